// luf_sf_drivers.inl -- one decoding cycle / one Frugal.run of Snowflake, written against a tile of
// nl cooperating lanes; included by luf_sf.cuh once per tile shape (LUF_SF_NS, LUF_ST_PHASE,
// LUF_ST_ANY).  The per-node rules and the window bookkeeping they call live in luf_sf.cuh.
namespace luf {
namespace LUF_SF_NS {

#define LUF_SF_REGATHER() { LUF_ST_PHASE(if (lane == 0) s.acnt[0] = 0); LUF_ST_PHASE(sf_gather_awake(g, s, lane, nl)); }

// Snowflake.merge (:324-353): adds the 'all' and 'unrooting' runtimes of one merge loop.
LUF_DEV void sf_merge(const SfView& g, const SfShot& s, bool whole, bool slow, int& rt_all, int& rt_unr, int lane, int nl)
{
    (void)lane;
    rt_all += 1;
    LUF_NOUNROLL_CA
    for (int guard = 0; guard < (1 << 20); ++guard) {
#if defined(LUF_ST_BLOCK_VOTES)
        // Block tile: the two votes of a timestep (any busy, any RESET) are one OR into a word of shared memory --
        // two alternating words, the next one cleared by lane 0 while this one is filled -- so that a timestep costs
        // two block barriers instead of four (phase A, phase B, and a __syncthreads_or per vote).
        uint32_t* vote = &s.acnt[1 + (guard & 1)];
        if (guard == 0 && lane == 0) *vote = 0u;
        LUF_ST_PHASE(sf_merge_a(g, s, whole, slow, lane, nl));
        {
            const uint32_t mine = sf_merge_b(g, s, lane, nl);
            if (mine) luf_atomic_or(vote, mine);
            if (lane == 0) s.acnt[1 + ((guard + 1) & 1)] = 0u;
        }
        __syncthreads();
        const uint32_t acc = *vote;
        rt_all += 1;
        if (acc & 2u) rt_unr += 1;
        if (!(acc & 1u)) break;
#else
        LUF_ST_PHASE(sf_merge_a(g, s, whole, slow, lane, nl));
        uint32_t acc = 0;
        LUF_ST_PHASE(acc |= sf_merge_b(g, s, lane, nl));
        rt_all += 1;
        if (LUF_ST_ANY(acc & 2u)) rt_unr += 1;
        if (!LUF_ST_ANY(acc & 1u)) break;
#endif
    }
}

// One decoding cycle after the fresh sheet is in place: Frugal.advance's tail
// (_schemes.py:524-530) = Snowflake.decode (:216-281), then get_logical_error
// (_schemes.py:262-276).  `floor_out` (optional, ELw words, global) receives the
// floor-sheet correction committed by this cycle's drop.  `m` is only
// meaningful on lane 0 (in the emulation: the single accumulated value).
// `drop_only`: decode(defects_possible=False), snowflake/main.py:244-264 -- the window moves, nothing grows or
// merges; the runtime is 1 timestep for time_only = 'all' and 0 otherwise (rt_all = 1, rt_unr = 0).
LUF_DEV void sf_cycle_tail(const SfView& g, const SfShot& s, int flags, uint32_t* floor_out,
                           int& m, int& rt_all, int& rt_unr, int lane, int nl, bool drop_only = false)
{
    const bool slow = flags & SF_FLAG_SLOW, one_one = flags & SF_FLAG_ONE_ONE;
    (void)lane;
    rt_all = 0; rt_unr = 0;
    // floor corrections to Pairs (FloorContact.drop), then the window moves
    // (one phase: the lane that owns Pairs walks the floor sheet itself -- a dozen words, mostly empty -- instead of a
    // vote of the tile on whether there is anything to load)
    LUF_ST_PHASE(if (floor_out) for (int c = lane; c < g.ELw; c += nl) floor_out[c] = s.corr[(g.h - 1) * g.ELw + c];
                 if (lane == 0) sf_pairs_load_block(g, s, s.corr + (g.h - 1) * g.ELw, m));
    LUF_ST_PHASE(sf_drop(g, s, lane, nl));
    LUF_SF_REGATHER();                               // the window moved
    if (drop_only) {
        rt_all = 1;
    } else if (one_one) {
        LUF_ST_PHASE(sf_grow(g, s, 0, lane, nl));
        LUF_SF_REGATHER();                           // growth wakes nodes
        sf_merge(g, s, true, slow, rt_all, rt_unr, lane, nl);
    } else {
        LUF_ST_PHASE(sf_grow(g, s, 1, lane, nl));
        LUF_SF_REGATHER();
        sf_merge(g, s, true, slow, rt_all, rt_unr, lane, nl);
        LUF_ST_PHASE(sf_grow(g, s, 2, lane, nl));
        LUF_SF_REGATHER();
        sf_merge(g, s, false, slow, rt_all, rt_unr, lane, nl);
    }
    LUF_ST_PHASE(sf_pairs_retire(g, s, lane, nl));
    LUF_ST_PHASE(sf_pairs_lower(g, s, lane, nl));
}

// After a decoding cycle: the window's growth (h * 2 * ELw words) and (min_defect_height, min_active_layer).
LUF_DEV void sf_cycle_metrics(const SfView& g, const SfShot& s, uint32_t* growth_out, int32_t* mins_out, int lane, int nl)
{
    (void)lane;
    LUF_ST_PHASE(if (lane == 0) { s.acnt[1] = (uint32_t)g.h; s.acnt[2] = (uint32_t)g.h; });
    LUF_ST_PHASE(sf_metrics_scan(g, s, growth_out, lane, nl));
    LUF_ST_PHASE(if (lane == 0) { mins_out[0] = (int32_t)s.acnt[1]; mins_out[1] = (int32_t)s.acnt[2]; });
}

// Frugal._raise_window's Pairs half: floor errors in ascending edge id (one lane), then shift.
LUF_DEV void sf_cycle_head(const SfView& g, const SfShot& s, int& m, int lane, int nl)
{
    (void)lane;
    LUF_ST_PHASE(if (lane == 0) sf_pairs_load_block(g, s, s.err + (g.h - 1) * g.ELw, m));
    LUF_ST_PHASE(sf_shift_errors(g, s, lane, nl));
    LUF_ST_PHASE(sf_begin_sheet(g, s, lane, nl));
}

// Frugal.run (_schemes.py:566-649) for one stream with device-sampled noise:
// transient h cycles, n*d timed cycles, d-1 more, one without future-boundary
// faults, 2h noiseless cleansing cycles.  `steps_out` (optional, [n][2]) gets the
// 'all' / 'unrooting' runtimes summed over each timed block of d cycles.
// mgrowth / mcycle (optional, [n*d][h*2*ELw] and [n*d][4]): after every timed cycle the window's growth and
// (min_defect_height, min_active_layer, runtime 'all', runtime 'unrooting') -- Frugal.run(metrics=...).
// m / runtimes are meaningful on lane 0.
LUF_DEV void sf_run(const SfView& g, const SfShot& s, const SampleParams& sp, int flags, uint64_t stream, int n,
                    int& m_total, long long& cycles, long long& rt_all_timed, long long& rt_unr_timed,
                    int32_t* steps_out, uint32_t* mgrowth, int32_t* mcycle, int lane, int nl)
{
    const int d = g.d, h = g.h;
    const int total = h + n * d + (d - 1) + 1 + 2 * h;
    LUF_ST_PHASE(sf_reset(g, s, lane, nl));
    int blk_all = 0, blk_unr = 0;
    LUF_NOUNROLL_CA
    for (int c = 0; c < total; ++c) {
        const bool cleanse = c >= total - 2 * h, excl = c == total - 2 * h - 1;
        const bool timed = c >= h && c < h + n * d;
        int m = 0, ra = 0, ru = 0;
        sf_cycle_head(g, s, m, lane, nl);
        // 32 sampling lanes whatever the tile: a stream's noise does not depend on the tile shape
        if (!cleanse) { LUF_ST_PHASE(if (lane < 32) sf_sample_sheet(g, s, sp, stream, (uint32_t)c, excl, lane, nl < 32 ? nl : 32)); }
        sf_cycle_tail(g, s, flags, nullptr, m, ra, ru, lane, nl);
        m_total += m;
        if (timed && mgrowth) {
            const size_t tc = (size_t)(c - h);
            sf_cycle_metrics(g, s, mgrowth + tc * (size_t)(h * 2 * g.ELw), mcycle + 4 * tc, lane, nl);
            if (lane == 0) { mcycle[4 * tc + 2] = ra; mcycle[4 * tc + 3] = ru; }
        }
        if (timed) {
            rt_all_timed += ra; rt_unr_timed += ru; blk_all += ra; blk_unr += ru;
            if ((c - h) % d == d - 1) {
                if (steps_out && lane == 0) { steps_out[2 * ((c - h) / d)] = blk_all; steps_out[2 * ((c - h) / d) + 1] = blk_unr; }
                blk_all = 0; blk_unr = 0;
            }
        }
    }
    cycles += total;
}

// decoder.syndrome non-empty? (snowflake/main.py:143-149: any detector of the window with a defect -- boundary
// nodes do not count; defects only sit on awake nodes)
LUF_DEV bool sf_window_has_defect(const SfView& g, const SfShot& s, int lane, int nl)
{
    uint32_t found = 0;
    (void)lane;
    LUF_ST_PHASE(sf_for_awake(g, s, lane, nl, [&](int, int q, int v) { if (q >= g.NLb && (s.fl[v] & SF_DEFECT)) found = 1u; }));
    return LUF_ST_ANY(found);
}

// K3b parity path for one window: ncycles consecutive cycles whose fresh errors (or, with
// SF_FLAG_SYNDROME, top-sheet syndromes) are GIVEN as rows of ELw words.  rt [ncycles][2] = decode's
// return value for time_only = 'all' / 'unrooting'; floor_bits (optional) [ncycles][ELw]; m_out [ncycles];
// mgrowth / mins (optional, [ncycles][h*2*ELw] and [ncycles][2]): the window's growth and
// (min_defect_height, min_active_layer) after every cycle; drop_only (optional, [ncycles]): cycles on which the
// decoder only drops (defects_possible = False).
// SF_FLAG_LATENCY_TAIL: the last h - 1 cycles are the noiseless tail of Frugal.sample_latency.
LUF_DEV void sf_decode_stream(const SfView& g, const SfShot& s, int flags, int ncycles, const uint32_t* fresh,
                              int32_t* rt, uint32_t* floor_bits, int32_t* m_out, uint32_t* mgrowth, int32_t* mins,
                              const uint8_t* drop_only, int lane, int nl)
{
    LUF_ST_PHASE(sf_reset(g, s, lane, nl));
    bool possible = true;                        // defects_possible of Frugal.sample_latency
    LUF_NOUNROLL_CA
    for (int c = 0; c < ncycles; ++c) {
        int m = 0, ra = 0, ru = 0;
        if ((flags & SF_FLAG_LATENCY_TAIL) && possible && c >= ncycles - (g.h - 1))
            possible = sf_window_has_defect(g, s, lane, nl);
        if (!possible) {                         // the decoder only drops: 1 timestep ('all'), nothing to unroot
            if (lane == 0) { rt[2 * c] = 1; rt[2 * c + 1] = 0; m_out[c] = 0; }
            if (floor_bits) { LUF_ST_PHASE(for (int w = lane; w < g.ELw; w += nl) floor_bits[(size_t)c * g.ELw + w] = 0); }
            continue;
        }
        sf_cycle_head(g, s, m, lane, nl);
        if (flags & SF_FLAG_SYNDROME) { LUF_ST_PHASE(sf_given_syndrome(g, s, fresh + (size_t)c * g.ELw, lane, nl)); }
        else { LUF_ST_PHASE(sf_given_sheet(g, s, fresh + (size_t)c * g.ELw, lane, nl)); }
        sf_cycle_tail(g, s, flags, floor_bits ? floor_bits + (size_t)c * g.ELw : nullptr, m, ra, ru, lane, nl,
                      drop_only && drop_only[c]);
        if (lane == 0) { rt[2 * c] = ra; rt[2 * c + 1] = ru; m_out[c] = m; }
        if (mgrowth) sf_cycle_metrics(g, s, mgrowth + (size_t)c * (size_t)(g.h * 2 * g.ELw), mins + 2 * (size_t)c, lane, nl);
    }
}

// Frugal.sample_latency (_schemes.py:651-698): h + 1 noisy cycles, one more without future-boundary
// faults, then h - 1 noiseless ones; the latency is the runtime of the last h cycles.  From the first
// noiseless cycle that starts with no defect in the window on, the decoder only drops
// (defects_possible = False, snowflake/main.py:258-264): 1 timestep each for time_only = 'all', 0 otherwise.
// lat[0..2] = latency for time_only 'all' / 'merging' / 'unrooting' (meaningful on lane 0).
LUF_DEV void sf_latency(const SfView& g, const SfShot& s, const SampleParams& sp, int flags, uint64_t stream,
                        int lat[3], int lane, int nl)
{
    const int h = g.h, total = 2 * h + 1, loops = (flags & SF_FLAG_ONE_ONE) ? 1 : 2;
    LUF_ST_PHASE(sf_reset(g, s, lane, nl));
    lat[0] = lat[1] = lat[2] = 0;
    LUF_NOUNROLL_CA
    for (int c = 0; c < total; ++c) {
        const bool noisy = c <= h + 1, counted = c >= h + 1;
        if (!noisy && !sf_window_has_defect(g, s, lane, nl)) { lat[0] += total - c; break; }
        int m = 0, ra = 0, ru = 0;
        sf_cycle_head(g, s, m, lane, nl);
        if (noisy) { LUF_ST_PHASE(if (lane < 32) sf_sample_sheet(g, s, sp, stream, (uint32_t)c, c == h + 1, lane, nl < 32 ? nl : 32)); }
        sf_cycle_tail(g, s, flags, nullptr, m, ra, ru, lane, nl);
        if (counted) { lat[0] += ra; lat[1] += ra - 2 * loops; lat[2] += ru; }
    }
}

#undef LUF_SF_REGATHER

}  // namespace LUF_SF_NS
}  // namespace luf
