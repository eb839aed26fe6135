// luf_ca_drivers.inl -- LUF.decode for Macar / Actis against a tile of nl cooperating lanes; included by
// luf_ca.cuh once per tile shape (LUF_CA_NS, LUF_CT_PHASE, LUF_CT_ANY).  The node rules and the per-step
// phases it calls live in luf_ca.cuh.
namespace luf {
namespace LUF_CA_NS {

// LUF.decode = validate (luf/main.py:112-145) then peel (:147-166).  Writes
// (tSV, tBP) to steps; exports the growth after validation if growth_out.
// Returns the west parity of the correction (per lane on the device).
template <bool ACTIS>
LUF_DEV uint32_t ca_decode(const CaView& g, const CaShot& s, int flags, int* t_sv, int* t_bp,
                           uint32_t* growth_out, int lane, int nl)
{
    CaCtl c;
    c.stage = ST_GROWING; c.step = 0;
    c.n_busy = 0; c.n_valid = ACTIS ? 0 : 1; c.n_countdown = 0; c.n_busy_signal = 0; c.n_active_signal = 0;
    c.n_next_active_signal = 0;
    int tsv = 0, tbp = 0, exported = 0;
    uint32_t parity = 0;
    (void)lane;
    if (!ACTIS) { LUF_CT_PHASE(ca_gather_touched(g, s, lane, nl)); }     // glob[2] == 0 after ca_reset
    LUF_NOUNROLL_CA
    while (c.stage != ST_DONE) {
        if (c.stage == ST_BURNING && !exported) {
            exported = 1;
            if (growth_out) { LUF_CT_PHASE(for (int w = lane; w < g.GW; w += nl) growth_out[w] = s.growth[w]); }
        }
        const int in_sv = c.stage < ST_BURNING;
        const int grew = !ACTIS && c.stage == ST_GROWING;                // this step may touch new nodes
        int busy, valid;
        if (ACTIS) {
            LUF_CT_PHASE(actis_phase_a(g, s, c, lane, nl));
            // what node 0 handed to the collection this timestep (ControllerFriendship.relay_signals)
            const int next_busy_signal = (int)s.glob[0];
            c.n_next_active_signal |= (int)s.glob[1];
            actis_waiter_advance(g, c, flags);                       // ActisNodes.advance, :565-567
            LUF_CT_PHASE(actis_phase_b(g, s, c, lane, nl));
            c.n_busy_signal = next_busy_signal;                      // _update_unphysicals_for_actis, :575-581
            c.n_active_signal = c.n_next_active_signal;
            busy = c.n_busy; valid = c.n_valid;
        } else {
#if defined(LUF_CT_WORD_VOTES)
            // Multi-warp tiles: the two votes of a timestep (any busy, any invalid) are one OR into a word of the
            // shot's state -- glob[0] / glob[1] in turns (Actis's inbox, idle under Macar), the next one cleared by
            // lane 0 while this one is filled -- so that a timestep costs two tile barriers instead of four.
            uint32_t* vote = &s.glob[c.step & 1];
            LUF_CT_PHASE(parity ^= macar_phase_a(g, s, c.stage, lane, nl));
            {
                const uint32_t mine = macar_phase_b(g, s, c.stage, lane, nl);
                if (mine) luf_atomic_or(vote, mine);
                if (lane == 0) s.glob[(c.step + 1) & 1] = 0u;
            }
            LUF_CT_SYNC();
            const uint32_t acc = *vote;
            busy = (acc & 1u) != 0; valid = !(acc & 2u);
#else
            uint32_t acc = 0;
            LUF_CT_PHASE(parity ^= macar_phase_a(g, s, c.stage, lane, nl));
            LUF_CT_PHASE(acc |= macar_phase_b(g, s, c.stage, lane, nl));
            busy = LUF_CT_ANY(acc & 1u); valid = !LUF_CT_ANY(acc & 2u);
#endif
        }
        const int growth_changed = ca_controller_advance(c, busy, valid);
        if (grew) {                                                      // newly touched nodes join the list
            LUF_CT_PHASE(if (lane == 0) s.glob[2] = 0);
            LUF_CT_PHASE(ca_gather_touched(g, s, lane, nl));
        }
        if (growth_changed) { LUF_CT_PHASE(ca_snapshot_access<ACTIS>(g, s, lane, nl)); }
        c.step += 1;
        if (in_sv) ++tsv; else ++tbp;
        if (tsv + tbp >= CA_MAX_STEPS) { tsv = -1; tbp = -1; break; }
    }
    *t_sv = tsv; *t_bp = tbp;
    return parity;
}


}  // namespace LUF_CA_NS
}  // namespace luf
