// luf_fw.cuh -- the forward (sliding-window) decoding scheme of
// localuf/_schemes.py:282-450 around the batch decoders (UF, Macar): one warp
// owns one stream; the window's error bits, the previous commit leftover and
// the string pairs of the logical counter stay in the warp's shared-memory
// slice across decoding cycles.  Per cycle (Forward.run, :433-447):
//   error    = buffer leftover lowered by the commit height  |  fresh sheet      (:302-324)
//   syndrome = syndrome(error) ^ artificial defects of the previous commit     (:351-368)
//   decoder.reset(); decoder.decode(syndrome)
//   commit leftover = (error ^ correction) & commit region; buffer leftover = error & buffer region (:326-339)
//   Pairs.load(commit leftover) in ascending edge id; LogicalCounter.count      (:341-349; _pairs.py)
// Edge / node indices are those of the flat graph.  Dual compilation as in luf_core.cuh.
#pragma once
#include "luf_core.cuh"
#include "luf_ca.cuh"

namespace luf {

struct FwView {
    int N, B, E, WE, WD, c;
    const uint32_t* edge_uv;        // [E]
    const uint16_t* lower_edge;     // [E]   0xFFFF: leaves the window
    const uint32_t* commit_mask;    // [WE]
    const uint16_t* art_node;       // [2E]  0xFFFF: none
    const uint32_t* fb_mask;        // [WE]  edges touching the future boundary
    const uint16_t* node_lower;     // [N]   0xFFFF: leaves the window
    const uint8_t* node_side;       // [N]   1 west, 2 east
};

struct FwLayout { uint32_t buf, commit, partner, partner2, bytes; };   // appended to the decoder's slice

struct FwState {
    uint32_t *buf, *commit;
    int16_t *partner, *partner2;
};

LUF_DEV FwState fw_bind(const FwLayout& l, unsigned char* b)
{
    FwState f;
    f.buf = (uint32_t*)(b + l.buf); f.commit = (uint32_t*)(b + l.commit);
    f.partner = (int16_t*)(b + l.partner); f.partner2 = (int16_t*)(b + l.partner2);
    return f;
}

static const int FWP_NONE = -1, FWP_WEST = -2, FWP_EAST = -3, FWP_DEAD = -4;

LUF_DEV void fw_reset(const FwView& g, const FwState& f, const uint32_t* init_buf, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) { f.buf[w] = init_buf ? init_buf[w] : 0u; f.commit[w] = 0; }
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) f.partner[v] = (int16_t)FWP_NONE;
}

// toggle the defect bits of an edge's detector endpoints (_base_classes.py:427-450)
LUF_DEV void fw_toggle_ends(const FwView& g, uint32_t* defect, uint32_t e)
{
    const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
    const uint32_t u = uv & 0xFFFFu, v = uv >> 16;
    if (u >= (uint32_t)g.B) luf_atomic_xor(&defect[(u - g.B) >> 5], 1u << ((u - g.B) & 31u));
    if (v >= (uint32_t)g.B) luf_atomic_xor(&defect[(v - g.B) >> 5], 1u << ((v - g.B) & 31u));
}

// error := lowered buffer leftover; syndrome of it; artificial defects.  `err`
// and `defect` must be zero on entry (the decoder's reset).
LUF_DEV void fw_begin_cycle(const FwView& g, const FwState& f, uint32_t* err, uint32_t* defect, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) {
        uint32_t x = f.buf[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            const uint32_t le = LUF_LDG(&g.lower_edge[32 * w + b]);
            if (le == 0xFFFFu) continue;
            luf_atomic_or(&err[le >> 5], 1u << (le & 31u));
            fw_toggle_ends(g, defect, le);
        }
        x = f.commit[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            LUF_NOUNROLL_CA
            for (int k = 0; k < 2; ++k) {
                const uint32_t a = LUF_LDG(&g.art_node[2 * (32 * w + b) + k]);
                if (a != 0xFFFFu) luf_atomic_xor(&defect[(a - g.B) >> 5], 1u << ((a - g.B) & 31u));
            }
        }
    }
}

// fresh sheet given as bits over code edge ids (parity path)
LUF_DEV void fw_given_sheet(const FwView& g, const uint32_t* fresh, uint32_t* err, uint32_t* defect, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) {
        uint32_t x = fresh[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            luf_atomic_or(&err[(32 * w + b) >> 5], 1u << b);
            fw_toggle_ends(g, defect, (uint32_t)(32 * w + b));
        }
    }
}

// Forward._get_leftover (:326-339); optionally exports the commit leftover
LUF_DEV void fw_leftover(const FwView& g, const FwState& f, const uint32_t* err, const uint32_t* corr,
                         uint32_t* commit_out, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) {
        const uint32_t m = LUF_LDG(&g.commit_mask[w]);
        const uint32_t cm = (err[w] ^ corr[w]) & m;
        f.commit[w] = cm;
        f.buf[w] = err[w] & ~m;
        if (commit_out) commit_out[w] = cm;
    }
}

LUF_DEV int fw_key(const FwView& g, uint32_t v)
{
    const uint32_t s = LUF_LDG(&g.node_side[v]);
    return s == 1 ? FWP_WEST : s == 2 ? FWP_EAST : (int)v;
}
LUF_DEV void fw_add(const FwState& f, int a, int b, int& count)
{
    if (a >= 0) f.partner[a] = (int16_t)b;
    if (b >= 0) f.partner[b] = (int16_t)a;
    if ((a == FWP_WEST && b == FWP_EAST) || (a == FWP_EAST && b == FWP_WEST)) ++count;
}
LUF_DEV int fw_take(const FwState& f, int u)
{
    const int w = f.partner[u];
    f.partner[u] = (int16_t)FWP_NONE;
    if (w >= 0) f.partner[w] = (int16_t)FWP_NONE;
    return w;
}

// Forward.get_logical_error (:341-349): commit leftover into Pairs, ascending edge id.  One lane.
LUF_DEV void fw_pairs_load(const FwView& g, const FwState& f, int& count)
{
    LUF_NOUNROLL_CA
    for (int w = 0; w < g.WE; ++w) {
        uint32_t x = f.commit[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            const uint32_t uv = LUF_LDG(&g.edge_uv[32 * w + b]);
            const int u = fw_key(g, uv & 0xFFFFu), v = fw_key(g, uv >> 16);
            if (u >= 0 && f.partner[u] != FWP_NONE) {                     // Pairs.load, _pairs.py:49-66
                const int p = fw_take(f, u);
                if (v >= 0 && f.partner[v] != FWP_NONE) fw_add(f, p, fw_take(f, v), count);
                else if (v != p || v < 0) fw_add(f, v, p, count);
            } else if (v >= 0 && f.partner[v] != FWP_NONE)
                fw_add(f, u, fw_take(f, v), count);
            else
                fw_add(f, u, v, count);
        }
    }
}

// LogicalCounter.count's lowering by the commit height (_pairs.py:82-111), three lane-parallel sweeps
LUF_DEV void fw_lower_clear(const FwView& g, const FwState& f, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) f.partner2[v] = (int16_t)FWP_NONE;
}
LUF_DEV void fw_lower_scatter(const FwView& g, const FwState& f, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) {
        const int p = f.partner[v];
        if (p == FWP_NONE) continue;
        const uint32_t lv = LUF_LDG(&g.node_lower[v]);
        if (lv == 0xFFFFu) continue;
        int q = p;
        if (p >= 0) { const uint32_t lp = LUF_LDG(&g.node_lower[p]); q = lp == 0xFFFFu ? FWP_DEAD : (int)lp; }
        f.partner2[lv] = (int16_t)q;
    }
}
LUF_DEV void fw_lower_commit(const FwView& g, const FwState& f, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) f.partner[v] = f.partner2[v];
}

// tail shared by both decoders: leftover, pairs, count, lowering.  Returns the count on lane 0.
LUF_DEV int fw_end_cycle(const FwView& g, const FwState& f, const uint32_t* err, const uint32_t* corr,
                         uint32_t* commit_out, int lane, int nl)
{
    int count = 0;
    (void)lane;
    LUF_PHASE(fw_leftover(g, f, err, corr, commit_out, lane, nl); fw_lower_clear(g, f, lane, nl));
    LUF_PHASE(if (lane == 0) fw_pairs_load(g, f, count));
    LUF_PHASE(fw_lower_scatter(g, f, lane, nl));
    LUF_PHASE(fw_lower_commit(g, f, lane, nl));
    return count;
}

// ------------------------------------------------ decoder policies for the driver
struct FwUf {                       // decoders/uf.py on the window graph
    const GraphView& g; const Layout& l; Shot s; int flags;
    LUF_MEM uint32_t* err() const { return s.err; }
    LUF_MEM uint32_t* defect() const { return s.defect; }
    LUF_MEM uint32_t* corr() const { return s.corr; }
};
struct FwMacar {                    // decoders/luf/main.py (Macar) on the window graph
    const CaView& g; CaShot s;
    LUF_MEM uint32_t* err() const { return s.err; }
    LUF_MEM uint32_t* defect() const { return s.defect; }
    LUF_MEM uint32_t* corr() const { return s.corr; }
};

LUF_DEV void fw_dec_reset(const FwUf& d, int lane, int nl) { ph_reset(d.g, d.l, d.s, lane, nl); }
LUF_DEV void fw_dec_reset(const FwMacar& d, int lane, int nl) { ca_reset<false>(d.g, d.s, lane, nl); }

// decoder.decode(syndrome): steps = (rounds, 0) for UF, (tSV, tBP) for Macar
LUF_DEV void fw_dec_decode(const FwUf& d, int& s0, int& s1, int lane, int nl)
{
    LUF_PHASE(ph_load(d.g, d.s, lane, nl));
    s0 = uf_validate(d.g, d.s, d.flags, lane, nl);
    s1 = 0;
    uf_peel(d.g, d.s, lane, nl);
}
LUF_DEV void fw_dec_decode(const FwMacar& d, int& s0, int& s1, int lane, int nl)
{
    (void)lane;
    LUF_PHASE(ca_load(d.g, d.s, lane, nl));
    ca_decode<false>(d.g, d.s, 0, &s0, &s1, (uint32_t*)0, lane, nl);
}

// One cycle on GIVEN fresh errors.  Returns the logical errors counted (lane 0).
template <class Dec>
LUF_DEV int fw_cycle_given(const FwView& g, const FwState& f, const Dec& d, const uint32_t* fresh,
                           uint32_t* commit_out, int& s0, int& s1, int lane, int nl)
{
    (void)lane;
    LUF_PHASE(fw_dec_reset(d, lane, nl));
    LUF_PHASE(fw_begin_cycle(g, f, d.err(), d.defect(), lane, nl); fw_given_sheet(g, fresh, d.err(), d.defect(), lane, nl));
    fw_dec_decode(d, s0, s1, lane, nl);
    return fw_end_cycle(g, f, d.err(), d.corr(), commit_out, lane, nl);
}

// Forward.run (:392-450) with device-sampled noise: buffer region filled first
// (:370-390), then n timed cycles, one without future-boundary faults, h//c
// noiseless ones.  steps_out (optional) [n][2].  m on lane 0.
template <class Dec>
LUF_DEV void fw_run(const GraphView& gv, const FwView& g, const FwState& f, const Dec& d, const SampleTab& buffer_tab,
                    const SampleParams& sp, uint64_t stream, int n, int n_cleanse, int& m_total, long long& cycles,
                    long long& sum0, long long& sum1, int32_t* steps_out, int lane, int nl)
{
    SampleTab fresh;
    fresh.F = gv.F; fresh.n_classes = gv.n_classes; fresh.perm = gv.fresh_perm; fresh.class_end = gv.class_end; fresh.perm32 = 0;
    Shot samp;                                   // the sampler writes the decoder's error / defect bitmaps
    samp.err = d.err(); samp.defect = d.defect();
    // _make_error_in_buffer_region: every buffer edge flips once with its class probability
    LUF_PHASE(fw_dec_reset(d, lane, nl); fw_reset(g, f, (const uint32_t*)0, lane, nl));
    LUF_PHASE(ph_sample_tab(gv, buffer_tab, samp, sp, stream, 0u, (const uint32_t*)0, lane, nl));
    LUF_PHASE(for (int w = lane; w < g.WE; w += nl) f.buf[w] = d.err()[w]);
    const int total = n + 1 + n_cleanse;
    LUF_NOUNROLL_CA
    for (int c = 0; c < total; ++c) {
        int s0 = 0, s1 = 0;
        LUF_PHASE(fw_dec_reset(d, lane, nl));
        LUF_PHASE(fw_begin_cycle(g, f, d.err(), d.defect(), lane, nl);
                  if (c <= n) ph_sample_tab(gv, fresh, samp, sp, stream, (uint32_t)(c + 1), c == n ? g.fb_mask : (const uint32_t*)0, lane, nl));
        fw_dec_decode(d, s0, s1, lane, nl);
        m_total += fw_end_cycle(g, f, d.err(), d.corr(), (uint32_t*)0, lane, nl);
        if (c < n) {
            sum0 += s0; sum1 += s1;
            if (steps_out && lane == 0) { steps_out[2 * c] = s0; steps_out[2 * c + 1] = s1; }
        }
    }
    cycles += total;
}

}  // namespace luf
