// luf_fw_wide.cuh -- the forward (sliding-window) scheme of localuf/_schemes.py:282-450 around the wide instance
// of the UF decoder (luf_uf_wide.cuh): windows beyond the 16-bit layout of luf_fw.cuh (Surface(23 | 25,
// 'circuit-level') and taller buffers).  Same cycle as luf_fw.cuh --
//   error    = buffer leftover lowered by the commit height  |  fresh sheet      (:302-324)
//   syndrome = syndrome(error) ^ artificial defects of the previous commit     (:351-368)
//   decoder.reset(); decoder.decode(syndrome)
//   commit leftover = (error ^ correction) & commit region; buffer leftover = error & buffer region (:326-339)
//   Pairs.load(commit leftover) in ascending edge id; LogicalCounter.count      (:341-349; _pairs.py)
// -- with 32-bit tables, one thread block per stream, and the stream's state (decoder slab + window bitmaps +
// partner maps) wherever the kernel binds it (shared memory or the L2-resident global slab).  The lane that
// loads the string pairs walks a one-bit-per-word summary of the commit leftover instead of the leftover
// itself (a window of 10^5 edges keeps a few dozen).  Dual compilation as in luf_core.cuh.
#pragma once
#include "luf_uf_wide.cuh"

#if defined(LUF_HOST_EMU)
  #define LUF_W_PHASE(stmt) LUF_PHASE(stmt)
#else
  #define LUF_W_PHASE(stmt) { stmt; } __syncthreads();
#endif

namespace luf {
namespace fw_wide {

static const uint32_t W_NONE = 0xFFFFFFFFu;

struct FwView {
    int N, B, E, WE, WD, c;
    const uint64_t* edge_uv;        // [E]   u | v << 32
    const uint32_t* lower_edge;     // [E]   W_NONE: leaves the window
    const uint32_t* commit_mask;    // [WE]
    const uint32_t* art_node;       // [2E]  W_NONE: none
    const uint32_t* fb_mask;        // [WE]
    const uint32_t* node_lower;     // [N]   W_NONE: leaves the window
    const uint8_t* node_side;       // [N]   1 west, 2 east
};

struct FwLayout { uint32_t buf, commit, nz, partner, partner2, bytes; };   // appended to the decoder's slab

struct FwState {
    uint32_t *buf, *commit, *nz;    // nz: bit w set iff commit[w] != 0
    int32_t *partner, *partner2;
};

LUF_DEV FwState fw_bind(const FwLayout& l, unsigned char* b)
{
    FwState f;
    f.buf = (uint32_t*)(b + l.buf); f.commit = (uint32_t*)(b + l.commit); f.nz = (uint32_t*)(b + l.nz);
    f.partner = (int32_t*)(b + l.partner); f.partner2 = (int32_t*)(b + l.partner2);
    return f;
}

static const int FWP_NONE = -1, FWP_WEST = -2, FWP_EAST = -3, FWP_DEAD = -4;

LUF_DEV void fw_reset(const FwView& g, const FwState& f, const uint32_t* init_buf, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) { f.buf[w] = init_buf ? init_buf[w] : 0u; f.commit[w] = 0; }
    LUF_NOUNROLL_CA
    for (int w = lane; w < (g.WE + 31) / 32; w += nl) f.nz[w] = 0;
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) f.partner[v] = FWP_NONE;
}

LUF_DEV void fw_toggle_ends(const FwView& g, uint32_t* defect, uint32_t e)
{
    const uint64_t uv = LUF_LDG64(&g.edge_uv[e]);
    const uint32_t u = (uint32_t)uv, v = (uint32_t)(uv >> 32);
    if (u >= (uint32_t)g.B) luf_atomic_xor(&defect[(u - g.B) >> 5], 1u << ((u - g.B) & 31u));
    if (v >= (uint32_t)g.B) luf_atomic_xor(&defect[(v - g.B) >> 5], 1u << ((v - g.B) & 31u));
}

// error := lowered buffer leftover; syndrome of it; artificial defects.  `err` and `defect` zero on entry.
LUF_DEV void fw_begin_cycle(const FwView& g, const FwState& f, uint32_t* err, uint32_t* defect, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) {
        uint32_t x = f.buf[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            const uint32_t le = LUF_LDG(&g.lower_edge[32 * w + b]);
            if (le == W_NONE) continue;
            luf_atomic_or(&err[le >> 5], 1u << (le & 31u));
            fw_toggle_ends(g, defect, le);
        }
        x = f.commit[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            LUF_NOUNROLL_CA
            for (int k = 0; k < 2; ++k) {
                const uint32_t a = LUF_LDG(&g.art_node[2 * (size_t)(32 * w + b) + k]);
                if (a != W_NONE) luf_atomic_xor(&defect[(a - g.B) >> 5], 1u << ((a - g.B) & 31u));
            }
        }
    }
}

LUF_DEV void fw_given_sheet(const FwView& g, const uint32_t* fresh, uint32_t* err, uint32_t* defect, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) {
        uint32_t x = fresh[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            luf_atomic_or(&err[w], 1u << b);
            fw_toggle_ends(g, defect, (uint32_t)(32 * w + b));
        }
    }
}

// Forward._get_leftover (:326-339); optionally exports the commit leftover.  f.nz is zero on entry.
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
        if (cm) luf_atomic_or(&f.nz[w >> 5], 1u << (w & 31));
    }
}

LUF_DEV int fw_key(const FwView& g, uint32_t v)
{
    const uint32_t s = LUF_LDG(&g.node_side[v]);
    return s == 1 ? FWP_WEST : s == 2 ? FWP_EAST : (int)v;
}
LUF_DEV void fw_add(const FwState& f, int a, int b, int& count)
{
    if (a >= 0) f.partner[a] = b;
    if (b >= 0) f.partner[b] = a;
    if ((a == FWP_WEST && b == FWP_EAST) || (a == FWP_EAST && b == FWP_WEST)) ++count;
}
LUF_DEV int fw_take(const FwState& f, int u)
{
    const int w = f.partner[u];
    f.partner[u] = FWP_NONE;
    if (w >= 0) f.partner[w] = FWP_NONE;
    return w;
}

// Forward.get_logical_error (:341-349): commit leftover into Pairs, ascending edge id.  One lane; clears f.nz.
LUF_DEV void fw_pairs_load(const FwView& g, const FwState& f, int& count)
{
    LUF_NOUNROLL_CA
    for (int sw = 0; sw < (g.WE + 31) / 32; ++sw) {
        uint32_t words = f.nz[sw];
        if (!words) continue;
        f.nz[sw] = 0;
        LUF_NOUNROLL_CA
        while (words) {
            const int wb = luf_ffs(words) - 1; words &= words - 1;
            const int w = 32 * sw + wb;
            uint32_t x = f.commit[w];
            LUF_NOUNROLL_CA
            while (x) {
                const int b = luf_ffs(x) - 1; x &= x - 1;
                const uint64_t uv = LUF_LDG64(&g.edge_uv[32 * w + b]);
                const int u = fw_key(g, (uint32_t)uv), v = fw_key(g, (uint32_t)(uv >> 32));
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
}

// LogicalCounter.count's lowering by the commit height (_pairs.py:82-111), three lane-parallel sweeps
LUF_DEV void fw_lower_clear(const FwView& g, const FwState& f, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) f.partner2[v] = FWP_NONE;
}
LUF_DEV void fw_lower_scatter(const FwView& g, const FwState& f, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) {
        const int p = f.partner[v];
        if (p == FWP_NONE) continue;
        const uint32_t lv = LUF_LDG(&g.node_lower[v]);
        if (lv == W_NONE) continue;
        int q = p;
        if (p >= 0) { const uint32_t lp = LUF_LDG(&g.node_lower[p]); q = lp == W_NONE ? FWP_DEAD : (int)lp; }
        f.partner2[lv] = q;
    }
}
LUF_DEV void fw_lower_commit(const FwView& g, const FwState& f, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) f.partner[v] = f.partner2[v];
}

// leftover, pairs, count, lowering.  Returns the count on lane 0.
LUF_DEV int fw_end_cycle(const FwView& g, const FwState& f, const uint32_t* err, const uint32_t* corr,
                         uint32_t* commit_out, int lane, int nl)
{
    int count = 0;
    (void)lane;
    LUF_W_PHASE(fw_leftover(g, f, err, corr, commit_out, lane, nl); fw_lower_clear(g, f, lane, nl));
    LUF_W_PHASE(if (lane == 0) fw_pairs_load(g, f, count));
    LUF_W_PHASE(fw_lower_scatter(g, f, lane, nl));
    LUF_W_PHASE(fw_lower_commit(g, f, lane, nl));
    return count;
}

struct FwUf {                       // decoders/uf.py on the window graph, wide instance
    const uf_wide::GraphView& g; const uf_wide::Layout& l; uf_wide::Shot s; int flags;
};

LUF_DEV void fw_dec_decode(const FwUf& d, int& s0, int& s1, int lane, int nl)
{
    LUF_W_PHASE(uf_wide::ph_load(d.g, d.s, lane, nl));
    s0 = uf_wide::uf_validate_t<-1>(d.g, d.s, d.flags, lane, nl);
    s1 = 0;
    uf_wide::uf_peel(d.g, d.s, lane, nl);
}

// One cycle on GIVEN fresh errors.  Returns the logical errors counted (lane 0).
LUF_DEV int fw_cycle_given(const FwView& g, const FwState& f, const FwUf& d, const uint32_t* fresh,
                           uint32_t* commit_out, int& s0, int& s1, int lane, int nl)
{
    (void)lane;
    LUF_W_PHASE(uf_wide::ph_reset(d.g, d.l, d.s, lane, nl));
    LUF_W_PHASE(fw_begin_cycle(g, f, d.s.err, d.s.defect, lane, nl); fw_given_sheet(g, fresh, d.s.err, d.s.defect, lane, nl));
    fw_dec_decode(d, s0, s1, lane, nl);
    return fw_end_cycle(g, f, d.s.err, d.s.corr, commit_out, lane, nl);
}

// Forward.run (:392-450) with device-sampled noise (as luf::fw_run; the first 32 lanes are the sampler's).
LUF_DEV void fw_run(const FwView& g, const FwState& f, const FwUf& d, const SampleTab& buffer_tab,
                    const SampleParams& sp, uint64_t stream, int n, int n_cleanse, int& m_total, long long& cycles,
                    long long& sum0, long long& sum1, int32_t* steps_out, int lane, int nl)
{
    SampleTab fresh;
    fresh.F = d.g.F; fresh.n_classes = d.g.n_classes; fresh.perm = 0; fresh.class_end = d.g.class_end; fresh.perm32 = d.g.fresh_perm32;
    (void)lane;
    // _make_error_in_buffer_region: every buffer edge flips once with its class probability
    LUF_W_PHASE(uf_wide::ph_reset(d.g, d.l, d.s, lane, nl); fw_reset(g, f, (const uint32_t*)0, lane, nl));
    LUF_W_PHASE(if (lane < 32) uf_wide::ph_sample_tab(d.g, buffer_tab, d.s, sp, stream, 0u, (const uint32_t*)0, lane, 32));
    LUF_W_PHASE(for (int w = lane; w < g.WE; w += nl) f.buf[w] = d.s.err[w]);
    const int total = n + 1 + n_cleanse;
    LUF_NOUNROLL_CA
    for (int c = 0; c < total; ++c) {
        int s0 = 0, s1 = 0;
        LUF_W_PHASE(uf_wide::ph_reset(d.g, d.l, d.s, lane, nl));
        LUF_W_PHASE(fw_begin_cycle(g, f, d.s.err, d.s.defect, lane, nl);
                    if (c <= n && lane < 32) uf_wide::ph_sample_tab(d.g, fresh, d.s, sp, stream, (uint32_t)(c + 1), c == n ? g.fb_mask : (const uint32_t*)0, lane, 32));
        fw_dec_decode(d, s0, s1, lane, nl);
        m_total += fw_end_cycle(g, f, d.s.err, d.s.corr, (uint32_t*)0, lane, nl);
        if (c < n) {
            sum0 += s0; sum1 += s1;
            if (steps_out && lane == 0) { steps_out[2 * c] = s0; steps_out[2 * c + 1] = s1; }
        }
    }
    cycles += total;
}

}  // namespace fw_wide
}  // namespace luf
