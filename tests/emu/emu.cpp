// Host emulation harness for the kernel phases of localuf_b200/csrc/luf_core.cuh.
//
// TEST INFRASTRUCTURE ONLY.  Compiles the *same* phase source the CUDA kernels
// use with -DLUF_HOST_EMU, runs the 32 lanes of a tile one after the other
// inside each phase, and exposes the result through a small C interface so that
// the CPU test-suite can compare the kernel logic with the oracle without a
// GPU.  It checks logic, not concurrency; the GPU parity tests do the rest.
#define LUF_HOST_EMU 1
#include <cstring>
#include <string>
#include <vector>

#include "../../localuf_b200/csrc/luf_host.h"

using namespace luf;

static GraphView view_of(const PackedGraph& p)
{
    GraphView g;
    g.N = p.N; g.B = p.B; g.E = p.E;
    g.WE = words32(p.E); g.WD = words32(p.N - p.B); g.WN = words32(p.N);
    g.DEG = p.DEG; g.adj = p.adj.data(); g.edge_uv = p.edge_uv.data();
    g.west_mask = p.west_mask.data(); g.node_long = p.node_long.data();
    g.F = p.F; g.n_classes = p.n_classes; g.parity_ok = p.parity_ok ? 1 : 0; g.leaf_bnd = 0;
    g.fresh_perm = p.perm_identity ? nullptr : p.fresh_perm.data(); g.class_end = p.class_end.data();
    return g;
}

static void export_growth(const GraphView& g, const Shot& s, uint32_t* out)
{
    std::memcpy(out, s.growth, 4u * ((g.E + 15) / 16));
}

extern "C" {

// Staged decode: syndromes in, corrections (+ growth, rounds) out.
int emu_decode_uf(const luf_graph_desc* d, int flags, long long nshots, const uint32_t* syn,
                  uint32_t* corr, uint32_t* growth2b, int32_t* steps)
{
    PackedGraph p;
    if (!pack_graph(*d, p).empty()) return -1;
    const GraphView g = view_of(p);
    const Layout l = make_layout(p.N, p.B, p.E, false, true);
    std::vector<unsigned char> mem(l.bytes);
    const Shot s = bind(l, mem.data());
    const int nl = 32, GW = (g.E + 15) / 16;
    for (long long k = 0; k < nshots; ++k) {
        LUF_PHASE(ph_reset(g, l, s, lane, nl));
        for (int w = 0; w < g.WD; ++w) s.defect[w] = syn[k * g.WD + w];
        LUF_PHASE(ph_load(g, s, lane, nl));
        const int rounds = uf_validate(g, s, flags, 0, nl);
        if (growth2b) export_growth(g, s, growth2b + k * GW);
        uf_peel(g, s, 0, nl);
        std::memcpy(corr + k * g.WE, s.corr, 4u * g.WE);
        if (steps) { steps[2 * k] = rounds; steps[2 * k + 1] = 0; }
    }
    return 0;
}

// The same with the block-per-shot instance of the phases (luf::uf_cta): a tile of `nl` lanes.
int emu_decode_uf_cta(const luf_graph_desc* d, int flags, int nl, long long nshots, const uint32_t* syn,
                      uint32_t* corr, uint32_t* growth2b, int32_t* steps)
{
    PackedGraph p;
    if (!pack_graph(*d, p).empty() || nl < 32 || nl > 1024) return -1;
    const GraphView g = view_of(p);
    const Layout l = make_layout(p.N, p.B, p.E, false, true, LUF_CTA_LIST_CAP, LUF_CTA_LIST2_CAP, LUF_CTA_CLAIM_SLOTS);
    std::vector<unsigned char> mem(l.bytes);
    const Shot s = bind(l, mem.data());
    const int GW = (g.E + 15) / 16;
    for (long long k = 0; k < nshots; ++k) {
        LUF_PHASE(uf_cta::ph_reset(g, l, s, lane, nl));
        for (int w = 0; w < g.WD; ++w) s.defect[w] = syn[k * g.WD + w];
        LUF_PHASE(uf_cta::ph_load(g, s, lane, nl));
        const int rounds = uf_cta::uf_validate_t<-1>(g, s, flags, 0, nl);
        if (growth2b) export_growth(g, s, growth2b + k * GW);
        uf_cta::uf_peel(g, s, 0, nl);
        std::memcpy(corr + k * g.WE, s.corr, 4u * g.WE);
        if (steps) { steps[2 * k] = rounds; steps[2 * k + 1] = 0; }
    }
    return 0;
}

// The wide instance of the phases (luf::uf_wide, luf_uf_wide.cuh: 64-bit node words, 32-bit indices): a tile
// of `nl` lanes, the state slab in host memory.
static uf_wide::GraphView wide_view_of(const PackedWide& p)
{
    uf_wide::GraphView g;
    g.N = p.N; g.B = p.B; g.E = p.E;
    g.WE = words32(p.E); g.WD = words32(p.N - p.B); g.WN = words32(p.N);
    g.DEG = p.DEG; g.adj = p.adj.data(); g.edge_uv = p.edge_uv.data();
    g.west_mask = p.west_mask.data(); g.node_long = p.node_long.data();
    g.F = p.F; g.n_classes = p.n_classes; g.parity_ok = p.parity_ok ? 1 : 0; g.leaf_bnd = 0;
    g.fresh_perm32 = p.perm_identity ? nullptr : p.fresh_perm32.data(); g.class_end = p.class_end.data();
    return g;
}

int emu_decode_uf_wide(const luf_graph_desc* d, int flags, int nl, long long nshots, const uint32_t* syn,
                       uint32_t* corr, uint32_t* growth2b, int32_t* steps)
{
    PackedWide p;
    if (!pack_wide(*d, p).empty() || nl < 32 || nl > 1024) return -1;
    const uf_wide::GraphView g = wide_view_of(p);
    const uf_wide::Layout l = make_wide_layout(p.N, p.B, p.E, false, true);
    std::vector<unsigned char> mem(l.bytes);
    const uf_wide::Shot s = uf_wide::bind(l, mem.data());
    const int GW = (g.E + 15) / 16;
    for (long long k = 0; k < nshots; ++k) {
        LUF_PHASE(uf_wide::ph_reset(g, l, s, lane, nl));
        for (int w = 0; w < g.WD; ++w) s.defect[w] = syn[k * g.WD + w];
        LUF_PHASE(uf_wide::ph_load(g, s, lane, nl));
        const int rounds = uf_wide::uf_validate_t<-1>(g, s, flags, 0, nl);
        if (growth2b) std::memcpy(growth2b + k * GW, s.growth, 4u * GW);
        uf_wide::uf_peel(g, s, 0, nl);
        std::memcpy(corr + k * g.WE, s.corr, 4u * g.WE);
        if (steps) { steps[2 * k] = rounds; steps[2 * k + 1] = 0; }
    }
    return 0;
}

// Fused shot loop of the wide instance; exports as emu_mc_uf.  parity_only: the logical bit through
// uf_peel_parity (what k_mc_uf_wide runs) instead of the full peel.
int emu_mc_uf_wide(const luf_graph_desc* d, int flags, int nl, int parity_only, const double* class_p, unsigned long long seed,
                   unsigned long long shot0, long long nshots, unsigned long long counts[2],
                   uint32_t* err_out, uint32_t* syn_out, uint32_t* corr_out, uint8_t* logical_out)
{
    PackedWide p;
    PackedFast pf;
    if (!pack_wide(*d, p).empty() || nl < 32 || nl > 1024) return -1;
    pack_fast(*d, pf);
    uf_wide::GraphView g = wide_view_of(p);
    g.leaf_bnd = pf.ok || (p.parity_ok && [&] { for (int v = 0; v < d->n_boundary; ++v) if (d->inc_off[v + 1] - d->inc_off[v] != 1) return false; return true; }()) ? 1 : 0;
    const uf_wide::Layout l = make_wide_layout(p.N, p.B, p.E, true, true);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes);
    const uf_wide::Shot s = uf_wide::bind(l, mem.data());
    unsigned long long fails = 0;
    for (long long k = 0; k < nshots; ++k) {
        uint32_t parity = 0;
        LUF_PHASE(uf_wide::ph_reset(g, l, s, lane, nl));
        LUF_PHASE(if (lane < 32) parity ^= uf_wide::ph_sample(g, s, sp, shot0 + k, lane, 32));
        if (err_out) std::memcpy(err_out + k * g.WE, s.err, 4u * g.WE);
        if (syn_out) std::memcpy(syn_out + k * g.WD, s.defect, 4u * g.WD);
        LUF_PHASE(uf_wide::ph_load(g, s, lane, nl));
        uf_wide::uf_validate_t<-1>(g, s, flags, 0, nl);
        parity ^= parity_only ? uf_wide::uf_peel_parity(g, s, 0, nl) : uf_wide::uf_peel(g, s, 0, nl);
        if (corr_out) std::memcpy(corr_out + k * g.WE, s.corr, 4u * g.WE);
        if (logical_out) logical_out[k] = (uint8_t)(parity & 1u);
        fails += parity & 1u;
    }
    counts[0] += fails; counts[1] += (unsigned long long)nshots;
    return 0;
}

// Fixed-weight sampler (ph_force): errors and syndromes of shots [shot0, shot0 + nshots).
int emu_sample_weight(const luf_graph_desc* d, int weight, unsigned long long seed, unsigned long long shot0,
                      long long nshots, uint32_t* err_out, uint32_t* syn_out)
{
    PackedGraph p;
    if (!pack_graph(*d, p).empty()) return -1;
    const GraphView g = view_of(p);
    const Layout l = make_layout(p.N, p.B, p.E, true, true);
    const double zero[MAX_CLASSES] = {0};
    const SampleParams sp = make_sample_params(zero, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes);
    const Shot s = bind(l, mem.data());
    const int nl = 32;
    for (long long k = 0; k < nshots; ++k) {
        LUF_PHASE(ph_reset(g, l, s, lane, nl));
        LUF_PHASE(ph_force(g, s, sp, (uint32_t)weight, shot0 + k, lane));
        std::memcpy(err_out + k * g.WE, s.err, 4u * g.WE);
        std::memcpy(syn_out + k * g.WD, s.defect, 4u * g.WD);
    }
    return 0;
}

// Fused Monte-Carlo shot loop; optionally exports the sampled errors, their
// syndromes and the corrections so that the oracle can re-decode them.
int emu_mc_uf(const luf_graph_desc* d, int flags, const double* class_p, unsigned long long seed,
              unsigned long long shot0, long long nshots, unsigned long long counts[2],
              uint32_t* err_out, uint32_t* syn_out, uint32_t* corr_out)
{
    PackedGraph p;
    if (!pack_graph(*d, p).empty()) return -1;
    const GraphView g = view_of(p);
    const Layout l = make_layout(p.N, p.B, p.E, true, true);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes);
    const Shot s = bind(l, mem.data());
    const int nl = 32;
    unsigned long long fails = 0;
    for (long long k = 0; k < nshots; ++k) {
        uint32_t parity = 0;
        LUF_PHASE(ph_reset(g, l, s, lane, nl));
        LUF_PHASE(parity ^= ph_sample(g, s, sp, shot0 + k, lane, nl));
        if (err_out) std::memcpy(err_out + k * g.WE, s.err, 4u * g.WE);
        if (syn_out) std::memcpy(syn_out + k * g.WD, s.defect, 4u * g.WD);
        LUF_PHASE(ph_load(g, s, lane, nl));
        uf_validate(g, s, flags, 0, nl);
        parity ^= uf_peel(g, s, 0, nl);
        if (corr_out) std::memcpy(corr_out + k * g.WE, s.corr, 4u * g.WE);
        fails += parity & 1u;
    }
    counts[0] += fails; counts[1] += (unsigned long long)nshots;
    return 0;
}

// The fused kernels' parity-only peel (uf_peel_parity) next to the full peel on the same sampled shots:
// counts[0] += failures by the full peel, [1] += failures by the parity path, [2] += shots on which the two
// differ, [3] += shots that hold a cluster with boundary nodes of both sides (they take the full peel).
// Whether pack_graph found the graph fit for the parity-only peel.
int emu_parity_ok(const luf_graph_desc* d)
{
    PackedGraph p;
    if (!pack_graph(*d, p).empty()) return -1;
    return p.parity_ok ? 1 : 0;
}

int emu_mc_uf_parity(const luf_graph_desc* d, int flags, const double* class_p, unsigned long long seed,
                     unsigned long long shot0, long long nshots, unsigned long long counts[4], int nl)
{
    const bool cta = nl != 32;
    PackedGraph p;
    if (!pack_graph(*d, p).empty()) return -1;
    const GraphView g = view_of(p);
    const Layout l = cta ? make_layout(p.N, p.B, p.E, true, true, LUF_CTA_LIST_CAP, LUF_CTA_LIST2_CAP, LUF_CTA_CLAIM_SLOTS)
                         : make_layout(p.N, p.B, p.E, true, true);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes), saved;
    const Shot s = bind(l, mem.data());
    for (long long k = 0; k < nshots; ++k) {
        uint32_t parity = 0;
        if (cta) { LUF_PHASE(uf_cta::ph_reset(g, l, s, lane, nl)); } else { LUF_PHASE(ph_reset(g, l, s, lane, nl)); }
        LUF_PHASE(if (lane < 32) parity ^= ph_sample(g, s, sp, shot0 + k, lane, 32));
        if (cta) { LUF_PHASE(uf_cta::ph_load(g, s, lane, nl)); } else { LUF_PHASE(ph_load(g, s, lane, nl)); }
        uint32_t fast, full;
        if (cta) {
            uf_cta::uf_validate(g, s, flags, 0, nl);
            counts[3] += s.cnt[4] ? 1 : 0;
            if (!g.parity_ok) saved = mem;                                    // the guard sends every shot to the (destructive) full peel
            fast = s.cnt[4] ? 2u : uf_cta::uf_peel_parity(g, s, 0, nl);       // read-only on unmixed shots
            if (!g.parity_ok) mem = saved;
            full = uf_cta::uf_peel(g, s, 0, nl);
        } else {
            uf_validate(g, s, flags, 0, nl);
            counts[3] += s.cnt[4] ? 1 : 0;
            if (!g.parity_ok) saved = mem;
            fast = s.cnt[4] ? 2u : uf_peel_parity(g, s, 0, nl);
            if (!g.parity_ok) mem = saved;
            full = uf_peel(g, s, 0, nl);
        }
        if (fast == 2u) fast = full;
        counts[0] += (parity ^ full) & 1u; counts[1] += (parity ^ fast) & 1u; counts[2] += (fast ^ full) & 1u;
    }
    return 0;
}

// Canonical validate + the pass over the defects (uf_peel_parity with g.leaf_bnd) next to the full peel on EVERY
// shot, also those with a cluster on both sides: counts[0] += failures by the full peel, [1] += failures by the
// pass over the defects, [2] += shots where they differ, [3] += shots with a cluster on both sides; -3 when the
// graph has a boundary node with more than one edge (pack_fast).
int emu_mc_uf_sides(const luf_graph_desc* d, int flags, const double* class_p, unsigned long long seed,
                    unsigned long long shot0, long long nshots, unsigned long long counts[4], int nl)
{
    const bool cta = nl != 32;
    PackedGraph p;
    PackedFast pf;
    if (!pack_graph(*d, p).empty()) return -1;
    pack_fast(*d, pf);
    if (!pf.ok) return -3;
    GraphView g = view_of(p);
    g.leaf_bnd = 1;
    const Layout l = cta ? make_layout(p.N, p.B, p.E, true, true, LUF_CTA_LIST_CAP, LUF_CTA_LIST2_CAP, LUF_CTA_CLAIM_SLOTS)
                         : make_layout(p.N, p.B, p.E, true, true);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes);
    const Shot s = bind(l, mem.data());
    for (long long k = 0; k < nshots; ++k) {
        uint32_t parity = 0, sides, full;
        if (cta) { LUF_PHASE(uf_cta::ph_reset(g, l, s, lane, nl)); } else { LUF_PHASE(ph_reset(g, l, s, lane, nl)); }
        LUF_PHASE(if (lane < 32) parity ^= ph_sample(g, s, sp, shot0 + k, lane, 32));
        if (cta) { LUF_PHASE(uf_cta::ph_load(g, s, lane, nl)); } else { LUF_PHASE(ph_load(g, s, lane, nl)); }
        if (cta) {
            uf_cta::uf_validate(g, s, flags, 0, nl);
            counts[3] += s.cnt[4] ? 1 : 0;
            sides = uf_cta::uf_peel_parity(g, s, 0, nl);         // read-only with leaf_bnd
            full = uf_cta::uf_peel(g, s, 0, nl);
        } else {
            uf_validate(g, s, flags, 0, nl);
            counts[3] += s.cnt[4] ? 1 : 0;
            sides = uf_peel_parity(g, s, 0, nl);
            full = uf_peel(g, s, 0, nl);
        }
        counts[0] += (parity ^ full) & 1u; counts[1] += (parity ^ sides) & 1u; counts[2] += (sides ^ full) & 1u;
    }
    return 0;
}

// The fused kernels' order-free merge (uf_validate_fast + ph_defect_sides) next to the canonical decode on the
// same sampled shots: counts[0] += failures canonical, [1] += failures by the fast path (shots it flags as
// mixed take the canonical result, as in the kernel), [2] += shots where the two differ, [3] += mixed shots,
// [4] += shots whose numbers of growth rounds differ.  flags: west (2), node (4), node + bucket (12) and combinations.
#define LUF_EMU_FAST(F) (flags == 0 ? F<0>(g, s, 0, nl) : flags == 2 ? F<2>(g, s, 0, nl) : flags == 4 ? F<4>(g, s, 0, nl) : \
                         flags == 6 ? F<6>(g, s, 0, nl) : flags == 12 ? F<12>(g, s, 0, nl) : F<14>(g, s, 0, nl))
int emu_mc_uf_fast(const luf_graph_desc* d, int flags, const double* class_p, unsigned long long seed,
                   unsigned long long shot0, long long nshots, unsigned long long counts[5], int nl)
{
    const bool cta = nl != 32;
    PackedGraph p;
    if (!pack_graph(*d, p).empty() || (flags & 1) || (flags & 12) == 8) return -1;
    const GraphView g = view_of(p);
    const Layout l = cta ? make_layout(p.N, p.B, p.E, true, true, LUF_CTA_LIST_CAP, LUF_CTA_LIST2_CAP, LUF_CTA_CLAIM_SLOTS)
                         : make_layout(p.N, p.B, p.E, true, true);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes);
    const Shot s = bind(l, mem.data());
    for (long long k = 0; k < nshots; ++k) {
        uint32_t pf = 0, pc = 0;
        int rf, rc;
        bool mixed;
        if (cta) {
            LUF_PHASE(uf_cta::ph_reset(g, l, s, lane, nl));
            LUF_PHASE(if (lane < 32) pf ^= ph_sample(g, s, sp, shot0 + k, lane, 32));
            LUF_PHASE(uf_cta::ph_load(g, s, lane, nl));
            rf = LUF_EMU_FAST(uf_cta::uf_validate_fast);
            mixed = s.cnt[4] != 0;
            LUF_PHASE(pf ^= uf_cta::ph_defect_sides(g, s, lane, nl));
            LUF_PHASE(uf_cta::ph_reset(g, l, s, lane, nl));
            LUF_PHASE(if (lane < 32) pc ^= ph_sample(g, s, sp, shot0 + k, lane, 32));
            LUF_PHASE(uf_cta::ph_load(g, s, lane, nl));
            rc = uf_cta::uf_validate(g, s, flags, 0, nl);
            pc ^= uf_cta::uf_peel(g, s, 0, nl);
        } else {
            LUF_PHASE(ph_reset(g, l, s, lane, nl));
            LUF_PHASE(pf ^= ph_sample(g, s, sp, shot0 + k, lane, nl));
            LUF_PHASE(ph_load(g, s, lane, nl));
            rf = LUF_EMU_FAST(uf_validate_fast);
            mixed = s.cnt[4] != 0;
            LUF_PHASE(pf ^= ph_defect_sides(g, s, lane, nl));
            LUF_PHASE(ph_reset(g, l, s, lane, nl));
            LUF_PHASE(pc ^= ph_sample(g, s, sp, shot0 + k, lane, nl));
            LUF_PHASE(ph_load(g, s, lane, nl));
            rc = uf_validate(g, s, flags, 0, nl);
            pc ^= uf_peel(g, s, 0, nl);
        }
        if (mixed) pf = pc;
        counts[0] += pc & 1u; counts[1] += pf & 1u; counts[2] += (pf ^ pc) & 1u; counts[3] += mixed; counts[4] += rf != rc;
    }
    return 0;
}

// The compact fast path of the fused kernels (luf_uf_fast.cuh) with a tile of `nl` lanes: out[k] = 0 / 1 (the
// logical bit) or 2 (HARD: a cluster touches both sides under the default inclination -- the kernel queues such
// shots for the canonical path; never under the west inclination); stats (optional) [nshots][4] = rounds, touched nodes, most newly FULL edges of a round,
// most active nodes of a round.  mu < 0: list capacities from the class probabilities, as the library does.
int emu_mc_uf_compact(const luf_graph_desc* d, int west, const double* class_p, unsigned long long seed, unsigned long long shot0,
                      long long nshots, int nl, double mu, uint8_t* out, uint32_t* stats)
{
    PackedFast pf;
    pack_fast(*d, pf);
    if (!pf.ok) return -2;
    struct { int E, n_classes; const std::vector<uint32_t>& class_end; } p = {pf.E, pf.n_classes, pf.class_end};
    if (mu < 0.0) {
        mu = 0.0;
        for (int c = 0, lo = 0; c < p.n_classes; ++c) { mu += class_p[c] * ((int)p.class_end[c] - lo); lo = (int)p.class_end[c]; }
    }
    fast::FView g;
    g.D = pf.D; g.E = pf.E; g.WD = words32(pf.D); g.DEG = pf.DEG; g.deg_magic = (uint32_t)(0x100000000ull / (uint64_t)pf.DEG) + 1u; g.fadj = pf.fadj.data(); g.einfo = pf.einfo.data();
    g.ebits = pf.ebits; g.emask = pf.emask; g.omask = pf.omask;
    g.F = pf.F; g.n_classes = pf.n_classes;
    g.fresh_perm = nullptr; g.fresh_perm32 = pf.perm_identity ? nullptr : pf.fresh_perm32.data(); g.class_end = pf.class_end.data();
    g.west = west;
    const fast::FLayout l = make_fast_layout(pf.D, p.E, pf.DEG, mu);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes + 16);
    const fast::FShot s = fast::fbind(l, mem.data());
    for (long long k = 0; k < nshots; ++k)
        out[k] = (uint8_t)fast::fast_shot<false>(g, l, s, sp, shot0 + k, 0, nl, 0, stats ? stats + 4 * k : nullptr);
    return 0;
}

// The compact fast path with the merge log (default inclination): a shot with odd clusters on both sides exports
// their merges and hard_resolve replays them in canonical order -- out[k] = the logical bit, or 2 when the log or
// the export buffer could not hold the shot (the kernel then queues it for the canonical launch).  nkeys
// (optional) [nshots]: exported keys (0: the shot had no such cluster).  log_scale < 0: the library's log size.
int emu_mc_uf_compact_log(const luf_graph_desc* d, const double* class_p, unsigned long long seed, unsigned long long shot0,
                          long long nshots, int nl, double mu, int lcap, int onepass, uint8_t* out, uint32_t* nkeys, uint32_t* nlog)
{
    PackedFast pf;
    pack_fast(*d, pf);
    if (!pf.ok) return -2;
    if (mu < 0.0) {
        mu = 0.0;
        for (int c = 0, lo = 0; c < pf.n_classes; ++c) { mu += class_p[c] * ((int)pf.class_end[c] - lo); lo = (int)pf.class_end[c]; }
    }
    fast::FView g;
    g.D = pf.D; g.E = pf.E; g.WD = words32(pf.D); g.DEG = pf.DEG; g.deg_magic = (uint32_t)(0x100000000ull / (uint64_t)pf.DEG) + 1u; g.fadj = pf.fadj.data(); g.einfo = pf.einfo.data();
    g.ebits = pf.ebits; g.emask = pf.emask; g.omask = pf.omask;
    g.F = pf.F; g.n_classes = pf.n_classes;
    g.fresh_perm = nullptr; g.fresh_perm32 = pf.perm_identity ? nullptr : pf.fresh_perm32.data(); g.class_end = pf.class_end.data();
    g.west = 0;
    fast::FLayout l = make_fast_layout(pf.D, pf.E, pf.DEG, mu, true);
    if (lcap >= 0) {                 // a log of the caller's size (rounded up to a power of two) behind the state
        uint32_t p2 = 1; while (p2 < (uint32_t)lcap) p2 <<= 1;
        l.lcap = p2; l.elog = l.bytes; l.bytes += 4u * p2 + 16u;
    }
    const SampleParams sp = make_sample_params(class_p, pf.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes + 16);
    fast::FShot s = fast::fbind(l, mem.data());
    std::vector<uint32_t> glog(l.lcap + 1), sortbuf;
    if (onepass) s.elog = glog.data();               // one-pass mode: the log outside the tile's state (global memory on the device)
    std::vector<uint32_t> xbuf((size_t)pf.E + 8), S(pf.D, 0xDEADBEEFu);
    uint64_t cursor = 0;
    uint32_t q[2], qn = 0, res = 0;
    for (long long k = 0; k < nshots; ++k) {
        cursor = 0;
        fast::FExport x{xbuf.data(), &cursor, (uint32_t)xbuf.size(), q, &qn};
        x.unsorted = onepass ? 1u : 0u; x.max_keys = (uint32_t)pf.E;
        uint32_t r = fast::fast_shot<false>(g, l, s, sp, shot0 + k, 0, nl, 0, nullptr, &x);
        if (nkeys) nkeys[k] = 0;
        if (nlog) nlog[k] = s.cnt[fast::C_MLIST];
        if (r == fast::HARD_LOGGED) {
            const uint32_t base = s.cnt[fast::C_XBASE];
            if (nkeys) nkeys[k] = xbuf[base] & 0x7FFFFFFFu;
            if (onepass) {
                uint32_t np = 1;
                while (np < (xbuf[base] & 0x7FFFFFFFu)) np <<= 1;
                sortbuf.assign(np, 0);
                fast::hard_sort(g, xbuf.data() + base, sortbuf.data(), 0, nl > 32 ? 256 : 32, 0);
            }
            r = fast::hard_resolve(xbuf.data() + base, S.data(), &res, 0, 32, 0);
        }
        out[k] = (uint8_t)r;
    }
    return 0;
}

// Staged sampler on the compact tables (f_sample_record): errors and syndromes of shots [shot0, shot0 + nshots).
int emu_sample_compact(const luf_graph_desc* d, const double* class_p, unsigned long long seed, unsigned long long shot0,
                       long long nshots, uint32_t* err_out, uint32_t* syn_out)
{
    PackedFast pf;
    pack_fast(*d, pf);
    if (!pf.ok) return -2;
    fast::FView g;
    g.D = pf.D; g.E = pf.E; g.WD = words32(pf.D); g.DEG = pf.DEG; g.deg_magic = (uint32_t)(0x100000000ull / (uint64_t)pf.DEG) + 1u; g.fadj = pf.fadj.data(); g.einfo = pf.einfo.data();
    g.ebits = pf.ebits; g.emask = pf.emask; g.omask = pf.omask; g.F = pf.F; g.n_classes = pf.n_classes;
    g.fresh_perm = nullptr; g.fresh_perm32 = pf.perm_identity ? nullptr : pf.fresh_perm32.data(); g.class_end = pf.class_end.data();
    g.west = 0;
    const SampleParams sp = make_sample_params(class_p, pf.n_classes, seed);
    const int WE = words32(pf.E), nl = 32;
    for (long long k = 0; k < nshots; ++k) {
        std::memset(err_out + k * WE, 0, 4u * WE);
        std::memset(syn_out + k * g.WD, 0, 4u * g.WD);
        LUF_PHASE(fast::f_sample_record(g, err_out + k * WE, syn_out + k * g.WD, sp, shot0 + k, lane, nl));
    }
    return 0;
}

// Macar (decoder 1) / Actis (decoder 2): syndromes in, (tSV, tBP), growth after
// validation and correction out.
int emu_decode_ca_tile(const luf_graph_desc* d, int decoder, int flags, long long nshots, const uint32_t* syn,
                       uint32_t* corr, uint32_t* growth2b, int32_t* steps, int nl)
{
    const bool cta = nl != 32;
    PackedCa p;
    if (!pack_ca(*d, p).empty()) return -1;
    const bool actis = decoder == 2;
    if (actis && !p.actis_ok) return -2;
    std::vector<uint32_t> west(d->west_edge_mask, d->west_edge_mask + words32(p.E));
    CaView g;
    g.N = p.N; g.B = p.B; g.E = p.E; g.K = p.K;
    g.WE = words32(p.E); g.WD = words32(p.N - p.B); g.WN = words32(p.N); g.GW = (p.E + 15) / 16;
    g.nbr = p.nbr.data(); g.west_mask = west.data(); g.span = p.span.data(); g.relay = p.relay.data();
    g.global_span = p.global_span; g.nbr_shared = 0;
    const CaLayout l = cta ? make_ca_layout(p.N, p.B, p.E, actis, false, CA_CTA_TLIST_CAP) : make_ca_layout(p.N, p.B, p.E, actis, false);
    std::vector<unsigned char> mem(l.bytes);
    const CaShot s = ca_bind(l, mem.data());
    for (long long k = 0; k < nshots; ++k) {
        if (actis) { LUF_PHASE(ca_reset<true>(g, s, lane, nl)); } else { LUF_PHASE(ca_reset<false>(g, s, lane, nl)); }
        for (int w = 0; w < g.WD; ++w) s.defect[w] = syn[k * g.WD + w];
        LUF_PHASE(ca_load(g, s, lane, nl));
        int tsv = 0, tbp = 0;
        uint32_t* gout = growth2b ? growth2b + k * g.GW : nullptr;
        if (cta) { if (actis) ca_cta::ca_decode<true>(g, s, flags, &tsv, &tbp, gout, 0, nl); else ca_cta::ca_decode<false>(g, s, flags, &tsv, &tbp, gout, 0, nl); }
        else { if (actis) ca_warp::ca_decode<true>(g, s, flags, &tsv, &tbp, gout, 0, nl); else ca_warp::ca_decode<false>(g, s, flags, &tsv, &tbp, gout, 0, nl); }
        std::memcpy(corr + k * g.WE, s.corr, 4u * g.WE);
        if (steps) { steps[2 * k] = tsv; steps[2 * k + 1] = tbp; }
    }
    return 0;
}

int emu_decode_ca(const luf_graph_desc* d, int decoder, int flags, long long nshots, const uint32_t* syn,
                  uint32_t* corr, uint32_t* growth2b, int32_t* steps)
{
    return emu_decode_ca_tile(d, decoder, flags, nshots, syn, corr, growth2b, steps, 32);
}

// Forward scheme on given fresh errors (one stream); decoder 0 = UF, 1 = Macar.
int emu_fw_stream(const luf_graph_desc* gd, const luf_forward_desc* fd, int decoder, int flags, int ncycles,
                  const uint32_t* init_buf, const uint32_t* fresh, int32_t* steps, uint32_t* commit_out, int32_t* m_out)
{
    PackedGraph pg; PackedCa pc; PackedForward pf;
    if (!pack_graph(*gd, pg).empty()) return -1;
    if (decoder == 1 && !pack_ca(*gd, pc).empty()) return -2;
    const luf_graph_desc_dims dims = {pg.N, pg.B, pg.E, pg.n_classes};
    if (!pack_forward(dims, *fd, pf).empty()) return -3;
    const GraphView gv = view_of(pg);
    FwView g;
    g.N = pg.N; g.B = pg.B; g.E = pg.E; g.WE = words32(pg.E); g.WD = words32(pg.N - pg.B); g.c = pf.c;
    g.edge_uv = pg.edge_uv.data(); g.lower_edge = pf.lower_edge.data(); g.commit_mask = pf.commit_mask.data();
    g.art_node = pf.art_node.data(); g.fb_mask = pf.fb_mask.data(); g.node_lower = pf.node_lower.data();
    g.node_side = pf.node_side.data();
    const int nl = 32;
    if (decoder == 0) {
        const Layout ul = make_layout(pg.N, pg.B, pg.E, true, true);
        const FwLayout fl = make_fw_layout(pg.N, pg.E, ul.bytes);
        std::vector<unsigned char> mem(fl.bytes);
        const FwState f = fw_bind(fl, mem.data());
        FwUf d{gv, ul, bind(ul, mem.data()), flags};
        LUF_PHASE(fw_reset(g, f, init_buf, lane, nl));
        for (int c = 0; c < ncycles; ++c)
            m_out[c] = fw_cycle_given(g, f, d, fresh + (size_t)c * g.WE, commit_out + (size_t)c * g.WE,
                                      steps[2 * c], steps[2 * c + 1], 0, nl);
    } else {
        std::vector<uint32_t> west(gd->west_edge_mask, gd->west_edge_mask + words32(pg.E));
        CaView cv;
        cv.N = pc.N; cv.B = pc.B; cv.E = pc.E; cv.K = pc.K;
        cv.WE = words32(pc.E); cv.WD = words32(pc.N - pc.B); cv.WN = words32(pc.N); cv.GW = (pc.E + 15) / 16;
        cv.nbr = pc.nbr.data(); cv.west_mask = west.data(); cv.span = pc.span.data(); cv.relay = pc.relay.data();
        cv.global_span = pc.global_span; cv.nbr_shared = 0;
        const CaLayout cl = make_ca_layout(pc.N, pc.B, pc.E, false, true);
        const FwLayout fl = make_fw_layout(pg.N, pg.E, cl.bytes);
        std::vector<unsigned char> mem(fl.bytes);
        const FwState f = fw_bind(fl, mem.data());
        FwMacar d{cv, ca_bind(cl, mem.data())};
        LUF_PHASE(fw_reset(g, f, init_buf, lane, nl));
        for (int c = 0; c < ncycles; ++c)
            m_out[c] = fw_cycle_given(g, f, d, fresh + (size_t)c * g.WE, commit_out + (size_t)c * g.WE,
                                      steps[2 * c], steps[2 * c + 1], 0, nl);
    }
    return 0;
}

// Forward scheme on the wide instance (luf_fw_wide.cuh), UF, one stream, a tile of `nl` lanes.
int emu_fw_stream_wide(const luf_graph_desc* gd, const luf_forward_desc* fd, int flags, int nl, int ncycles,
                       const uint32_t* init_buf, const uint32_t* fresh, int32_t* steps, uint32_t* commit_out, int32_t* m_out)
{
    PackedWide pg; PackedForwardWide pf;
    if (!pack_wide(*gd, pg).empty()) return -1;
    const luf_graph_desc_dims dims = {pg.N, pg.B, pg.E, pg.n_classes};
    if (!pack_forward_wide(dims, *fd, pf).empty()) return -3;
    const uf_wide::GraphView gv = wide_view_of(pg);
    fw_wide::FwView g;
    g.N = pg.N; g.B = pg.B; g.E = pg.E; g.WE = words32(pg.E); g.WD = words32(pg.N - pg.B); g.c = pf.c;
    g.edge_uv = pg.edge_uv.data(); g.lower_edge = pf.lower_edge.data(); g.commit_mask = pf.commit_mask.data();
    g.art_node = pf.art_node.data(); g.fb_mask = pf.fb_mask.data(); g.node_lower = pf.node_lower.data();
    g.node_side = pf.node_side.data();
    const uf_wide::Layout ul = make_wide_layout(pg.N, pg.B, pg.E, true, true);
    const fw_wide::FwLayout fl = make_fw_wide_layout(pg.N, pg.E, ul.bytes);
    std::vector<unsigned char> mem(fl.bytes);
    const fw_wide::FwState f = fw_wide::fw_bind(fl, mem.data());
    fw_wide::FwUf d{gv, ul, uf_wide::bind(ul, mem.data()), flags};
    LUF_PHASE(fw_wide::fw_reset(g, f, init_buf, lane, nl));
    for (int c = 0; c < ncycles; ++c)
        m_out[c] = fw_wide::fw_cycle_given(g, f, d, fresh + (size_t)c * g.WE, commit_out + (size_t)c * g.WE,
                                           steps[2 * c], steps[2 * c + 1], 0, nl);
    return 0;
}

// Snowflake + frugal bookkeeping on given fresh errors (one stream).
int emu_sf_stream_metrics(const luf_window_desc* d, int flags, int ncycles, const uint32_t* fresh,
                          int32_t* rt, uint32_t* floor_bits, int32_t* m_out, uint32_t* mgrowth, int32_t* mins,
                          const uint8_t* drop_only, int nl)
{
    const bool cta = nl != 32;
    PackedWindow p;
    if (!pack_window(*d, p).empty()) return -1;
    SfView g;
    g.d = p.d; g.h = p.h; g.NLb = p.NLb; g.NLd = p.NLd; g.NL = p.NL; g.EL = p.EL; g.K = p.K;
    g.ELw = words32(p.EL); g.WLn = words32(p.NL); g.N = p.NL * p.h; g.n_classes = p.n_classes;
    g.nbr = p.nbr.data(); g.edge_ends = p.edge_ends.data(); g.pos_long = p.pos_long.data();
    g.fresh_perm = p.fresh_perm.data(); g.class_end = p.class_end.data();
    const SfLayout l = cta ? make_sf_layout(p, SF_CTA_ALIST_CAP) : make_sf_layout(p);
    std::vector<unsigned char> mem(l.bytes);
    const SfShot s = sf_bind(l, mem.data());
    if (cta) sf_cta::sf_decode_stream(g, s, flags, ncycles, fresh, rt, floor_bits, m_out, mgrowth, mins, drop_only, 0, nl);
    else sf_warp::sf_decode_stream(g, s, flags, ncycles, fresh, rt, floor_bits, m_out, mgrowth, mins, drop_only, 0, nl);
    return 0;
}

int emu_sf_stream_tile(const luf_window_desc* d, int flags, int ncycles, const uint32_t* fresh,
                       int32_t* rt, uint32_t* floor_bits, int32_t* m_out, int nl)
{
    return emu_sf_stream_metrics(d, flags, ncycles, fresh, rt, floor_bits, m_out, nullptr, nullptr, nullptr, nl);
}

// Frugal.sample_latency streams with the device sampler (host arithmetic): lat [nstreams][3].
int emu_sf_latency(const luf_window_desc* d, int flags, const double* class_p, unsigned long long seed,
                   unsigned long long stream0, long long nstreams, int32_t* lat, int nl)
{
    const bool cta = nl != 32;
    PackedWindow p;
    if (!pack_window(*d, p).empty()) return -1;
    SfView g;
    g.d = p.d; g.h = p.h; g.NLb = p.NLb; g.NLd = p.NLd; g.NL = p.NL; g.EL = p.EL; g.K = p.K;
    g.ELw = words32(p.EL); g.WLn = words32(p.NL); g.N = p.NL * p.h; g.n_classes = p.n_classes;
    g.nbr = p.nbr.data(); g.edge_ends = p.edge_ends.data(); g.pos_long = p.pos_long.data();
    g.fresh_perm = p.fresh_perm.data(); g.class_end = p.class_end.data();
    const SfLayout l = cta ? make_sf_layout(p, SF_CTA_ALIST_CAP) : make_sf_layout(p);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes);
    const SfShot s = sf_bind(l, mem.data());
    for (long long k = 0; k < nstreams; ++k) {
        int v[3];
        if (cta) sf_cta::sf_latency(g, s, sp, flags, stream0 + k, v, 0, nl);
        else sf_warp::sf_latency(g, s, sp, flags, stream0 + k, v, 0, nl);
        lat[3 * k] = v[0]; lat[3 * k + 1] = v[1]; lat[3 * k + 2] = v[2];
    }
    return 0;
}

int emu_sf_stream(const luf_window_desc* d, int flags, int ncycles, const uint32_t* fresh,
                  int32_t* rt, uint32_t* floor_bits, int32_t* m_out)
{
    return emu_sf_stream_tile(d, flags, ncycles, fresh, rt, floor_bits, m_out, 32);
}

// Frugal.run(n) streams with the device sampler (host arithmetic).
int emu_sf_mc_tile(const luf_window_desc* d, int flags, const double* class_p, unsigned long long seed,
                   unsigned long long stream0, long long nstreams, int n, unsigned long long counts[4], int32_t* steps,
                   int nl)
{
    const bool cta = nl != 32;
    PackedWindow p;
    if (!pack_window(*d, p).empty()) return -1;
    SfView g;
    g.d = p.d; g.h = p.h; g.NLb = p.NLb; g.NLd = p.NLd; g.NL = p.NL; g.EL = p.EL; g.K = p.K;
    g.ELw = words32(p.EL); g.WLn = words32(p.NL); g.N = p.NL * p.h; g.n_classes = p.n_classes;
    g.nbr = p.nbr.data(); g.edge_ends = p.edge_ends.data(); g.pos_long = p.pos_long.data();
    g.fresh_perm = p.fresh_perm.data(); g.class_end = p.class_end.data();
    const SfLayout l = cta ? make_sf_layout(p, SF_CTA_ALIST_CAP) : make_sf_layout(p);
    const SampleParams sp = make_sample_params(class_p, p.n_classes, seed);
    std::vector<unsigned char> mem(l.bytes);
    const SfShot s = sf_bind(l, mem.data());
    for (long long k = 0; k < nstreams; ++k) {
        int m = 0; long long cyc = 0, ra = 0, ru = 0;
        int32_t* st = steps ? steps + (size_t)k * n * 2 : nullptr;
        if (cta) sf_cta::sf_run(g, s, sp, flags, stream0 + k, n, m, cyc, ra, ru, st, nullptr, nullptr, 0, nl);
        else sf_warp::sf_run(g, s, sp, flags, stream0 + k, n, m, cyc, ra, ru, st, nullptr, nullptr, 0, nl);
        counts[0] += m; counts[1] += cyc; counts[2] += ra; counts[3] += ru;
    }
    return 0;
}

int emu_sf_mc(const luf_window_desc* d, int flags, const double* class_p, unsigned long long seed,
              unsigned long long stream0, long long nstreams, int n, unsigned long long counts[4], int32_t* steps)
{
    return emu_sf_mc_tile(d, flags, class_p, seed, stream0, nstreams, n, counts, steps, 32);
}

}  // extern "C"
