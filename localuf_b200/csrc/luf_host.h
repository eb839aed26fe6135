// luf_host.h -- host-side preparation shared by the CUDA library and by the
// host emulation harness of the kernel phases (tests/emu): packing the flat
// graph description into the tables of luf::GraphView and laying out one
// shot's state.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/luf_b200.h"
#include "luf_core.cuh"
#include "luf_uf_fast.cuh"
#include "luf_uf_wide.cuh"
#include "luf_ca.cuh"
#include "luf_sf.cuh"
#include "luf_fw.cuh"
#include "luf_fw_wide.cuh"

namespace luf {

inline int words32(int nbits) { return (nbits + 31) / 32; }

struct luf_graph_desc_dims { int N, B, E, n_classes; };

struct PackedGraph {
    int N = 0, B = 0, E = 0, F = 0, n_classes = 0;
    int DEG = 0;
    bool perm_identity = false;
    bool parity_ok = true;        // see GraphView::parity_ok
    std::vector<uint32_t> adj, edge_uv, west_mask, class_end;
    std::vector<int16_t> node_long;
    std::vector<uint16_t> fresh_perm;
};

// Returns "" on success, else why the graph cannot be packed.
inline std::string pack_graph(const luf_graph_desc& d, PackedGraph& out)
{
    const int N = d.n_nodes, B = d.n_boundary, E = d.n_edges;
    if (N <= 0 || E <= 0 || B < 0 || B > N) return "empty or inconsistent graph";
    if (N > 32767 || E > 0xFFFD)
        return "graph too large for the 16-bit state layout (N <= 32767, E <= 65533)";
    if (d.n_classes < 1 || d.n_classes > MAX_CLASSES) return "n_classes out of range";
    out.N = N; out.B = B; out.E = E; out.n_classes = d.n_classes;
    out.edge_uv.assign(E, 0);
    out.node_long.assign(N, 0);
    for (int e = 0; e < E; ++e) {
        const int u = d.edge_u[e], v = d.edge_v[e];
        if (u < 0 || u >= N || v < 0 || v >= N) return "edge endpoint out of range";
        out.edge_uv[e] = (uint32_t)u | ((uint32_t)v << 16);
    }
    for (int v = 0; v < N; ++v) {
        if (((d.node_flags[v] & 1) != 0) != (v < B)) return "boundary nodes must be node indices [0, B)";
        out.node_long[v] = (int16_t)d.node_long[v];
    }
    if (d.inc_off[N] != 2 * E) return "CSR size mismatch";
    int deg = 2;
    for (int v = 0; v < N; ++v) deg = std::max(deg, d.inc_off[v + 1] - d.inc_off[v]);
    out.DEG = (deg + 1) & ~1;                     // rows of adj are read as 64-bit pairs
    out.adj.assign((size_t)N * out.DEG, ADJ_PAD);
    for (int v = 0; v < N; ++v) {
        int prev = -1;
        for (int k = d.inc_off[v]; k < d.inc_off[v + 1]; ++k) {
            const int e = d.inc_edge[k], o = d.inc_other[k];
            if (e <= prev) return "incident edges must ascend in edge id";
            prev = e;
            const bool first = d.edge_u[e] == v;
            if ((first ? d.edge_v[e] : d.edge_u[e]) != o || (!first && d.edge_v[e] != v))
                return "CSR entry does not match the edge list";
            out.adj[(size_t)v * out.DEG + (k - d.inc_off[v])] = (uint32_t)e | ((uint32_t)o << 16) | (first ? 0x80000000u : 0u);
        }
    }
    out.west_mask.assign(d.west_edge_mask, d.west_edge_mask + words32(E));
    // What the fused kernels' parity-only peel assumes of the graph (every graph of Surface / Repetition has it):
    // no edge joins two boundary nodes, and the logical mask is exactly the edges whose FIRST endpoint is a west
    // boundary node, none of which is ever a second endpoint.  Any other graph takes the full peel.
    out.parity_ok = true;
    for (int e = 0; e < E; ++e) {
        const int u = d.edge_u[e], v = d.edge_v[e];
        const bool west_first = u < B && d.node_long[u] == -1, west_second = v < B && d.node_long[v] == -1;
        const bool masked = (d.west_edge_mask[e >> 5] >> (e & 31)) & 1u;
        if ((u < B && v < B) || west_second || masked != west_first) { out.parity_ok = false; break; }
    }
    // fresh edges, stably sorted by probability class
    out.class_end.assign(d.n_classes, 0);
    out.fresh_perm.clear();
    for (int c = 0; c < d.n_classes; ++c) {
        for (int e = 0; e < E; ++e)
            if (d.edge_class[e] == c) out.fresh_perm.push_back((uint16_t)e);
        out.class_end[c] = (uint32_t)out.fresh_perm.size();
    }
    out.F = (int)out.fresh_perm.size();
    out.perm_identity = out.F == E;
    for (int k = 0; k < out.F && out.perm_identity; ++k) out.perm_identity = out.fresh_perm[k] == k;
    return "";
}

inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// list_cap / list2_cap / claim_slots: the work-list sizes of the tile shape the layout is for
// (defaults: warp tiles; luf::uf_cta uses LUF_CTA_*).
inline Layout make_layout(int N, int B, int E, bool keep_err, bool keep_corr, uint32_t list_cap = LIST_CAP,
                          uint32_t list2_cap = LIST2_CAP, uint32_t claim_slots = CLAIM_SLOTS)
{
    Layout l;
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off = align16(off + bytes); return o; };
    const uint32_t WE = words32(E), WD = words32(N - B) ? words32(N - B) : 1, WN = words32(N);
    l.S = take(16u * ((N + 3) / 4));
    l.zero_off = off;                                 // ---- cleared by ph_reset
    l.growth = take(4u * ((E + 15) / 16));
    l.newfull = take(4u * (WE > WN ? WE : WN));       // tree-root bitmap of the peel (after the last merge round)
    l.defect = take(4u * WD); l.full1 = take(4u * WN); l.full2 = take(4u * WN);
    l.err = keep_err ? take(4u * WE) : 0;
    l.corr = keep_corr ? take(4u * WE) : 0;
    l.zero_bytes = off - l.zero_off;                  // ----
    l.claim = take(4u * (claim_slots > WN ? claim_slots : WN));   // visited bitmap of the peel
    l.list = take(4u * list_cap); l.list2 = take(4u * list2_cap); l.cnt = take(32);
    l.bytes = off;
    return l;
}

// ------------------------------------------------- wide instance (luf_uf_wide.cuh)
struct PackedWide {
    int N = 0, B = 0, E = 0, F = 0, n_classes = 0, DEG = 0;
    bool perm_identity = false, parity_ok = true;
    std::vector<uint64_t> adj, edge_uv;
    std::vector<uint32_t> west_mask, class_end, fresh_perm32;
    std::vector<int16_t> node_long;
};

// pack_graph for graphs of any size the description's 32-bit integers can name.  Returns "" on success.
inline std::string pack_wide(const luf_graph_desc& d, PackedWide& out)
{
    const int N = d.n_nodes, B = d.n_boundary, E = d.n_edges;
    if (N <= 0 || E <= 0 || B < 0 || B > N) return "empty or inconsistent graph";
    if (N > 0x7FFFFFF0 || E > 0x7FFFFFF0) return "graph too large";
    if (d.n_classes < 1 || d.n_classes > MAX_CLASSES) return "n_classes out of range";
    out.N = N; out.B = B; out.E = E; out.n_classes = d.n_classes;
    out.edge_uv.assign(E, 0);
    out.node_long.assign(N, 0);
    for (int e = 0; e < E; ++e) {
        const int u = d.edge_u[e], v = d.edge_v[e];
        if (u < 0 || u >= N || v < 0 || v >= N) return "edge endpoint out of range";
        out.edge_uv[e] = (uint64_t)(uint32_t)u | ((uint64_t)(uint32_t)v << 32);
    }
    for (int v = 0; v < N; ++v) {
        if (((d.node_flags[v] & 1) != 0) != (v < B)) return "boundary nodes must be node indices [0, B)";
        if (d.node_long[v] < -32768 || d.node_long[v] > 32767) return "node coordinate out of range";
        out.node_long[v] = (int16_t)d.node_long[v];
    }
    if (d.inc_off[N] != 2 * E) return "CSR size mismatch";
    int deg = 1;
    for (int v = 0; v < N; ++v) deg = std::max(deg, d.inc_off[v + 1] - d.inc_off[v]);
    out.DEG = deg;
    out.adj.assign((size_t)N * out.DEG, uf_wide::ADJ_PAD64);
    for (int v = 0; v < N; ++v) {
        int prev = -1;
        for (int k = d.inc_off[v]; k < d.inc_off[v + 1]; ++k) {
            const int e = d.inc_edge[k], o = d.inc_other[k];
            if (e <= prev || e >= E) return "incident edges must ascend in edge id";
            prev = e;
            const bool first = d.edge_u[e] == v;
            if ((first ? d.edge_v[e] : d.edge_u[e]) != o || (!first && d.edge_v[e] != v))
                return "CSR entry does not match the edge list";
            out.adj[(size_t)v * out.DEG + (k - d.inc_off[v])] =
                (uint64_t)(uint32_t)e | ((uint64_t)(uint32_t)o << 32) | (first ? 0x8000000000000000ull : 0ull);
        }
    }
    out.west_mask.assign(d.west_edge_mask, d.west_edge_mask + words32(E));
    out.parity_ok = true;               // as pack_graph
    for (int e = 0; e < E; ++e) {
        const int u = d.edge_u[e], v = d.edge_v[e];
        const bool west_first = u < B && d.node_long[u] == -1, west_second = v < B && d.node_long[v] == -1;
        const bool masked = (d.west_edge_mask[e >> 5] >> (e & 31)) & 1u;
        if ((u < B && v < B) || west_second || masked != west_first) { out.parity_ok = false; break; }
    }
    out.class_end.assign(d.n_classes, 0);
    out.fresh_perm32.clear();
    for (int c = 0; c < d.n_classes; ++c) {
        for (int e = 0; e < E; ++e)
            if (d.edge_class[e] == c) out.fresh_perm32.push_back((uint32_t)e);
        out.class_end[c] = (uint32_t)out.fresh_perm32.size();
    }
    out.F = (int)out.fresh_perm32.size();
    out.perm_identity = out.F == E;
    for (int k = 0; k < out.F && out.perm_identity; ++k) out.perm_identity = out.fresh_perm32[k] == (uint32_t)k;
    return "";
}

inline uf_wide::Layout make_wide_layout(int N, int B, int E, bool keep_err, bool keep_corr, uint32_t list_cap = LUF_WIDE_LIST_CAP,
                                        uint32_t list2_cap = LUF_WIDE_LIST2_CAP, uint32_t claim_slots = LUF_WIDE_CLAIM_SLOTS)
{
    uf_wide::Layout l;
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off = (off + bytes + 15u) & ~15u; return o; };
    const uint32_t WE = words32(E), WD = words32(N - B) ? words32(N - B) : 1, WN = words32(N);
    l.S = take(8u * (uint32_t)N);
    l.zero_off = off;                                 // ---- cleared by ph_reset
    l.growth = take(4u * ((E + 15) / 16));
    l.newfull = take(4u * (WE > WN ? WE : WN));
    l.defect = take(4u * WD); l.full1 = take(4u * WN); l.full2 = take(4u * WN);
    l.err = keep_err ? take(4u * WE) : 0;
    l.corr = keep_corr ? take(4u * WE) : 0;
    l.zero_bytes = off - l.zero_off;                  // ----
    l.claim = take(std::max(8u * claim_slots, 4u * WN));   // visited bitmap of the peel
    l.list = take(8u * list_cap); l.list2 = take(8u * list2_cap); l.cnt = take(32);
    l.bytes = (off + 127u) & ~127u;                   // slabs of consecutive blocks start on their own cache lines
    return l;
}

// ------------------------------------------------- compact fast path (luf_uf_fast.cuh)
struct PackedFast {
    bool ok = false;              // the graph allows the compact fast path
    int D = 0, DEG = 0, E = 0, F = 0, n_classes = 0;
    uint32_t ebits = 16, emask = 0xFFFFu, omask = 0x7FFFu;
    std::vector<uint32_t> fadj, einfo;
    // sampler tables of graphs too large for the 16-bit layout (pack_graph refuses them): the compact path is
    // then the only one, and brings its own
    std::vector<uint32_t> fresh_perm32, class_end, west_mask;
    bool perm_identity = false;
};

// The compact tables index detectors only.  Needs what the parity-only peel needs (pack_graph's parity_ok: no
// edge joins two boundary nodes, the logical mask is exactly the edges whose FIRST endpoint is a west boundary
// node, none of which is ever a second endpoint) and, on top, that every boundary node has exactly one edge and
// lies on the west (-1) side or on one other side.  Every graph of codes.Surface / codes.Repetition has all of
// it.  Works from the description alone, so that it also serves graphs beyond the 16-bit layout of the
// canonical kernels (up to 32766 detectors; the edge / other-end split of an adjacency entry moves with the
// edge count: Surface(25, 'circuit-level') has 85297 edges and 15000 detectors -> 17 + 14 bits).
inline void pack_fast(const luf_graph_desc& d, PackedFast& out)
{
    const int N = d.n_nodes, B = d.n_boundary, E = d.n_edges, D = N - B;
    out.ok = false;
    if (N <= 0 || E <= 0 || B < 0 || B > N || D < 1 || D > fast::MAX_DETECTORS) return;
    if (d.n_classes < 1 || d.n_classes > MAX_CLASSES || d.inc_off[N] != 2 * E) return;
    uint32_t ebits = 16;
    while (ebits < 31 && ((uint64_t)1 << ebits) < (uint64_t)E + 1) ++ebits;
    const uint32_t omask = (1u << (31 - ebits)) - 1u;
    if ((uint64_t)D + 2 > (uint64_t)omask + 1) return;               // detector indices and the two boundary codes
    for (int v = 0; v < N; ++v)
        if (((d.node_flags[v] & 1) != 0) != (v < B)) return;
    int east_long = 0;
    bool have_east = false;
    for (int v = 0; v < B; ++v) {
        if (d.inc_off[v + 1] - d.inc_off[v] != 1) return;
        if (d.node_long[v] == -1) continue;
        if (have_east && d.node_long[v] != east_long) return;
        have_east = true; east_long = d.node_long[v];
    }
    auto end_code = [&](int v) -> uint32_t { return v >= B ? (uint32_t)(v - B) : (d.node_long[v] == -1 ? fast::END_WEST : fast::END_EAST); };
    out.einfo.assign(E, 0);
    for (int e = 0; e < E; ++e) {
        const int u = d.edge_u[e], v = d.edge_v[e];
        if (u < 0 || u >= N || v < 0 || v >= N) return;
        const bool west_first = u < B && d.node_long[u] == -1, west_second = v < B && d.node_long[v] == -1;
        const bool masked = (d.west_edge_mask[e >> 5] >> (e & 31)) & 1u;
        if ((u < B && v < B) || west_second || masked != west_first) return;
        out.einfo[e] = end_code(u) | (end_code(v) << 16);
    }
    int deg = 2;
    for (int v = B; v < N; ++v) deg = std::max(deg, d.inc_off[v + 1] - d.inc_off[v]);
    out.D = D; out.E = E; out.DEG = (deg + 1) & ~1;
    out.ebits = ebits; out.emask = (1u << ebits) - 1u; out.omask = omask;
    out.fadj.assign((size_t)D * out.DEG, ADJ_PAD);
    for (int v = B; v < N; ++v) {
        int prev = -1;
        for (int k = d.inc_off[v]; k < d.inc_off[v + 1]; ++k) {
            const int e = d.inc_edge[k], o = d.inc_other[k];
            if (e <= prev || e >= E) return;
            prev = e;
            const bool first = d.edge_u[e] == v;
            if ((first ? d.edge_v[e] : d.edge_u[e]) != o || (!first && d.edge_v[e] != v)) return;
            const uint32_t oc = o >= B ? (uint32_t)(o - B) : (d.node_long[o] == -1 ? omask : omask - 1u);
            out.fadj[(size_t)(v - B) * out.DEG + (k - d.inc_off[v])] = (uint32_t)e | (oc << ebits) | (first ? 0x80000000u : 0u);
        }
    }
    out.class_end.assign(d.n_classes, 0);
    out.fresh_perm32.clear();
    for (int c = 0; c < d.n_classes; ++c) {
        for (int e = 0; e < E; ++e)
            if (d.edge_class[e] == c) out.fresh_perm32.push_back((uint32_t)e);
        out.class_end[c] = (uint32_t)out.fresh_perm32.size();
    }
    out.F = (int)out.fresh_perm32.size(); out.n_classes = d.n_classes;
    out.perm_identity = out.F == E;
    for (int k = 0; k < out.F && out.perm_identity; ++k) out.perm_identity = out.fresh_perm32[k] == (uint32_t)k;
    out.west_mask.assign(d.west_edge_mask, d.west_edge_mask + words32(E));
    out.ok = true;
}

// List capacities follow the expected number of faults per shot `mu` (the launch knows the noise level): the
// touched nodes (two defects per fault and their neighbours in later rounds), the active nodes of a round and
// the newly FULL edges of a round.  A shot that overflows one of them is HARD (the canonical kernel decodes
// it), so the capacities trade shared memory for the share of such shots.
// `log`: room for the merge log (every newly FULL edge of a shot: about two per fault near threshold) -- what lets
// shots with odd clusters on both sides skip the canonical kernel (hard_resolve, luf_uf_fast.cuh).
inline fast::FLayout make_fast_layout(int D, int E, int DEG, double mu, bool log = false)
{
    fast::FLayout l;
    auto cap = [&](double want, int most) { uint32_t c = (uint32_t)std::min<double>(want, most); return (c + 7u) & ~7u; };
    // fitted to the 99.9th percentiles of the host emulation (Surface d = 9 .. 17, p from 0.005 to threshold):
    // touched nodes <= 5.6 mu, newly FULL edges of a round <= 4.8 mu, active nodes of a round <= 1.9 mu there
    const double sq = std::sqrt(mu > 0.0 ? mu : 0.0);
    // (degree 6; the diagonal edges of circuit-level graphs, degree 12, fill twice as many edges per round)
    // LUF_FAST_CAP_SCALE (measurements): an overflowing list is not fatal (its consumers read the bitmaps), so the
    // capacities trade shared memory per shot against how often that slower path runs.  cfg3 on a B200, M shots/s
    // at scale 1.0 / 0.6 / 0.4 (99.9th / ~99th / ~95th percentile of the list sizes): 175.8 / 181.2 / 176.3
    static const double scale = getenv("LUF_FAST_CAP_SCALE") ? atof(getenv("LUF_FAST_CAP_SCALE")) : 0.6;
    const double f = (DEG > 6 ? DEG / 6.0 : 1.0) * scale;
    l.tcap = cap((5.2 * mu + 10.0 * sq + 32.0) * (scale + f) / 2.0, D); l.acap = cap((mu + 3.0 * sq + 16.0) * scale, D);
    l.mcap = cap((4.0 * mu + 8.0 * sq + 24.0) * f, E);
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off = align16(off + bytes); return o; };
    const uint32_t WD = (uint32_t)words32(D);
    l.P = take(2u * (uint32_t)D);
    l.fill_words = (off - l.P) / 4u;
    l.zero_off = off;
    l.growth = take(4u * (uint32_t)((E + 15) / 16)); l.defect = take(4u * WD); l.touched = take(4u * WD);
    l.zero_bytes = off - l.zero_off;
    l.tlist = take(2u * l.tcap); l.alist = take(4u * l.acap); l.mlist = take(4u * l.mcap); l.cnt = take(32);
    // (host emulation, Surface d = 5 .. 17 near threshold: the shots that need the log make 6 mu newly FULL edges on
    // average and 10 mu at their 99th percentile -- every edge inside a cluster becomes FULL at some point; the
    // diagonal edges of circuit-level graphs make that 37 mu.  Only the second pass over the HARD shots carries a log.)
    // An edge becomes FULL once, so E entries always suffice; a shot beyond the log goes to the canonical launch, whose
    // latency for ONE such shot is milliseconds (5 ms for 11 shots of d = 17 at p = 0.026): be generous.
    l.lcap = log ? cap((16.0 * mu + 40.0 * sq + 128.0) * (DEG > 6 ? 3.0 : 1.0), E) : 0u;
    if (l.lcap) { uint32_t p2 = 64; while (p2 < l.lcap) p2 <<= 1; l.lcap = p2; }      // sorted in place by a bitonic network
    l.elog = l.lcap ? take(4u * l.lcap) : 0u;
    l.bytes = off;
    return l;
}

inline SampleParams make_sample_params(const double* class_p, int n_classes, uint64_t seed)
{
    SampleParams sp;
    for (int c = 0; c < MAX_CLASSES; ++c) { sp.p[c] = 0.f; sp.inv_log2_q[c] = 0.f; }
    for (int c = 0; c < n_classes; ++c) {
        const double p = class_p[c];
        sp.p[c] = p >= 1.0 ? 1.0f : (p <= 0.0 ? 0.0f : (float)p);
        if (p > 0.0 && p < 1.0) {
            if (sp.p[c] < 1.17549435e-38f) sp.p[c] = 1.17549435e-38f;   // keep tiny p "on": FLT_MIN, not a denormal (--use_fast_math flushes those to 0)
            if (sp.p[c] >= 1.0f) sp.p[c] = 0.99999994f;
            sp.inv_log2_q[c] = (float)(1.0 / (std::log1p(-p) / std::log(2.0)));
        }
    }
    sp.key0 = (uint32_t)seed; sp.key1 = (uint32_t)(seed >> 32);
    return sp;
}

// ------------------------------------------------------------- Macar / Actis
struct PackedCa {
    int N = 0, B = 0, E = 0, K = 0, global_span = 0;
    bool actis_ok = false;
    std::vector<uint32_t> nbr;
    std::vector<uint8_t> span;
    std::vector<uint16_t> relay;
};

// Returns "" on success, else why the cellular-automaton decoders cannot run on this graph.
inline std::string pack_ca(const luf_graph_desc& d, PackedCa& out)
{
    const int N = d.n_nodes, K = d.n_dirs;
    if (K < 2 || K > 12 || (K & 1)) return "neighbour table width must be even and at most 12";
    if (N > 0xFFFE || d.n_edges > 0xFFFE) return "graph too large for the 16-bit cellular-automaton layout";
    out.N = N; out.B = d.n_boundary; out.E = d.n_edges; out.K = K; out.global_span = d.actis_global_span;
    out.nbr.assign((size_t)N * K, NBR_NONE);
    for (int v = 0; v < N; ++v) {
        if (d.node_id[v] != v) return "Macar/Actis need node index == index_to_id (surface-code graphs)";
        for (int k = 0; k < K; ++k) {
            const int e = d.nbr_edge[v * K + k], w = d.nbr_node[v * K + k];
            if (e < 0) continue;
            if ((d.nbr_owned[v * K + k] != 0) != (k >= K / 2)) return "unexpected edge ownership in the neighbour table";
            out.nbr[(size_t)v * K + k] = (uint32_t)e | ((uint32_t)w << 16);
        }
    }
    out.span.assign(N, 0); out.relay.assign(N, 0xFFFF);
    out.actis_ok = d.actis_global_span > 0 && d.actis_global_span < 255;
    for (int v = 0; v < N && out.actis_ok; ++v) {
        const int sp = d.actis_span[v], r = d.actis_relay[v];
        if (sp < 0 || sp > 255 || r < -1 || r >= N) { out.actis_ok = false; break; }
        out.span[v] = (uint8_t)sp;
        out.relay[v] = r < 0 ? (uint16_t)0xFFFF : (uint16_t)r;
    }
    return "";
}

inline CaLayout make_ca_layout(int N, int B, int E, bool actis, bool keep_err, uint32_t tlist_cap = CA_TLIST_CAP)
{
    CaLayout l;
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off = align16(off + bytes); return o; };
    const uint32_t WE = words32(E), WD = words32(N - B) ? words32(N - B) : 1, WN = words32(N), GW = (E + 15) / 16;
    l.cid = take(2u * N); l.ncid = take(2u * N);
    l.acp = take(2u * N); l.fl = take(N);
    l.ast = actis ? take(N) : 0; l.cnt = actis ? take(N) : 0;
    l.growth = take(4u * GW);
    l.rx_anyon = take(4u * WN); l.rx_defect = take(4u * WN);
    l.rx_busy = actis ? take(8u * WN) : 0; l.rx_active = actis ? take(8u * WN) : 0;
    l.defect = take(4u * WD);
    l.corr = take(4u * WE);
    l.glob = take(16);
    l.touched = take(4u * WN);
    l.tlist_cap = tlist_cap;
    l.tlist = take(2u * tlist_cap);
    l.err = keep_err ? take(4u * WE) : 0;
    l.bytes = off;
    return l;
}

// ------------------------------------------------------------ forward scheme
struct PackedForward {
    int c = 0, n_cleanse = 0, n_buffer = 0;
    std::vector<uint16_t> lower_edge, art_node, node_lower, buffer_perm;
    std::vector<uint32_t> commit_mask, fb_mask, buffer_class_end;
    std::vector<uint8_t> node_side;
};

inline std::string pack_forward(const luf_graph_desc_dims& dims, const luf_forward_desc& d, PackedForward& out)
{
    const int N = dims.N, E = dims.E, B = dims.B, C = dims.n_classes;
    if (d.commit_height < 1 || d.n_cleanse < 1) return "bad commit height";
    out.c = d.commit_height; out.n_cleanse = d.n_cleanse;
    out.lower_edge.assign(E, 0xFFFF); out.art_node.assign(2 * (size_t)E, 0xFFFF);
    out.commit_mask.assign(d.commit_mask, d.commit_mask + words32(E));
    out.fb_mask.assign(words32(E), 0);
    for (int e = 0; e < E; ++e) {
        if (d.lower_edge[e] >= E) return "lower_edge out of range";
        if (d.lower_edge[e] >= 0) out.lower_edge[e] = (uint16_t)d.lower_edge[e];
        for (int x = 0; x < 2; ++x) {
            const int a = d.art_node[2 * e + x];
            if (a >= N || (a >= 0 && a < B)) return "artificial defect must be a detector";
            if (a >= 0) out.art_node[2 * e + x] = (uint16_t)a;
        }
        if (d.edge_fb[e]) out.fb_mask[e >> 5] |= 1u << (e & 31);
    }
    out.node_lower.assign(N, 0xFFFF); out.node_side.assign(d.node_side, d.node_side + N);
    for (int v = 0; v < N; ++v) {
        if (d.node_lower[v] >= N) return "node_lower out of range";
        if (d.node_lower[v] >= 0) out.node_lower[v] = (uint16_t)d.node_lower[v];
    }
    out.buffer_class_end.assign(C, 0); out.buffer_perm.clear();
    for (int c = 0; c < C; ++c) {
        for (int e = 0; e < E; ++e)
            if (d.buffer_class[e] == c) out.buffer_perm.push_back((uint16_t)e);
        out.buffer_class_end[c] = (uint32_t)out.buffer_perm.size();
    }
    out.n_buffer = (int)out.buffer_perm.size();
    return "";
}

inline FwLayout make_fw_layout(int N, int E, uint32_t base)
{
    FwLayout l;
    uint32_t off = base;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off = align16(off + bytes); return o; };
    l.buf = take(4u * words32(E)); l.commit = take(4u * words32(E));
    l.partner = take(2u * N); l.partner2 = take(2u * N);
    l.bytes = off;
    return l;
}

// forward scheme on the wide instance (luf_fw_wide.cuh): the same tables with 32-bit entries
struct PackedForwardWide {
    int c = 0, n_cleanse = 0, n_buffer = 0;
    std::vector<uint32_t> lower_edge, art_node, node_lower, buffer_perm, commit_mask, fb_mask, buffer_class_end;
    std::vector<uint8_t> node_side;
};

inline std::string pack_forward_wide(const luf_graph_desc_dims& dims, const luf_forward_desc& d, PackedForwardWide& out)
{
    const int N = dims.N, E = dims.E, B = dims.B, C = dims.n_classes;
    if (d.commit_height < 1 || d.n_cleanse < 1) return "bad commit height";
    out.c = d.commit_height; out.n_cleanse = d.n_cleanse;
    out.lower_edge.assign(E, fw_wide::W_NONE); out.art_node.assign(2 * (size_t)E, fw_wide::W_NONE);
    out.commit_mask.assign(d.commit_mask, d.commit_mask + words32(E));
    out.fb_mask.assign(words32(E), 0);
    for (int e = 0; e < E; ++e) {
        if (d.lower_edge[e] >= E) return "lower_edge out of range";
        if (d.lower_edge[e] >= 0) out.lower_edge[e] = (uint32_t)d.lower_edge[e];
        for (int x = 0; x < 2; ++x) {
            const int a = d.art_node[2 * e + x];
            if (a >= N || (a >= 0 && a < B)) return "artificial defect must be a detector";
            if (a >= 0) out.art_node[2 * (size_t)e + x] = (uint32_t)a;
        }
        if (d.edge_fb[e]) out.fb_mask[e >> 5] |= 1u << (e & 31);
    }
    out.node_lower.assign(N, fw_wide::W_NONE); out.node_side.assign(d.node_side, d.node_side + N);
    for (int v = 0; v < N; ++v) {
        if (d.node_lower[v] >= N) return "node_lower out of range";
        if (d.node_lower[v] >= 0) out.node_lower[v] = (uint32_t)d.node_lower[v];
    }
    out.buffer_class_end.assign(C, 0); out.buffer_perm.clear();
    for (int c = 0; c < C; ++c) {
        for (int e = 0; e < E; ++e)
            if (d.buffer_class[e] == c) out.buffer_perm.push_back((uint32_t)e);
        out.buffer_class_end[c] = (uint32_t)out.buffer_perm.size();
    }
    out.n_buffer = (int)out.buffer_perm.size();
    return "";
}

inline fw_wide::FwLayout make_fw_wide_layout(int N, int E, uint32_t base)
{
    fw_wide::FwLayout l;
    uint32_t off = base;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off = align16(off + bytes); return o; };
    l.buf = take(4u * words32(E)); l.commit = take(4u * words32(E)); l.nz = take(4u * ((words32(E) + 31) / 32));
    l.partner = take(4u * N); l.partner2 = take(4u * N);
    l.bytes = (off + 127u) & ~127u;
    return l;
}

// ----------------------------------------------------- Snowflake / frugal
struct PackedWindow {
    int d = 0, h = 0, NLb = 0, NLd = 0, NL = 0, EL = 0, K = 0, n_classes = 0;
    std::vector<uint32_t> nbr, edge_ends, class_end;
    std::vector<int8_t> pos_long;
    std::vector<uint16_t> fresh_perm;
};

inline std::string pack_window(const luf_window_desc& d, PackedWindow& out)
{
    const int NL = d.n_boundary_per_layer + d.n_detectors_per_layer, K = d.n_dirs, EL = d.edges_per_layer;
    if (d.h < 3 || NL <= 0 || EL <= 0 || K <= 0 || K > 12) return "inconsistent window (at most 12 neighbour slots)";
    if (NL > 1023 || EL > 4095) return "window layer too large for the packed tables (NL <= 1023, EL <= 4095)";
    if ((long long)NL * d.h > 0xFFFE) return "window too large for 16-bit cluster ids";
    if (d.n_classes < 1 || d.n_classes > MAX_CLASSES) return "n_classes out of range";
    out.d = d.d; out.h = d.h; out.NLb = d.n_boundary_per_layer; out.NLd = d.n_detectors_per_layer;
    out.NL = NL; out.EL = EL; out.K = K; out.n_classes = d.n_classes;
    out.nbr.assign((size_t)NL * K, 0);
    for (int x = 0; x < NL * K; ++x) {
        if (!d.nb_valid[x]) continue;
        const int dt = d.nb_dt[x], et = d.nb_et[x];
        if (dt < -1 || dt > 1 || et < -1 || et > 0 || d.nb_q[x] < 0 || d.nb_q[x] >= NL || d.nb_el[x] < 0 || d.nb_el[x] >= EL)
            return "neighbour table entry out of range";
        out.nbr[x] = 1u | ((uint32_t)d.nb_q[x] << 1) | ((uint32_t)d.nb_el[x] << 11) |
                     ((uint32_t)(dt + 1) << 23) | ((uint32_t)(et + 1) << 25);
    }
    // the slot under which the neighbour sees the same edge (same local edge; its block seen from
    // the other end: et' = et - dt): needed to mark access at both ends when an edge becomes fully grown
    for (int x = 0; x < NL * K; ++x) {
        if (!d.nb_valid[x]) continue;
        const int q = x / K, q2 = d.nb_q[x];
        int back = -1;
        for (int k2 = 0; k2 < K; ++k2) {
            const int y = q2 * K + k2;
            if (d.nb_valid[y] && d.nb_q[y] == q && d.nb_dt[y] == -d.nb_dt[x] && d.nb_el[y] == d.nb_el[x] &&
                d.nb_et[y] == d.nb_et[x] - d.nb_dt[x]) { back = k2; break; }
        }
        out.nbr[x] |= (uint32_t)(back < 0 ? 15 : back) << 26;    // 15: the far end does not scan this direction
    }
    out.edge_ends.assign(EL, 0);
    for (int l = 0; l < EL; ++l) {
        for (int x = 0; x < 2; ++x) {
            const int q = d.edge_q[2 * l + x], dt = d.edge_dt[2 * l + x];
            if (q < 0 || q >= NL || dt < 0 || dt > 1) return "edge endpoint out of range";
            out.edge_ends[l] |= ((uint32_t)q | ((uint32_t)dt << 10)) << (11 * x);
        }
        if (d.edge_fb[l]) out.edge_ends[l] |= 1u << 22;
    }
    out.pos_long.assign(NL, 0);
    for (int q = 0; q < NL; ++q) out.pos_long[q] = (int8_t)d.pos_long[q];
    out.class_end.assign(d.n_classes, 0);
    out.fresh_perm.clear();
    for (int c = 0; c < d.n_classes; ++c) {
        for (int l = 0; l < EL; ++l)
            if (d.edge_class[l] == c) out.fresh_perm.push_back((uint16_t)l);
        out.class_end[c] = (uint32_t)out.fresh_perm.size();
    }
    if ((int)out.fresh_perm.size() != EL) return "every edge of the fresh sheet needs a probability class";
    return "";
}

inline SfLayout make_sf_layout(const PackedWindow& p, uint32_t alist_cap = SF_ALIST_CAP)
{
    SfLayout l;
    uint32_t off = 0;
    auto take = [&](uint32_t bytes) { uint32_t o = off; off = align16(off + bytes); return o; };
    const uint32_t N = (uint32_t)p.NL * p.h, ELw = words32(p.EL), WLn = words32(p.NL);
    l.cid = take(2u * N); l.ncid = take(2u * N); l.fl = take(2u * N); l.acp = take(2u * (N + 1u));
    l.ndef = take(4u * p.h * WLn); l.awake = take(4u * p.h * WLn);
    l.growth = take(4u * p.h * 2u * ELw);
    l.corr = take(4u * p.h * ELw); l.err = take(4u * p.h * ELw);
    l.top = take(4u * WLn); l.fb = take(4u * WLn);
    l.partner = take(2u * 2u * p.NL);
    l.alist_cap = alist_cap;
    l.alist = take(2u * alist_cap); l.acnt = take(16);
    l.bytes = off;
    return l;
}

}  // namespace luf
