// luf_uf_wide.cuh -- the wide instance of the Union-Find phases: graphs beyond the 16-bit state layout of
// luf_core.cuh (more than 32767 nodes or 65533 edges: Surface(23 | 25, 'circuit-level'), codes.py:155-251, the
// windows of the global batch scheme with h = d * n layers, _schemes.py:166-191, forward windows of large codes).
//
// The same phase source (luf_uf_phases.inl) compiled a third time, as namespace luf::uf_wide, with 64-bit node
// words and 32-bit indices:
//   root      bit 31 set, bits 0..30 = cluster size, bit 63 = odd defect parity,
//             bits 32..62 = boundary node of the cluster + 1 (0 = none)
//   non-root  bit 31 clear, bits 0..30 = parent, bits 32..63 = the edge whose union made it a non-root
//   list / reservation entries: lower half | upper half << 32 likewise
//   adjacency entry (64 bits): edge | other end << 32 | this is the edge's first endpoint << 63
// One thread block owns one shot at a time.  A shot's state is a slab of `Layout::bytes` bytes that the kernel
// binds wherever it fits: the block's shared memory, or a per-block slab in global memory that stays in L2 (the
// state-spill tile: every access of the phases is through generic pointers, and a block's barriers order its
// global accesses like its shared ones).  Same algorithm, same canonical order, same random stream as the
// 16-bit instances; dual compilation (-DLUF_HOST_EMU) as in luf_core.cuh.
#pragma once
#include "luf_core.cuh"

namespace luf {
namespace uf_wide {

typedef uint64_t word_t;
typedef uint64_t ent_t;
static const int HSH = 32;
static const word_t LO_MASK = 0xFFFFFFFFull, IDX_NONE = 0xFFFFFFFFull, TREE_ROOT_W = 0xFFFFFFFEull;
static const word_t ROOT_BIT = 0x80000000ull, ODD_BIT = 0x8000000000000000ull;
static const word_t BND_MASK = 0x7FFFFFFF00000000ull, SIZE_MASK = 0x7FFFFFFFull;
static const uint64_t ADJ_PAD64 = ~(uint64_t)0;

struct GraphView {
    int N, B, E;
    int WE, WD, WN;
    int DEG;                    // row length of adj (maximum degree)
    const uint64_t* adj;        // [N][DEG] ascending edge id; ADJ_PAD64 fills a row
    const uint64_t* edge_uv;    // [E]   u | v << 32
    const uint32_t* west_mask;  // [WE]
    const int16_t* node_long;   // [N]
    int parity_ok, leaf_bnd;    // as luf::GraphView
    int F, n_classes;
    const uint32_t* fresh_perm32;   // [F]; null when it is the identity
    const uint32_t* class_end;      // [n_classes]
};

// as luf::Layout; S, claim, list, list2 hold 64-bit words
struct Layout {
    uint32_t S, growth, newfull, defect, full1, full2, err, corr, zero_off, zero_bytes, claim, list, list2, cnt, bytes;
};

struct Shot {
    word_t *S, *claim, *list, *list2;
    uint32_t *growth, *newfull, *defect, *full1, *full2, *err, *corr, *cnt;
};

LUF_DEV Shot bind(const Layout& l, unsigned char* base)
{
    Shot s;
    s.S = (word_t*)(base + l.S);
    s.growth = (uint32_t*)(base + l.growth); s.newfull = (uint32_t*)(base + l.newfull);
    s.defect = (uint32_t*)(base + l.defect); s.full1 = (uint32_t*)(base + l.full1);
    s.full2 = (uint32_t*)(base + l.full2);
    s.err = l.err ? (uint32_t*)(base + l.err) : 0;
    s.corr = l.corr ? (uint32_t*)(base + l.corr) : 0;
    s.claim = (word_t*)(base + l.claim);
    s.list = (word_t*)(base + l.list); s.list2 = (word_t*)(base + l.list2); s.cnt = (uint32_t*)(base + l.cnt);
    return s;
}

LUF_DEV uint32_t growth_of(const Shot& s, uint32_t e) { return (s.growth[e >> 4] >> ((e & 15u) * 2u)) & 3u; }

LUF_DEV uint32_t ent_edge(ent_t ent) { return (uint32_t)ent; }
LUF_DEV uint32_t other_end(ent_t ent, bool& first)
{
    first = (ent >> 63) != 0;
    return (uint32_t)(ent >> 32) & 0x7FFFFFFFu;
}
LUF_DEV void edge_ends(const GraphView& g, uint32_t e, uint32_t& u, uint32_t& v)
{
    const uint64_t uv = LUF_LDG64(&g.edge_uv[e]);
    u = (uint32_t)uv; v = (uint32_t)(uv >> 32);
}

template <class F>
LUF_DEV void for_row(const GraphView& g, uint32_t v, F f)
{
    const uint64_t* row = g.adj + (size_t)v * (uint32_t)g.DEG;
    LUF_NOUNROLL
    for (int k = 0; k < g.DEG; ++k) {
        const uint64_t ent = LUF_LDG64(row + k);
        if (ent == ADJ_PAD64) break;
        f(ent);
    }
}

// _base_classes.py:427-450 (luf::flip_edge on the wide tables)
LUF_DEV uint32_t flip_edge(const GraphView& g, const Shot& s, uint32_t e)
{
    if (s.err) luf_atomic_xor(&s.err[e >> 5], 1u << (e & 31u));
    uint32_t u, v;
    edge_ends(g, e, u, v);
    if (u >= (uint32_t)g.B) luf_atomic_xor(&s.defect[(u - g.B) >> 5], 1u << ((u - g.B) & 31u));
    if (v >= (uint32_t)g.B) luf_atomic_xor(&s.defect[(v - g.B) >> 5], 1u << ((v - g.B) & 31u));
    return (LUF_LDG(&g.west_mask[e >> 5]) >> (e & 31u)) & 1u;
}

// One sampler, one random stream (luf::ph_sample_core): `lane` of the 32 sampler lanes of a shot.
LUF_DEV uint32_t ph_sample_tab(const GraphView& g, const SampleTab& t, const Shot& s, const SampleParams& sp,
                               uint64_t shot, uint32_t cycle, const uint32_t* skip_mask, int lane, int nl)
{
    return ph_sample_core(t, sp, shot, cycle, skip_mask, lane, nl, [&](uint32_t e) { return flip_edge(g, s, e); });
}

LUF_DEV uint32_t ph_sample(const GraphView& g, const Shot& s, const SampleParams& sp, uint64_t shot, int lane, int nl)
{
    SampleTab t;
    t.F = g.F; t.n_classes = g.n_classes; t.perm = 0; t.class_end = g.class_end; t.perm32 = g.fresh_perm32;
    return ph_sample_tab(g, t, s, sp, shot, 0u, (const uint32_t*)0, lane, nl);
}

// luf::ph_force on the wide tables (noise/main.py:116-122)
LUF_DEV void ph_force(const GraphView& g, const Shot& s, const SampleParams& sp, uint32_t weight, uint64_t shot, int lane)
{
    if (lane != 0) return;
    Philox rng;
    rng.c0 = (uint32_t)shot; rng.c1 = (uint32_t)(shot >> 32); rng.c2 = 0xF0CEu; rng.c3 = 0u;
    rng.k0 = sp.key0; rng.k1 = sp.key1; rng.have = 0;
    rng.o0 = rng.o1 = rng.o2 = rng.o3 = 0;
    const uint32_t F = (uint32_t)g.F;
    if (weight > F) weight = F;
    LUF_NOUNROLL
    for (uint32_t j = F - weight; j < F; ++j) {
        const uint32_t t = luf_mulhi(philox_next(rng), j + 1u);
        const uint32_t et = g.fresh_perm32 ? LUF_LDG(&g.fresh_perm32[t]) : t;
        const uint32_t e = !((s.err[et >> 5] >> (et & 31u)) & 1u) ? et : g.fresh_perm32 ? LUF_LDG(&g.fresh_perm32[j]) : j;
        flip_edge(g, s, e);
    }
}

}  // namespace uf_wide
}  // namespace luf

// Tile = one thread block (as luf::uf_cta); the work lists live in the slab, so they can be long.
#ifndef LUF_WIDE_LIST_CAP
  #define LUF_WIDE_LIST_CAP 4096
#endif
#ifndef LUF_WIDE_LIST2_CAP
  #define LUF_WIDE_LIST2_CAP 2048
#endif
#define LUF_WIDE_CLAIM_SLOTS 2048
#if defined(LUF_HOST_EMU)
  #define LUF_T_PHASE(stmt) LUF_PHASE(stmt)
  #define LUF_T_ANY(x) LUF_ANY(x)
  #define LUF_T_SUM(x) LUF_SUM(x)
  #define LUF_T_EXSCAN(c, lane, nl, pos, total) luf_exscan(c, lane, nl, pos, total)
  #define LUF_T_SCAN_PREP(expr) LUF_SCAN_PREP(expr)
#else
  #define LUF_T_PHASE(stmt) { stmt; } __syncthreads();
  #define LUF_T_ANY(x) __syncthreads_or((x) != 0)
  #define LUF_T_SUM(x) luf_block_sum(x)
  #define LUF_T_EXSCAN(c, lane, nl, pos, total) luf_block_exscan(c, lane, nl, pos, total)
  #define LUF_T_SCAN_PREP(expr)
#endif
#define LUF_T_WIDE 1
#define LUF_UF_NS uf_wide
#define LUF_T_LIST_CAP ((uint32_t)LUF_WIDE_LIST_CAP)
#define LUF_T_LIST2_CAP ((uint32_t)LUF_WIDE_LIST2_CAP)
#define LUF_T_CLAIM_SLOTS LUF_WIDE_CLAIM_SLOTS
#include "luf_uf_phases.inl"
#undef LUF_T_WIDE
#undef LUF_UF_NS
#undef LUF_T_LIST_CAP
#undef LUF_T_LIST2_CAP
#undef LUF_T_CLAIM_SLOTS
#undef LUF_T_PHASE
#undef LUF_T_ANY
#undef LUF_T_SUM
#undef LUF_T_EXSCAN
#undef LUF_T_SCAN_PREP
