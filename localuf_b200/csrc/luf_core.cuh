// luf_core.cuh -- per-shot phases of the B200 Union-Find decoding engine.
//
// One *tile* of NL cooperating lanes (a warp in the shipped kernels) owns one
// shot at a time; the shot's whole decoder state lives in that tile's slice of
// shared memory.  A shot is a sequence of phases separated by tile barriers;
// inside a phase every lane runs the same function with its own lane index and
// only touches shared state through atomics or through data it exclusively
// owns.  The phases are written against a tiny portability layer (LUF_* macros
// below) so that the very same source also compiles as plain C++
// (-DLUF_HOST_EMU, tests/emu/): there the lanes of a phase simply run one after
// the other.  That build is a test harness for the kernel logic, never a
// product path -- the product has no CPU fallback.
//
// Algorithm (reference: localuf/decoders/uf.py, canonical order of
// oracle/ref_canonical.py):
//   K1  sample + syndrome   noise/main.py:114-115,309-314; _base_classes.py:427-450
//   K2a validate            uf.py:200-301  (grow a half-edge per round, merge in ascending edge id)
//   K2b peel                uf.py:305-344  (LIFO spanning tree in ascending edge id, leaf-up peeling)
//   K4  logical parity      _schemes.py:123-129
#pragma once
#include <stdint.h>

// LUF_OPT_NOUNROLL (default 1): the data-dependent loops of the UF phases stay rolled.  nvcc's default unrolling
// made k_mc_uf 3152 instructions / 71 registers and k_decode_uf 96 registers; rolled: 2432 instructions / 63
// registers, cfg3 fused 58.4 -> 59.3 M shots/s, staged pipeline 48.0 -> 53.7 M shots/s (instruction-cache misses
// were 14 % of the stalls).  LUF_OPT_NOUNROLL_CA (default 1) does the same for the cellular-automaton sources: Macar
// 20.7 -> 21.9 M shots/s, Snowflake cfg4 27.2 -> 33.3 M decoding cycles/s.
#ifndef LUF_OPT_NOUNROLL
  #define LUF_OPT_NOUNROLL 1
#endif
#ifndef LUF_OPT_NOUNROLL_CA
  #define LUF_OPT_NOUNROLL_CA 1
#endif
#ifndef LUF_OPT_PHILOX_UNROLL
  #define LUF_OPT_PHILOX_UNROLL 0
#endif
#if LUF_OPT_NOUNROLL && !defined(LUF_HOST_EMU)
  #define LUF_NOUNROLL _Pragma("unroll 1")
#else
  #define LUF_NOUNROLL
#endif
#if LUF_OPT_NOUNROLL_CA && !defined(LUF_HOST_EMU)
  #define LUF_NOUNROLL_CA _Pragma("unroll 1")
#else
  #define LUF_NOUNROLL_CA
#endif

#if defined(LUF_HOST_EMU)
  #include <math.h>
  #define LUF_DEV static inline
  #define LUF_MEM inline
  static inline uint32_t luf_atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
  static inline uint32_t luf_atomic_sub(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o - v; return o; }
  static inline uint32_t luf_atomic_or(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o | v; return o; }
  static inline uint32_t luf_atomic_xor(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o ^ v; return o; }
  static inline uint32_t luf_atomic_and(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o & v; return o; }
  static inline int luf_ffs(uint32_t x) { return __builtin_ffs((int)x); }
  static inline uint32_t luf_mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
  static inline float luf_log2(float x) { return log2f(x); }
  static inline int luf_popc(uint32_t x) { return __builtin_popcount(x); }
  struct luf_u32x2 { uint32_t x, y; };
  struct luf_u32x4 { uint32_t x, y, z, w; };
  #define LUF_LDG2(p) (*(const luf_u32x2*)(p))
  // lanes run one after the other: the counts are stored by one phase and summed in the next
  static uint32_t luf_scan_scratch[1024];
  #define LUF_SCAN_PREP(expr) LUF_PHASE(luf_scan_scratch[lane] = (expr))
  static inline void luf_exscan(uint32_t c, int lane, int nl, uint32_t& pos, uint32_t& total)
  {
      (void)c; pos = 0; total = 0;
      LUF_NOUNROLL
      for (int j = 0; j < nl; ++j) { if (j < lane) pos += luf_scan_scratch[j]; total += luf_scan_scratch[j]; }
  }
  #define LUF_LDG(p) (*(p))
  // a phase = every lane of the tile runs `stmt`, then the tile synchronises
  #define LUF_PHASE(stmt) for (int lane = 0; lane < nl; ++lane) { stmt; }
  // tile-wide reductions of a value the lanes accumulated during a phase (the
  // emulation accumulates into one variable, so they are identities there)
  #define LUF_ANY(x) ((x) != 0)
  #define LUF_XOR(x) (x)
  #define LUF_SUM(x) (x)
  static inline uint32_t luf_atomic_min(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
  static inline uint32_t luf_atomic_cas(uint32_t* p, uint32_t expect, uint32_t v) { uint32_t o = *p; if (o == expect) *p = v; return o; }
  // 64-bit words of the wide layout (luf_uf_wide.cuh)
  static inline uint64_t luf_atomic_add(uint64_t* p, uint64_t v) { uint64_t o = *p; *p = o + v; return o; }
  static inline uint64_t luf_atomic_min(uint64_t* p, uint64_t v) { uint64_t o = *p; if (v < o) *p = v; return o; }
  static inline uint64_t luf_atomic_cas(uint64_t* p, uint64_t expect, uint64_t v) { uint64_t o = *p; if (o == expect) *p = v; return o; }
  #define LUF_LDG64(p) (*(p))
#else
  #define LUF_DEV __device__ __forceinline__
  #define LUF_MEM __device__ __forceinline__
  LUF_DEV uint32_t luf_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
  LUF_DEV uint32_t luf_atomic_sub(uint32_t* p, uint32_t v) { return atomicSub(p, v); }
  LUF_DEV uint32_t luf_atomic_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
  LUF_DEV uint32_t luf_atomic_xor(uint32_t* p, uint32_t v) { return atomicXor(p, v); }
  LUF_DEV uint32_t luf_atomic_and(uint32_t* p, uint32_t v) { return atomicAnd(p, v); }
  LUF_DEV int luf_ffs(uint32_t x) { return __ffs((int)x); }
  LUF_DEV uint32_t luf_mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
  LUF_DEV float luf_log2(float x) { return __log2f(x); }
  LUF_DEV int luf_popc(uint32_t x) { return __popc(x); }
  typedef uint2 luf_u32x2;
  typedef uint4 luf_u32x4;
  #define LUF_LDG2(p) __ldg((const uint2*)(p))
  #define LUF_SCAN_PREP(expr)
  // exclusive prefix sum of c over the lanes of the warp
  LUF_DEV void luf_exscan(uint32_t c, int lane, int nl, uint32_t& pos, uint32_t& total)
  {
      uint32_t x = c;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
      pos = x - c; total = __shfl_sync(0xffffffffu, x, 31); (void)nl;
  }
  #define LUF_LDG(p) __ldg(p)
  #define LUF_PHASE(stmt) { stmt; } __syncwarp();
  #define LUF_ANY(x) __any_sync(0xffffffffu, (x) != 0)
  #define LUF_XOR(x) __reduce_xor_sync(0xffffffffu, (x))
  #define LUF_SUM(x) __reduce_add_sync(0xffffffffu, (x))
  LUF_DEV uint32_t luf_atomic_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
  LUF_DEV uint32_t luf_atomic_cas(uint32_t* p, uint32_t expect, uint32_t v) { return atomicCAS(p, expect, v); }
  // 64-bit words of the wide layout (luf_uf_wide.cuh)
  LUF_DEV uint64_t luf_atomic_add(uint64_t* p, uint64_t v) { return atomicAdd((unsigned long long*)p, (unsigned long long)v); }
  LUF_DEV uint64_t luf_atomic_min(uint64_t* p, uint64_t v) { return atomicMin((unsigned long long*)p, (unsigned long long)v); }
  LUF_DEV uint64_t luf_atomic_cas(uint64_t* p, uint64_t expect, uint64_t v) { return atomicCAS((unsigned long long*)p, (unsigned long long)expect, (unsigned long long)v); }
  #define LUF_LDG64(p) ((uint64_t)__ldg((const unsigned long long*)(p)))
#endif

// Optional event counters of the host emulation (tools/uf_stats.py); compiled out everywhere else.
#if defined(LUF_HOST_EMU) && defined(LUF_STATS)
  extern "C" { unsigned long long luf_stats[32]; }
  #define LUF_STAT(i, x) (luf_stats[i] += (unsigned long long)(x))
#else
  #define LUF_STAT(i, x) ((void)0)
#endif

namespace luf {

typedef uint16_t idx_t;                 // node / edge index inside a shot's state
// One 32-bit word per node, S[v]:
//   root      bit 15 set, bits 0..14 = cluster size, bit 31 = odd defect parity,
//             bits 16..30 = boundary node of the cluster + 1 (0 = none)
//   non-root  bit 15 clear, bits 0..14 = parent, bits 16..31 = the edge whose
//             union made it a non-root (the whole tree of a two-node cluster)
static const uint32_t ROOT_BIT = 0x8000u;
static const uint32_t ODD_BIT = 0x80000000u;
static const uint32_t BND_MASK = 0x7FFF0000u;
static const uint32_t SIZE_MASK = 0x7FFFu;
static const uint32_t NONE16 = 0xFFFFu;
static const uint32_t TREE_ROOT = 0xFFFEu;
static const uint32_t ADJ_PAD = 0xFFFFFFFFu;
static const int MAX_CLASSES = 32;
static const int CLAIM_SLOTS = 128;      // power of two (warp tiles)
#ifndef LUF_LIST_CAP
  #define LUF_LIST_CAP 96
#endif
#ifndef LUF_LIST2_CAP
  #define LUF_LIST2_CAP 64
#endif
static const uint32_t LIST_CAP = LUF_LIST_CAP;     // compacted bitmap (defects / touched nodes / newly full edges)
static const uint32_t LIST2_CAP = LUF_LIST2_CAP;   // active nodes of a round / tree roots (tests shrink both to force the overflow paths)
static const uint32_t LIST_DONE = 0xFFFFFFFFu;

enum { G_UNGROWN = 0, G_HALF = 1, G_FULL = 2, G_BURNT = 3 };
enum { FLAG_DYNAMIC = 1, FLAG_WEST = 2, FLAG_NODE = 4, FLAG_BUCKET = 8 };

// Read-only graph tables (global memory, read through L1).
struct GraphView {
    int N, B, E;            // nodes, boundary nodes (indices [0,B)), edges
    int WE, WD, WN;         // 32-bit words of an edge / detector / node bitmap
    int DEG;                // row length of adj (maximum degree rounded up to even)
    const uint32_t* adj;        // [N][DEG] edge | other<<16 | first<<31, ascending edge id; ADJ_PAD fills a row
    const uint32_t* edge_uv;    // [E]   u | v<<16
    const uint32_t* west_mask;  // [WE]  edges whose first endpoint is a west boundary node
    const int16_t* node_long;   // [N]   coordinate along the long axis
    int parity_ok;              // the graph allows the parity-only peel of the fused kernels (pack_graph, luf_host.h)
    int leaf_bnd;               // ... and every boundary node has exactly one edge (pack_fast): then the peel is never
                                // needed for the logical bit, not even when a cluster holds boundary nodes of both sides
    // sampler tables: fresh edges sorted by probability class
    int F, n_classes;
    const uint16_t* fresh_perm; // [F]; null when it is the identity (one class, every edge fresh)
    const uint32_t* class_end;  // [n_classes] exclusive end position of each class in fresh_perm
};

// Per-class sampling parameters of one launch (kernel argument, by value).
struct SampleParams {
    float p[MAX_CLASSES];           // flip probability
    float inv_log2_q[MAX_CLASSES];  // 1 / log2(1 - p)  (negative); unused when p <= 0 or p >= 1
    uint32_t key0, key1;            // Philox key = seed
};

// Byte offsets of one shot's state inside its shared-memory slice.
struct Layout {
    uint32_t S;                         // uint32[4*ceil(N/4)]  node words (the peel re-uses them: tree edge | stack link << 16)
    uint32_t growth;                    // uint32[ceil(E/16)]  2 bits per edge
    uint32_t newfull;                   // uint32[max(WE,WN)]  edges that reached FULL this round; tree roots during the peel
    uint32_t defect;                    // uint32[WD]
    uint32_t full1, full2;              // uint32[WN] each: touched nodes (a defect, or an end of a fully grown edge) /
                                        // nodes with two or more fully grown edges (defects: one or more)
    uint32_t err, corr;                 // uint32[WE]  only in the staged kernels (0 = absent)
    uint32_t zero_off, zero_bytes;      // growth .. corr are contiguous: what ph_reset clears
    uint32_t claim;                     // uint32[max(CLAIM_SLOTS,WN)]  merge reservations; visited bitmap of the peel
    uint32_t list, list2, cnt;          // uint32[LIST_CAP], uint32[LIST2_CAP] work lists; uint32[8]: [0..3] their fill counters,
                                        // [4] set once a union joined clusters whose boundary nodes lie on different sides
    uint32_t bytes;                     // slice size, multiple of 16
};

struct Shot {
    uint32_t *S, *growth, *newfull, *defect, *full1, *full2, *err, *corr, *claim, *list, *list2, *cnt;
};

LUF_DEV Shot bind(const Layout& l, unsigned char* base)
{
    Shot s;
    s.S = (uint32_t*)(base + l.S);
    s.growth = (uint32_t*)(base + l.growth); s.newfull = (uint32_t*)(base + l.newfull);
    s.defect = (uint32_t*)(base + l.defect); s.full1 = (uint32_t*)(base + l.full1);
    s.full2 = (uint32_t*)(base + l.full2);
    s.err = l.err ? (uint32_t*)(base + l.err) : 0;
    s.corr = l.corr ? (uint32_t*)(base + l.corr) : 0;
    s.claim = (uint32_t*)(base + l.claim);
    s.list = (uint32_t*)(base + l.list); s.list2 = (uint32_t*)(base + l.list2); s.cnt = (uint32_t*)(base + l.cnt);
    return s;
}

// ------------------------------------------------------------------ helpers
LUF_DEV uint32_t growth_of(const Shot& s, uint32_t e) { return (s.growth[e >> 4] >> ((e & 15u) * 2u)) & 3u; }
LUF_DEV bool bit_of(const uint32_t* bm, uint32_t i) { return (bm[i >> 5] >> (i & 31u)) & 1u; }

// --------------------------------------------------- phase 1: sample + syndrome
struct Philox {
    uint32_t c0, c1, c2, c3, k0, k1;    // counter, key
    uint32_t o0, o1, o2, o3;            // buffered outputs (shifted out, no indexing -> registers)
    int have;
};

// Philox4x32-10 (Salmon et al., SC'11): counter-based, so shot k's stream does
// not depend on which warp, launch or GPU draws it.
LUF_DEV void philox_block(Philox& r)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t c0 = r.c0, c1 = r.c1, c2 = r.c2, c3 = r.c3, k0 = r.k0, k1 = r.k1;
    // LUF_OPT_PHILOX_UNROLL: the ten rounds as straight-line code (8 instructions a round instead of 12 with the
    // loop counter and branch)
#if LUF_OPT_PHILOX_UNROLL && !defined(LUF_HOST_EMU)
    #pragma unroll
#else
    LUF_NOUNROLL
#endif
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = luf_mulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = luf_mulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r.o0 = c0; r.o1 = c1; r.o2 = c2; r.o3 = c3;
    r.c3 += 1u;             // block counter
    r.have = 4;
}

LUF_DEV uint32_t philox_next(Philox& r)
{
    if (r.have == 0) philox_block(r);
    const uint32_t v = r.o0;
    r.o0 = r.o1; r.o1 = r.o2; r.o2 = r.o3;
    --r.have;
    return v;
}

// Flip edge e: record it, toggle the defect bit of its detector endpoints
// (_base_classes.py:427-450), and return whether it counts for the logical
// parity (_schemes.py:127-128).
LUF_DEV uint32_t flip_edge(const GraphView& g, const Shot& s, uint32_t e)
{
    if (s.err) luf_atomic_xor(&s.err[e >> 5], 1u << (e & 31u));
    const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
    const uint32_t u = uv & 0xFFFFu, v = uv >> 16;
    if (u >= (uint32_t)g.B) luf_atomic_xor(&s.defect[(u - g.B) >> 5], 1u << ((u - g.B) & 31u));
    if (v >= (uint32_t)g.B) luf_atomic_xor(&s.defect[(v - g.B) >> 5], 1u << ((v - g.B) & 31u));
    return (LUF_LDG(&g.west_mask[e >> 5]) >> (e & 31u)) & 1u;
}

// A list of edges sorted by probability class (the freshly exposed edges of a
// shot / cycle, or the buffer region of a forward window).
struct SampleTab {
    int F, n_classes;
    const uint16_t* perm;       // [F]
    const uint32_t* class_end;  // [n_classes]
    const uint32_t* perm32;     // [F] instead of perm, for graphs of more than 65535 edges (compact fast path only); else null
};

// Each lane owns a contiguous range of the class-sorted edge list and walks it
// with geometric skips: the gap to the next flipped edge of a class with flip
// probability p is floor(log(U)/log(1-p)).  Same distribution as the
// reference's one Bernoulli draw per edge, ~p*F + classes draws per lane
// instead of F/NL.  `cycle` decorrelates the sheets of a streaming window;
// edges whose bit is set in `skip_mask` (optional) are never flipped
// (exclude_future_boundary, _base_classes.py:418-424).  Returns this lane's
// parity of flipped west-boundary edges.
template <class Flip>
LUF_DEV uint32_t ph_sample_core(const SampleTab& t, const SampleParams& sp, uint64_t shot, uint32_t cycle,
                                const uint32_t* skip_mask, int lane, int nl, Flip flip)
{
    const uint32_t a = (uint32_t)(((uint64_t)t.F * (uint32_t)lane) / (uint32_t)nl);
    const uint32_t b = (uint32_t)(((uint64_t)t.F * (uint32_t)(lane + 1)) / (uint32_t)nl);
    if (a >= b) return 0;
    Philox rng;
    rng.c0 = (uint32_t)shot; rng.c1 = (uint32_t)(shot >> 32); rng.c2 = (uint32_t)lane; rng.c3 = cycle << 6;
    rng.k0 = sp.key0; rng.k1 = sp.key1; rng.have = 0;
    rng.o0 = rng.o1 = rng.o2 = rng.o3 = 0;
    uint32_t parity = 0, start = 0;
    LUF_NOUNROLL
    for (int c = 0; c < t.n_classes; ++c) {
        const uint32_t end = LUF_LDG(&t.class_end[c]);
        const uint32_t lo = a > start ? a : start, hi = b < end ? b : end;
        start = end;
        if (lo >= hi) { if (end >= b) break; continue; }
        const float p = sp.p[c];
        if (p <= 0.0f) continue;
        uint32_t pos = lo;
        LUF_NOUNROLL
        for (;;) {
            if (p < 1.0f) {
                const float u = ((float)philox_next(rng) + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
                const float gf = luf_log2(u) * sp.inv_log2_q[c];
                const uint32_t gap = gf >= 1.0e9f ? 1000000000u : (uint32_t)gf;
                if (gap >= hi - pos) break;
                pos += gap;
            }
            const uint32_t e = t.perm32 ? LUF_LDG(&t.perm32[pos]) : t.perm ? (uint32_t)LUF_LDG(&t.perm[pos]) : pos;   // no table when the order is the identity
            if (!skip_mask || !((LUF_LDG(&skip_mask[e >> 5]) >> (e & 31u)) & 1u)) parity ^= flip(e);
            if (++pos >= hi) break;
        }
    }
    return parity;
}

// `flip(e)` records a flipped edge and returns whether it counts for the logical parity: flip_edge on the
// full tables here, fast::f_flip on the compact ones (luf_uf_fast.cuh) -- one sampler, one random stream.
LUF_DEV uint32_t ph_sample_tab(const GraphView& g, const SampleTab& t, const Shot& s, const SampleParams& sp,
                               uint64_t shot, uint32_t cycle, const uint32_t* skip_mask, int lane, int nl)
{
    return ph_sample_core(t, sp, shot, cycle, skip_mask, lane, nl, [&](uint32_t e) { return flip_edge(g, s, e); });
}

LUF_DEV uint32_t ph_sample(const GraphView& g, const Shot& s, const SampleParams& sp,
                           uint64_t shot, int lane, int nl)
{
    SampleTab t;
    t.F = g.F; t.n_classes = g.n_classes; t.perm = g.fresh_perm; t.class_end = g.class_end; t.perm32 = 0;
    return ph_sample_tab(g, t, s, sp, shot, 0u, (const uint32_t*)0, lane, nl);
}

// Fixed-weight sampling (noise/main.py:116-122, _Uniform.force_error =
// random.sample(FRESH_EDGES, weight)): exactly `weight` distinct fresh edges,
// every subset equally likely.  Floyd's algorithm: for j = F - weight .. F - 1
// draw t uniformly from [0, j]; take t, or j if t was taken already.  The error
// bitmap itself records what was taken, so the staged sampler layout suffices.
// One lane (the picks are a dependent chain; weights of interest are small).
// Index draws are mulhi(r32, j + 1): relative bias <= (j + 1) / 2^32 < 2e-5.
LUF_DEV void ph_force(const GraphView& g, const Shot& s, const SampleParams& sp, uint32_t weight,
                      uint64_t shot, int lane)
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
        const uint32_t et = g.fresh_perm ? (uint32_t)LUF_LDG(&g.fresh_perm[t]) : t;
        const uint32_t e = !((s.err[et >> 5] >> (et & 31u)) & 1u) ? et : g.fresh_perm ? (uint32_t)LUF_LDG(&g.fresh_perm[j]) : j;
        flip_edge(g, s, e);
    }
}

}  // namespace luf

// ------------------------------------------------------------- the UF phases
// Tile = one warp (the default everywhere: luf::ph_reset, luf::uf_validate, ...).
#define LUF_UF_NS uf_warp
#define LUF_T_LIST_CAP ((uint32_t)LUF_LIST_CAP)
#define LUF_T_LIST2_CAP ((uint32_t)LUF_LIST2_CAP)
#define LUF_T_CLAIM_SLOTS 128
#define LUF_T_PHASE(stmt) LUF_PHASE(stmt)
#define LUF_T_ANY(x) LUF_ANY(x)
#define LUF_T_SUM(x) LUF_SUM(x)
#define LUF_T_EXSCAN(c, lane, nl, pos, total) luf_exscan(c, lane, nl, pos, total)
#define LUF_T_SCAN_PREP(expr) LUF_SCAN_PREP(expr)
#include "luf_uf_phases.inl"
#undef LUF_UF_NS
#undef LUF_T_LIST_CAP
#undef LUF_T_LIST2_CAP
#undef LUF_T_CLAIM_SLOTS
#undef LUF_T_PHASE
#undef LUF_T_ANY
#undef LUF_T_SUM
#undef LUF_T_EXSCAN
#undef LUF_T_SCAN_PREP
namespace luf { using namespace uf_warp; }

// Tile = one thread block, one shot per block: graphs whose shot state would
// leave only a few warps per SM.  Barriers, votes and scans span the block.
#ifndef LUF_CTA_LIST_CAP
  #define LUF_CTA_LIST_CAP 2048
#endif
#ifndef LUF_CTA_LIST2_CAP
  #define LUF_CTA_LIST2_CAP 1024
#endif
#define LUF_CTA_CLAIM_SLOTS 1024
#if defined(LUF_HOST_EMU)
  #define LUF_T_PHASE(stmt) LUF_PHASE(stmt)
  #define LUF_T_ANY(x) LUF_ANY(x)
  #define LUF_T_SUM(x) LUF_SUM(x)
  #define LUF_T_EXSCAN(c, lane, nl, pos, total) luf_exscan(c, lane, nl, pos, total)
  #define LUF_T_SCAN_PREP(expr) LUF_SCAN_PREP(expr)
#else
  // block-wide sum / exclusive scan through a small static shared array (at most 32 warps)
  LUF_DEV uint32_t luf_block_sum(uint32_t x)
  {
      __shared__ uint32_t red[32];
      x = __reduce_add_sync(0xffffffffu, x);
      if ((threadIdx.x & 31u) == 0) red[threadIdx.x >> 5] = x;
      __syncthreads();
      uint32_t t = 0;
      LUF_NOUNROLL
      for (unsigned w = 0; w < (blockDim.x + 31u) >> 5; ++w) t += red[w];
      __syncthreads();
      return t;
  }
  LUF_DEV void luf_block_exscan(uint32_t c, int lane, int nl, uint32_t& pos, uint32_t& total)
  {
      __shared__ uint32_t red[32];
      uint32_t x = c;
      #pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((lane & 31) >= d) x += y; }
      if ((lane & 31) == 31) red[lane >> 5] = x;
      __syncthreads();
      uint32_t before = 0, all = 0;
      LUF_NOUNROLL
      for (int w = 0; w < (nl + 31) >> 5; ++w) { const uint32_t r = red[w]; if (w < (lane >> 5)) before += r; all += r; }
      __syncthreads();
      pos = before + x - c; total = all;
  }
  #define LUF_T_PHASE(stmt) { stmt; } __syncthreads();
  #define LUF_T_ANY(x) __syncthreads_or((x) != 0)
  #define LUF_T_SUM(x) luf_block_sum(x)
  #define LUF_T_EXSCAN(c, lane, nl, pos, total) luf_block_exscan(c, lane, nl, pos, total)
  #define LUF_T_SCAN_PREP(expr)
#endif
#define LUF_UF_NS uf_cta
#define LUF_T_LIST_CAP ((uint32_t)LUF_CTA_LIST_CAP)
#define LUF_T_LIST2_CAP ((uint32_t)LUF_CTA_LIST2_CAP)
#define LUF_T_CLAIM_SLOTS LUF_CTA_CLAIM_SLOTS
#include "luf_uf_phases.inl"
#undef LUF_UF_NS
#undef LUF_T_LIST_CAP
#undef LUF_T_LIST2_CAP
#undef LUF_T_CLAIM_SLOTS
#undef LUF_T_PHASE
#undef LUF_T_ANY
#undef LUF_T_SUM
#undef LUF_T_EXSCAN
#undef LUF_T_SCAN_PREP
