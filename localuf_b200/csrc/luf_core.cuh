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
  #define LUF_LDG(p) (*(p))
  // a phase = every lane of the tile runs `stmt`, then the tile synchronises
  #define LUF_PHASE(stmt) for (int lane = 0; lane < nl; ++lane) { stmt; }
  // tile-wide reductions of a value the lanes accumulated during a phase (the
  // emulation accumulates into one variable, so they are identities there)
  #define LUF_ANY(x) ((x) != 0)
  #define LUF_XOR(x) (x)
  #define LUF_BALLOT(var, pred) { var = 0; for (int lane = 0; lane < nl; ++lane) if (pred) var |= 1u << lane; }
  static inline uint32_t luf_atomic_min(uint32_t* p, uint32_t v) { uint32_t o = *p; if (v < o) *p = v; return o; }
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
  #define LUF_LDG(p) __ldg(p)
  #define LUF_PHASE(stmt) { stmt; } __syncwarp();
  #define LUF_ANY(x) __any_sync(0xffffffffu, (x) != 0)
  #define LUF_XOR(x) __reduce_xor_sync(0xffffffffu, (x))
  #define LUF_BALLOT(var, pred) { var = __ballot_sync(0xffffffffu, (pred)); }
  LUF_DEV uint32_t luf_atomic_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
#endif

namespace luf {

typedef uint16_t idx_t;                 // node / edge index inside a shot's state
static const uint32_t ROOT_BIT = 0x8000u;   // P[v]: root marker; low 15 bits = cluster size
static const uint32_t ODD_BIT = 0x8000u;    // R[root]: odd defect parity; low 15 bits = boundary node + 1
static const uint32_t NONE16 = 0xFFFFu;
static const uint32_t TREE_ROOT = 0xFFFEu;
static const int MAX_CLASSES = 32;
static const int CLAIM_SLOTS = 128;      // power of two
#ifndef LUF_LIST_CAP
  #define LUF_LIST_CAP 80
#endif
static const uint32_t LIST_CAP = LUF_LIST_CAP;   // work-list capacity (tests shrink it to force the overflow paths)
static const uint32_t LIST_DONE = 0xFFFFFFFFu;

enum { G_UNGROWN = 0, G_HALF = 1, G_FULL = 2, G_BURNT = 3 };
enum { FLAG_DYNAMIC = 1, FLAG_WEST = 2 };

// Read-only graph tables (global memory, read through L1).
struct GraphView {
    int N, B, E;            // nodes, boundary nodes (indices [0,B)), edges
    int WE, WD, WN;         // 32-bit words of an edge / detector / node bitmap
    const uint32_t* inc_off;    // [N+1] CSR offsets
    const uint32_t* inc;        // [2E]  edge | other<<16 | first<<31, ascending edge id per node
    const uint32_t* edge_uv;    // [E]   u | v<<16
    const uint32_t* west_mask;  // [WE]  edges whose first endpoint is a west boundary node
    const int16_t* node_long;   // [N]   coordinate along the long axis
    // sampler tables: fresh edges sorted by probability class
    int F, n_classes;
    const uint16_t* fresh_perm; // [F]
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
    uint32_t P, R;                      // idx_t[N] each (aliased by the peel: tree edge, stack link)
    uint32_t growth;                    // uint32[ceil(E/16)]  2 bits per edge
    uint32_t newfull;                   // uint32[WE]  edges that reached FULL this round
    uint32_t defect;                    // uint32[WD]
    uint32_t touched;                   // uint32[WN]  nodes that are a defect or belong to a cluster of size > 1
    uint32_t troot;                     // uint32[WN]  tree roots to peel from (aliases newfull, which is idle by then)
    uint32_t deg;                       // uint32[ceil(N/16)]  per node: has >= 1 / >= 2 fully grown incident edges
    uint32_t err, corr;                 // uint32[WE]  only in the staged kernels (0 = absent)
    uint32_t claim;                     // uint32[CLAIM_SLOTS]  merge reservations (see ph_merge_claim)
    uint32_t list, cnt;                 // uint32[LIST_CAP] work list shared by the phases that balance work over lanes; uint32[4] its fill counter
    uint32_t bytes;                     // slice size, multiple of 16
};

struct Shot {
    idx_t *P, *R;
    uint32_t *growth, *newfull, *defect, *touched, *troot, *deg, *err, *corr, *claim, *list, *cnt;
};

LUF_DEV Shot bind(const Layout& l, unsigned char* base)
{
    Shot s;
    s.P = (idx_t*)(base + l.P); s.R = (idx_t*)(base + l.R);
    s.growth = (uint32_t*)(base + l.growth); s.newfull = (uint32_t*)(base + l.newfull);
    s.defect = (uint32_t*)(base + l.defect); s.touched = (uint32_t*)(base + l.touched);
    s.troot = (uint32_t*)(base + l.troot); s.deg = (uint32_t*)(base + l.deg);
    s.err = l.err ? (uint32_t*)(base + l.err) : 0;
    s.corr = l.corr ? (uint32_t*)(base + l.corr) : 0;
    s.claim = (uint32_t*)(base + l.claim);
    s.list = (uint32_t*)(base + l.list); s.cnt = (uint32_t*)(base + l.cnt);
    return s;
}

// ------------------------------------------------------------------ helpers
LUF_DEV uint32_t growth_of(const Shot& s, uint32_t e) { return (s.growth[e >> 4] >> ((e & 15u) * 2u)) & 3u; }

// find without compression: safe while other lanes read the forest
LUF_DEV uint32_t find_ro(const Shot& s, uint32_t v)
{
    uint32_t p;
    while (!((p = s.P[v]) & ROOT_BIT)) v = p;
    return v;
}

// find that also shortcuts the start node to the root it found.  Safe next to
// concurrent finds and unions of other lanes: a non-root node never becomes a
// root again and the value written is one of its ancestors (16-bit stores are
// atomic), so every interleaving leaves a valid forest with the same roots.
LUF_DEV uint32_t find_pc(const Shot& s, uint32_t v)
{
    uint32_t p = s.P[v];
    if (p & ROOT_BIT) return v;
    uint32_t r = p, q;
    int hops = 0;
    while (!((q = s.P[r]) & ROOT_BIT)) { r = q; ++hops; }
    if (hops) s.P[v] = (idx_t)r;
    return r;
}

// uf.py:247-254 -- find with full path compression (single lane only)
LUF_DEV uint32_t find_compress(const Shot& s, uint32_t v)
{
    uint32_t r = v, p;
    while (!((p = s.P[r]) & ROOT_BIT)) r = p;
    while (v != r) { p = s.P[v]; s.P[v] = (idx_t)r; v = p; }
    return r;
}

// ------------------------------------------------------------ phase 0: reset
// uf.py:79-96,496-519 -- every node a singleton cluster; a boundary node is its
// own cluster boundary.  (The reference re-allocates N Python objects here,
// 40-58 % of its time; this is ~N*4 bytes of shared-memory stores.)
LUF_DEV void ph_reset(const GraphView& g, const Layout& l, const Shot& s, int lane, int nl)
{
    uint32_t* P2 = (uint32_t*)s.P;
    uint32_t* R2 = (uint32_t*)s.R;
    const uint32_t root1 = (ROOT_BIT | 1u) * 0x10001u;               // two singleton roots per 32-bit store
    for (int v = 2 * lane; v < g.N; v += 2 * nl) {
        P2[v >> 1] = root1;
        R2[v >> 1] = (v < g.B ? (uint32_t)v + 1u : 0u) | ((v + 1 < g.B ? (uint32_t)v + 2u : 0u) << 16);
    }
    const int GW = (g.E + 15) >> 4;
    for (int w = lane; w < GW; w += nl) s.growth[w] = 0;
    for (int w = lane; w < g.WE; w += nl) {
        s.newfull[w] = 0;
        if (s.err) s.err[w] = 0;
        if (s.corr) s.corr[w] = 0;
    }
    for (int w = lane; w < g.WD; w += nl) s.defect[w] = 0;
    for (int w = lane; w < g.WN; w += nl) { s.troot[w] = 0; s.touched[w] = 0; }
    for (int w = lane; w < ((g.N + 15) >> 4); w += nl) s.deg[w] = 0;
    for (int w = lane; w < CLAIM_SLOTS; w += nl) s.claim[w] = 0xFFFFFFFFu;
    if (lane < 4) s.cnt[lane] = 0;
    (void)l;
}

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
#if !defined(LUF_HOST_EMU)
    #pragma unroll
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
};

// Each lane owns a contiguous range of the class-sorted edge list and walks it
// with geometric skips: the gap to the next flipped edge of a class with flip
// probability p is floor(log(U)/log(1-p)).  Same distribution as the
// reference's one Bernoulli draw per edge, ~p*F + classes draws per lane
// instead of F/NL.  `cycle` decorrelates the sheets of a streaming window;
// edges whose bit is set in `skip_mask` (optional) are never flipped
// (exclude_future_boundary, _base_classes.py:418-424).  Returns this lane's
// parity of flipped west-boundary edges.
LUF_DEV uint32_t ph_sample_tab(const GraphView& g, const SampleTab& t, const Shot& s, const SampleParams& sp,
                               uint64_t shot, uint32_t cycle, const uint32_t* skip_mask, int lane, int nl)
{
    const uint32_t a = (uint32_t)(((uint64_t)t.F * (uint32_t)lane) / (uint32_t)nl);
    const uint32_t b = (uint32_t)(((uint64_t)t.F * (uint32_t)(lane + 1)) / (uint32_t)nl);
    if (a >= b) return 0;
    Philox rng;
    rng.c0 = (uint32_t)shot; rng.c1 = (uint32_t)(shot >> 32); rng.c2 = (uint32_t)lane; rng.c3 = cycle << 6;
    rng.k0 = sp.key0; rng.k1 = sp.key1; rng.have = 0;
    rng.o0 = rng.o1 = rng.o2 = rng.o3 = 0;
    uint32_t parity = 0, start = 0;
    for (int c = 0; c < t.n_classes; ++c) {
        const uint32_t end = LUF_LDG(&t.class_end[c]);
        const uint32_t lo = a > start ? a : start, hi = b < end ? b : end;
        start = end;
        if (lo >= hi) { if (end >= b) break; continue; }
        const float p = sp.p[c];
        if (p <= 0.0f) continue;
        uint32_t pos = lo;
        for (;;) {
            if (p < 1.0f) {
                const float u = ((float)philox_next(rng) + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
                const float gf = luf_log2(u) * sp.inv_log2_q[c];
                const uint32_t gap = gf >= 1.0e9f ? 1000000000u : (uint32_t)gf;
                if (gap >= hi - pos) break;
                pos += gap;
            }
            const uint32_t e = LUF_LDG(&t.perm[pos]);
            if (!skip_mask || !((LUF_LDG(&skip_mask[e >> 5]) >> (e & 31u)) & 1u)) parity ^= flip_edge(g, s, e);
            if (++pos >= hi) break;
        }
    }
    return parity;
}

LUF_DEV uint32_t ph_sample(const GraphView& g, const Shot& s, const SampleParams& sp,
                           uint64_t shot, int lane, int nl)
{
    SampleTab t;
    t.F = g.F; t.n_classes = g.n_classes; t.perm = g.fresh_perm; t.class_end = g.class_end;
    return ph_sample_tab(g, t, s, sp, shot, 0u, (const uint32_t*)0, lane, nl);
}

// uf.py:98-104 load: every defect is an odd singleton cluster (hence active).
LUF_DEV void ph_load(const GraphView& g, const Shot& s, int lane, int nl)
{
    for (int w = lane; w < g.WD; w += nl) {
        uint32_t bits = s.defect[w];
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t v = (uint32_t)(g.B + 32 * w + b);
            s.R[v] = (idx_t)ODD_BIT;
            luf_atomic_or(&s.touched[v >> 5], 1u << (v & 31u));
        }
    }
}

// ------------------------------------------------------- phase 2a: grow a round
// uf.py:221-222,234-245.  vision(C) = edges incident to C's members that are not
// yet FULL/BURNT; a cluster is active iff it is odd and has no boundary
// (uf.py:296-301), i.e. R[root] == ODD_BIT.  Members are not kept in lists:
// every *touched* node (defect, or member of a cluster of size > 1 -- singleton
// non-defects are never active) looks its root up and grows its own incident
// edges if the root is active.  An edge inside one cluster is in that cluster's
// vision once: only its first endpoint adds.  Two different active clusters
// each add 1.  The 2-bit counters saturate at FULL.  Returns whether this lane
// saw an active node.
// Two flag bits per node: "has a fully grown incident edge" / "has at least
// two".  Set with atomicOr only, so concurrent lanes cannot overflow a field.
LUF_DEV void deg_bump(const Shot& s, uint32_t v)
{
    const uint32_t sh = (v & 15u) * 2u;
    if ((luf_atomic_or(&s.deg[v >> 4], 1u << sh) >> sh) & 1u) luf_atomic_or(&s.deg[v >> 4], 2u << sh);
}
LUF_DEV bool deg_is_one(const Shot& s, uint32_t v) { return ((s.deg[v >> 4] >> ((v & 15u) * 2u)) & 3u) == 1u; }

LUF_DEV void grow_node(const GraphView& g, const Shot& s, uint32_t u, uint32_t r)
{
    const uint32_t k1 = LUF_LDG(&g.inc_off[u + 1]);
    for (uint32_t k = LUF_LDG(&g.inc_off[u]); k < k1; ++k) {
        const uint32_t ent = LUF_LDG(&g.inc[k]);
        const uint32_t e = ent & 0xFFFFu;
        const uint32_t sh = (e & 15u) * 2u;
        if (((s.growth[e >> 4] >> sh) & 3u) >= G_FULL) continue;
        if (!(ent >> 31)) {                       // second endpoint: skip if same cluster
            const uint32_t o = (ent >> 16) & 0x7FFFu;
            if (o == r || find_ro(s, o) == r) continue;
        }
        const uint32_t old = (luf_atomic_add(&s.growth[e >> 4], 1u << sh) >> sh) & 3u;
        if (old >= G_FULL) luf_atomic_sub(&s.growth[e >> 4], 1u << sh);     // saturate
        else if (old == G_HALF) {                                           // newly FULL
            luf_atomic_or(&s.newfull[e >> 5], 1u << (e & 31u));
            deg_bump(s, u); deg_bump(s, (ent >> 16) & 0x7FFFu);
        }
    }
}

// Gather the active nodes into the work list (cheap per touched node), so that
// the edge loops of ph_grow_exec are spread evenly over the lanes however the
// defects are distributed over the bitmap words.
LUF_DEV void ph_grow_gather(const GraphView& g, const Shot& s, int lane, int nl)
{
    for (int w = lane; w < g.WN; w += nl) {
        uint32_t bits = s.touched[w];
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t u = (uint32_t)(32 * w + b);
            const uint32_t r = find_pc(s, u);
            if (s.R[r] != ODD_BIT) continue;
            const uint32_t idx = luf_atomic_add(&s.cnt[0], 1u);
            if (idx < LIST_CAP) s.list[idx] = u | (r << 16);
            else grow_node(g, s, u, r);
        }
    }
}

// Returns whether any node was active this round.
LUF_DEV uint32_t ph_grow_exec(const GraphView& g, const Shot& s, int lane, int nl)
{
    const uint32_t total = s.cnt[0];
    const uint32_t n = total < LIST_CAP ? total : LIST_CAP;
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const uint32_t it = s.list[i];
        grow_node(g, s, it & 0xFFFFu, it >> 16);
    }
    return total;
}

// ------------------------------------------------------------ phase 2b: merge
// uf.py:223-229,256-301 in canonical order: this round's newly FULL edges are
// merged in ascending edge id.  Union by size, ties keep the cluster of the
// edge's first endpoint; parity XOR; boundary by inclination.  The smaller
// root's R slot is dead after the union and keeps the union edge (used by the
// peel for two-node clusters).
LUF_DEV void union_roots(const GraphView& g, const Shot& s, int flags, uint32_t e, uint32_t cu, uint32_t cv)
{
    if (flags & FLAG_DYNAMIC) {                                   // uf.py:261-266
        if (cu == cv || ((s.R[cu] & 0x7FFFu) && (s.R[cv] & 0x7FFFu))) {
            luf_atomic_or(&s.growth[e >> 4], 3u << ((e & 15u) * 2u));     // FULL(2) -> BURNT(3)
            return;
        }
    } else if (cu == cv) return;                                  // uf.py:256-259
    const uint32_t su = s.P[cu] & 0x7FFFu, sv = s.P[cv] & 0x7FFFu;
    const uint32_t L = su >= sv ? cu : cv, S = su >= sv ? cv : cu;    // uf.py:270-273
    const uint32_t rl = s.R[L], rs = s.R[S];
    uint32_t bl = rl & 0x7FFFu;
    const uint32_t bs = rs & 0x7FFFu;
    if (flags & FLAG_WEST) {                                      // uf.py:544-549
        if (bs && (!bl || LUF_LDG(&g.node_long[bs - 1]) < LUF_LDG(&g.node_long[bl - 1]))) bl = bs;
    } else if (!bl && bs) bl = bs;                                // uf.py:536-538
    s.R[L] = (idx_t)(((rl ^ rs) & ODD_BIT) | bl);
    s.R[S] = (idx_t)e;
    s.P[L] = (idx_t)(ROOT_BIT | (su + sv));
    s.P[S] = (idx_t)L;
    if (su == 1u) luf_atomic_or(&s.touched[cu >> 5], 1u << (cu & 31u));   // no longer a singleton
    if (sv == 1u) luf_atomic_or(&s.touched[cv >> 5], 1u << (cv & 31u));
}

// The serial order is reproduced in parallel by deterministic reservations:
// every pending edge writes (iteration tag, edge id) with atomicMin into a
// small table slot of each of its two current clusters; an edge that reads its
// own id back from both slots is the lowest-id pending edge at both clusters.
// No lower-id pending edge can touch -- or, through still lower edges, come to
// touch -- those clusters before it, so running it now commutes with
// everything the serial order would have run before it.  Winners of one
// iteration touch pairwise disjoint clusters and run concurrently; losers
// retry.  Slots are hashed (collisions only add ordering, never remove it) and
// tagged with a decreasing iteration number, so stale entries always lose and
// the table needs no clearing.
LUF_DEV uint32_t claim_slot(uint32_t root) { return (root * 0x9E5u >> 3) & (uint32_t)(CLAIM_SLOTS - 1); }

LUF_DEV void ph_merge_claim(const GraphView& g, const Shot& s, uint32_t tag, int lane, int nl)
{
    for (int w = lane; w < g.WE; w += nl) {
        uint32_t bits = s.newfull[w];
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t e = (uint32_t)(32 * w + b);
            const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
            const uint32_t cu = find_pc(s, uv & 0xFFFFu), cv = find_pc(s, uv >> 16);
            luf_atomic_min(&s.claim[claim_slot(cu)], tag | e);
            luf_atomic_min(&s.claim[claim_slot(cv)], tag | e);
        }
    }
}

// Returns whether this lane still has pending edges.
LUF_DEV uint32_t ph_merge_exec(const GraphView& g, const Shot& s, int flags, uint32_t tag, int lane, int nl)
{
    uint32_t pending = 0;
    for (int w = lane; w < g.WE; w += nl) {
        uint32_t bits = s.newfull[w], keep = 0;
        if (!bits) continue;
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t e = (uint32_t)(32 * w + b);
            const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
            const uint32_t cu = find_ro(s, uv & 0xFFFFu), cv = find_ro(s, uv >> 16);
            if (s.claim[claim_slot(cu)] == (tag | e) && s.claim[claim_slot(cv)] == (tag | e))
                union_roots(g, s, flags, e, cu, cv);
            else
                keep |= 1u << b;
        }
        s.newfull[w] = keep;
        pending |= keep;
    }
    return pending;
}

// What the first reservation round could not place (stars around one defect,
// chains, hash collisions) continues as further reservation rounds over a
// compact work list, one pending edge per lane: the lowest-id pending edge of
// every cluster wins each round, so a star of k edges takes k rounds, all
// stars side by side.
LUF_DEV void ph_rest_gather(const GraphView& g, const Shot& s, int lane, int nl)
{
    for (int w = lane; w < g.WE; w += nl) {
        uint32_t bits = s.newfull[w];
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t idx = luf_atomic_add(&s.cnt[1], 1u);
            if (idx < LIST_CAP) s.list[idx] = (uint32_t)(32 * w + b);
        }
    }
}

LUF_DEV void ph_rest_claim(const GraphView& g, const Shot& s, uint32_t n, uint32_t tag, int lane, int nl)
{
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const uint32_t it = s.list[i];
        if (it == LIST_DONE) continue;
        const uint32_t e = it & 0xFFFFu;
        const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
        const uint32_t cu = find_pc(s, uv & 0xFFFFu), cv = find_pc(s, uv >> 16);
        luf_atomic_min(&s.claim[claim_slot(cu)], tag | e);
        luf_atomic_min(&s.claim[claim_slot(cv)], tag | e);
        s.list[i] = e | (cu << 16);           // the root of the first endpoint cannot change before this edge runs
    }
}

LUF_DEV uint32_t ph_rest_exec(const GraphView& g, const Shot& s, int flags, uint32_t n, uint32_t tag, int lane, int nl)
{
    uint32_t pending = 0;
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const uint32_t it = s.list[i];
        if (it == LIST_DONE) continue;
        const uint32_t e = it & 0xFFFFu, cu = it >> 16;
        if (s.claim[claim_slot(cu)] != (tag | e)) { pending = 1; continue; }
        const uint32_t cv = find_ro(s, LUF_LDG(&g.edge_uv[e]) >> 16);
        if (s.claim[claim_slot(cv)] == (tag | e)) {
            union_roots(g, s, flags, e, cu, cv);
            s.list[i] = LIST_DONE;
        } else pending = 1;
    }
    return pending;
}

// Overflow path (more pending edges than the list holds): one lane merges them
// in ascending edge id, with path compression.  `mask` tells it which of the
// 32 words from `base` on still hold edges.
LUF_DEV void merge_rest(const GraphView& g, const Shot& s, int flags, int base, uint32_t mask)
{
    while (mask) {
        const int w = base + luf_ffs(mask) - 1; mask &= mask - 1;
        uint32_t bits = s.newfull[w];
        s.newfull[w] = 0;
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t e = (uint32_t)(32 * w + b);
            const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
            const uint32_t cu = find_compress(s, uv & 0xFFFFu), cv = find_compress(s, uv >> 16);
            union_roots(g, s, flags, e, cu, cv);
        }
    }
}

// ------------------------------------------------------------- phase 3: peel
// uf.py:314-324: one tree per cluster, rooted at cluster.boundary or else the
// union-find root.  Clusters without a defect cannot contribute a correction
// edge, so only clusters of defects are rooted.
//
// Two-node clusters (the common case at low noise: one flipped edge) need no
// traversal: their only tree is the union edge, kept in the R slot of the
// member that is not the root, and it is in the correction iff the cluster
// holds a defect.  Exactly one member acts: the non-root member if it is a
// defect; else the root, whose partner is then the cluster's boundary node
// (an odd cluster that stopped growing has one).
// Returns this lane's parity of pair corrections on the west boundary and
// (bit 1) whether any cluster needs the general traversal.
LUF_DEV uint32_t ph_tree_roots(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t out = 0;
    for (int w = lane; w < g.WD; w += nl) {
        uint32_t bits = s.defect[w];
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t dn = (uint32_t)(g.B + 32 * w + b);
            const uint32_t r = find_ro(s, dn);
            const uint32_t rr = s.R[r];
            if ((s.P[r] & 0x7FFFu) == 2u) {
                uint32_t x = dn;                                  // the member whose R slot holds the union edge
                if (dn == r) {
                    if (!(rr & ODD_BIT)) continue;                // both are defects: the other member acts
                    x = (rr & 0x7FFFu) - 1u;                      // single defect: its partner is the boundary node
                }
                const uint32_t e = s.R[x];
                if (s.corr) luf_atomic_xor(&s.corr[e >> 5], 1u << (e & 31u));
                out ^= (LUF_LDG(&g.west_mask[e >> 5]) >> (e & 31u)) & 1u;
                continue;
            }
            const uint32_t bn = rr & 0x7FFFu;
            const uint32_t t = bn ? bn - 1u : r;
            luf_atomic_or(&s.troot[t >> 5], 1u << (t & 31u));
            out |= 2u;
        }
    }
    return out;
}

// The union-find arrays are dead now; reuse them for the traversal:
//   P -> tree edge towards the root (NONE16 = undiscovered)    R -> link of the LIFO stack
// uf.py:326-344: the reference's LIFO traversal (neighbours in ascending edge
// id) from a tree root; only the tree (the edge towards the root of every
// discovered node) is kept.
LUF_DEV void build_tree(const GraphView& g, const Shot& s, uint32_t root)
{
    idx_t* tedge = s.P; idx_t* stk = s.R;
    tedge[root] = (idx_t)TREE_ROOT;
    stk[root] = (idx_t)NONE16;
    uint32_t top = root;
    while (top != NONE16) {
        const uint32_t u = top;
        top = stk[u];
        const uint32_t k1 = LUF_LDG(&g.inc_off[u + 1]);
        for (uint32_t k = LUF_LDG(&g.inc_off[u]); k < k1; ++k) {
            const uint32_t ent = LUF_LDG(&g.inc[k]);
            const uint32_t e = ent & 0xFFFFu;
            if (growth_of(s, e) != G_FULL) continue;
            const uint32_t o = (ent >> 16) & 0x7FFFu;
            if (tedge[o] != NONE16) continue;
            tedge[o] = (idx_t)e;
            if (deg_is_one(s, o)) continue;       // a leaf: nothing to discover from it
            stk[o] = (idx_t)top; top = o;
        }
    }
}

// one lane per tree root, roots handed out through the work list
LUF_DEV void ph_peel_gather(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t* P2 = (uint32_t*)s.P;
    for (int v = lane; 2 * v < g.N; v += nl) P2[v] = 0xFFFFFFFFu;        // every node undiscovered
    for (int w = lane; w < g.WN; w += nl) {
        uint32_t bits = s.troot[w], keep = 0;
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t idx = luf_atomic_add(&s.cnt[2], 1u);
            if (idx < LIST_CAP) s.list[idx] = (uint32_t)(32 * w + b);
            else keep |= 1u << b;                                        // overflow: stays in the bitmap
        }
        s.troot[w] = keep;
    }
}

LUF_DEV void ph_peel_build(const GraphView& g, const Shot& s, int lane, int nl)
{
    const uint32_t total = s.cnt[2];
    const uint32_t n = total < LIST_CAP ? total : LIST_CAP;
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) build_tree(g, s, s.list[i]);
    if (total > LIST_CAP)                                              // overflow roots, by bitmap word
        for (int w = lane; w < g.WN; w += nl) {
            uint32_t bits = s.troot[w];
            s.troot[w] = 0;
            while (bits) {
                const int b = luf_ffs(bits) - 1; bits &= bits - 1;
                build_tree(g, s, (uint32_t)(32 * w + b));
            }
        }
}

// uf.py:305-312: peeling the forest leaf-up puts a tree edge in the correction
// iff an odd number of defects hang below it -- the XOR over all defects of
// their paths to the root.  Every defect of a traversed cluster walks its own
// path (defects of two-node clusters were handled above and are undiscovered
// here).  Returns this lane's parity of correction edges on the west boundary.
LUF_DEV uint32_t ph_peel_paths(const GraphView& g, const Shot& s, int lane, int nl)
{
    const idx_t* tedge = s.P;
    uint32_t parity = 0;
    for (int w = lane; w < g.WD; w += nl) {
        uint32_t bits = s.defect[w];
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            uint32_t v = (uint32_t)(g.B + 32 * w + b);
            uint32_t e = tedge[v];
            if (e == NONE16) continue;
            while (e != TREE_ROOT) {
                if (s.corr) luf_atomic_xor(&s.corr[e >> 5], 1u << (e & 31u));
                parity ^= (LUF_LDG(&g.west_mask[e >> 5]) >> (e & 31u)) & 1u;
                const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
                v = (uv & 0xFFFFu) == v ? (uv >> 16) : (uv & 0xFFFFu);
                e = tedge[v];
            }
        }
    }
    return parity;
}

// ------------------------------------------------------------------ drivers
// uf.py:200-209 validate: rounds of grow / merge until no cluster is active.
// Returns the number of growth rounds.  `lane` is this thread's lane on the
// device; the host emulation iterates it inside LUF_PHASE.
LUF_DEV int uf_validate(const GraphView& g, const Shot& s, int flags, int lane, int nl)
{
    int rounds = 0;
    uint32_t iter = 0xFFFFu;                 // decreasing reservation tag
    (void)lane;
    for (;;) {
        uint32_t active = 0;
        LUF_PHASE(ph_grow_gather(g, s, lane, nl));
        LUF_PHASE(active |= ph_grow_exec(g, s, lane, nl));
        if (!LUF_ANY(active)) break;
        // one reservation round over the bitmap places the independent merges ...
        uint32_t tag = iter-- << 16, pending = 0;
        LUF_PHASE(ph_merge_claim(g, s, tag, lane, nl); if (lane == 0) s.cnt[0] = 0);
        LUF_PHASE(pending |= ph_merge_exec(g, s, flags, tag, lane, nl));
        if (LUF_ANY(pending)) {              // ... the rest goes on over a compact list
            LUF_PHASE(ph_rest_gather(g, s, lane, nl));
            uint32_t n = 0;
            LUF_PHASE(n = s.cnt[1]);
            if (n <= LIST_CAP) {
                LUF_PHASE(for (int w = lane; w < g.WE; w += nl) s.newfull[w] = 0; if (lane == 0) s.cnt[1] = 0);
                for (uint32_t guard = 0; guard <= n; ++guard) {       // at least one edge is placed per round
                    tag = iter-- << 16; pending = 0;
                    LUF_PHASE(ph_rest_claim(g, s, n, tag, lane, nl));
                    LUF_PHASE(pending |= ph_rest_exec(g, s, flags, n, tag, lane, nl));
                    if (!LUF_ANY(pending)) break;
                }
            } else {                         // overflow: one lane, ascending edge id
                LUF_PHASE(if (lane == 0) s.cnt[1] = 0);
                for (int base = 0; base < g.WE; base += 32) {
                    uint32_t mask;
                    LUF_BALLOT(mask, base + lane < g.WE && s.newfull[base + lane] != 0);
                    if (mask) { LUF_PHASE(if (lane == 0) merge_rest(g, s, flags, base, mask)); }
                }
            }
        }
        ++rounds;
    }
    LUF_PHASE(if (lane == 0) s.cnt[0] = 0);
    return rounds;
}

// uf.py:305-312 peel.  Returns (per lane on the device, total in the
// emulation) the parity of correction edges on the west boundary.
LUF_DEV uint32_t uf_peel(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t parity = 0, general = 0;
    (void)lane;
    LUF_PHASE(const uint32_t r = ph_tree_roots(g, s, lane, nl); parity ^= r & 1u; general |= r >> 1);
    if (LUF_ANY(general)) {
        LUF_PHASE(ph_peel_gather(g, s, lane, nl));
        LUF_PHASE(ph_peel_build(g, s, lane, nl));
        LUF_PHASE(parity ^= ph_peel_paths(g, s, lane, nl); if (lane == 0) s.cnt[2] = 0);
    }
    return parity;
}

}  // namespace luf
