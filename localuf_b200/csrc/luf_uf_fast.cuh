// luf_uf_fast.cuh -- the compact fast path of the fused Monte-Carlo kernel for static UF
// (decoders/uf.py:200-301 validate; _schemes.py:116-155 Batch.run), default or west inclination.
//
// What a Monte-Carlo shot needs is one bit: the parity of error ^ correction on the west boundary.  As long as
// no cluster comes to hold boundary nodes of BOTH sides, that bit does not depend on the order of the merges
// nor on the peeling forest (luf_uf_phases.inl, "order-free merge" and uf_peel_parity): it is the XOR, over the
// clusters that touch the west boundary, of their defect parity.  So this path keeps per cluster only
// (odd, touches west, touches east) and per node only a 16-bit word:
//
//   P[v]   (detectors only; boundary nodes have no entry: they have one edge each and an active cluster has
//          no boundary, so a boundary node is never looked up)
//          root      bit 15 set | bit 0 odd | bit 1 touches west | bit 2 touches east
//          non-root  the parent (a detector index < 32768)
//   growth 2 bits per edge (0 ungrown, 1 half, 2 full), as in the canonical path
//   tlist  the touched detectors in order of appearance (defects first): what later growth rounds visit
//   alist  the active touched nodes of a round (node | root << 16)
//   mlist  the edges that became fully grown this round, as endpoint pairs (detector | other end << 16, the
//          other end being a detector or END_WEST / END_EAST)
//
// No union by size: a root is hooked below the root of smaller index with a compare-and-swap (a static strict
// order, so no interleaving closes a cycle), taking its (odd, west, east) bits along; bits are added to live
// roots with compare-and-swap too, so a payload always ends up at the root of its tree whatever the
// interleaving.  Every change of a root word is one atomic transition, and each one adjusts the count of
// active clusters (odd, no boundary) by the difference it makes -- the round loop ends when that count is 0.
//
// A shot that ends with an ODD cluster touching both sides is reported as HARD under the default inclination
// (which boundary node such a cluster keeps, and so where its one string to the boundary goes, depends on the
// merge order; an even cluster on both sides adds nothing either way): with a merge log (further down) the unions
// inside those clusters are exported and replayed in canonical order, without one the caller queues the shot -- for
// a second pass with a log, or for the canonical kernel.  Under the west inclination (uf.py:544-549) such a cluster keeps a west boundary node; that node has
// one edge, every defect's path to the root runs through it, so the cluster adds its defect parity like any
// other that touches the west side -- no HARD shots at all.  A work list that overflows is not fatal either:
// its consumers walk the bitmap behind it (touched nodes, defects) or the growth words themselves (fully grown
// edges; joining an edge twice is harmless).  The state of a cfg3 shot is about 5 KB instead of 8.9 KB, and
// nothing in a shot's loops reads global memory when the adjacency table is staged in shared memory.
//
// The tile of `nl` cooperating lanes may be a part of a warp (8 / 16 lanes: several shots advance through the
// same instructions), a warp, or several warps of a block (named barrier); the only tile operation is the
// barrier between phases -- counts and votes go through the shot's counters in shared memory.
#pragma once
#include "luf_core.cuh"

namespace luf {
namespace fast {

static const uint32_t P_ROOT = 0x8000u, P_ODD = 1u, P_WEST = 2u, P_EAST = 4u;
static const uint32_t P_FRESH = P_ROOT;                  // an untouched detector: even singleton, no boundary
static const uint32_t P_ACTIVE = P_ROOT | P_ODD;         // odd and no boundary (uf.py:296-301)
static const uint32_t END_WEST = 0xFFFFu, END_EAST = 0xFFFEu;         // edge table / merge list: the end is a boundary node
static const int MAX_DETECTORS = 0x7FFE;                              // a non-root word is a parent index below 0x8000
// cnt[]: fill counters of the three lists (they only grow: a round's entries start at the previous total; a
// counter beyond its capacity says the list overflowed), the number of active clusters, the parity
// accumulator, and the HARD flag (an odd cluster on both sides; written by the last phase only, read after the
// barrier that ends it).
enum { C_TOUCHED = 0, C_ALIST = 1, C_MLIST = 2, C_ACTIVE = 3, C_XBASE = 4, C_PARITY = 5, C_HARD_M = 6, C_XCOUNT = 7 };
static const int C_XPOS = C_ALIST;        // (the active list is done with when a shot is exported)
static const uint32_t HARD = 2u;                         // result of a shot the canonical kernel has to decode
static const uint32_t HARD_LOGGED = 3u;                  // ... or whose odd clusters on both sides were exported for hard_resolve
static const uint32_t KEY_EDGE_BITS = 20u, KEY_EDGE_MASK = 0xFFFFFu;      // merge-log key = growth round << 20 | edge id

// Where the HARD clusters of a launch go (global memory): per shot `count | fast bit << 31` followed by `count` keys,
// reserved with one atomic add on *cursor; (shot, offset) pairs in `queue`, *qcount of them.
struct FExport {
    uint32_t* buf;
    uint64_t* cursor;           // (64 bits: reservations that find no room still add to it)
    uint32_t cap;               // words of buf
    uint32_t* queue;            // [2 * shots of the launch]
    uint32_t* qcount;
    uint32_t unsorted;          // 1: the log lies in global memory (a slab per resident tile, bound to FShot::elog by the
                                //    kernel) and the keys go out as they lie -- k_hard_sort orders them; 0: the log is
                                //    in the tile's shared memory, sorted there and exported as endpoint pairs
    uint32_t max_keys;          // unsorted: most keys per shot (the sort kernel's buffer)
};

// Read-only tables (global memory; fadj optionally staged in shared memory by the kernel).
struct FView {
    int D, E, WD, DEG;          // detectors, edges, 32-bit words of a detector bitmap, row length of fadj (even)
    uint32_t deg_magic;         // floor(2^32 / DEG) + 1: w / DEG == mulhi(w, deg_magic) for the item counts of a shot
    const uint32_t* fadj;       // [D][DEG] edge | other end << ebits | this is the edge's first endpoint << 31; other end =
                                // detector index, or the two top values of its field: omask = west, omask - 1 = east
    uint32_t ebits, emask, omask;   // 16 / 0xFFFF / 0x7FFF up to 65533 edges; graphs with more edges and fewer detectors shift the split
    const uint32_t* einfo;      // [E] first end | second end << 16 (detector index, END_WEST, END_EAST)
    int F, n_classes;
    const uint16_t* fresh_perm; // as GraphView
    const uint32_t* fresh_perm32;   // instead, for more than 65535 edges
    const uint32_t* class_end;
    int west;                   // west inclination: a cluster on both sides is not HARD
};

struct FLayout {
    uint32_t P, growth, defect, touched;        // P: uint16[D]; growth: uint32[ceil(E/16)]; bitmaps: uint32[WD]
    uint32_t fill_words, zero_off, zero_bytes;  // P is filled with P_FRESH pairs; growth .. touched are cleared
    uint32_t tlist, alist, mlist, cnt;          // uint16[tcap], uint32[acap], uint32[mcap], uint32[8]
    uint32_t tcap, acap, mcap;
    uint32_t elog, lcap;                        // uint32[lcap]: the merge log (every newly FULL edge of the shot, keyed by round); lcap = 0: none
    uint32_t bytes;                             // slice size, multiple of 16
};

struct FShot {
    uint16_t* P;
    uint32_t *growth, *defect, *touched;
    uint16_t* tlist;
    uint32_t *alist, *mlist, *cnt, *elog;
};

LUF_DEV FShot fbind(const FLayout& l, unsigned char* base)
{
    FShot s;
    s.P = (uint16_t*)(base + l.P);
    s.growth = (uint32_t*)(base + l.growth); s.defect = (uint32_t*)(base + l.defect);
    s.touched = (uint32_t*)(base + l.touched);
    s.tlist = (uint16_t*)(base + l.tlist);
    s.alist = (uint32_t*)(base + l.alist); s.mlist = (uint32_t*)(base + l.mlist); s.cnt = (uint32_t*)(base + l.cnt);
    s.elog = (uint32_t*)(base + l.elog);
    return s;
}

// ---------------------------------------------------------------- node words
// Compare-and-swap on a 16-bit node word through its 32-bit container (shared memory has no 16-bit
// compare-and-swap; nvcc emulates atomicCAS(unsigned short*) the same way).  Fails at once when the word is
// not `expect`; a change of the neighbouring half only repeats the attempt.
LUF_DEV bool cas16(uint16_t* P, uint32_t v, uint32_t expect, uint32_t val)
{
    uint32_t* wp = (uint32_t*)P + (v >> 1);
    const uint32_t sh = (v & 1u) * 16u;
    uint32_t cur = *(volatile uint32_t*)wp;
    LUF_NOUNROLL
    for (;;) {
        if (((cur >> sh) & 0xFFFFu) != expect) return false;
        const uint32_t nw = (cur & ~(0xFFFFu << sh)) | (val << sh);
        const uint32_t old = luf_atomic_cas(wp, cur, nw);
        if (old == cur) return true;
        cur = old;
    }
}

LUF_DEV uint32_t pload(const FShot& s, uint32_t v) { return *(volatile uint16_t*)&s.P[v]; }

// find without writes; `wr` = the root's word
LUF_DEV uint32_t ffind_ro(const FShot& s, uint32_t v, uint32_t& wr)
{
    uint32_t w;
    LUF_STAT(20, 1);
    LUF_NOUNROLL
    while (!((w = pload(s, v)) & P_ROOT)) { v = w; LUF_STAT(21, 1); }
    wr = w;
    return v;
}

// find that also points the start node at the root it found: safe next to concurrent finds, hooks and payload
// updates (a non-root never becomes a root again; the value stored is one of its ancestors)
LUF_DEV uint32_t ffind(const FShot& s, uint32_t v, uint32_t& wr)
{
    const uint32_t w = pload(s, v);
    LUF_STAT(22, 1);
    if (w & P_ROOT) { wr = w; return v; }
    uint32_t r = w, q;
    int hops = 0;
    LUF_STAT(23, 1);
    LUF_NOUNROLL
    while (!((q = pload(s, r)) & P_ROOT)) { r = q; ++hops; LUF_STAT(23, 1); }
    if (hops) s.P[v] = (uint16_t)r;
    wr = q;
    return r;
}

// Add (odd, west, east) bits to the root of v's tree.
LUF_DEV void add_bits(const FShot& s, uint32_t v, uint32_t bits, int west)
{
    LUF_NOUNROLL
    for (;;) {
        uint32_t w;
        const uint32_t r = ffind_ro(s, v, w);
        const uint32_t nw = (w ^ (bits & P_ODD)) | (bits & (P_WEST | P_EAST));
        if (nw == w) return;
        if (cas16(s.P, r, w, nw)) {
            const int delta = (nw == P_ACTIVE ? 1 : 0) - (w == P_ACTIVE ? 1 : 0);
            if (delta) luf_atomic_add(&s.cnt[C_ACTIVE], (uint32_t)delta);
            return;
        }
        v = r;
    }
}

// uf.py:256-280 without the order: join the clusters of detectors a and b.
LUF_DEV void join(const FShot& s, uint32_t a, uint32_t b, int west)
{
    LUF_NOUNROLL
    for (;;) {
        uint32_t wa, wb;
        uint32_t ra = ffind(s, a, wa), rb = ffind(s, b, wb);
        if (ra == rb) return;
        if (ra > rb) { uint32_t t = ra; ra = rb; rb = t; wb = wa; }      // the root of larger index goes below the other
        if (!cas16(s.P, rb, wb, ra)) continue;
        if (wb == P_ACTIVE) luf_atomic_sub(&s.cnt[C_ACTIVE], 1u);
        if (wb & (P_ODD | P_WEST | P_EAST)) add_bits(s, ra, wb, west);
        return;
    }
}

// ------------------------------------------------------------------- phases
LUF_DEV void f_reset(const FLayout& l, const FShot& s, int lane, int nl)
{
    luf_u32x4* P4 = (luf_u32x4*)s.P;
    luf_u32x4 f; f.x = f.y = f.z = f.w = P_FRESH | (P_FRESH << 16);
    LUF_NOUNROLL
    for (int i = lane; i < (int)(l.fill_words >> 2); i += nl) P4[i] = f;
    luf_u32x4* Z4 = (luf_u32x4*)((unsigned char*)s.P + (l.zero_off - l.P));
    luf_u32x4 z; z.x = z.y = z.z = z.w = 0u;
    LUF_NOUNROLL
    for (int i = lane; i < (int)(l.zero_bytes >> 4); i += nl) Z4[i] = z;
    if (lane == 0) { luf_u32x4* c4 = (luf_u32x4*)s.cnt; c4[0] = z; c4[1] = z; }       // lane 0 alone: it read the previous shot's result
}

// _base_classes.py:427-450 on the compact tables; returns whether the edge counts for the logical parity
LUF_DEV uint32_t f_flip(const FView& g, const FShot& s, uint32_t e)
{
    const uint32_t ei = LUF_LDG(&g.einfo[e]);
    const uint32_t a = ei & 0xFFFFu, b = ei >> 16;
    if (a < END_EAST) luf_atomic_xor(&s.defect[a >> 5], 1u << (a & 31u));
    if (b < END_EAST) luf_atomic_xor(&s.defect[b >> 5], 1u << (b & 31u));
    return a == END_WEST ? 1u : 0u;
}

// Staged K1 on the compact tables (graphs beyond the 16-bit layout have no other): the 32 virtual sampler lanes
// record the flipped edges in `err` (W(E) words) and toggle the defects (noise/main.py:114-115,309-314;
// _base_classes.py:427-450).  Same random stream as fast_shot and, where both exist, as ph_sample.
LUF_DEV void f_sample_record(const FView& g, uint32_t* err, uint32_t* defect, const SampleParams& sp, uint64_t shot,
                             int lane, int nl)
{
    SampleTab t; t.F = g.F; t.n_classes = g.n_classes; t.perm = g.fresh_perm; t.class_end = g.class_end; t.perm32 = g.fresh_perm32;
    for (int vl = lane; vl < 32; vl += nl)
        ph_sample_core(t, sp, shot, 0u, (const uint32_t*)0, vl, 32, [&](uint32_t e) {
            luf_atomic_xor(&err[e >> 5], 1u << (e & 31u));
            const uint32_t ei = LUF_LDG(&g.einfo[e]);
            const uint32_t a = ei & 0xFFFFu, b = ei >> 16;
            if (a < END_EAST) luf_atomic_xor(&defect[a >> 5], 1u << (a & 31u));
            if (b < END_EAST) luf_atomic_xor(&defect[b >> 5], 1u << (b & 31u));
            return 0u;
        });
}

// uf.py:98-104: every defect an odd singleton; the defects open the list of touched nodes
LUF_DEV void f_load(const FView& g, const FLayout& l, const FShot& s, int lane, int nl)
{
    LUF_NOUNROLL
    for (int w = lane; w < g.WD; w += nl) {
        uint32_t bits = s.defect[w];
        if (!bits) continue;
        s.touched[w] = bits;
        uint32_t pos = luf_atomic_add(&s.cnt[C_TOUCHED], (uint32_t)luf_popc(bits));
        LUF_NOUNROLL
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t v = (uint32_t)(32 * w + b);
            s.P[v] = (uint16_t)P_ACTIVE;
            if (pos < l.tcap) s.tlist[pos] = (uint16_t)v;
            ++pos;
        }
    }
}

template <bool ADJ_SHARED, class F>
LUF_DEV void f_row(const FView& g, uint32_t v, F f)
{
    const uint32_t* row = g.fadj + (size_t)v * (uint32_t)g.DEG;
    LUF_NOUNROLL
    for (int k = 0; k < g.DEG; k += 2) {
        luf_u32x2 p;
        if (ADJ_SHARED) p = *(const luf_u32x2*)(row + k); else p = LUF_LDG2(row + k);
        if (p.x == ADJ_PAD) break;
        f(p.x);
        if (p.y != ADJ_PAD) f(p.y);
    }
}

// `other16`: the other end as a detector index, END_WEST or END_EAST
// `key`: growth round << 20 | edge id, for the merge log (the order of the canonical merges: uf.py:223-229 takes a
// round's newly FULL edges in ascending edge id)
LUF_DEV void f_newfull(const FLayout& l, const FShot& s, uint32_t u, uint32_t other16, uint32_t mbase, uint32_t key)
{
    const uint32_t tot = luf_atomic_add(&s.cnt[C_MLIST], 1u), idx = tot - mbase;
    if (tot < l.lcap) s.elog[tot] = key;
    if (idx < l.mcap) s.mlist[idx] = u | (other16 << 16);
    // else: the counter alone says the list overflowed; the merge phase then reads the growth words
}

// Run f(v) for the first n entries of the touched list -- or, when the list could not hold them (total beyond
// its capacity), for every set bit of the bitmap behind it.
template <class F>
LUF_DEV void f_for_nodes(const FView& g, const FLayout& l, const FShot& s, const uint32_t* bitmap, uint32_t n, uint32_t total,
                         int lane, int nl, F f)
{
    if (total <= l.tcap) {
        LUF_NOUNROLL
        for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) f((uint32_t)s.tlist[i]);
    } else {
        LUF_NOUNROLL
        for (int w = lane; w < g.WD; w += nl) {
            uint32_t bits = bitmap[w];
            LUF_NOUNROLL
            while (bits) { const int b = luf_ffs(bits) - 1; bits &= bits - 1; f((uint32_t)(32 * w + b)); }
        }
    }
}

template <bool ADJ_SHARED>
LUF_DEV uint32_t f_entry(const FView& g, uint32_t v, uint32_t k)
{
    const uint32_t* p = g.fadj + (size_t)v * (uint32_t)g.DEG + k;
    return ADJ_SHARED ? *p : LUF_LDG(p);
}

LUF_DEV void f_grow_first_entry(const FView& g, const FLayout& l, const FShot& s, uint32_t v, uint32_t ent)
{
    const uint32_t e = ent & g.emask, sh = (e & 15u) * 2u;
    const uint32_t old = (luf_atomic_add(&s.growth[e >> 4], 1u << sh) >> sh) & 3u;
    if (old == G_HALF) f_newfull(l, s, v, (ent >> g.ebits) & g.omask, 0u, (1u << KEY_EDGE_BITS) | e);       // (the other end grew too: a detector)
}

// Round 1 (uf.py:234-245): every defect adds 1 to each of its edges, no look-ups; an edge is newly FULL iff the
// other end (a defect too, then) got there first.  One lane per (defect, adjacency slot): a shot's defects
// number about two per fault, fewer than a warp's lanes at the noise levels of interest, their slots several
// times more.
template <bool ADJ_SHARED>
LUF_DEV void f_grow_first(const FView& g, const FLayout& l, const FShot& s, uint32_t nd, int lane, int nl)
{
    if (nd <= l.tcap) {
        const uint32_t items = nd * (uint32_t)g.DEG;
        LUF_NOUNROLL
        for (uint32_t w = (uint32_t)lane; w < items; w += (uint32_t)nl) {
            const uint32_t i = luf_mulhi(w, g.deg_magic), v = s.tlist[i];
            const uint32_t ent = f_entry<ADJ_SHARED>(g, v, w - i * (uint32_t)g.DEG);
            if (ent != ADJ_PAD) f_grow_first_entry(g, l, s, v, ent);
        }
        return;
    }
    f_for_nodes(g, l, s, s.defect, nd, nd, lane, nl, [&](uint32_t v) {
        f_row<ADJ_SHARED>(g, v, [&](uint32_t ent) { f_grow_first_entry(g, l, s, v, ent); });
    });
}

LUF_DEV void f_touch(const FLayout& l, const FShot& s, uint32_t v)
{
    const uint32_t bit = 1u << (v & 31u);
    if (luf_atomic_or(&s.touched[v >> 5], bit) & bit) return;
    const uint32_t idx = luf_atomic_add(&s.cnt[C_TOUCHED], 1u);
    if (idx < l.tcap) s.tlist[idx] = (uint16_t)v;          // else: the consumers walk the touched bitmap
}

// Later rounds: a touched node whose cluster (root r) is active adds 1 to each of its edges that is not FULL;
// an edge inside one cluster is counted once (only its first endpoint adds); the 2-bit counters saturate.
LUF_DEV void f_grow_entry(const FView& g, const FLayout& l, const FShot& s, uint32_t u, uint32_t r, uint32_t ent, uint32_t mbase, uint32_t rkey)
{
    const uint32_t e = ent & g.emask, sh = (e & 15u) * 2u;
    if (((s.growth[e >> 4] >> sh) & 3u) >= G_FULL) return;
    const uint32_t o = (ent >> g.ebits) & g.omask;
    const bool detector = o < g.omask - 1u;
    if (!(ent >> 31) && detector) {                // second endpoint of an edge between detectors: same cluster?
        uint32_t wr;
        if (o == r || ffind_ro(s, o, wr) == r) return;
    }
    const uint32_t old = (luf_atomic_add(&s.growth[e >> 4], 1u << sh) >> sh) & 3u;
    if (old >= G_FULL) luf_atomic_sub(&s.growth[e >> 4], 1u << sh);
    else if (old == G_HALF) {
        f_newfull(l, s, u, detector ? o : (o == g.omask ? END_WEST : END_EAST), mbase, rkey | e);
        if (detector) f_touch(l, s, o);
    }
}

template <bool ADJ_SHARED>
LUF_DEV void f_grow_node(const FView& g, const FLayout& l, const FShot& s, uint32_t u, uint32_t r, uint32_t mbase, uint32_t rkey)
{
    f_row<ADJ_SHARED>(g, u, [&](uint32_t ent) { f_grow_entry(g, l, s, u, r, ent, mbase, rkey); });
}

template <bool ADJ_SHARED>
LUF_DEV void f_grow_filter(const FView& g, const FLayout& l, const FShot& s, uint32_t nt, uint32_t abase, uint32_t mbase,
                           uint32_t rkey, int lane, int nl)
{
    f_for_nodes(g, l, s, s.touched, nt, nt, lane, nl, [&](uint32_t u) {
        uint32_t wr;
        const uint32_t r = ffind(s, u, wr);
        if (wr != P_ACTIVE) return;
        const uint32_t idx = luf_atomic_add(&s.cnt[C_ALIST], 1u) - abase;
        if (idx < l.acap) s.alist[idx] = u | (r << 16);
        else f_grow_node<ADJ_SHARED>(g, l, s, u, r, mbase, rkey);       // the list is full: grow right away (roots are stable here)
    });
}

// one lane per (active node, adjacency slot)
template <bool ADJ_SHARED>
LUF_DEV void f_grow_exec(const FView& g, const FLayout& l, const FShot& s, uint32_t na, uint32_t mbase, uint32_t rkey, int lane, int nl)
{
    if (na > l.acap) na = l.acap;
    const uint32_t items = na * (uint32_t)g.DEG;
    LUF_NOUNROLL
    for (uint32_t w = (uint32_t)lane; w < items; w += (uint32_t)nl) {
        const uint32_t i = luf_mulhi(w, g.deg_magic), it = s.alist[i];
        const uint32_t ent = f_entry<ADJ_SHARED>(g, it & 0xFFFFu, w - i * (uint32_t)g.DEG);
        if (ent != ADJ_PAD) f_grow_entry(g, l, s, it & 0xFFFFu, it >> 16, ent, mbase, rkey);
    }
}

// uf.py:223-229 in any order, two phases over the list of newly FULL edges.
//   hook   one lane per edge: the root of larger index goes below the other (compare-and-swap on its word, which
//          is read back as the payload: the bits the hooked cluster carried); an edge to a boundary node has its
//          side as the payload.  The list entry becomes (node to credit | payload << 16 | valid).
//   apply  roots do not move in this phase, so a payload is added to the root of its node with plain atomic
//          OR / XOR on the word's 32-bit container -- no compare-and-swap loop -- and the old value each atomic
//          returns gives the change in the number of active clusters.
static const uint32_t PAY_VALID = 0x80000000u;

LUF_DEV void f_hook(const FShot& s, uint32_t nm, int lane, int nl)
{
    int dact = 0;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < nm; i += (uint32_t)nl) {
        const uint32_t it = s.mlist[i];
        const uint32_t a = it & 0xFFFFu, b = it >> 16;
        if (b >= END_EAST) { s.mlist[i] = a | ((b == END_WEST ? P_WEST : P_EAST) << 16) | PAY_VALID; continue; }
        uint32_t out = 0;
        LUF_NOUNROLL
        for (;;) {
            uint32_t wa, wb;
            uint32_t ra = ffind(s, a, wa), rb = ffind(s, b, wb);
            if (ra == rb) break;
            if (ra > rb) { uint32_t t = ra; ra = rb; rb = t; wb = wa; }
            if (!cas16(s.P, rb, wb, ra)) continue;
            if (wb == P_ACTIVE) --dact;
            if (wb & (P_ODD | P_WEST | P_EAST)) out = ra | ((wb & (P_ODD | P_WEST | P_EAST)) << 16) | PAY_VALID;
            break;
        }
        s.mlist[i] = out;
    }
    if (dact) luf_atomic_add(&s.cnt[C_ACTIVE], (uint32_t)dact);
}

LUF_DEV void f_apply(const FView& g, const FShot& s, uint32_t nm, int lane, int nl)
{
    int dact = 0;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < nm; i += (uint32_t)nl) {
        const uint32_t it = s.mlist[i];
        if (!(it & PAY_VALID)) continue;
        const uint32_t bits = (it >> 16) & (P_ODD | P_WEST | P_EAST);
        uint32_t w;
        const uint32_t r = ffind_ro(s, it & 0xFFFFu, w);
        uint32_t* wp = (uint32_t*)s.P + (r >> 1);
        const uint32_t sh = (r & 1u) * 16u;
        if (bits & (P_WEST | P_EAST)) {
            const uint32_t old = (luf_atomic_or(wp, (bits & (P_WEST | P_EAST)) << sh) >> sh) & 0xFFFFu;
            const uint32_t nw = old | (bits & (P_WEST | P_EAST));
            dact += (nw == P_ACTIVE ? 1 : 0) - (old == P_ACTIVE ? 1 : 0);
        }
        if (bits & P_ODD) {
            const uint32_t old = (luf_atomic_xor(wp, P_ODD << sh) >> sh) & 0xFFFFu;
            dact += ((old ^ P_ODD) == P_ACTIVE ? 1 : 0) - (old == P_ACTIVE ? 1 : 0);
        }
    }
    if (dact) luf_atomic_add(&s.cnt[C_ACTIVE], (uint32_t)dact);
}

// A round with more newly FULL edges than the list holds joins every fully grown edge of the growth words in one
// phase instead (compare-and-swap payloads; an edge joined before is inside one cluster by now: nothing
// happens), endpoints from the edge table.
LUF_DEV void f_merge_pair(const FView& g, const FShot& s, uint32_t a, uint32_t b)
{
    if (b >= END_EAST) add_bits(s, a, b == END_WEST ? P_WEST : P_EAST, g.west);
    else join(s, a, b, g.west);
}

LUF_DEV void f_merge_overflow(const FView& g, const FShot& s, int lane, int nl)
{
    const int GW = (g.E + 15) >> 4;
    LUF_NOUNROLL
    for (int w = lane; w < GW; w += nl) {
        const uint32_t x = s.growth[w];
        uint32_t full = (x >> 1) & 0x55555555u;                   // bit 2k: edge 16 w + k is FULL
        LUF_NOUNROLL
        while (full) {
            const int b = luf_ffs(full) - 1; full &= full - 1;
            const uint32_t ei = LUF_LDG(&g.einfo[16 * w + (b >> 1)]);
            const uint32_t u = ei & 0xFFFFu, v = ei >> 16;
            if (u >= END_EAST) f_merge_pair(g, s, v, u); else f_merge_pair(g, s, u, v);
        }
    }
}

// every defect adds 1 iff its cluster touches the west boundary (uf_peel_parity, luf_uf_phases.inl)
LUF_DEV uint32_t f_defect_sides(const FView& g, const FLayout& l, const FShot& s, uint32_t nd, int lane, int nl)
{
    uint32_t parity = 0;
    f_for_nodes(g, l, s, s.defect, nd, nd, lane, nl, [&](uint32_t v) {
        uint32_t wr;
        ffind_ro(s, v, wr);
        parity ^= (wr >> 1) & 1u;
        // an ODD cluster on both sides adds its parity iff the boundary node it kept is a west one: under the
        // default inclination that depends on the order of the merges (an even one adds 0 either way, and
        // under the west inclination the node kept is a west one: the line above is right)
        if (!g.west && (wr & (P_ODD | P_WEST | P_EAST)) == (P_ODD | P_WEST | P_EAST)) s.cnt[C_HARD_M] = 1u;
    });
    return parity;
}

// ---------------------------------------------------------------- HARD clusters without the canonical kernel
// Which boundary node an ODD cluster on both sides keeps under the default inclination (uf.py:536-538: the larger
// cluster of a union keeps its own) depends on the sizes of the clusters at every union, i.e. on the canonical
// order of the merges INSIDE that cluster -- and on nothing else: growth, the partition after every round and
// every other cluster's contribution to the logical bit are order-free.  So a shot with such clusters is not
// decoded again: the merge log of the shot (every newly FULL edge, keyed by round and edge id) is filtered to the
// edges inside those clusters, sorted and exported; hard_resolve replays those unions alone in canonical
// order (union by size, ties keep the cluster of the edge's first endpoint; every boundary edge brings a
// boundary node of its own: size 1, its side) and reports how many of the clusters keep an EAST node -- for each
// of them the fast bit, which counted the cluster as west, is flipped.
LUF_DEV bool f_in_hard_cluster(const FView& g, const FShot& s, uint32_t key)
{
    const uint32_t ei = LUF_LDG(&g.einfo[key & KEY_EDGE_MASK]);
    const uint32_t a = ei & 0xFFFFu, b = ei >> 16;
    uint32_t wr;
    ffind_ro(s, a < END_EAST ? a : b, wr);
    return (wr & (P_ODD | P_WEST | P_EAST)) == (P_ODD | P_WEST | P_EAST);
}

// The shot is decoded; its log becomes the export.  Entries outside the HARD clusters are blanked (all ones: they
// sort to the end), the rest is counted ...
LUF_DEV void f_export_mark(const FView& g, const FShot& s, uint32_t ntot, uint32_t np, int lane, int nl)
{
    uint32_t c = 0;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < np; i += (uint32_t)nl) {
        if (i < ntot && f_in_hard_cluster(g, s, s.elog[i])) ++c; else s.elog[i] = 0xFFFFFFFFu;
    }
    if (c) luf_atomic_add(&s.cnt[C_XCOUNT], c);
}

// The same when the log fills at most half of its array: the entries of the HARD clusters are gathered into the
// upper half (C_XCOUNT hands out the places), so that the network sorts them alone and not the whole log.
LUF_DEV void f_export_gather(const FView& g, const FLayout& l, const FShot& s, uint32_t ntot, int lane, int nl)
{
    uint32_t* dst = s.elog + (l.lcap >> 1);
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < ntot; i += (uint32_t)nl) {
        const uint32_t key = s.elog[i];
        if (f_in_hard_cluster(g, s, key)) dst[luf_atomic_add(&s.cnt[C_XCOUNT], 1u)] = key;
    }
}

// One-pass mode (the log in global memory): count the entries of the HARD clusters, then write the keys as they lie.
LUF_DEV void f_export_count(const FView& g, const FShot& s, uint32_t ntot, int lane, int nl)
{
    uint32_t c = 0;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < ntot; i += (uint32_t)nl) c += f_in_hard_cluster(g, s, s.elog[i]) ? 1u : 0u;
    if (c) luf_atomic_add(&s.cnt[C_XCOUNT], c);
}

LUF_DEV void f_export_keys(const FView& g, const FShot& s, const FExport& x, uint32_t ntot, int lane, int nl)
{
    const uint32_t base = s.cnt[C_XBASE];
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < ntot; i += (uint32_t)nl) {
        const uint32_t key = s.elog[i];
        if (f_in_hard_cluster(g, s, key)) x.buf[base + 1u + luf_atomic_add(&s.cnt[C_XPOS], 1u)] = key;
    }
}

// ... sorted by (round, edge id) in place, a bitonic network over the tile's lanes ...
LUF_DEV void r_sort_step(uint32_t* key, uint32_t np, uint32_t k, uint32_t j, int lane, int nl)
{
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < np; i += (uint32_t)nl) {
        const uint32_t o = i ^ j;
        if (o <= i) continue;
        const uint32_t a = key[i], b = key[o];
        if ((a > b) == ((i & k) == 0u)) { key[i] = b; key[o] = a; }
    }
}

// ... one lane reserves room for `count | bit << 31` and the edges; C_XBASE = where, or 0xFFFFFFFF (no room) ...
LUF_DEV void f_export_reserve(const FShot& s, const FExport& x, uint32_t fast_bit)
{
    const uint32_t n = s.cnt[C_XCOUNT];
    const uint64_t at = luf_atomic_add(x.cursor, (uint64_t)n + 1u);
    uint32_t base = 0xFFFFFFFFu;
    if (at + n + 1u <= (uint64_t)x.cap) { base = (uint32_t)at; x.buf[base] = n | (fast_bit << 31); }
    s.cnt[C_XBASE] = base;
}

// ... and the edges go out in canonical order as endpoint pairs (first end | second end << 16: hard_resolve needs
// no table).
LUF_DEV void f_export_write(const FView& g, const FShot& s, const uint32_t* sorted, const FExport& x, int lane, int nl)
{
    const uint32_t base = s.cnt[C_XBASE], n = s.cnt[C_XCOUNT];
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) x.buf[base + 1u + i] = LUF_LDG(&g.einfo[sorted[i] & KEY_EDGE_MASK]);
}

// hard_resolve's node words (one per detector, in the warp's shared memory): root = R_ROOT | kept side << 28 |
// counted << 27 | size, non-root = parent
static const uint32_t R_ROOT = 0x80000000u, R_SIDE_SHIFT = 28u, R_SIDE_MASK = 3u << 28, R_COUNTED = 1u << 27, R_SIZE_MASK = (1u << 27) - 1u;
static const uint32_t R_WEST = 1u, R_EAST = 2u;

// One dependent load per hop (union by size keeps the trees shallow); a start node two or more hops from its
// root is pointed at it.  `wr` = the root's word.
LUF_DEV uint32_t r_find(uint32_t* S, uint32_t v, uint32_t& wr)
{
    uint32_t w = S[v];
    if (w & R_ROOT) { wr = w; return v; }
    const uint32_t start = v;
    int hops = 0;
    LUF_NOUNROLL
    do { v = w; w = S[v]; ++hops; } while (!(w & R_ROOT));
    if (hops > 1) S[start] = v;
    wr = w;
    return v;
}

// every detector the edges name starts as a singleton without a boundary
LUF_DEV void r_init(const uint32_t* ends, uint32_t n, uint32_t* S, int lane, int nl)
{
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const uint32_t ei = ends[i];
        const uint32_t a = ei & 0xFFFFu, b = ei >> 16;
        if (a < END_EAST) S[a] = R_ROOT | 1u;
        if (b < END_EAST) S[b] = R_ROOT | 1u;
    }
}

// uf.py:256-301 for one edge (endpoint pair) on the scratch forest
LUF_DEV void r_union(uint32_t* S, uint32_t ei)
{
    const uint32_t a = ei & 0xFFFFu, b = ei >> 16;
    if (a >= END_EAST || b >= END_EAST) {            // a boundary node of its own (one edge each): size 1, boundary = itself
        const uint32_t side = (a >= END_EAST ? a : b) == END_WEST ? R_WEST : R_EAST;
        uint32_t w;
        const uint32_t r = r_find(S, a >= END_EAST ? b : a, w);
        w += 1u;
        if (!(w & R_SIDE_MASK)) w |= side << R_SIDE_SHIFT;     // uf.py:536-538 (a detector cluster that is no larger has no boundary yet)
        S[r] = w;
        return;
    }
    uint32_t wa, wb;
    const uint32_t ra = r_find(S, a, wa), rb = r_find(S, b, wb);
    if (ra == rb) return;
    const bool a_large = (wa & R_SIZE_MASK) >= (wb & R_SIZE_MASK);             // uf.py:270-273: ties keep the first endpoint's cluster
    const uint32_t L = a_large ? ra : rb, Sm = a_large ? rb : ra, wl = a_large ? wa : wb, ws = a_large ? wb : wa;
    const uint32_t side = (wl & R_SIDE_MASK) ? (wl & R_SIDE_MASK) : (ws & R_SIDE_MASK);
    S[L] = R_ROOT | side | ((wl & R_SIZE_MASK) + (ws & R_SIZE_MASK));
    S[Sm] = L;
}

// All lanes: every cluster once (the roots are stable now); *out ^= 1 per cluster that keeps an east node.
LUF_DEV void r_count_east(const uint32_t* ends, uint32_t n, uint32_t* S, uint32_t* out, int lane, int nl)
{
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const uint32_t ei = ends[i];
        uint32_t v = (ei & 0xFFFFu) < END_EAST ? (ei & 0xFFFFu) : (ei >> 16), w;
        LUF_NOUNROLL
        while (!((w = S[v]) & R_ROOT)) v = w;
        if (luf_atomic_or(&S[v], R_COUNTED) & R_COUNTED) continue;
        if (((w & R_SIDE_MASK) >> R_SIDE_SHIFT) == R_EAST) luf_atomic_xor(out, 1u);
    }
}

}  // namespace fast
}  // namespace luf

// ------------------------------------------------------------------ the shot driver
// FT_PHASE(stmt): every lane of the tile runs stmt, then the tile synchronises (device: FT_SYNC() must be
// defined by the including translation unit before fast_shot is instantiated; host emulation: the lanes run
// one after the other).
#if defined(LUF_HOST_EMU)
  #define FT_PHASE(stmt) for (int lane = 0; lane < nl; ++lane) { stmt; }
#else
  #define FT_PHASE(stmt) { stmt; } sync();
#endif

namespace luf {
namespace fast {

// One shot: reset, sample (32 virtual sampler lanes whatever the tile, so that a shot's noise is the one
// luf_sample writes for the same seed), load, rounds of grow / merge, parity.  Returns 0 / 1 = the logical bit
// (_schemes.py:143-155), or HARD.  `sync` is the tile barrier.  stats (emulation only): [0] rounds,
// [1] touched nodes, [2] most newly FULL edges of a round, [3] most active nodes of a round.
template <bool ADJ_SHARED, class Sync>
LUF_DEV uint32_t fast_shot(const FView& g, const FLayout& l, const FShot& s, const SampleParams& sp, uint64_t shot,
                           int lane, int nl, Sync sync, uint32_t* stats = nullptr, const FExport* x = nullptr)
{
    (void)lane; (void)sync;
    FT_PHASE((void)0);                                  // every lane is done with the previous shot's counters
    FT_PHASE(f_reset(l, s, lane, nl));
    FT_PHASE(
        SampleTab t; t.F = g.F; t.n_classes = g.n_classes; t.perm = g.fresh_perm; t.class_end = g.class_end; t.perm32 = g.fresh_perm32;
        uint32_t par = 0;
        for (int vl = lane; vl < 32; vl += nl)
            par ^= ph_sample_core(t, sp, shot, 0u, (const uint32_t*)0, vl, 32, [&](uint32_t e) { return f_flip(g, s, e); });
        if (par & 1u) luf_atomic_xor(&s.cnt[C_PARITY], 1u));
    FT_PHASE(f_load(g, l, s, lane, nl));
    const uint32_t nd = s.cnt[C_TOUCHED];
    uint32_t rounds = 0, abase = 0, mbase = 0, most_m = 0, most_a = 0;
    if (nd) {
        FT_PHASE(if (lane == 0) s.cnt[C_ACTIVE] = nd; f_grow_first<ADJ_SHARED>(g, l, s, nd, lane, nl));
        LUF_NOUNROLL
        for (;;) {
            ++rounds;
            const uint32_t nm = s.cnt[C_MLIST] - mbase;
            if (nm > most_m) most_m = nm;
            if (nm <= l.mcap) {
                FT_PHASE(f_hook(s, nm, lane, nl));
                FT_PHASE(f_apply(g, s, nm, lane, nl));
            } else {
                FT_PHASE(f_merge_overflow(g, s, lane, nl));
            }
            mbase += nm;
            if (!s.cnt[C_ACTIVE]) break;
            const uint32_t nt = s.cnt[C_TOUCHED];
            const uint32_t rkey = (rounds + 1u) << KEY_EDGE_BITS;
            FT_PHASE(f_grow_filter<ADJ_SHARED>(g, l, s, nt, abase, mbase, rkey, lane, nl));
            const uint32_t na = s.cnt[C_ALIST] - abase;
            if (na > most_a) most_a = na;
            FT_PHASE(f_grow_exec<ADJ_SHARED>(g, l, s, na, mbase, rkey, lane, nl));
            abase += na;
        }
        FT_PHASE(const uint32_t par = f_defect_sides(g, l, s, nd, lane, nl); if (par) luf_atomic_xor(&s.cnt[C_PARITY], 1u));
        if (s.cnt[C_HARD_M]) {                           // (uniform: every lane reads the same word after the barrier)
            const uint32_t ntot = s.cnt[C_MLIST];
            if (!x || !x->buf || !l.lcap || ntot > l.lcap) return HARD;           // no log, or it could not hold the shot's merges
            if (x->unsorted) {                           // one-pass mode: the keys as they lie; k_hard_sort orders them
                FT_PHASE(f_export_count(g, s, ntot, lane, nl); if (lane == 0) s.cnt[C_XPOS] = 0u);
                if (s.cnt[C_XCOUNT] > x->max_keys) return HARD;
                FT_PHASE(if (lane == 0) f_export_reserve(s, *x, s.cnt[C_PARITY] & 1u));
                if (s.cnt[C_XBASE] == 0xFFFFFFFFu) return HARD;
                FT_PHASE(f_export_keys(g, s, *x, ntot, lane, nl));
                return HARD_LOGGED;
            }
            uint32_t np = 1;
            uint32_t* sorted = s.elog;
            if (2u * ntot <= l.lcap) {                   // room to gather the HARD clusters' entries apart: a smaller sort
                sorted = s.elog + (l.lcap >> 1);
                FT_PHASE(f_export_gather(g, l, s, ntot, lane, nl));
                const uint32_t nh = s.cnt[C_XCOUNT];
                while (np < nh) np <<= 1;
                FT_PHASE(for (uint32_t i = nh + (uint32_t)lane; i < np; i += (uint32_t)nl) sorted[i] = 0xFFFFFFFFu);
            } else {
                while (np < ntot) np <<= 1;              // (lcap is a power of two)
                FT_PHASE(f_export_mark(g, s, ntot, np, lane, nl));
            }
            LUF_NOUNROLL
            for (uint32_t k = 2; k <= np; k <<= 1) {
                LUF_NOUNROLL
                for (uint32_t j = k >> 1; j > 0; j >>= 1) { FT_PHASE(r_sort_step(sorted, np, k, j, lane, nl)); }
            }
            FT_PHASE(if (lane == 0) f_export_reserve(s, *x, s.cnt[C_PARITY] & 1u));
            if (s.cnt[C_XBASE] == 0xFFFFFFFFu) return HARD;
            FT_PHASE(f_export_write(g, s, sorted, *x, lane, nl));
            return HARD_LOGGED;                          // the caller queues (shot, s.cnt[C_XBASE])
        }
    }
    if (stats) { stats[0] = rounds; stats[1] = s.cnt[C_TOUCHED]; stats[2] = most_m; stats[3] = most_a; }
    return s.cnt[C_PARITY] & 1u;
}

// One exported shot of the one-pass mode, one tile: the keys of rec[1 .. n] sorted through `key` (np = the power
// of two >= n words of shared memory) and written back as endpoint pairs, which is what hard_resolve reads.
template <class Sync>
LUF_DEV void hard_sort(const FView& g, uint32_t* rec, uint32_t* key, int lane, int nl, Sync sync)
{
    (void)lane; (void)sync;
    const uint32_t n = rec[0] & 0x7FFFFFFFu;
    uint32_t np = 1;
    while (np < n) np <<= 1;
    FT_PHASE(for (uint32_t i = (uint32_t)lane; i < np; i += (uint32_t)nl) key[i] = i < n ? rec[1u + i] : 0xFFFFFFFFu);
    LUF_NOUNROLL
    for (uint32_t k = 2; k <= np; k <<= 1) {
        LUF_NOUNROLL
        for (uint32_t j = k >> 1; j > 0; j >>= 1) { FT_PHASE(r_sort_step(key, np, k, j, lane, nl)); }
    }
    FT_PHASE(for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) rec[1u + i] = LUF_LDG(&g.einfo[key[i] & KEY_EDGE_MASK]));
}

// One exported shot, one warp: rec[0] = count | fast bit << 31, then the edges of the shot's HARD clusters in
// canonical order.  `S` = one word per detector, `out` = one word (both the warp's own).  The unions are one
// dependent chain; every lane of the warp runs it on the same values (same cost as one lane, and the warp can
// fetch the edges 32 at a time).  Returns the shot's logical bit.
template <class Sync>
LUF_DEV uint32_t hard_resolve(const uint32_t* rec, uint32_t* S, uint32_t* out, int lane, int nl, Sync sync)
{
    (void)lane; (void)sync;
    const uint32_t n = rec[0] & 0x7FFFFFFFu, bit = rec[0] >> 31;
    const uint32_t* ends = rec + 1;
    FT_PHASE(r_init(ends, n, S, lane, nl); if (lane == 0) *out = bit);
#if defined(LUF_HOST_EMU)
    for (uint32_t i = 0; i < n; ++i) r_union(S, ends[i]);
#else
    LUF_NOUNROLL
    for (uint32_t c = 0; c < n; c += 32u) {
        const uint32_t mine = c + (uint32_t)lane < n ? __ldg(&ends[c + (uint32_t)lane]) : 0u;
        const int m = (int)(n - c < 32u ? n - c : 32u);
        LUF_NOUNROLL
        for (int j = 0; j < m; ++j) { r_union(S, __shfl_sync(0xffffffffu, mine, j)); __syncwarp(); }
    }
    sync();
#endif
    FT_PHASE(r_count_east(ends, n, S, out, lane, nl));
    return *out;
}

}  // namespace fast
}  // namespace luf
