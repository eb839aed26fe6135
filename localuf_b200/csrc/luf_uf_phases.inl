// luf_uf_phases.inl -- the per-shot phases of the Union-Find decoder
// (validate + peel), written once against a *tile* of nl cooperating lanes and
// compiled once per tile shape by luf_core.cuh:
//   namespace luf::uf_warp   tile = one warp, small work lists (graphs whose shot state leaves many warps per SM)
//   namespace luf::uf_cta    tile = one thread block, large work lists (big graphs: one shot per block)
// The including file defines LUF_UF_NS, the list sizes LUF_T_LIST_CAP /
// LUF_T_LIST2_CAP / LUF_T_CLAIM_SLOTS and the tile operations LUF_T_PHASE /
// LUF_T_ANY / LUF_T_SUM / LUF_T_EXSCAN / LUF_T_SCAN_PREP.
//
// Algorithm (reference: localuf/decoders/uf.py, canonical order of
// oracle/ref_canonical.py):
//   K2a validate            uf.py:200-301  (grow a half-edge per round, merge in ascending edge id)
//   K2b peel                uf.py:305-344  (LIFO spanning tree in ascending edge id, leaf-up peeling)
//
// Word width.  The two 16-bit instances pack a node word into 32 bits (indices below 2^15 / 2^16: luf_core.cuh);
// the wide instance (luf_uf_wide.cuh, LUF_T_WIDE: graphs beyond that, state in global or shared memory) packs the
// same fields into 64 bits.  word_t = a node word / list entry / reservation; HSH = where its upper half starts;
// LO_MASK = its lower half; IDX_NONE / TREE_ROOT = the two reserved index values of the peel; ent_t = an
// adjacency entry.  The wide instance defines them, and its own GraphView / Shot / Layout, before including this.
namespace luf {
namespace LUF_UF_NS {

#if !defined(LUF_T_WIDE)
typedef uint32_t word_t;
typedef uint32_t ent_t;
static const int HSH = 16;
static const word_t LO_MASK = 0xFFFFu, IDX_NONE = NONE16, TREE_ROOT_W = TREE_ROOT;
LUF_DEV uint32_t ent_edge(ent_t ent) { return ent & 0xFFFFu; }
LUF_DEV void edge_ends(const GraphView& g, uint32_t e, uint32_t& u, uint32_t& v)
{
    const uint32_t uv = LUF_LDG(&g.edge_uv[e]);
    u = uv & 0xFFFFu; v = uv >> 16;
}
#endif
static const word_t WORD_ONES = ~(word_t)0;
LUF_DEV uint32_t* claim_bits(const Shot& s) { return (uint32_t*)s.claim; }      // the reservation table doubles as a node bitmap

// find without writes: safe while other lanes read or extend the forest.  `wr` = the root's word.
LUF_DEV uint32_t find_ro(const Shot& s, uint32_t v, word_t& wr)
{
    word_t w;
    LUF_NOUNROLL
    while (!((w = s.S[v]) & ROOT_BIT)) v = (uint32_t)(w & SIZE_MASK);
    wr = w;
    return v;
}

// find that also shortcuts the start node to the root it found.  Safe next to
// concurrent finds and unions of other lanes: a non-root node never becomes a
// root again and the value written is one of its ancestors (32-bit stores are
// atomic; the union edge in the upper half is kept), so every interleaving
// leaves a valid forest with the same roots.
LUF_DEV uint32_t find_pc(const Shot& s, uint32_t v, word_t& wr)
{
    const word_t w = s.S[v];
    if (w & ROOT_BIT) { wr = w; return v; }
    uint32_t r = (uint32_t)(w & SIZE_MASK);
    word_t q;
    int hops = 0;
    LUF_NOUNROLL
    while (!((q = s.S[r]) & ROOT_BIT)) { r = (uint32_t)(q & SIZE_MASK); ++hops; }
    if (hops) s.S[v] = (w & ~LO_MASK) | r;
    wr = q;
    return r;
}

// uf.py:247-254 -- find with full path compression (single lane only)
LUF_DEV uint32_t find_compress(const Shot& s, uint32_t v)
{
    uint32_t r = v;
    word_t w;
    LUF_NOUNROLL
    while (!((w = s.S[r]) & ROOT_BIT)) r = (uint32_t)(w & SIZE_MASK);
    LUF_NOUNROLL
    while (v != r) { w = s.S[v]; s.S[v] = (w & ~LO_MASK) | r; v = (uint32_t)(w & SIZE_MASK); }
    return r;
}

// Compact the set bits of a bitmap into the work list: every lane owns a
// contiguous range of words, reserves room for its bits with one atomic add
// and writes `id0 + bit index` for each.  When the total exceeds the list the
// consumers walk the bitmap itself (same per-item code, unbalanced lanes).
// `word(w)` yields word w of the bitmap.
template <class W>
LUF_DEV void gather_bits(W word, int nwords, uint32_t id0, word_t* list, uint32_t* counter, int lane, int nl)
{
    const int wpl = (nwords + nl - 1) / nl;
    const int w0 = lane * wpl, w1 = w0 + wpl < nwords ? w0 + wpl : nwords;
    uint32_t c = 0;
    LUF_NOUNROLL
    for (int w = w0; w < w1; ++w) c += (uint32_t)luf_popc(word(w));
    if (!c) return;
    uint32_t pos = luf_atomic_add(counter, c);
    if (pos + c > LUF_T_LIST_CAP) return;
    LUF_NOUNROLL
    for (int w = w0; w < w1; ++w) {
        uint32_t bits = word(w);
        LUF_NOUNROLL
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            list[pos++] = id0 + (uint32_t)(32 * w + b);
        }
    }
}

// Run f(id) for every set bit, straight from the bitmap (short bodies: cheaper than compacting first).
template <class W, class F>
LUF_DEV void for_bits(W word, int nwords, uint32_t id0, int lane, int nl, F f)
{
    LUF_NOUNROLL
    for (int w = lane; w < nwords; w += nl) {
        uint32_t bits = word(w);
        LUF_NOUNROLL
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            f(id0 + (uint32_t)(32 * w + b));
        }
    }
}

// Run f(id) for every item: from the list if it held them all, else from the bitmap.
template <class W, class F>
LUF_DEV void for_items(W word, int nwords, uint32_t id0, const word_t* list, uint32_t total,
                       int lane, int nl, F f)
{
    if (total <= LUF_T_LIST_CAP) {
        LUF_NOUNROLL
        for (uint32_t i = (uint32_t)lane; i < total; i += (uint32_t)nl) f((uint32_t)list[i]);
    } else for_bits(word, nwords, id0, lane, nl, f);
}

// Row v of the adjacency table, read as 64-bit pairs; f(entry) for every real entry
// (entry = edge | other end << 16 | this is the edge's first endpoint << 31).
#if !defined(LUF_T_WIDE)
template <class F>
LUF_DEV void for_row(const GraphView& g, uint32_t v, F f)
{
    const uint32_t* row = g.adj + (size_t)v * (uint32_t)g.DEG;
    LUF_NOUNROLL
    for (int k = 0; k < g.DEG; k += 2) {
        const luf_u32x2 p = LUF_LDG2(row + k);
        if (p.x == ADJ_PAD) break;
        f(p.x);
        if (p.y != ADJ_PAD) f(p.y);
    }
}

LUF_DEV uint32_t other_end(uint32_t ent, bool& first)
{
    first = (ent >> 31) != 0;
    return (ent >> 16) & 0x7FFFu;
}
#endif

// ------------------------------------------------------------ phase 0: reset
// uf.py:79-96,496-519 -- every node a singleton cluster; a boundary node is its
// own cluster boundary.  (The reference re-allocates N Python objects here,
// 40-58 % of its time; this is N*4 + ~3 KB of 16-byte shared-memory stores.)
LUF_DEV word_t fresh_word(const GraphView& g, int v) { return ROOT_BIT | 1u | (v < g.B ? ((word_t)v + 1u) << HSH : (word_t)0); }

LUF_DEV void ph_reset(const GraphView& g, const Layout& l, const Shot& s, int lane, int nl)
{
#if defined(LUF_T_WIDE)
    LUF_NOUNROLL
    for (int v = lane; v < g.N; v += nl) s.S[v] = fresh_word(g, v);
#else
    luf_u32x4* S4 = (luf_u32x4*)s.S;
    LUF_NOUNROLL
    for (int v = 4 * lane; v < g.N; v += 4 * nl) {
        luf_u32x4 w;
        w.x = fresh_word(g, v); w.y = fresh_word(g, v + 1); w.z = fresh_word(g, v + 2); w.w = fresh_word(g, v + 3);
        S4[v >> 2] = w;
    }
#endif
    luf_u32x4* Z4 = (luf_u32x4*)((unsigned char*)s.S + (l.zero_off - l.S));
    luf_u32x4 z; z.x = z.y = z.z = z.w = 0u;
    LUF_NOUNROLL
    for (int i = lane; i < (int)(l.zero_bytes >> 4); i += nl) Z4[i] = z;
    LUF_NOUNROLL
    for (int w = lane; w < LUF_T_CLAIM_SLOTS; w += nl) s.claim[w] = WORD_ONES;
    if (lane < 8) s.cnt[lane] = 0;
}

// uf.py:98-104 load: every defect is an odd singleton cluster (hence active).
// Also compacts the defects into the work list for the first growth round
// (cnt[0] = their number; must be 0 on entry, ph_reset leaves it so).
LUF_DEV void ph_load(const GraphView& g, const Shot& s, int lane, int nl)
{
    const int wpl = (g.WD + nl - 1) / nl;
    const int w0 = lane * wpl, w1 = w0 + wpl < g.WD ? w0 + wpl : g.WD;
    uint32_t c = 0;
    LUF_NOUNROLL
    for (int w = w0; w < w1; ++w) c += (uint32_t)luf_popc(s.defect[w]);
    if (!c) return;
    uint32_t pos = luf_atomic_add(&s.cnt[0], c);
    const bool fits = pos + c <= LUF_T_LIST_CAP;
    LUF_NOUNROLL
    for (int w = w0; w < w1; ++w) {
        uint32_t bits = s.defect[w];
        LUF_NOUNROLL
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t v = (uint32_t)(g.B + 32 * w + b);
            s.S[v] = ODD_BIT | ROOT_BIT | 1u;
            luf_atomic_or(&s.full1[v >> 5], 1u << (v & 31u));
            if (fits) s.list[pos++] = v;
        }
    }
}

// ------------------------------------------------------- phase 2a: grow a round
// uf.py:221-222,234-245.  vision(C) = edges incident to C's members that are not
// yet FULL/BURNT; a cluster is active iff it is odd and has no boundary
// (uf.py:296-301), i.e. the upper half of its root word is exactly the odd
// bit.  Members are not kept in lists: every *touched* node (a defect, or an
// end of a fully grown edge -- other singletons are never active) looks its
// root up and grows its own incident edges if the root is active.  An edge
// inside one cluster is in that cluster's vision once: only its first endpoint
// adds.  Two different active clusters each add 1.  The 2-bit counters
// saturate at FULL.
//
// Node flags, set with atomicOr only.  full1 = touched: a defect, or an end of
// a fully grown edge.  full2 = a second fully grown edge arrived at a touched
// node, i.e. two or more such edges (one or more for defects); it only decides
// whether the peel traversal may skip a leaf.
LUF_DEV void bump(const Shot& s, uint32_t v)
{
    const uint32_t bit = 1u << (v & 31u);
    if (luf_atomic_or(&s.full1[v >> 5], bit) & bit) luf_atomic_or(&s.full2[v >> 5], bit);
}

LUF_DEV void became_full(const Shot& s, uint32_t e, uint32_t u, uint32_t o)
{
    luf_atomic_or(&s.newfull[e >> 5], 1u << (e & 31u));
    bump(s, u); bump(s, o);
}

// First round: every cluster is a singleton and every edge ungrown, so each
// defect adds 1 to all its edges with no look-ups; an edge is newly FULL iff
// the other end got there first.
LUF_DEV void grow_first(const GraphView& g, const Shot& s, uint32_t v)
{
    for_row(g, v, [&](ent_t ent) {
        const uint32_t e = ent_edge(ent), sh = (e & 15u) * 2u;
        const uint32_t old = (luf_atomic_add(&s.growth[e >> 4], 1u << sh) >> sh) & 3u;
        if (old == G_HALF) { bool first; became_full(s, e, v, other_end(ent, first)); }
    });
}

// Returns the number of defects (0 = nothing to decode).
LUF_DEV uint32_t ph_grow_first(const GraphView& g, const Shot& s, int lane, int nl)
{
    const uint32_t total = s.cnt[0];
    for_items([&](int w) { return s.defect[w]; }, g.WD, (uint32_t)g.B, s.list, total, lane, nl,
              [&](uint32_t v) { grow_first(g, s, v); });
    return total;
}

// `node`: NodeUF (node_uf.py:45-67) -- every frontier node adds 1 to each of its active edges, also
// inside its own cluster.
LUF_DEV void grow_node(const GraphView& g, const Shot& s, uint32_t u, uint32_t r, bool node)
{
    for_row(g, u, [&](ent_t ent) {
        const uint32_t e = ent_edge(ent), sh = (e & 15u) * 2u;
        if (((s.growth[e >> 4] >> sh) & 3u) >= G_FULL) return;
        bool first;
        const uint32_t o = other_end(ent, first);
        if (!node && !first) {                        // second endpoint: skip if same cluster
            word_t wr;
            if (o == r || find_ro(s, o, wr) == r) return;
        }
        const uint32_t old = (luf_atomic_add(&s.growth[e >> 4], 1u << sh) >> sh) & 3u;
        if (old >= G_FULL) luf_atomic_sub(&s.growth[e >> 4], 1u << sh);     // saturate
        else if (old == G_HALF) became_full(s, e, u, o);
    });
}

// Later rounds: compact the touched nodes (cnt[0]; cnt[0] = cnt[1] = 0 on entry) ...
LUF_DEV void ph_grow_gather(const GraphView& g, const Shot& s, int lane, int nl)
{
    gather_bits([&](int w) { return s.full1[w]; }, g.WN, 0u, s.list, &s.cnt[0], lane, nl);
    if (lane == 0) { s.cnt[2] = 0; s.cnt[3] = 0xFFFFFFFFu; }
}

// ... keep those whose cluster is active (cnt[1], list2), one per lane ...
// Bucket mode (NodeBUF, node_buf.py:35-61: always the smallest active clusters first): cnt[3]
// collects the smallest active cluster size (0xFFFFFFFF on entry); nothing grows in this phase.
template <int TF>
LUF_DEV void ph_grow_filter(const GraphView& g, const Shot& s, int flags, int lane, int nl)
{
    const int f = TF >= 0 ? TF : flags;
    const bool node = (f & FLAG_NODE) != 0, bucket = (f & FLAG_BUCKET) != 0;
    for_items([&](int w) { return s.full1[w]; }, g.WN, 0u, s.list, s.cnt[0], lane, nl, [&](uint32_t u) {
        word_t wr;
        const uint32_t r = find_pc(s, u, wr);
        if ((wr >> HSH) != (ODD_BIT >> HSH)) return;
        const uint32_t idx = luf_atomic_add(&s.cnt[1], 1u);
        if (bucket) luf_atomic_min(&s.cnt[3], (uint32_t)(wr & SIZE_MASK));
        if (idx < LUF_T_LIST2_CAP) s.list2[idx] = (word_t)u | ((word_t)r << HSH);
        else if (!bucket) grow_node(g, s, u, r, node);
    });
}

// ... and grow them, the edge loops spread evenly over the lanes.  Returns the
// number of active nodes of the round.
template <int TF>
LUF_DEV uint32_t ph_grow_exec(const GraphView& g, const Shot& s, int flags, int lane, int nl)
{
    const int f = TF >= 0 ? TF : flags;
    const bool node = (f & FLAG_NODE) != 0, bucket = (f & FLAG_BUCKET) != 0;
    const uint32_t total = s.cnt[1];
    const uint32_t smallest = s.cnt[3];
    if (bucket && total > LUF_T_LIST2_CAP) {          // the list could not hold them: look the active nodes up again
        for_items([&](int w) { return s.full1[w]; }, g.WN, 0u, s.list, s.cnt[0], lane, nl, [&](uint32_t u) {
            word_t wr;
            const uint32_t r = find_ro(s, u, wr);
            if ((wr >> HSH) == (ODD_BIT >> HSH) && (uint32_t)(wr & SIZE_MASK) == smallest) grow_node(g, s, u, r, node);
        });
        return total;
    }
    const uint32_t n = total < LUF_T_LIST2_CAP ? total : LUF_T_LIST2_CAP;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const word_t it = s.list2[i];
        if (bucket && (uint32_t)(s.S[it >> HSH] & SIZE_MASK) != smallest) continue;
        grow_node(g, s, (uint32_t)(it & LO_MASK), (uint32_t)(it >> HSH), node);
    }
    return total;
}

// BUF (buf.py:64-78): in every round only the active clusters with the shortest vision grow;
// len(vision) = the distinct not-yet-FULL edges at the cluster's nodes (an edge inside the cluster
// counted once).  The rounds of this mode walk the touched nodes five times: mark the active roots
// in a bitmap (the reservation table, idle during growth), add up the vision lengths in the
// boundary field of the root words (an active cluster has no boundary; < 2^15 edges), take the
// minimum, grow the clusters that attain it, put the words back.
LUF_DEV uint32_t vision_edges(const GraphView& g, const Shot& s, uint32_t u, uint32_t r)
{
    uint32_t c = 0;
    for_row(g, u, [&](ent_t ent) {
        if (growth_of(s, ent_edge(ent)) >= G_FULL) return;
        bool first;
        const uint32_t o = other_end(ent, first);
        word_t wr;
        if (!first && (o == r || find_ro(s, o, wr) == r)) return;
        ++c;
    });
    return c;
}

template <class F>
LUF_DEV void for_active(const GraphView& g, const Shot& s, int lane, int nl, F f)
{
    const uint32_t* act = claim_bits(s);
    for_items([&](int w) { return s.full1[w]; }, g.WN, 0u, s.list, s.cnt[0], lane, nl, [&](uint32_t u) {
        word_t wr;
        const uint32_t r = find_ro(s, u, wr);
        if (bit_of(act, r)) f(u, r, wr);
    });
}

LUF_DEV void ph_buf_mark(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t* act = claim_bits(s);
    for_items([&](int w) { return s.full1[w]; }, g.WN, 0u, s.list, s.cnt[0], lane, nl, [&](uint32_t u) {
        word_t wr;
        const uint32_t r = find_pc(s, u, wr);
        if ((wr >> HSH) != (ODD_BIT >> HSH)) return;
        luf_atomic_or(&act[r >> 5], 1u << (r & 31u));
        luf_atomic_add(&s.cnt[1], 1u);
    });
}

LUF_DEV void ph_buf_count(const GraphView& g, const Shot& s, int lane, int nl)
{
    for_active(g, s, lane, nl, [&](uint32_t u, uint32_t r, word_t) {
        const uint32_t c = vision_edges(g, s, u, r);
        if (c) luf_atomic_add(&s.S[r], (word_t)c << HSH);
    });
}

LUF_DEV void ph_buf_min(const GraphView& g, const Shot& s, int lane, int nl)
{
    for_active(g, s, lane, nl, [&](uint32_t, uint32_t, word_t wr) { luf_atomic_min(&s.cnt[3], (uint32_t)((wr & BND_MASK) >> HSH)); });
}

// Returns the number of active nodes of the round.
LUF_DEV uint32_t ph_buf_exec(const GraphView& g, const Shot& s, int lane, int nl)
{
    const uint32_t shortest = s.cnt[3];
    for_active(g, s, lane, nl, [&](uint32_t u, uint32_t r, word_t wr) {
        if ((uint32_t)((wr & BND_MASK) >> HSH) == shortest) grow_node(g, s, u, r, false);
    });
    return s.cnt[1];
}

LUF_DEV void ph_buf_restore(const GraphView& g, const Shot& s, int lane, int nl)
{
    for_active(g, s, lane, nl, [&](uint32_t, uint32_t r, word_t wr) { s.S[r] = (wr & LO_MASK) | ODD_BIT; });
}

// ------------------------------------------------------------ phase 2b: merge
// uf.py:223-229,256-301 in canonical order: this round's newly FULL edges are
// merged in ascending edge id.  Union by size, ties keep the cluster of the
// edge's first endpoint; parity XOR; boundary by inclination.  TF >= 0 fixes
// the decoder flags at compile time (the fused kernel), TF < 0 reads `flags`.
// a cluster is active iff it is odd and has no boundary (uf.py:296-301)
LUF_DEV uint32_t is_active(word_t root_word) { return (root_word >> HSH) == (ODD_BIT >> HSH) ? 1u : 0u; }

template <int TF>
LUF_DEV void union_roots(const GraphView& g, const Shot& s, int flags, uint32_t e, uint32_t cu, uint32_t cv)
{
    const int f = TF >= 0 ? TF : flags;
    const word_t wu = s.S[cu], wv = s.S[cv];
    if (f & FLAG_DYNAMIC) {                                       // uf.py:261-266
        if (cu == cv || ((wu & BND_MASK) && (wv & BND_MASK))) {
            luf_atomic_or(&s.growth[e >> 4], 3u << ((e & 15u) * 2u));     // FULL(2) -> BURNT(3)
            return;
        }
    } else if (cu == cv) return;                                  // uf.py:256-259
    const uint32_t su = (uint32_t)(wu & SIZE_MASK), sv = (uint32_t)(wv & SIZE_MASK);
    const bool ul = su >= sv;                                     // uf.py:270-273
    const uint32_t L = ul ? cu : cv, Sm = ul ? cv : cu;
    const word_t wl = ul ? wu : wv, ws = ul ? wv : wu;
    word_t bl = wl & BND_MASK;
    const word_t bs = ws & BND_MASK;
    // both clusters reach a boundary: remember if the two boundary nodes lie on different sides (uf_peel_parity)
    if (bl && bs && LUF_LDG(&g.node_long[(uint32_t)(bs >> HSH) - 1]) != LUF_LDG(&g.node_long[(uint32_t)(bl >> HSH) - 1])) s.cnt[4] = 1u;
    if (f & FLAG_WEST) {                                          // uf.py:544-549
        if (bs && (!bl || LUF_LDG(&g.node_long[(uint32_t)(bs >> HSH) - 1]) < LUF_LDG(&g.node_long[(uint32_t)(bl >> HSH) - 1]))) bl = bs;
    } else if (!bl) bl = bs;                                      // uf.py:536-538
    const word_t neww = ((wl ^ ws) & ODD_BIT) | bl | ROOT_BIT | (word_t)(su + sv);
    s.S[L] = neww;
    s.S[Sm] = (word_t)L | ((word_t)e << HSH);
    const int delta = (int)is_active(neww) - (int)is_active(wl) - (int)is_active(ws);      // cnt[5]: active clusters
    if (delta) luf_atomic_add(&s.cnt[5], (uint32_t)delta);
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
// the tables need no clearing.
//
LUF_DEV uint32_t claim_slot(uint32_t root) { return (root * 0x9E5u >> 3) & (uint32_t)(LUF_T_CLAIM_SLOTS - 1); }

LUF_DEV uint32_t merge_count(const GraphView& g, const Shot& s, int lane, int nl)
{
    const int wpl = (g.WE + nl - 1) / nl;
    const int w1 = (lane + 1) * wpl < g.WE ? (lane + 1) * wpl : g.WE;
    uint32_t c = 0;
    LUF_NOUNROLL
    for (int w = lane * wpl; w < w1; ++w) c += (uint32_t)luf_popc(s.newfull[w]);
    return c;
}

// Compact this round's newly FULL edges into the list in ascending edge id and clear their bits:
// cnt[2] = how many were taken (at most the list's capacity), cnt[3] = how many stay in the bitmap for
// the next chunk.  Chunks are merged one after the other, lowest edge ids first, which is the serial order.
LUF_DEV void ph_merge_gather(const GraphView& g, const Shot& s, int lane, int nl)
{
    const int wpl = (g.WE + nl - 1) / nl;
    const int w0 = lane * wpl, w1 = w0 + wpl < g.WE ? w0 + wpl : g.WE;
    uint32_t pos, total;
    LUF_T_EXSCAN(merge_count(g, s, lane, nl), lane, nl, pos, total);
    const uint32_t take = total < LUF_T_LIST_CAP ? total : LUF_T_LIST_CAP;
    if (lane == 0) { s.cnt[0] = 0; s.cnt[1] = 0; s.cnt[2] = take; s.cnt[3] = total - take; }
    LUF_NOUNROLL
    for (int w = w0; w < w1 && pos < take; ++w) {
        uint32_t bits = s.newfull[w];
        if (!bits) continue;
        LUF_NOUNROLL
        while (bits && pos < take) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            s.list[pos++] = (uint32_t)(32 * w + b);
        }
        s.newfull[w] = bits;
    }
}

// Returns how many edges this lane placed (edges inside one cluster).
template <int TF>
LUF_DEV uint32_t ph_merge_claim(const GraphView& g, const Shot& s, int flags, uint32_t n, word_t tag, bool first, int lane, int nl)
{
    uint32_t placed = 0;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const word_t it = s.list[i];
        if (it == WORD_ONES) continue;
        const uint32_t e = (uint32_t)(it & LO_MASK);
        word_t wu, wv;
        uint32_t u, v;
        edge_ends(g, e, u, v);
        const uint32_t cu = find_pc(s, first ? u : (uint32_t)(it >> HSH), wu), cv = find_pc(s, v, wv);
        if (cu == cv) {                               // inside one cluster whatever runs before it
            union_roots<TF>(g, s, flags, e, cu, cv);
            s.list[i] = WORD_ONES;
            ++placed;
            continue;
        }
        luf_atomic_min(&s.claim[claim_slot(cu)], tag | e);
        luf_atomic_min(&s.claim[claim_slot(cv)], tag | e);
        s.list[i] = (word_t)e | ((word_t)cu << HSH);                   // the root of the first endpoint cannot change before this edge runs
    }
    return placed;
}

// Returns how many edges this lane placed.
template <int TF>
LUF_DEV uint32_t ph_merge_exec(const GraphView& g, const Shot& s, int flags, uint32_t n, word_t tag, int lane, int nl)
{
    uint32_t placed = 0;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) {
        const word_t it = s.list[i];
        if (it == WORD_ONES) continue;
        const uint32_t e = (uint32_t)(it & LO_MASK), r0 = (uint32_t)((it >> HSH) & SIZE_MASK);
        if (s.claim[claim_slot(r0)] != (tag | e)) continue;
        word_t wr;
        uint32_t eu, ev;
        edge_ends(g, e, eu, ev);
        const uint32_t cv = find_ro(s, ev, wr);
        if (s.claim[claim_slot(cv)] != (tag | e)) continue;
        union_roots<TF>(g, s, flags, e, r0, cv);
        s.list[i] = WORD_ONES;
        ++placed;
    }
    return placed;
}

// ------------------------------------------------------------- phase 3: peel
// uf.py:314-324: one tree per cluster, rooted at cluster.boundary or else the
// union-find root.  Clusters without a defect cannot contribute a correction
// edge, so only clusters of defects are rooted.
//
// Two-node clusters (the common case at low noise: one flipped edge) need no
// traversal: their only tree is the union edge, kept in the word of the member
// that is not the root, and it is in the correction iff the cluster holds a
// defect.  Exactly one member acts: the non-root member if it is a defect;
// else the root, whose partner is then the cluster's boundary node (an odd
// cluster that stopped growing has one).
//
// Entry state: cnt[0..3] = 0 and the visited bitmap (aliasing the claim table) cleared.
template <class F>
LUF_DEV void for_defects(const GraphView& g, const Shot& s, int lane, int nl, F f)
{
    for_items([&](int w) { return s.defect[w]; }, g.WD, (uint32_t)g.B, s.list, s.cnt[0], lane, nl, f);
}

LUF_DEV void ph_peel_gather(const GraphView& g, const Shot& s, int lane, int nl)
{
    gather_bits([&](int w) { return s.defect[w]; }, g.WD, (uint32_t)g.B, s.list, &s.cnt[0], lane, nl);
}

// Returns this lane's parity of pair corrections on the west boundary and
// (bit 1) whether any cluster needs the general traversal.  Tree roots go to
// list2 (cnt[1]); their bit in the troot bitmap (= newfull, idle by now)
// de-duplicates them and serves the overflow path.
LUF_DEV uint32_t ph_tree_roots(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t out = 0;
    uint32_t* troot = s.newfull;
    for_defects(g, s, lane, nl, [&](uint32_t dn) {
        word_t rr;
        const uint32_t r = find_ro(s, dn, rr);
        if ((rr & SIZE_MASK) == 2u) {
            uint32_t x = dn;                                  // the member whose word holds the union edge
            if (dn == r) {
                if (!(rr & ODD_BIT)) return;                  // both are defects: the other member acts
                x = (uint32_t)((rr & BND_MASK) >> HSH) - 1u;  // single defect: its partner is the boundary node
            }
            const uint32_t e = (uint32_t)(s.S[x] >> HSH);
            if (s.corr) luf_atomic_xor(&s.corr[e >> 5], 1u << (e & 31u));
            out ^= (LUF_LDG(&g.west_mask[e >> 5]) >> (e & 31u)) & 1u;
            return;
        }
        const uint32_t bn = (uint32_t)((rr & BND_MASK) >> HSH);
        const uint32_t t = bn ? bn - 1u : r;
        out |= 2u;
        const uint32_t bit = 1u << (t & 31u);
        if (luf_atomic_or(&troot[t >> 5], bit) & bit) return;
        const uint32_t idx = luf_atomic_add(&s.cnt[1], 1u);
        if (idx < LUF_T_LIST2_CAP) s.list2[idx] = t;
    });
    return out;
}

// The union-find forest is dead now; the node words are re-used by the traversal:
//   bits 0..15 tree edge towards the root    bits 16..31 link of the LIFO stack
// uf.py:326-344: the reference's LIFO traversal (neighbours in ascending edge
// id) from a tree root; only the tree (the edge towards the root of every
// discovered node) is kept.  `visited` marks discovered nodes.
LUF_DEV void build_tree(const GraphView& g, const Shot& s, uint32_t root)
{
    uint32_t* visited = claim_bits(s);
    luf_atomic_or(&visited[root >> 5], 1u << (root & 31u));
    s.S[root] = TREE_ROOT_W | (IDX_NONE << HSH);
    uint32_t top = root;
    LUF_NOUNROLL
    while (top != (uint32_t)IDX_NONE) {
        const uint32_t u = top;
        top = (uint32_t)(s.S[u] >> HSH);
        for_row(g, u, [&](ent_t ent) {
            const uint32_t e = ent_edge(ent);
            if (growth_of(s, e) != G_FULL) return;
            bool first;
            const uint32_t o = other_end(ent, first);
            const uint32_t bit = 1u << (o & 31u);
            if (visited[o >> 5] & bit) return;
            luf_atomic_or(&visited[o >> 5], bit);
            if (!(s.full2[o >> 5] & bit)) { s.S[o] = (word_t)e | (IDX_NONE << HSH); return; }     // a leaf: nothing to discover from it
            s.S[o] = (word_t)e | ((word_t)top << HSH); top = o;
        });
    }
}

// one lane per tree root
LUF_DEV void ph_peel_build(const GraphView& g, const Shot& s, int lane, int nl)
{
    const uint32_t total = s.cnt[1];
    const uint32_t n = total < LUF_T_LIST2_CAP ? total : LUF_T_LIST2_CAP;
    LUF_NOUNROLL
    for (uint32_t i = (uint32_t)lane; i < n; i += (uint32_t)nl) build_tree(g, s, (uint32_t)s.list2[i]);
}

// overflow: roots that did not fit the list are still undiscovered; by bitmap word
LUF_DEV void ph_peel_build_rest(const GraphView& g, const Shot& s, int lane, int nl)
{
    const uint32_t* troot = s.newfull;
    LUF_NOUNROLL
    for (int w = lane; w < g.WN; w += nl) {
        uint32_t bits = troot[w] & ~claim_bits(s)[w];
        LUF_NOUNROLL
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
    uint32_t parity = 0;
    const uint32_t* visited = claim_bits(s);
    for_defects(g, s, lane, nl, [&](uint32_t v) {
        if (!bit_of(visited, v)) return;
        uint32_t e = (uint32_t)(s.S[v] & LO_MASK);
        LUF_NOUNROLL
        while (e != (uint32_t)TREE_ROOT_W) {
            if (s.corr) luf_atomic_xor(&s.corr[e >> 5], 1u << (e & 31u));
            parity ^= (LUF_LDG(&g.west_mask[e >> 5]) >> (e & 31u)) & 1u;
            uint32_t eu, ev;
            edge_ends(g, e, eu, ev);
            v = eu == v ? ev : eu;
            e = (uint32_t)(s.S[v] & LO_MASK);
        }
    });
    return parity;
}

// ------------------------------------------------------------------ drivers
// uf.py:200-209 validate: rounds of grow / merge until no cluster is active.
// Returns the number of growth rounds.  `lane` is this thread's lane on the
// device; the host emulation iterates it inside LUF_PHASE.  Expects ph_load
// to have run just before (defects in the work list).
template <int TF>
LUF_DEV int uf_validate_t(const GraphView& g, const Shot& s, int flags, int lane, int nl)
{
    int rounds = 0;
    const bool vision_buckets = ((TF >= 0 ? TF : flags) & (FLAG_BUCKET | FLAG_NODE)) == FLAG_BUCKET;      // BUF
    uint32_t iter = 0xFFFFu;                 // decreasing reservation tag
    (void)lane;
    LUF_NOUNROLL
    for (;;) {
        uint32_t active = 0;
        if (vision_buckets) {
            if (rounds == 0) { LUF_T_PHASE(if (lane == 0) s.cnt[0] = 0); }      // ph_load's list of defects is not used
            LUF_T_PHASE(ph_grow_gather(g, s, lane, nl); for (int w = lane; w < g.WN; w += nl) claim_bits(s)[w] = 0);
            LUF_T_PHASE(ph_buf_mark(g, s, lane, nl));
            LUF_T_PHASE(ph_buf_count(g, s, lane, nl));
            LUF_T_PHASE(ph_buf_min(g, s, lane, nl));
            LUF_T_PHASE(active |= ph_buf_exec(g, s, lane, nl));
            LUF_T_PHASE(ph_buf_restore(g, s, lane, nl));
            LUF_T_PHASE(for (int w = lane; w < g.WN; w += nl) claim_bits(s)[w] = 0xFFFFFFFFu);
        } else if (rounds == 0) {              // every defect is an active singleton
            LUF_T_PHASE(active |= ph_grow_first(g, s, lane, nl); if (lane == 0) s.cnt[5] = s.cnt[0]);
        } else {
            if (!s.cnt[5]) break;              // union_roots keeps count of the active clusters: none left
            LUF_T_PHASE(ph_grow_gather(g, s, lane, nl));
            LUF_T_PHASE(ph_grow_filter<TF>(g, s, flags, lane, nl));
            LUF_T_PHASE(active |= ph_grow_exec<TF>(g, s, flags, lane, nl));
        }
        if (!LUF_T_ANY(active)) break;
        ++rounds;
        LUF_STAT(0, 1); LUF_STAT(rounds == 1 ? 1 : 2, active);
        LUF_NOUNROLL
        for (;;) {                             // chunks of at most LUF_T_LIST_CAP edges, lowest edge ids first
            uint32_t n = 0, left = 0;
            LUF_T_SCAN_PREP(merge_count(g, s, lane, nl));
            LUF_T_PHASE(ph_merge_gather(g, s, lane, nl));
            LUF_T_PHASE(n = s.cnt[2]; left = s.cnt[3]);
            LUF_STAT(3, n); LUF_STAT(4, n > 32); LUF_STAT(rounds == 1 ? 8 : 9, n);
            uint32_t done = 0;
            LUF_NOUNROLL
            for (uint32_t guard = 0; guard <= n; ++guard) {           // at least one edge is placed per iteration
                if (iter == 0) {                                      // tags used up (huge graphs): start over
                    LUF_T_PHASE(for (int w = lane; w < LUF_T_CLAIM_SLOTS; w += nl) s.claim[w] = WORD_ONES);
                    iter = 0xFFFFu;
                }
                const word_t tag = (word_t)(iter--) << HSH;
                uint32_t placed = 0;
                LUF_T_PHASE(placed += ph_merge_claim<TF>(g, s, flags, n, tag, guard == 0, lane, nl));
                LUF_T_PHASE(placed += ph_merge_exec<TF>(g, s, flags, n, tag, lane, nl));
                done += LUF_T_SUM(placed);
                LUF_STAT(5, 1); LUF_STAT(rounds == 1 ? 6 : 7, 1);
                if (done >= n) break;
            }
            if (!left) break;
        }
    }
    // hand over to the peel: counters cleared, visited bitmap (in the claim table) cleared
    LUF_T_PHASE(if (lane < 4) s.cnt[lane] = 0; for (int w = lane; w < g.WN; w += nl) claim_bits(s)[w] = 0);
    return rounds;
}

// ---------------------------------------------------- order-free merge (fused kernels, static UF)
// What a growth round needs from the merges before it is the partition into clusters and, per cluster, its
// size, its defect parity and whether it touches a boundary -- none of which depends on the order of the
// merges.  The order only decides which node is the union-find root and which boundary node represents the
// cluster, i.e. where the peel roots its trees.  The fused Monte-Carlo kernels do not peel as long as no
// cluster holds boundary nodes of both sides (uf_peel_parity), so they merge a round's newly FULL edges in
// any order: one lane per edge hooks the smaller root below the larger with a compare-and-swap (strict
// order by (size at the start of the trip, index): no cycles whatever the interleaving), then adds the
// hooked root's size / parity / boundary to the root it ends up under.  Two phases per 32 (or blockDim)
// edges instead of two per reservation iteration (19 iterations per cfg3 shot).  A shot in which a cluster
// comes to hold boundary nodes of different sides (cnt[4]) is decoded again in canonical order.
LUF_DEV void ph_fmerge_hook(const GraphView& g, const Shot& s, uint32_t i, int lane)
{
    const uint32_t e = (uint32_t)(s.list[i] & LO_MASK);
    uint32_t u, v;
    edge_ends(g, e, u, v);
    s.list2[lane] = 0u;                                           // nothing hooked (a payload always has ROOT_BIT)
    LUF_NOUNROLL
    for (;;) {
        word_t wu, wv;
        const uint32_t ru = find_pc(s, u, wu), rv = find_pc(s, v, wv);
        if (ru == rv) return;
        const uint32_t su = (uint32_t)(wu & SIZE_MASK), sv = (uint32_t)(wv & SIZE_MASK);
        const bool u_small = su < sv || (su == sv && ru > rv);
        const uint32_t small = u_small ? ru : rv, large = u_small ? rv : ru;
        const word_t wsmall = u_small ? wu : wv;
        if (luf_atomic_cas(&s.S[small], wsmall, (word_t)large | ((word_t)e << HSH)) == wsmall) {
            s.list[i] = small;
            s.list2[lane] = wsmall;
            return;
        }
    }
}

template <int TF>
LUF_DEV void ph_fmerge_apply(const GraphView& g, const Shot& s, uint32_t i, int lane)
{
    const word_t pay = s.list2[lane];
    if (!pay) return;
    word_t wr;
    const uint32_t R = find_ro(s, (uint32_t)s.list[i], wr);       // roots are stable in this phase
    const word_t bs = pay & BND_MASK;
    LUF_NOUNROLL
    for (;;) {
        word_t bl = wr & BND_MASK;
        if (bl && bs) {
            const int ll = LUF_LDG(&g.node_long[(uint32_t)(bl >> HSH) - 1]), ls = LUF_LDG(&g.node_long[(uint32_t)(bs >> HSH) - 1]);
            if (ll != ls) s.cnt[4] = 1u;
            if ((TF & FLAG_WEST) && ls < ll) bl = bs;
        } else if (!bl) bl = bs;
        const word_t neww = ((wr ^ pay) & ODD_BIT) | bl | ROOT_BIT | ((wr & SIZE_MASK) + (pay & SIZE_MASK));
        const word_t old = luf_atomic_cas(&s.S[R], wr, neww);
        if (old == wr) {                                          // cnt[5]: the number of active clusters
            const int delta = (int)is_active(neww) - (int)is_active(wr) - (int)is_active(pay);
            if (delta) luf_atomic_add(&s.cnt[5], (uint32_t)delta);
            return;
        }
        wr = old;
    }
}

// This round's newly FULL edges into the list, in no particular order (one atomic add per lane instead of the
// ordered compaction's scan).  cnt[6] (0 on entry) hands out the positions and ends up as the number of newly FULL
// edges; those beyond the list's capacity stay in the bitmap for the next chunk.
LUF_DEV void ph_fmerge_gather(const GraphView& g, const Shot& s, int lane, int nl)
{
    const int wpl = (g.WE + nl - 1) / nl;
    const int w0 = lane * wpl, w1 = w0 + wpl < g.WE ? w0 + wpl : g.WE;
    if (lane == 0) { s.cnt[0] = 0; s.cnt[1] = 0; }                // what ph_grow_gather expects
    uint32_t c = 0;
    LUF_NOUNROLL
    for (int w = w0; w < w1; ++w) c += (uint32_t)luf_popc(s.newfull[w]);
    if (!c) return;
    uint32_t pos = luf_atomic_add(&s.cnt[6], c);
    LUF_NOUNROLL
    for (int w = w0; w < w1 && pos < LUF_T_LIST_CAP; ++w) {
        uint32_t bits = s.newfull[w];
        LUF_NOUNROLL
        while (bits && pos < LUF_T_LIST_CAP) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            s.list[pos++] = (uint32_t)(32 * w + b);
        }
        s.newfull[w] = bits;
    }
}

// uf_validate_t with the order-free merge.  TF: any of FLAG_WEST, FLAG_NODE, FLAG_NODE | FLAG_BUCKET (cluster
// sizes, the bucket key of NodeBUF, do not depend on the merge order either).
template <int TF>
LUF_DEV int uf_validate_fast(const GraphView& g, const Shot& s, int lane, int nl)
{
    int rounds = 0;
    (void)lane;
    const uint32_t width = (uint32_t)nl < LUF_T_LIST2_CAP ? (uint32_t)nl : LUF_T_LIST2_CAP;
    LUF_NOUNROLL
    for (;;) {
        uint32_t active = 0;
        if (rounds == 0) {                     // every defect is an active singleton
            LUF_T_PHASE(active |= ph_grow_first(g, s, lane, nl); if (lane == 0) s.cnt[5] = s.cnt[0]);
        } else {
            // the merges keep count of the active clusters (cnt[5]): no sweep over the touched nodes just to
            // find that none is left
            if (!s.cnt[5]) break;
            LUF_T_PHASE(ph_grow_gather(g, s, lane, nl));
            LUF_T_PHASE(ph_grow_filter<TF>(g, s, TF, lane, nl));
            LUF_T_PHASE(active |= ph_grow_exec<TF>(g, s, TF, lane, nl));
        }
        if (!LUF_T_ANY(active)) break;
        ++rounds;
        LUF_NOUNROLL
        for (;;) {                             // chunks of at most LUF_T_LIST_CAP edges
            uint32_t n = 0, left = 0;
            LUF_T_PHASE(ph_fmerge_gather(g, s, lane, nl));
            LUF_T_PHASE(const uint32_t total = s.cnt[6]; n = total < LUF_T_LIST_CAP ? total : LUF_T_LIST_CAP; left = total - n);
            LUF_T_PHASE(if (lane == 0) s.cnt[6] = 0);
            LUF_NOUNROLL
            for (uint32_t base = 0; base < n; base += width) {
                LUF_T_PHASE(if ((uint32_t)lane < width && base + (uint32_t)lane < n) ph_fmerge_hook(g, s, base + (uint32_t)lane, lane));
                LUF_T_PHASE(if ((uint32_t)lane < width && base + (uint32_t)lane < n) ph_fmerge_apply<TF>(g, s, base + (uint32_t)lane, lane));
            }
            if (!left) break;
        }
    }
    LUF_T_PHASE(if (lane < 4) s.cnt[lane] = 0);
    return rounds;
}

LUF_DEV int uf_validate(const GraphView& g, const Shot& s, int flags, int lane, int nl)
{
    return uf_validate_t<-1>(g, s, flags, lane, nl);
}

// uf.py:305-312 peel.  Returns (per lane on the device, total in the
// emulation) the parity of correction edges on the west boundary.
LUF_DEV uint32_t uf_peel(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t parity = 0, general = 0;
    (void)lane;
    LUF_T_PHASE(ph_peel_gather(g, s, lane, nl));
    LUF_T_PHASE(const uint32_t r = ph_tree_roots(g, s, lane, nl); parity ^= r & 1u; general |= r >> 1);
    if (LUF_T_ANY(general)) {
        uint32_t roots = 0;
        LUF_T_PHASE(ph_peel_build(g, s, lane, nl); roots = s.cnt[1]);
        LUF_STAT(10, 1); LUF_STAT(11, roots);
        if (roots > LUF_T_LIST2_CAP) { LUF_T_PHASE(ph_peel_build_rest(g, s, lane, nl)); }
        LUF_T_PHASE(parity ^= ph_peel_paths(g, s, lane, nl));
    }
    return parity;
}

// The west parity of the correction without building it (the fused Monte-Carlo kernels need nothing else).
// As long as no cluster holds boundary nodes of both sides, every correction inside the fully grown edges of
// a cluster has the same parity on the west boundary: two of them differ by closed cycles (two west edges at
// every west boundary node they pass) and by paths between boundary nodes of one side (two west edges, or
// none).  That parity is the number of the cluster's defects if its boundary nodes are west ones (an odd
// cluster sends one string to the boundary, an even one an even number), else 0 -- so every defect adds 1 iff
// its cluster's boundary node is a west one; no forest, no traversal.  cnt[4] (union_roots) says whether a
// cluster with boundary nodes of different sides exists; such shots take the full peel -- unless every
// boundary node of the graph has exactly one edge (g.leaf_bnd; every Surface / Repetition graph): the peel
// roots a cluster's tree at cluster.boundary (uf.py:314-324), every other boundary node of the cluster is then
// a leaf that is no defect, so its edge is never in the correction (uf.py:305-312), and the root's single
// edge is iff the cluster is odd.  The west parity of the correction is the cluster's defect parity if
// cluster.boundary is a west node, else 0 -- the same pass over the defects, whatever sides the cluster
// touches, as long as the merges ran in canonical order (which decides which boundary node the cluster keeps).
LUF_DEV uint32_t ph_defect_sides(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t parity = 0;
    LUF_NOUNROLL
    for (int w = lane; w < g.WD; w += nl) {
        uint32_t bits = s.defect[w];
        LUF_NOUNROLL
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            word_t rr;
            find_ro(s, (uint32_t)(g.B + 32 * w + b), rr);
            const uint32_t bn = (uint32_t)((rr & BND_MASK) >> HSH);
            if (bn && LUF_LDG(&g.node_long[bn - 1u]) == -1) parity ^= 1u;
        }
    }
    return parity;
}

LUF_DEV uint32_t uf_peel_parity(const GraphView& g, const Shot& s, int lane, int nl)
{
    uint32_t mixed = 0, parity = 0;
    (void)lane;
    LUF_T_PHASE(mixed = (g.leaf_bnd ? 0u : s.cnt[4]) | (g.parity_ok ? 0u : 1u));
    if (LUF_T_ANY(mixed)) return uf_peel(g, s, lane, nl);
    LUF_T_PHASE(parity ^= ph_defect_sides(g, s, lane, nl));
    return parity;
}

}  // namespace LUF_UF_NS
}  // namespace luf
