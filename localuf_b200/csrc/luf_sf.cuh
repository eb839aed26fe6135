// luf_sf.cuh -- Snowflake (localuf/decoders/snowflake/main.py) under the frugal
// scheme (localuf/_schemes.py:480-649): one warp owns one stream; the sliding
// window's whole decoder state, the window's error bits and the string pairs
// of the logical counter stay resident in the warp's shared-memory slice for
// the lifetime of the stream.
//
// Layout (localuf_b200/_window.py): nodes are indexed r*NL + q with r = h-1-t
// the layer counted from the top and q the in-layer position; per-node bit
// fields are bitmaps of WLn words per layer; edges are indexed
// b*ELp + l with b = h-1-tmin the layer block from the top, l the local index
// (ELp = EL rounded up to 32).  "Drop one layer" (snowflake/main.py:298-310) is
// then a copy of every layer to the next one; cluster ids keep the reference's
// own time-reversed numbering (snowflake/main.py:159-199) because merging
// compares them.
//
// A merging timestep is two phases, as for Macar: A = every node floods and
// syncs (snowflake/main.py:1097-1129) reading committed neighbour state and
// writing only its own next_* fields (toggles of a pointee's next_defect and
// of an edge's correction go through shared-memory atomics); B = commit
// (update_after_merging, :1131-1137) and warp votes.  Dual compilation as in
// luf_core.cuh.
#pragma once
#include "luf_core.cuh"

namespace luf {

static const uint32_t SF_RESET = 0xFFFFu;     // snowflake/constants.py:5 (RESET = -1)
static const uint32_t SF_PTR_C = 0xFu;       // pointer 'C' (no neighbour) in the 4-bit pointer field
// per-node flag word (written by the owning lane only)
enum { SF_ACTIVE = 1, SF_WHOLE = 2, SF_DEFECT = 4, SF_GROWN = 8, SF_UNROOTED = 16, SF_BUSY = 32,
       SF_N_ACTIVE = 64, SF_N_GROWN = 128, SF_N_UNROOTED = 256 };
enum { SF_FLAG_SLOW = 1, SF_FLAG_ONE_ONE = 2, SF_FLAG_SYNDROME = 4, SF_FLAG_LATENCY_TAIL = 8 };
// string-pair codes of an endpoint's partner (see oracle/luf_oracle_snowflake.c)
static const int SFP_NONE = -1, SFP_WEST = -2, SFP_EAST = -3, SFP_DEAD = -4;

struct SfView {
    int d, h, NLb, NLd, NL, EL, ELw, K, WLn;   // ELw = words per edge block, WLn = words per node-layer bitmap
    int N, n_classes;
    const uint32_t* nbr;        // [NL*K]  valid:1 | q2:10 | el:12 | dt+1:2 | et+1:1 | slot of this edge at the neighbour:4   (see sf_nbr)
    const uint32_t* edge_ends;  // [EL]    q0:10 | dt0:1 | q1:10 | dt1:1 | fb:1
    const int8_t* pos_long;     // [NL]
    const uint16_t* fresh_perm; // [EL]    local edges sorted by probability class
    const uint32_t* class_end;  // [n_classes]
};

#ifndef LUF_SF_ALIST_CAP
  #define LUF_SF_ALIST_CAP 384
#endif
static const uint32_t SF_ALIST_CAP = LUF_SF_ALIST_CAP;      // warp tiles; tests shrink it to force the bitmap sweep
#ifndef LUF_SF_CTA_ALIST_CAP
  #define LUF_SF_CTA_ALIST_CAP 2048
#endif
static const uint32_t SF_CTA_ALIST_CAP = LUF_SF_CTA_ALIST_CAP;   // block tiles (large windows)

struct SfLayout {
    uint32_t cid, ncid;         // u16[N]
    uint32_t fl;                // u16[N]
    uint32_t acp;               // u16[N]  bits 0..11: neighbour slots whose edge is fully grown (access); bits 12..15: pointer
    uint32_t ndef;              // u32[h*WLn]  next_defect (toggled by neighbours)
    uint32_t awake;             // u32[h*WLn]  nodes that ever saw a defect or a fully grown edge while in the window
    uint32_t growth;            // u32[h*2*ELw]  2 bits per edge: block b at word b*2*ELw
    uint32_t corr, err;         // u32[h*ELw]
    uint32_t top, fb;           // u32[WLn]  this cycle's top-sheet syndrome / defects on the future boundary
    uint32_t partner;           // i16[2*NL]
    uint32_t alist_cap;         // entries of alist (SF_ALIST_CAP for warp tiles, SF_CTA_ALIST_CAP for block tiles)
    uint32_t alist, acnt;       // u16[alist_cap] the awake nodes, compacted as sheet << 10 | position (rebuilt after
                                // every drop and grow step); u32[4] its fill
    uint32_t bytes;
};

struct SfShot {
    uint16_t *cid, *ncid, *fl;
    uint16_t* acp;
    uint32_t *ndef, *awake, *growth, *corr, *err, *top, *fb, *acnt;
    uint16_t* alist;
    int16_t* partner;
    uint32_t acap;              // capacity of alist
};

LUF_DEV SfShot sf_bind(const SfLayout& l, unsigned char* b)
{
    SfShot s;
    s.cid = (uint16_t*)(b + l.cid); s.ncid = (uint16_t*)(b + l.ncid); s.fl = (uint16_t*)(b + l.fl);
    s.acp = (uint16_t*)(b + l.acp);
    s.ndef = (uint32_t*)(b + l.ndef); s.awake = (uint32_t*)(b + l.awake); s.growth = (uint32_t*)(b + l.growth);
    s.alist = (uint16_t*)(b + l.alist); s.acnt = (uint32_t*)(b + l.acnt);
    s.corr = (uint32_t*)(b + l.corr); s.err = (uint32_t*)(b + l.err);
    s.top = (uint32_t*)(b + l.top); s.fb = (uint32_t*)(b + l.fb);
    s.partner = (int16_t*)(b + l.partner);
    s.acap = l.alist_cap;
    return s;
}

// Snowflake's own id of node (r, q): snowflake/main.py:159-179
LUF_DEV uint32_t sf_id(const SfView& g, int r, int q)
{
    return q < g.NLb ? (uint32_t)(g.NLb * r + q) : (uint32_t)(g.NLb * g.h + g.NLd * r + (q - g.NLb));
}

// Neighbour slot k of node (r, q).  Returns false if absent in the window.
struct SfNbr { int node, r2, q2; uint32_t e, back; };   // node index, layered edge bit index, this edge's slot at the neighbour
LUF_DEV bool sf_nbr(const SfView& g, int r, int q, int k, SfNbr& o)
{
    const uint32_t ent = LUF_LDG(&g.nbr[q * g.K + k]);
    if (!(ent & 1u)) return false;
    const int dt = (int)((ent >> 23) & 3u) - 1;
    const int r2 = r - dt;
    if (r2 < 0 || r2 >= g.h) return false;
    o.r2 = r2; o.q2 = (int)((ent >> 1) & 1023u);
    o.node = r2 * g.NL + o.q2;
    const int et = (int)((ent >> 25) & 1u) - 1;
    o.e = (uint32_t)((r - et) * g.ELw * 32) + ((ent >> 11) & 4095u);
    o.back = (ent >> 26) & 15u;
    return true;
}

LUF_DEV uint32_t sf_growth(const SfShot& s, uint32_t e) { return (s.growth[e >> 4] >> ((e & 15u) * 2u)) & 3u; }

// Access mask and pointer of node v share one 16-bit word.  Lanes other than the owner set access
// bits (the far end of an edge that just became fully grown), so every update made in a phase with
// such writers goes through a 32-bit atomic OR on the containing pair of nodes.
LUF_DEV uint32_t sf_ptr(const SfShot& s, int v) { return (uint32_t)s.acp[v] >> 12; }
LUF_DEV uint32_t sf_access(const SfShot& s, int v) { return (uint32_t)s.acp[v] & 0xFFFu; }
LUF_DEV void sf_set_ptr(const SfShot& s, int v, uint32_t ptr) { s.acp[v] = (uint16_t)((s.acp[v] & 0xFFFu) | (ptr << 12)); }
LUF_DEV void sf_or_acp(const SfShot& s, int v, uint32_t bits)
{
#if defined(LUF_HOST_EMU)
    s.acp[v] = (uint16_t)(s.acp[v] | bits);         // lanes run one after the other
#else
    luf_atomic_or((uint32_t*)(s.acp + (v & ~1)), bits << (16 * (v & 1)));
#endif
}

// Snowflake.reset + Frugal.reset (snowflake/main.py:201-214,1011-1031,1489-1496; _schemes.py:498-502)
LUF_DEV void sf_reset(const SfView& g, const SfShot& s, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int r = 0; r < g.h; ++r)
        LUF_NOUNROLL_CA
        for (int q = lane; q < g.NL; q += nl) {
            const int v = r * g.NL + q;
            const uint32_t id = sf_id(g, r, q);
            s.cid[v] = (uint16_t)id; s.ncid[v] = (uint16_t)id;
            s.fl[v] = (uint16_t)SF_WHOLE; s.acp[v] = (uint16_t)(SF_PTR_C << 12);
        }
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.h * g.WLn; w += nl) { s.ndef[w] = 0; s.awake[w] = 0; }
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.h * 2 * g.ELw; w += nl) s.growth[w] = 0;
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.h * g.ELw; w += nl) { s.corr[w] = 0; s.err[w] = 0; }
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WLn; w += nl) { s.top[w] = 0; s.fb[w] = 0; }
    LUF_NOUNROLL_CA
    for (int k = lane; k < 2 * g.NL; k += nl) s.partner[k] = (int16_t)SFP_NONE;
}

// ------------------------------------------------------------------- Pairs
// _pairs.py:4-66 on endpoint keys t*NL + q (t in {0,1}); an endpoint that can
// no longer be extended is only remembered by its side.  LogicalCounter.count
// (_pairs.py:90-111) is applied when a pair becomes final.
LUF_DEV int sf_endpoint_key(const SfView& g, uint32_t q, uint32_t t)
{
    if ((int)q < g.NLb) return LUF_LDG(&g.pos_long[q]) == -1 ? SFP_WEST : SFP_EAST;
    return (int)(t * (uint32_t)g.NL + q);
}

LUF_DEV void sf_pairs_add(const SfShot& s, int a, int b, int& count)
{
    if (a >= 0) s.partner[a] = (int16_t)b;
    if (b >= 0) s.partner[b] = (int16_t)a;
    if ((a == SFP_WEST && b == SFP_EAST) || (a == SFP_EAST && b == SFP_WEST)) ++count;
}

LUF_DEV int sf_pairs_take(const SfShot& s, int u)       // remove u's pair, return its partner
{
    const int w = s.partner[u];
    s.partner[u] = (int16_t)SFP_NONE;
    if (w >= 0) s.partner[w] = (int16_t)SFP_NONE;
    return w;
}

LUF_DEV void sf_pairs_load(const SfShot& s, int u, int v, int& count)
{
    if (u >= 0 && s.partner[u] != SFP_NONE) {
        const int w = sf_pairs_take(s, u);
        if (v >= 0 && s.partner[v] != SFP_NONE) sf_pairs_add(s, w, sf_pairs_take(s, v), count);
        else if (v != w || v < 0) sf_pairs_add(s, v, w, count);
    } else if (v >= 0 && s.partner[v] != SFP_NONE)
        sf_pairs_add(s, u, sf_pairs_take(s, v), count);
    else
        sf_pairs_add(s, u, v, count);
}

// Hand the set bits of one floor-sheet bitmap (block h-1) to Pairs in index
// order = ascending code edge id.  One lane.
LUF_DEV void sf_pairs_load_block(const SfView& g, const SfShot& s, const uint32_t* bits, int& count)
{
    LUF_NOUNROLL_CA
    for (int w = 0; w < g.ELw; ++w) {
        uint32_t x = bits[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            const uint32_t ee = LUF_LDG(&g.edge_ends[32 * w + b]);
            sf_pairs_load(s, sf_endpoint_key(g, ee & 1023u, (ee >> 10) & 1u),
                          sf_endpoint_key(g, (ee >> 11) & 1023u, (ee >> 21) & 1u), count);
        }
    }
}

// lowering of the kept pairs by one layer (_pairs.py:82-111), two phases over the detectors of a
// layer.  (a) strings that end in the committed layer 0 are final: both ends in layer 0 -> gone,
// other end in layer 1 -> that end can no longer be extended (DEAD); every layer-0 key is
// cleared by the lane that owns it, a layer-1 key is written only by its partner's lane.
LUF_DEV void sf_pairs_retire(const SfView& g, const SfShot& s, int lane, int nl)
{
    const int NL = g.NL;
    LUF_NOUNROLL_CA
    for (int q = g.NLb + lane; q < NL; q += nl) {
        const int p = s.partner[q];
        if (p == SFP_NONE) continue;
        s.partner[q] = (int16_t)SFP_NONE;
        if (p >= NL) s.partner[p] = (int16_t)SFP_DEAD;
    }
}

// (b) the keys of layer 1 move down to layer 0
LUF_DEV void sf_pairs_lower(const SfView& g, const SfShot& s, int lane, int nl)
{
    const int NL = g.NL;
    LUF_NOUNROLL_CA
    for (int q = g.NLb + lane; q < NL; q += nl) {
        const int p = s.partner[NL + q];
        if (p == SFP_NONE) continue;
        s.partner[NL + q] = (int16_t)SFP_NONE;
        s.partner[q] = (int16_t)(p >= NL ? p - NL : p);
    }
}

// ------------------------------------------------------- window bookkeeping
// Frugal._raise_window (_schemes.py:532-543): shift the error bits one layer down.
LUF_DEV void sf_shift_errors(const SfView& g, const SfShot& s, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int c = lane; c < g.ELw; c += nl)
        LUF_NOUNROLL_CA
        for (int b = g.h - 1; b >= 1; --b) s.err[b * g.ELw + c] = s.err[(b - 1) * g.ELw + c];
}

// Frugal._load (_schemes.py:545-564), first half: the carried future-boundary
// defects become this cycle's top-sheet syndrome.
LUF_DEV void sf_begin_sheet(const SfView& g, const SfShot& s, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WLn; w += nl) { s.top[w] = s.fb[w]; s.fb[w] = 0; }
    LUF_NOUNROLL_CA
    for (int c = lane; c < g.ELw; c += nl) s.err[c] = 0;
}

// A freshly flipped edge (local index l of the top block): record it and toggle
// its detector endpoints -- on the top sheet, or on the future boundary.
LUF_DEV void sf_flip_fresh(const SfView& g, const SfShot& s, uint32_t l, bool exclude_fb)
{
    const uint32_t ee = LUF_LDG(&g.edge_ends[l]);
    if (exclude_fb && ((ee >> 22) & 1u)) return;               // _base_classes.py:418-424
    luf_atomic_or(&s.err[l >> 5], 1u << (l & 31u));
    LUF_NOUNROLL_CA
    for (int x = 0; x < 2; ++x) {
        const uint32_t q = (ee >> (11 * x)) & 1023u, dt = (ee >> (11 * x + 10)) & 1u;
        if ((int)q < g.NLb) continue;
        luf_atomic_xor(dt ? &s.fb[q >> 5] : &s.top[q >> 5], 1u << (q & 31u));
    }
}

// Fresh errors given as bits (staged / parity path).
LUF_DEV void sf_given_sheet(const SfView& g, const SfShot& s, const uint32_t* fresh, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.ELw; w += nl) {
        uint32_t x = fresh[w];
        LUF_NOUNROLL_CA
        while (x) {
            const int b = luf_ffs(x) - 1; x &= x - 1;
            sf_flip_fresh(g, s, (uint32_t)(32 * w + b), false);
        }
    }
}

// The top sheet's syndrome given directly (Snowflake.decode(syndrome), snowflake/main.py:216-256;
// generate_output, :656-726): bit q of `syn` = defect at in-layer position q of the new layer.
LUF_DEV void sf_given_syndrome(const SfView& g, const SfShot& s, const uint32_t* syn, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WLn; w += nl) s.top[w] ^= syn[w];
}

// Fresh errors sampled on the device: geometric skips over the class-sorted
// local edges, Philox counter = (stream, lane, cycle).  Same scheme as ph_sample.
LUF_DEV void sf_sample_sheet(const SfView& g, const SfShot& s, const SampleParams& sp, uint64_t stream,
                             uint32_t cycle, bool exclude_fb, int lane, int nl)
{
    const uint32_t a = (uint32_t)(((uint64_t)g.EL * (uint32_t)lane) / (uint32_t)nl);
    const uint32_t b = (uint32_t)(((uint64_t)g.EL * (uint32_t)(lane + 1)) / (uint32_t)nl);
    if (a >= b) return;
    Philox rng;
    rng.c0 = (uint32_t)stream; rng.c1 = (uint32_t)(stream >> 32); rng.c2 = (uint32_t)lane; rng.c3 = cycle << 6;
    rng.k0 = sp.key0; rng.k1 = sp.key1 ^ 0x5F0F5F0Fu; rng.have = 0;
    rng.o0 = rng.o1 = rng.o2 = rng.o3 = 0;
    uint32_t start = 0;
    LUF_NOUNROLL_CA
    for (int c = 0; c < g.n_classes; ++c) {
        const uint32_t end = LUF_LDG(&g.class_end[c]);
        const uint32_t lo = a > start ? a : start, hi = b < end ? b : end;
        start = end;
        if (lo >= hi) { if (end >= b) break; continue; }
        const float p = sp.p[c];
        if (p <= 0.0f) continue;
        uint32_t pos = lo;
        LUF_NOUNROLL_CA
        for (;;) {
            if (p < 1.0f) {
                const float u = ((float)philox_next(rng) + 1.0f) * 2.3283064365386963e-10f;
                const float gf = luf_log2(u) * sp.inv_log2_q[c];
                const uint32_t gap = gf >= 1.0e9f ? 1000000000u : (uint32_t)gf;
                if (gap >= hi - pos) break;
                pos += gap;
            }
            sf_flip_fresh(g, s, LUF_LDG(&g.fresh_perm[pos]), exclude_fb);
            if (++pos >= hi) break;
        }
    }
}

// Snowflake.drop (snowflake/main.py:298-310): contacts (:1547-1564), friendships
// (:1161-1211), _load (:312-322), update_after_drop (:1033-1043,1498-1506).
// Every lane owns whole columns (an edge word / an in-layer position through
// all layers), so the in-place shift needs no barrier.
LUF_DEV void sf_drop(const SfView& g, const SfShot& s, int lane, int nl)
{
    const int h = g.h;
    LUF_NOUNROLL_CA
    for (int c = lane; c < g.ELw; c += nl) {
        LUF_NOUNROLL_CA
        for (int b = h - 1; b >= 1; --b) s.corr[b * g.ELw + c] = s.corr[(b - 1) * g.ELw + c];
        s.corr[c] = 0;
    }
    LUF_NOUNROLL_CA
    for (int c = lane; c < 2 * g.ELw; c += nl) {
        LUF_NOUNROLL_CA
        for (int b = h - 1; b >= 1; --b) s.growth[b * 2 * g.ELw + c] = s.growth[(b - 1) * 2 * g.ELw + c];
        s.growth[c] = 0;
    }
    LUF_NOUNROLL_CA
    for (int q = lane; q < g.NL; q += nl) {
        LUF_NOUNROLL_CA
        for (int r = h - 1; r >= 1; --r) {
            const int v = r * g.NL + q, u = v - g.NL;
            const uint32_t fu = s.fl[u], cu = s.cid[u];
            const uint32_t c = cu + (cu < (uint32_t)(g.NLb * h) ? (uint32_t)g.NLb : (uint32_t)g.NLd);   // id_below, :186-199
            uint32_t f = fu & (SF_ACTIVE | SF_WHOLE | SF_DEFECT);
            if (f & SF_ACTIVE) f |= SF_N_ACTIVE;
            s.fl[v] = (uint16_t)f;
            s.cid[v] = (uint16_t)c; s.ncid[v] = (uint16_t)c;
            s.acp[v] = s.acp[u];
        }
        // top sheet: TopSheetFriendship (:1201-1211); next_cid / next_pointer / next_whole keep their values
        const uint32_t syn = (q >= g.NLb) ? (s.top[q >> 5] >> (q & 31)) & 1u : 0u;
        s.fl[q] = (uint16_t)(SF_WHOLE | (syn ? SF_DEFECT : 0));
        s.cid[q] = s.ncid[q];
        s.acp[q] = (uint16_t)(SF_PTR_C << 12);
    }
    // next_defect == defect for every node before a drop (r.next_defect = node.defect, :1195),
    // so it shifts like everything else; the top sheet receives the new syndrome (:312-322)
    LUF_NOUNROLL_CA
    for (int c = lane; c < g.WLn; c += nl) {
        LUF_NOUNROLL_CA
        for (int r = h - 1; r >= 1; --r) {
            s.ndef[r * g.WLn + c] = s.ndef[(r - 1) * g.WLn + c];
            s.awake[r * g.WLn + c] = s.awake[(r - 1) * g.WLn + c];
        }
        s.ndef[c] = s.top[c];
        s.awake[c] = s.top[c];
    }
}

// Only awake nodes can act in a grow or merging step: an un-awake node has no
// defect, is inactive, has cid == ID, pointer C and no fully grown edge, so
// growing, flooding and syncing all leave it unchanged and idle.
template <class F>
LUF_DEV void sf_for_awake(const SfView& g, const SfShot& s, int lane, int nl, F&& f)
{
    const uint32_t total = s.acnt[0];
    if (total <= s.acap) {                          // balanced: one awake node per lane and trip
        LUF_NOUNROLL_CA
        for (uint32_t i = (uint32_t)lane; i < total; i += (uint32_t)nl) {
            const int it = (int)s.alist[i], r = it >> 10, q = it & 1023;
            f(r, q, r * g.NL + q);
        }
        return;
    }
    const int words = g.h * g.WLn;
    LUF_NOUNROLL_CA
    for (int x = lane; x < words; x += nl) {
        uint32_t bits = s.awake[x];
        if (!bits) continue;
        const int r = x / g.WLn, w = x - r * g.WLn;
        LUF_NOUNROLL_CA
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const int q = 32 * w + b;
            f(r, q, r * g.NL + q);
        }
    }
}

// Per-cycle metrics of Snowflake.decode (snowflake/main.py:266-297), the part that needs the node state:
// min_defect_height (:283-288: the lowest layer with a defect on a detector, h if none) into acnt[1] and
// min_active_layer (:290-296: the lowest layer with an active node, h if none) into acnt[2] (both preset to h);
// the window's growth goes out as it lies (swim distance and unclustered edge fraction are k_dcs's work).
LUF_DEV void sf_metrics_scan(const SfView& g, const SfShot& s, uint32_t* growth_out, int lane, int nl)
{
    const int words = g.h * 2 * g.ELw;
    LUF_NOUNROLL_CA
    for (int w = lane; w < words; w += nl) growth_out[w] = s.growth[w];
    sf_for_awake(g, s, lane, nl, [&](int r, int q, int v) {
        const uint32_t f = s.fl[v];
        if (q >= g.NLb && (f & SF_DEFECT)) luf_atomic_min(&s.acnt[1], (uint32_t)(g.h - 1 - r));
        if (f & SF_ACTIVE) luf_atomic_min(&s.acnt[2], (uint32_t)(g.h - 1 - r));
    });
}

// Compact the awake bitmap into alist (acnt[0] = number of awake nodes; must be 0 on entry).
// Windows taller than 63 sheets keep the bitmap sweep (acnt[0] is then left above the capacity).
LUF_DEV void sf_gather_awake(const SfView& g, const SfShot& s, int lane, int nl)
{
    if (g.h > 63) { if (lane == 0) s.acnt[0] = s.acap + 1u; return; }
    const int words = g.h * g.WLn;
    const int wpl = (words + nl - 1) / nl;
    const int w0 = lane * wpl, w1 = w0 + wpl < words ? w0 + wpl : words;
    uint32_t c = 0;
    LUF_NOUNROLL_CA
    for (int x = w0; x < w1; ++x) c += (uint32_t)luf_popc(s.awake[x]);
    if (!c) return;
    uint32_t pos = luf_atomic_add(&s.acnt[0], c);
    if (pos + c > s.acap) return;
    LUF_NOUNROLL_CA
    for (int x = w0; x < w1; ++x) {
        uint32_t bits = s.awake[x];
        if (!bits) continue;
        const int r = x / g.WLn, w = x - r * g.WLn;
        LUF_NOUNROLL_CA
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            s.alist[pos++] = (uint16_t)((r << 10) | (32 * w + b));
        }
    }
}

// _Node._grow, snowflake/main.py:1075-1084
LUF_DEV void sf_grow_node(const SfView& g, const SfShot& s, int r, int q)
{
    LUF_NOUNROLL_CA
    for (int k = 0; k < g.K; ++k) {
        SfNbr nb;
        if (!sf_nbr(g, r, q, k, nb)) continue;
        const uint32_t e = nb.e, sh = (e & 15u) * 2u;
        if (((s.growth[e >> 4] >> sh) & 3u) >= G_FULL) continue;
        const uint32_t old = (luf_atomic_add(&s.growth[e >> 4], 1u << sh) >> sh) & 3u;
        if (old >= G_FULL) luf_atomic_sub(&s.growth[e >> 4], 1u << sh);
        else if (old == G_HALF) {                                  // newly FULL: both ends have access from now on
            luf_atomic_or(&s.awake[r * g.WLn + (q >> 5)], 1u << (q & 31));
            luf_atomic_or(&s.awake[nb.r2 * g.WLn + (nb.q2 >> 5)], 1u << (nb.q2 & 31));
            sf_or_acp(s, r * g.NL + q, 1u << k);
            if (nb.back != 15u) sf_or_acp(s, nb.node, 1u << nb.back);
        }
    }
}

// grow / grow_whole / grow_half (snowflake/main.py:1051-1073) + find_broken_pointers
// (:1221-1223) + _FullUnrooter.start (:1387-1389).  mode: 0 = 1:1 grow, 1 = whole, 2 = half.
LUF_DEV void sf_grow(const SfView& g, const SfShot& s, int mode, int lane, int nl)
{
    sf_for_awake(g, s, lane, nl, [&](int r, int q, int v) {
        uint32_t f = s.fl[v];
        if (mode == 2) {
            f &= ~(uint32_t)(SF_UNROOTED | SF_N_UNROOTED);
            if ((f & SF_ACTIVE) && !(f & SF_WHOLE) && !(f & SF_GROWN)) { sf_grow_node(g, s, r, q); f ^= SF_WHOLE; }
        } else {
            if ((f & SF_ACTIVE) && (mode == 0 || (f & SF_WHOLE))) {
                sf_grow_node(g, s, r, q); f ^= SF_WHOLE;
                if (mode == 1) f |= SF_GROWN | SF_N_GROWN;
            }
            if (r == g.h - 1) {                                   // bottom sheet: a downward pointer is broken
                const uint32_t p = sf_ptr(s, v);
                if (p != SF_PTR_C && ((LUF_LDG(&g.nbr[q * g.K + p]) >> 23) & 3u) == 0u) {
                    s.cid[v] = (uint16_t)SF_RESET; sf_or_acp(s, v, SF_PTR_C << 12);     // C = all ones
                }
            }
        }
        s.fl[v] = (uint16_t)f;
    });
}

// _Node.flooding: _FullUnrooter.flooding_whole (:1394-1409) / flooding_half (:1359-1362)
LUF_DEV void sf_flooding(const SfView& g, const SfShot& s, int r, int q, int v, bool whole, uint32_t& f, uint32_t& ptr)
{
    const uint32_t cid = s.cid[v];
    uint32_t ncid = s.ncid[v];
    if (whole && cid == SF_RESET) {                                   // finish unrooting
        f |= SF_BUSY | SF_N_UNROOTED;
        s.ncid[v] = (uint16_t)sf_id(g, r, q);
        return;
    }
    LUF_NOUNROLL_CA
    for (uint32_t slots = sf_access(s, v); slots; slots &= slots - 1u) {     // ascending slot = the scan order
        const int k = luf_ffs(slots) - 1;
        SfNbr nb;
        if (!sf_nbr(g, r, q, k, nb)) continue;                        // the far end left the window
        const uint32_t c = s.cid[nb.node];
        if (whole && c == SF_RESET) {
            if (!(f & SF_UNROOTED)) { f |= SF_BUSY; ncid = SF_RESET; ptr = SF_PTR_C; break; }
        } else if (ncid == SF_RESET ? false : c < ncid) {             // _compare_cid, :1364-1369
            f |= SF_BUSY; ptr = (uint32_t)k; ncid = c;
        }
        if (whole && (s.fl[nb.node] & SF_GROWN) && !(f & SF_N_GROWN)) f |= SF_BUSY | SF_N_GROWN;   // _check_grown
    }
    s.ncid[v] = (uint16_t)ncid;
}

// _Node.syncing, snowflake/main.py:1106-1122
LUF_DEV void sf_syncing(const SfView& g, const SfShot& s, int r, int q, int v, uint32_t& f, uint32_t ptr)
{
    (void)v;
    const bool detector_defect = q >= g.NLb && (f & SF_DEFECT);
    uint32_t nact;
    SfNbr nb;
    if (ptr == SF_PTR_C || !sf_nbr(g, r, q, (int)ptr, nb)) nact = detector_defect ? 1u : 0u;
    else {
        nact = (s.fl[nb.node] & SF_ACTIVE) ? 1u : 0u;
        if (detector_defect) {                                        // push the defect along the pointer
            f |= SF_BUSY;
            luf_atomic_xor(&s.ndef[nb.r2 * g.WLn + (nb.q2 >> 5)], 1u << (nb.q2 & 31));
            luf_atomic_xor(&s.ndef[r * g.WLn + (q >> 5)], 1u << (q & 31));
            luf_atomic_xor(&s.corr[nb.e >> 5], 1u << (nb.e & 31u));
        }
    }
    f = (f & ~(uint32_t)SF_N_ACTIVE) | (nact ? SF_N_ACTIVE : 0);
    if (((f & SF_ACTIVE) != 0) != (nact != 0)) f |= SF_BUSY;
}

// One merging timestep, phase A: _Node.merging (:1097-1104) with the fast or slow merger (:1237-1256)
LUF_DEV void sf_merge_a(const SfView& g, const SfShot& s, bool whole, bool slow, int lane, int nl)
{
    sf_for_awake(g, s, lane, nl, [&](int r, int q, int v) {
        uint32_t f = s.fl[v] & ~(uint32_t)SF_BUSY, ptr = sf_ptr(s, v);
        if (slow) { sf_syncing(g, s, r, q, v, f, ptr); sf_flooding(g, s, r, q, v, whole, f, ptr); }
        else { sf_flooding(g, s, r, q, v, whole, f, ptr); sf_syncing(g, s, r, q, v, f, ptr); }
        s.fl[v] = (uint16_t)f; sf_set_ptr(s, v, ptr);
    });
}

// phase B: update_after_merging (:1131-1137).  Returns bit0 any busy, bit1 any cid == RESET.
LUF_DEV uint32_t sf_merge_b(const SfView& g, const SfShot& s, int lane, int nl)
{
    uint32_t acc = 0;
    sf_for_awake(g, s, lane, nl, [&](int r, int q, int v) {
        uint32_t f = s.fl[v];
        const uint32_t c = s.ncid[v];
        s.cid[v] = (uint16_t)c;
        const uint32_t nd = (s.ndef[r * g.WLn + (q >> 5)] >> (q & 31)) & 1u;
        f = (f & ~(uint32_t)(SF_DEFECT | SF_ACTIVE | SF_UNROOTED | SF_GROWN))
            | (nd ? SF_DEFECT : 0) | ((f & SF_N_ACTIVE) ? SF_ACTIVE : 0)
            | ((f & SF_N_UNROOTED) ? SF_UNROOTED : 0) | ((f & SF_N_GROWN) ? SF_GROWN : 0);
        s.fl[v] = (uint16_t)f;
        if (f & SF_BUSY) acc |= 1u;
        if (c == SF_RESET) acc |= 2u;
    });
    return acc;
}

}  // namespace luf

// ------------------------------------------------------------------ drivers
// The cycle / stream drivers are written against a tile (LUF_ST_PHASE / LUF_ST_ANY) and compiled
// once per tile shape: luf::sf_warp (one warp per window) and luf::sf_cta (one thread block per
// window: large windows, whose state leaves room for only a few per SM and whose sweeps visit
// hundreds of awake nodes).
#define LUF_SF_NS sf_warp
#define LUF_ST_PHASE(stmt) LUF_PHASE(stmt)
#define LUF_ST_ANY(x) LUF_ANY(x)
#include "luf_sf_drivers.inl"
#undef LUF_SF_NS
#undef LUF_ST_PHASE
#undef LUF_ST_ANY
namespace luf { using namespace sf_warp; }

#define LUF_SF_NS sf_cta
#if defined(LUF_HOST_EMU)
  #define LUF_ST_PHASE(stmt) LUF_PHASE(stmt)
  #define LUF_ST_ANY(x) LUF_ANY(x)
#else
  #define LUF_ST_PHASE(stmt) { stmt; } __syncthreads();
  #define LUF_ST_ANY(x) __syncthreads_or((x) != 0)
  #define LUF_ST_BLOCK_VOTES 1
#endif
#include "luf_sf_drivers.inl"
#undef LUF_ST_BLOCK_VOTES
#undef LUF_SF_NS
#undef LUF_ST_PHASE
#undef LUF_ST_ANY
