// luf_ca.cuh -- Macar / Actis: the local Union-Find cellular automata
// (reference: localuf/decoders/luf/main.py) as synchronous sweeps of one warp
// over one shot's lattice held in shared memory.
//
// One reference timestep (`LUF._advance`, luf/main.py:182-188) is two phases:
//   A  every node runs the action of the current stage, reading only committed
//      state of its neighbours and writing its own `next_*` fields; effects on
//      *other* nodes (a relayed anyon, a peeled defect, Actis signals) go to
//      receive-bitmaps with shared-memory atomics;
//   B  every node commits (`update_unphysicals`, luf/main.py:425-432,569-581),
//      the warp votes busy/active, and the (warp-uniform) controller advances
//      (luf/main.py:354-378).
// Within-timestep updates of the reference are order independent (SURVEY.md
// section 7, hard part 3), so the parallel sweep reproduces its timestep counts,
// erasure and correction exactly.  Same dual-compilation scheme as
// luf_core.cuh (LUF_HOST_EMU = lanes run sequentially; test harness only).
#pragma once
#include "luf_core.cuh"

namespace luf {

enum { ST_GROWING = 0, ST_MERGING = 1, ST_PRESYNCING = 2, ST_SYNCING = 3, ST_BURNING = 4, ST_PEELING = 5, ST_DONE = 6 };
enum { FL_DEFECT = 1, FL_ACTIVE = 2, FL_ANYON = 4, FL_BUSY = 8, FL_SEEN_ANYON = 16, FL_SEEN_DEFECT = 32 };
// Actis byte: stage (bits 0-1), next_stage (2-3), busy_signal (4), active_signal (5), own next_active_signal (6)
enum { AS_BUSY_SIG = 16, AS_ACTIVE_SIG = 32, AS_OWN_NAS = 64 };
static const uint32_t PTR_C = 0xFu;          // pointer 'C' (no neighbour) in the 4-bit pointer field
static const uint32_t NBR_NONE = 0xFFFFFFFFu;
enum { FLAG_ACTIS_UNOPTIMAL = 4 };
#ifndef LUF_CA_TLIST_CAP
  #define LUF_CA_TLIST_CAP 256
#endif
static const uint32_t CA_TLIST_CAP = LUF_CA_TLIST_CAP;    // warp tiles; tests shrink it to force the bitmap path
#ifndef LUF_CA_CTA_TLIST_CAP
  #define LUF_CA_CTA_TLIST_CAP 2048
#endif
static const uint32_t CA_CTA_TLIST_CAP = LUF_CA_CTA_TLIST_CAP;   // block tiles

struct CaView {
    int N, B, E, K;             // K directions; slots >= K/2 are the edges a node owns (it is their first endpoint)
    int WE, WD, WN, GW;
    const uint32_t* nbr;        // [N*K] edge | node<<16 in scan order (luf/main.py:748-777), NBR_NONE if absent
    const uint32_t* west_mask;  // [WE]
    const uint8_t* span;        // [N]   Actis per-node countdown start (luf/main.py:970-976)
    const uint16_t* relay;      // [N]   Actis neighbour towards node 0 (luf/main.py:1138-1163); 0xFFFF = controller
    int global_span;            // luf/main.py:517-523
    int nbr_shared;             // nbr points at a copy in shared memory (k_mc_ca_tiles): plain loads instead of ld.global.nc
};

struct CaLayout {
    uint32_t cid, ncid;         // u16[N]
    uint32_t acp;               // u16[N]  bits 0..11: the slots whose edge was fully grown at the last update_access
                                //         (luf/main.py:434-440; what the node may look through); bits 12..15: pointer
    uint32_t fl, ast, cnt;      // u8[N]   (ast, cnt only for Actis)
    uint32_t growth;            // u32[GW] live growth
    uint32_t rx_anyon, rx_defect;   // u32[WN] cumulative toggles received from neighbours
    uint32_t rx_busy, rx_active;    // u32[2*WN] Actis signals received, double-buffered by timestep parity
    uint32_t defect, err, corr; // u32[WD] / u32[WE] / u32[WE]
    uint32_t touched;           // u32[WN]  Macar: nodes that are a defect or touch a fully grown edge (the only ones that can act)
    uint32_t glob;              // u32[4]  Actis collection-level next_busy_signal / next_active_signal; Macar: [2] = fill of tlist
    uint32_t tlist_cap;         // entries of tlist
    uint32_t tlist;             // u16[tlist_cap]  Macar: the touched nodes, compacted (rebuilt after every growth step)
    uint32_t bytes;
};

struct CaShot {
    uint16_t *cid, *ncid;
    uint16_t* acp;
    uint8_t *fl, *ast, *cnt;
    uint32_t *growth, *rx_anyon, *rx_defect, *rx_busy, *rx_active, *defect, *err, *corr, *glob, *touched;
    uint16_t* tlist;
    uint32_t tcap;              // capacity of tlist
};

LUF_DEV CaShot ca_bind(const CaLayout& l, unsigned char* b)
{
    CaShot s;
    s.cid = (uint16_t*)(b + l.cid); s.ncid = (uint16_t*)(b + l.ncid);
    s.acp = (uint16_t*)(b + l.acp); s.fl = b + l.fl; s.ast = b + l.ast; s.cnt = b + l.cnt;
    s.growth = (uint32_t*)(b + l.growth);
    s.rx_anyon = (uint32_t*)(b + l.rx_anyon); s.rx_defect = (uint32_t*)(b + l.rx_defect);
    s.rx_busy = (uint32_t*)(b + l.rx_busy); s.rx_active = (uint32_t*)(b + l.rx_active);
    s.defect = (uint32_t*)(b + l.defect);
    s.err = l.err ? (uint32_t*)(b + l.err) : 0;
    s.corr = (uint32_t*)(b + l.corr);
    s.glob = (uint32_t*)(b + l.glob);
    s.touched = (uint32_t*)(b + l.touched);
    s.tlist = (uint16_t*)(b + l.tlist);
    s.tcap = l.tlist_cap;
    return s;
}

// Warp-uniform controller + (Actis) collection state; lives in registers.
struct CaCtl {
    int stage;                  // Controller.stage, luf/main.py:350-352
    int step;                   // timestep parity for the double-buffered signal bitmaps
    // ActisNodes, luf/main.py:545-553
    int n_busy, n_valid, n_countdown, n_busy_signal, n_active_signal, n_next_active_signal;
};
static const int CA_MAX_STEPS = 1 << 20;    // safety net: a shot that runs this long is reported with steps = -1

LUF_DEV uint32_t ca_nbr(const CaView& g, int idx) { return g.nbr_shared ? g.nbr[idx] : LUF_LDG(&g.nbr[idx]); }
LUF_DEV uint32_t g2(const uint32_t* a, uint32_t e) { return (a[e >> 4] >> ((e & 15u) * 2u)) & 3u; }
// access mask | pointer << 12; written by the node's own lane only
LUF_DEV uint32_t ca_ptr(const CaShot& s, int v) { return (uint32_t)s.acp[v] >> 12; }
LUF_DEV uint32_t ca_access(const CaShot& s, int v) { return (uint32_t)s.acp[v] & 0xFFFu; }

// LUF.reset / _Node.reset / ActisNode.reset (luf/main.py:173-180,799-809,1004-1012)
template <bool ACTIS>
LUF_DEV void ca_reset(const CaView& g, const CaShot& s, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) {
        s.cid[v] = (uint16_t)v; s.ncid[v] = (uint16_t)v;          // dense ids: node index == index_to_id
        s.acp[v] = (uint16_t)(PTR_C << 12); s.fl[v] = 0;
        if (ACTIS) { s.ast[v] = 0; s.cnt[v] = 0; }
    }
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.GW; w += nl) s.growth[w] = 0;
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WN; w += nl) {
        s.rx_anyon[w] = 0; s.rx_defect[w] = 0; s.touched[w] = 0;
        if (ACTIS) { s.rx_busy[w] = 0; s.rx_busy[g.WN + w] = 0; s.rx_active[w] = 0; s.rx_active[g.WN + w] = 0; }
    }
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WE; w += nl) { s.corr[w] = 0; if (s.err) s.err[w] = 0; }
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WD; w += nl) s.defect[w] = 0;
    if (lane < 4) s.glob[lane] = 0;
}

// Nodes.load / make_defect (luf/main.py:405-408,811-815)
LUF_DEV void ca_load(const CaView& g, const CaShot& s, int lane, int nl)
{
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WD; w += nl) {
        uint32_t bits = s.defect[w];
        LUF_NOUNROLL_CA
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            const uint32_t v = (uint32_t)(g.B + 32 * w + b);
            s.fl[v] = (uint8_t)(FL_DEFECT | FL_ACTIVE | FL_ANYON);
            luf_atomic_or(&s.touched[v >> 5], 1u << (v & 31u));
        }
    }
}

// _Node.growing, luf/main.py:817-826 (saturating, like the in-place loop of the reference)
LUF_DEV void ca_growing(const CaView& g, const CaShot& s, int v)
{
    if (!(s.fl[v] & FL_ACTIVE)) return;
    LUF_NOUNROLL_CA
    for (int k = 0; k < g.K; ++k) {
        const uint32_t ent = ca_nbr(g, v * g.K + k);
        if (ent == NBR_NONE) continue;
        const uint32_t e = ent & 0xFFFFu, sh = (e & 15u) * 2u;
        if (((s.growth[e >> 4] >> sh) & 3u) >= G_FULL) continue;
        const uint32_t old = (luf_atomic_add(&s.growth[e >> 4], 1u << sh) >> sh) & 3u;
        if (old >= G_FULL) luf_atomic_sub(&s.growth[e >> 4], 1u << sh);
        else if (old == G_HALF) {                                   // newly FULL: both ends can act from now on
            const uint32_t o = ent >> 16;
            luf_atomic_or(&s.touched[v >> 5], 1u << (v & 31));
            luf_atomic_or(&s.touched[o >> 5], 1u << (o & 31u));
        }
    }
}

// Macar only ever has work at touched nodes: an untouched node has no defect,
// no anyon, is inactive and sees no fully grown edge, so every stage action and
// every commit leaves it unchanged and idle.  (Actis nodes also relay stages
// and signals, so Actis sweeps all nodes.)
template <bool SPARSE, class F>
LUF_DEV void ca_for_nodes(const CaView& g, const CaShot& s, int lane, int nl, F&& f)
{
    if (SPARSE) {
        const uint32_t total = s.glob[2];
        if (total <= s.tcap) {                      // balanced: one touched node per lane and trip
            LUF_NOUNROLL_CA
            for (uint32_t i = (uint32_t)lane; i < total; i += (uint32_t)nl) f((int)s.tlist[i]);
        } else
            LUF_NOUNROLL_CA
            for (int w = lane; w < g.WN; w += nl) {
                uint32_t bits = s.touched[w];
                LUF_NOUNROLL_CA
                while (bits) {
                    const int b = luf_ffs(bits) - 1; bits &= bits - 1;
                    f(32 * w + b);
                }
            }
    } else
        LUF_NOUNROLL_CA
        for (int v = lane; v < g.N; v += nl) f(v);
}

// Compact the touched bitmap into tlist (glob[2] = number of touched nodes; 0 on entry).
LUF_DEV void ca_gather_touched(const CaView& g, const CaShot& s, int lane, int nl)
{
    const int wpl = (g.WN + nl - 1) / nl;
    const int w0 = lane * wpl, w1 = w0 + wpl < g.WN ? w0 + wpl : g.WN;
    uint32_t c = 0;
    LUF_NOUNROLL_CA
    for (int w = w0; w < w1; ++w) c += (uint32_t)luf_popc(s.touched[w]);
    if (!c) return;
    uint32_t pos = luf_atomic_add(&s.glob[2], c);
    if (pos + c > s.tcap) return;
    LUF_NOUNROLL_CA
    for (int w = w0; w < w1; ++w) {
        uint32_t bits = s.touched[w];
        LUF_NOUNROLL_CA
        while (bits) {
            const int b = luf_ffs(bits) - 1; bits &= bits - 1;
            s.tlist[pos++] = (uint16_t)(32 * w + b);
        }
    }
}

// _Node.merging, luf/main.py:839-854.  Returns busy.
LUF_DEV uint32_t ca_merging(const CaView& g, const CaShot& s, int v)
{
    const uint32_t access = ca_access(s, v);
    uint32_t fl = s.fl[v] & ~(uint32_t)FL_BUSY, ptr = ca_ptr(s, v), busy = 0;
    if (v >= g.B && ptr != PTR_C && (fl & FL_ANYON)) {             // relay the anyon along the pointer
        busy = 1;
        const uint32_t to = ca_nbr(g, v * g.K + ptr) >> 16;
        luf_atomic_xor(&s.rx_anyon[to >> 5], 1u << (to & 31u));
        fl &= ~(uint32_t)FL_ANYON;
    }
    uint32_t best = s.ncid[v];
    LUF_NOUNROLL_CA
    for (uint32_t slots = access; slots; slots &= slots - 1u) {      // ascending slot = the scan order
        const int k = luf_ffs(slots) - 1;
        const uint32_t c = s.cid[ca_nbr(g, v * g.K + k) >> 16];
        if (c < best) { best = c; ptr = (uint32_t)k; busy = 1; }
    }
    s.ncid[v] = (uint16_t)best; s.acp[v] = (uint16_t)(access | (ptr << 12));
    s.fl[v] = (uint8_t)(fl | (busy ? FL_BUSY : 0));
    return busy;
}

// _Node.presyncing, luf/main.py:856-866.  Returns the new `active`.
LUF_DEV uint32_t ca_presyncing(const CaView& g, const CaShot& s, int v)
{
    const uint32_t fl = s.fl[v];
    const uint32_t act = (v >= g.B && ca_ptr(s, v) == PTR_C && (fl & FL_ANYON)) ? 1u : 0u;
    s.fl[v] = (uint8_t)((fl & ~(uint32_t)FL_ACTIVE) | (act ? FL_ACTIVE : 0));
    return act;
}

// _Node.syncing, luf/main.py:880-884.  Returns busy.
LUF_DEV uint32_t ca_syncing(const CaView& g, const CaShot& s, int v)
{
    const uint32_t fl = s.fl[v] & ~(uint32_t)FL_BUSY;
    uint32_t busy = 0;
    if (!(fl & FL_ACTIVE))
        LUF_NOUNROLL_CA
        for (uint32_t slots = ca_access(s, v); slots; slots &= slots - 1u) {
            const uint32_t ent = ca_nbr(g, v * g.K + luf_ffs(slots) - 1);
            if (s.fl[ent >> 16] & FL_ACTIVE) { busy = 1; break; }
        }
    s.fl[v] = (uint8_t)(fl | (busy ? FL_BUSY : 0));
    return busy;
}

// _Node.burning, luf/main.py:890-899; OPPOSITE[slot k] = slot K-1-k
LUF_DEV void ca_burning(const CaView& g, const CaShot& s, int v)
{
    const uint32_t ptr = ca_ptr(s, v);
    LUF_NOUNROLL_CA
    for (int k = g.K >> 1; k < g.K; ++k) {                          // owned edges
        const uint32_t ent = ca_nbr(g, v * g.K + k);
        if (ent == NBR_NONE) continue;
        const uint32_t e = ent & 0xFFFFu;
        if (g2(s.growth, e) == G_FULL && ptr != (uint32_t)k && ca_ptr(s, (int)(ent >> 16)) != (uint32_t)(g.K - 1 - k))
            luf_atomic_and(&s.growth[e >> 4], ~(3u << ((e & 15u) * 2u)));
    }
}

// _Node.peeling, luf/main.py:901-916.  Returns busy | west-parity << 1.
LUF_DEV uint32_t ca_peeling(const CaView& g, const CaShot& s, int v)
{
    uint32_t fl = s.fl[v] & ~(uint32_t)FL_BUSY;
    uint32_t out = 0;
    if (ca_ptr(s, v) != PTR_C) {
        const uint32_t access = ca_access(s, v);
        if (access && !(access & (access - 1u))) {                  // exactly one way out: a leaf
            const uint32_t only = ca_nbr(g, v * g.K + luf_ffs(access) - 1);
            const uint32_t e = only & 0xFFFFu, to = only >> 16;
            fl |= FL_BUSY; out = 1;
            luf_atomic_and(&s.growth[e >> 4], ~(3u << ((e & 15u) * 2u)));
            if (fl & FL_DEFECT) {
                fl &= ~(uint32_t)FL_DEFECT;
                luf_atomic_xor(&s.corr[e >> 5], 1u << (e & 31u));
                out |= ((LUF_LDG(&g.west_mask[e >> 5]) >> (e & 31u)) & 1u) << 1;
                luf_atomic_xor(&s.rx_defect[to >> 5], 1u << (to & 31u));
            }
        }
    }
    s.fl[v] = (uint8_t)fl;
    return out;
}

// Controller.advance, luf/main.py:354-378.  Returns growth_changed.
LUF_DEV int ca_controller_advance(CaCtl& c, int nodes_busy, int nodes_valid)
{
    int growth_changed = c.stage == ST_PEELING;
    if (!nodes_busy) {
        if (c.stage < 4) {
            if (c.stage == ST_SYNCING && nodes_valid) c.stage = ST_BURNING;
            else {
                if (c.stage == ST_GROWING) growth_changed = 1;
                c.stage = (c.stage + 1) & 3;
            }
        } else {
            growth_changed = 1;
            c.stage += 1;
        }
    }
    return growth_changed;
}

// ---------------------------------------------------------------- Macar step
// Phase A: MacarNode.advance, luf/main.py:929-942.  Returns west parity of this lane's peels.
LUF_DEV uint32_t macar_phase_a(const CaView& g, const CaShot& s, int stage, int lane, int nl)
{
    uint32_t parity = 0;
    ca_for_nodes<true>(g, s, lane, nl, [&](int v) {
        switch (stage) {
        case ST_GROWING: ca_growing(g, s, v); break;
        case ST_MERGING: ca_merging(g, s, v); break;
        case ST_PRESYNCING: ca_presyncing(g, s, v); break;
        case ST_SYNCING: ca_syncing(g, s, v); break;
        case ST_BURNING: ca_burning(g, s, v); break;
        case ST_PEELING: parity ^= ca_peeling(g, s, v) >> 1; break;
        }
    });
    return parity;
}

// Commit of one node: update_after_merge_step / update_after_sync_step
// (luf/main.py:872-878,886-888) and the toggles received from neighbours.
LUF_DEV uint32_t ca_commit_node(const CaShot& s, int stage, int v)
{
    uint32_t fl = s.fl[v];
    if (stage == ST_MERGING) {
        s.cid[v] = s.ncid[v];
        const uint32_t rx = (s.rx_anyon[v >> 5] >> (v & 31)) & 1u;
        if (rx != ((fl >> 4) & 1u)) fl ^= FL_ANYON | FL_SEEN_ANYON;
    } else if (stage == ST_SYNCING) {
        if (fl & FL_BUSY) fl |= FL_ACTIVE;
    } else if (stage == ST_PEELING) {
        const uint32_t rx = (s.rx_defect[v >> 5] >> (v & 31)) & 1u;
        if (rx != ((fl >> 5) & 1u)) fl ^= FL_DEFECT | FL_SEEN_DEFECT;
    }
    s.fl[v] = (uint8_t)fl;
    return fl;
}

// Phase B.  Returns bit0 any busy, bit1 any active (of this lane's nodes).
LUF_DEV uint32_t macar_phase_b(const CaView& g, const CaShot& s, int stage, int lane, int nl)
{
    uint32_t acc = 0;
    ca_for_nodes<true>(g, s, lane, nl, [&](int v) {
        const uint32_t fl = ca_commit_node(s, stage, v);
        if (fl & FL_BUSY) acc |= 1u;
        if (fl & FL_ACTIVE) acc |= 2u;
    });
    return acc;
}

// Nodes.update_access, luf/main.py:434-440: every node records which of its edges are fully grown
// now; until the next update that is what it looks through.  (Macar: an untouched node has none.)
template <bool ACTIS>
LUF_DEV void ca_snapshot_access(const CaView& g, const CaShot& s, int lane, int nl)
{
    ca_for_nodes<!ACTIS>(g, s, lane, nl, [&](int v) {
        uint32_t access = 0;
        LUF_NOUNROLL_CA
        for (int k = 0; k < g.K; ++k) {
            const uint32_t ent = ca_nbr(g, v * g.K + k);
            if (ent != NBR_NONE && g2(s.growth, ent & 0xFFFFu) == G_FULL) access |= 1u << k;
        }
        s.acp[v] = (uint16_t)((s.acp[v] & 0xF000u) | access);
    });
}

// ---------------------------------------------------------------- Actis step
// ActisNode.advance (luf/main.py:1014-1049) + Friendship (luf/main.py:1090-1176)
LUF_DEV void actis_phase_a(const CaView& g, const CaShot& s, const CaCtl& c, int lane, int nl)
{
    const int cur = c.step & 1;
    uint32_t* rxb = s.rx_busy + (cur ? g.WN : 0);
    uint32_t* rxa = s.rx_active + (cur ? g.WN : 0);
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) {
        uint32_t a = s.ast[v];
        const uint32_t st = a & 3u;
        if (LUF_LDG(&g.relay[v]) == 0xFFFFu) { s.glob[0] = 0; s.glob[1] = 0; }   // this timestep's collection inbox
        if (st == ST_GROWING || st == ST_PRESYNCING) {               // advance_definite
            const uint32_t cd = s.cnt[v];
            if (cd == 0) {
                if (st == ST_GROWING) ca_growing(g, s, v);
                else if (ca_presyncing(g, s, v)) a |= AS_OWN_NAS;
                a = (a & ~12u) | ((st + 1u) << 2);
            } else s.cnt[v] = (uint8_t)(cd - 1);
        } else {                                                     // advance_indefinite
            const uint32_t relay = LUF_LDG(&g.relay[v]);
            const uint32_t rs = relay == 0xFFFFu ? (uint32_t)c.stage : (s.ast[relay] & 3u);
            if ((st == ST_MERGING && rs == ST_PRESYNCING) || (st == ST_SYNCING && rs == ST_GROWING)) {
                a = (a & ~12u) | (rs << 2);
                s.cnt[v] = LUF_LDG(&g.span[v]);
            } else {
                const uint32_t busy = st == ST_MERGING ? ca_merging(g, s, v) : ca_syncing(g, s, v);
                if (relay == 0xFFFFu) {                               // ControllerFriendship.relay_signals
                    s.glob[0] = (a >> 4) & 1u;
                    if (a & AS_ACTIVE_SIG) s.glob[1] = 1;
                } else {                                              // NodeFriendship.relay_signals
                    if (busy) a |= AS_BUSY_SIG;
                    if (a & AS_BUSY_SIG) luf_atomic_or(&rxb[relay >> 5], 1u << (relay & 31u));
                    if (a & AS_ACTIVE_SIG) luf_atomic_or(&rxa[relay >> 5], 1u << (relay & 31u));
                }
            }
        }
        s.ast[v] = (uint8_t)a;
    }
}

// Nodes.update_unphysicals + update_unphysicals_for_actis (luf/main.py:569-581,1051-1059)
LUF_DEV void actis_phase_b(const CaView& g, const CaShot& s, const CaCtl& c, int lane, int nl)
{
    const int cur = c.step & 1;
    uint32_t* rxb = s.rx_busy + (cur ? g.WN : 0);
    uint32_t* rxa = s.rx_active + (cur ? g.WN : 0);
    uint32_t* nxb = s.rx_busy + (cur ? 0 : g.WN);
    uint32_t* nxa = s.rx_active + (cur ? 0 : g.WN);
    LUF_NOUNROLL_CA
    for (int v = lane; v < g.N; v += nl) {
        ca_commit_node(s, c.stage, v);
        const uint32_t old = s.ast[v];
        const uint32_t nst = (old >> 2) & 3u;
        uint32_t a = nst | (nst << 2);
        if ((rxb[v >> 5] >> (v & 31)) & 1u) a |= AS_BUSY_SIG;
        if (((rxa[v >> 5] >> (v & 31)) & 1u) || (old & AS_OWN_NAS)) a |= AS_ACTIVE_SIG;
        s.ast[v] = (uint8_t)a;
    }
    LUF_NOUNROLL_CA
    for (int w = lane; w < g.WN; w += nl) { nxb[w] = 0; nxa[w] = 0; }   // buffers of the next timestep
}

// Waiter.advance, luf/main.py:613-679
LUF_DEV void actis_waiter_advance(const CaView& g, CaCtl& c, int flags)
{
    const int optimal = !(flags & FLAG_ACTIS_UNOPTIMAL);
    c.n_busy = 1;
    if (c.n_countdown == 0) {
        c.n_busy = 0;
        c.n_countdown = g.global_span;
        if (c.stage == ST_GROWING) c.n_countdown += 1;
        else if (c.stage == ST_SYNCING) {                            // ActisNodes.update_valid, :559-563
            c.n_valid = !c.n_active_signal; c.n_active_signal = 0; c.n_next_active_signal = 0;
        }
    } else if (optimal ? (c.n_busy_signal && (c.n_countdown == 1 || c.n_countdown == 2)) : c.n_busy_signal)
        c.n_countdown = optimal ? 2 : g.global_span + 1;
    else
        c.n_countdown -= 1;
}

}  // namespace luf

// ------------------------------------------------------------------- driver
// The decode driver is compiled once per tile shape (LUF_CT_PHASE / LUF_CT_ANY): luf::ca_warp (one warp per
// shot) and luf::ca_cta (one thread block per shot: Actis, whose every timestep visits every node, and
// large graphs).
#define LUF_CA_NS ca_warp
#define LUF_CT_PHASE(stmt) LUF_PHASE(stmt)
#define LUF_CT_ANY(x) LUF_ANY(x)
#include "luf_ca_drivers.inl"
#undef LUF_CA_NS
#undef LUF_CT_PHASE
#undef LUF_CT_ANY
namespace luf { using namespace ca_warp; }

#define LUF_CA_NS ca_cta
#if defined(LUF_HOST_EMU)
  #define LUF_CT_PHASE(stmt) LUF_PHASE(stmt)
  #define LUF_CT_ANY(x) LUF_ANY(x)
#else
  #define LUF_CT_PHASE(stmt) { stmt; } __syncthreads();
  #define LUF_CT_ANY(x) __syncthreads_or((x) != 0)
  #define LUF_CT_SYNC() __syncthreads()
  #define LUF_CT_WORD_VOTES 1
#endif
#include "luf_ca_drivers.inl"
#undef LUF_CA_NS
#undef LUF_CT_PHASE
#undef LUF_CT_ANY
#undef LUF_CT_SYNC
#undef LUF_CT_WORD_VOTES

// Tile = nl threads (a multiple of 32) of a block that holds several shots next to one copy of the neighbour
// table (k_mc_ca_tiles): named barrier 1 + threadIdx.x / nl.  Macar only (its votes go through a word of the
// shot's state, LUF_CT_WORD_VOTES); device only.
#if !defined(LUF_HOST_EMU)
#define LUF_CA_NS ca_tile
#define LUF_CT_SYNC() asm volatile("bar.sync %0, %1;" ::"r"(1u + threadIdx.x / (unsigned)nl), "r"(nl) : "memory")
#define LUF_CT_PHASE(stmt) { stmt; } LUF_CT_SYNC();
#define LUF_CT_ANY(x) ((x) != 0)          /* (only on the Actis path, which this tile does not run) */
#define LUF_CT_WORD_VOTES 1
#include "luf_ca_drivers.inl"
#undef LUF_CA_NS
#undef LUF_CT_PHASE
#undef LUF_CT_ANY
#undef LUF_CT_SYNC
#undef LUF_CT_WORD_VOTES
#endif
