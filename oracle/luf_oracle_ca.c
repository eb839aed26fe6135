/*
 * luf_oracle_ca.c -- CPU restatement of the reference's local Union-Find
 * cellular automata: Macar and Actis (localuf/decoders/luf/main.py).
 * TEST INFRASTRUCTURE ONLY; see luf_oracle.h.  Literal, sequential emulation
 * of one `_advance` per timestep; citations are to /root/reference.
 */
#include "luf_oracle.h"

#include <stdlib.h>
#include <string.h>

#define WORDS(n) (((n) + 31) / 32)
#define GETBIT(a, k) (((a)[(k) >> 5] >> ((k) & 31)) & 1u)
#define FLIPBIT(a, k) ((a)[(k) >> 5] ^= (1u << ((k) & 31)))

enum { UNGROWN = 0, HALF = 1, FULL = 2 };
/* decoders/luf/constants.py:17-41 */
enum { GROWING = 0, MERGING = 1, PRESYNCING = 2, SYNCING = 3, BURNING = 4, PEELING = 5, DONE = 6 };
#define PTR_C (-1)

typedef struct {
    const oracle_graph *g;
    int N, K;
    /* _Node state, luf/main.py:799-809 */
    uint8_t *defect, *active, *anyon, *next_anyon, *busy;
    int32_t *cid, *next_cid;
    int8_t *pointer;        /* direction slot, PTR_C = 'C' */
    uint16_t *access;       /* bit k: neighbour in slot k reachable over a FULL edge (luf/main.py:828-837) */
    uint8_t *growth;
    uint32_t *corr;
    /* ActisNode state, luf/main.py:1004-1012 */
    int32_t *countdown;
    uint8_t *stage, *next_stage, *busy_signal, *next_busy_signal, *active_signal, *next_active_signal;
    /* Controller (luf/main.py:350-352) and ActisNodes (luf/main.py:545-553) */
    int c_stage;
    int actis, optimal;
    int n_busy, n_valid, n_countdown, n_busy_signal, n_next_busy_signal, n_active_signal, n_next_active_signal;
} ca;

static ca *ca_new(const oracle_graph *g)
{
    ca *c = calloc(1, sizeof(ca));
    const int N = g->n_nodes;
    c->g = g; c->N = N; c->K = g->n_dirs;
    c->defect = malloc(N); c->active = malloc(N); c->anyon = malloc(N); c->next_anyon = malloc(N); c->busy = malloc(N);
    c->cid = malloc(4 * N); c->next_cid = malloc(4 * N); c->pointer = malloc(N); c->access = malloc(2 * N);
    c->growth = malloc(g->n_edges); c->corr = malloc(4 * WORDS(g->n_edges));
    c->countdown = malloc(4 * N);
    c->stage = malloc(N); c->next_stage = malloc(N); c->busy_signal = malloc(N); c->next_busy_signal = malloc(N);
    c->active_signal = malloc(N); c->next_active_signal = malloc(N);
    return c;
}

static void ca_free(ca *c)
{
    free(c->defect); free(c->active); free(c->anyon); free(c->next_anyon); free(c->busy);
    free(c->cid); free(c->next_cid); free(c->pointer); free(c->access); free(c->growth); free(c->corr);
    free(c->countdown); free(c->stage); free(c->next_stage); free(c->busy_signal); free(c->next_busy_signal);
    free(c->active_signal); free(c->next_active_signal); free(c);
}

/* LUF.reset (luf/main.py:173-180), Controller.reset (:350-352), _Node.reset (:799-809),
 * ActisNode.reset (:1004-1012), ActisNodes.reset (:545-553) */
static void ca_reset(ca *c)
{
    const oracle_graph *g = c->g;
    for (int v = 0; v < c->N; ++v) {
        c->defect[v] = c->active[v] = c->anyon[v] = c->next_anyon[v] = c->busy[v] = 0;
        c->cid[v] = c->next_cid[v] = g->node_id[v];
        c->pointer[v] = PTR_C; c->access[v] = 0;
        c->countdown[v] = 0; c->stage[v] = c->next_stage[v] = GROWING;
        c->busy_signal[v] = c->next_busy_signal[v] = c->active_signal[v] = c->next_active_signal[v] = 0;
    }
    memset(c->growth, UNGROWN, g->n_edges);
    memset(c->corr, 0, 4 * WORDS(g->n_edges));
    c->c_stage = GROWING;
    c->n_busy = 0; c->n_valid = 1; c->n_countdown = 0;
    c->n_busy_signal = c->n_next_busy_signal = c->n_active_signal = c->n_next_active_signal = 0;
}

/* _Node.growing, luf/main.py:817-826 */
static void node_growing(ca *c, int v)
{
    const oracle_graph *g = c->g;
    if (!c->active[v]) return;
    for (int k = g->inc_off[v]; k < g->inc_off[v + 1]; ++k) {
        const int e = g->inc_edge[k];
        if (c->growth[e] == UNGROWN || c->growth[e] == HALF) c->growth[e] += 1;
    }
}

/* _Node.update_access, luf/main.py:828-837 */
static void node_update_access(ca *c, int v)
{
    const oracle_graph *g = c->g;
    uint16_t a = 0;
    for (int k = 0; k < c->K; ++k) {
        const int e = g->nbr_edge[v * c->K + k];
        if (e >= 0 && c->growth[e] == FULL) a |= (uint16_t)(1u << k);
    }
    c->access[v] = a;
}

/* _Node.merging, luf/main.py:839-854 */
static void node_merging(ca *c, int v)
{
    const oracle_graph *g = c->g;
    c->busy[v] = 0;
    if (!(g->node_flags[v] & 1) && c->pointer[v] != PTR_C && c->anyon[v]) {
        c->busy[v] = 1;
        c->next_anyon[g->nbr_node[v * c->K + c->pointer[v]]] ^= 1;      /* relay along pointer */
        c->anyon[v] = 0;
    }
    for (int k = 0; k < c->K; ++k) {
        if (!(c->access[v] >> k & 1)) continue;
        const int nb = g->nbr_node[v * c->K + k];
        if (c->cid[nb] < c->next_cid[v]) {
            c->busy[v] = 1; c->pointer[v] = (int8_t)k; c->next_cid[v] = c->cid[nb];
        }
    }
}

/* _Node.presyncing, luf/main.py:856-866 */
static void node_presyncing(ca *c, int v)
{
    c->active[v] = !(c->g->node_flags[v] & 1) && c->pointer[v] == PTR_C && c->anyon[v];
}

/* _Node.syncing, luf/main.py:880-884 */
static void node_syncing(ca *c, int v)
{
    int any = 0;
    for (int k = 0; k < c->K; ++k)
        if ((c->access[v] >> k & 1) && c->active[c->g->nbr_node[v * c->K + k]]) any = 1;
    c->busy[v] = !c->active[v] && any;
}

/* _Node.burning, luf/main.py:890-899; OPPOSITE[slot k] = slot K-1-k (luf/main.py:720-733) */
static void node_burning(ca *c, int v)
{
    const oracle_graph *g = c->g;
    for (int k = 0; k < c->K; ++k) {
        const int e = g->nbr_edge[v * c->K + k];
        if (e < 0 || !g->nbr_owned[v * c->K + k]) continue;
        if (c->growth[e] == FULL && c->pointer[v] != k &&
            c->pointer[g->nbr_node[v * c->K + k]] != c->K - 1 - k)
            c->growth[e] = UNGROWN;
    }
}

/* _Node.peeling, luf/main.py:901-916 */
static void node_peeling(ca *c, int v)
{
    const oracle_graph *g = c->g;
    c->busy[v] = 0;
    if (c->pointer[v] != PTR_C && __builtin_popcount(c->access[v]) == 1) {
        const int k = __builtin_ctz(c->access[v]);
        const int e = g->nbr_edge[v * c->K + k];
        c->busy[v] = 1;
        c->growth[e] = UNGROWN;
        if (c->defect[v]) {
            c->defect[v] = 0;
            FLIPBIT(c->corr, e);
            c->defect[g->nbr_node[v * c->K + k]] ^= 1;
        }
    }
}

/* Controller.advance, luf/main.py:354-378 */
static int controller_advance(ca *c, int nodes_busy, int nodes_valid)
{
    int growth_changed = c->c_stage == PEELING;
    if (!nodes_busy) {
        if (c->c_stage < 4) {
            if (c->c_stage == SYNCING && nodes_valid) c->c_stage = BURNING;
            else {
                if (c->c_stage == GROWING) growth_changed = 1;
                c->c_stage = (c->c_stage + 1) % 4;
            }
        } else {
            growth_changed = 1;
            c->c_stage += 1;
        }
    }
    return growth_changed;
}

/* LUF._advance for Macar, luf/main.py:182-188 with MacarNode.advance (:929-942),
 * Nodes.update_unphysicals (:425-432), MacarNodes.busy/valid (:480-486) */
static void macar_advance(ca *c)
{
    const int N = c->N;
    for (int v = 0; v < N; ++v)
        switch (c->c_stage) {
        case GROWING: node_growing(c, v); break;
        case MERGING: node_merging(c, v); break;
        case PRESYNCING: node_presyncing(c, v); break;
        case SYNCING: node_syncing(c, v); break;
        case BURNING: node_burning(c, v); break;
        case PEELING: node_peeling(c, v); break;
        }
    if (c->c_stage == MERGING)
        for (int v = 0; v < N; ++v) { c->cid[v] = c->next_cid[v]; c->anyon[v] ^= c->next_anyon[v]; c->next_anyon[v] = 0; }
    else if (c->c_stage == SYNCING)
        for (int v = 0; v < N; ++v) c->active[v] |= c->busy[v];
    int busy = 0, active = 0;
    for (int v = 0; v < N; ++v) { busy |= c->busy[v]; active |= c->active[v]; }
    if (controller_advance(c, busy, !active))
        for (int v = 0; v < N; ++v) node_update_access(c, v);
}

/* Friendship.update_stage_helper, luf/main.py:1090-1103 */
static int actis_update_stage(ca *c, int v)
{
    const int relay = c->g->actis_relay[v];
    const int rs = relay < 0 ? c->c_stage : c->stage[relay];
    if ((c->next_stage[v] == MERGING && rs == PRESYNCING) || (c->next_stage[v] == SYNCING && rs == GROWING)) {
        c->next_stage[v] = (uint8_t)rs;
        return 1;
    }
    return 0;
}

/* ControllerFriendship.relay_signals (:1122-1124), NodeFriendship.relay_signals (:1173-1176) */
static void actis_relay_signals(ca *c, int v)
{
    const int relay = c->g->actis_relay[v];
    if (relay < 0) {
        c->n_next_busy_signal = c->busy_signal[v];
        c->n_next_active_signal |= c->active_signal[v];
    } else {
        c->busy_signal[v] |= c->busy[v];
        c->next_busy_signal[relay] |= c->busy_signal[v];
        c->next_active_signal[relay] |= c->active_signal[v];
    }
}

/* ActisNode.advance, luf/main.py:1014-1049 */
static void actis_node_advance(ca *c, int v)
{
    switch (c->stage[v]) {
    case GROWING:
    case PRESYNCING:                     /* advance_definite */
        if (c->countdown[v] == 0) {
            if (c->stage[v] == GROWING) node_growing(c, v);
            else { node_presyncing(c, v); c->next_active_signal[v] = c->active[v]; }
            c->next_stage[v] += 1;
        } else c->countdown[v] -= 1;
        break;
    case MERGING:
    case SYNCING:                        /* advance_indefinite */
        if (actis_update_stage(c, v)) c->countdown[v] = c->g->actis_span[v];
        else {
            if (c->stage[v] == MERGING) node_merging(c, v); else node_syncing(c, v);
            actis_relay_signals(c, v);
        }
        break;
    }
}

/* Waiter.advance (:613-643), OptimalWaiter (:646-662), UnoptimalWaiter (:665-679) */
static void actis_waiter_advance(ca *c)
{
    const int span = c->g->actis_global_span;
    c->n_busy = 1;
    if (c->n_countdown == 0) {
        c->n_busy = 0;
        c->n_countdown = span;
        if (c->c_stage == GROWING) c->n_countdown += 1;
        else if (c->c_stage == SYNCING) {          /* ActisNodes.update_valid, :559-563 */
            c->n_valid = !c->n_active_signal;
            c->n_active_signal = 0; c->n_next_active_signal = 0;
        }
    } else if (c->optimal ? (c->n_busy_signal && (c->n_countdown == 1 || c->n_countdown == 2)) : c->n_busy_signal)
        c->n_countdown = c->optimal ? 2 : span + 1;
    else
        c->n_countdown -= 1;
}

/* LUF._advance for Actis: ActisNodes.advance (:565-567), update_unphysicals (:569-581) */
static void actis_advance(ca *c)
{
    const int N = c->N;
    for (int v = 0; v < N; ++v) actis_node_advance(c, v);
    actis_waiter_advance(c);
    if (c->c_stage == MERGING)
        for (int v = 0; v < N; ++v) { c->cid[v] = c->next_cid[v]; c->anyon[v] ^= c->next_anyon[v]; c->next_anyon[v] = 0; }
    else if (c->c_stage == SYNCING)
        for (int v = 0; v < N; ++v) c->active[v] |= c->busy[v];
    for (int v = 0; v < N; ++v) {        /* ActisNode.update_unphysicals_for_actis, :1051-1059 */
        c->stage[v] = c->next_stage[v];
        c->busy_signal[v] = c->next_busy_signal[v]; c->next_busy_signal[v] = 0;
        c->active_signal[v] = c->next_active_signal[v]; c->next_active_signal[v] = 0;
    }
    c->n_busy_signal = c->n_next_busy_signal; c->n_next_busy_signal = 0;
    c->n_active_signal = c->n_next_active_signal;
    if (controller_advance(c, c->n_busy, c->n_valid))
        for (int v = 0; v < N; ++v) node_update_access(c, v);
}

/* LUF.decode = validate (:112-145) then peel (:147-166) */
static void ca_decode(ca *c, int decoder, int flags, const uint32_t *syn_bits, uint32_t *corr_bits,
                      uint8_t *growth_out, int32_t *steps)
{
    const oracle_graph *g = c->g;
    const int B = g->n_boundary;
    c->actis = decoder == ORACLE_DEC_ACTIS;
    c->optimal = !(flags & 4);
    ca_reset(c);
    /* Nodes.load (:405-408), make_defect (:811-815); ActisNodes.load (:555-557) */
    for (int k = 0; k < g->n_nodes - B; ++k)
        if (GETBIT(syn_bits, k)) { c->defect[B + k] = c->active[B + k] = c->anyon[B + k] = 1; }
    if (c->actis) c->n_valid = 0;
    int tSV = 0, tBP = 0;
    while (c->c_stage != BURNING) { if (c->actis) actis_advance(c); else macar_advance(c); ++tSV; }
    if (growth_out) memcpy(growth_out, c->growth, g->n_edges);
    while (c->c_stage != DONE) { if (c->actis) actis_advance(c); else macar_advance(c); ++tBP; }
    memcpy(corr_bits, c->corr, 4 * WORDS(g->n_edges));
    if (steps) { steps[0] = tSV; steps[1] = tBP; }
}

int oracle_ca_decode_batch(const oracle_graph *g, int decoder, int flags, int64_t nshots,
                           const uint32_t *syn_bits, uint32_t *corr_bits, uint8_t *growth_out, int32_t *steps)
{
    if (decoder != ORACLE_DEC_MACAR && decoder != ORACLE_DEC_ACTIS) return -1;
    if (g->n_dirs > 16) return -1;
    const int WD = WORDS(g->n_nodes - g->n_boundary), WE = WORDS(g->n_edges);
    ca *c = ca_new(g);
    for (int64_t s = 0; s < nshots; ++s)
        ca_decode(c, decoder, flags, syn_bits + s * WD, corr_bits + s * WE,
                  growth_out ? growth_out + s * g->n_edges : NULL, steps ? steps + 2 * s : NULL);
    ca_free(c);
    return 0;
}
