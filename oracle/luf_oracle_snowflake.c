/*
 * luf_oracle_snowflake.c -- CPU restatement of the reference's streaming path:
 * the Snowflake decoder (localuf/decoders/snowflake/main.py) driven by the
 * frugal scheme (localuf/_schemes.py:480-649) with string pairing and logical
 * counting (localuf/_pairs.py).  TEST INFRASTRUCTURE ONLY; see luf_oracle.h.
 * Literal, sequential emulation; citations are to /root/reference.
 */
#include "luf_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define WORDS(n) (((n) + 31) / 32)
#define GETBIT(a, k) (((a)[(k) >> 5] >> ((k) & 31)) & 1u)
#define SETBIT(a, k) ((a)[(k) >> 5] |= (1u << ((k) & 31)))

enum { UNGROWN = 0, HALF = 1, FULL = 2 };
#define RESET (-1)      /* snowflake/constants.py:5 */
#define PTR_C (-1)

typedef struct {
    const oracle_window *w;
    int N, E, NL;
    /* _Node state, snowflake/main.py:1011-1031 */
    uint8_t *active, *whole, *defect, *grown, *unrooted, *busy;
    uint8_t *n_active, *n_whole, *n_defect, *n_grown, *n_unrooted;
    int32_t *cid, *n_cid;
    int8_t *pointer, *n_pointer;
    uint16_t *access;
    /* _Edge state, snowflake/main.py:1489-1496 */
    uint8_t *growth, *corr, *n_growth, *n_corr;
    /* Frugal state, _schemes.py:490-502 */
    uint8_t *error;             /* over all h*EL window edges (incl. future-boundary edges) */
    uint8_t *fb_syndrome;       /* [NLd] defects on the future boundary */
    /* Pairs (_pairs.py:4-66): endpoint -> partner, keyed by (q, t) with t in {0, 1};
     * an endpoint that can no longer be extended only keeps its side. */
    int32_t *partner;           /* [2*NL]: -1 none, >= 0 live key, -2 west, -3 east, -4 dead elsewhere */
} sf;

#define P_NONE (-1)
#define P_WEST (-2)
#define P_EAST (-3)
#define P_DEAD (-4)

static int node_id(const oracle_window *w, int q, int t)
{
    return q < w->NLb ? w->NLb * (w->h - 1 - t) + q : w->NLb * w->h + w->NLd * (w->h - 1 - t) + (q - w->NLb);
}
static int node_layer(const oracle_window *w, int id)
{
    return id < w->NLb * w->h ? w->h - 1 - id / w->NLb : w->h - 1 - (id - w->NLb * w->h) / w->NLd;
}
static int node_pos(const oracle_window *w, int id)
{
    return id < w->NLb * w->h ? id % w->NLb : w->NLb + (id - w->NLb * w->h) % w->NLd;
}

static sf *sf_new(const oracle_window *w)
{
    sf *s = calloc(1, sizeof(sf));
    s->w = w; s->NL = w->NLb + w->NLd; s->N = s->NL * w->h; s->E = w->EL * w->h;
    const int N = s->N, E = s->E;
    uint8_t **nb[] = {&s->active, &s->whole, &s->defect, &s->grown, &s->unrooted, &s->busy,
                      &s->n_active, &s->n_whole, &s->n_defect, &s->n_grown, &s->n_unrooted};
    for (unsigned i = 0; i < sizeof(nb) / sizeof(nb[0]); ++i) *nb[i] = malloc(N);
    s->cid = malloc(4 * N); s->n_cid = malloc(4 * N);
    s->pointer = malloc(N); s->n_pointer = malloc(N); s->access = malloc(2 * N);
    s->growth = malloc(E); s->corr = malloc(E); s->n_growth = malloc(E); s->n_corr = malloc(E);
    s->error = malloc(E); s->fb_syndrome = malloc(w->NLd);
    s->partner = malloc(4 * 2 * s->NL);
    return s;
}

static void sf_free(sf *s)
{
    free(s->active); free(s->whole); free(s->defect); free(s->grown); free(s->unrooted); free(s->busy);
    free(s->n_active); free(s->n_whole); free(s->n_defect); free(s->n_grown); free(s->n_unrooted);
    free(s->cid); free(s->n_cid); free(s->pointer); free(s->n_pointer); free(s->access);
    free(s->growth); free(s->corr); free(s->n_growth); free(s->n_corr); free(s->error); free(s->fb_syndrome);
    free(s->partner); free(s);
}

/* Snowflake.reset (:201-214), _Node.reset (:1011-1031), _Edge.reset (:1489-1496), Frugal.reset (_schemes.py:498-502) */
static void sf_reset(sf *s)
{
    for (int v = 0; v < s->N; ++v) {
        s->active[v] = 0; s->whole[v] = 1; s->cid[v] = v; s->defect[v] = 0; s->pointer[v] = PTR_C;
        s->grown[v] = 0; s->unrooted[v] = 0;
        s->n_active[v] = 0; s->n_whole[v] = 1; s->n_cid[v] = v; s->n_defect[v] = 0; s->n_pointer[v] = PTR_C;
        s->n_grown[v] = 0; s->n_unrooted[v] = 0;
        s->busy[v] = 0; s->access[v] = 0;
    }
    memset(s->growth, UNGROWN, s->E); memset(s->corr, 0, s->E);
    memset(s->n_growth, UNGROWN, s->E); memset(s->n_corr, 0, s->E);
    memset(s->error, 0, s->E); memset(s->fb_syndrome, 0, s->w->NLd);
    for (int k = 0; k < 2 * s->NL; ++k) s->partner[k] = P_NONE;
}

/* neighbour of node v in slot k: returns 0 if absent; else fills node id and layered edge index */
static int nbr(const sf *s, int v, int k, int *node, int *edge)
{
    const oracle_window *w = s->w;
    const int q = node_pos(w, v), t = node_layer(w, v), x = q * w->K + k;
    if (!w->nb_valid[x]) return 0;
    const int t2 = t + w->nb_dt[x];
    if (t2 < 0 || t2 > w->h - 1) return 0;
    *node = node_id(w, w->nb_q[x], t2);
    *edge = (w->h - 1 - (t + w->nb_et[x])) * w->EL + w->nb_el[x];
    return 1;
}

/* ------------------------------------------------------------------ Pairs */
/* endpoint of a floor edge: key of a live endpoint, or its side if it is a space-boundary node */
static int endpoint_key(const sf *s, int q, int t)
{
    if (q < s->w->NLb) return s->w->pos_long[q] == -1 ? P_WEST : P_EAST;
    return t * s->NL + q;
}

static void pairs_add(sf *s, int a, int b, int *count)
{
    /* _pairs.py:39-42, with LogicalCounter.count (:90-111) applied to pairs that can
     * no longer change: west-east counts one logical error, same side is dropped */
    if (a >= 0) s->partner[a] = b;
    if (b >= 0) s->partner[b] = a;
    if (a < 0 && b < 0 && ((a == P_WEST && b == P_EAST) || (a == P_EAST && b == P_WEST))) ++*count;
}

/* Pairs.load, _pairs.py:49-66 */
static void pairs_load(sf *s, int u, int v, int *count)
{
    const int u_in = u >= 0 && s->partner[u] != P_NONE;
    if (u_in) {
        const int wv = s->partner[u];
        s->partner[u] = P_NONE; if (wv >= 0) s->partner[wv] = P_NONE;      /* remove(u) */
        if (v >= 0 && s->partner[v] != P_NONE) {
            const int x = s->partner[v];
            s->partner[v] = P_NONE; if (x >= 0) s->partner[x] = P_NONE;    /* remove(v) */
            pairs_add(s, wv, x, count);
        } else if (v != wv || v < 0)
            pairs_add(s, v, wv, count);
        /* else: the edge closed a loop */
    } else if (v >= 0 && s->partner[v] != P_NONE) {
        const int x = s->partner[v];
        s->partner[v] = P_NONE; if (x >= 0) s->partner[x] = P_NONE;
        pairs_add(s, u, x, count);
    } else
        pairs_add(s, u, v, count);
}

/* LogicalCounter.count's lowering of the kept pairs (_pairs.py:82-111): endpoints on
 * layer 1 move to layer 0, endpoints on layer 0 fall out of reach for good. */
static void pairs_lower(sf *s)
{
    const int NL = s->NL;
    for (int q = 0; q < NL; ++q) {                  /* layer-0 endpoints die */
        const int p = s->partner[q];
        if (p == P_NONE) continue;
        s->partner[q] = P_NONE;
        if (p >= 0) s->partner[p] = P_DEAD;         /* its partner now ends on a dead detector */
    }
    for (int q = 0; q < NL; ++q) {                  /* layer 1 -> layer 0 */
        const int p = s->partner[NL + q];
        s->partner[NL + q] = P_NONE;
        if (p == P_NONE) continue;
        const int p2 = p >= NL ? p - NL : p;        /* a live partner is on layer 1 as well by now */
        s->partner[q] = p2;
    }
}

/* --------------------------------------------------------------- Snowflake */
/* _Node._grow, :1075-1084 */
static void node_grow(sf *s, int v)
{
    for (int k = 0; k < s->w->K; ++k) {
        int nb, e;
        if (!nbr(s, v, k, &nb, &e)) continue;
        if (s->growth[e] == UNGROWN || s->growth[e] == HALF) s->growth[e] += 1;
    }
    s->whole[v] ^= 1;
}

/* _Node.update_access, :1086-1095 */
static void update_access_all(sf *s)
{
    for (int v = 0; v < s->N; ++v) {
        uint16_t a = 0;
        for (int k = 0; k < s->w->K; ++k) {
            int nb, e;
            if (nbr(s, v, k, &nb, &e) && s->growth[e] == FULL) a |= (uint16_t)(1u << k);
        }
        s->access[v] = a;
    }
}

/* _Unrooter._compare_cid (:1364-1369), _check_grown (:1371-1375) */
static void compare_cid(sf *s, int v, int k, int nb)
{
    if (s->cid[nb] < s->n_cid[v]) { s->busy[v] = 1; s->pointer[v] = (int8_t)k; s->n_cid[v] = s->cid[nb]; }
}
static void check_grown(sf *s, int v, int nb)
{
    if (s->grown[nb] && !s->n_grown[v]) { s->busy[v] = 1; s->n_grown[v] = 1; }
}

/* _Node.flooding (:1124-1129) with _FullUnrooter (:1394-1409) / flooding_half (:1359-1362) */
static void node_flooding(sf *s, int v, int whole)
{
    if (whole && s->cid[v] == RESET) {              /* finish unrooting */
        s->busy[v] = 1; s->n_cid[v] = v; s->n_unrooted[v] = 1;
        return;
    }
    for (int k = 0; k < s->w->K; ++k) {
        if (!(s->access[v] >> k & 1)) continue;
        int nb, e;
        nbr(s, v, k, &nb, &e);
        if (!whole) { compare_cid(s, v, k, nb); continue; }
        if (s->cid[nb] == RESET) {
            if (!s->unrooted[v]) { s->busy[v] = 1; s->n_cid[v] = RESET; s->pointer[v] = PTR_C; break; }
        } else
            compare_cid(s, v, k, nb);
        check_grown(s, v, nb);
    }
}

/* _Node.syncing, :1106-1122 */
static void node_syncing(sf *s, int v)
{
    const int q = node_pos(s->w, v);
    const int detector_defect = q >= s->w->NLb && s->defect[v];
    if (s->pointer[v] == PTR_C) s->n_active[v] = (uint8_t)detector_defect;
    else {
        int nb, e;
        nbr(s, v, s->pointer[v], &nb, &e);
        s->n_active[v] = s->active[nb];
        if (detector_defect) {
            s->busy[v] = 1;
            s->n_defect[nb] ^= 1; s->n_defect[v] ^= 1;
            s->corr[e] ^= 1;
        }
    }
    if (s->active[v] != s->n_active[v]) s->busy[v] = 1;
}

/* Snowflake.merge, :324-353.  Adds the 'all' and 'unrooting' runtimes. */
static void sf_merge(sf *s, int whole, int *rt_all, int *rt_unrooting)
{
    const int N = s->N;
    *rt_all += 1;
    for (;;) {
        for (int v = 0; v < N; ++v) {               /* _Node.merging (:1097-1104) + mergers (:1237-1256) */
            s->busy[v] = 0;
            if (s->w->merger_slow) { node_syncing(s, v); node_flooding(s, v, whole); }
            else { node_flooding(s, v, whole); node_syncing(s, v); }
        }
        int any_reset = 0, any_busy = 0;
        for (int v = 0; v < N; ++v) {               /* update_after_merging, :1131-1137 */
            s->cid[v] = s->n_cid[v]; s->defect[v] = s->n_defect[v]; s->active[v] = s->n_active[v];
            s->unrooted[v] = s->n_unrooted[v]; s->grown[v] = s->n_grown[v];
            any_reset |= s->cid[v] == RESET; any_busy |= s->busy[v];
        }
        *rt_all += 1; *rt_unrooting += any_reset;
        if (!any_busy) break;
    }
}

/* Snowflake.drop (:298-310) with contacts (:1547-1564) and friendships (:1161-1211); the
 * floor-sheet corrections go to Pairs in edge order and are reported in floor_bits. */
static void sf_drop(sf *s, const uint8_t *top_syndrome, uint32_t *floor_bits, int *count)
{
    const oracle_window *w = s->w;
    const int EL = w->EL, h = w->h, N = s->N, E = s->E;
    for (int e = 0; e < E; ++e) {                   /* edge.CONTACT.drop() */
        if (e + EL < E) { s->n_growth[e + EL] = s->growth[e]; s->n_corr[e + EL] = s->corr[e]; }
    }
    for (int l = 0; l < EL; ++l) {                  /* FloorContact.drop: floor block = last EL indices */
        const int e = (h - 1) * EL + l;
        if (!s->corr[e]) continue;
        SETBIT(floor_bits, l);
        pairs_load(s, endpoint_key(s, w->edge_q[2 * l], w->edge_dt[2 * l]),
                   endpoint_key(s, w->edge_q[2 * l + 1], w->edge_dt[2 * l + 1]), count);
    }
    for (int e = 0; e < E; ++e) { s->growth[e] = s->n_growth[e]; s->corr[e] = s->n_corr[e]; }   /* update_after_drop */
    for (int v = 0; v < N; ++v) {                   /* node.FRIENDSHIP.drop() */
        s->grown[v] = s->n_grown[v] = s->unrooted[v] = s->n_unrooted[v] = 0;
        const int t = node_layer(w, v);
        if (t > 0) {
            const int r = v + (v < w->NLb * h ? w->NLb : w->NLd);
            s->n_defect[r] = s->defect[v]; s->n_active[r] = s->active[v];
            s->n_cid[r] = s->cid[v] + (s->cid[v] < w->NLb * h ? w->NLb : w->NLd);          /* id_below, :186-199 */
            s->n_pointer[r] = s->pointer[v]; s->n_whole[r] = s->whole[v];
        }
        if (t == h - 1) { s->n_defect[v] = 0; s->n_active[v] = 0; }                         /* TopSheetFriendship */
    }
    for (int q = w->NLb; q < s->NL; ++q)            /* Snowflake._load, :312-322 */
        if (top_syndrome[q - w->NLb]) s->n_defect[node_id(w, q, h - 1)] ^= 1;
    for (int v = 0; v < N; ++v) {                   /* update_after_drop, :1033-1043 */
        s->defect[v] = s->n_defect[v]; s->active[v] = s->n_active[v]; s->cid[v] = s->n_cid[v];
        s->pointer[v] = s->n_pointer[v]; s->whole[v] = s->n_whole[v];
    }
}

/* NothingFriendship.find_broken_pointers (:1221-1223) + _FullUnrooter.start (:1387-1389) */
static void find_broken_pointers(sf *s, int v)
{
    if (node_layer(s->w, v) != 0 || s->pointer[v] == PTR_C) return;
    if (s->w->nb_dt[node_pos(s->w, v) * s->w->K + s->pointer[v]] == -1) {   /* 'D' in pointer */
        s->cid[v] = RESET; s->pointer[v] = PTR_C;
    }
}

/* Snowflake.decode (:216-281): drop, then _TwoOne / _OneOne.finish_decode (:1290-1336) */
static void sf_decode(sf *s, const uint8_t *top_syndrome, uint32_t *floor_bits, int *count,
                      int *rt_all, int *rt_unrooting)
{
    const int N = s->N;
    *rt_all = 0; *rt_unrooting = 0;
    sf_drop(s, top_syndrome, floor_bits, count);
    if (s->w->schedule_11) {
        for (int v = 0; v < N; ++v) {               /* _Node.grow, :1051-1058 */
            if (s->active[v]) node_grow(s, v);
            find_broken_pointers(s, v);
        }
        update_access_all(s);
        sf_merge(s, 1, rt_all, rt_unrooting);
        return;
    }
    for (int v = 0; v < N; ++v) {                   /* grow_whole, :1060-1066 */
        if (s->active[v] && s->whole[v]) { node_grow(s, v); s->grown[v] = 1; s->n_grown[v] = 1; }
        find_broken_pointers(s, v);
    }
    update_access_all(s);
    sf_merge(s, 1, rt_all, rt_unrooting);
    for (int v = 0; v < N; ++v) {                   /* grow_half, :1068-1073 */
        s->unrooted[v] = 0; s->n_unrooted[v] = 0;
        if (s->active[v] && !s->whole[v] && !s->grown[v]) node_grow(s, v);
    }
    update_access_all(s);
    sf_merge(s, 0, rt_all, rt_unrooting);
}

/* Frugal.advance (_schemes.py:504-530) + get_logical_error (:262-276) for one cycle,
 * fresh errors given as bits over the top block's local edge index. */
static void sf_cycle(sf *s, const uint32_t *fresh_bits, uint32_t *floor_bits, int *m, int *rt_all, int *rt_unrooting)
{
    const oracle_window *w = s->w;
    const int EL = w->EL, h = w->h, E = s->E;
    int count = 0;
    /* _raise_window (:532-543): floor errors to Pairs (ascending edge id), the rest one layer down */
    for (int l = 0; l < EL; ++l)
        if (s->error[(h - 1) * EL + l])
            pairs_load(s, endpoint_key(s, w->edge_q[2 * l], w->edge_dt[2 * l]),
                       endpoint_key(s, w->edge_q[2 * l + 1], w->edge_dt[2 * l + 1]), &count);
    memmove(s->error + EL, s->error, E - EL);
    /* _load (:545-564): the fresh sheet; syndrome = carried future-boundary defects xor new detector flips */
    uint8_t *top = calloc(w->NLd, 1);
    memcpy(top, s->fb_syndrome, w->NLd);
    memset(s->fb_syndrome, 0, w->NLd);
    for (int l = 0; l < EL; ++l) {
        const int flipped = GETBIT(fresh_bits, l) != 0;
        s->error[l] = (uint8_t)flipped;
        if (!flipped) continue;
        for (int x = 0; x < 2; ++x) {
            const int q = w->edge_q[2 * l + x];
            if (q < w->NLb) continue;                               /* space boundary */
            if (w->edge_dt[2 * l + x] == 1) s->fb_syndrome[q - w->NLb] ^= 1;   /* on the future boundary t = h */
            else top[q - w->NLb] ^= 1;
        }
    }
    sf_decode(s, top, floor_bits, &count, rt_all, rt_unrooting);
    free(top);
    pairs_lower(s);
    *m = count;
}

int oracle_sf_stream(const oracle_window *w, int32_t ncycles, const uint32_t *fresh_bits,
                     int32_t *rt_all, int32_t *rt_unrooting, uint32_t *floor_bits, int32_t *m_cycle)
{
    if (w->K > 16) return -1;
    sf *s = sf_new(w);
    sf_reset(s);
    const int W = WORDS(w->EL);
    memset(floor_bits, 0, sizeof(uint32_t) * W * ncycles);
    for (int c = 0; c < ncycles; ++c) {
        int m, a, u;
        sf_cycle(s, fresh_bits + (size_t)c * W, floor_bits + (size_t)c * W, &m, &a, &u);
        rt_all[c] = a; rt_unrooting[c] = u; m_cycle[c] = m;
    }
    sf_free(s);
    return 0;
}

/* ------------------------------------------------------------- Monte Carlo */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(uint64_t s[4])
{
    const uint64_t r = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return r;
}
static uint64_t sm64(uint64_t *x)
{
    uint64_t z = (*x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

typedef struct { const oracle_window *w; const double *class_p; uint64_t seed; int64_t nshots; int n, tid; int64_t m, cycles; } mcjob;

static void *sf_mc_worker(void *arg)
{
    mcjob *j = arg;
    const oracle_window *w = j->w;
    const int W = WORDS(w->EL), d = w->d, h = w->h;
    sf *s = sf_new(w);
    uint32_t *fresh = malloc(4 * W), *floor_bits = malloc(4 * W);
    uint64_t x = j->seed + 0x51a2c3d4e5f60718ull * (uint64_t)(j->tid + 1), rng[4];
    for (int k = 0; k < 4; ++k) rng[k] = sm64(&x);
    /* Frugal.run phases, _schemes.py:626-632 */
    const int total = h + j->n * d + (d - 1) + 1 + 2 * h;
    for (int64_t shot = 0; shot < j->nshots; ++shot) {
        sf_reset(s);
        for (int c = 0; c < total; ++c) {
            const int cleanse = c >= total - 2 * h, excl = c == total - 2 * h - 1;
            memset(fresh, 0, 4 * W);
            if (!cleanse)
                for (int l = 0; l < w->EL; ++l) {     /* noise/main.py:114-115,309-314 + _base_classes.py:418-424 */
                    const double u = (double)(xo_next(rng) >> 11) * 0x1.0p-53;
                    if (u < j->class_p[w->edge_class[l]] && !(excl && w->edge_fb[l])) SETBIT(fresh, l);
                }
            int m, a, u2;
            memset(floor_bits, 0, 4 * W);
            sf_cycle(s, fresh, floor_bits, &m, &a, &u2);
            j->m += m;
        }
        j->cycles += total;
    }
    free(fresh); free(floor_bits); sf_free(s);
    return NULL;
}

int64_t oracle_sf_mc(const oracle_window *w, const double *class_p, uint64_t seed, int64_t nshots,
                     int32_t n, int nthreads, int64_t *cycles)
{
    if (nthreads < 1) nthreads = 1;
    mcjob *jobs = calloc(nthreads, sizeof(mcjob));
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = (mcjob){w, class_p, seed, nshots / nthreads + (t < nshots % nthreads), n, t, 0, 0};
        pthread_create(&th[t], NULL, sf_mc_worker, &jobs[t]);
    }
    int64_t m = 0;
    for (int t = 0; t < nthreads; ++t) { pthread_join(th[t], NULL); m += jobs[t].m; if (cycles) *cycles += jobs[t].cycles; }
    free(jobs); free(th);
    return m;
}
