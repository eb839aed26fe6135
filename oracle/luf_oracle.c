/*
 * luf_oracle.c -- CPU restatement of the reference's Monte-Carlo decode path.
 * TEST INFRASTRUCTURE ONLY; see luf_oracle.h.  Citations are to /root/reference.
 */
#include "luf_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define WORDS(n) (((n) + 31) / 32)
#define GETBIT(a, k) (((a)[(k) >> 5] >> ((k) & 31)) & 1u)
#define FLIPBIT(a, k) ((a)[(k) >> 5] ^= (1u << ((k) & 31)))

enum { UNGROWN = 0, HALF = 1, FULL = 2, BURNT = 3 };
enum { F_BOUNDARY = 1, F_WEST = 2, F_EAST = 4, F_FUTURE = 8 };

/* ------------------------------------------------------------------ RNG */
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static inline uint64_t xoshiro_next(uint64_t s[4])
{
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}

static inline double uniform01(uint64_t s[4]) { return (double)(xoshiro_next(s) >> 11) * 0x1.0p-53; }

/* noise/main.py:114-115 (_Uniform.make_error) and :309-314 (CircuitLevel.make_error):
 * one uniform draw per fresh edge, flipped iff draw < its class probability. */
void oracle_sample(const oracle_graph *g, const double *class_p, uint64_t rng[4], uint32_t *err_bits)
{
    memset(err_bits, 0, sizeof(uint32_t) * WORDS(g->n_edges));
    for (int e = 0; e < g->n_edges; ++e) {
        const uint8_t c = g->edge_class[e];
        if (c == 255) continue;
        if (uniform01(rng) < class_p[c]) FLIPBIT(err_bits, e);
    }
}

/* _base_classes.py:436-450 (XOR endpoints) and :434 (drop boundary nodes). */
void oracle_syndrome(const oracle_graph *g, const uint32_t *err_bits, uint32_t *syn_bits)
{
    const int B = g->n_boundary;
    memset(syn_bits, 0, sizeof(uint32_t) * WORDS(g->n_nodes - B));
    for (int e = 0; e < g->n_edges; ++e) {
        if (!GETBIT(err_bits, e)) continue;
        const int u = g->edge_u[e], v = g->edge_v[e];
        if (u >= B) FLIPBIT(syn_bits, u - B);
        if (v >= B) FLIPBIT(syn_bits, v - B);
    }
}

/* _schemes.py:123-129 */
int oracle_logical_batch(const oracle_graph *g, const uint32_t *err_bits, const uint32_t *corr_bits)
{
    uint32_t acc = 0;
    for (int w = 0; w < WORDS(g->n_edges); ++w)
        acc ^= (err_bits[w] ^ corr_bits[w]) & g->west_edge_mask[w];
    return __builtin_parity(acc);
}

/* ------------------------------------------------------------------- UF */
typedef struct {
    int32_t *parent;      /* uf.py:63-64  parents */
    int32_t *size;        /* uf.py:498    cluster.size   (valid at roots) */
    uint8_t *odd;         /* uf.py:499    cluster.odd    (valid at roots) */
    int32_t *bnd;         /* uf.py:500    cluster.boundary or -1 (valid at roots) */
    int32_t *next, *tail; /* members of a cluster as a linked list headed by its root;
                             vision(C) = not-yet-full edges incident to members (uf.py:519,278) */
    uint8_t *growth;      /* _base_uf.py:47 */
    int32_t *seen_round, *seen_root; /* which cluster already grew an edge this round */
    int32_t *active;      /* uf.py:69-70 active_clusters, as a list of roots */
    int32_t *merge_ls;    /* uf.py:219 */
    int32_t *stack, *tree_edge, *tree_node;
    uint8_t *discovered, *defect;
} uf_ws;

static uf_ws *uf_ws_new(const oracle_graph *g)
{
    const int N = g->n_nodes, E = g->n_edges;
    uf_ws *w = (uf_ws *)calloc(1, sizeof(uf_ws));
    w->parent = malloc(sizeof(int32_t) * N); w->size = malloc(sizeof(int32_t) * N);
    w->odd = malloc(N); w->bnd = malloc(sizeof(int32_t) * N);
    w->next = malloc(sizeof(int32_t) * N); w->tail = malloc(sizeof(int32_t) * N);
    w->growth = malloc(E);
    w->seen_round = malloc(sizeof(int32_t) * E); w->seen_root = malloc(sizeof(int32_t) * E);
    w->active = malloc(sizeof(int32_t) * N); w->merge_ls = malloc(sizeof(int32_t) * E);
    w->stack = malloc(sizeof(int32_t) * N);
    w->tree_edge = malloc(sizeof(int32_t) * N); w->tree_node = malloc(sizeof(int32_t) * N);
    w->discovered = malloc(N); w->defect = malloc(N);
    return w;
}

static void uf_ws_free(uf_ws *w)
{
    free(w->parent); free(w->size); free(w->odd); free(w->bnd); free(w->next); free(w->tail);
    free(w->growth); free(w->seen_round); free(w->seen_root); free(w->active); free(w->merge_ls);
    free(w->stack); free(w->tree_edge); free(w->tree_node); free(w->discovered); free(w->defect);
    free(w);
}

/* uf.py:247-254 -- find with full path compression (iterative form) */
static int uf_find(uf_ws *w, int v)
{
    int r = v;
    while (w->parent[r] != r) r = w->parent[r];
    while (w->parent[v] != r) { int n = w->parent[v]; w->parent[v] = r; v = n; }
    return r;
}

static int cmp_i32(const void *a, const void *b)
{
    const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

static int uf_decode_ws(const oracle_graph *g, uf_ws *w, int flags, const uint32_t *syn_bits,
                        uint32_t *corr_bits, uint8_t *growth_out)
{
    const int N = g->n_nodes, B = g->n_boundary, E = g->n_edges;
    const int dynamic = flags & ORACLE_UF_DYNAMIC, west = flags & ORACLE_UF_WEST, node = flags & ORACLE_UF_NODE,
              bucket = flags & ORACLE_UF_BUCKET;
    int n_active = 0, rounds = 0;

    /* uf.py:79-96 reset + :496-519 one singleton cluster per node */
    for (int v = 0; v < N; ++v) {
        w->parent[v] = v; w->size[v] = 1; w->odd[v] = 0;
        w->bnd[v] = (g->node_flags[v] & F_BOUNDARY) ? v : -1;
        w->next[v] = -1; w->tail[v] = v; w->defect[v] = 0;
    }
    memset(w->growth, UNGROWN, E);
    for (int e = 0; e < E; ++e) w->seen_round[e] = -1;
    /* uf.py:98-104 load */
    for (int k = 0; k < N - B; ++k)
        if (GETBIT(syn_bits, k)) {
            const int v = B + k;
            w->odd[v] = 1; w->defect[v] = 1; w->active[n_active++] = v;
        }

    /* uf.py:208-209 validate */
    while (n_active > 0) {
        int n_merge = 0;
        /* uf.py:221-222,234-245: every active cluster adds one growth increment to
         * every edge of its vision; an edge seen by two different active clusters
         * gets two.  Saturating (Growth(3) would raise in the reference). */
        /* Bucket decoders: only the active clusters of the smallest metric grow this round.
         * NodeBUF (node_buf.py:42-47): metric = cluster size.  BUF (buf.py:64-78): metric = len(vision),
         * the number of distinct not-yet-FULL edges at the cluster's nodes. */
        int smallest = 0;
        if (bucket) {
            smallest = 0x7FFFFFFF;
            for (int a = 0; a < n_active; ++a) {
                const int r = w->active[a];
                int metric = w->size[r];
                if (!node) {
                    const int mark = -2 - rounds;
                    metric = 0;
                    for (int u = r; u != -1; u = w->next[u])
                        for (int k = g->inc_off[u]; k < g->inc_off[u + 1]; ++k) {
                            const int e = g->inc_edge[k];
                            if (w->growth[e] >= FULL) continue;
                            if (w->seen_round[e] == mark && w->seen_root[e] == r) continue;
                            w->seen_round[e] = mark; w->seen_root[e] = r;
                            ++metric;
                        }
                }
                w->stack[a] = metric;                                   /* scratch: the stack is idle during validation */
                if (metric < smallest) smallest = metric;
            }
        }
        for (int a = 0; a < n_active; ++a) {
            const int r = w->active[a];
            if (bucket && w->stack[a] != smallest) continue;
            for (int u = r; u != -1; u = w->next[u])
                for (int k = g->inc_off[u]; k < g->inc_off[u + 1]; ++k) {
                    const int e = g->inc_edge[k];
                    if (w->growth[e] >= FULL) continue;                 /* not in any vision */
                    /* UF: a vision (a set) holds e once.  NodeUF (node_uf.py:45-67): every frontier node
                     * grows its own active edges, so an edge inside the cluster gets one increment per end. */
                    if (!node && w->seen_round[e] == rounds && w->seen_root[e] == r) continue;
                    w->seen_round[e] = rounds; w->seen_root[e] = r;
                    if (++w->growth[e] == FULL) w->merge_ls[n_merge++] = e;
                }
        }
        ++rounds;
        /* canonical order: ascending edge id (oracle/ref_canonical.py) */
        qsort(w->merge_ls, n_merge, sizeof(int32_t), cmp_i32);
        /* uf.py:223-229 */
        for (int m = 0; m < n_merge; ++m) {
            const int e = w->merge_ls[m];
            const int cu = uf_find(w, g->edge_u[e]), cv = uf_find(w, g->edge_v[e]);
            if (dynamic) {  /* uf.py:261-266 */
                if (cu == cv || (w->bnd[cu] >= 0 && w->bnd[cv] >= 0)) { w->growth[e] = BURNT; continue; }
            } else if (cu == cv) continue; /* uf.py:256-259 */
            /* uf.py:268-280 union by size, ties keep the cluster of e[0] */
            int larger, smaller;
            if (w->size[cu] >= w->size[cv]) { larger = cu; smaller = cv; } else { larger = cv; smaller = cu; }
            w->parent[smaller] = larger;
            w->size[larger] += w->size[smaller];
            w->odd[larger] ^= w->odd[smaller];
            w->next[w->tail[larger]] = smaller; w->tail[larger] = w->tail[smaller];
            if (west) {   /* uf.py:544-549: keep the boundary node that is further west */
                if (w->bnd[smaller] >= 0 && (w->bnd[larger] < 0 ||
                        g->node_long[w->bnd[smaller]] < g->node_long[w->bnd[larger]]))
                    w->bnd[larger] = w->bnd[smaller];
            } else {      /* uf.py:536-538 */
                if (w->bnd[larger] < 0 && w->bnd[smaller] >= 0) w->bnd[larger] = w->bnd[smaller];
            }
        }
        /* uf.py:290-301: active <=> odd and no boundary (roots only) */
        {
            int n_new = 0;
            for (int a = 0; a < n_active; ++a) {
                const int r = uf_find(w, w->active[a]);
                if (!w->odd[r] || w->bnd[r] >= 0) continue;
                int dup = 0;
                for (int b = 0; b < n_new; ++b) if (w->active[b] == r) { dup = 1; break; }
                /* safe to compact in place: n_new <= a */
                if (!dup) w->active[n_new++] = r;
            }
            n_active = n_new;
        }
    }

    if (growth_out) memcpy(growth_out, w->growth, E);
    memset(corr_bits, 0, sizeof(uint32_t) * WORDS(E));

    /* uf.py:305-344 peel: per surviving cluster, a spanning tree from
     * cluster.boundary (or its root), then leaf-to-root peeling. */
    memset(w->discovered, 0, N);
    for (int v = 0; v < N; ++v) {
        if (w->parent[v] != v) continue;                 /* one tree per cluster (uf.py:321-323) */
        const int root = w->bnd[v] >= 0 ? w->bnd[v] : v;
        int sp = 0, nt = 0;
        w->discovered[root] = 1; w->stack[sp++] = root;
        while (sp > 0) {                                 /* uf.py:335-343, LIFO */
            const int u = w->stack[--sp];
            for (int k = g->inc_off[u]; k < g->inc_off[u + 1]; ++k) {   /* canonical: ascending edge id */
                const int e = g->inc_edge[k], o = g->inc_other[k];
                if (w->growth[e] != FULL || w->discovered[o]) continue;
                w->discovered[o] = 1; w->stack[sp++] = o;
                w->tree_edge[nt] = e; w->tree_node[nt] = o; ++nt;
            }
        }
        for (int t = nt - 1; t >= 0; --t) {              /* uf.py:309-312 */
            const int e = w->tree_edge[t], c = w->tree_node[t];
            if (!w->defect[c]) continue;
            FLIPBIT(corr_bits, e);
            w->defect[g->edge_u[e]] ^= 1; w->defect[g->edge_v[e]] ^= 1;
        }
    }
    return rounds;
}

int oracle_uf_decode(const oracle_graph *g, int flags, const uint32_t *syn_bits,
                     uint32_t *corr_bits, uint8_t *growth_out)
{
    uf_ws *w = uf_ws_new(g);
    const int r = uf_decode_ws(g, w, flags, syn_bits, corr_bits, growth_out);
    uf_ws_free(w);
    return r;
}

/* ----------------------------------------------------------- batch drivers */
#ifdef ORACLE_HAVE_CA
int oracle_ca_decode_batch(const oracle_graph *g, int decoder, int flags, int64_t nshots,
                           const uint32_t *syn_bits, uint32_t *corr_bits,
                           uint8_t *growth_out, int32_t *steps);
#endif

int oracle_decode_batch(const oracle_graph *g, int decoder, int flags, int64_t nshots,
                        const uint32_t *syn_bits, uint32_t *corr_bits,
                        uint8_t *growth_out, int32_t *steps)
{
    const int WD = WORDS(g->n_nodes - g->n_boundary), WE = WORDS(g->n_edges);
    if (decoder == ORACLE_DEC_UF) {
        uf_ws *w = uf_ws_new(g);
        for (int64_t s = 0; s < nshots; ++s) {
            const int r = uf_decode_ws(g, w, flags, syn_bits + s * WD, corr_bits + s * WE,
                                       growth_out ? growth_out + s * g->n_edges : NULL);
            if (steps) { steps[2 * s] = r; steps[2 * s + 1] = 0; }
        }
        uf_ws_free(w);
        return 0;
    }
#ifdef ORACLE_HAVE_CA
    return oracle_ca_decode_batch(g, decoder, flags, nshots, syn_bits, corr_bits, growth_out, steps);
#else
    return -1;
#endif
}

typedef struct {
    const oracle_graph *g; int decoder, flags; const double *class_p;
    uint64_t seed; int64_t nshots; int tid; int64_t fails;
} mc_job;

/* _schemes.py:131-155: make_error -> get_syndrome -> reset+decode -> leftover -> logical */
static void *mc_worker(void *arg)
{
    mc_job *j = (mc_job *)arg;
    const oracle_graph *g = j->g;
    const int WD = WORDS(g->n_nodes - g->n_boundary), WE = WORDS(g->n_edges);
    uint32_t *err = malloc(4 * WE), *corr = malloc(4 * WE), *syn = malloc(4 * (WD ? WD : 1));
    uint64_t sm = j->seed + 0x632be59bd9b4e019ull * (uint64_t)(j->tid + 1), rng[4];
    for (int k = 0; k < 4; ++k) rng[k] = splitmix64(&sm);
    uf_ws *w = j->decoder == ORACLE_DEC_UF ? uf_ws_new(g) : NULL;
    int64_t fails = 0;
    for (int64_t s = 0; s < j->nshots; ++s) {
        oracle_sample(g, j->class_p, rng, err);
        oracle_syndrome(g, err, syn);
        if (w) uf_decode_ws(g, w, j->flags, syn, corr, NULL);
        else if (oracle_decode_batch(g, j->decoder, j->flags, 1, syn, corr, NULL, NULL)) { fails = -1; break; }
        fails += oracle_logical_batch(g, err, corr);
    }
    if (w) uf_ws_free(w);
    free(err); free(corr); free(syn);
    j->fails = fails;
    return NULL;
}

int64_t oracle_mc_batch(const oracle_graph *g, int decoder, int flags, const double *class_p,
                        uint64_t seed, int64_t nshots, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    mc_job *jobs = calloc(nthreads, sizeof(mc_job));
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        jobs[t] = (mc_job){g, decoder, flags, class_p, seed, nshots / nthreads + (t < nshots % nthreads), t, 0};
        pthread_create(&th[t], NULL, mc_worker, &jobs[t]);
    }
    int64_t total = 0; int bad = 0;
    for (int t = 0; t < nthreads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].fails < 0) bad = 1; else total += jobs[t].fails;
    }
    free(jobs); free(th);
    return bad ? -1 : total;
}

/* ---------------------------------------------------------- forward scheme */
#define FP_NONE (-1)
#define FP_WEST (-2)
#define FP_EAST (-3)
#define FP_DEAD (-4)

static int fw_key(const oracle_forward *f, int v)
{
    return f->node_side[v] == 1 ? FP_WEST : f->node_side[v] == 2 ? FP_EAST : v;
}
static void fw_add(int32_t *partner, int a, int b, int *count)
{
    if (a >= 0) partner[a] = b;
    if (b >= 0) partner[b] = a;
    if ((a == FP_WEST && b == FP_EAST) || (a == FP_EAST && b == FP_WEST)) ++*count;   /* _pairs.py:103-105 */
}
static int fw_take(int32_t *partner, int u)
{
    const int w = partner[u];
    partner[u] = FP_NONE;
    if (w >= 0) partner[w] = FP_NONE;
    return w;
}
/* Pairs.load, _pairs.py:49-66 */
static void fw_load(int32_t *partner, int u, int v, int *count)
{
    if (u >= 0 && partner[u] != FP_NONE) {
        const int w = fw_take(partner, u);
        if (v >= 0 && partner[v] != FP_NONE) fw_add(partner, w, fw_take(partner, v), count);
        else if (v != w || v < 0) fw_add(partner, v, w, count);
    } else if (v >= 0 && partner[v] != FP_NONE)
        fw_add(partner, u, fw_take(partner, v), count);
    else
        fw_add(partner, u, v, count);
}

int oracle_forward_stream(const oracle_graph *g, const oracle_forward *f, int decoder, int flags, int32_t ncycles,
                          const uint32_t *init_buffer_bits, const uint32_t *fresh_bits,
                          int32_t *steps, uint32_t *commit_bits, int32_t *m_cycle)
{
    const int N = g->n_nodes, B = g->n_boundary, E = g->n_edges, WE = WORDS(E), WD = WORDS(N - B);
    uint32_t *error = calloc(WE, 4), *buffer = calloc(WE, 4), *commit = calloc(WE, 4), *corr = calloc(WE, 4);
    uint32_t *syn = calloc(WD ? WD : 1, 4);
    int32_t *partner = malloc(4 * N), *np2 = malloc(4 * N);
    for (int v = 0; v < N; ++v) partner[v] = FP_NONE;
    memcpy(buffer, init_buffer_bits, 4 * WE);
    int rc = 0;
    for (int c = 0; c < ncycles && !rc; ++c) {
        /* Forward._make_error (:302-324): lower the buffer leftover by the commit height, add the fresh sheet */
        memcpy(error, fresh_bits + (size_t)c * WE, 4 * WE);
        for (int e = 0; e < E; ++e)
            if (GETBIT(buffer, e) && f->lower_edge[e] >= 0) error[f->lower_edge[e] >> 5] |= 1u << (f->lower_edge[e] & 31);
        /* syndrome xor artificial defects of the previous commit (:351-368) */
        oracle_syndrome(g, error, syn);
        for (int e = 0; e < E; ++e)
            if (GETBIT(commit, e))
                for (int x = 0; x < 2; ++x) {
                    const int a = f->art_node[2 * e + x];
                    if (a >= 0) FLIPBIT(syn, a - B);
                }
        /* decoder.reset(); decoder.decode(syndrome) (:440-441) */
        rc = oracle_decode_batch(g, decoder, flags, 1, syn, corr, NULL, steps + 2 * c);
        /* Forward._get_leftover (:326-339) */
        for (int w = 0; w < WE; ++w) {
            commit[w] = (error[w] ^ corr[w]) & f->commit_mask[w];
            buffer[w] = error[w] & ~f->commit_mask[w];
        }
        memcpy(commit_bits + (size_t)c * WE, commit, 4 * WE);
        /* Forward.get_logical_error (:341-349), ascending edge id */
        int count = 0;
        for (int e = 0; e < E; ++e)
            if (GETBIT(commit, e)) fw_load(partner, fw_key(f, g->edge_u[e]), fw_key(f, g->edge_v[e]), &count);
        m_cycle[c] = count;
        /* LogicalCounter.count's lowering (_pairs.py:82-111) */
        for (int v = 0; v < N; ++v) np2[v] = FP_NONE;
        for (int v = 0; v < N; ++v) {
            const int p = partner[v];
            if (p == FP_NONE) continue;
            const int lv = f->node_lower[v];
            if (lv < 0) continue;                               /* this end is out of reach now */
            np2[lv] = p >= 0 ? (f->node_lower[p] >= 0 ? f->node_lower[p] : FP_DEAD) : p;
        }
        memcpy(partner, np2, 4 * N);
    }
    free(error); free(buffer); free(commit); free(corr); free(syn); free(partner); free(np2);
    return rc;
}
