/*
 * luf_oracle_dcs.c -- CPU restatement of the reference's decoder confidence scores.
 * TEST INFRASTRUCTURE ONLY; see luf_oracle.h.  Citations are to /root/reference.
 *
 *   unclustered_edge_fraction   localuf/decoders/_base_uf.py:101-121
 *   swim_distance               localuf/decoders/_base_uf.py:250-316 (networkx.bidirectional_dijkstra over
 *                               _cached_swim_graph with the weights of _swim_graph; restated as a plain
 *                               Dijkstra from the west meta boundary -- the two searches return the same
 *                               path length up to the rounding of a different summation order)
 *   unclustered_node_fraction   localuf/decoders/uf.py:188-197 (a node is in a cluster of size > 1 iff it
 *                               has a fully grown edge: every such edge has been merged; a BURNT edge of
 *                               dynamic UF, uf.py:263-266, joins nothing)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double d; int v; } heap_item;

static void heap_push(heap_item *h, int *n, double d, int v)
{
    int i = (*n)++;
    while (i > 0) {
        const int p = (i - 1) / 2;
        if (h[p].d <= d) break;
        h[i] = h[p]; i = p;
    }
    h[i].d = d; h[i].v = v;
}

static heap_item heap_pop(heap_item *h, int *n)
{
    const heap_item top = h[0], last = h[--(*n)];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && h[c + 1].d < h[c].d) ++c;
        if (last.d <= h[c].d) break;
        h[i] = h[c]; i = c;
    }
    h[i] = last;
    return top;
}

/* growth: [nshots][n_edges] codes 0 ungrown, 1 half, 2 full, 3 burnt (Growth, constants.py:17-31).
 * weight / in_growth may be NULL (all 1).  side: 1 west, 2 east.  Outputs optional. */
int oracle_dcs_batch(int n_nodes, int n_edges, const int32_t *edge_u, const int32_t *edge_v, const double *weight,
                     const uint8_t *in_growth, const uint8_t *side, long long nshots, const uint8_t *growth,
                     double *swim, double *uef, int32_t *clustered)
{
    /* CSR adjacency */
    int32_t *off = calloc((size_t)n_nodes + 1, sizeof *off), *adj = malloc(sizeof *adj * 2 * (size_t)(n_edges ? n_edges : 1));
    double *dist = malloc(sizeof *dist * (size_t)n_nodes);
    uint8_t *done = malloc((size_t)n_nodes), *mark = malloc((size_t)n_nodes);
    heap_item *heap = malloc(sizeof *heap * (2 * (size_t)n_edges + (size_t)n_nodes + 1));
    if (!off || !adj || !dist || !done || !mark || !heap) return -1;
    for (int e = 0; e < n_edges; ++e) { off[edge_u[e] + 1]++; off[edge_v[e] + 1]++; }
    for (int v = 0; v < n_nodes; ++v) off[v + 1] += off[v];
    {
        int32_t *pos = malloc(sizeof *pos * (size_t)n_nodes);
        memcpy(pos, off, sizeof *pos * (size_t)n_nodes);
        for (int e = 0; e < n_edges; ++e) { adj[pos[edge_u[e]]++] = e; adj[pos[edge_v[e]]++] = e; }
        free(pos);
    }
    double total = 0.0;                                   /* _total_edge_weight, :101-104 */
    for (int e = 0; e < n_edges; ++e)
        if (!in_growth || in_growth[e]) total += weight ? weight[e] : 1.0;
    for (long long k = 0; k < nshots; ++k) {
        const uint8_t *g = growth + (size_t)k * n_edges;
        if (uef) {                                        /* :117-121 */
            double num = 0.0;
            for (int e = 0; e < n_edges; ++e)
                if (!in_growth || in_growth[e])
                    num += (weight ? weight[e] : 1.0) * (g[e] == 0 ? 0.0 : g[e] == 1 ? 0.5 : 1.0);   /* Growth.as_float */
            uef[k] = 1 - num / total;
        }
        if (clustered) {
            memset(mark, 0, (size_t)n_nodes);
            int c = 0;
            for (int e = 0; e < n_edges; ++e)
                if (g[e] == 2) { mark[edge_u[e]] = 1; mark[edge_v[e]] = 1; }
            for (int v = 0; v < n_nodes; ++v) c += mark[v];
            clustered[k] = c;
        }
        if (swim) {
            int hn = 0;
            for (int v = 0; v < n_nodes; ++v) {
                done[v] = 0;
                dist[v] = side[v] == 1 ? 0.0 : INFINITY;  /* west meta boundary: weight-0 edges, :312-313 */
                if (side[v] == 1) heap_push(heap, &hn, 0.0, v);
            }
            double best = INFINITY;
            while (hn) {
                const heap_item it = heap_pop(heap, &hn);
                if (done[it.v]) continue;
                done[it.v] = 1;
                if (side[it.v] == 2) { best = it.d; break; }          /* east meta boundary at distance + 0 */
                for (int i = off[it.v]; i < off[it.v + 1]; ++i) {
                    const int e = adj[i], o = edge_u[e] == it.v ? edge_v[e] : edge_u[e];
                    const double w = weight ? weight[e] : 1.0;
                    const double cost = g[e] == 0 ? w : g[e] == 1 ? w / 2 : 0.0;    /* _swim_graph, :294-303 */
                    const double x = it.d + cost;
                    if (x < dist[o]) { dist[o] = x; heap_push(heap, &hn, x, o); }
                }
            }
            swim[k] = best;
        }
    }
    free(off); free(adj); free(dist); free(done); free(mark); free(heap);
    return 0;
}
