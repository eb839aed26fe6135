// luf_dcs.cuh -- decoder confidence scores of a batch of growth states
// (localuf/decoders/_base_uf.py:101-121 unclustered_edge_fraction, :250-316 swim_distance;
// localuf/decoders/uf.py:188-197 unclustered_node_fraction).
//
// One thread block owns one shot.  The swim distance -- the shortest west-to-east path where an ungrown
// edge costs its weight, a half-grown edge half of it and a fully grown or burnt edge nothing
// (_swim_graph, :281-304) -- is an edge-parallel Bellman-Ford in shared memory: dist[] holds the bit
// patterns of non-negative doubles, which order like unsigned integers, so a relaxation is one 64-bit
// atomicMin.  dist[v] is the minimum over paths of the left-to-right sum of the edge costs, the value
// Dijkstra's algorithm returns.  A sweep visits the edges in the order of their hop distance from the west
// boundary (a table made once per graph): the 256 threads then work on one lattice column after the other,
// relaxations made earlier in a sweep are seen later in the same sweep, and the wave crosses the lattice in
// one sweep instead of one column per sweep when the warps keep pace.  Measured on cfg3 (65536 growth states):
// 0.5 ms for everything but the sweeps, 0.55 ms per sweep, about 10 sweeps until nothing changes -- shortest
// paths wind through the zero-cost clusters in every direction, so a block barrier after every 256 edges (strict
// column order) reached the east side in one sweep but converged no sooner and cost 0.9 ms per sweep; not kept.  The unclustered edge fraction is one weighted sum over the edges; the
// clustered node count marks the ends of fully grown edges in a bitmap (every fully grown edge has been merged;
// a burnt edge, dynamic UF, was not).
#pragma once
#include <stdint.h>

namespace luf {

struct DcsView {
    int N, E, WN;
    const uint32_t* uv;          // [E] u | v << 16
    const double* weight;        // [E]
    const uint8_t* in_growth;    // [E] edge is a key of decoder.growth (counts towards the edge fraction)
    const uint8_t* side;         // [N] 1 west boundary, 2 east boundary
    const uint4* rec;            // [E_relax] the real edges sorted by their hop distance from the west boundary:
    int E_relax;                 //           x = u | v << 16, y = edge id, (z, w) = the bits of its weight
    double total_weight;         // _total_edge_weight (:101-104), summed on the host in edge order
};

static const unsigned long long DCS_INF = 0x7FF0000000000000ull;

__device__ __forceinline__ double dcs_block_sum(double x, double* scratch)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = x;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w];
    return t;
}

// growth2b: [nshots][stride] words of 2-bit fields (0 ungrown, 1 half, 2 full, 3 burnt); edges beyond
// 16 * growth_words are ungrown.  Outputs are optional.
__global__ void __launch_bounds__(256) k_dcs(DcsView g, long long nshots, const uint32_t* __restrict__ growth2b,
                                             long long stride, int growth_words, int max_passes, double* __restrict__ swim,
                                             double* __restrict__ uef, int32_t* __restrict__ clustered)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* dist = (unsigned long long*)smem;                 // [N]
    uint32_t* mark = (uint32_t*)(smem + 8u * (size_t)g.N);                // [WN]
    uint32_t* gw = mark + g.WN;                                           // [ceil(E / 16)]
    __shared__ double scratch[16];
    __shared__ unsigned long long best;
    __shared__ int cnt;
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x, GW = (g.E + 15) >> 4;
    for (long long k = blockIdx.x; k < nshots; k += gridDim.x) {
        const uint32_t* row = growth2b + k * stride;
        for (int w = tid; w < GW; w += nt) gw[w] = w < growth_words ? __ldg(&row[w]) : 0u;
        for (int v = tid; v < g.N; v += nt) dist[v] = __ldg(&g.side[v]) == 1 ? 0ull : DCS_INF;
        for (int w = tid; w < g.WN; w += nt) mark[w] = 0u;
        if (tid == 0) { best = DCS_INF; cnt = 0; }
        __syncthreads();
        // one pass over the edges: the weighted sum and the clustered nodes
        double num = 0.0;
        for (int e = tid; e < g.E; e += nt) {
            const uint32_t c = (gw[e >> 4] >> ((e & 15) * 2)) & 3u;
            if (c && __ldg(&g.in_growth[e])) num += __ldg(&g.weight[e]) * (c == 1u ? 0.5 : 1.0);
            if (c == 2u) {                   // a burnt edge joins nothing: the two ends may be singleton boundary nodes
                const uint32_t p = __ldg(&g.uv[e]), u = p & 0xFFFFu, v = p >> 16;
                atomicOr(&mark[u >> 5], 1u << (u & 31)); atomicOr(&mark[v >> 5], 1u << (v & 31));
            }
        }
        __syncthreads();
        if (swim) {
            for (int it = 0; it < max_passes; ++it) {
                int changed = 0;
                for (int i = tid; i < g.E_relax; i += nt) {
                    const uint4 r = __ldg(&g.rec[i]);                      // one 16-byte load per edge and sweep
                    const uint32_t u = r.x & 0xFFFFu, v = r.x >> 16, e = r.y;
                    const unsigned long long bu = dist[u], bv = dist[v];
                    if (bu == bv) continue;                                // both unreached, or nothing to gain
                    const uint32_t c = (gw[e >> 4] >> ((e & 15) * 2)) & 3u;
                    const double w = __hiloint2double((int)r.w, (int)r.z);
                    const double cost = c == 0u ? w : c == 1u ? w / 2 : 0.0;
                    if (bu < bv) {
                        const double x = __longlong_as_double((long long)bu) + cost;
                        if (x < __longlong_as_double((long long)bv)) { atomicMin(&dist[v], (unsigned long long)__double_as_longlong(x)); changed = 1; }
                    } else {
                        const double x = __longlong_as_double((long long)bv) + cost;
                        if (x < __longlong_as_double((long long)bu)) { atomicMin(&dist[u], (unsigned long long)__double_as_longlong(x)); changed = 1; }
                    }
                }
                if (!__syncthreads_or(changed)) break;
            }
            unsigned long long mine = DCS_INF;
            for (int v = tid; v < g.N; v += nt)
                if (__ldg(&g.side[v]) == 2 && dist[v] < mine) mine = dist[v];
            if (mine != DCS_INF) atomicMin(&best, mine);
        }
        if (clustered) {
            int c = 0;
            for (int w = tid; w < g.WN; w += nt) c += __popc(mark[w]);
            if (c) atomicAdd(&cnt, c);
        }
        const double total = uef ? dcs_block_sum(num, scratch) : 0.0;
        __syncthreads();
        if (tid == 0) {
            if (swim) swim[k] = __longlong_as_double((long long)best);
            if (uef) uef[k] = 1 - total / g.total_weight;
            if (clustered) clustered[k] = cnt;
        }
        __syncthreads();
    }
}

}  // namespace luf
