// luf_kernels.cu -- CUDA kernels (sm_100a) and C ABI of the B200 batched
// Union-Find decoding engine.  See include/luf_b200.h for the interface and
// luf_core.cuh for the per-shot phases.
//
// Execution model: a tile of lanes -- a warp, a part of a block on a named
// barrier, or a whole block -- owns one shot at a time; the grid is persistent
// (a multiple of the SM count) and the tiles stride over the shot range.  All
// decoder state of a shot lives in the tile's slice of dynamic shared memory
// (the wide instance may bind it to an L2-resident slab of global memory
// instead); the read-only graph tables sit in global memory and are served by
// L1/L2, or are staged once per block in shared memory (compact fast path,
// Macar tiles).  There is no tensor-core work in this path (no dense
// contraction) and nothing to tile with TMA: HBM sees only the bit-packed
// inputs/outputs of the staged kernels, and nothing at all in the fused
// Monte-Carlo kernels.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <new>
#include <string>
#include <vector>

#include "luf_host.h"
#include "luf_dcs.cuh"

using namespace luf;

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

static int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            return fail(LUF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ------------------------------------------------------------------ kernels
static const int WARP = 32;
#define LUF_TILE_SYNC(tile, tw) asm volatile("bar.sync %0, %1;" ::"r"(1u + (unsigned)(tile)), "r"(tw) : "memory")

__device__ __forceinline__ unsigned char* warp_slice(const Layout& l)
{
    extern __shared__ __align__(16) unsigned char smem[];
    return smem + (size_t)(threadIdx.x >> 5) * l.bytes;
}

// Decoder flags whose fused kernels take the order-free merge (luf_uf_phases.inl): not dynamic UF (which merges
// are burnt depends on their order), not BUF (vision-length buckets have their own round structure).
#define LUF_FAST_TF(TF) ((TF) >= 0 && !((TF) & FLAG_DYNAMIC) && (((TF) & (FLAG_BUCKET | FLAG_NODE)) != FLAG_BUCKET))

// Fused K1 + K2 + K4 for the UF decoder, batch scheme (_schemes.py:116-155).
// QUEUE: the second launch behind k_mc_uf_fast -- the shots listed in queue[0 .. *qcount) (indices relative to
// shot0), all in canonical order with the full peel where a cluster spans both sides; the fast launch counted
// the shots already.
template <int TF, bool QUEUE>
__global__ void __launch_bounds__(448, 2) k_mc_uf(GraphView g, Layout l, SampleParams sp, int flags,
                        unsigned long long shot0, long long nshots, unsigned long long* counts,
                        uint8_t* __restrict__ logical, const uint32_t* __restrict__ queue,
                        const uint32_t* __restrict__ qcount)
{
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const Shot s = bind(l, warp_slice(l));
    const long long stride = (long long)gridDim.x * wpb;
    unsigned int fails = 0;
    if (QUEUE) nshots = (long long)*qcount;
    for (long long i = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); i < nshots; i += stride) {
        const long long k = QUEUE ? (long long)queue[i] : i;
        uint32_t parity = 0;
        LUF_PHASE(ph_reset(g, l, s, lane, nl));
        LUF_PHASE(parity ^= ph_sample(g, s, sp, shot0 + (unsigned long long)k, lane, nl));
        LUF_PHASE(ph_load(g, s, lane, nl));
        if (!QUEUE && LUF_FAST_TF(TF) && g.parity_ok) {                   // static UF / NodeUF / NodeBUF: merges in any order, no peel ...
            uf_validate_fast<(LUF_FAST_TF(TF) ? TF : 0)>(g, s, lane, nl);
            if (__any_sync(0xffffffffu, s.cnt[4] != 0u)) { // ... unless a cluster spans both sides: canonical order
                parity = 0;
                __syncwarp();
                LUF_PHASE(ph_reset(g, l, s, lane, nl));
                LUF_PHASE(parity ^= ph_sample(g, s, sp, shot0 + (unsigned long long)k, lane, nl));
                LUF_PHASE(ph_load(g, s, lane, nl));
                uf_validate_t<TF>(g, s, flags, lane, nl);
                parity ^= uf_peel_parity(g, s, lane, nl);      // (no peel at all on graphs whose boundary nodes are leaves)
            } else {
                parity ^= ph_defect_sides(g, s, lane, nl);
                __syncwarp();
            }
        } else {
            uf_validate_t<TF>(g, s, flags, lane, nl);      // TF >= 0: the flags are compile-time constants
            parity ^= uf_peel_parity(g, s, lane, nl);
        }
        const uint32_t bit = LUF_XOR(parity) & 1u;
        fails += bit;
        if (logical && lane == 0) logical[k] = (uint8_t)bit;          // per-shot outcome (_schemes.py:143-155), optional
    }
    if (lane == 0 && fails) atomicAdd(&counts[0], (unsigned long long)fails);
    if (!QUEUE && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)nshots);
}

// The compact fast path (luf_uf_fast.cuh): static UF, default or west inclination.  A block holds the
// adjacency table (ADJ_SHARED) and `blockDim.x / TW` shots, one per tile of TW threads: a part of a warp
// (several shots advance through the same instructions), a warp, or TW / 32 warps synchronised by a named
// barrier.  A shot the path cannot finish (HARD) is appended to `queue` for the canonical launch.
template <int TW>
struct TileSync {
    uint32_t x;          // lane mask of the tile (TW <= 32) or its named barrier (TW > 32; 0 = the whole block)
    __device__ __forceinline__ void operator()() const
    {
        if (TW <= 32) __syncwarp(x);
        else if (x == 0u) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(x), "n"(TW) : "memory");
    }
};

// LOGQ: the second pass, over the HARD shots the first one queued (`qin`, *qin_count): the same decode with
// the merge log (luf_uf_fast.cuh); the merges inside the odd clusters on both sides are exported for
// k_hard_resolve (x.queue), and only a shot whose log or export did not fit goes on to `queue` -- the canonical launch.
// (48 registers: two blocks of 640 threads -- five shots of 128 lanes, the d = 17 plan -- have to fit an SM's register
// file, which is allocated in units of 8 per thread: with 50 the second block did not, 10.4 -> 6.2 M shots/s there;
// the warp tile of cfg3 is 8 % slower when capped -- 38 registers instead of 46 -- and keeps the 64 of its launch shape)
// LOGQ = 2: one pass -- every shot keeps its merge log in a slab of global memory (`glog`, lcap words per resident
// tile: written once per newly FULL edge, never read unless the shot turns out HARD); a HARD shot exports the keys
// of its clusters as they lie, k_hard_sort orders them.  Chosen at noise levels where HARD shots are common.
template <int TW, bool ADJ_SHARED, int LOGQ>
__global__ void __maxnreg__(TW >= 64 ? 48 : 64) k_mc_uf_fast(fast::FView g, fast::FLayout l, uint32_t adj_bytes, SampleParams sp,
                        unsigned long long shot0, long long nshots, unsigned long long* counts,
                        uint8_t* __restrict__ logical, uint32_t* __restrict__ queue, uint32_t* __restrict__ qcount,
                        fast::FExport x, const uint32_t* __restrict__ qin, const uint32_t* __restrict__ qin_count,
                        uint32_t* __restrict__ glog)
{
    extern __shared__ __align__(16) unsigned char smem[];
    if (ADJ_SHARED) {
        const uint4* src = (const uint4*)g.fadj;
        uint4* dst = (uint4*)smem;
        for (uint32_t i = threadIdx.x; i < adj_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
        g.fadj = (const uint32_t*)smem;
    }
    const int tile = (int)threadIdx.x / TW, lane = (int)threadIdx.x % TW, tpb = (int)blockDim.x / TW;
    fast::FShot s = fast::fbind(l, smem + adj_bytes + (size_t)tile * l.bytes);
    if (LOGQ == 2) s.elog = glog + ((size_t)blockIdx.x * tpb + tile) * (size_t)l.lcap;
    TileSync<TW> sync;
    if (TW <= 32) sync.x = TW == 32 ? 0xffffffffu : ((1u << (TW & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(TW - 1));
    else sync.x = tpb == 1 ? 0u : (uint32_t)(1 + tile);
    unsigned int fails = 0;
    if (LOGQ == 1) nshots = (long long)*qin_count;
    for (long long i = (long long)blockIdx.x * tpb + tile; i < nshots; i += (long long)gridDim.x * tpb) {
        const long long k = LOGQ == 1 ? (long long)qin[i] : i;
        const uint32_t r = fast::fast_shot<ADJ_SHARED>(g, l, s, sp, shot0 + (unsigned long long)k, lane, TW, sync, nullptr,
                                                       LOGQ ? &x : nullptr);
        if (lane == 0) {
            if (r == fast::HARD) queue[atomicAdd(qcount, 1u)] = (uint32_t)k;
            else if (r == fast::HARD_LOGGED) {
                const uint32_t pos = atomicAdd(x.qcount, 1u);
                x.queue[2u * pos] = (uint32_t)k; x.queue[2u * pos + 1u] = s.cnt[fast::C_XBASE];
            } else { fails += r; if (logical) logical[k] = (uint8_t)r; }
        }
    }
    if (lane == 0 && fails) atomicAdd(&counts[0], (unsigned long long)fails);
    if (LOGQ != 1 && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)nshots);
}

// One-pass mode: the exported keys of every queued shot, sorted in place into endpoint pairs (fast::hard_sort).
// One block per shot at a time; shared memory: keycap words.
__global__ void __launch_bounds__(256) k_hard_sort(fast::FView g, fast::FExport x)
{
    extern __shared__ __align__(16) unsigned char smem[];
    TileSync<256> sync;
    sync.x = 0u;
    const uint32_t n = *x.qcount;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        fast::hard_sort(g, x.buf + x.queue[2u * i + 1u], (uint32_t*)smem, (int)threadIdx.x, 256, sync);
        __syncthreads();
    }
}

// The exported HARD clusters of a launch: one warp per shot replays the shot's unions (fast::hard_resolve).
// Shared memory per warp: one word per detector and the result word.
__global__ void k_hard_resolve(int D, fast::FExport x, unsigned long long* counts, uint8_t* __restrict__ logical)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    uint32_t* S = (uint32_t*)smem + (size_t)(threadIdx.x >> 5) * (size_t)(D + 4);
    uint32_t* out = S + D;
    TileSync<32> sync;
    sync.x = 0xffffffffu;
    unsigned int fails = 0;
    const uint32_t n = *x.qcount;
    for (uint32_t i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const uint32_t k = x.queue[2u * i], base = x.queue[2u * i + 1u];
        const uint32_t bit = fast::hard_resolve(x.buf + base, S, out, lane, WARP, sync);
        if (lane == 0) { fails += bit; if (logical) logical[k] = (uint8_t)bit; }
        __syncwarp();
    }
    if (lane == 0 && fails) atomicAdd(&counts[0], (unsigned long long)fails);
}

// Staged K2: syndromes in, corrections (+ growth, rounds) out (uf.py:106-119).
__global__ void k_decode_uf(GraphView g, Layout l, int flags, long long nshots,
                            const uint32_t* __restrict__ syn, uint32_t* __restrict__ corr,
                            uint32_t* __restrict__ growth2b, int32_t* __restrict__ steps)
{
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const Shot s = bind(l, warp_slice(l));
    const long long stride = (long long)gridDim.x * wpb;
    const int GW = (g.E + 15) >> 4;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        LUF_PHASE(ph_reset(g, l, s, lane, nl));
        LUF_PHASE(for (int w = lane; w < g.WD; w += nl) s.defect[w] = __ldg(&syn[k * g.WD + w]));
        LUF_PHASE(ph_load(g, s, lane, nl));
        const int rounds = uf_validate(g, s, flags, lane, nl);
        if (growth2b)
            for (int w = lane; w < GW; w += nl) growth2b[k * GW + w] = s.growth[w];
        uf_peel(g, s, lane, nl);
        for (int w = lane; w < g.WE; w += nl) corr[k * g.WE + w] = s.corr[w];
        if (steps && lane == 0) { steps[2 * k] = rounds; steps[2 * k + 1] = 0; }
        __syncwarp();
    }
}

// The same two kernels with one thread block per shot (luf::uf_cta): graphs
// whose shot state leaves only a few warps per SM (Surface d >= 13 with
// measurement rounds).  The first warp samples, so that a shot's noise does not
// depend on the tile shape; every other phase runs on all threads of the block.
template <int TF, bool QUEUE>
__global__ void __launch_bounds__(448, 2) k_mc_uf_cta(GraphView g, Layout l, SampleParams sp, int flags,
                        unsigned long long shot0, long long nshots, unsigned long long* counts,
                        uint8_t* __restrict__ logical, const uint32_t* __restrict__ queue,
                        const uint32_t* __restrict__ qcount)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x, nl = blockDim.x;
    const Shot s = bind(l, smem);
    unsigned int fails = 0;
    if (QUEUE) nshots = (long long)*qcount;
    for (long long i = blockIdx.x; i < nshots; i += gridDim.x) {
        const long long k = QUEUE ? (long long)queue[i] : i;
        uint32_t parity = 0;
        uf_cta::ph_reset(g, l, s, lane, nl);
        __syncthreads();
        if (lane < WARP) parity ^= ph_sample(g, s, sp, shot0 + (unsigned long long)k, lane, WARP);
        __syncthreads();
        uf_cta::ph_load(g, s, lane, nl);
        __syncthreads();
        if (!QUEUE && LUF_FAST_TF(TF) && g.parity_ok) {
            uf_cta::uf_validate_fast<(LUF_FAST_TF(TF) ? TF : 0)>(g, s, lane, nl);
            if (__syncthreads_or(s.cnt[4] != 0u)) {
                parity = 0;
                uf_cta::ph_reset(g, l, s, lane, nl);
                __syncthreads();
                if (lane < WARP) parity ^= ph_sample(g, s, sp, shot0 + (unsigned long long)k, lane, WARP);
                __syncthreads();
                uf_cta::ph_load(g, s, lane, nl);
                __syncthreads();
                uf_cta::uf_validate_t<TF>(g, s, flags, lane, nl);
                parity ^= uf_cta::uf_peel_parity(g, s, lane, nl);
            } else {
                parity ^= uf_cta::ph_defect_sides(g, s, lane, nl);
                __syncthreads();
            }
        } else {
            uf_cta::uf_validate_t<TF>(g, s, flags, lane, nl);
            parity ^= uf_cta::uf_peel_parity(g, s, lane, nl);
        }
        const unsigned int bit = (unsigned int)__syncthreads_count((int)(parity & 1u)) & 1u;
        fails += bit;
        if (logical && threadIdx.x == 0) logical[k] = (uint8_t)bit;
    }
    if (threadIdx.x == 0 && fails) atomicAdd(&counts[0], (unsigned long long)fails);
    if (!QUEUE && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)nshots);
}

__global__ void __launch_bounds__(448, 2) k_decode_uf_cta(GraphView g, Layout l, int flags, long long nshots,
                            const uint32_t* __restrict__ syn, uint32_t* __restrict__ corr,
                            uint32_t* __restrict__ growth2b, int32_t* __restrict__ steps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x, nl = blockDim.x;
    const Shot s = bind(l, smem);
    const int GW = (g.E + 15) >> 4;
    for (long long k = blockIdx.x; k < nshots; k += gridDim.x) {
        uf_cta::ph_reset(g, l, s, lane, nl);
        __syncthreads();
        for (int w = lane; w < g.WD; w += nl) s.defect[w] = __ldg(&syn[k * g.WD + w]);
        __syncthreads();
        uf_cta::ph_load(g, s, lane, nl);
        __syncthreads();
        const int rounds = uf_cta::uf_validate_t<-1>(g, s, flags, lane, nl);
        if (growth2b)
            for (int w = lane; w < GW; w += nl) growth2b[k * GW + w] = s.growth[w];
        uf_cta::uf_peel(g, s, lane, nl);
        for (int w = lane; w < g.WE; w += nl) corr[k * g.WE + w] = s.corr[w];
        if (steps && lane == 0) { steps[2 * k] = rounds; steps[2 * k + 1] = 0; }
        __syncthreads();
    }
}

// Staged K3a: Macar / Actis (decoders/luf/main.py:83-166): syndromes in;
// correction, growth after validation and (tSV, tBP) out.
template <bool ACTIS>
__global__ void k_decode_ca(CaView g, CaLayout l, int flags, long long nshots,
                            const uint32_t* __restrict__ syn, uint32_t* __restrict__ corr,
                            uint32_t* __restrict__ growth2b, int32_t* __restrict__ steps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const CaShot s = ca_bind(l, smem + (size_t)(threadIdx.x >> 5) * l.bytes);
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        LUF_PHASE(ca_reset<ACTIS>(g, s, lane, nl));
        LUF_PHASE(for (int w = lane; w < g.WD; w += nl) s.defect[w] = __ldg(&syn[k * g.WD + w]));
        LUF_PHASE(ca_load(g, s, lane, nl));
        int tsv, tbp;
        ca_decode<ACTIS>(g, s, flags, &tsv, &tbp, growth2b ? growth2b + k * g.GW : nullptr, lane, nl);
        for (int w = lane; w < g.WE; w += nl) corr[k * g.WE + w] = s.corr[w];
        if (steps && lane == 0) { steps[2 * k] = tsv; steps[2 * k + 1] = tbp; }
        __syncwarp();
    }
}

// Fused K1 + K3a + K4 (batch scheme): counts[0] += failures, [1] += shots,
// [2] += sum of tSV, [3] += sum of tBP.
template <bool ACTIS>
__global__ void k_mc_ca(GraphView gv, CaView g, CaLayout l, SampleParams sp, int flags,
                        unsigned long long shot0, long long nshots, unsigned long long* counts,
                        uint8_t* __restrict__ logical)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const CaShot s = ca_bind(l, smem + (size_t)(threadIdx.x >> 5) * l.bytes);
    Shot samp;                       // the sampler only touches the error / defect bitmaps
    samp.err = nullptr; samp.defect = s.defect;
    const long long stride = (long long)gridDim.x * wpb;
    unsigned int fails = 0;
    unsigned long long sum_sv = 0, sum_bp = 0;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        uint32_t parity = 0;
        LUF_PHASE(ca_reset<ACTIS>(g, s, lane, nl));
        LUF_PHASE(parity ^= ph_sample(gv, samp, sp, shot0 + (unsigned long long)k, lane, nl));
        LUF_PHASE(ca_load(g, s, lane, nl));
        int tsv, tbp;
        parity ^= ca_decode<ACTIS>(g, s, flags, &tsv, &tbp, nullptr, lane, nl);
        const uint32_t bit = LUF_XOR(parity) & 1u;
        fails += bit;
        if (logical && lane == 0) logical[k] = (uint8_t)bit;
        sum_sv += (unsigned long long)tsv; sum_bp += (unsigned long long)tbp;
    }
    if (lane == 0) {
        if (fails) atomicAdd(&counts[0], (unsigned long long)fails);
        atomicAdd(&counts[2], sum_sv); atomicAdd(&counts[3], sum_bp);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)nshots);
}

// The same two kernels with one thread block per shot (luf::ca_cta).
template <bool ACTIS>
__global__ void k_decode_ca_cta(CaView g, CaLayout l, int flags, long long nshots,
                                const uint32_t* __restrict__ syn, uint32_t* __restrict__ corr,
                                uint32_t* __restrict__ growth2b, int32_t* __restrict__ steps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = (int)threadIdx.x, nl = (int)blockDim.x;
    const CaShot s = ca_bind(l, smem);
    for (long long k = blockIdx.x; k < nshots; k += gridDim.x) {
        ca_reset<ACTIS>(g, s, lane, nl);
        __syncthreads();
        for (int w = lane; w < g.WD; w += nl) s.defect[w] = __ldg(&syn[k * g.WD + w]);
        __syncthreads();
        ca_load(g, s, lane, nl);
        __syncthreads();
        int tsv, tbp;
        ca_cta::ca_decode<ACTIS>(g, s, flags, &tsv, &tbp, growth2b ? growth2b + k * g.GW : nullptr, lane, nl);
        for (int w = lane; w < g.WE; w += nl) corr[k * g.WE + w] = s.corr[w];
        if (steps && lane == 0) { steps[2 * k] = tsv; steps[2 * k + 1] = tbp; }
        __syncthreads();
    }
}

template <bool ACTIS>
__global__ void k_mc_ca_cta(GraphView gv, CaView g, CaLayout l, SampleParams sp, int flags,
                            unsigned long long shot0, long long nshots, unsigned long long* counts,
                            uint8_t* __restrict__ logical)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = (int)threadIdx.x, nl = (int)blockDim.x;
    const CaShot s = ca_bind(l, smem);
    Shot samp;
    samp.err = nullptr; samp.defect = s.defect;
    unsigned int fails = 0;
    unsigned long long sum_sv = 0, sum_bp = 0;
    for (long long k = blockIdx.x; k < nshots; k += gridDim.x) {
        uint32_t parity = 0;
        ca_reset<ACTIS>(g, s, lane, nl);
        __syncthreads();
        // 32 sampling lanes whatever the tile: a shot's noise does not depend on the tile shape
        if (lane < 32) parity ^= ph_sample(gv, samp, sp, shot0 + (unsigned long long)k, lane, 32);
        __syncthreads();
        ca_load(g, s, lane, nl);
        __syncthreads();
        int tsv, tbp;
        parity ^= ca_cta::ca_decode<ACTIS>(g, s, flags, &tsv, &tbp, nullptr, lane, nl);
        const unsigned int bit = (unsigned int)__syncthreads_count((int)(parity & 1u)) & 1u;
        fails += bit;
        if (logical && lane == 0) logical[k] = (uint8_t)bit;
        sum_sv += (unsigned long long)tsv; sum_bp += (unsigned long long)tbp;
    }
    if (lane == 0) {
        if (fails) atomicAdd(&counts[0], (unsigned long long)fails);
        atomicAdd(&counts[2], sum_sv); atomicAdd(&counts[3], sum_bp);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)nshots);
}

// Macar, several shots per block next to ONE copy of the neighbour table in shared memory (35 KB for cfg3: in
// global memory it hit L1 55 % of the time under the shot states, ncu r01 v13: long-scoreboard stalls 4.6 per
// issue).  A tile of `tw` threads per shot, named barrier 1 + tile (luf::ca_tile); the first warp of a tile samples.
__global__ void __launch_bounds__(1024, 1) k_mc_ca_tiles(GraphView gv, CaView g, CaLayout l, uint32_t table_bytes, int tw, SampleParams sp, int flags,
                              unsigned long long shot0, long long nshots, unsigned long long* counts,
                              uint8_t* __restrict__ logical)
{
    extern __shared__ __align__(16) unsigned char smem[];
    {
        const uint4* src = (const uint4*)g.nbr;
        uint4* dst = (uint4*)smem;
        for (uint32_t i = threadIdx.x; i < table_bytes / 16u; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
        g.nbr = (const uint32_t*)smem; g.nbr_shared = 1;
    }
    const int tile = (int)threadIdx.x / tw, lane = (int)threadIdx.x % tw, nl = tw, tpb = (int)blockDim.x / tw;
    const CaShot s = ca_bind(l, smem + table_bytes + (size_t)tile * l.bytes);
    Shot samp;
    samp.err = nullptr; samp.defect = s.defect;
    unsigned int fails = 0;
    unsigned long long sum_sv = 0, sum_bp = 0;
    uint32_t* red = s.glob + 3;                      // the tile's parity word (idle under Macar)
    for (long long k = (long long)blockIdx.x * tpb + tile; k < nshots; k += (long long)gridDim.x * tpb) {
        uint32_t parity = 0;
        ca_reset<false>(g, s, lane, nl);
        LUF_TILE_SYNC(tile, tw);
        if (lane < 32) parity ^= ph_sample(gv, samp, sp, shot0 + (unsigned long long)k, lane, 32);
        LUF_TILE_SYNC(tile, tw);
        ca_load(g, s, lane, nl);
        LUF_TILE_SYNC(tile, tw);
        int tsv, tbp;
        parity ^= ca_tile::ca_decode<false>(g, s, flags, &tsv, &tbp, nullptr, lane, nl);
        parity = __reduce_xor_sync(0xffffffffu, parity & 1u);
        if ((lane & 31) == 0 && parity) atomicXor(red, 1u);
        LUF_TILE_SYNC(tile, tw);
        const unsigned int bit = *red & 1u;
        LUF_TILE_SYNC(tile, tw);                     // (ca_reset clears the word for the next shot)
        fails += bit;
        if (logical && lane == 0) logical[k] = (uint8_t)bit;
        sum_sv += (unsigned long long)tsv; sum_bp += (unsigned long long)tbp;
    }
    if (lane == 0) {
        if (fails) atomicAdd(&counts[0], (unsigned long long)fails);
        atomicAdd(&counts[2], sum_sv); atomicAdd(&counts[3], sum_bp);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)nshots);
}

// Forward scheme (_schemes.py:392-450) around UF (MACAR = false) or Macar.
struct FwArgs {
    GraphView gv; Layout ul; CaView cv; CaLayout cl; FwView fw; FwLayout fl; SampleTab buffer_tab;
    int flags, n_cleanse;
};

template <bool MACAR, class F>
__device__ __forceinline__ void fw_with_decoder(const FwArgs& a, unsigned char* slice, F&& body)
{
    if constexpr (MACAR) { FwMacar d{a.cv, ca_bind(a.cl, slice)}; body(d); }
    else { FwUf d{a.gv, a.ul, bind(a.ul, slice), a.flags}; body(d); }
}

template <bool MACAR>
__global__ void k_fw_decode(FwArgs a, long long nstreams, int ncycles, const uint32_t* __restrict__ init_buf,
                            const uint32_t* __restrict__ fresh, int32_t* __restrict__ steps,
                            uint32_t* __restrict__ commit_out, int32_t* __restrict__ m_out)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    unsigned char* slice = smem + (size_t)(threadIdx.x >> 5) * a.fl.bytes;
    const FwState f = fw_bind(a.fl, slice);
    const long long stride = (long long)gridDim.x * wpb;
    const int WE = a.fw.WE;
    fw_with_decoder<MACAR>(a, slice, [&](auto& d) {
        for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nstreams; k += stride) {
            LUF_PHASE(fw_reset(a.fw, f, init_buf + (size_t)k * WE, lane, nl));
            for (int c = 0; c < ncycles; ++c) {
                const size_t row = (size_t)k * ncycles + c;
                int s0 = 0, s1 = 0;
                const int m = fw_cycle_given(a.fw, f, d, fresh + row * WE, commit_out ? commit_out + row * WE : nullptr,
                                             s0, s1, lane, nl);
                if (lane == 0) { steps[2 * row] = s0; steps[2 * row + 1] = s1; m_out[row] = m; }
            }
        }
    });
}

template <bool MACAR>
__global__ void k_fw_mc(FwArgs a, SampleParams sp, unsigned long long stream0, long long nstreams, int n,
                        unsigned long long* counts, int32_t* steps)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    unsigned char* slice = smem + (size_t)(threadIdx.x >> 5) * a.fl.bytes;
    const FwState f = fw_bind(a.fl, slice);
    const long long stride = (long long)gridDim.x * wpb;
    int m_total = 0;
    long long cycles = 0, s0 = 0, s1 = 0;
    fw_with_decoder<MACAR>(a, slice, [&](auto& d) {
        for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nstreams; k += stride)
            fw_run(a.gv, a.fw, f, d, a.buffer_tab, sp, stream0 + (unsigned long long)k, n, a.n_cleanse, m_total, cycles,
                   s0, s1, steps ? steps + (size_t)k * n * 2 : nullptr, lane, nl);
    });
    if (lane == 0) {
        atomicAdd(&counts[0], (unsigned long long)m_total); atomicAdd(&counts[1], (unsigned long long)cycles);
        atomicAdd(&counts[2], (unsigned long long)s0); atomicAdd(&counts[3], (unsigned long long)s1);
    }
}

// K3b parity path: Snowflake + frugal bookkeeping on GIVEN fresh errors.
__global__ void k_sf_decode(SfView g, SfLayout l, int flags, long long nstreams, int ncycles,
                            const uint32_t* __restrict__ fresh, int32_t* __restrict__ rt,
                            uint32_t* __restrict__ floor_bits, int32_t* __restrict__ m_out,
                            uint32_t* __restrict__ mgrowth, int32_t* __restrict__ mins, const uint8_t* __restrict__ drop_only)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const SfShot s = sf_bind(l, smem + (size_t)(threadIdx.x >> 5) * l.bytes);
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nstreams; k += stride) {
        const size_t row0 = (size_t)k * ncycles;
        sf_warp::sf_decode_stream(g, s, flags, ncycles, fresh + row0 * g.ELw, rt + 2 * row0,
                                  floor_bits ? floor_bits + row0 * g.ELw : nullptr, m_out + row0,
                                  mgrowth ? mgrowth + row0 * (size_t)(g.h * 2 * g.ELw) : nullptr, mins + 2 * row0, drop_only ? drop_only + row0 : nullptr, lane, nl);
    }
}

// Fused K1 + K3b + K4: one Frugal.run(n) memory experiment per warp.
__global__ void k_sf_mc(SfView g, SfLayout l, SampleParams sp, int flags, unsigned long long stream0,
                        long long nstreams, int n, unsigned long long* counts, int32_t* steps,
                        uint32_t* __restrict__ mgrowth, int32_t* __restrict__ mcycle)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const SfShot s = sf_bind(l, smem + (size_t)(threadIdx.x >> 5) * l.bytes);
    const long long stride = (long long)gridDim.x * wpb;
    int m_total = 0;
    long long cycles = 0, ra = 0, ru = 0;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nstreams; k += stride)
        sf_run(g, s, sp, flags, stream0 + (unsigned long long)k, n, m_total, cycles, ra, ru,
               steps ? steps + (size_t)k * n * 2 : nullptr,
               mgrowth ? mgrowth + (size_t)k * n * g.d * (size_t)(g.h * 2 * g.ELw) : nullptr, mcycle + (size_t)k * n * g.d * 4, lane, nl);
    if (lane == 0) {
        atomicAdd(&counts[0], (unsigned long long)m_total); atomicAdd(&counts[1], (unsigned long long)cycles);
        atomicAdd(&counts[2], (unsigned long long)ra); atomicAdd(&counts[3], (unsigned long long)ru);
    }
}

// The same two kernels with one thread block per window (luf::sf_cta): large windows.
__global__ void k_sf_decode_cta(SfView g, SfLayout l, int flags, long long nstreams, int ncycles,
                                const uint32_t* __restrict__ fresh, int32_t* __restrict__ rt,
                                uint32_t* __restrict__ floor_bits, int32_t* __restrict__ m_out,
                                uint32_t* __restrict__ mgrowth, int32_t* __restrict__ mins, const uint8_t* __restrict__ drop_only)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = (int)threadIdx.x, nl = (int)blockDim.x;
    const SfShot s = sf_bind(l, smem);
    for (long long k = blockIdx.x; k < nstreams; k += gridDim.x) {
        const size_t row0 = (size_t)k * ncycles;
        sf_cta::sf_decode_stream(g, s, flags, ncycles, fresh + row0 * g.ELw, rt + 2 * row0,
                                 floor_bits ? floor_bits + row0 * g.ELw : nullptr, m_out + row0,
                                 mgrowth ? mgrowth + row0 * (size_t)(g.h * 2 * g.ELw) : nullptr, mins + 2 * row0, drop_only ? drop_only + row0 : nullptr, lane, nl);
    }
}

__global__ void k_sf_mc_cta(SfView g, SfLayout l, SampleParams sp, int flags, unsigned long long stream0,
                            long long nstreams, int n, unsigned long long* counts, int32_t* steps,
                            uint32_t* __restrict__ mgrowth, int32_t* __restrict__ mcycle)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = (int)threadIdx.x, nl = (int)blockDim.x;
    const SfShot s = sf_bind(l, smem);
    int m_total = 0;
    long long cycles = 0, ra = 0, ru = 0;
    for (long long k = blockIdx.x; k < nstreams; k += gridDim.x)
        sf_cta::sf_run(g, s, sp, flags, stream0 + (unsigned long long)k, n, m_total, cycles, ra, ru,
                       steps ? steps + (size_t)k * n * 2 : nullptr,
                       mgrowth ? mgrowth + (size_t)k * n * g.d * (size_t)(g.h * 2 * g.ELw) : nullptr, mcycle + (size_t)k * n * g.d * 4, lane, nl);
    if (lane == 0) {
        atomicAdd(&counts[0], (unsigned long long)m_total); atomicAdd(&counts[1], (unsigned long long)cycles);
        atomicAdd(&counts[2], (unsigned long long)ra); atomicAdd(&counts[3], (unsigned long long)ru);
    }
}

// Frugal.sample_latency for nstreams independent memory experiments: lat [nstreams][3].
__global__ void k_sf_latency(SfView g, SfLayout l, SampleParams sp, int flags, unsigned long long stream0,
                             long long nstreams, int32_t* __restrict__ lat)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const SfShot s = sf_bind(l, smem + (size_t)(threadIdx.x >> 5) * l.bytes);
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nstreams; k += stride) {
        int v[3];
        sf_warp::sf_latency(g, s, sp, flags, stream0 + (unsigned long long)k, v, lane, nl);
        if (lane == 0) { lat[3 * k] = v[0]; lat[3 * k + 1] = v[1]; lat[3 * k + 2] = v[2]; }
    }
}

__global__ void k_sf_latency_cta(SfView g, SfLayout l, SampleParams sp, int flags, unsigned long long stream0,
                                 long long nstreams, int32_t* __restrict__ lat)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = (int)threadIdx.x, nl = (int)blockDim.x;
    const SfShot s = sf_bind(l, smem);
    for (long long k = blockIdx.x; k < nstreams; k += gridDim.x) {
        int v[3];
        sf_cta::sf_latency(g, s, sp, flags, stream0 + (unsigned long long)k, v, lane, nl);
        if (lane == 0) { lat[3 * k] = v[0]; lat[3 * k + 1] = v[1]; lat[3 * k + 2] = v[2]; }
    }
}

// Staged K1: bit-packed errors and syndromes to HBM
// (noise/main.py:114-115,309-314; _base_classes.py:427-450).
// A warp takes four consecutive shots at a time: their rows are 4 W(E) and 4 W(D) contiguous words of `err` and
// `syn` -- multiples of 16 bytes whatever W is -- so the warp's slice of shared memory (err[4][WE] | syn[4][WD]) is
// cleared and written out with 16-byte accesses (streaming stores: the rows are not read again by this kernel);
// round 1 cleared and stored word by word, 22 % of the kernel's instructions (ncu r2m).  VEC = false: the
// buffers are not 16-byte aligned.
template <bool VEC>
__global__ void k_sample(GraphView g, SampleParams sp, unsigned long long shot0,
                         long long nshots, uint32_t* __restrict__ err, uint32_t* __restrict__ syn)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const int WE = g.WE, WD = g.WD, words = 4 * (WE + WD);
    uint32_t* slice = (uint32_t*)smem + (size_t)(threadIdx.x >> 5) * words;
    const long long stride = (long long)gridDim.x * wpb, ngroups = (nshots + 3) >> 2;
    for (long long grp = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); grp < ngroups; grp += stride) {
        {
            uint4* z = (uint4*)slice;
            for (int i = lane; i < WE + WD; i += nl) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncwarp();
        const int in_group = (int)(nshots - 4 * grp < 4 ? nshots - 4 * grp : 4);
        for (int j = 0; j < in_group; ++j) {
            Shot s;
            s.err = slice + j * WE; s.defect = slice + 4 * WE + j * WD;
            ph_sample(g, s, sp, shot0 + (unsigned long long)(4 * grp + j), lane, nl);
        }
        __syncwarp();
        uint32_t* eo = err + 4 * grp * WE;
        uint32_t* so = syn + 4 * grp * WD;
        if (VEC && in_group == 4) {
            const uint4* e4 = (const uint4*)slice;
            const uint4* d4 = (const uint4*)(slice + 4 * WE);
            for (int i = lane; i < WE; i += nl) __stcs((uint4*)eo + i, e4[i]);
            for (int i = lane; i < WD; i += nl) __stcs((uint4*)so + i, d4[i]);
        } else {
            for (int i = lane; i < in_group * WE; i += nl) eo[i] = slice[i];
            for (int i = lane; i < in_group * WD; i += nl) so[i] = slice[4 * WE + i];
        }
        __syncwarp();
    }
}

// Staged fixed-weight sampler (subset sampling, noise/main.py:116-122): every
// shot flips exactly `weight` distinct fresh edges.
__global__ void k_sample_weight(GraphView g, Layout l, SampleParams sp, uint32_t weight, unsigned long long shot0,
                                long long nshots, uint32_t* __restrict__ err, uint32_t* __restrict__ syn)
{
    const int lane = threadIdx.x & 31, nl = WARP;
    const int wpb = blockDim.x >> 5;
    const Shot s = bind(l, warp_slice(l));
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        LUF_PHASE(for (int w = lane; w < g.WE; w += nl) s.err[w] = 0;
                  for (int w = lane; w < g.WD; w += nl) s.defect[w] = 0);
        LUF_PHASE(ph_force(g, s, sp, weight, shot0 + (unsigned long long)k, lane));
        for (int w = lane; w < g.WE; w += nl) err[k * g.WE + w] = s.err[w];
        for (int w = lane; w < g.WD; w += nl) syn[k * g.WD + w] = s.defect[w];
        __syncwarp();
    }
}

// Staged K1 on the compact tables: graphs beyond the 16-bit layout (fast::f_sample_record).
__global__ void k_sample_fast(fast::FView g, int WE, uint32_t slice_bytes, SampleParams sp, unsigned long long shot0,
                              long long nshots, uint32_t* __restrict__ err, uint32_t* __restrict__ syn)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    uint32_t* e = (uint32_t*)(smem + (size_t)(threadIdx.x >> 5) * slice_bytes);
    uint32_t* d = e + WE;
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        for (int w = lane; w < WE + g.WD; w += WARP) e[w] = 0;
        __syncwarp();
        fast::f_sample_record(g, e, d, sp, shot0 + (unsigned long long)k, lane, WARP);
        __syncwarp();
        for (int w = lane; w < WE; w += WARP) err[k * WE + w] = e[w];
        for (int w = lane; w < g.WD; w += WARP) syn[k * g.WD + w] = d[w];
        __syncwarp();
    }
}

// Staged K4: leftover parity on the west boundary (_schemes.py:123-129,154-155).
__global__ void k_check(int WE, long long nshots, const uint32_t* __restrict__ west_mask,
                        const uint32_t* __restrict__ err, const uint32_t* __restrict__ corr,
                        uint8_t* __restrict__ logical, unsigned long long* fail_count)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const long long stride = (long long)gridDim.x * wpb;
    unsigned int fails = 0;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        uint32_t acc = 0;
        for (int w = lane; w < WE; w += WARP)
            acc ^= ((err ? __ldg(&err[k * WE + w]) : 0u) ^ __ldg(&corr[k * WE + w])) & __ldg(&west_mask[w]);      // err == nullptr: the frame of the correction alone
        acc = __reduce_xor_sync(0xffffffffu, acc);
        const unsigned int bit = __popc(acc) & 1u;
        if (lane == 0 && logical) logical[k] = (uint8_t)bit;
        fails += bit;
    }
    if (lane == 0 && fails) atomicAdd(fail_count, (unsigned long long)fails);
}

// Staged K4, streaming form: a warp takes four consecutive shots at a time -- 4*WE words of `err`
// and of `corr`, contiguous and 16-byte aligned -- with 16-byte loads that are all in flight before
// the first is used (3.5 KB per warp instead of 0.9 KB: the kernel is bound by how many bytes an SM
// keeps in flight).  A word's shot is its index / WE, by multiplication with `magic` = ceil(2^32 / WE)
// (exact below 2^16 words).  The last nshots % 4 shots go through k_check.
template <int V>        // uint4 loads per lane and array: ceil(WE / 32)
__global__ void __launch_bounds__(256) k_check4(int WE, uint32_t magic, long long ngroups, const uint32_t* __restrict__ west_mask,
                                                const uint4* __restrict__ err, const uint4* __restrict__ corr,
                                                uint8_t* __restrict__ logical, unsigned long long* fail_count)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const long long stride = (long long)gridDim.x * wpb;
    unsigned int fails = 0;
    for (long long grp = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); grp < ngroups; grp += stride) {
        const uint4* e4 = err + grp * WE;       // 4*WE words = WE uint4
        const uint4* c4 = corr + grp * WE;
        uint4 e[V], c[V];
        #pragma unroll
        for (int i = 0; i < V; ++i) {
            const int q = lane + 32 * i;
            if (q < WE) { e[i] = __ldcs(e4 + q); c[i] = __ldcs(c4 + q); }      // streamed once: evict first
            else { e[i] = make_uint4(0, 0, 0, 0); c[i] = e[i]; }
        }
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        #pragma unroll
        for (int i = 0; i < V; ++i) {
            const uint32_t x[4] = {e[i].x ^ c[i].x, e[i].y ^ c[i].y, e[i].z ^ c[i].z, e[i].w ^ c[i].w};
            #pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!x[k]) continue;                                            // errors and corrections are sparse
                const uint32_t j = (uint32_t)(4 * (lane + 32 * i) + k);
                const uint32_t row = __umulhi(j, magic);
                const uint32_t m = x[k] & __ldg(&west_mask[j - row * (uint32_t)WE]);
                a0 ^= row == 0 ? m : 0u; a1 ^= row == 1 ? m : 0u; a2 ^= row == 2 ? m : 0u; a3 ^= row == 3 ? m : 0u;
            }
        }
        a0 = __reduce_xor_sync(0xffffffffu, a0); a1 = __reduce_xor_sync(0xffffffffu, a1);
        a2 = __reduce_xor_sync(0xffffffffu, a2); a3 = __reduce_xor_sync(0xffffffffu, a3);
        const uint32_t bits = (__popc(a0) & 1u) | (__popc(a1) & 1u) << 8 | (__popc(a2) & 1u) << 16 | (__popc(a3) & 1u) << 24;
        if (lane == 0 && logical) *(uint32_t*)(logical + 4 * grp) = bits;       // four result bytes at once
        fails += __popc(bits);
    }
    if (lane == 0 && fails) atomicAdd(fail_count, (unsigned long long)fails);
}

// Global.get_logical_error (_schemes.py:194-203) for a batch of leftovers = err ^ corr: the edges are loaded
// into Pairs in ascending edge id (_pairs.py:49-66) and the strings that join the two sides are counted.  One
// warp per shot: the partner map (one int16 per node; a boundary node has one edge, so the two sides are two
// keys) lives in the warp's shared memory; the lanes fetch 32 leftover words at a time, lane 0 splices their
// edges in order (the pairing where three or more leftover edges meet depends on it), and all lanes clear what
// the shot touched.
static const int STR_NONE = -32768, STR_WEST = -1, STR_EAST = -2;

__device__ __forceinline__ int str_key(const GraphView& g, uint32_t v)
{
    return v >= (uint32_t)g.B ? (int)v : (__ldg(&g.node_long[v]) == -1 ? STR_WEST : STR_EAST);
}
__device__ __forceinline__ void str_add(int16_t* partner, int a, int b, int& count)
{
    if (a >= 0) partner[a] = (int16_t)b;
    if (b >= 0) partner[b] = (int16_t)a;
    if ((a == STR_WEST && b == STR_EAST) || (a == STR_EAST && b == STR_WEST)) ++count;
}
__device__ __forceinline__ int str_take(int16_t* partner, int u)
{
    const int w = partner[u];
    partner[u] = (int16_t)STR_NONE;
    if (w >= 0) partner[w] = (int16_t)STR_NONE;
    return w;
}

__global__ void k_count_strings(GraphView g, long long nshots, const uint32_t* __restrict__ err,
                                const uint32_t* __restrict__ corr, int32_t* __restrict__ per_shot,
                                unsigned long long* total, uint32_t slice_bytes)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    int16_t* partner = (int16_t*)(smem + (size_t)(threadIdx.x >> 5) * slice_bytes);
    for (int v = lane; v < g.N; v += 32) partner[v] = (int16_t)STR_NONE;
    __syncwarp();
    unsigned long long sum = 0;
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        int count = 0;
        for (int base = 0; base < g.WE; base += 32) {
            const int w = base + lane;
            const uint32_t x = w < g.WE ? (__ldg(&err[k * g.WE + w]) ^ __ldg(&corr[k * g.WE + w])) : 0u;
            uint32_t lanes = __ballot_sync(0xffffffffu, x != 0u);
            while (lanes) {                                   // ascending words, ascending bits: ascending edge id
                const int src = __ffs((int)lanes) - 1; lanes &= lanes - 1;
                uint32_t bits = __shfl_sync(0xffffffffu, x, src);
                if (lane == 0) {
                    while (bits) {
                        const int b = __ffs((int)bits) - 1; bits &= bits - 1;
                        const uint32_t uv = __ldg(&g.edge_uv[32 * (base + src) + b]);
                        const int u = str_key(g, uv & 0xFFFFu), v = str_key(g, uv >> 16);
                        if (u >= 0 && partner[u] != STR_NONE) {
                            const int p = str_take(partner, u);
                            if (v >= 0 && partner[v] != STR_NONE) str_add(partner, p, str_take(partner, v), count);
                            else if (v != p || v < 0) str_add(partner, v, p, count);
                        } else if (v >= 0 && partner[v] != STR_NONE)
                            str_add(partner, u, str_take(partner, v), count);
                        else
                            str_add(partner, u, v, count);
                    }
                }
            }
            __syncwarp();
        }
        // put the map back: every entry the shot wrote belongs to an end of a leftover edge
        for (int base = 0; base < g.WE; base += 32) {
            const int w = base + lane;
            uint32_t bits = w < g.WE ? (__ldg(&err[k * g.WE + w]) ^ __ldg(&corr[k * g.WE + w])) : 0u;
            while (bits) {
                const int b = __ffs((int)bits) - 1; bits &= bits - 1;
                const uint32_t uv = __ldg(&g.edge_uv[32 * w + b]);
                if ((uv & 0xFFFFu) >= (uint32_t)g.B) partner[uv & 0xFFFFu] = (int16_t)STR_NONE;
                if ((uv >> 16) >= (uint32_t)g.B) partner[uv >> 16] = (int16_t)STR_NONE;
            }
        }
        __syncwarp();
        if (lane == 0) { if (per_shot) per_shot[k] = count; sum += (unsigned long long)count; }
    }
    if (lane == 0 && sum) atomicAdd(total, sum);
}

// ------------------------------------------------------------ wide instance (luf_uf_wide.cuh)
// Graphs beyond the 16-bit layout: one thread block per shot, 64-bit node words, the shot's state slab in the
// block's dynamic shared memory (slab == nullptr) or in global memory at slab + blockIdx.x * l.bytes -- the
// state-spill tile; a persistent grid keeps the slabs of all blocks in L2.  The first warp samples (32 sampler
// lanes whatever the tile: the random stream of the 16-bit kernels).
__device__ __forceinline__ unsigned char* wide_state(const uf_wide::Layout& l, unsigned char* slab)
{
    extern __shared__ __align__(16) unsigned char smem[];
    return slab ? slab + (size_t)blockIdx.x * l.bytes : smem;
}

// Staged K2 (uf.py:106-119): syndromes in, corrections (+ growth, rounds) out.
__global__ void __launch_bounds__(512, 2) k_decode_uf_wide(uf_wide::GraphView g, uf_wide::Layout l, int flags, long long nshots,
                            const uint32_t* __restrict__ syn, uint32_t* __restrict__ corr,
                            uint32_t* __restrict__ growth2b, int32_t* __restrict__ steps, unsigned char* slab)
{
    const int lane = threadIdx.x, nl = blockDim.x;
    const uf_wide::Shot s = uf_wide::bind(l, wide_state(l, slab));
    const int GW = (g.E + 15) >> 4;
    for (long long k = blockIdx.x; k < nshots; k += gridDim.x) {
        uf_wide::ph_reset(g, l, s, lane, nl);
        __syncthreads();
        for (int w = lane; w < g.WD; w += nl) s.defect[w] = __ldg(&syn[k * g.WD + w]);
        __syncthreads();
        uf_wide::ph_load(g, s, lane, nl);
        __syncthreads();
        const int rounds = uf_wide::uf_validate_t<-1>(g, s, flags, lane, nl);
        if (growth2b)
            for (int w = lane; w < GW; w += nl) growth2b[k * GW + w] = s.growth[w];
        uf_wide::uf_peel(g, s, lane, nl);
        for (int w = lane; w < g.WE; w += nl) corr[k * g.WE + w] = s.corr[w];
        if (steps && lane == 0) { steps[2 * k] = rounds; steps[2 * k + 1] = 0; }
        __syncthreads();
    }
}

// Fused K1 + K2 + K4 (_schemes.py:116-155), canonical order.  QUEUE: the launch behind k_mc_uf_fast (as k_mc_uf).
template <bool QUEUE>
__global__ void __launch_bounds__(512, 2) k_mc_uf_wide(uf_wide::GraphView g, uf_wide::Layout l, SampleParams sp, int flags,
                        unsigned long long shot0, long long nshots, unsigned long long* counts,
                        uint8_t* __restrict__ logical, const uint32_t* __restrict__ queue,
                        const uint32_t* __restrict__ qcount, unsigned char* slab)
{
    const int lane = threadIdx.x, nl = blockDim.x;
    const uf_wide::Shot s = uf_wide::bind(l, wide_state(l, slab));
    unsigned int fails = 0;
    if (QUEUE) nshots = (long long)*qcount;
    for (long long i = blockIdx.x; i < nshots; i += gridDim.x) {
        const long long k = QUEUE ? (long long)queue[i] : i;
        uint32_t parity = 0;
        uf_wide::ph_reset(g, l, s, lane, nl);
        __syncthreads();
        if (lane < WARP) parity ^= uf_wide::ph_sample(g, s, sp, shot0 + (unsigned long long)k, lane, WARP);
        __syncthreads();
        uf_wide::ph_load(g, s, lane, nl);
        __syncthreads();
        uf_wide::uf_validate_t<-1>(g, s, flags, lane, nl);
        parity ^= uf_wide::uf_peel_parity(g, s, lane, nl);
        const unsigned int bit = (unsigned int)__syncthreads_count((int)(parity & 1u)) & 1u;
        fails += bit;
        if (logical && threadIdx.x == 0) logical[k] = (uint8_t)bit;
    }
    if (threadIdx.x == 0 && fails) atomicAdd(&counts[0], (unsigned long long)fails);
    if (!QUEUE && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&counts[1], (unsigned long long)nshots);
}

// Staged K1 on the wide tables: a warp per shot flips the bits of the shot's rows in HBM directly (a row of a
// large graph is tens of KB; its few hundred faults are scattered atomics either way).
template <bool WEIGHT>
__global__ void k_sample_wide(uf_wide::GraphView g, SampleParams sp, uint32_t weight, unsigned long long shot0, long long nshots,
                              uint32_t* __restrict__ err, uint32_t* __restrict__ syn)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        uf_wide::Shot s;
        s.err = err + k * g.WE; s.defect = syn + k * g.WD;
        for (int w = lane; w < g.WE; w += WARP) s.err[w] = 0;
        for (int w = lane; w < g.WD; w += WARP) s.defect[w] = 0;
        __syncwarp();
        if (WEIGHT) uf_wide::ph_force(g, s, sp, weight, shot0 + (unsigned long long)k, lane);
        else uf_wide::ph_sample(g, s, sp, shot0 + (unsigned long long)k, lane, WARP);
        __syncwarp();
    }
}

// Forward scheme (_schemes.py:392-450) around the wide UF instance: one block per stream (luf_fw_wide.cuh).
struct FwWideArgs {
    uf_wide::GraphView gv; uf_wide::Layout ul; fw_wide::FwView fw; fw_wide::FwLayout fl; SampleTab buffer_tab;
    int flags, n_cleanse;
};

__device__ __forceinline__ unsigned char* fw_wide_state(const fw_wide::FwLayout& l, unsigned char* slab)
{
    extern __shared__ __align__(16) unsigned char smem[];
    return slab ? slab + (size_t)blockIdx.x * l.bytes : smem;
}

__global__ void __launch_bounds__(512, 2) k_fw_decode_wide(FwWideArgs a, long long nstreams, int ncycles, const uint32_t* __restrict__ init_buf,
                            const uint32_t* __restrict__ fresh, int32_t* __restrict__ steps,
                            uint32_t* __restrict__ commit_out, int32_t* __restrict__ m_out, unsigned char* slab)
{
    const int lane = threadIdx.x, nl = blockDim.x;
    unsigned char* base = fw_wide_state(a.fl, slab);
    const fw_wide::FwState f = fw_wide::fw_bind(a.fl, base);
    const fw_wide::FwUf d{a.gv, a.ul, uf_wide::bind(a.ul, base), a.flags};
    const int WE = a.fw.WE;
    for (long long k = blockIdx.x; k < nstreams; k += gridDim.x) {
        fw_wide::fw_reset(a.fw, f, init_buf + (size_t)k * WE, lane, nl);
        __syncthreads();
        for (int c = 0; c < ncycles; ++c) {
            const size_t row = (size_t)k * ncycles + c;
            int s0 = 0, s1 = 0;
            const int m = fw_wide::fw_cycle_given(a.fw, f, d, fresh + row * WE, commit_out ? commit_out + row * WE : nullptr,
                                                  s0, s1, lane, nl);
            if (lane == 0) { steps[2 * row] = s0; steps[2 * row + 1] = s1; m_out[row] = m; }
        }
    }
}

__global__ void __launch_bounds__(512, 2) k_fw_mc_wide(FwWideArgs a, SampleParams sp, unsigned long long stream0, long long nstreams, int n,
                        unsigned long long* counts, int32_t* steps, unsigned char* slab)
{
    const int lane = threadIdx.x, nl = blockDim.x;
    unsigned char* base = fw_wide_state(a.fl, slab);
    const fw_wide::FwState f = fw_wide::fw_bind(a.fl, base);
    const fw_wide::FwUf d{a.gv, a.ul, uf_wide::bind(a.ul, base), a.flags};
    int m_total = 0;
    long long cycles = 0, s0 = 0, s1 = 0;
    for (long long k = blockIdx.x; k < nstreams; k += gridDim.x)
        fw_wide::fw_run(a.fw, f, d, a.buffer_tab, sp, stream0 + (unsigned long long)k, n, a.n_cleanse, m_total, cycles,
                        s0, s1, steps ? steps + (size_t)k * n * 2 : nullptr, lane, nl);
    if (lane == 0) {
        atomicAdd(&counts[0], (unsigned long long)m_total); atomicAdd(&counts[1], (unsigned long long)cycles);
        atomicAdd(&counts[2], (unsigned long long)s0); atomicAdd(&counts[3], (unsigned long long)s1);
    }
}

// k_count_strings on the wide tables: the partner map (one int32 per node) of a warp lives in global memory.
static const int WSTR_NONE = INT32_MIN;
__device__ __forceinline__ int wstr_key(const uf_wide::GraphView& g, uint32_t v)
{
    return v >= (uint32_t)g.B ? (int)v : (__ldg(&g.node_long[v]) == -1 ? STR_WEST : STR_EAST);
}
__device__ __forceinline__ void wstr_add(int32_t* partner, int a, int b, int& count)
{
    if (a >= 0) partner[a] = b;
    if (b >= 0) partner[b] = a;
    if ((a == STR_WEST && b == STR_EAST) || (a == STR_EAST && b == STR_WEST)) ++count;
}
__device__ __forceinline__ int wstr_take(int32_t* partner, int u)
{
    const int w = partner[u];
    partner[u] = WSTR_NONE;
    if (w >= 0) partner[w] = WSTR_NONE;
    return w;
}

__global__ void k_count_strings_wide(uf_wide::GraphView g, long long nshots, const uint32_t* __restrict__ err,
                                     const uint32_t* __restrict__ corr, int32_t* __restrict__ per_shot,
                                     unsigned long long* total, int32_t* maps)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    int32_t* partner = maps + ((size_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * (size_t)g.N;
    for (int v = lane; v < g.N; v += 32) partner[v] = WSTR_NONE;
    __syncwarp();
    unsigned long long sum = 0;
    const long long stride = (long long)gridDim.x * wpb;
    for (long long k = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); k < nshots; k += stride) {
        int count = 0;
        for (int base = 0; base < g.WE; base += 32) {
            const int w = base + lane;
            const uint32_t x = w < g.WE ? (__ldg(&err[k * g.WE + w]) ^ __ldg(&corr[k * g.WE + w])) : 0u;
            uint32_t lanes = __ballot_sync(0xffffffffu, x != 0u);
            while (lanes) {                                   // ascending words, ascending bits: ascending edge id
                const int src = __ffs((int)lanes) - 1; lanes &= lanes - 1;
                uint32_t bits = __shfl_sync(0xffffffffu, x, src);
                if (lane == 0) {
                    while (bits) {
                        const int b = __ffs((int)bits) - 1; bits &= bits - 1;
                        uint32_t eu, ev;
                        uf_wide::edge_ends(g, (uint32_t)(32 * (base + src) + b), eu, ev);
                        const int u = wstr_key(g, eu), v = wstr_key(g, ev);
                        if (u >= 0 && partner[u] != WSTR_NONE) {       // Pairs.load, _pairs.py:49-66
                            const int p = wstr_take(partner, u);
                            if (v >= 0 && partner[v] != WSTR_NONE) wstr_add(partner, p, wstr_take(partner, v), count);
                            else if (v != p || v < 0) wstr_add(partner, v, p, count);
                        } else if (v >= 0 && partner[v] != WSTR_NONE)
                            wstr_add(partner, u, wstr_take(partner, v), count);
                        else
                            wstr_add(partner, u, v, count);
                    }
                }
            }
            __syncwarp();
        }
        for (int base = 0; base < g.WE; base += 32) {        // put the map back
            const int w = base + lane;
            uint32_t bits = w < g.WE ? (__ldg(&err[k * g.WE + w]) ^ __ldg(&corr[k * g.WE + w])) : 0u;
            while (bits) {
                const int b = __ffs((int)bits) - 1; bits &= bits - 1;
                uint32_t eu, ev;
                uf_wide::edge_ends(g, (uint32_t)(32 * w + b), eu, ev);
                if (eu >= (uint32_t)g.B) partner[eu] = WSTR_NONE;
                if (ev >= (uint32_t)g.B) partner[ev] = WSTR_NONE;
            }
        }
        __syncwarp();
        if (lane == 0) { if (per_shot) per_shot[k] = count; sum += (unsigned long long)count; }
    }
    if (lane == 0 && sum) atomicAdd(total, sum);
}

// ------------------------------------------------------------------- graph
struct luf_graph {
    int device = 0, sm_count = 0, max_smem = 0;
    PackedGraph host;
    GraphView view{};
    std::vector<void*> allocs;
    Layout lay_fused{}, lay_decode{}, lay_sample{};
    Layout lay_fused_cta{}, lay_decode_cta{};     // one block per shot (luf::uf_cta work-list sizes)
    PackedCa ca_host;
    std::string ca_why;           // why Macar/Actis cannot run on this graph ("" = they can)
    CaView ca_view{};
    CaLayout lay_macar{}, lay_actis{}, lay_macar_cta{}, lay_actis_cta{};
    // forward scheme (luf_forward_attach)
    bool has_forward = false;
    PackedForward fw_host;
    FwView fw_view{};
    SampleTab fw_buffer_tab{};
    Layout fw_lay_uf{};
    CaLayout fw_lay_macar{};
    FwLayout fw_ext_uf{}, fw_ext_macar{};
    // grow-only scratch for the *_host entry points
    void* scratch[6] = {0, 0, 0, 0, 0, 0};
    size_t scratch_bytes[6] = {0, 0, 0, 0, 0, 0};
    std::atomic<long long> launches{0};
    cudaStream_t pipe[3] = {0, 0, 0};     // copy / compute pipeline of the *_host batch entry points
    std::mutex mu;
    std::map<std::pair<const void*, uint32_t>, std::pair<int, int>> plans;   // (kernel, slice bytes) -> (warps per block, blocks per SM)
    // compact fast path of the fused UF kernel (luf_uf_fast.cuh)
    bool canonical_ok = true;     // false: a graph beyond the 16-bit layout -- only the compact path's entry points run
    PackedFast fast_host;
    fast::FView fview{};
    std::map<void*, uint32_t*> queues;       // per stream: uint32[FAST_CHUNK] shot indices + the fill counter behind them
    // wide instance (luf_uf_wide.cuh): what runs when the graph is beyond the 16-bit layout (or LUF_FORCE_WIDE=1)
    bool wide = false;
    PackedWide wide_host;
    uf_wide::GraphView wview{};
    uf_wide::Layout wlay_fused{}, wlay_decode{}, wfw_lay_uf{};
    PackedForwardWide wfw_host;
    fw_wide::FwView wfw_view{};
    fw_wide::FwLayout wfw_ext{};
    SampleTab wfw_buffer_tab{};
    std::map<std::pair<void*, int>, std::pair<void*, size_t>> slabs;    // per (stream, use): the state slabs of a persistent grid
};

static const long long FAST_CHUNK = 1ll << 22;       // shots per fast launch (the queue of HARD shots holds as many)

template <class T>
static int upload(luf_graph* g, const std::vector<T>& v, const T** out)
{
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, v.size() * sizeof(T) + 16));
    g->allocs.push_back(p);
    CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return 0;
}

static int scratch_get(luf_graph* g, int slot, size_t bytes, void** out)
{
    if (bytes > g->scratch_bytes[slot]) {
        if (g->scratch[slot]) CUDA_TRY(cudaFree(g->scratch[slot]));
        g->scratch[slot] = nullptr; g->scratch_bytes[slot] = 0;
        CUDA_TRY(cudaMalloc(&g->scratch[slot], bytes));
        g->scratch_bytes[slot] = bytes;
    }
    *out = g->scratch[slot];
    return 0;
}

struct LaunchCfg { int grid, block; size_t smem; };

// The opt-in shared-memory limit of a kernel is process-wide state of the function (per device): it is raised
// once, to the device maximum, and never lowered, so concurrent callers (ctypes releases the GIL) with graphs
// of different slice sizes cannot undercut each other between a set-attribute and a launch.  The same mutex
// guards the plan caches of the graphs: the *_dev entry points take no per-graph lock.
static std::mutex g_plan_mu;
static std::set<std::pair<int, const void*>> g_smem_raised;

template <class K>
static int raise_smem_limit(int device, int max_smem, K kernel)
{
    std::lock_guard<std::mutex> lock(g_plan_mu);
    const std::pair<int, const void*> key(device, (const void*)kernel);
    if (g_smem_raised.count(key)) return 0;
    cudaFuncAttributes attr;
    CUDA_TRY(cudaFuncGetAttributes(&attr, kernel));           // static shared memory counts against the same limit
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - (int)attr.sharedSizeBytes));
    g_smem_raised.insert(key);
    return 0;
}

// Warps per block and a persistent grid: as many blocks per SM as the
// shared-memory budget allows, never more blocks than there is work.
template <class K>
static int plan_launch(luf_graph* g, K kernel, uint32_t slice_bytes, long long nshots, LaunchCfg* cfg, int max_wpb = 32)
{
    struct { uint32_t bytes; } l = {slice_bytes};
    if (l.bytes > (uint32_t)g->max_smem)
        return fail(LUF_ERR_UNSUPPORTED, "one shot's decoder state (" + std::to_string(l.bytes) +
                    " B) exceeds the shared memory of an SM");
    // Tunables (for sweeps; the defaults are what bench.py measures):
    //   LUF_WARPS_PER_BLOCK  warps (= concurrent shots) per block
    //   LUF_SMEM_KB_PER_SM   shared-memory budget per SM for shot state; the
    //                        rest of the 228 KB stays L1 for the graph tables
    static const int env_wpb = getenv("LUF_WARPS_PER_BLOCK") ? atoi(getenv("LUF_WARPS_PER_BLOCK")) : 0;
    static const int env_kb = getenv("LUF_SMEM_KB_PER_SM") ? atoi(getenv("LUF_SMEM_KB_PER_SM")) : 0;
    // pick the block size that keeps the most warps (= shots) resident per SM (cached per kernel)
    const int budget_kb = env_kb > 0 ? env_kb : 228;
    int wpb = 1, per_sm = 1, best = 0;
    const std::pair<const void*, uint32_t> key((const void*)kernel, slice_bytes);
    if (int rc = raise_smem_limit(g->device, g->max_smem, kernel)) return rc;
    bool cached = false;
    {
        std::lock_guard<std::mutex> lock(g_plan_mu);
        auto hit = g->plans.find(key);
        if (hit != g->plans.end()) { wpb = hit->second.first; per_sm = hit->second.second; best = wpb * per_sm; cached = true; }
    }
    if (!cached) {
        // the occupancy query honours the carve-out preference a previous plan left on the kernel
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        for (int cand = 1; cand <= max_wpb; ++cand) {
            if (env_wpb > 0 && cand != env_wpb) continue;
            const size_t sm = (size_t)cand * l.bytes;
            if (sm > (size_t)g->max_smem) break;
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, cand * WARP, sm));
            const int by_budget = (int)((size_t)budget_kb * 1024 / (sm + 1024));
            if (occ > by_budget) occ = by_budget;
            if (occ * cand >= best && occ > 0) { best = occ * cand; wpb = cand; per_sm = occ; }   // ties: fewer, larger blocks
        }
        if (best == 0) return fail(LUF_ERR_UNSUPPORTED, "kernel does not fit on an SM");
        if (getenv("LUF_DEBUG_PLAN"))
            fprintf(stderr, "[luf] launch plan: %u B per shot, %d warps per block, %d blocks per SM\n", slice_bytes, wpb, per_sm);
        std::lock_guard<std::mutex> lock(g_plan_mu);
        g->plans[key] = std::make_pair(wpb, per_sm);
    }
    const size_t smem = (size_t)wpb * l.bytes;
    {   // ask for just enough shared memory for per_sm blocks, leave the rest to L1 (a hint: harmless if another
        // thread's launch of the same kernel overwrites it)
        int pct = (int)(((smem + 1024) * per_sm * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (pct > 100) pct = 100;
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    long long blocks = (long long)g->sm_count * per_sm;
    const long long need = (nshots + wpb - 1) / wpb;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    cfg->grid = (int)blocks; cfg->block = wpb * WARP; cfg->smem = smem;
    return 0;
}

// One block per shot: the block size that keeps the most warps resident.
template <class K>
static int plan_cta_launch(luf_graph* g, K kernel, uint32_t slice_bytes, long long nshots, LaunchCfg* cfg)
{
    if (slice_bytes > (uint32_t)g->max_smem)
        return fail(LUF_ERR_UNSUPPORTED, "one shot's decoder state (" + std::to_string(slice_bytes) +
                    " B) exceeds the shared memory of an SM");
    static const int env_threads = getenv("LUF_CTA_THREADS") ? atoi(getenv("LUF_CTA_THREADS")) : 0;
    const std::pair<const void*, uint32_t> key((const void*)kernel, slice_bytes);
    if (int rc = raise_smem_limit(g->device, g->max_smem, kernel)) return rc;
    int threads = 0, per_sm = 0;
    {
        std::lock_guard<std::mutex> lock(g_plan_mu);
        auto hit = g->plans.find(key);
        if (hit != g->plans.end()) { threads = hit->second.first; per_sm = hit->second.second; }
    }
    if (!threads) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        int best = 0;
        for (int cand = 448; cand >= 64; cand -= 32) {
            if (env_threads > 0 && cand != env_threads) continue;
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, cand, slice_bytes));
            if (occ * cand > best) { best = occ * cand; threads = cand; per_sm = occ; }    // ties: larger blocks
        }
        if (best == 0) return fail(LUF_ERR_UNSUPPORTED, "kernel does not fit on an SM");
        if (getenv("LUF_DEBUG_PLAN"))
            fprintf(stderr, "[luf] launch plan (block per shot): %u B per shot, %d threads per block, %d blocks per SM\n",
                    slice_bytes, threads, per_sm);
        std::lock_guard<std::mutex> lock(g_plan_mu);
        g->plans[key] = std::make_pair(threads, per_sm);
    }
    long long blocks = (long long)g->sm_count * per_sm;
    if (blocks > nshots) blocks = nshots;
    if (blocks < 1) blocks = 1;
    cfg->grid = (int)blocks; cfg->block = threads; cfg->smem = slice_bytes;
    return 0;
}

// Warp tiles while a shot's state leaves at least 10 warps per SM, else one
// block per shot (LUF_TILE=warp|cta overrides, for measurements).  Measured on
// phenomenological Surface codes: warps win up to d = 15 (28.7 vs 18.4 M shots/s
// at p = 0.005, 1.9 vs 1.4 at 0.026), blocks from d = 17 on (16.0 vs 11.3; d = 25: 5.0 vs 0.73).
static bool use_cta_tiles(const luf_graph* g, uint32_t warp_slice_bytes, uint32_t cta_slice_bytes)
{
    static const char* env = getenv("LUF_TILE");
    if (cta_slice_bytes > (uint32_t)g->max_smem) return false;
    if (env && !strcmp(env, "warp")) return false;
    if (env && !strcmp(env, "cta")) return true;
    return (size_t)10 * warp_slice_bytes > (size_t)g->max_smem;
}

// Macar: blocks from about a thousand nodes on (cfg3, N = 1452: 20.8 M shots/s with two warps per shot
// against 18.3 M with one; d = 21 p = 0.01: 1.53 M against 0.094 M).  Actis visits every node in every
// timestep, so a block per shot pays from a few hundred nodes on.  LUF_CA_TILE=warp|cta overrides.
static bool use_ca_cta(const luf_graph* g, bool actis)
{
    const char* env = getenv("LUF_CA_TILE");
    const uint32_t warp_bytes = actis ? g->lay_actis.bytes : g->lay_macar.bytes;
    const uint32_t cta_bytes = actis ? g->lay_actis_cta.bytes : g->lay_macar_cta.bytes;
    if (cta_bytes > (uint32_t)g->max_smem) return false;
    if (env && !strcmp(env, "warp")) return false;
    if (env && !strcmp(env, "cta")) return true;
    (void)warp_bytes;
    return g->ca_view.N >= (actis ? 256 : 1024);
}

// One block per shot for the cellular automata: `want` threads (a multiple of 32); keep_resident caps them so
// that no fewer shots stay resident than shared memory allows (Macar: few touched nodes per shot; Actis
// sweeps all nodes every timestep and prefers the threads: cfg3 328 k shots/s with 128, 279 k with 64).
template <class K>
static int plan_ca_cta_launch(luf_graph* g, K kernel, uint32_t bytes, long long nshots, LaunchCfg* cfg, int want,
                              bool keep_resident = true)
{
    if (bytes > (uint32_t)g->max_smem)
        return fail(LUF_ERR_UNSUPPORTED, "one shot's decoder state (" + std::to_string(bytes) + " B) exceeds the shared memory of an SM");
    const int env_threads = getenv("LUF_CA_CTA_THREADS") ? atoi(getenv("LUF_CA_CTA_THREADS")) : 0;
    if (int rc = raise_smem_limit(g->device, g->max_smem, kernel)) return rc;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cudaFuncAttributes attr;
    CUDA_TRY(cudaFuncGetAttributes(&attr, kernel));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 64, bytes));
    if (per_sm < 1) return fail(LUF_ERR_UNSUPPORTED, "kernel does not fit on an SM");
    int most = keep_resident ? (2048 / per_sm) & ~31 : 1024;
    const int by_regs = (65536 / (attr.numRegs > 0 ? attr.numRegs : 64) / per_sm) & ~31;
    if (keep_resident && most > by_regs) most = by_regs;
    if (most > attr.maxThreadsPerBlock) most = attr.maxThreadsPerBlock & ~31;
    if (most > 1024) most = 1024;
    if (most < 64) most = 64;
    int threads = (want + 31) & ~31;
    if (threads < 64) threads = 64;
    if (threads > most) threads = most;
    if (env_threads >= 32 && env_threads <= 1024) threads = env_threads & ~31;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, bytes));
    if (per_sm < 1) return fail(LUF_ERR_UNSUPPORTED, "kernel does not fit on an SM");
    if (getenv("LUF_DEBUG_PLAN"))
        fprintf(stderr, "[luf] CA launch plan (block per shot): %u B per shot, %d threads per block, %d blocks per SM\n", bytes, threads, per_sm);
    long long blocks = (long long)g->sm_count * per_sm;
    if (blocks > nshots) blocks = nshots;
    if (blocks < 1) blocks = 1;
    cfg->grid = (int)blocks; cfg->block = threads; cfg->smem = bytes;
    return 0;
}

// ------------------------------------------------------- compact fast path: launch plan
struct FastPlan { int tw, spb, bps; bool adj_shared; uint32_t adj_bytes; };

// Tile width, shots per block, blocks per SM and whether the adjacency table is staged in shared memory.
// The most resident shots win; a tile is as narrow as still leaves about 24 warps per SM (a narrow tile wastes
// fewer lanes on the short lists of a shot, but the SM needs the warps to hide shared-memory latency).
// LUF_FAST_TW / LUF_FAST_ADJ / LUF_FAST_SPB / LUF_FAST_BPS override (measurements).  false = does not fit.
static bool fast_plan(const luf_graph* g, const fast::FLayout& l, FastPlan* out)
{
    static const int env_tw = getenv("LUF_FAST_TW") ? atoi(getenv("LUF_FAST_TW")) : 0;
    static const int env_adj = getenv("LUF_FAST_ADJ") ? atoi(getenv("LUF_FAST_ADJ")) : -1;
    static const int env_spb = getenv("LUF_FAST_SPB") ? atoi(getenv("LUF_FAST_SPB")) : 0;
    static const int env_bps = getenv("LUF_FAST_BPS") ? atoi(getenv("LUF_FAST_BPS")) : 0;
    static const int env_warps = getenv("LUF_FAST_WARPS") ? atoi(getenv("LUF_FAST_WARPS")) : 24;
    static const int env_l1 = getenv("LUF_FAST_L1_KB") ? atoi(getenv("LUF_FAST_L1_KB")) : -1;
    const uint32_t adj_all = align16((uint32_t)g->fast_host.fadj.size() * 4u);
    const int sm_smem = 228 * 1024;
    FastPlan best{};
    int best_shots = 0;
    // Measured on cfg3 (B200, M shots/s; adjacency 29 KB): shared, 1 block of 32 shots 175.8; shared, 2 x 16 173.6;
    // global with 84 KB of L1 left (1 x 32) 165.3; global with 28 KB of L1 left (2 x 21 = 42 shots) 119.8 -- the
    // table has to stay on chip, in shared memory or in L1, before more resident shots pay.  Tiles narrower
    // than a warp lose outright (16 lanes 72, 8 lanes 29: the shots of a warp serialise each other's branches).
    for (int adj = 1; adj >= 0; --adj) {
        if (env_adj >= 0 && adj != env_adj) continue;
        if (adj && env_adj < 0 && adj_all > 40u * 1024u) continue;       // a large table is worth more as resident shots
        const uint32_t adj_bytes = adj ? adj_all : 0u;
        // adjacency in global memory: leave it room in L1 (up to 64 KB) when that costs less than half the shots
        int l1_kb = adj ? 0 : (int)std::min<uint32_t>((adj_all + 16u * 1024u) / 1024u, 64u);
        if (!adj && adj_all > 64u * 1024u) l1_kb = 0;                     // cannot be kept anyway
        if (env_l1 >= 0) l1_kb = adj ? 0 : env_l1;
        for (int bps = 1; bps <= 2; ++bps) {
            if (env_bps > 0 && bps != env_bps) continue;
            long long budget = (sm_smem - l1_kb * 1024) / bps - 1024;
            if (budget > g->max_smem) budget = g->max_smem;
            if (budget < (long long)adj_bytes + (long long)l.bytes) continue;
            const int fit = (int)((budget - adj_bytes) / l.bytes);
            // the narrowest tile, a warp at least, that keeps the warp count: fit shots * tw / 32 ~ env_warps / bps
            int tw = 32;
            while (tw < 256 && (long long)fit * bps * tw < 32ll * env_warps) tw <<= 1;
            if (env_tw > 0) tw = env_tw;
            int spb = fit;
            if (spb > 1024 / tw) spb = 1024 / tw;
            if (tw > 32 && spb > 15) spb = 15;                 // named barriers 1 .. 15
            if (env_spb > 0 && env_spb < spb) spb = env_spb;
            if (spb < 1) continue;
            // shots first; a table in shared memory is worth a fifth more of them (cfg3: 32 shots with it 175.8,
            // 34 without 162.9); then fewer blocks
            const int score = spb * bps * (adj ? 12 : 10) * 4 + (2 - bps);
            if (score > best_shots) { best_shots = score; best = FastPlan{tw, spb, bps, adj != 0, adj_bytes}; }
        }
    }
    if (!best_shots) return false;
    *out = best;
    return true;
}

static int fast_queue(luf_graph* g, void* stream, uint32_t** out)
{
    std::lock_guard<std::mutex> lock(g_plan_mu);       // (not g->mu: the *_host entry points hold that one while they call in here)
    auto it = g->queues.find(stream);
    if (it == g->queues.end()) {
        void* p = nullptr;
        // Q1 [FAST_CHUNK] | counters [16] | Q3 [FAST_CHUNK] | Q2 pairs [2 * FAST_CHUNK]  (fast_queue_layout)
        CUDA_TRY(cudaMalloc(&p, ((size_t)4 * FAST_CHUNK + 16) * 4));
        it = g->queues.emplace(stream, (uint32_t*)p).first;
    }
    *out = it->second;
    return 0;
}

struct FastLogArgs { fast::FExport x; const uint32_t* qin; const uint32_t* qin_count; uint32_t* glog; int mode; };     // mode: 1 second pass, 2 one pass

template <int TW, bool ADJ, int LOGQ>
static int fast_launch_t(luf_graph* g, const fast::FView& fv, const FastPlan& plan, const fast::FLayout& fl, const SampleParams& sp,
                         unsigned long long shot0, long long nshots, unsigned long long* counts, uint8_t* logical,
                         uint32_t* queue, uint32_t* qcount, const FastLogArgs& la, cudaStream_t stream)
{
    if (int rc = raise_smem_limit(g->device, g->max_smem, k_mc_uf_fast<TW, ADJ, LOGQ>)) return rc;
    {
        static std::atomic<bool> carved{false};
        if (!carved.exchange(true)) CUDA_TRY(cudaFuncSetAttribute(k_mc_uf_fast<TW, ADJ, LOGQ>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    const size_t smem = (size_t)plan.adj_bytes + (size_t)plan.spb * fl.bytes;
    long long blocks = (long long)g->sm_count * plan.bps;
    const long long need = (nshots + plan.spb - 1) / plan.spb;
    if (blocks > need) blocks = need;
    if (getenv("LUF_DEBUG_PLAN"))
        fprintf(stderr, "[luf] fast plan: %u B per shot (lists %u / %u / %u), tile %d, %d shots per block, %d blocks per SM, adjacency %s (%u B), %zu B of shared memory per block\n",
                fl.bytes, fl.tcap, fl.acap, fl.mcap, plan.tw, plan.spb, plan.bps, plan.adj_shared ? "shared" : "global", plan.adj_bytes, smem);
    k_mc_uf_fast<TW, ADJ, LOGQ><<<(int)blocks, plan.spb * TW, smem, stream>>>(fv, fl, plan.adj_bytes, sp, shot0, nshots, counts,
                                                                            logical, queue, qcount, la.x, la.qin, la.qin_count, la.glog);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// `la` (optional): the second pass over the queued HARD shots, with the merge log
static int fast_launch(luf_graph* g, const fast::FView& fv, const FastPlan& plan, const fast::FLayout& fl, const SampleParams& sp,
                       unsigned long long shot0, long long nshots, unsigned long long* counts, uint8_t* logical,
                       uint32_t* queue, uint32_t* qcount, cudaStream_t stream, const FastLogArgs* la = nullptr)
{
    static const FastLogArgs none{};
#define LUF_FAST_TW(T) case T: return la && la->mode == 2 ? (plan.adj_shared ? fast_launch_t<T, true, 2>(g, fv, plan, fl, sp, shot0, nshots, counts, logical, queue, qcount, *la, stream) \
                                                            : fast_launch_t<T, false, 2>(g, fv, plan, fl, sp, shot0, nshots, counts, logical, queue, qcount, *la, stream)) \
                                  : la ? (plan.adj_shared ? fast_launch_t<T, true, 1>(g, fv, plan, fl, sp, shot0, nshots, counts, logical, queue, qcount, *la, stream) \
                                                            : fast_launch_t<T, false, 1>(g, fv, plan, fl, sp, shot0, nshots, counts, logical, queue, qcount, *la, stream)) \
                                         : (plan.adj_shared ? fast_launch_t<T, true, 0>(g, fv, plan, fl, sp, shot0, nshots, counts, logical, queue, qcount, none, stream) \
                                                            : fast_launch_t<T, false, 0>(g, fv, plan, fl, sp, shot0, nshots, counts, logical, queue, qcount, none, stream));
    switch (plan.tw) {
        LUF_FAST_TW(16) LUF_FAST_TW(32) LUF_FAST_TW(64) LUF_FAST_TW(128) LUF_FAST_TW(256)
        default: return fail(LUF_ERR_INVALID, "fast path: tile width must be 16, 32, 64, 128 or 256");
    }
#undef LUF_FAST_TW
}

// ------------------------------------------------------- wide instance: launch plan
struct WidePlan { int grid, threads; size_t smem; unsigned char* slab; };

// Grow-only device scratch per (stream, use).  cudaFree waits for the device, so a slab still read by an earlier
// launch on the stream is never pulled away.
static int wide_slab(luf_graph* g, void* stream, int use, size_t bytes, void** out)
{
    std::lock_guard<std::mutex> lock(g_plan_mu);
    auto& e = g->slabs[std::make_pair(stream, use)];
    if (bytes > e.second) {
        if (e.first) CUDA_TRY(cudaFree(e.first));
        e.first = nullptr; e.second = 0;
        CUDA_TRY(cudaMalloc(&e.first, bytes));
        e.second = bytes;
    }
    *out = e.first;
    return 0;
}

// One block per shot.  The state slab goes to shared memory when two blocks' slabs fit on an SM, else to global
// memory with as many resident blocks as the registers allow while the slabs of the grid stay within
// LUF_WIDE_L2_MB (default 192: Surface(23, 'circuit-level'), 197 KB per shot, 820 k shots/s with four blocks of
// 256 threads per SM = 116 MB of slabs against 710 k with three and 330 k with one; the B200's L2 holds 126 MB
// and what spills is streamed from HBM).  LUF_WIDE_STATE=smem|global, LUF_WIDE_THREADS override (measurements).
template <class K>
static int wide_plan_bytes(luf_graph* g, K kernel, uint32_t slab_bytes, long long nshots, void* stream, WidePlan* out)
{
    struct { uint32_t bytes; } l = {slab_bytes};
    const char* env_state = getenv("LUF_WIDE_STATE");
    const int env_threads = getenv("LUF_WIDE_THREADS") ? atoi(getenv("LUF_WIDE_THREADS")) : 0;
    const int env_l2 = getenv("LUF_WIDE_L2_MB") ? atoi(getenv("LUF_WIDE_L2_MB")) : 192;
    bool in_smem = (size_t)2 * l.bytes + 2048 <= (size_t)g->max_smem;
    if (env_state && !strcmp(env_state, "global")) in_smem = false;
    if (env_state && !strcmp(env_state, "smem")) in_smem = l.bytes <= (uint32_t)g->max_smem;
    if (int rc = raise_smem_limit(g->device, g->max_smem, kernel)) return rc;
    int threads = env_threads >= 32 && env_threads <= 512 ? env_threads & ~31 : 256;
    const size_t smem = in_smem ? l.bytes : 0;
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    if (per_sm < 1) return fail(LUF_ERR_UNSUPPORTED, "wide kernel does not fit on an SM");
    if (!in_smem) {
        const long long by_l2 = std::max<long long>(1, (long long)env_l2 * 1024 * 1024 / ((long long)g->sm_count * l.bytes));
        if (per_sm > by_l2) per_sm = (int)by_l2;
        if (!env_threads && per_sm <= 3) {          // few resident blocks: wider ones
            threads = 512;
            int occ = 0;
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
            if (occ < 1) threads = 256; else if (per_sm > occ) per_sm = occ;
        }
    }
    long long blocks = (long long)g->sm_count * per_sm;
    if (blocks > nshots) blocks = nshots;
    if (blocks < 1) blocks = 1;
    out->grid = (int)blocks; out->threads = threads; out->smem = smem; out->slab = nullptr;
    if (!in_smem) {
        void* p = nullptr;
        if (int rc = wide_slab(g, stream, 0, (size_t)g->sm_count * per_sm * l.bytes, &p)) return rc;
        out->slab = (unsigned char*)p;
    }
    if (getenv("LUF_DEBUG_PLAN"))
        fprintf(stderr, "[luf] wide plan: %u B per shot in %s memory, %d threads per block, %d blocks per SM\n", l.bytes,
                in_smem ? "shared" : "global", threads, per_sm);
    return 0;
}

template <class K>
static int wide_plan(luf_graph* g, K kernel, const uf_wide::Layout& l, long long nshots, void* stream, WidePlan* out)
{
    return wide_plan_bytes(g, kernel, l.bytes, nshots, stream, out);
}

// *_host entry points on stream 0: no asynchronous copy into a caller's buffer outlives the call, error or not
struct DrainDefaultStream { ~DrainDefaultStream() { cudaStreamSynchronize(0); } };

static int check_device(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(LUF_ERR_NO_DEVICE, "no CUDA device visible; the engine has no CPU fallback");
    }
    return 0;
}

// ------------------------------------------------------------------ streams
struct luf_stream {
    int device = 0, sm_count = 0, max_smem = 0;
    PackedWindow host;
    SfView view{};
    SfLayout lay{}, lay_cta{};        // one warp / one thread block per window
    std::vector<void*> allocs;
    void* scratch[5] = {0, 0, 0, 0, 0};
    size_t scratch_bytes[5] = {0, 0, 0, 0, 0};
    std::atomic<long long> launches{0};
    std::mutex mu;
};

template <class T>
static int upload_s(luf_stream* g, const std::vector<T>& v, const T** out)
{
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, v.size() * sizeof(T) + 16));
    g->allocs.push_back(p);
    CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return 0;
}

static int scratch_s(luf_stream* g, int slot, size_t bytes, void** out)
{
    if (bytes > g->scratch_bytes[slot]) {
        if (g->scratch[slot]) CUDA_TRY(cudaFree(g->scratch[slot]));
        g->scratch[slot] = nullptr; g->scratch_bytes[slot] = 0;
        CUDA_TRY(cudaMalloc(&g->scratch[slot], bytes));
        g->scratch_bytes[slot] = bytes;
    }
    *out = g->scratch[slot];
    return 0;
}

// one warp per stream; as many blocks per SM as shared memory allows
template <class K>
static int plan_stream_launch(luf_stream* g, K kernel, long long nstreams, LaunchCfg* cfg)
{
    const uint32_t bytes = g->lay.bytes;
    if (bytes > (uint32_t)g->max_smem)
        return fail(LUF_ERR_UNSUPPORTED, "one window's decoder state (" + std::to_string(bytes) +
                    " B) exceeds the shared memory of an SM");
    int wpb = g->max_smem / (int)bytes;
    if (wpb > 4) wpb = 4;
    const size_t smem = (size_t)wpb * bytes;
    if (int rc = raise_smem_limit(g->device, g->max_smem, kernel)) return rc;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, wpb * WARP, smem));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)g->sm_count * per_sm;
    const long long need = (nstreams + wpb - 1) / wpb;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    cfg->grid = (int)blocks; cfg->block = wpb * WARP; cfg->smem = smem;
    return 0;
}

// One thread block per window.  `awake_estimate` = the number of awake nodes a sweep is expected to
// visit (0 = unknown): the block gets about one thread per two of them, at least 128, and at most what
// fills the SM's thread slots next to the other resident windows (measured, B200: d=13 at p=0.003
// 27 M cycles/s with 128 threads against 21 with 256; d=25 at p=0.008 the other way round).
template <class K>
static int plan_stream_cta_launch(luf_stream* g, K kernel, long long nstreams, LaunchCfg* cfg, double awake_estimate = 0.0)
{
    const uint32_t bytes = g->lay_cta.bytes;
    if (bytes > (uint32_t)g->max_smem)
        return fail(LUF_ERR_UNSUPPORTED, "one window's decoder state (" + std::to_string(bytes) +
                    " B) exceeds the shared memory of an SM");
    const int env_threads = getenv("LUF_SF_CTA_THREADS") ? atoi(getenv("LUF_SF_CTA_THREADS")) : 0;
    if (int rc = raise_smem_limit(g->device, g->max_smem, kernel)) return rc;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    cudaFuncAttributes attr;
    CUDA_TRY(cudaFuncGetAttributes(&attr, kernel));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, bytes));
    if (per_sm < 1) return fail(LUF_ERR_UNSUPPORTED, "kernel does not fit on an SM");
    int most = (2048 / per_sm) & ~31;                                  // thread slots of an SM ...
    const int by_regs = (65536 / (attr.numRegs > 0 ? attr.numRegs : 64) / per_sm) & ~31;   // ... and its registers:
    if (most > by_regs) most = by_regs;                                // never fewer resident windows than shared memory allows
    if (most > attr.maxThreadsPerBlock) most = attr.maxThreadsPerBlock & ~31;
    if (most > 1024) most = 1024;
    if (most < 128) most = 128;
    int threads = awake_estimate > 0.0 ? (((int)(awake_estimate / 2.0) + 31) & ~31) : most;
    if (threads < 128) threads = 128;
    if (threads > most) threads = most;
    if (env_threads >= 32 && env_threads <= 1024) threads = env_threads & ~31;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, bytes));
    if (per_sm < 1) return fail(LUF_ERR_UNSUPPORTED, "kernel does not fit on an SM");
    if (getenv("LUF_DEBUG_PLAN"))
        fprintf(stderr, "[luf] stream launch plan (block per window): %u B per window, %d threads per block, %d blocks per SM\n",
                bytes, threads, per_sm);
    long long blocks = (long long)g->sm_count * per_sm;
    if (blocks > nstreams) blocks = nstreams;
    if (blocks < 1) blocks = 1;
    cfg->grid = (int)blocks; cfg->block = threads; cfg->smem = bytes;
    return 0;
}

// Warp tiles while a window's state leaves at least 16 warps per SM, else one block per window
// (LUF_SF_TILE=warp|cta overrides, for measurements and tests).
static bool use_stream_cta(const luf_stream* g)
{
    const char* env = getenv("LUF_SF_TILE");
    if (g->lay_cta.bytes > (uint32_t)g->max_smem) return false;
    if (env && !strcmp(env, "warp")) return false;
    if (env && !strcmp(env, "cta")) return true;
    return (size_t)16 * g->lay.bytes > (size_t)g->max_smem;
}

extern "C" {

int luf_stream_create(const luf_window_desc* desc, int device, luf_stream** out)
{
    if (!desc || !out) return fail(LUF_ERR_INVALID, "null argument");
    *out = nullptr;
    if (int rc = check_device()) return rc;
    luf_stream* g = new (std::nothrow) luf_stream();
    if (!g) return fail(LUF_ERR_INVALID, "out of host memory");
    const std::string why = pack_window(*desc, g->host);
    if (!why.empty()) {
        delete g;
        return fail(why.find("too large") != std::string::npos ? LUF_ERR_UNSUPPORTED : LUF_ERR_INVALID, why);
    }
    g->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete g; return fail(LUF_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&g->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const PackedWindow& p = g->host;
    SfView& v = g->view;
    v.d = p.d; v.h = p.h; v.NLb = p.NLb; v.NLd = p.NLd; v.NL = p.NL; v.EL = p.EL; v.K = p.K;
    v.ELw = words32(p.EL); v.WLn = words32(p.NL); v.N = p.NL * p.h; v.n_classes = p.n_classes;
    int rc = 0;
    if ((rc = upload_s(g, p.nbr, &v.nbr)) || (rc = upload_s(g, p.edge_ends, &v.edge_ends)) ||
        (rc = upload_s(g, p.pos_long, &v.pos_long)) || (rc = upload_s(g, p.fresh_perm, &v.fresh_perm)) ||
        (rc = upload_s(g, p.class_end, &v.class_end))) {
        luf_stream_destroy(g);
        return rc;
    }
    g->lay = make_sf_layout(p);
    // block tiles: the awake-node list takes whatever shared memory the windows resident on an SM leave over
    {
        int sm_smem = 0;
        cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
        const uint32_t min_cap = SF_CTA_ALIST_CAP < 256u ? SF_CTA_ALIST_CAP : 256u;
        const uint32_t base = make_sf_layout(p, min_cap).bytes + 1024u;       // + the per-block reservation
        uint32_t cap = SF_CTA_ALIST_CAP;
        if (SF_CTA_ALIST_CAP >= 256u && base <= (uint32_t)sm_smem) {           // (test builds keep their tiny list)
            const uint32_t blocks = (uint32_t)sm_smem / base;
            uint32_t budget = (uint32_t)sm_smem / blocks - 1024u;
            if (budget > (uint32_t)g->max_smem) budget = (uint32_t)g->max_smem;
            cap = min_cap + (budget - (base - 1024u)) / 2u;
            cap &= ~7u;
            if (cap > (uint32_t)v.N) cap = ((uint32_t)v.N + 7u) & ~7u;
            if (cap > 16384u) cap = 16384u;
        }
        g->lay_cta = make_sf_layout(p, cap);
    }
    *out = g;
    return 0;
}

int luf_stream_destroy(luf_stream* g)
{
    if (!g) return 0;
    cudaSetDevice(g->device);
    for (void* p : g->allocs) cudaFree(p);
    for (void* p : g->scratch) if (p) cudaFree(p);
    delete g;
    return 0;
}

int64_t luf_stream_launch_count(const luf_stream* g) { return g ? g->launches.load() : 0; }

int luf_stream_decode(luf_stream* g, int flags, int64_t nstreams, int32_t ncycles, const uint32_t* fresh_bits_dev,
                      int32_t* rt_dev, uint32_t* floor_bits_dev, int32_t* m_dev, void* stream)
{
    return luf_stream_decode_metrics(g, flags, nstreams, ncycles, fresh_bits_dev, rt_dev, floor_bits_dev, m_dev,
                                     nullptr, nullptr, nullptr, stream);
}

int luf_stream_growth_words(const luf_stream* g) { return g ? g->view.h * 2 * g->view.ELw : 0; }

int luf_stream_decode_metrics(luf_stream* g, int flags, int64_t nstreams, int32_t ncycles, const uint32_t* fresh_bits_dev,
                              int32_t* rt_dev, uint32_t* floor_bits_dev, int32_t* m_dev, uint32_t* growth2b_dev,
                              int32_t* mins_dev, const uint8_t* drop_only_dev, void* stream)
{
    if (!g || !fresh_bits_dev || !rt_dev || !m_dev || nstreams < 0 || ncycles < 0 || (!growth2b_dev) != (!mins_dev))
        return fail(LUF_ERR_INVALID, "luf_stream_decode: bad argument");
    if ((growth2b_dev || drop_only_dev) && (flags & LUF_SF_LATENCY_TAIL))
        return fail(LUF_ERR_INVALID, "per-cycle metrics / drop-only flags do not combine with the tail of sample_latency");
    if (flags & ~(LUF_SF_SLOW_MERGER | LUF_SF_ONE_ONE | LUF_SF_SYNDROME | LUF_SF_LATENCY_TAIL))
        return fail(LUF_ERR_INVALID, "unknown Snowflake flag");
    if (nstreams == 0 || ncycles == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    LaunchCfg c;
    if (use_stream_cta(g)) {
        if (int rc = plan_stream_cta_launch(g, k_sf_decode_cta, nstreams, &c)) return rc;
        k_sf_decode_cta<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_cta, flags, nstreams, ncycles,
                                                                            fresh_bits_dev, rt_dev, floor_bits_dev, m_dev,
                                                                            growth2b_dev, mins_dev, drop_only_dev);
    } else {
        if (int rc = plan_stream_launch(g, k_sf_decode, nstreams, &c)) return rc;
        k_sf_decode<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay, flags, nstreams, ncycles,
                                                                        fresh_bits_dev, rt_dev, floor_bits_dev, m_dev,
                                                                        growth2b_dev, mins_dev, drop_only_dev);
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_stream_mc(luf_stream* g, int flags, const double* class_p_host, uint64_t seed, uint64_t stream0,
                  int64_t nstreams, int32_t n, unsigned long long* counts_dev, int32_t* steps_dev, void* stream)
{
    return luf_stream_mc_metrics(g, flags, class_p_host, seed, stream0, nstreams, n, counts_dev, steps_dev, nullptr, nullptr, stream);
}

int luf_stream_mc_metrics(luf_stream* g, int flags, const double* class_p_host, uint64_t seed, uint64_t stream0,
                          int64_t nstreams, int32_t n, unsigned long long* counts_dev, int32_t* steps_dev,
                          uint32_t* growth2b_dev, int32_t* cycle_dev, void* stream)
{
    if (!g || !class_p_host || !counts_dev || nstreams < 0 || (!growth2b_dev) != (!cycle_dev))
        return fail(LUF_ERR_INVALID, "luf_stream_mc: bad argument");
    if (n < 1) return fail(LUF_ERR_INVALID, "n must be positive integer.");
    if (flags & ~(LUF_SF_SLOW_MERGER | LUF_SF_ONE_ONE)) return fail(LUF_ERR_INVALID, "unknown Snowflake flag");
    if (nstreams == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const SampleParams sp = make_sample_params(class_p_host, g->view.n_classes, seed);
    LaunchCfg c;
    if (use_stream_cta(g)) {
        // expected awake nodes of a sweep: about 4 per fresh fault and sheet (two defects, and as many woken neighbours)
        double faults = 0.0;
        for (int cl = 0, lo = 0; cl < g->view.n_classes; ++cl) {
            const int hi = (int)g->host.class_end[cl];
            faults += class_p_host[cl] * (hi - lo);
            lo = hi;
        }
        if (int rc = plan_stream_cta_launch(g, k_sf_mc_cta, nstreams, &c, 4.0 * faults * g->view.h + 1.0)) return rc;
        k_sf_mc_cta<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_cta, sp, flags, stream0, nstreams, n,
                                                                        counts_dev, steps_dev, growth2b_dev, cycle_dev);
    } else {
        if (int rc = plan_stream_launch(g, k_sf_mc, nstreams, &c)) return rc;
        k_sf_mc<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay, sp, flags, stream0, nstreams, n,
                                                                    counts_dev, steps_dev, growth2b_dev, cycle_dev);
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_stream_decode_host(luf_stream* g, int flags, int64_t nstreams, int32_t ncycles, const uint32_t* fresh_bits_host,
                           int32_t* rt_host, uint32_t* floor_bits_host, int32_t* m_host)
{
    if (!g || !fresh_bits_host || !rt_host || !m_host || nstreams < 0 || ncycles < 0)
        return fail(LUF_ERR_INVALID, "luf_stream_decode_host: bad argument");
    if (nstreams == 0 || ncycles == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t rows = (size_t)nstreams * ncycles, W = g->view.ELw;
    void *fr, *rt, *fl = nullptr, *mm;
    int rc;
    if ((rc = scratch_s(g, 0, rows * W * 4, &fr)) || (rc = scratch_s(g, 1, rows * 8, &rt)) ||
        (rc = scratch_s(g, 2, rows * 4, &mm))) return rc;
    if (floor_bits_host && (rc = scratch_s(g, 3, rows * W * 4, &fl))) return rc;
    CUDA_TRY(cudaMemcpyAsync(fr, fresh_bits_host, rows * W * 4, cudaMemcpyHostToDevice, 0));
    if ((rc = luf_stream_decode(g, flags, nstreams, ncycles, (const uint32_t*)fr, (int32_t*)rt, (uint32_t*)fl,
                                (int32_t*)mm, 0))) return rc;
    CUDA_TRY(cudaMemcpyAsync(rt_host, rt, rows * 8, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpyAsync(m_host, mm, rows * 4, cudaMemcpyDeviceToHost, 0));
    if (fl) CUDA_TRY(cudaMemcpyAsync(floor_bits_host, fl, rows * W * 4, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

int luf_stream_mc_host(luf_stream* g, int flags, const double* class_p_host, uint64_t seed, uint64_t stream0,
                       int64_t nstreams, int32_t n, unsigned long long counts_host[4], int32_t* steps_host)
{
    if (!g || !counts_host || nstreams < 0) return fail(LUF_ERR_INVALID, "luf_stream_mc_host: bad argument");
    if (n < 1) return fail(LUF_ERR_INVALID, "n must be positive integer.");
    if (nstreams == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    void *cnt, *st = nullptr;
    int rc;
    if ((rc = scratch_s(g, 4, 32, &cnt))) return rc;
    if (steps_host && (rc = scratch_s(g, 1, (size_t)nstreams * n * 8, &st))) return rc;
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 32, 0));
    if ((rc = luf_stream_mc(g, flags, class_p_host, seed, stream0, nstreams, n, (unsigned long long*)cnt,
                            (int32_t*)st, 0))) return rc;
    unsigned long long tmp[4];
    if (st) CUDA_TRY(cudaMemcpyAsync(steps_host, st, (size_t)nstreams * n * 8, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpy(tmp, cnt, 32, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) counts_host[i] += tmp[i];
    return 0;
}

int luf_stream_latency(luf_stream* g, int flags, const double* class_p_host, uint64_t seed, uint64_t stream0,
                       int64_t nstreams, int32_t* lat_dev, void* stream)
{
    if (!g || !class_p_host || !lat_dev || nstreams < 0) return fail(LUF_ERR_INVALID, "luf_stream_latency: bad argument");
    if (flags & ~(LUF_SF_SLOW_MERGER | LUF_SF_ONE_ONE)) return fail(LUF_ERR_INVALID, "unknown Snowflake flag");
    if (nstreams == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const SampleParams sp = make_sample_params(class_p_host, g->view.n_classes, seed);
    LaunchCfg c;
    if (use_stream_cta(g)) {
        double faults = 0.0;
        for (int cl = 0, lo = 0; cl < g->view.n_classes; ++cl) {
            const int hi = (int)g->host.class_end[cl];
            faults += class_p_host[cl] * (hi - lo);
            lo = hi;
        }
        if (int rc = plan_stream_cta_launch(g, k_sf_latency_cta, nstreams, &c, 4.0 * faults * g->view.h + 1.0)) return rc;
        k_sf_latency_cta<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_cta, sp, flags, stream0, nstreams, lat_dev);
    } else {
        if (int rc = plan_stream_launch(g, k_sf_latency, nstreams, &c)) return rc;
        k_sf_latency<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay, sp, flags, stream0, nstreams, lat_dev);
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_stream_latency_host(luf_stream* g, int flags, const double* class_p_host, uint64_t seed, uint64_t stream0,
                            int64_t nstreams, int32_t* lat_host)
{
    if (!g || !lat_host || nstreams < 0) return fail(LUF_ERR_INVALID, "luf_stream_latency_host: bad argument");
    if (nstreams == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    void* lat;
    int rc;
    if ((rc = scratch_s(g, 1, (size_t)nstreams * 12, &lat))) return rc;
    if ((rc = luf_stream_latency(g, flags, class_p_host, seed, stream0, nstreams, (int32_t*)lat, 0))) return rc;
    CUDA_TRY(cudaMemcpy(lat_host, lat, (size_t)nstreams * 12, cudaMemcpyDeviceToHost));
    return 0;
}

const char* luf_last_error(void) { return g_last_error.c_str(); }

int luf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int luf_graph_create(const luf_graph_desc* desc, int device, luf_graph** out)
{
    if (!desc || !out) return fail(LUF_ERR_INVALID, "null argument");
    *out = nullptr;
    if (int rc = check_device()) return rc;
    luf_graph* g = new (std::nothrow) luf_graph();
    if (!g) return fail(LUF_ERR_INVALID, "out of host memory");
    const std::string why = pack_graph(*desc, g->host);
    pack_fast(*desc, g->fast_host);
    g->canonical_ok = why.empty();
    // beyond the 16-bit layout -- or LUF_FORCE_WIDE=1 (tests: any graph through the wide kernels)
    g->wide = (!why.empty() && why.find("too large") != std::string::npos) || (getenv("LUF_FORCE_WIDE") && atoi(getenv("LUF_FORCE_WIDE")) != 0);
    if (!why.empty() && !g->wide) { delete g; return fail(LUF_ERR_INVALID, why); }
    {   // the wide tables are kept for every graph: a forward window or a decode state that does not fit the
        // shared memory of an SM in the 16-bit layout runs through the wide instance too
        const std::string wwhy = pack_wide(*desc, g->wide_host);
        if (!wwhy.empty()) { delete g; return fail(LUF_ERR_INVALID, wwhy); }
    }
    g->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete g; return fail(LUF_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
    cudaDeviceGetAttribute(&g->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&g->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const PackedGraph& p = g->host;
    GraphView& v = g->view;
    int rc = 0;
    if (g->fast_host.ok) {
        const PackedFast& q = g->fast_host;
        fast::FView& f = g->fview;
        f.D = q.D; f.E = q.E; f.WD = words32(q.D); f.DEG = q.DEG; f.deg_magic = (uint32_t)(0x100000000ull / (uint64_t)q.DEG) + 1u; f.ebits = q.ebits; f.emask = q.emask; f.omask = q.omask;
        f.F = q.F; f.n_classes = q.n_classes; f.fresh_perm = nullptr; f.fresh_perm32 = nullptr; f.west = 0;
        if ((rc = upload(g, q.fadj, &f.fadj)) || (rc = upload(g, q.einfo, &f.einfo)) || (rc = upload(g, q.class_end, &f.class_end)) ||
            (!q.perm_identity && (rc = upload(g, q.fresh_perm32, &f.fresh_perm32)))) {
            luf_graph_destroy(g);
            return rc;
        }
    }
    {
        const PackedWide& q = g->wide_host;
        uf_wide::GraphView& w = g->wview;
        w.N = q.N; w.B = q.B; w.E = q.E; w.F = q.F; w.n_classes = q.n_classes; w.DEG = q.DEG;
        w.WE = words32(q.E); w.WD = words32(q.N - q.B); w.WN = words32(q.N);
        w.parity_ok = q.parity_ok ? 1 : 0; w.leaf_bnd = g->fast_host.ok ? 1 : 0; w.fresh_perm32 = nullptr;
        if ((rc = upload(g, q.adj, &w.adj)) || (rc = upload(g, q.edge_uv, &w.edge_uv)) || (rc = upload(g, q.west_mask, &w.west_mask)) ||
            (rc = upload(g, q.node_long, &w.node_long)) || (rc = upload(g, q.class_end, &w.class_end)) ||
            (!q.perm_identity && (rc = upload(g, q.fresh_perm32, &w.fresh_perm32)))) {
            luf_graph_destroy(g);
            return rc;
        }
        g->wlay_fused = make_wide_layout(q.N, q.B, q.E, false, false);
        g->wlay_decode = make_wide_layout(q.N, q.B, q.E, false, true);
    }
    if (!g->canonical_ok) {
        // beyond the 16-bit layout: UF runs through the wide instance (luf_uf_wide.cuh) and, for the fused
        // Monte-Carlo entry points, the compact fast path; Macar / Actis answer LUF_ERR_UNSUPPORTED
        memset(&v, 0, sizeof(v));
        v.N = desc->n_nodes; v.B = desc->n_boundary; v.E = desc->n_edges; v.F = g->fast_host.F; v.n_classes = desc->n_classes;
        v.WE = words32(v.E); v.WD = words32(v.N - v.B); v.WN = words32(v.N);
        v.west_mask = g->wview.west_mask;
        g->ca_why = "graph too large for the 16-bit cellular-automaton layout";
        *out = g;
        return 0;
    }
    v.N = p.N; v.B = p.B; v.E = p.E; v.F = p.F; v.n_classes = p.n_classes; v.DEG = p.DEG;
    v.parity_ok = p.parity_ok ? 1 : 0;
    v.leaf_bnd = g->fast_host.ok ? 1 : 0;
    v.WE = words32(p.E); v.WD = words32(p.N - p.B); v.WN = words32(p.N);
    if ((rc = upload(g, p.adj, &v.adj)) ||
        (rc = upload(g, p.edge_uv, &v.edge_uv)) || (rc = upload(g, p.west_mask, &v.west_mask)) ||
        (rc = upload(g, p.node_long, &v.node_long)) ||
        (!p.perm_identity && (rc = upload(g, p.fresh_perm, &v.fresh_perm))) ||
        (rc = upload(g, p.class_end, &v.class_end))) {
        luf_graph_destroy(g);
        return rc;
    }
    g->lay_fused = make_layout(p.N, p.B, p.E, false, false);
    g->lay_decode = make_layout(p.N, p.B, p.E, false, true);
    g->lay_fused_cta = make_layout(p.N, p.B, p.E, false, false, LUF_CTA_LIST_CAP, LUF_CTA_LIST2_CAP, LUF_CTA_CLAIM_SLOTS);
    g->lay_decode_cta = make_layout(p.N, p.B, p.E, false, true, LUF_CTA_LIST_CAP, LUF_CTA_LIST2_CAP, LUF_CTA_CLAIM_SLOTS);
    if (g->lay_decode.bytes > (uint32_t)g->max_smem && g->lay_decode_cta.bytes > (uint32_t)g->max_smem) g->wide = true;   // no 16-bit tile holds a shot
    {   // the sampler only needs the error and defect bitmaps
        Layout l; memset(&l, 0, sizeof(l));
        uint32_t off = 0;
        l.defect = off; off = align16(off + 4u * (uint32_t)(v.WD ? v.WD : 1));
        l.err = off; off = align16(off + 4u * (uint32_t)v.WE);
        l.bytes = off;
        // keep every pointer inside the slice (unused arrays alias the start)
        g->lay_sample = l;
    }
    g->ca_why = pack_ca(*desc, g->ca_host);
    if (g->ca_why.empty()) {
        CaView& c = g->ca_view;
        const PackedCa& q = g->ca_host;
        c.N = q.N; c.B = q.B; c.E = q.E; c.K = q.K; c.global_span = q.global_span;
        c.WE = v.WE; c.WD = v.WD; c.WN = v.WN; c.GW = (q.E + 15) / 16;
        c.west_mask = v.west_mask;
        if ((rc = upload(g, q.nbr, &c.nbr)) || (rc = upload(g, q.span, &c.span)) || (rc = upload(g, q.relay, &c.relay))) {
            luf_graph_destroy(g);
            return rc;
        }
        g->lay_macar = make_ca_layout(q.N, q.B, q.E, false, false);
        g->lay_actis = make_ca_layout(q.N, q.B, q.E, true, false);
        {   // block tiles: the touched-node list covers every node when that costs less than a tenth of the state
            uint32_t cap = CA_CTA_TLIST_CAP;
            if (CA_CTA_TLIST_CAP >= 256u && (uint32_t)q.N > cap && 2u * (uint32_t)q.N < g->lay_macar.bytes / 10u) cap = ((uint32_t)q.N + 7u) & ~7u;
            g->lay_macar_cta = make_ca_layout(q.N, q.B, q.E, false, false, cap);
            g->lay_actis_cta = make_ca_layout(q.N, q.B, q.E, true, false, 8);
        }
    }
    *out = g;
    return 0;
}

int luf_graph_destroy(luf_graph* g)
{
    if (!g) return 0;
    cudaSetDevice(g->device);
    for (void* p : g->allocs) cudaFree(p);
    for (void* p : g->scratch) if (p) cudaFree(p);
    for (auto& q : g->queues) cudaFree(q.second);
    for (auto& q : g->slabs) if (q.second.first) cudaFree(q.second.first);
    for (cudaStream_t st : g->pipe) if (st) cudaStreamDestroy(st);
    delete g;
    return 0;
}

int luf_graph_words(const luf_graph* g, int32_t* edge_words, int32_t* detector_words, int32_t* growth_words)
{
    if (!g) return fail(LUF_ERR_INVALID, "null graph");
    if (edge_words) *edge_words = g->view.WE;
    if (detector_words) *detector_words = g->view.WD;
    if (growth_words) *growth_words = (g->view.E + 15) / 16;
    return 0;
}

int64_t luf_launch_count(const luf_graph* g) { return g ? g->launches.load() : 0; }

static int check_decoder(const luf_graph* g, int decoder, int flags)
{
    if (decoder == LUF_DEC_UF) {
        if (flags & ~(LUF_UF_DYNAMIC | LUF_UF_WEST | LUF_UF_NODE | LUF_UF_BUCKET)) return fail(LUF_ERR_INVALID, "unknown UF flag");
        if ((flags & LUF_UF_BUCKET) && !(flags & LUF_UF_NODE) && g->view.E > 32767 && !g->wide)
            return fail(LUF_ERR_UNSUPPORTED, "BUF keeps vision lengths in 15 bits: graphs of up to 32767 edges");
        return 0;
    }
    if (decoder == LUF_DEC_MACAR || decoder == LUF_DEC_ACTIS) {
        if (!g->ca_why.empty()) return fail(LUF_ERR_UNSUPPORTED, g->ca_why);
        if (decoder == LUF_DEC_MACAR && flags) return fail(LUF_ERR_INVALID, "Macar takes no flags");
        if (decoder == LUF_DEC_ACTIS) {
            if (flags & ~LUF_ACTIS_UNOPTIMAL) return fail(LUF_ERR_INVALID, "unknown Actis flag");
            if (!g->ca_host.actis_ok) return fail(LUF_ERR_UNSUPPORTED, "Actis needs a batch-scheme surface-code graph");
        }
        return 0;
    }
    if (decoder == LUF_DEC_SNOWFLAKE)
        return fail(LUF_ERR_UNSUPPORTED, "Snowflake runs through the streaming entry points");
    return fail(LUF_ERR_INVALID, "unknown decoder");
}

int luf_sample(luf_graph* g, const double* class_p_host, uint64_t seed, uint64_t shot0, int64_t nshots,
               uint32_t* err_bits_dev, uint32_t* syn_bits_dev, void* stream)
{
    if (!g || !class_p_host || !err_bits_dev || !syn_bits_dev || nshots < 0)
        return fail(LUF_ERR_INVALID, "luf_sample: bad argument");
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const SampleParams sp = make_sample_params(class_p_host, g->view.n_classes, seed);
    LaunchCfg c;
    if (g->wide && (g->canonical_ok || !g->fast_host.ok)) {       // the sampler of the wide tables (same random stream)
        const int wpb = 8;
        long long blocks = std::min<long long>((long long)g->sm_count * 8, (nshots + wpb - 1) / wpb);
        k_sample_wide<false><<<(int)blocks, wpb * WARP, 0, (cudaStream_t)stream>>>(g->wview, sp, 0u, shot0, nshots, err_bits_dev, syn_bits_dev);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        return 0;
    }
    if (!g->canonical_ok) {               // the sampler of the compact tables (same random stream)
        const uint32_t slice = align16(4u * (uint32_t)(g->view.WE + g->view.WD));
        if (int rc = raise_smem_limit(g->device, g->max_smem, k_sample_fast)) return rc;
        int wpb = std::max(1, std::min(8, g->max_smem / (int)slice));
        long long blocks = std::min<long long>((long long)g->sm_count * 4, (nshots + wpb - 1) / wpb);
        k_sample_fast<<<(int)blocks, wpb * WARP, (size_t)wpb * slice, (cudaStream_t)stream>>>(g->fview, g->view.WE, slice, sp, shot0,
                                                                                       nshots, err_bits_dev, syn_bits_dev);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        return 0;
    }
    const uint32_t group_bytes = 16u * (uint32_t)(g->view.WE + g->view.WD);            // four shots per warp at a time
    const long long ngroups = (nshots + 3) / 4;
    if (!(((uintptr_t)err_bits_dev | (uintptr_t)syn_bits_dev) & 15u)) {
        if (int rc = plan_launch(g, k_sample<true>, group_bytes, ngroups, &c)) return rc;
        k_sample<true><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, sp, shot0, nshots, err_bits_dev, syn_bits_dev);
    } else {
        if (int rc = plan_launch(g, k_sample<false>, group_bytes, ngroups, &c)) return rc;
        k_sample<false><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, sp, shot0, nshots, err_bits_dev, syn_bits_dev);
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_sample_weight(luf_graph* g, int32_t weight, uint64_t seed, uint64_t shot0, int64_t nshots,
                      uint32_t* err_bits_dev, uint32_t* syn_bits_dev, void* stream)
{
    if (!g || !err_bits_dev || !syn_bits_dev || nshots < 0)
        return fail(LUF_ERR_INVALID, "luf_sample_weight: bad argument");
    if (weight < 0 || weight > g->view.F)
        return fail(LUF_ERR_INVALID, "Sample larger than population or is negative");       // random.sample's message
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const double zero[MAX_CLASSES] = {0};
    const SampleParams sp = make_sample_params(zero, g->view.n_classes, seed);
    LaunchCfg c;
    if (g->wide) {
        const int wpb = 8;
        long long blocks = std::min<long long>((long long)g->sm_count * 8, (nshots + wpb - 1) / wpb);
        k_sample_wide<true><<<(int)blocks, wpb * WARP, 0, (cudaStream_t)stream>>>(g->wview, sp, (uint32_t)weight, shot0, nshots, err_bits_dev, syn_bits_dev);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        return 0;
    }
    if (int rc = plan_launch(g, k_sample_weight, g->lay_sample.bytes, nshots, &c)) return rc;
    k_sample_weight<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_sample, sp, (uint32_t)weight, shot0,
                                                                        nshots, err_bits_dev, syn_bits_dev);
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_decode_batch(luf_graph* g, int decoder, int flags, int64_t nshots, const uint32_t* syn_bits_dev,
                     uint32_t* corr_bits_dev, uint32_t* growth2b_dev, int32_t* steps_dev, void* stream)
{
    if (!g || !syn_bits_dev || !corr_bits_dev || nshots < 0)
        return fail(LUF_ERR_INVALID, "luf_decode_batch: bad argument");
    if (int rc = check_decoder(g, decoder, flags)) return rc;
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    LaunchCfg c;
    if (decoder == LUF_DEC_UF && g->wide) {
        WidePlan w;
        if (int rc = wide_plan(g, k_decode_uf_wide, g->wlay_decode, nshots, stream, &w)) return rc;
        k_decode_uf_wide<<<w.grid, w.threads, w.smem, (cudaStream_t)stream>>>(g->wview, g->wlay_decode, flags, nshots, syn_bits_dev,
                                                                           corr_bits_dev, growth2b_dev, steps_dev, w.slab);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        return 0;
    }
    if (decoder == LUF_DEC_MACAR) {
        if (use_ca_cta(g, false)) {
            if (int rc = plan_ca_cta_launch(g, k_decode_ca_cta<false>, g->lay_macar_cta.bytes, nshots, &c, g->ca_view.N / 16)) return rc;
            k_decode_ca_cta<false><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(
                g->ca_view, g->lay_macar_cta, flags, nshots, syn_bits_dev, corr_bits_dev, growth2b_dev, steps_dev);
        } else {
        if (int rc = plan_launch(g, k_decode_ca<false>, g->lay_macar.bytes, nshots, &c)) return rc;
        k_decode_ca<false><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(
            g->ca_view, g->lay_macar, flags, nshots, syn_bits_dev, corr_bits_dev, growth2b_dev, steps_dev);
        }
    } else if (decoder == LUF_DEC_ACTIS) {
        if (use_ca_cta(g, true)) {
            if (int rc = plan_ca_cta_launch(g, k_decode_ca_cta<true>, g->lay_actis_cta.bytes, nshots, &c, g->ca_view.N / 12, false)) return rc;
            k_decode_ca_cta<true><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(
                g->ca_view, g->lay_actis_cta, flags, nshots, syn_bits_dev, corr_bits_dev, growth2b_dev, steps_dev);
        } else {
        if (int rc = plan_launch(g, k_decode_ca<true>, g->lay_actis.bytes, nshots, &c)) return rc;
        k_decode_ca<true><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(
            g->ca_view, g->lay_actis, flags, nshots, syn_bits_dev, corr_bits_dev, growth2b_dev, steps_dev);
        }
    } else {
        if (use_cta_tiles(g, g->lay_decode.bytes, g->lay_decode_cta.bytes)) {
            if (int rc = plan_cta_launch(g, k_decode_uf_cta, g->lay_decode_cta.bytes, nshots, &c)) return rc;
            k_decode_uf_cta<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_decode_cta, flags, nshots,
                                                                                syn_bits_dev, corr_bits_dev, growth2b_dev, steps_dev);
        } else {
            if (int rc = plan_launch(g, k_decode_uf, g->lay_decode.bytes, nshots, &c)) return rc;
            k_decode_uf<<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_decode, flags, nshots,
                                                                            syn_bits_dev, corr_bits_dev, growth2b_dev, steps_dev);
        }
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_check(luf_graph* g, int64_t nshots, const uint32_t* err_bits_dev, const uint32_t* corr_bits_dev,
              uint8_t* logical_dev, unsigned long long* fail_count_dev, void* stream)
{
    if (!g || !err_bits_dev || !corr_bits_dev || !fail_count_dev || nshots < 0)
        return fail(LUF_ERR_INVALID, "luf_check: bad argument");
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const int block = 256, wpb = block / WARP, WE = g->view.WE;
    const long long cap = (long long)g->sm_count * 8;
    long long done = 0;
    // groups of four shots through the streaming kernel when the buffers allow 16-byte loads
    const bool aligned = !(((uintptr_t)err_bits_dev | (uintptr_t)corr_bits_dev) & 15u) && !((uintptr_t)logical_dev & 3u);
    if (aligned && nshots >= 4 && WE >= 2 && WE <= 256) {          // (WE == 1: the magic number would be 2^32)
        const long long ngroups = nshots / 4;
        long long blocks = (ngroups + wpb - 1) / wpb;
        if (blocks > cap) blocks = cap;
        const uint32_t magic = (uint32_t)((0x100000000ull + (uint64_t)WE - 1) / (uint64_t)WE);
#define LUF_CHECK4(V) k_check4<V><<<(int)blocks, block, 0, (cudaStream_t)stream>>>(WE, magic, ngroups, g->view.west_mask, \
            (const uint4*)err_bits_dev, (const uint4*)corr_bits_dev, logical_dev, fail_count_dev)
        switch ((WE + 31) / 32) {
            case 1: LUF_CHECK4(1); break;
            case 2: LUF_CHECK4(2); break;
            case 3: LUF_CHECK4(3); break;
            case 4: LUF_CHECK4(4); break;
            case 5: LUF_CHECK4(5); break;
            case 6: LUF_CHECK4(6); break;
            case 7: LUF_CHECK4(7); break;
            default: LUF_CHECK4(8); break;
        }
#undef LUF_CHECK4
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        done = 4 * ngroups;
    }
    if (done < nshots) {
        const long long rest = nshots - done;
        long long blocks = (rest + wpb - 1) / wpb;
        if (blocks > cap) blocks = cap;
        k_check<<<(int)blocks, block, 0, (cudaStream_t)stream>>>(WE, rest, g->view.west_mask, err_bits_dev + done * WE,
                                                                  corr_bits_dev + done * WE,
                                                                  logical_dev ? logical_dev + done : nullptr, fail_count_dev);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
    }
    return 0;
}

int luf_count_strings(luf_graph* g, int64_t nshots, const uint32_t* err_bits_dev, const uint32_t* corr_bits_dev,
                      int32_t* count_dev, unsigned long long* total_dev, void* stream)
{
    if (!g || !err_bits_dev || !corr_bits_dev || !total_dev || nshots < 0)
        return fail(LUF_ERR_INVALID, "luf_count_strings: bad argument");
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    if (g->wide) {
        const int wpb = 4;
        long long blocks = std::min<long long>((long long)g->sm_count * 4, (nshots + wpb - 1) / wpb);
        void* maps = nullptr;
        if (int rc = wide_slab(g, stream, 1, (size_t)blocks * wpb * (size_t)g->wview.N * 4, &maps)) return rc;
        k_count_strings_wide<<<(int)blocks, wpb * WARP, 0, (cudaStream_t)stream>>>(g->wview, nshots, err_bits_dev, corr_bits_dev,
                                                                              count_dev, total_dev, (int32_t*)maps);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        return 0;
    }
    const uint32_t slice = align16(2u * (uint32_t)g->view.N);
    int wpb = g->max_smem / (int)slice;
    if (wpb < 1) return fail(LUF_ERR_UNSUPPORTED, "luf_count_strings: the partner map of one shot exceeds the shared memory of an SM");
    if (wpb > 8) wpb = 8;
    if (int rc = raise_smem_limit(g->device, g->max_smem, k_count_strings)) return rc;
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_count_strings, wpb * WARP, (size_t)wpb * slice));
    if (per_sm < 1) per_sm = 1;
    long long blocks = (long long)g->sm_count * per_sm;
    const long long need = (nshots + wpb - 1) / wpb;
    if (blocks > need) blocks = need;
    k_count_strings<<<(int)blocks, wpb * WARP, (size_t)wpb * slice, (cudaStream_t)stream>>>(
        g->view, nshots, err_bits_dev, corr_bits_dev, count_dev, total_dev, slice);
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_mc_run(luf_graph* g, int decoder, int flags, const double* class_p_host, uint64_t seed,
               uint64_t shot0, int64_t nshots, unsigned long long* counts_dev, void* stream)
{
    return luf_mc_run_logical(g, decoder, flags, class_p_host, seed, shot0, nshots, counts_dev, nullptr, stream);
}

int luf_mc_run_logical(luf_graph* g, int decoder, int flags, const double* class_p_host, uint64_t seed,
                       uint64_t shot0, int64_t nshots, unsigned long long* counts_dev, uint8_t* logical_dev, void* stream)
{
    if (!g || !class_p_host || !counts_dev || nshots < 0) return fail(LUF_ERR_INVALID, "luf_mc_run: bad argument");
    if (int rc = check_decoder(g, decoder, flags)) return rc;
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const SampleParams sp = make_sample_params(class_p_host, g->view.n_classes, seed);
    LaunchCfg c;
    if (decoder == LUF_DEC_MACAR) {
        // several shots per block around one copy of the neighbour table in shared memory, when it leaves room for at
        // least half the shots that would be resident without it (LUF_CA_TILES=0: the one-shot blocks)
        static const bool env_tiles = !(getenv("LUF_CA_TILES") && atoi(getenv("LUF_CA_TILES")) == 0);
        const uint32_t table_bytes = align16((uint32_t)g->ca_host.nbr.size() * 4u);
        const uint32_t shot_bytes = g->lay_macar_cta.bytes;
        if (env_tiles && use_ca_cta(g, false) && (size_t)table_bytes + 2 * (size_t)shot_bytes <= (size_t)g->max_smem) {
            int tw = ((g->ca_view.N / 16) + 31) & ~31;
            tw = std::max(64, std::min(256, tw));
            if (getenv("LUF_CA_TILE_TW")) tw = std::max(32, std::min(256, atoi(getenv("LUF_CA_TILE_TW")) & ~31));
            int tiles = (int)(((size_t)g->max_smem - table_bytes) / shot_bytes);
            tiles = std::min(tiles, std::min(15, 1024 / tw));
            const int without = (int)std::min<size_t>(2048 / tw, (size_t)(228 * 1024) / (shot_bytes + 1024));
            if (tiles >= 8 && tiles * 3 >= without * 2) {       // (d = 17: 2 shots next to a 132 KB table lose to 4 without it)
                if (int rc = raise_smem_limit(g->device, g->max_smem, k_mc_ca_tiles)) return rc;
                const size_t smem = (size_t)table_bytes + (size_t)tiles * shot_bytes;
                long long blocks = std::min<long long>(g->sm_count, (nshots + tiles - 1) / tiles);
                if (getenv("LUF_DEBUG_PLAN"))
                    fprintf(stderr, "[luf] Macar plan: %d shots of %d threads per block around the neighbour table (%u B), %zu B of shared memory\n", tiles, tw, table_bytes, smem);
                k_mc_ca_tiles<<<(int)blocks, tiles * tw, smem, (cudaStream_t)stream>>>(g->view, g->ca_view, g->lay_macar_cta, table_bytes, tw, sp, flags,
                                                                                      shot0, nshots, counts_dev, logical_dev);
                CUDA_TRY(cudaGetLastError());
                ++g->launches;
                return 0;
            }
        }
        if (use_ca_cta(g, false)) {
            if (int rc = plan_ca_cta_launch(g, k_mc_ca_cta<false>, g->lay_macar_cta.bytes, nshots, &c, g->ca_view.N / 16)) return rc;
            k_mc_ca_cta<false><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->ca_view, g->lay_macar_cta, sp, flags,
                                                                                   shot0, nshots, counts_dev, logical_dev);
        } else {
            if (int rc = plan_launch(g, k_mc_ca<false>, g->lay_macar.bytes, nshots, &c)) return rc;
            k_mc_ca<false><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->ca_view, g->lay_macar, sp, flags,
                                                                               shot0, nshots, counts_dev, logical_dev);
        }
    } else if (decoder == LUF_DEC_ACTIS) {
        if (use_ca_cta(g, true)) {
            if (int rc = plan_ca_cta_launch(g, k_mc_ca_cta<true>, g->lay_actis_cta.bytes, nshots, &c, g->ca_view.N / 12, false)) return rc;
            k_mc_ca_cta<true><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->ca_view, g->lay_actis_cta, sp, flags,
                                                                                  shot0, nshots, counts_dev, logical_dev);
        } else {
            if (int rc = plan_launch(g, k_mc_ca<true>, g->lay_actis.bytes, nshots, &c)) return rc;
            k_mc_ca<true><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->ca_view, g->lay_actis, sp, flags,
                                                                              shot0, nshots, counts_dev, logical_dev);
        }
    } else {
        const bool cta = g->canonical_ok && use_cta_tiles(g, g->lay_fused.bytes, g->lay_fused_cta.bytes);
        // the decoder flags are compile-time constants of the fused kernel
#define LUF_MC_UF(TF, Q, SHOT0, NSHOTS, LOGICAL, QUEUE, QCOUNT) { \
        if (int rc = plan_launch(g, k_mc_uf<TF, Q>, g->lay_fused.bytes, Q ? FAST_CHUNK : (long long)(NSHOTS), &c, 14)) return rc; \
        k_mc_uf<TF, Q><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_fused, sp, flags, SHOT0, NSHOTS, counts_dev, LOGICAL, QUEUE, QCOUNT); }
#define LUF_MC_UF_CTA(TF, Q, SHOT0, NSHOTS, LOGICAL, QUEUE, QCOUNT) { \
        if (int rc = plan_cta_launch(g, k_mc_uf_cta<TF, Q>, g->lay_fused_cta.bytes, Q ? FAST_CHUNK : (long long)(NSHOTS), &c)) return rc; \
        k_mc_uf_cta<TF, Q><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(g->view, g->lay_fused_cta, sp, flags, SHOT0, NSHOTS, counts_dev, LOGICAL, QUEUE, QCOUNT); }
#define LUF_MC_UF_ANY(TF, Q, SHOT0, NSHOTS, LOGICAL, QUEUE, QCOUNT) { \
        if (cta) LUF_MC_UF_CTA(TF, Q, SHOT0, NSHOTS, LOGICAL, QUEUE, QCOUNT) else LUF_MC_UF(TF, Q, SHOT0, NSHOTS, LOGICAL, QUEUE, QCOUNT) }
        // static UF, default / west inclination: the compact fast launch, then the canonical launch over the
        // shots it queued (LUF_FAST=0: the one-launch kernels of round 1)
        static const bool env_fast = !(getenv("LUF_FAST") && atoi(getenv("LUF_FAST")) == 0);
        if (env_fast && g->fast_host.ok && (flags & ~LUF_UF_WEST) == 0) {
            double mu = 0.0;
            for (int cl = 0, lo = 0; cl < g->view.n_classes; ++cl) {
                const int hi = (int)g->fast_host.class_end[cl];
                mu += std::min(1.0, std::max(0.0, class_p_host[cl])) * (hi - lo);
                lo = hi;
            }
            const fast::FLayout fl = make_fast_layout(g->fview.D, g->fview.E, g->fview.DEG, mu);
            FastPlan plan;
            if (fast_plan(g, fl, &plan)) {
                uint32_t* queue = nullptr;
                if (int rc = fast_queue(g, stream, &queue)) return rc;
                // Q1: the HARD shots of the first pass; Q3: those the second pass could not log (canonical launch);
                // Q2: (shot, offset) of the exported ones; counters: |Q1|, |Q3|, |Q2|, -, export cursor (64 bits)
                uint32_t* cnt = queue + FAST_CHUNK;
                uint32_t* q3 = cnt + 16;
                uint32_t* q2 = q3 + FAST_CHUNK;
                uint32_t* qcount = cnt;
                // Second pass over the HARD shots with the merge log, then k_hard_resolve -- instead of the canonical
                // launch, whose reservation merge needs about one iteration per edge of a lattice-spanning cluster (5 ms
                // for ONE shot of d = 17 at p = 0.026).  LUF_FAST_LOG=0 sends the HARD shots straight to the canonical launch.
                const bool env_log = !(getenv("LUF_FAST_LOG") && atoi(getenv("LUF_FAST_LOG")) == 0);
                const fast::FLayout fl2 = make_fast_layout(g->fview.D, g->fview.E, g->fview.DEG, mu, true);
                FastPlan plan2;
                const size_t resolve_warp = 4 * ((size_t)g->fview.D + 4);         // hard_resolve: one word per detector per warp
                const bool log_pass = env_log && !(flags & LUF_UF_WEST) && fast_plan(g, fl2, &plan2) && resolve_warp <= (size_t)g->max_smem;
                // One pass instead (LUF_FAST_ONEPASS=0 / 1 forces it off / on): where HARD shots are common -- from 0.06
                // expected faults per detector on (p = 0.021 on the phenomenological surface code: a percent of the
                // shots) -- every shot writes its merge log to a slab of global memory (one store per newly FULL edge,
                // read back only by a HARD shot: + 3 - 6 % on the first pass, - 2 % on cfg3 at p = 0.01, which therefore
                // keeps the two passes), the HARD shots export their clusters' keys at once and k_hard_sort orders
                // them: no second decode of the heaviest shots.  p = 0.026, M shots/s, one pass / two passes /
                // canonical launch (profiles/r02_hard_resolve.txt): d = 11 41.5 / 38.4 / 37.0, d = 13 20.7 / 18.5 / 17.1,
                // d = 17 7.84 / 6.96 / 3.80, d = 21 2.51 / 2.21 / 1.19, d = 25 1.25 / 1.07 / 0.59.
                const int env_one = getenv("LUF_FAST_ONEPASS") ? atoi(getenv("LUF_FAST_ONEPASS")) : -1;
                uint32_t sortcap = 64;
                while (sortcap < fl2.lcap) sortcap <<= 1;
                const bool one_pass = log_pass && (env_one < 0 ? mu >= 0.06 * g->fview.D : env_one != 0) && 4 * (size_t)sortcap <= (size_t)g->max_smem;
                fast::FLayout fl1 = fl;
                if (one_pass) fl1.lcap = fl2.lcap;           // (the log itself lies in global memory: fl1.elog is not used)
                for (int64_t done = 0; done < nshots; done += FAST_CHUNK) {
                    const long long m = std::min<long long>(FAST_CHUNK, nshots - done);
                    const unsigned long long first = shot0 + (unsigned long long)done;
                    uint8_t* lg = logical_dev ? logical_dev + done : nullptr;
                    fast::FView fv = g->fview;
                    fv.west = (flags & LUF_UF_WEST) ? 1 : 0;
                    if (!fv.west) CUDA_TRY(cudaMemsetAsync(cnt, 0, 32, (cudaStream_t)stream));
                    // LUF_DEBUG_TIMES=1: CUDA events around the launches of a chunk (measurements; synchronises)
                    const bool env_times = getenv("LUF_DEBUG_TIMES") && atoi(getenv("LUF_DEBUG_TIMES")) != 0;
                    cudaEvent_t ev[5] = {0, 0, 0, 0, 0};
                    int nev = 0;
                    auto mark = [&]() { if (env_times && nev < 5) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], (cudaStream_t)stream); ++nev; } };
                    mark();
                    uint32_t *cq = queue, *cqc = qcount;         // what the canonical launch takes
                    FastLogArgs la;
                    memset(&la, 0, sizeof(la));
                    if (log_pass) {
                        // export buffer: 64 words per shot of the chunk, between 2^22 and 2^27 (a shot that finds no room goes on
                        // to the canonical launch)
                        size_t xcap = (size_t)std::min<long long>(1ll << 27, std::max<long long>(1ll << 22, (long long)m * 64));
                        if (getenv("LUF_FAST_XCAP_WORDS")) xcap = (size_t)std::max(64, atoi(getenv("LUF_FAST_XCAP_WORDS")));     // (tests: the no-room path)
                        void* xbuf = nullptr;
                        if (int rc = wide_slab(g, stream, 4, xcap * 4, &xbuf)) return rc;
                        la.x.buf = (uint32_t*)xbuf; la.x.cursor = (uint64_t*)(cnt + 4); la.x.cap = (uint32_t)xcap;
                        la.x.queue = q2; la.x.qcount = cnt + 2;
                    }
                    if (one_pass) {
                        const long long tiles = (long long)g->sm_count * plan.bps * plan.spb;
                        void* glog = nullptr;
                        if (int rc = wide_slab(g, stream, 5, (size_t)tiles * fl1.lcap * 4, &glog)) return rc;
                        la.x.unsorted = 1u; la.x.max_keys = sortcap; la.glog = (uint32_t*)glog; la.mode = 2;
                        if (int rc = fast_launch(g, fv, plan, fl1, sp, first, m, counts_dev, lg, queue, qcount, (cudaStream_t)stream, &la)) return rc;
                        ++g->launches;
                        mark();
                        if (int rc = raise_smem_limit(g->device, g->max_smem, k_hard_sort)) return rc;
                        const size_t ssmem = 4 * (size_t)sortcap;
                        const long long sper = std::max<long long>(1, std::min<long long>(8, (long long)(228 * 1024) / (long long)(ssmem + 1024)));
                        k_hard_sort<<<(int)(g->sm_count * sper), 256, ssmem, (cudaStream_t)stream>>>(fv, la.x);
                        CUDA_TRY(cudaGetLastError());
                        ++g->launches;
                    } else {
                        if (int rc = fast_launch(g, fv, plan, fl, sp, first, m, counts_dev, lg, queue, qcount, (cudaStream_t)stream)) return rc;
                        ++g->launches;
                        mark();
                    }
                    if (log_pass) {
                        if (!one_pass) {                     // second pass over the queued HARD shots, the log in shared memory
                            la.qin = queue; la.qin_count = cnt; la.mode = 1;
                            if (int rc = fast_launch(g, fv, plan2, fl2, sp, first, m, counts_dev, lg, q3, cnt + 1, (cudaStream_t)stream, &la)) return rc;
                            ++g->launches;
                            cq = q3; cqc = cnt + 1;          // only what the second pass could not log
                        }
                        mark();
                        // (the replay of a shot is one dependent chain: as many resident warps as shared memory allows)
                        {
                            if (int rc = raise_smem_limit(g->device, g->max_smem, k_hard_resolve)) return rc;
                            const int rwpb = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)g->max_smem / resolve_warp));
                            const size_t rsmem = (size_t)rwpb * resolve_warp;
                            const long long per_sm = std::max<long long>(1, std::min<long long>(64 / rwpb, (long long)(228 * 1024) / (long long)(rsmem + 1024)));
                            k_hard_resolve<<<(int)(g->sm_count * per_sm), rwpb * WARP, rsmem, (cudaStream_t)stream>>>(g->fview.D, la.x, counts_dev, lg);
                        }
                        CUDA_TRY(cudaGetLastError());
                        ++g->launches;
                        mark();
                    }
                    if (!fv.west && g->wide) {  // the HARD shots of a graph beyond the 16-bit layout: the wide instance
                        WidePlan w;
                        if (int rc = wide_plan(g, k_mc_uf_wide<true>, g->wlay_fused, 4096, stream, &w)) return rc;
                        k_mc_uf_wide<true><<<w.grid, w.threads, w.smem, (cudaStream_t)stream>>>(g->wview, g->wlay_fused, sp, flags, first, m,
                                                                                             counts_dev, lg, cq, cqc, w.slab);
                        CUDA_TRY(cudaGetLastError());
                        ++g->launches;
                    } else if (!fv.west) {      // (the west inclination leaves no HARD shots: luf_uf_fast.cuh)
                        LUF_MC_UF_ANY(0, true, first, m, lg, cq, cqc)
                        CUDA_TRY(cudaGetLastError());
                        ++g->launches;
                    }
                    mark();
                    if (env_times) {
                        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
                        float t[4] = {0, 0, 0, 0};
                        for (int i = 0; i + 1 < nev; ++i) cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]);
                        for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
                        if (log_pass) fprintf(stderr, "[luf] chunk of %lld shots: first pass %.3f ms, %s %.3f ms, hard_resolve %.3f ms, canonical launch %.3f ms\n", m, t[0], one_pass ? "hard_sort" : "log pass", t[1], t[2], t[3]);
                        else fprintf(stderr, "[luf] chunk of %lld shots: first pass %.3f ms, canonical launch %.3f ms\n", m, t[0], t[1]);
                    }
                    if (!fv.west && getenv("LUF_DEBUG_PLAN")) {
                        uint32_t h[6];
                        CUDA_TRY(cudaMemcpyAsync(h, cnt, 24, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
                        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
                        fprintf(stderr, "[luf] HARD shots of %lld: %u queued by the first pass, %u exported for hard_resolve (%llu words), %u left to the canonical launch\n",
                                m, one_pass ? h[0] + h[2] : h[0], h[2], (unsigned long long)h[4] | ((unsigned long long)h[5] << 32), log_pass && !one_pass ? h[1] : h[0]);
                    }
                }
                return 0;
            }
        }
        if (g->wide) {
            WidePlan w;
            if (int rc = wide_plan(g, k_mc_uf_wide<false>, g->wlay_fused, nshots, stream, &w)) return rc;
            k_mc_uf_wide<false><<<w.grid, w.threads, w.smem, (cudaStream_t)stream>>>(g->wview, g->wlay_fused, sp, flags, shot0, nshots,
                                                                                  counts_dev, logical_dev, nullptr, nullptr, w.slab);
            CUDA_TRY(cudaGetLastError());
            ++g->launches;
            return 0;
        }
        // the common flag sets are compile-time constants of the fused kernel; the rest run the generic instance
        switch (flags <= 7 ? flags : -1) {
            case 0: LUF_MC_UF_ANY(0, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            case 1: LUF_MC_UF_ANY(1, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            case 2: LUF_MC_UF_ANY(2, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            case 3: LUF_MC_UF_ANY(3, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            case 4: LUF_MC_UF_ANY(4, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            case 5: LUF_MC_UF_ANY(5, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            case 6: LUF_MC_UF_ANY(6, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            case 7: LUF_MC_UF_ANY(7, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
            default: LUF_MC_UF_ANY(-1, false, shot0, nshots, logical_dev, nullptr, nullptr) break;
        }
#undef LUF_MC_UF
#undef LUF_MC_UF_CTA
#undef LUF_MC_UF_ANY
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

// ------------------------------------------------------------ forward scheme
int luf_forward_attach(luf_graph* g, const luf_forward_desc* d)
{
    if (!g || !d) return fail(LUF_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(g->device));
    const luf_graph_desc_dims dims = {g->view.N, g->view.B, g->view.E, g->view.n_classes};
    {                               // the forward scheme around the wide UF instance (luf_fw_wide.cuh)
        const std::string wwhy = pack_forward_wide(dims, *d, g->wfw_host);
        if (!wwhy.empty()) return fail(LUF_ERR_INVALID, wwhy);
        const PackedForwardWide& p = g->wfw_host;
        fw_wide::FwView& v = g->wfw_view;
        v.N = dims.N; v.B = dims.B; v.E = dims.E; v.WE = g->view.WE; v.WD = g->view.WD; v.c = p.c;
        v.edge_uv = g->wview.edge_uv;
        int rc = 0;
        g->wfw_buffer_tab.perm = nullptr;
        if ((rc = upload(g, p.lower_edge, &v.lower_edge)) || (rc = upload(g, p.commit_mask, &v.commit_mask)) ||
            (rc = upload(g, p.art_node, &v.art_node)) || (rc = upload(g, p.fb_mask, &v.fb_mask)) ||
            (rc = upload(g, p.node_lower, &v.node_lower)) || (rc = upload(g, p.node_side, &v.node_side)) ||
            (rc = upload(g, p.buffer_perm, &g->wfw_buffer_tab.perm32)) ||
            (rc = upload(g, p.buffer_class_end, &g->wfw_buffer_tab.class_end))) return rc;
        g->wfw_buffer_tab.F = p.n_buffer; g->wfw_buffer_tab.n_classes = dims.n_classes;
        g->wfw_lay_uf = make_wide_layout(dims.N, dims.B, dims.E, true, true);
        g->wfw_ext = make_fw_wide_layout(dims.N, dims.E, g->wfw_lay_uf.bytes);
        g->has_forward = true;
        if (!g->canonical_ok) return 0;
    }
    const std::string why = pack_forward(dims, *d, g->fw_host);
    if (!why.empty()) return fail(LUF_ERR_INVALID, why);
    const PackedForward& p = g->fw_host;
    FwView& v = g->fw_view;
    v.N = dims.N; v.B = dims.B; v.E = dims.E; v.WE = g->view.WE; v.WD = g->view.WD; v.c = p.c;
    v.edge_uv = g->view.edge_uv;
    int rc = 0;
    if ((rc = upload(g, p.lower_edge, &v.lower_edge)) || (rc = upload(g, p.commit_mask, &v.commit_mask)) ||
        (rc = upload(g, p.art_node, &v.art_node)) || (rc = upload(g, p.fb_mask, &v.fb_mask)) ||
        (rc = upload(g, p.node_lower, &v.node_lower)) || (rc = upload(g, p.node_side, &v.node_side)) ||
        (rc = upload(g, p.buffer_perm, &g->fw_buffer_tab.perm)) ||
        (rc = upload(g, p.buffer_class_end, &g->fw_buffer_tab.class_end))) return rc;
    g->fw_buffer_tab.F = p.n_buffer; g->fw_buffer_tab.n_classes = dims.n_classes;
    g->fw_lay_uf = make_layout(dims.N, dims.B, dims.E, true, true);
    g->fw_ext_uf = make_fw_layout(dims.N, dims.E, g->fw_lay_uf.bytes);
    g->fw_lay_macar = make_ca_layout(dims.N, dims.B, dims.E, false, true);
    g->fw_ext_macar = make_fw_layout(dims.N, dims.E, g->fw_lay_macar.bytes);
    g->has_forward = true;
    return 0;
}

static int fw_args(luf_graph* g, int decoder, int flags, FwArgs* a)
{
    if (!g->has_forward) return fail(LUF_ERR_INVALID, "luf_forward_attach was not called for this graph");
    if (decoder != LUF_DEC_UF && decoder != LUF_DEC_MACAR)
        return fail(LUF_ERR_UNSUPPORTED, "the forward scheme runs UF or Macar (decoders/uf.py:48-49, luf/__init__.py:22-23)");
    if (int rc = check_decoder(g, decoder, flags)) return rc;
    a->gv = g->view; a->ul = g->fw_lay_uf; a->cv = g->ca_view; a->cl = g->fw_lay_macar; a->fw = g->fw_view;
    a->fl = decoder == LUF_DEC_MACAR ? g->fw_ext_macar : g->fw_ext_uf;
    a->buffer_tab = g->fw_buffer_tab; a->flags = flags; a->n_cleanse = g->fw_host.n_cleanse;
    return 0;
}

// UF streams take the wide instance when the graph does (luf_graph::wide) or when a stream's state in the 16-bit
// layout exceeds the shared memory of an SM (the 16-bit forward kernels keep a warp's slice there)
static bool fw_is_wide(const luf_graph* g, int decoder)
{
    return g->has_forward && decoder == LUF_DEC_UF && (g->wide || g->fw_ext_uf.bytes > (uint32_t)g->max_smem);
}

static int fw_wide_args(luf_graph* g, int flags, FwWideArgs* a)
{
    if (int rc = check_decoder(g, LUF_DEC_UF, flags)) return rc;
    a->gv = g->wview; a->ul = g->wfw_lay_uf; a->fw = g->wfw_view; a->fl = g->wfw_ext; a->buffer_tab = g->wfw_buffer_tab;
    a->flags = flags; a->n_cleanse = g->wfw_host.n_cleanse;
    return 0;
}

int luf_forward_decode(luf_graph* g, int decoder, int flags, int64_t nstreams, int32_t ncycles,
                       const uint32_t* init_buffer_dev, const uint32_t* fresh_bits_dev, int32_t* steps_dev,
                       uint32_t* commit_bits_dev, int32_t* m_dev, void* stream)
{
    if (!g || !init_buffer_dev || !fresh_bits_dev || !steps_dev || !m_dev || nstreams < 0 || ncycles < 0)
        return fail(LUF_ERR_INVALID, "luf_forward_decode: bad argument");
    if (fw_is_wide(g, decoder)) {
        FwWideArgs wa;
        if (int rc = fw_wide_args(g, flags, &wa)) return rc;
        if (nstreams == 0 || ncycles == 0) return 0;
        CUDA_TRY(cudaSetDevice(g->device));
        WidePlan w;
        if (int rc = wide_plan_bytes(g, k_fw_decode_wide, wa.fl.bytes, nstreams, stream, &w)) return rc;
        k_fw_decode_wide<<<w.grid, w.threads, w.smem, (cudaStream_t)stream>>>(wa, nstreams, ncycles, init_buffer_dev, fresh_bits_dev,
                                                                           steps_dev, commit_bits_dev, m_dev, w.slab);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        return 0;
    }
    FwArgs a;
    if (int rc = fw_args(g, decoder, flags, &a)) return rc;
    if (nstreams == 0 || ncycles == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    LaunchCfg c;
    if (decoder == LUF_DEC_MACAR) {
        if (int rc = plan_launch(g, k_fw_decode<true>, a.fl.bytes, nstreams, &c)) return rc;
        k_fw_decode<true><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(a, nstreams, ncycles, init_buffer_dev,
                                                                           fresh_bits_dev, steps_dev, commit_bits_dev, m_dev);
    } else {
        if (int rc = plan_launch(g, k_fw_decode<false>, a.fl.bytes, nstreams, &c)) return rc;
        k_fw_decode<false><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(a, nstreams, ncycles, init_buffer_dev,
                                                                            fresh_bits_dev, steps_dev, commit_bits_dev, m_dev);
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_forward_mc(luf_graph* g, int decoder, int flags, const double* class_p_host, uint64_t seed, uint64_t stream0,
                   int64_t nstreams, int32_t n, unsigned long long* counts_dev, int32_t* steps_dev, void* stream)
{
    if (!g || !class_p_host || !counts_dev || nstreams < 0) return fail(LUF_ERR_INVALID, "luf_forward_mc: bad argument");
    if (n < 1) return fail(LUF_ERR_INVALID, "n must be positive integer.");
    if (fw_is_wide(g, decoder)) {
        FwWideArgs wa;
        if (int rc = fw_wide_args(g, flags, &wa)) return rc;
        if (nstreams == 0) return 0;
        CUDA_TRY(cudaSetDevice(g->device));
        const SampleParams wsp = make_sample_params(class_p_host, g->view.n_classes, seed);
        WidePlan w;
        if (int rc = wide_plan_bytes(g, k_fw_mc_wide, wa.fl.bytes, nstreams, stream, &w)) return rc;
        k_fw_mc_wide<<<w.grid, w.threads, w.smem, (cudaStream_t)stream>>>(wa, wsp, stream0, nstreams, n, counts_dev, steps_dev, w.slab);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
        return 0;
    }
    FwArgs a;
    if (int rc = fw_args(g, decoder, flags, &a)) return rc;
    if (nstreams == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const SampleParams sp = make_sample_params(class_p_host, g->view.n_classes, seed);
    LaunchCfg c;
    if (decoder == LUF_DEC_MACAR) {
        if (int rc = plan_launch(g, k_fw_mc<true>, a.fl.bytes, nstreams, &c)) return rc;
        k_fw_mc<true><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(a, sp, stream0, nstreams, n, counts_dev, steps_dev);
    } else {
        if (int rc = plan_launch(g, k_fw_mc<false>, a.fl.bytes, nstreams, &c)) return rc;
        k_fw_mc<false><<<c.grid, c.block, c.smem, (cudaStream_t)stream>>>(a, sp, stream0, nstreams, n, counts_dev, steps_dev);
    }
    CUDA_TRY(cudaGetLastError());
    ++g->launches;
    return 0;
}

int luf_forward_decode_host(luf_graph* g, int decoder, int flags, int64_t nstreams, int32_t ncycles,
                            const uint32_t* init_buffer_host, const uint32_t* fresh_bits_host, int32_t* steps_host,
                            uint32_t* commit_bits_host, int32_t* m_host)
{
    if (!g || !init_buffer_host || !fresh_bits_host || !steps_host || !m_host || nstreams < 0 || ncycles < 0)
        return fail(LUF_ERR_INVALID, "luf_forward_decode_host: bad argument");
    if (nstreams == 0 || ncycles == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t rows = (size_t)nstreams * ncycles, W = g->view.WE;
    void *ib, *fr, *st, *cm = nullptr, *mm;
    int rc;
    if ((rc = scratch_get(g, 1, (size_t)nstreams * W * 4, &ib)) || (rc = scratch_get(g, 2, rows * W * 4, &fr)) ||
        (rc = scratch_get(g, 4, rows * 8, &st)) || (rc = scratch_get(g, 5, rows * 4, &mm))) return rc;
    if (commit_bits_host && (rc = scratch_get(g, 3, rows * W * 4, &cm))) return rc;
    CUDA_TRY(cudaMemcpyAsync(ib, init_buffer_host, (size_t)nstreams * W * 4, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemcpyAsync(fr, fresh_bits_host, rows * W * 4, cudaMemcpyHostToDevice, 0));
    if ((rc = luf_forward_decode(g, decoder, flags, nstreams, ncycles, (const uint32_t*)ib, (const uint32_t*)fr,
                                 (int32_t*)st, (uint32_t*)cm, (int32_t*)mm, 0))) return rc;
    CUDA_TRY(cudaMemcpyAsync(steps_host, st, rows * 8, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpyAsync(m_host, mm, rows * 4, cudaMemcpyDeviceToHost, 0));
    if (cm) CUDA_TRY(cudaMemcpyAsync(commit_bits_host, cm, rows * W * 4, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

int luf_forward_mc_host(luf_graph* g, int decoder, int flags, const double* class_p_host, uint64_t seed,
                        uint64_t stream0, int64_t nstreams, int32_t n, unsigned long long counts_host[4],
                        int32_t* steps_host)
{
    if (!g || !counts_host || nstreams < 0) return fail(LUF_ERR_INVALID, "luf_forward_mc_host: bad argument");
    if (n < 1) return fail(LUF_ERR_INVALID, "n must be positive integer.");
    if (nstreams == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    void *cnt, *st = nullptr;
    int rc;
    if ((rc = scratch_get(g, 0, 32, &cnt))) return rc;
    if (steps_host && (rc = scratch_get(g, 4, (size_t)nstreams * n * 8, &st))) return rc;
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 32, 0));
    if ((rc = luf_forward_mc(g, decoder, flags, class_p_host, seed, stream0, nstreams, n, (unsigned long long*)cnt,
                             (int32_t*)st, 0))) return rc;
    unsigned long long tmp[4];
    if (st) CUDA_TRY(cudaMemcpyAsync(steps_host, st, (size_t)nstreams * n * 8, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpy(tmp, cnt, 32, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) counts_host[i] += tmp[i];
    return 0;
}

// ------------------------------------------------------- host-buffer variants
int luf_mc_run_host(luf_graph* g, int decoder, int flags, const double* class_p_host, uint64_t seed,
                    uint64_t shot0, int64_t nshots, unsigned long long counts_host[4])
{
    return luf_mc_run_logical_host(g, decoder, flags, class_p_host, seed, shot0, nshots, counts_host, nullptr);
}

int luf_mc_run_logical_host(luf_graph* g, int decoder, int flags, const double* class_p_host, uint64_t seed,
                            uint64_t shot0, int64_t nshots, unsigned long long counts_host[4], uint8_t* logical_host)
{
    if (!g || !counts_host || nshots < 0) return fail(LUF_ERR_INVALID, "luf_mc_run_host: bad argument");
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    void *d = nullptr, *lg = nullptr;
    if (int rc = scratch_get(g, 0, 32, &d)) return rc;
    if (logical_host && nshots > 0)
        if (int rc = scratch_get(g, 3, (size_t)nshots, &lg)) return rc;
    CUDA_TRY(cudaMemsetAsync(d, 0, 32, 0));
    if (int rc = luf_mc_run_logical(g, decoder, flags, class_p_host, seed, shot0, nshots, (unsigned long long*)d,
                                    (uint8_t*)lg, 0)) {
        cudaStreamSynchronize(0);
        return rc;
    }
    unsigned long long tmp[4] = {0, 0, 0, 0};
    if (lg) CUDA_TRY(cudaMemcpyAsync(logical_host, lg, (size_t)nshots, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpy(tmp, d, 32, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) counts_host[i] += tmp[i];
    return 0;
}

int luf_decode_batch_host(luf_graph* g, int decoder, int flags, int64_t nshots, const uint32_t* syn_bits_host,
                          uint32_t* corr_bits_host, uint32_t* growth2b_host, int32_t* steps_host)
{
    if (!g || !syn_bits_host || !corr_bits_host || nshots < 0)
        return fail(LUF_ERR_INVALID, "luf_decode_batch_host: bad argument");
    if (nshots == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t n = (size_t)nshots, WE = g->view.WE, WD = g->view.WD, GW = (g->view.E + 15) / 16;
    void *syn, *corr, *gr = nullptr, *st_dev = nullptr;
    int rc;
    if ((rc = scratch_get(g, 1, n * WD * 4, &syn)) || (rc = scratch_get(g, 2, n * WE * 4, &corr))) return rc;
    if (growth2b_host && (rc = scratch_get(g, 3, n * GW * 4, &gr))) return rc;
    if (steps_host && (rc = scratch_get(g, 4, n * 8, &st_dev))) return rc;
    // Chunks of shots flow through three streams: while one chunk decodes, the next one's
    // syndromes arrive and the previous one's corrections leave (full-duplex PCIe; the copies
    // are asynchronous when the caller's buffers are page-locked, staged by the driver otherwise).
    size_t chunk = (size_t)1 << 16;
    if (getenv("LUF_HOST_CHUNK")) chunk = (size_t)std::max(1024, atoi(getenv("LUF_HOST_CHUNK")));
    for (cudaStream_t& st : g->pipe) if (!st) CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    // whatever happens after the first enqueue, every pipe stream is drained before the call returns: the
    // copies in flight write into the caller's buffers and into scratch the next call may reallocate
    struct Drain {
        luf_graph* g;
        ~Drain() { for (cudaStream_t st : g->pipe) if (st) cudaStreamSynchronize(st); }
    } drain{g};
    int k = 0;
    for (size_t lo = 0; lo < n; lo += chunk, ++k) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        cudaStream_t st = g->pipe[k % 3];
        uint32_t* d_syn = (uint32_t*)syn + lo * WD;
        uint32_t* d_corr = (uint32_t*)corr + lo * WE;
        uint32_t* d_gr = gr ? (uint32_t*)gr + lo * GW : nullptr;
        int32_t* d_st = st_dev ? (int32_t*)st_dev + lo * 2 : nullptr;
        CUDA_TRY(cudaMemcpyAsync(d_syn, syn_bits_host + lo * WD, m * WD * 4, cudaMemcpyHostToDevice, st));
        if ((rc = luf_decode_batch(g, decoder, flags, (int64_t)m, d_syn, d_corr, d_gr, d_st, st))) return rc;
        CUDA_TRY(cudaMemcpyAsync(corr_bits_host + lo * WE, d_corr, m * WE * 4, cudaMemcpyDeviceToHost, st));
        if (d_gr) CUDA_TRY(cudaMemcpyAsync(growth2b_host + lo * GW, d_gr, m * GW * 4, cudaMemcpyDeviceToHost, st));
        if (d_st) CUDA_TRY(cudaMemcpyAsync(steps_host + lo * 2, d_st, m * 8, cudaMemcpyDeviceToHost, st));
    }
    for (cudaStream_t st : g->pipe) CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// Decode into device scratch and keep one byte per shot: the parity of the correction on the logical mask.
int luf_decode_frames(luf_graph* g, int decoder, int flags, int64_t nshots, const uint32_t* syn_bits_dev,
                      uint8_t* frame_dev, int32_t* steps_dev, void* stream)
{
    if (!g || !syn_bits_dev || !frame_dev || nshots < 0) return fail(LUF_ERR_INVALID, "luf_decode_frames: bad argument");
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    void *corr = nullptr, *cnt = nullptr;
    const size_t chunk = (size_t)1 << 18, WE = g->view.WE, WD = g->view.WD;
    if (int rc = wide_slab(g, stream, 2, std::min<size_t>(chunk, (size_t)nshots) * WE * 4 + 16, &corr)) return rc;
    if (int rc = wide_slab(g, stream, 3, 16, &cnt)) return rc;
    const int block = 256, wpb = block / WARP;
    for (size_t lo = 0; lo < (size_t)nshots; lo += chunk) {
        const size_t m = std::min<size_t>(chunk, (size_t)nshots - lo);
        if (int rc = luf_decode_batch(g, decoder, flags, (int64_t)m, syn_bits_dev + lo * WD, (uint32_t*)corr, nullptr,
                                      steps_dev ? steps_dev + 2 * lo : nullptr, stream)) return rc;
        long long blocks = std::min<long long>((long long)g->sm_count * 8, ((long long)m + wpb - 1) / wpb);
        k_check<<<(int)blocks, block, 0, (cudaStream_t)stream>>>((int)WE, (long long)m, g->view.west_mask, nullptr, (const uint32_t*)corr,
                                                                  frame_dev + lo, (unsigned long long*)cnt);
        CUDA_TRY(cudaGetLastError());
        ++g->launches;
    }
    return 0;
}

int luf_decode_frames_host(luf_graph* g, int decoder, int flags, int64_t nshots, const uint32_t* syn_bits_host,
                           uint8_t* frame_host, int32_t* steps_host)
{
    if (!g || !syn_bits_host || !frame_host || nshots < 0) return fail(LUF_ERR_INVALID, "luf_decode_frames_host: bad argument");
    if (nshots == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t n = (size_t)nshots, WD = g->view.WD;
    void *syn, *fr, *st_dev = nullptr;
    int rc;
    if ((rc = scratch_get(g, 1, n * WD * 4, &syn)) || (rc = scratch_get(g, 3, n, &fr))) return rc;
    if (steps_host && (rc = scratch_get(g, 4, n * 8, &st_dev))) return rc;
    const size_t chunk = (size_t)1 << 18;           // as luf_decode_batch_host: copies of one chunk overlap the decode of another
    for (cudaStream_t& st : g->pipe) if (!st) CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct Drain {
        luf_graph* g;
        ~Drain() { for (cudaStream_t st : g->pipe) if (st) cudaStreamSynchronize(st); }
    } drain{g};
    int k = 0;
    for (size_t lo = 0; lo < n; lo += chunk, ++k) {
        const size_t m = n - lo < chunk ? n - lo : chunk;
        cudaStream_t st = g->pipe[k % 3];
        uint32_t* d_syn = (uint32_t*)syn + lo * WD;
        int32_t* d_st = st_dev ? (int32_t*)st_dev + lo * 2 : nullptr;
        CUDA_TRY(cudaMemcpyAsync(d_syn, syn_bits_host + lo * WD, m * WD * 4, cudaMemcpyHostToDevice, st));
        if ((rc = luf_decode_frames(g, decoder, flags, (int64_t)m, d_syn, (uint8_t*)fr + lo, d_st, st))) return rc;
        CUDA_TRY(cudaMemcpyAsync(frame_host + lo, (uint8_t*)fr + lo, m, cudaMemcpyDeviceToHost, st));
        if (d_st) CUDA_TRY(cudaMemcpyAsync(steps_host + lo * 2, d_st, m * 8, cudaMemcpyDeviceToHost, st));
    }
    for (cudaStream_t st : g->pipe) CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int luf_sample_host(luf_graph* g, const double* class_p_host, uint64_t seed, uint64_t shot0, int64_t nshots,
                    uint32_t* err_bits_host, uint32_t* syn_bits_host)
{
    if (!g || !err_bits_host || !syn_bits_host || nshots < 0) return fail(LUF_ERR_INVALID, "luf_sample_host: bad argument");
    if (nshots == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t n = (size_t)nshots, WE = g->view.WE, WD = g->view.WD;
    void *err, *syn;
    int rc;
    if ((rc = scratch_get(g, 1, n * WD * 4, &syn)) || (rc = scratch_get(g, 2, n * WE * 4, &err))) return rc;
    if ((rc = luf_sample(g, class_p_host, seed, shot0, nshots, (uint32_t*)err, (uint32_t*)syn, 0))) return rc;
    CUDA_TRY(cudaMemcpyAsync(err_bits_host, err, n * WE * 4, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpyAsync(syn_bits_host, syn, n * WD * 4, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

int luf_sample_weight_host(luf_graph* g, int32_t weight, uint64_t seed, uint64_t shot0, int64_t nshots,
                           uint32_t* err_bits_host, uint32_t* syn_bits_host)
{
    if (!g || !err_bits_host || !syn_bits_host || nshots < 0) return fail(LUF_ERR_INVALID, "luf_sample_weight_host: bad argument");
    if (nshots == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t n = (size_t)nshots, WE = g->view.WE, WD = g->view.WD;
    void *err, *syn;
    int rc;
    if ((rc = scratch_get(g, 1, n * WD * 4, &syn)) || (rc = scratch_get(g, 2, n * WE * 4, &err))) return rc;
    if ((rc = luf_sample_weight(g, weight, seed, shot0, nshots, (uint32_t*)err, (uint32_t*)syn, 0))) return rc;
    CUDA_TRY(cudaMemcpyAsync(err_bits_host, err, n * WE * 4, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpyAsync(syn_bits_host, syn, n * WD * 4, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

int luf_check_host(luf_graph* g, int64_t nshots, const uint32_t* err_bits_host, const uint32_t* corr_bits_host,
                   uint8_t* logical_host, unsigned long long* fail_count_host)
{
    if (!g || !err_bits_host || !corr_bits_host || !fail_count_host || nshots < 0)
        return fail(LUF_ERR_INVALID, "luf_check_host: bad argument");
    if (nshots == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t n = (size_t)nshots, WE = g->view.WE;
    void *err, *corr, *lg, *cnt;
    int rc;
    if ((rc = scratch_get(g, 1, n * WE * 4, &err)) || (rc = scratch_get(g, 2, n * WE * 4, &corr)) ||
        (rc = scratch_get(g, 3, n, &lg)) || (rc = scratch_get(g, 0, 32, &cnt))) return rc;
    CUDA_TRY(cudaMemcpyAsync(err, err_bits_host, n * WE * 4, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemcpyAsync(corr, corr_bits_host, n * WE * 4, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemsetAsync(cnt, 0, 16, 0));
    if ((rc = luf_check(g, nshots, (const uint32_t*)err, (const uint32_t*)corr, (uint8_t*)lg,
                        (unsigned long long*)cnt, 0))) return rc;
    unsigned long long tmp = 0;
    if (logical_host) CUDA_TRY(cudaMemcpyAsync(logical_host, lg, n, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaMemcpy(&tmp, cnt, 8, cudaMemcpyDeviceToHost));
    *fail_count_host += tmp;
    return 0;
}

// ------------------------------------------------------------------ confidence scores
struct luf_dcs {
    int device = 0, sm_count = 0, max_smem = 0;
    DcsView view{};
    std::vector<void*> allocs;
    void* scratch[4] = {0, 0, 0, 0};
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    std::mutex mu;
};

}  // extern "C"
template <class T>
static int upload_d(luf_dcs* g, const std::vector<T>& v, const T** out)
{
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, v.size() * sizeof(T) + 16));
    g->allocs.push_back(p);
    CUDA_TRY(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (const T*)p;
    return 0;
}
extern "C" {

int luf_dcs_create(const luf_dcs_desc* d, int device, luf_dcs** out)
{
    if (!d || !out || !d->edge_u || !d->edge_v || !d->node_side || d->n_nodes < 1 || d->n_edges < 0)
        return fail(LUF_ERR_INVALID, "luf_dcs_create: bad argument");
    if (d->n_nodes > 65535) return fail(LUF_ERR_UNSUPPORTED, "confidence scores: more than 65535 nodes");
    if (int rc = check_device()) return rc;
    CUDA_TRY(cudaSetDevice(device));
    const int N = d->n_nodes, E = d->n_edges;
    std::vector<uint32_t> uv((size_t)E + 1);
    std::vector<double> w((size_t)E + 1);
    std::vector<uint8_t> ing((size_t)E + 1), side(d->node_side, d->node_side + N);
    double total = 0.0;
    for (int e = 0; e < E; ++e) {
        if (d->edge_u[e] < 0 || d->edge_u[e] >= N || d->edge_v[e] < 0 || d->edge_v[e] >= N)
            return fail(LUF_ERR_INVALID, "luf_dcs_create: edge endpoint out of range");
        uv[e] = (uint32_t)d->edge_u[e] | ((uint32_t)d->edge_v[e] << 16);
        w[e] = d->weight ? d->weight[e] : 1.0;
        if (!(w[e] >= 0.0)) return fail(LUF_ERR_INVALID, "luf_dcs_create: negative edge weight (noise level above 1/2)");
        ing[e] = d->in_growth ? (d->in_growth[e] ? 1 : 0) : 1;
        if (ing[e]) total += w[e];                       // _total_edge_weight: edge order, left to right
    }
    // sweep order of the relaxations: real edges by hop distance from the west meta boundary (breadth first)
    std::vector<uint4> order;
    {
        std::vector<int> off((size_t)N + 1, 0), adj, hop((size_t)N, -1), queue;
        for (int e = 0; e < E; ++e) if (d->edge_u[e] != d->edge_v[e]) { ++off[d->edge_u[e] + 1]; ++off[d->edge_v[e] + 1]; }
        for (int v = 0; v < N; ++v) off[v + 1] += off[v];
        adj.resize((size_t)off[N]);
        std::vector<int> pos(off.begin(), off.end() - 1);
        for (int e = 0; e < E; ++e) if (d->edge_u[e] != d->edge_v[e]) { adj[pos[d->edge_u[e]]++] = d->edge_v[e]; adj[pos[d->edge_v[e]]++] = d->edge_u[e]; }
        for (int v = 0; v < N; ++v) if (side[v] == 1) { hop[v] = 0; queue.push_back(v); }
        for (size_t q = 0; q < queue.size(); ++q) {
            const int v = queue[q];
            for (int i = off[v]; i < off[v + 1]; ++i) if (hop[adj[i]] < 0) { hop[adj[i]] = hop[v] + 1; queue.push_back(adj[i]); }
        }
        std::vector<std::pair<int, uint32_t>> keyed;
        for (int e = 0; e < E; ++e) {
            if (d->edge_u[e] == d->edge_v[e]) continue;
            const int hu = hop[d->edge_u[e]], hv = hop[d->edge_v[e]];
            if (hu < 0 && hv < 0) continue;                // not connected to the west boundary: never relaxed
            keyed.push_back({hu < 0 ? hv : hv < 0 ? hu : (hu < hv ? hu : hv), (uint32_t)e});
        }
        std::sort(keyed.begin(), keyed.end());
        for (auto& k : keyed) {
            const uint32_t e = k.second;
            uint64_t bits;
            memcpy(&bits, &w[e], 8);
            order.push_back(uint4{uv[e], e, (uint32_t)bits, (uint32_t)(bits >> 32)});
        }
        order.push_back(uint4{0, 0, 0, 0});
    }
    luf_dcs* g = new (std::nothrow) luf_dcs;
    if (!g) return fail(LUF_ERR_CUDA, "out of host memory");
    g->device = device;
    cudaDeviceProp prop;
    cudaError_t ce = cudaGetDeviceProperties(&prop, device);
    if (ce != cudaSuccess) { delete g; return fail(LUF_ERR_CUDA, cudaGetErrorString(ce)); }
    g->sm_count = prop.multiProcessorCount; g->max_smem = (int)prop.sharedMemPerBlockOptin;
    g->view.N = N; g->view.E = E; g->view.WN = (N + 31) / 32; g->view.total_weight = total;
    int rc;
    g->view.E_relax = (int)order.size() - 1;
    if ((rc = upload_d(g, order, &g->view.rec)) ||
        (rc = upload_d(g, uv, &g->view.uv)) || (rc = upload_d(g, w, &g->view.weight)) ||
        (rc = upload_d(g, ing, &g->view.in_growth)) || (rc = upload_d(g, side, &g->view.side))) {
        luf_dcs_destroy(g);
        return rc;
    }
    *out = g;
    return 0;
}

int luf_dcs_destroy(luf_dcs* g)
{
    if (!g) return 0;
    cudaSetDevice(g->device);
    for (void* p : g->allocs) cudaFree(p);
    for (void* p : g->scratch) if (p) cudaFree(p);
    delete g;
    return 0;
}

int luf_dcs_run(luf_dcs* g, int64_t nshots, const uint32_t* growth2b_dev, int64_t stride_words, int32_t growth_words,
                double* swim_dev, double* uef_dev, int32_t* clustered_nodes_dev, void* stream)
{
    if (!g || !growth2b_dev || nshots < 0 || growth_words < 0 || stride_words < growth_words)
        return fail(LUF_ERR_INVALID, "luf_dcs_run: bad argument");
    if (nshots == 0) return 0;
    CUDA_TRY(cudaSetDevice(g->device));
    const size_t bytes = 8u * (size_t)g->view.N + 4u * (size_t)g->view.WN + 4u * (size_t)((g->view.E + 15) / 16) + 16u;
    if (bytes > (size_t)g->max_smem)
        return fail(LUF_ERR_UNSUPPORTED, "confidence scores: one shot's distances (" + std::to_string(bytes) + " B) exceed the shared memory of an SM");
    if (int rc = raise_smem_limit(g->device, g->max_smem, k_dcs)) return rc;
    int per_sm = 0;
    // 128 threads per growth state measured best on cfg3 (4.8 ms per 65536 states; 64: 5.9, 256: 5.8)
    int threads = getenv("LUF_DCS_THREADS") ? atoi(getenv("LUF_DCS_THREADS")) : 128;
    if (threads < 32 || threads > 256 || (threads & 31)) threads = 128;
    const int max_passes = getenv("LUF_DCS_MAX_PASSES") ? atoi(getenv("LUF_DCS_MAX_PASSES")) : g->view.N + 1;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dcs, threads, bytes));
    if (per_sm < 1) return fail(LUF_ERR_UNSUPPORTED, "k_dcs does not fit on an SM");
    long long blocks = (long long)g->sm_count * per_sm;
    if (blocks > nshots) blocks = nshots;
    k_dcs<<<(int)blocks, threads, bytes, (cudaStream_t)stream>>>(g->view, nshots, growth2b_dev, stride_words, growth_words,
                                                                max_passes, swim_dev, uef_dev, clustered_nodes_dev);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int luf_dcs_run_host(luf_dcs* g, int64_t nshots, const uint32_t* growth2b_host, int64_t stride_words, int32_t growth_words,
                     double* swim_host, double* uef_host, int32_t* clustered_nodes_host)
{
    if (!g || !growth2b_host || nshots < 0) return fail(LUF_ERR_INVALID, "luf_dcs_run_host: bad argument");
    if (nshots == 0) return 0;
    std::lock_guard<std::mutex> lock(g->mu);
    CUDA_TRY(cudaSetDevice(g->device));
    DrainDefaultStream drain_on_exit;
    const size_t n = (size_t)nshots;
    const size_t want[4] = {n * (size_t)stride_words * 4, n * 8, n * 8, n * 4};
    for (int i = 0; i < 4; ++i)
        if (want[i] > g->scratch_bytes[i]) {
            if (g->scratch[i]) CUDA_TRY(cudaFree(g->scratch[i]));
            g->scratch[i] = nullptr; g->scratch_bytes[i] = 0;
            CUDA_TRY(cudaMalloc(&g->scratch[i], want[i]));
            g->scratch_bytes[i] = want[i];
        }
    CUDA_TRY(cudaMemcpyAsync(g->scratch[0], growth2b_host, want[0], cudaMemcpyHostToDevice, 0));
    if (int rc = luf_dcs_run(g, nshots, (const uint32_t*)g->scratch[0], stride_words, growth_words,
                             swim_host ? (double*)g->scratch[1] : nullptr, uef_host ? (double*)g->scratch[2] : nullptr,
                             clustered_nodes_host ? (int32_t*)g->scratch[3] : nullptr, 0)) return rc;
    if (swim_host) CUDA_TRY(cudaMemcpyAsync(swim_host, g->scratch[1], n * 8, cudaMemcpyDeviceToHost, 0));
    if (uef_host) CUDA_TRY(cudaMemcpyAsync(uef_host, g->scratch[2], n * 8, cudaMemcpyDeviceToHost, 0));
    if (clustered_nodes_host) CUDA_TRY(cudaMemcpyAsync(clustered_nodes_host, g->scratch[3], n * 4, cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return 0;
}

}  // extern "C"
