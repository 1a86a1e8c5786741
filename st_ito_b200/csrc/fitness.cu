// Fitness epilogue (K4): NaN scrub + L2 normalisation of the raw mid/side embeddings
// (get_param_embeds, st_ito/utils.py:492-501) and the cosine fitness of evaluate()
// (st_ito/style_transfer.py:544-571): fitness = mean over {mid, side} of -cos(out, target).
#include <cfloat>
#include <cstdio>

#include "stito_internal.h"

namespace stito {

namespace {

__device__ __forceinline__ float block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.0f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

// flags[0] = any NaN in mid, flags[1] = any NaN in side (whole batch, like torch.isnan(x).any())
__global__ void nan_flags_kernel(const float *mid, const float *side, int64_t n, int *flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (isnan(mid[i])) atomicOr(flags + 0, 1);
    if (isnan(side[i])) atomicOr(flags + 1, 1);
}

__device__ __forceinline__ float nan_to_num(float v) {
    if (isnan(v)) return 0.0f;
    if (isinf(v)) return v > 0 ? FLT_MAX : -FLT_MAX;
    return v;
}

// one CTA per item: scrub (mid first, ELSE side: utils.py:492-497), then x / max(||x||, 1e-12)
__global__ void __launch_bounds__(128) embed_normalize_kernel(float *mid, float *side, int E,
                                                              const int *flags) {
    __shared__ float red[4];
    const int b = blockIdx.x;
    const bool scrub_mid = flags[0] != 0;
    const bool scrub_side = !scrub_mid && flags[1] != 0;
    for (int h = 0; h < 2; ++h) {
        float *v = (h ? side : mid) + (int64_t)b * E;
        const bool scrub = h ? scrub_side : scrub_mid;
        float ss = 0.0f;
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            float t = v[e];
            if (scrub) { t = nan_to_num(t); v[e] = t; }
            ss = fmaf(t, t, ss);
        }
        ss = block_sum(ss, red);
        const float denom = fmaxf(sqrtf(ss), 1e-12f);
        for (int e = threadIdx.x; e < E; e += blockDim.x) v[e] = v[e] / denom;
    }
}

// torch.cosine_similarity(x, y, dim=-1, eps=1e-8) = x.y / max(||x|| * ||y||, eps)
__global__ void __launch_bounds__(128) fitness_kernel(const float *mid, const float *side,
                                                      const float *tgt_mid, const float *tgt_side, int E,
                                                      float *fitness) {
    __shared__ float red[4];
    const int b = blockIdx.x;
    float d[2];
    for (int h = 0; h < 2; ++h) {
        const float *x = (h ? side : mid) + (int64_t)b * E;
        const float *y = h ? tgt_side : tgt_mid;
        float xy = 0.f, xx = 0.f, yy = 0.f;
        for (int e = threadIdx.x; e < E; e += blockDim.x) {
            const float a = x[e], c = y[e];
            xy = fmaf(a, c, xy);
            xx = fmaf(a, a, xx);
            yy = fmaf(c, c, yy);
        }
        xy = block_sum(xy, red);
        xx = block_sum(xx, red);
        yy = block_sum(yy, red);
        d[h] = -(xy / fmaxf(sqrtf(xx * yy), 1e-8f));
    }
    if (threadIdx.x == 0) fitness[b] = (d[0] + d[1]) / 2.0f;
}

// ---- fitness + all-gather over NVLink peer memory (multi-GPU: the population is sharded over the ranks) -------------
// The reference has no collective on this path; sharding the population needs exactly one per generation: every rank needs
// all P fitness values for its replica of the CMA-ES.  Instead of a NCCL all-gather after the fitness kernel, the fitness
// kernel itself STORES each value into the gather buffer of every peer GPU (peer pointers from CUDA IPC handles; P2P stores
// over NVLink / NVSwitch), and its last CTA publishes "rank r, epoch e is complete" to every peer (st.release.sys after a
// system fence).  A one-warp kernel then waits (ld.acquire.sys, bounded) until all ranks' flags carry the epoch; the gathered
// vector is local memory by then.  Buffers are double-buffered by epoch parity: a rank can be at most one generation ahead
// of a peer (it needs that peer's values of the current generation to proceed).
__global__ void __launch_bounds__(128) fitness_scatter_kernel(const float *mid, const float *side, const float *tgt_mid,
                                                              const float *tgt_side, int E, int n_local, int lo,
                                                              GatherPeers peers, int parity, int epoch, int *done_counter) {
    __shared__ float red[4];
    __shared__ int is_last;
    const int b = blockIdx.x;
    if (b < n_local) {
        float d[2];
        for (int h = 0; h < 2; ++h) {
            const float *x = (h ? side : mid) + (int64_t)b * E;
            const float *y = h ? tgt_side : tgt_mid;
            float xy = 0.f, xx = 0.f, yy = 0.f;
            for (int e = threadIdx.x; e < E; e += blockDim.x) {
                const float a = x[e], c = y[e];
                xy = fmaf(a, c, xy);
                xx = fmaf(a, a, xx);
                yy = fmaf(c, c, yy);
            }
            xy = block_sum(xy, red);
            xx = block_sum(xx, red);
            yy = block_sum(yy, red);
            d[h] = -(xy / fmaxf(sqrtf(xx * yy), 1e-8f));
        }
        const float f = (d[0] + d[1]) / 2.0f;  // same arithmetic as fitness_kernel
        if ((int)threadIdx.x < peers.world)    // one lane per destination rank (own rank included): P2P store
            peers.buf[threadIdx.x][(size_t)parity * peers.capacity + lo + b] = f;
    }
    // last CTA to finish publishes the epoch to every peer (an empty shard launches one CTA that only does this)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(done_counter, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __threadfence_system();
        if ((int)threadIdx.x < peers.world) {
            int *flag = peers.flag[threadIdx.x] + parity * kGatherMaxWorld + peers.rank;
            asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
        }
        if (threadIdx.x == 0) *done_counter = 0;
    }
}

__global__ void gather_wait_kernel(const int *flags /* local [2][kGatherMaxWorld] */, int world, int parity, int epoch) {
    const int r = threadIdx.x;
    if (r >= world) return;
    const int *flag = flags + parity * kGatherMaxWorld + r;
    unsigned long long t0 = 0;
    for (;;) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == epoch) return;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t0 == 0) t0 = t1;
        if (t1 - t0 > 20000000000ull) {  // 20 s: a peer rank died or fell out of step -- fail loudly, never hang the GPU
            printf("libstito: fitness gather timeout waiting for rank %d (epoch %d, saw %d)\n", r, epoch, v);
            __trap();
        }
        __nanosleep(500);
    }
}

}  // namespace

cudaError_t launch_fitness_gather(cudaStream_t st, const float *mid, const float *side, const float *tgt_mid,
                                  const float *tgt_side, int n_local, int E, int lo, const GatherPeers &peers, int parity,
                                  int epoch, int *done_counter, const int *local_flags, int *launches) {
    fitness_scatter_kernel<<<n_local > 0 ? n_local : 1, 128, 0, st>>>(mid, side, tgt_mid, tgt_side, E, n_local, lo, peers, parity,
                                                                     epoch, done_counter);
    gather_wait_kernel<<<1, 32, 0, st>>>(local_flags, peers.world, parity, epoch);
    *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_embed_normalize(cudaStream_t st, float *mid, float *side, int B, int E, int *flags,
                                   int *launches) {
    cudaError_t e = cudaMemsetAsync(flags, 0, 2 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int64_t n = (int64_t)B * E;
    nan_flags_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mid, side, n, flags);
    embed_normalize_kernel<<<B, 128, 0, st>>>(mid, side, E, flags);
    *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_fitness(cudaStream_t st, const float *mid, const float *side, const float *tgt_mid,
                           const float *tgt_side, int B, int E, float *fitness, int *launches) {
    fitness_kernel<<<B, 128, 0, st>>>(mid, side, tgt_mid, tgt_side, E, fitness);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace stito
