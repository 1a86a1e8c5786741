// AFx-Rep encoder body on 5th-generation tensor cores (precision mode 1).
//
// Every 3x3 convolution with Cin >= 64 is an implicit GEMM   D[pixels, Cout] = A[pixels, 9*Cin] * B[9*Cin, Cout]
// issued as tcgen05.mma (kind::f16, fp32 accumulation in TMEM) by one elected thread per CTA:
//   * A is never materialised: for tap (kh, kw) and a 64-channel slab, the 128-pixel x 64-channel operand
//     tile is ONE 4-D TMA box (C, W, H, N) at spatial offset (kh-1, kw-1); TMA's out-of-bounds zero fill
//     IS the convolution's zero padding.  B tiles ([Cout][9*Cin], K-major) are 2-D TMA boxes.  Both land
//     in shared memory in the 128-byte-swizzled K-major layout the UMMA descriptors expect.
//   * error-compensated fp16x3: activations and (pre-scaled) weights are stored as hi + lo fp16 pairs and
//     every K-slab issues  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  into an fp32 TMEM accumulator: ~22 mantissa
//     bits of each operand (single-pass fp16/bf16 misses the 1e-4 embedding gate).
//     Activations are stored pre-multiplied by 2^6 (exact) so their fp16 lo parts stay out of the
//     subnormal range.
//   * the tensor core's fp32 accumulation truncates, so the K loop is chunked: each chunk of a few K-slabs
//     accumulates in one half of a 2-deep TMEM ring and is then added, round-to-nearest, to fp32 register
//     accumulators by the epilogue warps while the next chunk is being multiplied (see the kernel comment).
//   * persistent CTAs (one per SM) walk the tile list; warp 0 = TMA producer, warp 1 = TMEM allocator + MMA
//     issuer, warps 2-9 = drain/epilogue; smem full/empty and TMEM full/empty mbarrier rings.
// Replaces ConvBlock / Cnn14.forward's conv stack (st_ito/models/panns.py:25-80, 250-261).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <map>
#include <string>
#include <utility>

#include "encoder_tc.h"

namespace stito {

namespace {

thread_local std::string g_tc_err;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must fail loudly (trap -> cudaErrorLaunchFailure), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try_wait(bar, parity)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) {  // 4 s
            printf("libstito: mbarrier timeout (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 operands, fp32 accumulate), one CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor: K-major, 128-byte swizzle, 128-byte rows, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30) = 0, SBO>>4 [32,46) = 64, version [46,48) = 1,
//  layout_type [61,64) = 2 (SWIZZLE_128B)).
// SLABK = 32 uses the 64-byte swizzle instead (64-byte rows, 8-row groups 512 B apart, layout_type = 4).
template <int SLABK>
__device__ __forceinline__ uint64_t make_sdesc_k(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((8 * SLABK * 2) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(SLABK == 64 ? 2 : 4) << 61;
    return d;
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=f16 [7,10)=0, B=f16 [10,13)=0,
// A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Activations are stored as fp16 hi/lo pairs of (value * 2^shift): the shift is per layer (TcWorkspace::act_shift,
// default 6) and calibrated from the measured per-layer maxima so that the pairs neither overflow (65504) nor lose
// their lo parts to the subnormal range, whatever the scale of a checkpoint's BatchNorm statistics.
constexpr float kHalfMax = 65504.0f;
constexpr int kWeightTop = 13;  // weights are stored as fp16 hi/lo pairs of w * 2^shift with max |w| * 2^shift in (2^12, 2^13]

__device__ __forceinline__ void publish_amax(unsigned *amax, float mx) {
    // warp maximum of the (non-negative, already scaled) outputs -> one atomic per warp; float bits order like unsigned
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (amax != nullptr && (threadIdx.x & 31) == 0 && mx > 0.0f) atomicMax(amax, __float_as_uint(mx));
}

struct ConvTcParams {
    const float *bias;   // [Cout]
    float unscale;       // 2^-shift of the weight pre-scaling / 2^shift of the input activations
    float out_scale;     // 2^shift of this layer's output pairs (x 0.25 with the fused average pool); 1 for fp32 output
    __half *out_hi, *out_lo;  // NHWC fp16 pair, or
    float *out_f32;           // NHWC fp32 (when non-null)
    int N, H, W, Cin, Cout;
    int BW, BH, IPT;          // spatial tile: BW x BH pixels x IPT images (<= 128 rows)
    int tilesW, tilesH;
    int mtiles, ntiles;       // tile grid; the persistent CTAs walk tile = m + mtiles * n
    int chunk_slabs;          // K-slabs accumulated inside TMEM before the fp32 drain (see below)
    int pool;                 // 1: fuse the block's 2x2 average pool into the epilogue (output [N][H/2][W/2][Cout])
    // GEMM mode (Winograd): 16 independent products M_x[Tp, Cout] = V_x[Tp, Cin] * U_x[Cout, Cin]^T; A / B are 2-D maps
    // over [16 * Tp][Cin] and [16 * Cout][Cin]; the raw fp32 accumulators go to out_f32[(x * Tp + row) * Cout + col]
    int gemm, gemm_Tp;
    // Compensation of the tensor core's truncating fp32 accumulate: every drained chunk partial sum is multiplied by
    // 1 + trunc_comp * (number of K = 16 MMA steps accumulated into it); see chunk_comp().
    float trunc_comp;
    // fp16 range guard: anything above kHalfMax is clamped and *overflow (device int, nullable) is raised; *amax (device,
    // nullable) receives the largest stored (scaled) value of the layer as float bits.  The caller re-calibrates the
    // per-layer shifts from amax and redoes the evaluation after an overflow.
    int *overflow;
    unsigned *amax;
    // Two-pass variants (per-layer developer knobs STITO_TC_DROP_ALO / STITO_TC_DROP_BLO, bit l = conv layer l): skip the
    // A_lo * B_hi (activation low parts) or the A_hi * B_lo (weight low parts) correction MMAs of this layer.
    int drop_alo, drop_blo;
};

constexpr int kTileM = 128;
constexpr int kSlabK = 64;                       // fp16 elements = one 128-byte swizzle row
constexpr int kEpiWarps = 8;                     // 4 TMEM lane quarters x 2 column halves
constexpr int kConvThreads = 32 * (2 + kEpiWarps);
constexpr int kAccStages = 2;                    // TMEM accumulator ring

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Final epilogue of one tile for one epilogue thread (one output pixel row x kCols channels held in acc):
// * unscale + bias, ReLU, optional fused 2x2 average pool, fp16 hi/lo split (or fp32) and the NHWC store.
template <int kCols>
__device__ __forceinline__ void store_tile(const ConvTcParams &p, const float (&acc)[kCols], int nt, int mt,
                                           int tiles_per_group, int img, int rr, int cc, int BN, int col0,
                                           const float *sbias, float &mx) {
    const int grp = mt / tiles_per_group, rem = mt - grp * tiles_per_group;
    const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
    const int n0 = grp * p.IPT, h0 = th * p.BH, w0 = tw * p.BW;
    const int hh = h0 + rr, ww = w0 + cc;
    const int cbase = nt * BN + col0;
    const float *bias = sbias + cbase;
    bool store;
    int64_t obase;
    if (p.pool) {
        // 2x2 average pool (floor): the window (rr, rr+1) x (cc, cc+1) lives in lanes {l, l^1, l^BW, l^BW^1}
        // of this warp (tile rows are BW <= 16 pixels wide); the lane of the even corner stores.
        const int Ho = p.H >> 1, Wo = p.W >> 1;
        store = img < p.IPT && (n0 + img) < p.N && !(rr & 1) && !(cc & 1) && (hh >> 1) < Ho && (ww >> 1) < Wo;
        obase = (((int64_t)(n0 + img) * Ho + (hh >> 1)) * Wo + (ww >> 1)) * p.Cout + cbase;
    } else {
        store = img < p.IPT && (n0 + img) < p.N && hh < p.H && ww < p.W;
        obase = (((int64_t)(n0 + img) * p.H + hh) * p.W + ww) * p.Cout + cbase;
    }
#pragma unroll
    for (int j0 = 0; j0 < kCols; j0 += 8) {
        float y[8];
        const float4 bq0 = *reinterpret_cast<const float4 *>(bias + j0);
        const float4 bq1 = *reinterpret_cast<const float4 *>(bias + j0 + 4);
        const float bv[8] = {bq0.x, bq0.y, bq0.z, bq0.w, bq1.x, bq1.y, bq1.z, bq1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = fmaxf(fmaf(acc[j0 + j], p.unscale, bv[j]), 0.0f);
            if (p.pool) {
                v = v + __shfl_xor_sync(0xffffffffu, v, 1);
                v = v + __shfl_xor_sync(0xffffffffu, v, p.BW);
            }
            y[j] = v * p.out_scale;
        }
        if (!store) continue;
        if (p.out_f32 != nullptr) {
            float4 *dst = reinterpret_cast<float4 *>(p.out_f32 + obase + j0);
            dst[0] = make_float4(y[0], y[1], y[2], y[3]);
            dst[1] = make_float4(y[4], y[5], y[6], y[7]);
        } else {
            uint32_t hi[4], lo[4];
            bool ovf = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // post-ReLU: only the upper bound can be hit (NaN compares false and passes through)
                mx = fmaxf(mx, y[j]);
                ovf |= y[j] > kHalfMax;
                y[j] = fminf(y[j], kHalfMax);
            }
            if (ovf && p.overflow != nullptr) atomicOr(p.overflow, 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half h0v = __float2half_rn(y[2 * j]), h1v = __float2half_rn(y[2 * j + 1]);
                const __half l0v = __float2half_rn(y[2 * j] - __half2float(h0v));
                const __half l1v = __float2half_rn(y[2 * j + 1] - __half2float(h1v));
                hi[j] = (uint32_t)__half_as_ushort(h0v) | ((uint32_t)__half_as_ushort(h1v) << 16);
                lo[j] = (uint32_t)__half_as_ushort(l0v) | ((uint32_t)__half_as_ushort(l1v) << 16);
            }
            *reinterpret_cast<uint4 *>(p.out_hi + obase + j0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(p.out_lo + obase + j0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// Persistent, warp-specialised implicit-GEMM convolution.
//
//   warp 0      TMA producer: for every tile and K-slab (tap, 64-channel slab) loads A_hi, A_lo (4-D boxes of
//               the NHWC activations at the tap's spatial offset; out-of-bounds zero fill = conv padding) and
//               B_hi, B_lo into a STAGES-deep shared-memory ring; runs ahead across tile boundaries.
//   warp 1      MMA issuer: 12 tcgen05.mma per slab (A_lo*B_hi, A_hi*B_lo, A_hi*B_hi) into ONE fp32 TMEM
//               accumulator.  The tensor core's fp32 accumulate truncates (round-toward-zero): over the
//               1152 sequential adds of a K = 18432 layer that bias reaches ~5e-5 relative.  So the K loop
//               is cut into chunks of `chunk_slabs` slabs; every chunk starts a fresh accumulator in the
//               other half of a 2-deep TMEM ring ...
//   warps 2-9   ... and the epilogue warps drain each finished chunk (tcgen05.ld) and add it to per-thread
//               fp32 register accumulators with round-to-nearest on the CUDA cores while the next chunk is
//               being multiplied.  After the last chunk: * unscale + bias, ReLU, fp16 hi/lo split (or fp32)
//               and the NHWC store.  The drain of tile i's last chunk overlaps the MMAs of tile i+1.
template <int BN, int STAGES, int SLABK, int MT>
__global__ void __launch_bounds__(kConvThreads, 1) conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmAh,
                                                                     const __grid_constant__ CUtensorMap tmAl,
                                                                     const __grid_constant__ CUtensorMap tmBh,
                                                                     const __grid_constant__ CUtensorMap tmBl,
                                                                     const ConvTcParams p) {
    constexpr int kABytes = kTileM * SLABK * 2;
    constexpr int kBBytes = BN * SLABK * 2;
    constexpr int kStageBytes = MT * 2 * kABytes + 2 * kBBytes;
    constexpr int kAccCols = MT * BN;             // TMEM columns of one accumulator stage
    constexpr int kCols = MT == 1 ? BN / 2 : BN;  // accumulator columns owned by one epilogue thread
    static_assert(kAccStages * kAccCols <= 512 && kCols <= 128, "TMEM / register budget");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * kStageBytes);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;
    uint64_t *acc_empty = acc_full + kAccStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccStages);
    float *sbias = reinterpret_cast<float *>(smem + STAGES * kStageBytes + 256);  // [Cout] (<= 2048 floats)
    if (!p.gemm)
        for (int i = threadIdx.x; i < p.Cout; i += kConvThreads) sbias[i] = __ldg(p.bias + i);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_group = p.tilesW * p.tilesH;
    const int mgroups = (p.mtiles + MT - 1) / MT;   // work item = MT consecutive M tiles x one N tile (they share B)
    const int total_tiles = mgroups * p.ntiles * (p.gemm ? 16 : 1);
    const int cpt = p.Cin / SLABK;  // channel slabs per tap
    const int nslabs = (p.gemm ? 1 : 9) * cpt;
    const int nchunks = (nslabs + p.chunk_slabs - 1) / p.chunk_slabs;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kAccStages * kAccCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---------------- TMA producer
            const uint32_t a_box_bytes = (uint32_t)(p.BW * p.BH * p.IPT) * SLABK * 2;
            const uint32_t tx_bytes = MT * 2 * a_box_bytes + 2 * kBBytes;
            uint32_t g = 0;  // slabs issued so far (ring position), continues across tiles
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                if (p.gemm) {  // work item = (x, nt, mt), mt fastest
                    const int per_x = mgroups * p.ntiles;
                    const int xi = tile / per_x, r2 = tile - xi * per_x;
                    const int nt = r2 / mgroups, mt = r2 - nt * mgroups;
                    for (int s = 0; s < nslabs; ++s, ++g) {
                        const uint32_t stage = g % STAGES, it = g / STAGES;
                        mbar_wait(&empty[stage], (it & 1) ^ 1);
                        uint8_t *sb = smem + stage * kStageBytes;
                        mbar_expect_tx(&full[stage], 2 * kABytes + 2 * kBBytes);
                        tma_load_2d(&tmAh, &full[stage], sb, s * SLABK, xi * p.gemm_Tp + mt * kTileM);
                        tma_load_2d(&tmAl, &full[stage], sb + kABytes, s * SLABK, xi * p.gemm_Tp + mt * kTileM);
                        tma_load_2d(&tmBh, &full[stage], sb + MT * 2 * kABytes, s * SLABK, xi * p.Cout + nt * BN);
                        tma_load_2d(&tmBl, &full[stage], sb + MT * 2 * kABytes + kBBytes, s * SLABK, xi * p.Cout + nt * BN);
                    }
                    continue;
                }
                const int nt = tile / mgroups, mg = tile - nt * mgroups;
                int n0[MT], h0[MT], w0[MT];
#pragma unroll
                for (int i = 0; i < MT; ++i) {  // an M tile past the end (odd tile count) lands out of bounds: zero fill
                    const int mt = mg * MT + i;
                    const int grp = mt / tiles_per_group, rem = mt - grp * tiles_per_group;
                    const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
                    n0[i] = grp * p.IPT; h0[i] = th * p.BH; w0[i] = tw * p.BW;
                }
                for (int s = 0; s < nslabs; ++s, ++g) {
                    const uint32_t stage = g % STAGES, it = g / STAGES;
                    mbar_wait(&empty[stage], (it & 1) ^ 1);  // passes at once on the first lap
                    uint8_t *sb = smem + stage * kStageBytes;
                    mbar_expect_tx(&full[stage], tx_bytes);
                    const int tap = s / cpt, c0 = (s - tap * cpt) * SLABK;
                    const int kh = tap / 3, kw = tap - kh * 3;
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        tma_load_4d(&tmAh, &full[stage], sb + i * 2 * kABytes, c0, w0[i] + kw - 1, h0[i] + kh - 1, n0[i]);
                        tma_load_4d(&tmAl, &full[stage], sb + i * 2 * kABytes + kABytes, c0, w0[i] + kw - 1, h0[i] + kh - 1, n0[i]);
                    }
                    tma_load_2d(&tmBh, &full[stage], sb + MT * 2 * kABytes, tap * p.Cin + c0, nt * BN);
                    tma_load_2d(&tmBl, &full[stage], sb + MT * 2 * kABytes + kBBytes, tap * p.Cin + c0, nt * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---------------- MMA issuer
            constexpr uint32_t idesc = make_idesc(kTileM, BN);
            uint32_t g = 0, c = 0;  // slab / chunk counters (ring positions)
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int s0 = 0; s0 < nslabs; s0 += p.chunk_slabs, ++c) {
                    const uint32_t a = c % kAccStages;
                    mbar_wait(&acc_empty[a], ((c / kAccStages) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d0 = tmem_base + a * kAccCols;
                    const int s1 = min(s0 + p.chunk_slabs, nslabs);
                    for (int s = s0; s < s1; ++s, ++g) {
                        const uint32_t stage = g % STAGES, it = g / STAGES;
                        mbar_wait(&full[stage], it & 1);
                        tc_fence_after();
                        const uint32_t sb = smem_u32(smem + stage * kStageBytes);
                        const uint64_t b_hi = make_sdesc_k<SLABK>(sb + MT * 2 * kABytes);
                        const uint64_t b_lo = make_sdesc_k<SLABK>(sb + MT * 2 * kABytes + kBBytes);
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
                            const uint32_t d = d0 + i * BN;
                            const uint64_t a_hi = make_sdesc_k<SLABK>(sb + i * 2 * kABytes);
                            const uint64_t a_lo = make_sdesc_k<SLABK>(sb + i * 2 * kABytes + kABytes);
                            // small terms first: they meet the accumulator while it is small
                            uint32_t acc_on = s > s0 ? 1u : 0u;
                            if (!p.drop_alo) {
#pragma unroll
                                for (int k = 0; k < SLABK / 16; ++k) {  // +32 B per K step of 16 inside the swizzle atom
                                    umma_f16(d, a_lo + 2 * k, b_hi + 2 * k, idesc, acc_on);
                                    acc_on = 1u;
                                }
                            }
                            if (!p.drop_blo) {
#pragma unroll
                                for (int k = 0; k < SLABK / 16; ++k) {
                                    umma_f16(d, a_hi + 2 * k, b_lo + 2 * k, idesc, acc_on);
                                    acc_on = 1u;
                                }
                            }
#pragma unroll
                            for (int k = 0; k < SLABK / 16; ++k) {
                                umma_f16(d, a_hi + 2 * k, b_hi + 2 * k, idesc, acc_on);
                                acc_on = 1u;
                            }
                        }
                        umma_commit(&empty[stage]);  // frees the smem slot once these MMAs have read it
                    }
                    umma_commit(&acc_full[a]);  // chunk complete -> epilogue may drain it
                }
            }
        }
    } else {  // ---------------- epilogue warps 2..9
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;     // MT == 1: column half; MT == 2: which of the two M tiles
        const int m = q * 32 + lane;
        const int per_img = p.BW * p.BH;
        const int img = m / per_img;
        const int r = m - img * per_img;
        const int rr = r / p.BW, cc = r - rr * p.BW;
        uint32_t c = 0;
        float amax_run = 0.0f;  // largest value this thread stored (for the activation-scale calibration)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            float acc[kCols];
#pragma unroll
            for (int j = 0; j < kCols; ++j) acc[j] = 0.0f;
            for (int ch = 0; ch < nchunks; ++ch, ++c) {
                const uint32_t a = c % kAccStages;
                mbar_wait(&acc_full[a], (c / kAccStages) & 1);
                tc_fence_after();
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + a * kAccCols + half * kCols;
                const int slabs_here = min(p.chunk_slabs, nslabs - ch * p.chunk_slabs);
                const float comp = 1.0f + p.trunc_comp * (float)(slabs_here * (SLABK / 16) * 3);
#pragma unroll
                for (int j0 = 0; j0 < kCols; j0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(t0 + j0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j0 + j] = fmaf(__uint_as_float(v[j]), comp, acc[j0 + j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a]);
            }
            if (p.gemm) {  // raw accumulators of M_x: row = my TMEM lane, kCols consecutive columns
                const int per_x = mgroups * p.ntiles;
                const int xi = tile / per_x, r2 = tile - xi * per_x;
                const int gnt = r2 / mgroups, gmt = r2 - gnt * mgroups;
                float4 *dst = reinterpret_cast<float4 *>(
                    p.out_f32 + ((int64_t)xi * p.gemm_Tp + gmt * kTileM + m) * p.Cout + gnt * BN + half * kCols);
#pragma unroll
                for (int j = 0; j < kCols / 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                continue;
            }
            const int nt = tile / mgroups, mg = tile - nt * mgroups;
            if (MT == 1) store_tile<kCols>(p, acc, nt, mg, tiles_per_group, img, rr, cc, BN, half * kCols, sbias, amax_run);
            else store_tile<kCols>(p, acc, nt, mg * MT + half, tiles_per_group, img, rr, cc, BN, 0, sbias, amax_run);
        }
        // one atomic per warp and KERNEL (a per-tile atomicMax on one address from 60 k warps-tiles cost ~5 % of block 1)
        if (!p.gemm && p.out_f32 == nullptr) publish_amax(p.amax, amax_run);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kAccStages * kAccCols);
}

// ---------------------------------------------------------------- 64 -> 64 channels (block 1, conv2)
// The generic kernel is bound by the L2 -> SM port on this layer (48 KB of operands per 384 tensor cycles).  Here
//   * all weights stay resident in shared memory: 9 taps x (B_hi | B_lo) = 144 KB, loaded once per CTA;
//   * the three taps of one kernel COLUMN kw share one A load: a (BH + 2)-row box at column offset kw - 1 holds
//     the tiles of kh = 0, 1, 2 at row offsets kh * BW pixels (BW = 16 -> 2 KB steps, swizzle-phase aligned), so
//     a tile costs 3 x 40 KB of A traffic instead of 9 x 32 KB + 9 x 16 KB;
//   * per K step two MMAs instead of three:  A_hi x [B_hi ; B_lo] (N = 128: main | correction columns) and
//     A_lo x B_hi (N = 64, into the correction columns); A_hi is read from shared memory once, not twice.
// One ring stage = one kw = one accumulation chunk (12 K steps) of the 2-deep TMEM ring.
constexpr int kC64BoxRows = 10;                                // BH + 2
constexpr int kC64ABytes = kC64BoxRows * 16 * kSlabK * 2;      // 20 KB per hi / lo box
constexpr int kC64StageBytes = 2 * kC64ABytes;
constexpr int kC64BTap = 2 * 64 * kSlabK * 2;                  // B_hi | B_lo of one tap: 16 KB
constexpr int kC64Stages = 2;
constexpr int kC64Smem = 9 * kC64BTap + kC64Stages * kC64StageBytes + 1024 + 256 + 64 * 4;

__global__ void __launch_bounds__(kConvThreads, 1) conv3x3_c64_kernel(const __grid_constant__ CUtensorMap tmAh,
                                                                      const __grid_constant__ CUtensorMap tmAl,
                                                                      const __grid_constant__ CUtensorMap tmBh,
                                                                      const __grid_constant__ CUtensorMap tmBl,
                                                                      const ConvTcParams p) {
    constexpr int BN = 64, kCols = BN / 2, kAccCols = 2 * BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *bres = smem;                                   // [9][B_hi 8 KB | B_lo 8 KB]
    uint8_t *ring = smem + 9 * kC64BTap;                    // [stages][A_hi | A_lo]
    uint64_t *full = reinterpret_cast<uint64_t *>(ring + kC64Stages * kC64StageBytes);
    uint64_t *empty = full + kC64Stages;
    uint64_t *acc_full = empty + kC64Stages;
    uint64_t *acc_empty = acc_full + kAccStages;
    uint64_t *wfull = acc_empty + kAccStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wfull + 1);
    float *sbias = reinterpret_cast<float *>(ring + kC64Stages * kC64StageBytes + 256);
    if (threadIdx.x < 64) sbias[threadIdx.x] = __ldg(p.bias + threadIdx.x);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_group = p.tilesW * p.tilesH;
    const int total_tiles = p.mtiles;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmAh); tma_prefetch_desc(&tmAl); tma_prefetch_desc(&tmBh); tma_prefetch_desc(&tmBl);
        for (int i = 0; i < kC64Stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kAccStages * kAccCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ---------------- TMA producer
            mbar_expect_tx(wfull, 9 * kC64BTap);
            for (int t = 0; t < 9; ++t) {
                tma_load_2d(&tmBh, wfull, bres + t * kC64BTap, t * 64, 0);
                tma_load_2d(&tmBl, wfull, bres + t * kC64BTap + kC64BTap / 2, t * 64, 0);
            }
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int grp = tile / tiles_per_group, rem = tile - grp * tiles_per_group;
                const int th = rem / p.tilesW, tw = rem - th * p.tilesW;
                const int h0 = th * p.BH, w0 = tw * p.BW;
                for (int kw = 0; kw < 3; ++kw, ++g) {
                    const uint32_t stage = g % kC64Stages, it = g / kC64Stages;
                    mbar_wait(&empty[stage], (it & 1) ^ 1);
                    uint8_t *sb = ring + stage * kC64StageBytes;
                    mbar_expect_tx(&full[stage], kC64StageBytes);
                    tma_load_4d(&tmAh, &full[stage], sb, 0, w0 + kw - 1, h0 - 1, grp);
                    tma_load_4d(&tmAl, &full[stage], sb + kC64ABytes, 0, w0 + kw - 1, h0 - 1, grp);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---------------- MMA issuer
            constexpr uint32_t idesc_cat = make_idesc(kTileM, 2 * BN), idesc_hi = make_idesc(kTileM, BN);
            mbar_wait(wfull, 0);
            tc_fence_after();
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int kw = 0; kw < 3; ++kw, ++g) {
                    const uint32_t a = g % kAccStages;
                    mbar_wait(&acc_empty[a], ((g / kAccStages) & 1) ^ 1);
                    const uint32_t stage = g % kC64Stages, it = g / kC64Stages;
                    mbar_wait(&full[stage], it & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + a * kAccCols;
                    const uint32_t sb = smem_u32(ring + stage * kC64StageBytes);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint64_t a_hi = make_sdesc(sb + kh * (16 * kSlabK * 2));
                        const uint64_t a_lo = make_sdesc(sb + kC64ABytes + kh * (16 * kSlabK * 2));
                        const uint64_t b_cat = make_sdesc(smem_u32(bres + (kh * 3 + kw) * kC64BTap));
#pragma unroll
                        for (int k = 0; k < kSlabK / 16; ++k) {
                            const uint32_t acc_on = (kh > 0 || k > 0) ? 1u : 0u;
                            // without the weight low parts the first MMA is N = 64 and leaves the correction columns untouched
                            umma_f16(d, a_hi + 2 * k, b_cat + 2 * k, p.drop_blo ? idesc_hi : idesc_cat, acc_on);
                            if (!p.drop_alo) umma_f16(d + BN, a_lo + 2 * k, b_cat + 2 * k, idesc_hi, p.drop_blo ? acc_on : 1u);
                        }
                    }
                    umma_commit(&empty[stage]);
                    umma_commit(&acc_full[a]);
                }
            }
        }
    } else {  // ---------------- epilogue warps 2..9
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int m = q * 32 + lane;
        const int rr = m / p.BW, cc = m - rr * p.BW;
        uint32_t c = 0;
        float amax_run = 0.0f;  // largest value this thread stored (for the activation-scale calibration)
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            float acc[kCols];
#pragma unroll
            for (int j = 0; j < kCols; ++j) acc[j] = 0.0f;
            for (int ch = 0; ch < 3; ++ch, ++c) {
                const uint32_t a = c % kAccStages;
                mbar_wait(&acc_full[a], (c / kAccStages) & 1);
                tc_fence_after();
                const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + a * kAccCols + half * kCols;
                uint32_t v[32], u[32];
                tmem_ld32(t0, v);
                tmem_ld32(t0 + BN, u);
                tmem_ld_wait();
                const float comp = 1.0f + p.trunc_comp * 12.0f;  // the main columns saw 12 accumulation steps
                const bool no_corr = p.drop_alo && p.drop_blo;  // nothing was accumulated into the correction columns
#pragma unroll
                for (int j = 0; j < kCols; ++j)
                    acc[j] = __fadd_rn(acc[j], fmaf(__uint_as_float(v[j]), comp, no_corr ? 0.0f : __uint_as_float(u[j])));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[a]);
            }
            store_tile<kCols>(p, acc, 0, tile, tiles_per_group, 0, rr, cc, BN, half * kCols, sbias, amax_run);
        }
        if (p.out_f32 == nullptr) publish_amax(p.amax, amax_run);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kAccStages * kAccCols);
}

// ---------------------------------------------------------------- Winograd F(2x2, 3x3) for the deep layers
// Y = A^T [ (G g G^T) . (B^T d B) ] A turns the 3x3 convolution of a 4x4 input tile d (stride 2) into 16 independent
// channel contractions -- 16 GEMMs  M_x[tiles, Cout] = V_x[tiles, Cin] U_x[Cout, Cin]^T  with 4/9 of the direct
// convolution's MACs.  The price is a 4x larger transformed activation (V) and an fp32 M round trip, so it pays only
// where the activations are small next to the weights: b5c1, b5c2, b6c1, b6c2 (see wino_layer()).  V and U are split into fp16
// hi/lo pairs like every other operand and the GEMMs run in conv3x3_tc_kernel's GEMM mode (same fp16x3 MMAs, same
// chunked fp32 accumulation); CPU emulation of this exact scheme: 4.5e-6 relative embedding error on the
// centred-head fixture (tests/dev/dev_emulate_winograd.py).
constexpr int kWinoMinCin = 512;   // U is prepared from here on; the dispatch rule is wino_layer()
constexpr float kWinoVScale = 0.25f;  // V is stored as B^T (64 d) B / 4 = 16 (B^T d B): head-room for the 4-term sums

// in (hi, lo) NHWC [N][H][W][C] (values * 64) -> V (hi, lo) [16][Tp][C] (values * 16); thread = (tile, 4 channels):
// 8-byte loads / stores, a warp covers 256 contiguous bytes of every pixel / transformed component
__global__ void __launch_bounds__(256) wino_in_kernel(const __half *__restrict__ in_hi, const __half *__restrict__ in_lo,
                                                      __half *__restrict__ v_hi, __half *__restrict__ v_lo, int N, int H,
                                                      int W, int C, int th, int tw, int Tp) {
    const int C4 = C >> 2;
    const int64_t total = (int64_t)N * th * tw * C4;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % C4);
        const int tile = (int)(idx / C4);
        const int tx = tile % tw, ty = (tile / tw) % th, n = tile / (tw * th);
        uint2 rh[16], rl[16];  // raw hi / lo of the 4 x 4 patch (4 channels each): all loads in flight before any use
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int hh = 2 * ty - 1 + i, ww = 2 * tx - 1 + j;
                uint2 a = make_uint2(0u, 0u), b = make_uint2(0u, 0u);
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
                    const size_t o = (((size_t)n * H + hh) * W + ww) * C + 4 * c4;
                    a = __ldg(reinterpret_cast<const uint2 *>(in_hi + o));
                    b = __ldg(reinterpret_cast<const uint2 *>(in_lo + o));
                }
                rh[i * 4 + j] = a; rl[i * 4 + j] = b;
            }
        uint2 oh[16], ol[16];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            float d[4][4], t[4][4];
#pragma unroll
            for (int x = 0; x < 16; ++x) {
                const uint32_t wh = (ch < 2) ? rh[x].x : rh[x].y, wl = (ch < 2) ? rl[x].x : rl[x].y;
                const __half2 h = *reinterpret_cast<const __half2 *>(&wh), l = *reinterpret_cast<const __half2 *>(&wl);
                d[x >> 2][x & 3] = (ch & 1) ? (__high2float(h) + __high2float(l)) : (__low2float(h) + __low2float(l));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // B^T d
                t[0][j] = d[0][j] - d[2][j];
                t[1][j] = d[1][j] + d[2][j];
                t[2][j] = d[2][j] - d[1][j];
                t[3][j] = d[1][j] - d[3][j];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // (.) B
                d[i][0] = (t[i][0] - t[i][2]) * kWinoVScale;
                d[i][1] = (t[i][1] + t[i][2]) * kWinoVScale;
                d[i][2] = (t[i][2] - t[i][1]) * kWinoVScale;
                d[i][3] = (t[i][1] - t[i][3]) * kWinoVScale;
            }
#pragma unroll
            for (int x = 0; x < 16; ++x) {
                const float v = d[x >> 2][x & 3];
                const __half hv = __float2half_rn(v);
                const __half lv = __float2half_rn(v - __half2float(hv));
                const uint32_t hb = (uint32_t)__half_as_ushort(hv) << ((ch & 1) * 16);
                const uint32_t lb = (uint32_t)__half_as_ushort(lv) << ((ch & 1) * 16);
                if (ch == 0) { oh[x].x = hb; ol[x].x = lb; }
                else if (ch == 1) { oh[x].x |= hb; ol[x].x |= lb; }
                else if (ch == 2) { oh[x].y = hb; ol[x].y = lb; }
                else { oh[x].y |= hb; ol[x].y |= lb; }
            }
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) {
            const size_t o = ((size_t)x * Tp + tile) * C + 4 * c4;
            *reinterpret_cast<uint2 *>(v_hi + o) = oh[x];
            *reinterpret_cast<uint2 *>(v_lo + o) = ol[x];
        }
    }
}

// M [16][Tp][Cout] fp32 (raw accumulators) -> A^T M A * unscale + bias -> ReLU -> (2x2 average pool: exactly one
// output tile) -> fp16 hi/lo NHWC (values * 64) or fp32 NHWC; thread = (tile, channel pair)
__global__ void __launch_bounds__(256) wino_out_kernel(const float *__restrict__ M, const float *__restrict__ bias,
                                                       float unscale, __half *__restrict__ out_hi,
                                                       __half *__restrict__ out_lo, float *__restrict__ out_f32, int N,
                                                       int H, int W, int Cout, int th, int tw, int Tp, int pool,
                                                       float out_scale, int *__restrict__ overflow,
                                                       unsigned *__restrict__ amax) {
    float mx = 0.0f;
    const int C2 = Cout >> 1;
    const int64_t total = (int64_t)N * th * tw * C2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c2 = (int)(idx % C2);
        const int tile = (int)(idx / C2);
        const int tx = tile % tw, ty = (tile / tw) % th, n = tile / (tw * th);
        float2 m[4][4];
#pragma unroll
        for (int x = 0; x < 16; ++x)
            m[x >> 2][x & 3] = __ldg(reinterpret_cast<const float2 *>(M + ((size_t)x * Tp + tile) * Cout + 2 * c2));
        const float2 bv = __ldg(reinterpret_cast<const float2 *>(bias + 2 * c2));
        float y[2][2][2];  // [row][col][channel]
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float t[2][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {  // A^T m
                const float m0 = ch ? m[0][j].y : m[0][j].x, m1 = ch ? m[1][j].y : m[1][j].x;
                const float m2 = ch ? m[2][j].y : m[2][j].x, m3 = ch ? m[3][j].y : m[3][j].x;
                t[0][j] = (m0 + m1) + m2;
                t[1][j] = (m1 - m2) - m3;
            }
            const float b = ch ? bv.y : bv.x;
#pragma unroll
            for (int i = 0; i < 2; ++i) {  // (.) A, then unscale + bias + ReLU
                y[i][0][ch] = fmaxf(fmaf((t[i][0] + t[i][1]) + t[i][2], unscale, b), 0.0f);
                y[i][1][ch] = fmaxf(fmaf((t[i][1] - t[i][2]) - t[i][3], unscale, b), 0.0f);
            }
        }
        if (pool) {
            const int Ho = H >> 1, Wo = W >> 1;
            if (ty >= Ho || tx >= Wo) continue;
            float o[2];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
                o[ch] = ((y[0][0][ch] + y[0][1][ch]) + (y[1][0][ch] + y[1][1][ch])) * (0.25f * out_scale);
                mx = fmaxf(mx, o[ch]);
                if (o[ch] > kHalfMax) { o[ch] = kHalfMax; if (overflow != nullptr) atomicOr(overflow, 1); }
            }
            const size_t oo = (((size_t)n * Ho + ty) * Wo + tx) * Cout + 2 * c2;
            const __half h0 = __float2half_rn(o[0]), h1 = __float2half_rn(o[1]);
            *reinterpret_cast<__half2 *>(out_hi + oo) = __halves2half2(h0, h1);
            *reinterpret_cast<__half2 *>(out_lo + oo) = __halves2half2(__float2half_rn(o[0] - __half2float(h0)),
                                                                      __float2half_rn(o[1] - __half2float(h1)));
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int hh = 2 * ty + i, ww = 2 * tx + j;
                    if (hh >= H || ww >= W) continue;
                    const size_t oo = (((size_t)n * H + hh) * W + ww) * Cout + 2 * c2;
                    if (out_f32 != nullptr) {
                        *reinterpret_cast<float2 *>(out_f32 + oo) = make_float2(y[i][j][0], y[i][j][1]);
                    } else {
                        float a = y[i][j][0] * out_scale, b = y[i][j][1] * out_scale;
                        mx = fmaxf(mx, fmaxf(a, b));
                        if (a > kHalfMax || b > kHalfMax) {
                            a = fminf(a, kHalfMax); b = fminf(b, kHalfMax);
                            if (overflow != nullptr) atomicOr(overflow, 1);
                        }
                        const __half h0 = __float2half_rn(a), h1 = __float2half_rn(b);
                        *reinterpret_cast<__half2 *>(out_hi + oo) = __halves2half2(h0, h1);
                        *reinterpret_cast<__half2 *>(out_lo + oo) = __halves2half2(__float2half_rn(a - __half2float(h0)),
                                                                                  __float2half_rn(b - __half2float(h1)));
                    }
                }
        }
    }
    publish_amax(amax, mx);
}

// ------------------------------------------------------------ SIMT helpers of the fp16x3 path
// first conv (Cin == 1, K = 9: memory-bound, exact fp32): feat [N][H][W] -> hi/lo NHWC [N][H][W][64].
// Thread = (4 horizontally adjacent pixels, group of 8 output channels).  The channel group is fixed per thread
// over the grid-stride loop (stride % 8 == 0), so its 72 weights + 8 biases live in registers; per iteration a
// thread reads a 3 x 6 input patch (cached loads, shared by its 4 pixels), does 288 FMAs and writes 4 x 2 x 16 B.
constexpr int kC1Px = 4;
__global__ void __launch_bounds__(256) tc_conv_first_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                            const float *__restrict__ bias, __half *__restrict__ yh,
                                                            __half *__restrict__ yl, int N, int H, int W,
                                                            float out_scale, int *__restrict__ overflow,
                                                            unsigned *__restrict__ amax) {
    float mx = 0.0f;
    const int g = threadIdx.x & 7;
    float wr[9][8], br[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int c = 0; c < 8; ++c) wr[t][c] = __ldg(w + t * 64 + g * 8 + c);
#pragma unroll
    for (int c = 0; c < 8; ++c) br[c] = __ldg(bias + g * 8 + c);
    // 32-bit index arithmetic (N*H*W < 2^31 is checked by the launcher): 64-bit div/mod would dominate the loop
    const uint32_t Wq = (uint32_t)(W + kC1Px - 1) / kC1Px;          // pixel quads per row
    const uint32_t nquads = (uint32_t)N * (uint32_t)H * Wq;
    const uint32_t qstep = (gridDim.x * blockDim.x) >> 3;
    for (uint32_t quad = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; quad < nquads; quad += qstep) {
        const uint32_t row = quad / Wq;                               // n * H + h
        const int w0 = (int)(quad - row * Wq) * kC1Px;
        const uint32_t n = row / (uint32_t)H;
        const int hq = (int)(row - n * (uint32_t)H);
        const float *xn = x + (size_t)n * H * W;
        float v[3][kC1Px + 2];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hh = hq + kh - 1;
            const bool hok = hh >= 0 && hh < H;
#pragma unroll
            for (int i = 0; i < kC1Px + 2; ++i) {
                const int ww = w0 + i - 1;
                v[kh][i] = (hok && ww >= 0 && ww < W) ? __ldg(xn + hh * W + ww) : 0.0f;
            }
        }
#pragma unroll
        for (int px = 0; px < kC1Px; ++px) {
            if (w0 + px >= W) break;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {  // same accumulation order as the single-pixel form (t = 0..8)
                        a0 = fmaf(v[kh][px + kw], wr[kh * 3 + kw][2 * c2], a0);
                        a1 = fmaf(v[kh][px + kw], wr[kh * 3 + kw][2 * c2 + 1], a1);
                    }
                a0 = fmaxf(a0 + br[2 * c2], 0.f) * out_scale;
                a1 = fmaxf(a1 + br[2 * c2 + 1], 0.f) * out_scale;
                mx = fmaxf(mx, fmaxf(a0, a1));
                if (a0 > kHalfMax || a1 > kHalfMax) {
                    a0 = fminf(a0, kHalfMax); a1 = fminf(a1, kHalfMax);
                    if (overflow != nullptr) atomicOr(overflow, 1);
                }
                const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
                const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
                hi[c2] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                lo[c2] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
            }
            const size_t o = ((size_t)row * W + w0 + px) * 64 + g * 8;
            *reinterpret_cast<uint4 *>(yh + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(yl + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
    publish_amax(amax, mx);
}

// ---------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int tc_fail(const char *what, const char *detail) {
    g_tc_err = std::string(what) + ": " + detail;
    return -1;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (function, device): done once each, not per launch
// (a generation used to make ~25 of these driver calls).
}  // namespace
cudaError_t ensure_dyn_smem(const void *func, int bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> done;  // (kernel, device) -> bytes granted so far
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find({func, dev});
    if (it != done.end() && it->second >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done[{func, dev}] = bytes;
    return e;
}
namespace {

const CUtensorMap *cached_map(TcWorkspace &ws, const TcMapKey &key, bool *hit) {
    for (TcMapEntry *e : ws.maps)
        if (e->key == key) { *hit = true; return reinterpret_cast<const CUtensorMap *>(e->map); }
    if (ws.maps.size() >= 1024) {  // unbounded shape churn: start over
        for (TcMapEntry *e : ws.maps) delete e;
        ws.maps.clear();
    }
    TcMapEntry *e = new TcMapEntry();
    e->key = key;
    ws.maps.push_back(e);
    *hit = false;
    return reinterpret_cast<const CUtensorMap *>(e->map);
}
void drop_last_map(TcWorkspace &ws) {
    delete ws.maps.back();
    ws.maps.pop_back();
}

// activations: fp16 NHWC viewed as 4-D (C, W, H, N), box (64, BW, BH, IPT), 128B swizzle, zero OOB fill
const CUtensorMap *make_act_map(TcWorkspace &ws, const void *base, int N, int H, int W, int C, int BW, int BH, int IPT,
                                int slabk = kSlabK) {
    const TcMapKey key{base, {N, H, W, C}, {slabk, BW, BH, IPT}};
    bool hit = false;
    const CUtensorMap *cm = cached_map(ws, key, &hit);
    if (hit) return cm;
    CUtensorMap *m = const_cast<CUtensorMap *>(cm);
    EncodeTiledFn enc = get_encode();
    if (!enc) { drop_last_map(ws); tc_fail("cuTensorMapEncodeTiled", "driver entry point not found"); return nullptr; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)slabk, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)IPT};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, slabk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof(buf), "CUresult %d for activation map N=%d H=%d W=%d C=%d box=(64,%d,%d,%d)", (int)r, N, H, W, C, BW, BH, IPT);
        drop_last_map(ws);
        tc_fail("cuTensorMapEncodeTiled", buf);
        return nullptr;
    }
    return cm;
}

// weights: fp16 [Cout][K] K-major, box (64, BN)
const CUtensorMap *make_w_map(TcWorkspace &ws, const void *base, int Cout, int K, int BN, int slabk = kSlabK) {
    const TcMapKey key{base, {Cout, K, 0, 0}, {slabk, BN, 0, 0}};
    bool hit = false;
    const CUtensorMap *cm = cached_map(ws, key, &hit);
    if (hit) return cm;
    CUtensorMap *m = const_cast<CUtensorMap *>(cm);
    EncodeTiledFn enc = get_encode();
    if (!enc) { drop_last_map(ws); tc_fail("cuTensorMapEncodeTiled", "driver entry point not found"); return nullptr; }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)slabk, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, slabk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[128];
        snprintf(buf, sizeof(buf), "CUresult %d for weight map Cout=%d K=%d BN=%d", (int)r, Cout, K, BN);
        drop_last_map(ws);
        tc_fail("cuTensorMapEncodeTiled", buf);
        return nullptr;
    }
    return cm;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int BN, int STAGES, int SLABK, int MT = 1>
int launch_conv_tc_t(cudaStream_t st, const CUtensorMap &ah, const CUtensorMap &al, const CUtensorMap &bh,
                     const CUtensorMap &bl, const ConvTcParams &p) {
    constexpr int smem = STAGES * (MT * 2 * kTileM * SLABK * 2 + 2 * BN * SLABK * 2) + 1024 + 256 + 2048 * 4;  // ring, align, barriers, bias
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void *>(&conv3x3_tc_kernel<BN, STAGES, SLABK, MT>), smem);
        if (e != cudaSuccess) return tc_fail("cudaFuncSetAttribute", cudaGetErrorString(e));
    }
    const int total = ((p.mtiles + MT - 1) / MT) * p.ntiles * (p.gemm ? 16 : 1);
    const int grid = total < num_sms() ? total : num_sms();  // persistent: one CTA per SM
    conv3x3_tc_kernel<BN, STAGES, SLABK, MT><<<grid, kConvThreads, smem, st>>>(ah, al, bh, bl, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return tc_fail("conv3x3_tc_kernel launch", cudaGetErrorString(e));
    return 0;
}

int ws_ensure(TcWorkspace &ws, int i, size_t bytes) {
    if (bytes <= ws.cap[i]) return 0;
    if (ws.buf[i]) cudaFree(ws.buf[i]);
    ws.buf[i] = nullptr;
    ws.cap[i] = 0;
    cudaError_t e = cudaMalloc(&ws.buf[i], bytes);
    if (e != cudaSuccess) return tc_fail("cudaMalloc(workspace)", cudaGetErrorString(e));
    ws.cap[i] = bytes;
    return 0;
}

// K-slabs (of 64) accumulated in TMEM between fp32 drains; STITO_TC_CHUNK overrides (developer knob)
int chunk_slabs(int BN = 256) {
    // K = 128 per chunk on the N = 256 layers: the 1e-4 gate then holds WITHOUT the accumulate compensation (see chunk_comp()).
    // The N = 128 layers (block 2) drain twice as often per MMA cycle, K = 128 chunks cost them +19 % / +11 %, and their chunk
    // length hardly matters for the error (measured with the compensation off, short fixture of tests/dev/dev_margins2.py,
    // Xavier / heavy-tailed weights: K = 128: 7.1e-5 / 7.6e-5, K = 192: 7.4e-5 / 9.1e-5, K = 256: 8.6e-5 / 8.9e-5; with the
    // compensation 2-3e-5 throughout): they keep K = 256.  STITO_TC_CHUNK overrides both, STITO_TC_CHUNK128 the N <= 128 value.
    static int v = 0, v128 = 0;
    if (v == 0) {
        v = 2;
        v128 = 4;
        if (const char *e = getenv("STITO_TC_CHUNK")) { const int t = atoi(e); if (t > 0) v = v128 = t; }
        if (const char *e = getenv("STITO_TC_CHUNK128")) { const int t = atoi(e); if (t > 0) v128 = t; }
    }
    return BN <= 128 ? v128 : v;
}

int slab_k(int cin) {
    // measured on B200 (P = 64, 10 s): K = 32 wins on the deep layers (Cin >= 1024: -9 % b5c2, -8 % b6c1, -10 % b6c2),
    // K = 64 on the wide, shallow ones (fewer TMA / barrier round trips per byte)
    static int forced = -1;
    if (forced < 0) {
        forced = 0;
        if (const char *e = getenv("STITO_TC_SLABK")) { const int t = atoi(e); if (t == 32 || t == 64) forced = t; }
    }
    if (forced) return forced;
    return cin >= 1024 ? 32 : 64;
}

bool use_mt2() {  // STITO_TC_MT2=0: one M tile per work item on the Cout = 128 layers (developer knob)
    static int v = -1;
    if (v < 0) { const char *e = getenv("STITO_TC_MT2"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v != 0;
}

// Winograd pays when the weights are large next to the activations.  Saved MACs and transform traffic both scale with
// the pixel count, so the break-even depends on Cin*Cout/(Cin+Cout) only; measured on B200 (P = 64, 10 s):
// 512->512 (256): +8 %, 512->1024 (341): -11 %, 1024->1024 (512): -24 %, 1024->2048 (683): -37 %, 2048->2048 (1024): -47 %.
// Threshold 340 selects the last four; STITO_TC_WINO_MIN overrides it (developer knob).
bool wino_layer(const ConvLayer &l) {
    static int thr = -1;
    if (thr < 0) {
        thr = 340;
        if (const char *e = getenv("STITO_TC_WINO_MIN")) { const int t = atoi(e); if (t > 0) thr = t; }
    }
    return l.u_hi != nullptr && (int64_t)l.cin * l.cout >= (int64_t)thr * (l.cin + l.cout);
}

bool use_wino() {  // STITO_TC_WINOGRAD=0: direct implicit-GEMM convolution on every layer (developer knob)
    static int v = -1;
    if (v < 0) { const char *e = getenv("STITO_TC_WINOGRAD"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v != 0;
}

int drop_mask(const char *name) {  // bit l set: conv layer l (1..11) runs without that correction pass
    const char *e = getenv(name);
    return e ? (int)strtol(e, nullptr, 0) : 0;
}

bool use_c64() {  // STITO_TC_C64=0 falls back to the generic kernel for block 1 (developer knob)
    static int v = -1;
    if (v < 0) { const char *e = getenv("STITO_TC_C64"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v != 0;
}

// The tensor core's fp32 accumulate rounds toward zero: every K = 16 MMA step shrinks the running sum by a tiny,
// systematic relative amount.  Measured on B200 by scanning this factor against the fp32 oracle (centred-head golden
// fixture, K = 256 chunks): embedding error 5.9e-5 at 0, 1.2e-5 at 0.25, 6.5e-5 at 0.5, 1.9e-4 at 1.0 (units of 2^-24
// per step) -- a clean V with its minimum at ~0.24, where the error reaches the level of the fp32 CUDA-core mode
// (9e-6).  Every drained chunk is therefore multiplied by 1 + 0.25 * 2^-24 * steps.  STITO_TC_COMP overrides.
float chunk_comp() {
    static float v = -1.0f;
    if (v < 0.0f) {
        float a = 0.25f;
        if (const char *e = getenv("STITO_TC_COMP")) a = (float)atof(e);
        v = a * 5.9604644775390625e-08f;
    }
    return v;
}

inline int blocks_for(int64_t total, int threads) {
    int64_t b = (total + threads - 1) / threads;
    const int64_t cap = 148 * 32;
    return (int)(b < cap ? b : cap);
}

}  // namespace

bool tc_available() { return true; }
const char *tc_last_error() { return g_tc_err.c_str(); }

int tc_prepare_layer(const float *wf, int cin, int cout, ConvLayer *cl, std::vector<void *> *owned) {
    // Pre-scale by a power of two so that the fp16 lo parts stay clear of the subnormal range: max |w| * 2^shift lands in
    // (2^12, 2^13], i.e. every weight above ~2e-5 of the largest keeps its full hi + lo = 22 bits.  (Round 1 scaled the
    // maximum to 1: fine for Xavier-uniform weights, but with heavy-tailed weights / BatchNorm scales spread over two
    // decades the typical weight sat at ~0.02, its lo part in the subnormal range -> 18-bit weights and a 2e-4 embedding
    // error on the second fixture of tests/dev/dev_margins2.py.)
    double mx = 0.0;
    const size_t n = (size_t)9 * cin * cout;
    for (size_t i = 0; i < n; ++i) mx = std::fmax(mx, std::fabs((double)wf[i]));
    int shift = 0;
    if (mx > 0) shift = kWeightTop - (int)std::ceil(std::log2(mx));
    if (shift > 60) shift = 60;
    if (shift < -40) shift = -40;
    const float scale = std::ldexp(1.0f, shift);
    const int K = 9 * cin;
    std::vector<__half> hi((size_t)cout * K), lo((size_t)cout * K);
    for (int t = 0; t < 9; ++t)
        for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co) {
                const float v = wf[((size_t)t * cin + ci) * cout + co] * scale;
                const __half h = __float2half_rn(v);
                const size_t o = (size_t)co * K + (size_t)t * cin + ci;
                hi[o] = h;
                lo[o] = __float2half_rn(v - __half2float(h));
            }
    void *dh = nullptr, *dl = nullptr;
    cudaError_t e = cudaMalloc(&dh, hi.size() * sizeof(__half));
    if (e != cudaSuccess) return tc_fail("cudaMalloc(weights)", cudaGetErrorString(e));
    owned->push_back(dh);
    e = cudaMalloc(&dl, lo.size() * sizeof(__half));
    if (e != cudaSuccess) return tc_fail("cudaMalloc(weights)", cudaGetErrorString(e));
    owned->push_back(dl);
    e = cudaMemcpy(dh, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dl, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return tc_fail("cudaMemcpy(weights)", cudaGetErrorString(e));
    cl->w_hi = dh;
    cl->w_lo = dl;
    cl->w_unscale = std::ldexp(1.0f, -shift);
    if (cin < kWinoMinCin) return 0;

    // Winograd-domain weights U = G g G^T, [16][cout][cin], computed in double from the BN-folded fp32 weights
    static const double G[4][3] = {{1, 0, 0}, {0.5, 0.5, 0.5}, {0.5, -0.5, 0.5}, {0, 0, 1}};
    const size_t nu = (size_t)16 * cout * cin;
    std::vector<float> U(nu);
    double umx = 0.0;
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci) {
            double g[3][3], t[4][3];
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) g[a][b] = wf[((size_t)(a * 3 + b) * cin + ci) * cout + co];
            for (int x = 0; x < 4; ++x)
                for (int b = 0; b < 3; ++b) t[x][b] = G[x][0] * g[0][b] + G[x][1] * g[1][b] + G[x][2] * g[2][b];
            for (int x = 0; x < 4; ++x)
                for (int y = 0; y < 4; ++y) {
                    const double u = t[x][0] * G[y][0] + t[x][1] * G[y][1] + t[x][2] * G[y][2];
                    U[((size_t)(x * 4 + y) * cout + co) * cin + ci] = (float)u;
                    umx = std::fmax(umx, std::fabs(u));
                }
        }
    int ushift = 0;
    if (umx > 0) ushift = kWeightTop - (int)std::ceil(std::log2(umx));
    if (ushift > 60) ushift = 60;
    if (ushift < -40) ushift = -40;
    const float uscale = std::ldexp(1.0f, ushift);
    std::vector<__half> uh(nu), ul(nu);
    for (size_t i = 0; i < nu; ++i) {
        const float v = U[i] * uscale;
        const __half h = __float2half_rn(v);
        uh[i] = h;
        ul[i] = __float2half_rn(v - __half2float(h));
    }
    void *duh = nullptr, *dul = nullptr;
    e = cudaMalloc(&duh, nu * sizeof(__half));
    if (e != cudaSuccess) return tc_fail("cudaMalloc(winograd weights)", cudaGetErrorString(e));
    owned->push_back(duh);
    e = cudaMalloc(&dul, nu * sizeof(__half));
    if (e != cudaSuccess) return tc_fail("cudaMalloc(winograd weights)", cudaGetErrorString(e));
    owned->push_back(dul);
    e = cudaMemcpy(duh, uh.data(), nu * sizeof(__half), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dul, ul.data(), nu * sizeof(__half), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return tc_fail("cudaMemcpy(winograd weights)", cudaGetErrorString(e));
    cl->u_hi = duh;
    cl->u_lo = dul;
    cl->u_unscale = std::ldexp(1.0f, -ushift);
    return 0;
}

void tc_workspace_release(TcWorkspace *ws) {
    for (TcMapEntry *e : ws->maps) delete e;
    ws->maps.clear();
    for (int i = 0; i < TcWorkspace::kBufs; ++i) {
        if (ws->buf[i]) cudaFree(ws->buf[i]);
        ws->buf[i] = nullptr;
        ws->cap[i] = 0;
    }
}

// One conv layer on tensor cores: in (hi, lo) NHWC [N][H][W][Cin] -> out fp16 pair or fp32.
static int conv_tc(cudaStream_t st, const ConvLayer &l, TcWorkspace &ws, int li, const __half *in_hi, const __half *in_lo,
                   __half *out_hi, __half *out_lo, float *out_f32, int N, int H, int W, bool pool, int *launches) {
    const float in_scale = std::ldexp(1.0f, ws.act_shift[li - 1]), out_scale = std::ldexp(1.0f, ws.act_shift[li]);
    ConvTcParams p{};
    p.bias = l.bias; p.unscale = l.w_unscale / in_scale;
    p.out_scale = out_f32 ? 1.0f : (pool ? 0.25f * out_scale : out_scale);  // exact powers of two
    p.amax = ws.amax ? ws.amax + li : nullptr;
    p.pool = pool ? 1 : 0;
    p.trunc_comp = chunk_comp();
    p.overflow = ws.overflow_flag;
    p.drop_alo = (drop_mask("STITO_TC_DROP_ALO") >> li) & 1;
    p.drop_blo = (drop_mask("STITO_TC_DROP_BLO") >> li) & 1;
    p.out_hi = out_hi; p.out_lo = out_lo; p.out_f32 = out_f32;
    p.N = N; p.H = H; p.W = W; p.Cin = l.cin; p.Cout = l.cout;
    p.BW = W < 64 ? W : 64;
    if (pool && p.BW > 16) p.BW = 16;  // keeps every 2x2 window inside one warp of the epilogue
    if (pool && (p.BW & (p.BW - 1)) != 0) return tc_fail("conv_tc", "pooled layers need a power-of-two tile width");
    int bh = kTileM / p.BW;
    if (bh > H) bh = H;
    p.BH = bh;
    int ipt = kTileM / (p.BW * p.BH);
    if (ipt > N) ipt = N;
    if (ipt < 1) ipt = 1;
    p.IPT = ipt;
    p.tilesW = (W + p.BW - 1) / p.BW;
    p.tilesH = (H + p.BH - 1) / p.BH;
    const int groups = (N + p.IPT - 1) / p.IPT;
    const int BN = l.cout >= 256 ? 256 : l.cout;
    p.mtiles = groups * p.tilesW * p.tilesH;
    p.ntiles = l.cout / BN;
    p.chunk_slabs = chunk_slabs();
    if (l.cin % kSlabK != 0 || l.cout % BN != 0 || l.cout > 2048 || (BN != 64 && BN != 128 && BN != 256))
        return tc_fail("conv_tc", "unsupported channel counts");
    const CUtensorMap *ah, *al, *bh_, *bl;
    if (l.cin == 64 && l.cout == 64 && p.BW == 16 && p.BH == 8 && p.IPT == 1 && use_c64()) {
        // resident weights + kh-shared A boxes (conv3x3_c64_kernel)
        if (!(ah = make_act_map(ws, in_hi, N, H, W, 64, 16, kC64BoxRows, 1))) return -1;
        if (!(al = make_act_map(ws, in_lo, N, H, W, 64, 16, kC64BoxRows, 1))) return -1;
        if (!(bh_ = make_w_map(ws, l.w_hi, 64, 9 * 64, 64))) return -1;
        if (!(bl = make_w_map(ws, l.w_lo, 64, 9 * 64, 64))) return -1;
        {
            cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void *>(&conv3x3_c64_kernel), kC64Smem);
            if (e != cudaSuccess) return tc_fail("cudaFuncSetAttribute", cudaGetErrorString(e));
        }
        const int grid = p.mtiles < num_sms() ? p.mtiles : num_sms();
        conv3x3_c64_kernel<<<grid, kConvThreads, kC64Smem, st>>>(*ah, *al, *bh_, *bl, p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return tc_fail("conv3x3_c64_kernel launch", cudaGetErrorString(e));
        *launches += 1;
        return 0;
    }
    // K-slab per ring stage: 32 (64-byte swizzle, twice as many stages in the same shared memory: the producer runs
    // further ahead of the MMA issuer) or 64 (128-byte swizzle), chosen per layer; STITO_TC_SLABK overrides
    const int slabk = slab_k(l.cin);
    p.chunk_slabs = chunk_slabs(BN) * (64 / slabk);
    if (!(ah = make_act_map(ws, in_hi, N, H, W, l.cin, p.BW, p.BH, p.IPT, slabk))) return -1;
    if (!(al = make_act_map(ws, in_lo, N, H, W, l.cin, p.BW, p.BH, p.IPT, slabk))) return -1;
    if (!(bh_ = make_w_map(ws, l.w_hi, l.cout, 9 * l.cin, BN, slabk))) return -1;
    if (!(bl = make_w_map(ws, l.w_lo, l.cout, 9 * l.cin, BN, slabk))) return -1;
    int rc;
    if (slabk == 32) {
        if (BN == 64) rc = launch_conv_tc_t<64, 8, 32>(st, *ah, *al, *bh_, *bl, p);
        else if (BN == 128) rc = launch_conv_tc_t<128, 6, 32>(st, *ah, *al, *bh_, *bl, p);
        else rc = launch_conv_tc_t<256, 4, 32>(st, *ah, *al, *bh_, *bl, p);
    } else {
        if (BN == 64) rc = launch_conv_tc_t<64, 4, 64>(st, *ah, *al, *bh_, *bl, p);
        else if (BN == 128 && l.cin >= 128 && use_mt2()) rc = launch_conv_tc_t<128, 2, 64, 2>(st, *ah, *al, *bh_, *bl, p);  // 2 M tiles share B (-4.5 % on b2c2; b2c1, K = 576, prefers the deeper 3-stage ring)
        else if (BN == 128) rc = launch_conv_tc_t<128, 3, 64>(st, *ah, *al, *bh_, *bl, p);
        else rc = launch_conv_tc_t<256, 2, 64>(st, *ah, *al, *bh_, *bl, p);
    }
    if (rc == 0) *launches += 1;
    return rc;
}

// One deep conv layer through Winograd F(2x2,3x3): input transform -> 16 GEMMs (conv3x3_tc_kernel, GEMM mode) ->
// output transform (+ bias, ReLU, optional 2x2 pool, hi/lo split or fp32).
static int conv_wino(cudaStream_t st, const ConvLayer &l, TcWorkspace &ws, int li, const __half *in_hi, const __half *in_lo,
                     __half *out_hi, __half *out_lo, float *out_f32, int N, int H, int W, bool pool, int *launches) {
    const float in_scale = std::ldexp(1.0f, ws.act_shift[li - 1]), out_scale = std::ldexp(1.0f, ws.act_shift[li]);
    const int th = (H + 1) / 2, tw = (W + 1) / 2;
    const int T = N * th * tw;
    const int Tp = (T + kTileM - 1) / kTileM * kTileM;
    if (ws_ensure(ws, 4, (size_t)16 * Tp * l.cin * sizeof(__half))) return -1;
    if (ws_ensure(ws, 5, (size_t)16 * Tp * l.cin * sizeof(__half))) return -1;
    if (ws_ensure(ws, 6, (size_t)16 * Tp * l.cout * sizeof(float))) return -1;
    __half *v_hi = (__half *)ws.buf[4], *v_lo = (__half *)ws.buf[5];
    float *M = (float *)ws.buf[6];
    wino_in_kernel<<<blocks_for((int64_t)T * (l.cin / 4), 256), 256, 0, st>>>(in_hi, in_lo, v_hi, v_lo, N, H, W, l.cin, th, tw, Tp);
    ConvTcParams p{};
    p.N = N; p.H = H; p.W = W; p.Cin = l.cin; p.Cout = l.cout;
    p.BW = 16; p.BH = 8; p.IPT = 1; p.tilesW = 1; p.tilesH = 1;  // unused in GEMM mode
    p.mtiles = Tp / kTileM;
    p.ntiles = l.cout / 256;
    p.gemm = 1; p.gemm_Tp = Tp;
    p.trunc_comp = chunk_comp();
    p.out_f32 = M;
    p.unscale = 1.0f; p.out_scale = 1.0f;
    const int slabk = slab_k(l.cin);
    p.chunk_slabs = chunk_slabs() * (64 / slabk);
    const CUtensorMap *ah, *al, *bh_, *bl;
    if (!(ah = make_w_map(ws, v_hi, 16 * Tp, l.cin, kTileM, slabk))) return -1;   // V as [16 * Tp][Cin], box (K slab, 128 rows)
    if (!(al = make_w_map(ws, v_lo, 16 * Tp, l.cin, kTileM, slabk))) return -1;
    if (!(bh_ = make_w_map(ws, l.u_hi, 16 * l.cout, l.cin, 256, slabk))) return -1;
    if (!(bl = make_w_map(ws, l.u_lo, 16 * l.cout, l.cin, 256, slabk))) return -1;
    int rc = slabk == 32 ? launch_conv_tc_t<256, 4, 32>(st, *ah, *al, *bh_, *bl, p) : launch_conv_tc_t<256, 2, 64>(st, *ah, *al, *bh_, *bl, p);
    if (rc) return rc;
    const float unscale = l.u_unscale / (in_scale * kWinoVScale);
    wino_out_kernel<<<blocks_for((int64_t)T * (l.cout / 2), 256), 256, 0, st>>>(M, l.bias, unscale, out_hi, out_lo, out_f32, N, H, W,
                                                                                  l.cout, th, tw, Tp, pool ? 1 : 0, out_scale, ws.overflow_flag,
                                                                                  ws.amax ? ws.amax + li : nullptr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return tc_fail("winograd transform launch", cudaGetErrorString(e));
    *launches += 3;
    return 0;
}

int tc_encoder_forward(cudaStream_t st, const EncoderDev &enc, TcWorkspace &ws, const float *feat, int N, int T,
                       int mel, float *pooled, int *launches, cudaEvent_t *ev) {
    // workspace: [0],[1] = hi/lo block inputs (pooled), [2] = hi+lo conv1 outputs, [3] = fp32 output of block 6
    const size_t px0 = (size_t)N * T * mel;
    if (ws_ensure(ws, 0, px0 * 16 * sizeof(__half))) return -1;  // block input (pooled) hi: largest is block 2, px0/4 x 64
    if (ws_ensure(ws, 1, px0 * 16 * sizeof(__half))) return -1;  // lo
    if (ws_ensure(ws, 2, 2 * px0 * 64 * sizeof(__half))) return -1;  // conv1 output hi | lo
    if (ws_ensure(ws, 3, (size_t)N * (T >> 5) * (mel >> 5) * 2048 * sizeof(float))) return -1;  // block-6 output fp32
    __half *in_hi = (__half *)ws.buf[0], *in_lo = (__half *)ws.buf[1];
    float *c2 = (float *)ws.buf[3];
    int H = T, W = mel;
    for (int b = 0; b < 6; ++b) {
        const ConvLayer &l1 = enc.conv[2 * b], &l2 = enc.conv[2 * b + 1];
        const size_t px = (size_t)N * H * W;
        __half *m_hi = (__half *)ws.buf[2], *m_lo = m_hi + px * l1.cout;
        if (ev) cudaEventRecord(ev[2 * b], st);
        if (b == 0) {
            if ((uint64_t)px >= (1ull << 31)) return tc_fail("tc_encoder_forward", "micro-batch too large for 32-bit pixel indices");
            tc_conv_first_kernel<<<blocks_for((int64_t)px * 8 / kC1Px, 256), 256, 0, st>>>(feat, l1.w, l1.bias, m_hi, m_lo, N, H, W, std::ldexp(1.0f, ws.act_shift[0]),
                                                                                           ws.overflow_flag, ws.amax);
            *launches += 1;
        } else {
            if (wino_layer(l1) && use_wino()) {
                if (conv_wino(st, l1, ws, 2 * b, in_hi, in_lo, m_hi, m_lo, nullptr, N, H, W, false, launches)) return -1;
            } else if (conv_tc(st, l1, ws, 2 * b, in_hi, in_lo, m_hi, m_lo, nullptr, N, H, W, false, launches)) return -1;
        }
        if (ev) cudaEventRecord(ev[2 * b + 1], st);
        const bool wino2 = wino_layer(l2) && use_wino();
        if (b < 5) {  // conv2 + ReLU + 2x2 average pool + hi/lo split in one kernel -> next block's input
            if (wino2) {
                if (conv_wino(st, l2, ws, 2 * b + 1, m_hi, m_lo, in_hi, in_lo, nullptr, N, H, W, true, launches)) return -1;
            } else if (conv_tc(st, l2, ws, 2 * b + 1, m_hi, m_lo, in_hi, in_lo, nullptr, N, H, W, true, launches)) return -1;
            H /= 2;
            W /= 2;
        } else {
            if (wino2) {
                if (conv_wino(st, l2, ws, 2 * b + 1, m_hi, m_lo, nullptr, nullptr, c2, N, H, W, false, launches)) return -1;
            } else if (conv_tc(st, l2, ws, 2 * b + 1, m_hi, m_lo, nullptr, nullptr, c2, N, H, W, false, launches)) return -1;
        }
    }
    if (ev) cudaEventRecord(ev[12], st);
    cudaError_t e = launch_global_pool(st, c2, pooled, N, H, W, 2048, launches);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return tc_fail("tc_encoder_forward", cudaGetErrorString(e));
    return 0;
}

}  // namespace stito
