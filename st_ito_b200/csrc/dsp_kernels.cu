// DSP kernels of the effect chain (K1): parametric EQ, compressor, distortion, delay, Freeverb,
// peak tracking.  Replaces the per-candidate CPU loop of the reference's process_audio
// (st_ito/style_transfer.py:45-115) and the Basic* plugins (st_ito/effects.py:800-959).
//
// Parallelisation (candidate x channel x TIME where the recurrence allows it):
//   EQ          6 cascaded biquads are LTI: the signal is cut into 512-sample chunks, every chunk
//               is filtered from a zero state in parallel (fp64), the 12-dim chunk-boundary states
//               are stitched with the 12x12 state-transition matrix, then every chunk is re-run from
//               its true initial state.  Arithmetic per sample is scipy.signal.lfilter's order.
//   compressor  the ballistics filter is a data-dependent (non-linear) recurrence: 64-sample chunks run
//               in parallel from guessed states and a Newton / policy iteration (one affine scan over the
//               chunk boundaries per step) makes the guesses consistent in a handful of passes.
//   Freeverb    delay lines live in shared memory; comb feedback has a lag >= 1214 samples and all-pass
//               >= 244, so super-steps of 1120 samples (combs) / sub-blocks of 224 (all-passes) are
//               time-parallel; the only lag-1 recurrence (the comb damping one-pole) is a 32-lane affine
//               scan.  Comb warps and all-pass warps of a CTA run concurrently (see reverb_core_kernel).
//   delay       lag-d feedback: the d residue classes are independent serial chains.
// All float arithmetic uses explicit _rn intrinsics where the oracle (compiled with
// -ffp-contract=off) rounds each operation separately.
#include <cstdio>
#include <cstdlib>

#include "stito_internal.h"

namespace stito {

namespace {

__device__ __forceinline__ float load_in(const SigView &v, int p, int c, int64_t n) {
    return __ldg(v.base + (int64_t)p * v.stride_p + (int64_t)c * v.stride_c + n);
}

__device__ __forceinline__ float clip_peak(const float *in_peak, int p) {
    // np.clip(np.max(np.abs(x)), a_min=1e-8): peaks are stored as float bits (atomicMax on unsigned)
    return fmaxf(in_peak[p], 1e-8f);
}

__device__ __forceinline__ void atomic_peak(unsigned *peak, int p, float v) {
    // v >= 0: IEEE ordering of non-negative floats equals unsigned ordering of their bits
    if (v > 0.0f) atomicMax(peak + p, __float_as_uint(v));
}

// ---- streaming hand-off between two concurrently resident kernels (small populations: see run_chain in stito_api.cu)
// The producer publishes one flag per (stream, granule of 2^kGranuleShift samples) once that part of its output is in
// global memory; the consumer, which walks time in order, polls the flag of the next granule it needs.
constexpr int kGranuleShift = 15;  // 32768 samples = one compressor super-block

__device__ __forceinline__ void publish_granule(int *flags, int idx) {  // one thread, after a CTA barrier
    __threadfence();
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + idx), "r"(1) : "memory");
}
__device__ __forceinline__ void await_granule(const int *flags, int idx) {  // one thread; bounded: a lost producer must
    int v = 0;                                                                // fail loudly, never hang the GPU
    unsigned long long t0 = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + idx) : "memory");
        if (v != 0) return;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t0 == 0) t0 = t1;
        if (t1 - t0 > 4000000000ull) {  // 4 s
            printf("libstito: streaming hand-off timeout (block %d flag %d)\n", blockIdx.x, idx);
            __trap();
        }
        __nanosleep(200);
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------ EQ
// One sample through the cascade, scipy.signal.lfilter evaluation order:
//   y = z0 + b0*x;  z0 = (z1 + b1*x) - a1*y;  z1 = b2*x - a2*y       cf = {b0,b1,b2,a1,a2}
template <bool EXACT = true>
__device__ __forceinline__ double eq_step(double v, const double (&cf)[6][5], double (&z0)[6],
                                          double (&z1)[6]) {
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        if (EXACT) {  // lfilter's operation order, every product and sum rounded separately: bit-identical output
            const double y = __dadd_rn(z0[s], __dmul_rn(cf[s][0], v));
            z0[s] = __dsub_rn(__dadd_rn(z1[s], __dmul_rn(cf[s][1], v)), __dmul_rn(cf[s][3], y));
            z1[s] = __dsub_rn(__dmul_rn(cf[s][2], v), __dmul_rn(cf[s][4], y));
            v = y;
        } else {      // fused multiply-adds (5 instead of 9 fp64 instructions): only for the zero-state pre-pass,
                      // whose result seeds chunk states that are rounding-level approximations anyway
            const double y = fma(cf[s][0], v, z0[s]);
            z0[s] = fma(-cf[s][3], y, fma(cf[s][1], v, z1[s]));
            z1[s] = fma(-cf[s][4], y, cf[s][2] * v);
            v = y;
        }
    }
    return v;
}

__device__ __forceinline__ void eq_load_coefs(const double *coefs, int p, double (&cf)[6][5]) {
#pragma unroll
    for (int s = 0; s < 6; ++s)
#pragma unroll
        for (int k = 0; k < 5; ++k) cf[s][k] = __ldg(coefs + ((int64_t)p * 6 + s) * 5 + k);
}

// APPLY = false: zero-state run, writes the chunk's final 12-state to state[stream][i][k].
// APPLY = true : starts from state[stream][i][k], writes the filtered chunk and tracks the peak.
template <bool APPLY>
__global__ void __launch_bounds__(128) eq_chunk_kernel(SigView in, const float *in_peak, float *out,
                                                       int chs, int64_t L, int K, const double *coefs,
                                                       double *state, unsigned *out_peak) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int stream = blockIdx.y;
    const int p = stream / chs, c = stream - p * chs;
    const bool active = k < K;
    float pk = 0.0f;
    if (active) {
        double cf[6][5], z0[6], z1[6];
        eq_load_coefs(coefs, p, cf);
        double *st = state + (int64_t)stream * kEqStates * K + k;
        if (APPLY) {
#pragma unroll
            for (int s = 0; s < 6; ++s) { z0[s] = st[(int64_t)(2 * s) * K]; z1[s] = st[(int64_t)(2 * s + 1) * K]; }
        } else {
#pragma unroll
            for (int s = 0; s < 6; ++s) { z0[s] = 0.0; z1[s] = 0.0; }
        }
        const int64_t n0 = (int64_t)k * kEqChunk;
        const int len = (int)min((int64_t)kEqChunk, L - n0);
        const float *src = in.base + (int64_t)p * in.stride_p + (int64_t)c * in.stride_c + n0;
        float *dst = APPLY ? out + (int64_t)stream * L + n0 : nullptr;
        const bool has_div = in_peak != nullptr;
        const float div = has_div ? clip_peak(in_peak, p) : 1.0f;
        const bool vec = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (len == kEqChunk) &&
                         (!APPLY || (reinterpret_cast<uintptr_t>(dst) & 15) == 0);
        if (vec) {
            for (int i = 0; i < kEqChunk; i += 4) {
                float4 x = __ldg(reinterpret_cast<const float4 *>(src + i));
                if (has_div) { x.x = x.x / div; x.y = x.y / div; x.z = x.z / div; x.w = x.w / div; }
                float4 y;
                y.x = (float)eq_step<APPLY>((double)x.x, cf, z0, z1);
                y.y = (float)eq_step<APPLY>((double)x.y, cf, z0, z1);
                y.z = (float)eq_step<APPLY>((double)x.z, cf, z0, z1);
                y.w = (float)eq_step<APPLY>((double)x.w, cf, z0, z1);
                if (APPLY) {
                    *reinterpret_cast<float4 *>(dst + i) = y;
                    pk = fmaxf(pk, fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w))));
                }
            }
        } else {
            for (int i = 0; i < len; ++i) {
                float x = __ldg(src + i);
                if (has_div) x = x / div;
                const float y = (float)eq_step<APPLY>((double)x, cf, z0, z1);
                if (APPLY) { dst[i] = y; pk = fmaxf(pk, fabsf(y)); }
            }
        }
        if (!APPLY) {
#pragma unroll
            for (int s = 0; s < 6; ++s) { st[(int64_t)(2 * s) * K] = z0[s]; st[(int64_t)(2 * s + 1) * K] = z1[s]; }
        }
    }
    if (APPLY && out_peak != nullptr) {
        pk = warp_max(pk);
        if ((threadIdx.x & 31) == 0) atomic_peak(out_peak, p, pk);
    }
}

// Stitch: S[k+1] = M * S[k] + F[k], S[0] = 0, where M (12 x 12) is the zero-input transition of the 12-state
// cascade over one full chunk and F[k] the zero-state final state of chunk k.  The scan itself is sequential
// (lane i < 12 of warp 0 owns row i of M and component i of S; ~100 cycles per chunk): the cascade's transition
// matrix is far from normal (states of a 20 Hz biquad are huge and cancel), and re-associating the scan through
// powers of M costs ~2 decimal digits, enough to flip the float32 rounding of a third of the output samples.
// What made the old one-warp kernel slow (0.35 ms) was not arithmetic but one dependent global load of F per
// step; here the CTA stages F in shared memory in tiles of 1024 chunks (coalesced), overlapped with the
// construction of M, and S overwrites F in place and is written back coalesced.
constexpr int kStitchThreads = 512;
constexpr int kStitchTile = 1024;  // chunks per shared-memory tile: 12 x 1024 x 8 B = 96 KB

__device__ __forceinline__ double stitch_matvec(const double (&row)[kEqStates], double s, double f) {
    double a0 = f, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int j = 0; j < kEqStates; j += 3) {
        a0 = fma(row[j], __shfl_sync(0xffffffffu, s, j), a0);
        a1 = fma(row[j + 1], __shfl_sync(0xffffffffu, s, j + 1), a1);
        a2 = fma(row[j + 2], __shfl_sync(0xffffffffu, s, j + 2), a2);
    }
    return a0 + (a1 + a2);
}

__global__ void __launch_bounds__(kStitchThreads) eq_stitch_kernel(int chs, int K, const double *coefs,
                                                                    const double *zs_final, double *init) {
    extern __shared__ double fs[];                      // [kEqStates][kStitchTile]
    __shared__ double Msm[kEqStates][kEqStates + 1];
    const int stream = blockIdx.x;
    const int p = stream / chs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *F = zs_final + (int64_t)stream * kEqStates * K;
    double *S = init + (int64_t)stream * kEqStates * K;
    auto load_tile = [&](int t0, int nt, int first, int nthreads) {
        for (int idx = first; idx < kEqStates * nt; idx += nthreads) {
            const int r = idx / nt, k = idx - r * nt;
            fs[r * kStitchTile + k] = F[(int64_t)r * K + t0 + k];
        }
    };
    if (warp == 0) {
        if (lane < kEqStates) {  // column `lane` of M: propagate the basis state e_lane for one chunk
            double cf[6][5];
            eq_load_coefs(coefs, p, cf);
            double z0[6], z1[6];
#pragma unroll
            for (int s = 0; s < 6; ++s) {
                z0[s] = (lane == 2 * s) ? 1.0 : 0.0;
                z1[s] = (lane == 2 * s + 1) ? 1.0 : 0.0;
            }
            for (int i = 0; i < kEqChunk; ++i) (void)eq_step<false>(0.0, cf, z0, z1);
#pragma unroll
            for (int s = 0; s < 6; ++s) { Msm[2 * s][lane] = z0[s]; Msm[2 * s + 1][lane] = z1[s]; }
        }
    } else {
        load_tile(0, min(kStitchTile, K), tid - 32, kStitchThreads - 32);  // meanwhile, warps 1.. stage the first tile
    }
    __syncthreads();
    const int r = lane < kEqStates ? lane : 0;
    double row[kEqStates];
#pragma unroll
    for (int j = 0; j < kEqStates; ++j) row[j] = Msm[r][j];
    double s = 0.0;
    for (int t0 = 0; t0 < K; t0 += kStitchTile) {
        const int nt = min(kStitchTile, K - t0);
        if (t0 > 0) {
            load_tile(t0, nt, tid, kStitchThreads);
            __syncthreads();
        }
        if (warp == 0) {
            double *fr = fs + r * kStitchTile;
            double fnext = fr[0];
            for (int k = 0; k < nt; ++k) {
                const double f = fnext;
                if (k + 1 < nt) fnext = fr[k + 1];
                if (lane < kEqStates) fr[k] = s;  // S[k] replaces F[k]
                s = stitch_matvec(row, s, f);
            }
        }
        __syncthreads();
        for (int idx = tid; idx < kEqStates * nt; idx += kStitchThreads) {
            const int rr = idx / nt, k = idx - rr * nt;
            S[(int64_t)rr * K + t0 + k] = fs[rr * kStitchTile + k];
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------- compressor
// juce::dsp::Compressor<float> restated in oracle/dsp_oracle.c: oracle_compressor.
//     env[n] = a + c*(env[n-1] - a),  a = |x[n]|,  c = a > env[n-1] ? cteAT : cteRL        (ballistics)
//     y[n]   = x[n] * (env[n] < thr ? 1 : (env[n]/thr)^(1/ratio - 1))                       (gain computer)
// The ballistics filter is a NON-linear first-order recurrence (the coefficient depends on the state), so it
// cannot be scanned like the biquads.  A sample-serial lane needs ~22 cycles/sample = 5.6 ms for 10 s of
// audio no matter how many streams are in flight, which made it the longest kernel of a generation.
//
// Time-parallel formulation (policy iteration / Newton on a piecewise-linear map).  One CTA per stream walks
// the signal in super-blocks of 512 chunks x 64 samples staged (transposed, conflict-free pitch 65) in shared
// memory.  Every thread owns one chunk and repeatedly
//   1. runs the exact recurrence (the oracle's operation order) over its chunk from its current guess of the
//      incoming state s_in[t], producing the outgoing state s_out[t] and the slope m[t] = prod c_n of the chunk
//      map along that trajectory (with the branch decisions frozen the chunk map is affine with that slope);
//   2. the CTA solves the linearised consistency equations  d[t] = (s_out[t-1] - s_in[t]) + m[t-1]*d[t-1],
//      d[0] = 0, with one affine scan over the 512 chunks and corrects s_in += d.
// Each step re-evaluates the decisions against the corrected trajectory (policy improvement); it converges in
// 4-7 iterations for typical settings (<= 12 for 0.1 ms attack / 1 s release).  Iteration stops when every
// chunk boundary matches to 2^-20 relative; the last pass applies the gain computer and writes the output.
// The result is the serial recurrence up to float32 rounding: any evaluation order of this filter differs from
// another by ~1e-6 relative because the contraction 1 - c ~ 1e-4 amplifies each rounding ~50x (tests: 5e-6).
constexpr int kCsT = 512;               // chunks (= threads) per super-block
constexpr int kCsC = 64;                // samples per chunk
constexpr int kCsPitch = kCsC + 1;      // shared-memory row pitch: bank = (t + j) % 32, conflict-free
constexpr int kCsSB = kCsT * kCsC;      // samples per super-block
constexpr int kCsMaxIter = 16;
constexpr size_t kCsSmem = (size_t)(kCsT * kCsPitch + 2 * kCsT + 2 * (kCsT / 32) + 4) * sizeof(float);

__global__ void __launch_bounds__(kCsT, 1) compressor_scan_kernel(SigView in, const float *in_peak, float *out,
                                                                   int chs, int64_t L, const CompParams *prm,
                                                                   unsigned *out_peak, int *noconv, int *ready) {
    extern __shared__ float cs_sm[];
    float *xs = cs_sm;                       // [kCsT][kCsPitch] samples, later overwritten by the output
    float *sout_s = xs + kCsT * kCsPitch;    // [kCsT] outgoing state of each chunk
    float *slope_s = sout_s + kCsT;          // [kCsT] slope of each chunk map
    float *wM = slope_s + kCsT;              // [16] per-warp scan totals
    float *wE = wM + kCsT / 32;
    float *carry_s = wE + kCsT / 32;         // state entering the next super-block

    const int stream = blockIdx.x;
    const int p = stream / chs, c = stream - p * chs;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const CompParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? clip_peak(in_peak, p) : 1.0f;
    const float *src = in.base + (int64_t)p * in.stride_p + (int64_t)c * in.stride_c;
    float *dst = out + (int64_t)stream * L;
    float *row = xs + t * kCsPitch;
    float pk = 0.0f;
    if (t == 0) *carry_s = 0.0f;

    // The next super-block is prefetched into registers (16 x 16 B per thread = the whole 128 KB super-block
    // in flight at once) while the current one is iterated on; float4 loads when the stream is 16-byte aligned.
    constexpr int kPre = kCsSB / 4 / kCsT;  // 16
    float4 pre[kPre];
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    auto issue_loads = [&](int64_t b) {
        const int n = (int)max((int64_t)0, min((int64_t)kCsSB, L - b));
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int i4 = 4 * (t + k * kCsT);
            if (aligned && i4 + 3 < n) {
                pre[k] = __ldg(reinterpret_cast<const float4 *>(src + b + i4));
            } else {
                pre[k].x = i4 + 0 < n ? __ldg(src + b + i4 + 0) : 0.0f;
                pre[k].y = i4 + 1 < n ? __ldg(src + b + i4 + 1) : 0.0f;
                pre[k].z = i4 + 2 < n ? __ldg(src + b + i4 + 2) : 0.0f;
                pre[k].w = i4 + 3 < n ? __ldg(src + b + i4 + 3) : 0.0f;
            }
        }
    };
    issue_loads(0);

    for (int64_t b0 = 0; b0 < L; b0 += kCsSB) {
        const int nb = (int)min((int64_t)kCsSB, L - b0);
        // park the prefetched super-block in shared memory, transposed into chunk rows
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int f = t + k * kCsT;
            float *d4 = xs + (f >> 4) * kCsPitch + (f & 15) * 4;
            float4 v = pre[k];
            if (has_div) { v.x = v.x / div; v.y = v.y / div; v.z = v.z / div; v.w = v.w / div; }
            d4[0] = v.x; d4[1] = v.y; d4[2] = v.z; d4[3] = v.w;
        }
        __syncthreads();
        issue_loads(b0 + kCsSB);
        const int len = max(0, min(kCsC, nb - t * kCsC));
        float s_in = *carry_s;

        bool converged = false;
        for (int it = 0; it < kCsMaxIter; ++it) {
            float env = s_in, slope = 1.0f;
            for (int j = 0; j < len; ++j) {
                const float a = fabsf(row[j]);
                const float d = __fsub_rn(env, a);
                const bool at = a > env;
                env = __fadd_rn(a, at ? __fmul_rn(q.cte_at, d) : __fmul_rn(q.cte_rl, d));
                slope *= at ? q.cte_at : q.cte_rl;
            }
            sout_s[t] = env;
            slope_s[t] = slope;
            __syncthreads();
            // mismatch at my incoming boundary and the affine map d[t-1] -> d[t]
            float M = 0.0f, E = 0.0f;
            bool tight = true, loose = true;
            if (t > 0) {
                E = __fsub_rn(sout_s[t - 1], s_in);
                M = slope_s[t - 1];
                const float ref = fmaxf(fabsf(s_in), 1e-20f);
                tight = fabsf(E) <= 9.5367431640625e-07f * ref;   // 2^-20
                loose = fabsf(E) <= 7.62939453125e-06f * ref;     // 2^-17
            }
            const int all_tight = __syncthreads_and(tight);
            const int all_loose = __syncthreads_and(loose);
            if (all_tight || (it >= 8 && all_loose)) { converged = true; break; }
            // inclusive scan of the affine maps (composition: later o earlier)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float Mp = __shfl_up_sync(0xffffffffu, M, o);
                const float Ep = __shfl_up_sync(0xffffffffu, E, o);
                if (lane >= o) { E = fmaf(M, Ep, E); M = M * Mp; }
            }
            if (lane == 31) { wM[warp] = M; wE[warp] = E; }
            __syncthreads();
            if (warp > 0) {
                float PM = wM[0], PE = wE[0];
                for (int w = 1; w < warp; ++w) { PE = fmaf(wM[w], PE, wE[w]); PM = wM[w] * PM; }
                E = fmaf(M, PE, E);
            }
            s_in += E;
            __syncthreads();  // wM / wE / sout_s / slope_s are rewritten next iteration
        }

        if (!converged) {
            // The policy iteration is monotone and finite in exact arithmetic; in float32 a decision can keep flipping at
            // rounding level.  Never hand out an unconverged state silently: thread 0 walks the super-block serially (the
            // oracle's own loop, ~0.4 ms) and the event is counted in *noconv (stito_timing.comp_fallbacks).
            __syncthreads();
            if (t == 0) {
                float env = *carry_s;
                for (int tt = 0; tt < kCsT; ++tt) {
                    sout_s[tt] = env;  // state entering chunk tt
                    const float *r = xs + tt * kCsPitch;
                    const int ln = max(0, min(kCsC, nb - tt * kCsC));
                    for (int j = 0; j < ln; ++j) {
                        const float a = fabsf(r[j]);
                        const float d = __fsub_rn(env, a);
                        env = __fadd_rn(a, (a > env) ? __fmul_rn(q.cte_at, d) : __fmul_rn(q.cte_rl, d));
                    }
                }
                if (noconv != nullptr) atomicAdd(noconv, 1);
            }
            __syncthreads();
            s_in = sout_s[t];
            __syncthreads();
        }
        // final pass: exact recurrence from the converged state + gain computer; output replaces the input row
        {
            float env = s_in;
            for (int j = 0; j < len; ++j) {
                const float x = row[j];
                const float a = fabsf(x);
                const float d = __fsub_rn(env, a);
                env = __fadd_rn(a, (a > env) ? __fmul_rn(q.cte_at, d) : __fmul_rn(q.cte_rl, d));
                // (env / thr)^expo as exp2(expo * log2(.)): CUDA's full-accuracy powf is ~150 instructions and was
                // 70 % of this kernel; log2f / exp2f are <= 1-2 ulp, and the exponent (|.| <= 13.3 for the -80 dB
                // threshold floor) adds <= 8e-7 absolute, i.e. <= 6e-7 relative on the gain -- inside this filter's
                // float32 conditioning noise (see the header comment)
                const float g = (env < q.thr) ? 1.0f : exp2f(__fmul_rn(q.expo, log2f(__fmul_rn(env, q.thr_inv))));
                const float y = __fmul_rn(g, x);
                row[j] = y;
                pk = fmaxf(pk, fabsf(y));
            }
            if (t == kCsT - 1) *carry_s = env;
        }
        __syncthreads();
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll 4
            for (int k = 0; k < kPre; ++k) {
                const int f = t + k * kCsT, i4 = 4 * f;
                const float *s4 = xs + (f >> 4) * kCsPitch + (f & 15) * 4;
                if (i4 + 3 < nb) {
                    *reinterpret_cast<float4 *>(dst + b0 + i4) = make_float4(s4[0], s4[1], s4[2], s4[3]);
                } else {
                    for (int e = 0; e < 4; ++e) if (i4 + e < nb) dst[b0 + i4 + e] = s4[e];
                }
            }
        } else {
#pragma unroll 8
            for (int idx = t; idx < nb; idx += kCsT) dst[b0 + idx] = xs[(idx >> 6) * kCsPitch + (idx & 63)];
        }
        __syncthreads();
        // streaming: this super-block of the output is complete -> the reverb CTA of this stream may consume it
        if (ready != nullptr && t == 0)
            publish_granule(ready, stream * (int)((L + kCsSB - 1) / kCsSB) + (int)(b0 / kCsSB));
    }
    if (out_peak != nullptr) {
        pk = warp_max(pk);
        if (lane == 0) atomic_peak(out_peak, p, pk);
    }
}

// --------------------------------------------------------------------- distortion
__global__ void __launch_bounds__(256) distortion_kernel(SigView in, const float *in_peak, float *out,
                                                         int chs, int64_t L, const DistParams *prm,
                                                         unsigned *out_peak) {
    const int stream = blockIdx.y;
    const int p = stream / chs, c = stream - p * chs;
    const DistParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? clip_peak(in_peak, p) : 1.0f;
    float pk = 0.0f;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < L; n += (int64_t)gridDim.x * blockDim.x) {
        float x = load_in(in, p, c, n);
        if (has_div) x = x / div;
        const float y = __fmul_rn(tanhf(__fmul_rn(x, q.drive)), q.out_gain);
        out[(int64_t)stream * L + n] = y;
        pk = fmaxf(pk, fabsf(y));
    }
    if (out_peak != nullptr) {
        pk = warp_max(pk);
        if ((threadIdx.x & 31) == 0) atomic_peak(out_peak, p, pk);
    }
}

// -------------------------------------------------------------------------- delay
// line[n] = x[n] + fb*line[n-d];  y[n] = dry*x[n] + mix*line[n-d]   (oracle_delay)
__global__ void __launch_bounds__(256) delay_kernel(SigView in, const float *in_peak, float *out, int chs,
                                                    int64_t L, const DelayParams *prm,
                                                    unsigned *out_peak) {
    const int stream = blockIdx.y;
    const int p = stream / chs, c = stream - p * chs;
    const DelayParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? clip_peak(in_peak, p) : 1.0f;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    float pk = 0.0f;
    if (r < q.d) {
        float prev = 0.0f;
        for (int64_t n = r; n < L; n += q.d) {
            float x = load_in(in, p, c, n);
            if (has_div) x = x / div;
            const float o = prev;
            prev = __fadd_rn(x, __fmul_rn(q.feedback, o));
            const float y = __fadd_rn(__fmul_rn(q.dry, x), __fmul_rn(q.mix, o));
            out[(int64_t)stream * L + n] = y;
            pk = fmaxf(pk, fabsf(y));
        }
    }
    if (out_peak != nullptr) {
        pk = warp_max(pk);
        if ((threadIdx.x & 31) == 0) atomic_peak(out_peak, p, pk);
    }
}

// ------------------------------------------------------------------------- reverb
// JUCE_UNDENORMALISE on x86: x += 0.1f; x -= 0.1f
__device__ __forceinline__ float undenorm(float v) { return __fsub_rn(__fadd_rn(v, 0.1f), 0.1f); }

constexpr int kRevMaxSeg = 8;  // samples per lane per block (block <= 256)

// juce::Reverb (Freeverb) restated in oracle/dsp_oracle.c: oracle_reverb.  NCH = 2: processStereo,
// one CTA per candidate; NCH = 1: processMono, one CTA per (candidate, channel).
template <int NCH>
__global__ void __launch_bounds__(NCH * 256) reverb_kernel(SigView in, const float *in_peak, float *out,
                                                           int chs, int64_t L, ReverbGeom g,
                                                           const ReverbParams *prm, unsigned *out_peak) {
    extern __shared__ float sm[];
    float *ring = sm;
    float *mixin = sm + g.total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int inst = blockIdx.x;
    const int p = NCH == 2 ? inst : inst / chs;
    const int c0 = NCH == 2 ? 0 : inst - p * chs;
    const int B = g.block;
    const int seg = B >> 5;
    for (int i = tid; i < g.total; i += blockDim.x) ring[i] = 0.0f;
    const ReverbParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? clip_peak(in_peak, p) : 1.0f;
    const float keep = __fsub_rn(1.0f, q.damp);

    // phase-1 role: thread -> (sample nloc, channel c)
    const int c = NCH == 2 ? (tid & 1) : 0;
    const int nloc = NCH == 2 ? (tid >> 1) : tid;
    const bool p1 = tid < NCH * B;
    int cb[8], ab[4];  // ring positions of sample n0 for the 8 combs / 4 all-passes of channel c
#pragma unroll
    for (int j = 0; j < 8; ++j) cb[j] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) ab[j] = 0;
    // comb role: warp -> (channel cc, comb j)
    const int cc = warp >> 3, cj = warp & 7;
    const bool comb_role = warp < NCH * 8;
    const int csize = g.comb_size[comb_role ? cc : 0][cj];
    float *cring = ring + g.comb_off[comb_role ? cc : 0][cj];
    int mycb = 0;
    float fstore = 0.0f;  // comb filterstore after the previous block
    float dpow = 1.0f;    // damp^seg
    for (int i = 0; i < seg; ++i) dpow = __fmul_rn(dpow, q.damp);

    float *dst = out + ((int64_t)p * chs + (NCH == 2 ? c : c0)) * L;
    const int cin = NCH == 2 ? c : c0;
    float pk = 0.0f;
    float xnext = (p1 && nloc < L) ? load_in(in, p, cin, nloc) : 0.0f;
    __syncthreads();

    for (int64_t n0 = 0; n0 < L; n0 += B) {
        const int nb = (int)min((int64_t)B, L - n0);
        if (p1) {
            const bool act = nloc < nb;
            float xin = xnext;
            if (has_div) xin = xin / div;
            {   // prefetch the next block's sample; consumed after the comb phase
                const int64_t nn = n0 + B + nloc;
                xnext = nn < L ? load_in(in, p, cin, nn) : 0.0f;
            }
            float inp;
            if (NCH == 2) {
                const float other = __shfl_xor_sync(0xffffffffu, xin, 1);
                inp = __fmul_rn(__fadd_rn(xin, other), 0.015f);
            } else {
                inp = __fmul_rn(xin, 0.015f);
            }
            if (c == 0 && act) mixin[nloc] = inp;
            float v = 0.0f;  // sum of the 8 comb outputs (delayed samples: independent of this block)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int pos = cb[j] + nloc;
                const int sz = g.comb_size[c][j];
                if (pos >= sz) pos -= sz;
                v = __fadd_rn(v, ring[g.comb_off[c][j] + pos]);
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) {  // series all-passes: lag >= block, so sample-parallel
                int pos = ab[s] + nloc;
                const int sz = g.ap_size[c][s];
                if (pos >= sz) pos -= sz;
                float *slot = ring + g.ap_off[c][s] + pos;
                const float bv = *slot;
                const float t = undenorm(__fadd_rn(v, __fmul_rn(bv, 0.5f)));
                if (act) *slot = t;
                v = __fsub_rn(bv, v);
            }
            float y;
            if (NCH == 2) {
                const float vo = __shfl_xor_sync(0xffffffffu, v, 1);
                y = __fadd_rn(__fadd_rn(__fmul_rn(v, q.wet1), __fmul_rn(vo, q.wet2)), __fmul_rn(xin, q.dry));
            } else {
                y = __fadd_rn(__fmul_rn(v, q.wet1), __fmul_rn(xin, q.dry));
            }
            if (act) { dst[n0 + nloc] = y; pk = fmaxf(pk, fabsf(y)); }
#pragma unroll
            for (int j = 0; j < 8; ++j) { cb[j] += B; if (cb[j] >= g.comb_size[c][j]) cb[j] -= g.comb_size[c][j]; }
#pragma unroll
            for (int s = 0; s < 4; ++s) { ab[s] += B; if (ab[s] >= g.ap_size[c][s]) ab[s] -= g.ap_size[c][s]; }
        }
        __syncthreads();
        if (comb_role) {
            const int i0 = lane * seg;
            float o[kRevMaxSeg];
            int pos0 = mycb + i0;
            if (pos0 >= csize) pos0 -= csize;
#pragma unroll
            for (int i = 0; i < kRevMaxSeg; ++i) {
                int pos = pos0 + i;
                if (pos >= csize) pos -= csize;
                o[i] = (i < seg) ? cring[pos] : 0.0f;
            }
            float z = 0.0f;  // zero-state response of the damping one-pole over my segment
#pragma unroll
            for (int i = 0; i < kRevMaxSeg; ++i)
                if (i < seg) z = undenorm(__fadd_rn(__fmul_rn(o[i], keep), __fmul_rn(z, q.damp)));
            float A = dpow, Bv = z;  // affine map of my segment: s -> A*s + Bv; inclusive scan
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float Ap = __shfl_up_sync(0xffffffffu, A, d);
                const float Bp = __shfl_up_sync(0xffffffffu, Bv, d);
                if (lane >= d) { Bv = fmaf(Bp, A, Bv); A = A * Ap; }
            }
            const float s_out = fmaf(A, fstore, Bv);
            float s = __shfl_up_sync(0xffffffffu, s_out, 1);
            if (lane == 0) s = fstore;
#pragma unroll
            for (int i = 0; i < kRevMaxSeg; ++i) {
                if (i < seg) {
                    s = undenorm(__fadd_rn(__fmul_rn(o[i], keep), __fmul_rn(s, q.damp)));
                    const float t = undenorm(__fadd_rn(mixin[i0 + i], __fmul_rn(s, q.fb)));
                    int pos = pos0 + i;
                    if (pos >= csize) pos -= csize;
                    if (i0 + i < nb) cring[pos] = t;
                }
            }
            fstore = __shfl_sync(0xffffffffu, s, 31);
            mycb += B;
            if (mycb >= csize) mycb -= csize;
        }
        __syncthreads();
    }
    if (out_peak != nullptr) {
        pk = warp_max(pk);
        if (lane == 0) atomic_peak(out_peak, p, pk);
    }
}

// ---------------------------------------------------------------- reverb, fast path
// The two channels of juce::Reverb share only the mono input sum and the final wet cross-mix, so the
// comb/all-pass network of every (candidate, channel) runs in its own CTA and a trivial element-wise
// kernel does the mix.  All-pass delay lines are power-of-two rings indexed with (n - delay) & mask; the comb
// lines are 3-super-step rings with a mirrored head (see kCombRing).
//
// Time is cut into super-steps of S = 32*seg samples, S <= shortest comb delay, so that inside a
// super-step every delayed read refers to samples of EARLIER super-steps:
//   * 8 comb warps (one comb each): lane l owns `seg` consecutive samples; the only lag-1 recurrence (the
//     damping one-pole) is a zero-state pass + 32-lane affine scan + replay, all in registers;
//   * 7 all-pass warps (224 threads = one sub-block <= shortest all-pass delay): sum the 8 comb outputs
//     (= ring values of earlier super-steps, read straight from the comb rings) and run the 4 series
//     all-passes sample-parallel, sub-block after sub-block, synchronised by a named barrier of their own.
// The two groups do not depend on each other inside a super-step (the comb rings keep a full super-step of
// history beyond the longest delay, so the comb writes cannot clobber what the all-pass group still has to
// read) and run concurrently; one __syncthreads per super-step (429 for 10 s) is the only
// CTA-wide barrier.  The next super-step's input is prefetched into a double buffer meanwhile.
constexpr int kRevMaxSegF = 35;          // instantiated for seg = 35 (>= 44.3 kHz: S = 1120 = 5 sub-blocks) and 32
constexpr int kRevSub = 224;             // all-pass sub-block: <= shortest all-pass line at >= 44.1 kHz
constexpr int kApRing = 1024;
// Comb rings hold 3 super-steps (RL = 3 S >= longest delay + S) plus a mirror of the first super-step behind the
// end, so that a lane's SEG consecutive reads / writes are one contiguous run: base pointer + immediate offsets
// instead of an add / mask / scale per access (a third of the comb role's instructions).
constexpr int kCombRing = 4 * 32 * kRevMaxSegF;  // per-comb allocation: 3 S + mirror S at the largest S
constexpr int kRevCombThreads = 256, kRevThreads = kRevCombThreads + kRevSub;
constexpr int kRevMaxS = 32 * kRevMaxSegF;

struct ReverbFastGeom {
    int comb_delay[2][8];
    int ap_delay[2][4];
};

// inst = (p, c).  PAIR (stereo reverb on stereo audio): input (l + r) * 0.015, tunings of channel c, and the two
// CTAs of a candidate form a thread-block cluster: every wet sample is stored into the own AND (through
// distributed shared memory) the peer CTA's buffer, followed by a release-arrive on an mbarrier in the peer CTA; the
// peer's all-pass group acquires it at the start of the next super-step and mixes its own output channel
// y = wet_own*wet1 + wet_peer*wet2 + x*dry  -- the wet signal never touches HBM, and the per-super-step barrier stays a
// plain CTA barrier (a full cluster barrier with release/acquire semantics cost ~30 % of the kernel in fence stalls).  !PAIR: independent channels, input x_c * 0.015, left tunings, y = wet*wet1 + x*dry.
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int SEG, bool PAIR>
__global__ void __launch_bounds__(kRevThreads, 1) reverb_core_kernel(SigView in, const float *in_peak, float *out,
                                                                     int chs, int64_t L, ReverbFastGeom g,
                                                                     const ReverbParams *prm, unsigned *out_peak,
                                                                     const int *ready) {
    extern __shared__ float sm[];
    float *comb = sm;                        // [8][kCombRing]
    float *ap = sm + 8 * kCombRing;          // [4][kApRing]
    float *inbuf = ap + 4 * kApRing;         // [3][kRevMaxS] reverb input of super-steps k-1, k, k+1 (k % 3)
    float *xraw = inbuf + 3 * kRevMaxS;      // [3][kRevMaxS] own-channel dry input, same indexing
    float *wetb = xraw + 3 * kRevMaxS;       // [2 (super-step parity)][2 (own, peer)][kRevMaxS]
    // PAIR: "the peer's wet samples of super-step k have landed": two mbarriers, k & 1 selects, phase parity (k >> 1) & 1.
    // (With a single barrier the peer could complete phase k+1 before a delayed thread here has waited on phase k.)
    uint64_t *xbar = reinterpret_cast<uint64_t *>(wetb + 4 * kRevMaxS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int inst = blockIdx.x;
    const int p = inst / chs, c = inst - p * chs;
    const int tune = PAIR ? c : 0;
    for (int i = tid; i < 8 * kCombRing + 4 * kApRing; i += kRevThreads) sm[i] = 0.0f;
    const ReverbParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? clip_peak(in_peak, p) : 1.0f;
    constexpr int seg = SEG, S = 32 * SEG;
    constexpr int nsub = (S + kRevSub - 1) / kRevSub;
    constexpr int kPre = (kRevMaxS + kRevThreads - 1) / kRevThreads;

    // raw samples of one super-step -> (reverb input, own dry sample)
    auto park = [&](const float (&pl)[kPre], const float (&pr)[kPre], int b) {
        float *nin = inbuf + b * kRevMaxS, *nx = xraw + b * kRevMaxS;
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int i = tid + k * kRevThreads;
            float l = pl[k], r = pr[k];
            if (has_div) { l = l / div; r = r / div; }
            if (i < S) {
                nin[i] = __fmul_rn(PAIR ? __fadd_rn(l, r) : l, 0.015f);
                nx[i] = (PAIR && c == 1) ? r : l;
            }
        }
    };
    auto fetch = [&](float (&pl)[kPre], float (&pr)[kPre], int64_t base) {
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int64_t n = base + tid + k * kRevThreads;
            const bool ok = (tid + k * kRevThreads) < S && n < L;
            // ld.cg, not the read-only path: in streaming mode the producer kernel is still writing this buffer
            pl[k] = ok ? __ldcg(in.base + (int64_t)p * in.stride_p + (int64_t)(PAIR ? 0 : c) * in.stride_c + n) : 0.0f;
            pr[k] = (ok && PAIR) ? __ldcg(in.base + (int64_t)p * in.stride_p + in.stride_c + n) : 0.0f;
        }
    };
    // streaming consumer state: samples [0, avail) of my input stream(s) are known to be complete
    constexpr int S_ = 32 * SEG;
    const int ngran = (int)((L + (1 << kGranuleShift) - 1) >> kGranuleShift);
    int64_t avail = ready != nullptr ? 0 : L;
    auto need_input = [&](int64_t upto) {  // tid 0 only; followed by a CTA barrier before anybody loads
        upto = min(upto, L);
        while (avail < upto) {
            const int gidx = (int)(avail >> kGranuleShift);
            if (PAIR) { await_granule(ready, (p * 2) * ngran + gidx); await_granule(ready, (p * 2 + 1) * ngran + gidx); }
            else await_granule(ready, inst * ngran + gidx);
            avail += (1 << kGranuleShift);
        }
    };
    if (ready != nullptr) {
        if (tid == 0) need_input(2 * (int64_t)S_);
        __syncthreads();
    }
    {
        float pl[kPre], pr[kPre];
        fetch(pl, pr, 0);
        park(pl, pr, 0);
    }
    uint32_t peer_wet = 0, peer_bar = 0;  // shared::cluster addresses of the peer CTA's wetb / xbar
    if (PAIR) {
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(wetb);
        const uint32_t lbar = (uint32_t)__cvta_generic_to_shared(xbar);
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_wet) : "r"(local), "r"((uint32_t)(c ^ 1)));
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_bar) : "r"(lbar), "r"((uint32_t)(c ^ 1)));
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lbar), "r"((uint32_t)kRevSub));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lbar + 8), "r"((uint32_t)kRevSub));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_barrier();  // the peer is resident and initialised before anything is stored into it
    } else {
        __syncthreads();
    }

    const bool comb_role = tid < kRevCombThreads;
    // comb role state
    const int my_delay = g.comb_delay[tune][warp & 7];
    float *my_ring = comb + (warp & 7) * kCombRing;
    const float keep = __fsub_rn(1.0f, q.damp);
    float fstore = 0.0f, dpow = 1.0f;
#pragma unroll
    for (int i = 0; i < SEG; ++i) dpow = __fmul_rn(dpow, q.damp);
    // all-pass role state
    const int a = tid - kRevCombThreads;
    int cd[8], ad[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) cd[j] = g.comb_delay[tune][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) ad[j] = g.ap_delay[tune][j];
    float *dst = out + (int64_t)inst * L;
    float pk = 0.0f;

    // mix of one finished super-step (both wet channels are in shared memory): y = wet_own*wet1 + wet_peer*wet2 + x*dry
    auto mix = [&](int64_t m0, int par, int xb, int first, int nthreads) {
        const float *wo = wetb + par * 2 * kRevMaxS, *wpeer = wo + kRevMaxS, *xr = xraw + xb * kRevMaxS;
        const int cnt = (int)min((int64_t)S, L - m0);
        for (int i = first; i < cnt; i += nthreads) {
            float y;
            if (PAIR) y = __fadd_rn(__fadd_rn(__fmul_rn(wo[i], q.wet1), __fmul_rn(wpeer[i], q.wet2)), __fmul_rn(xr[i], q.dry));
            else y = __fadd_rn(__fmul_rn(wo[i], q.wet1), __fmul_rn(xr[i], q.dry));
            dst[m0 + i] = y;
            pk = fmaxf(pk, fabsf(y));
        }
    };

    constexpr int RL = 3 * S;  // comb ring length; [RL, RL + S) mirrors [0, S)
    int par = 0, ib = 0, wbase = 0;  // wbase = ring position of this super-step's first sample: 0, S, 2S, 0, ...
    int step = 0;
    for (int64_t n0 = 0; n0 < L; n0 += S, ++step, par ^= 1, ib = (ib == 2) ? 0 : ib + 1, wbase = (wbase == 2 * S) ? 0 : wbase + S) {
        const int nbase = (int)(n0 & (kApRing * 1024 - 1));  // only the low bits matter for the all-pass masks
        // prefetch the next super-step's input into registers now (the loads fly while this super-step is
        // computed); it is parked in the third input buffer before the barrier below
        float pre_l[kPre], pre_r[kPre];  // raw samples: no arithmetic before the end of the super-step
        fetch(pre_l, pre_r, n0 + S);
        float *wown = wetb + par * 2 * kRevMaxS;
        if (comb_role) {
            const float *inb = inbuf + ib * kRevMaxS;
            const int i0 = lane * seg;
            int rb = wbase - my_delay;
            if (rb < 0) rb += RL;
            const float *rp = my_ring + rb + i0;   // SEG contiguous delayed samples (runs into the mirror, never wraps)
            float *wp = my_ring + wbase + i0;
            float o[SEG];
#pragma unroll
            for (int i = 0; i < SEG; ++i) o[i] = rp[i];
            float z = 0.0f;  // zero-state response of the damping one-pole over my segment
#pragma unroll
            for (int i = 0; i < SEG; ++i) z = undenorm(__fadd_rn(__fmul_rn(o[i], keep), __fmul_rn(z, q.damp)));
            float A = dpow, Bv = z;  // affine map of my segment: s -> A*s + Bv; inclusive scan
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float Ap = __shfl_up_sync(0xffffffffu, A, d);
                const float Bp = __shfl_up_sync(0xffffffffu, Bv, d);
                if (lane >= d) { Bv = fmaf(Bp, A, Bv); A = A * Ap; }
            }
            const float s_out = fmaf(A, fstore, Bv);
            float sv = __shfl_up_sync(0xffffffffu, s_out, 1);
            if (lane == 0) sv = fstore;
#pragma unroll
            for (int i = 0; i < SEG; ++i) {
                sv = undenorm(__fadd_rn(__fmul_rn(o[i], keep), __fmul_rn(sv, q.damp)));
                const float tv = undenorm(__fadd_rn(inb[i0 + i], __fmul_rn(sv, q.fb)));
                wp[i] = tv;
                if (wbase == 0) wp[RL + i] = tv;  // keep the mirror of the ring head current
            }
            fstore = __shfl_sync(0xffffffffu, sv, 31);
        } else {
            // the all-pass group has slack against the comb chains: it also mixes and writes the PREVIOUS super-step.
            // First thing in the super-step, so that its global stores have drained long before the release fence of
            // the cluster barrier below (issued right before the barrier they cost ~30 % of the kernel in fence stalls).
            if (n0 > 0) {
                if (PAIR) {  // the peer's wet samples of the previous super-step have landed in my wetb (acquire)
                    const uint32_t lbar = (uint32_t)__cvta_generic_to_shared(xbar) + 8u * (uint32_t)((step - 1) & 1);
                    const uint32_t parity = (uint32_t)(((step - 1) >> 1) & 1);
                    uint32_t ok = 0;
                    while (!ok)
                        asm volatile("{\n\t.reg .pred p;\n\t"
                                     "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(lbar), "r"(parity) : "memory");
                }
                mix(n0 - S, par ^ 1, ib == 0 ? 2 : ib - 1, a, kRevSub);
            }
            const float *cp[8];  // delayed comb outputs of this super-step: contiguous runs
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int rb = wbase - cd[j];
                if (rb < 0) rb += RL;
                cp[j] = comb + j * kCombRing + rb + a;
            }
            for (int sb = 0; sb < nsub; ++sb) {
                const int off = sb * kRevSub + a;
                if (off < S) {
                    const int n = nbase + off;
                    float v = 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) v = __fadd_rn(v, cp[j][sb * kRevSub]);
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const float bv = ap[s * kApRing + ((n - ad[s]) & (kApRing - 1))];
                        const float tv = undenorm(__fadd_rn(v, __fmul_rn(bv, 0.5f)));
                        ap[s * kApRing + (n & (kApRing - 1))] = tv;
                        v = __fsub_rn(bv, v);
                    }
                    wown[off] = v;
                    if (PAIR)  // the same sample into the peer CTA's "peer" half (distributed shared memory)
                        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(peer_wet + (uint32_t)((par * 2 + 1) * kRevMaxS + off) * 4u), "f"(v) : "memory");
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kRevSub) : "memory");  // all-pass group only
            }
            if (PAIR)  // my wet samples of this super-step are in the peer's buffer: release them (one arrive per thread)
                asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(peer_bar + 8u * (uint32_t)(step & 1)) : "memory");
        }
        park(pre_l, pre_r, ib == 2 ? 0 : ib + 1);
        if (ready != nullptr && tid == 0) need_input(n0 + 3 * (int64_t)S);  // the next iteration prefetches [n0 + 2S, n0 + 3S)
        __syncthreads();
    }
    if (PAIR) {  // last super-step: wait for the peer's samples; nobody leaves while the peer may still store into it
        const uint32_t lbar = (uint32_t)__cvta_generic_to_shared(xbar) + 8u * (uint32_t)((step - 1) & 1);
        const uint32_t parity = (uint32_t)(((step - 1) >> 1) & 1);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(lbar), "r"(parity) : "memory");
    }
    {   // the last super-step is mixed by everybody
        const int64_t nsteps = (L + S - 1) / S;
        const int64_t m0 = (nsteps - 1) * S;
        mix(m0, (int)((nsteps - 1) & 1), (int)((nsteps - 1) % 3), tid, kRevThreads);
    }
    if (PAIR) cluster_barrier();  // keep my shared memory alive until the peer has finished with it
    if (out_peak != nullptr) {
        pk = warp_max(pk);
        if (lane == 0) atomic_peak(out_peak, p, pk);
    }
}

// ------------------------------------------- reverb, split over a cluster (small populations)
// With P * chs <= ~40 streams the kernel above leaves three quarters of the GPU idle while every CTA is bound by the
// dependent chains of its 8 comb warps (2.2 us per super-step, 429 super-steps for 10 s).  Here one stereo candidate is a
// cluster of 8 CTAs, 4 per channel (c = rank / 4, g = rank % 4), and every stage of the Freeverb gets its own SM(s):
//   g == 0  "home": the ALL-PASS group only (224 threads).  Per super-step it waits (transaction barrier full[slot]) for the
//           three producer rows, adds them, runs the 4 all-passes over 5 sub-blocks, hands the wet row to the two mixer
//           groups by bulk copy and frees the slot.  It is the serial stage of the pipeline, so it shares its SM with nobody:
//           with two comb filters next to it the SM's issue slots were the bottleneck (1.55 us per super-step, v11 below).
//   g == 1..3  PRODUCERS: 2, 3 and 3 comb filters (WPC warps each, SEGL samples per lane, cross-warp affine scan of the
//           damping filter).  After super-step k a producer stages ONE row -- the sum of its combs' delayed outputs that the
//           all-pass chain of super-step k + 1 needs, dly[g][slot][i] = sum_j ring_j[n - delay_j] -- and one lane hands it to
//           the bulk-copy engine (cp.async.bulk shared::cta -> shared::cluster, complete_tx on the home's full[slot]).  A ring
//           of kRsDepth slots decouples producers and consumer; free[slot] (home -> producers) returns the credit.
//   g == 1  also hosts the MIXER group of the channel (224 threads): y = wet_own * wet1 + wet_peer * wet2 + x * dry from the
//           wet rows of BOTH channels' homes (bulk copies completing on mixfull[ws], ring of kRsWet rows, credit wfree) and the
//           dry samples re-read from global memory one super-step ahead.
// The sum of the 8 comb outputs is formed as ((c0 + c1) + ((c2 + c3) + c4)) + ((c5 + c6) + c7): within an ulp of
// reverb_core_kernel's left-to-right sum.  Development log with ncu source-level measurements: profiles/r02e_reverb_split.md.
constexpr int kRsDepth = 6;  // slots of the producer -> all-pass ring (one row per producer CTA and slot)
constexpr int kRsWet = 4;    // slots of the all-pass -> mixer ring (wet rows)
constexpr int kRsProd = 3;   // producer CTAs per channel
constexpr int kRsMaxCw = 3;  // comb filters per producer CTA: 2, 3, 3
template <int WPC> struct RsCfg {
    static constexpr int kCombThreads = 32 * WPC * kRsMaxCw;
    static constexpr int kThreads = kCombThreads + kRevSub;
    static constexpr int kFloats = kRsMaxCw * kCombRing + kRsProd * kRsDepth * kRevMaxS + 4 * kApRing + 2 * kRevMaxS +
                                   kRsWet * kRevMaxS + kRsDepth * kRevMaxS +      // wet rows / delayed rows staged for the bulk copies
                                   2 * kRsMaxCw * WPC * 2 + 2 * kRsMaxCw + 2;     // per-warp scan totals (float2, double-buffered) + carries
    static constexpr size_t kSmem = (size_t)kFloats * sizeof(float) + (2 * kRsDepth + 2 * kRsWet) * sizeof(uint64_t) + 16;
};

// mbarrier waits of reverb_split_kernel, bounded (a protocol bug must trap, not hang the GPU).
//   kTx      : the barrier completes through the bulk-copy engine's complete_tx on THIS CTA's barrier -- the canonical
//              transaction-barrier wait (acquire at CTA scope), also when the copy was issued by another CTA of the cluster;
//   kControl : control only (relaxed): the data the hand-off protects is ordered by the CTA barrier that follows.
// An acquire at CLUSTER scope compiles to CCTL.IVALL -- an L1 invalidation per waiter and per poll; with all 320 comb threads
// waiting that way it was 20 % of the comb group's time (profiles/r02e_reverb_split.md, v10).
enum MbarWait { kTx, kControl };
template <MbarWait MODE> __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    unsigned long long t0 = 0;
    while (!ok) {
        if (MODE == kTx)
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && (++spins & 1023u) == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t0 == 0) t0 = t1;
            if (t1 - t0 > 4000000000ull) {
                printf("libstito: reverb_split mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
                __trap();
            }
        }
    }
}

// SEGL samples per lane, WPC warps per comb filter: S = 32 * WPC * SEGL samples per super-step (7 x 5 -> 1120, 8 x 4 -> 1024)
template <int SEGL, int WPC>
__global__ void __launch_bounds__(RsCfg<WPC>::kThreads, 1) reverb_split_kernel(SigView in, float *out, int64_t L, ReverbFastGeom g,
                                                                              const ReverbParams *prm, unsigned *out_peak,
                                                                              const int *ready) {
    using Cfg = RsCfg<WPC>;
    constexpr int S = 32 * WPC * SEGL, RL = 3 * S;
    constexpr int nsub = (S + kRevSub - 1) / kRevSub;
    extern __shared__ float sm[];
    float *ring = sm;                                     // [kRsMaxCw][kCombRing] my comb filters (producers)
    float *dly = ring + kRsMaxCw * kCombRing;             // home: [kRsProd][kRsDepth][kRevMaxS] delayed rows; g == 1: the mixer's wet[kRsWet][2][kRevMaxS]
    float *ap = dly + kRsProd * kRsDepth * kRevMaxS;      // [4][kApRing] (home)
    float *inbuf = ap + 4 * kApRing;                      // [2][kRevMaxS] reverb input (l + r) * 0.015 of super-steps k, k + 1 (producers)
    float *wetb = inbuf + 2 * kRevMaxS;                   // [kRsWet][kRevMaxS] wet rows staged for the mixers (home)
    float *stage = wetb + kRsWet * kRevMaxS;              // [kRsDepth][kRevMaxS] delayed rows staged for the bulk copy (producers)
    float *wtot = stage + kRsDepth * kRevMaxS;            // [2][kRsMaxCw][WPC] float2: affine map of each warp's part of the damping scan
    float *carry = wtot + 2 * kRsMaxCw * WPC * 2;         // [2][kRsMaxCw] damping-filter state entering the next super-step
    uint64_t *bars = reinterpret_cast<uint64_t *>(carry + 2 * kRsMaxCw + 2);  // full[kRsDepth], free[kRsDepth], wfree[kRsWet], mixfull[kRsWet]
    static_assert((Cfg::kFloats % 2) == 0 && ((kRsMaxCw * kCombRing) % 4) == 0 && (kRevMaxS % 4) == 0 && (kApRing % 4) == 0, "alignment");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = blockIdx.x % 8;
    const int p = blockIdx.x / 8;
    const int c = rank / 4, gq = rank % 4;
    const bool home = gq == 0;
    const int ncomb = home ? 0 : (gq == 1 ? 2 : 3);                 // comb filters of this CTA ...
    const int comb0 = gq == 1 ? 0 : (gq == 2 ? 2 : 5);              // ... starting at this global index
    const uint32_t home_rank = (uint32_t)(c * 4), peer_rank = (uint32_t)((c ^ 1) * 4);
    for (int i = tid; i < Cfg::kFloats; i += Cfg::kThreads) sm[i] = 0.0f;
    const ReverbParams q = prm[p];
    const int64_t nsteps = (L + S - 1) / S;
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    auto map_to = [](uint32_t local, uint32_t target_rank) {
        uint32_t r;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(target_rank));
        return r;
    };
    if (tid == 0) {
        for (int s2 = 0; s2 < kRsDepth; ++s2) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * s2), "r"(1u));                // full (home): kRsProd rows of bytes
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (kRsDepth + s2)), "r"(1u));   // free (producers): the home's all-pass group
        }
        for (int s2 = 0; s2 < kRsWet; ++s2) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (2 * kRsDepth + s2)), "r"(2u));           // wfree (home): both mixers
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8u * (2 * kRsDepth + kRsWet + s2)), "r"(1u));  // mixfull (mixer): 2 wet rows of bytes
        }
        if (home)  // arm the first phases: one delayed row of S floats per producer and slot
            for (int s2 = 0; s2 < kRsDepth; ++s2)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * s2), "r"((uint32_t)(kRsProd * S * 4)) : "memory");
        if (gq == 1)  // the mixer: one wet row from each channel's home per slot
            for (int s2 = 0; s2 < kRsWet; ++s2)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * (2 * kRsDepth + kRsWet + s2)), "r"((uint32_t)(2 * S * 4)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_barrier();  // every CTA of the cluster is resident, zeroed and initialised before anything is stored into it

    const int ngran = (int)((L + (1 << kGranuleShift) - 1) >> kGranuleShift);
    int64_t avail = ready != nullptr ? 0 : L;
    auto need_input = [&](int64_t upto) {  // one thread per group; followed by that group's barrier
        upto = min(upto, L);
        while (avail < upto) {
            const int gidx = (int)(avail >> kGranuleShift);
            await_granule(ready, (p * 2) * ngran + gidx);
            await_granule(ready, (p * 2 + 1) * ngran + gidx);
            avail += (1 << kGranuleShift);
        }
    };
    const float *xl = in.base + (int64_t)p * in.stride_p;
    const float *xr = xl + in.stride_c;

    // Roles of this CTA's threads.  Producer CTAs: comb warps (cw < ncomb) and a STAGER group that builds and ships the
    // rows for the all-pass chain concurrently with the comb warps' next super-step -- g == 1: the unused third comb slot
    // (32 * WPC threads; the 224 threads behind the comb slots are the mixer); g == 2, 3: the 224 threads behind the comb slots.
    const bool comb_slot = tid < Cfg::kCombThreads;
    const int cw = warp / WPC, ww = warp % WPC;
    const bool is_comb = comb_slot && cw < ncomb;
    const bool is_stager = !home && (gq == 1 ? (comb_slot && cw == 2) : !comb_slot);
    const int nthr = 32 * WPC * ncomb;                     // comb threads of this CTA (320 or 480 at WPC = 5)
    const int nst = gq == 1 ? 32 * WPC : kRevSub;          // stager threads
    auto comb_bar = [&]() { asm volatile("bar.sync 2, %0;" ::"r"(nthr) : "memory"); };                  // (A): comb warps
    auto ring_bar = [&]() { asm volatile("bar.sync 3, %0;" ::"r"(nthr + nst) : "memory"); };            // (B): comb warps + stagers
    auto stager_bar = [&]() { asm volatile("bar.sync 4, %0;" ::"r"(nst) : "memory"); };

    if (is_comb) {
        // ------------------------------------------------------------------ comb group (producer CTAs)
        // One comb filter = WPC warps; lane l = ww * 32 + lane of the comb owns SEGL consecutive samples of the super-step.
        // (One warp per comb with 35 samples per lane is a ~980-instruction dependent chain per super-step -- 2 us, which
        // is what bounds reverb_core_kernel as well.)  The loop is bound by its own dependent chains (clock trace in
        // profiles/r02e_reverb_split.md: ~930 cycles to barrier (A), ~720 to barrier (B)), so everything that is not the
        // recurrence -- building and shipping rows, flag polling, credit waits -- lives in the stager group.
        const int my_delay = g.comb_delay[c][comb0 + cw];
        float *my_ring = ring + cw * kCombRing;
        const float keep = __fsub_rn(1.0f, q.damp);
        float dpow = 1.0f;
#pragma unroll
        for (int i = 0; i < SEGL; ++i) dpow = __fmul_rn(dpow, q.damp);
        const int lc = ww * 32 + lane;                 // lane index inside the comb
        const int i0 = lc * SEGL;
        // stage the reverb input of one super-step: thread t handles flat indices t, t + nthr, ... of [S] (l, r summed)
        constexpr int kPerT = (kRevMaxS + 32 * WPC * 2 - 1) / (32 * WPC * 2);  // sized for the smallest group (2 combs)
        float pl[kPerT], pr[kPerT];
        auto fetch_in = [&](int64_t base) {
#pragma unroll
            for (int k = 0; k < kPerT; ++k) {
                const int i = tid + k * nthr;
                const int64_t n = base + i;
                const bool ok = i < S && n < L;
                pl[k] = ok ? __ldcg(xl + n) : 0.0f;
                pr[k] = ok ? __ldcg(xr + n) : 0.0f;
            }
        };
        auto park_in = [&](int b) {
            float *dst = inbuf + b * kRevMaxS;
#pragma unroll
            for (int k = 0; k < kPerT; ++k) {
                const int i = tid + k * nthr;
                if (i < S) dst[i] = __fmul_rn(__fadd_rn(pl[k], pr[k]), 0.015f);
            }
        };
        ring_bar();    // (B) of the prologue: the stager's thread 0 has acquired the input up to 3 S (streaming pair)
        fetch_in(0);
        park_in(0);
        fetch_in(S);   // registers: the input of super-step 1, parked at the start of super-step 0
        if (lc == 0) { carry[cw] = 0.0f; carry[kRsMaxCw + cw] = 0.0f; }
        comb_bar();
        int wbase = 0;
        for (int64_t k = 0; k < nsteps; ++k, wbase = (wbase == 2 * S) ? 0 : wbase + S) {
            const int par = (int)(k & 1);
            // input pipeline: the registers hold super-step k + 1 (loaded a whole super-step ago: an L2 round trip no longer
            // shows up as a stall); park it -- inbuf[par ^ 1] was last read before barrier (B) of super-step k - 1 -- and
            // start the loads of super-step k + 2 (the stager acquired them before that barrier)
            park_in(par ^ 1);
            fetch_in((k + 2) * S);
            const float *inb = inbuf + par * kRevMaxS;
            int rb = wbase - my_delay;
            if (rb < 0) rb += RL;
            const float *rp = my_ring + rb + i0;
            float *wp = my_ring + wbase + i0;
            float o[SEGL];
#pragma unroll
            for (int i = 0; i < SEGL; ++i) o[i] = rp[i];
            float z = 0.0f;  // zero-state response of the damping one-pole over my segment
#pragma unroll
            for (int i = 0; i < SEGL; ++i) z = undenorm(__fadd_rn(__fmul_rn(o[i], keep), __fmul_rn(z, q.damp)));
            float A = dpow, Bv = z;  // affine map of my segment: s -> A * s + Bv; inclusive scan over the warp
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float Ap = __shfl_up_sync(0xffffffffu, A, d);
                const float Bp = __shfl_up_sync(0xffffffffu, Bv, d);
                if (lane >= d) { Bv = fmaf(Bp, A, Bv); A = A * Ap; }
            }
            float2 *wt = reinterpret_cast<float2 *>(wtot) + (par * kRsMaxCw + cw) * WPC;
            if (lane == 31) wt[ww] = make_float2(A, Bv);
            comb_bar();                 // (A) warp totals + last step's carry visible
            float s_in = carry[par * kRsMaxCw + cw];  // state entering this super-step, then through the warps before mine
#pragma unroll
            for (int w2 = 0; w2 < WPC - 1; ++w2)
                if (w2 < ww) { const float2 t = wt[w2]; s_in = fmaf(t.x, s_in, t.y); }
            const float s_out = fmaf(A, s_in, Bv);
            float sv = __shfl_up_sync(0xffffffffu, s_out, 1);
            if (lane == 0) sv = s_in;
#pragma unroll
            for (int i = 0; i < SEGL; ++i) {
                sv = undenorm(__fadd_rn(__fmul_rn(o[i], keep), __fmul_rn(sv, q.damp)));
                const float tv = undenorm(__fadd_rn(inb[i0 + i], __fmul_rn(sv, q.fb)));
                wp[i] = tv;
                if (wbase == 0) wp[RL + i] = tv;  // keep the mirror of the ring head current
            }
            if (ww == WPC - 1 && lane == 31) carry[(par ^ 1) * kRsMaxCw + cw] = sv;
            ring_bar();                 // (B) ring writes of this super-step visible to the group and to the stagers
        }
    } else if (is_stager) {
        // ------------------------------------------------------------------ stager group (producer CTAs)
        // Delayed outputs for the all-pass chain of super-step m: for every comb the S samples are one contiguous run of its
        // ring (it continues into the mirror, never wraps), complete after barrier (B) of super-step m - 1 and not
        // overwritten before super-step m + 1 has passed (the delays exceed S).  While the comb warps run super-step m the
        // stagers add the combs' runs into a 16-byte aligned staging row and ONE lane hands the row to the bulk-copy engine,
        // which writes the home CTA's dly row through distributed shared memory and completes the transaction on the
        // home's full[slot] mbarrier -- no remote stores and no fences on the comb filters' critical path.  The staging row
        // is reused for super-step m + depth, i.e. after free[slot] said that the all-pass group has consumed this one.
        const int s = gq == 1 ? tid - 2 * 32 * WPC : tid - Cfg::kCombThreads;
        int dl[kRsMaxCw];
#pragma unroll
        for (int j = 0; j < kRsMaxCw; ++j) dl[j] = g.comb_delay[c][comb0 + (j < ncomb ? j : 0)];
        const uint32_t dly_home = map_to((uint32_t)__cvta_generic_to_shared(dly), home_rank) + (uint32_t)((gq - 1) * kRsDepth * kRevMaxS) * 4u;
        const uint32_t full_home = map_to(bar0, home_rank);
        auto ship_row = [&](int slot) {  // one lane, after the stagers' barrier that follows the staging writes + proxy fences
            const float *srow = stage + slot * kRevMaxS;
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dly_home + (uint32_t)(slot * kRevMaxS) * 4u), "r"((uint32_t)__cvta_generic_to_shared(srow)),
                           "r"((uint32_t)(S * 4)), "r"(full_home + 8u * slot) : "memory");
        };
        // row 0: nothing has been written to the rings yet -- all zeros
        for (int i = s; i < S; i += nst) stage[i] = 0.0f;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        stager_bar();
        if (s == 0) {
            // streaming pair: acquire BEFORE the first row leaves -- the mixers take "row m has arrived" as proof that the
            // input up to (m + 3) S is published (their dry loads have no flag polling of their own)
            if (ready != nullptr) need_input(3 * (int64_t)S);
            ship_row(0);
        }
        ring_bar();  // (B) of the prologue
        int slot = 0, wb = 0;
        uint32_t free_par = 1;  // the first wait (row kRsDepth) is for phase 0
        for (int64_t k = 0; k < nsteps; ++k) {
            if (s == 0 && ready != nullptr) need_input((k + 4) * (int64_t)S);  // what the comb warps prefetch after this barrier
            ring_bar();  // (B) of super-step k: its ring writes are visible
            if (++slot == kRsDepth) { slot = 0; free_par ^= 1u; }
            wb = (wb == 2 * S) ? 0 : wb + S;
            if (k + 1 >= nsteps) break;
            // row k + 1 -> slot (k + 1) % depth, ring position ((k + 1) % 3) * S; the slot is free once the home's all-pass group
            // has finished super-step k + 1 - depth (control only: the bulk copy out of the staging row is then long complete)
            if (k + 1 >= kRsDepth) mbar_wait<kControl>(bar0 + 8u * (kRsDepth + slot), free_par);
            float *srow = stage + slot * kRevMaxS;
            for (int v4 = s; v4 < S / 4; v4 += nst) {
                float4 acc;
#pragma unroll
                for (int cq = 0; cq < kRsMaxCw; ++cq) {
                    if (cq < ncomb) {
                        int rb = wb - dl[cq];
                        if (rb < 0) rb += RL;
                        const float *src = ring + cq * kCombRing + rb + 4 * v4;
                        if (cq == 0) acc = make_float4(src[0], src[1], src[2], src[3]);
                        else { acc.x = __fadd_rn(acc.x, src[0]); acc.y = __fadd_rn(acc.y, src[1]); acc.z = __fadd_rn(acc.z, src[2]); acc.w = __fadd_rn(acc.w, src[3]); }
                    }
                }
                *reinterpret_cast<float4 *>(srow + 4 * v4) = acc;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my staging writes -> visible to the async proxy
            stager_bar();
            if (s == 0) ship_row(slot);
        }
    } else if (home && !comb_slot) {
        // ------------------------------------------------------------------ all-pass group of the home CTA (consumer)
        const int a = tid - Cfg::kCombThreads;
        int ad[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ad[j] = g.ap_delay[c][j];
        const uint32_t mix_own = map_to((uint32_t)__cvta_generic_to_shared(dly), home_rank + 1);        // wet[.][0] of my channel's mixer
        const uint32_t mix_peer = map_to((uint32_t)__cvta_generic_to_shared(dly), peer_rank + 1);       // wet[.][1] of the other channel's
        const uint32_t mixfull_own = map_to(bar0 + 8u * (2 * kRsDepth + kRsWet), home_rank + 1);
        const uint32_t mixfull_peer = map_to(bar0 + 8u * (2 * kRsDepth + kRsWet), peer_rank + 1);
        uint32_t free_of[kRsProd];
#pragma unroll
        for (int gg = 0; gg < kRsProd; ++gg) free_of[gg] = map_to(bar0 + 8u * kRsDepth, home_rank + 1 + gg);
        auto ap_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kRevSub) : "memory"); };
        int slot = 0, ws = 0;
        uint32_t full_par = 0, wfree_par = 1;  // wfree: the first wait (k = kRsWet) is for phase 0
        int nbase = 0;
        for (int64_t k = 0; k < nsteps; ++k) {
            // the staging row of the wet samples is free once BOTH mixers have consumed super-step k - kRsWet (which also
            // means the two bulk copies out of it are complete); control only, and long since true in steady state
            if (k >= kRsWet && ws == 0) wfree_par ^= 1u;
            if (k >= kRsWet) mbar_wait<kControl>(bar0 + 8u * (2 * kRsDepth + ws), wfree_par);
            mbar_wait<kTx>(bar0 + 8u * slot, full_par);  // the kRsProd delayed rows of this super-step
            const float *row = dly + slot * kRevMaxS;
            float *wown = wetb + ws * kRevMaxS;
            for (int sb = 0; sb < nsub; ++sb) {
                const int off = sb * kRevSub + a;
                if (off < S) {
                    const int n = nbase + off;
                    // all 7 shared-memory loads up front: the four all-pass rings are distinct arrays and each stage reads
                    // >= 244 samples behind what this sub-block writes, but the compiler cannot know and would serialise every
                    // stage's load behind the previous stage's store (4 x ~30 cycles of latency on the critical chain)
                    float cr[kRsProd], bvv[4];
#pragma unroll
                    for (int j = 0; j < kRsProd; ++j) cr[j] = row[j * kRsDepth * kRevMaxS + off];
#pragma unroll
                    for (int s2 = 0; s2 < 4; ++s2) bvv[s2] = ap[s2 * kApRing + ((n - ad[s2]) & (kApRing - 1))];
                    float v = cr[0];  // the producers pre-add their combs
#pragma unroll
                    for (int j = 1; j < kRsProd; ++j) v = __fadd_rn(v, cr[j]);
#pragma unroll
                    for (int s2 = 0; s2 < 4; ++s2) {
                        const float tv = undenorm(__fadd_rn(v, __fmul_rn(bvv[s2], 0.5f)));
                        ap[s2 * kApRing + (n & (kApRing - 1))] = tv;
                        v = __fsub_rn(bvv[s2], v);
                    }
                    wown[off] = v;
                }
                if (sb == nsub - 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // wet samples -> visible to the bulk copies
                ap_bar();
            }
            // Every all-pass thread has passed the barrier above: the wet row of this super-step is complete and dly[.][slot]
            // has been read.  ONE thread hands the wet row to the mixer of this channel (as its "own" row) and to the mixer of
            // the other channel (as its "peer" row) and re-arms the slot; kRsProd threads tell one producer CTA each that the
            // slot (and its staging row) is free.
            if (a == 0) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(wown);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(mix_own + (uint32_t)((ws * 2 + 0) * kRevMaxS) * 4u), "r"(src), "r"((uint32_t)(S * 4)), "r"(mixfull_own + 8u * ws) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(mix_peer + (uint32_t)((ws * 2 + 1) * kRevMaxS) * 4u), "r"(src), "r"((uint32_t)(S * 4)), "r"(mixfull_peer + 8u * ws) : "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * slot), "r"((uint32_t)(kRsProd * S * 4)) : "memory");
            }
            __syncwarp();
            // relaxed: what has to be ordered -- every thread's READS of dly[.][slot] -- completed before the barrier above
            if (a < kRsProd)
                asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(free_of[a] + 8u * slot) : "memory");
            if (++slot == kRsDepth) { slot = 0; full_par ^= 1u; }
            if (++ws == kRsWet) ws = 0;
            nbase = (nbase + S) & (kApRing * 1024 - 1);
        }
    } else if (gq == 1 && !comb_slot) {
        // ------------------------------------------------------------------ mixer group (CTA g == 1 of each channel)
        // y = wet_own * wet1 + wet_peer * wet2 + x * dry.  The two wet rows of a super-step arrive by bulk copy from the two
        // home CTAs (complete_tx on mixfull[ws]); the dry samples are re-read from global memory (L2 hits), prefetched one
        // super-step ahead -- in the streaming pair that is safe: when mixfull of super-step m completes, every producer
        // had acquired the input up to (m + 3) S before it staged its row m.  This CTA is not a home, so its dly region is
        // free: wet[kRsWet][2][kRevMaxS] lives there.
        const int a = tid - Cfg::kCombThreads;
        const float *wet = dly;
        float *dst = out + ((int64_t)p * 2 + c) * L;
        const float *xc = c ? xr : xl;
        float pk = 0.0f;
        const uint32_t mixfull0 = bar0 + 8u * (2 * kRsDepth + kRsWet);
        const uint32_t wfree_own = map_to(bar0 + 8u * (2 * kRsDepth), home_rank);
        const uint32_t wfree_peer = map_to(bar0 + 8u * (2 * kRsDepth), peer_rank);
        auto mx_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kRevSub) : "memory"); };
        constexpr int kPerA = (kRevMaxS + kRevSub - 1) / kRevSub;
        auto load_dry = [&](int64_t m, float (&xd)[kPerA]) {
            const int64_t m0 = m * S;
#pragma unroll
            for (int t = 0; t < kPerA; ++t) {
                const int i = a + t * kRevSub;
                xd[t] = (i < S && m0 + i < L) ? __ldcg(xc + m0 + i) : 0.0f;
            }
        };
        float xd[kPerA], xn[kPerA];
        int ws = 0;
        uint32_t par = 0;
        for (int64_t m = 0; m < nsteps; ++m) {
            mbar_wait<kTx>(mixfull0 + 8u * ws, par);
            if (m == 0) load_dry(0, xd);
            load_dry(m + 1, xn);
            const float *wo = wet + ws * 2 * kRevMaxS, *wpeer = wo + kRevMaxS;
            const int64_t m0 = m * S;
            const int cnt = (int)min((int64_t)S, L - m0);
#pragma unroll
            for (int t = 0; t < kPerA; ++t) {
                const int i = a + t * kRevSub;
                if (i < cnt) {
                    const float y = __fadd_rn(__fadd_rn(__fmul_rn(wo[i], q.wet1), __fmul_rn(wpeer[i], q.wet2)), __fmul_rn(xd[t], q.dry));
                    dst[m0 + i] = y;
                    pk = fmaxf(pk, fabsf(y));
                }
            }
#pragma unroll
            for (int t = 0; t < kPerA; ++t) xd[t] = xn[t];
            mx_bar();  // every mixer thread has read wet[ws]
            if (a == 0) {  // re-arm the slot, then return the credit to both homes (relaxed: control only, see above)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mixfull0 + 8u * ws), "r"((uint32_t)(2 * S * 4)) : "memory");
                asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(wfree_own + 8u * ws) : "memory");
                asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(wfree_peer + 8u * ws) : "memory");
            }
            if (++ws == kRsWet) { ws = 0; par ^= 1u; }
        }
        if (out_peak != nullptr) {
            pk = warp_max(pk);
            if (lane == 0) atomic_peak(out_peak, p, pk);
        }
    }
    cluster_barrier();  // nobody leaves while a peer may still store into its shared memory or arrive on its barriers
}

// ------------------------------------------------------------------- copy / peak
__global__ void __launch_bounds__(256) copy_kernel(SigView in, const float *in_peak, float *out, int chs,
                                                   int64_t L, unsigned *out_peak) {
    const int stream = blockIdx.y;
    const int p = stream / chs, c = stream - p * chs;
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? clip_peak(in_peak, p) : 1.0f;
    float pk = 0.0f;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < L; n += (int64_t)gridDim.x * blockDim.x) {
        float x = load_in(in, p, c, n);
        if (has_div) x = x / div;
        if (out != nullptr) out[(int64_t)stream * L + n] = x;
        pk = fmaxf(pk, fabsf(x));
    }
    if (out_peak != nullptr) {
        pk = warp_max(pk);
        if ((threadIdx.x & 31) == 0) atomic_peak(out_peak, p, pk);
    }
}

__global__ void __launch_bounds__(256) normalize_kernel(const float *x, const unsigned *peak, float *y,
                                                        int chs, int64_t L) {
    const int stream = blockIdx.y;
    const int p = stream / chs;
    const float div = fmaxf(__uint_as_float(peak[p]), 1e-8f);
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < L; n += (int64_t)gridDim.x * blockDim.x)
        y[(int64_t)stream * L + n] = x[(int64_t)stream * L + n] / div;
}

inline int grid_for(int64_t L, int threads, int streams) {
    // enough CTAs to fill 148 SMs a few times over without exceeding the work
    int64_t want = (L + threads - 1) / threads;
    int64_t cap = (148 * 16 + streams - 1) / streams;
    if (cap < 1) cap = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace

size_t eq_scratch_doubles(int P, int chs, int64_t L) {
    const int64_t K = (L + kEqChunk - 1) / kEqChunk;
    return (size_t)P * chs * kEqStates * K;
}

cudaError_t launch_eq(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                      int64_t L, const double *coefs, double *scratch_f, double *scratch_s,
                      unsigned *out_peak, int *launches) {
    const int K = (int)((L + kEqChunk - 1) / kEqChunk);
    const int streams = P * chs;
    dim3 grid((K + 127) / 128, streams);
    eq_chunk_kernel<false><<<grid, 128, 0, st>>>(in, in_peak, nullptr, chs, L, K, coefs, scratch_f, nullptr);
    {
        constexpr int smem = kEqStates * kStitchTile * (int)sizeof(double);
        {
            cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void *>(&eq_stitch_kernel), smem);
            if (e != cudaSuccess) return e;
        }
        eq_stitch_kernel<<<streams, kStitchThreads, smem, st>>>(chs, K, coefs, scratch_f, scratch_s);
    }
    eq_chunk_kernel<true><<<grid, 128, 0, st>>>(in, in_peak, out, chs, L, K, coefs, scratch_s, out_peak);
    *launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_compressor(cudaStream_t st, SigView in, const float *in_peak, float *out, int P,
                              int chs, int64_t L, const CompParams *prm, unsigned *out_peak, int *noconv,
                              int *ready, int *launches) {
    {
        cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void *>(&compressor_scan_kernel), (int)kCsSmem);
        if (e != cudaSuccess) return e;
    }
    static_assert(kCsSB == (1 << kGranuleShift), "the hand-off granule is the compressor super-block");
    compressor_scan_kernel<<<P * chs, kCsT, kCsSmem, st>>>(in, in_peak, out, chs, L, prm, out_peak, noconv, ready);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_distortion(cudaStream_t st, SigView in, const float *in_peak, float *out, int P,
                              int chs, int64_t L, const DistParams *prm, unsigned *out_peak,
                              int *launches) {
    dim3 grid(grid_for(L, 256, P * chs), P * chs);
    distortion_kernel<<<grid, 256, 0, st>>>(in, in_peak, out, chs, L, prm, out_peak);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_delay(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                         int64_t L, const DelayParams *prm, int max_d, unsigned *out_peak,
                         int *launches) {
    dim3 grid((max_d + 255) / 256, P * chs);
    delay_kernel<<<grid, 256, 0, st>>>(in, in_peak, out, chs, L, prm, out_peak);
    *launches += 1;
    return cudaGetLastError();
}

void reverb_geometry(double sample_rate, ReverbGeom *g) {
    static const int comb_t[8] = {1116, 1188, 1277, 1356, 1422, 1491, 1557, 1617};
    static const int ap_t[4] = {556, 441, 341, 225};
    const int isr = (int)sample_rate;
    int off = 0, min_ap = 1 << 30;
    for (int c = 0; c < 2; ++c) {
        for (int j = 0; j < 8; ++j) {
            g->comb_size[c][j] = (int)(((int64_t)isr * (comb_t[j] + (c ? 23 : 0))) / 44100);
            g->comb_off[c][j] = off;
            off += g->comb_size[c][j];
        }
        for (int j = 0; j < 4; ++j) {
            g->ap_size[c][j] = (int)(((int64_t)isr * (ap_t[j] + (c ? 23 : 0))) / 44100);
            g->ap_off[c][j] = off;
            off += g->ap_size[c][j];
            if (g->ap_size[c][j] < min_ap) min_ap = g->ap_size[c][j];
        }
    }
    g->total = off;
    int b = (min_ap / 32) * 32;
    if (b > 256) b = 256;
    g->block = b;
}

bool reverb_can_stream(const ReverbGeom &g) {  // the fast path (reverb_core_kernel) is the one that can consume flags
    int max_comb = 0, max_ap = 0, min_ap = 1 << 30, min_comb = 1 << 30;
    for (int c = 0; c < 2; ++c) {
        for (int j = 0; j < 8; ++j) { max_comb = max(max_comb, g.comb_size[c][j]); min_comb = min(min_comb, g.comb_size[c][j]); }
        for (int j = 0; j < 4; ++j) { max_ap = max(max_ap, g.ap_size[c][j]); min_ap = min(min_ap, g.ap_size[c][j]); }
    }
    const int seg = min_comb >= 32 * kRevMaxSegF ? kRevMaxSegF : 32;
    return g.block >= 32 && min_comb >= 32 * seg && max_comb <= 2 * 32 * seg && max_ap + kRevSub <= kApRing && min_ap >= kRevSub;
}

cudaError_t launch_reverb(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                          int stereo, int64_t L, const ReverbGeom &g, const ReverbParams *prm,
                          unsigned *out_peak, const int *ready, int sm_budget, int *launches) {
    if (g.block < 32) return cudaErrorInvalidValue;  // sample rate too low for the block scheme
    cudaError_t e;
    // fast path: one CTA per (candidate, channel), L / R CTAs paired in a cluster (reverb_core_kernel)
    int max_comb = 0, max_ap = 0, min_ap = 1 << 30;
    for (int c = 0; c < 2; ++c) {
        for (int j = 0; j < 8; ++j) max_comb = g.comb_size[c][j] > max_comb ? g.comb_size[c][j] : max_comb;
        for (int j = 0; j < 4; ++j) {
            max_ap = g.ap_size[c][j] > max_ap ? g.ap_size[c][j] : max_ap;
            min_ap = g.ap_size[c][j] < min_ap ? g.ap_size[c][j] : min_ap;
        }
    }
    int min_comb = 1 << 30;
    for (int c = 0; c < 2; ++c)
        for (int j = 0; j < 8; ++j) min_comb = g.comb_size[c][j] < min_comb ? g.comb_size[c][j] : min_comb;
    const int seg = min_comb >= 32 * kRevMaxSegF ? kRevMaxSegF : 32;
    if (min_comb >= 32 * seg && max_comb <= 2 * 32 * seg &&
        max_ap + kRevSub <= kApRing && min_ap >= kRevSub) {
        ReverbFastGeom fg;
        for (int c = 0; c < 2; ++c) {
            for (int j = 0; j < 8; ++j) fg.comb_delay[c][j] = g.comb_size[c][j];
            for (int j = 0; j < 4; ++j) fg.ap_delay[c][j] = g.ap_size[c][j];
        }
        const bool pair = stereo != 0;
        // small populations: one candidate = a cluster of 8 CTAs (reverb_split_kernel) when that many SMs are free
        static const bool split_on = !(getenv("STITO_REVERB_SPLIT") && atoi(getenv("STITO_REVERB_SPLIT")) == 0);
        if (split_on && pair && in_peak == nullptr && P * 8 <= sm_budget) {
            using KernS = void (*)(SigView, float *, int64_t, ReverbFastGeom, const ReverbParams *, unsigned *, const int *);
            constexpr int kWpcA = 5, kWpcB = 4;  // 7 samples x 160 lanes = 1120; 8 x 128 = 1024
            static_assert(32 * kWpcA * 7 == 32 * kRevMaxSegF && 32 * kWpcB * 8 == 32 * 32, "super-step lengths");
            const bool big = seg == kRevMaxSegF;
            KernS ks = big ? reverb_split_kernel<7, kWpcA> : reverb_split_kernel<8, kWpcB>;
            const size_t ssm = big ? RsCfg<kWpcA>::kSmem : RsCfg<kWpcB>::kSmem;
            e = ensure_dyn_smem(reinterpret_cast<const void *>(ks), (int)ssm);
            if (e != cudaSuccess) return e;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(P * 8);
            cfg.blockDim = dim3(big ? RsCfg<kWpcA>::kThreads : RsCfg<kWpcB>::kThreads);
            cfg.dynamicSmemBytes = ssm;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 8;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            e = cudaLaunchKernelEx(&cfg, ks, in, out, L, fg, prm, out_peak, ready);
            if (e != cudaSuccess) return e;
            *launches += 1;
            return cudaGetLastError();
        }
        const size_t smem = (size_t)(8 * kCombRing + 4 * kApRing + 10 * kRevMaxS) * sizeof(float) + 16;
        using Kern = void (*)(SigView, const float *, float *, int, int64_t, ReverbFastGeom, const ReverbParams *, unsigned *,
                              const int *);
        Kern kern = pair ? (seg == kRevMaxSegF ? reverb_core_kernel<kRevMaxSegF, true> : reverb_core_kernel<32, true>)
                         : (seg == kRevMaxSegF ? reverb_core_kernel<kRevMaxSegF, false> : reverb_core_kernel<32, false>);
        e = ensure_dyn_smem(reinterpret_cast<const void *>(kern), (int)smem);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(P * chs);
        cfg.blockDim = dim3(kRevThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;  // the L / R CTAs of a candidate exchange their wet signals
        attr[0].val.clusterDim.x = pair ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, in, in_peak, out, chs, L, fg, prm, out_peak, ready);
        if (e != cudaSuccess) return e;
        *launches += 1;
        return cudaGetLastError();
    }
    if (ready != nullptr) return cudaErrorInvalidValue;  // only the fast path streams (callers check reverb_can_stream)
    const size_t smem = (size_t)(g.total + g.block) * sizeof(float);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    if (stereo) {
        e = ensure_dyn_smem(reinterpret_cast<const void *>(&reverb_kernel<2>), (int)smem);
        if (e != cudaSuccess) return e;
        reverb_kernel<2><<<P, 512, smem, st>>>(in, in_peak, out, 2, L, g, prm, out_peak);
    } else {
        e = ensure_dyn_smem(reinterpret_cast<const void *>(&reverb_kernel<1>), (int)smem);
        if (e != cudaSuccess) return e;
        reverb_kernel<1><<<P * chs, 256, smem, st>>>(in, in_peak, out, chs, L, g, prm, out_peak);
    }
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_copy(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                        int64_t L, unsigned *out_peak, int *launches) {
    dim3 grid(grid_for(L, 256, P * chs), P * chs);
    copy_kernel<<<grid, 256, 0, st>>>(in, in_peak, out, chs, L, out_peak);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_peak(cudaStream_t st, SigView in, int P, int chs, int64_t L, unsigned *peak,
                        int *launches) {
    return launch_copy(st, in, nullptr, nullptr, P, chs, L, peak, launches);
}

cudaError_t launch_normalize(cudaStream_t st, const float *x, const unsigned *peak, float *y, int P,
                             int chs, int64_t L, int *launches) {
    dim3 grid(grid_for(L, 256, P * chs), P * chs);
    normalize_kernel<<<grid, 256, 0, st>>>(x, peak, y, chs, L);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace stito
