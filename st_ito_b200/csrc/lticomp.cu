// Feed-forward compressor with LTI gain smoothing (STITO_FX_LTI_COMPRESSOR): the variant the reference reaches through
// apply_compressor (st_ito/effects.py:623-648) / apply_random_compressor (st_ito/dsp.py:49-78), i.e. dasp-pytorch's
// `compressor` -- summed side-chain, soft-knee gain computer in dB, ONE-pole smoothing with the attack constant only,
// make-up gain, look-ahead by delaying the input.  CPU restatement: oracle/lticomp.py.
//
// The smoothing y[n] = alpha y[n-1] + (1 - alpha) g_c[n] is linear, so time is cut into chunks of 4096 samples that are
// processed in parallel (HBM-bound: the side-chain is read twice, the output written once):
//   pass 1  every CTA folds its chunk into one affine map s -> A s + B (A = alpha^valid, B = zero-state response);
//   pass 2  every CTA combines the maps of the chunks before it (<= a few hundred terms, block reduction) into its
//           entering state, adds the wrap-around term of the reference's frequency-sampled filter (below), re-computes the
//           gain computer for its samples and writes x[n - lookahead] * 10^((y[n] + makeup) / 20).
// Inside a chunk a warp owns 16 rows of 32 consecutive samples (coalesced loads); a row is an inclusive warp scan of affine
// maps in fp64, rows are chained through lane 31.
//
// Wrap-around: the reference filters by frequency sampling (dasp_pytorch.signal.lfilter_via_fsm, n_fft =
// 2^ceil(log2(2L - 1))), which is the circular convolution with the n_fft-periodic impulse response = the same recursion
// started from y[-1] = y0[L-1] alpha^(n_fft - L) / (1 - alpha^n_fft).  Zero for the ES configurations, not for short clips
// with a long attack, so it is reproduced.
#include <cmath>

#include "stito_internal.h"

namespace stito {

namespace {

constexpr int kLtiThreads = 256;
constexpr int kLtiRows = 16;                          // rows of 32 samples per warp
constexpr int kLtiChunk = kLtiThreads * kLtiRows;     // 4096 samples per CTA
constexpr int kLtiWarpSpan = 32 * kLtiRows;           // 512 consecutive samples per warp

struct Aff { double a, b; };  // s -> a * s + b
__device__ __forceinline__ Aff then(const Aff &first, const Aff &second) {
    return {second.a * first.a, fma(second.a, first.b, second.b)};
}
__device__ __forceinline__ Aff shfl_up(const Aff &v, int d) {
    return {__shfl_up_sync(0xffffffffu, v.a, d), __shfl_up_sync(0xffffffffu, v.b, d)};
}

// gain computer (oracle/lticomp.py gain_computer_db; float32, the reference's operation order)
__device__ __forceinline__ float gain_db(float side, const LtiCompParams &q) {
    const float x_db = __fmul_rn(20.0f, log10f(fmaxf(fabsf(side), 1e-8f)));
    const float half = __fdiv_rn(q.knee, 2.0f);
    float x_sc = x_db;
    if (x_db > __fadd_rn(q.thr, half)) {
        x_sc = __fadd_rn(q.thr, __fdiv_rn(__fsub_rn(x_db, q.thr), q.ratio));
    } else if (x_db >= __fsub_rn(q.thr, half)) {
        const float u = __fadd_rn(__fsub_rn(x_db, q.thr), half);
        const float slope = __fsub_rn(__fdiv_rn(1.0f, q.ratio), 1.0f);
        x_sc = __fadd_rn(x_db, __fdiv_rn(__fmul_rn(slope, __fmul_rn(u, u)), __fmul_rn(2.0f, q.knee)));
    }
    return __fsub_rn(x_sc, x_db);
}

// side-chain sample n of stream (p, c0): sum over `link` channels, each divided by the candidate's input peak if given
__device__ __forceinline__ float side_at(const SigView &v, int p, int c0, int link, int64_t n, float div, bool has_div) {
    float s = __ldg(v.base + (int64_t)p * v.stride_p + (int64_t)c0 * v.stride_c + n);
    if (has_div) s = s / div;
    if (link == 2) {
        float r = __ldg(v.base + (int64_t)p * v.stride_p + (int64_t)(c0 + 1) * v.stride_c + n);
        if (has_div) r = r / div;
        s = __fadd_rn(s, r);
    }
    return s;
}

// Inclusive scan over the 32 lanes of a row: on return v maps the state entering the row to the state after my sample.
__device__ __forceinline__ Aff row_scan(Aff v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Aff u = shfl_up(v, d);
        if (lane >= d) v = then(u, v);
    }
    return v;
}

// grid (chunks, streams): streams = P * (chs / link); maps[stream][chunk]
__global__ void __launch_bounds__(kLtiThreads) lti_fold_kernel(SigView in, const float *in_peak, int chs, int link, int64_t L,
                                                               const LtiCompParams *prm, Aff *maps) {
    __shared__ Aff wmap[kLtiThreads / 32];
    const int groups = chs / link;
    const int stream = blockIdx.y, p = stream / groups, c0 = (stream % groups) * link;
    const LtiCompParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? fmaxf(in_peak[p], 1e-8f) : 1.0f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * kLtiChunk + (int64_t)warp * kLtiWarpSpan;
    Aff acc = {1.0, 0.0};  // my warp's 512 samples
    float side[kLtiRows];
#pragma unroll
    for (int r = 0; r < kLtiRows; ++r) {
        const int64_t n = base + r * 32 + lane;
        side[r] = n < L ? side_at(in, p, c0, link, n, div, has_div) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < kLtiRows; ++r) {
        const int64_t n = base + r * 32 + lane;
        Aff v = {1.0, 0.0};  // past the end: identity, so that the last chunk's map ends exactly at sample L - 1
        if (n < L) v = {q.alpha, q.b0 * (double)gain_db(side[r], q)};
        v = row_scan(v, lane);
        const Aff row = {__shfl_sync(0xffffffffu, v.a, 31), __shfl_sync(0xffffffffu, v.b, 31)};
        acc = then(acc, row);
    }
    if (lane == 0) wmap[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        Aff m = wmap[0];
        for (int w = 1; w < kLtiThreads / 32; ++w) m = then(m, wmap[w]);
        maps[(int64_t)stream * gridDim.x + blockIdx.x] = m;
    }
}

__global__ void __launch_bounds__(kLtiThreads) lti_apply_kernel(SigView in, const float *in_peak, float *out, int chs, int link,
                                                                int64_t L, int lookahead, const LtiCompParams *prm,
                                                                const Aff *maps, unsigned *out_peak) {
    __shared__ Aff red[kLtiThreads / 32];
    __shared__ Aff wmap[kLtiThreads / 32];
    __shared__ double s_enter;
    const int groups = chs / link;
    const int stream = blockIdx.y, p = stream / groups, c0 = (stream % groups) * link;
    const int nchunks = gridDim.x, chunk = blockIdx.x;
    const LtiCompParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? fmaxf(in_peak[p], 1e-8f) : 1.0f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Aff *mp = maps + (int64_t)stream * nchunks;

    // zero-state value entering this chunk (x) and at the very end of the signal (y): sum_j B_j * prod_{i>j} A_i.  Every
    // chunk but the last is full, so prod A_i = alpha^(4096 * count) for the chunks strictly inside; the last chunk's own A
    // (alpha^valid) multiplies the terms of `y` only.
    {
        const Aff last = mp[nchunks - 1];
        double x = 0.0, y = 0.0;
        for (int j = threadIdx.x; j < nchunks; j += kLtiThreads) {
            const double bj = mp[j].b;
            if (j < chunk) x += bj * exp(q.ln_alpha * (double)((int64_t)(chunk - 1 - j) * kLtiChunk));
            if (j < nchunks - 1) y += bj * exp(q.ln_alpha * (double)((int64_t)(nchunks - 2 - j) * kLtiChunk)) * last.a;
            else y += bj;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x += __shfl_xor_sync(0xffffffffu, x, o);
            y += __shfl_xor_sync(0xffffffffu, y, o);
        }
        if (lane == 0) red[warp] = {x, y};
        __syncthreads();
        if (threadIdx.x == 0) {
            double sx = 0.0, sy = 0.0;
            for (int w = 0; w < kLtiThreads / 32; ++w) { sx += red[w].a; sy += red[w].b; }
            // periodic steady state of the frequency-sampled filter: y[-1] = y0[L-1] * wrap; it enters sample n0 as alpha^n0
            const double y_init = sy * q.wrap;
            s_enter = sx + (y_init != 0.0 ? y_init * exp(q.ln_alpha * (double)((int64_t)chunk * kLtiChunk)) : 0.0);
        }
    }

    // my warp's rows: gain computer, row scans kept in registers (the zero-state prefix inside the warp)
    const int64_t base = (int64_t)chunk * kLtiChunk + (int64_t)warp * kLtiWarpSpan;
    Aff pre[kLtiRows];  // map from the state entering my WARP's span to the state after my sample of row r
    Aff acc = {1.0, 0.0};
    {
        float side[kLtiRows];
#pragma unroll
        for (int r = 0; r < kLtiRows; ++r) {
            const int64_t n = base + r * 32 + lane;
            side[r] = n < L ? side_at(in, p, c0, link, n, div, has_div) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < kLtiRows; ++r) {
            const int64_t n = base + r * 32 + lane;
            Aff v = {1.0, 0.0};
            if (n < L) v = {q.alpha, q.b0 * (double)gain_db(side[r], q)};
            v = row_scan(v, lane);
            pre[r] = then(acc, v);
            const Aff row = {__shfl_sync(0xffffffffu, v.a, 31), __shfl_sync(0xffffffffu, v.b, 31)};
            acc = then(acc, row);
        }
    }
    if (lane == 0) wmap[warp] = acc;
    __syncthreads();
    double s = s_enter;  // state entering my warp's span
    for (int w = 0; w < warp; ++w) s = fma(wmap[w].a, s, wmap[w].b);

    float pk = 0.0f;
#pragma unroll
    for (int r = 0; r < kLtiRows; ++r) {
        const int64_t n = base + r * 32 + lane;
        if (n >= L) continue;
        const float g_s = (float)fma(pre[r].a, s, pre[r].b);
        const float g_lin = exp10f(__fdiv_rn(__fadd_rn(g_s, q.makeup), 20.0f));
        const int64_t m = n - lookahead;
        for (int c = 0; c < link; ++c) {
            float x = 0.0f;
            if (m >= 0) {
                x = __ldg(in.base + (int64_t)p * in.stride_p + (int64_t)(c0 + c) * in.stride_c + m);
                if (has_div) x = x / div;
            }
            const float y = __fmul_rn(x, g_lin);
            out[((int64_t)p * chs + c0 + c) * L + n] = y;
            pk = fmaxf(pk, fabsf(y));
        }
    }
    if (out_peak != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        if (lane == 0 && pk > 0.0f) atomicMax(out_peak + p, __float_as_uint(pk));
    }
}

}  // namespace

size_t lticomp_scratch_bytes(int P, int chs, int link, int64_t L) {
    const int64_t nchunks = (L + kLtiChunk - 1) / kLtiChunk;
    return (size_t)P * (chs / link) * nchunks * sizeof(Aff);
}

// Host: the per-candidate constants (oracle/lticomp.py attack_alpha, smooth_gain_recursive)
void lticomp_design(double sample_rate, int64_t L, float threshold_db, float ratio, float attack_ms, float knee_db,
                    float makeup_db, LtiCompParams *q) {
    const float nat = (float)sample_rate * (attack_ms / 1e3f);
    const float log9 = (float)std::log((double)9.0f);
    const float arg = -log9 / nat;
    const float alpha = (float)std::exp((double)arg);  // the correctly rounded float32 of the reference's float32 formula
    q->alpha = (double)alpha;
    q->b0 = (double)(1.0f - alpha);
    q->ln_alpha = alpha > 0.0f ? std::log((double)alpha) : -1.0e300;
    int64_t n_fft = 1;
    while (n_fft < 2 * L - 1) n_fft <<= 1;
    q->wrap = alpha > 0.0f ? std::exp(q->ln_alpha * (double)(n_fft - L)) / (-std::expm1(q->ln_alpha * (double)n_fft)) : 0.0;
    q->thr = threshold_db;
    q->ratio = ratio;
    q->knee = knee_db;
    q->makeup = makeup_db;
}

cudaError_t launch_lticomp(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs, int link, int64_t L,
                           int lookahead, const LtiCompParams *prm, void *scratch, unsigned *out_peak, int *launches) {
    if (P <= 0 || L <= 0) return cudaSuccess;
    const int64_t nchunks = (L + kLtiChunk - 1) / kLtiChunk;
    const dim3 grid((unsigned)nchunks, (unsigned)(P * (chs / link)));
    Aff *maps = reinterpret_cast<Aff *>(scratch);
    lti_fold_kernel<<<grid, kLtiThreads, 0, st>>>(in, in_peak, chs, link, L, prm, maps);
    lti_apply_kernel<<<grid, kLtiThreads, 0, st>>>(in, in_peak, out, chs, link, L, lookahead, prm, maps, out_peak);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace stito
