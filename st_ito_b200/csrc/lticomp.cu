// Feed-forward compressor with LTI gain smoothing (STITO_FX_LTI_COMPRESSOR): the variant the reference reaches through
// apply_compressor (st_ito/effects.py:623-648) / apply_random_compressor (st_ito/dsp.py:49-78), i.e. dasp-pytorch's
// `compressor` -- summed side-chain, soft-knee gain computer in dB, ONE-pole smoothing with the attack constant only,
// make-up gain, look-ahead by delaying the input.  CPU restatement: oracle/lticomp.py.
//
// The smoothing y[n] = alpha y[n-1] + (1 - alpha) g_c[n] is linear, so time is cut into chunks of 4096 samples that are
// processed in parallel (HBM-bound: x is read twice, g_c written and read once, the output written once):
//   pass 1  every CTA computes the gain computer for its chunk (coalesced loads, float32), keeps g_c in shared memory and
//           in a scratch row, and folds the chunk into one affine map s -> A s + B (A = alpha^valid, B = zero-state response):
//           a thread owns 16 CONSECUTIVE samples (a 16-step fp64 recurrence out of a padded, conflict-free shared-memory
//           layout), the 256 thread maps are combined by an ordered shuffle reduction;
//   pass 2  every CTA combines the maps of the chunks before it (<= a few hundred terms, block reduction -- no serial
//           stitch kernel) into its entering state, adds the wrap-around term of the reference's frequency-sampled filter
//           (below), re-runs the 16-step recurrence per thread from the scanned thread states, turns the smoothed gain into
//           a linear factor (shared memory again) and writes x[n - lookahead] * 10^((y[n] + makeup) / 20) coalesced.
// (The first version scanned rows of 32 samples with warp shuffles -- 5 steps x 4 shuffles per sample row, 640 instructions
// per row in the end: 114 + 284 us for 16 stereo candidates, 4 - 8 % of the HBM roofline; profiles/r02_lticomp_*.)
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
constexpr int kLtiPer = 16;                           // consecutive samples per thread in the recurrence
constexpr int kLtiChunk = kLtiThreads * kLtiPer;      // 4096 samples per CTA
constexpr int kLtiPadded = kLtiChunk + kLtiChunk / 16;
__device__ __forceinline__ int lti_pad(int j) { return j + (j >> 4); }  // thread t reads 17 t + i: conflict-free

struct Aff { double a, b; };  // s -> a * s + b
__device__ __forceinline__ Aff then(const Aff &first, const Aff &second) {
    return {second.a * first.a, fma(second.a, first.b, second.b)};
}

// gain computer (oracle/lticomp.py gain_computer_db; float32, the reference's operation order -- its three divisions are
// kept as IEEE divisions: a reciprocal multiply would differ from torch by an ulp of a value of up to 160 dB)
__device__ __forceinline__ float gain_db(float side, const LtiCompParams &q) {
    const float x_db = __fmul_rn(20.0f, log10f(fmaxf(fabsf(side), 1e-8f)));
    const float half = __fmul_rn(q.knee, 0.5f);
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

// my 16 consecutive samples out of shared memory: zero-state fold (valid = how many of them lie below L)
__device__ __forceinline__ Aff fold16(const float *gs, int t, int valid, const LtiCompParams &q) {
    Aff m = {1.0, 0.0};
#pragma unroll
    for (int i = 0; i < kLtiPer; ++i)
        if (i < valid) { m.b = fma(q.alpha, m.b, q.b0 * (double)gs[17 * t + i]); m.a *= q.alpha; }
    return m;
}

// grid (chunks, streams): streams = P * (chs / link); maps[stream][chunk], G[stream][chunk * 4096 + j]
__global__ void __launch_bounds__(kLtiThreads) lti_fold_kernel(SigView in, const float *in_peak, int chs, int link, int64_t L,
                                                               const LtiCompParams *prm, Aff *maps, float *G) {
    __shared__ float gs[kLtiPadded];
    __shared__ Aff wmap[kLtiThreads / 32];
    const int groups = chs / link;
    const int stream = blockIdx.y, p = stream / groups, c0 = (stream % groups) * link;
    const LtiCompParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? fmaxf(in_peak[p], 1e-8f) : 1.0f;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t n0 = (int64_t)blockIdx.x * kLtiChunk;
    const float *x0 = in.base + (int64_t)p * in.stride_p + (int64_t)c0 * in.stride_c + n0;
    const float *x1 = x0 + in.stride_c;
    float *grow = G + ((int64_t)stream * gridDim.x + blockIdx.x) * kLtiChunk;
    float side[kLtiPer];
#pragma unroll
    for (int i = 0; i < kLtiPer; ++i) {  // coalesced: sample i * 256 + t; the side-chain is the SUM of the linked channels
        const int j = i * kLtiThreads + t;
        float v = 0.0f;
        if (n0 + j < L) {
            v = __ldg(x0 + j);
            if (has_div) v = v / div;
            if (link == 2) {
                float r = __ldg(x1 + j);
                if (has_div) r = r / div;
                v = __fadd_rn(v, r);
            }
        }
        side[i] = v;
    }
#pragma unroll
    for (int i = 0; i < kLtiPer; ++i) {
        const int j = i * kLtiThreads + t;
        const float gdb = gain_db(side[i], q);
        gs[lti_pad(j)] = gdb;
        grow[j] = gdb;  // (samples past L: harmless values, never used)
    }
    __syncthreads();
    const int64_t left = L - (n0 + (int64_t)t * kLtiPer);
    Aff v = fold16(gs, t, (int)max((int64_t)0, min((int64_t)kLtiPer, left)), q);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {  // ordered reduction: lane l ends up with the map of lanes l .. l + 2d - 1
        const Aff u = {__shfl_down_sync(0xffffffffu, v.a, d), __shfl_down_sync(0xffffffffu, v.b, d)};
        if (lane + d < 32) v = then(v, u);
    }
    if (lane == 0) wmap[warp] = v;
    __syncthreads();
    if (t == 0) {
        Aff m = wmap[0];
        for (int w = 1; w < kLtiThreads / 32; ++w) m = then(m, wmap[w]);
        maps[(int64_t)stream * gridDim.x + blockIdx.x] = m;
    }
}

__global__ void __launch_bounds__(kLtiThreads) lti_apply_kernel(SigView in, const float *in_peak, float *out, int chs, int link,
                                                                int64_t L, int lookahead, const LtiCompParams *prm,
                                                                const Aff *maps, const float *G, unsigned *out_peak) {
    __shared__ float gs[kLtiPadded];
    __shared__ Aff red[kLtiThreads / 32];
    __shared__ Aff wmap[kLtiThreads / 32];
    __shared__ double s_enter;
    const int groups = chs / link;
    const int stream = blockIdx.y, p = stream / groups, c0 = (stream % groups) * link;
    const int nchunks = gridDim.x, chunk = blockIdx.x;
    const LtiCompParams q = prm[p];
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? fmaxf(in_peak[p], 1e-8f) : 1.0f;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const Aff *mp = maps + (int64_t)stream * nchunks;
    const int64_t n0 = (int64_t)chunk * kLtiChunk;

    // my 16 consecutive g_c values (the scratch rows are 16-byte aligned) -> shared memory, padded layout
    {
        const float4 *grow = reinterpret_cast<const float4 *>(G + ((int64_t)stream * nchunks + chunk) * kLtiChunk) + 4 * t;
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
            const float4 g4 = __ldg(grow + v4);
            float *dst = gs + 17 * t + 4 * v4;
            dst[0] = g4.x; dst[1] = g4.y; dst[2] = g4.z; dst[3] = g4.w;
        }
    }
    // zero-state value entering this chunk (x) and at the very end of the signal (y): sum_j B_j * prod_{i>j} A_i.  Every
    // chunk but the last is full, so prod A_i = alpha^(4096 * count) for the chunks strictly inside; the last chunk's own A
    // (alpha^valid) multiplies the terms of `y` only.
    {
        const Aff last = mp[nchunks - 1];
        double x = 0.0, y = 0.0;
        for (int j = t; j < nchunks; j += kLtiThreads) {
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
    }
    // thread maps (zero state) and their exclusive scan over the block
    const int64_t left = L - (n0 + (int64_t)t * kLtiPer);
    const int valid = (int)max((int64_t)0, min((int64_t)kLtiPer, left));
    Aff inc = fold16(gs, t, valid, q);  // reads only my own 16 values: no barrier needed yet
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Aff u = {__shfl_up_sync(0xffffffffu, inc.a, d), __shfl_up_sync(0xffffffffu, inc.b, d)};
        if (lane >= d) inc = then(u, inc);
    }
    if (lane == 31) wmap[warp] = inc;
    Aff exc = {__shfl_up_sync(0xffffffffu, inc.a, 1), __shfl_up_sync(0xffffffffu, inc.b, 1)};
    if (lane == 0) exc = {1.0, 0.0};
    __syncthreads();
    if (t == 0) {
        double sx = 0.0, sy = 0.0;
        for (int w = 0; w < kLtiThreads / 32; ++w) { sx += red[w].a; sy += red[w].b; }
        // periodic steady state of the frequency-sampled filter: y[-1] = y0[L-1] * wrap; it enters sample n0 as alpha^n0
        const double y_init = sy * q.wrap;
        s_enter = sx + (y_init != 0.0 ? y_init * exp(q.ln_alpha * (double)n0) : 0.0);
    }
    __syncthreads();
    double s = s_enter;  // state entering my warp, then my thread
    for (int w = 0; w < warp; ++w) s = fma(wmap[w].a, s, wmap[w].b);
    s = fma(exc.a, s, exc.b);
    // the recurrence over my 16 samples; the linear gain factor replaces g_c in shared memory
#pragma unroll
    for (int i = 0; i < kLtiPer; ++i) {
        if (i < valid) {
            s = fma(q.alpha, s, q.b0 * (double)gs[17 * t + i]);
            gs[17 * t + i] = exp10f(__fdiv_rn(__fadd_rn((float)s, q.makeup), 20.0f));
        }
    }
    __syncthreads();
    // coalesced output: sample i * 256 + t of the chunk, every linked channel; the INPUT is delayed by the look-ahead
    float pk = 0.0f;
#pragma unroll 4
    for (int i = 0; i < kLtiPer; ++i) {
        const int j = i * kLtiThreads + t;
        const int64_t n = n0 + j;
        if (n >= L) break;
        const float g_lin = gs[lti_pad(j)];
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

static size_t lti_maps_bytes(int P, int chs, int link, int64_t L) {
    const int64_t nchunks = (L + kLtiChunk - 1) / kLtiChunk;
    return (((size_t)P * (chs / link) * nchunks * sizeof(Aff)) + 255) & ~(size_t)255;
}
size_t lticomp_scratch_bytes(int P, int chs, int link, int64_t L) {  // chunk maps + one row of g_c per stream
    const int64_t nchunks = (L + kLtiChunk - 1) / kLtiChunk;
    return lti_maps_bytes(P, chs, link, L) + (size_t)P * (chs / link) * nchunks * kLtiChunk * sizeof(float);
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
    float *G = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(scratch) + lti_maps_bytes(P, chs, link, L));
    lti_fold_kernel<<<grid, kLtiThreads, 0, st>>>(in, in_peak, chs, link, L, prm, maps, G);
    lti_apply_kernel<<<grid, kLtiThreads, 0, st>>>(in, in_peak, out, chs, link, L, lookahead, prm, maps, G, out_peak);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace stito
