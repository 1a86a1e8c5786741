// Front-end kernel (K2): mid/side -> framing (centre, reflect pad) -> periodic-Hann window ->
// 2048-point FFT -> power spectrum -> Slaney mel (128) -> 10*log10 -> min-max normalisation.
// Replaces torchlibrosa's Spectrogram (two dense-DFT Conv1d) + LogmelFilterBank as used by the
// reference's Cnn14 (st_ito/models/panns.py:147-168, 219-245); restated on the CPU in
// oracle/frontend.py.
//
// One CTA per (item, frame).  Mid and side of a stereo item are transformed by ONE complex FFT
// (z = mid + i*side) and separated with the conjugate-symmetry identities, so a stereo frame costs
// a single 2048-point complex transform.  The FFT is a mixed-radix (4,4,4,4,4,2) Stockham autosort
// in shared memory; the final peak normalisation of process_audio (style_transfer.py:113) is applied
// to the samples as they are loaded, so the normalised waveform never has to exist in HBM.
#include "stito_internal.h"

namespace stito {

namespace {

constexpr int kNfft = 2048;
constexpr int kThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <int NS>
__device__ __forceinline__ void radix4_stage(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                             const float2 *__restrict__ tw, int tid) {
    constexpr int Q = kNfft / 4;
#pragma unroll
    for (int it = 0; it < Q / kThreads; ++it) {
        const int j = tid + it * kThreads;
        const int k = j & (NS - 1);
        float2 v0 = src[j], v1 = src[j + Q], v2 = src[j + 2 * Q], v3 = src[j + 3 * Q];
        if (NS > 1) {
            constexpr int step = kNfft / (NS * 4);
            v1 = cmul(v1, __ldg(tw + k * step));
            v2 = cmul(v2, __ldg(tw + 2 * k * step));
            v3 = cmul(v3, __ldg(tw + 3 * k * step));
        }
        const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
        const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));  // (v1 - v3) * (-i)
        const int d = ((j - k) << 2) + k;  // (j / NS) * NS * 4 + k
        dst[d] = make_float2(t0.x + t2.x, t0.y + t2.y);
        dst[d + NS] = make_float2(t1.x + t3.x, t1.y + t3.y);
        dst[d + 2 * NS] = make_float2(t0.x - t2.x, t0.y - t2.y);
        dst[d + 3 * NS] = make_float2(t1.x - t3.x, t1.y - t3.y);
    }
}

__device__ __forceinline__ void radix2_last_stage(const float2 *__restrict__ src, float2 *__restrict__ dst,
                                                  const float2 *__restrict__ tw, int tid) {
    constexpr int H = kNfft / 2;  // NS == H: k = j, d = j
#pragma unroll
    for (int it = 0; it < H / kThreads; ++it) {
        const int j = tid + it * kThreads;
        const float2 v0 = src[j];
        const float2 v1 = cmul(src[j + H], __ldg(tw + j));
        dst[j] = make_float2(v0.x + v1.x, v0.y + v1.y);
        dst[j + H] = make_float2(v0.x - v1.x, v0.y - v1.y);
    }
}

constexpr int kFramesPerCta = 1;  // frames per CTA (frame i+1 prefetched during frame i); 4 measured 24 % slower than 1 on B200

__global__ void __launch_bounds__(kThreads) logmel_kernel(SigView in, const unsigned *peak, int chs,
                                                          int64_t L, int T, FrontendTables tb,
                                                          float *feat) {
    __shared__ float2 bufA[kNfft];
    __shared__ float2 bufB[kNfft];
    const int tid = threadIdx.x;
    const int item = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int f1 = min(f0 + kFramesPerCta, T);
    const bool has_div = peak != nullptr;
    const float div = has_div ? fmaxf(__uint_as_float(peak[item]), 1e-8f) : 1.0f;
    const float *b0 = in.base + (int64_t)item * in.stride_p;
    const float *b1 = b0 + in.stride_c;
    constexpr int kPer = kNfft / kThreads;  // 8 samples per thread and channel
    float wl[kPer], wr[kPer];               // raw samples of the frame about to be transformed
    auto fetch = [&](int frame) {
        const int64_t s0 = (int64_t)frame * tb.hop - kNfft / 2;
#pragma unroll
        for (int it = 0; it < kPer; ++it) {
            int64_t s = s0 + tid + it * kThreads;
            if (s < 0) s = -s;                      // reflect (no edge repeat), torch pad_mode="reflect"
            if (s >= L) s = 2 * (L - 1) - s;
            wl[it] = __ldg(b0 + s);
            wr[it] = chs == 2 ? __ldg(b1 + s) : 0.0f;
        }
    };
    fetch(f0);
    for (int frame = f0; frame < f1; ++frame) {
#pragma unroll
        for (int it = 0; it < kPer; ++it) {
            const int n = tid + it * kThreads;
            const float w = __ldg(tb.window + n);
            float2 z;
            float l = wl[it], r = wr[it];
            if (chs == 2) {
                if (has_div) { l = l / div; r = r / div; }
                z.x = __fadd_rn(l, r) * 0.5f * w;  // mid  (panns.py:220)
                z.y = __fsub_rn(l, r) * 0.5f * w;  // side (panns.py:221)
            } else {
                if (has_div) l = l / div;
                z.x = l * w;
                z.y = 0.0f;
            }
            bufA[n] = z;
        }
        if (frame + 1 < f1) fetch(frame + 1);  // in flight while this frame is transformed
        __syncthreads();
        radix4_stage<1>(bufA, bufB, tb.twiddle, tid);   __syncthreads();
        radix4_stage<4>(bufB, bufA, tb.twiddle, tid);   __syncthreads();
        radix4_stage<16>(bufA, bufB, tb.twiddle, tid);  __syncthreads();
        radix4_stage<64>(bufB, bufA, tb.twiddle, tid);  __syncthreads();
        radix4_stage<256>(bufA, bufB, tb.twiddle, tid); __syncthreads();
        radix2_last_stage(bufB, bufA, tb.twiddle, tid); __syncthreads();

        // power spectra of the two real signals packed in z: M = (Z[k] + conj Z[N-k]) / 2,
        // S = (Z[k] - conj Z[N-k]) / (2i)
        float *pw_mid = reinterpret_cast<float *>(bufB);
        float *pw_side = pw_mid + (kNfft / 2 + 1);
        for (int k = tid; k <= kNfft / 2; k += kThreads) {
            const float2 a = bufA[k];
            const float2 b = bufA[(kNfft - k) & (kNfft - 1)];
            if (chs == 2) {
                const float mr = 0.5f * (a.x + b.x), mi = 0.5f * (a.y - b.y);
                const float sr = 0.5f * (a.y + b.y), si = 0.5f * (b.x - a.x);
                pw_mid[k] = mr * mr + mi * mi;
                pw_side[k] = sr * sr + si * si;
            } else {
                pw_mid[k] = a.x * a.x + a.y * a.y;
            }
        }
        __syncthreads();
        const int n_mels = tb.n_mels;
        for (int o = tid; o < n_mels * chs; o += kThreads) {
            const int q = o / n_mels, m = o - q * n_mels;
            const float *pw = q ? pw_side : pw_mid;
            const int st = tb.mel_start[m], cnt = tb.mel_count[m];
            const float *wt = tb.mel_wt + tb.mel_off[m];
            float acc = 0.0f;
            for (int i = 0; i < cnt; ++i) acc = fmaf(pw[st + i], __ldg(wt + i), acc);
            // LogmelFilterBank: 10*log10(clamp(mel, 1e-10)) - 10*log10(max(amin, ref=1)) (= 0)
            float db = 10.0f * log10f(fmaxf(acc, 1e-10f));
            // input_norm == "minmax" (panns.py:238-241)
            db = fminf(fmaxf(db, -80.0f), 40.0f);
            const float v = ((db + 80.0f) / 120.0f) * 2.0f - 1.0f;
            feat[(((int64_t)item * chs + q) * T + frame) * n_mels + m] = v;
        }
        __syncthreads();  // bufA / bufB are rewritten by the next frame
    }
}

}  // namespace

cudaError_t launch_logmel(cudaStream_t st, SigView in, const unsigned *peak, int B, int chs, int64_t L,
                          int T, const FrontendTables &tb, float *feat, int *launches) {
    if (tb.n_fft != kNfft || L < kNfft / 2 + 1) return cudaErrorInvalidValue;
    dim3 grid((T + kFramesPerCta - 1) / kFramesPerCta, B);
    logmel_kernel<<<grid, kThreads, 0, st>>>(in, peak, chs, L, T, tb, feat);
    *launches += 1;
    return cudaGetLastError();
}

}  // namespace stito
