// Noise-shaped convolution reverb (BASELINE config 4 "2 s-IR conv reverb"): the arithmetic the
// reference's apply_reverb (st_ito/effects.py:558-620, st_ito/dsp.py:26-46) obtains from
// dasp_pytorch.noise_shaped_reverberation, restated on the CPU in oracle/convreverb.py:
//     IR[ch][n] = 1/12 * sum_b gain_b * exp(-(10 decay_b + 1) * t_n) * band_b(noise_ch)[n],   t_n = n / (N_ir - 1)
//     y[ch]     = x[ch] (*) IR[ch]   (causal, truncated to the input length),   out = (1 - mix) x + mix y
// A 96 000-tap FIR cannot be sample-serial (96 k MAC per sample); it runs as a uniformly partitioned overlap-save
// convolution in the frequency domain.  B = 8192-sample partitions, 16384-point FFTs held entirely in shared memory:
//   setup (once per handle / (sample rate, N_ir, seed)): the candidate-INDEPENDENT part -- seeded white noise through the
//       12 octave-band FIR filters (1023 taps, scipy.signal.firwin restated on the host) -> bands[2][12][N_ir];
//   K1 crv_ir_fft    per (candidate, partition): build the partition of the candidate's stereo IR from its 12 gains /
//                    decays, FFT (L + iR packed in one complex transform), store the unpacked spectra HL, HR;
//   K2 crv_x_fft     per (candidate, block j): FFT of input samples [(j-1)B, (j+1)B), both channels packed;
//   K3 crv_mac       frequency-domain delay line: Y_j = sum_p H_p . X_{j-p}; a thread owns one bin pair (k, N-k) for all
//                    blocks, keeps H (K partitions) and the last K input spectra in registers, reads every X once and
//                    writes Y in place;
//   K4 crv_ifft_mix  inverse FFT of Y_j, keep the B valid samples, out = (1 - mix) x + mix y, peak tracking.
// HBM traffic per stereo candidate: 4 * 2L (read x) + 3 * 16 L (spectra written, read+written, read) + 2 * 8L ~ 72 L bytes
// against 16 L algorithmic (read x, write out): the spectra round trips are the price of the partitioned form.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "stito_internal.h"

namespace stito {

namespace {

constexpr int kCrvN = 16384;          // FFT length
constexpr int kCrvB = kCrvN / 2;      // partition / hop
constexpr int kCrvThreads = 1024;
constexpr int kCrvBands = 12;
constexpr int kCrvTaps = 1023;
constexpr size_t kCrvSmem = (size_t)(kCrvN + kCrvN / 4) * sizeof(float2);  // data + quarter-circle twiddles

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 16384-point complex FFT (forward, e^{-i...}) of buf[] in shared memory, natural order in and out: seven radix-4
// Stockham passes.  One buffer only (two would not fit): every pass loads its 16 inputs per thread into registers,
// synchronises, then writes its 16 outputs.  Twiddles: the three factors of a butterfly are w, w^2, w^3 with
// w = exp(-2 pi i k / (4 ns)), an angle in the first quadrant -> ONE read of a 4096-entry quarter-circle table that lives in
// shared memory next to the data (tws), the other two by complex multiplication.  (The first version read all three from the
// 128 KB global table: 84 L2-latency loads per thread and transform, 42-52 % long-scoreboard stalls in ncu, 46 us per FFT.)
constexpr int kCrvTw = kCrvN / 4;  // quarter-circle twiddle table entries

__device__ __forceinline__ void load_twiddles(float2 *tws, const float2 *__restrict__ tw, int tid) {
    for (int i = tid; i < kCrvTw; i += kCrvThreads) tws[i] = __ldg(tw + i);  // tw[i] = exp(-2 pi i * i / N), i < N / 4
}

__device__ __forceinline__ void fft16k(float2 *buf, const float2 *tws, int tid) {
    constexpr int Q = kCrvN / 4;
    constexpr int PER = Q / kCrvThreads;  // 4 butterflies per thread and pass
#pragma unroll 1
    for (int ns = 1; ns < kCrvN; ns <<= 2) {
        float2 v[PER][4];
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int j = tid + it * kCrvThreads;
            v[it][0] = buf[j]; v[it][1] = buf[j + Q]; v[it][2] = buf[j + 2 * Q]; v[it][3] = buf[j + 3 * Q];
        }
        __syncthreads();
        const int step = kCrvN / (ns * 4);
#pragma unroll
        for (int it = 0; it < PER; ++it) {
            const int j = tid + it * kCrvThreads;
            const int k = j & (ns - 1);
            float2 v0 = v[it][0], v1 = v[it][1], v2 = v[it][2], v3 = v[it][3];
            if (ns > 1) {
                const float2 w1 = tws[k * step];  // k * step < N / 4
                const float2 w2 = cmulf(w1, w1), w3 = cmulf(w2, w1);
                v1 = cmulf(v1, w1);
                v2 = cmulf(v2, w2);
                v3 = cmulf(v3, w3);
            }
            const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
            const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
            const float2 t3 = make_float2(v1.y - v3.y, -(v1.x - v3.x));  // (v1 - v3) * (-i)
            const int d = ((j - k) << 2) + k;
            buf[d] = make_float2(t0.x + t2.x, t0.y + t2.y);
            buf[d + ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
            buf[d + 2 * ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
            buf[d + 3 * ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
        }
        __syncthreads();
    }
}

// ---- setup: bands[ch][b][n] = sum_t wn[ch][b][n + t] * filt[b][t]   (conv1d = "valid" correlation; fp64 accumulate)
__global__ void __launch_bounds__(256) crv_bands_kernel(const float *__restrict__ wn, const float *__restrict__ filt,
                                                        float *__restrict__ bands, int n_ir) {
    __shared__ float fs[kCrvTaps + 1];
    const int cb = blockIdx.y;  // ch * 12 + band
    const int b = cb % kCrvBands;
    for (int i = threadIdx.x; i < kCrvTaps; i += blockDim.x) fs[i] = filt[b * kCrvTaps + i];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_ir) return;
    const float *src = wn + (size_t)cb * (n_ir + kCrvTaps - 1) + n;
    double acc = 0.0;
    for (int t = 0; t < kCrvTaps; ++t) acc = fma((double)__ldg(src + t), (double)fs[t], acc);
    bands[(size_t)cb * n_ir + n] = (float)acc;
}

// ---- K1: candidate IR partition -> spectrum.  H layout: [P][K][N/2 + 1] float4 = (HL.re, HL.im, HR.re, HR.im)
__global__ void __launch_bounds__(kCrvThreads, 1) crv_ir_fft_kernel(const float *__restrict__ bands, int n_ir, int K,
                                                                     const ConvRevParams *__restrict__ prm,
                                                                     const float2 *__restrict__ tw,
                                                                     float4 *__restrict__ H) {
    extern __shared__ float2 crv_buf[];
    float2 *tws = crv_buf + kCrvN;
    const int part = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
    load_twiddles(tws, tw, tid);
    const ConvRevParams q = prm[p];
    float dec[kCrvBands];
#pragma unroll
    for (int b = 0; b < kCrvBands; ++b) dec[b] = __fadd_rn(__fmul_rn(q.decay[b], 10.0f), 1.0f);  // band_decays * 10 + 1
    const double tden = (double)(n_ir - 1);
    for (int i = tid; i < kCrvN; i += kCrvThreads) {
        float2 z = make_float2(0.0f, 0.0f);
        const int n = part * kCrvB + i;
        if (i < kCrvB && n < n_ir) {
            const float t = (float)((double)n / tden);  // linspace(0, 1, n_ir)
            float al = 0.0f, ar = 0.0f;
#pragma unroll
            for (int b = 0; b < kCrvBands; ++b) {
                const float c = __fmul_rn(expf(-__fmul_rn(dec[b], t)), q.gain[b]);  // env * gain
                al = __fadd_rn(al, __fmul_rn(__ldg(bands + (size_t)b * n_ir + n), c));
                ar = __fadd_rn(ar, __fmul_rn(__ldg(bands + (size_t)(kCrvBands + b) * n_ir + n), c));
            }
            z = make_float2(al / (float)kCrvBands, ar / (float)kCrvBands);  // mean over the bands
        }
        crv_buf[i] = z;
    }
    __syncthreads();
    fft16k(crv_buf, tws, tid);
    float4 *dst = H + ((size_t)p * K + part) * (kCrvN / 2 + 1);
    for (int k = tid; k <= kCrvN / 2; k += kCrvThreads) {
        const float2 a = crv_buf[k], b = crv_buf[(kCrvN - k) & (kCrvN - 1)];
        // two real signals packed as l + i r:  L = (Z[k] + conj Z[N-k]) / 2,  R = (Z[k] - conj Z[N-k]) / (2i)
        dst[k] = make_float4(0.5f * (a.x + b.x), 0.5f * (a.y - b.y), 0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
    }
}

// ---- K2: input block j = samples [(j-1)B, (j+1)B) of both channels -> spectrum X[p][j][N] (packed l + i r)
__global__ void __launch_bounds__(kCrvThreads, 1) crv_x_fft_kernel(SigView in, const float *in_peak, int64_t L,
                                                                    int nblocks, const float2 *__restrict__ tw,
                                                                    float2 *__restrict__ X) {
    extern __shared__ float2 crv_buf[];
    float2 *tws = crv_buf + kCrvN;
    const int j = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
    load_twiddles(tws, tw, tid);
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? fmaxf(in_peak[p], 1e-8f) : 1.0f;
    const float *xl = in.base + (int64_t)p * in.stride_p;
    const float *xr = xl + in.stride_c;
    const int64_t n0 = (int64_t)(j - 1) * kCrvB;
    for (int i = tid; i < kCrvN; i += kCrvThreads) {
        const int64_t n = n0 + i;
        float2 z = make_float2(0.0f, 0.0f);
        if (n >= 0 && n < L) {
            z.x = __ldg(xl + n);
            z.y = __ldg(xr + n);
            if (has_div) { z.x = z.x / div; z.y = z.y / div; }
        }
        crv_buf[i] = z;
    }
    __syncthreads();
    fft16k(crv_buf, tws, tid);
    float2 *dst = X + ((size_t)p * nblocks + j) * kCrvN;
    for (int k = tid; k < kCrvN; k += kCrvThreads) dst[k] = crv_buf[k];
}

// ---- K3: Y_j[k] = sum_{p < K} H_p[k] * X_{j-p}[k] per channel, in place over X
template <int K>
__global__ void __launch_bounds__(128) crv_mac_kernel(const float4 *__restrict__ H, float2 *X, int nblocks) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (k > kCrvN / 2) return;
    const int partner = (kCrvN - k) & (kCrvN - 1);
    float4 h[K];
#pragma unroll
    for (int q = 0; q < K; ++q) h[q] = __ldg(H + ((size_t)p * K + q) * (kCrvN / 2 + 1) + k);
    float4 hist[K];  // (XL.re, XL.im, XR.re, XR.im) of blocks j, j-1, ...
#pragma unroll
    for (int q = 0; q < K; ++q) hist[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 *base = X + (size_t)p * nblocks * kCrvN;
    float2 a = base[k], b = base[partner];
    for (int j = 0; j < nblocks; ++j) {
        float2 an = a, bn = b;
        if (j + 1 < nblocks) { an = base[(size_t)(j + 1) * kCrvN + k]; bn = base[(size_t)(j + 1) * kCrvN + partner]; }
#pragma unroll
        for (int q = K - 1; q > 0; --q) hist[q] = hist[q - 1];
        hist[0] = make_float4(0.5f * (a.x + b.x), 0.5f * (a.y - b.y), 0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
        float ylr = 0.f, yli = 0.f, yrr = 0.f, yri = 0.f;
#pragma unroll
        for (int q = 0; q < K; ++q) {
            ylr = fmaf(h[q].x, hist[q].x, ylr); ylr = fmaf(-h[q].y, hist[q].y, ylr);
            yli = fmaf(h[q].x, hist[q].y, yli); yli = fmaf(h[q].y, hist[q].x, yli);
            yrr = fmaf(h[q].z, hist[q].z, yrr); yrr = fmaf(-h[q].w, hist[q].w, yrr);
            yri = fmaf(h[q].z, hist[q].w, yri); yri = fmaf(h[q].w, hist[q].z, yri);
        }
        // Y = YL + i YR;  Y[N-k] = conj(YL) + i conj(YR)
        base[(size_t)j * kCrvN + k] = make_float2(ylr - yri, yli + yrr);
        if (partner != k) base[(size_t)j * kCrvN + partner] = make_float2(ylr + yri, yrr - yli);
        a = an; b = bn;
    }
}

// ---- K4: inverse FFT of Y_j (ifft(Y) = conj(fft(conj Y)) / N), samples [B, 2B) are block j's outputs; mix; peak
__global__ void __launch_bounds__(kCrvThreads, 1) crv_ifft_mix_kernel(SigView in, const float *in_peak, float *out,
                                                                       int64_t L, int nblocks,
                                                                       const ConvRevParams *__restrict__ prm,
                                                                       const float2 *__restrict__ tw,
                                                                       const float2 *__restrict__ Y,
                                                                       unsigned *out_peak) {
    extern __shared__ float2 crv_buf[];
    float2 *tws = crv_buf + kCrvN;
    const int j = blockIdx.x, p = blockIdx.y, tid = threadIdx.x;
    load_twiddles(tws, tw, tid);
    const float2 *src = Y + ((size_t)p * nblocks + j) * kCrvN;
    for (int k = tid; k < kCrvN; k += kCrvThreads) {
        const float2 v = src[k];
        crv_buf[k] = make_float2(v.x, -v.y);
    }
    __syncthreads();
    fft16k(crv_buf, tws, tid);
    const bool has_div = in_peak != nullptr;
    const float div = has_div ? fmaxf(in_peak[p], 1e-8f) : 1.0f;
    const float mix = prm[p].mix, dry = __fsub_rn(1.0f, mix);
    const float *xl = in.base + (int64_t)p * in.stride_p;
    const float *xr = xl + in.stride_c;
    float *ol = out + (int64_t)p * 2 * L, *orr = ol + L;
    const float inv = 1.0f / (float)kCrvN;
    float pk = 0.0f;
    for (int i = tid; i < kCrvB; i += kCrvThreads) {
        const int64_t n = (int64_t)j * kCrvB + i;
        if (n < L) {
            const float2 r = crv_buf[kCrvB + i];
            float l = __ldg(xl + n), rr = __ldg(xr + n);
            if (has_div) { l = l / div; rr = rr / div; }
            const float yl = __fadd_rn(__fmul_rn(dry, l), __fmul_rn(mix, r.x * inv));
            const float yr = __fadd_rn(__fmul_rn(dry, rr), __fmul_rn(mix, -r.y * inv));
            ol[n] = yl;
            orr[n] = yr;
            pk = fmaxf(pk, fmaxf(fabsf(yl), fabsf(yr)));
        }
    }
    if (out_peak != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
        if ((tid & 31) == 0 && pk > 0.0f) atomicMax(out_peak + p, __float_as_uint(pk));
    }
}

// ------------------------------------------------------------------------------------------------ host side
uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// element i of the seeded white noise: Box-Muller of two uniforms hashed from (seed, i)  (oracle/convreverb.py)
void white_noise(uint64_t seed, size_t count, float *out) {
    const uint64_t key = splitmix64(seed);
    for (size_t i = 0; i < count; ++i) {
        const uint64_t a = splitmix64(key + 2 * (uint64_t)i), b = splitmix64(key + 2 * (uint64_t)i + 1);
        const double u1 = ((double)(a >> 11) + 1.0) * 0x1.0p-53, u2 = (double)(b >> 11) * 0x1.0p-53;
        out[i] = (float)(std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2));
    }
}

// scipy.signal.firwin(numtaps, cutoffs, window="hamming", pass_zero=..., scale=True, fs=fs), bands given as pass-band
// edges normalised to Nyquist
void firwin(int numtaps, double left, double right, double *h) {
    const double alpha = 0.5 * (numtaps - 1);
    auto sinc = [](double x) { return x == 0.0 ? 1.0 : std::sin(M_PI * x) / (M_PI * x); };
    for (int n = 0; n < numtaps; ++n) {
        const double m = n - alpha;
        const double fac = -M_PI + 2.0 * M_PI * n / (numtaps - 1);
        const double win = 0.54 + 0.46 * std::cos(fac);  // general_hamming(M, 0.54, sym=True)
        h[n] = (right * sinc(right * m) - left * sinc(left * m)) * win;
    }
    const double scale_frequency = left == 0.0 ? 0.0 : (right == 1.0 ? 1.0 : 0.5 * (left + right));
    double s = 0.0;
    for (int n = 0; n < numtaps; ++n) s += h[n] * std::cos(M_PI * (n - alpha) * scale_frequency);
    for (int n = 0; n < numtaps; ++n) h[n] /= s;
}

void octave_filterbank(double fs, std::vector<float> &out) {
    static const double centres[10] = {31.5, 63, 125, 250, 500, 1000, 2000, 4000, 8000, 16000};
    const double nyq = 0.5 * fs;
    out.resize((size_t)kCrvBands * kCrvTaps);
    std::vector<double> h(kCrvTaps);
    for (int b = 0; b < kCrvBands; ++b) {
        double left, right;
        if (b == 0) { left = 0.0; right = 12.0 / nyq; }                 // low-pass 12 Hz
        else if (b == kCrvBands - 1) { left = 18000.0 / nyq; right = 1.0; }  // high-pass 18 kHz
        else {
            const double fc = centres[b - 1];
            double fmax = fc * std::sqrt(2.0);
            const double cap = nyq * 0.999;
            if (fmax > cap) fmax = cap;
            left = (fc / std::sqrt(2.0)) / nyq;
            right = fmax / nyq;
        }
        firwin(kCrvTaps, left, right, h.data());
        for (int t = 0; t < kCrvTaps; ++t) out[(size_t)b * kCrvTaps + t] = (float)h[t];
    }
}

}  // namespace

// host-side pieces of the setup, exported through the C ABI for the CPU test-suite (tests/test_convreverb_oracle.py)
void convreverb_host_filterbank(double sample_rate, float *out /*[12][1023]*/) {
    std::vector<float> f;
    octave_filterbank(sample_rate, f);
    for (size_t i = 0; i < f.size(); ++i) out[i] = f[i];
}
void convreverb_host_noise(uint64_t seed, size_t count, float *out) { white_noise(seed, count, out); }

void convreverb_release(ConvReverbState *s) {
    void **ptrs[] = {(void **)&s->bands, (void **)&s->twiddle, (void **)&s->H, (void **)&s->X};
    for (void **p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    s->H_cap = s->X_cap = 0;
    s->n_ir = 0;
}

cudaError_t convreverb_prepare(cudaStream_t st, ConvReverbState *s, double sample_rate, int n_ir, int seed) {
    if (s->bands && s->n_ir == n_ir && s->seed == seed && s->sample_rate == sample_rate) return cudaSuccess;
    if (n_ir < 2 || n_ir > 16 * kCrvB || sample_rate < 2.0 * 18000.0 / 0.999) return cudaErrorInvalidValue;
    cudaError_t e;
    if (!s->twiddle) {
        std::vector<float2> tw(kCrvN);
        for (int k = 0; k < kCrvN; ++k) {
            const double a = -2.0 * M_PI * (double)k / (double)kCrvN;
            tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        if ((e = cudaMalloc((void **)&s->twiddle, kCrvN * sizeof(float2))) != cudaSuccess) return e;
        if ((e = cudaMemcpy(s->twiddle, tw.data(), kCrvN * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    }
    if (s->bands) cudaFree(s->bands);
    s->bands = nullptr;
    const size_t span = (size_t)n_ir + kCrvTaps - 1, nwn = 2 * kCrvBands * span;
    std::vector<float> wn(nwn), filt;
    white_noise((uint64_t)(int64_t)seed, nwn, wn.data());
    octave_filterbank(sample_rate, filt);
    float *dwn = nullptr, *dfilt = nullptr;
    if ((e = cudaMalloc((void **)&dwn, nwn * sizeof(float))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void **)&dfilt, filt.size() * sizeof(float))) != cudaSuccess) { cudaFree(dwn); return e; }
    if ((e = cudaMalloc((void **)&s->bands, (size_t)2 * kCrvBands * n_ir * sizeof(float))) != cudaSuccess) { cudaFree(dwn); cudaFree(dfilt); return e; }
    e = cudaMemcpyAsync(dwn, wn.data(), nwn * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dfilt, filt.data(), filt.size() * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        dim3 grid((n_ir + 255) / 256, 2 * kCrvBands);
        crv_bands_kernel<<<grid, 256, 0, st>>>(dwn, dfilt, s->bands, n_ir);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // wn / filt live on the host stack until here
    cudaFree(dwn);
    cudaFree(dfilt);
    if (e != cudaSuccess) return e;
    s->n_ir = n_ir; s->seed = seed; s->sample_rate = sample_rate;
    return cudaSuccess;
}

cudaError_t launch_convreverb(cudaStream_t st, ConvReverbState *s, SigView in, const float *in_peak, float *out, int P,
                              int64_t L, const ConvRevParams *prm, unsigned *out_peak, int *launches) {
    const int K = (s->n_ir + kCrvB - 1) / kCrvB;
    const int Kt = K <= 8 ? 8 : (K <= 12 ? 12 : 16);  // register-resident partitions of crv_mac_kernel (unused ones are zero)
    const int nblocks = (int)((L + kCrvB - 1) / kCrvB);
    cudaError_t e;
    const size_t hbytes = (size_t)P * Kt * (kCrvN / 2 + 1) * sizeof(float4);
    const size_t xbytes = (size_t)P * nblocks * kCrvN * sizeof(float2);
    if (hbytes > s->H_cap) {
        if (s->H) cudaFree(s->H);
        s->H = nullptr; s->H_cap = 0;
        if ((e = cudaMalloc((void **)&s->H, hbytes)) != cudaSuccess) return e;
        s->H_cap = hbytes;
    }
    if (xbytes > s->X_cap) {
        if (s->X) cudaFree(s->X);
        s->X = nullptr; s->X_cap = 0;
        if ((e = cudaMalloc((void **)&s->X, xbytes)) != cudaSuccess) return e;
        s->X_cap = xbytes;
    }
    if ((e = ensure_dyn_smem(reinterpret_cast<const void *>(&crv_ir_fft_kernel), (int)kCrvSmem)) != cudaSuccess) return e;
    if ((e = ensure_dyn_smem(reinterpret_cast<const void *>(&crv_x_fft_kernel), (int)kCrvSmem)) != cudaSuccess) return e;
    if ((e = ensure_dyn_smem(reinterpret_cast<const void *>(&crv_ifft_mix_kernel), (int)kCrvSmem)) != cudaSuccess) return e;
    // partitions K .. Kt-1 lie beyond the IR: crv_ir_fft_kernel writes all-zero spectra for them
    crv_ir_fft_kernel<<<dim3(Kt, P), kCrvThreads, kCrvSmem, st>>>(s->bands, s->n_ir, Kt, prm, s->twiddle, s->H);
    crv_x_fft_kernel<<<dim3(nblocks, P), kCrvThreads, kCrvSmem, st>>>(in, in_peak, L, nblocks, s->twiddle, s->X);
    dim3 mg((kCrvN / 2 + 1 + 127) / 128, P);
    if (Kt == 8) crv_mac_kernel<8><<<mg, 128, 0, st>>>(s->H, s->X, nblocks);
    else if (Kt == 12) crv_mac_kernel<12><<<mg, 128, 0, st>>>(s->H, s->X, nblocks);
    else crv_mac_kernel<16><<<mg, 128, 0, st>>>(s->H, s->X, nblocks);
    crv_ifft_mix_kernel<<<dim3(nblocks, P), kCrvThreads, kCrvSmem, st>>>(in, in_peak, out, L, nblocks, prm, s->twiddle,
                                                                        s->X, out_peak);
    *launches += 4;
    return cudaGetLastError();
}

}  // namespace stito
