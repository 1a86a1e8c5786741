// C-ABI layer of libstito.so (declared in include/stito.h): handle management, the host side of the
// effect-chain compilation (parameter de-normalisation and filter design run here on the host in
// fp64/libm so they match the reference's numpy/scipy arithmetic), kernel orchestration, timing.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "encoder_tc.h"
#include "stito_internal.h"

using namespace stito;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(STITO_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),      \
                        __FILE__, __LINE__);                                                      \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {  // pinned staging
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// value ranges of the Basic* plugins' Parameter objects (st_ito/effects.py:823-840, 884-889,
// 902-905, 921-925, 945-950), in the order of each plugin's `.parameters` dict
struct Range { double lo, hi; };
const Range kEqRanges[18] = {
    {-24, 24}, {20, 4000}, {0.1, 4},  {-24, 24}, {20, 10000}, {0.1, 4}, {-24, 24}, {20, 10000}, {0.1, 4},
    {-24, 24}, {20, 10000}, {0.1, 4}, {-24, 24}, {20, 10000}, {0.1, 4}, {-24, 24}, {200, 18000}, {0.1, 4}};
const Range kCompRanges[4] = {{-80, 0}, {1, 20}, {0.1, 100}, {10, 1000}};
const Range kDistRanges[2] = {{-48, 48}, {-24, 24}};
const Range kDelayRanges[3] = {{0.01, 1.0}, {0.05, 1.0}, {0.0, 1.0}};
const Range kReverbRanges[4] = {{0, 1}, {0, 1}, {0, 1}, {0, 1}};
const Range kLtiCompRanges[6] = {{-60, 0}, {1, 20}, {0.1, 250}, {10, 2000}, {1, 24}, {0, 24}};  // effects.py:629-634
const Range kUnitRanges[STITO_MAX_FX_PARAMS] = {{0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1},
                                                {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1},
                                                {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}};

int fx_num_params(int kind) {
    switch (kind) {
        case STITO_FX_EQ: return 18;
        case STITO_FX_COMPRESSOR: return 4;
        case STITO_FX_DISTORTION: return 2;
        case STITO_FX_DELAY: return 3;
        case STITO_FX_REVERB: return 4;
        case STITO_FX_CONV_REVERB: return 25;
        case STITO_FX_LTI_COMPRESSOR: return 6;
    }
    return -1;
}
const Range *fx_ranges(int kind) {
    switch (kind) {
        case STITO_FX_EQ: return kEqRanges;
        case STITO_FX_COMPRESSOR: return kCompRanges;
        case STITO_FX_DISTORTION: return kDistRanges;
        case STITO_FX_DELAY: return kDelayRanges;
        case STITO_FX_REVERB: return kReverbRanges;
        case STITO_FX_CONV_REVERB: return kUnitRanges;
        case STITO_FX_LTI_COMPRESSOR: return kLtiCompRanges;
    }
    return nullptr;
}

// Parameter.get_value(): raw * (max - min) + min in Python floats (effects.py:795-797)
inline double denorm(double raw, const Range &r) { return raw * (r.hi - r.lo) + r.lo; }

// biqaud() of st_ito/effects.py:395-450; out = {b0,b1,b2,a1,a2} / a0
void biquad_design(double gain_db, double fc, double q, double fs, int type, double *out) {
    const double A = std::pow(10.0, gain_db / 40.0);
    const double w0 = 2.0 * M_PI * (fc / fs);
    const double alpha = std::sin(w0) / (2.0 * q);
    const double c = std::cos(w0);
    const double sA = std::sqrt(A);
    double b0, b1, b2, a0, a1, a2;
    if (type == 2) {  // high shelf
        b0 = A * ((A + 1) + (A - 1) * c + 2 * sA * alpha);
        b1 = -2 * A * ((A - 1) + (A + 1) * c);
        b2 = A * ((A + 1) + (A - 1) * c - 2 * sA * alpha);
        a0 = (A + 1) - (A - 1) * c + 2 * sA * alpha;
        a1 = 2 * ((A - 1) - (A + 1) * c);
        a2 = (A + 1) - (A - 1) * c - 2 * sA * alpha;
    } else if (type == 0) {  // low shelf
        b0 = A * ((A + 1) - (A - 1) * c + 2 * sA * alpha);
        b1 = 2 * A * ((A - 1) - (A + 1) * c);
        b2 = A * ((A + 1) - (A - 1) * c - 2 * sA * alpha);
        a0 = (A + 1) + (A - 1) * c + 2 * sA * alpha;
        a1 = -2 * ((A - 1) + (A + 1) * c);
        a2 = (A + 1) + (A - 1) * c - 2 * sA * alpha;
    } else {  // peaking
        b0 = 1 + alpha * A;
        b1 = -2 * c;
        b2 = 1 - alpha * A;
        a0 = 1 + alpha / A;
        a1 = -2 * c;
        a2 = 1 - alpha / A;
    }
    out[0] = b0 / a0; out[1] = b1 / a0; out[2] = b2 / a0; out[3] = a1 / a0; out[4] = a2 / a0;
}

constexpr int kNumEvents = 8;
constexpr int kNumFlags = 16;  // see stito_handle::flags

}  // namespace

struct stito_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    stito_chain_desc chain{};
    bool has_encoder = false;
    int precision = 0;
    int microbatch = 64;

    // input
    DevBuf input;  // [chs][cap_len] zero padded
    int in_chs = 0;
    int64_t in_len = 0, in_cap = 0;
    // target embeddings [2][E]
    DevBuf target;
    bool has_target = false;

    // encoder
    EncoderDev enc{};
    FrontendTables ft{};
    std::vector<void *> owned;  // device allocations freed on destroy
    int n_fft = 2048, hop = 1024, n_mels = 128, embed_dim = 512;

    // work buffers
    DevBuf audio[3], eq_f, eq_s, params, peaks, Wdev, feat, act[3], pooled, emb, fit, flags, xin, ready;
    // compressor -> reverb streaming (small populations): second stream + fork / join events
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    HostBuf hparams, hW, hflags;
    size_t hparams_cursor = 0;  // every micro-batch of a call designs into its own slice of the pinned staging buffer
    // flags (device ints): [0], [1] NaN in mid / side embeddings; [2] an activation left the fp16 range of the fp16x3
    // encoder; [3] compressor super-blocks redone serially (Newton iteration did not converge)
    //        [4..15] per conv layer: float bits of the largest stored activation (tensor-core path)
    bool tc_calibrated = false;  // per-layer activation scales (tcws.act_shift) have been measured on this handle
    int tc_attempts = 0;
    int comp_fallbacks_total = 0;
    ReverbGeom rgeom{};
    ConvReverbState crv;
    // multi-GPU fitness exchange over peer memory (stito_gather_*)
    void *gather_block = nullptr;          // exported: [2][capacity] floats + [2][kGatherMaxWorld] ints
    int gather_capacity = 0, gather_epoch = 0;
    bool gather_attached = false;
    GatherPeers gather_peers{};
    void *gather_opened[kGatherMaxWorld] = {};
    int *gather_done = nullptr;
    TcWorkspace tcws;

    // timing
    cudaEvent_t ev[kNumEvents] = {};
    cudaEvent_t ev_conv[13] = {};
    stito_timing timing{};
    bool timing_pending = false;
    double timing_scale = 1.0;
};

// After a synchronised pass of the tensor-core encoder: calibrate the per-layer storage scales of the fp16 hi/lo
// activation pairs from the measured per-layer maxima (hflags[4..15], float bits of the largest STORED value).
//   * first pass on this handle: shift_l = floor(log2(16384 / max_l)) -- a factor 4 of head-room below the fp16 maximum,
//     lo parts far above the subnormal range; if any shift changed the pass is redone once with the new scales;
//   * later: only an overflow (hflags[2]: a stored value exceeded 65504 and was clamped) re-calibrates -- maxima only ever
//     widen the range -- and the pass is redone; layers downstream of a clamped one are re-measured by that pass.
// Returns true when the caller must redo the pass.  After 14 fruitless attempts the handle drops to the fp32 encoder.
static bool tc_after_pass(stito_handle *h) {
    if (h->precision != 1) return false;
    const int *hf = h->hflags.as<int>();
    const bool overflow = hf[2] != 0;
    if (h->tc_calibrated && !overflow) { h->tc_attempts = 0; return false; }
    if (++h->tc_attempts > 14) {
        fprintf(stderr, "libstito: could not find fp16x3 activation scales without overflow; this handle now uses the fp32 "
                        "CUDA-core encoder\n");
        h->precision = 0;
        h->tc_attempts = 0;
        return true;
    }
    bool changed = false;
    for (int l = 0; l < 11; ++l) {  // layer 11 (block 6 conv 2) writes fp32
        float stored;
        memcpy(&stored, &hf[4 + l], sizeof(float));
        if (!(stored > 0.0f) || !std::isfinite(stored)) continue;
        const double true_max = (double)stored / std::ldexp(1.0, h->tcws.act_shift[l]);
        int ns = (int)std::floor(std::log2(16384.0 / true_max));
        if (ns > 24) ns = 24;
        if (ns < -12) ns = -12;
        if (h->tc_calibrated && ns > h->tcws.act_shift[l]) ns = h->tcws.act_shift[l];  // after calibration: only widen
        if (ns != h->tcws.act_shift[l]) { h->tcws.act_shift[l] = ns; changed = true; }
    }
    h->tc_calibrated = !overflow;
    if (!changed && !overflow) { h->tc_attempts = 0; return false; }
    return true;
}

namespace {

cudaError_t dev_alloc_copy(stito_handle *h, const void *src, size_t bytes, void **out) {
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
    h->owned.push_back(p);
    if (src) e = cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    *out = p;
    return e;
}

int validate_chain(const stito_chain_desc *c) {
    if (!c) return fail(STITO_EINVAL, "chain descriptor is NULL");
    if (c->num_fx < 0 || c->num_fx > STITO_MAX_FX) return fail(STITO_EINVAL, "num_fx %d out of range", c->num_fx);
    if (!(c->sample_rate > 0)) return fail(STITO_EINVAL, "sample_rate must be positive");
    for (int f = 0; f < c->num_fx; ++f) {
        const stito_fx_desc &d = c->fx[f];
        const int np = fx_num_params(d.kind);
        if (np < 0) return fail(STITO_EINVAL, "effect %d: unknown kind %d", f, d.kind);
        if (d.num_params != np) return fail(STITO_EINVAL, "effect %d: kind %d takes %d parameters, got %d", f, d.kind, np, d.num_params);
        if (d.num_channels != 1 && d.num_channels != 2) return fail(STITO_EINVAL, "effect %d: num_channels must be 1 or 2", f);
        if (d.kind == STITO_FX_CONV_REVERB) {
            if (d.num_channels != 2) return fail(STITO_EINVAL, "effect %d: the convolution reverb is a 2-channel plugin", f);
            if (d.iopt[0] < 2 || d.iopt[0] > 131072) return fail(STITO_EINVAL, "effect %d: impulse-response length %d outside [2, 131072]", f, d.iopt[0]);
            if (c->sample_rate < 36100.0) return fail(STITO_EINVAL, "effect %d: the 18 kHz band of the convolution reverb needs a sample rate above 36.1 kHz", f);
        }
        if (d.kind == STITO_FX_LTI_COMPRESSOR && (d.iopt[0] < 0 || d.iopt[0] > (1 << 20)))
            return fail(STITO_EINVAL, "effect %d: look-ahead of %d samples outside [0, 2^20]", f, d.iopt[0]);
        for (int k = 0; k < np; ++k)
            if (d.w_index[k] >= c->num_w) return fail(STITO_EINVAL, "effect %d parameter %d: w index %d >= D=%d", f, k, d.w_index[k], c->num_w);
    }
    return STITO_OK;
}

bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// slab layout of the per-population parameter block (bytes per candidate, per effect slot)
constexpr size_t kParamSlot = 6 * 5 * sizeof(double);  // largest: EQ coefficients

// Host: de-normalise w and design every effect's constants for P candidates into hparams
// ([fx][P][kParamSlot]); returns max delay length through *max_d.
void design_params(const stito_chain_desc &c, const double *W, int P, int D, int64_t L, uint8_t *hp, int *max_d) {
    const double fs = c.sample_rate;
    *max_d = 1;
    for (int f = 0; f < c.num_fx; ++f) {
        const stito_fx_desc &d = c.fx[f];
        const Range *rg = fx_ranges(d.kind);
        for (int p = 0; p < P; ++p) {
            double v[STITO_MAX_FX_PARAMS];
            for (int k = 0; k < d.num_params; ++k) {
                const double raw = d.w_index[k] >= 0 ? W[(size_t)p * D + d.w_index[k]] : d.fixed_raw[k];
                v[k] = denorm(raw, rg[k]);
            }
            // effect f owns the block [f*P*kParamSlot, (f+1)*P*kParamSlot); inside it the kernels index a
            // dense array of the effect's own parameter struct
            uint8_t *block = hp + (size_t)f * P * kParamSlot;
            switch (d.kind) {
                case STITO_FX_EQ: {
                    double *cf = reinterpret_cast<double *>(block) + (size_t)p * 30;
                    for (int s = 0; s < 6; ++s)
                        biquad_design(v[3 * s], v[3 * s + 1], v[3 * s + 2], fs, s == 0 ? 0 : (s == 5 ? 2 : 1), cf + 5 * s);
                    break;
                }
                case STITO_FX_COMPRESSOR: {  // oracle_compressor
                    CompParams *q = reinterpret_cast<CompParams *>(block) + p;
                    const float thr_db = (float)v[0], ratio = (float)v[1], at = (float)v[2], rl = (float)v[3];
                    const double ef = -2.0 * M_PI * 1000.0 / fs;
                    q->cte_at = at < 1.0e-3f ? 0.0f : (float)std::exp(ef / (double)at);
                    q->cte_rl = rl < 1.0e-3f ? 0.0f : (float)std::exp(ef / (double)rl);
                    q->thr = thr_db > -200.0f ? powf(10.0f, thr_db * 0.05f) : 0.0f;
                    q->thr_inv = 1.0f / q->thr;
                    q->expo = 1.0f / ratio - 1.0f;
                    break;
                }
                case STITO_FX_DISTORTION: {  // oracle_distortion
                    DistParams *q = reinterpret_cast<DistParams *>(block) + p;
                    const float dr = (float)v[0], og = (float)v[1];
                    q->drive = dr > -100.0f ? powf(10.0f, dr * 0.05f) : 0.0f;
                    q->out_gain = og > -100.0f ? powf(10.0f, og * 0.05f) : 0.0f;
                    break;
                }
                case STITO_FX_DELAY: {  // oracle_delay
                    DelayParams *q = reinterpret_cast<DelayParams *>(block) + p;
                    const float ds = (float)v[0];
                    int dd = (int)((double)ds * fs);
                    if (dd < 1) dd = 1;
                    q->d = dd;
                    q->feedback = (float)v[1];
                    q->mix = (float)v[2];
                    q->dry = 1.0f - q->mix;
                    if (dd > *max_d) *max_d = dd;
                    break;
                }
                case STITO_FX_CONV_REVERB: {  // apply_reverb (effects.py:564-588): parameters are used raw
                    ConvRevParams *q = reinterpret_cast<ConvRevParams *>(block) + p;
                    for (int b = 0; b < 12; ++b) { q->gain[b] = (float)v[b]; q->decay[b] = (float)v[12 + b]; }
                    q->mix = (float)v[24];
                    break;
                }
                case STITO_FX_LTI_COMPRESSOR: {  // apply_compressor (effects.py:629-634); release_ms (v[3]) is unused upstream
                    LtiCompParams *q = reinterpret_cast<LtiCompParams *>(block) + p;
                    lticomp_design(fs, L, (float)v[0], (float)v[1], (float)v[2], (float)v[4], (float)v[5], q);
                    break;
                }
                case STITO_FX_REVERB: {  // BasicReverb.process (effects.py:952-959) + oracle_reverb
                    ReverbParams *q = reinterpret_cast<ReverbParams *>(block) + p;
                    const float room = (float)v[0], damping = (float)v[1];
                    const float wet_level = (float)v[2], dry_level = (float)(1 - v[2]), width = (float)v[3];
                    const float wet = wet_level * 3.0f;
                    q->dry = dry_level * 2.0f;
                    q->wet1 = 0.5f * wet * (1.0f + width);
                    q->wet2 = 0.5f * wet * (1.0f - width);
                    q->damp = damping * 0.4f;
                    const float t = room * 0.28f;
                    q->fb = t + 0.7f;
                    break;
                }
            }
        }
    }
}

int out_channels(const stito_chain_desc &c, int chs) {
    for (int f = 0; f < c.num_fx; ++f)
        if (c.fx[f].num_channels == 2 && chs == 1) chs = 2;
    return chs;
}

// Runs the chain for P candidates on `in` (chs channels, L samples).  Result: *res points at
// [P][out_chs][L] un-normalised audio in one of h->audio[], *res_peak at its per-candidate peaks.
int run_chain(stito_handle *h, cudaStream_t st, SigView in, int chs, int64_t L, const double *W_host, int P,
              int D, const float **res, const unsigned **res_peak, int *res_chs, int *launches) {
    const stito_chain_desc &c = h->chain;
    const int ochs = out_channels(c, chs);
    const size_t abytes = (size_t)P * ochs * L * sizeof(float);
    CU(h->audio[0].ensure(abytes));
    if (c.num_fx > 1) CU(h->audio[1].ensure(abytes));
    CU(h->peaks.ensure((size_t)(c.num_fx + 1) * P * sizeof(unsigned)));
    CU(cudaMemsetAsync(h->peaks.p, 0, (size_t)(c.num_fx + 1) * P * sizeof(unsigned), st));
    unsigned *peaks = h->peaks.as<unsigned>();

    if (c.num_fx == 0) {
        CU(launch_copy(st, in, nullptr, h->audio[0].as<float>(), P, chs, L, peaks, launches));
        *res = h->audio[0].as<float>();
        *res_peak = peaks;
        *res_chs = chs;
        return STITO_OK;
    }
    // parameters: host design -> one H2D copy
    const size_t pbytes = (size_t)c.num_fx * P * kParamSlot;
    // The H2D copy below is asynchronous: with three or more micro-batches in flight the host would otherwise overwrite the
    // pinned staging block of micro-batch i + 1 (whose copy is queued behind the kernels of micro-batch i) with the parameters
    // of micro-batch i + 2.  Every micro-batch of a call therefore gets its own slice (the caller reserved the capacity and
    // reset the cursor); if the slices are exhausted, drain the stream and start over.
    if (h->hparams_cursor + pbytes > h->hparams.cap) {
        CU(cudaStreamSynchronize(st));
        h->hparams_cursor = 0;
        CU(h->hparams.ensure(pbytes));
    }
    uint8_t *hslice = h->hparams.as<uint8_t>() + h->hparams_cursor;
    h->hparams_cursor += (pbytes + 255) & ~(size_t)255;
    CU(h->params.ensure(pbytes));
    int max_d = 1;
    design_params(c, W_host, P, D, L, hslice, &max_d);
    CU(cudaMemcpyAsync(h->params.p, hslice, pbytes, cudaMemcpyHostToDevice, st));

    SigView cur = in;
    int cur_chs = chs;
    const float *in_peak = nullptr;
    // Small populations leave most SMs idle while each of these latency-bound kernels walks its streams one after the
    // other: a compressor directly followed by the Freeverb then runs as a STREAMING pair -- both kernels resident at the
    // same time (main + auxiliary stream; 2 * P * chs CTAs must fit the SMs), the reverb consuming super-blocks of 32768
    // samples as the compressor publishes them (flags in h->ready) -- so the pair costs max(comp, reverb), not the sum.
    static const bool streaming_on = !(getenv("STITO_DSP_STREAMING") && atoi(getenv("STITO_DSP_STREAMING")) == 0);
    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, h->device);
    float *stream_out = nullptr;  // output buffer override for the reverb of a streaming pair
    const int *stream_ready = nullptr;
    for (int f = 0; f < c.num_fx; ++f) {
        const stito_fx_desc &d = c.fx[f];
        if (d.num_channels == 2 && cur_chs == 1) {  // np.concatenate((x, x)) style_transfer.py:94-95
            cur.stride_c = 0;
            cur_chs = 2;
        }
        float *out = stream_out ? stream_out : h->audio[f & 1].as<float>();
        stream_out = nullptr;
        const bool last = f == c.num_fx - 1;
        unsigned *opk = (last || c.normalize_stages) ? peaks + (size_t)f * P : nullptr;
        const uint8_t *slot = h->params.as<uint8_t>() + (size_t)f * P * kParamSlot;
        switch (d.kind) {
            case STITO_FX_EQ: {
                const size_t nd = eq_scratch_doubles(P, cur_chs, L);
                CU(h->eq_f.ensure(nd * sizeof(double)));
                CU(h->eq_s.ensure(nd * sizeof(double)));
                CU(launch_eq(st, cur, in_peak, out, P, cur_chs, L, reinterpret_cast<const double *>(slot),
                             h->eq_f.as<double>(), h->eq_s.as<double>(), opk, launches));
                break;
            }
            case STITO_FX_COMPRESSOR: {
                int *ready = nullptr;
                // (Tried for 15..18 candidates, where the cluster-split Freeverb fits the GPU alone but not next to the
                // compressor's CTAs: compressor THEN split reverb.  Slower -- 16 clusters of 8 do not all become resident at
                // once (a cluster must sit inside one GPC), the last one runs as a second wave: DSP 1.59 vs 1.32 ms.)
                const bool pair = streaming_on && !c.normalize_stages && f + 1 < c.num_fx && cur_chs == 2 &&
                                  c.fx[f + 1].kind == STITO_FX_REVERB && c.fx[f + 1].num_channels == 2 &&
                                  reverb_can_stream(h->rgeom) && 2 * P * cur_chs + 16 <= sm_count;
                if (pair) {
                    const size_t nflags = (size_t)P * cur_chs * ((L + kStreamGranule - 1) / kStreamGranule);
                    CU(h->ready.ensure(nflags * sizeof(int)));
                    CU(cudaMemsetAsync(h->ready.p, 0, nflags * sizeof(int), st));
                    CU(h->audio[2].ensure(abytes));
                    ready = h->ready.as<int>();
                    stream_ready = ready;
                    stream_out = h->audio[2].as<float>();  // the reverb must not write where the compressor still reads
                    CU(cudaEventRecord(h->ev_fork, st));   // everything up to here (EQ output, flags, parameters) ...
                    CU(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));  // ... precedes the reverb on the second stream
                }
                CU(launch_compressor(st, cur, in_peak, out, P, cur_chs, L, reinterpret_cast<const CompParams *>(slot), opk,
                                     h->flags.as<int>() + 3, ready, launches));
                break;
            }
            case STITO_FX_DISTORTION:
                CU(launch_distortion(st, cur, in_peak, out, P, cur_chs, L, reinterpret_cast<const DistParams *>(slot), opk, launches));
                break;
            case STITO_FX_DELAY:
                CU(launch_delay(st, cur, in_peak, out, P, cur_chs, L, reinterpret_cast<const DelayParams *>(slot), max_d, opk, launches));
                break;
            case STITO_FX_CONV_REVERB: {
                CU(convreverb_prepare(st, &h->crv, c.sample_rate, d.iopt[0], d.iopt[1]));
                CU(launch_convreverb(st, &h->crv, cur, in_peak, out, P, L, reinterpret_cast<const ConvRevParams *>(slot), opk, launches));
                break;
            }
            case STITO_FX_LTI_COMPRESSOR: {
                const int link = (cur_chs == 2 && d.num_channels == 2) ? 2 : 1;
                CU(h->eq_f.ensure(lticomp_scratch_bytes(P, cur_chs, link, L)));
                CU(launch_lticomp(st, cur, in_peak, out, P, cur_chs, link, L, d.iopt[0], reinterpret_cast<const LtiCompParams *>(slot),
                                  h->eq_f.p, opk, launches));
                break;
            }
            case STITO_FX_REVERB: {
                const int stereo = (cur_chs == 2 && d.num_channels == 2) ? 1 : 0;
                cudaStream_t rst = stream_ready ? h->aux_stream : st;
                cudaError_t e = launch_reverb(rst, cur, in_peak, out, P, cur_chs, stereo, L, h->rgeom,
                                              reinterpret_cast<const ReverbParams *>(slot), opk, stream_ready,
                                              sm_count - 4 - (stream_ready ? P * cur_chs : 0), launches);
                if (e == cudaErrorInvalidValue) return fail(STITO_EINVAL, "reverb: unsupported sample rate %.1f", c.sample_rate);
                CU(e);
                if (stream_ready) {  // join: the main stream continues after the reverb
                    CU(cudaEventRecord(h->ev_join, h->aux_stream));
                    CU(cudaStreamWaitEvent(st, h->ev_join, 0));
                    stream_ready = nullptr;
                }
                break;
            }
        }
        cur.base = out;
        cur.stride_p = (int64_t)cur_chs * L;
        cur.stride_c = L;
        in_peak = c.normalize_stages ? reinterpret_cast<const float *>(opk) : nullptr;
    }
    *res = cur.base;
    *res_peak = peaks + (size_t)(c.num_fx - 1) * P;
    *res_chs = cur_chs;
    return STITO_OK;
}

// Encoder forward for B items whose audio is described by `in` (+ optional per-item peaks):
// writes RAW mid/side [B][E] to the device buffers.
int encoder_forward(stito_handle *h, cudaStream_t st, SigView in, const unsigned *peak, int B, int chs,
                    int64_t L, float *mid, float *side, int *launches, bool time_layers) {
    if (!h->has_encoder) return fail(STITO_ESTATE, "handle was created without encoder weights");
    const int T = (int)(L / h->hop) + 1;
    if (T < 32) return fail(STITO_EINVAL, "audio too short for the encoder: %lld samples give %d frames, 32 needed", (long long)L, T);
    const int N = B * chs;
    CU(h->feat.ensure((size_t)N * T * h->n_mels * sizeof(float)));
    if (time_layers) CU(cudaEventRecord(h->ev[2], st));
    CU(launch_logmel(st, in, peak, B, chs, L, T, h->ft, h->feat.as<float>(), launches));
    if (time_layers) CU(cudaEventRecord(h->ev[3], st));
    CU(h->pooled.ensure((size_t)N * 2048 * sizeof(float)));
    if (h->precision == 1) {
        int rc = tc_encoder_forward(st, h->enc, h->tcws, h->feat.as<float>(), N, T, h->n_mels, h->pooled.as<float>(), launches,
                                    time_layers ? h->ev_conv : nullptr);
        if (rc != 0) return fail(STITO_ECUDA, "tensor-core encoder failed: %s", tc_last_error());
    } else {
        const size_t a0 = (size_t)N * T * h->n_mels * 64 * sizeof(float);
        for (int i = 0; i < 3; ++i) CU(h->act[i].ensure(a0));
        float *X = h->act[0].as<float>(), *Y = h->act[1].as<float>(), *Z = h->act[2].as<float>();
        int H = T, Wd = h->n_mels;
        for (int b = 0; b < 6; ++b) {
            const ConvLayer &c1 = h->enc.conv[2 * b], &c2 = h->enc.conv[2 * b + 1];
            if (time_layers) CU(cudaEventRecord(h->ev_conv[2 * b], st));
            if (b == 0) CU(launch_conv_first(st, h->feat.as<float>(), c1, Y, N, H, Wd, launches));
            else CU(launch_conv_simt(st, X, c1, Y, N, H, Wd, launches));
            if (time_layers) CU(cudaEventRecord(h->ev_conv[2 * b + 1], st));
            CU(launch_conv_simt(st, Y, c2, Z, N, H, Wd, launches));
            if (b < 5) {
                CU(launch_avgpool(st, Z, X, N, H, Wd, c2.cout, launches));
                H /= 2;
                Wd /= 2;
            }
        }
        if (time_layers) CU(cudaEventRecord(h->ev_conv[12], st));
        CU(launch_global_pool(st, Z, h->pooled.as<float>(), N, H, Wd, 2048, launches));
    }
    CU(launch_heads(st, h->pooled.as<float>(), h->enc, B, chs, mid, side, launches));
    if (time_layers) CU(cudaEventRecord(h->ev[4], st));
    return STITO_OK;
}

double encoder_flops(int N, int T, int mel) {
    static const int ch[7] = {1, 64, 128, 256, 512, 1024, 2048};
    double total = 0;
    int hh = T, ww = mel;
    for (int i = 0; i < 6; ++i) {
        total += (double)hh * ww * ch[i + 1] * 9.0 * ch[i] + (double)hh * ww * ch[i + 1] * 9.0 * ch[i + 1];
        if (i < 5) { hh /= 2; ww /= 2; }
    }
    return 2.0 * total * N;
}

}  // namespace

extern "C" {

const char *stito_last_error(void) { return g_err.c_str(); }
int stito_version(void) { return 200; }

int stito_create(const stito_chain_desc *chain, const stito_encoder_weights *wts, int device,
                 stito_handle **out) {
    if (!out) return fail(STITO_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = validate_chain(chain);
    if (rc) return rc;
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(STITO_EINVAL, "device %d not present (%d CUDA devices)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(STITO_ECUDA, "libstito is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    stito_handle *h = new stito_handle();
    h->device = device;
    h->chain = *chain;
    reverb_geometry(chain->sample_rate, &h->rgeom);
    if (const char *mb = getenv("STITO_MICROBATCH")) {
        const int v = atoi(mb);
        if (v > 0) h->microbatch = v;
    }
    auto bail = [&](int code) { stito_destroy(h); return code; };
#define CUB(expr)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return bail(fail(STITO_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),  \
                             __FILE__, __LINE__));                                                 \
    } while (0)
    CUB(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < kNumEvents; ++i) CUB(cudaEventCreate(&h->ev[i]));
    for (int i = 0; i < 13; ++i) CUB(cudaEventCreate(&h->ev_conv[i]));
    CUB(h->flags.ensure(kNumFlags * sizeof(int)));
    CUB(cudaMemset(h->flags.p, 0, kNumFlags * sizeof(int)));
    CUB(h->hflags.ensure(kNumFlags * sizeof(int)));
    h->tcws.overflow_flag = h->flags.as<int>() + 2;
    h->tcws.amax = h->flags.as<unsigned>() + 4;

    if (wts) {
        if (wts->n_fft != 2048 || wts->hop <= 0 || wts->n_mels != 128 || wts->embed_dim <= 0 ||
            wts->embed_dim > STITO_EMBED_DIM_MAX)
            return bail(fail(STITO_EINVAL, "unsupported encoder geometry n_fft=%d hop=%d n_mels=%d embed_dim=%d (AFx-Rep: 2048/1024/128/512)",
                             wts->n_fft, wts->hop, wts->n_mels, wts->embed_dim));
        h->n_fft = wts->n_fft; h->hop = wts->hop; h->n_mels = wts->n_mels; h->embed_dim = wts->embed_dim;
        static const int ch[7] = {1, 64, 128, 256, 512, 1024, 2048};
        for (int l = 0; l < 12; ++l) {
            const int cin = (l & 1) ? ch[l / 2 + 1] : ch[l / 2], cout = ch[l / 2 + 1];
            if (!wts->conv_w[l] || !wts->bn_weight[l] || !wts->bn_bias[l] || !wts->bn_mean[l] || !wts->bn_var[l])
                return bail(fail(STITO_EINVAL, "encoder weights: layer %d has NULL tensors", l));
            // fold eval-mode BatchNorm: w' = w * g / sqrt(var + eps), b' = beta - mean * g / sqrt(var + eps)
            std::vector<float> wf((size_t)9 * cin * cout), bf(cout);
            std::vector<double> sc(cout);
            for (int co = 0; co < cout; ++co) {
                sc[co] = (double)wts->bn_weight[l][co] / std::sqrt((double)wts->bn_var[l][co] + (double)wts->bn_eps);
                bf[co] = (float)((double)wts->bn_bias[l][co] - (double)wts->bn_mean[l][co] * sc[co]);
            }
            for (int co = 0; co < cout; ++co)
                for (int ci = 0; ci < cin; ++ci)
                    for (int t = 0; t < 9; ++t)
                        wf[((size_t)t * cin + ci) * cout + co] =
                            (float)((double)wts->conv_w[l][((size_t)co * cin + ci) * 9 + t] * sc[co]);
            ConvLayer &cl = h->enc.conv[l];
            cl.cin = cin; cl.cout = cout;
            CUB(dev_alloc_copy(h, wf.data(), wf.size() * sizeof(float), (void **)&cl.w));
            CUB(dev_alloc_copy(h, bf.data(), bf.size() * sizeof(float), (void **)&cl.bias));
            if (cin >= 64) {
                int rc2 = tc_prepare_layer(wf.data(), cin, cout, &cl, &h->owned);
                if (rc2 != 0) return bail(fail(STITO_ECUDA, "tensor-core weight preparation failed: %s", tc_last_error()));
            }
        }
        const int E = wts->embed_dim;
        for (int hd = 0; hd < 2; ++hd) {
            const float *w = hd ? wts->fc_side_w : wts->fc_mid_w;
            const float *b = hd ? wts->fc_side_b : wts->fc_mid_b;
            if (!w || !b) return bail(fail(STITO_EINVAL, "encoder weights: fc tensors are NULL"));
            CUB(dev_alloc_copy(h, w, (size_t)2048 * E * sizeof(float), (void **)&h->enc.fc_w[hd]));
            CUB(dev_alloc_copy(h, b, (size_t)E * sizeof(float), (void **)&h->enc.fc_b[hd]));
        }
        h->enc.embed_dim = E;
        // front-end tables
        if (!wts->mel_w) return bail(fail(STITO_EINVAL, "encoder weights: mel_w is NULL"));
        const int nf = wts->n_fft, nb = nf / 2 + 1, nm = wts->n_mels;
        std::vector<float2> tw(nf);
        std::vector<float> win(nf);
        for (int k = 0; k < nf; ++k) {
            const double a = -2.0 * M_PI * (double)k / (double)nf;
            tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
            win[k] = (float)(0.5 - 0.5 * std::cos(2.0 * M_PI * (double)k / (double)nf));
        }
        std::vector<int> ms(nm), mc(nm), mo(nm);
        std::vector<float> mw;
        for (int m = 0; m < nm; ++m) {
            int lo = nb, hi = -1;
            for (int k = 0; k < nb; ++k)
                if (wts->mel_w[(size_t)k * nm + m] != 0.0f) { if (k < lo) lo = k; hi = k; }
            if (hi < 0) { lo = 0; hi = -1; }
            ms[m] = lo; mc[m] = hi - lo + 1; mo[m] = (int)mw.size();
            for (int k = lo; k <= hi; ++k) mw.push_back(wts->mel_w[(size_t)k * nm + m]);
        }
        if (mw.empty()) mw.push_back(0.0f);
        CUB(dev_alloc_copy(h, tw.data(), tw.size() * sizeof(float2), (void **)&h->ft.twiddle));
        CUB(dev_alloc_copy(h, win.data(), win.size() * sizeof(float), (void **)&h->ft.window));
        CUB(dev_alloc_copy(h, ms.data(), ms.size() * sizeof(int), (void **)&h->ft.mel_start));
        CUB(dev_alloc_copy(h, mc.data(), mc.size() * sizeof(int), (void **)&h->ft.mel_count));
        CUB(dev_alloc_copy(h, mo.data(), mo.size() * sizeof(int), (void **)&h->ft.mel_off));
        CUB(dev_alloc_copy(h, mw.data(), mw.size() * sizeof(float), (void **)&h->ft.mel_wt));
        h->ft.n_fft = nf; h->ft.hop = wts->hop; h->ft.n_mels = nm;
        h->has_encoder = true;
        h->precision = tc_available() ? 1 : 0;
        if (const char *pm = getenv("STITO_PRECISION")) h->precision = atoi(pm) ? 1 : 0;
    }
#undef CUB
    *out = h;
    return STITO_OK;
}

void stito_destroy(stito_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    cudaDeviceSynchronize();
    for (void *p : h->owned) cudaFree(p);
    if (h->aux_stream) { cudaStreamSynchronize(h->aux_stream); cudaStreamDestroy(h->aux_stream); }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    DevBuf *bufs[] = {&h->input, &h->target, &h->audio[0], &h->audio[1], &h->audio[2], &h->ready, &h->eq_f, &h->eq_s, &h->params,
                      &h->peaks, &h->Wdev, &h->feat, &h->act[0], &h->act[1], &h->act[2], &h->pooled,
                      &h->emb, &h->fit, &h->flags, &h->xin};
    for (DevBuf *b : bufs) b->release();
    h->hparams.release();
    h->hW.release();
    h->hflags.release();
    tc_workspace_release(&h->tcws);
    convreverb_release(&h->crv);
    for (int r = 0; r < kGatherMaxWorld; ++r)
        if (h->gather_opened[r]) cudaIpcCloseMemHandle(h->gather_opened[r]);
    if (h->gather_block) cudaFree(h->gather_block);
    if (h->gather_done) cudaFree(h->gather_done);
    for (int i = 0; i < kNumEvents; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 13; ++i) if (h->ev_conv[i]) cudaEventDestroy(h->ev_conv[i]);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int stito_set_chain(stito_handle *h, const stito_chain_desc *chain) {
    if (!h) return fail(STITO_EINVAL, "handle is NULL");
    int rc = validate_chain(chain);
    if (rc) return rc;
    h->chain = *chain;
    reverb_geometry(chain->sample_rate, &h->rgeom);
    return STITO_OK;
}

int stito_set_precision(stito_handle *h, int precision) {
    if (!h) return fail(STITO_EINVAL, "handle is NULL");
    if (precision != 0 && precision != 1) return fail(STITO_EINVAL, "precision must be 0 (fp32 SIMT) or 1 (fp16x3 tcgen05)");
    if (precision == 1 && !tc_available()) return fail(STITO_EINVAL, "tensor-core encoder not available in this build");
    h->precision = precision;
    return STITO_OK;
}

int stito_set_input(stito_handle *h, const float *x, int chs, int64_t L, int64_t min_len) {
    if (!h || !x) return fail(STITO_EINVAL, "NULL argument");
    if (chs != 1 && chs != 2) return fail(STITO_EINVAL, "Invalid number of channels: %d", chs);
    if (L <= 0) return fail(STITO_EINVAL, "empty input");
    CU(cudaSetDevice(h->device));
    const int64_t cap = L > min_len ? L : min_len;
    CU(h->input.ensure((size_t)chs * cap * sizeof(float)));
    cudaStream_t st = h->own_stream;
    CU(cudaMemsetAsync(h->input.p, 0, (size_t)chs * cap * sizeof(float), st));
    for (int c = 0; c < chs; ++c)
        CU(cudaMemcpyAsync(h->input.as<float>() + (size_t)c * cap, x + (size_t)c * L, (size_t)L * sizeof(float), cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    h->in_chs = chs; h->in_len = L; h->in_cap = cap;
    return STITO_OK;
}

int stito_set_target_embeds(stito_handle *h, const float *mid, const float *side, int embed_dim) {
    if (!h || !mid || !side) return fail(STITO_EINVAL, "NULL argument");
    if (embed_dim != h->embed_dim) return fail(STITO_EINVAL, "embed_dim %d != %d", embed_dim, h->embed_dim);
    CU(cudaSetDevice(h->device));
    const size_t eb = (size_t)embed_dim * sizeof(float);
    CU(h->target.ensure(2 * eb));
    CU(cudaMemcpy(h->target.p, mid, eb, cudaMemcpyDefault));
    CU(cudaMemcpy(h->target.as<float>() + embed_dim, side, eb, cudaMemcpyDefault));
    h->has_target = true;
    return STITO_OK;
}

int stito_set_target(stito_handle *h, const float *target, int chs, int64_t L) {
    if (!h || !target) return fail(STITO_EINVAL, "NULL argument");
    if (chs != 1 && chs != 2) return fail(STITO_EINVAL, "Invalid number of channels: %d", chs);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->own_stream;
    const int E = h->embed_dim;
    CU(h->xin.ensure((size_t)chs * L * sizeof(float)));
    CU(cudaMemcpyAsync(h->xin.p, target, (size_t)chs * L * sizeof(float), cudaMemcpyDefault, st));
    CU(h->peaks.ensure(sizeof(unsigned)));
    CU(cudaMemsetAsync(h->peaks.p, 0, sizeof(unsigned), st));
    SigView v{h->xin.as<float>(), (int64_t)chs * L, L};
    int launches = 0;
    CU(launch_peak(st, v, 1, chs, L, h->peaks.as<unsigned>(), &launches));
    CU(h->target.ensure((size_t)2 * E * sizeof(float)));
    float *mid = h->target.as<float>(), *side = mid + E;
    for (;;) {
        CU(cudaMemsetAsync(h->flags.as<int>() + 2, 0, (kNumFlags - 2) * sizeof(int), st));
        int rc = encoder_forward(h, st, v, h->peaks.as<unsigned>(), 1, chs, L, mid, side, &launches, false);
        if (rc) return rc;
        CU(cudaMemcpyAsync(h->hflags.p, h->flags.p, kNumFlags * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (!tc_after_pass(h)) break;
    }
    CU(launch_embed_normalize(st, mid, side, 1, E, h->flags.as<int>(), &launches));
    CU(cudaStreamSynchronize(st));
    h->has_target = true;
    return STITO_OK;
}

int stito_out_channels(const stito_handle *h, int chs) {
    if (!h) return fail(STITO_EINVAL, "handle is NULL");
    return out_channels(h->chain, chs);
}

static int fetch_W(stito_handle *h, const double *W, int P, int D, cudaStream_t st, const double **Wh) {
    if (is_device_ptr(W)) {
        CU(h->hW.ensure((size_t)P * D * sizeof(double)));
        CU(cudaMemcpyAsync(h->hW.p, W, (size_t)P * D * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        *Wh = h->hW.as<double>();
    } else {
        *Wh = W;
    }
    return STITO_OK;
}

// One pass of evaluate() over the population, enqueued on `st` (no synchronisation): chain -> log-mel -> encoder
// (h->precision) -> normalise -> fitness, results copied to the caller's buffers, flags [2], [3] to h->hflags.
static int enqueue_population(stito_handle *h, cudaStream_t st, const double *Wh, int P, int D, int64_t start,
                              int64_t len, float *fitness, float *embeds, float *audio, int *launches_out) {
    const int E = h->embed_dim;
    const int chs = h->in_chs;
    const int ochs = out_channels(h->chain, chs);
    CU(h->emb.ensure((size_t)2 * P * E * sizeof(float)));
    CU(h->fit.ensure((size_t)P * sizeof(float)));
    float *mid_all = h->emb.as<float>(), *side_all = mid_all + (size_t)P * E;
    int launches = 0;
    SigView in{h->input.as<float>() + start, 0, h->in_cap};
    CU(cudaMemsetAsync(h->flags.as<int>() + 2, 0, (kNumFlags - 2) * sizeof(int), st));
    {   // pinned parameter staging for ALL micro-batches of this call (see run_chain); nothing of it is in flight here
        const size_t nmb = (size_t)(P + h->microbatch - 1) / h->microbatch;
        const size_t per = (((size_t)h->chain.num_fx * (P < h->microbatch ? P : h->microbatch) * kParamSlot) + 255) & ~(size_t)255;
        CU(h->hparams.ensure(nmb * per + 256));
        h->hparams_cursor = 0;
    }
    CU(cudaEventRecord(h->ev[0], st));
    for (int p0 = 0; p0 < P; p0 += h->microbatch) {
        const int pb = (P - p0) < h->microbatch ? (P - p0) : h->microbatch;
        const bool tl = p0 == 0;
        const float *y = nullptr;
        const unsigned *pk = nullptr;
        int ych = 0;
        int rc = run_chain(h, st, in, chs, len, Wh + (size_t)p0 * D, pb, D, &y, &pk, &ych, &launches);
        if (rc) return rc;
        if (tl) CU(cudaEventRecord(h->ev[1], st));
        if (audio) {
            float *dst = audio + (size_t)p0 * ochs * len;
            if (is_device_ptr(audio)) {
                CU(launch_normalize(st, y, pk, dst, pb, ych, len, &launches));
            } else {
                const size_t bytes = (size_t)pb * ych * len * sizeof(float);
                CU(h->audio[0].ensure(bytes));  // no-ops for the buffer holding y (already >= bytes)
                CU(h->audio[1].ensure(bytes));
                float *tmp = (y == h->audio[0].as<float>()) ? h->audio[1].as<float>() : h->audio[0].as<float>();
                CU(launch_normalize(st, y, pk, tmp, pb, ych, len, &launches));
                CU(cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, st));
            }
        }
        SigView yv{y, (int64_t)ych * len, len};
        rc = encoder_forward(h, st, yv, pk, pb, ych, len, mid_all + (size_t)p0 * E, side_all + (size_t)p0 * E, &launches, tl);
        if (rc) return rc;
    }
    CU(launch_embed_normalize(st, mid_all, side_all, P, E, h->flags.as<int>(), &launches));
    CU(cudaEventRecord(h->ev[5], st));
    if (fitness) CU(launch_fitness(st, mid_all, side_all, h->target.as<float>(), h->target.as<float>() + E, P, E, h->fit.as<float>(), &launches));
    CU(cudaEventRecord(h->ev[6], st));
    if (fitness) CU(cudaMemcpyAsync(fitness, h->fit.p, (size_t)P * sizeof(float), cudaMemcpyDefault, st));
    if (embeds) CU(cudaMemcpyAsync(embeds, h->emb.p, (size_t)2 * P * E * sizeof(float), cudaMemcpyDefault, st));
    CU(cudaMemcpyAsync(h->hflags.p, h->flags.p, kNumFlags * sizeof(int), cudaMemcpyDeviceToHost, st));
    *launches_out += launches;
    return STITO_OK;
}

int stito_eval_population(stito_handle *h, const double *W, int P, int D, int64_t start, int64_t len,
                          float *fitness, float *embeds, float *audio, void *stream) {
    if (!h || (!W && P > 0)) return fail(STITO_EINVAL, "NULL argument");
    if (P < 0) return fail(STITO_EINVAL, "negative population size");
    if (D != h->chain.num_w) return fail(STITO_EINVAL, "parameter vectors have %d entries, chain expects %d", D, h->chain.num_w);
    if (h->in_chs == 0) return fail(STITO_ESTATE, "stito_set_input has not been called");
    if (!h->has_encoder) return fail(STITO_ESTATE, "handle was created without encoder weights");
    if (fitness && !h->has_target) return fail(STITO_ESTATE, "no target set (stito_set_target / stito_set_target_embeds)");
    if (start < 0 || len <= 0 || start + len > h->in_cap) return fail(STITO_EINVAL, "view [%lld, %lld) outside the padded input of %lld samples", (long long)start, (long long)(start + len), (long long)h->in_cap);
    if (P == 0) {  // an empty shard of a sharded population (more ranks than candidates): nothing to do, nothing written
        h->timing_pending = false;
        memset(&h->timing, 0, sizeof(h->timing));
        h->timing.precision = h->precision;
        return STITO_OK;
    }
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const double *Wh = nullptr;
    int rc = fetch_W(h, W, P, D, st, &Wh);
    if (rc) return rc;
    const int chs = h->in_chs;
    const int ochs = out_channels(h->chain, chs);
    int launches = 0;
    // An activation that exceeds what the fp16 hi/lo pairs of the tensor-core path can hold is clamped, i.e. the result is
    // wrong: tc_after_pass() re-calibrates the per-layer scales (or, as a last resort, drops to the fp32 encoder) and the
    // evaluation is redone.  Steady state: one pass.
    int overflowed = 0, comp_fallbacks = 0;
    for (;;) {
        rc = enqueue_population(h, st, Wh, P, D, start, len, fitness, embeds, audio, &launches);
        if (rc) return rc;
        CU(cudaStreamSynchronize(st));
        comp_fallbacks = h->hflags.as<int>()[3];
        if (!tc_after_pass(h)) break;
        ++overflowed;
    }
    h->comp_fallbacks_total += comp_fallbacks;
    // Timing of this call: the 17 cudaEventElapsedTime queries (~50 us of host time, every generation) are deferred to
    // stito_get_timing(); only what cannot be recomputed later is stored here.
    stito_timing &t = h->timing;
    memset(&t, 0, sizeof(t));
    t.launches = launches;
    t.precision = h->precision;
    t.comp_fallbacks = comp_fallbacks;
    t.act_overflow = overflowed;
    const int T = (int)(len / h->hop) + 1;
    t.encoder_flop = encoder_flops(P * ochs, T, h->n_mels);
    t.dsp_bytes = 4.0 * chs * len + 4.0 * ochs * len * P;
    t.frontend_bytes = P * (4.0 * ochs * len + (double)ochs * T * h->n_mels * 4.0);
    h->timing_pending = true;
    h->timing_scale = (double)P / (P < h->microbatch ? P : h->microbatch);  // stage times are measured on the first micro-batch
    return STITO_OK;
}

/* ---- multi-GPU: fitness all-gather over NVLink peer memory, fused into the fitness kernel (fitness.cu) ---- */
int stito_gather_export(stito_handle *h, int capacity, void *ipc_handle_out) {
    if (!h || !ipc_handle_out || capacity <= 0) return fail(STITO_EINVAL, "bad argument");
    CU(cudaSetDevice(h->device));
    if (h->gather_block) return fail(STITO_ESTATE, "gather block already exported");
    const size_t bytes = (size_t)2 * capacity * sizeof(float) + (size_t)2 * kGatherMaxWorld * sizeof(int);
    CU(cudaMalloc(&h->gather_block, bytes));
    CU(cudaMemset(h->gather_block, 0, bytes));
    CU(cudaMalloc((void **)&h->gather_done, sizeof(int)));
    CU(cudaMemset(h->gather_done, 0, sizeof(int)));
    CU(cudaDeviceSynchronize());
    h->gather_capacity = capacity;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t hd;
    CU(cudaIpcGetMemHandle(&hd, h->gather_block));
    memcpy(ipc_handle_out, &hd, sizeof(hd));
    return STITO_OK;
}

int stito_gather_attach(stito_handle *h, int rank, int world, const void *ipc_handles) {
    if (!h || !ipc_handles) return fail(STITO_EINVAL, "NULL argument");
    if (!h->gather_block) return fail(STITO_ESTATE, "stito_gather_export has not been called");
    if (world < 1 || world > kGatherMaxWorld || rank < 0 || rank >= world) return fail(STITO_EINVAL, "rank %d / world %d out of range (max %d ranks)", rank, world, kGatherMaxWorld);
    CU(cudaSetDevice(h->device));
    GatherPeers &g = h->gather_peers;
    g.world = world; g.rank = rank; g.capacity = h->gather_capacity;
    for (int r = 0; r < world; ++r) {
        void *base = h->gather_block;
        if (r != rank) {
            cudaIpcMemHandle_t hd;
            memcpy(&hd, (const char *)ipc_handles + (size_t)r * sizeof(hd), sizeof(hd));
            CU(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
            h->gather_opened[r] = base;
        }
        g.buf[r] = reinterpret_cast<float *>(base);
        g.flag[r] = reinterpret_cast<int *>(reinterpret_cast<char *>(base) + (size_t)2 * g.capacity * sizeof(float));
    }
    h->gather_attached = true;
    h->gather_epoch = 0;
    return STITO_OK;
}

int stito_eval_population_gather(stito_handle *h, const double *W, int P_local, int D, int64_t start, int64_t len, int lo,
                                 int P_total, float *fitness_all, void *stream) {
    if (!h || (!W && P_local > 0) || !fitness_all) return fail(STITO_EINVAL, "NULL argument");
    if (!h->gather_attached) return fail(STITO_ESTATE, "stito_gather_attach has not been called");
    if (P_local < 0 || lo < 0 || lo + P_local > P_total || P_total > h->gather_capacity)
        return fail(STITO_EINVAL, "shard [%d, %d) of a population of %d does not fit (capacity %d)", lo, lo + P_local, P_total, h->gather_capacity);
    if (D != h->chain.num_w) return fail(STITO_EINVAL, "parameter vectors have %d entries, chain expects %d", D, h->chain.num_w);
    if (h->in_chs == 0) return fail(STITO_ESTATE, "stito_set_input has not been called");
    if (!h->has_encoder) return fail(STITO_ESTATE, "handle was created without encoder weights");
    if (!h->has_target) return fail(STITO_ESTATE, "no target set (stito_set_target / stito_set_target_embeds)");
    if (start < 0 || len <= 0 || start + len > h->in_cap) return fail(STITO_EINVAL, "view outside the padded input");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const int E = h->embed_dim;
    int launches = 0, overflowed = 0, comp_fallbacks = 0;
    if (P_local > 0) {
        const double *Wh = nullptr;
        int rc = fetch_W(h, W, P_local, D, st, &Wh);
        if (rc) return rc;
        // phase 1: embeddings of my shard (redone if the activation scales had to be re-calibrated); nothing is published yet
        for (;;) {
            rc = enqueue_population(h, st, Wh, P_local, D, start, len, nullptr, nullptr, nullptr, &launches);
            if (rc) return rc;
            CU(cudaStreamSynchronize(st));
            comp_fallbacks = h->hflags.as<int>()[3];
            if (!tc_after_pass(h)) break;
            ++overflowed;
        }
    }
    // phase 2: fitness of my shard, stored straight into every rank's gather buffer; wait for everybody's; read the vector
    const int epoch = ++h->gather_epoch, parity = epoch & 1;
    const float *mid_all = h->emb.as<float>(), *side_all = mid_all + (size_t)P_local * E;
    GatherPeers &g = h->gather_peers;
    CU(launch_fitness_gather(st, mid_all, side_all, h->target.as<float>(), h->target.as<float>() + E, P_local, E, lo, g, parity,
                             epoch, h->gather_done, g.flag[g.rank], &launches));
    CU(cudaMemcpyAsync(fitness_all, g.buf[g.rank] + (size_t)parity * g.capacity, (size_t)P_total * sizeof(float), cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    h->comp_fallbacks_total += comp_fallbacks;
    stito_timing &t = h->timing;
    const bool had_pass = P_local > 0;
    memset(&t, 0, sizeof(t));
    t.launches = launches;
    t.precision = h->precision;
    t.comp_fallbacks = comp_fallbacks;
    t.act_overflow = overflowed;
    if (had_pass) {
        const int chs = h->in_chs, ochs = out_channels(h->chain, chs);
        const int T = (int)(len / h->hop) + 1;
        t.encoder_flop = encoder_flops(P_local * ochs, T, h->n_mels);
        t.dsp_bytes = 4.0 * chs * len + 4.0 * ochs * len * P_local;
        t.frontend_bytes = P_local * (4.0 * ochs * len + (double)ochs * T * h->n_mels * 4.0);
        h->timing_scale = (double)P_local / (P_local < h->microbatch ? P_local : h->microbatch);
    }
    h->timing_pending = had_pass;
    return STITO_OK;
}

int stito_process(stito_handle *h, const float *x, int chs, int64_t L, const double *W, int P, int D,
                  int final_normalize, float *y, void *stream) {
    if (!h || !x || !y || (!W && h->chain.num_w > 0)) return fail(STITO_EINVAL, "NULL argument");
    if (chs != 1 && chs != 2) return fail(STITO_EINVAL, "Invalid number of channels: %d", chs);
    if (P <= 0 || L <= 0) return fail(STITO_EINVAL, "empty input");
    if (D != h->chain.num_w) return fail(STITO_EINVAL, "parameter vectors have %d entries, chain expects %d", D, h->chain.num_w);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const double *Wh = nullptr;
    int rc = D > 0 ? fetch_W(h, W, P, D, st, &Wh) : STITO_OK;
    if (rc) return rc;
    CU(h->xin.ensure((size_t)chs * L * sizeof(float)));
    CU(cudaMemcpyAsync(h->xin.p, x, (size_t)chs * L * sizeof(float), cudaMemcpyDefault, st));
    SigView in{h->xin.as<float>(), 0, L};
    const int ochs = out_channels(h->chain, chs);
    const bool ydev = is_device_ptr(y);
    int launches = 0;
    CU(cudaMemsetAsync(h->flags.as<int>() + 2, 0, (kNumFlags - 2) * sizeof(int), st));
    for (int p0 = 0; p0 < P; p0 += h->microbatch) {
        const int pb = (P - p0) < h->microbatch ? (P - p0) : h->microbatch;
        const float *res = nullptr;
        const unsigned *pk = nullptr;
        int ych = 0;
        h->hparams_cursor = 0;  // the previous micro-batch was synchronised below
        rc = run_chain(h, st, in, chs, L, Wh ? Wh + (size_t)p0 * D : nullptr, pb, D, &res, &pk, &ych, &launches);
        if (rc) return rc;
        float *dst = y + (size_t)p0 * ochs * L;
        const size_t bytes = (size_t)pb * ych * L * sizeof(float);
        if (final_normalize) {
            if (ydev) {
                CU(launch_normalize(st, res, pk, dst, pb, ych, L, &launches));
            } else {
                CU(h->audio[1].ensure(bytes));
                CU(h->audio[0].ensure(bytes));
                float *tmp = (res == h->audio[0].as<float>()) ? h->audio[1].as<float>() : h->audio[0].as<float>();
                CU(launch_normalize(st, res, pk, tmp, pb, ych, L, &launches));
                CU(cudaMemcpyAsync(dst, tmp, bytes, cudaMemcpyDeviceToHost, st));
            }
        } else {
            CU(cudaMemcpyAsync(dst, res, bytes, cudaMemcpyDefault, st));
        }
        CU(cudaStreamSynchronize(st));
    }
    CU(cudaMemcpyAsync(h->hflags.p, h->flags.p, kNumFlags * sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    h->timing_pending = false;
    memset(&h->timing, 0, sizeof(h->timing));
    h->timing.launches = launches;
    h->timing.precision = h->precision;
    h->timing.comp_fallbacks = h->hflags.as<int>()[3];
    h->comp_fallbacks_total += h->timing.comp_fallbacks;
    return STITO_OK;
}

int stito_embed(stito_handle *h, const float *x, int B, int chs, int64_t L, int peak_normalize, float *mid,
                float *side, void *stream) {
    if (!h || !x || !mid || !side) return fail(STITO_EINVAL, "NULL argument");
    if (chs != 1 && chs != 2) return fail(STITO_EINVAL, "Invalid number of channels: %d", chs);
    if (B <= 0) return fail(STITO_EINVAL, "empty batch");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const int E = h->embed_dim;
    const float *xd = x;
    if (!is_device_ptr(x)) {
        CU(h->xin.ensure((size_t)B * chs * L * sizeof(float)));
        CU(cudaMemcpyAsync(h->xin.p, x, (size_t)B * chs * L * sizeof(float), cudaMemcpyHostToDevice, st));
        xd = h->xin.as<float>();
    }
    CU(h->emb.ensure((size_t)2 * B * E * sizeof(float)));
    float *mid_d = h->emb.as<float>(), *side_d = mid_d + (size_t)B * E;
    CU(h->peaks.ensure((size_t)B * sizeof(unsigned)));
    int launches = 0;
    SigView all{xd, (int64_t)chs * L, L};
    if (peak_normalize) {
        CU(cudaMemsetAsync(h->peaks.p, 0, (size_t)B * sizeof(unsigned), st));
        CU(launch_peak(st, all, B, chs, L, h->peaks.as<unsigned>(), &launches));
    }
    int overflowed = 0;
    for (;;) {
        CU(cudaMemsetAsync(h->flags.as<int>() + 2, 0, (kNumFlags - 2) * sizeof(int), st));
        for (int b0 = 0; b0 < B; b0 += h->microbatch) {
            const int bb = (B - b0) < h->microbatch ? (B - b0) : h->microbatch;
            SigView v{xd + (size_t)b0 * chs * L, (int64_t)chs * L, L};
            int rc = encoder_forward(h, st, v, peak_normalize ? h->peaks.as<unsigned>() + b0 : nullptr, bb, chs, L,
                                     mid_d + (size_t)b0 * E, side_d + (size_t)b0 * E, &launches, false);
            if (rc) return rc;
        }
        CU(cudaMemcpyAsync(mid, mid_d, (size_t)B * E * sizeof(float), cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(side, side_d, (size_t)B * E * sizeof(float), cudaMemcpyDefault, st));
        CU(cudaMemcpyAsync(h->hflags.p, h->flags.p, kNumFlags * sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (!tc_after_pass(h)) break;  // else: activation scales re-calibrated -> once more
        ++overflowed;
    }
    h->timing_pending = false;
    memset(&h->timing, 0, sizeof(h->timing));
    h->timing.launches = launches;
    h->timing.precision = h->precision;
    h->timing.act_overflow = overflowed;
    return STITO_OK;
}

int stito_logmel(stito_handle *h, const float *x, int B, int chs, int64_t L, float *out, void *stream) {
    if (!h || !x || !out) return fail(STITO_EINVAL, "NULL argument");
    if (!h->has_encoder) return fail(STITO_ESTATE, "handle was created without encoder weights");
    if (chs != 1 && chs != 2) return fail(STITO_EINVAL, "Invalid number of channels: %d", chs);
    CU(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->own_stream;
    const float *xd = x;
    if (!is_device_ptr(x)) {
        CU(h->xin.ensure((size_t)B * chs * L * sizeof(float)));
        CU(cudaMemcpyAsync(h->xin.p, x, (size_t)B * chs * L * sizeof(float), cudaMemcpyHostToDevice, st));
        xd = h->xin.as<float>();
    }
    const int T = (int)(L / h->hop) + 1;
    const size_t fb = (size_t)B * chs * T * h->n_mels * sizeof(float);
    CU(h->feat.ensure(fb));
    int launches = 0;
    SigView v{xd, (int64_t)chs * L, L};
    CU(launch_logmel(st, v, nullptr, B, chs, L, T, h->ft, h->feat.as<float>(), &launches));
    CU(cudaMemcpyAsync(out, h->feat.p, fb, cudaMemcpyDefault, st));
    CU(cudaStreamSynchronize(st));
    h->timing.launches = launches;
    return STITO_OK;
}

/* Host-side setup pieces of the convolution reverb, for the CPU tests: the 12 x 1023 octave-band FIR bank
 * (scipy.signal.firwin restated) and n samples of the seeded white noise.  No GPU involved. */
int stito_crv_host_filterbank(double sample_rate, float *out) {
    if (!out || !(sample_rate > 36100.0)) return fail(STITO_EINVAL, "bad argument");
    convreverb_host_filterbank(sample_rate, out);
    return STITO_OK;
}
int stito_crv_host_noise(uint64_t seed, int64_t n, float *out) {
    if (!out || n < 0) return fail(STITO_EINVAL, "bad argument");
    convreverb_host_noise(seed, (size_t)n, out);
    return STITO_OK;
}

/* Host-side design of STITO_FX_LTI_COMPRESSOR's smoothing filter, for the CPU tests: out = {alpha, 1 - alpha, wrap}. */
int stito_lticomp_host_design(double sample_rate, int64_t L, float attack_ms, double *out) {
    if (!out || !(sample_rate > 0) || L <= 0 || !(attack_ms > 0)) return fail(STITO_EINVAL, "bad argument");
    LtiCompParams q;
    lticomp_design(sample_rate, L, 0.0f, 1.0f, attack_ms, 1.0f, 0.0f, &q);
    out[0] = q.alpha;
    out[1] = q.b0;
    out[2] = q.wrap;
    return STITO_OK;
}

int stito_get_timing(const stito_handle *hc, stito_timing *out) {
    if (!hc || !out) return fail(STITO_EINVAL, "NULL argument");
    stito_handle *h = const_cast<stito_handle *>(hc);
    if (h->timing_pending) {  // events of the last stito_eval_population (it synchronised the stream before returning)
        stito_timing &t = h->timing;
        const float scale = (float)h->timing_scale;
        cudaEventElapsedTime(&t.ms_dsp, h->ev[0], h->ev[1]);
        cudaEventElapsedTime(&t.ms_frontend, h->ev[2], h->ev[3]);
        cudaEventElapsedTime(&t.ms_encoder, h->ev[3], h->ev[4]);
        cudaEventElapsedTime(&t.ms_fitness, h->ev[5], h->ev[6]);
        cudaEventElapsedTime(&t.ms_total, h->ev[0], h->ev[6]);
        t.ms_dsp *= scale; t.ms_frontend *= scale; t.ms_encoder *= scale;
        for (int l = 0; l < 12; ++l) {
            cudaEventElapsedTime(&t.ms_conv[l], h->ev_conv[l], h->ev_conv[l + 1]);
            t.ms_conv[l] *= scale;
        }
        cudaGetLastError();  // precision 0 records no per-layer events
        h->timing_pending = false;
    }
    *out = h->timing;
    return STITO_OK;
}

}  // extern "C"
