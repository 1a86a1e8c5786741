// Internal declarations shared by the kernels and the C-ABI layer of libstito.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "stito.h"

namespace stito {

// A read-only view of audio: element (p, c, n) lives at base[p*stride_p + c*stride_c + n].
// stride 0 broadcasts: the shared input waveform has stride_p == 0; the mono->stereo
// up-mix of style_transfer.py:94-95 is stride_c == 0.
struct SigView {
    const float *base;
    int64_t stride_p;
    int64_t stride_c;
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), once per (kernel, device) (encoder_tc.cu)
cudaError_t ensure_dyn_smem(const void *func, int bytes);

// ---------------------------------------------------------------- DSP (dsp_kernels.cu)
constexpr int kEqChunk = 512;  // samples per time-chunk of the chunk-parallel biquad cascade
constexpr int kEqStates = 12;  // 6 biquads x 2 DF-II-transposed states

struct CompParams { float cte_at, cte_rl, thr, thr_inv, expo; };
struct DistParams { float drive, out_gain; };
struct DelayParams { int d; float feedback, mix, dry; };
struct ReverbParams { float damp, fb, wet1, wet2, dry; };
struct ConvRevParams { float gain[12]; float decay[12]; float mix; };  // raw [0, 1] values (effects.py:564-588)
struct ReverbGeom {
    int comb_size[2][8], comb_off[2][8];
    int ap_size[2][4], ap_off[2][4];
    int total;  // floats of delay-line storage for two channels
    int block;  // samples per block: multiple of 32, <= shortest all-pass line, <= 256
};

// All launch_* enqueue on `st`, bump *launches by the number of kernels launched and return
// the cudaError_t of the launch.
cudaError_t launch_eq(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                      int64_t L, const double *coefs /*[P][6][5]*/, double *scratch_f,
                      double *scratch_s, unsigned *out_peak, int *launches);
size_t eq_scratch_doubles(int P, int chs, int64_t L);  // per scratch buffer

// *noconv (nullable, device) counts super-blocks whose Newton iteration did not converge and were redone serially
cudaError_t launch_compressor(cudaStream_t st, SigView in, const float *in_peak, float *out, int P,
                              int chs, int64_t L, const CompParams *prm, unsigned *out_peak, int *noconv,
                              int *ready, int *launches);
cudaError_t launch_distortion(cudaStream_t st, SigView in, const float *in_peak, float *out, int P,
                              int chs, int64_t L, const DistParams *prm, unsigned *out_peak,
                              int *launches);
cudaError_t launch_delay(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                         int64_t L, const DelayParams *prm, int max_d, unsigned *out_peak,
                         int *launches);
// stereo != 0: one joint stereo Freeverb per candidate (chs must be 2); else one mono Freeverb
// per (candidate, channel).
cudaError_t launch_reverb(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                          int stereo, int64_t L, const ReverbGeom &g, const ReverbParams *prm,
                          unsigned *out_peak, const int *ready, int sm_budget, int *launches);
// Streaming hand-off compressor -> Freeverb: `ready` = one int per (stream, 32768-sample granule), zeroed before the
// launch; the compressor sets a flag when that super-block of its output is in memory, the reverb (launched on a second
// stream, concurrently resident) waits for the flags of the samples it is about to read.
bool reverb_can_stream(const ReverbGeom &g);
constexpr int kStreamGranule = 32768;
cudaError_t launch_copy(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs,
                        int64_t L, unsigned *out_peak, int *launches);
// peak[i] = max |x[i, :, :]| as float bits (buffer must be zeroed first).
cudaError_t launch_peak(cudaStream_t st, SigView in, int P, int chs, int64_t L, unsigned *peak,
                        int *launches);
// y = x / clip(peak, 1e-8)
cudaError_t launch_normalize(cudaStream_t st, const float *x, const unsigned *peak, float *y, int P,
                             int chs, int64_t L, int *launches);
void reverb_geometry(double sample_rate, ReverbGeom *g);

// Noise-shaped convolution reverb (convreverb.cu).  The state caches what does not depend on the candidate -- the
// filtered-noise bands of (sample rate, IR length, seed) and the FFT twiddles -- plus the spectra work buffers.
struct ConvReverbState {
    float *bands = nullptr;     // [2][12][n_ir]
    float2 *twiddle = nullptr;  // [16384]
    float4 *H = nullptr;        // [P][K][8193] candidate IR spectra
    float2 *X = nullptr;        // [P][blocks][16384] input / output spectra
    size_t H_cap = 0, X_cap = 0;
    int n_ir = 0, seed = 0;
    double sample_rate = 0.0;
};
cudaError_t convreverb_prepare(cudaStream_t st, ConvReverbState *s, double sample_rate, int n_ir, int seed);
// in: stereo view (mono is up-mixed by stride_c == 0); out [P][2][L]
cudaError_t launch_convreverb(cudaStream_t st, ConvReverbState *s, SigView in, const float *in_peak, float *out, int P,
                              int64_t L, const ConvRevParams *prm, unsigned *out_peak, int *launches);
void convreverb_release(ConvReverbState *s);
void convreverb_host_filterbank(double sample_rate, float *out);
void convreverb_host_noise(uint64_t seed, size_t count, float *out);

// Compressor with LTI gain smoothing (lticomp.cu).  link = channels summed into one side-chain (1 or 2; chs % link == 0).
struct LtiCompParams {
    double alpha, b0, ln_alpha;  // one-pole smoothing y = alpha y + b0 g (alpha, b0 = float32 values of the reference)
    double wrap;                 // y[-1] = y0[L-1] * wrap: periodic steady state of the frequency-sampled filter
    float thr, ratio, knee, makeup;
};
void lticomp_design(double sample_rate, int64_t L, float threshold_db, float ratio, float attack_ms, float knee_db,
                    float makeup_db, LtiCompParams *q);
size_t lticomp_scratch_bytes(int P, int chs, int link, int64_t L);
cudaError_t launch_lticomp(cudaStream_t st, SigView in, const float *in_peak, float *out, int P, int chs, int link, int64_t L,
                           int lookahead, const LtiCompParams *prm, void *scratch, unsigned *out_peak, int *launches);

// ------------------------------------------------------- front-end (frontend_kernels.cu)
struct FrontendTables {
    float2 *twiddle;   // [n_fft] exp(-2*pi*i*k/n_fft)
    float *window;     // [n_fft] periodic Hann
    int *mel_start;    // [n_mels] first FFT bin with a non-zero weight
    int *mel_count;    // [n_mels]
    int *mel_off;      // [n_mels] offset into mel_wt
    float *mel_wt;     // compacted non-zero columns of melW
    int n_fft, hop, n_mels;
};
// items = B; feat [B*chs][T][n_mels]; peak nullable (per item, float bits): samples are divided by
// clip(peak, 1e-8) on load.
cudaError_t launch_logmel(cudaStream_t st, SigView in, const unsigned *peak, int B, int chs, int64_t L,
                          int T, const FrontendTables &tb, float *feat, int *launches);

// ------------------------------------------------------------ encoder (encoder_simt.cu)
struct ConvLayer {
    int cin, cout;
    float *w;      // fp32 [9][cin][cout], BatchNorm scale folded in
    float *bias;   // fp32 [cout], BatchNorm shift
    // tensor-core operands (encoder_tc.cu): fp16 hi/lo split of w * 2^shift, [cout][9*cin] K-major
    void *w_hi, *w_lo;
    float w_unscale;  // 2^-shift
    // Winograd F(2x2,3x3) operands (deep layers only): fp16 hi/lo split of (G g G^T) * 2^shift, [16][cout][cin]
    void *u_hi = nullptr, *u_lo = nullptr;
    float u_unscale = 1.0f;
};
struct EncoderDev {
    ConvLayer conv[12];
    float *fc_w[2];  // [embed_dim][2048] (reference layout)
    float *fc_b[2];
    int embed_dim;
};
// x [N][H][W] (C == 1) -> y [N][H][W][64], relu(conv + bias)
cudaError_t launch_conv_first(cudaStream_t st, const float *x, const ConvLayer &l, float *y, int N, int H,
                              int W, int *launches);
// NHWC fp32 3x3 conv + bias + relu on CUDA cores
cudaError_t launch_conv_simt(cudaStream_t st, const float *x, const ConvLayer &l, float *y, int N, int H,
                             int W, int *launches);
// 2x2 average pool, floor semantics, NHWC
cudaError_t launch_avgpool(cudaStream_t st, const float *x, float *y, int N, int H, int W, int C,
                           int *launches);
// x [N][H][W][C] -> pooled [N][C]: mean over W, then max over H + mean over H (panns.py:262-266)
cudaError_t launch_global_pool(cudaStream_t st, const float *x, float *y, int N, int H, int W, int C,
                               int *launches);
// heads: pooled [B*chs][2048] -> mid [B][E], side [B][E] (panns.py:269-279; mono: side = mid)
cudaError_t launch_heads(cudaStream_t st, const float *pooled, const EncoderDev &enc, int B, int chs,
                         float *mid, float *side, int *launches);

// ------------------------------------------------------------------ fitness (fitness.cu)
// raw mid/side [B][E] -> NaN scrub (utils.py:492-497) -> L2 normalise (utils.py:500-501), in place
cudaError_t launch_embed_normalize(cudaStream_t st, float *mid, float *side, int B, int E, int *flags,
                                   int *launches);
// fitness[b] = mean(-cos(mid_b, tgt_mid), -cos(side_b, tgt_side))  (style_transfer.py:544-571)
cudaError_t launch_fitness(cudaStream_t st, const float *mid, const float *side, const float *tgt_mid,
                           const float *tgt_side, int B, int E, float *fitness, int *launches);

// fitness + all-gather over peer memory (multi-GPU).  buf[r] / flag[r] point into rank r's exported gather block:
// [2][capacity] floats followed by [2][kGatherMaxWorld] ints.
constexpr int kGatherMaxWorld = 16;
struct GatherPeers {
    float *buf[kGatherMaxWorld];
    int *flag[kGatherMaxWorld];
    int world, rank, capacity;
};
cudaError_t launch_fitness_gather(cudaStream_t st, const float *mid, const float *side, const float *tgt_mid,
                                  const float *tgt_side, int n_local, int E, int lo, const GatherPeers &peers, int parity,
                                  int epoch, int *done_counter, const int *local_flags, int *launches);

}  // namespace stito
