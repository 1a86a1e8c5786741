/*
 * stito.h -- C ABI of libstito.so, the B200 (sm_100a) implementation of st-ito's
 * CMA-ES population evaluation path.
 *
 * Every entry point replaces a piece of the reference's Python path (citations are
 * file:line in csteinmetz1/st-ito); the reference-side binding is a ctypes stub,
 * shown in INTEGRATION.md and implemented in st_ito_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types;
 *   - every function returns 0 on success or a negative STITO_E* code; the message
 *     is available from stito_last_error() (thread-local);
 *   - audio is float32, row-major [chs, L] or [P, chs, L]; parameter vectors are
 *     float64 on [0,1] exactly as pycma hands them to evaluate();
 *   - data pointers may be HOST or DEVICE pointers (detected with
 *     cudaPointerGetAttributes); the library never frees caller memory;
 *   - `stream` is a cudaStream_t passed as void*.  NULL = the handle's own (non-blocking) stream, which is NOT
 *     ordered against the legacy default stream: callers whose device inputs are produced on the legacy default
 *     stream pass cudaStreamLegacy ((void*)1), on a per-thread default stream cudaStreamPerThread ((void*)2).
 *     Every call returns after its work on `stream` has completed (outputs are valid on return);
 *   - one handle = one host thread at a time (the reference's plugin objects are
 *     stateful and not re-entrant either, style_transfer.py:76-92).
 */
#ifndef STITO_H
#define STITO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define STITO_API __attribute__((visibility("default")))
#else
#define STITO_API
#endif

#define STITO_OK 0
#define STITO_EINVAL (-1)   /* malformed descriptor / argument                       */
#define STITO_ECUDA (-2)    /* CUDA runtime error (message holds cudaGetErrorString) */
#define STITO_ESTATE (-3)   /* call order: no input / target / encoder set           */
#define STITO_ENOMEM (-4)

#define STITO_MAX_FX 8
#define STITO_MAX_FX_PARAMS 32
#define STITO_EMBED_DIM_MAX 1024

/* Effects of the reference's built-in ("Basic*") plugin family, st_ito/effects.py:800-959. */
enum stito_fx_kind {
    STITO_FX_EQ = 0,         /* BasicParametricEQ  effects.py:800-873 (18 parameters)            */
    STITO_FX_COMPRESSOR = 1, /* BasicCompressor    effects.py:876-897 (4)                        */
    STITO_FX_DISTORTION = 2, /* BasicDistortion    effects.py:900-914 (2)                        */
    STITO_FX_DELAY = 3,      /* BasicDelay         effects.py:917-934 (3)                        */
    STITO_FX_REVERB = 4,     /* BasicReverb        effects.py:937-959 (4)                        */
    /* Noise-shaped convolution reverb: the arithmetic of apply_reverb, effects.py:558-620 -> dasp-pytorch's
     * noise_shaped_reverberation (25 parameters used raw: 12 band gains, 12 band decays, mix); num_channels must be 2.
     * iopt[0] = impulse-response length in samples (65536 = dasp default; 96000 = the 2 s IR of BASELINE config 4),
     * iopt[1] = seed of the white noise (the reference draws fresh noise per call; here it is fixed per plugin). */
    STITO_FX_CONV_REVERB = 5,
    /* Compressor with LTI gain smoothing: the arithmetic of apply_compressor, effects.py:623-648 -> dasp-pytorch's
     * compressor (6 parameters: threshold_db -60..0, ratio 1..20, attack_ms 0.1..250, release_ms 10..2000 [accepted and
     * unused, as upstream], knee_db 1..24, makeup_gain_db 0..24).  The side-chain is the SUM of the channels the plugin
     * sees (num_channels 2: stereo-linked; 1: every channel on its own).  iopt[0] = look-ahead in samples (512 at
     * effects.py:646; 0 at dsp.py:66-75). */
    STITO_FX_LTI_COMPRESSOR = 6
};

/* One entry of the reference's ordered `plugins` dict (run_optim.py:376-407,
 * style_transfer.py:17-42), flattened.  Effect parameters are listed in the order of the
 * plugin's `.parameters` dict; `our_bypass` slots consume a w index but map to nothing
 * (style_transfer.py:88-92 never skips the plugin). */
typedef struct {
    int32_t kind;                             /* enum stito_fx_kind                                   */
    int32_t num_channels;                     /* plugins[name]["num_channels"]: 1 or 2                */
    int32_t num_params;                       /* effect parameters described below (excl. our_bypass) */
    int32_t w_index[STITO_MAX_FX_PARAMS];     /* index into w, or -1: use fixed_raw                   */
    double fixed_raw[STITO_MAX_FX_PARAMS];    /* raw_value of fixed_parameters (Parameter.set_value)  */
    int32_t iopt[4];                          /* per-kind integer options (see enum stito_fx_kind)    */
} stito_fx_desc;

typedef struct {
    int32_t num_fx;
    int32_t num_w;                /* D: length of a parameter vector                          */
    int32_t normalize_stages;     /* process_audio(normalize_stages=...) style_transfer.py:106 */
    int32_t reserved;
    double sample_rate;
    stito_fx_desc fx[STITO_MAX_FX];
} stito_chain_desc;

/* AFx-Rep encoder (st_ito/models/panns.py:121-207) weights, HOST pointers, float32, in the
 * layouts of the reference state_dict.  Index 2*b is conv_block{b+1}.conv1
 * (+bn1), 2*b+1 is conv2 (+bn2).  The library folds eval-mode BatchNorm into the convolutions. */
typedef struct {
    int32_t n_fft, hop, n_mels, embed_dim; /* 2048, 1024, 128, 512                              */
    float bn_eps;                          /* 1e-5                                               */
    int32_t reserved;
    const float *conv_w[12];               /* [Cout, Cin, 3, 3]                                  */
    const float *bn_weight[12], *bn_bias[12], *bn_mean[12], *bn_var[12]; /* [Cout]               */
    const float *fc_mid_w, *fc_mid_b;      /* [embed_dim, 2048], [embed_dim]                     */
    const float *fc_side_w, *fc_side_b;
    const float *mel_w;                    /* logmel_extractor.melW [n_fft/2+1, n_mels]          */
} stito_encoder_weights;

typedef struct stito_handle stito_handle;

/* Per-stage device times of the last stito_eval_population call (CUDA events on the launch
 * stream), the kernels it launched and the algorithmic work it did. */
typedef struct {
    float ms_dsp, ms_frontend, ms_encoder, ms_fitness, ms_total;
    int32_t launches;            /* kernels of this library launched by the call                */
    int32_t precision;           /* encoder arithmetic used: 0 = fp32 SIMT, 1 = fp16x3 tcgen05  */
    double encoder_flop;         /* 2*MACs of the 12 convolutions + heads                       */
    double dsp_bytes, frontend_bytes; /* algorithmic HBM bytes                      */
    float ms_conv[12];           /* per conv layer                                              */
    /* -- appended in ABI 2.0.0 -- */
    int32_t comp_fallbacks;      /* compressor super-blocks whose time-parallel Newton iteration did not converge
                                    and were recomputed by the exact serial loop (results stay correct)          */
    int32_t act_overflow;        /* times the call was redone because the per-layer storage scales of the fp16x3
                                    encoder were (re-)calibrated: 1 on a handle's first encoder pass, > 0 later only
                                    when an activation left the fp16 range (results are those of the last pass)  */
} stito_timing;

/* Create an evaluator for (chain, encoder) on CUDA device `device`.  `weights` may be NULL:
 * the handle then only renders audio (plugin.process / process_audio).
 * Replaces: load_plugins + model construction, style_transfer.py:17-42, utils.py:511-551. */
STITO_API int stito_create(const stito_chain_desc *chain, const stito_encoder_weights *weights, int device,
                 stito_handle **out);
STITO_API void stito_destroy(stito_handle *h);

/* Swap the effect chain (keeps weights, input and target). */
STITO_API int stito_set_chain(stito_handle *h, const stito_chain_desc *chain);

/* Encoder arithmetic: 0 = fp32 CUDA cores (bit-for-bit conv semantics of the oracle up to
 * summation order), 1 = error-compensated fp16x3 on tcgen05 tensor cores (default).  Mode 1 stores activations as
 * fp16 hi/lo pairs of value * 2^shift with a per-layer shift that the library calibrates from the measured
 * per-layer maxima on the handle's first encoder pass and widens whenever a later input would overflow. */
STITO_API int stito_set_precision(stito_handle *h, int precision);

/* Upload the input waveform x[chs, L].  The device copy is zero-padded to max(L, min_len) so
 * that evaluate()'s "pad to 262144" policy (style_transfer.py:518) is a view, not a copy.
 * Replaces: x = input_audio.clone(), style_transfer.py:623. */
STITO_API int stito_set_input(stito_handle *h, const float *x, int chs, int64_t L, int64_t min_len);

/* Target embeddings: either computed here from audio (peak-normalise per item, encoder, L2
 * normalise: embed_func(target_audio), style_transfer.py:456-460) or supplied directly. */
STITO_API int stito_set_target(stito_handle *h, const float *target, int chs, int64_t L);
STITO_API int stito_set_target_embeds(stito_handle *h, const float *mid, const float *side, int embed_dim);

/* evaluate(W, x, ...) of style_transfer.py:474-573 for one population:
 *   for each w: process_audio(x[:, start:start+len], w)  -> peak-normalised audio
 *   get_param_embeds -> L2-normalised mid/side embeddings -> fitness = mean_k(-cos(out_k, tgt_k)).
 * W is [P, D] float64 row-major (host or device).  Outputs (each nullable, host or device):
 *   fitness [P]; embeds [2, P, embed_dim] (mid then side); audio [P, out_chs, len]. */
STITO_API int stito_eval_population(stito_handle *h, const double *W, int P, int D, int64_t start,
                          int64_t len, float *fitness, float *embeds, float *audio, void *stream);

/* ---- multi-GPU (one process per GPU; the population is sharded over the ranks) --------------------------------
 * Every rank's replica of the CMA-ES needs all P fitness values each generation.  Instead of a NCCL all-gather after the
 * fitness kernel, the fitness kernel itself stores each value into the gather buffer of EVERY rank through NVLink peer
 * memory and publishes an epoch flag; a one-warp kernel waits for all ranks' flags.  Setup, once per handle:
 *   stito_gather_export(h, capacity >= P_total, handle[64])   -> a cudaIpcMemHandle_t of this rank's gather block;
 *   (exchange the 64-byte handles between the ranks, e.g. torch.distributed.all_gather_object)
 *   stito_gather_attach(h, rank, world, handles[world][64])    -> opens the peers' blocks (world <= 16, same node).
 * Per generation: stito_eval_population_gather scores the shard W [P_local][D] = candidates [lo, lo + P_local) of a population
 * of P_total and returns ALL P_total fitness values (host or device pointer).  Collective: every rank must call it once per
 * generation (P_local may be 0); a rank that never arrives makes the others fail after a bounded wait (20 s). */
STITO_API int stito_gather_export(stito_handle *h, int capacity, void *ipc_handle_out);
STITO_API int stito_gather_attach(stito_handle *h, int rank, int world, const void *ipc_handles);
STITO_API int stito_eval_population_gather(stito_handle *h, const double *W, int P_local, int D, int64_t start, int64_t len,
                                           int lo, int P_total, float *fitness_all, void *stream);

/* process_audio (style_transfer.py:45-115) for P parameter vectors on an arbitrary signal:
 * x[chs, L] -> y[P, out_chs, L].  final_normalize=1 applies the closing peak normalisation
 * (style_transfer.py:113); 0 returns the raw chain output (what plugin.process returns). */
STITO_API int stito_process(stito_handle *h, const float *x, int chs, int64_t L, const double *W, int P,
                  int D, int final_normalize, float *y, void *stream);

/* Channels produced by the chain for a chs-channel input (mono is up-mixed by 2-channel plugins,
 * style_transfer.py:94-95). */
STITO_API int stito_out_channels(const stito_handle *h, int chs);

/* Cnn14.forward (panns.py:209-281) on x[B, chs, L]; peak_normalize=1 first divides every item by
 * its peak (utils.py:473-474).  Writes RAW (un-normalised) mid/side embeddings [B, embed_dim]. */
STITO_API int stito_embed(stito_handle *h, const float *x, int B, int chs, int64_t L, int peak_normalize,
                float *mid, float *side, void *stream);

/* Normalised log-mel features of x[B, chs, L] (panns.py:219-245): out [B*chs, T, n_mels],
 * T = L / hop + 1, row order b0-mid, b0-side, b1-mid, ... */
STITO_API int stito_logmel(stito_handle *h, const float *x, int B, int chs, int64_t L, float *out,
                 void *stream);

STITO_API int stito_get_timing(const stito_handle *h, stito_timing *out);

/* Host-side setup pieces of STITO_FX_CONV_REVERB, exported for the CPU test-suite (no GPU involved): the 12 x 1023
 * octave-band FIR bank (scipy.signal.firwin restated; out [12][1023]) and n samples of the seeded white noise. */
STITO_API int stito_crv_host_filterbank(double sample_rate, float *out);
STITO_API int stito_crv_host_noise(uint64_t seed, int64_t n, float *out);
/* Host-side design of STITO_FX_LTI_COMPRESSOR's smoothing filter (same purpose): out[0] = alpha = the float32
 * exp(-log(9) / (fs * attack_ms / 1000)) of dasp-pytorch's compressor, out[1] = float32(1 - alpha), out[2] = the factor
 * of the frequency-sampling wrap-around, y[-1] = y0[L-1] * out[2] (n_fft = 2^ceil(log2(2L - 1))). */
STITO_API int stito_lticomp_host_design(double sample_rate, int64_t L, float attack_ms, double *out);

/* ---- native CMA-ES (host, fp64): what run_es obtains from pycma, style_transfer.py:614-673 -----------------------------
 * cma.CMAEvolutionStrategy(x0, sigma0, {"bounds": [lower, upper], "popsize": P}) -> stito_cma_create (lower >= upper: no
 * bounds); es.ask() -> stito_cma_ask (X [P][D], feasible); es.tell(X, f) -> stito_cma_tell; es.result -> stito_cma_result.
 * (mu/mu_w, lambda)-CMA-ES, rank-one + rank-mu update, CSA, pycma's BoxConstraintsLinQuadTransformation.  Gaussian
 * draws: element n of the stream = Box-Muller (cos, sin branches alternate) of the uniforms splitmix64(key + 2m),
 * splitmix64(key + 2m + 1), key = splitmix64(seed).  These return STITO_E* codes and set no message. */
typedef struct stito_cma stito_cma;
STITO_API int stito_cma_create(const double *x0, int D, double sigma0, int popsize, double lower, double upper,
                               uint64_t seed, stito_cma **out);
STITO_API void stito_cma_destroy(stito_cma *es);
STITO_API int stito_cma_eig(const double *A, int n, double *V, double *d); /* symmetric eigensolver of the update (tests) */
STITO_API int stito_cma_ask(stito_cma *es, double *X);
STITO_API int stito_cma_geno(const stito_cma *es, double *G); /* search points of the last ask before the box map (tests) */
STITO_API int stito_cma_tell(stito_cma *es, const double *X, const double *f);
STITO_API int stito_cma_result(const stito_cma *es, double *xbest, double *fbest, int *has_best, double *xfavorite,
                               double *sigma, double *stds, int64_t *evals_best, int64_t *evaluations,
                               int64_t *iterations, double *axis_ratio, double *last_f_range);

/* Thread-local message of the last failing call. */
STITO_API const char *stito_last_error(void);

/* Library/ABI version: major*10000 + minor*100 + patch. */
STITO_API int stito_version(void);

#ifdef __cplusplus
}
#endif
#endif /* STITO_H */
