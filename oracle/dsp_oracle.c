/*
 * oracle/dsp_oracle.c -- CPU restatement of the st-ito effect chain arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under st_ito_b200/ may link, load or call
 * this file; it exists so tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs have something to check the CUDA path
 * against (and to time on host cores).
 *
 * What is restated, and from where:
 *   - RBJ biquad design + 6-stage cascade           reference st_ito/effects.py:395-450, 453-512
 *     (scipy.signal.lfilter direct-form-II-transposed evaluation order, fp64)
 *   - feed-forward peak compressor                   reference st_ito/effects.py:876-897 ->
 *     pedalboard.Compressor -> juce::dsp::Compressor<float> + BallisticsFilter  [recollection]
 *   - Freeverb                                       reference st_ito/effects.py:937-959 ->
 *     pedalboard.Reverb -> juce::Reverb                                       [recollection]
 *   - tanh distortion + gain, feedback delay         reference st_ito/effects.py:900-934 ->
 *     pedalboard.Distortion/Gain/Delay                                        [recollection]
 *
 * Parity status: the EQ is pinned against the reference's own code (executed by
 * tests/golden/make_golden.py, fixtures in tests/golden/).  pedalboard/JUCE are
 * not vendored by the reference and are absent here, so compressor, reverb,
 * distortion and delay are "parity unpinned": this file is the written spec.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: scipy/JUCE x86-64 builds use separate mul/add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ EQ ---- */

/* type: 0 = low_shelf, 1 = peaking, 2 = high_shelf.  Outputs are already
 * divided by a0 (effects.py:447-448), b = {b0,b1,b2}, a = {1,a1,a2}. */
ORACLE_API void oracle_biquad_coefs(double gain_db, double cutoff, double q, double fs,
                                    int type, double *b, double *a)
{
    const double A = pow(10.0, gain_db / 40.0);
    const double w0 = 2.0 * M_PI * (cutoff / fs);
    const double alpha = sin(w0) / (2.0 * q);
    const double c = cos(w0);
    const double sA = sqrt(A);
    double b0, b1, b2, a0, a1, a2;
    if (type == 2) {
        b0 = A * ((A + 1) + (A - 1) * c + 2 * sA * alpha);
        b1 = -2 * A * ((A - 1) + (A + 1) * c);
        b2 = A * ((A + 1) + (A - 1) * c - 2 * sA * alpha);
        a0 = (A + 1) - (A - 1) * c + 2 * sA * alpha;
        a1 = 2 * ((A - 1) - (A + 1) * c);
        a2 = (A + 1) - (A - 1) * c - 2 * sA * alpha;
    } else if (type == 0) {
        b0 = A * ((A + 1) - (A - 1) * c + 2 * sA * alpha);
        b1 = 2 * A * ((A - 1) - (A + 1) * c);
        b2 = A * ((A + 1) - (A - 1) * c - 2 * sA * alpha);
        a0 = (A + 1) + (A - 1) * c + 2 * sA * alpha;
        a1 = -2 * ((A - 1) + (A + 1) * c);
        a2 = (A + 1) + (A - 1) * c - 2 * sA * alpha;
    } else {
        b0 = 1 + alpha * A;
        b1 = -2 * c;
        b2 = 1 - alpha * A;
        a0 = 1 + alpha / A;
        a1 = -2 * c;
        a2 = 1 - alpha / A;
    }
    b[0] = b0 / a0; b[1] = b1 / a0; b[2] = b2 / a0;
    a[0] = a0 / a0; a[1] = a1 / a0; a[2] = a2 / a0;
}

/* One channel through the 6-biquad cascade.  p[18] holds the de-normalised
 * (gain_db, cutoff, q) triples in the order low-shelf, band0..3, high-shelf
 * (effects.py:823-840).  Every stage is evaluated exactly like scipy's
 * lfilter inner loop:  y = z0 + b0*x;  z0 = (z1 + b1*x) - a1*y;  z1 = b2*x - a2*y
 * in fp64 with one final cast to fp32 (effects.py:512). */
ORACLE_API void oracle_eq(const float *x, int64_t n, const double *p, double fs, float *y)
{
    double b[6][3], a[6][3], z0[6] = {0}, z1[6] = {0};
    for (int s = 0; s < 6; ++s) {
        int type = (s == 0) ? 0 : (s == 5 ? 2 : 1);
        oracle_biquad_coefs(p[3 * s], p[3 * s + 1], p[3 * s + 2], fs, type, b[s], a[s]);
    }
    for (int64_t i = 0; i < n; ++i) {
        double v = (double)x[i];
        for (int s = 0; s < 6; ++s) {
            double out = z0[s] + b[s][0] * v;
            z0[s] = (z1[s] + b[s][1] * v) - a[s][1] * out;
            z1[s] = b[s][2] * v - a[s][2] * out;
            v = out;
        }
        y[i] = (float)v;
    }
}

/* ---------------------------------------------------------- compressor ---- */

/* juce::dsp::Compressor<float> (peak BallisticsFilter + VCA), zero initial
 * envelope, one independent envelope per channel [recollection]:
 *   expFactor = -2*pi*1000/fs                      (double)
 *   cte(t_ms) = t_ms < 1e-3 ? 0 : (float)exp(expFactor / t_ms)
 *   a = |x|;  c = a > env ? cteAT : cteRL;  env = a + c*(env - a)
 *   gain = env < thr ? 1 : powf(env * (1/thr), 1/ratio - 1);   y = gain*x
 *   thr = powf(10, thr_db * 0.05f)  (decibelsToGain with a -200 dB floor)
 */
ORACLE_API void oracle_compressor(const float *x, int64_t n, float thr_db, float ratio,
                                  float attack_ms, float release_ms, double fs, float *y)
{
    const double exp_factor = -2.0 * M_PI * 1000.0 / fs;
    const float cte_at = attack_ms < 1.0e-3f ? 0.0f : (float)exp(exp_factor / (double)attack_ms);
    const float cte_rl = release_ms < 1.0e-3f ? 0.0f : (float)exp(exp_factor / (double)release_ms);
    const float thr = thr_db > -200.0f ? powf(10.0f, thr_db * 0.05f) : 0.0f;
    const float thr_inv = 1.0f / thr;
    const float expo = 1.0f / ratio - 1.0f;
    float env = 0.0f;
    for (int64_t i = 0; i < n; ++i) {
        const float in = x[i];
        const float a = fabsf(in);
        const float c = (a > env) ? cte_at : cte_rl;
        env = a + c * (env - a);
        const float g = (env < thr) ? 1.0f : powf(env * thr_inv, expo);
        y[i] = g * in;
    }
}

/* -------------------------------------------------------------- reverb ---- */

typedef struct { float *buf; int size, idx; float last; } comb_t;
typedef struct { float *buf; int size, idx; } allpass_t;

static const short k_comb_tunings[8] = {1116, 1188, 1277, 1356, 1422, 1491, 1557, 1617};
static const short k_allpass_tunings[4] = {556, 441, 341, 225};
#define STEREO_SPREAD 23

/* x86 JUCE builds define JUCE_UNDENORMALISE(x) as { x += 0.1f; x -= 0.1f; } */
#define UNDENORM(v) do { volatile float t_ = (v) + 0.1f; (v) = t_ - 0.1f; } while (0)

static inline float comb_process(comb_t *c, float input, float damp, float fb)
{
    const float out = c->buf[c->idx];
    c->last = (out * (1.0f - damp)) + (c->last * damp);
    UNDENORM(c->last);
    float t = input + (c->last * fb);
    UNDENORM(t);
    c->buf[c->idx] = t;
    c->idx = (c->idx + 1) % c->size;
    return out;
}

static inline float allpass_process(allpass_t *p, float input)
{
    const float bv = p->buf[p->idx];
    float t = input + (bv * 0.5f);
    UNDENORM(t);
    p->buf[p->idx] = t;
    p->idx = (p->idx + 1) % p->size;
    return bv - input;
}

/* juce::Reverb (Freeverb).  chs == 2: processStereo on (l, r); chs == 1:
 * processMono on l (r ignored).  Parameters are constant over the block (the
 * SmoothedValues are reset to their targets by prepare()) [recollection]. */
ORACLE_API void oracle_reverb(float *l, float *r, int64_t n, int chs, float room, float damping,
                              float wet_level, float dry_level, float width, double fs)
{
    comb_t comb[2][8];
    allpass_t ap[2][4];
    const int isr = (int)fs;
    for (int c = 0; c < 2; ++c) {
        for (int j = 0; j < 8; ++j) {
            comb[c][j].size = (isr * (k_comb_tunings[j] + (c ? STEREO_SPREAD : 0))) / 44100;
            comb[c][j].buf = (float *)calloc((size_t)comb[c][j].size, sizeof(float));
            comb[c][j].idx = 0; comb[c][j].last = 0.0f;
        }
        for (int j = 0; j < 4; ++j) {
            ap[c][j].size = (isr * (k_allpass_tunings[j] + (c ? STEREO_SPREAD : 0))) / 44100;
            ap[c][j].buf = (float *)calloc((size_t)ap[c][j].size, sizeof(float));
            ap[c][j].idx = 0;
        }
    }
    const float wet = wet_level * 3.0f;
    const float dry = dry_level * 2.0f;
    const float wet1 = 0.5f * wet * (1.0f + width);
    const float wet2 = 0.5f * wet * (1.0f - width);
    const float gain = 0.015f;
    const float damp = damping * 0.4f;
    const float fb = room * 0.28f + 0.7f;

    if (chs == 2) {
        for (int64_t i = 0; i < n; ++i) {
            const float input = (l[i] + r[i]) * gain;
            float outl = 0.0f, outr = 0.0f;
            for (int j = 0; j < 8; ++j) {
                outl += comb_process(&comb[0][j], input, damp, fb);
                outr += comb_process(&comb[1][j], input, damp, fb);
            }
            for (int j = 0; j < 4; ++j) {
                outl = allpass_process(&ap[0][j], outl);
                outr = allpass_process(&ap[1][j], outr);
            }
            const float li = l[i], ri = r[i];
            l[i] = outl * wet1 + outr * wet2 + li * dry;
            r[i] = outr * wet1 + outl * wet2 + ri * dry;
        }
    } else {
        for (int64_t i = 0; i < n; ++i) {
            const float input = l[i] * gain;
            float out = 0.0f;
            for (int j = 0; j < 8; ++j) out += comb_process(&comb[0][j], input, damp, fb);
            for (int j = 0; j < 4; ++j) out = allpass_process(&ap[0][j], out);
            l[i] = out * wet1 + l[i] * dry;
        }
    }
    for (int c = 0; c < 2; ++c) {
        for (int j = 0; j < 8; ++j) free(comb[c][j].buf);
        for (int j = 0; j < 4; ++j) free(ap[c][j].buf);
    }
}

/* ---------------------------------------------------------- distortion ---- */

/* pedalboard.Distortion = juce::dsp::Gain(drive_db) -> tanh waveshaper, then
 * pedalboard.Gain(output_gain_db) (effects.py:900-914) [recollection]. */
ORACLE_API void oracle_distortion(const float *x, int64_t n, float drive_db, float out_gain_db,
                                  float *y)
{
    const float drive = drive_db > -100.0f ? powf(10.0f, drive_db * 0.05f) : 0.0f;
    const float og = out_gain_db > -100.0f ? powf(10.0f, out_gain_db * 0.05f) : 0.0f;
    for (int64_t i = 0; i < n; ++i) y[i] = tanhf(x[i] * drive) * og;
}

/* --------------------------------------------------------------- delay ---- */

/* pedalboard.Delay: juce::dsp::DelayLine<float, Linear> with feedback and a
 * dry/wet mix, independent per channel (effects.py:917-934) [recollection]:
 *   d    = (int)(delay_seconds * fs)          whole-sample delay
 *   out  = line[n - d]                         (0 before the line fills)
 *   line[n] = x[n] + feedback * out
 *   y[n] = (1 - mix) * x[n] + mix * out
 */
ORACLE_API void oracle_delay(const float *x, int64_t n, float delay_seconds, float feedback,
                             float mix, double fs, float *y)
{
    int d = (int)((double)delay_seconds * fs);
    if (d < 1) d = 1;
    float *line = (float *)calloc((size_t)d, sizeof(float));
    int idx = 0;
    const float dry = 1.0f - mix;
    for (int64_t i = 0; i < n; ++i) {
        const float out = line[idx];
        const float in = x[i];
        line[idx] = in + feedback * out;
        idx = (idx + 1) % d;
        y[i] = dry * in + mix * out;
    }
    free(line);
}

/* ---------------------------------------------------------------- misc ---- */

/* x /= clip(max|x|, 1e-8) over n samples (style_transfer.py:113); returns the peak. */
ORACLE_API float oracle_peak_normalize(float *x, int64_t n)
{
    float pk = 0.0f;
    for (int64_t i = 0; i < n; ++i) { const float a = fabsf(x[i]); if (a > pk) pk = a; }
    const float d = pk < 1e-8f ? 1e-8f : pk;
    for (int64_t i = 0; i < n; ++i) x[i] /= d;
    return pk;
}
