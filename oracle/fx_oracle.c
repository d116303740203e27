/*
 * fx_oracle.c -- plain-C restatement ("port") of the reference's per-frame analysis path.
 * TEST INFRASTRUCTURE ONLY: the checker for the CUDA path.  Nothing under feature-extractor_b200/
 * may include, link, load or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED against outputs of the reference itself -- oracle/_ref/libfxref.so is the
 * reference's own headers compiled headless (ref_driver.cpp); tests/test_oracle.py checks this port
 * against it bit for bit on seeded signals, and against tests/golden/ fixtures generated from it
 * (tests/golden/make_golden.py).  The reference ships no tests or golden vectors of its own
 * (SURVEY.md section 4), and its JUCE 4.2.3 dependency (juce::FFT, AudioSampleBuffer) is restated
 * from JUCE's published algorithm, so parity at the JUCE boundary is "unpinned" (see DESIGN.md).
 *
 * All file:line citations are relative to /root/reference/Source/.
 * Build: gcc -O2 -std=c99 (no -ffast-math, no -march=native: FMA contraction would change results).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdint.h>
#include "fx_oracle_api.h"

#define FX_PI_D 3.1415926535897932384626433832795
#define FX_PI_F 3.14159265358979323846f

/* ------------------------------------------------------------------------------------------------
 * juce::FFT (JUCE 4.2.3 juce_audio_basics, restated): radix-4 then radix-2 decimation in time,
 * fp32 arithmetic, twiddles evaluated in double and rounded to float, forward un-normalised.
 * Reached from RealTimeFFT (RealTimeAudioAnalysis.h:159-189). */
typedef struct { float r, i; } cpx;

typedef struct {
    int  n, inverse, n_stages;
    int  radix[32], length[32];
    cpx* tw;
} fft_plan;

static void fft_plan_init (fft_plan* p, int n, int inverse)
{
    int i, m = n;
    p->n = n; p->inverse = inverse; p->n_stages = 0;
    p->tw = (cpx*) malloc (sizeof (cpx) * (size_t) n);
    /* juce_FFT.cpp (4.2.x) [JUCE-recall]: phase = i * inverseFactor, inverseFactor = (isInverse ? 2.0 : -2.0) * double_Pi / fftSize */
    const double inverse_factor = (inverse ? 2.0 : -2.0) * FX_PI_D / n;
    for (i = 0; i < n; ++i) {
        const double phase = i * inverse_factor;
        p->tw[i].r = (float) cos (phase);
        p->tw[i].i = (float) sin (phase);
    }
    while (m > 1) {
        const int radix = (m % 4 == 0) ? 4 : 2;
        m /= radix;
        p->radix[p->n_stages] = radix;
        p->length[p->n_stages] = m;
        p->n_stages++;
    }
}

static void fft_plan_free (fft_plan* p) { free (p->tw); p->tw = NULL; }

static cpx cmul (cpx a, cpx b) { cpx c; c.r = a.r * b.r - a.i * b.i; c.i = a.r * b.i + a.i * b.r; return c; }
static cpx cadd (cpx a, cpx b) { cpx c; c.r = a.r + b.r; c.i = a.i + b.i; return c; }
static cpx csub (cpx a, cpx b) { cpx c; c.r = a.r - b.r; c.i = a.i - b.i; return c; }

static void fft_combine2 (const fft_plan* p, cpx* d, int stride, int length)
{
    int i;
    for (i = 0; i < length; ++i) {
        const cpx s = cmul (d[length + i], p->tw[i * stride]);
        d[length + i] = csub (d[i], s);
        d[i] = cadd (d[i], s);
    }
}

static void fft_combine4 (const fft_plan* p, cpx* d, int stride, int length)
{
    int i;
    for (i = 0; i < length; ++i) {
        const cpx s0 = cmul (d[i + length],     p->tw[i * stride]);
        const cpx s1 = cmul (d[i + 2 * length], p->tw[i * stride * 2]);
        const cpx s2 = cmul (d[i + 3 * length], p->tw[i * stride * 3]);
        const cpx s3 = cadd (s0, s2);
        const cpx s4 = csub (s0, s2);
        const cpx s5 = csub (d[i], s1);
        cpx a = cadd (d[i], s1);
        d[i + 2 * length] = csub (a, s3);
        d[i] = cadd (a, s3);
        if (p->inverse) {
            d[i + length].r     = s5.r - s4.i;  d[i + length].i     = s5.i + s4.r;
            d[i + 3 * length].r = s5.r + s4.i;  d[i + 3 * length].i = s5.i - s4.r;
        } else {
            d[i + length].r     = s5.r + s4.i;  d[i + length].i     = s5.i - s4.r;
            d[i + 3 * length].r = s5.r - s4.i;  d[i + 3 * length].i = s5.i + s4.r;
        }
    }
}

static void fft_recurse (const fft_plan* p, const cpx* in, cpx* out, int stride, int level)
{
    int q;
    const int radix = p->radix[level], length = p->length[level];
    if (length == 1) {
        for (q = 0; q < radix; ++q) out[q] = in[q * stride];
    } else {
        for (q = 0; q < radix; ++q)
            fft_recurse (p, in + q * stride, out + q * length, stride * radix, level + 1);
    }
    if (radix == 4) fft_combine4 (p, out, stride, length);
    else            fft_combine2 (p, out, stride, length);
}

static void fft_perform (const fft_plan* p, const cpx* in, cpx* out)
{
    if (p->n == 1) { out[0] = in[0]; return; }
    fft_recurse (p, in, out, 1, 0);
}

/* FFT::performRealOnlyForwardTransform: d holds N reals in, 2N interleaved floats out */
static void fft_real_forward (const fft_plan* p, float* d, cpx* scratch)
{
    int i;
    for (i = 0; i < p->n; ++i) { scratch[i].r = d[i]; scratch[i].i = 0.0f; }
    fft_perform (p, scratch, (cpx*) d);
}

/* FFT::performRealOnlyInverseTransform: 2N interleaved in; d[i] = Re/N, d[i+N] = Im/N out */
static void fft_real_inverse (const fft_plan* p, float* d, cpx* scratch)
{
    int i;
    const float scale = 1.0f / p->n;
    fft_perform (p, (const cpx*) d, scratch);
    for (i = 0; i < p->n; ++i) {
        d[i]        = scratch[i].r * scale;
        d[i + p->n] = scratch[i].i * scale;
    }
}

/* ------------------------------------------------------------------------------------------------
 * ValueHistory (RealTimeAudioAnalysis.h:40-96) */
#define FX_MAX_HIST 64
typedef struct { float h[FX_MAX_HIST]; int len, recorded; } history;

static void hist_set_length (history* v, int len)
{
    int i;
    if (len > FX_MAX_HIST) len = FX_MAX_HIST;
    v->len = len; v->recorded = 0;
    for (i = 0; i < FX_MAX_HIST; ++i) v->h[i] = 0.0f;
}

static void hist_push (history* v, float x)            /* :59-71 */
{
    int i;
    for (i = 0; i < v->len - 1; ++i) v->h[i] = v->h[i + 1];
    v->h[v->len - 1] = x;
    if (v->recorded < v->len) v->recorded++;
}

static float hist_total (const history* v)              /* :49-57 */
{
    int i; float total = 0.0f;
    for (i = 0; i < v->len; ++i) total += v->h[i];
    return total;
}

/* AudioFeatures::getValue (RealTimeAnalyser.h:84-88): NaN before the first push (0/0) */
static float hist_mean (const history* v) { return hist_total (v) / v->recorded; }

/* ------------------------------------------------------------------------------------------------ */
static float relmargin (double a, double b)
{
    double m = fabs (a) > fabs (b) ? fabs (a) : fabs (b);
    if (! (m > 0.0)) return 0.0f;
    return (float) (fabs (a - b) / m);
}
static void margin_min (float* dst, float m) { if (m < *dst) *dst = m; }

/* ------------------------------------------------------------------------------------------------
 * per-track state */
typedef struct {
    fxo_config cfg;
    int      n, m;                 /* window, numMagnitudes = N/2 */
    fft_plan fwd, inv;
    cpx*     scratch;
    float   *frame, *win, *filt, *buf_s, *buf_f, *buf_r, *ac;   /* N, N, N, 2N, 2N, 2N, 2N */
    float   *cnd;                  /* 2N */
    double  *prev_mag, *mag, *fc;  /* M each */
    float   *normed;               /* M */
    int     *peaks;                /* M */
    history  feat[FXO_NUM_FEATURES];
    history  onset_flux, onset_amp;
    float    diag[FXO_NUM_DIAG];
} track;

static void track_init (track* t, const fxo_config* cfg)
{
    int f;
    const int n = cfg->window, m = n / 2;
    memset (t, 0, sizeof (*t));
    t->cfg = *cfg; t->n = n; t->m = m;
    fft_plan_init (&t->fwd, n, 0);
    fft_plan_init (&t->inv, n, 1);
    t->scratch = (cpx*) calloc ((size_t) n, sizeof (cpx));
    t->frame = (float*) calloc ((size_t) n, sizeof (float));
    t->win   = (float*) calloc ((size_t) n, sizeof (float));
    t->filt  = (float*) calloc ((size_t) n, sizeof (float));
    t->buf_s = (float*) calloc ((size_t) 2 * n, sizeof (float));
    t->buf_f = (float*) calloc ((size_t) 2 * n, sizeof (float));
    t->buf_r = (float*) calloc ((size_t) 2 * n, sizeof (float));
    t->ac    = (float*) calloc ((size_t) 2 * n, sizeof (float));
    t->cnd   = (float*) calloc ((size_t) 2 * n, sizeof (float));
    t->prev_mag = (double*) calloc ((size_t) m, sizeof (double));   /* SpectralCharacteristics.h:34-38 */
    t->mag   = (double*) calloc ((size_t) m, sizeof (double));
    t->fc    = (double*) calloc ((size_t) m, sizeof (double));
    t->normed = (float*) calloc ((size_t) m, sizeof (float));
    t->peaks = (int*) calloc ((size_t) m, sizeof (int));
    /* AudioFeatures ctor (RealTimeAnalyser.h:70-74): depth 1 for onset and flux, 10 otherwise */
    for (f = 0; f < FXO_NUM_FEATURES; ++f)
        hist_set_length (&t->feat[f], (f == FXO_ONSET || f == FXO_FLUX) ? 1 : 10);
    hist_set_length (&t->onset_flux, cfg->onset_hist);              /* SpectralCharacteristics.h:237-241 */
    hist_set_length (&t->onset_amp,  cfg->onset_hist);
}

static void track_free (track* t)
{
    fft_plan_free (&t->fwd); fft_plan_free (&t->inv);
    free (t->scratch); free (t->frame); free (t->win); free (t->filt);
    free (t->buf_s); free (t->buf_f); free (t->buf_r); free (t->ac); free (t->cnd);
    free (t->prev_mag); free (t->mag); free (t->fc); free (t->normed); free (t->peaks);
}

/* AudioSampleBuffer::getRMSLevel [JUCE]: fp32 square, fp64 accumulate; then RealTimeAnalyser.h:148-149 / :207-208 */
static float log_rms (const float* x, int n)
{
    int i; double sum = 0.0; float rms;
    for (i = 0; i < n; ++i) { const float s = x[i]; sum += s * s; }
    rms = (float) sqrt (sum / n);
    return log10f (rms * 9.0f + 1.0f);
}

/* RealTimeWindower::scaleBufferWithBartlettWindowing (RealTimeAudioAnalysis.h:141-151) over
 * AudioSampleBuffer::applyGainRamp [JUCE]: gain accumulated additively in fp32 */
static void bartlett (float* x, int n)
{
    int i, half = n / 2;
    float g, inc;
    g = 0.0f; inc = (1.0f - 0.0f) / half;
    for (i = 0; i < half; ++i) { x[i] *= g; g += inc; }
    g = 1.0f; inc = (0.0f - 1.0f) / half;
    for (i = 0; i < half; ++i) { x[half + i] *= g; g += inc; }
}

/* AudioFilter::filterAudio (RealTimeAudioAnalysis.h:106-125), m = 2 (:127) */
static void one_pole (const float* in, float* out, int n)
{
    int i;
    const float m = 2.0f;
    if (n > 0) out[0] = in[0];
    for (i = 1; i < n; ++i)
        out[i] = ((FX_PI_F / m) * in[i]) + (expf (-FX_PI_F / m) * out[i - 1]);
}

/* FFTAnalyser::getFrequencyData (RealTimeAudioAnalysis.h:255-278) */
static void frequency_data (track* t, const float* time, float* buf2n)
{
    memcpy (buf2n, time, sizeof (float) * (size_t) t->n);
    memset (buf2n + t->n, 0, sizeof (float) * (size_t) t->n);
    fft_real_forward (&t->fwd, buf2n, t->scratch);
}

/* SpectralCharacteristicsAnalyser::calculateSpectralCharacteristics (+ fillIntermediateValues,
 * ...FromIntermediates) -- SpectralCharacteristics.h:62-143.  out5 = centroid, spread, flatness, ler, flux */
static void spectral_characteristics (track* t, const float* buf, double rms, double nyquist, float* out5)
{
    const int M = t->m;
    int i;
    const double frpb = nyquist / M;
    const int lower_portion = M / 5;
    const double eps = 0.01 * rms;                                              /* :108 */
    double weighted = 0.0, var = 0.0, mag_sum = 0.0, product = 1.0, flat_sum = 0.0, flux = 0.0, lhr = 0.0, count = 0.0;
    float max_flux, centroid, flatness, log_flatness, c, log_centroid, max_spread, spread;
    double inv;
    int state = 0;

    for (i = 0; i < M; ++i) {                                                   /* :66-96 */
        const double fcv = (double) i * frpb + (frpb / 2.0);
        const double v = (double) buf[2 * i];
        const double mag = v * v;
        const double diff = fabs (mag) - fabs (t->prev_mag[i]);
        const double rect = (diff + fabs (diff)) / 2.0;
        t->fc[i] = fcv;
        if (diff > 0.0) flux += rect;
        t->mag[i] = mag;
        mag_sum += mag;
        if (i == lower_portion) lhr = mag_sum;
        if (mag > eps) {
            flat_sum += mag;
            product *= mag;
            count += 1.0;
            if (state == 0 && product == 0.0) state = 1;
            if (state == 0 && isinf (product)) state = 2;
        }
        margin_min (&t->diag[FXO_DIAG_FLAT_MARGIN], relmargin (mag, eps));
        weighted += fcv * mag;
    }
    max_flux = (M * (M + 1)) / 2.0f;                                            /* :111 */
    flux /= max_flux;
    t->diag[FXO_DIAG_FLAT_COUNT] = (float) count;
    /* Margin of the gate decisions (diagnostic only).  The gate compares Re^2 with eps, and Re carries the absolute rounding
     * noise of an fp32 FFT -- this transform's as much as any other implementation's -- taken as 1e-6 of the spectrum's rms
     * (the same floor the pitch margin below discounts; an fp32 radix-4 FFT of this length is good to a few 1e-7 of the rms).
     * For every bin the relative gap shrinks by at most 2 e / sqrt (eps) + e^2 / eps: a bin whose gap is inside that band is
     * decided by the transform's rounding, in the reference itself too. */
    if (eps > 0.0) {
        const double e_abs = 1.0e-6 * sqrt (mag_sum / (double) M);
        const double fm = (double) t->diag[FXO_DIAG_FLAT_MARGIN] - (2.0 * e_abs / sqrt (eps) + e_abs * e_abs / eps);
        t->diag[FXO_DIAG_FLAT_MARGIN] = fm > 0.0 ? (float) fm : 0.0f;
    }
    margin_min (&t->diag[FXO_DIAG_GATE_MARGIN], relmargin (mag_sum, 0.05));

    if (! (mag_sum > 0.05)) {                                                   /* :121-123: prev NOT updated */
        out5[0] = out5[1] = out5[2] = out5[3] = out5[4] = 0.0f;
        t->diag[FXO_DIAG_FLAT_STATE] = 3.0f;
        return;
    }
    t->diag[FXO_DIAG_FLAT_STATE] = (float) state;
    lhr /= mag_sum;
    centroid = (float) (weighted / mag_sum);
    inv = 1.0 / (count > 0.0 ? count : 1.0);
    flatness = flat_sum > eps ? (float) (pow (product, inv) / (inv * flat_sum)) : 0.0f;     /* :57-60 */
    log_flatness = (float) log10 (flatness * 9.0 + 1.0);
    c = centroid / (float) (nyquist / 2.0);
    log_centroid = log10f (c * 9.0f + 1.0f);
    for (i = 0; i < M; ++i) {                                                   /* :135-139 */
        var += pow ((t->fc[i] / nyquist) - (centroid / nyquist), 2.0) * t->mag[i];
        t->prev_mag[i] = t->mag[i];
    }
    max_spread = (float) ((centroid / nyquist) * (1.0 - (centroid / nyquist)));
    spread = (float) ((var / mag_sum) / max_spread);
    out5[0] = log_centroid; out5[1] = spread; out5[2] = log_flatness; out5[3] = (float) lhr; out5[4] = (float) flux;
}

/* SpectralCharacteristicsAnalyser::calculateNormalisedSpectralSlope -- SpectralCharacteristics.h:145-200 */
static float spectral_slope (track* t, const float* buf)
{
    const int M = t->m;
    int i;
    double mean_bin = 0.5, mean_energy = 0.0, prod_sum = 0.0, bin_var = 0.0, energy_var = 0.0;
    double max_mag, bin_std, energy_std, r, grad, di;
    float mx = 0.0f;
    /* AudioSampleBuffer::getMagnitude (channel, 0, M): max |x| over the first M RAW interleaved floats (:153) */
    for (i = 0; i < M; ++i) { const float a = fabsf (buf[i]); if (a > mx) mx = a; }
    max_mag = mx;
    for (i = 0; i < M; ++i) {
        const double v = buf[2 * i];
        const double mag = v * v;
        t->mag[i] = mag;
        if (mag > max_mag) max_mag = mag;
    }
    margin_min (&t->diag[FXO_DIAG_GATE_MARGIN], relmargin (max_mag, 0.0001));
    if (! (max_mag > 0.0001)) return 0.0f;                                      /* :165-167 */
    for (i = 0; i < M; ++i) {
        const double e = t->mag[i] / max_mag;
        mean_energy += e;
        prod_sum += (double) i * e;
    }
    mean_energy /= (double) M;
    for (di = 0.0; di < M; di++) {                                              /* :182-188 */
        const double ni = di / (double) M;
        const double e = t->mag[(int) di] / max_mag;
        bin_var += (ni - mean_bin) * (ni - mean_bin);
        energy_var += (e - mean_energy) * (e - mean_energy);
    }
    bin_var /= (double) M;
    energy_var /= (double) M;
    bin_std = sqrt (bin_var);
    energy_std = sqrt (energy_var);
    r = (prod_sum - (M * mean_energy * mean_bin)) / (M - 1.0f) * energy_std * bin_std;      /* :195 */
    grad = r * (bin_std / energy_std);                                           /* :198 */
    return (float) grad;
}

/* OnsetDetector::detectOnset -- SpectralCharacteristics.h:249-306 */
static int detect_onset (track* t)
{
    const history* sf = &t->onset_flux;
    const history* am = &t->onset_amp;
    const int type = t->cfg.onset_type;
    const float mult = t->cfg.onset_multiplier;
    float mean_sf, mean_amp, cand_sf, cand_amp;
    int cand, i, onset_sf, onset_amp;

    if (am->recorded == 0 || sf->recorded == 0) return 0;
    if (sf->recorded < sf->len || am->recorded < am->len) return 0;
    mean_sf  = hist_total (sf) / sf->recorded;
    mean_amp = hist_total (am) / am->recorded;
    cand = sf->len - 1;
    if (type == 0 || type == 2) cand = sf->len / 2;
    cand_sf = sf->h[cand];
    cand_amp = am->h[cand];
    margin_min (&t->diag[FXO_DIAG_ONSET_MARGIN], relmargin (cand_amp, 0.01f));
    if (cand_amp < 0.01f) return 0;
    for (i = 0; i < sf->len; ++i) {
        if (i != cand) {
            if (type == 1 || type == 2) margin_min (&t->diag[FXO_DIAG_ONSET_MARGIN], relmargin (am->h[i], cand_amp));
            if (am->h[i] >= cand_amp && (type == 1 || type == 2)) return 0;
            if (type == 0 || type == 2) margin_min (&t->diag[FXO_DIAG_ONSET_MARGIN], relmargin (sf->h[i], cand_sf));
            if (sf->h[i] >= cand_sf && (type == 0 || type == 2)) return 0;
        }
    }
    onset_sf  = cand_sf  > mean_sf  * mult;
    onset_amp = cand_amp > mean_amp * mult;
    if (type != 0) margin_min (&t->diag[FXO_DIAG_ONSET_MARGIN], relmargin (cand_amp, mean_amp * mult));
    if (type != 1) margin_min (&t->diag[FXO_DIAG_ONSET_MARGIN], relmargin (cand_sf, mean_sf * mult));
    switch (type) {
        case 1:  return onset_amp;
        case 0:  return onset_sf;
        case 2:  return onset_amp && onset_sf;
        default: return 0;
    }
}

/* Margin of a comparison between two cnd values whose fp32 uncertainties are ua, ub: the part of the relative
 * gap that exceeds the uncertainty.  0 means "decided by rounding noise". */
static float noisy_margin (float a, float ua, float b, float ub)
{
    const float gap = fabsf (a - b) - (ua + ub);
    const float m = fabsf (a) > fabsf (b) ? fabsf (a) : fabsf (b);
    if (! (gap > 0.0f) || ! (m > 0.0f)) return 0.0f;
    return gap / m;
}

/* PitchAnalyser::estimatePitch -- PitchAnalyser.h:24-59, :83-217.  Returns f0 (Hz), sets lag + margin.
 * The margin (diagnostic only) discounts the fp32 noise floor of the autocorrelation: d[s] carries an absolute
 * error of about 1e-6 d[0] after two fp32 FFTs, so cnd[s] = d[s]^2 s / sum carries u[s] = (2 |d[s]| e + e^2) s / sum. */
static double estimate_pitch (track* t, const float* freq, double nyquist)
{
    const int N = t->n, two_n = 2 * t->n;
    int s, lag = -1;
    float sum = 0.0f, global_min = 100.0f, global_min_index = -1.0f, second_min = 100.0f, lag_estimate;
    float margin = 1.0f, e_abs;
    float* unc = (float*) t->scratch;          /* N floats of scratch: uncertainty of cnd[s] */
    int crossed = 0;

    /* getComplexConjugateMultiplication (:83-108): Re^2, imaginary cleared */
    for (s = 0; s < two_n; s += 2) {
        const float re = freq[s];
        t->ac[s] = re * re;
        t->ac[s + 1] = 0.0f;
    }
    /* getAutoCorrelationFromConjugateMultiplication (:110-127) */
    fft_real_inverse (&t->inv, t->ac, t->scratch);
    e_abs = 1.0e-6f * fabsf (t->ac[0]);
    for (s = 0; s < N; ++s) unc[s] = fabsf (t->ac[s]);       /* |d[s]| for now */
    for (s = 0; s < two_n; ++s) t->ac[s] = t->ac[s] * t->ac[s] * s;
    /* getCumulativeNormalisedDifferenceFromAutoCorrelationBuffer (:129-159): sequential fp32 sum */
    t->cnd[0] = 1.0f;
    for (s = 1; s < two_n; ++s) {
        const float value = t->ac[s];
        sum += value;
        t->cnd[s] = (sum != 0.0f) ? value / sum : 0.0f;
        if (s < N) unc[s] = (sum != 0.0f) ? (2.0f * unc[s] * e_abs + e_abs * e_abs) * (float) s / sum : 0.0f;
    }
    /* getLagEstimateFromCumulativeDifference (:161-190), threshold 0.01 */
    for (s = 2; s < N; ++s) {
        if (t->cnd[s] < global_min) { second_min = global_min; global_min_index = (float) s; global_min = t->cnd[s]; }
        else if (t->cnd[s] < second_min) second_min = t->cnd[s];
        margin_min (&margin, noisy_margin (t->cnd[s], unc[s], 0.01f, 0.0f));
        if (t->cnd[s] < 0.01f) {
            int right;
            while (s + 1 < N && t->cnd[s + 1] < t->cnd[s]) {
                margin_min (&margin, noisy_margin (t->cnd[s + 1], unc[s + 1], t->cnd[s], unc[s]));
                s++;
            }
            if (s + 1 < N) margin_min (&margin, noisy_margin (t->cnd[s + 1], unc[s + 1], t->cnd[s], unc[s]));
            /* getInterpolatedValleyFromCumulativeDifferenceLagEstimate (:192-217): leftNeighbour == lagEstimate
             * always, so the parabolic branch is unreachable and the result is an integer lag */
            right = s + ((s < N + 1) ? 1 : 0);
            lag = (t->cnd[s] <= t->cnd[right]) ? s : right;
            crossed = 1;
            break;
        }
    }
    if (! crossed) {
        margin_min (&margin, relmargin (global_min, second_min));
        lag_estimate = global_min_index;
    } else {
        lag_estimate = (float) lag;
    }
    t->diag[FXO_DIAG_LAG] = lag_estimate;
    t->diag[FXO_DIAG_PITCH_MARGIN] = margin;
    return (nyquist * 2.0f) / lag_estimate;                                      /* :57 */
}

static int bin_for_frequency (double freq, double frpb) { return (int) (floor (freq / frpb)); }     /* HarmonicCharacteristics.h:246-249 */

static double frequency_ratio (double f1, double f2)                             /* :251-259 */
{
    double higher, lower;
    if (f1 == f2) return 1.0;
    higher = f1 > f2 ? f1 : f2;
    lower = higher == f1 ? f2 : f1;
    return higher / lower;
}

static double max_bin_in_neighbourhood (const float* normed, int num_bins, int centre, int range)   /* :200-210 */
{
    int b;
    const int start = centre - range >= 0 ? centre - range : 0;
    const int end   = centre + range < num_bins ? centre + range : num_bins;
    double mx = normed[centre];
    for (b = start; b < end; ++b)
        if (normed[b] > mx) mx = normed[b];
    return mx;
}

/* HarmonicCharacteristicsAnalyser::calculateHarmonicCharacteristics -- HarmonicCharacteristics.h:46-106.
 * out3 = logHER, logOER, logInharm */
static void harmonic_characteristics (track* t, const float* buf, double f0, double nyquist, float* out3)
{
    const int M = t->m;
    int i, n_peaks = 0, f0_bin;
    double mean_mag, mag_sum = 0.0, max_mag = 0.0, sum_normed = 0.0, frpb;
    double score = 0.0, even = 0.0, odd = 0.0, her, oer, inharm = 0.0, lower, harmonic;
    float her_f, oer_f;

    for (i = 0; i < M; ++i) {                                                   /* :61-69 */
        const double v = (double) buf[2 * i];
        const double mag = v * v;
        t->mag[i] = mag;
        mag_sum += mag;
        if (mag > max_mag) max_mag = mag;
    }
    for (i = 0; i < M; ++i) {                                                   /* :71-78 */
        const double nm = t->mag[i] / max_mag;
        t->normed[i] = (float) nm;
        sum_normed += nm;
    }
    mean_mag = mag_sum / (double) M;
    margin_min (&t->diag[FXO_DIAG_GATE_MARGIN], relmargin (mag_sum, 0.005));
    if (mag_sum < 0.005) { out3[0] = out3[1] = out3[2] = 0.0f; t->diag[FXO_DIAG_NUM_PEAKS] = 0.0f; return; }   /* :88-89 */

    /* fillPeakBins / binIsPeak (:115-145): neighbours bin-2, bin-1, bin+1 (loop end exclusive) */
    for (i = 0; i < M; ++i) {
        const double mag = t->mag[i];
        int left_off, right_off, nb, is_peak = 1;
        margin_min (&t->diag[FXO_DIAG_PEAK_MARGIN], relmargin (mag, mean_mag));
        if (mag <= mean_mag) continue;
        left_off  = i < 2 ? 2 - i : 0;
        right_off = i >= M - 2 ? 2 - ((M - 1) - i) : 0;
        for (nb = i - (2 - left_off); nb < i + (2 - right_off); nb++) {
            if (nb != i) {
                margin_min (&t->diag[FXO_DIAG_PEAK_MARGIN], relmargin (t->mag[nb], mag));
                if (t->mag[nb] > mag) { is_peak = 0; break; }
            }
        }
        if (is_peak) t->peaks[n_peaks++] = i;
    }
    t->diag[FXO_DIAG_NUM_PEAKS] = (float) n_peaks;

    frpb = nyquist / (double) M;
    /* calculateHarmonicEnergyCharacteristics (:147-198) called with numLower = 15, numHarmonics = 3 (:94) */
    f0_bin = bin_for_frequency (f0, frpb);
    for (lower = 1.0; lower < 15.0 + 1.0; ++lower) {
        const double lf = f0 / pow (2.0, lower);
        const int lb = bin_for_frequency (lf, frpb);
        if (lb == f0_bin) continue;
        score += max_bin_in_neighbourhood (t->normed, M, lb, 2);
    }
    for (harmonic = 1.0; harmonic < 3.0 + 1.0; harmonic++) {
        const double hf = f0 * harmonic;
        const int hb = bin_for_frequency (hf, frpb);
        double bm;
        if (hb >= M) break;
        bm = max_bin_in_neighbourhood (t->normed, M, hb, 2);
        if ((int) harmonic % 2 == 0) even += bm; else odd += bm;
        score += bm;
    }
    her = score / sum_normed;
    if (her > 1.0) her = 1.0;
    if (her < 0.0) her = 0.0;
    oer = 1.0;
    if (odd > 0.0) oer = even / odd;
    if (oer > 1.0) oer = 1.0;
    if (oer < 0.0) oer = 0.0;
    her_f = (float) her; oer_f = (float) oer;                                    /* struct of floats (:197) */

    if (f0 > 0.0) {                                                              /* calculateInharmonicity (:212-244) */
        int p;
        for (p = 0; p < n_peaks; ++p) {
            const int bin = t->peaks[p];
            double start_f, end_f, r0, r1, ratio, prop;
            if (f0_bin == bin) continue;
            start_f = bin * frpb;
            if (start_f == 0.0) start_f = frpb * 0.5;
            end_f = (double) (bin + 1) * frpb;
            r0 = frequency_ratio (start_f, f0);
            r1 = frequency_ratio (end_f, f0);
            if (floor (r0) != floor (r1)) continue;
            ratio = r0 < r1 ? r0 : r1;
            prop = ratio - floor (ratio);
            inharm += prop * (t->mag[bin] / mag_sum);
        }
    }
    out3[0] = (float) log10 ((double) her_f * 9.0 + 1.0);                        /* :101-103 */
    out3[1] = (float) log10 ((double) oer_f * 9.0 + 1.0);
    out3[2] = (float) log10 (inharm * 9.0 + 1.0);
}

/* one hop: RealTimeSpectralAnalyser::run body (RealTimeAnalyser.h:205-229) then
 * RealTimeHarmonicAnalyser::run body (:145-172) */
static void analyse_frame (track* t)
{
    const int N = t->n;
    const double nyquist = t->cfg.sample_rate / 2.0;
    float log_r, sc[5], slope, hc[3], cur_flux, cur_amp;
    double f0;
    int d, onset;

    for (d = 0; d < FXO_NUM_DIAG; ++d) t->diag[d] = 1.0f;
    t->diag[FXO_DIAG_TRUE_OER] = 0.0f;

    /* ---- spectral body ---- */
    log_r = log_rms (t->frame, N);
    hist_push (&t->feat[FXO_RMS], log_r);
    memcpy (t->win, t->frame, sizeof (float) * (size_t) N);
    bartlett (t->win, N);
    frequency_data (t, t->win, t->buf_s);
    spectral_characteristics (t, t->buf_s, log_r, nyquist, sc);
    hist_push (&t->feat[FXO_CENTROID], sc[0]);
    hist_push (&t->feat[FXO_FLATNESS], sc[2]);
    hist_push (&t->feat[FXO_LER],      sc[3]);
    hist_push (&t->feat[FXO_SPREAD],   sc[1]);
    hist_push (&t->feat[FXO_FLUX],     sc[4]);
    slope = spectral_slope (t, t->buf_s);
    hist_push (&t->feat[FXO_SLOPE], slope);
    /* RealTimeSpectralAnalyser::detectOnset (:236-242): smoothed flux + smoothed RMS into the detector */
    cur_flux = hist_mean (&t->feat[FXO_FLUX]);
    cur_amp  = hist_mean (&t->feat[FXO_RMS]);
    hist_push (&t->onset_flux, cur_flux);
    hist_push (&t->onset_amp,  cur_amp);
    onset = detect_onset (t);
    hist_push (&t->feat[FXO_ONSET], onset ? 1.0f : 0.0f);

    /* ---- harmonic body ---- */
    log_r = log_rms (t->frame, N);
    if (t->cfg.rms_pushes >= 2)
        hist_push (&t->feat[FXO_RMS], log_r);
    one_pole (t->frame, t->filt, N);
    bartlett (t->filt, N);
    frequency_data (t, t->filt, t->buf_f);
    frequency_data (t, t->frame, t->buf_r);                                      /* raw, UN-windowed (:161) */
    f0 = estimate_pitch (t, t->buf_f, nyquist);
    hist_push (&t->feat[FXO_F0], (float) (f0 / 5000.0));
    harmonic_characteristics (t, t->buf_r, f0, nyquist, hc);
    hist_push (&t->feat[FXO_HER],    hc[0]);
    hist_push (&t->feat[FXO_OER],    hc[0]);                                     /* :171 stores HER in the OER slot */
    hist_push (&t->feat[FXO_INHARM], hc[2]);
    t->diag[FXO_DIAG_TRUE_OER] = hc[1];
}

static long analyse_one (const fxo_config* cfg, const float* audio, long n_samples,
                         float* raw, float* smooth, float* diag, long max_frames)
{
    track t;
    const int N = cfg->window, H = cfg->hop;
    long f, frames = n_samples / H;
    int i, k;
    if (frames > max_frames) frames = max_frames;
    if (cfg->mode == 0 && H != N / 2) return -1;
    track_init (&t, cfg);
    for (f = 0; f < frames; ++f) {
        /* RealTimeAudioDataOverlapper::getNextBuffer (RealTimeAudioAnalysis.h:205-219) generalised to hop H;
         * the collector's gain multiply is AudioDataCollector.h:88 */
        for (i = 0; i + H < N; ++i) t.frame[i] = t.frame[i + H];
        for (i = 0; i < H; ++i)     t.frame[N - H + i] = audio[f * H + i] * cfg->gain;
        analyse_frame (&t);
        for (k = 0; k < FXO_NUM_FEATURES; ++k) {
            if (raw)    raw[f * FXO_NUM_FEATURES + k]    = t.feat[k].h[t.feat[k].len - 1];
            if (smooth) smooth[f * FXO_NUM_FEATURES + k] = hist_mean (&t.feat[k]);
        }
        if (diag) memcpy (diag + f * FXO_NUM_DIAG, t.diag, sizeof (t.diag));
    }
    track_free (&t);
    return frames;
}

/* ------------------------------------------------------------------------------------------------ */
void fxo_default_config (fxo_config* cfg)
{
    cfg->window = 2048;            /* AnalyserTrackController.h:20-21 */
    cfg->hop = 1024;               /* RealTimeAudioAnalysis.h:207 */
    cfg->sample_rate = 48000.0;    /* RealTimeAnalyser.h:100 */
    cfg->gain = 1.0f;              /* AudioDataCollector.h:129 */
    cfg->onset_type = 1;           /* SpectralCharacteristics.h:240 */
    cfg->onset_hist = 5;           /* SpectralCharacteristics.h:238-239 */
    cfg->onset_multiplier = 1.7f;  /* SpectralCharacteristics.h:311 */
    cfg->rms_pushes = 2;           /* RealTimeAnalyser.h:150,209 */
    cfg->mode = 1;
}

const char* fxo_kind (void) { return "port"; }

long fxo_analyse_track (const fxo_config* cfg, const float* audio, long n_samples,
                        float* raw, float* smooth, float* diag, long max_frames)
{
    return analyse_one (cfg, audio, n_samples, raw, smooth, diag, max_frames);
}

typedef struct {
    const fxo_config* cfg; const float* audio; long t0, t1, stride, n_samples, frames;
    float *raw, *smooth, *diag;
} job;

static void* worker (void* arg)
{
    job* j = (job*) arg;
    long t;
    for (t = j->t0; t < j->t1; ++t)
        analyse_one (j->cfg, j->audio + t * j->stride, j->n_samples,
                     j->raw    ? j->raw    + t * j->frames * FXO_NUM_FEATURES : NULL,
                     j->smooth ? j->smooth + t * j->frames * FXO_NUM_FEATURES : NULL,
                     j->diag   ? j->diag   + t * j->frames * FXO_NUM_DIAG     : NULL, j->frames);
    return NULL;
}

long fxo_analyse_tracks (const fxo_config* cfg, const float* audio, long n_tracks, long track_stride, long n_samples,
                         float* raw, float* smooth, float* diag, long max_frames, int n_threads)
{
    long frames = n_samples / cfg->hop;
    int w;
    pthread_t* th;
    job* jobs;
    if (frames > max_frames) frames = max_frames;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_tracks) n_threads = (int) n_tracks;
    if (n_threads < 1) return frames;
    th = (pthread_t*) malloc (sizeof (pthread_t) * (size_t) n_threads);
    jobs = (job*) malloc (sizeof (job) * (size_t) n_threads);
    for (w = 0; w < n_threads; ++w) {
        jobs[w].cfg = cfg; jobs[w].audio = audio; jobs[w].stride = track_stride; jobs[w].n_samples = n_samples;
        jobs[w].frames = frames; jobs[w].raw = raw; jobs[w].smooth = smooth; jobs[w].diag = diag;
        jobs[w].t0 = n_tracks * w / n_threads; jobs[w].t1 = n_tracks * (w + 1) / n_threads;
        pthread_create (&th[w], NULL, worker, &jobs[w]);
    }
    for (w = 0; w < n_threads; ++w) pthread_join (th[w], NULL);
    free (th); free (jobs);
    return frames;
}

void fxo_fft_forward (const float* frame, int n, float* out_2n)
{
    fft_plan p; cpx* scratch = (cpx*) malloc (sizeof (cpx) * (size_t) n);
    fft_plan_init (&p, n, 0);
    memcpy (out_2n, frame, sizeof (float) * (size_t) n);
    memset (out_2n + n, 0, sizeof (float) * (size_t) n);
    fft_real_forward (&p, out_2n, scratch);
    fft_plan_free (&p); free (scratch);
}

void fxo_fft_inverse (float* inout_2n, int n)
{
    fft_plan p; cpx* scratch = (cpx*) malloc (sizeof (cpx) * (size_t) n);
    fft_plan_init (&p, n, 1);
    fft_real_inverse (&p, inout_2n, scratch);
    fft_plan_free (&p); free (scratch);
}

/* ---- file ingest: PCM decode ------------------------------------------------------------------------------
 * juce::AudioFormatReader::read [JUCE-recall, juce_audio_formats 4.2.3; not vendored under /root/reference]: the WAV /
 * AIFF readers left-justify integer samples into int32 (8-bit WAV is unsigned with a 128 offset) and the float conversion
 * multiplies (float) int32 by 1.0f / 0x7fffffff, which is exactly 2^-31 in fp32.  Parity unpinned: the reference holds
 * no test or fixture for this conversion; AudioFilePlayer (AudioFilePlayer.h:41-60) only hands the file to JUCE. */
long fxo_pcm_decode (const void* pcm, int format, int n_channels, int channel, long n_samples, float* out)
{
    static const int bytes[11] = { 0, 1, 1, 2, 2, 3, 3, 4, 4, 4, 4 };
    if (! pcm || ! out || format < 1 || format > 10 || n_channels < 1 || channel < 0 || channel >= n_channels || n_samples < 0) return -1;
    const int bps = bytes[format];
    const unsigned char* b = (const unsigned char*) pcm;
    const float scale = 1.0f / 0x7fffffff;
    for (long i = 0; i < n_samples; ++i)
    {
        const unsigned char* q = b + ((size_t) i * (size_t) n_channels + (size_t) channel) * (size_t) bps;
        uint32_t u = 0;
        switch (format)
        {
            case 1:  u = (uint32_t) (q[0] ^ 0x80u) << 24; break;                                        /* offset binary */
            case 2:  u = (uint32_t) q[0] << 24; break;
            case 3:  u = ((uint32_t) q[0] << 16) | ((uint32_t) q[1] << 24); break;
            case 4:  u = ((uint32_t) q[1] << 16) | ((uint32_t) q[0] << 24); break;
            case 5:  u = ((uint32_t) q[0] << 8) | ((uint32_t) q[1] << 16) | ((uint32_t) q[2] << 24); break;
            case 6:  u = ((uint32_t) q[2] << 8) | ((uint32_t) q[1] << 16) | ((uint32_t) q[0] << 24); break;
            case 7:  case 9:  u = (uint32_t) q[0] | ((uint32_t) q[1] << 8) | ((uint32_t) q[2] << 16) | ((uint32_t) q[3] << 24); break;
            default: u = (uint32_t) q[3] | ((uint32_t) q[2] << 8) | ((uint32_t) q[1] << 16) | ((uint32_t) q[0] << 24); break;
        }
        if (format >= 9) { float f; memcpy (&f, &u, 4); out[i] = f; }
        else             out[i] = (float) (int32_t) u * scale;
    }
    return n_samples;
}

/* ================================================================================================
 * Legacy offline analyser -- AudioAnalysis.h (struct AudioAnalyser), restated.  See fx_oracle_api.h for what is driven
 * and why (the feature block is commented out at the reference's own call site, :219-247; the member functions are
 * intact).  Pinned bit for bit against oracle/_ref (the reference's own functions) in tests/test_legacy.py.
 * Slot FXL_MARGIN (this port only; oracle/_ref reports -1) is the smallest relative margin of the decisions that feed the
 * harmonic features of the frame: peak tests (:366-381), the choice of the histogram's best candidate (:428-433), the
 * previousF0 hysteresis (:279-298), the silence gates (:271, :497, :575).
 * ================================================================================================ */
typedef struct { int interval, count; double her, freq; } l_cand;

static double l_ratio (double f1, double f2)                                       /* getFrequencyRatio :340-348 */
{
    double higher, lower;
    if (f1 == f2) return 1.0;
    higher = f1 > f2 ? f1 : f2;
    lower = higher == f1 ? f2 : f1;
    return higher / lower;
}

/* F0Candidate::updateHarmonicEnergyRatio :79-98 */
static double l_her (const float* mag, int nb, double frequency, double frpb, double total, double num_harmonics)
{
    double score = 0.0, harmonic;
    for (harmonic = 1.0; harmonic < num_harmonics + 1.0; harmonic++) {
        const double hf = frequency * harmonic;
        const int bin = (int) (ceil (hf / frpb));
        if (bin >= nb) break;
        score += (double) mag[bin];
    }
    return score / total;
}

static int l_is_peak (const float* mag, int nb, int bin, double mean, float* margin)     /* binIsPeak :366-387 */
{
    const double m = (double) mag[bin];
    int left, right, n;
    margin_min (margin, relmargin (m, mean));
    if (m <= mean) return 0;
    left = bin < 2 ? 2 - bin : 0;
    right = bin >= nb - 2 ? 2 - ((nb - 1) - bin) : 0;
    for (n = bin - (2 - left); n < bin + (2 - right); n++) {
        const double nm = (double) mag[n];
        if (n != bin) margin_min (margin, relmargin (nm, m));
        if (n != bin && nm > m) return 0;
    }
    return 1;
}

long fxo_legacy_analyse (int window, double sample_rate, const float* audio, long n_samples, int n_frames,
                         float* out, float* log_attack)
{
    const int N = window, nb = N / 2 + 1;
    const double nyquist = sample_rate / 2.0;
    const double frpb = nyquist / (double) nb;
    fft_plan plan;
    float *fft_in, *temp, *mag, *envelope;
    cpx* scratch;
    double* prev;
    int* peaks;
    l_cand* hist;
    double previous_f0 = 0.0;
    int frame, i, step;
    if (N < 16 || (N & (N - 1)) != 0 || n_frames < 1 || n_samples < n_frames) return -1;
    step = (int) (n_samples / n_frames);                                          /* :130 */
    fft_plan_init (&plan, N, 0);
    fft_in = (float*) calloc ((size_t) N, sizeof (float));
    temp = (float*) calloc ((size_t) 2 * N, sizeof (float));
    mag = (float*) calloc ((size_t) nb, sizeof (float));
    envelope = (float*) calloc ((size_t) n_frames, sizeof (float));
    scratch = (cpx*) calloc ((size_t) N, sizeof (cpx));
    prev = (double*) calloc ((size_t) nb, sizeof (double));                       /* previousBinMagnitudes, zeros (:115-116) */
    peaks = (int*) calloc ((size_t) nb, sizeof (int));
    hist = (l_cand*) calloc ((size_t) nb + 1, sizeof (l_cand));

    for (frame = 0; frame < n_frames; ++frame) {
        float* o = out + (long) frame * FXL_NUM;
        float margin = 1.0f;
        int range_start = -N / 2, range_end = N / 2, offset = N / 2, s;
        memset (fft_in, 0, (size_t) N * sizeof (float));                          /* :144 */
        if (n_frames == 1) { range_start = 0; range_end = N; offset = 0; }        /* :155-161 */
        for (s = range_start; s < range_end; ++s) {                               /* :165-179 */
            const long k = (long) frame * step + s;
            if (k < 0 || k >= n_samples) fft_in[s + N / 2] = 0.0f;
            else fft_in[s + offset] = audio[k];
        }
        /* scaleBufferWithBartlettWindowing :663-672: two applyGainRamp calls, gain accumulated additively in fp32 */
        {
            float g = 0.0f; const float inc = (1.0f - 0.0f) / (float) (N / 2);
            for (i = 0; i < N / 2; ++i) { fft_in[i] *= g; g += inc; }
            g = 1.0f;
            { const float dec = (0.0f - 1.0f) / (float) (N / 2); for (i = 0; i < N / 2; ++i) { fft_in[N / 2 + i] *= g; g += dec; } }
        }
        /* FFT::performFrequencyOnlyForwardTransform (:208) [JUCE-recall]: full complex transform, then juce_hypot per bin */
        for (i = 0; i < N; ++i) { temp[i] = fft_in[i]; temp[N + i] = 0.0f; }
        fft_real_forward (&plan, temp, scratch);
        for (i = 0; i < nb; ++i) mag[i] = (float) sqrt ((double) temp[2 * i] * temp[2 * i] + (double) temp[2 * i + 1] * temp[2 * i + 1]);

        /* ---- calculateSpectralCharacteristics :463-515 ------------------------------------------------ */
        {
            double weighted = 0.0, var = 0.0, sum = 0.0, product = 1.0, flux = 0.0;
            for (i = 0; i < nb; ++i) {
                const double fc = (double) i * frpb + (frpb / 2.0);
                const double m = (double) mag[i];
                const double diff = fabs (m) - fabs (prev[i]);
                const double rect = (diff + fabs (diff)) / 2.0;
                if (diff > 0.0) flux += rect;
                sum += m;
                product *= m;
                weighted += fc * m;
            }
            margin_min (&margin, relmargin (sum, 0.001));
            if (! (sum > 0.001)) { o[FXL_CENTROID] = o[FXL_SPREAD] = o[FXL_FLATNESS] = o[FXL_FLUX] = 0.0f; }
            else {
                const float centroid = (float) (weighted / sum);
                const double inv = 1.0 / (double) nb;
                const float flatness = (float) (pow (product, inv) / (inv * sum));
                float max_spread;
                for (i = 0; i < nb; ++i) {
                    const double fc = (double) i * frpb + (frpb / 2.0);
                    var += pow ((fc / nyquist) - (centroid / nyquist), 2.0) * (double) mag[i];
                    prev[i] = (double) mag[i];
                }
                max_spread = (float) ((centroid / nyquist) * (1.0 - (centroid / nyquist)));
                o[FXL_CENTROID] = centroid / (float) nyquist;
                o[FXL_SPREAD] = (float) ((var / sum) / max_spread);
                o[FXL_FLATNESS] = flatness;
                o[FXL_FLUX] = (float) flux;
            }
        }
        /* ---- calculateNormalisedSpectralSlope :566-609 ------------------------------------------------- */
        {
            const double num_bins = (double) nb, mean_bin = 0.5;
            double mean_energy = 0.0, prod_sum = 0.0, bin_var = 0.0, energy_var = 0.0, di;
            float mx = 0.0f;
            for (i = 0; i < nb; ++i) { const float a = fabsf (mag[i]); if (a > mx) mx = a; }       /* getMagnitude */
            margin_min (&margin, relmargin ((double) mx, 0.0001));
            if (! ((double) mx > 0.0001)) o[FXL_SLOPE] = 0.0f;
            else {
                const double fm = (double) mx;
                double bin_std, energy_std, r;
                for (i = 0; i < nb; ++i) { const double e = mag[i] / fm; mean_energy += e; prod_sum += (double) i * e; }
                mean_energy /= num_bins;
                for (di = 0.0; di < num_bins; di++) {
                    const double ni = di / num_bins;
                    const double e = mag[(int) di] / fm;
                    bin_var += (ni - mean_bin) * (ni - mean_bin);
                    energy_var += (e - mean_energy) * (e - mean_energy);
                }
                bin_var /= num_bins; energy_var /= num_bins;
                bin_std = sqrt (bin_var); energy_std = sqrt (energy_var);
                r = (prod_sum - (num_bins * mean_energy * mean_bin)) / (num_bins - 1.0f) * energy_std * bin_std;
                o[FXL_SLOPE] = (float) (r * (bin_std / energy_std));
            }
        }
        /* ---- calculateHarmonicCharacteristics :253-306 ------------------------------------------------- */
        {
            double sum = 0.0, mean;
            int n_peaks = 0, n_hist = 0;
            o[FXL_F0] = o[FXL_HER] = o[FXL_INHARM] = 0.0f; o[FXL_NUM_PEAKS] = 0.0f;
            for (i = 0; i < nb; ++i) sum += (double) mag[i];
            mean = sum / (double) nb;
            margin_min (&margin, relmargin (sum, 0.001));
            if (! (sum < 0.001)) {
                double f0 = 0.0, her = 0.0, max_weighted = 0.0, second_weighted = 0.0, inharm = 0.0;
                int b, p, c;
                /* fillPeakBinsAndFrequencyHistogram :350-363, addNewPeakBinAndUpdateHistogram :389-403 */
                for (b = 0; b < nb; ++b) {
                    if (! l_is_peak (mag, nb, b, mean, &margin)) continue;
                    peaks[n_peaks++] = b;
                    for (p = 0; p < n_peaks - 1; ++p) {
                        const int interval = b - peaks[p];
                        int pos = -1;
                        for (c = 0; c < n_hist; ++c) if (hist[c].interval == interval) { pos = c; break; }
                        if (pos != -1) hist[pos].count++;
                        else { hist[n_hist].interval = interval; hist[n_hist].count = 1; n_hist++; }
                    }
                }
                /* estimateF0AndHERFromFrequencyHistogram :415-436 */
                for (c = 0; c < n_hist; ++c) {
                    double w;
                    hist[c].freq = (double) hist[c].interval * frpb;
                    hist[c].her = l_her (mag, nb, hist[c].freq, frpb, sum, 15.0);
                    w = (double) hist[c].count * hist[c].her;
                    if (w > max_weighted) { second_weighted = max_weighted; max_weighted = w; f0 = hist[c].freq; her = hist[c].her; }
                    else if (w > second_weighted) second_weighted = w;
                }
                if (n_hist > 0) margin_min (&margin, relmargin (max_weighted, second_weighted));
                /* previousF0 hysteresis :279-298 */
                if (previous_f0 != f0 && previous_f0 > 10.0) {
                    const double top = previous_f0 > f0 ? previous_f0 : f0;
                    const double bottom = top == previous_f0 ? f0 : previous_f0;
                    const double ratio = top / bottom;
                    margin_min (&margin, relmargin (ratio, 2.0));
                    if (ratio > 2.0) {
                        const double frac = ratio - floor (ratio);
                        margin_min (&margin, relmargin (frac, 0.1));
                        if (frac < 0.1) { f0 = previous_f0; her = l_her (mag, nb, f0, frpb, sum, 15.0); }
                    }
                }
                previous_f0 = f0;
                /* calculateInharmonicity :308-338 */
                if (f0 > 0.0) {
                    const int f0_bin = (int) (ceil (f0 / frpb));                    /* getBinForFrequency :438-443 */
                    for (p = 0; p < n_peaks; ++p) {
                        const int bin = peaks[p];
                        double start, end, ra, rb, r;
                        if (f0_bin == bin) continue;
                        start = bin * frpb;
                        if (start == 0.0) start = frpb * 0.5;
                        end = (double) (bin + 1) * frpb;
                        ra = l_ratio (start, f0); rb = l_ratio (end, f0);
                        if (floor (ra) != floor (rb)) continue;
                        r = ra < rb ? ra : rb;
                        inharm += (r - floor (r)) * ((double) mag[bin] / sum);
                    }
                }
                o[FXL_F0] = (float) f0; o[FXL_HER] = (float) her; o[FXL_INHARM] = (float) inharm; o[FXL_NUM_PEAKS] = (float) n_peaks;
            }
        }
        /* energy envelope :249 (sumAccrossChannels: fp32 running sum in bin order) */
        { float e = 0.0f; for (i = 0; i < nb; ++i) e += mag[i]; envelope[frame] = e; o[FXL_ENERGY] = e; }
        o[FXL_MARGIN] = margin;
    }
    /* analyseNormalisedZeroCrosses :517-541 */
    for (frame = 0; frame < n_frames; ++frame) {
        float crosses = 0.0f;
        const long base = (long) frame * step;
        int s;
        for (s = 0; s < step - 1; ++s) {
            const float first = audio[base + s], second = audio[base + s + 1];
            if ((first > 0.0f && (first - second) > first) || (first < 0.0f && (first - second) < first)) crosses++;
        }
        out[(long) frame * FXL_NUM + FXL_ZCR] = crosses * 2.0f / (float) step;
    }
    /* setLogAttackTime :611-622 */
    if (log_attack != NULL) {
        float mx = envelope[0];
        const double ms_per_sample = 1.0 / (double) ((int) sample_rate / 1000);
        for (i = 1; i < n_frames; ++i) if (envelope[i] > mx) mx = envelope[i];
        i = 0;
        while (i < n_frames && envelope[i] != mx) i++;
        *log_attack = (float) (log10 ((double) (float) (i * step) * (float) ms_per_sample));
    }
    fft_plan_free (&plan);
    free (fft_in); free (temp); free (mag); free (envelope); free (scratch); free (prev); free (peaks); free (hist);
    return n_frames;
}
