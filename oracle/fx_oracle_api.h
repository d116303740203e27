/*
 * fx_oracle_api.h -- C API shared by the two CPU checkers under oracle/.
 * TEST INFRASTRUCTURE ONLY: nothing under feature-extractor_b200/ may include, link or load this.
 *
 *   oracle/_ref/libfxref.so   the reference's own headers (the .h files under /root/reference/Source), compiled
 *                             headless against oracle/juce_shim/JuceHeader.h (ref_driver.cpp)
 *   oracle/libfxoracle.so     plain-C restatement of the same path (fx_oracle.c)
 *
 * Both export the same entry points so the tests can swap one for the other.
 */
#ifndef FX_ORACLE_API_H
#define FX_ORACLE_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* feature slots follow AudioFeatures::eAudioFeature (RealTimeAnalyser.h:17-32) */
enum {
    FXO_ONSET = 0, FXO_RMS, FXO_F0, FXO_CENTROID, FXO_SPREAD, FXO_FLATNESS, FXO_LER, FXO_FLUX,
    FXO_SLOPE, FXO_HER, FXO_OER, FXO_INHARM, FXO_NUM_FEATURES
};

/* per-frame diagnostics (float[FXO_NUM_DIAG]) */
enum {
    FXO_DIAG_TRUE_OER = 0,     /* log-mapped odd/even ratio the reference computes but never stores (RealTimeAnalyser.h:171) */
    FXO_DIAG_LAG,              /* integer pitch lag (PitchAnalyser.h:161-190) */
    FXO_DIAG_PITCH_MARGIN,     /* smallest relative margin of any comparison that decided the lag  (port only, else -1) */
    FXO_DIAG_NUM_PEAKS,        /* number of peak bins (HarmonicCharacteristics.h:115-145)          (port only, else -1) */
    FXO_DIAG_PEAK_MARGIN,      /* smallest relative margin of any peak / bin-index decision        (port only, else -1) */
    FXO_DIAG_FLAT_COUNT,       /* bins gated into the flatness product (SpectralCharacteristics.h:89-94) (port only, else -1) */
    FXO_DIAG_FLAT_MARGIN,      /* smallest relative distance of a bin magnitude to the gate eps    (port only, else -1) */
    FXO_DIAG_GATE_MARGIN,      /* smallest relative margin of the silence gates (0.05, 0.005, 1e-4) (port only, else -1) */
    FXO_DIAG_ONSET_MARGIN,     /* smallest relative margin of the onset detector's comparisons     (port only, else -1) */
    FXO_DIAG_FLAT_STATE,       /* flatness product: 0 finite, 1 underflowed to 0, 2 overflowed to inf, 3 frame gated silent (port only, else -1) */
    FXO_NUM_DIAG
};

typedef struct fxo_config {
    int    window;            /* N, power of two                                  */
    int    hop;               /* H; mode 0 requires H == N/2                      */
    double sample_rate;
    float  gain;              /* AudioDataCollector::setGain                       */
    int    onset_type;        /* OnsetDetector::eOnsetDetectionType: 0 spectral, 1 amplitude (default), 2 combination */
    int    onset_hist;        /* 5                                                */
    float  onset_multiplier;  /* 1.7                                              */
    int    rms_pushes;        /* 2 = both analyser bodies push RMS (reference wiring), 1 = spectral only */
    int    mode;              /* 0 = A: verbatim collector/overlapper/run() bodies; 1 = B: re-sequenced per-frame calls, any hop */
} fxo_config;

void fxo_default_config (fxo_config* cfg);

/* which checker is this: "reference" or "port" */
const char* fxo_kind (void);

/* Analyse one track.  Outputs are [frames][12] raw (values as pushed into AudioFeatures), [frames][12]
 * smoothed (AudioFeatures::getValue after both analyser bodies of the hop), [frames][FXO_NUM_DIAG].
 * Any output pointer may be NULL.  Returns the number of frames written (min (n_samples / hop, max_frames)). */
long fxo_analyse_track (const fxo_config* cfg, const float* audio, long n_samples,
                        float* raw, float* smooth, float* diag, long max_frames);

/* Same over n_tracks rows of `audio` (row stride in samples), contiguous track ranges per thread. */
long fxo_analyse_tracks (const fxo_config* cfg, const float* audio, long n_tracks, long track_stride, long n_samples,
                         float* raw, float* smooth, float* diag, long max_frames, int n_threads);

/* The FFT the reference reaches through RealTimeFFT (RealTimeAudioAnalysis.h:167-177):
 * forward: N reals -> 2N interleaved floats; inverse: 2N interleaved -> d[0..N) = Re/N, d[N..2N) = Im/N. */
void fxo_fft_forward (const float* frame, int n, float* out_2n);
void fxo_fft_inverse (float* inout_2n, int n);

/* File ingest (port only): decode n_samples frames of interleaved PCM, taking one channel, to the fp32 samples the
 * analysis consumes.  Restates what juce::AudioFormatReader::read does for the formats registerBasicFormats() offers to
 * AudioFilePlayer (AudioFilePlayer.h:17,47) [JUCE-recall; JUCE is not under /root/reference]: integer samples are
 * left-justified into int32 and scaled by 1.0f / 0x7fffffff (= 2^-31 in fp32); 8-bit WAV is offset binary; floats pass through.
 * Formats: 1 U8, 2 S8, 3 S16LE, 4 S16BE, 5 S24LE, 6 S24BE, 7 S32LE, 8 S32BE, 9 F32LE, 10 F32BE.  Returns n_samples or -1. */
long fxo_pcm_decode (const void* pcm, int format, int n_channels, int channel, long n_samples, float* out);

/* ---- legacy offline analyser (AudioAnalysis.h: AudioAnalyser) ------------------------------------------------------
 * The application never instantiates AudioAnalyser, and the block of performSpectralAnalysis that would produce features is
 * commented out at its call site (AudioAnalysis.h:219-247); the member functions themselves are intact.  This entry point
 * drives them the way that block does, frame by frame: the framing and windowing of performSpectralAnalysis (:134-181: frame i is
 * centred on sample i * stepSize, zero padded at the ends, symmetric Bartlett ramps), performFrequencyOnlyForwardTransform
 * (:208, true magnitudes), fftOut <- the first N/2 + 1 magnitudes (:226-227), calculateSpectralCharacteristics (:463-515),
 * calculateNormalisedSpectralSlope (:566-609), calculateHarmonicCharacteristics (:253-306: histogram of peak intervals, previousF0
 * hysteresis, inharmonicity), energyEnvelope (:249), then analyseNormalisedZeroCrosses (:517-541) and setLogAttackTime (:611-622).
 * One channel.  out is [n_frames][FXL_NUM]; *log_attack receives estimatedLogAttackTime.  Returns n_frames or -1. */
enum {
    FXL_CENTROID = 0,   /* centroid / nyquist                      (:514) */
    FXL_SPREAD, FXL_FLATNESS, FXL_FLUX,
    FXL_SLOPE,          /* calculateNormalisedSpectralSlope         */
    FXL_F0, FXL_HER, FXL_INHARM,
    FXL_ZCR,            /* zero crossings * 2 / stepSize            (:538) */
    FXL_ENERGY,         /* energy envelope: sum of the frame's magnitudes (:249) */
    FXL_NUM_PEAKS,      /* diagnostics: number of peak bins (:361-369) */
    FXL_MARGIN,         /* diagnostics: smallest relative margin of the frame's decisions (port only, else -1) */
    FXL_NUM
};
long fxo_legacy_analyse (int window, double sample_rate, const float* audio, long n_samples, int n_frames,
                         float* out, float* log_attack);

#ifdef __cplusplus
}
#endif
#endif
