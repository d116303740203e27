/*
 * fx_engine.h -- C ABI of the B200-native per-frame analysis engine (libfxb200.so).
 *
 * This is the drop-in boundary for the reference's hot path L1a..L2:
 * everything between AudioDataCollector::audioDeviceIOCallback (Source/AudioDataCollector.h:36)
 * and AudioFeatures::getValue (Source/RealTimeAnalyser.h:84).  The reference has no FFI of its own --
 * its boundary is the C++ surface of those classes -- so each entry point below cites the reference
 * interface it replaces; feature-extractor_b200/host/ rebuilds that C++ surface (same class names and
 * method signatures) on top of these calls, and INTEGRATION.md shows the binding a maintainer adds.
 *
 * Plain pointers and sizes only; no exceptions cross this boundary; every call returns fx_status
 * (0 = OK, negative = error) and fx_last_error() gives the text.  There is NO CPU fallback: if the
 * CUDA device or kernels are unavailable, fx_engine_create fails.
 *
 * All file:line citations are relative to /root/reference/Source/.
 */
#ifndef FX_ENGINE_H
#define FX_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fx_engine fx_engine;
typedef int fx_status;

enum {
    FX_OK                =  0,
    FX_ERR_INVALID_ARG   = -1,
    FX_ERR_CUDA          = -2,
    FX_ERR_UNSUPPORTED   = -3,   /* window not in {1024, 2048, 4096}, hop not a multiple of 16 dividing the window, ... */
    FX_ERR_OVERRUN       = -4,   /* fx_push_block would overwrite samples that were not analysed yet */
    FX_ERR_NO_DEVICE     = -5
};

/* Feature slots: AudioFeatures::eAudioFeature (RealTimeAnalyser.h:17-32). */
enum {
    FX_ONSET = 0, FX_RMS, FX_F0, FX_CENTROID, FX_SPREAD, FX_FLATNESS, FX_LER, FX_FLUX,
    FX_SLOPE, FX_HER, FX_OER, FX_INHARM, FX_NUM_FEATURES
};

/* Per-frame diagnostics: integer decisions and the smallest relative margin that decided them
 * (BASELINE.json north_star: "frames whose decision margin falls below that tolerance ... are counted and reported"). */
enum {
    FX_DIAG_TRUE_OER = 0,    /* log-mapped odd/even ratio (HarmonicCharacteristics.h:190-195); the reference stores HER in its slot (RealTimeAnalyser.h:171) */
    FX_DIAG_LAG,             /* integer pitch lag (PitchAnalyser.h:161-190) */
    FX_DIAG_PITCH_MARGIN,
    FX_DIAG_NUM_PEAKS,       /* HarmonicCharacteristics.h:115-145 */
    FX_DIAG_PEAK_MARGIN,
    FX_DIAG_FLAT_COUNT,      /* bins gated into the flatness product (SpectralCharacteristics.h:89-94) */
    FX_DIAG_FLAT_MARGIN,
    FX_DIAG_GATE_MARGIN,     /* silence gates 0.05 / 0.005 / 1e-4 (SpectralCharacteristics.h:121-123,165-167; HarmonicCharacteristics.h:88) */
    FX_DIAG_ONSET_MARGIN,    /* SpectralCharacteristics.h:271-292 */
    FX_DIAG_FLAT_STATE,      /* 0 finite, 1 product underflowed to 0, 2 overflowed to inf, 3 frame gated silent */
    FX_NUM_DIAG
};

/* The order OSCFeatureAnalysisOutput::sendSpectralFeaturesViaOSC puts on the wire (OSCFeatureAnalysisOutput.h:107):
 * onset, rms, f0, centroid, slope, spread, flatness, ler, flux, her, oer, inharm (12 floats);
 * README.md:55-57 documents the same list without ler and oer (10 floats). */
#define FX_OSC_FLOATS_CODE   12
#define FX_OSC_FLOATS_README 10

typedef struct fx_config {
    int    n_tracks;            /* one AnalyserTrackController per track (AnalyserTrackController.h:17) */
    int    window;              /* N: 1024, 2048 (reference: AnalyserTrackController.h:20-21) or 4096 */
    int    hop;                 /* H: reference N/2 (RealTimeAudioAnalysis.h:207); any multiple of 16 dividing N */
    double sample_rate;         /* RealTimeAnalyser.h:100, sampleRateChanged :111-114 */
    int    device;              /* CUDA device ordinal */
    int    rms_pushes_per_frame;/* 2 = both analyser bodies push RMS (RealTimeAnalyser.h:150,209), 1 = spectral only */
    int    onset_type;          /* OnsetDetector::eOnsetDetectionType (SpectralCharacteristics.h:213-219), default 1 = amplitude */
    int    onset_hist;          /* SpectralCharacteristics.h:237-239, default 5, <= 16 */
    float  onset_multiplier;    /* meanThresholdMultiplier, SpectralCharacteristics.h:311, default 1.7 */
    float  gain;                /* AudioDataCollector::setGain (AudioDataCollector.h:124), default 1 */
    long   max_frames_per_call; /* capacity of the engine-owned device result buffers (frames per track per call) */
    int    ring_hops;           /* streaming: device + pinned ring length in hops per track (>= window/hop + 2) */
    int    tracks_per_group;    /* streaming: tracks per CUDA stream / pinned-ring group; 0 = all tracks in one group */
} fx_config;

void        fx_default_config (fx_config* cfg);

/* Replaces constructing AnalyserTrackController x n_tracks (AnalyserTrackController.h:17-45) + prepareToPlay (:175-188). */
fx_status   fx_engine_create  (const fx_config* cfg, fx_engine** out);
fx_status   fx_engine_destroy (fx_engine* e);
const char* fx_last_error     (const fx_engine* e);      /* e may be NULL: last create error */
const char* fx_version        (void);

/* Runtime parameter surface (applies from the next analysed frame).
 * AudioDataCollector::setGain (AudioDataCollector.h:124);
 * RealTimeSpectralAnalyser::setOnsetDetectionType / setOnsetWindowLength / setOnsetDetectionSensitivity
 * (RealTimeAnalyser.h:244-258; multiplier = 1 + sensitivity).  track = -1 applies to every track.
 * As in the reference, changing the window length clears the onset histories. */
fx_status   fx_set_gain  (fx_engine* e, int track, float gain);
fx_status   fx_set_onset (fx_engine* e, int track, int type, int hist_len, float multiplier);

/* Forget all per-track state (overlap buffer, previous spectrum, feature and onset histories): a fresh
 * RealTimeAudioDataOverlapper (RealTimeAudioAnalysis.h:197-203), SpectralCharacteristicsAnalyser
 * (SpectralCharacteristics.h:34-38) and AudioFeatures (RealTimeAnalyser.h:70-74). */
fx_status   fx_reset (fx_engine* e);

/* ---- offline / batch analysis: tracks x frames in one call ---------------------------------------
 * Analyses the NEXT n_samples of every track (state carries over from earlier calls; call fx_reset
 * for a fresh stream).  audio is [n_tracks][track_stride] fp32; frames = n_samples / hop per track.
 * Outputs are [n_tracks][frames][12] raw (values as pushed into AudioFeatures::updateFeature,
 * RealTimeAnalyser.h:76), [n_tracks][frames][12] smoothed (AudioFeatures::getValue, :84-88, after both
 * analyser bodies of the hop) and [n_tracks][frames][FX_NUM_DIAG]; any may be NULL.
 *
 * fx_analyse_host: HOST pointers (pinned or pageable); host<->device copies are inside the call,
 *                  pipelined over track groups; returns when the results are in host memory.
 * fx_analyse_device: DEVICE pointers, asynchronous on `stream` (a cudaStream_t; NULL = the engine's own). */
fx_status   fx_analyse_host   (fx_engine* e, const float* audio, long track_stride, long n_samples,
                               float* raw, float* smooth, float* diag, long* n_frames);
fx_status   fx_analyse_device (fx_engine* e, const float* d_audio, long track_stride, long n_samples,
                               float* d_raw, float* d_smooth, float* d_diag, void* stream, long* n_frames);

/* ---- file ingest: PCM in, features out ---------------------------------------------------------------
 * Replaces the file route of the reference: AudioFilePlayer::loadFileIntoTransport (AudioFilePlayer.h:41-60) hands the
 * file to JUCE's format readers, the transport plays it through the device output and AudioDataCollector re-captures one
 * output channel (AudioDataCollector.h:42-43).  Here the bytes of the file's data chunk go to the GPU as they are
 * (interleaved sample frames, WAV little endian / AIFF big endian), kernel k_pcm_decode converts the one channel each
 * track analyses to fp32 exactly as JUCE's readers do ((float) left-justified int32 * (1.0f / 0x7fffffff); 8-bit WAV is
 * offset binary; float data passes through), and the analysis runs on the result.  Integer samples of <= 24 bits
 * convert exactly, so a 16-bit file crosses PCIe at 2 bytes per sample and still yields the fp32 input the
 * float API would have been given. */
enum {
    FX_PCM_U8 = 1, FX_PCM_S8, FX_PCM_S16LE, FX_PCM_S16BE, FX_PCM_S24LE, FX_PCM_S24BE,
    FX_PCM_S32LE, FX_PCM_S32BE, FX_PCM_F32LE, FX_PCM_F32BE
};
int         fx_pcm_bytes_per_sample (int format);        /* 0 for an unknown format */
/* pcm holds n_tracks rows, track_stride_bytes apart; each row is n_samples sample frames of n_channels interleaved
 * samples, of which `channel` is analysed.  channel = -1: track t analyses channel t % n_channels (the reference's
 * tracks each pick one channel of the same device stream); track_stride_bytes = 0: every track reads the same row (one
 * multichannel file feeding all tracks).  Otherwise as fx_analyse_host (HOST pointers, pipelined copies, state carries
 * over).  fx_decode_pcm_device is the conversion alone on DEVICE pointers, asynchronous on `stream`. */
fx_status   fx_analyse_host_pcm  (fx_engine* e, const void* pcm, int format, int n_channels, int channel,
                                  long track_stride_bytes, long n_samples,
                                  float* raw, float* smooth, float* diag, long* n_frames);
fx_status   fx_decode_pcm_device (fx_engine* e, const void* d_pcm, int format, int n_channels, int channel,
                                  long track_stride_bytes, long n_samples, long n_tracks,
                                  float* d_audio, long audio_stride, void* stream);

/* ---- real-time path --------------------------------------------------------------------------------
 * fx_push_block replaces AudioDataCollector::audioDeviceIOCallback (AudioDataCollector.h:36-70) for a
 * range of tracks: copies n_samples of each channel into the pinned host ring (wait-free: memcpy + index
 * publish; no allocation, no CUDA call).  channels[i] is track first_track + i.
 * fx_process replaces the analyser threads' wake-up (RealTimeAnalyser.h:141-177, :201-234): streams the
 * new samples to the device (cudaMemcpyAsync on the group's stream), analyses every hop that became
 * complete and brings the smoothed features back to host memory.  Returns the number of new frames per
 * track in *n_new_frames (may be 0).  fx_poll_features replaces AudioFeatures::getValue x 12 as read by the
 * OSC / GUI timers (OSCFeatureAnalysisOutput.h:91-104): lock-free snapshot of the latest smoothed vector. */
fx_status   fx_push_block    (fx_engine* e, int first_track, int n_tracks, const float* const* channels, int n_samples);
fx_status   fx_process       (fx_engine* e, long* n_new_frames);
fx_status   fx_poll_features (fx_engine* e, int track, float out12[FX_NUM_FEATURES], uint64_t* frame_index);
fx_status   fx_flush         (fx_engine* e);             /* synchronise every group stream */

/* OSC argument order (OSCFeatureAnalysisOutput.h:107): reorder one 12-slot feature vector.
 * n_out = 12 (as the code sends) or 10 (as README.md:55-57 documents). */
fx_status   fx_osc_order (const float in12[FX_NUM_FEATURES], float* out, int n_out);

/* ---- measurement support (bench.py; not part of the reference surface) -------------------------------
 * Fill d_audio [n_tracks][track_stride] with the synthetic workload of SURVEY.md section 8d:
 * A sin(2 pi f_t n / sr + phi_t) + sigma_t u_t[n], Philox-4x32-10 keyed by (seed, track), with the
 * silence / burst insertions.  first_track offsets the track index (multi-GPU sharding by track range). */
fx_status   fx_synth_device (fx_engine* e, float* d_audio, long track_stride, long n_samples,
                             long first_track, uint64_t seed, void* stream);
/* number of kernels this engine has launched since creation */
uint64_t    fx_kernel_launches (const fx_engine* e);
/* Per-kernel device timing: when enabled, every call brackets the analysis kernel (k_analyse) and the small
 * post kernels with CUDA events on the stream they are launched on.  fx_profile_read synchronises those events
 * and returns the accumulated milliseconds and the number of bracketed calls since the last read. */
fx_status   fx_profile_enable (fx_engine* e, int on);
fx_status   fx_profile_read   (fx_engine* e, double* ms_analyse, double* ms_post, long* n_calls);
/* FP32 FMA microbenchmark (registers only) for the compute roofline denominator: achieved TFLOP/s on `device`. */
fx_status   fx_measure_fp32_peak (int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
