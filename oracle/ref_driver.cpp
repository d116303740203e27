/*
 * ref_driver.cpp -- builds oracle/_ref/libfxref.so: the REFERENCE'S OWN analysis classes, compiled
 * headless.  TEST INFRASTRUCTURE ONLY.
 *
 * The six hot-path headers are #included verbatim from where they lie under /root/reference/Source
 * (passed with -I by oracle/Makefile; never copied into this repository), in the order of the
 * reference's unity include (Source/include.h:27-34), on top of oracle/juce_shim/JuceHeader.h.
 *
 * Two drive modes (fxo_config.mode):
 *   A (0)  verbatim: AudioDataCollector::audioDeviceIOCallback -> RealTimeAudioDataOverlapper ->
 *          RealTimeSpectralAnalyser::run / RealTimeHarmonicAnalyser::run single-stepped, wired the way
 *          AnalyserTrackController.h:18-41 wires them (two collectors, two analysers, ONE AudioFeatures).
 *          Hop is N/2 by construction (RealTimeAudioAnalysis.h:207).
 *   B (1)  re-sequenced: the same calls the two run() bodies make (RealTimeAnalyser.h:145-172, :205-229),
 *          in the same order, on frames this driver cuts at an arbitrary hop.  Identical to A at hop N/2
 *          (tests/test_oracle.py checks it bit for bit).
 *
 * Determinism decision: per hop the spectral body runs first, then the harmonic body,
 * on one thread; both push RMS into the shared AudioFeatures exactly as in the app.
 */
#include "juce_shim/JuceHeader.h"

#include <thread>
#include <memory>

// expose the analysers' internals (raw pushed values, private helpers) to this driver only
#define private public
#include "AudioDataCollector.h"
#include "RealTimeAudioAnalysis.h"
#include "PitchAnalyser.h"
#include "SpectralCharacteristics.h"
#include "HarmonicCharacteristics.h"
#include "RealTimeAnalyser.h"
// the legacy offline analyser (never instantiated by the application; driven by fxo_legacy_analyse below)
#include "AudioFeatures.h"
#include "AudioAnalysis.h"
#undef private

#include "fx_oracle_api.h"

namespace
{
struct TrackRig
{
    TrackRig (const fxo_config& c)
        : cfg (c),
          collectorHarm (0), collectorSpec (0),
          harm (collectorHarm, features, c.window),
          spec (collectorSpec, features, c.window)
    {
        // AnalyserTrackController::prepareToPlay (AnalyserTrackController.h:175-188)
        harm.sampleRateChanged (c.sample_rate);
        spec.sampleRateChanged (c.sample_rate);
        collectorHarm.setExpectedSamplesPerBlock (c.window / 2);
        collectorSpec.setExpectedSamplesPerBlock (c.window / 2);
        collectorHarm.setGain (c.gain);
        collectorSpec.setGain (c.gain);
        // onset parameter surface (RealTimeAnalyser.h:244-258); only touched when non-default so that the
        // default path is exactly what the app does (the GUI never creates the onset widgets)
        if (c.onset_multiplier != 1.7f)
            spec.getOnsetDetector().meanThresholdMultiplier = c.onset_multiplier;
        if (c.onset_hist != 5)
            spec.setOnsetWindowLength (c.onset_hist);
        if (c.onset_type != (int) OnsetDetector::enAmplitude)
            spec.setOnsetDetectionType ((OnsetDetector::eOnsetDetectionType) c.onset_type);
    }

    // ---- mode A --------------------------------------------------------------------------------
    void hopVerbatim (const float* newSamples)
    {
        const int h = cfg.window / 2;
        const float* in[1] = { newSamples };
        float* out[1] = { nullptr };
        collectorSpec.audioDeviceIOCallback (in, 1, out, 0, h);
        collectorHarm.audioDeviceIOCallback (in, 1, out, 0, h);
        spec.stepOnce();
        harm.stepOnce();
    }

    // ---- mode B --------------------------------------------------------------------------------
    void spectralBody (const AudioSampleBuffer& frame)
    {
        // RealTimeAnalyser.h:205-229
        AudioSampleBuffer audioWindow (frame);
        float rms = audioWindow.getRMSLevel (0, 0, audioWindow.getNumSamples());
        float logRMS = log10 (rms * 9.0f + 1.0f);
        spec.getFeatures().updateFeature (AudioFeatures::enRMS, logRMS);
        RealTimeWindower::scaleBufferWithBartlettWindowing (audioWindow);
        AudioSampleBuffer frequencyBuffer = spec.getFFTAnalyser().getFrequencyData (audioWindow);
        SpectralCharacteristics sf = spec.getSpectralAnalyser().calculateSpectralCharacteristics (frequencyBuffer, logRMS, 0, spec.getFFTAnalyser().getNyquist());
        spec.getFeatures().updateFeature (AudioFeatures::enCentroid, sf.centroid);
        spec.getFeatures().updateFeature (AudioFeatures::enFlatness, sf.flatness);
        spec.getFeatures().updateFeature (AudioFeatures::enLER,      sf.ler);
        spec.getFeatures().updateFeature (AudioFeatures::enSpread,   sf.spread);
        spec.getFeatures().updateFeature (AudioFeatures::enFlux,     sf.flux);
        spec.getFeatures().updateFeature (AudioFeatures::enSlope,    spec.getSpectralAnalyser().calculateNormalisedSpectralSlope (frequencyBuffer, 0));
        spec.getFeatures().updateFeature (AudioFeatures::enOnset,    spec.detectOnset());
    }

    void harmonicBody (const AudioSampleBuffer& frame)
    {
        // RealTimeAnalyser.h:145-172
        FFTAnalyser& fftAnalyser = harm.getFFTAnalyser();
        AudioSampleBuffer audioWindow (frame);
        float rms = audioWindow.getRMSLevel (0, 0, audioWindow.getNumSamples());
        float logRMS = log10 (rms * 9.0f + 1.0f);
        if (cfg.rms_pushes >= 2)
            harm.getFeatures().updateFeature (AudioFeatures::enRMS, logRMS);
        AudioSampleBuffer filteredAudio (audioWindow);
        filteredAudio.clear();
        harm.filter.filterAudio (audioWindow, filteredAudio);
        RealTimeWindower::scaleBufferWithBartlettWindowing (filteredAudio);
        AudioSampleBuffer filteredFrequencyBuffer = fftAnalyser.getFrequencyData (filteredAudio);
        AudioSampleBuffer frequencyBuffer = fftAnalyser.getFrequencyData (audioWindow);
        double f0Estimate = harm.getPitchAnalyser().estimatePitch (filteredFrequencyBuffer);
        const double f0NormalisationFactor = 5000.0;
        harm.getFeatures().updateFeature (AudioFeatures::enF0, (float) (f0Estimate / f0NormalisationFactor));
        HarmonicCharacteristics hf = harm.getHarmonicAnalyser().calculateHarmonicCharacteristics (frequencyBuffer, f0Estimate, fftAnalyser.getNyquist(), 0);
        harm.getFeatures().updateFeature (AudioFeatures::enHarmonicEnergyRatio,  hf.harmonicEnergyRatio);
        harm.getFeatures().updateFeature (AudioFeatures::enOddEvenHarmonicRatio, hf.harmonicEnergyRatio);
        harm.getFeatures().updateFeature (AudioFeatures::enInharmonicity,        hf.inharmonicity);
        lastTrueOER = hf.oddEvenHarmonicRatio;
    }

    void snapshot (float* raw, float* smooth, float* diag)
    {
        for (int f = 0; f < FXO_NUM_FEATURES; ++f)
        {
            if (raw != nullptr)    raw[f]    = features.smoothedFeatures[(size_t) f].history.back();
            if (smooth != nullptr) smooth[f] = features.getValue ((AudioFeatures::eAudioFeature) f);
        }
        if (diag != nullptr)
        {
            for (int d = 0; d < FXO_NUM_DIAG; ++d) diag[d] = -1.0f;
            const float f0 = features.smoothedFeatures[AudioFeatures::enF0].history.back() * 5000.0f;
            diag[FXO_DIAG_LAG] = (float) floor (cfg.sample_rate / (double) f0 + 0.5);
            diag[FXO_DIAG_TRUE_OER] = lastTrueOER;
        }
    }

    fxo_config               cfg;
    AudioFeatures            features;
    AudioDataCollector       collectorHarm;
    AudioDataCollector       collectorSpec;
    RealTimeHarmonicAnalyser harm;
    RealTimeSpectralAnalyser spec;
    float                    lastTrueOER = -1.0f;
};

long analyseOne (const fxo_config& cfg, const float* audio, long nSamples,
                 float* raw, float* smooth, float* diag, long maxFrames)
{
    const int N = cfg.window, H = cfg.hop;
    long frames = nSamples / H;
    if (frames > maxFrames) frames = maxFrames;
    TrackRig rig (cfg);

    if (cfg.mode == 0)
    {
        if (H != N / 2) return -1;
        for (long f = 0; f < frames; ++f)
        {
            rig.hopVerbatim (audio + f * H);
            // mode A cannot observe the un-stored OER
            rig.snapshot (raw ? raw + f * FXO_NUM_FEATURES : nullptr,
                          smooth ? smooth + f * FXO_NUM_FEATURES : nullptr,
                          diag ? diag + f * FXO_NUM_DIAG : nullptr);
        }
        return frames;
    }

    // mode B: the overlapper's shift-and-append (RealTimeAudioAnalysis.h:205-219) at hop H, with the
    // collector's gain multiply (AudioDataCollector.h:88)
    AudioSampleBuffer frame (1, N);
    frame.clear();
    for (long f = 0; f < frames; ++f)
    {
        float* d = frame.getWritePointer (0);
        for (int i = 0; i + H < N; ++i) d[i] = d[i + H];
        for (int i = 0; i < H; ++i)     d[N - H + i] = audio[f * H + i] * cfg.gain;
        rig.spectralBody (frame);
        rig.harmonicBody (frame);
        rig.snapshot (raw ? raw + f * FXO_NUM_FEATURES : nullptr,
                      smooth ? smooth + f * FXO_NUM_FEATURES : nullptr,
                      diag ? diag + f * FXO_NUM_DIAG : nullptr);
    }
    return frames;
}
} // namespace

extern "C" {

void fxo_default_config (fxo_config* cfg)
{
    cfg->window = 2048;            // AnalyserTrackController.h:20-21
    cfg->hop = 1024;               // RealTimeAudioAnalysis.h:207
    cfg->sample_rate = 48000.0;    // RealTimeAnalyser.h:100
    cfg->gain = 1.0f;              // AudioDataCollector.h:129
    cfg->onset_type = 1;           // SpectralCharacteristics.h:240
    cfg->onset_hist = 5;           // SpectralCharacteristics.h:238-239
    cfg->onset_multiplier = 1.7f;  // SpectralCharacteristics.h:311
    cfg->rms_pushes = 2;           // RealTimeAnalyser.h:150,209
    cfg->mode = 1;
}

const char* fxo_kind (void) { return "reference"; }

long fxo_analyse_track (const fxo_config* cfg, const float* audio, long n_samples,
                        float* raw, float* smooth, float* diag, long max_frames)
{
    return analyseOne (*cfg, audio, n_samples, raw, smooth, diag, max_frames);
}

long fxo_analyse_tracks (const fxo_config* cfg, const float* audio, long n_tracks, long track_stride, long n_samples,
                         float* raw, float* smooth, float* diag, long max_frames, int n_threads)
{
    long frames = n_samples / cfg->hop;
    if (frames > max_frames) frames = max_frames;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_tracks) n_threads = (int) n_tracks;
    std::vector<std::thread> pool;
    for (int w = 0; w < n_threads; ++w)
    {
        const long t0 = n_tracks * w / n_threads, t1 = n_tracks * (w + 1) / n_threads;
        pool.emplace_back ([=]()
        {
            for (long t = t0; t < t1; ++t)
                analyseOne (*cfg, audio + t * track_stride, n_samples,
                            raw    ? raw    + t * frames * FXO_NUM_FEATURES : nullptr,
                            smooth ? smooth + t * frames * FXO_NUM_FEATURES : nullptr,
                            diag   ? diag   + t * frames * FXO_NUM_DIAG     : nullptr, frames);
        });
    }
    for (auto& th : pool) th.join();
    return frames;
}

long fxo_legacy_analyse (int window, double sample_rate, const float* audio, long n_samples, int n_frames,
                         float* out, float* log_attack)
{
    if (window < 16 || (window & (window - 1)) != 0 || n_frames < 1 || n_samples < n_frames) return -1;
    const int N = window;
    AudioAnalyser analyser (N, 1, sample_rate / 2.0, true, true);
    ConcatenatedFeatureBuffer features (1, (int) n_samples, n_frames, N / 2 + 1, 1000.0 * (double) n_samples / sample_rate, sample_rate);
    for (long i = 0; i < n_samples; ++i) features.audioOutput.setSample (0, (int) i, audio[i]);

    // ---- performSpectralAnalysis (AudioAnalysis.h:121-251) with the commented-out feature block (:219-247) reinstated -------
    const int numInputSamples = features.audioOutput.getNumSamples();
    const int stepSize = numInputSamples / features.numDownsamples;                     // :130
    const int numFFTInputSamples = analyser.fftIn.getNumSamples();
    features.nyquistFrequency = features.sampleRate / 2.0;                              // :139
    features.energyEnvelope.setSize (1, n_frames);                                     // :140
    std::vector<float> temp ((size_t) 2 * N);                // (the reference's float temp[4096] is too small for its own call at N > 2048)
    std::vector<int> numPeaks ((size_t) n_frames, 0);
    for (int frame = 0; frame < n_frames; ++frame)
    {
        analyser.fftIn.clear();                                                         // :144
        int rangeStart = -numFFTInputSamples / 2, rangeEnd = numFFTInputSamples / 2, offsetToAdd = numFFTInputSamples / 2;   // :150-152
        if (n_frames == 1) { rangeStart = 0; rangeEnd = numFFTInputSamples; offsetToAdd = 0; }                               // :155-161
        for (int sample = rangeStart; sample < rangeEnd; ++sample)                      // :165-179
        {
            const int i = frame * stepSize + sample;
            if (i < 0 || i >= numInputSamples) analyser.fftIn.getWritePointer (0)[sample + numFFTInputSamples / 2] = 0.0f;
            else                               analyser.fftIn.getWritePointer (0)[sample + offsetToAdd] = features.audioOutput.getReadPointer (0)[i];
        }
        AudioAnalyser::scaleBufferWithBartlettWindowing (analyser.fftIn);              // :181
        for (int i = 0; i < N; ++i) { temp[(size_t) i] = analyser.fftIn.getReadPointer (0)[i]; temp[(size_t) (N + i)] = 0.0f; }
        analyser.fft.performFrequencyOnlyForwardTransform (temp.data());                // :208 / :224
        for (int i = 0; i <= N / 2; ++i) analyser.fftOut.setSample (0, i, temp[(size_t) i]);          // :226-227: N/2 + 1 amplitudes
        float* o = out + (long) frame * FXL_NUM;
        const AudioAnalyser::SpectralCharacteristics sc = analyser.calculateSpectralCharacteristics (analyser.fftOut, 0);   // :231
        o[FXL_CENTROID] = sc.centroid; o[FXL_SPREAD] = sc.spread; o[FXL_FLATNESS] = sc.flatness; o[FXL_FLUX] = sc.flux;   // :232-235
        o[FXL_SLOPE] = analyser.calculateNormalisedSpectralSlope (analyser.fftOut, 0);                                        // :236
        {
            // (the peak count is a diagnostic: the same test calculateHarmonicCharacteristics applies, :361-369)
            double sum = 0.0;
            for (int b = 0; b <= N / 2; ++b) sum += (double) analyser.fftOut.getSample (0, b);
            int np = 0;
            if (! (sum < 0.001)) for (int b = 0; b <= N / 2; ++b) np += analyser.binIsPeak (b, analyser.fftOut, 0, sum / (double) (N / 2 + 1)) ? 1 : 0;
            numPeaks[(size_t) frame] = np;
        }
        const AudioAnalyser::HarmonicCharacteristics hc = analyser.calculateHarmonicCharacteristics (analyser.fftOut, 0);   // :242
        o[FXL_F0] = hc.f0; o[FXL_HER] = hc.harmonicEnergyRatio; o[FXL_INHARM] = hc.inharmonicity;                          // :243-245
        o[FXL_NUM_PEAKS] = (float) numPeaks[(size_t) frame];
        o[FXL_MARGIN] = -1.0f;
        features.energyEnvelope.setSample (0, frame, AudioAnalyser::sumAccrossChannels (analyser.fftOut));                   // :249
        o[FXL_ENERGY] = features.energyEnvelope.getSample (0, frame);
    }
    analyser.analyseNormalisedZeroCrosses (features);                                   // :517-541
    for (int frame = 0; frame < n_frames; ++frame)
        out[(long) frame * FXL_NUM + FXL_ZCR] = features.getFeatureSample (ConcatenatedFeatureBuffer::ZeroCrosses, 0, frame);
    analyser.setLogAttackTime (features);                                               // :611-622
    if (log_attack != nullptr) *log_attack = features.estimatedLogAttackTime;
    return n_frames;
}

void fxo_fft_forward (const float* frame, int n, float* out_2n)
{
    RealTimeFFT fft (n);
    for (int i = 0; i < n; ++i) { out_2n[i] = frame[i]; out_2n[n + i] = 0.0f; }
    fft.performForward (out_2n, 2 * n);
}

void fxo_fft_inverse (float* inout_2n, int n)
{
    RealTimeFFT fft (n);
    fft.performInverse (inout_2n, 2 * n);
}

} // extern "C"
