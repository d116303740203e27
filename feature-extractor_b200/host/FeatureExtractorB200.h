// FeatureExtractorB200.h -- C++ host facade: the reference's per-track analyser surface, rebuilt over the C ABI
// of libfxb200.so (include/fx_engine.h).  Header-only; link with -lfxb200.
//
// Same class names, method names and argument meaning as the reference (all citations relative to
// /root/reference/Source/), so code written against
//     AudioDataCollector / RealTimeAnalyser / RealTimeSpectralAnalyser / RealTimeHarmonicAnalyser /
//     AudioFeatures / OSCFeatureAnalysisOutput / AnalyserTrackController
// keeps compiling when it switches `#include "include.h"` for this header and `juce::` device plumbing for
// fxb200::AudioDeviceManager.  What changed underneath:
//   * AudioDataCollector's 4096-float ring + spinning consumer (AudioDataCollector.h:24,72-94) is the engine's
//     pinned host ring, streamed to the GPU with cudaMemcpyAsync per track group (fx_push_block / fx_process);
//   * the two analyser threads per track (AnalyserTrackController.h:184-185) are one GPU launch sequence for
//     all tracks: startThread / notify / stopThread keep their meaning (start analysing / new data / stop),
//     but no CPU thread spins;
//   * AudioFeatures::getValue reads the engine's latest smoothed vector (fx_poll_features).
// Everything lives in namespace fxb200 so that it can coexist with JUCE.
#pragma once

#include "../../include/fx_engine.h"

#include <arpa/inet.h>
#include <netdb.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fxb200 {

using String = std::string;

// ------------------------------------------------------------------------------------------------------
// The slice of juce::AudioSampleBuffer the scope callback of AudioDataCollector hands out (one channel).
class AudioSampleBuffer
{
public:
    AudioSampleBuffer() = default;
    AudioSampleBuffer (int numChannels, int numSamples) : data ((size_t) (numChannels > 0 ? numSamples : 0), 0.0f) {}
    int getNumChannels() const noexcept { return 1; }
    int getNumSamples() const noexcept  { return (int) data.size(); }
    const float* getReadPointer (int) const noexcept { return data.data(); }
    float* getWritePointer (int) noexcept            { return data.data(); }
    float getSample (int, int i) const noexcept      { return data[(size_t) i]; }
    void setSample (int, int i, float v) noexcept    { data[(size_t) i] = v; }
private:
    std::vector<float> data;
};

// juce::AudioIODeviceCallback, as far as AudioDataCollector uses it (AudioDataCollector.h:18,28-38)
class AudioIODevice;
class AudioIODeviceCallback
{
public:
    virtual ~AudioIODeviceCallback() = default;
    virtual void audioDeviceIOCallback (const float** inputChannelData, int numInputChannels,
                                        float** outputChannelData, int numOutputChannels, int numSamples) = 0;
    virtual void audioDeviceAboutToStart (AudioIODevice*) {}
    virtual void audioDeviceStopped() {}
};

// ------------------------------------------------------------------------------------------------------
// AudioDeviceManager: stands where juce::AudioDeviceManager stood in AnalyserTrackController's constructor
// (AnalyserTrackController.h:17,35,41).  It owns the GPU engine for all its input channels.  The host's audio
// thread calls processBlock() with the device's channel pointers; the manager fans the block out to the
// registered collectors exactly as JUCE's device manager calls each AudioIODeviceCallback, then wakes the analysis.
class AudioDeviceManager
{
public:
    struct AudioDeviceSetup { int bufferSize = 512; double sampleRate = 48000.0; };

    AudioDeviceManager (int numInputChannels, double sampleRate = 48000.0, int bufferSize = 512,
                        int windowSize = 2048, int device = 0, int tracksPerGroup = 0)
    {
        fx_config cfg;
        fx_default_config (&cfg);
        cfg.n_tracks = numInputChannels;
        cfg.window = windowSize;                 // AnalyserTrackController.h:20-21
        cfg.hop = windowSize / 2;                // RealTimeAudioAnalysis.h:207
        cfg.sample_rate = sampleRate;
        cfg.device = device;
        cfg.tracks_per_group = tracksPerGroup;
        cfg.ring_hops = 16;
        setup.bufferSize = bufferSize;
        setup.sampleRate = sampleRate;
        numChannels = numInputChannels;
        hopSize = cfg.hop;
        if (fx_engine_create (&cfg, &engine) != FX_OK)
            throw std::runtime_error (String ("fx_engine_create: ") + fx_last_error (nullptr));   // no CPU fallback
    }

    ~AudioDeviceManager() { if (engine) fx_engine_destroy (engine); }
    AudioDeviceManager (const AudioDeviceManager&) = delete;
    AudioDeviceManager& operator= (const AudioDeviceManager&) = delete;

    void addAudioCallback (AudioIODeviceCallback* cb)    { callbacks.push_back (cb); }
    void removeAudioCallback (AudioIODeviceCallback* cb)
    {
        for (size_t i = 0; i < callbacks.size(); ++i)
            if (callbacks[i] == cb) { callbacks.erase (callbacks.begin() + (long) i); break; }
    }
    void getAudioDeviceSetup (AudioDeviceSetup& s) const { s = setup; }

    // The audio thread's entry point: one device block for every channel.  Collectors copy into the pinned ring
    // and ask for a wake-up (Thread::notify in the reference); the wake-up itself happens once per block.
    void processBlock (const float** inputChannelData, int numInputChannels, int numSamples)
    {
        for (auto* cb : callbacks)
            cb->audioDeviceIOCallback (inputChannelData, numInputChannels, nullptr, 0, numSamples);
        if (wakeRequested)
        {
            wakeRequested = false;
            if (pump() > 0)
                for (auto& h : afterAnalysis) h.second();
        }
    }
    void requestWake() noexcept { wakeRequested = true; }
    void addAfterAnalysisHook (const void* owner, std::function<void()> f) { afterAnalysis.emplace_back (owner, std::move (f)); }
    void removeAfterAnalysisHooks (const void* owner)
    {
        for (size_t i = afterAnalysis.size(); i-- > 0;)
            if (afterAnalysis[i].first == owner) afterAnalysis.erase (afterAnalysis.begin() + (long) i);
    }

    // The analysis wake-up (what Thread::notify() led to in the reference).  Returns the number of new hops analysed.
    long pump()
    {
        long n = 0;
        if (analysing && fx_process (engine, &n) != FX_OK)
            throw std::runtime_error (String ("fx_process: ") + fx_last_error (engine));
        return n;
    }

    fx_engine* getEngine() const noexcept { return engine; }
    int getNumInputChannels() const noexcept { return numChannels; }
    long getHopSize() const noexcept { return hopSize; }
    void setAnalysing (bool on) noexcept { analysing = on; }

    // one collector per track feeds the ring; the reference's second collector per track carries the same samples
    bool claimTrack (int track, const void* owner)
    {
        if (track < 0 || track >= numChannels) return false;
        if (owners.size() < (size_t) numChannels) owners.resize ((size_t) numChannels, nullptr);
        if (owners[(size_t) track] == nullptr) owners[(size_t) track] = owner;
        return owners[(size_t) track] == owner;
    }
    void releaseTrack (int track, const void* owner)
    {
        if (track >= 0 && (size_t) track < owners.size() && owners[(size_t) track] == owner) owners[(size_t) track] = nullptr;
    }

private:
    fx_engine* engine = nullptr;
    AudioDeviceSetup setup;
    int numChannels = 0;
    long hopSize = 0;
    bool analysing = true;
    bool wakeRequested = false;
    std::vector<std::pair<const void*, std::function<void()>>> afterAnalysis;
    std::vector<AudioIODeviceCallback*> callbacks;
    std::vector<const void*> owners;
};

// ------------------------------------------------------------------------------------------------------
// AudioFeatures (RealTimeAnalyser.h:14-92): same enum, same accessors.  Bound to an engine track it reads the
// GPU's smoothed vector; unbound it is the reference's host-side store of ValueHistory moving averages.
struct AudioFeatures
{
    enum eAudioFeature
    {
        enOnset = 0, enRMS, enF0, enCentroid, enSpread, enFlatness, enLER, enFlux, enSlope,
        enHarmonicEnergyRatio, enOddEvenHarmonicRatio, enInharmonicity, numFeatures
    };

    static String getFeatureName (eAudioFeature f)
    {
        static const char* names[] = { "Onset", "Amp.", "Pitch", "Centroid", "Spread", "Flatness", "L.E.R", "Flux",
                                       "Slope", "H.E.R", "O.E.R", "Inharm." };
        return f >= 0 && f < numFeatures ? names[f] : "";
    }
    static float getMaxValueForFeature (eAudioFeature) { return 1.0f; }

    AudioFeatures()
    {
        for (int f = 0; f < numFeatures; ++f)
            local.emplace_back ((f == enOnset || f == enFlux) ? 1 : 10);      // RealTimeAnalyser.h:72-73
    }

    void bind (fx_engine* e, int trackIndex) { engine = e; track = trackIndex; }

    void updateFeature (eAudioFeature f, float v)                              // RealTimeAnalyser.h:76-82
    {
        History& h = local[(size_t) f];
        for (size_t i = 0; i + 1 < h.v.size(); ++i) h.v[i] = h.v[i + 1];
        h.v.back() = v;
        if (h.recorded < (int) h.v.size()) h.recorded++;
    }

    float getValue (eAudioFeature f) const                                     // RealTimeAnalyser.h:84-88
    {
        if (engine != nullptr)
        {
            float v[FX_NUM_FEATURES];
            if (fx_poll_features (engine, track, v, nullptr) == FX_OK) return v[(int) f];
            return NAN;
        }
        const History& h = local[(size_t) f];
        float total = 0.0f;
        for (float x : h.v) total += x;
        return total / (float) h.recorded;
    }

    // all 12 at once + the index of the hop they belong to
    bool snapshot (float out12[FX_NUM_FEATURES], uint64_t* frameIndex = nullptr) const
    {
        return engine != nullptr && fx_poll_features (engine, track, out12, frameIndex) == FX_OK;
    }

private:
    struct History { explicit History (int n) : v ((size_t) n, 0.0f) {} std::vector<float> v; int recorded = 0; };
    std::vector<History> local;
    fx_engine* engine = nullptr;
    int track = -1;
};

// ------------------------------------------------------------------------------------------------------
// AudioDataCollector (AudioDataCollector.h:18-138)
class AudioDataCollector : public AudioIODeviceCallback
{
public:
    explicit AudioDataCollector (int audioChannelToCollect) : channelToCollect (audioChannelToCollect) {}
    ~AudioDataCollector() override { if (manager) manager->releaseTrack (channelToCollect, this); }
    AudioDataCollector (const AudioDataCollector&) = delete;
    AudioDataCollector& operator= (const AudioDataCollector&) = delete;

    void attach (AudioDeviceManager& m) { manager = &m; primary = m.claimTrack (channelToCollect, this); }

    // Audio thread.  Wait-free: a memcpy into the pinned ring and an index publish (fx_push_block).
    void audioDeviceIOCallback (const float** inputChannelData, int numInputChannels,
                                float** outputChannelData, int numOutputChannels, int numberOfSamples) override
    {
        const float* const* channelData = collectInput ? inputChannelData : (const float* const*) outputChannelData;   // :42
        const int available = collectInput ? numInputChannels : numOutputChannels;
        if (channelData == nullptr || channelToCollect < 0 || channelToCollect >= available || manager == nullptr) return;
        const float* src = channelData[channelToCollect];
        if (primary)
        {
            const float* one[1] = { src };
            lastStatus = fx_push_block (manager->getEngine(), channelToCollect, 1, one, numberOfSamples);
        }
        if (bufferToDrawUpdated)                                                                       // :96-102
        {
            AudioSampleBuffer b (1, numberOfSamples);
            for (int i = 0; i < numberOfSamples; ++i) b.setSample (0, i, src[i] * gain);
            bufferToDrawUpdated (b);
        }
        if (notifyAnalysisThread) notifyAnalysisThread();                                              // :68-69
    }

    void setBufferToDrawUpdatedCallback  (std::function<void (AudioSampleBuffer&)> f) { bufferToDrawUpdated = std::move (f); }
    void setNotifyAnalysisThreadCallback (std::function<void()> f)                    { notifyAnalysisThread = std::move (f); }

    void toggleCollectInput (bool shouldCollectInput) noexcept { clearBuffer(); collectInput = shouldCollectInput; }   // :119
    void setExpectedSamplesPerBlock (int spb) noexcept         { expectedSamplesPerBlock = spb; }                    // :120
    void clearBuffer() {}                          // :122 -- the pinned ring only ever exposes samples that were pushed
    void setChannelToCollect (int c)
    {
        if (manager) manager->releaseTrack (channelToCollect, this);
        channelToCollect = c;
        if (manager) primary = manager->claimTrack (c, this);
    }
    void setGain (float g)                                                                             // :124
    {
        gain = g;
        if (manager) fx_set_gain (manager->getEngine(), channelToCollect, g);
    }
    int  getChannel() const noexcept   { return channelToCollect; }
    int  getLastStatus() const noexcept { return lastStatus; }      // FX_ERR_OVERRUN when the analysis fell behind the producer
    AudioDeviceManager* getManager() const noexcept { return manager; }

private:
    AudioDeviceManager* manager = nullptr;
    std::function<void (AudioSampleBuffer&)> bufferToDrawUpdated;
    std::function<void()> notifyAnalysisThread;
    float gain = 1.0f;
    int expectedSamplesPerBlock = 512;
    int channelToCollect = 0;
    bool collectInput = true;
    bool primary = false;
    int lastStatus = FX_OK;
};

// ------------------------------------------------------------------------------------------------------
struct OnsetDetector
{
    enum eOnsetDetectionType { enSpectral = 0, enAmplitude, enCombination, enNumTypes };     // SpectralCharacteristics.h:213-219
    static String getStringForDetectionType (eOnsetDetectionType t)
    {
        switch (t) { case enSpectral: return "Spectral"; case enAmplitude: return "Amplitude"; case enCombination: return "Combination"; default: return "UNKNOWN"; }
    }
};

// RealTimeAnalyser (RealTimeAnalyser.h:97-127): the juce::Thread surface is kept, the thread is the GPU.
class RealTimeAnalyser
{
public:
    RealTimeAnalyser (AudioDataCollector& adc, AudioFeatures& featuresRef, int windowSize, double sampleRate = 48000.0)
        : audioDataCollector (adc), features (featuresRef), window (windowSize), rate (sampleRate) {}
    virtual ~RealTimeAnalyser() = default;
    RealTimeAnalyser (const RealTimeAnalyser&) = delete;
    RealTimeAnalyser& operator= (const RealTimeAnalyser&) = delete;

    void sampleRateChanged (double newSampleRate) { rate = newSampleRate; }      // :111-114 (the engine's rate is fixed at creation)
    void startThread (int /*priority*/ = 5)       { running = true; }
    bool stopThread (int /*timeoutMs*/)           { running = false; return true; }
    bool isThreadRunning() const noexcept         { return running; }
    // new audio has arrived: ask the manager to analyse whatever hops are complete (all tracks at once, once per block)
    void notify()
    {
        if (running && audioDataCollector.getManager() != nullptr) audioDataCollector.getManager()->requestWake();
    }
    AudioFeatures& getFeatures() { return features; }
    int getWindowSize() const noexcept { return window; }
    double getSampleRate() const noexcept { return rate; }

protected:
    fx_engine* engineOrNull() const { return audioDataCollector.getManager() ? audioDataCollector.getManager()->getEngine() : nullptr; }
    AudioDataCollector& audioDataCollector;
    AudioFeatures& features;
    int window;
    double rate;
    bool running = false;
};

class RealTimeHarmonicAnalyser : public RealTimeAnalyser       // RealTimeAnalyser.h:133-188
{
public:
    using RealTimeAnalyser::RealTimeAnalyser;
};

class RealTimeSpectralAnalyser : public RealTimeAnalyser       // RealTimeAnalyser.h:193-269
{
public:
    using RealTimeAnalyser::RealTimeAnalyser;

    void setOnsetDetectionSensitivity (float s)                                  // :244-248
    {
        multiplier = 1.0f + s;
        apply();
    }
    void setOnsetWindowLength (int length)                                       // :250-254
    {
        histLen = length;
        apply();
    }
    void setOnsetDetectionType (OnsetDetector::eOnsetDetectionType t)            // :258
    {
        type = t;
        apply();
    }
    void setOnsetDetectedCallback (std::function<void()> f) { onsetDetectedCallback = std::move (f); }     // :256

    // call after notify(): fires the callback when the newest hop carries an onset (:228-229)
    void dispatchOnsetCallback()
    {
        if (onsetDetectedCallback && features.getValue (AudioFeatures::enOnset) > 0.0f) onsetDetectedCallback();
    }

private:
    void apply()
    {
        if (fx_engine* e = engineOrNull())
            fx_set_onset (e, audioDataCollector.getChannel(), (int) type, histLen, multiplier);
    }
    OnsetDetector::eOnsetDetectionType type = OnsetDetector::enAmplitude;        // SpectralCharacteristics.h:240
    int histLen = 5;                                                             // :238-239
    float multiplier = 1.7f;                                                     // :311
    std::function<void()> onsetDetectedCallback;
};

// ------------------------------------------------------------------------------------------------------
// OSCFeatureAnalysisOutput (OSCFeatureAnalysisOutput.h:25-145): OSC 1.0 message over UDP,
// address pattern = bundleAddress, type tags ",ffffffffffff", twelve big-endian floats in the order of :107.
class OSCFeatureAnalysisOutput
{
public:
    OSCFeatureAnalysisOutput (AudioFeatures& rta, String ip, String bundle)
        : realTimeAudioFeatures (rta), address (std::move (ip)), bundleAddress (std::move (bundle))
    {
        if (! bundleAddress.empty()) connectToAddress (address);                  // :80-81
    }
    ~OSCFeatureAnalysisOutput() { if (sock >= 0) ::close (sock); }
    OSCFeatureAnalysisOutput (const OSCFeatureAnalysisOutput&) = delete;
    OSCFeatureAnalysisOutput& operator= (const OSCFeatureAnalysisOutput&) = delete;

    void timerCallback() { sendSpectralFeaturesViaOSC (true); }                   // :84-87 (60 Hz in the reference, :133)

    // builds the message; returns the bytes that went (or would go) on the wire
    std::vector<uint8_t> encode() const
    {
        float v[FX_NUM_FEATURES], o[FX_OSC_FLOATS_CODE];
        for (int f = 0; f < FX_NUM_FEATURES; ++f) v[f] = realTimeAudioFeatures.getValue ((AudioFeatures::eAudioFeature) f);
        fx_osc_order (v, o, FX_OSC_FLOATS_CODE);
        std::vector<uint8_t> m;
        auto padded = [&m] (const String& s)
        {
            m.insert (m.end(), s.begin(), s.end());
            m.push_back (0);
            while (m.size() % 4) m.push_back (0);
        };
        padded (bundleAddress);
        padded (String (",") + String ((size_t) FX_OSC_FLOATS_CODE, 'f'));
        for (int i = 0; i < FX_OSC_FLOATS_CODE; ++i)
        {
            uint32_t bits;
            std::memcpy (&bits, &o[i], 4);
            bits = htonl (bits);
            const uint8_t* b = reinterpret_cast<const uint8_t*> (&bits);
            m.insert (m.end(), b, b + 4);
        }
        return m;
    }

    void sendSpectralFeaturesViaOSC (bool updateHarmonicFeatures)                 // :89-113
    {
        if (! updateHarmonicFeatures) return;                                     // the reference sends nothing in that branch (:109-112)
        const std::vector<uint8_t> m = encode();
        if (sock >= 0) ::send (sock, m.data(), m.size(), 0);
    }

    bool connectToAddress (String newAddress)                                     // :115-136
    {
        int port = 9000;
        const size_t sep = newAddress.find_last_of (':');
        if (sep != String::npos) port = std::atoi (newAddress.substr (sep + 1).c_str());
        address = newAddress.substr (0, newAddress.find (':'));
        if (sock >= 0) { ::close (sock); sock = -1; }
        addrinfo hints{}; hints.ai_family = AF_INET; hints.ai_socktype = SOCK_DGRAM;
        addrinfo* res = nullptr;
        if (getaddrinfo (address.c_str(), std::to_string (port).c_str(), &hints, &res) != 0 || res == nullptr) return false;
        sock = ::socket (res->ai_family, res->ai_socktype, res->ai_protocol);
        const bool ok = sock >= 0 && ::connect (sock, res->ai_addr, res->ai_addrlen) == 0;
        freeaddrinfo (res);
        if (! ok && sock >= 0) { ::close (sock); sock = -1; }
        return ok;
    }

    float getAudioFeature (AudioFeatures::eAudioFeature f) const { return realTimeAudioFeatures.getValue (f); }

    AudioFeatures& realTimeAudioFeatures;
    String address;
    String bundleAddress { "/Audio/Features" };

private:
    int sock = -1;
};

// ------------------------------------------------------------------------------------------------------
// AudioFormatReader: what AudioFormatManager::createReaderFor (AudioFilePlayer.h:47, after registerBasicFormats :17) yields
// for the two basic formats the reference can open -- WAV (RIFF/WAVE: PCM, IEEE float, WAVE_FORMAT_EXTENSIBLE) and AIFF /
// AIFF-C ('NONE', 'sowt', 'fl32').  Only the container is parsed on the host; the sample data stays as it lies in the file
// and is converted on the GPU (kernel k_pcm_decode behind fx_analyse_host_pcm).
class AudioFormatReader
{
public:
    double  sampleRate = 0.0;
    int     numChannels = 0;
    int     bitsPerSample = 0;
    bool    usesFloatingPointData = false;
    long    lengthInSamples = 0;                  // sample frames
    int     pcmFormat = 0;                        // FX_PCM_*
    String  formatName;                           // "WAV file" / "AIFF file" as JUCE names them

    const uint8_t* data() const noexcept { return file.data() + dataOffset; }
    size_t dataBytes() const noexcept { return (size_t) lengthInSamples * (size_t) numChannels * (size_t) fx_pcm_bytes_per_sample (pcmFormat); }

    // nullptr when the file cannot be opened or is not a WAV / AIFF this reader understands (as createReaderFor does)
    static std::unique_ptr<AudioFormatReader> createReaderFor (const String& path)
    {
        FILE* f = fopen (path.c_str(), "rb");
        if (! f) return nullptr;
        std::unique_ptr<AudioFormatReader> r (new AudioFormatReader);
        fseek (f, 0, SEEK_END);
        const long size = ftell (f);
        fseek (f, 0, SEEK_SET);
        if (size < 12) { fclose (f); return nullptr; }
        r->file.resize ((size_t) size);
        const bool ok = fread (r->file.data(), 1, (size_t) size, f) == (size_t) size;
        fclose (f);
        if (! ok) return nullptr;
        return (r->parseWav() || r->parseAiff()) ? std::move (r) : nullptr;
    }
    static std::unique_ptr<AudioFormatReader> createReaderFor (std::vector<uint8_t> bytes)
    {
        std::unique_ptr<AudioFormatReader> r (new AudioFormatReader);
        r->file = std::move (bytes);
        return (r->file.size() >= 12 && (r->parseWav() || r->parseAiff())) ? std::move (r) : nullptr;
    }

private:
    std::vector<uint8_t> file;
    size_t dataOffset = 0;

    bool tag (size_t at, const char* four) const { return at + 4 <= file.size() && memcmp (file.data() + at, four, 4) == 0; }
    uint32_t le32 (size_t at) const { return (uint32_t) file[at] | ((uint32_t) file[at + 1] << 8) | ((uint32_t) file[at + 2] << 16) | ((uint32_t) file[at + 3] << 24); }
    uint16_t le16 (size_t at) const { return (uint16_t) (file[at] | (file[at + 1] << 8)); }
    uint32_t be32 (size_t at) const { return (uint32_t) file[at + 3] | ((uint32_t) file[at + 2] << 8) | ((uint32_t) file[at + 1] << 16) | ((uint32_t) file[at] << 24); }
    uint16_t be16 (size_t at) const { return (uint16_t) (file[at + 1] | (file[at] << 8)); }

    bool finish (size_t offset, size_t bytes, int bits, bool isFloat, bool bigEndian, bool eightBitSigned)
    {
        int fmt = 0;
        if (isFloat)        fmt = bits == 32 ? (bigEndian ? FX_PCM_F32BE : FX_PCM_F32LE) : 0;
        else if (bits == 8) fmt = eightBitSigned ? FX_PCM_S8 : FX_PCM_U8;
        else if (bits == 16) fmt = bigEndian ? FX_PCM_S16BE : FX_PCM_S16LE;
        else if (bits == 24) fmt = bigEndian ? FX_PCM_S24BE : FX_PCM_S24LE;
        else if (bits == 32) fmt = bigEndian ? FX_PCM_S32BE : FX_PCM_S32LE;
        if (fmt == 0 || numChannels < 1 || sampleRate <= 0.0 || offset > file.size()) return false;
        if (bytes > file.size() - offset) bytes = file.size() - offset;           // truncated file: read what is there
        pcmFormat = fmt; bitsPerSample = bits; usesFloatingPointData = isFloat; dataOffset = offset;
        lengthInSamples = (long) (bytes / ((size_t) numChannels * (size_t) fx_pcm_bytes_per_sample (fmt)));
        return true;
    }

    bool parseWav()
    {
        if (! (tag (0, "RIFF") && tag (8, "WAVE"))) return false;
        size_t at = 12;
        int formatTag = 0, bits = 0;
        bool haveFmt = false;
        while (at + 8 <= file.size())
        {
            const size_t len = le32 (at + 4), body = at + 8;
            if (tag (at, "fmt ") && body + 16 <= file.size())
            {
                formatTag = le16 (body); numChannels = le16 (body + 2); sampleRate = (double) le32 (body + 4); bits = le16 (body + 14);
                if (formatTag == 0xFFFE && len >= 40 && body + 26 <= file.size()) formatTag = le16 (body + 24);      // sub-format GUID, first two bytes
                haveFmt = true;
            }
            else if (tag (at, "data") && haveFmt)
            {
                formatName = "WAV file";
                return (formatTag == 1 || formatTag == 3) && finish (body, len, bits, formatTag == 3, false, false);
            }
            at = body + len + (len & 1u);                                         // chunks are word aligned
        }
        return false;
    }

    static double extended80 (const uint8_t* b)                                    // the AIFF sample rate
    {
        const int exponent = ((b[0] & 0x7f) << 8) | b[1];
        uint64_t mant = 0;
        for (int i = 0; i < 8; ++i) mant = (mant << 8) | b[2 + i];
        if (exponent == 0 && mant == 0) return 0.0;
        return std::ldexp ((double) mant, exponent - 16383 - 63) * ((b[0] & 0x80) ? -1.0 : 1.0);
    }

    bool parseAiff()
    {
        if (! (tag (0, "FORM") && (tag (8, "AIFF") || tag (8, "AIFC")))) return false;
        const bool aifc = tag (8, "AIFC");
        size_t at = 12;
        int bits = 0; uint32_t frames = 0; bool haveComm = false, littleEndian = false, isFloat = false;
        while (at + 8 <= file.size())
        {
            const size_t len = be32 (at + 4), body = at + 8;
            if (tag (at, "COMM") && body + 18 <= file.size())
            {
                numChannels = be16 (body); frames = be32 (body + 2); bits = be16 (body + 6); sampleRate = extended80 (file.data() + body + 8);
                if (aifc && len >= 22 && body + 22 <= file.size())
                {
                    if (tag (body + 18, "sowt")) littleEndian = true;
                    else if (tag (body + 18, "fl32") || tag (body + 18, "FL32")) isFloat = true;
                    else if (! tag (body + 18, "NONE")) return false;             // compressed AIFF-C
                }
                haveComm = true;
            }
            else if (tag (at, "SSND") && haveComm && body + 8 <= file.size())
            {
                const size_t offset = be32 (body);
                formatName = "AIFF file";
                const size_t bps = (size_t) ((bits + 7) / 8);
                return finish (body + 8 + offset, (size_t) frames * (size_t) numChannels * bps, (int) bps * 8, isFloat, ! littleEndian, true);
            }
            at = body + len + (len & 1u);
        }
        return false;
    }
};

// ------------------------------------------------------------------------------------------------------
// AudioFilePlayer (AudioFilePlayer.h:14-78).  The reference plays the file through the device output and the collectors
// re-capture one output channel per track (AudioDataCollector.h:42-43); here the loaded file is analysed in one call:
// track t of the engine takes channel t % numChannels of the file, the PCM goes to the GPU at its file width.
// The transport surface (play / pause / stop / restart / hasFile) is kept; it selects what analyseLoadedFile() covers.
class AudioFilePlayer
{
public:
    AudioFilePlayer() = default;
    AudioFilePlayer (const AudioFilePlayer&) = delete;
    AudioFilePlayer& operator= (const AudioFilePlayer&) = delete;

    void setupAudioCallback (AudioDeviceManager& deviceManager) { manager = &deviceManager; }           // :29-33

    void loadFileIntoTransport (const String& audioFile)                                              // :41-60
    {
        playing = false; position = 0;
        currentAudioFileSource = AudioFormatReader::createReaderFor (audioFile);
    }
    void play()    { playing = true; }                                                                  // :62
    void pause()   { playing = false; }                                                                 // :63
    void stop()    { pause(); position = 0; }                                                           // :64
    void restart() { position = 0; }                                                                    // :65
    bool hasFile() const { return currentAudioFileSource != nullptr; }                                  // :67
    bool isPlaying() const noexcept { return playing; }
    const AudioFormatReader* getReader() const noexcept { return currentAudioFileSource.get(); }
    void setPosition (long sampleFrame) { position = sampleFrame < 0 ? 0 : sampleFrame; }

    // Analyse the loaded file from the transport position to its end (complete hops only) on the manager's engine.
    // smoothed receives [tracks][frames][12] in AudioFeatures order; returns the number of frames per track, -1 on error.
    long analyseLoadedFile (std::vector<float>& smoothed, std::vector<float>* raw = nullptr)
    {
        if (! manager || ! currentAudioFileSource) return -1;
        const AudioFormatReader& r = *currentAudioFileSource;
        fx_engine* e = manager->getEngine();
        const long remaining = r.lengthInSamples - position;
        if (remaining <= 0) return 0;
        const size_t frameBytes = (size_t) r.numChannels * (size_t) fx_pcm_bytes_per_sample (r.pcmFormat);
        const int T = manager->getNumInputChannels();
        const long hop = manager->getHopSize();
        const long frames = remaining / hop;
        smoothed.assign ((size_t) T * (size_t) frames * FX_NUM_FEATURES, 0.0f);
        if (raw) raw->assign (smoothed.size(), 0.0f);
        if (frames == 0) return 0;
        long got = 0;
        const fx_status st = fx_analyse_host_pcm (e, r.data() + (size_t) position * frameBytes, r.pcmFormat, r.numChannels, -1, 0, remaining,
                                                  raw ? raw->data() : nullptr, smoothed.data(), nullptr, &got);
        if (st != FX_OK || got != frames) return -1;
        position += frames * hop;
        return frames;
    }

private:
    AudioDeviceManager* manager = nullptr;
    std::unique_ptr<AudioFormatReader> currentAudioFileSource;
    bool playing = false;
    long position = 0;
};

// ------------------------------------------------------------------------------------------------------
// AnalyserTrackController (AnalyserTrackController.h:14-212) without the GUI link and the file player.
class AnalyserTrackController
{
public:
    AnalyserTrackController (AudioDeviceManager& deviceManagerRef, int channelToAnalyse, String nameOfInputChannel,
                             String ip, String secondaryIP, String bundle)
        : audioDataCollectorHarm (channelToAnalyse), audioDataCollectorSpec (channelToAnalyse),
          audioAnalyserHarm (audioDataCollectorHarm, features, 2048), audioAnalyserSpec (audioDataCollectorSpec, features, 2048),
          oscFeatureSender (features, std::move (ip), bundle), secondaryOSCFeatureSender (features, std::move (secondaryIP), bundle),
          deviceManager (deviceManagerRef), channelName (std::move (nameOfInputChannel))
    {
        enabled = channelToAnalyse >= 0;                                          // :27
        if (enabled)
        {
            features.bind (deviceManager.getEngine(), channelToAnalyse);
            audioDataCollectorSpec.attach (deviceManager);                        // feeds the ring
            audioDataCollectorHarm.attach (deviceManager);                        // same samples: second collector of the reference (:35)
            audioDataCollectorHarm.setNotifyAnalysisThreadCallback ([this]() { audioAnalyserHarm.notify(); });
            deviceManager.addAudioCallback (&audioDataCollectorHarm);
            audioDataCollectorSpec.setNotifyAnalysisThreadCallback ([this]() { audioAnalyserSpec.notify(); });
            deviceManager.addAudioCallback (&audioDataCollectorSpec);
            deviceManager.addAfterAnalysisHook (this, [this]() { audioAnalyserSpec.dispatchOnsetCallback(); });      // RealTimeAnalyser.h:228-229
        }
    }

    ~AnalyserTrackController()
    {
        if (enabled)
        {
            deviceManager.removeAudioCallback (&audioDataCollectorHarm);
            deviceManager.removeAudioCallback (&audioDataCollectorSpec);
            deviceManager.removeAfterAnalysisHooks (this);
        }
        stopAnalysis();
        audioDataCollectorHarm.setNotifyAnalysisThreadCallback (nullptr);
        audioDataCollectorSpec.setNotifyAnalysisThreadCallback (nullptr);
    }
    AnalyserTrackController (const AnalyserTrackController&) = delete;
    AnalyserTrackController& operator= (const AnalyserTrackController&) = delete;

    void clearAnalysisBuffers() { audioDataCollectorHarm.clearBuffer(); audioDataCollectorSpec.clearBuffer(); }    // :167-171
    float getAudioFeature (AudioFeatures::eAudioFeature f) const { return features.getValue (f); }                 // :173

    void prepareToPlay (int samplesPerBlockExpected, double sampleRate)           // :175-188
    {
        stopAnalysis();
        audioAnalyserHarm.sampleRateChanged (sampleRate);
        audioAnalyserSpec.sampleRateChanged (sampleRate);
        audioAnalyserHarm.startThread (4);
        audioAnalyserSpec.startThread (4);
        audioDataCollectorHarm.setExpectedSamplesPerBlock (samplesPerBlockExpected);
        audioDataCollectorSpec.setExpectedSamplesPerBlock (samplesPerBlockExpected);
    }
    void stopAnalysis() { audioAnalyserHarm.stopThread (100); audioAnalyserSpec.stopThread (100); }                // :190-194

    // the setters AnalyserTrack's GUI callbacks reach (:126-134)
    void setGain (float g)                       { audioDataCollectorHarm.setGain (g); audioDataCollectorSpec.setGain (g); }
    void setOnsetDetectionSensitivity (float s)  { audioAnalyserSpec.setOnsetDetectionSensitivity (s); }
    void setOnsetWindowLength (int n)            { audioAnalyserSpec.setOnsetWindowLength (n); }
    void setOnsetDetectionType (OnsetDetector::eOnsetDetectionType t) { audioAnalyserSpec.setOnsetDetectionType (t); }
    void setOnsetDetectedCallback (std::function<void()> f) { audioAnalyserSpec.setOnsetDetectedCallback (std::move (f)); }

    String getChannelName() const noexcept { return channelName; }
    bool isEnabled() const noexcept { return enabled; }
    AudioFeatures& getFeatures() { return features; }
    OSCFeatureAnalysisOutput& getOSCSender() { return oscFeatureSender; }
    OSCFeatureAnalysisOutput& getSecondaryOSCSender() { return secondaryOSCFeatureSender; }

private:
    AudioFeatures            features;
    AudioDataCollector       audioDataCollectorHarm;
    AudioDataCollector       audioDataCollectorSpec;
    RealTimeHarmonicAnalyser audioAnalyserHarm;
    RealTimeSpectralAnalyser audioAnalyserSpec;
    OSCFeatureAnalysisOutput oscFeatureSender;
    OSCFeatureAnalysisOutput secondaryOSCFeatureSender;
    AudioDeviceManager&      deviceManager;
    String                   channelName;
    bool                     enabled { true };
};

} // namespace fxb200
