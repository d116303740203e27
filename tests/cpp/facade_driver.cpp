// facade_driver.cpp -- test driver for the C++ host facade (feature-extractor_b200/host/FeatureExtractorB200.h).
// Written the way an application uses the reference classes: one AnalyserTrackController per input channel on a
// device manager (MainComponent.cpp:137-171), audio arriving in device blocks, features read back per track.
// Usage: facade_driver <audio.f32> <n_tracks> <n_samples> <block> <sample_rate> <out.f32> [osc_port]
// Reads [n_tracks][n_samples] fp32, feeds it in `block`-sample device blocks from this ("audio") thread -- processBlock only
// copies -- and after every block that completed a hop waits until the engine's workers have published it, then appends
// [hop index, 12 features] x n_tracks to out.f32.  Onset callbacks arrive on the worker threads.
#include "../../feature-extractor_b200/host/FeatureExtractorB200.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <memory>

using namespace fxb200;

int main (int argc, char** argv)
{
    if (argc < 7) { fprintf (stderr, "usage\n"); return 2; }
    const int T = atoi (argv[2]);
    const long S = atol (argv[3]);
    const int block = atoi (argv[4]);
    const double sr = atof (argv[5]);
    std::vector<float> audio ((size_t) T * (size_t) S);
    FILE* f = fopen (argv[1], "rb");
    if (! f || fread (audio.data(), sizeof (float), audio.size(), f) != audio.size()) { fprintf (stderr, "read failed\n"); return 2; }
    fclose (f);
    FILE* out = fopen (argv[6], "wb");

    OSCFeatureAnalysisOutput::timerAutoStart() = false;          // this driver steps the senders by hand, hop by hop
    AudioDeviceManager deviceManager (T, sr, block, 2048, 0, T > 2 ? 2 : T);      // several track groups
    std::vector<std::unique_ptr<AnalyserTrackController>> tracks;
    std::atomic<int> onsetCallbacks { 0 };
    const String osc = argc > 7 ? String ("127.0.0.1:") + argv[7] : String ("127.0.0.1:9000");
    for (int ch = 0; ch < T; ++ch)
    {
        tracks.emplace_back (new AnalyserTrackController (deviceManager, ch, "Input " + std::to_string (ch), osc, osc,
                                                          "/Audio/A" + std::to_string (ch)));           // MainComponent.cpp:168-171
        tracks.back()->setOnsetDetectedCallback ([&onsetCallbacks]() { ++onsetCallbacks; });
    }
    AudioDeviceManager::AudioDeviceSetup setup;
    deviceManager.getAudioDeviceSetup (setup);
    for (auto& t : tracks) t->prepareToPlay (setup.bufferSize, setup.sampleRate);

    std::vector<const float*> chans ((size_t) T);
    uint64_t lastHop = 0;
    const long hopSize = deviceManager.getHopSize();
    for (long pos = 0; pos + block <= S; pos += block)
    {
        for (int ch = 0; ch < T; ++ch) chans[(size_t) ch] = audio.data() + (size_t) ch * (size_t) S + pos;
        deviceManager.processBlock (chans.data(), T, block);
        float v[FX_NUM_FEATURES];
        uint64_t hop = 0;
        const uint64_t due = (uint64_t) ((pos + block) / hopSize);
        if (due != lastHop)
        {
            for (int ch = 0; ch < T; ++ch)
                if (! deviceManager.waitForHop (ch, due)) { fprintf (stderr, "hop %llu of track %d never arrived\n", (unsigned long long) due, ch); return 4; }
            lastHop = due;
            for (int ch = 0; ch < T; ++ch)
            {
                tracks[(size_t) ch]->getFeatures().snapshot (v, &hop);
                const float h = (float) hop;
                fwrite (&h, sizeof (float), 1, out);
                for (int k = 0; k < FX_NUM_FEATURES; ++k)
                {
                    const float one = tracks[(size_t) ch]->getAudioFeature ((AudioFeatures::eAudioFeature) k);
                    if (! (one == v[k] || (one != one && v[k] != v[k]))) { fprintf (stderr, "getAudioFeature != snapshot\n"); return 3; }
                }
                fwrite (v, sizeof (float), FX_NUM_FEATURES, out);
                if (argc > 7) tracks[(size_t) ch]->getOSCSender().timerCallback();
            }
        }
    }
    fclose (out);
    // MainComponent destroys and re-creates controllers as channels toggle (MainComponent.cpp:137-171): the new controller of
    // channel 0 starts from empty histories while the other channels carry on
    usleep (100000);                                             // the callbacks of the last hop run right after it was published
    const int onsetsOfTheRun = onsetCallbacks.load();
    uint64_t recreatedHops = 0, othersHops = 0;
    if (T > 1)
    {
        tracks[0].reset();
        tracks[0].reset (new AnalyserTrackController (deviceManager, 0, "Input 0 again", osc, osc, "/Audio/A0"));
        tracks[0]->prepareToPlay (setup.bufferSize, setup.sampleRate);
        const long extraHops = 3;
        for (long pos = 0; pos + block <= extraHops * hopSize; pos += block)
        {
            for (int ch = 0; ch < T; ++ch) chans[(size_t) ch] = audio.data() + (size_t) ch * (size_t) S + pos;
            deviceManager.processBlock (chans.data(), T, block);
        }
        deviceManager.waitForHop (0, (uint64_t) extraHops);
        deviceManager.waitForHop (1, lastHop + (uint64_t) extraHops);
        float v[FX_NUM_FEATURES];
        tracks[0]->getFeatures().snapshot (v, &recreatedHops);
        tracks[1]->getFeatures().snapshot (v, &othersHops);
    }
    const std::vector<uint8_t> msg = tracks[0]->getOSCSender().encode();
    for (auto& t : tracks) t->stopAnalysis();
    printf ("hops %llu onset_callbacks %d osc_bytes %zu push_errors %ld recreated_hops %llu others_hops %llu\n", (unsigned long long) lastHop,
            onsetsOfTheRun, msg.size(), deviceManager.getPushErrorCount(), (unsigned long long) recreatedHops, (unsigned long long) othersHops);
    return 0;
}
