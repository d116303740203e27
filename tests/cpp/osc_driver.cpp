// osc_driver.cpp -- the 60 Hz OSC output of many tracks at once: one AnalyserTrackController per
// channel, each with its two OSCFeatureAnalysisOutput senders (AnalyserTrackController.h:22-23) whose connectToAddress starts
// the 60 Hz timer (OSCFeatureAnalysisOutput.h:133).  All timers share one thread that encodes every sender's datagram in one
// pass over the engine's published block (fx_osc_encode_tracks) and ships them with sendmmsg.
// Usage: osc_driver <n_tracks> <seconds> <portA> <portB> <out.f32>
// Feeds paced 256-sample blocks of synthetic audio for <seconds>, then stops the analysis, lets one more timer tick go out by
// hand and writes the final smoothed vector of every track ([n_tracks][12] fp32, AudioFeatures order) to out.f32.
#include "../../feature-extractor_b200/host/FeatureExtractorB200.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>

using namespace fxb200;

int main (int argc, char** argv)
{
    if (argc < 6) { fprintf (stderr, "usage\n"); return 2; }
    const int T = atoi (argv[1]);
    const double seconds = atof (argv[2]);
    const String ipA = String ("127.0.0.1:") + argv[3], ipB = String ("127.0.0.1:") + argv[4];
    const double sr = 48000.0;
    const int block = 256;

    AudioDeviceManager deviceManager (T, sr, block, 2048);
    std::vector<std::unique_ptr<AnalyserTrackController>> tracks;
    for (int ch = 0; ch < T; ++ch)
        tracks.emplace_back (new AnalyserTrackController (deviceManager, ch, "Input " + std::to_string (ch), ipA, ipB, "/Audio/A" + std::to_string (ch)));
    for (auto& t : tracks) t->prepareToPlay (block, sr);

    // 16 blocks of audio per channel, replayed (the content only has to be non-trivial)
    const int loop = 16;
    std::vector<float> audio ((size_t) T * loop * block);
    for (int ch = 0; ch < T; ++ch)
        for (int i = 0; i < loop * block; ++i)
            audio[(size_t) ch * loop * block + (size_t) i] = (float) (0.4 * std::sin (2.0 * M_PI * (110.0 + 3.0 * ch) * i / sr) + 0.01 * ((i * 2654435761u >> 8) % 1000) / 1000.0);
    std::vector<const float*> chans ((size_t) T);
    const long n_blocks = (long) (seconds * sr / block);
    const auto t0 = std::chrono::steady_clock::now();
    for (long b = 0; b < n_blocks; ++b)
    {
        for (int ch = 0; ch < T; ++ch) chans[(size_t) ch] = audio.data() + (size_t) ch * loop * block + (size_t) (b % loop) * block;
        std::this_thread::sleep_until (t0 + std::chrono::duration_cast<std::chrono::steady_clock::duration> (std::chrono::duration<double> ((double) (b + 1) * block / sr)));
        deviceManager.processBlock (chans.data(), T, block);
    }
    const uint64_t hops = (uint64_t) (n_blocks * block / deviceManager.getHopSize());
    for (int ch = 0; ch < T; ++ch)
        if (! deviceManager.waitForHop (ch, hops)) { fprintf (stderr, "track %d never reached hop %llu\n", ch, (unsigned long long) hops); return 4; }
    for (auto& t : tracks) t->stopAnalysis();
    const unsigned long ticks = OSCTimerThread::instance().ticks(), sent = OSCTimerThread::instance().datagramsSent();
    // the timers are still running: stop them, then send the final state once more by hand
    for (auto& t : tracks) { t->getOSCSender().stopTimer(); t->getSecondaryOSCSender().stopTimer(); }
    std::this_thread::sleep_for (std::chrono::milliseconds (100));
    for (auto& t : tracks) { t->getOSCSender().timerCallback(); t->getSecondaryOSCSender().timerCallback(); }
    FILE* out = fopen (argv[5], "wb");
    for (int ch = 0; ch < T; ++ch)
    {
        float v[FX_NUM_FEATURES];
        tracks[(size_t) ch]->getFeatures().snapshot (v);
        fwrite (v, sizeof (float), FX_NUM_FEATURES, out);
    }
    fclose (out);
    printf ("tracks %d hops %llu ticks %lu datagrams %lu push_errors %ld\n", T, (unsigned long long) hops, ticks, sent, deviceManager.getPushErrorCount());
    return 0;
}
