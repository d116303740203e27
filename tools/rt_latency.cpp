// rt_latency.cpp -- BASELINE.json configs[3]: real-time mode, 512 live tracks, 256-sample callback blocks at 48 kHz
// streamed through the pinned ring; p50 / p99 / max block-to-features latency.
// "Latency" = host wall clock from the moment the audio thread hands the block that completes a hop to
// fx_push_block until fx_process returns with the smoothed features of that hop visible in host memory.
// Build: g++ -O2 -std=c++17 tools/rt_latency.cpp -Iinclude -Lfeature-extractor_b200/lib -lfxb200 -o rt_latency
// Usage: rt_latency [tracks=512] [block=256] [seconds=20] [paced=1] [tracks_per_group=128] [window=2048]
#include "fx_engine.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

int main (int argc, char** argv)
{
    const int tracks = argc > 1 ? atoi (argv[1]) : 512;
    const int block = argc > 2 ? atoi (argv[2]) : 256;
    const double seconds = argc > 3 ? atof (argv[3]) : 20.0;
    const int paced = argc > 4 ? atoi (argv[4]) : 1;
    const int per_group = argc > 5 ? atoi (argv[5]) : 128;
    const int window = argc > 6 ? atoi (argv[6]) : 2048;
    const double sr = 48000.0;

    fx_config cfg;
    fx_default_config (&cfg);
    cfg.n_tracks = tracks; cfg.window = window; cfg.hop = window / 2; cfg.sample_rate = sr;
    cfg.tracks_per_group = per_group; cfg.ring_hops = 8;
    fx_engine* e = nullptr;
    if (fx_engine_create (&cfg, &e) != FX_OK) { fprintf (stderr, "create: %s\n", fx_last_error (nullptr)); return 1; }

    // one block of synthetic audio per track, refreshed with a rotating phase so the content changes over time
    std::vector<float> audio ((size_t) tracks * block);
    std::vector<const float*> chans ((size_t) tracks);
    for (int t = 0; t < tracks; ++t) chans[(size_t) t] = audio.data() + (size_t) t * block;

    const long n_blocks = (long) (seconds * sr / block);
    std::vector<double> lat; lat.reserve ((size_t) n_blocks);
    using clk = std::chrono::steady_clock;
    const auto t_start = clk::now();
    long hops_total = 0, overruns = 0;
    unsigned rng = 12345u;
    for (long b = 0; b < n_blocks; ++b)
    {
        for (int t = 0; t < tracks; ++t)
        {
            const double f = 110.0 * std::pow (2.0, (t % 48) / 12.0);
            float* dst = audio.data() + (size_t) t * block;
            for (int i = 0; i < block; ++i)
            {
                rng = rng * 1664525u + 1013904223u;
                const double n = (double) (b * block + i);
                dst[i] = (float) (0.5 * std::sin (2.0 * M_PI * f * n / sr) + 0.05 * ((double) (rng >> 8) / 8388608.0 - 1.0));
            }
        }
        if (paced)
        {
            const auto due = t_start + std::chrono::duration_cast<clk::duration> (std::chrono::duration<double> ((b + 1) * block / sr));
            std::this_thread::sleep_until (due);
        }
        const auto t0 = clk::now();
        if (fx_push_block (e, 0, tracks, chans.data(), block) != FX_OK) { ++overruns; continue; }
        long hops = 0;
        if (fx_process (e, &hops) != FX_OK) { fprintf (stderr, "process: %s\n", fx_last_error (e)); return 1; }
        const auto t1 = clk::now();
        if (hops > 0)
        {
            hops_total += hops;
            lat.push_back (std::chrono::duration<double, std::micro> (t1 - t0).count());
        }
    }
    const double wall = std::chrono::duration<double> (clk::now() - t_start).count();
    float v[FX_NUM_FEATURES]; uint64_t idx = 0;
    fx_poll_features (e, 0, v, &idx);
    std::sort (lat.begin(), lat.end());
    auto pct = [&] (double p) { return lat.empty() ? 0.0 : lat[std::min (lat.size() - 1, (size_t) (p * lat.size()))]; };
    printf ("{\"mode\": \"realtime\", \"tracks\": %d, \"block\": %d, \"window\": %d, \"hop\": %d, \"sample_rate\": %.0f, \"paced\": %d, "
            "\"tracks_per_group\": %d, \"seconds_of_audio\": %.1f, \"wall_s\": %.2f, \"hops\": %ld, \"frames\": %ld, \"overruns\": %ld, "
            "\"latency_us\": {\"p50\": %.1f, \"p90\": %.1f, \"p99\": %.1f, \"max\": %.1f, \"n\": %zu}, \"block_period_us\": %.1f, "
            "\"last_frame_index\": %llu, \"rms_feature\": %.5f}\n",
            tracks, block, window, window / 2, sr, paced, per_group, seconds, wall, hops_total, hops_total * tracks, overruns,
            pct (0.50), pct (0.90), pct (0.99), lat.empty() ? 0.0 : lat.back(), lat.size(), 1e6 * block / sr,
            (unsigned long long) idx, v[FX_RMS]);
    fx_engine_destroy (e);
    return 0;
}
