// rt_latency.cpp -- BASELINE.json configs[3]: real-time mode, 512 live tracks, 256-sample callback blocks at 48 kHz
// streamed through the pinned ring; p50 / p99 / max block-to-features latency.
//
// The threads are the reference's (AudioDataCollector.h:36-70, RealTimeAnalyser.h:97-127): THIS thread is the audio
// thread -- paced at the block period it only calls fx_push_block -- and the engine's group workers analyse and publish.
//   audio_thread_us        time spent inside fx_push_block per block (what the audio callback costs)
//   block_to_features_us   from the moment the audio thread hands over the block that completes a hop until the features
//                          of that hop are published for that track group (host wall clock, taken in the features callback
//                          on the worker thread)
// One line of JSON on stdout.
// Build: g++ -O2 -std=c++17 tools/rt_latency.cpp -Iinclude -Lfeature-extractor_b200/lib -lfxb200 -lpthread -o rt_latency
// Usage: rt_latency [tracks=512] [block=256] [seconds=60] [paced=1] [tracks_per_group=128] [window=2048]
#include "fx_engine.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

using clk = std::chrono::steady_clock;

struct Shared
{
    int tracks = 0, per_group = 0, n_groups = 0;
    long max_hops = 0;
    std::vector<std::atomic<int64_t>> push_ns;        // [hop]: when the completing block was handed over (ns since start)
    std::vector<std::vector<double>> lat_us;          // [group][...]
    std::vector<std::atomic<long>> hops_seen;         // [group]
    std::vector<long> multi_hop_batches;              // [group]
    clk::time_point t_start;
};

static void on_features (void* user, int first_track, int /*n_tracks*/, uint64_t frame_index, int n_new)
{
    Shared& s = *static_cast<Shared*> (user);
    const int64_t now = std::chrono::duration_cast<std::chrono::nanoseconds> (clk::now() - s.t_start).count();
    const int g = first_track / s.per_group;
    if (n_new > 1) ++s.multi_hop_batches[(size_t) g];
    for (uint64_t h = frame_index - (uint64_t) n_new + 1; h <= frame_index; ++h)        // 1-based hop numbers
    {
        if ((long) h > s.max_hops) continue;
        const int64_t t_push = s.push_ns[(size_t) (h - 1)].load (std::memory_order_acquire);
        if (t_push >= 0) s.lat_us[(size_t) g].push_back ((double) (now - t_push) * 1e-3);
    }
    s.hops_seen[(size_t) g].store ((long) frame_index, std::memory_order_release);
}

static double pct (std::vector<double>& v, double p) { return v.empty() ? 0.0 : v[std::min (v.size() - 1, (size_t) (p * (double) v.size()))]; }

int main (int argc, char** argv)
{
    const int tracks = argc > 1 ? atoi (argv[1]) : 512;
    const int block = argc > 2 ? atoi (argv[2]) : 256;
    const double seconds = argc > 3 ? atof (argv[3]) : 60.0;
    const int paced = argc > 4 ? atoi (argv[4]) : 1;
    const int per_group = argc > 5 ? atoi (argv[5]) : 128;
    const int window = argc > 6 ? atoi (argv[6]) : 2048;
    const double sr = 48000.0;
    const int hop = window / 2;

    fx_config cfg;
    fx_default_config (&cfg);
    cfg.n_tracks = tracks; cfg.window = window; cfg.hop = hop; cfg.sample_rate = sr;
    cfg.tracks_per_group = per_group; cfg.ring_hops = 8;
    fx_engine* e = nullptr;
    if (fx_engine_create (&cfg, &e) != FX_OK) { fprintf (stderr, "create: %s\n", fx_last_error (nullptr)); return 1; }

    const long n_blocks = (long) (seconds * sr / block);
    Shared sh;
    sh.tracks = tracks; sh.per_group = per_group; sh.n_groups = (tracks + per_group - 1) / per_group;
    sh.max_hops = n_blocks * block / hop;
    sh.push_ns = std::vector<std::atomic<int64_t>> ((size_t) sh.max_hops + 1);
    for (auto& p : sh.push_ns) p.store (-1);
    sh.lat_us.resize ((size_t) sh.n_groups);
    for (auto& v : sh.lat_us) v.reserve ((size_t) sh.max_hops + 16);
    sh.hops_seen = std::vector<std::atomic<long>> ((size_t) sh.n_groups);
    for (auto& h : sh.hops_seen) h.store (0);
    sh.multi_hop_batches.assign ((size_t) sh.n_groups, 0);

    // synthetic audio per track, prepared up front (the audio thread of a real host receives its blocks from the device;
    // generating them here would be charged to the callback)
    // (16 blocks per track, 8 MB in all: like a device's block buffers they stay cache resident, the content does not matter to the timing)
    const long loop_blocks = std::min<long> (n_blocks, 16);
    const long loop_len = loop_blocks * block;
    std::vector<float> audio ((size_t) tracks * (size_t) loop_len);
    unsigned rng = 12345u;
    for (int t = 0; t < tracks; ++t)
    {
        const double f = 110.0 * std::pow (2.0, (t % 48) / 12.0);
        float* dst = audio.data() + (size_t) t * (size_t) loop_len;
        for (long i = 0; i < loop_len; ++i)
        {
            rng = rng * 1664525u + 1013904223u;
            dst[i] = (float) (0.5 * std::sin (2.0 * M_PI * f * (double) i / sr) + 0.05 * ((double) (rng >> 8) / 8388608.0 - 1.0));
        }
    }
    std::vector<const float*> chans ((size_t) tracks);

    fx_set_features_callback (e, on_features, &sh);
    // warm-up outside the measurement: first launches, clocks
    if (fx_rt_start (e) != FX_OK) { fprintf (stderr, "rt_start: %s\n", fx_last_error (e)); return 1; }
    sh.t_start = clk::now();
    {
        const long warm_blocks = 8L * hop / block;
        for (long b = 0; b < warm_blocks; ++b)
        {
            for (int t = 0; t < tracks; ++t) chans[(size_t) t] = audio.data() + (size_t) t * (size_t) loop_len + (b % loop_blocks) * block;
            while (fx_push_block (e, 0, tracks, chans.data(), block) == FX_ERR_OVERRUN) std::this_thread::sleep_for (std::chrono::microseconds (200));
            std::this_thread::sleep_for (std::chrono::microseconds ((long) (1e6 * block / sr)));
        }
        std::this_thread::sleep_for (std::chrono::milliseconds (50));
    }
    fx_rt_stop (e);
    fx_reset (e);
    fx_rt_stats st0; fx_rt_get_stats (e, &st0, 1);
    for (auto& v : sh.lat_us) v.clear();
    std::fill (sh.multi_hop_batches.begin(), sh.multi_hop_batches.end(), 0L);
    if (fx_rt_start (e) != FX_OK) { fprintf (stderr, "rt_start: %s\n", fx_last_error (e)); return 1; }

    std::vector<double> push_us; push_us.reserve ((size_t) n_blocks);
    long overruns = 0, late_blocks = 0;
    sh.t_start = clk::now();
    const auto t_start = sh.t_start;
    for (long b = 0; b < n_blocks; ++b)
    {
        for (int t = 0; t < tracks; ++t) chans[(size_t) t] = audio.data() + (size_t) t * (size_t) loop_len + (b % loop_blocks) * block;
        if (paced)
        {
            const auto due = t_start + std::chrono::duration_cast<clk::duration> (std::chrono::duration<double> ((double) (b + 1) * block / sr));
            if (clk::now() > due + std::chrono::microseconds ((long) (1e6 * block / sr))) ++late_blocks;      // this thread itself fell a period behind
            std::this_thread::sleep_until (due);
        }
        const auto t0 = clk::now();
        const long hop_done = ((b + 1) * block) / hop, hop_before = (b * block) / hop;
        if (hop_done != hop_before)
            sh.push_ns[(size_t) (hop_done - 1)].store (std::chrono::duration_cast<std::chrono::nanoseconds> (t0 - t_start).count(), std::memory_order_release);
        fx_status s = fx_push_block (e, 0, tracks, chans.data(), block);
        if (! paced) while (s == FX_ERR_OVERRUN) { std::this_thread::yield(); s = fx_push_block (e, 0, tracks, chans.data(), block); }
        const auto t1 = clk::now();
        if (s != FX_OK) ++overruns;
        push_us.push_back (std::chrono::duration<double, std::micro> (t1 - t0).count());
    }
    // let the workers drain
    for (int i = 0; i < 2000; ++i)
    {
        bool done = true;
        for (int g = 0; g < sh.n_groups; ++g) done = done && sh.hops_seen[(size_t) g].load (std::memory_order_acquire) >= sh.max_hops;
        if (done || overruns > 0) break;
        std::this_thread::sleep_for (std::chrono::milliseconds (1));
    }
    const double wall = std::chrono::duration<double> (clk::now() - t_start).count();
    if (fx_rt_stop (e) != FX_OK) { fprintf (stderr, "worker failed: %s\n", fx_last_error (e)); return 1; }
    fx_rt_stats st; fx_rt_get_stats (e, &st, 0);

    std::vector<double> all;
    for (auto& v : sh.lat_us) all.insert (all.end(), v.begin(), v.end());
    std::sort (all.begin(), all.end());
    std::sort (push_us.begin(), push_us.end());
    long multi = 0; for (long m : sh.multi_hop_batches) multi += m;
    const double period_us = 1e6 * block / sr;
    long over_period = 0; for (double v : all) if (v > period_us) ++over_period;
    float v12[FX_NUM_FEATURES]; uint64_t idx = 0;
    fx_poll_features (e, 0, v12, &idx);
    printf ("{\"mode\": \"realtime\", \"tracks\": %d, \"block\": %d, \"window\": %d, \"hop\": %d, \"sample_rate\": %.0f, \"paced\": %d, "
            "\"tracks_per_group\": %d, \"groups\": %d, \"seconds_of_audio\": %.1f, \"wall_s\": %.2f, \"hops_per_track\": %ld, \"frames\": %ld, "
            "\"overruns\": %ld, \"late_audio_blocks\": %ld, \"multi_hop_batches\": %ld, \"block_period_us\": %.1f, "
            "\"block_to_features_us\": {\"p50\": %.1f, \"p90\": %.1f, \"p99\": %.1f, \"p999\": %.1f, \"max\": %.1f, \"n\": %zu, \"over_block_period\": %ld}, "
            "\"audio_thread_us\": {\"p50\": %.2f, \"p99\": %.2f, \"max\": %.2f, \"n\": %zu}, "
            "\"engine_batches\": %llu, \"engine_batch_ms\": {\"mean\": %.4f, \"max\": %.4f}, "
            "\"last_frame_index\": %llu, \"rms_feature\": %.5f}\n",
            tracks, block, window, hop, sr, paced, per_group, sh.n_groups, seconds, wall, (long) idx, (long) idx * tracks,
            overruns, late_blocks, multi, period_us,
            pct (all, 0.50), pct (all, 0.90), pct (all, 0.99), pct (all, 0.999), all.empty() ? 0.0 : all.back(), all.size(), over_period,
            pct (push_us, 0.50), pct (push_us, 0.99), push_us.empty() ? 0.0 : push_us.back(), push_us.size(),
            (unsigned long long) st.batches, st.batch_ms_mean, st.batch_ms_max,
            (unsigned long long) idx, v12[FX_RMS]);
    fx_engine_destroy (e);
    return overruns == 0 ? 0 : 3;
}
