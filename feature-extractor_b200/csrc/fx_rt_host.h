// fx_rt_host.h -- host-side primitives of the real-time path (no CUDA in here: the same code is compiled into
// libfxb200.so and, with -fsanitize=thread, into tests/cpp/ring_tsan.cpp).
//
// They replace the thread plumbing of the reference (all citations relative to /root/reference/Source/):
//   TrackRings   AudioDataCollector's 4096-float circleBuffer with its non-atomic writeIndex / readIndex and the
//                analysisBufferUpdating busy flag (AudioDataCollector.h:24,36-94): one ring per track in one (pinned)
//                allocation, a monotonic sample count per track published with release / acquire, and the consumer's
//                position per track GROUP.  The producer side is wait-free: bounds check, memcpy, one atomic store.
//   WakeWord     Thread::notify() / wait (-1) (AudioDataCollector.h:68-69, RealTimeAnalyser.h:175,232): a futex word; the
//                producer only enters the kernel when the consumer is actually asleep.
//   SeqBlock     AudioFeatures::getValue read unsynchronised from the OSC / GUI timer threads while the analysis threads
//                write (RealTimeAnalyser.h:76-88, OSCFeatureAnalysisOutput.h:91-104): here a seqlock over relaxed atomic
//                words, so a reader never sees a torn feature vector and never blocks the writer.
#pragma once

#include <atomic>
#include <cerrno>
#include <climits>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <memory>

#include <linux/futex.h>
#include <sched.h>
#if defined (__SSE2__) && ! defined (__SANITIZE_THREAD__)
 #include <emmintrin.h>
 #define FX_RING_STREAMING_STORES 1      // ThreadSanitizer does not see intrinsic stores: its build copies with memcpy
#else
 #define FX_RING_STREAMING_STORES 0
#endif
#include <sys/syscall.h>
#include <unistd.h>

namespace fx {

// The ring is written once by the audio thread and read next by the GPU's copy engine: streaming (non-temporal) stores
// skip the read-for-ownership of the destination lines and leave the audio thread's cache to the host application.  They
// are weakly ordered, so ring_fence() goes between the copies of a push and the publishing stores.
inline void ring_copy (float* dst, const float* src, size_t n) noexcept
{
#if FX_RING_STREAMING_STORES
    if (((reinterpret_cast<uintptr_t> (dst) | reinterpret_cast<uintptr_t> (src)) & 15u) == 0 && (n & 3u) == 0)
    {
        for (size_t i = 0; i < n; i += 4)
            _mm_stream_si128 (reinterpret_cast<__m128i*> (dst + i), _mm_load_si128 (reinterpret_cast<const __m128i*> (src + i)));
        return;
    }
#endif
    std::memcpy (dst, src, n * sizeof (float));
}
inline void ring_fence() noexcept
{
#if FX_RING_STREAMING_STORES
    _mm_sfence();
#endif
}

// ---------------------------------------------------------------------------------------------------------
class WakeWord
{
public:
    // producer (audio thread): publish "something changed", wake the consumer if it sleeps.  No lock, no allocation.
    void signal() noexcept
    {
        word.fetch_add (1u, std::memory_order_seq_cst);
        if (sleeping.load (std::memory_order_seq_cst) != 0u)
            syscall (SYS_futex, reinterpret_cast<uint32_t*> (&word), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
    }
    // consumer: take a ticket before looking for work ...
    uint32_t observe() const noexcept { return word.load (std::memory_order_seq_cst); }
    // ... and sleep until the word moves past it (or the timeout elapses).  Returns at once if a signal came in between.
    void wait (uint32_t seen, int timeout_ms) noexcept
    {
        sleeping.store (1u, std::memory_order_seq_cst);
        if (word.load (std::memory_order_seq_cst) == seen)
        {
            timespec ts;
            ts.tv_sec = timeout_ms / 1000;
            ts.tv_nsec = (long) (timeout_ms % 1000) * 1000000L;
            syscall (SYS_futex, reinterpret_cast<uint32_t*> (&word), FUTEX_WAIT_PRIVATE, seen, &ts, nullptr, 0);
        }
        sleeping.store (0u, std::memory_order_seq_cst);
    }

private:
    std::atomic<uint32_t> word { 0u };
    std::atomic<uint32_t> sleeping { 0u };
};

// ---------------------------------------------------------------------------------------------------------
// One ring of ring_len floats per track inside `base` (owned by the caller: cudaHostAlloc in the engine, malloc in the
// TSAN test).  wpos[t] counts the samples ever pushed into track t (written by the one producer of that track),
// rpos[g] the samples the consumer of group g has finished reading (written by that group's consumer).
class TrackRings
{
public:
    void init (float* memory, long n_tracks, long ring_length, long tracks_per_group)
    {
        base = memory; T = n_tracks; L = ring_length; per = tracks_per_group;
        G = (T + per - 1) / per;
        wpos.reset (new std::atomic<long>[(size_t) T]);
        active.reset (new std::atomic<uint32_t>[(size_t) T]);
        clear_upto.reset (new std::atomic<long>[(size_t) T]);
        rpos.reset (new std::atomic<long>[(size_t) G]);
        for (long t = 0; t < T; ++t) { wpos[(size_t) t].store (0); active[(size_t) t].store (1u); clear_upto[(size_t) t].store (-1L); }
        for (long g = 0; g < G; ++g) rpos[(size_t) g].store (0);
    }
    // back to the state after init (no producer and no consumer may be running)
    void reset() noexcept
    {
        for (long t = 0; t < T; ++t) { wpos[(size_t) t].store (0); clear_upto[(size_t) t].store (-1L); }
        for (long g = 0; g < G; ++g) rpos[(size_t) g].store (0);
    }

    long group_of (long track) const noexcept { return track / per; }
    float* row (long track) const noexcept { return base + (size_t) track * (size_t) L; }

    // Producer.  Returns false (and copies NOTHING for any track) when a block would overwrite samples its group's
    // consumer has not read yet.  Inactive tracks are skipped.  crossed (optional, [G]) is set for every group in which
    // some track completed a hop, so that the caller signals each group once.
    bool push (long first_track, long n, const float* const* channels, long n_samples, long hop, unsigned char* crossed) noexcept
    {
        // (calls for more tracks than the decision bitmap holds are split: each part is all-or-nothing on its own)
        for (long at = 0; at < n; at += kMaxPushTracks)
            if (! push_part (first_track + at, n - at < kMaxPushTracks ? n - at : kMaxPushTracks, channels + at, n_samples, hop, crossed)) return false;
        return true;
    }

    static constexpr long kMaxPushTracks = 16384;

    bool push_part (long first_track, long n, const float* const* channels, long n_samples, long hop, unsigned char* crossed) noexcept
    {
        Inside inside (producers_inside);              // deactivate() waits for the producers that may have seen the old flag
        uint64_t take[kMaxPushTracks / 64];            // which tracks this call feeds: each track's flag is read exactly once
        for (long i = 0; i < (n + 63) / 64; ++i) take[i] = 0;
        for (long i = 0; i < n; ++i)
        {
            const long t = first_track + i;
            if (active[(size_t) t].load (std::memory_order_seq_cst) == 0u) continue;
            const long w = wpos[(size_t) t].load (std::memory_order_relaxed);
            const long r = rpos[(size_t) group_of (t)].load (std::memory_order_acquire);
            if (w + n_samples - r > L) return false;
            take[i >> 6] |= 1ull << (i & 63);
        }
        for (long i = 0; i < n; ++i)
        {
            if (! ((take[i >> 6] >> (i & 63)) & 1ull)) continue;
            const long t = first_track + i;
            const long w = wpos[(size_t) t].load (std::memory_order_relaxed);
            float* ring = row (t);
            const long o = w % L;
            const long first = (o + n_samples <= L) ? n_samples : L - o;
            ring_copy (ring + o, channels[i], (size_t) first);
            if (first < n_samples) ring_copy (ring, channels[i] + first, (size_t) (n_samples - first));
        }
        ring_fence();
        for (long i = 0; i < n; ++i)
        {
            if (! ((take[i >> 6] >> (i & 63)) & 1ull)) continue;
            const long t = first_track + i;
            const long w = wpos[(size_t) t].load (std::memory_order_relaxed);
            wpos[(size_t) t].store (w + n_samples, std::memory_order_release);           // publish
            if (crossed != nullptr && (w / hop) != ((w + n_samples) / hop)) crossed[group_of (t)] = 1;
        }
        return true;
    }

    // Consumer of group g: complete hops available on every ACTIVE track of [t0, t0 + n) (0 when no track is active).
    long hops_available (long g, long t0, long n, long hop) const noexcept
    {
        const long r = rpos[(size_t) g].load (std::memory_order_relaxed);
        long avail = -1;
        for (long t = t0; t < t0 + n; ++t)
        {
            if (active[(size_t) t].load (std::memory_order_acquire) == 0u) continue;
            const long a = wpos[(size_t) t].load (std::memory_order_acquire) - r;
            if (avail < 0 || a < avail) avail = a;
        }
        return avail <= 0 ? 0 : avail / hop;
    }
    // ---- track life cycle (callers serialise these against the group's consumer, e.g. with its batch mutex) -------------
    // The track stops gating its group and its ring becomes silence.  Never blocks a producer: it waits for them instead.
    void deactivate (long t) noexcept
    {
        active[(size_t) t].store (0u, std::memory_order_seq_cst);
        while (producers_inside.load (std::memory_order_seq_cst) != 0) sched_yield();
        std::memset (row (t), 0, (size_t) L * sizeof (float));
        clear_upto[(size_t) t].store (-1L, std::memory_order_relaxed);
    }
    // The track's stream (re)starts at its group's read position.
    void activate (long t) noexcept
    {
        if (active[(size_t) t].load (std::memory_order_relaxed) != 0u) return;
        wpos[(size_t) t].store (rpos[(size_t) group_of (t)].load (std::memory_order_relaxed), std::memory_order_relaxed);
        active[(size_t) t].store (1u, std::memory_order_seq_cst);
    }

    // ---- AudioDataCollector::clearBuffer (AudioDataCollector.h:122) ------------------------------------------------------
    // Any thread: everything pushed into track t so far and not consumed yet is to read as silence; positions do not move.
    void request_clear (long t) noexcept
    {
        clear_upto[(size_t) t].store (wpos[(size_t) t].load (std::memory_order_acquire), std::memory_order_release);
    }
    // Consumer of the track's group, before it reads samples [r, r + n) of the track: those samples are published and not
    // released yet, so nobody else touches them.
    void apply_clear (long t, long r, long n) noexcept
    {
        const long c = clear_upto[(size_t) t].load (std::memory_order_acquire);
        if (c < 0) return;
        const long upto = c < r + n ? c : r + n;
        float* ring = row (t);
        for (long a = r; a < upto;)
        {
            const long o = a % L, run = (o + (upto - a) <= L) ? upto - a : L - o;
            std::memset (ring + o, 0, (size_t) run * sizeof (float));
            a += run;
        }
        if (c <= r + n) { long expect = c; clear_upto[(size_t) t].compare_exchange_strong (expect, -1L); }     // a newer request stays
    }

    long read_pos (long g) const noexcept { return rpos[(size_t) g].load (std::memory_order_relaxed); }
    void consumed (long g, long n_samples) noexcept
    {
        rpos[(size_t) g].store (rpos[(size_t) g].load (std::memory_order_relaxed) + n_samples, std::memory_order_release);
    }

    float* base = nullptr;
    long T = 0, L = 0, per = 1, G = 0;
    std::unique_ptr<std::atomic<long>[]> wpos;
    std::unique_ptr<std::atomic<uint32_t>[]> active;
    std::unique_ptr<std::atomic<long>[]> rpos;
    std::unique_ptr<std::atomic<long>[]> clear_upto;     // per track: samples pushed before the last request_clear (-1: none pending)
    std::atomic<int> producers_inside { 0 };

private:
    struct Inside
    {
        explicit Inside (std::atomic<int>& c) noexcept : count (c) { count.fetch_add (1, std::memory_order_seq_cst); }
        ~Inside() { count.fetch_sub (1, std::memory_order_seq_cst); }
        std::atomic<int>& count;
    };
};

// ---------------------------------------------------------------------------------------------------------
// A block of 32-bit words with one writer and any number of readers: a seqlock without stand-alone fences (every
// ordering hangs on an atomic access, which is also what ThreadSanitizer models).
//   writer  seq = odd (relaxed); words with RELEASE stores (the odd store cannot sink below them); seq = even (release)
//   reader  seq (acquire); words with ACQUIRE loads (the re-check cannot rise above them); seq again (relaxed)
// A reader that saw any word of write k has synchronised with that store, so its re-check sees write k's odd value or
// later and retries.  On x86 all of these are plain moves.
class SeqBlock
{
public:
    void init (size_t n_words) { words.reset (new std::atomic<uint32_t>[n_words]); n = n_words; for (size_t i = 0; i < n; ++i) words[i].store (0u); }
    size_t size() const noexcept { return n; }

    // writer (one thread at a time)
    void write_begin() noexcept { seq.store (seq.load (std::memory_order_relaxed) + 1u, std::memory_order_relaxed); }
    void put (size_t i, uint32_t v) noexcept { words[i].store (v, std::memory_order_release); }
    void write_end() noexcept { seq.store (seq.load (std::memory_order_relaxed) + 1u, std::memory_order_release); }

    // reader: copies words [first, first + count) consistently (retries while a write is in flight)
    void read (size_t first, size_t count, uint32_t* out) const noexcept
    {
        for (;;)
        {
            const uint32_t s1 = seq.load (std::memory_order_acquire);
            if (s1 & 1u) continue;
            for (size_t i = 0; i < count; ++i) out[i] = words[first + i].load (std::memory_order_acquire);
            if (seq.load (std::memory_order_relaxed) == s1) return;
        }
    }

private:
    std::atomic<uint32_t> seq { 0u };
    std::unique_ptr<std::atomic<uint32_t>[]> words;
    size_t n = 0;
};

} // namespace fx
