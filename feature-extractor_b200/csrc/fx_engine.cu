// fx_engine.cu -- host side of libfxb200.so: the C ABI of include/fx_engine.h over the kernels in
// fx_analyse.cu / fx_post.cu.  No CPU fallback: every analysis call runs the CUDA path or fails.
//
// Per-track state kept resident in HBM between calls (all ping-ponged so that chunks of one call may read
// the carry-in while another chunk writes the carry-out):
//   tail   [T][N-H]   fp32  overlap of the framer            (RealTimeAudioDataOverlapper, RealTimeAudioAnalysis.h:194-242)
//   prev   [T][M]     fp32  Re spectrum of the last non-silent frame (previousBinMagnitudes, SpectralCharacteristics.h:203)
//   hist   [T][32][12] fp32 last raw feature rows            (AudioFeatures / ValueHistory, RealTimeAnalyser.h:70-92)
// Real-time ingest: a pinned host ring per track (AudioDataCollector's 4096-float ring, AudioDataCollector.h:24,
// becomes a pinned ring of ring_hops hops), streamed with cudaMemcpy2DAsync on per-track-group streams.
#include "fx_kernels.cuh"
#include "fx_fft.cuh"
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {
thread_local std::string g_create_error;
}

struct fx_group
{
    int          t0 = 0, n = 0;          // track range
    cudaStream_t stream = nullptr;
    float*       d_stage = nullptr;      // [n][stage_hops * H] device staging of newly complete hops
};

struct fx_engine
{
    fx_config cfg{};
    int N = 0, H = 0, M = 0, NB = 0, log2_hop = 0, sm_count = 148, ctas_per_sm = 2;
    cudaStream_t stream = nullptr;
    std::string err;
    std::atomic<uint64_t> launches{0};

    float2 *d_tw1 = nullptr, *d_tw2 = nullptr, *d_tw1f = nullptr;
    double *d_ex_tab = nullptr;
    int    *d_ex_off = nullptr;
    short  *d_her_tab = nullptr;
    short   f0bin_pow2[16] = {};
    double bin_var = 0.0;
    float  iir_c1 = 0.0f, iir_c2 = 0.0f;

    // per-track parameters
    std::vector<float> h_gain, h_mult;
    std::vector<int>   h_type, h_hist;
    std::vector<long>  h_reset;
    float *d_gain = nullptr, *d_mult = nullptr;
    int   *d_type = nullptr, *d_hist = nullptr;
    long  *d_reset = nullptr;
    bool  params_dirty = true;

    // carried state
    float* d_tail[2] = { nullptr, nullptr };
    float* d_prev[2] = { nullptr, nullptr };
    float* d_hrows[2] = { nullptr, nullptr };
    int    flip = 0;
    long   frames_done = 0;             // hops analysed per track since the last reset

    // chunk boundary spectra
    float *d_first_spec = nullptr, *d_last_spec = nullptr;
    int   *d_first_idx = nullptr;
    long  chunk_capacity = 0;           // in (track, chunk) pairs

    // engine-owned result buffers (host API, streaming)
    float *d_raw = nullptr, *d_smooth = nullptr, *d_diag = nullptr;
    long  result_capacity = 0;          // frames per track
    fx::FrameRec* d_rec = nullptr;      // K1 -> K1b hand-over, [T][rec_capacity]
    long  rec_capacity = 0;
    float *d_latest = nullptr, *h_latest = nullptr;      // [T][14]

    // host API pipeline
    float* d_audio_slot[3] = { nullptr, nullptr, nullptr };
    long   audio_slot_floats = 0;
    unsigned char* d_pcm_slot[3] = { nullptr, nullptr, nullptr };     // file ingest: raw PCM rows of one track group
    long   pcm_slot_bytes = 0;
    cudaStream_t pipe_stream[3] = { nullptr, nullptr, nullptr };      // upload, analysis, download
    cudaEvent_t  ev_ready[3] = { nullptr, nullptr, nullptr }, ev_free[3] = { nullptr, nullptr, nullptr };

    // streaming
    float* h_ring = nullptr;            // pinned [T][ring_len]
    long   ring_len = 0;
    std::vector<std::atomic<long>> wpos;   // samples written per track (monotonic)
    long   rpos = 0;                       // samples consumed per track (all tracks advance together)
    std::vector<fx_group> groups;
    long   stage_hops = 0;

    // optional per-kernel timing (bench.py roofline)
    bool   profiling = false;
    struct ProfRec { cudaEvent_t a, b, c; };
    std::vector<ProfRec> prof_pending, prof_free;
};

namespace {

using namespace fx;

bool fail (fx_engine* e, fx_status&, const char* what, cudaError_t ce)
{
    char buf[512];
    snprintf (buf, sizeof (buf), "%s: %s", what, cudaGetErrorString (ce));
    if (e) e->err = buf; else g_create_error = buf;
    return false;
}

#define FX_CUDA(e, call)                                                             \
    do { cudaError_t ce_ = (call); if (ce_ != cudaSuccess) { fx_status st_ = FX_ERR_CUDA; fail ((e), st_, #call, ce_); return FX_ERR_CUDA; } } while (0)

int ilog2 (int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// twiddle tables in the layout fx_fft.cuh expects; evaluated in double, rounded to fp32 (as juce::FFT does)
void build_twiddles (int N, std::vector<float2>& tw1, std::vector<float2>& tw2, std::vector<float2>& tw1f)
{
    const int R1 = N / 256;
    const double pi = 3.14159265358979323846;
    tw1.assign ((size_t) (R1 - 1) * 32, make_float2 (1.0f, 0.0f));
    for (int k1 = 1; k1 < R1; ++k1)
        for (int i = 0; i < 16; ++i)
        {
            const double pa = -2.0 * pi * (double) ((long) 16 * i * k1 % N) / (double) N;     // W_N^(16 mh k1)
            const double pb = -2.0 * pi * (double) (i * k1) / (double) N;                      // W_N^(ml k1)
            tw1[(size_t) (k1 - 1) * 32 + i]      = make_float2 ((float) cos (pa), (float) sin (pa));
            tw1[(size_t) (k1 - 1) * 32 + 16 + i] = make_float2 ((float) cos (pb), (float) sin (pb));
        }
    // the full stage-1 table: W_N^(m k1) = W_N^(16 mh k1) * W_N^(ml k1), m = 16 mh + ml, as the fp32 product of the two
    // factors above with the roundings of the kernel's former in-line product (one rounded product, one FMA per part)
    tw1f.assign ((size_t) (R1 - 1) * 256, make_float2 (1.0f, 0.0f));
    for (int k1 = 1; k1 < R1; ++k1)
        for (int m = 0; m < 256; ++m)
        {
            const float2 wa = tw1[(size_t) (k1 - 1) * 32 + (m >> 4)], wb = tw1[(size_t) (k1 - 1) * 32 + 16 + (m & 15)];
            volatile float pyy = wa.y * wb.y, pxy = wa.x * wb.y;
            tw1f[(size_t) (k1 - 1) * 256 + m] = make_float2 (fmaf (wa.x, wb.x, -pyy), fmaf (wa.y, wb.x, pxy));
        }
    tw2.assign (15 * 16, make_float2 (1.0f, 0.0f));
    for (int k2 = 1; k2 < 16; ++k2)
        for (int n3 = 0; n3 < 16; ++n3)
        {
            const double ph = -2.0 * pi * (double) (n3 * k2) / 256.0;
            tw2[(size_t) (k2 - 1) * 16 + n3] = make_float2 ((float) cos (ph), (float) sin (ph));
        }
}

// f0 and the harmonic bins as functions of the integer lag, in the reference's arithmetic (see AnalyseParams)
void build_lag_tables (int N, double sample_rate, std::vector<short>& her_tab)
{
    const int M = N / 2;
    const double nyquist = sample_rate / 2.0;
    const double frpb = nyquist / (double) M;                       // HarmonicCharacteristics.h:53
    her_tab.assign ((size_t) (N + 1) * FX_HER_TAB_STRIDE, (short) -1);
    auto clamp_short = [] (double v) { return (short) (v > 32767.0 ? 32767.0 : (v < -32768.0 ? -32768.0 : v)); };
    for (int slot = 0; slot <= N; ++slot)
    {
        const double lag = slot == 0 ? -1.0 : (double) slot;
        const double f0 = (nyquist * 2.0) / lag;                    // PitchAnalyser.h:57
        short* row = her_tab.data() + (size_t) slot * FX_HER_TAB_STRIDE;
        const double f0_bin_d = floor (f0 / frpb);                  // :246-249
        row[18] = clamp_short (f0_bin_d);
        const int f0_bin = (int) f0_bin_d;
        for (int l = 0; l < 15; ++l)
        {
            const double fr = f0 * ldexp (1.0, -(l + 1));           // f0 / 2^(l+1), exact scaling (:160-161)
            const double hb = floor (fr / frpb);
            // a sub-octave landing in f0's own bin is skipped (:163-164)
            if (hb >= 0.0 && hb < (double) M && (int) hb != f0_bin) row[l] = (short) hb;
        }
        for (int h = 1; h <= 3; ++h)
        {
            const double fr = f0 * (double) h;                      // :171-172
            const double hb = floor (fr / frpb);
            // harmonics stop at the first bin >= M (:174-175); they ascend, so skipping every bin >= M is the same
            if (hb >= 0.0 && hb < (double) M) row[14 + h] = (short) hb;
        }
    }
}

// Inharmonicity fractions for the (lag, bin) pairs whose edge ratios may be exact integers -- there the reference's own fp64
// rounding decides on which side of the integer a ratio lands, so the value is taken from the reference's arithmetic, evaluated
// here (HarmonicCharacteristics.h:223-236, :251-259), instead of being derived on the GPU.  Layout per lag (step = N / gcd (lag, N)):
//   [ex_off[lag] + 2 q]     bin = q step      (start edge: bin lag is a multiple of N)
//   [ex_off[lag] + 2 q + 1] bin = q step - 1  (end edge: (bin + 1) lag is a multiple of N),   q = 0 .. M / step
// and, for lag = 2^b only, 2 * 13 entries in front of them for the bins below f0's:
//   [ex_off[lag] - 26 + 2 a] bin = 2^a,  [ex_off[lag] - 26 + 2 a + 1] bin = 2^a - 1   (bin lag or (bin + 1) lag divides N)
void build_exact_ratio_table (int N, double sample_rate, std::vector<double>& ex_tab, std::vector<int>& ex_off)
{
    const int M = N / 2;
    const double nyquist = sample_rate / 2.0;
    const double frpb = nyquist / (double) M;
    auto ratio = [] (double f1, double f2)                         // getFrequencyRatio (:251-259)
    {
        if (f1 == f2) return 1.0;
        const double higher = f1 > f2 ? f1 : f2;
        const double lower = higher == f1 ? f2 : f1;
        return higher / lower;
    };
    auto fraction = [&] (int bin, double f0)                       // :223-236
    {
        if (bin < 0 || bin >= M) return 0.0;
        double start = (double) bin * frpb;
        if (start == 0.0) start = frpb * 0.5;
        const double end = (double) (bin + 1) * frpb;
        const double ra = ratio (start, f0), rb = ratio (end, f0);
        if (floor (ra) != floor (rb)) return 0.0;
        const double r = ra < rb ? ra : rb;
        return r - floor (r);
    };
    ex_tab.clear();
    ex_off.assign ((size_t) N + 1, 0);
    for (int lag = 1; lag <= N; ++lag)
    {
        const double f0 = (nyquist * 2.0) / (double) lag;           // PitchAnalyser.h:57
        const int g = lag & -lag, step = N / g;
        if ((lag & (lag - 1)) == 0)
            for (int a = 0; a < 13; ++a)
            {
                ex_tab.push_back (fraction (1 << a, f0));
                ex_tab.push_back (fraction ((1 << a) - 1, f0));
            }
        ex_off[(size_t) lag] = (int) ex_tab.size();
        for (int q = 0; q <= M / step; ++q)
        {
            ex_tab.push_back (fraction (q * step, f0));
            ex_tab.push_back (fraction (q * step - 1, f0));
        }
    }
}

fx_status upload_params (fx_engine* e, cudaStream_t s)
{
    if (! e->params_dirty) return FX_OK;
    const size_t T = (size_t) e->cfg.n_tracks;
    FX_CUDA (e, cudaMemcpyAsync (e->d_gain,  e->h_gain.data(),  T * sizeof (float), cudaMemcpyHostToDevice, s));
    FX_CUDA (e, cudaMemcpyAsync (e->d_mult,  e->h_mult.data(),  T * sizeof (float), cudaMemcpyHostToDevice, s));
    FX_CUDA (e, cudaMemcpyAsync (e->d_type,  e->h_type.data(),  T * sizeof (int),   cudaMemcpyHostToDevice, s));
    FX_CUDA (e, cudaMemcpyAsync (e->d_hist,  e->h_hist.data(),  T * sizeof (int),   cudaMemcpyHostToDevice, s));
    FX_CUDA (e, cudaMemcpyAsync (e->d_reset, e->h_reset.data(), T * sizeof (long),  cudaMemcpyHostToDevice, s));
    // the vectors must not change until the copies have been issued from pageable memory: cudaMemcpyAsync from
    // pageable memory stages synchronously, so they are safe to modify once it returns
    e->params_dirty = false;
    return FX_OK;
}

fx_status ensure_chunks (fx_engine* e, long pairs)
{
    if (pairs <= e->chunk_capacity) return FX_OK;
    // grow on the engine stream's timeline: earlier work that used the old buffers must have finished
    FX_CUDA (e, cudaDeviceSynchronize());
    cudaFree (e->d_first_spec); cudaFree (e->d_last_spec); cudaFree (e->d_first_idx);
    e->d_first_spec = e->d_last_spec = nullptr; e->d_first_idx = nullptr; e->chunk_capacity = 0;
    FX_CUDA (e, cudaMalloc (&e->d_first_spec, (size_t) pairs * e->M * sizeof (float)));
    FX_CUDA (e, cudaMalloc (&e->d_last_spec,  (size_t) pairs * e->M * sizeof (float)));
    FX_CUDA (e, cudaMalloc (&e->d_first_idx,  (size_t) pairs * sizeof (int)));
    e->chunk_capacity = pairs;
    return FX_OK;
}

fx_status ensure_results (fx_engine* e, long frames)
{
    if (frames <= e->result_capacity) return FX_OK;
    FX_CUDA (e, cudaDeviceSynchronize());
    cudaFree (e->d_raw); cudaFree (e->d_smooth); cudaFree (e->d_diag);
    e->d_raw = e->d_smooth = e->d_diag = nullptr; e->result_capacity = 0;
    const size_t rows = (size_t) e->cfg.n_tracks * (size_t) frames;
    FX_CUDA (e, cudaMalloc (&e->d_raw,    rows * FX_NUM_FEATURES * sizeof (float)));
    FX_CUDA (e, cudaMalloc (&e->d_smooth, rows * FX_NUM_FEATURES * sizeof (float)));
    FX_CUDA (e, cudaMalloc (&e->d_diag,   rows * FX_NUM_DIAG * sizeof (float)));
    e->result_capacity = frames;
    return FX_OK;
}

fx_status ensure_records (fx_engine* e, long frames)
{
    if (frames <= e->rec_capacity) return FX_OK;
    FX_CUDA (e, cudaDeviceSynchronize());
    cudaFree (e->d_rec);
    e->d_rec = nullptr; e->rec_capacity = 0;
    FX_CUDA (e, cudaMalloc (&e->d_rec, (size_t) e->cfg.n_tracks * (size_t) frames * sizeof (fx::FrameRec)));
    e->rec_capacity = frames;
    return FX_OK;
}

// How many chunks per track.  All CTAs carry equal work, so the launch takes ceil (CTAs / resident slots) rounds and an almost
// empty last round costs a whole chunk's time: 4096 tracks on 444 slots are 9.2 rounds as whole tracks but 36.9 as quarters.
// Against that, every chunk start refills the whole window and repeats the prologue (about two frames' worth).  Among the
// chunk counts that give at least a few rounds over the machine take the one with the best fill x (1 - start-up share); never
// chunks shorter than 8 frames.
int choose_chunks (const fx_engine* e, long n_tracks, long frames)
{
    if (frames <= 8 || n_tracks <= 0) return 1;
    if (const char* f = getenv ("FXB200_CHUNKS")) { const long k = atol (f); if (k >= 1 && k <= (frames + 7) / 8) return (int) k; }   // tuning aid
    const long slots = (long) e->sm_count * e->ctas_per_sm;
    const long target = slots * 6;
    const long max_c = (frames + 7) / 8;
    long c = (target + n_tracks - 1) / n_tracks;
    if (c > max_c) c = max_c;
    if (c < 1) c = 1;
    const double startup_frames = 2.0;
    long best_c = c; double best_score = 0.0;
    for (long k = c; k <= c + 8 && k <= max_c; ++k)
    {
        const long ctas = n_tracks * k, rounds = (ctas + slots - 1) / slots;
        const double fill = (double) ctas / (double) (rounds * slots);
        const double fpc = (double) ((frames + k - 1) / k);
        const double score = fill * fpc / (fpc + startup_frames);
        if (score > best_score + 1e-9) { best_score = score; best_c = k; }
    }
    return (int) best_c;
}

// K1 -> K2 -> K3 for tracks [t0, t0 + nt) on `s`.  Output pointers are already offset to track t0's rows.
// Does NOT advance the carried-state bookkeeping (the caller flips once per call after all groups).
fx_status run_range (fx_engine* e, int t0, int nt, int n_chunks, const float* d_audio, long track_stride, long frames,
                     float* d_raw, float* d_smooth, float* d_diag, float* d_latest, cudaStream_t s)
{
    const int in = e->flip, out = e->flip ^ 1;
    const int fpc = (int) ((frames + n_chunks - 1) / n_chunks);

    AnalyseParams a{};
    a.audio = d_audio; a.track_stride = track_stride;
    a.tail_in = e->d_tail[in] + (size_t) t0 * (e->N - e->H);
    a.tail_out = e->d_tail[out] + (size_t) t0 * (e->N - e->H);
    a.first_hop = e->frames_done;
    a.n_frames = (int) frames; a.frames_per_chunk = fpc; a.n_chunks = n_chunks;
    a.hop = e->H; a.log2_hop = e->log2_hop;
    a.use_bulk = (((uintptr_t) d_audio & 15u) == 0 && (track_stride % 4) == 0 && (e->H % 4) == 0) ? 1 : 0;
    a.gain = e->d_gain + t0;
    a.sample_rate = e->cfg.sample_rate; a.bin_var = e->bin_var; a.iir_c1 = e->iir_c1; a.iir_c2 = e->iir_c2;
    a.tw1 = e->d_tw1; a.tw2 = e->d_tw2; a.tw1f = e->d_tw1f;
    a.her_tab = e->d_her_tab; a.ex_tab = e->d_ex_tab; a.ex_off = e->d_ex_off;
    for (int k = 0; k < 16; ++k) a.f0bin_pow2[k] = e->f0bin_pow2[k];
    a.rec = e->d_rec + (size_t) t0 * (size_t) frames;
    // chunk buffers are indexed by (local track, chunk); each range uses its own slice keyed by t0
    const size_t coff = (size_t) t0 * (size_t) n_chunks;
    a.first_spec = e->d_first_spec + coff * e->M;
    a.last_spec  = e->d_last_spec + coff * e->M;
    a.first_idx  = e->d_first_idx + coff;
    fx_engine::ProfRec pr{};
    if (e->profiling)
    {
        if (! e->prof_free.empty()) { pr = e->prof_free.back(); e->prof_free.pop_back(); }
        else { FX_CUDA (e, cudaEventCreate (&pr.a)); FX_CUDA (e, cudaEventCreate (&pr.b)); FX_CUDA (e, cudaEventCreate (&pr.c)); }
        FX_CUDA (e, cudaEventRecord (pr.a, s));
    }
    FX_CUDA (e, launch_analyse (e->N, nt, a, s));
    if (e->profiling) FX_CUDA (e, cudaEventRecord (pr.b, s));

    FinalizeParams fz{};
    fz.rec = a.rec; fz.n_rows = (long) nt * frames; fz.window = e->N; fz.sample_rate = e->cfg.sample_rate; fz.bin_var = e->bin_var;
    fz.raw = d_raw; fz.diag = d_diag;
    FX_CUDA (e, launch_finalize (fz, s));

    FluxFixParams fp{};
    fp.n_frames = (int) frames; fp.n_chunks = n_chunks; fp.m = e->M;
    fp.first_spec = a.first_spec; fp.last_spec = a.last_spec; fp.first_idx = a.first_idx;
    fp.prev_in = e->d_prev[in] + (size_t) t0 * e->M;
    fp.prev_out = e->d_prev[out] + (size_t) t0 * e->M;
    fp.raw = d_raw;
    FX_CUDA (e, launch_flux_fix (nt, fp, s));

    SmoothParams sp{};
    sp.n_frames = (int) frames; sp.frames_before = e->frames_done; sp.rms_pushes = e->cfg.rms_pushes_per_frame;
    sp.raw = d_raw; sp.smooth = d_smooth; sp.diag = d_diag;
    sp.hist_in = e->d_hrows[in] + (size_t) t0 * kHistRows * FX_NUM_FEATURES;
    sp.hist_out = e->d_hrows[out] + (size_t) t0 * kHistRows * FX_NUM_FEATURES;
    sp.onset_type = e->d_type + t0; sp.onset_hist = e->d_hist + t0; sp.onset_mult = e->d_mult + t0; sp.onset_reset = e->d_reset + t0;
    sp.latest = d_latest;
    FX_CUDA (e, launch_smooth (nt, sp, s));
    if (e->profiling) { FX_CUDA (e, cudaEventRecord (pr.c, s)); e->prof_pending.push_back (pr); }
    e->launches += 5;
    return FX_OK;
}

void free_engine (fx_engine* e)
{
    if (! e) return;
    cudaSetDevice (e->cfg.device);
    cudaDeviceSynchronize();
    cudaFree (e->d_tw1); cudaFree (e->d_tw2); cudaFree (e->d_tw1f); cudaFree (e->d_her_tab); cudaFree (e->d_ex_tab); cudaFree (e->d_ex_off);
    cudaFree (e->d_gain); cudaFree (e->d_mult); cudaFree (e->d_type); cudaFree (e->d_hist); cudaFree (e->d_reset);
    for (int i = 0; i < 2; ++i) { cudaFree (e->d_tail[i]); cudaFree (e->d_prev[i]); cudaFree (e->d_hrows[i]); }
    cudaFree (e->d_first_spec); cudaFree (e->d_last_spec); cudaFree (e->d_first_idx);
    cudaFree (e->d_raw); cudaFree (e->d_smooth); cudaFree (e->d_diag); cudaFree (e->d_latest); cudaFree (e->d_rec);
    if (e->h_latest) cudaFreeHost (e->h_latest);
    for (int i = 0; i < 3; ++i) cudaFree (e->d_pcm_slot[i]);
    for (int i = 0; i < 3; ++i) { cudaFree (e->d_audio_slot[i]); if (e->pipe_stream[i]) cudaStreamDestroy (e->pipe_stream[i]); }
    for (int i = 0; i < 3; ++i) { if (e->ev_ready[i]) cudaEventDestroy (e->ev_ready[i]); if (e->ev_free[i]) cudaEventDestroy (e->ev_free[i]); }
    if (e->h_ring) cudaFreeHost (e->h_ring);
    for (auto& g : e->groups) { cudaFree (g.d_stage); if (g.stream) cudaStreamDestroy (g.stream); }
    for (auto* v : { &e->prof_pending, &e->prof_free })
        for (auto& pr : *v) { cudaEventDestroy (pr.a); cudaEventDestroy (pr.b); cudaEventDestroy (pr.c); }
    if (e->stream) cudaStreamDestroy (e->stream);
    delete e;
}

fx_status clear_state (fx_engine* e, cudaStream_t s)
{
    const size_t T = (size_t) e->cfg.n_tracks;
    for (int i = 0; i < 2; ++i)
    {
        FX_CUDA (e, cudaMemsetAsync (e->d_tail[i], 0, T * (size_t) (e->N - e->H) * sizeof (float), s));
        FX_CUDA (e, cudaMemsetAsync (e->d_prev[i], 0, T * (size_t) e->M * sizeof (float), s));
        FX_CUDA (e, cudaMemsetAsync (e->d_hrows[i], 0, T * kHistRows * FX_NUM_FEATURES * sizeof (float), s));
    }
    FX_CUDA (e, cudaMemsetAsync (e->d_latest, 0, T * (FX_NUM_FEATURES + 2) * sizeof (float), s));
    e->flip = 0;
    e->frames_done = 0;
    return FX_OK;
}

} // namespace

extern "C" {

void fx_default_config (fx_config* c)
{
    c->n_tracks = 1;
    c->window = 2048;                 // AnalyserTrackController.h:20-21
    c->hop = 1024;                    // RealTimeAudioAnalysis.h:207
    c->sample_rate = 48000.0;         // RealTimeAnalyser.h:100
    c->device = 0;
    c->rms_pushes_per_frame = 2;      // RealTimeAnalyser.h:150,209
    c->onset_type = 1;                // SpectralCharacteristics.h:240
    c->onset_hist = 5;                // SpectralCharacteristics.h:238-239
    c->onset_multiplier = 1.7f;       // SpectralCharacteristics.h:311
    c->gain = 1.0f;                   // AudioDataCollector.h:129
    c->max_frames_per_call = 0;       // grow on demand
    c->ring_hops = 16;
    c->tracks_per_group = 0;
}

const char* fx_version (void) { return "fxb200 0.1 (sm_100a)"; }

const char* fx_last_error (const fx_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

fx_status fx_engine_create (const fx_config* cfg, fx_engine** out)
{
    if (! cfg || ! out) { g_create_error = "null argument"; return FX_ERR_INVALID_ARG; }
    *out = nullptr;
    const int N = cfg->window, H = cfg->hop;
    if (N != 1024 && N != 2048 && N != 4096) { g_create_error = "window must be 1024, 2048 or 4096"; return FX_ERR_UNSUPPORTED; }
    if (H < 16 || H > N || (N % H) != 0 || (H & (H - 1)) != 0) { g_create_error = "hop must be a power of two >= 16 dividing the window"; return FX_ERR_UNSUPPORTED; }
    if (cfg->n_tracks < 1) { g_create_error = "n_tracks must be >= 1"; return FX_ERR_INVALID_ARG; }
    if (cfg->onset_hist < 1 || cfg->onset_hist > kMaxOnsetHist) { g_create_error = "onset_hist must be in 1..16"; return FX_ERR_INVALID_ARG; }
    if (cfg->rms_pushes_per_frame != 1 && cfg->rms_pushes_per_frame != 2) { g_create_error = "rms_pushes_per_frame must be 1 or 2"; return FX_ERR_INVALID_ARG; }

    int ndev = 0;
    if (cudaGetDeviceCount (&ndev) != cudaSuccess || ndev == 0) { g_create_error = "no CUDA device (this library has no CPU fallback)"; return FX_ERR_NO_DEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return FX_ERR_INVALID_ARG; }

    fx_engine* e = new fx_engine();
    e->cfg = *cfg; e->N = N; e->H = H; e->M = N / 2; e->NB = N / H; e->log2_hop = ilog2 (H);
    e->ctas_per_sm = (N == 4096) ? 3 : (N == 2048 ? 6 : 12);
    const size_t T = (size_t) cfg->n_tracks;

#define FX_CREATE(call) do { cudaError_t ce_ = (call); if (ce_ != cudaSuccess) { fx_status st_ = FX_ERR_CUDA; fail (nullptr, st_, #call, ce_); free_engine (e); return FX_ERR_CUDA; } } while (0)
    FX_CREATE (cudaSetDevice (cfg->device));
    cudaDeviceProp prop{};
    FX_CREATE (cudaGetDeviceProperties (&prop, cfg->device));
    e->sm_count = prop.multiProcessorCount;
    FX_CREATE (configure_analyse (N));
    FX_CREATE (cudaStreamCreateWithFlags (&e->stream, cudaStreamNonBlocking));

    std::vector<float2> tw1, tw2, tw1f;
    build_twiddles (N, tw1, tw2, tw1f);
    FX_CREATE (cudaMalloc (&e->d_tw1, tw1.size() * sizeof (float2)));
    FX_CREATE (cudaMalloc (&e->d_tw2, tw2.size() * sizeof (float2)));
    FX_CREATE (cudaMemcpy (e->d_tw1, tw1.data(), tw1.size() * sizeof (float2), cudaMemcpyHostToDevice));
    FX_CREATE (cudaMemcpy (e->d_tw2, tw2.data(), tw2.size() * sizeof (float2), cudaMemcpyHostToDevice));
    FX_CREATE (cudaMalloc (&e->d_tw1f, tw1f.size() * sizeof (float2)));
    FX_CREATE (cudaMemcpy (e->d_tw1f, tw1f.data(), tw1f.size() * sizeof (float2), cudaMemcpyHostToDevice));
    {
        std::vector<short> her_tab;
        build_lag_tables (N, cfg->sample_rate, her_tab);
        for (int k = 0; k < 16 && (1 << k) <= N; ++k) e->f0bin_pow2[k] = her_tab[(size_t) (1 << k) * FX_HER_TAB_STRIDE + 18];
        std::vector<double> ex_tab; std::vector<int> ex_off;
        build_exact_ratio_table (N, cfg->sample_rate, ex_tab, ex_off);
        FX_CREATE (cudaMalloc (&e->d_ex_tab, ex_tab.size() * sizeof (double)));
        FX_CREATE (cudaMalloc (&e->d_ex_off, ex_off.size() * sizeof (int)));
        FX_CREATE (cudaMemcpy (e->d_ex_tab, ex_tab.data(), ex_tab.size() * sizeof (double), cudaMemcpyHostToDevice));
        FX_CREATE (cudaMemcpy (e->d_ex_off, ex_off.data(), ex_off.size() * sizeof (int), cudaMemcpyHostToDevice));
        FX_CREATE (cudaMalloc (&e->d_her_tab, her_tab.size() * sizeof (short)));
        FX_CREATE (cudaMemcpy (e->d_her_tab, her_tab.data(), her_tab.size() * sizeof (short), cudaMemcpyHostToDevice));
    }

    // SpectralCharacteristics.h:180-189: binVar accumulated sequentially in double
    {
        double bv = 0.0;
        const int M = e->M;
        for (double i = 0.0; i < M; i++) { const double ni = i / (double) M; bv += (ni - 0.5) * (ni - 0.5); }
        e->bin_var = bv / (double) M;
    }
    // RealTimeAudioAnalysis.h:122,127: (float_Pi / m) and exp (-float_Pi / m), m = 2, in fp32
    {
        const float pi_f = 3.14159265358979323846f, m = 2.0f;
        e->iir_c1 = pi_f / m;
        e->iir_c2 = expf (-pi_f / m);
    }

    e->h_gain.assign (T, cfg->gain); e->h_mult.assign (T, cfg->onset_multiplier);
    e->h_type.assign (T, cfg->onset_type); e->h_hist.assign (T, cfg->onset_hist); e->h_reset.assign (T, 0);
    FX_CREATE (cudaMalloc (&e->d_gain, T * sizeof (float)));
    FX_CREATE (cudaMalloc (&e->d_mult, T * sizeof (float)));
    FX_CREATE (cudaMalloc (&e->d_type, T * sizeof (int)));
    FX_CREATE (cudaMalloc (&e->d_hist, T * sizeof (int)));
    FX_CREATE (cudaMalloc (&e->d_reset, T * sizeof (long)));
    for (int i = 0; i < 2; ++i)
    {
        FX_CREATE (cudaMalloc (&e->d_tail[i], T * (size_t) (N - H + 4) * sizeof (float)));
        FX_CREATE (cudaMalloc (&e->d_prev[i], T * (size_t) e->M * sizeof (float)));
        FX_CREATE (cudaMalloc (&e->d_hrows[i], T * kHistRows * FX_NUM_FEATURES * sizeof (float)));
    }
    FX_CREATE (cudaMalloc (&e->d_latest, T * (FX_NUM_FEATURES + 2) * sizeof (float)));
    FX_CREATE (cudaHostAlloc (&e->h_latest, T * (FX_NUM_FEATURES + 2) * sizeof (float), cudaHostAllocDefault));
    memset (e->h_latest, 0, T * (FX_NUM_FEATURES + 2) * sizeof (float));
    {
        // AudioFeatures::getValue before the first push is 0/0 (RealTimeAnalyser.h:87)
        for (size_t t = 0; t < T; ++t) for (int k = 0; k < FX_NUM_FEATURES; ++k) e->h_latest[t * (FX_NUM_FEATURES + 2) + k] = NAN;
    }
    if (clear_state (e, e->stream) != FX_OK) { g_create_error = e->err; free_engine (e); return FX_ERR_CUDA; }
    if (cfg->max_frames_per_call > 0 && ensure_results (e, cfg->max_frames_per_call) != FX_OK) { g_create_error = e->err; free_engine (e); return FX_ERR_CUDA; }

    // streaming plumbing
    {
        const int rh = cfg->ring_hops >= e->NB + 2 ? cfg->ring_hops : e->NB + 2;
        e->ring_len = (long) rh * H;
        e->stage_hops = rh;
        FX_CREATE (cudaHostAlloc (&e->h_ring, T * (size_t) e->ring_len * sizeof (float), cudaHostAllocDefault));
        memset (e->h_ring, 0, T * (size_t) e->ring_len * sizeof (float));
        e->wpos = std::vector<std::atomic<long>> (T);
        for (auto& w : e->wpos) w.store (0);
        const int per = cfg->tracks_per_group > 0 ? cfg->tracks_per_group : cfg->n_tracks;
        for (int t0 = 0; t0 < cfg->n_tracks; t0 += per)
        {
            fx_group g;
            g.t0 = t0; g.n = (t0 + per <= cfg->n_tracks) ? per : cfg->n_tracks - t0;
            FX_CREATE (cudaStreamCreateWithFlags (&g.stream, cudaStreamNonBlocking));
            FX_CREATE (cudaMalloc (&g.d_stage, (size_t) g.n * (size_t) e->stage_hops * H * sizeof (float)));
            e->groups.push_back (g);
        }
    }
    FX_CREATE (cudaStreamSynchronize (e->stream));
#undef FX_CREATE
    *out = e;
    return FX_OK;
}

fx_status fx_engine_destroy (fx_engine* e)
{
    free_engine (e);
    return FX_OK;
}

fx_status fx_set_gain (fx_engine* e, int track, float gain)
{
    if (! e || track < -1 || track >= e->cfg.n_tracks) return FX_ERR_INVALID_ARG;
    if (track < 0) std::fill (e->h_gain.begin(), e->h_gain.end(), gain); else e->h_gain[(size_t) track] = gain;
    e->params_dirty = true;
    return FX_OK;
}

fx_status fx_set_onset (fx_engine* e, int track, int type, int hist_len, float multiplier)
{
    if (! e || track < -1 || track >= e->cfg.n_tracks) return FX_ERR_INVALID_ARG;
    if (type < 0 || type > 2 || hist_len < 1 || hist_len > kMaxOnsetHist) { e->err = "onset type must be 0..2 and hist_len 1..16"; return FX_ERR_INVALID_ARG; }
    const int a = track < 0 ? 0 : track, b = track < 0 ? e->cfg.n_tracks : track + 1;
    for (int t = a; t < b; ++t)
    {
        e->h_type[(size_t) t] = type;
        e->h_mult[(size_t) t] = multiplier;
        if (e->h_hist[(size_t) t] != hist_len)
        {
            // RealTimeSpectralAnalyser::setOnsetWindowLength -> ValueHistory::setHistoryLength clears (RealTimeAudioAnalysis.h:73-81)
            e->h_hist[(size_t) t] = hist_len;
            e->h_reset[(size_t) t] = e->frames_done;
        }
    }
    e->params_dirty = true;
    return FX_OK;
}

fx_status fx_reset (fx_engine* e)
{
    if (! e) return FX_ERR_INVALID_ARG;
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    FX_CUDA (e, cudaDeviceSynchronize());
    fx_status st = clear_state (e, e->stream);
    if (st != FX_OK) return st;
    std::fill (e->h_reset.begin(), e->h_reset.end(), 0L);
    e->params_dirty = true;
    for (auto& w : e->wpos) w.store (0);
    e->rpos = 0;
    const size_t T = (size_t) e->cfg.n_tracks;
    for (size_t t = 0; t < T; ++t)
    {
        for (int k = 0; k < FX_NUM_FEATURES; ++k) e->h_latest[t * (FX_NUM_FEATURES + 2) + k] = NAN;
        e->h_latest[t * (FX_NUM_FEATURES + 2) + FX_NUM_FEATURES] = 0.0f;
        e->h_latest[t * (FX_NUM_FEATURES + 2) + FX_NUM_FEATURES + 1] = 0.0f;
    }
    FX_CUDA (e, cudaStreamSynchronize (e->stream));
    return FX_OK;
}

fx_status fx_analyse_device (fx_engine* e, const float* d_audio, long track_stride, long n_samples,
                             float* d_raw, float* d_smooth, float* d_diag, void* stream, long* n_frames)
{
    if (! e || ! d_audio || n_samples < 0 || track_stride < n_samples) return FX_ERR_INVALID_ARG;
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    cudaStream_t s = stream ? (cudaStream_t) stream : e->stream;
    const long frames = n_samples / e->H;
    if (n_frames) *n_frames = frames;
    if (frames == 0) return FX_OK;
    if (frames > 0x7fffffffL / (FX_NUM_FEATURES * 4)) { e->err = "too many frames in one call"; return FX_ERR_INVALID_ARG; }
    fx_status st = upload_params (e, s);
    if (st != FX_OK) return st;
    const long T = e->cfg.n_tracks;
    const int n_chunks = choose_chunks (e, T, frames);
    st = ensure_chunks (e, T * n_chunks);
    if (st != FX_OK) return st;
    st = ensure_records (e, frames);
    if (st != FX_OK) return st;
    if (! d_raw)
    {
        st = ensure_results (e, frames);
        if (st != FX_OK) return st;
        d_raw = e->d_raw;
    }
    st = run_range (e, 0, (int) T, n_chunks, d_audio, track_stride, frames, d_raw, d_smooth, d_diag, e->d_latest, s);
    if (st != FX_OK) return st;
    e->flip ^= 1;
    e->frames_done += frames;
    return FX_OK;
}

namespace {

// Shared body of fx_analyse_host / fx_analyse_host_pcm.  format == 0: `src` is fp32, row stride in bytes = 4 * track_stride.
fx_status analyse_host_impl (fx_engine* e, const unsigned char* src, long row_stride_bytes, int format, int n_channels, int channel,
                             long n_samples, float* raw, float* smooth, float* diag, long* n_frames)
{
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    const long frames = n_samples / e->H;
    if (n_frames) *n_frames = frames;
    if (frames == 0) return FX_OK;
    const long T = e->cfg.n_tracks;
    const long used = frames * e->H;                                  // samples per track actually analysed
    const long frame_bytes = format ? (long) n_channels * pcm_bytes_per_sample (format) : (long) sizeof (float);
    const long row_bytes = used * frame_bytes;                        // source bytes per track that cross PCIe
    const bool shared_row = format != 0 && row_stride_bytes == 0;     // every track reads the same interleaved stream

    // Track groups: about thirty per call (the first group's upload and the last group's analysis are the only parts of
    // the pipeline that do not overlap), but never below ~64 MiB of fp32 audio -- small groups would have to cut every
    // track into short chunks to fill the GPU, and each chunk start refills a whole window
    long want_groups = 32;
    if (const char* env = getenv ("FXB200_PIPE_GROUPS")) { const long v = atol (env); if (v > 0) want_groups = v; }
    long per = (T + want_groups - 1) / want_groups;
    const long per_min = (64L << 20) / (used * (long) sizeof (float));
    if (per < per_min) per = per_min;
    if (per < 1) per = 1;
    if (per > T) per = T;
    const long n_groups = (T + per - 1) / per;
    const long slot_floats = per * used;
    if (slot_floats > e->audio_slot_floats)
    {
        FX_CUDA (e, cudaDeviceSynchronize());
        for (int i = 0; i < 3; ++i) { cudaFree (e->d_audio_slot[i]); e->d_audio_slot[i] = nullptr; }
        e->audio_slot_floats = 0;
        for (int i = 0; i < 3; ++i) FX_CUDA (e, cudaMalloc (&e->d_audio_slot[i], (size_t) slot_floats * sizeof (float)));
        e->audio_slot_floats = slot_floats;
    }
    // PCM rows keep a 16-byte aligned pitch on the device so that the vector decode path applies whenever the format allows
    const long pcm_pitch = (row_bytes + 15) & ~15L;
    if (format && per * pcm_pitch > e->pcm_slot_bytes)
    {
        FX_CUDA (e, cudaDeviceSynchronize());
        for (int i = 0; i < 3; ++i) { cudaFree (e->d_pcm_slot[i]); e->d_pcm_slot[i] = nullptr; }
        e->pcm_slot_bytes = 0;
        for (int i = 0; i < 3; ++i) FX_CUDA (e, cudaMalloc (&e->d_pcm_slot[i], (size_t) (per * pcm_pitch)));
        e->pcm_slot_bytes = per * pcm_pitch;
    }
    for (int i = 0; i < 3; ++i)
        if (! e->pipe_stream[i]) FX_CUDA (e, cudaStreamCreateWithFlags (&e->pipe_stream[i], cudaStreamNonBlocking));

    fx_status st = upload_params (e, e->stream);
    if (st != FX_OK) return st;
    FX_CUDA (e, cudaStreamSynchronize (e->stream));
    st = ensure_results (e, frames);
    if (st != FX_OK) return st;
    st = ensure_records (e, frames);
    if (st != FX_OK) return st;
    const int n_chunks = choose_chunks (e, per, frames);
    st = ensure_chunks (e, T * n_chunks);
    if (st != FX_OK) return st;

    // Three streams, three slots: uploads, kernels and result downloads each run in group order on their own stream and
    // meet through events, so that group g + 1's upload overlaps group g's analysis and the analysis kernels run back to
    // back.  (Launching each group on its own stream lets the hardware interleave the CTAs of three analysis kernels: all
    // three then finish together, their slots free together, and every third upload is exposed.)
    cudaStream_t s_up = e->pipe_stream[0], s_run = e->pipe_stream[1], s_down = e->pipe_stream[2];
    for (int i = 0; i < 3; ++i)
    {
        if (! e->ev_ready[i]) FX_CUDA (e, cudaEventCreateWithFlags (&e->ev_ready[i], cudaEventDisableTiming));
        if (! e->ev_free[i])  FX_CUDA (e, cudaEventCreateWithFlags (&e->ev_free[i], cudaEventDisableTiming));
    }
    // FXB200_PIPE_TRACE=1: per-group device timeline of the pipeline on stderr (diagnostics)
    const bool trace = getenv ("FXB200_PIPE_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&] (cudaStream_t st_) { if (trace) { cudaEvent_t ev; cudaEventCreate (&ev); cudaEventRecord (ev, st_); tev.push_back (ev); } };
    for (long g = 0; g < n_groups; ++g)
    {
        const int slot = (int) (g % 3);
        const long t0 = g * per, nt = (t0 + per <= T) ? per : T - t0;
        // upload: the slot's previous occupant (group g - 3) must have been analysed
        if (g >= 3) FX_CUDA (e, cudaStreamWaitEvent (s_up, e->ev_free[slot], 0));
        mark (s_up);
        if (format == 0)
            FX_CUDA (e, cudaMemcpy2DAsync (e->d_audio_slot[slot], (size_t) row_bytes, src + t0 * row_stride_bytes, (size_t) row_stride_bytes,
                                            (size_t) row_bytes, (size_t) nt, cudaMemcpyHostToDevice, s_up));
        else if (shared_row)       // one interleaved stream feeds every track (track t <- channel t % n_channels): upload it once per group
            FX_CUDA (e, cudaMemcpyAsync (e->d_pcm_slot[slot], src, (size_t) row_bytes, cudaMemcpyHostToDevice, s_up));
        else
            FX_CUDA (e, cudaMemcpy2DAsync (e->d_pcm_slot[slot], (size_t) pcm_pitch, src + t0 * row_stride_bytes, (size_t) row_stride_bytes,
                                            (size_t) row_bytes, (size_t) nt, cudaMemcpyHostToDevice, s_up));
        mark (s_up);
        FX_CUDA (e, cudaEventRecord (e->ev_ready[slot], s_up));
        // analysis
        FX_CUDA (e, cudaStreamWaitEvent (s_run, e->ev_ready[slot], 0));
        if (format != 0)
        {
            PcmParams pp{};
            pp.pcm = e->d_pcm_slot[slot]; pp.track_stride_bytes = shared_row ? 0 : pcm_pitch; pp.format = format; pp.n_channels = n_channels;
            pp.channel = channel; pp.first_track = t0;
            pp.n_samples = used; pp.n_tracks = nt; pp.audio = e->d_audio_slot[slot]; pp.audio_stride = used;
            FX_CUDA (e, launch_pcm_decode (pp, s_run));
            e->launches += (uint64_t) pcm_launch_count (pp);
        }
        const size_t roff = (size_t) t0 * (size_t) frames;
        float* dr = e->d_raw + roff * FX_NUM_FEATURES;
        float* ds = e->d_smooth + roff * FX_NUM_FEATURES;
        float* dd = e->d_diag + roff * FX_NUM_DIAG;
        st = run_range (e, (int) t0, (int) nt, n_chunks, e->d_audio_slot[slot], used, frames, dr, smooth ? ds : nullptr, diag ? dd : nullptr,
                        e->d_latest + (size_t) t0 * (FX_NUM_FEATURES + 2), s_run);
        if (st != FX_OK) return st;
        mark (s_run);
        FX_CUDA (e, cudaEventRecord (e->ev_free[slot], s_run));
        // results (engine-owned result buffers are indexed by track: no reuse inside a call)
        FX_CUDA (e, cudaStreamWaitEvent (s_down, e->ev_free[slot], 0));
        if (raw)    FX_CUDA (e, cudaMemcpyAsync (raw + roff * FX_NUM_FEATURES, dr, (size_t) nt * frames * FX_NUM_FEATURES * sizeof (float), cudaMemcpyDeviceToHost, s_down));
        if (smooth) FX_CUDA (e, cudaMemcpyAsync (smooth + roff * FX_NUM_FEATURES, ds, (size_t) nt * frames * FX_NUM_FEATURES * sizeof (float), cudaMemcpyDeviceToHost, s_down));
        if (diag)   FX_CUDA (e, cudaMemcpyAsync (diag + roff * FX_NUM_DIAG, dd, (size_t) nt * frames * FX_NUM_DIAG * sizeof (float), cudaMemcpyDeviceToHost, s_down));
        mark (s_down);
    }
    for (int i = 0; i < 3; ++i) FX_CUDA (e, cudaStreamSynchronize (e->pipe_stream[i]));
    if (trace)
    {
        for (size_t i = 0; i + 3 < tev.size(); i += 4)
        {
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime (&a, tev[0], tev[i]); cudaEventElapsedTime (&b, tev[0], tev[i + 1]);
            cudaEventElapsedTime (&c, tev[0], tev[i + 2]); cudaEventElapsedTime (&d, tev[0], tev[i + 3]);
            fprintf (stderr, "pipe group %3zu: upload %8.3f .. %8.3f  analysed %8.3f  results out %8.3f ms\n", i / 4, a, b, c, d);
        }
        for (auto ev : tev) cudaEventDestroy (ev);
    }
    e->flip ^= 1;
    e->frames_done += frames;
    return FX_OK;
}

} // namespace

fx_status fx_analyse_host (fx_engine* e, const float* audio, long track_stride, long n_samples,
                           float* raw, float* smooth, float* diag, long* n_frames)
{
    if (! e || ! audio || n_samples < 0 || track_stride < n_samples) return FX_ERR_INVALID_ARG;
    return analyse_host_impl (e, reinterpret_cast<const unsigned char*> (audio), track_stride * (long) sizeof (float), 0, 1, 0,
                              n_samples, raw, smooth, diag, n_frames);
}

// ---- file ingest ----------------------------------------------------------------------------------------
int fx_pcm_bytes_per_sample (int format) { return fx::pcm_bytes_per_sample (format); }

fx_status fx_analyse_host_pcm (fx_engine* e, const void* pcm, int format, int n_channels, int channel,
                               long track_stride_bytes, long n_samples,
                               float* raw, float* smooth, float* diag, long* n_frames)
{
    const int bps = fx::pcm_bytes_per_sample (format);
    if (! e || ! pcm || n_samples < 0 || n_channels < 1 || channel < -1 || channel >= n_channels) return FX_ERR_INVALID_ARG;
    if (bps == 0) return FX_ERR_UNSUPPORTED;
    if (track_stride_bytes != 0 && track_stride_bytes < n_samples * (long) n_channels * bps) return FX_ERR_INVALID_ARG;
    return analyse_host_impl (e, static_cast<const unsigned char*> (pcm), track_stride_bytes, format, n_channels, channel,
                              n_samples, raw, smooth, diag, n_frames);
}

fx_status fx_decode_pcm_device (fx_engine* e, const void* d_pcm, int format, int n_channels, int channel,
                                long track_stride_bytes, long n_samples, long n_tracks,
                                float* d_audio, long audio_stride, void* stream)
{
    const int bps = fx::pcm_bytes_per_sample (format);
    if (! e || ! d_pcm || ! d_audio || n_samples < 0 || n_tracks < 0 || n_channels < 1 || channel < -1 || channel >= n_channels
        || audio_stride < n_samples) return FX_ERR_INVALID_ARG;
    if (bps == 0) return FX_ERR_UNSUPPORTED;
    if (n_tracks > 1 && track_stride_bytes != 0 && track_stride_bytes < n_samples * (long) n_channels * bps) return FX_ERR_INVALID_ARG;
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    PcmParams pp{};
    pp.pcm = static_cast<const unsigned char*> (d_pcm); pp.track_stride_bytes = track_stride_bytes; pp.format = format;
    pp.n_channels = n_channels; pp.channel = channel; pp.n_samples = n_samples; pp.n_tracks = n_tracks;
    pp.audio = d_audio; pp.audio_stride = audio_stride;
    FX_CUDA (e, launch_pcm_decode (pp, stream ? static_cast<cudaStream_t> (stream) : e->stream));
    e->launches += (uint64_t) pcm_launch_count (pp);
    return FX_OK;
}

// ---- real-time path ---------------------------------------------------------------------------------------
fx_status fx_push_block (fx_engine* e, int first_track, int n_tracks, const float* const* channels, int n_samples)
{
    if (! e || ! channels || first_track < 0 || n_tracks < 0 || first_track + n_tracks > e->cfg.n_tracks || n_samples < 0)
        return FX_ERR_INVALID_ARG;
    const long L = e->ring_len;
    for (int i = 0; i < n_tracks; ++i)
    {
        const size_t t = (size_t) (first_track + i);
        const long w = e->wpos[t].load (std::memory_order_relaxed);
        if (w + n_samples - e->rpos > L) return FX_ERR_OVERRUN;          // would overwrite samples not analysed yet
        float* ring = e->h_ring + t * (size_t) L;
        const long o = w % L;
        const long first = (o + n_samples <= L) ? n_samples : L - o;
        memcpy (ring + o, channels[i], (size_t) first * sizeof (float));
        if (first < n_samples) memcpy (ring, channels[i] + first, (size_t) (n_samples - first) * sizeof (float));
        e->wpos[t].store (w + n_samples, std::memory_order_release);     // publish
    }
    return FX_OK;
}

fx_status fx_process (fx_engine* e, long* n_new_frames)
{
    if (! e) return FX_ERR_INVALID_ARG;
    if (n_new_frames) *n_new_frames = 0;
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    // every track advances together: the number of complete hops is set by the slowest producer
    long avail = -1;
    for (auto& w : e->wpos)
    {
        const long a = w.load (std::memory_order_acquire) - e->rpos;
        if (avail < 0 || a < avail) avail = a;
    }
    long hops = avail / e->H;
    if (hops > e->stage_hops) hops = e->stage_hops;
    if (hops <= 0) return FX_OK;
    const long L = e->ring_len, H = e->H;
    const long n = hops * H;

    fx_status st = ensure_results (e, hops);
    if (st != FX_OK) return st;
    st = ensure_records (e, hops);
    if (st != FX_OK) return st;
    // parameter upload and the chunk buffers are shared by the groups: settle them before fanning out
    st = upload_params (e, e->stream);
    if (st != FX_OK) return st;
    FX_CUDA (e, cudaStreamSynchronize (e->stream));
    const int n_chunks = choose_chunks (e, e->groups.empty() ? 1 : e->groups[0].n, hops);
    st = ensure_chunks (e, (long) e->cfg.n_tracks * n_chunks);
    if (st != FX_OK) return st;

    const size_t LW = FX_NUM_FEATURES + 2;
    for (auto& g : e->groups)
    {
        const long o = e->rpos % L;
        const long first = (o + n <= L) ? n : L - o;
        const float* src = e->h_ring + (size_t) g.t0 * (size_t) L;
        FX_CUDA (e, cudaMemcpy2DAsync (g.d_stage, (size_t) n * sizeof (float), src + o, (size_t) L * sizeof (float),
                                        (size_t) first * sizeof (float), (size_t) g.n, cudaMemcpyHostToDevice, g.stream));
        if (first < n)
            FX_CUDA (e, cudaMemcpy2DAsync (g.d_stage + first, (size_t) n * sizeof (float), src, (size_t) L * sizeof (float),
                                            (size_t) (n - first) * sizeof (float), (size_t) g.n, cudaMemcpyHostToDevice, g.stream));
        const size_t roff = (size_t) g.t0 * (size_t) hops;
        st = run_range (e, g.t0, g.n, n_chunks, g.d_stage, n, hops, e->d_raw + roff * FX_NUM_FEATURES, e->d_smooth + roff * FX_NUM_FEATURES,
                        nullptr, e->d_latest + (size_t) g.t0 * LW, g.stream);
        if (st != FX_OK) return st;
        FX_CUDA (e, cudaMemcpyAsync (e->h_latest + (size_t) g.t0 * LW, e->d_latest + (size_t) g.t0 * LW,
                                      (size_t) g.n * LW * sizeof (float), cudaMemcpyDeviceToHost, g.stream));
    }
    for (auto& g : e->groups) FX_CUDA (e, cudaStreamSynchronize (g.stream));
    e->rpos += n;
    e->flip ^= 1;
    e->frames_done += hops;
    if (n_new_frames) *n_new_frames = hops;
    return FX_OK;
}

fx_status fx_poll_features (fx_engine* e, int track, float out12[FX_NUM_FEATURES], uint64_t* frame_index)
{
    if (! e || track < 0 || track >= e->cfg.n_tracks || ! out12) return FX_ERR_INVALID_ARG;
    const float* row = e->h_latest + (size_t) track * (FX_NUM_FEATURES + 2);
    for (int k = 0; k < FX_NUM_FEATURES; ++k) out12[k] = row[k];
    if (frame_index)
    {
        uint32_t lo, hi;
        memcpy (&lo, row + FX_NUM_FEATURES, 4); memcpy (&hi, row + FX_NUM_FEATURES + 1, 4);
        *frame_index = ((uint64_t) hi << 32) | lo;
    }
    return FX_OK;
}

fx_status fx_flush (fx_engine* e)
{
    if (! e) return FX_ERR_INVALID_ARG;
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    for (auto& g : e->groups) FX_CUDA (e, cudaStreamSynchronize (g.stream));
    for (int i = 0; i < 3; ++i) if (e->pipe_stream[i]) FX_CUDA (e, cudaStreamSynchronize (e->pipe_stream[i]));
    FX_CUDA (e, cudaStreamSynchronize (e->stream));
    return FX_OK;
}

fx_status fx_osc_order (const float in12[FX_NUM_FEATURES], float* out, int n_out)
{
    if (! in12 || ! out) return FX_ERR_INVALID_ARG;
    // OSCFeatureAnalysisOutput.h:107
    static const int code12[12]   = { FX_ONSET, FX_RMS, FX_F0, FX_CENTROID, FX_SLOPE, FX_SPREAD, FX_FLATNESS, FX_LER, FX_FLUX, FX_HER, FX_OER, FX_INHARM };
    // README.md:55-57
    static const int readme10[10] = { FX_ONSET, FX_RMS, FX_F0, FX_CENTROID, FX_SLOPE, FX_SPREAD, FX_FLATNESS, FX_FLUX, FX_HER, FX_INHARM };
    if (n_out == 12) { for (int i = 0; i < 12; ++i) out[i] = in12[code12[i]]; return FX_OK; }
    if (n_out == 10) { for (int i = 0; i < 10; ++i) out[i] = in12[readme10[i]]; return FX_OK; }
    return FX_ERR_INVALID_ARG;
}

fx_status fx_synth_device (fx_engine* e, float* d_audio, long track_stride, long n_samples,
                           long first_track, uint64_t seed, void* stream)
{
    if (! e || ! d_audio) return FX_ERR_INVALID_ARG;
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    FX_CUDA (e, launch_synth (d_audio, track_stride, n_samples, e->cfg.n_tracks, first_track, e->cfg.sample_rate, seed,
                              stream ? (cudaStream_t) stream : e->stream));
    e->launches += 1;
    return FX_OK;
}

uint64_t fx_kernel_launches (const fx_engine* e) { return e ? e->launches.load() : 0; }

fx_status fx_profile_enable (fx_engine* e, int on)
{
    if (! e) return FX_ERR_INVALID_ARG;
    e->profiling = on != 0;
    return FX_OK;
}

fx_status fx_profile_read (fx_engine* e, double* ms_analyse, double* ms_post, long* n_calls)
{
    if (! e) return FX_ERR_INVALID_ARG;
    FX_CUDA (e, cudaSetDevice (e->cfg.device));
    double ka = 0.0, kp = 0.0;
    long n = 0;
    for (auto& pr : e->prof_pending)
    {
        FX_CUDA (e, cudaEventSynchronize (pr.c));
        float m1 = 0.0f, m2 = 0.0f;
        FX_CUDA (e, cudaEventElapsedTime (&m1, pr.a, pr.b));
        FX_CUDA (e, cudaEventElapsedTime (&m2, pr.b, pr.c));
        ka += m1; kp += m2; ++n;
        e->prof_free.push_back (pr);
    }
    e->prof_pending.clear();
    if (ms_analyse) *ms_analyse = ka;
    if (ms_post) *ms_post = kp;
    if (n_calls) *n_calls = n;
    return FX_OK;
}

fx_status fx_measure_fp32_peak (int device, double* tflops)
{
    if (! tflops) return FX_ERR_INVALID_ARG;
    if (cudaSetDevice (device) != cudaSuccess) return FX_ERR_NO_DEVICE;
    return fx::measure_fp32_peak (tflops) == cudaSuccess ? FX_OK : FX_ERR_CUDA;
}

} // extern "C"
