// fx_analyse.cu -- K1, the per-frame analysis kernel (sm_100a).
//
// One CTA walks a chunk of consecutive frames of one track.  Per frame it reproduces, on the GPU, the two
// analyser bodies of the reference (all citations relative to /root/reference/Source/):
//   RealTimeSpectralAnalyser::run  (RealTimeAnalyser.h:205-229)   RMS, Bartlett window, FFT, spectral features, slope
//   RealTimeHarmonicAnalyser::run  (RealTimeAnalyser.h:145-172)   RMS, one-pole filter, window, 2 FFTs, pitch, harmonic features
// including their quirks: "magnitude" = Re(X)^2, asymmetric Bartlett window, harmonic
// features on the un-windowed frame, integer pitch lag, raw fp64 flatness product with IEEE under/overflow,
// previous spectrum not updated on silent frames.
//
// Data movement: the hop's new samples arrive by cp.async.bulk (TMA bulk copy, mbarrier completion) into a
// shared-memory ring holding the N-sample window, so each input sample crosses HBM once per chunk; the next
// hop is prefetched while the current frame's FFTs run.  Three complex FFTs per frame:
//   FFT1  z = x + i (x * w)      -> Re A (raw frame), Re B / Im B (windowed frame) by conjugate symmetry
//   FFT2  c = onepole (x) * w    -> P[k] = Re C[k]^2
//   FFT3  inverse of P (chained in registers from FFT2) -> d[s], only the real part is used (PitchAnalyser.h:163)
// All feature reductions accumulate in fp64 like the reference.  No tensor cores: nothing here is a GEMM.
#include "fx_fft.cuh"
#include "fx_kernels.cuh"
#include <math.h>

namespace fx {

// ---------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA, non-tensor form)
__device__ __forceinline__ uint32_t smem_u32 (const void* p) { return (uint32_t) __cvta_generic_to_shared (p); }

__device__ __forceinline__ void mbar_init (uint64_t* bar, uint32_t count)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx (uint64_t* bar, uint32_t bytes)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive (uint64_t* bar)
{
    asm volatile ("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint64_t* bar, uint32_t parity)
{
    asm volatile (
        "{\n"
        ".reg .pred p;\n"
        "FX_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FX_WAIT_DONE;\n"
        "bra FX_WAIT_LOOP;\n"
        "FX_WAIT_DONE:\n"
        "}\n" :: "r"(smem_u32 (bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s (void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :: "r"(smem_u32 (dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile ("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------
// small numeric helpers
__device__ __forceinline__ float relmargin (double a, double b)
{
    const double m = fmax (fabs (a), fabs (b));
    return (m > 0.0) ? (float) (fabs (a - b) / m) : 0.0f;
}

// extended-range product: value = m * 2^e with m in [0.5, 1)
struct ME { double m; int e; };
__device__ __forceinline__ ME me_one() { ME r; r.m = 0.5; r.e = 1; return r; }
__device__ __forceinline__ ME me_from (double x) { ME r; r.m = frexp (x, &r.e); return r; }
__device__ __forceinline__ ME me_mul (ME a, ME b)
{
    ME r; int ex;
    r.m = frexp (a.m * b.m, &ex);
    r.e = a.e + b.e + ex;
    return r;
}

template <int K> __device__ __forceinline__ void warp_sum (double (&v)[K])
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        #pragma unroll
        for (int k = 0; k < K; ++k) v[k] += __shfl_xor_sync (0xffffffffu, v[k], off);
}
__device__ __forceinline__ double warp_max (double v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax (v, __shfl_xor_sync (0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ float warp_minf (float v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fminf (v, __shfl_xor_sync (0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ unsigned warp_minu (unsigned v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = min (v, __shfl_xor_sync (0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ unsigned long long warp_minull (unsigned long long v)
{
    #pragma unroll
    for (int off = 16; off > 0; off >>= 1) { const unsigned long long o = __shfl_xor_sync (0xffffffffu, v, off); v = o < v ? o : v; }
    return v;
}

// ---------------------------------------------------------------------------------------------------------
template <int R1> struct Smem
{
    using D = FftDims<R1>;
    static constexpr int N = D::N, M = N / 2, T = D::T, NW = T / 32;
    static constexpr int kRed = 12;

    float2   ex[D::EX_LEN];          // FFT exchange buffer; doubles as the fp32 work array (skewed, N*17/16 floats)
    float2   tw1[D::TW1_LEN];
    float2   tw2[D::TW2_LEN];
    float    ring[N];                // ring[a & (N-1)] = absolute sample a of the track
    float    specA[M];               // Re FFT (raw frame)        -> harmonic features
    float    specB[2][M];            // Re FFT (windowed frame)   -> spectral features; [cur ^ 1] = previous non-silent frame
    double   red[2][kRed][NW];       // block-reduction partials, double buffered by phase parity
    double   scan_m[NW];             // flatness product scan, warp totals
    int      scan_e[NW];
    double   pscan[NW];              // pitch cumulative sum scan, warp totals
    unsigned long long keys[NW];
    unsigned ucodes[2][NW];
    float    fmins[2][NW];
    double   flat_prod;
    double   f0;
    float    her_terms[18];
    int      her_bins[18];
    uint64_t mbar;
};

template <int R1>
__global__ void __launch_bounds__ (16 * R1, (R1 == 16 ? 2 : (R1 == 8 ? 4 : 8)))
k_analyse (const AnalyseParams p)
{
    using D = FftDims<R1>;
    using S = Smem<R1>;
    constexpr int N = D::N, M = N / 2, T = D::T, NW = T / 32, Q1 = D::Q1;

    extern __shared__ __align__ (128) unsigned char smem_raw[];
    S& sm = *reinterpret_cast<S*> (smem_raw);
    float* workf = reinterpret_cast<float*> (sm.ex);

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long cta = blockIdx.x;
    const int chunk = (int) (cta % p.n_chunks);
    const long track = cta / p.n_chunks;
    const int f_begin = chunk * p.frames_per_chunk;
    const int f_end = min (p.n_frames, f_begin + p.frames_per_chunk);
    if (f_begin >= f_end)
    {
        if (t == 0 && f_begin < p.n_chunks * p.frames_per_chunk) p.first_idx[track * p.n_chunks + chunk] = -1;
        return;
    }

    const int H = p.hop, NB = N >> p.log2_hop;
    const double nyquist = p.sample_rate / 2.0;
    const double frpb = nyquist / (double) M;
    const float gain = p.gain[track];
    const float* src = p.audio + track * p.track_stride;
    const float* tail = p.tail_in + track * (long) (N - H);

    for (int i = t; i < D::TW1_LEN; i += T) sm.tw1[i] = p.tw1[i];
    for (int i = t; i < D::TW2_LEN; i += T) sm.tw2[i] = p.tw2[i];
    for (int i = t; i < M; i += T) { sm.specB[0][i] = 0.0f; sm.specB[1][i] = 0.0f; }
    if (t == 0) { mbar_init (&sm.mbar, 1); }
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // ---- fill the ring with the window of the chunk's first frame ------------------------------------
    // frame f (call-relative) ends with hop block f of this call = absolute hop first_hop + f
    {
        const long j_new = p.first_hop + f_begin;
        uint32_t bulk_bytes = 0;
        for (int b = 0; b < NB; ++b)
        {
            const long j = j_new - (NB - 1) + b;                 // absolute hop index
            float* dst = sm.ring + (int) ((j * H) & (N - 1));
            const float* g = nullptr;
            if (j >= p.first_hop)                  g = src + (j - p.first_hop) * H;
            else if (j >= 0)                       g = tail + (j - (p.first_hop - (NB - 1))) * H;
            if (g == nullptr)      { for (int i = t; i < H; i += T) dst[i] = 0.0f; }      // before the stream started (RealTimeAudioAnalysis.h:202)
            else if (! p.use_bulk) { for (int i = t; i < H; i += T) dst[i] = g[i]; }
            else                   bulk_bytes += (uint32_t) H * 4u;
        }
        if (p.use_bulk && t == 0)
        {
            fence_proxy_async();
            mbar_expect_tx (&sm.mbar, bulk_bytes);
            for (int b = 0; b < NB; ++b)
            {
                const long j = j_new - (NB - 1) + b;
                float* dst = sm.ring + (int) ((j * H) & (N - 1));
                if (j >= p.first_hop)      bulk_g2s (dst, src + (j - p.first_hop) * H, (uint32_t) H * 4u, &sm.mbar);
                else if (j >= 0)           bulk_g2s (dst, tail + (j - (p.first_hop - (NB - 1))) * H, (uint32_t) H * 4u, &sm.mbar);
            }
        }
        if (! p.use_bulk && t == 0) mbar_arrive (&sm.mbar);
        __syncthreads();
    }

    uint32_t phase = 0;
    int cur = 0;                      // specB[cur] receives this frame, specB[cur ^ 1] is the previous non-silent frame
    bool have_prev = false;           // false until the chunk's first non-silent frame (its flux is fixed up by K2)
    int first_nonsilent = -1;
    const float max_flux = (float) (M * (M + 1)) / 2.0f;            // SpectralCharacteristics.h:111
    const int lower_portion = M / 5;                                // :65

    for (int f = f_begin; f < f_end; ++f)
    {
        const long j_new = p.first_hop + f;
        const long a0 = (j_new - (NB - 1)) * (long) H;              // absolute sample index of window sample 0
        float* out = p.raw + (track * p.n_frames + f) * FX_NUM_FEATURES;
        float* dg  = p.diag ? p.diag + (track * p.n_frames + f) * FX_NUM_DIAG : nullptr;

        mbar_wait (&sm.mbar, phase);
        phase ^= 1u;

        // =========================== FFT1: z = x + i (x * bartlett) ====================================
        float2 v[16];
        double rms_part = 0.0;
        {
            #pragma unroll
            for (int q = 0; q < Q1; ++q)
                #pragma unroll
                for (int n1 = 0; n1 < R1; ++n1)
                {
                    const int n = n1 * 256 + t + T * q;
                    const float x = __fmul_rn (sm.ring[(int) ((a0 + n) & (N - 1))], gain);       // AudioDataCollector.h:88
                    // RealTimeAudioAnalysis.h:148-149: w[n] = n * 2/N, w[N/2 + n] = 1 - n * 2/N (exact in fp32)
                    const float w = (n < M) ? (float) n * (2.0f / N) : 1.0f - (float) (n - M) * (2.0f / N);
                    v[q * R1 + n1] = make_float2 (x, __fmul_rn (x, w));
                    rms_part += (double) __fmul_rn (x, x);                                       // getRMSLevel: fp32 square, fp64 sum
                }
        }
        fft_stage1_store<R1, false> (v, t, sm.ex, sm.tw1);
        {
            double r1[1] = { rms_part };
            warp_sum<1> (r1);
            if (lane == 0) sm.red[0][0][warp] = r1[0];
        }
        __syncthreads();                                                                          // (1)
        fft_stage2<R1, false> (t, sm.ex, sm.tw2);
        __syncthreads();                                                                          // (2)
        fft_stage3<R1, false> (t, sm.ex, v);
        __syncthreads();                                                                          // (3)
        {
            const int kl = klow<R1> (t);
            #pragma unroll
            for (int s = 0; s < 16; ++s) sm.ex[phys (kl + T * out_index<16> (s))] = v[s];
        }
        __syncthreads();                                                                          // (4)

        // split the packed spectrum: A = FFT (x), B = FFT (x w); keep Re A, Re B; Im B only for the slope quirk
        float rawmax = 0.0f;     // SpectralCharacteristics.h:153: max |buf[j]|, j < M, over the interleaved Re/Im floats = bins k < M/2
        #pragma unroll
        for (int j = 0; j < 8; ++j)
        {
            const int k = t + T * j;
            const float2 zk = sm.ex[phys (k)];
            const float2 zn = sm.ex[phys ((N - k) & (N - 1))];
            const float reA = 0.5f * (zk.x + zn.x);
            const float reB = 0.5f * (zk.y + zn.y);
            const float imB = 0.5f * (zn.x - zk.x);
            sm.specA[k] = reA;
            sm.specB[cur][k] = reB;
            if (k < M / 2) rawmax = fmaxf (rawmax, fmaxf (fabsf (reB), fabsf (imB)));
        }
        __syncthreads();                                                                          // (5)

        // RMS (RealTimeAnalyser.h:207-208)
        double rms_sum = 0.0;
        #pragma unroll
        for (int w = 0; w < NW; ++w) rms_sum += sm.red[0][0][w];
        const float rms = (float) sqrt (rms_sum / (double) N);
        const float log_rms = (float) log10 ((double) __fadd_rn (__fmul_rn (rms, 9.0f), 1.0f));
        const double eps = 0.01 * (double) log_rms;                                               // SpectralCharacteristics.h:108

        // =========================== spectral features, pass 1 ========================================
        double mag[8];
        const int b0 = 8 * t;
        ME lprod = me_one();
        {
            const float4 c0 = *reinterpret_cast<const float4*> (&sm.specB[cur][b0]);
            const float4 c1 = *reinterpret_cast<const float4*> (&sm.specB[cur][b0 + 4]);
            const float4 p0 = *reinterpret_cast<const float4*> (&sm.specB[cur ^ 1][b0]);
            const float4 p1 = *reinterpret_cast<const float4*> (&sm.specB[cur ^ 1][b0 + 4]);
            const float cr[8] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w };
            const float pr[8] = { p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w };
            double mag_sum = 0.0, weighted = 0.0, flux = 0.0, lhr = 0.0, flat_sum = 0.0, count = 0.0, maxmag = 0.0;
            float fmargin = 1.0f;
            #pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const int bin = b0 + j;
                const double re = (double) cr[j];
                const double mg = re * re;                                                       // :72-73  Re^2
                const double pm = (double) pr[j] * (double) pr[j];
                const double diff = mg - pm;                                                     // :76-79
                if (diff > 0.0) flux += diff;
                mag[j] = mg;
                mag_sum += mg;
                if (bin <= lower_portion) lhr += mg;                                             // :86-87
                if (mg > eps)                                                                    // :89-94
                {
                    flat_sum += mg;
                    count += 1.0;
                    lprod = me_mul (lprod, me_from (mg));
                }
                fmargin = fminf (fmargin, relmargin (mg, eps));
                const double fc = (double) bin * frpb + (frpb / 2.0);                            // :70
                weighted += fc * mg;
                maxmag = fmax (maxmag, mg);
            }
            double s6[6] = { mag_sum, weighted, flux, lhr, flat_sum, count };
            warp_sum<6> (s6);
            const double wmax = warp_max (maxmag);
            const double wraw = warp_max ((double) rawmax);
            const float wmar = warp_minf (fmargin);
            // inclusive warp scan of the extended-range product, in bin order
            ME inc = lprod;
            #pragma unroll
            for (int off = 1; off < 32; off <<= 1)
            {
                ME o; o.m = __shfl_up_sync (0xffffffffu, inc.m, off); o.e = __shfl_up_sync (0xffffffffu, inc.e, off);
                if (lane >= off) inc = me_mul (o, inc);
            }
            ME exc; exc.m = __shfl_up_sync (0xffffffffu, inc.m, 1); exc.e = __shfl_up_sync (0xffffffffu, inc.e, 1);
            if (lane == 0) exc = me_one();
            lprod = exc;                                                                         // lane-exclusive prefix within the warp
            if (lane == 31) { sm.scan_m[warp] = inc.m; sm.scan_e[warp] = inc.e; }
            if (lane == 0)
            {
                #pragma unroll
                for (int k = 0; k < 6; ++k) sm.red[1][k][warp] = s6[k];
                sm.red[1][6][warp] = wmax;
                sm.red[1][7][warp] = wraw;
                sm.fmins[0][warp] = wmar;
            }
        }
        __syncthreads();                                                                          // (6)
        double mag_sum = 0.0, weighted = 0.0, flux = 0.0, lhr = 0.0, flat_sum = 0.0, count = 0.0, maxmag = 0.0, rawmax_d = 0.0;
        float flat_margin = 1.0f;
        ME total = me_one(), prefix = me_one();
        #pragma unroll
        for (int w = 0; w < NW; ++w)
        {
            mag_sum += sm.red[1][0][w]; weighted += sm.red[1][1][w]; flux += sm.red[1][2][w];
            lhr += sm.red[1][3][w]; flat_sum += sm.red[1][4][w]; count += sm.red[1][5][w];
            maxmag = fmax (maxmag, sm.red[1][6][w]); rawmax_d = fmax (rawmax_d, sm.red[1][7][w]);
            flat_margin = fminf (flat_margin, sm.fmins[0][w]);
            ME wt; wt.m = sm.scan_m[w]; wt.e = sm.scan_e[w];
            if (w < warp) prefix = me_mul (prefix, wt);
            total = me_mul (total, wt);
        }
        prefix = me_mul (prefix, lprod);
        const bool silent = ! (mag_sum > 0.05);                                                   // :121-123
        const float centroid = (float) (weighted / mag_sum);                                      // :127
        const double max_e = fmax (rawmax_d, maxmag);                                             // :153-163
        const bool slope_gated = ! (max_e > 0.0001);                                              // :165-167

        // =========================== pass 2: spread, slope sums, flatness range events ================
        unsigned ev_code = 0xffffffffu;
        double ev_before = 1.0;
        {
            double var = 0.0, se = 0.0, sie = 0.0;
            ME run = prefix;
            #pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const int bin = b0 + j;
                const double fc = (double) bin * frpb + (frpb / 2.0);
                const double dv = (fc / nyquist) - ((double) centroid / nyquist);                 // :137
                var += (dv * dv) * mag[j];
                const double e = mag[j] / max_e;                                                  // :172
                se += e;
                sie += (double) bin * e;                                                          // :175
                if (mag[j] > eps)
                {
                    const ME nxt = me_mul (run, me_from (mag[j]));
                    if (ev_code == 0xffffffffu && (nxt.e >= 1025 || nxt.e <= -1022))
                    {
                        ev_code = (unsigned) bin * 2u + (nxt.e >= 1025 ? 1u : 0u);
                        ev_before = ldexp (run.m, run.e);
                    }
                    run = nxt;
                }
            }
            double s3[3] = { var, se, sie };
            warp_sum<3> (s3);
            const unsigned wev = warp_minu (ev_code);
            if (lane == 0)
            {
                sm.red[0][1][warp] = s3[0]; sm.red[0][2][warp] = s3[1]; sm.red[0][3][warp] = s3[2];
                sm.ucodes[0][warp] = wev;
            }
        }
        __syncthreads();                                                                          // (7)
        double var = 0.0, se = 0.0, sie = 0.0;
        unsigned ev = 0xffffffffu;
        #pragma unroll
        for (int w = 0; w < NW; ++w)
        {
            var += sm.red[0][1][w]; se += sm.red[0][2][w]; sie += sm.red[0][3][w];
            ev = min (ev, sm.ucodes[0][w]);
        }
        const double mean_e = se / (double) M;                                                    // :177
        {
            // pass 3: energy variance (:182-190)
            double evar = 0.0;
            #pragma unroll
            for (int j = 0; j < 8; ++j) { const double d = mag[j] / max_e - mean_e; evar += d * d; }
            double s1[1] = { evar };
            warp_sum<1> (s1);
            if (lane == 0) sm.red[1][0][warp] = s1[0];
            // flatness product left the normal fp64 range at bin ev >> 1: the owner of that bin replays the
            // reference's sequential IEEE multiply (gradual underflow included) from there (:92)
            if (ev != 0xffffffffu && ev == ev_code)
            {
                double prod;
                if (ev & 1u) prod = INFINITY;
                else
                {
                    prod = ev_before;
                    for (int b = (int) (ev >> 1); b < M; ++b)
                    {
                        const double re = (double) sm.specB[cur][b];
                        const double mg = re * re;
                        if (mg > eps)
                        {
                            prod *= mg;
                            if (prod == 0.0 || isinf (prod)) break;
                        }
                    }
                }
                sm.flat_prod = prod;
            }
        }
        __syncthreads();                                                                          // (8)
        if (t == 0)
        {
            double evar = 0.0;
            #pragma unroll
            for (int w = 0; w < NW; ++w) evar += sm.red[1][0][w];
            float gate_margin = fminf (relmargin (mag_sum, 0.05), relmargin (max_e, 0.0001));
            float o_centroid = 0.0f, o_spread = 0.0f, o_flat = 0.0f, o_ler = 0.0f, o_flux = 0.0f, o_slope = 0.0f;
            float flat_state = 3.0f;
            if (! silent)
            {
                double product;
                if (ev == 0xffffffffu) { product = ldexp (total.m, total.e); flat_state = 0.0f; }
                else
                {
                    product = sm.flat_prod;
                    flat_state = (product == 0.0) ? 1.0f : (isinf (product) ? 2.0f : 0.0f);
                }
                const double inv = 1.0 / (count > 0.0 ? count : 1.0);                             // :130
                const float flat = flat_sum > eps ? (float) (pow (product, inv) / (inv * flat_sum)) : 0.0f;    // :57-60
                o_flat = (float) log10 ((double) flat * 9.0 + 1.0);                               // :132
                const float c = __fdiv_rn (centroid, (float) (nyquist / 2.0));                    // :133
                o_centroid = (float) log10 ((double) __fadd_rn (__fmul_rn (c, 9.0f), 1.0f));      // :134
                const float max_spread = (float) (((double) centroid / nyquist) * (1.0 - ((double) centroid / nyquist)));   // :140
                o_spread = (float) ((var / mag_sum) / (double) max_spread);                       // :141
                o_ler = (float) (lhr / mag_sum);                                                  // :125
                o_flux = have_prev ? (float) (flux / (double) max_flux) : 0.0f;                   // :112 (K2 fixes the chunk's first non-silent frame)
            }
            if (! slope_gated)
            {
                const double energy_var = evar / (double) M;
                const double bin_std = sqrt (p.bin_var), energy_std = sqrt (energy_var);
                const double r = (sie - ((double) M * mean_e * 0.5)) / (double) ((float) M - 1.0f) * energy_std * bin_std;   // :195
                o_slope = (float) (r * (bin_std / energy_std));                                   // :198
            }
            out[FX_RMS] = log_rms;
            out[FX_CENTROID] = o_centroid; out[FX_SPREAD] = o_spread; out[FX_FLATNESS] = o_flat;
            out[FX_LER] = o_ler; out[FX_FLUX] = o_flux; out[FX_SLOPE] = o_slope; out[FX_ONSET] = 0.0f;
            if (dg)
            {
                dg[FX_DIAG_FLAT_COUNT] = (float) count;
                dg[FX_DIAG_FLAT_MARGIN] = flat_margin;
                dg[FX_DIAG_FLAT_STATE] = flat_state;
                dg[FX_DIAG_GATE_MARGIN] = gate_margin;
                dg[FX_DIAG_ONSET_MARGIN] = 1.0f;
            }
        }
        if (! silent)
        {
            if (! have_prev)
            {
                have_prev = true;
                first_nonsilent = f;
                float* fs = p.first_spec + (track * p.n_chunks + chunk) * (long) M;
                for (int i = t; i < M; i += T) fs[i] = sm.specB[cur][i];
            }
            cur ^= 1;                                                                             // :138 prev <- current
        }

        // =========================== one-pole filter + window -> work array ============================
        // AudioFilter::filterAudio (RealTimeAudioAnalysis.h:106-125): y[0] = x[0]; y[n] = (pi/2) x[n] + e^(-pi/2) y[n-1].
        // Each thread owns 16 consecutive samples; the recurrence is warmed up over the 16 samples before them
        // (e^(-pi/2)^16 = 1.2e-11 of the state survives, far below fp32 resolution), then run in the reference's arithmetic.
        {
            const int n0 = 16 * t;
            float y = 0.0f;
            if (t > 0)
            {
                const int w0 = n0 - 16;
                #pragma unroll
                for (int j = 0; j < 16; ++j)
                {
                    const float x = __fmul_rn (sm.ring[(int) ((a0 + w0 + j) & (N - 1))], gain);
                    y = (w0 + j == 0) ? x : __fadd_rn (__fmul_rn (p.iir_c1, x), __fmul_rn (p.iir_c2, y));
                }
            }
            #pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                const int n = n0 + j;
                const float x = __fmul_rn (sm.ring[(int) ((a0 + n) & (N - 1))], gain);
                y = (n == 0) ? x : __fadd_rn (__fmul_rn (p.iir_c1, x), __fmul_rn (p.iir_c2, y));
                const float w = (n < M) ? (float) n * (2.0f / N) : 1.0f - (float) (n - M) * (2.0f / N);
                workf[17 * t + j] = __fmul_rn (y, w);                                             // phys (16 t + j)
            }
        }
        __syncthreads();                                                                          // (9) ring is free: prefetch the next hop
        if (f + 1 < f_end)
        {
            const long jn = j_new + 1;
            float* dst = sm.ring + (int) ((jn * H) & (N - 1));
            const float* g = src + (jn - p.first_hop) * H;
            if (p.use_bulk)
            {
                if (t == 0)
                {
                    fence_proxy_async();
                    mbar_expect_tx (&sm.mbar, (uint32_t) H * 4u);
                    bulk_g2s (dst, g, (uint32_t) H * 4u, &sm.mbar);
                }
            }
            else
            {
                for (int i = t; i < H; i += T) dst[i] = g[i];       // visible after the barriers below
                if (t == 0) mbar_arrive (&sm.mbar);
            }
        }

        // =========================== FFT2 (filtered, windowed) chained into FFT3 (inverse) =============
        #pragma unroll
        for (int q = 0; q < Q1; ++q)
            #pragma unroll
            for (int n1 = 0; n1 < R1; ++n1)
                v[q * R1 + n1] = make_float2 (workf[phys (n1 * 256 + t + T * q)], 0.0f);
        __syncthreads();                                                                          // (10)
        fft_stage1_store<R1, false> (v, t, sm.ex, sm.tw1);
        __syncthreads();                                                                          // (11)
        fft_stage2<R1, false> (t, sm.ex, sm.tw2);
        __syncthreads();                                                                          // (12)
        {
            float2 c3[16];
            fft_stage3<R1, false> (t, sm.ex, c3);
            // PitchAnalyser::getComplexConjugateMultiplication (PitchAnalyser.h:97-104): Re^2, imaginary cleared
            #pragma unroll
            for (int s = 0; s < 16; ++s) c3[s] = make_float2 (__fmul_rn (c3[s].x, c3[s].x), 0.0f);
            chain_permute<R1> (c3, v);
        }
        __syncthreads();                                                                          // (13)
        fft_stage1_store<R1, true> (v, klow<R1> (t), sm.ex, sm.tw1);
        __syncthreads();                                                                          // (14)
        fft_stage2<R1, true> (t, sm.ex, sm.tw2);
        __syncthreads();                                                                          // (15)
        fft_stage3<R1, true> (t, sm.ex, v);
        __syncthreads();                                                                          // (16)
        {
            // performRealOnlyInverseTransform scales by 1/N; only the real half d[0..N) can reach the lag search
            const int kl = klow<R1> (t);
            #pragma unroll
            for (int s = 0; s < 16; ++s) workf[phys (kl + T * out_index<16> (s))] = __fmul_rn (v[s].x, 1.0f / N);
        }
        __syncthreads();                                                                          // (17)

        // =========================== pitch: cumulative normalised difference + lag search ==============
        float ac[16];
        double lsum[16];
        double seg_exc;
        {
            double run = 0.0;
            #pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                const int s = 16 * t + j;
                const float d = workf[17 * t + j];
                ac[j] = __fmul_rn (__fmul_rn (d, d), (float) s);                                 // PitchAnalyser.h:122-123
                if (s >= 1) run += (double) ac[j];
                lsum[j] = run;
            }
            // block scan of the segment totals
            double inc = run;
            #pragma unroll
            for (int off = 1; off < 32; off <<= 1)
            {
                const double o = __shfl_up_sync (0xffffffffu, inc, off);
                if (lane >= off) inc += o;
            }
            seg_exc = __shfl_up_sync (0xffffffffu, inc, 1);
            if (lane == 0) seg_exc = 0.0;
            if (lane == 31) sm.pscan[warp] = inc;
        }
        __syncthreads();                                                                          // (18)
        unsigned first_cross = 0xffffffffu;
        {
            double base = seg_exc;
            #pragma unroll
            for (int w = 0; w < NW; ++w) if (w < warp) base += sm.pscan[w];
            unsigned long long key = ((unsigned long long) __float_as_uint (100.0f) << 32) | 0xffffffffull;
            #pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                const int s = 16 * t + j;
                const float sumf = (float) (base + lsum[j]);                                     // :145 (fp32 running sum in the reference)
                float c = (sumf != 0.0f) ? __fdiv_rn (ac[j], sumf) : 0.0f;                       // :146-154
                if (s == 0) c = 1.0f;                                                             // :141
                workf[17 * t + j] = c;
                if (s >= 2)
                {
                    if (c < 0.01f && first_cross == 0xffffffffu) first_cross = (unsigned) s;      // :176
                    const unsigned long long k2 = ((unsigned long long) __float_as_uint (c) << 32) | (unsigned) s;
                    if (k2 < key && c >= 0.0f) key = k2;                                          // :171-175 first strict minimum
                }
                ac[j] = c;
            }
            const unsigned wfc = warp_minu (first_cross);
            const unsigned long long wkey = warp_minull (key);
            if (lane == 0) { sm.ucodes[1][warp] = wfc; sm.keys[warp] = wkey; }
        }
        // harmonic pass A (independent of the pitch): sum and max of Re A ^2 (HarmonicCharacteristics.h:61-69)
        {
            const float4 a0v = *reinterpret_cast<const float4*> (&sm.specA[b0]);
            const float4 a1v = *reinterpret_cast<const float4*> (&sm.specA[b0 + 4]);
            const float ar[8] = { a0v.x, a0v.y, a0v.z, a0v.w, a1v.x, a1v.y, a1v.z, a1v.w };
            double hsum = 0.0, hmax = 0.0;
            #pragma unroll
            for (int j = 0; j < 8; ++j) { const double re = (double) ar[j]; mag[j] = re * re; hsum += mag[j]; hmax = fmax (hmax, mag[j]); }
            double s1[1] = { hsum };
            warp_sum<1> (s1);
            const double wm = warp_max (hmax);
            if (lane == 0) { sm.red[0][4][warp] = s1[0]; sm.red[0][5][warp] = wm; }
        }
        __syncthreads();                                                                          // (19)
        unsigned s0 = 0xffffffffu;
        unsigned long long gkey = ~0ull;
        double hsum = 0.0, hmax = 0.0;
        #pragma unroll
        for (int w = 0; w < NW; ++w)
        {
            s0 = min (s0, sm.ucodes[1][w]);
            gkey = sm.keys[w] < gkey ? sm.keys[w] : gkey;
            hsum += sm.red[0][4][w]; hmax = fmax (hmax, sm.red[0][5][w]);
        }
        const bool crossed = (s0 != 0xffffffffu);
        {
            // phase C: end of the descending run that starts at s0 (:178-181), margins of the threshold tests,
            // runner-up of the global minimum (margin only)
            unsigned send = 0xffffffffu;
            float pm = 1.0f;
            float second = 100.0f;
            const unsigned gidx = (unsigned) (gkey & 0xffffffffull);
            #pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                const unsigned s = (unsigned) (16 * t + j);
                if (s < 2u) continue;
                if (crossed)
                {
                    if (s <= s0) pm = fminf (pm, relmargin (ac[j], 0.01));
                    if (s >= s0 && send == 0xffffffffu)
                    {
                        const bool has_next = (s + 1u < (unsigned) N);
                        const float nxt = has_next ? workf[phys ((int) s + 1)] : 0.0f;
                        if (! (has_next && nxt < ac[j])) send = s;
                    }
                }
                else
                {
                    pm = fminf (pm, relmargin (ac[j], 0.01));
                    if (s != gidx) second = fminf (second, ac[j]);
                }
            }
            const unsigned wsend = warp_minu (send);
            const float wpm = warp_minf (pm);
            const float wsec = warp_minf (second);
            if (lane == 0) { sm.ucodes[0][warp] = wsend; sm.fmins[0][warp] = wpm; sm.fmins[1][warp] = wsec; }
        }
        __syncthreads();                                                                          // (20)
        if (t == 0)
        {
            unsigned send = 0xffffffffu; float pm = 1.0f, second = 100.0f;
            #pragma unroll
            for (int w = 0; w < NW; ++w) { send = min (send, sm.ucodes[0][w]); pm = fminf (pm, sm.fmins[0][w]); second = fminf (second, sm.fmins[1][w]); }
            float lag;
            if (crossed)
            {
                // getInterpolatedValleyFromCumulativeDifferenceLagEstimate (:192-203): the parabolic branch is unreachable
                const int s_end = (int) send;
                const int right = s_end + 1;
                const float c_end = workf[phys (s_end)];
                const float c_right = (right < N) ? workf[phys (right)] : 0.0f;      // cnd[N] = Im part of lag 0 = 0
                for (int s = (int) s0; s < s_end; ++s) pm = fminf (pm, relmargin (workf[phys (s + 1)], workf[phys (s)]));
                pm = fminf (pm, relmargin (c_end, c_right));
                lag = (c_end <= c_right) ? (float) s_end : (float) right;
            }
            else
            {
                const unsigned gidx = (unsigned) (gkey & 0xffffffffull);
                lag = (gidx == 0xffffffffu) ? -1.0f : (float) gidx;                               // :165,188
                pm = fminf (pm, relmargin (__uint_as_float ((unsigned) (gkey >> 32)), second));
            }
            const double f0 = (nyquist * 2.0) / (double) lag;                                     // :57
            sm.f0 = f0;
            out[FX_F0] = (float) (f0 / 5000.0);                                                   // RealTimeAnalyser.h:165-166
            if (dg) { dg[FX_DIAG_LAG] = lag; dg[FX_DIAG_PITCH_MARGIN] = pm; }
        }
        __syncthreads();                                                                          // (21)

        // =========================== harmonic features (HarmonicCharacteristics.h:46-106) =============
        const double f0 = sm.f0;
        const bool hsilent = hsum < 0.005;                                                        // :88
        const double mean_mag = hsum / (double) M;                                                // :86
        const int f0_bin = (int) floor (f0 / frpb);                                               // :246-249
        {
            double sum_normed = 0.0, inharm = 0.0, npeaks = 0.0;
            float pkm = 1.0f;
            // neighbours bin-2, bin-1, bin+1 (:136-143, loop end exclusive)
            const double l2 = (b0 >= 2) ? (double) sm.specA[b0 - 2] * (double) sm.specA[b0 - 2] : 0.0;
            const double l1 = (b0 >= 1) ? (double) sm.specA[b0 - 1] * (double) sm.specA[b0 - 1] : 0.0;
            const double r1 = (b0 + 8 < M) ? (double) sm.specA[b0 + 8] * (double) sm.specA[b0 + 8] : 0.0;
            float nm[8];
            #pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const int bin = b0 + j;
                const double e = mag[j] / hmax;                                                   // :75
                nm[j] = (float) e;
                sum_normed += e;
                const double mg = mag[j];
                pkm = fminf (pkm, relmargin (mg, mean_mag));
                if (mg > mean_mag)
                {
                    const double m2 = (j >= 2) ? mag[j - 2] : (j == 1 ? l1 : l2);
                    const double m1 = (j >= 1) ? mag[j - 1] : l1;
                    const double p1 = (j < 7) ? mag[j + 1] : r1;
                    // edge clamps (:136-137): the neighbour window is [max (bin-2, 0), min (bin+2, M-1))
                    bool peak = true;
                    const int lo = bin - 2 > 0 ? bin - 2 : 0;
                    const int hi = bin + 2 < M - 1 ? bin + 2 : M - 1;                             // exclusive
                    if (bin - 2 >= lo && bin - 2 < hi) { pkm = fminf (pkm, relmargin (m2, mg)); if (m2 > mg) peak = false; }
                    if (peak && bin - 1 >= lo && bin - 1 < hi) { pkm = fminf (pkm, relmargin (m1, mg)); if (m1 > mg) peak = false; }
                    if (peak && bin + 1 < hi) { pkm = fminf (pkm, relmargin (p1, mg)); if (p1 > mg) peak = false; }
                    if (peak)
                    {
                        npeaks += 1.0;
                        if (f0 > 0.0 && bin != f0_bin)                                            // :98, :220
                        {
                            double start_f = (double) bin * frpb;                                 // :223
                            if (start_f == 0.0) start_f = frpb * 0.5;
                            const double end_f = (double) (bin + 1) * frpb;
                            const double ra = (start_f == f0) ? 1.0 : (start_f > f0 ? start_f / f0 : f0 / start_f);    // :251-259
                            const double rb = (end_f == f0) ? 1.0 : (end_f > f0 ? end_f / f0 : f0 / end_f);
                            if (floor (ra) == floor (rb))                                         // :232
                            {
                                const double ratio = ra < rb ? ra : rb;
                                inharm += (ratio - floor (ratio)) * (mg / hsum);                  // :235-239
                            }
                        }
                    }
                }
            }
            *reinterpret_cast<float4*> (&workf[b0])     = make_float4 (nm[0], nm[1], nm[2], nm[3]);
            *reinterpret_cast<float4*> (&workf[b0 + 4]) = make_float4 (nm[4], nm[5], nm[6], nm[7]);
            double s3[3] = { sum_normed, inharm, npeaks };
            warp_sum<3> (s3);
            const float wpk = warp_minf (pkm);
            if (lane == 0) { sm.red[1][1][warp] = s3[0]; sm.red[1][2][warp] = s3[1]; sm.red[1][3][warp] = s3[2]; sm.fmins[0][warp] = wpk; }
        }
        __syncthreads();                                                                          // (22)
        if (warp == 0)
        {
            // calculateHarmonicEnergyCharacteristics (:147-198), numLower = 15, numHarmonics = 3 (:94)
            float term = 0.0f; int hb = -1;
            if (lane < 18 && ! hsilent)
            {
                const double fr = (lane < 15) ? f0 / ldexp (1.0, lane + 1) : f0 * (double) (lane - 14);
                hb = (int) floor (fr / frpb);
                if (hb >= 0 && hb < M)
                {
                    const int st = hb - 2 >= 0 ? hb - 2 : 0;
                    const int en = hb + 2 < M ? hb + 2 : M;
                    float mx = workf[hb];                                                         // :200-210
                    for (int b = st; b < en; ++b) mx = fmaxf (mx, workf[b]);
                    term = mx;
                }
            }
            if (lane < 18) { sm.her_terms[lane] = term; sm.her_bins[lane] = hb; }
            __syncwarp();
            if (lane == 0)
            {
                double sum_normed = 0.0, inharm = 0.0, npeaks = 0.0; float pkm = 1.0f;
                #pragma unroll
                for (int w = 0; w < NW; ++w) { sum_normed += sm.red[1][1][w]; inharm += sm.red[1][2][w]; npeaks += sm.red[1][3][w]; pkm = fminf (pkm, sm.fmins[0][w]); }
                float o_her = 0.0f, o_oer = 0.0f, o_inh = 0.0f;
                if (! hsilent)
                {
                    double score = 0.0, even = 0.0, odd = 0.0;
                    for (int l = 0; l < 15; ++l)
                    {
                        if (sm.her_bins[l] == f0_bin) continue;                                   // :163-164
                        score += (double) sm.her_terms[l];
                    }
                    for (int h = 1; h <= 3; ++h)
                    {
                        if (sm.her_bins[14 + h] >= M) break;                                      // :174-175
                        const double bm = (double) sm.her_terms[14 + h];
                        if (h % 2 == 0) even += bm; else odd += bm;
                        score += bm;
                    }
                    double her = score / sum_normed;
                    her = her > 1.0 ? 1.0 : her; her = her < 0.0 ? 0.0 : her;
                    double oer = 1.0;
                    if (odd > 0.0) oer = even / odd;
                    oer = oer > 1.0 ? 1.0 : oer; oer = oer < 0.0 ? 0.0 : oer;
                    o_her = (float) log10 ((double) (float) her * 9.0 + 1.0);                     // :101-103
                    o_oer = (float) log10 ((double) (float) oer * 9.0 + 1.0);
                    o_inh = (float) log10 (inharm * 9.0 + 1.0);
                }
                else { npeaks = 0.0; }
                out[FX_HER] = o_her;
                out[FX_OER] = o_her;                                                              // RealTimeAnalyser.h:171 stores HER in the OER slot
                out[FX_INHARM] = o_inh;
                if (dg)
                {
                    dg[FX_DIAG_TRUE_OER] = o_oer;
                    dg[FX_DIAG_NUM_PEAKS] = (float) npeaks;
                    dg[FX_DIAG_PEAK_MARGIN] = pkm;
                    dg[FX_DIAG_GATE_MARGIN] = fminf (dg[FX_DIAG_GATE_MARGIN], relmargin (hsum, 0.005));
                }
            }
        }
        __syncthreads();                                                                          // (23) work array and reduction slots are reused by the next frame
    }

    // ---- chunk epilogue ----------------------------------------------------------------------------------
    if (t == 0) p.first_idx[track * p.n_chunks + chunk] = first_nonsilent;
    if (have_prev)
    {
        float* ls = p.last_spec + (track * p.n_chunks + chunk) * (long) M;
        for (int i = t; i < M; i += T) ls[i] = sm.specB[cur ^ 1][i];
    }
    if (f_end == p.n_frames && p.tail_out != nullptr)
    {
        // the newest N - H samples of the stream become the next call's overlap
        const long a_end = (p.first_hop + f_end) * (long) H;
        float* to = p.tail_out + track * (long) (N - H);
        for (int i = t; i < N - H; i += T) to[i] = sm.ring[(int) ((a_end - (N - H) + i) & (N - 1))];
    }
}

// ---------------------------------------------------------------------------------------------------------
template <int R1> static cudaError_t launch_t (long n_tracks, const AnalyseParams& p, cudaStream_t stream)
{
    const long grid = n_tracks * p.n_chunks;
    if (grid <= 0) return cudaSuccess;
    k_analyse<R1><<<(unsigned) grid, 16 * R1, sizeof (Smem<R1>), stream>>> (p);
    return cudaGetLastError();
}

size_t analyse_smem_bytes (int window)
{
    switch (window)
    {
        case 1024: return sizeof (Smem<4>);
        case 2048: return sizeof (Smem<8>);
        case 4096: return sizeof (Smem<16>);
        default:   return 0;
    }
}

cudaError_t configure_analyse (int window)
{
    switch (window)
    {
        case 1024: return cudaFuncSetAttribute (k_analyse<4>,  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof (Smem<4>));
        case 2048: return cudaFuncSetAttribute (k_analyse<8>,  cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof (Smem<8>));
        case 4096: return cudaFuncSetAttribute (k_analyse<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof (Smem<16>));
        default:   return cudaErrorInvalidValue;
    }
}

cudaError_t launch_analyse (int window, long n_tracks, const AnalyseParams& p, cudaStream_t stream)
{
    switch (window)
    {
        case 1024: return launch_t<4>  (n_tracks, p, stream);
        case 2048: return launch_t<8>  (n_tracks, p, stream);
        case 4096: return launch_t<16> (n_tracks, p, stream);
        default:   return cudaErrorInvalidValue;
    }
}

} // namespace fx
