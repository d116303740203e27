// fx_pcm.cu -- K0 k_pcm_decode: file ingest.  Interleaved integer / float PCM as it lies in a WAV (little endian) or
// AIFF (big endian) data chunk -> the track-major fp32 samples K1 consumes, one channel per track.
//
// Replaces the sample-format conversion the reference reaches through AudioFilePlayer::loadFileIntoTransport
// (Source/AudioFilePlayer.h:41-60: AudioFormatManager::createReaderFor -> AudioFormatReaderSource -> AudioTransportSource)
// and the channel pick of AudioDataCollector (Source/AudioDataCollector.h:42-43,119).  JUCE's readers left-justify integer
// samples into int32 and convert with (float) int32 * (1.0f / 0x7fffffff), which is 2^-31 in fp32 [JUCE-recall]; 8-bit WAV
// is offset binary.  The conversion is exact for <= 24-bit samples, so host PCM can cross PCIe at its file width
// (2 bytes per sample for 16-bit audio instead of 4) and still give bit-identical fp32 input to the analysis.
//
// HBM-bound byte work: one pass, every input byte read once, every output float written once; the mono 16-bit case
// moves 16 bytes in / 32 bytes out per thread with vector accesses, the general case 4 samples per thread.
#include "fx_kernels.cuh"
#include <stdint.h>

namespace fx {

__device__ __forceinline__ float fixed_to_float (uint32_t left_justified)
{
    return __fmul_rn (__int2float_rn ((int) left_justified), 4.656612873077392578125e-10f);      // 1.0f / 0x7fffffff == 2^-31
}

__device__ __forceinline__ float decode_sample (const uint8_t* q, int format)
{
    switch (format)
    {
        case FX_PCM_U8:    return fixed_to_float ((uint32_t) (q[0] ^ 0x80u) << 24);
        case FX_PCM_S8:    return fixed_to_float ((uint32_t) q[0] << 24);
        case FX_PCM_S16LE: return fixed_to_float (((uint32_t) q[0] << 16) | ((uint32_t) q[1] << 24));
        case FX_PCM_S16BE: return fixed_to_float (((uint32_t) q[1] << 16) | ((uint32_t) q[0] << 24));
        case FX_PCM_S24LE: return fixed_to_float (((uint32_t) q[0] << 8) | ((uint32_t) q[1] << 16) | ((uint32_t) q[2] << 24));
        case FX_PCM_S24BE: return fixed_to_float (((uint32_t) q[2] << 8) | ((uint32_t) q[1] << 16) | ((uint32_t) q[0] << 24));
        case FX_PCM_S32LE: return fixed_to_float ((uint32_t) q[0] | ((uint32_t) q[1] << 8) | ((uint32_t) q[2] << 16) | ((uint32_t) q[3] << 24));
        case FX_PCM_S32BE: return fixed_to_float ((uint32_t) q[3] | ((uint32_t) q[2] << 8) | ((uint32_t) q[1] << 16) | ((uint32_t) q[0] << 24));
        case FX_PCM_F32LE: return __uint_as_float ((uint32_t) q[0] | ((uint32_t) q[1] << 8) | ((uint32_t) q[2] << 16) | ((uint32_t) q[3] << 24));
        default:           return __uint_as_float ((uint32_t) q[3] | ((uint32_t) q[2] << 8) | ((uint32_t) q[1] << 16) | ((uint32_t) q[0] << 24));
    }
}

// mono little-endian 16-bit, rows 16-byte aligned on both sides: 8 samples per thread, LDG.128 in, 2 x STG.128 out
__global__ void __launch_bounds__ (256) k_pcm_decode_s16 (const PcmParams p)
{
    const long octs = p.n_samples >> 3;
    const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n_tracks * octs) return;
    const long track = idx / octs, o = idx % octs;
    const uint4 v = __ldcs (reinterpret_cast<const uint4*> (p.pcm + track * p.track_stride_bytes) + o);       // streamed: read once
    float4* dst = reinterpret_cast<float4*> (p.audio + track * p.audio_stride) + 2 * o;
    dst[0] = make_float4 (fixed_to_float (v.x << 16), fixed_to_float (v.x & 0xffff0000u), fixed_to_float (v.y << 16), fixed_to_float (v.y & 0xffff0000u));
    dst[1] = make_float4 (fixed_to_float (v.z << 16), fixed_to_float (v.z & 0xffff0000u), fixed_to_float (v.w << 16), fixed_to_float (v.w & 0xffff0000u));
}

// any format / channel count: 4 consecutive samples of one track per thread
__global__ void __launch_bounds__ (256) k_pcm_decode (const PcmParams p, long first_sample)
{
    const long n = p.n_samples - first_sample;
    const long quads = (n + 3) >> 2;
    const long idx = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n_tracks * quads) return;
    const long track = idx / quads, i0 = first_sample + ((idx % quads) << 2);
    const int bps = pcm_bytes_per_sample (p.format);
    const long frame_bytes = (long) p.n_channels * bps;
    const int ch = p.channel >= 0 ? p.channel : (int) ((p.first_track + track) % p.n_channels);
    const uint8_t* src = p.pcm + track * p.track_stride_bytes + (long) ch * bps;
    float* dst = p.audio + track * p.audio_stride;
    #pragma unroll
    for (int u = 0; u < 4; ++u)
    {
        const long i = i0 + u;
        if (i < p.n_samples) dst[i] = decode_sample (src + i * frame_bytes, p.format);
    }
}

cudaError_t launch_pcm_decode (const PcmParams& p, cudaStream_t stream)
{
    if (p.n_tracks <= 0 || p.n_samples <= 0) return cudaSuccess;
    long done = 0;
    const bool fast = p.format == FX_PCM_S16LE && p.n_channels == 1 && (p.track_stride_bytes & 15) == 0 && (p.audio_stride & 3) == 0
                      && ((uintptr_t) p.pcm & 15) == 0 && ((uintptr_t) p.audio & 15) == 0;
    if (fast && (p.n_samples >> 3) > 0)
    {
        const long total = p.n_tracks * (p.n_samples >> 3);
        k_pcm_decode_s16<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>> (p);
        done = (p.n_samples >> 3) << 3;
    }
    if (done < p.n_samples)
    {
        const long total = p.n_tracks * ((p.n_samples - done + 3) >> 2);
        k_pcm_decode<<<(unsigned) ((total + 255) / 256), 256, 0, stream>>> (p, done);
    }
    return cudaGetLastError();
}

int pcm_launch_count (const PcmParams& p)
{
    if (p.n_tracks <= 0 || p.n_samples <= 0) return 0;
    const bool fast = p.format == FX_PCM_S16LE && p.n_channels == 1 && (p.track_stride_bytes & 15) == 0 && (p.audio_stride & 3) == 0
                      && ((uintptr_t) p.pcm & 15) == 0 && ((uintptr_t) p.audio & 15) == 0 && (p.n_samples >> 3) > 0;
    return fast ? (((p.n_samples & 7) != 0) ? 2 : 1) : 1;
}

} // namespace fx
