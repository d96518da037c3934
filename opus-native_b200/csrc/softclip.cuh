// softclip.cuh -- pcm_soft_clip (src/lib.rs:526-632) for many independent buffers.
// The algorithm is a data-dependent serial scan per channel (zero-crossing search, peak search),
// so one thread owns one (buffer, channel); it only runs on the generic Decoder::decode<S> path.
// Restated as written in the reference, including its search-loop quirk (lib.rs:556-566 leaves
// pos == frame_size-1, so the last region of each channel always goes through the non-linearity).
#pragma once
#include "opn_device.cuh"
#include "opn_internal.h"

namespace opn {

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// One channel of pcm_soft_clip (lib.rs:540-630) on an interleaved buffer of `ch` channels; `a` is the
// channel's declip memory (in/out).  pcm may point to global or shared memory.
__device__ __forceinline__ void soft_clip_channel(float *pcm, int frame_size, int ch, int c, float &a)
{
    for (int i = 0; i < frame_size; i++) {
        const int off = c + i * ch;
        if (pcm[off] * a >= 0.0f) break;
        pcm[off] += a * pcm[off] * pcm[off];
    }
    int curr = 0;
    const float x0 = pcm[c];
    for (;;) {
        int pos = 0;
        for (int i = curr; i < frame_size; i++) {
            pos = i;
            if (pcm[c + pos * ch] > 1.0f || pcm[c + pos * ch] < -1.0f) break;
        }
        if (pos == frame_size) {
            a = 0.0f;
            break;
        }
        int peak_pos = pos, start = pos, end = pos;
        float maxval = fabsf(pcm[c + pos * ch]);
        while (start > 0 && pcm[c + pos * ch] * pcm[c + (start - 1) * ch] >= 0.0f) start -= 1;
        while (end < frame_size && pcm[c + pos * ch] * pcm[c + end * ch] >= 0.0f) {
            if (fabsf(pcm[c + end * ch]) > maxval) {
                maxval = fabsf(pcm[c + end * ch]);
                peak_pos = end;
            }
            end += 1;
        }
        const bool special = start == 0 && (pcm[c + pos * ch] * pcm[c]) >= 0.0f;
        a = (maxval - 1.0f) / (maxval * maxval);
        a += a * 2.4e-7f;
        if (pcm[c + pos * ch] > 0.0f) a = -a;
        for (int i = start; i < end; i++) {
            const int off = c + i * ch;
            pcm[off] += a * pcm[off] * pcm[off];
        }
        if (special && peak_pos >= 2) {
            float offset = x0 - pcm[c];
            const float delta = offset / (float)peak_pos;
            for (int i = curr; i < peak_pos; i++) {
                const int off = c + i * ch;
                offset -= delta;
                pcm[off] += offset;
                pcm[off] = clampf(pcm[off], -1.0f, 1.0f);
            }
        }
        curr = end;
        if (curr == frame_size) break;
    }
}

__global__ void k_op_soft_clip(float *__restrict__ pcm_all, size_t row_stride, size_t row_len, int channels, uint32_t n_rows,
                               float *__restrict__ mem_all)
{
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_rows * (uint32_t)channels) return;
    const uint32_t row = gid / (uint32_t)channels;
    const int c = (int)(gid - row * (uint32_t)channels);
    float *pcm = pcm_all + (size_t)row * row_stride;
    const int ch = channels;
    const int frame_size = (int)(row_len / (size_t)channels);
    if (row_len == 0) return;
    // saturate to +-2 (lib.rs:538); samples past frame_size*channels are clamped by channel 0's thread
    for (int i = 0; i < frame_size; i++) pcm[c + i * ch] = clampf(pcm[c + i * ch], -2.0f, 2.0f);
    if (c == 0)
        for (size_t i = (size_t)frame_size * ch; i < row_len; i++) pcm[i] = clampf(pcm[i], -2.0f, 2.0f);
    float a = mem_all[(size_t)row * channels + c];
    soft_clip_channel(pcm, frame_size, ch, c, a);
    mem_all[(size_t)row * channels + c] = a;
}

// Sample::from_f32 (lib.rs:63-107) as the crate writes it: scale (and offset), clamp, then a Rust `as` cast, i.e.
// truncation toward zero that saturates and maps NaN to 0 -- exactly what cvt.rzi does.  f32::clamp keeps NaN.
template <typename T> __device__ __forceinline__ T sample_from_f32(float f);
template <> __device__ __forceinline__ float sample_from_f32<float>(float f) { return f; }
template <> __device__ __forceinline__ double sample_from_f32<double>(float f) { return (double)f; }
template <> __device__ __forceinline__ int16_t sample_from_f32<int16_t>(float f)
{
    f = clampf(f * 32768.0f, -32768.0f, 32767.0f);
    return (int16_t)__float2int_rz(f);
}
template <> __device__ __forceinline__ int32_t sample_from_f32<int32_t>(float f)
{
    f = clampf(f * 2147483648.0f, -2147483648.0f, 2147483648.0f);  // the crate's 2_147_483_647.0 is 2^31 as an f32
    return __float2int_rz(f);                                        // saturates at i32::MAX
}
template <> __device__ __forceinline__ uint16_t sample_from_f32<uint16_t>(float f)
{
    f = clampf(f * 32768.0f + 32768.0f, 0.0f, 32768.0f);
    return (uint16_t)__float2uint_rz(f);
}
template <> __device__ __forceinline__ uint32_t sample_from_f32<uint32_t>(float f)
{
    f = clampf(f * 2147483648.0f + 2147483648.0f, 0.0f, 2147483648.0f);
    return __float2uint_rz(f);
}

// Decoder::decode::<S> epilogue for a batch (decoder.rs:177-189): one warp per stream stages the stream's
// interleaved float row in shared memory, lanes 0..C-1 run pcm_soft_clip on their channel (a serial scan),
// then the warp converts the row with Sample::from_f32 and stores 8 samples per lane and step.
// clip_len[row] is the slice length the reference hands to pcm_soft_clip (its sample_count: the per-channel
// count, decoder.rs:415-419 omits the "x channels"); <= 0 skips the clip (error rows are all zeros).
template <typename T>
__global__ void __launch_bounds__(32)
k_softclip_convert(const float *__restrict__ dense, size_t dense_stride, const int32_t *__restrict__ clip_len, int channels,
                   uint32_t row_floats, uint32_t first_row, float *__restrict__ mem_all, T *__restrict__ out, size_t out_stride)
{
    extern __shared__ __align__(16) float s_row[];
    const uint32_t row = first_row + blockIdx.x, lane = threadIdx.x;
    const float4 *src = reinterpret_cast<const float4 *>(dense + (size_t)row * dense_stride);
    for (uint32_t i = lane; i < row_floats / 4u; i += 32u) reinterpret_cast<float4 *>(s_row)[i] = src[i];
    __syncwarp();
    const int rl = clip_len[row];
    if (rl > 0) {
        const int frame_size = rl / channels;
        for (int i = (int)lane; i < rl; i += 32) s_row[i] = clampf(s_row[i], -2.0f, 2.0f);  // lib.rs:538
        __syncwarp();
        if ((int)lane < channels && frame_size > 0) {
            float a = mem_all[(size_t)row * 2 + lane];
            soft_clip_channel(s_row, frame_size, channels, (int)lane, a);
            mem_all[(size_t)row * 2 + lane] = a;
        }
        __syncwarp();
    }
    constexpr uint32_t NV = sizeof(T) * 8u / 16u;  // 16-byte stores per 8 samples
    uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)row * out_stride);
    for (uint32_t i = lane; i < row_floats / 8u; i += 32u) {
        union {
            T v[8];
            uint4 q[NV];
        } pk;
#pragma unroll
        for (int e = 0; e < 8; e++) pk.v[e] = sample_from_f32<T>(s_row[8u * i + e]);
#pragma unroll
        for (uint32_t q = 0; q < NV; q++) dst[i * NV + q] = pk.q[q];
    }
}

// smooth_fade_into_in1 / smooth_fade_into_in2, src/decoder.rs:833-865: the cross-fade of a mode transition,
// out = w^2 * in2 + (1 - w^2) * in1 over the first `overlap` samples of interleaved rows, w = WINDOW[i * 48000 / fs].
// `out` may be in1 or in2 (the crate's two in-place forms).  One thread per sample of a row, one row per blockIdx.y.
__global__ void __launch_bounds__(128) k_op_smooth_fade(const float *in1, const float *in2, float *out, size_t row_stride, int overlap,
                                                         int channels, int inc)
{
    const int j = blockIdx.x * 128 + threadIdx.x;  // interleaved index: sample i = j / channels
    if (j >= overlap * channels) return;
    const size_t at = (size_t)blockIdx.y * row_stride + (size_t)j;
    const float wv = g_tab.window[(j / channels) * inc];
    const float w = wv * wv;
    out[at] = (w * in2[at]) + ((1.0f - w) * in1[at]);
}

// The cross-fade decode_frame applies when a stream switches from CELT to SILK (src/decoder.rs:519-543, 674-676, 765-788): the old
// decoder conceals 5 ms into a transition buffer (here: the stream's 240 C floats behind the n_rows dense rows); the first 2.5 ms of
// the new frame are replaced by it, the next 2.5 ms are smooth_fade_into_in2(transition, samples) = w^2 samples + (1 - w^2) transition.
// One CTA per listed stream, f2_5 = 120 samples at 48 kHz.
__global__ void __launch_bounds__(128) k_transition_fade(float *dense, size_t dense_stride, const uint32_t *streams, uint32_t n_rows, int channels)
{
    const uint32_t stream = streams[blockIdx.x];
    float *row = dense + (size_t)stream * dense_stride;
    const float *tr = dense + (size_t)n_rows * dense_stride + (size_t)stream * 240 * channels;  // the tail area behind all rows
    const int n = 120 * channels;
    for (int j = (int)threadIdx.x; j < n; j += 128) {
        const float wv = g_tab.window[j / channels];
        const float w = wv * wv;
        row[n + j] = (w * row[n + j]) + ((1.0f - w) * tr[n + j]);
        row[j] = tr[j];
    }
}

}  // namespace opn
