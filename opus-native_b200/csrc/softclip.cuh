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
    mem_all[(size_t)row * channels + c] = a;
}

}  // namespace opn
