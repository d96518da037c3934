// imdct.cuh -- kernel 1: IMDCT + TDAC overlap-add + comb post-filter + interleaved PCM store.
//
// Device mirror of Mdct::backward (src/celt/mdct.rs:159-260), KissFft::process and its
// butterflies (src/celt/kiss_fft.rs:24-243) and comb_filter_inplace
// (src/celt/comb_filter/mod.rs:130-193, scalar kernel fallback.rs:32-53).
//
// Arithmetic contract: every sum and product below is evaluated in the same order and with the
// same single roundings as the reference; this translation unit is compiled with -fmad=false so
// nvcc never contracts a*b+c (Rust does not either).  Data placement (registers / shared
// memory / which thread owns which butterfly) is free and is what the B200 design changes.
//
// One CTA decodes one stream: 128 threads per channel, both channels of a stereo stream in the
// same CTA so the epilogue can store interleaved float4 PCM.
#pragma once
#include "opn_device.cuh"
#include "opn_internal.h"

namespace opn {

constexpr int SY_FLOATS = 2048;     // HIST_CAP + 960 + 60, rounded up
constexpr int SF_CPLX = 480;

__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// src/math.rs:115-124
__device__ __forceinline__ float2 c_mul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 c_scale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 ldtw(int idx) { return __ldg(&g_tab.twiddles[idx]); }

#define OPN_FRAC_1_SQRT_2 0.70710678118654752440f

// kiss_fft.rs:129-147 (m == 1)
__device__ __forceinline__ void bfly4_m1(float2 *d)
{
    float2 s0 = c_sub(d[0], d[2]);
    float2 s1 = c_add(d[1], d[3]);
    float2 d0 = c_add(d[0], d[2]);
    d[2] = c_sub(d0, s1);
    d[0] = c_add(d0, s1);
    s1 = c_sub(d[1], d[3]);
    d[1] = make_float2(s0.x + s1.y, s0.y - s1.x);
    d[3] = make_float2(s0.x - s1.y, s0.y + s1.x);
}

// kiss_fft.rs:55-87: pair j (0..3) of one group of 8
__device__ __forceinline__ void bfly2_pair(float2 *d, int j)
{
    float2 x = d[4 + j], t;
    if (j == 0) t = x;
    else if (j == 1) t = make_float2((x.x + x.y) * OPN_FRAC_1_SQRT_2, (x.y - x.x) * OPN_FRAC_1_SQRT_2);
    else if (j == 2) t = make_float2(x.y, -x.x);
    else t = make_float2((x.y - x.x) * OPN_FRAC_1_SQRT_2, (-(x.y + x.x)) * OPN_FRAC_1_SQRT_2);
    float2 a = d[j];
    d[4 + j] = c_sub(a, t);
    d[j] = c_add(a, t);
}

// kiss_fft.rs:148-187
__device__ __forceinline__ void bfly4(float2 *d, int m, int u, int stride)
{
    float2 s0 = c_mul(d[m], ldtw(u * stride));
    float2 s1 = c_mul(d[2 * m], ldtw(2 * u * stride));
    float2 s2 = c_mul(d[3 * m], ldtw(3 * u * stride));
    float2 s5 = c_sub(d[0], s1);
    float2 d0 = c_add(d[0], s1);
    float2 s3 = c_add(s0, s2);
    float2 s4 = c_sub(s0, s2);
    d[2 * m] = c_sub(d0, s3);
    d[0] = c_add(d0, s3);
    d[m] = make_float2(s5.x + s4.y, s5.y - s4.x);
    d[3 * m] = make_float2(s5.x - s4.y, s5.y + s4.x);
}

// kiss_fft.rs:89-127
__device__ __forceinline__ void bfly3(float2 *d, int m, int u, int stride)
{
    const float2 epi3 = ldtw(stride * m);
    float2 s1 = c_mul(d[m], ldtw(u * stride));
    float2 s2 = c_mul(d[2 * m], ldtw(2 * u * stride));
    float2 s3 = c_add(s1, s2);
    float2 s0 = c_sub(s1, s2);
    float2 dm = c_sub(d[0], c_scale(s3, 0.5f));
    s0 = c_scale(s0, epi3.y);
    d[0] = c_add(d[0], s3);
    d[2 * m] = make_float2(dm.x + s0.y, dm.y - s0.x);
    d[m] = make_float2(dm.x - s0.y, dm.y + s0.x);
}

// kiss_fft.rs:190-243
__device__ __forceinline__ void bfly5(float2 *d, int m, int u, int stride)
{
    const float2 ya = ldtw(stride * m), yb = ldtw(stride * 2 * m);
    float2 s0 = d[0];
    float2 s1 = c_mul(d[m], ldtw(u * stride));
    float2 s2 = c_mul(d[2 * m], ldtw(2 * u * stride));
    float2 s3 = c_mul(d[3 * m], ldtw(3 * u * stride));
    float2 s4 = c_mul(d[4 * m], ldtw(4 * u * stride));
    float2 s7 = c_add(s1, s4), s10 = c_sub(s1, s4);
    float2 s8 = c_add(s2, s3), s9 = c_sub(s2, s3);
    d[0] = c_add(s0, c_add(s7, s8));
    float2 s5, s6, s11, s12;
    s5.x = s0.x + (s7.x * ya.x + s8.x * yb.x);
    s5.y = s0.y + (s7.y * ya.x + s8.y * yb.x);
    s6.x = s10.y * ya.y + s9.y * yb.y;
    s6.y = -(s10.x * ya.y + s9.x * yb.y);
    d[m] = c_sub(s5, s6);
    d[4 * m] = c_add(s5, s6);
    s11.x = s0.x + (s7.x * yb.x + s8.x * ya.x);
    s11.y = s0.y + (s7.y * yb.x + s8.y * ya.x);
    s12.x = s9.y * ya.y - s10.y * yb.y;
    s12.y = s10.x * yb.y - s9.x * ya.y;
    d[2 * m] = c_add(s11, s12);
    d[3 * m] = c_sub(s11, s12);
}

// Stage lists in execution order (kiss_fft.rs:38-52 walks `factors` last-to-first):
// nfft 480 = 4(m1) 2(m4) 4(m8) 3(m32) 5(m96); 240 = 4 4(m4) 3(m16) 5(m48);
// 120 = 4 2(m4) 3(m8) 5(m24); 60 = 4 3(m4) 5(m12).
__device__ __forceinline__ int fft_num_stages(int shift) { return shift == 0 ? 5 : shift == 3 ? 3 : 4; }
__device__ __forceinline__ void fft_stage(int shift, int s, int &radix, int &m)
{
    const int r0[5] = {4, 2, 4, 3, 5}, m0[5] = {1, 4, 8, 32, 96};
    const int r1[4] = {4, 4, 3, 5}, m1[4] = {1, 4, 16, 48};
    const int r2[4] = {4, 2, 3, 5}, m2[4] = {1, 4, 8, 24};
    const int r3[3] = {4, 3, 5}, m3[3] = {1, 4, 12};
    if (shift == 0) { radix = r0[s]; m = m0[s]; }
    else if (shift == 1) { radix = r1[s]; m = m1[s]; }
    else if (shift == 2) { radix = r2[s]; m = m2[s]; }
    else { radix = r3[s]; m = m3[s]; }
}

// KissFft::process on `nblk` consecutive transforms of size nfft = 480 >> shift held in d
// (already in bit-reversed order).  `nt` threads cooperate; ends with a __syncthreads().
__device__ __forceinline__ void fft_process(float2 *d, int shift, int nblk, int tid, int nt)
{
    const int nfft = 480 >> shift;
    const int ns = fft_num_stages(shift);
    for (int s = 0; s < ns; s++) {
        int radix, m;
        fft_stage(shift, s, radix, m);
        const int mm = radix * m;             // span of one group
        const int groups = nfft / mm;         // "n" in the reference
        const int stride = groups << shift;   // twiddle stride (kiss_fft.rs:41)
        if (radix == 4 && m == 1) {
            for (int b = tid; b < nblk * groups; b += nt) bfly4_m1(d + 4 * b);
        } else if (radix == 2) {
            for (int b = tid; b < nblk * groups * 4; b += nt) bfly2_pair(d + 8 * (b >> 2), b & 3);
        } else {
            const int per = groups * m;  // butterflies per transform
            for (int b = tid; b < nblk * per; b += nt) {
                const int blk = b / per, r = b - blk * per;
                const int g = r / m, u = r - g * m;
                float2 *p = d + blk * nfft + g * mm + u;
                if (radix == 4) bfly4(p, m, u, stride);
                else if (radix == 3) bfly3(p, m, u, stride);
                else bfly5(p, m, u, stride);
            }
        }
        __syncthreads();
    }
}

// Mdct::backward core for `nblk` interleaved blocks (mdct.rs:159-238).  x: coefficient row in
// shared memory (block b reads x[b + nblk*k]); sF: FFT scratch; out: output row, block b writes
// out[n2*b + 60 + j], j < n2.  Caller fills out[0..60) with the previous tail BEFORE
// calling tdac_mirror.  All threads of the channel group call this (contains __syncthreads).
__device__ __forceinline__ void imdct_core(const float *x, float2 *sF, float *out, int shift, int nblk, int tid, int nt)
{
    const int n = 1920 >> shift, n2 = n >> 1, n4 = n >> 2;
    int trigp = 0;
    for (int s = 0, nn = 1920; s < shift; s++) { nn >>= 1; trigp += nn; }
    const float *trig = g_tab.trig + trigp;
    const uint16_t *bitrev = g_tab.bitrev[shift];
    // pre-rotation (mdct.rs:184-200)
    for (int w = tid; w < nblk * n4; w += nt) {
        const int blk = w / n4, i = w - blk * n4;
        const float x0 = x[blk + nblk * (2 * i)];
        const float x1 = x[blk + nblk * (n2 - 1 - 2 * i)];
        const float t0 = __ldg(trig + i), t1 = __ldg(trig + n4 + i);
        const float re = (x1 * t0) + (x0 * t1);
        const float im = (x0 * t0) - (x1 * t1);
        sF[blk * n4 + __ldg(bitrev + i)] = make_float2(im, re);
    }
    __syncthreads();
    fft_process(sF, shift, nblk, tid, nt);
    // post-rotation and de-shuffle (mdct.rs:205-238)
    for (int w = tid; w < nblk * n4; w += nt) {
        const int blk = w / n4, i = w - blk * n4;
        float *o = out + n2 * blk + 60;
        const float2 c = sF[blk * n4 + i];
        const float2 c2 = sF[blk * n4 + n4 - 1 - i];
        const float e = (c.y * __ldg(trig + i)) + (c.x * __ldg(trig + n4 + i));
        const float od = (c2.y * __ldg(trig + n2 - 1 - i)) - (c2.x * __ldg(trig + n4 - 1 - i));
        *reinterpret_cast<float2 *>(o + 2 * i) = make_float2(e, od);
    }
    __syncthreads();
}

// TDAC mirror for `nblk` blocks (mdct.rs:241-259), overlap 120, window = mode::WINDOW.
__device__ __forceinline__ void tdac_mirror(float *out, int n2, int nblk, int tid, int nt)
{
    for (int w = tid; w < nblk * 60; w += nt) {
        const int blk = w / 60, i = w - blk * 60;
        float *o = out + n2 * blk;
        const float x0 = o[119 - i], x1 = o[i];
        const float w0 = __ldg(&g_tab.window[i]), w1 = __ldg(&g_tab.window[119 - i]);
        o[i] = (w1 * x1) - (w0 * x0);
        o[119 - i] = (w0 * x1) + (w1 * x0);
    }
    __syncthreads();
}

// comb_filter_inplace (comb_filter/mod.rs:130-193) on y[0..n) with history y[-T-2..0) in shared
// memory.  The filter is recursive (it reads samples it wrote T-2 or more positions earlier), so
// the `nt` threads sweep the frame in chunks of W = min(period) - 2 samples: inside a chunk every
// tap lies before the chunk start.  Uniform control flow per CTA (parameters are per stream).
__device__ __forceinline__ void comb_inplace_smem(float *y, int t0, int t1, int n, float g0, float g1, int tap0, int tap1,
                                                  int overlap, int tid, int nt)
{
    if (g0 == 0.0f && g1 == 0.0f) return;
    t0 = max(t0, 15);
    t1 = max(t1, 15);
    const float g00 = g0 * g_tab.comb_gains[tap0 * 3], g01 = g0 * g_tab.comb_gains[tap0 * 3 + 1],
                g02 = g0 * g_tab.comb_gains[tap0 * 3 + 2];
    const float g10 = g1 * g_tab.comb_gains[tap1 * 3], g11 = g1 * g_tab.comb_gains[tap1 * 3 + 1],
                g12 = g1 * g_tab.comb_gains[tap1 * 3 + 2];
    if (fabsf(g0 - g1) < 1.1920929e-7f && t0 == t1 && tap0 == tap1) overlap = 0;
    // cross-fade part (mod.rs:162-179)
    {
        const int W = min(min(t0, t1) - 2, nt);
        for (int base = 0; base < overlap; base += W) {
            const int i = base + tid;
            if (tid < W && i < overlap) {
                const float f = __ldg(&g_tab.window_sq[i]);
                const float x0 = y[i - t1 + 2], x1 = y[i - t1 + 1], x2 = y[i - t1], x3 = y[i - t1 - 1], x4 = y[i - t1 - 2];
                float acc = y[i];
                acc = acc + (((1.0f - f) * g00) * y[i - t0]);
                acc = acc + (((1.0f - f) * g01) * (y[i - t0 + 1] + y[i - t0 - 1]));
                acc = acc + (((1.0f - f) * g02) * (y[i - t0 + 2] + y[i - t0 - 2]));
                acc = acc + ((f * g10) * x2);
                acc = acc + ((f * g11) * (x1 + x3));
                acc = acc + ((f * g12) * (x0 + x4));
                y[i] = acc;
            }
            __syncthreads();
        }
    }
    if (g1 == 0.0f) return;
    // constant part (fallback.rs:32-53)
    {
        const int W = min(t1 - 2, nt);
        for (int base = overlap; base < n; base += W) {
            const int i = base + tid;
            if (tid < W && i < n) {
                const float x0 = y[i - t1 + 2], x1 = y[i - t1 + 1], x2 = y[i - t1], x3 = y[i - t1 - 1], x4 = y[i - t1 - 2];
                y[i] = y[i] + (g10 * x2) + (g11 * (x1 + x3)) + (g12 * (x0 + x4));
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(2 * IM_TPC) k_imdct_post(ImdctArgs A)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int C = A.channels;
    const int ch = threadIdx.x / IM_TPC, tid = threadIdx.x - ch * IM_TPC;
    float *sY = reinterpret_cast<float *>(smem_raw) + ch * SY_FLOATS;
    float2 *sF = reinterpret_cast<float2 *>(reinterpret_cast<float *>(smem_raw) + C * SY_FLOATS) + ch * SF_CPLX;

    const uint32_t item = blockIdx.x;
    const uint32_t stream = A.stream_idx ? A.stream_idx[item] : item;
    const int32_t status = A.status[stream];
    const int lm = A.lm, nf = 120 << lm;
    if (status < 0) {  // rejected packet: state untouched (decoder.rs:397)
        if (threadIdx.x == 0 && A.result) A.result[stream] = status;
        return;
    }
    const opn_synth_side *side = A.side + stream;
    const bool lost = status == ITEM_LOST;
    const int transient = side->transient;
    const int nblk = transient ? (1 << lm) : 1;
    const int shift = transient ? 3 : 3 - lm;

    // post-filter parameters: previous frame -> this frame
    const PfState old = A.pf[stream];
    int t1 = old.period, tap1 = old.tapset;
    float g1 = old.gain;
    if (!lost) {
        const int on = side->postfilter;
        t1 = on ? side->period : 0;
        g1 = on ? 0.09375f * (float)(side->gain_idx + 1) : 0.0f;
        tap1 = on ? side->tapset : 0;
    }
    const bool comb_on = A.postfilter && (old.gain != 0.0f || g1 != 0.0f);
    const uint32_t pos = A.ring_pos[stream];
    float *ring = A.ring + (size_t)stream * RING_SAMPLES * C;

    // stage coefficients into the frame region of sY (float4, coalesced)
    {
        const float4 *src = reinterpret_cast<const float4 *>(A.coef + ((size_t)stream * C + ch) * nf);
        float4 *dst = reinterpret_cast<float4 *>(sY + HIST_CAP);
        for (int i = tid; i < nf / 4; i += IM_TPC) dst[i] = __ldg(src + i);
    }
    // comb history: the last max(T0,T1)+2 output samples of this channel, from the PCM ring
    if (comb_on) {
        const int need = max(max(old.period, t1), 15) + 2;
        for (int j = tid; j < need; j += IM_TPC) {
            uint32_t p = pos + RING_SAMPLES - 1u - (uint32_t)j;
            if (p >= RING_SAMPLES) p -= RING_SAMPLES;
            sY[HIST_CAP - 1 - j] = ring[(size_t)p * C + ch];
        }
    }
    __syncthreads();
    imdct_core(sY + HIST_CAP, sF, sY + HIST_CAP, shift, nblk, tid, IM_TPC);
    // NOTE: imdct_core read every coefficient (pre-rotation) before its first barrier and wrote
    // out[60 ..] only after the FFT, so in-place use of the frame region is safe.
    float *carry = A.carry + ((size_t)stream * C + ch) * 60;
    if (tid < 60) sY[HIST_CAP + tid] = carry[tid];
    __syncthreads();
    tdac_mirror(sY + HIST_CAP, 960 >> shift, nblk, tid, IM_TPC);
    if (tid < 60) carry[tid] = sY[HIST_CAP + nf + tid];
    if (comb_on)
        comb_inplace_smem(sY + HIST_CAP, old.period, t1, nf, old.gain, g1, old.tapset, tap1, 120, tid, IM_TPC);
    __syncthreads();

    // epilogue: interleaved PCM -> ring (history + device-resident output) and optional dense rows
    const float *s0 = reinterpret_cast<float *>(smem_raw) + HIST_CAP;
    const float *s1 = s0 + SY_FLOATS;
    float *dense = A.dense ? A.dense + (size_t)stream * A.dense_stride + (A.dense_off ? A.dense_off[item] : 0u) : nullptr;
    if (C == 2) {
        // two samples x two channels per float4
        for (int i = threadIdx.x; i < nf / 2; i += 2 * IM_TPC) {
            float4 v = make_float4(s0[2 * i], s1[2 * i], s0[2 * i + 1], s1[2 * i + 1]);
            uint32_t p = pos + 2u * (uint32_t)i;
            if (p >= RING_SAMPLES) p -= RING_SAMPLES;  // nf | RING_SAMPLES and pos % 120 == 0: pairs never straddle the wrap
            *reinterpret_cast<float4 *>(ring + (size_t)p * 2) = v;
            if (dense) {
                const float g = A.gain;
                if (g != 1.0f) { v.x *= g; v.y *= g; v.z *= g; v.w *= g; }
                *reinterpret_cast<float4 *>(dense + 4 * i) = v;
            }
        }
    } else {
        for (int i = threadIdx.x; i < nf / 4; i += IM_TPC) {
            float4 v = *reinterpret_cast<const float4 *>(s0 + 4 * i);
            uint32_t p = pos + 4u * (uint32_t)i;
            if (p >= RING_SAMPLES) p -= RING_SAMPLES;
            *reinterpret_cast<float4 *>(ring + p) = v;
            if (dense) {
                const float g = A.gain;
                if (g != 1.0f) { v.x *= g; v.y *= g; v.z *= g; v.w *= g; }
                *reinterpret_cast<float4 *>(dense + 4 * i) = v;
            }
        }
    }
    if (threadIdx.x == 0) {
        uint32_t np = pos + (uint32_t)nf;
        if (np >= RING_SAMPLES) np -= RING_SAMPLES;
        A.ring_pos[stream] = np;
        PfState nw;
        nw.period = t1;
        nw.tapset = tap1;
        nw.gain = g1;
        nw.pad = 0;
        A.pf[stream] = nw;
        if (A.result) A.result[stream] = nf;
        if (A.final_range) A.final_range[stream] = lost ? 0u : side->final_rng;
    }
}

// ---------------------------------------------------------------------------------------------
// Operator-level kernels (tests call these through opn_op_*).

// Mdct::backward on independent rows: one CTA of IM_TPC threads per row.
__global__ void __launch_bounds__(IM_TPC)
k_op_imdct(const float *__restrict__ input, size_t in_stride, float *__restrict__ output, size_t out_stride, int shift,
           int nblk)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *sX = reinterpret_cast<float *>(smem_raw);            // 960 coefficients
    float *sO = sX + 960;                                       // 1024 output
    float2 *sF = reinterpret_cast<float2 *>(sO + 1024);
    const int tid = threadIdx.x;
    const int n2 = (960 >> shift);
    const int nin = n2 * nblk, nout = n2 * nblk + 60;
    const float *in = input + (size_t)blockIdx.x * in_stride;
    float *out = output + (size_t)blockIdx.x * out_stride;
    for (int i = tid; i < nin; i += IM_TPC) sX[i] = in[i];
    for (int i = tid; i < 60; i += IM_TPC) sO[i] = out[i];
    __syncthreads();
    imdct_core(sX, sF, sO, shift, nblk, tid, IM_TPC);
    tdac_mirror(sO, n2, nblk, tid, IM_TPC);
    for (int i = tid; i < nout; i += IM_TPC) out[i] = sO[i];
}

// comb_filter_inplace on rows; history must be present in the row before y_offset.
__global__ void __launch_bounds__(IM_TPC)
k_op_comb_inplace(float *__restrict__ y, size_t row_stride, int y_offset, int n, const int32_t *__restrict__ params4,
                  const float *__restrict__ gains2, int overlap)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *s = reinterpret_cast<float *>(smem_raw);
    const int tid = threadIdx.x;
    float *row = y + (size_t)blockIdx.x * row_stride;
    const int t0 = params4[4 * blockIdx.x], t1 = params4[4 * blockIdx.x + 1];
    const int tap0 = params4[4 * blockIdx.x + 2], tap1 = params4[4 * blockIdx.x + 3];
    const float g0 = gains2[2 * blockIdx.x], g1 = gains2[2 * blockIdx.x + 1];
    const int hist = min(y_offset, HIST_CAP + 2);
    for (int i = tid; i < hist + n; i += IM_TPC) s[i] = row[y_offset - hist + i];
    __syncthreads();
    comb_inplace_smem(s + hist, t0, t1, n, g0, g1, tap0, tap1, overlap, tid, IM_TPC);
    __syncthreads();
    for (int i = tid; i < n; i += IM_TPC) row[y_offset + i] = s[hist + i];
}

// comb_filter (out of place, FIR; comb_filter/mod.rs:59-127, fallback.rs:6-29): no recursion, so
// every output sample is independent.
__global__ void __launch_bounds__(IM_TPC)
k_op_comb(float *__restrict__ y, const float *__restrict__ x, size_t row_stride, int offset, int n,
          const int32_t *__restrict__ params4, const float *__restrict__ gains2, int overlap)
{
    const float *xr = x + (size_t)blockIdx.x * row_stride + offset;
    float *yr = y + (size_t)blockIdx.x * row_stride + offset;
    int t0 = params4[4 * blockIdx.x], t1 = params4[4 * blockIdx.x + 1];
    const int tap0 = params4[4 * blockIdx.x + 2], tap1 = params4[4 * blockIdx.x + 3];
    const float g0 = gains2[2 * blockIdx.x], g1 = gains2[2 * blockIdx.x + 1];
    if (g0 == 0.0f && g1 == 0.0f) {
        for (int i = threadIdx.x; i < n; i += IM_TPC) yr[i] = xr[i];
        return;
    }
    t0 = max(t0, 15);
    t1 = max(t1, 15);
    const float g00 = g0 * g_tab.comb_gains[tap0 * 3], g01 = g0 * g_tab.comb_gains[tap0 * 3 + 1],
                g02 = g0 * g_tab.comb_gains[tap0 * 3 + 2];
    const float g10 = g1 * g_tab.comb_gains[tap1 * 3], g11 = g1 * g_tab.comb_gains[tap1 * 3 + 1],
                g12 = g1 * g_tab.comb_gains[tap1 * 3 + 2];
    if (fabsf(g0 - g1) < 1.1920929e-7f && t0 == t1 && tap0 == tap1) overlap = 0;
    for (int i = threadIdx.x; i < n; i += IM_TPC) {
        float acc = xr[i];
        if (i < overlap) {
            const float f = __ldg(&g_tab.window_sq[i]);
            acc = acc + (((1.0f - f) * g00) * xr[i - t0]);
            acc = acc + (((1.0f - f) * g01) * (xr[i - t0 + 1] + xr[i - t0 - 1]));
            acc = acc + (((1.0f - f) * g02) * (xr[i - t0 + 2] + xr[i - t0 - 2]));
            acc = acc + ((f * g10) * xr[i - t1]);
            acc = acc + ((f * g11) * (xr[i - t1 + 1] + xr[i - t1 - 1]));
            acc = acc + ((f * g12) * (xr[i - t1 + 2] + xr[i - t1 - 2]));
        } else if (g1 != 0.0f) {
            acc = acc + (g10 * xr[i - t1]) + (g11 * (xr[i - t1 + 1] + xr[i - t1 - 1])) +
                  (g12 * (xr[i - t1 + 2] + xr[i - t1 - 2]));
        }
        yr[i] = acc;
    }
}

}  // namespace opn
