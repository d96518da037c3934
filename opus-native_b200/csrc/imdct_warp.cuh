// imdct_warp.cuh -- kernel 1, warp-per-stream generation: one WARP decodes one stream (both channels).
//
// Device mirror of Mdct::backward (src/celt/mdct.rs:159-260), KissFft::process and its
// butterflies (src/celt/kiss_fft.rs:24-243), comb_filter_inplace
// (src/celt/comb_filter/mod.rs:130-193, scalar kernel fallback.rs:32-53).
//
// Arithmetic contract (unchanged from imdct.cuh): every sum and product is evaluated in the
// reference's order with single roundings; the TU is compiled with -fmad=false.  What changes is
// the placement of the work, which is what the first version spent 80 % of its instructions on:
//
//   * The reference's stage list for nfft = 480 >> shift is  4(m=1) [2|4](m=4) [4](m=8) | 3 | 5.
//     Everything left of the bar only ever combines elements inside one aligned group of
//     GS = 32 >> shift positions, and the two stages right of it only combine the 15 elements
//     {u + GS*j}.  So the transform is two register-resident passes with ONE shared-memory
//     transpose between them and no index arithmetic at run time:
//       pass A  lane (channel, g), g < 15: pre-rotates its GS elements straight into registers
//               (the digit-reversal permutation is resolved at compile time: group g holds the
//               inputs i = r + 15 q with r = g/3 + 5 (g%3)), runs the radix-4/2/4 stages with
//               twiddles that are compile-time constants (constant-bank operands);
//       pass B  lane u < GS (x block x channel): 15 elements, radix-3 then radix-5, post-rotation
//               fused on the registers (FFT output k yields out[2k] and out[n2-1-2k] from the same
//               two trig values).
//   * A warp never waits for another warp: all hand-offs are __syncwarp().  A CTA is W_WPC
//     independent warps that only share read-only tables (trig pairs, twiddles, window) staged in
//     shared memory once per CTA.
//   * Coefficient rows arrive by TMA (cp.async.bulk, one 3840-byte row per channel) into the very
//     row that later holds the output; the transpose buffer aliases that row too (every lane has its
//     inputs in registers before the first transposed element is written), so a stream needs
//     C x (nf + 60) floats of shared memory and ~20 streams are resident per SM.
//   * The post-filter reads its history straight from the interleaved PCM ring in HBM/L2 (one
//     float2 = both channels of a tap).  History-only spans of the frame have no recursion and are
//     filtered in one parallel sweep, 4 samples per lane; only the truly recursive remainder is
//     swept in chunks, from shared memory.
#pragma once
#include "imdct.cuh"
#include "opn_tables.h"

namespace opn {

// Twiddles addressed with compile-time indices (src/celt/kiss_fft.rs:341-582) fold into instruction
// immediates: OPN_TWIDDLES is constexpr in C++ translation units.
template <int I> struct WTw {
    static constexpr float re = OPN_TWIDDLES[2 * I], im = OPN_TWIDDLES[2 * I + 1];
    __device__ __forceinline__ static float2 get() { return make_float2(re, im); }
};

#ifndef OPN_K1_MIN_CTAS
#define OPN_K1_MIN_CTAS 20
#endif
constexpr int W_K1_MIN_CTAS = OPN_K1_MIN_CTAS;  // kernel 1 register budget: 65536 / (32 * this) per thread
__host__ __device__ constexpr int w_ch_floats(int lm) { return (120 << lm) + 60; }
__host__ __device__ constexpr int trig_pair_off(int shift) { return shift == 0 ? 0 : shift == 1 ? 480 : shift == 2 ? 720 : 840; }
__host__ __device__ constexpr size_t w_smem_bytes(int lm, int channels) { return (size_t)channels * w_ch_floats(lm) * 4 + 16; }
__host__ __device__ constexpr size_t w_comb_smem_bytes(int lm, int channels) { return (size_t)channels * (HIST_CAP + (120 << lm)) * 4 + 16; }

// ---- TMA / mbarrier (single-CTA cluster) ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}

// ---- butterflies on registers, twiddles passed in (same operation order as imdct.cuh) --------
// kiss_fft.rs:148-187
__device__ __forceinline__ void r_bfly4(float2 &d0, float2 &d1, float2 &d2, float2 &d3, float2 w1, float2 w2, float2 w3)
{
    const float2 s0 = c_mul(d1, w1), s1 = c_mul(d2, w2), s2 = c_mul(d3, w3);
    const float2 s5 = c_sub(d0, s1);
    const float2 a0 = c_add(d0, s1);
    const float2 s3 = c_add(s0, s2), s4 = c_sub(s0, s2);
    d2 = c_sub(a0, s3);
    d0 = c_add(a0, s3);
    d1 = make_float2(s5.x + s4.y, s5.y - s4.x);
    d3 = make_float2(s5.x - s4.y, s5.y + s4.x);
}
// kiss_fft.rs:129-147 (m == 1)
__device__ __forceinline__ void r_bfly4_m1(float2 &d0, float2 &d1, float2 &d2, float2 &d3)
{
    const float2 s0 = c_sub(d0, d2);
    float2 s1 = c_add(d1, d3);
    const float2 a0 = c_add(d0, d2);
    d2 = c_sub(a0, s1);
    d0 = c_add(a0, s1);
    s1 = c_sub(d1, d3);
    d1 = make_float2(s0.x + s1.y, s0.y - s1.x);
    d3 = make_float2(s0.x - s1.y, s0.y + s1.x);
}
// kiss_fft.rs:55-87, pair J of a group of 8: (lo, hi) = (d[J], d[4+J])
template <int J> __device__ __forceinline__ void r_bfly2(float2 &lo, float2 &hi)
{
    const float2 x = hi;
    float2 t;
    if (J == 0) t = x;
    else if (J == 1) t = make_float2((x.x + x.y) * OPN_FRAC_1_SQRT_2, (x.y - x.x) * OPN_FRAC_1_SQRT_2);
    else if (J == 2) t = make_float2(x.y, -x.x);
    else t = make_float2((x.y - x.x) * OPN_FRAC_1_SQRT_2, (-(x.y + x.x)) * OPN_FRAC_1_SQRT_2);
    const float2 a = lo;
    hi = c_sub(a, t);
    lo = c_add(a, t);
}
// kiss_fft.rs:89-127
__device__ __forceinline__ void r_bfly3(float2 &d0, float2 &d1, float2 &d2, float2 w1, float2 w2, float epi3y)
{
    const float2 s1 = c_mul(d1, w1), s2 = c_mul(d2, w2);
    const float2 s3 = c_add(s1, s2);
    float2 s0 = c_sub(s1, s2);
    const float2 dm = c_sub(d0, c_scale(s3, 0.5f));
    s0 = c_scale(s0, epi3y);
    d0 = c_add(d0, s3);
    d2 = make_float2(dm.x + s0.y, dm.y - s0.x);
    d1 = make_float2(dm.x - s0.y, dm.y + s0.x);
}
// kiss_fft.rs:190-243
__device__ __forceinline__ void r_bfly5(float2 &d0, float2 &d1, float2 &d2, float2 &d3, float2 &d4, float2 w1, float2 w2,
                                        float2 w3, float2 w4, float2 ya, float2 yb)
{
    const float2 s0 = d0;
    const float2 s1 = c_mul(d1, w1), s2 = c_mul(d2, w2), s3 = c_mul(d3, w3), s4 = c_mul(d4, w4);
    const float2 s7 = c_add(s1, s4), s10 = c_sub(s1, s4);
    const float2 s8 = c_add(s2, s3), s9 = c_sub(s2, s3);
    d0 = c_add(s0, c_add(s7, s8));
    float2 s5, s6, s11, s12;
    s5.x = s0.x + (s7.x * ya.x + s8.x * yb.x);
    s5.y = s0.y + (s7.y * ya.x + s8.y * yb.x);
    s6.x = s10.y * ya.y + s9.y * yb.y;
    s6.y = -(s10.x * ya.y + s9.x * yb.y);
    d1 = c_sub(s5, s6);
    d4 = c_add(s5, s6);
    s11.x = s0.x + (s7.x * yb.x + s8.x * ya.x);
    s11.y = s0.y + (s7.y * yb.x + s8.y * ya.x);
    s12.x = s9.y * ya.y - s10.y * yb.y;
    s12.y = s10.x * yb.y - s9.x * ya.y;
    d2 = c_add(s11, s12);
    d3 = c_sub(s11, s12);
}

// radix-4 butterfly U of the last in-group stage: elements U + M*{0,1,2,3}, twiddle stride S
template <int U, int M, int S> __device__ __forceinline__ void w_bfly4_const(float2 *d)
{
    r_bfly4(d[U], d[U + M], d[U + 2 * M], d[U + 3 * M], WTw<S * U>::get(), WTw<2 * S * U>::get(), WTw<3 * S * U>::get());
}

// In-group position p (0 <= p < GS) -> q, where the group's inputs are i = r + 15 q
// (digit reversal of kiss_fft.rs:281-336 for the factor lists :251,259,267,275; checked against the
// generated tables in tests/test_host_logic.py::test_digit_reversal_closed_form).
template <int SHIFT> __host__ __device__ constexpr int w_qmap(int p)
{
    return SHIFT == 0 ? ((p >> 3) + 4 * ((p >> 2) & 1) + 8 * (p & 3))
         : SHIFT == 1 ? ((p >> 2) + 4 * (p & 3))
         : SHIFT == 2 ? ((p >> 2) + 2 * (p & 3))
                      : p;
}

// Stages that stay inside one group of GS positions (execution order, kiss_fft.rs:38-52).
template <int SHIFT> __device__ __forceinline__ void w_group_stages(float2 *d)
{
    constexpr int GS = 32 >> SHIFT;
#pragma unroll
    for (int b = 0; b < GS / 4; b++) r_bfly4_m1(d[4 * b], d[4 * b + 1], d[4 * b + 2], d[4 * b + 3]);
    if constexpr (SHIFT == 0 || SHIFT == 2) {  // radix 2, m = 4
#pragma unroll
        for (int g8 = 0; g8 < GS / 8; g8++) {
            r_bfly2<0>(d[8 * g8 + 0], d[8 * g8 + 4]);
            r_bfly2<1>(d[8 * g8 + 1], d[8 * g8 + 5]);
            r_bfly2<2>(d[8 * g8 + 2], d[8 * g8 + 6]);
            r_bfly2<3>(d[8 * g8 + 3], d[8 * g8 + 7]);
        }
    }
    if constexpr (SHIFT == 0) {  // radix 4, m = 8, twiddle stride 15
        w_bfly4_const<0, 8, 15>(d);
        w_bfly4_const<1, 8, 15>(d);
        w_bfly4_const<2, 8, 15>(d);
        w_bfly4_const<3, 8, 15>(d);
        w_bfly4_const<4, 8, 15>(d);
        w_bfly4_const<5, 8, 15>(d);
        w_bfly4_const<6, 8, 15>(d);
        w_bfly4_const<7, 8, 15>(d);
    }
    if constexpr (SHIFT == 1) {  // radix 4, m = 4, twiddle stride 30
        w_bfly4_const<0, 4, 30>(d);
        w_bfly4_const<1, 4, 30>(d);
        w_bfly4_const<2, 4, 30>(d);
        w_bfly4_const<3, 4, 30>(d);
    }
}

// Mdct::backward for the C channels of one stream, NBLK interleaved blocks of nfft = 480 >> SHIFT
// (NBLK == 1: one long block; SHIFT == 3 and NBLK == 2^LM: transient frame).  `o` is channel 0's
// row of CHF = nf + 60 floats (channel 1 follows): on entry it holds the coefficients, on exit
// out[0 .. nf+60) after the TDAC mirror (mdct.rs:241-259).  `carry` is this lane's float4 of the
// previous tail (lane = 15*ch + k -> out[4k .. 4k+4)).
template <int SHIFT, int NBLK, int C>
__device__ __forceinline__ void w_imdct(float *o, int lane, float4 carry, const float2 *tpair, const float2 *tw, const float *win)
{
    constexpr int GS = 32 >> SHIFT, N2 = 960 >> SHIFT;
    constexpr int NF = N2 * NBLK, E = GS * NBLK;
    constexpr int CHF = NF + 60;
    // Transpose buffer (aliases the row): element `a` (= position inside the transform, plus 60 per
    // short block) lives at a + a / PADG -- one float2 of padding per PADG elements keeps the
    // stride-GS writes of pass A and the unit-stride reads of pass B on distinct banks.  Short blocks
    // only fit the row with the coarser padding.
    constexpr int PADG = NBLK > 1 ? 16 : GS;
    constexpr int NFFT = 15 * GS;
    static_assert(2 * (NFFT * NBLK + (NFFT * NBLK + PADG - 1) / PADG) <= CHF, "transpose buffer must fit the row");

    // ---------------------------------------------------------------- pass A
    {
        const bool on = lane < 15 * C;
        const int ch = (C == 2 && lane >= 15) ? 1 : 0;
        const int g = on ? lane - 15 * ch : 0;
        const int j1 = g / 3, j2 = g - 3 * j1, r = j1 + 5 * j2;
        const float *x = o + ch * CHF;
        const float2 *tp = tpair + r;
        float2 d[NBLK * GS];
        if (on) {
#pragma unroll
            for (int blk = 0; blk < NBLK; blk++)
#pragma unroll
                for (int p = 0; p < GS; p++) {  // pre-rotation, mdct.rs:184-200
                    const int q = w_qmap<SHIFT>(p);
                    const float x0 = x[blk + NBLK * (2 * r) + NBLK * 30 * q];
                    const float x1 = x[blk + NBLK * (N2 - 1 - 2 * r) - NBLK * 30 * q];
                    const float2 t = __ldg(tp + 15 * q);  // (trig[i], trig[n4 + i])
                    const float re = (x1 * t.x) + (x0 * t.y);
                    const float im = (x0 * t.x) - (x1 * t.y);
                    d[blk * GS + p] = make_float2(im, re);
                }
        }
        __syncwarp();  // every coefficient is in a register: the row may be overwritten
        if (on) {
            float2 *xc = reinterpret_cast<float2 *>(o + ch * CHF);
#pragma unroll
            for (int blk = 0; blk < NBLK; blk++) {
                w_group_stages<SHIFT>(d + blk * GS);
                const int a = NFFT * blk + GS * g;  // p < GS <= PADG never crosses a padding boundary
#pragma unroll
                for (int p = 0; p < GS; p++) xc[a + a / PADG + p] = d[blk * GS + p];
            }
        }
    }
    __syncwarp();
    // ---------------------------------------------------------------- pass B
    {
        constexpr int S3 = 5 << SHIFT;  // twiddle stride of the radix-3 stage
        const int col = lane % E, ch0 = lane / E;
        const int blk = col / GS, u = col % GS;
        const float2 w31 = __ldg(tw + u * S3), w32 = __ldg(tw + 2 * u * S3);
        float2 w5[3][4];
#pragma unroll
        for (int jj = 0; jj < 3; jj++)
#pragma unroll
            for (int k = 0; k < 4; k++) w5[jj][k] = __ldg(tw + (((u + GS * jj) * (k + 1)) << SHIFT));
        constexpr float epi3y = WTw<160>::im;
        const float2 ya = WTw<96>::get(), yb = WTw<192>::get();
        constexpr int ITER = (C * E + 31) / 32;
#pragma unroll 1
        for (int it = 0; it < ITER; it++) {
            const int ch = ch0 + it * (32 / E);
            const bool on = ch < C;
            float2 d[15];
            if (on) {
                const float2 *xc = reinterpret_cast<const float2 *>(o + ch * CHF);
#pragma unroll
                for (int j = 0; j < 15; j++) {
                    if constexpr (NBLK == 1) {
                        d[j] = xc[u + GS * j + (GS * j) / PADG];  // u < GS = PADG adds no pad
                    } else {
                        const int a = NFFT * blk + GS * j + u;
                        d[j] = xc[a + a / PADG];
                    }
                }
            }
            __syncwarp();  // transposed elements are in registers: the row may be overwritten
            if (on) {
#pragma unroll
                for (int a = 0; a < 5; a++) r_bfly3(d[3 * a], d[3 * a + 1], d[3 * a + 2], w31, w32, epi3y);
#pragma unroll
                for (int jj = 0; jj < 3; jj++)
                    r_bfly5(d[jj], d[jj + 3], d[jj + 6], d[jj + 9], d[jj + 12], w5[jj][0], w5[jj][1], w5[jj][2], w5[jj][3], ya, yb);
                // post-rotation and de-shuffle, mdct.rs:205-238: FFT output k = u + GS*j
                float *ob = o + ch * CHF + N2 * blk + 60;
                const float2 *tp = tpair + u;
#pragma unroll
                for (int j = 0; j < 15; j++) {
                    const float2 t = __ldg(tp + GS * j);
                    ob[2 * u + 2 * GS * j] = (d[j].y * t.x) + (d[j].x * t.y);
                    ob[N2 - 1 - 2 * u - 2 * GS * j] = (d[j].y * t.y) - (d[j].x * t.x);
                }
            }
        }
    }
    // out[0..60) <- previous tail
    if (lane < 15 * C) {
        const int ch = (C == 2 && lane >= 15) ? 1 : 0;
        *reinterpret_cast<float4 *>(o + ch * CHF + 4 * (lane - 15 * ch)) = carry;
    }
    __syncwarp();
    // ---------------------------------------------------------------- TDAC mirror, mdct.rs:241-259
    for (int w = lane; w < C * NBLK * 60; w += 32) {
        int ch = 0, blk = 0, i = w;
        if constexpr (NBLK == 1) {
            if (C == 2 && w >= 60) { ch = 1; i = w - 60; }
        } else {
            ch = w / (NBLK * 60);
            const int rem = w - ch * (NBLK * 60);
            blk = rem / 60;
            i = rem - blk * 60;
        }
        float *ob = o + ch * CHF + N2 * blk;
        const float x0 = ob[119 - i], x1 = ob[i];
        const float w0 = __ldg(win + i), w1 = __ldg(win + 119 - i);
        ob[i] = (w1 * x1) - (w0 * x0);
        ob[119 - i] = (w0 * x1) + (w1 * x0);
    }
    __syncwarp();
}

// ---- post-filter -------------------------------------------------------------------------------
// comb_filter_const_inplace term order (fallback.rs:46-51): y + g0*x2 + g1*(x1+x3) + g2*(x0+x4)
__device__ __forceinline__ float comb5(float y, float x0, float x1, float x2, float x3, float x4, float g0, float g1, float g2)
{
    return y + (g0 * x2) + (g1 * (x1 + x3)) + (g2 * (x0 + x4));
}

// One sample of the C interleaved channels (the PCM ring layout): C == 2 moves both channels of a tap
// in one 64-bit access.
template <int C> struct WSmp;
template <> struct WSmp<1> {
    float a;
    __device__ __forceinline__ static WSmp ld(const float *p) { return WSmp{*p}; }
    __device__ __forceinline__ void st(float *p) const { *p = a; }
};
template <> struct WSmp<2> {
    float a, b;
    __device__ __forceinline__ static WSmp ld(const float *p)
    {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        return WSmp{v.x, v.y};
    }
    __device__ __forceinline__ void st(float *p) const { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
};
template <int C>
__device__ __forceinline__ WSmp<C> w_comb5(WSmp<C> y, WSmp<C> x0, WSmp<C> x1, WSmp<C> x2, WSmp<C> x3, WSmp<C> x4, float g0, float g1,
                                           float g2)
{
    WSmp<C> r;
    r.a = comb5(y.a, x0.a, x1.a, x2.a, x3.a, x4.a, g0, g1, g2);
    if constexpr (C == 2) r.b = comb5(y.b, x0.b, x1.b, x2.b, x3.b, x4.b, g0, g1, g2);
    return r;
}
// cross-fade accumulation order of comb_filter_inplace (mod.rs:166-177); a* = y[i-t0-2 .. i-t0+2],
// b* = y[i-t1-2 .. i-t1+2]
__device__ __forceinline__ float xfade1(float y, float a0, float a1, float a2, float a3, float a4, float b0, float b1, float b2, float b3,
                                        float b4, bool has0, bool has1, float f, float g00, float g01, float g02, float g10, float g11,
                                        float g12)
{
    float v = y;
    if (has0) {
        v = v + (((1.0f - f) * g00) * a2);
        v = v + (((1.0f - f) * g01) * (a3 + a1));
        v = v + (((1.0f - f) * g02) * (a4 + a0));
    }
    if (has1) {
        v = v + ((f * g10) * b2);
        v = v + ((f * g11) * (b3 + b1));
        v = v + ((f * g12) * (b4 + b0));
    }
    return v;
}
template <int C>
__device__ __forceinline__ WSmp<C> w_xfade(WSmp<C> y, const WSmp<C> *a, const WSmp<C> *b, bool has0, bool has1, float f, float g00,
                                           float g01, float g02, float g10, float g11, float g12)
{
    WSmp<C> r;
    r.a = xfade1(y.a, a[0].a, a[1].a, a[2].a, a[3].a, a[4].a, b[0].a, b[1].a, b[2].a, b[3].a, b[4].a, has0, has1, f, g00, g01, g02, g10,
                 g11, g12);
    if constexpr (C == 2)
        r.b = xfade1(y.b, a[0].b, a[1].b, a[2].b, a[3].b, a[4].b, b[0].b, b[1].b, b[2].b, b[3].b, b[4].b, has0, has1, f, g00, g01, g02,
                     g10, g11, g12);
    return r;
}

// comb_filter_inplace (comb_filter/mod.rs:130-193) on one stream: y points at sample 0 of the frame,
// sample i of channel c lives at y[C*i + c] and the history (the previous max(T)+2 samples) directly
// below, exactly as in the PCM ring.  The filter is recursive, y[i] depends on y[i-T-2 .. i-T+2];
// wherever those taps are history the samples are independent and are filtered in parallel, the rest
// is swept in chunks no longer than T-2.  A tap set whose gain is exactly zero contributes +-0 to
// every sum and is skipped.
// tap gains of the old and the new filter (comb_filter/mod.rs:45-55, 146-151); a separate step so that kernel 2 can
// have the table reads in flight while its samples are still arriving
struct CombGains {
    float g00, g01, g02, g10, g11, g12;
};
__device__ __forceinline__ CombGains w_comb_gains(float g0, float g1, int tap0, int tap1)
{
    CombGains k;
    k.g00 = g0 * g_tab.comb_gains[tap0 * 3];
    k.g01 = g0 * g_tab.comb_gains[tap0 * 3 + 1];
    k.g02 = g0 * g_tab.comb_gains[tap0 * 3 + 2];
    k.g10 = g1 * g_tab.comb_gains[tap1 * 3];
    k.g11 = g1 * g_tab.comb_gains[tap1 * 3 + 1];
    k.g12 = g1 * g_tab.comb_gains[tap1 * 3 + 2];
    return k;
}
template <int C>
__device__ __forceinline__ void w_comb(float *y, int t0, int t1, int n, float g0, float g1, int tap0, int tap1, int overlap, int lane,
                                       const float *win_sq, const CombGains &kg)
{
    using S = WSmp<C>;
    if (g0 == 0.0f && g1 == 0.0f) return;
    t0 = max(t0, 15);
    t1 = max(t1, 15);
    const float g00 = kg.g00, g01 = kg.g01, g02 = kg.g02, g10 = kg.g10, g11 = kg.g11, g12 = kg.g12;
    if (fabsf(g0 - g1) < 1.1920929e-7f && t0 == t1 && tap0 == tap1) overlap = 0;
    const bool has0 = g0 != 0.0f, has1 = g1 != 0.0f;

    // Independent samples are dealt to the lanes round-robin (sample = base + lane + 32 e): consecutive
    // lanes touch consecutive float2, so every shared-memory access is conflict-free.
    // ---- cross-fade part (mod.rs:162-179): samples [0, overlap)
    if (overlap > 0) {
        const int tmin = min(has0 ? t0 : 1 << 20, has1 ? t1 : 1 << 20);
        // [0, pre): every tap of every live set is history -> no recursion
        const int pre = min(overlap, tmin - 2);
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int i = lane + 32 * e;
            if (i < pre) {
                const float f = __ldg(win_sq + i);
                S a[5], b[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    a[k] = has0 ? S::ld(y + C * (i - t0 - 2 + k)) : S{};
                    b[k] = has1 ? S::ld(y + C * (i - t1 - 2 + k)) : S{};
                }
                w_xfade<C>(S::ld(y + C * i), a, b, has0, has1, f, g00, g01, g02, g10, g11, g12).st(y + C * i);
            }
        }
        __syncwarp();
        // [pre, overlap): chunks of W <= tmin - 2 samples, one per lane
        const int W = min(tmin - 2, 32);
        for (int base = pre; base < overlap; base += W) {
            const int i = base + lane;
            if (lane < W && i < overlap) {
                const float f = __ldg(win_sq + i);
                S a[5], b[5];
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    a[k] = has0 ? S::ld(y + C * (i - t0 - 2 + k)) : S{};
                    b[k] = has1 ? S::ld(y + C * (i - t1 - 2 + k)) : S{};
                }
                w_xfade<C>(S::ld(y + C * i), a, b, has0, has1, f, g00, g01, g02, g10, g11, g12).st(y + C * i);
            }
            __syncwarp();
        }
    }
    if (!has1) return;

    // ---- constant part (fallback.rs:32-53): samples [overlap, n)
    auto one = [&](int i) {
        const float *p = y + C * (i - t1);
        w_comb5<C>(S::ld(y + C * i), S::ld(p + 2 * C), S::ld(p + C), S::ld(p), S::ld(p - C), S::ld(p - 2 * C), g10, g11, g12).st(y + C * i);
    };
    // (1) history-only span [overlap, hend): no recursion, 128 samples per step
    int at = overlap;
    {
        const int hend = max(at, min(n, t1 - 2));
        for (int base = at; base < hend; base += 128) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int i = base + lane + 32 * e;
                if (i < hend) one(i);
            }
        }
        at = hend;
        __syncwarp();
    }
    // (2) recursive remainder [at, n): chunks of min(t1 - 2, 128) samples; inside a chunk every tap lies
    //     before the chunk
    if (t1 - 2 >= 128) {
        for (int base = at; base < n; base += 128) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int i = base + lane + 32 * e;
                if (i < n) one(i);
            }
            __syncwarp();
        }
    } else if (t1 - 2 >= 64) {
        for (int base = at; base < n; base += 64) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int i = base + lane + 32 * e;
                if (i < n) one(i);
            }
            __syncwarp();
        }
    } else {
        const int W = min(t1 - 2, 32);
        for (int base = at; base < n; base += W) {
            const int i = base + lane;
            if (lane < W && i < n) one(i);
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 1 (IMDCT + TDAC overlap-add + PCM store): one warp = one CTA = one stream (item).
// Shared memory per warp: C rows of nf+60 floats and one mbarrier.  The post-filter runs in kernel 2
// on the interleaved PCM this kernel leaves in the ring; kernel 1 only decides whether it is needed and
// leaves the parameters (old -> new) in `job`.
template <int LM, int C> __global__ void __launch_bounds__(32, W_K1_MIN_CTAS) k_imdct_post_w(ImdctArgs A)
{
    extern __shared__ __align__(16) float o[];
    constexpr int NF = 120 << LM;
    constexpr int CHF = w_ch_floats(LM);
    const int lane = threadIdx.x;
    uint64_t *bar = reinterpret_cast<uint64_t *>(o + C * CHF);

    const uint32_t item = blockIdx.x;
    const uint32_t stream = A.stream_idx ? A.stream_idx[item] : item;
    // coefficient rows -> output rows by TMA, before anything else (the row address only needs `stream`)
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, C * NF * 4);
#pragma unroll
        for (int c = 0; c < C; c++) bulk_g2s(o + c * CHF, A.coef + ((size_t)stream * C + c) * NF, NF * 4, bar);
    }
    // everything the frame needs from per-stream state, requested in one go
    const opn_synth_side *side = A.side + stream;
    const int32_t status = A.status[stream];
    const PfState old = A.pf[stream];
    const uint32_t pos = A.ring_pos[stream];
    const int s_transient = side->transient, s_on = side->postfilter, s_period = side->period, s_gain = side->gain_idx,
              s_tapset = side->tapset;
    const uint32_t s_final = side->final_rng;
    const uint32_t dense_off = (A.dense && A.dense_off) ? A.dense_off[item] : 0u;
    float *carry_g = A.carry + (size_t)stream * C * 60;
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < 15 * C) carry = *reinterpret_cast<const float4 *>(carry_g + 4 * lane);  // [C][60] = 15 float4 per channel
    __syncwarp();
    if (status < 0) {  // rejected packet: state untouched (decoder.rs:397)
        if (lane == 0) {
            if (A.result) A.result[stream] = status;
            A.job[item].on = 0;
        }
        mbar_wait(bar, 0);  // the rows are in flight: do not retire the CTA under them
        return;
    }
    const bool lost = status == ITEM_LOST;
    // post-filter parameters: previous frame -> this frame
    int t1 = old.period, tap1 = old.tapset;
    float g1 = old.gain;
    if (!lost) {
        t1 = s_on ? s_period : 0;
        g1 = s_on ? 0.09375f * (float)(s_gain + 1) : 0.0f;
        tap1 = s_on ? s_tapset : 0;
    }
    const bool comb_on = A.postfilter && (old.gain != 0.0f || g1 != 0.0f);
    float *ring = A.ring + (size_t)stream * RING_SAMPLES * C;

    mbar_wait(bar, 0);
    if constexpr (LM > 0) {
        if (s_transient) w_imdct<3, (1 << LM), C>(o, lane, carry, g_tab.trig_pair + trig_pair_off(3), g_tab.twiddles, g_tab.window);
        else w_imdct<3 - LM, 1, C>(o, lane, carry, g_tab.trig_pair + trig_pair_off(3 - LM), g_tab.twiddles, g_tab.window);
    } else {
        w_imdct<3, 1, C>(o, lane, carry, g_tab.trig_pair + trig_pair_off(3), g_tab.twiddles, g_tab.window);
    }

    // tail of this frame -> carry
    if (lane < 15 * C) {
        const int ch = (C == 2 && lane >= 15) ? 1 : 0;
        *reinterpret_cast<float4 *>(carry_g + 4 * lane) = *reinterpret_cast<const float4 *>(o + ch * CHF + NF + 4 * (lane - 15 * ch));
    }
    // interleaved PCM -> ring (history + device-resident output); dense rows only when no post-filter follows.
    // The frame is contiguous in the ring except when it wraps (pos is a multiple of 120).
    float *dense = (A.dense && !comb_on) ? A.dense + (size_t)stream * A.dense_stride + dense_off : nullptr;
    const float gain = A.gain;
    constexpr int VEC = C == 2 ? NF / 2 : NF / 4;         // float4 per frame
    constexpr int SPV = C == 2 ? 2 : 4;                   // samples per float4
    const int wrap_at = (int)(RING_SAMPLES - pos) / SPV;  // first float4 that lands at the ring start
    float4 *r0 = reinterpret_cast<float4 *>(ring + (size_t)pos * C);
    float4 *r1 = reinterpret_cast<float4 *>(ring) - wrap_at;
#pragma unroll
    for (int i = lane; i < VEC; i += 32) {
        float4 v;
        if (C == 2) {
            const float2 a = *reinterpret_cast<const float2 *>(o + 2 * i), b = *reinterpret_cast<const float2 *>(o + CHF + 2 * i);
            v = make_float4(a.x, b.x, a.y, b.y);
        } else {
            v = *reinterpret_cast<const float4 *>(o + 4 * i);
        }
        (i < wrap_at ? r0 : r1)[i] = v;
        if (dense) {
            if (gain != 1.0f) { v.x *= gain; v.y *= gain; v.z *= gain; v.w *= gain; }
            reinterpret_cast<float4 *>(dense)[i] = v;
        }
    }
    if (lane == 0) {
        uint32_t np = pos + (uint32_t)NF;
        if (np >= RING_SAMPLES) np -= RING_SAMPLES;
        A.ring_pos[stream] = np;
        PfState nw;
        nw.period = t1;
        nw.tapset = tap1;
        nw.gain = g1;
        nw.pad = 0;
        A.pf[stream] = nw;
        if (A.result) A.result[stream] = NF;
        if (A.final_range) A.final_range[stream] = lost ? 0u : s_final;
        CombJob j;
        j.on = comb_on ? 1 : 0;
        j.pos = pos;
        j.t0 = old.period;
        j.t1 = t1;
        j.tap0 = old.tapset;
        j.tap1 = tap1;
        j.g0 = old.gain;
        j.g1 = g1;
        A.job[item] = j;
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 2 (pitch comb post-filter, comb_filter_inplace): one warp = one CTA = one stream (item) whose job is on.
// The history (the T+2 samples before the frame) and the frame are one contiguous span of the
// interleaved ring: one TMA transfer (three when the span wraps) brings both into shared memory, the
// filter runs in place on float2 = (left, right) samples, and the frame goes back to the ring and, for
// host-buffer calls, to the dense output rows.
template <int LM, int C> __global__ void __launch_bounds__(32) k_comb_post_w(ImdctArgs A)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int NF = 120 << LM;
    const int lane = threadIdx.x;
    float *y = sm + C * HIST_CAP;  // sample 0 of the frame; history below
    uint64_t *bar = reinterpret_cast<uint64_t *>(y + C * NF);

    const uint32_t item = blockIdx.x;
    const CombJob j = A.job[item];
    if (!j.on) return;
    const uint32_t stream = A.stream_idx ? A.stream_idx[item] : item;
    float *ring = A.ring + (size_t)stream * RING_SAMPLES * C;
    const int need = (max(max(j.t0, j.t1), 15) + 2 + 3) & ~3;  // multiple of 4: every piece 16-byte sized and aligned
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (need + NF) * C * 4);
        // span [pos - need, pos + NF) of the ring, cut where it wraps
        int first = (int)j.pos - need, count = need + NF;
        float *dst = y - need * C;
        if (first < 0) {
            bulk_g2s(dst, ring + (size_t)(first + RING_SAMPLES) * C, -first * C * 4, bar);
            dst += -first * C;
            count += first;
            first = 0;
        }
        const int fit = min(count, RING_SAMPLES - first);
        bulk_g2s(dst, ring + (size_t)first * C, fit * C * 4, bar);
        if (count > fit) bulk_g2s(dst + fit * C, ring, (count - fit) * C * 4, bar);
    }
    const uint32_t dense_off = (A.dense && A.dense_off) ? A.dense_off[item] : 0u;
    const CombGains kg = w_comb_gains(j.g0, j.g1, j.tap0, j.tap1);
    __syncwarp();
    mbar_wait(bar, 0);
    w_comb<C>(y, j.t0, j.t1, NF, j.g0, j.g1, j.tap0, j.tap1, 120, lane, g_tab.window_sq, kg);
    __syncwarp();
    float *dense = A.dense ? A.dense + (size_t)stream * A.dense_stride + dense_off : nullptr;
    const float gain = A.gain;
    constexpr int VEC = NF * C / 4;
    constexpr int SPV = 4 / C;
    const int wrap_at = (int)(RING_SAMPLES - j.pos) / SPV;
    float4 *r0 = reinterpret_cast<float4 *>(ring + (size_t)j.pos * C);
    float4 *r1 = reinterpret_cast<float4 *>(ring) - wrap_at;
#pragma unroll
    for (int i = lane; i < VEC; i += 32) {
        float4 v = reinterpret_cast<const float4 *>(y)[i];
        (i < wrap_at ? r0 : r1)[i] = v;
        if (dense) {
            if (gain != 1.0f) { v.x *= gain; v.y *= gain; v.z *= gain; v.w *= gain; }
            reinterpret_cast<float4 *>(dense)[i] = v;
        }
    }
}

// Operator-level Mdct::backward on independent rows (tests; opn_op_imdct_tdac): one warp per row.
template <int SHIFT, int NBLK>
__global__ void __launch_bounds__(32)
k_op_imdct_w(const float *__restrict__ input, size_t in_stride, float *__restrict__ output, size_t out_stride)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int N2 = 960 >> SHIFT, NF = N2 * NBLK;
    const int lane = threadIdx.x;
    const float *in = input + (size_t)blockIdx.x * in_stride;
    float *out = output + (size_t)blockIdx.x * out_stride;
    for (int i = lane; i < NF; i += 32) sm[i] = in[i];
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < 15) carry = make_float4(out[4 * lane], out[4 * lane + 1], out[4 * lane + 2], out[4 * lane + 3]);
    __syncwarp();
    w_imdct<SHIFT, NBLK, 1>(sm, lane, carry, g_tab.trig_pair + trig_pair_off(SHIFT), g_tab.twiddles, g_tab.window);
    for (int i = lane; i < NF + 60; i += 32) out[i] = sm[i];
}

// Operator-level comb_filter_inplace on rows (tests; opn_op_comb_filter_inplace): one warp per row,
// the row's own prefix [y_offset - hist, y_offset) plays the role of the PCM ring.
__global__ void __launch_bounds__(32)
k_op_comb_inplace_w(float *__restrict__ y, size_t row_stride, int y_offset, int n, const int32_t *__restrict__ params4,
                    const float *__restrict__ gains2, int overlap)
{
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x;
    float *row = y + (size_t)blockIdx.x * row_stride;
    const int t0 = params4[4 * blockIdx.x], t1 = params4[4 * blockIdx.x + 1];
    const int tap0 = params4[4 * blockIdx.x + 2], tap1 = params4[4 * blockIdx.x + 3];
    const float g0 = gains2[2 * blockIdx.x], g1 = gains2[2 * blockIdx.x + 1];
    // shared memory: [HIST_CAP history | n samples]; the row's own prefix plays the role of the PCM ring
    const int need = min(max(max(t0, t1), 15) + 2, y_offset);
    float *ys = sm + HIST_CAP;
    for (int i = lane; i < n; i += 32) ys[i] = row[y_offset + i];
    for (int i = lane; i < need; i += 32) ys[-1 - i] = row[y_offset - 1 - i];
    __syncwarp();
    w_comb<1>(ys, t0, t1, n, g0, g1, tap0, tap1, overlap, lane, g_tab.window_sq, w_comb_gains(g0, g1, tap0, tap1));
    __syncwarp();
    for (int i = lane; i < n; i += 32) row[y_offset + i] = ys[i];
}

}  // namespace opn
