// imdct_warp.cuh -- Mdct::backward by one WARP for the channels of one stream (device function w_imdct).
//
// Device mirror of Mdct::backward (src/celt/mdct.rs:159-260), KissFft::process and its
// butterflies (src/celt/kiss_fft.rs:24-243).
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
//   * The kernels that call w_imdct (frame kernel, operator kernel) are in frame_warp.cuh.
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

__host__ __device__ constexpr int w_ch_floats(int lm) { return (120 << lm) + 60; }
__host__ __device__ constexpr int trig_pair_off(int shift) { return shift == 0 ? 0 : shift == 1 ? 480 : shift == 2 ? 720 : 840; }

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
    d1 = c_add_mrot(s5, s4);
    d3 = c_sub_mrot(s5, s4);
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
    d1 = c_add_mrot(s0, s1);
    d3 = c_sub_mrot(s0, s1);
}
// kiss_fft.rs:55-87, pair J of a group of 8: (lo, hi) = (d[J], d[4+J])
template <int J> __device__ __forceinline__ void r_bfly2(float2 &lo, float2 &hi)
{
    const float2 x = hi;
    const float2 a = lo;
    if (J == 2) {  // t = (x.y, -x.x)
        hi = c_sub_mrot(a, x);
        lo = c_add_mrot(a, x);
        return;
    }
    float2 t;
    if (J == 0) t = x;
    else if (J == 1) {
        const float2 u = c_add_mrot(x, x);  // (x.x + x.y, x.y - x.x)
        t = make_float2(u.x * OPN_FRAC_1_SQRT_2, u.y * OPN_FRAC_1_SQRT_2);
    } else {
        // (x.y - x.x, -(x.y + x.x)); negating both operands of a sum negates it exactly, signed zeros included
        const float2 u = p_add(make_float2(x.y, -x.y), make_float2(-x.x, -x.x));
        t = make_float2(u.x * OPN_FRAC_1_SQRT_2, u.y * OPN_FRAC_1_SQRT_2);
    }
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
    d2 = c_add_mrot(dm, s0);
    d1 = c_sub_mrot(dm, s0);
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
    // s5 = s0 + (s7 ya.x + s8 yb.x);  s6 = (s10.y ya.y + s9.y yb.y, -(s10.x ya.y + s9.x yb.y))
    const float2 s5 = p_add(s0, p_add(make_float2(s7.x * ya.x, s7.y * ya.x), make_float2(s8.x * yb.x, s8.y * yb.x)));
    const float2 q6 = p_add(make_float2(s10.y * ya.y, s10.x * ya.y), make_float2(s9.y * yb.y, s9.x * yb.y));  // (s6.x, -s6.y)
    d1 = p_add(s5, make_float2(-q6.x, q6.y));   // s5 - s6
    d4 = p_add(s5, make_float2(q6.x, -q6.y));   // s5 + s6
    // s11 = s0 + (s7 yb.x + s8 ya.x);  s12 = (s9.y ya.y - s10.y yb.y, s10.x yb.y - s9.x ya.y)
    const float2 s11 = p_add(s0, p_add(make_float2(s7.x * yb.x, s7.y * yb.x), make_float2(s8.x * ya.x, s8.y * ya.x)));
    const float2 s12 = p_add(make_float2(s9.y * ya.y, s10.x * yb.y), make_float2(-(s10.y * yb.y), -(s9.x * ya.y)));
    d2 = c_add(s11, s12);
    d3 = c_sub(s11, s12);
}

// radix-4 butterfly U of the last in-group stage: elements U + M*{0,1,2,3}, twiddle stride S
template <int U, int M, int S> __device__ __forceinline__ void w_bfly4_const(float2 *d)
{
    r_bfly4(d[U], d[U + M], d[U + 2 * M], d[U + 3 * M], WTw<S * U>::get(), WTw<2 * S * U>::get(), WTw<3 * S * U>::get());
}

// Half of the radix-4 butterfly U of the last in-group stage (m = 8, twiddle stride 15) when a group is split over two lanes
// (see w_imdct, pass A for one channel): this lane holds d[U] and d[8 + U].
template <int U> __device__ __forceinline__ void w_split_bfly4(float2 *d, int h)
{
    const float2 wy = h ? WTw<45 * U>::get() : WTw<30 * U>::get();
    const float2 xm = c_mul(d[U], WTw<15 * U>::get());
    const float2 X = h ? xm : d[U];
    const float2 Y = c_mul(d[8 + U], wy);
    const float2 S = c_add(X, Y), D = c_sub(X, Y);
    const float2 snd = h ? S : D;
    float2 rcv;
    rcv.x = __shfl_xor_sync(0x3FFFFFFFu, snd.x, 1);
    rcv.y = __shfl_xor_sync(0x3FFFFFFFu, snd.y, 1);
    const float2 P = h ? rcv : S;
    const float2 Q = h ? make_float2(D.y, -D.x) : rcv;
    d[U] = c_add(P, Q);
    d[8 + U] = c_sub(P, Q);
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
// TS: the tables (trig pairs, twiddles, window) are in shared memory (frame kernel) rather than global memory.
template <bool TS, class T> __device__ __forceinline__ T tab_ld(const T *p)
{
    if constexpr (TS) return *p;
    else return __ldg(p);
}
// BM: the coefficients of a transient frame arrive block-major (bin k of short block b at b * N2 + k) instead of
// interleaved (b + NBLK * k, the bitstream's order).  The frame kernel's expansion writes them that way: pass A reads bin
// 2r + 30q of every block, which in the interleaved layout is a stride of 16 words between lanes -- two banks for fifteen
// lanes -- and in the block-major one a stride of two.
template <int SHIFT, int NBLK, int C, bool TS = false, bool BM = false>
__device__ __forceinline__ void w_imdct(float *o, int lane, float4 carry, const float2 *tpair, const float2 *tw, const float *win)
{
    constexpr int GS = 32 >> SHIFT, N2 = 960 >> SHIFT;
    auto cin = [](int blk, int k) { return BM ? blk * N2 + k : blk + NBLK * k; };  // where input bin k of block blk lives
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
    if constexpr (SHIFT == 0 && NBLK == 1 && C == 1) {
        // One channel, TWO lanes per group (30 lanes): lane (g, h) owns the octets h and h+2 of group g, 16 elements
        // instead of 32 -- half the registers.  The stages with m = 1 and m = 4 stay inside an octet.  The last one
        // (radix 4, m = 8) combines element U of the four octets: lane 0 holds d0, d2, lane 1 holds d1, d3 of
        // kiss_fft.rs:148-187.  Each lane forms S = X + Y, D = X - Y of its own pair (lane 0: X = d0, Y = d2 w2, so
        // S = a0, D = s5; lane 1: X = d1 w1, Y = d3 w3, so S = s3, D = s4), they swap ONE complex value (lane 0 sends s5
        // and gets s3, lane 1 the reverse) and finish: lane 0 writes a0 +- s3 (outputs 0, 2), lane 1 s5 +- (s4.y, -s4.x)
        // (outputs 1, 3) -- again the octets h and h+2.  Same operations on the same values as one lane doing it all.
        const bool on = lane < 30;
        const int g = on ? lane >> 1 : 0, h = lane & 1;
        const int j1 = g / 3, j2 = g - 3 * j1, r = j1 + 5 * j2;
        const float *x0p = o + 2 * r + 30 * h, *x1p = o + (N2 - 1 - 2 * r) - 30 * h;
        const float2 *tp = tpair + r + 15 * h;
        float2 d[16];  // d[8e + pl]: position p = 8 (h + 2e) + pl of the group
        if (on) {
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
                for (int pl = 0; pl < 8; pl++) {  // pre-rotation, mdct.rs:184-200; q = w_qmap<0>(p) = q0 + h
                    const int q0 = 2 * e + 4 * ((pl >> 2) & 1) + 8 * (pl & 3);
                    const float x0 = x0p[30 * q0], x1 = x1p[-30 * q0];
                    const float2 t = tab_ld<TS>(tp + 15 * q0);
                    d[8 * e + pl] = p_add(make_float2(x0 * t.x, x1 * t.x), make_float2(-(x1 * t.y), x0 * t.y));
                }
        }
        __syncwarp();  // every coefficient is in a register: the row may be overwritten
        if (on) {
#pragma unroll
            for (int b = 0; b < 4; b++) r_bfly4_m1(d[4 * b], d[4 * b + 1], d[4 * b + 2], d[4 * b + 3]);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                r_bfly2<0>(d[8 * e + 0], d[8 * e + 4]);
                r_bfly2<1>(d[8 * e + 1], d[8 * e + 5]);
                r_bfly2<2>(d[8 * e + 2], d[8 * e + 6]);
                r_bfly2<3>(d[8 * e + 3], d[8 * e + 7]);
            }
            w_split_bfly4<0>(d, h);
            w_split_bfly4<1>(d, h);
            w_split_bfly4<2>(d, h);
            w_split_bfly4<3>(d, h);
            w_split_bfly4<4>(d, h);
            w_split_bfly4<5>(d, h);
            w_split_bfly4<6>(d, h);
            w_split_bfly4<7>(d, h);
            float2 *xc = reinterpret_cast<float2 *>(o) + 33 * g + 8 * h;  // a + a / PADG with a = 32 g + p
#pragma unroll
            for (int U = 0; U < 8; U++) {
                xc[U] = d[U];
                xc[16 + U] = d[8 + U];
            }
        }
    } else if constexpr (NBLK > 1 && C == 1) {
        // One channel of a transient frame: the NBLK x 15 groups of four are dealt to the lanes (up to four per lane)
        // instead of one lane holding a group of every block.
        constexpr int NT = NBLK * 15, TPL = (NT + 31) / 32;
        static_assert(SHIFT == 3, "short blocks are 120-bin transforms");
        float2 d[TPL * 4];
#pragma unroll
        for (int i = 0; i < TPL; i++) {
            const int t = lane + 32 * i;
            if (t < NT) {
                const int blk = t / 15, g = t - 15 * blk;
                const int j1 = g / 3, j2 = g - 3 * j1, r = j1 + 5 * j2;
#pragma unroll
                for (int p = 0; p < 4; p++) {  // w_qmap<3>(p) = p
                    const float x0 = o[cin(blk, 2 * r + 30 * p)];
                    const float x1 = o[cin(blk, N2 - 1 - 2 * r - 30 * p)];
                    const float2 tt = tab_ld<TS>(tpair + r + 15 * p);
                    d[4 * i + p] = p_add(make_float2(x0 * tt.x, x1 * tt.x), make_float2(-(x1 * tt.y), x0 * tt.y));
                }
            }
        }
        __syncwarp();
        float2 *xc = reinterpret_cast<float2 *>(o);
#pragma unroll
        for (int i = 0; i < TPL; i++) {
            const int t = lane + 32 * i;
            if (t < NT) {
                const int blk = t / 15, g = t - 15 * blk;
                r_bfly4_m1(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
                const int a = NFFT * blk + GS * g;
#pragma unroll
                for (int p = 0; p < 4; p++) xc[a + a / PADG + p] = d[4 * i + p];
            }
        }
    } else {
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
                    const float x0 = x[cin(blk, 2 * r + 30 * q)];
                    const float x1 = x[cin(blk, N2 - 1 - 2 * r - 30 * q)];
                    const float2 t = tab_ld<TS>(tp + 15 * q);  // (trig[i], trig[n4 + i])
                    // re = (x1 t.x) + (x0 t.y), im = (x0 t.x) - (x1 t.y); the element is (im, re)
                    d[blk * GS + p] = p_add(make_float2(x0 * t.x, x1 * t.x), make_float2(-(x1 * t.y), x0 * t.y));
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
        const float2 w31 = tab_ld<TS>(tw + u * S3), w32 = tab_ld<TS>(tw + 2 * u * S3);
        float2 w5[3][4];
#pragma unroll
        for (int jj = 0; jj < 3; jj++)
#pragma unroll
            for (int k = 0; k < 4; k++) w5[jj][k] = tab_ld<TS>(tw + (((u + GS * jj) * (k + 1)) << SHIFT));
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
                    const float2 t = tab_ld<TS>(tp + GS * j);
                    // out[2k] = (d.y t.x) + (d.x t.y), out[n2-1-2k] = (d.y t.y) - (d.x t.x)
                    const float2 r = p_add(make_float2(d[j].y * t.x, d[j].y * t.y), make_float2(d[j].x * t.y, -(d[j].x * t.x)));
                    ob[2 * u + 2 * GS * j] = r.x;
                    ob[N2 - 1 - 2 * u - 2 * GS * j] = r.y;
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
        const float w0 = tab_ld<TS>(win + i), w1 = tab_ld<TS>(win + 119 - i);
        const float2 r = p_add(make_float2(w1 * x1, w0 * x1), make_float2(-(w0 * x0), w1 * x0));
        ob[i] = r.x;        // (w1 x1) - (w0 x0)
        ob[119 - i] = r.y;  // (w0 x1) + (w1 x0)
    }
    __syncwarp();
}

}  // namespace opn
