// imdct_warp.cuh -- kernel 1, second generation: one WARP decodes one stream (both channels).
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
//   * A warp never waits for another warp: all hand-offs are __syncwarp(), CTAs are single warps,
//     13 streams are resident per SM and slip past each other (memory phases of one stream overlap
//     FP32 phases of the others).
//   * Coefficient rows arrive by TMA (cp.async.bulk, one 3840-byte row per channel) into the
//     region the output row later occupies; the comb history is staged from the interleaved PCM
//     ring with float4 loads into the region the transpose used.
//
// Shared memory per channel (floats):  [ A: 1024 | O: nf + 60 ]
//   A = transpose buffer (float2, 15*(GS+1) per block, padded so both passes are conflict-free),
//       later y[-1024 .. -1] (post-filter history; ends exactly where O starts, so a tap at any
//       signed index is one address)
//   O = coefficient row (TMA destination), later out[0 .. nf+60): carry-in/TDAC, PCM, carry-out.
#pragma once
#include "imdct.cuh"

namespace opn {

// Twiddles addressed with compile-time indices (src/celt/kiss_fft.rs:341-582): constant bank.
__constant__ float2 c_tw[480];

constexpr int WA_FLOATS = 1024;  // region A
__host__ __device__ constexpr int w_ch_floats(int lm) { return WA_FLOATS + (120 << lm) + 60; }
__host__ __device__ constexpr size_t w_smem_bytes(int lm, int channels) { return (size_t)channels * w_ch_floats(lm) * 4 + 16; }
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
template <int SHIFT> __device__ __forceinline__ void w_group_stages(float2 (&d)[32 >> SHIFT])
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
#pragma unroll
        for (int u = 0; u < 8; u++) r_bfly4(d[u], d[u + 8], d[u + 16], d[u + 24], c_tw[15 * u], c_tw[30 * u], c_tw[45 * u]);
    }
    if constexpr (SHIFT == 1) {  // radix 4, m = 4, twiddle stride 30
#pragma unroll
        for (int u = 0; u < 4; u++) r_bfly4(d[u], d[u + 4], d[u + 8], d[u + 12], c_tw[30 * u], c_tw[60 * u], c_tw[90 * u]);
    }
}

// Mdct::backward for the C channels of one stream, NBLK interleaved blocks of nfft = 480 >> SHIFT
// (NBLK == 1: one long block; SHIFT == 3 and NBLK == 2^LM: transient frame).  On entry O holds the
// coefficient rows; on exit O holds out[0 .. nf+60) after the TDAC mirror (mdct.rs:241-259).
// `carry` is this lane's float4 of the previous tail (lane = 15*ch + k -> out[4k .. 4k+4)).
template <int SHIFT, int NBLK, int C> __device__ __forceinline__ void w_imdct(float *sm, int lane, float4 carry)
{
    constexpr int GS = 32 >> SHIFT, N2 = 960 >> SHIFT;
    constexpr int NF = N2 * NBLK, E = GS * NBLK;
    constexpr int CHF = WA_FLOATS + NF + 60;
    // Transpose buffer: position pos of a block lives at pos + pos / PADG (one float2 of padding per
    // PADG positions keeps the stride-GS writes of pass A and the unit-stride reads of pass B on
    // distinct banks).  Eight short blocks only fit region A with the coarser padding.
    constexpr int PADG = (GS == 4 && NBLK == 8) ? 16 : GS;
    constexpr int XBLK = 15 * GS + (15 * GS + PADG - 1) / PADG;  // float2 per block
    static_assert(2 * XBLK * NBLK <= WA_FLOATS, "transpose buffer must fit region A");
    const float2 *tpair = g_tab.trig_pair + trig_pair_off(SHIFT);

    // ---------------------------------------------------------------- pass A
    if (lane < 15 * C) {
        const int ch = (C == 2 && lane >= 15) ? 1 : 0;
        const int g = lane - 15 * ch;
        const int j1 = g / 3, j2 = g - 3 * j1, r = j1 + 5 * j2;
        const float *x = sm + ch * CHF + WA_FLOATS;
        float2 *xc = reinterpret_cast<float2 *>(sm + ch * CHF) + GS * g + (GS * g) / PADG;  // p < GS <= PADG adds no pad
        const float2 *tp = tpair + r;
#pragma unroll
        for (int blk = 0; blk < NBLK; blk++) {
            float2 d[GS];
#pragma unroll
            for (int p = 0; p < GS; p++) {  // pre-rotation, mdct.rs:184-200
                const int q = w_qmap<SHIFT>(p);
                const float x0 = x[blk + NBLK * (2 * r) + NBLK * 30 * q];
                const float x1 = x[blk + NBLK * (N2 - 1 - 2 * r) - NBLK * 30 * q];
                const float2 t = __ldg(tp + 15 * q);  // (trig[i], trig[n4 + i])
                const float re = (x1 * t.x) + (x0 * t.y);
                const float im = (x0 * t.x) - (x1 * t.y);
                d[p] = make_float2(im, re);
            }
            w_group_stages<SHIFT>(d);
#pragma unroll
            for (int p = 0; p < GS; p++) xc[blk * XBLK + p] = d[p];
        }
    }
    __syncwarp();
    // coefficient rows are consumed: out[0..60) <- previous tail
    if (lane < 15 * C) {
        const int ch = (C == 2 && lane >= 15) ? 1 : 0;
        *reinterpret_cast<float4 *>(sm + ch * CHF + WA_FLOATS + 4 * (lane - 15 * ch)) = carry;
    }
    // ---------------------------------------------------------------- pass B
    {
        constexpr int S3 = 5 << SHIFT;  // twiddle stride of the radix-3 stage
        const int col = lane % E, ch0 = lane / E;
        const int blk = col / GS, u = col % GS;
        const float2 *tw = g_tab.twiddles;
        const float2 w31 = __ldg(tw + u * S3), w32 = __ldg(tw + 2 * u * S3);
        float2 w5[3][4];
#pragma unroll
        for (int jj = 0; jj < 3; jj++)
#pragma unroll
            for (int k = 0; k < 4; k++) w5[jj][k] = __ldg(tw + (((u + GS * jj) * (k + 1)) << SHIFT));
        const float epi3y = c_tw[160].y;
        const float2 ya = c_tw[96], yb = c_tw[192];
        constexpr int ITER = (C * E + 31) / 32;
#pragma unroll
        for (int it = 0; it < ITER; it++) {
            const int ch = ch0 + it * (32 / E);
            if (ch < C) {
                const float2 *xc = reinterpret_cast<const float2 *>(sm + ch * CHF) + blk * XBLK + u;
                float2 d[15];
#pragma unroll
                for (int j = 0; j < 15; j++) d[j] = xc[GS * j + (GS * j) / PADG];  // u < GS adds no pad
#pragma unroll
                for (int a = 0; a < 5; a++) r_bfly3(d[3 * a], d[3 * a + 1], d[3 * a + 2], w31, w32, epi3y);
#pragma unroll
                for (int jj = 0; jj < 3; jj++)
                    r_bfly5(d[jj], d[jj + 3], d[jj + 6], d[jj + 9], d[jj + 12], w5[jj][0], w5[jj][1], w5[jj][2], w5[jj][3], ya, yb);
                // post-rotation and de-shuffle, mdct.rs:205-238: FFT output k = u + GS*j
                float *o = sm + ch * CHF + WA_FLOATS + N2 * blk + 60;
                const float2 *tp = tpair + u;
#pragma unroll
                for (int j = 0; j < 15; j++) {
                    const float2 t = __ldg(tp + GS * j);
                    o[2 * u + 2 * GS * j] = (d[j].y * t.x) + (d[j].x * t.y);
                    o[N2 - 1 - 2 * u - 2 * GS * j] = (d[j].y * t.y) - (d[j].x * t.x);
                }
            }
        }
    }
    __syncwarp();
    // ---------------------------------------------------------------- TDAC mirror, mdct.rs:241-259
    for (int w = lane; w < C * NBLK * 60; w += 32) {
        const int ch = w / (NBLK * 60), rem = w - ch * (NBLK * 60);
        const int blk = rem / 60, i = rem - blk * 60;
        float *o = sm + ch * CHF + WA_FLOATS + N2 * blk;
        const float x0 = o[119 - i], x1 = o[i];
        const float w0 = __ldg(&g_tab.window[i]), w1 = __ldg(&g_tab.window[119 - i]);
        o[i] = (w1 * x1) - (w0 * x0);
        o[119 - i] = (w0 * x1) + (w1 * x0);
    }
    __syncwarp();
}

// comb_filter_inplace (comb_filter/mod.rs:130-193) for the C channels of one stream; channel ch's
// samples are y + ch*chf, its history directly below.  Recursive filter: the warp sweeps the frame
// in chunks of W = min(period) - 2 <= 32 samples, inside which every tap lies before the chunk.
template <int C>
__device__ __forceinline__ void w_comb(float *y, int chf, int t0, int t1, int n, float g0, float g1, int tap0, int tap1, int overlap,
                                       int lane)
{
    if (g0 == 0.0f && g1 == 0.0f) return;
    t0 = max(t0, 15);
    t1 = max(t1, 15);
    const float g00 = g0 * g_tab.comb_gains[tap0 * 3], g01 = g0 * g_tab.comb_gains[tap0 * 3 + 1],
                g02 = g0 * g_tab.comb_gains[tap0 * 3 + 2];
    const float g10 = g1 * g_tab.comb_gains[tap1 * 3], g11 = g1 * g_tab.comb_gains[tap1 * 3 + 1],
                g12 = g1 * g_tab.comb_gains[tap1 * 3 + 2];
    if (fabsf(g0 - g1) < 1.1920929e-7f && t0 == t1 && tap0 == tap1) overlap = 0;
    {  // cross-fade part (mod.rs:162-179)
        const int W = min(min(t0, t1) - 2, 32);
        for (int base = 0; base < overlap; base += W) {
            const int i = base + lane;
            if (lane < W && i < overlap) {
                const float f = __ldg(&g_tab.window_sq[i]);
                const float a0 = (1.0f - f) * g00, a1 = (1.0f - f) * g01, a2 = (1.0f - f) * g02;
                const float b0 = f * g10, b1 = f * g11, b2 = f * g12;
#pragma unroll
                for (int c = 0; c < C; c++) {
                    float *yc = y + c * chf;
                    const float x0 = yc[i - t1 + 2], x1 = yc[i - t1 + 1], x2 = yc[i - t1], x3 = yc[i - t1 - 1], x4 = yc[i - t1 - 2];
                    float acc = yc[i];
                    acc = acc + (a0 * yc[i - t0]);
                    acc = acc + (a1 * (yc[i - t0 + 1] + yc[i - t0 - 1]));
                    acc = acc + (a2 * (yc[i - t0 + 2] + yc[i - t0 - 2]));
                    acc = acc + (b0 * x2);
                    acc = acc + (b1 * (x1 + x3));
                    acc = acc + (b2 * (x0 + x4));
                    yc[i] = acc;
                }
            }
            __syncwarp();
        }
    }
    if (g1 == 0.0f) return;
    {  // constant part (fallback.rs:32-53)
        const int W = min(t1 - 2, 32);
        for (int base = overlap; base < n; base += W) {
            const int i = base + lane;
            if (lane < W && i < n) {
#pragma unroll
                for (int c = 0; c < C; c++) {
                    float *yc = y + c * chf;
                    const float x0 = yc[i - t1 + 2], x1 = yc[i - t1 + 1], x2 = yc[i - t1], x3 = yc[i - t1 - 1], x4 = yc[i - t1 - 2];
                    yc[i] = yc[i] + (g10 * x2) + (g11 * (x1 + x3)) + (g12 * (x0 + x4));
                }
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 1: grid = items (streams of this bucket), one warp per CTA.
template <int LM, int C> __global__ void __launch_bounds__(32) k_imdct_post_w(ImdctArgs A)
{
    extern __shared__ __align__(16) float sm[];
    constexpr int NF = 120 << LM;
    constexpr int CHF = w_ch_floats(LM);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + C * CHF);
    const int lane = threadIdx.x;

    const uint32_t item = blockIdx.x;
    const uint32_t stream = A.stream_idx ? A.stream_idx[item] : item;
    const int32_t status = A.status[stream];
    if (status < 0) {  // rejected packet: state untouched (decoder.rs:397)
        if (lane == 0 && A.result) A.result[stream] = status;
        return;
    }
    // coefficient rows -> region O by TMA
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, C * NF * 4);
#pragma unroll
        for (int c = 0; c < C; c++)
            bulk_g2s(sm + c * CHF + WA_FLOATS, A.coef + ((size_t)stream * C + c) * NF, NF * 4, bar);
    }
    const opn_synth_side *side = A.side + stream;
    const bool lost = status == ITEM_LOST;
    const int transient = side->transient;

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
    float *carry_g = A.carry + (size_t)stream * C * 60;
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < 15 * C) carry = *reinterpret_cast<const float4 *>(carry_g + 4 * lane);  // [C][60] = 15 float4 per channel

    __syncwarp();
    mbar_wait(bar, 0);
    if constexpr (LM > 0) {
        if (transient) w_imdct<3, (1 << LM), C>(sm, lane, carry);
        else w_imdct<3 - LM, 1, C>(sm, lane, carry);
    } else {
        w_imdct<3, 1, C>(sm, lane, carry);
    }

    // tail of this frame -> carry; post-filter history <- PCM ring (region A is free again)
    if (lane < 15 * C) {
        const int ch = (C == 2 && lane >= 15) ? 1 : 0;
        *reinterpret_cast<float4 *>(carry_g + 4 * lane) =
            *reinterpret_cast<const float4 *>(sm + ch * CHF + WA_FLOATS + NF + 4 * (lane - 15 * ch));
    }
    if (comb_on) {
        const int need = max(max(old.period, t1), 15) + 2;  // <= 1024
        if (C == 2) {
            // two samples x two channels per float4; pos is a multiple of 120, so pairs never straddle the wrap
            for (int j = lane; 2 * j < need; j += 32) {
                uint32_t p = pos + RING_SAMPLES - 2u - 2u * (uint32_t)j;
                if (p >= RING_SAMPLES) p -= RING_SAMPLES;
                const float4 v = *reinterpret_cast<const float4 *>(ring + (size_t)p * 2);
                float *h0 = sm + WA_FLOATS - 2 - 2 * j, *h1 = h0 + CHF;
                *reinterpret_cast<float2 *>(h0) = make_float2(v.x, v.z);
                *reinterpret_cast<float2 *>(h1) = make_float2(v.y, v.w);
            }
        } else {
            for (int j = lane; 4 * j < need; j += 32) {
                uint32_t p = pos + RING_SAMPLES - 4u - 4u * (uint32_t)j;
                if (p >= RING_SAMPLES) p -= RING_SAMPLES;
                *reinterpret_cast<float4 *>(sm + WA_FLOATS - 4 - 4 * j) = *reinterpret_cast<const float4 *>(ring + p);
            }
        }
        __syncwarp();
        w_comb<C>(sm + WA_FLOATS, CHF, old.period, t1, NF, old.gain, g1, old.tapset, tap1, 120, lane);
    }
    __syncwarp();

    // epilogue: interleaved PCM -> ring (history + device-resident output) and optional dense rows
    const float *s0 = sm + WA_FLOATS;
    const float *s1 = s0 + CHF;
    float *dense = A.dense ? A.dense + (size_t)stream * A.dense_stride + (A.dense_off ? A.dense_off[item] : 0u) : nullptr;
    const float gain = A.gain;
    if (C == 2) {
#pragma unroll 5
        for (int i = lane; i < NF / 2; i += 32) {
            const float2 a = *reinterpret_cast<const float2 *>(s0 + 2 * i), b = *reinterpret_cast<const float2 *>(s1 + 2 * i);
            float4 v = make_float4(a.x, b.x, a.y, b.y);
            uint32_t p = pos + 2u * (uint32_t)i;
            if (p >= RING_SAMPLES) p -= RING_SAMPLES;
            *reinterpret_cast<float4 *>(ring + (size_t)p * 2) = v;
            if (dense) {
                if (gain != 1.0f) { v.x *= gain; v.y *= gain; v.z *= gain; v.w *= gain; }
                *reinterpret_cast<float4 *>(dense + 4 * i) = v;
            }
        }
    } else {
        for (int i = lane; i < NF / 4; i += 32) {
            float4 v = *reinterpret_cast<const float4 *>(s0 + 4 * i);
            uint32_t p = pos + 4u * (uint32_t)i;
            if (p >= RING_SAMPLES) p -= RING_SAMPLES;
            *reinterpret_cast<float4 *>(ring + p) = v;
            if (dense) {
                if (gain != 1.0f) { v.x *= gain; v.y *= gain; v.z *= gain; v.w *= gain; }
                *reinterpret_cast<float4 *>(dense + 4 * i) = v;
            }
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
        if (A.final_range) A.final_range[stream] = lost ? 0u : side->final_rng;
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
    for (int i = lane; i < NF; i += 32) sm[WA_FLOATS + i] = in[i];
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < 15) carry = make_float4(out[4 * lane], out[4 * lane + 1], out[4 * lane + 2], out[4 * lane + 3]);
    __syncwarp();
    w_imdct<SHIFT, NBLK, 1>(sm, lane, carry);
    for (int i = lane; i < NF + 60; i += 32) out[i] = sm[WA_FLOATS + i];
}

}  // namespace opn
