/* CPU ORACLE (test infrastructure) -- mixed-radix FFT and (I)MDCT.
 * Restates /root/reference/src/celt/kiss_fft.rs:13-279 and src/celt/mdct.rs:17-260.
 * Complex arithmetic follows src/math.rs:79-158 (mul = 4 mul + 2 add, never fused). */
#include "oracle.h"
#include "oracle_tables.h"

typedef struct { float r, i; } cpx;

static inline cpx c_add(cpx a, cpx b) { cpx o = {a.r + b.r, a.i + b.i}; return o; }
static inline cpx c_sub(cpx a, cpx b) { cpx o = {a.r - b.r, a.i - b.i}; return o; }
/* math.rs:115-124 */
static inline cpx c_mul(cpx a, cpx b) { cpx o = {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; return o; }
static inline cpx c_scale(cpx a, float s) { cpx o = {a.r * s, a.i * s}; return o; }

#define FRAC_1_SQRT_2 0.707106781186547524400844362104849039f

static inline cpx tw(size_t idx) { cpx o = {ORC_TWIDDLES[2 * idx], ORC_TWIDDLES[2 * idx + 1]}; return o; }

static const uint8_t *fft_factors(int shift)
{
    switch (shift) {
    case 0: return ORC_FFT_FACTORS_480;
    case 1: return ORC_FFT_FACTORS_240;
    case 2: return ORC_FFT_FACTORS_120;
    default: return ORC_FFT_FACTORS_60;
    }
}

const uint16_t *orc_fft_bitrev(int shift)
{
    switch (shift) {
    case 0: return ORC_BITREV_480;
    case 1: return ORC_BITREV_240;
    case 2: return ORC_BITREV_120;
    default: return ORC_BITREV_60;
    }
}

/* kiss_fft.rs:249,257,265,273 */
float orc_fft_scale(int shift)
{
    static const float s[4] = {0.002083333f, 0.004166667f, 0.008333333f, 0.016666667f};
    return s[shift];
}

/* kiss_fft.rs:55-87 */
static void bfly2(cpx *d, size_t m, size_t n)
{
    (void)m; /* m == 4 */
    size_t off = 0;
    for (size_t k = 0; k < n; k++) {
        size_t o2 = off + 4;
        cpx t = d[o2];
        d[o2] = c_sub(d[off], t);
        d[off] = c_add(d[off], t);

        t.r = (d[o2 + 1].r + d[o2 + 1].i) * FRAC_1_SQRT_2;
        t.i = (d[o2 + 1].i - d[o2 + 1].r) * FRAC_1_SQRT_2;
        d[o2 + 1] = c_sub(d[off + 1], t);
        d[off + 1] = c_add(d[off + 1], t);

        t.r = d[o2 + 2].i;
        t.i = -d[o2 + 2].r;
        d[o2 + 2] = c_sub(d[off + 2], t);
        d[off + 2] = c_add(d[off + 2], t);

        t.r = (d[o2 + 3].i - d[o2 + 3].r) * FRAC_1_SQRT_2;
        t.i = (-(d[o2 + 3].i + d[o2 + 3].r)) * FRAC_1_SQRT_2;
        d[o2 + 3] = c_sub(d[off + 3], t);
        d[off + 3] = c_add(d[off + 3], t);
        off += 8;
    }
}

/* kiss_fft.rs:89-127 */
static void bfly3(cpx *d, size_t stride, size_t m, size_t n, size_t mm)
{
    size_t m2 = 2 * m;
    cpx epi3 = tw(stride * m);
    for (size_t i = 0; i < n; i++) {
        size_t off = i * mm, t1 = 0, t2 = 0;
        for (size_t k = m; k > 0; k--) {
            cpx s1 = c_mul(d[off + m], tw(t1));
            cpx s2 = c_mul(d[off + m2], tw(t2));
            cpx s3 = c_add(s1, s2);
            cpx s0 = c_sub(s1, s2);
            t1 += stride;
            t2 += stride * 2;
            d[off + m] = c_sub(d[off], c_scale(s3, 0.5f));
            s0 = c_scale(s0, epi3.i);
            d[off] = c_add(d[off], s3);
            d[off + m2].r = d[off + m].r + s0.i;
            d[off + m2].i = d[off + m].i - s0.r;
            d[off + m].r -= s0.i;
            d[off + m].i += s0.r;
            off += 1;
        }
    }
}

/* kiss_fft.rs:129-188 */
static void bfly4(cpx *d, size_t stride, size_t m, size_t n, size_t mm)
{
    if (m == 1) {
        size_t off = 0;
        for (size_t k = 0; k < n; k++) {
            cpx s0 = c_sub(d[off], d[off + 2]);
            cpx s1 = c_add(d[off + 1], d[off + 3]);
            d[off] = c_add(d[off], d[off + 2]);
            d[off + 2] = c_sub(d[off], s1);
            d[off] = c_add(d[off], s1);
            s1 = c_sub(d[off + 1], d[off + 3]);
            d[off + 1].r = s0.r + s1.i;
            d[off + 1].i = s0.i - s1.r;
            d[off + 3].r = s0.r - s1.i;
            d[off + 3].i = s0.i + s1.r;
            off += 4;
        }
        return;
    }
    size_t m2 = 2 * m, m3 = 3 * m;
    for (size_t i = 0; i < n; i++) {
        size_t off = i * mm, t1 = 0, t2 = 0, t3 = 0;
        for (size_t k = 0; k < m; k++) {
            cpx s0 = c_mul(d[off + m], tw(t1));
            cpx s1 = c_mul(d[off + m2], tw(t2));
            cpx s2 = c_mul(d[off + m3], tw(t3));
            cpx s5 = c_sub(d[off], s1);
            d[off] = c_add(d[off], s1);
            cpx s3 = c_add(s0, s2);
            cpx s4 = c_sub(s0, s2);
            d[off + m2] = c_sub(d[off], s3);
            t1 += stride;
            t2 += stride * 2;
            t3 += stride * 3;
            d[off] = c_add(d[off], s3);
            d[off + m].r = s5.r + s4.i;
            d[off + m].i = s5.i - s4.r;
            d[off + m3].r = s5.r - s4.i;
            d[off + m3].i = s5.i + s4.r;
            off += 1;
        }
    }
}

/* kiss_fft.rs:190-243 */
static void bfly5(cpx *d, size_t stride, size_t m, size_t n, size_t mm)
{
    cpx ya = tw(stride * m), yb = tw(stride * 2 * m);
    for (size_t i = 0; i < n; i++) {
        size_t o0 = i * mm, o1 = o0 + m, o2 = o0 + 2 * m, o3 = o0 + 3 * m, o4 = o0 + 4 * m;
        for (size_t u = 0; u < m; u++) {
            cpx s0 = d[o0];
            cpx s1 = c_mul(d[o1], tw(u * stride));
            cpx s2 = c_mul(d[o2], tw(2 * u * stride));
            cpx s3 = c_mul(d[o3], tw(3 * u * stride));
            cpx s4 = c_mul(d[o4], tw(4 * u * stride));
            cpx s7 = c_add(s1, s4), s10 = c_sub(s1, s4);
            cpx s8 = c_add(s2, s3), s9 = c_sub(s2, s3);
            d[o0] = c_add(d[o0], c_add(s7, s8));
            cpx s5, s6, s11, s12;
            s5.r = s0.r + (s7.r * ya.r + s8.r * yb.r);
            s5.i = s0.i + (s7.i * ya.r + s8.i * yb.r);
            s6.r = s10.i * ya.i + s9.i * yb.i;
            s6.i = -(s10.r * ya.i + s9.r * yb.i);
            d[o1] = c_sub(s5, s6);
            d[o4] = c_add(s5, s6);
            s11.r = s0.r + (s7.r * yb.r + s8.r * ya.r);
            s11.i = s0.i + (s7.i * yb.r + s8.i * ya.r);
            s12.r = s9.i * ya.i - s10.i * yb.i;
            s12.i = s10.r * yb.i - s9.r * ya.i;
            d[o2] = c_add(s11, s12);
            d[o3] = c_sub(s11, s12);
            o0++, o1++, o2++, o3++, o4++;
        }
    }
}

/* kiss_fft.rs:24-53 */
void orc_fft_process(int shift, float *data_f)
{
    cpx *data = (cpx *)data_f;
    const uint8_t *factors = fft_factors(shift);
    size_t strides[8];
    strides[0] = 1;
    size_t m = 0, l = 0;
    while (m != 1) {
        size_t p = factors[2 * l];
        m = factors[2 * l + 1];
        strides[l + 1] = strides[l] * p;
        l += 1;
    }
    m = factors[2 * l - 1];
    for (size_t i = l; i-- > 0;) {
        size_t m2 = i != 0 ? factors[2 * i - 1] : 1;
        size_t stride = strides[i] << shift;
        switch (factors[2 * i]) {
        case 2: bfly2(data, m, strides[i]); break;
        case 4: bfly4(data, stride, m, strides[i], m2); break;
        case 3: bfly3(data, stride, m, strides[i], m2); break;
        case 5: bfly5(data, stride, m, strides[i], m2); break;
        default: break;
        }
        m = m2;
    }
}

const float *orc_window(void) { return ORC_WINDOW; }
const float *orc_trig(void) { return ORC_TRIG; }

/* mdct.rs:159-260 */
void orc_mdct_backward(const float *input, float *output, const float *window, int overlap,
                       int shift, int stride)
{
    int n = ORC_MDCT_N, trigp = 0;
    for (int s = 0; s < shift; s++) {
        n >>= 1;
        trigp += n;
    }
    int n2 = n >> 1, n4 = n >> 2;
    cpx spc[480];
    const uint16_t *bitrev = orc_fft_bitrev(shift);
    const float *trig = ORC_TRIG + trigp;

    /* pre-rotation (mdct.rs:184-200) */
    {
        long ip0 = 0, ip1 = (long)stride * (n2 - 1);
        for (int i = 0; i < n4; i++) {
            float re = (input[ip1] * trig[i]) + (input[ip0] * trig[n4 + i]);
            float im = (input[ip0] * trig[i]) - (input[ip1] * trig[n4 + i]);
            spc[bitrev[i]].r = im;
            spc[bitrev[i]].i = re;
            ip0 += 2 * stride;
            ip1 -= 2 * stride;
        }
    }
    orc_fft_process(shift, (float *)spc);
    /* post-rotate and de-shuffle (mdct.rs:205-238) */
    {
        int ho = overlap >> 1;
        for (int i = 0; i < n4; i++) {
            cpx c = spc[i];
            output[ho + 2 * i] = (c.i * trig[i]) + (c.r * trig[n4 + i]);
        }
        for (int i = 0; i < n4; i++) {
            cpx c = spc[n4 - i - 1];
            float t0 = trig[n4 - i - 1], t1 = trig[n2 - i - 1];
            output[ho + 1 + 2 * i] = (c.i * t1) - (c.r * t0);
        }
    }
    /* TDAC mirror (mdct.rs:241-259) */
    {
        int op0 = 0, op1 = overlap - 1, wp0 = 0, wp1 = overlap - 1;
        for (int i = 0; i < overlap / 2; i++) {
            float x0 = output[op1], x1 = output[op0];
            output[op0] = (window[wp1] * x1) - (window[wp0] * x0);
            output[op1] = (window[wp0] * x1) + (window[wp1] * x0);
            op0++, op1--, wp0++, wp1--;
        }
    }
}

/* mdct.rs:37-156 (encoder side; used by tests and to build spectra of known signals) */
void orc_mdct_forward(const float *input, float *output, const float *window, int overlap,
                      int shift, int stride)
{
    int n = ORC_MDCT_N, trigp = 0;
    for (int s = 0; s < shift; s++) {
        n >>= 1;
        trigp += n;
    }
    int n2 = n >> 1, n4 = n >> 2;
    float spf[960];
    cpx spc[480];
    const uint16_t *bitrev = orc_fft_bitrev(shift);
    const float *trig = ORC_TRIG + trigp;
    float scale = orc_fft_scale(shift);
    {
        int overlap_offset = (overlap + 3) >> 2;
        long ip0 = overlap >> 1, ip1 = ip0 + n2 - 1, sp = 0, wp0 = overlap >> 1, wp1 = wp0 - 1;
        for (int i = 0; i < overlap_offset; i++) {
            spf[sp] = (window[wp1] * input[ip0 + n2]) + (window[wp0] * input[ip1]);
            spf[sp + 1] = (window[wp0] * input[ip0]) - (window[wp1] * input[ip1 - n2]);
            sp += 2, ip0 += 2, ip1 -= 2, wp0 += 2, wp1 -= 2;
        }
        wp0 = 0;
        wp1 = overlap - 1;
        for (int i = overlap_offset; i < n4 - overlap_offset; i++) {
            spf[sp] = input[ip1];
            spf[sp + 1] = input[ip0];
            sp += 2, ip0 += 2, ip1 -= 2;
        }
        for (int i = n4 - overlap_offset; i < n4; i++) {
            spf[sp] = -(window[wp0] * input[ip0 - n2]) + (window[wp1] * input[ip1]);
            spf[sp + 1] = (window[wp1] * input[ip0]) + (window[wp0] * input[ip1 + n2]);
            sp += 2, ip0 += 2, ip1 -= 2, wp0 += 2, wp1 -= 2;
        }
    }
    for (int i = 0; i < n4; i++) {
        float t0 = trig[i], t1 = trig[n4 + i];
        float re = spf[2 * i], im = spf[2 * i + 1];
        cpx t;
        t.r = (re * t0) - (im * t1);
        t.i = (im * t0) + (re * t1);
        t.r = t.r * scale;
        t.i = t.i * scale;
        spc[bitrev[i]] = t;
    }
    orc_fft_process(shift, (float *)spc);
    {
        long op0 = 0, op1 = (long)stride * (n2 - 1);
        for (int i = 0; i < n4; i++) {
            output[op0] = (spc[i].i * trig[n4 + i]) - (spc[i].r * trig[i]);
            output[op1] = (spc[i].r * trig[n4 + i]) + (spc[i].i * trig[i]);
            op0 += 2 * stride;
            op1 -= 2 * stride;
        }
    }
}
