/* CPU ORACLE (test infrastructure) -- pitch comb filter.
 * Restates /root/reference/src/celt/comb_filter/mod.rs:43-193 with the scalar kernels of
 * fallback.rs:6-53 (the SSE build, sse.rs, evaluates the same sums in the same order for
 * n % 4 == 0 and skips the n % 4 tail; the scalar semantics are the defining ones). */
#include "oracle.h"
#include "oracle_tables.h"

#include <math.h>
#include <string.h>

#define F32_EPSILON 1.1920929e-7f

/* fallback.rs:6-29 */
static void comb_const(float *y, size_t yo, const float *x, size_t xo, size_t t, size_t n,
                       float g10, float g11, float g12)
{
    float x4 = x[xo - t - 2], x3 = x[xo - t - 1], x2 = x[xo - t], x1 = x[xo - t + 1];
    for (size_t i = 0; i < n; i++) {
        float x0 = x[xo + i - t + 2];
        y[yo + i] = x[xo + i] + (g10 * x2) + (g11 * (x1 + x3)) + (g12 * (x0 + x4));
        x4 = x3, x3 = x2, x2 = x1, x1 = x0;
    }
}

/* fallback.rs:32-53 */
static void comb_const_inplace(float *y, size_t yo, size_t t, size_t n, float g10, float g11,
                               float g12)
{
    float x4 = y[yo - t - 2], x3 = y[yo - t - 1], x2 = y[yo - t], x1 = y[yo - t + 1];
    for (size_t i = 0; i < n; i++) {
        float x0 = y[yo + i - t + 2];
        y[yo + i] = y[yo + i] + (g10 * x2) + (g11 * (x1 + x3)) + (g12 * (x0 + x4));
        x4 = x3, x3 = x2, x2 = x1, x1 = x0;
    }
}

/* mod.rs:59-127 */
void orc_comb_filter(float *y, size_t yo, const float *x, size_t xo, size_t t0, size_t t1,
                     size_t n, float g0, float g1, size_t tapset0, size_t tapset1, size_t overlap)
{
    if (g0 == 0.0f && g1 == 0.0f) {
        memmove(y + yo, x + xo, n * sizeof(float));
        return;
    }
    if (t0 < ORC_COMB_MINPERIOD) t0 = ORC_COMB_MINPERIOD;
    if (t1 < ORC_COMB_MINPERIOD) t1 = ORC_COMB_MINPERIOD;
    const float *G = ORC_COMB_GAINS;
    float g00 = g0 * G[tapset0 * 3], g01 = g0 * G[tapset0 * 3 + 1], g02 = g0 * G[tapset0 * 3 + 2];
    float g10 = g1 * G[tapset1 * 3], g11 = g1 * G[tapset1 * 3 + 1], g12 = g1 * G[tapset1 * 3 + 2];
    float x1 = x[xo - t1 + 1], x2 = x[xo - t1], x3 = x[xo - t1 - 1], x4 = x[xo - t1 - 2];
    if (fabsf(g0 - g1) < F32_EPSILON && t0 == t1 && tapset0 == tapset1) overlap = 0;
    size_t j = 0;
    for (size_t i = 0; i < overlap; i++) {
        float x0 = x[xo + i - t1 + 2];
        float f = ORC_WINDOW[i] * ORC_WINDOW[i];
        y[yo + i] = x[xo + i]
            + (((1.0f - f) * g00) * x[xo + i - t0])
            + (((1.0f - f) * g01) * (x[xo + i - t0 + 1] + x[xo + i - t0 - 1]))
            + (((1.0f - f) * g02) * (x[xo + i - t0 + 2] + x[xo + i - t0 - 2]))
            + ((f * g10) * x2)
            + ((f * g11) * (x1 + x3))
            + ((f * g12) * (x0 + x4));
        x4 = x3, x3 = x2, x2 = x1, x1 = x0;
        j += 1;
    }
    if (g1 == 0.0f) {
        memmove(y + yo + overlap, x + xo + overlap, (n - overlap) * sizeof(float));
        return;
    }
    comb_const(y, yo + j, x, xo + j, t1, n - j, g10, g11, g12);
}

/* mod.rs:130-193 */
void orc_comb_filter_inplace(float *y, size_t yo, size_t t0, size_t t1, size_t n, float g0,
                             float g1, size_t tapset0, size_t tapset1, size_t overlap)
{
    if (g0 == 0.0f && g1 == 0.0f) return;
    if (t0 < ORC_COMB_MINPERIOD) t0 = ORC_COMB_MINPERIOD;
    if (t1 < ORC_COMB_MINPERIOD) t1 = ORC_COMB_MINPERIOD;
    const float *G = ORC_COMB_GAINS;
    float g00 = g0 * G[tapset0 * 3], g01 = g0 * G[tapset0 * 3 + 1], g02 = g0 * G[tapset0 * 3 + 2];
    float g10 = g1 * G[tapset1 * 3], g11 = g1 * G[tapset1 * 3 + 1], g12 = g1 * G[tapset1 * 3 + 2];
    float x1 = y[yo - t1 + 1], x2 = y[yo - t1], x3 = y[yo - t1 - 1], x4 = y[yo - t1 - 2];
    if (fabsf(g0 - g1) < F32_EPSILON && t0 == t1 && tapset0 == tapset1) overlap = 0;
    size_t j = 0;
    for (size_t i = 0; i < overlap; i++) {
        float x0 = y[yo + i - t1 + 2];
        float f = ORC_WINDOW[i] * ORC_WINDOW[i];
        y[yo + i] = y[yo + i]
            + (((1.0f - f) * g00) * y[yo + i - t0])
            + (((1.0f - f) * g01) * (y[yo + i - t0 + 1] + y[yo + i - t0 - 1]))
            + (((1.0f - f) * g02) * (y[yo + i - t0 + 2] + y[yo + i - t0 - 2]))
            + ((f * g10) * x2)
            + ((f * g11) * (x1 + x3))
            + ((f * g12) * (x0 + x4));
        x4 = x3, x3 = x2, x2 = x1, x1 = x0;
        j += 1;
    }
    if (g1 == 0.0f) return;
    comb_const_inplace(y, yo + j, t1, n - j, g10, g11, g12);
}
