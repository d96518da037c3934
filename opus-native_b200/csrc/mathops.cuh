// mathops.cuh -- bit-exact integer trigonometry (src/math.rs:51-75) on the device.
// CELT uses these Q15 routines for the theta bit split: the SYNTH-CELT/2 frame decode (celt2.cuh)
// calls them; the operator kernel below parity-checks them against the reference's checksums.
#pragma once
#include <stdint.h>

// The integer routines are also compiled for the host (packet generator, host_rangeenc.cpp).
#ifdef __CUDACC__
#define OPN_HD __host__ __device__ __forceinline__
#else
#define OPN_HD inline
#endif

namespace opn {

OPN_HD int32_t m_ilog(uint32_t x)  // math.rs:5-7
{
#ifdef __CUDA_ARCH__
    return 32 - __clz((int)x);
#else
    return x ? 32 - __builtin_clz(x) : 0;
#endif
}

// math.rs:72-75
OPN_HD int16_t frac_mul16(int16_t a, int16_t b)
{
    const int32_t x = (int32_t)a * (int32_t)b;
    return (int16_t)((16384 + x) >> 15);
}

// math.rs:51-55
OPN_HD int16_t bitexact_cos(int16_t x)
{
    const int32_t x2 = (int32_t)x * (int32_t)x;
    const int16_t y = (int16_t)((x2 + 4096) >> 13);
    return (int16_t)(1 + (32767 - y) + frac_mul16(y, (int16_t)(-7651 + frac_mul16(y, (int16_t)(8277 + frac_mul16(-626, y))))));
}

// math.rs:59-69
OPN_HD int32_t bitexact_log2tan(int32_t isin, int32_t icos)
{
    const int32_t ls = m_ilog((uint32_t)isin);
    const int32_t lc = m_ilog((uint32_t)icos);
    const int16_t c = (int16_t)(icos << (15 - lc));
    const int16_t s = (int16_t)(isin << (15 - ls));
    const int32_t a = frac_mul16(s, (int16_t)(frac_mul16(s, -2597) + 7932));
    const int32_t b = frac_mul16(c, (int16_t)(frac_mul16(c, -2597) + 7932));
    return (ls - lc) * (1 << 11) + a - b;
}

#ifdef __CUDACC__
// out_cos[i] = bitexact_cos(x[i]); out_l2t[i] = bitexact_log2tan(isin[i], icos[i]) (either pair may be null)
__global__ void k_op_bitexact_trig(const int16_t *__restrict__ x, int16_t *__restrict__ out_cos, uint32_t n_cos,
                                   const int32_t *__restrict__ isin, const int32_t *__restrict__ icos, int32_t *__restrict__ out_l2t,
                                   uint32_t n_l2t)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_cos) out_cos[i] = bitexact_cos(x[i]);
    if (i < n_l2t) out_l2t[i] = bitexact_log2tan(isin[i], icos[i]);
}
#endif

}  // namespace opn
