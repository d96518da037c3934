/* CPU ORACLE (test infrastructure) -- bit-exact integer trig.
 * Restates /root/reference/src/math.rs:51-75. */
#include "oracle.h"

/* math.rs:72-75 */
static int16_t frac_mul16(int16_t a, int16_t b)
{
    int32_t x = (int32_t)a * (int32_t)b;
    return (int16_t)((16384 + x) >> 15);
}

/* math.rs:51-55 */
int16_t orc_bitexact_cos(int16_t x)
{
    int32_t x2 = (int32_t)x * (int32_t)x;
    int16_t y = (int16_t)((x2 + 4096) >> 13);
    return (int16_t)(1 + (32767 - y) +
                     frac_mul16(y, (int16_t)(-7651 + frac_mul16(y, (int16_t)(8277 + frac_mul16(-626, y))))));
}

/* math.rs:59-69 */
int32_t orc_bitexact_log2tan(int32_t isin, int32_t icos)
{
    int32_t ls = (int32_t)orc_ilog((uint32_t)isin);
    int32_t lc = (int32_t)orc_ilog((uint32_t)icos);
    int16_t c = (int16_t)(icos << (15 - lc));
    int16_t s = (int16_t)(isin << (15 - ls));
    int32_t a = frac_mul16(s, (int16_t)(frac_mul16(s, -2597) + 7932));
    int32_t b = frac_mul16(c, (int16_t)(frac_mul16(c, -2597) + 7932));
    return (ls - lc) * (1 << 11) + a - b;
}
