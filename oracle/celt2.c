/* CPU ORACLE (test infrastructure) -- SYNTH-CELT/2: allocation-driven CELT frame decode.  PARITY UNPINNED.
 *
 * The reference's CeltDecoder::decode is `todo!()` (src/celt/decoder.rs:47-56).  What the reference DOES hold of a CELT
 * frame are its tables -- ALLOC_VECTORS, LOG_N, CACHE_INDEX / CACHE_BITS / CACHE_CAPS (src/celt/mode.rs:13-28, 70-111) --
 * its range decoder, decode_pulses and the integer trigonometry of the band split (bitexact_cos / bitexact_log2tan,
 * src/math.rs:51-75).  SYNTH-CELT/2 is the slice of the frame decode that those pieces define once they are wired
 * together the way RFC 6716 section 4.3 (libopus celt_decoder.c / rate.c / bands.c) wires them: the bit allocation is
 * COMPUTED per frame from the decoder's running tell_frac, so the shape (n, K) of every PVQ part is data dependent.
 * The wiring below is restated from the RFC's description and memory of libopus; neither is in this container, there
 * is no reference code and no reference test for it: the oracle defines truth for this layout and says so.
 *
 * Frame payload (after the TOC), `len` bytes, C channels, LM = log2(frame / 120):
 *   silence, post-filter parameters, transient, intra, coarse energies      as SYNTH-CELT/1 (oracle/synth.c)
 *   spread        = icdf({25,23,2,0}, 5)                                     decoded and reported only
 *   dynalloc      per band: boost flags bit_logp(6 -> 1), budget- and cap-limited (RFC 4.3.3 "band boost")
 *   alloc_trim    = icdf(trim table, 7) if 6 more bits fit, else 5
 *   allocation    compute_allocation(): static table interpolation, band skipping (skip flags bit_logp(1)),
 *                 intensity = uint(coded+1), dual_stereo = bit_logp(1) when stereo; fine-energy / PVQ bit split
 *   fine energy   bits(ebits[b]) per band and channel
 *   bands         per band b (and per channel: stereo bands are always coded as two mono bands, libopus's dual-stereo
 *                 path, whatever the decoded flag says): bits from pulses[b] and the running balance; a band whose
 *                 budget exceeds its largest codebook is split in halves with an angle theta (uniform pdf when the
 *                 frame is transient, triangular otherwise; gains from bitexact_cos, bit split from
 *                 bitexact_log2tan); a leaf decodes K = get_pulses(bits2pulses(b)) pulses with decode_pulses and
 *                 becomes coefficients y * gain / sqrt(yy), gain = 2^-5 x the product of its cos/sin factors
 *   anti-collapse bit (when reserved), final fine-energy bits by priority
 * Not in this slice (absent from the reference, float-heavy or table-less): tf_select / tf_change, spreading rotation,
 * folding and noise fill of empty bands, joint (mid/side) stereo bands, anti-collapse processing, applying the band
 * energies (denormalise_bands), de-emphasis.  After the coefficients the pipeline is SYNTH-CELT/1's (IMDCT, comb).
 *
 * The same function body encodes (generator: random symbol values through the oracle's range ENCODER) and decodes:
 * every decision is driven by tell_frac, which encoder and decoder agree on symbol by symbol. */
#include "oracle.h"
#include "oracle_tables.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BITRES 3
#define NBANDS 21
#define ALLOC_STEPS 6
#define FINE_OFFSET 21
#define MAX_FINE_BITS 8
#define QTHETA_OFFSET 4
#define LOG_MAX_PSEUDO 6

static const uint8_t TAPSET_ICDF[3] = {2, 1, 0};
static const uint8_t SPREAD_ICDF[4] = {25, 23, 2, 0};
static const uint8_t TRIM_ICDF[11] = {126, 124, 119, 109, 87, 41, 19, 9, 4, 2, 0};
static const int16_t EXP2_TABLE8[8] = {16384, 17866, 19483, 21247, 23170, 25267, 27554, 30048};

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------ symbol layer: one body, two directions */
typedef struct {
    uint64_t s;
} prng;
static uint64_t pr_next(prng *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t pr_below(prng *r, uint32_t n) { return (uint32_t)(((pr_next(r) >> 32) * (uint64_t)n) >> 32); }

typedef struct {
    orc_dec *d; /* decode */
    orc_enc *e; /* encode: symbol values are drawn from `rng` */
    prng rng;
    uint32_t transient_permille;
} coder;

static uint32_t c_tell_frac(coder *c) { return c->d ? orc_dec_tell_frac(c->d) : orc_enc_tell_frac(c->e); }
static uint32_t c_tell(coder *c) { return c->d ? orc_dec_tell(c->d) : orc_enc_tell(c->e); }
/* a flag with P(1) = 2^-logp in the bitstream; the generator draws 1 with probability p1/1000 */
static int c_bit_logp(coder *c, uint32_t logp, uint32_t p1_permille)
{
    if (c->d) return orc_dec_bit_logp(c->d, logp);
    int v = pr_below(&c->rng, 1000) < p1_permille;
    orc_enc_bit_logp(c->e, (uint32_t)v, logp);
    return v;
}
static uint32_t c_icdf(coder *c, const uint8_t *icdf, uint32_t ftb, uint32_t n_sym)
{
    if (c->d) return orc_dec_icdf(c->d, icdf, ftb);
    uint32_t v = pr_below(&c->rng, n_sym);
    orc_enc_icdf(c->e, v, icdf, ftb);
    return v;
}
static uint32_t c_uint(coder *c, uint32_t ft)
{
    if (c->d) return orc_dec_uint(c->d, ft);
    uint32_t v = pr_below(&c->rng, ft);
    orc_enc_uint(c->e, v, ft);
    return v;
}
static uint32_t c_bits(coder *c, uint32_t n)
{
    if (c->d) return orc_dec_bits(c->d, n);
    uint32_t v = pr_below(&c->rng, 1u << n);
    orc_enc_bits(c->e, v, n);
    return v;
}
static int32_t c_laplace(coder *c, uint32_t fs, uint32_t decay)
{
    if (c->d) return orc_dec_laplace(c->d, fs, decay);
    int32_t v = (int32_t)pr_below(&c->rng, 16) - 7;
    orc_enc_laplace(c->e, &v, fs, decay);
    return v;
}
/* the split angle with a triangular pdf over 0..qn (RFC 6716 4.3.4.? "theta", libopus compute_theta) */
static uint32_t isqrt32(uint32_t v)
{
    uint32_t g = 0, b = 1u << 15;
    for (int i = 0; i < 16; i++, b >>= 1)
        if ((uint64_t)(g + b) * (g + b) <= v) g += b;
    return g;
}
static uint32_t c_theta_tri(coder *c, uint32_t qn)
{
    const uint32_t h = qn >> 1, ft = (h + 1) * (h + 1);
    uint32_t itheta, fl, fs;
    if (c->d) {
        uint32_t fm = orc_dec_decode(c->d, ft);
        if (fm < ((h * (h + 1)) >> 1)) {
            itheta = (isqrt32(8 * fm + 1) - 1) >> 1;
            fs = itheta + 1;
            fl = (itheta * (itheta + 1)) >> 1;
        } else {
            itheta = (2 * (qn + 1) - isqrt32(8 * (ft - fm - 1) + 1)) >> 1;
            fs = qn + 1 - itheta;
            fl = ft - (((qn + 1 - itheta) * (qn + 2 - itheta)) >> 1);
        }
        orc_dec_update(c->d, fl, fl + fs, ft);
        return itheta;
    }
    itheta = pr_below(&c->rng, qn + 1);
    if (itheta <= h) {
        fs = itheta + 1;
        fl = (itheta * (itheta + 1)) >> 1;
    } else {
        fs = qn + 1 - itheta;
        fl = ft - (((qn + 1 - itheta) * (qn + 2 - itheta)) >> 1);
    }
    orc_enc_encode(c->e, fl, fl + fs, ft);
    return itheta;
}

/* ------------------------------------------------------------------ rate tables (libopus rate.h) */
static uint32_t get_pulses(uint32_t i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }
static const uint8_t *pulse_cache(int band, int lm) { return ORC_CACHE_BITS + ORC_CACHE_INDEX[(lm + 1) * NBANDS + band]; }
static int bits2pulses(int band, int lm, int bits)
{
    const uint8_t *cache = pulse_cache(band, lm);
    int lo = 0, hi = cache[0];
    bits--;
    for (int i = 0; i < LOG_MAX_PSEUDO; i++) {
        int mid = (lo + hi + 1) >> 1;
        if ((int)cache[mid] >= bits) hi = mid;
        else lo = mid;
    }
    return bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits ? lo : hi;
}
static int pulses2bits(int band, int lm, int pulses) { return pulses == 0 ? 0 : pulse_cache(band, lm)[pulses] + 1; }

/* ------------------------------------------------------------------ compute_allocation (libopus rate.c) */
static int interp_bits2pulses(coder *ec, int end, int skip_start, const int *bits1, const int *bits2, const int *thresh, const int *cap,
                              int32_t total, int32_t *balance_out, int skip_rsv, int *intensity, int intensity_rsv, int *dual_stereo,
                              int dual_stereo_rsv, int *bits, int *ebits, int *fine_priority, int C, int LM)
{
    const int start = 0, stereo = C > 1, alloc_floor = C << BITRES, logM = LM << BITRES;
    int32_t psum;
    int lo = 0, hi = 1 << ALLOC_STEPS, done, j, codedBands;
    for (int i = 0; i < ALLOC_STEPS; i++) {
        int mid = (lo + hi) >> 1;
        psum = 0;
        done = 0;
        for (j = end; j-- > start;) {
            int tmp = bits1[j] + (int)(((int32_t)mid * bits2[j]) >> ALLOC_STEPS);
            if (tmp >= thresh[j] || done) {
                done = 1;
                psum += imin(tmp, cap[j]);
            } else if (tmp >= alloc_floor)
                psum += alloc_floor;
        }
        if (psum > total) hi = mid;
        else lo = mid;
    }
    psum = 0;
    done = 0;
    for (j = end; j-- > start;) {
        int tmp = bits1[j] + (int)(((int32_t)lo * bits2[j]) >> ALLOC_STEPS);
        if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
        else done = 1;
        tmp = imin(tmp, cap[j]);
        bits[j] = tmp;
        psum += tmp;
    }
    /* band skipping, from the top */
    for (codedBands = end;; codedBands--) {
        j = codedBands - 1;
        if (j <= skip_start) {
            total += skip_rsv;
            break;
        }
        int32_t left = total - psum;
        int32_t percoeff = left / (ORC_E_BANDS[codedBands] - ORC_E_BANDS[start]);
        left -= (ORC_E_BANDS[codedBands] - ORC_E_BANDS[start]) * percoeff;
        int rem = imax((int)left - (ORC_E_BANDS[j] - ORC_E_BANDS[start]), 0);
        int band_width = ORC_E_BANDS[codedBands] - ORC_E_BANDS[j];
        int band_bits = (int)(bits[j] + percoeff * band_width + rem);
        if (band_bits >= imax(thresh[j], alloc_floor + (1 << BITRES))) {
            if (c_bit_logp(ec, 1, 850)) break; /* "this band is coded": the generator keeps most bands */
            psum += 1 << BITRES;
            band_bits -= 1 << BITRES;
        }
        psum -= bits[j] + intensity_rsv;
        if (intensity_rsv > 0) intensity_rsv = ORC_LOG2_FRAC_TABLE[j - start];
        psum += intensity_rsv;
        if (band_bits >= alloc_floor) {
            psum += alloc_floor;
            bits[j] = alloc_floor;
        } else
            bits[j] = 0;
    }
    if (intensity_rsv > 0) *intensity = start + (int)c_uint(ec, (uint32_t)(codedBands + 1 - start));
    else *intensity = 0;
    if (*intensity <= start) {
        total += dual_stereo_rsv;
        dual_stereo_rsv = 0;
    }
    if (dual_stereo_rsv > 0) *dual_stereo = c_bit_logp(ec, 1, 500);
    else *dual_stereo = 0;

    int32_t left = total - psum;
    int32_t percoeff = left / (ORC_E_BANDS[codedBands] - ORC_E_BANDS[start]);
    left -= (ORC_E_BANDS[codedBands] - ORC_E_BANDS[start]) * percoeff;
    for (j = start; j < codedBands; j++) bits[j] += (int)percoeff * (ORC_E_BANDS[j + 1] - ORC_E_BANDS[j]);
    for (j = start; j < codedBands; j++) {
        int tmp = (int)(left < ORC_E_BANDS[j + 1] - ORC_E_BANDS[j] ? left : ORC_E_BANDS[j + 1] - ORC_E_BANDS[j]);
        bits[j] += tmp;
        left -= tmp;
    }
    int32_t balance = 0;
    for (j = start; j < codedBands; j++) {
        int N0 = ORC_E_BANDS[j + 1] - ORC_E_BANDS[j], N = N0 << LM;
        int32_t bit = (int32_t)bits[j] + balance, excess;
        if (N > 1) {
            excess = bit - cap[j] > 0 ? bit - cap[j] : 0;
            bits[j] = bit - excess;
            int den = C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0);
            int NClogN = den * (ORC_LOG_N[j] + logM);
            int offset = (NClogN >> 1) - den * FINE_OFFSET;
            if (N == 2) offset += den << BITRES >> 2;
            if (bits[j] + offset < den * 2 << BITRES) offset += NClogN >> 2;
            else if (bits[j] + offset < den * 3 << BITRES) offset += NClogN >> 3;
            ebits[j] = imax(0, bits[j] + offset + (den << (BITRES - 1)));
            ebits[j] = (ebits[j] / den) >> BITRES;
            if (C * ebits[j] > (bits[j] >> BITRES)) ebits[j] = bits[j] >> stereo >> BITRES;
            ebits[j] = imin(ebits[j], MAX_FINE_BITS);
            fine_priority[j] = ebits[j] * (den << BITRES) >= bits[j] + offset;
            bits[j] -= C * ebits[j] << BITRES;
        } else {
            excess = bit - (C << BITRES) > 0 ? bit - (C << BITRES) : 0;
            bits[j] = bit - excess;
            ebits[j] = 0;
            fine_priority[j] = 1;
        }
        if (excess > 0) {
            int extra_fine = imin((int)(excess >> (stereo + BITRES)), MAX_FINE_BITS - ebits[j]);
            ebits[j] += extra_fine;
            int extra_bits = extra_fine * C << BITRES;
            fine_priority[j] = extra_bits >= excess - balance;
            excess -= extra_bits;
        }
        balance = excess;
    }
    *balance_out = balance;
    for (; j < end; j++) {
        ebits[j] = bits[j] >> stereo >> BITRES;
        bits[j] = 0;
        fine_priority[j] = ebits[j] < 1;
    }
    return codedBands;
}

static int compute_allocation(coder *ec, int end, const int *offsets, const int *cap, int alloc_trim, int *intensity, int *dual_stereo,
                              int32_t total, int32_t *balance, int *pulses, int *ebits, int *fine_priority, int C, int LM)
{
    const int start = 0, len = NBANDS;
    int thresh[NBANDS], trim_offset[NBANDS], bits1[NBANDS], bits2[NBANDS];
    int skip_start = start, j;
    if (total < 0) total = 0;
    int skip_rsv = total >= 1 << BITRES ? 1 << BITRES : 0;
    total -= skip_rsv;
    int intensity_rsv = 0, dual_stereo_rsv = 0;
    if (C == 2) {
        intensity_rsv = ORC_LOG2_FRAC_TABLE[end - start];
        if (intensity_rsv > total) intensity_rsv = 0;
        else {
            total -= intensity_rsv;
            dual_stereo_rsv = total >= 1 << BITRES ? 1 << BITRES : 0;
            total -= dual_stereo_rsv;
        }
    }
    for (j = start; j < end; j++) {
        int w = ORC_E_BANDS[j + 1] - ORC_E_BANDS[j];
        thresh[j] = imax(C << BITRES, (3 * w << LM << BITRES) >> 4);
        trim_offset[j] = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + BITRES)) >> 6;
        if (w << LM == 1) trim_offset[j] -= C << BITRES;
    }
    int lo = 1, hi = ORC_NB_ALLOC_VECTORS - 1;
    do {
        int done = 0, psum = 0, mid = (lo + hi) >> 1;
        for (j = end; j-- > start;) {
            int N = ORC_E_BANDS[j + 1] - ORC_E_BANDS[j];
            int bitsj = C * N * ORC_ALLOC_VECTORS[mid * len + j] << LM >> 2;
            if (bitsj > 0) bitsj = imax(0, bitsj + trim_offset[j]);
            bitsj += offsets[j];
            if (bitsj >= thresh[j] || done) {
                done = 1;
                psum += imin(bitsj, cap[j]);
            } else if (bitsj >= C << BITRES)
                psum += C << BITRES;
        }
        if (psum > total) hi = mid - 1;
        else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
    for (j = start; j < end; j++) {
        int N = ORC_E_BANDS[j + 1] - ORC_E_BANDS[j];
        int bits1j = C * N * ORC_ALLOC_VECTORS[lo * len + j] << LM >> 2;
        int bits2j = hi >= ORC_NB_ALLOC_VECTORS ? cap[j] : C * N * ORC_ALLOC_VECTORS[hi * len + j] << LM >> 2;
        if (bits1j > 0) bits1j = imax(0, bits1j + trim_offset[j]);
        if (bits2j > 0) bits2j = imax(0, bits2j + trim_offset[j]);
        if (lo > 0) bits1j += offsets[j];
        bits2j += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        bits2j = imax(0, bits2j - bits1j);
        bits1[j] = bits1j;
        bits2[j] = bits2j;
    }
    return interp_bits2pulses(ec, end, skip_start, bits1, bits2, thresh, cap, total, balance, skip_rsv, intensity, intensity_rsv,
                              dual_stereo, dual_stereo_rsv, pulses, ebits, fine_priority, C, LM);
}

/* ------------------------------------------------------------------ bands (libopus bands.c, mono partition only) */
typedef struct {
    coder *ec;
    int band, transient_blocks; /* B of the frame: 1, or 2^LM short blocks */
    int32_t remaining_bits;
    orc_celt2_side *side;
    orc_celt2_part *parts;
    float *coef;   /* channel's row */
    float *coef0;  /* channel 0's row (the part list records positions in the channel-major frame) */
    int32_t *y_out;
} band_ctx;

static int compute_qn(int N, int b, int offset, int pulse_cap)
{
    int N2 = 2 * N - 1;
    int qb = (b + N2 * offset) / N2; /* celt_sudiv: operands are non-negative here or the C division applies */
    qb = imin(b - pulse_cap - (4 << BITRES), qb);
    qb = imin(8 << BITRES, qb);
    if (qb < (1 << BITRES >> 1)) return 1;
    int qn = EXP2_TABLE8[qb & 7] >> (14 - (qb >> BITRES));
    return (qn + 1) >> 1 << 1;
}

static void quant_partition(band_ctx *ctx, int base, int N, int b, int B, int LM, float gain)
{
    const uint8_t *cache = pulse_cache(ctx->band, LM);
    if (LM != -1 && b > cache[cache[0]] + 12 && N > 2) {
        /* split in halves, compute_theta for a mono partition */
        const int B0 = B;
        N >>= 1;
        LM -= 1;
        B = (B + 1) >> 1;
        int pulse_cap = ORC_LOG_N[ctx->band] + LM * (1 << BITRES);
        int offset = (pulse_cap >> 1) - QTHETA_OFFSET;
        int qn = compute_qn(N, b, offset, pulse_cap);
        int32_t tell = (int32_t)c_tell_frac(ctx->ec);
        int itheta = 0;
        if (qn != 1) {
            itheta = B0 > 1 ? (int)c_uint(ctx->ec, (uint32_t)qn + 1) : (int)c_theta_tri(ctx->ec, (uint32_t)qn);
            itheta = (int)(((int32_t)itheta * 16384) / qn);
        }
        int32_t qalloc = (int32_t)c_tell_frac(ctx->ec) - tell;
        b -= qalloc;
        int imid, iside, delta;
        if (itheta == 0) {
            imid = 32767;
            iside = 0;
            delta = -16384;
        } else if (itheta == 16384) {
            imid = 0;
            iside = 32767;
            delta = 16384;
        } else {
            imid = orc_bitexact_cos((int16_t)itheta);
            iside = orc_bitexact_cos((int16_t)(16384 - itheta));
            /* FRAC_MUL16((N-1)<<7, bitexact_log2tan(iside, imid)) */
            delta = (16384 + (int32_t)(int16_t)((N - 1) << 7) * (int32_t)(int16_t)orc_bitexact_log2tan(iside, imid)) >> 15;
        }
        ctx->side->n_splits += 1;
        ctx->side->theta_sum += (uint32_t)itheta;
        const float mid = (1.0f / 32768.0f) * (float)imid, side = (1.0f / 32768.0f) * (float)iside;
        if (B0 > 1 && (itheta & 0x3fff)) {
            if (itheta > 8192) delta -= delta >> (4 - LM);
            else delta = imin(0, delta + (N << BITRES >> (5 - LM)));
        }
        int mbits = imax(0, imin(b, (b - delta) / 2));
        int sbits = b - mbits;
        ctx->remaining_bits -= qalloc;
        int32_t rebalance = ctx->remaining_bits;
        if (mbits >= sbits) {
            quant_partition(ctx, base, N, mbits, B, LM, gain * mid);
            rebalance = mbits - (rebalance - ctx->remaining_bits);
            if (rebalance > 3 << BITRES && itheta != 0) sbits += rebalance - (3 << BITRES);
            quant_partition(ctx, base + N, N, sbits, B, LM, gain * side);
        } else {
            quant_partition(ctx, base + N, N, sbits, B, LM, gain * side);
            rebalance = sbits - (rebalance - ctx->remaining_bits);
            if (rebalance > 3 << BITRES && itheta != 16384) mbits += rebalance - (3 << BITRES);
            quant_partition(ctx, base, N, mbits, B, LM, gain * mid);
        }
        return;
    }
    /* leaf */
    int q = bits2pulses(ctx->band, LM, b);
    int curr_bits = pulses2bits(ctx->band, LM, q);
    ctx->remaining_bits -= curr_bits;
    while (ctx->remaining_bits < 0 && q > 0) {
        ctx->remaining_bits += curr_bits;
        q--;
        curr_bits = pulses2bits(ctx->band, LM, q);
        ctx->remaining_bits -= curr_bits;
    }
    if (q == 0) return; /* no pulses: the part stays zero (no folding in this slice) */
    const uint32_t K = get_pulses((uint32_t)q);
    const uint32_t ft = orc_pvq_v((uint32_t)N, K);
    int32_t y[176];
    uint32_t index;
    if (ctx->ec->d) {
        index = orc_dec_uint(ctx->ec->d, ft);
    } else {
        index = pr_below(&ctx->ec->rng, ft);
        orc_enc_uint(ctx->ec->e, index, ft);
    }
    const float yy = orc_cwrsi(y, (uint32_t)N, K, index);
    const float g = gain / sqrtf(yy);
    for (int j = 0; j < N; j++) {
        ctx->coef[base + j] = (float)y[j] * g;
        if (ctx->y_out) ctx->y_out[base + j] = y[j];
    }
    orc_celt2_side *sd = ctx->side;
    if (sd->n_parts < ORC_CELT2_MAX_PARTS) {
        orc_celt2_part *p = &ctx->parts[sd->n_parts];
        p->base = (uint16_t)((base + (int)(ctx->coef - ctx->coef0)) | ctx->band << 11);
        p->n = (uint8_t)N;
        p->k = (uint8_t)K;
        p->index = index;
        p->gain = gain;
    }
    sd->n_parts += 1;
    sd->n_pulses += K;
}

/* one frame, either direction; coef: [C][120<<LM] zeroed by the caller.  Returns 0 or ORC_ERR_*. */
static int celt2_frame(coder *ec, uint32_t len, int LM, int C, orc_celt2_side *sd, orc_celt2_part *parts, float *coef, int32_t *y_out)
{
    const int nf = 120 << LM, end = NBANDS, M = 1 << LM;
    memset(sd, 0, sizeof(*sd));
    int32_t total_bits = (int32_t)len * 8;
    uint32_t n_one_bin = 0;
    sd->silence = c_bit_logp(ec, 15, 0);
    if (sd->silence) goto done;
    sd->postfilter = c_bit_logp(ec, 1, 500);
    if (sd->postfilter) {
        sd->octave = (int32_t)c_uint(ec, 6);
        sd->period = (16 << sd->octave) + (int32_t)c_bits(ec, 4 + (uint32_t)sd->octave) - 1;
        sd->gain_idx = (int32_t)c_bits(ec, 3);
        sd->tapset = (int32_t)c_icdf(ec, TAPSET_ICDF, 2, 3);
    }
    sd->transient = c_bit_logp(ec, 3, ec->transient_permille);
    sd->intra = c_bit_logp(ec, 3, 125);
    for (int b = 0; b < NBANDS; b++)
        for (int c = 0; c < C; c++) {
            uint32_t decay = 6000u + 400u * (uint32_t)b;
            sd->coarse[c][b] = c_laplace(ec, orc_laplace_start_freq(decay), decay);
        }
    sd->spread = (int32_t)c_icdf(ec, SPREAD_ICDF, 5, 4);
    /* band boosts */
    int cap[NBANDS], offsets[NBANDS];
    for (int i = 0; i < NBANDS; i++) {
        int N = (ORC_E_BANDS[i + 1] - ORC_E_BANDS[i]) << LM;
        cap[i] = (ORC_CACHE_CAPS[NBANDS * (2 * LM + C - 1) + i] + 64) * C * N >> 2;
    }
    {
        int dynalloc_logp = 6;
        int32_t total_frac = total_bits << BITRES;
        int32_t tell = (int32_t)c_tell_frac(ec);
        for (int i = 0; i < end; i++) {
            int width = C * (ORC_E_BANDS[i + 1] - ORC_E_BANDS[i]) << LM;
            int quanta = imin(width << BITRES, imax(6 << BITRES, width));
            int loop_logp = dynalloc_logp, boost = 0;
            while (tell + (loop_logp << BITRES) < total_frac && boost < cap[i]) {
                int flag = c_bit_logp(ec, (uint32_t)loop_logp, 30);
                tell = (int32_t)c_tell_frac(ec);
                if (!flag) break;
                boost += quanta;
                total_frac -= quanta;
                loop_logp = 1;
            }
            offsets[i] = boost;
            sd->offsets[i] = boost;
            if (boost > 0) dynalloc_logp = imax(2, dynalloc_logp - 1);
        }
        sd->alloc_trim = tell + (6 << BITRES) <= total_frac ? (int32_t)c_icdf(ec, TRIM_ICDF, 7, 11) : 5;
    }
    int32_t bits = (total_bits << BITRES) - (int32_t)c_tell_frac(ec) - 1;
    int anti_collapse_rsv = sd->transient && LM >= 2 && bits >= ((LM + 2) << BITRES) ? (1 << BITRES) : 0;
    bits -= anti_collapse_rsv;
    int pulses[NBANDS], ebits[NBANDS], fine_priority[NBANDS], intensity = 0, dual_stereo = 0;
    int32_t balance = 0;
    int codedBands = compute_allocation(ec, end, offsets, cap, sd->alloc_trim, &intensity, &dual_stereo, bits, &balance, pulses, ebits,
                                        fine_priority, C, LM);
    sd->coded_bands = codedBands;
    sd->intensity = intensity;
    sd->dual_stereo = dual_stereo;
    sd->balance = balance;
    for (int i = 0; i < NBANDS; i++) {
        sd->pulses[i] = pulses[i];
        sd->ebits[i] = ebits[i];
        sd->fine_priority[i] = fine_priority[i];
    }
    /* fine energy */
    for (int i = 0; i < end; i++)
        if (ebits[i] > 0)
            for (int c = 0; c < C; c++) sd->fine[c][i] = (int32_t)c_bits(ec, (uint32_t)ebits[i]);
    /* bands */
    {
        band_ctx ctx;
        ctx.ec = ec;
        ctx.transient_blocks = sd->transient ? M : 1;
        ctx.side = sd;
        ctx.parts = parts;
        ctx.coef0 = coef;
        const int32_t band_total = (total_bits << BITRES) - anti_collapse_rsv;
        for (int i = 0; i < end; i++) {
            int32_t tell = (int32_t)c_tell_frac(ec);
            if (i != 0) balance -= tell;
            int32_t remaining_bits = band_total - tell - 1;
            int b = 0;
            if (i <= codedBands - 1) {
                int32_t curr_balance = balance / imin(3, codedBands - i);
                b = imax(0, imin(16383, imin((int)remaining_bits + 1, pulses[i] + (int)curr_balance)));
            }
            const int N = (ORC_E_BANDS[i + 1] - ORC_E_BANDS[i]) << LM, base = ORC_E_BANDS[i] << LM;
            ctx.band = i;
            ctx.remaining_bits = remaining_bits;
            for (int c = 0; c < C; c++) {
                ctx.coef = coef + c * nf;
                ctx.y_out = y_out ? y_out + c * nf : NULL;
                const int bc = b / C;
                if (N == 1) { /* quant_band_n1: one sign bit if it fits */
                    int sign = 0;
                    if (ctx.remaining_bits >= 1 << BITRES) {
                        sign = (int)c_bits(ec, 1);
                        ctx.remaining_bits -= 1 << BITRES;
                    }
                    ctx.coef[base] = sign ? -0.03125f : 0.03125f;
                    if (ctx.y_out) ctx.y_out[base] = sign ? -1 : 1;
                    sd->n_pulses += 1;
                    n_one_bin += 1;
                } else {
                    quant_partition(&ctx, base + c * 0, N, bc, ctx.transient_blocks, LM, 0.03125f);
                }
            }
            balance += pulses[i] + tell;
        }
    }
    if (anti_collapse_rsv > 0) sd->anti_collapse = (int32_t)c_bits(ec, 1);
    /* unquant_energy_finalise: left-over whole bits, priority 0 bands first */
    {
        int bits_left = (int)total_bits - (int)c_tell(ec);
        for (int prio = 0; prio < 2; prio++)
            for (int i = 0; i < end && bits_left >= C; i++) {
                if (ebits[i] >= MAX_FINE_BITS || fine_priority[i] != prio) continue;
                for (int c = 0; c < C; c++) {
                    sd->fine_final[c][i] = 1 + (int32_t)c_bits(ec, 1); /* 0 = none, 1 / 2 = the bit */
                    bits_left--;
                }
            }
    }
    /* band energies (unquant_coarse_energy without prediction, unquant_fine_energy, unquant_energy_finalise; RFC 6716 4.3.2)
     * in 1/512 of a doubling, then denormalise_bands: every coefficient of the band times 2^energy */
    for (int c = 0; c < C; c++)
        for (int i = 0; i < end; i++) {
            int q = sd->coarse[c][i];
            int e = (q < -6 ? -6 : q > 2 ? 2 : q) * 512;
            if (ebits[i] > 0) e += (int)((2u * (uint32_t)sd->fine[c][i] + 1u) << (8 - ebits[i])) - 256; /* (v + 1/2) 2^-fq - 1/2 */
            if (sd->fine_final[c][i]) e += (2 * (sd->fine_final[c][i] - 1) - 1) * (1 << (7 - ebits[i]));   /* (v - 1/2) 2^-(fq+1) */
            sd->energy_q9[c][i] = e;
            const int fr = e & 511, ex = (e - fr) / 512;
            const float g = ORC_EXP2_Q9[fr] * ldexpf(1.0f, ex);
            for (int j = ORC_E_BANDS[i] << LM; j < ORC_E_BANDS[i + 1] << LM; j++) coef[c * nf + j] *= g;
        }
done:
    sd->tell_frac = c_tell_frac(ec);
    sd->final_rng = ec->d ? ec->d->rng : ec->e->rng;
    /* the product keeps PVQ leaves and one-bin bands in one list of ORC_CELT2_MAX_PARTS entries: a frame that needs more is rejected */
    if (sd->n_parts + n_one_bin > ORC_CELT2_MAX_PARTS) return ORC_ERR_INTERNAL;
    return ec->e ? ec->e->error : 0;
}

int orc_celt2_decode_symbols(const uint8_t *payload, uint32_t len, int lm, int channels, orc_celt2_side *side, orc_celt2_part *parts,
                             int32_t *y_out, float *coef_out)
{
    if (lm < 0 || lm > 3 || channels < 1 || channels > 2 || !payload || len < 2 || !side || !coef_out) return ORC_ERR_BAD_ARG;
    orc_celt2_part local[ORC_CELT2_MAX_PARTS];
    const int nf = 120 << lm;
    memset(coef_out, 0, sizeof(float) * (size_t)nf * (size_t)channels);
    if (y_out) memset(y_out, 0, sizeof(int32_t) * (size_t)nf * (size_t)channels);
    orc_dec d;
    orc_dec_init(&d, payload, len);
    coder c;
    memset(&c, 0, sizeof(c));
    c.d = &d;
    return celt2_frame(&c, len, lm, channels, side, parts ? parts : local, coef_out, y_out);
}

/* One frame of a stream: symbols -> coefficients -> PCM through the shared back end (oracle/synth.c); payloads of 0 or 1
 * byte are lost frames (src/decoder.rs:467).  A frame whose part list overflows is rejected: returns the error, state untouched. */
int orc_celt2_decode_frame(orc_synth_state *st, const uint8_t *payload, uint32_t len, int lm, int channels, int apply_comb,
                           orc_celt2_side *side, float *pcm_out)
{
    if (lm < 0 || lm > 3 || channels < 1 || channels > 2) return ORC_ERR_BAD_ARG;
    float coef[2 * 960];
    orc_celt2_side local;
    if (!side) side = &local;
    const int lost = len <= 1;
    if (lost) {
        memset(side, 0, sizeof(*side));
        memset(coef, 0, sizeof(coef));
    } else {
        int rc = orc_celt2_decode_symbols(payload, len, lm, channels, side, NULL, NULL, coef);
        if (rc) return rc;
    }
    return orc_synth_finish_frame(st, coef, lm, channels, apply_comb, lost, side->postfilter, side->period, side->gain_idx, side->tapset,
                                  side->transient, pcm_out);
}

/* The same with a packet of stream_channels channels in a decoder of `channels` channels (orc_map_channels, oracle/synth.c). */
int orc_celt2_decode_frame_mapped(orc_synth_state *st, const uint8_t *payload, uint32_t len, int lm, int stream_channels, int channels,
                                  int apply_comb, orc_celt2_side *side, float *pcm_out)
{
    if (lm < 0 || lm > 3 || channels < 1 || channels > 2 || stream_channels < 1 || stream_channels > 2) return ORC_ERR_BAD_ARG;
    float coef[2 * 960];
    orc_celt2_side local;
    if (!side) side = &local;
    const int lost = len <= 1;
    memset(coef, 0, sizeof(coef));
    if (lost) {
        memset(side, 0, sizeof(*side));
    } else {
        int rc = orc_celt2_decode_symbols(payload, len, lm, stream_channels, side, NULL, NULL, coef);
        if (rc) return rc;
        orc_map_channels(coef, 120 << lm, stream_channels, channels);
    }
    return orc_synth_finish_frame(st, coef, lm, channels, apply_comb, lost, side->postfilter, side->period, side->gain_idx, side->tapset,
                                  side->transient, pcm_out);
}

/* TOC + SYNTH-CELT/2 payload of exactly pkt_bytes, symbol values drawn from splitmix64(stream, frame). */
int orc_celt2_packet(uint64_t stream_id, uint64_t frame_idx, int lm, int channels, uint32_t pkt_bytes, uint32_t transient_permille,
                     uint8_t *out, orc_celt2_side *truth)
{
    if (!out || lm < 0 || lm > 3 || channels < 1 || channels > 2 || pkt_bytes < 8 || pkt_bytes > 1276) return ORC_ERR_BAD_ARG;
    orc_celt2_side local;
    orc_celt2_part parts[ORC_CELT2_MAX_PARTS];
    float coef[2 * 960];
    out[0] = (uint8_t)(0x80 | 0x60 | (lm << 3) | (channels == 2 ? 0x4 : 0));
    orc_enc e;
    orc_enc_init(&e, out + 1, pkt_bytes - 1);
    coder c;
    memset(&c, 0, sizeof(c));
    c.e = &e;
    c.rng.s = 4242ull + 1000003ull * stream_id + 0xD1B54A32D192ED03ull * frame_idx;
    c.transient_permille = transient_permille;
    memset(coef, 0, sizeof(coef));
    int rc = celt2_frame(&c, pkt_bytes - 1, lm, channels, truth ? truth : &local, parts, coef, NULL);
    if (rc) return rc;
    if (orc_enc_tell(&e) > 8u * (pkt_bytes - 1u)) return ORC_ERR_BUFFER_TOO_SMALL;
    orc_enc_done(&e);
    return e.error ? e.error : (int)pkt_bytes;
}
