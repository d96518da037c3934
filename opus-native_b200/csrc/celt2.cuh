// celt2.cuh -- SYNTH-CELT/2: the allocation-driven part of a CELT frame decode (first slice of CeltDecoder::decode,
// src/celt/decoder.rs:47-56, which is `todo!()` in the reference).  PARITY UNPINNED: see DESIGN.md section 3b.
//
// What the reference holds of a CELT frame -- the allocation tables ALLOC_VECTORS / LOG_N / CACHE_INDEX / CACHE_BITS /
// CACHE_CAPS (src/celt/mode.rs:13-28, 70-111), the range decoder, decode_pulses and the integer trigonometry of the band
// split (src/math.rs:51-75) -- wired together as RFC 6716 section 4.3 describes: band boosts, allocation trim,
// compute_allocation driven by the decoder's running tell_frac (skip flags, intensity, dual-stereo flag decoded on the
// way), fine-energy bits, and per band a recursive split in halves with an angle theta (bitexact_cos gains,
// bitexact_log2tan bit split) down to leaves of K = get_pulses(bits2pulses(b)) pulses.  So the shape (n, K) of every
// PVQ part and its alphabet are computed per frame, from the bitstream, on the device.
// Not in this slice: tf_select / tf_change, spreading rotation, folding / noise fill of empty bands, joint stereo bands
// (stereo bands are always two mono bands), anti-collapse processing, applying the band energies, de-emphasis.
//
// One body, two users: the lane-per-packet decode kernel (Coder = a LaneDec wrapper) and the host packet generator
// (Coder = the host range ENCODER drawing random symbol values; every decision is driven by tell_frac, which encoder
// and decoder agree on symbol by symbol).  The CPU oracle has its own, independent C restatement (oracle/celt2.c).
#pragma once
#include <stdint.h>

#include "mathops.cuh"

namespace opn {

constexpr int C2_BITRES = 3, C2_NBANDS = 21, C2_ALLOC_STEPS = 6, C2_FINE_OFFSET = 21, C2_MAX_FINE_BITS = 8, C2_QTHETA_OFFSET = 4,
              C2_LOG_MAX_PSEUDO = 6, C2_NB_ALLOC_VECTORS = 11;
constexpr int CELT2_MAX_PARTS = 192;

struct Celt2Tabs {  // device: g_tab members; host: the OPN_* tables of opn_tables.h
    const uint8_t *e_bands, *log_n, *alloc_vectors, *cache_bits, *cache_caps, *log2_frac;
    const int16_t *cache_index;
    const uint32_t *pvq_u;    // U(n,k) rows back to back (pvc.rs:315-429)
    const uint16_t *pvq_row;  // row offsets (pvc.rs:301-303)
};

// One PVQ leaf: coefficients [pos, pos+n) of the channel-major frame = (cwrsi(n, k, index) * (gain / sqrt(yy))) * band gain,
// pos = base & C2_POS_MASK, band = base >> C2_BAND_SHIFT (the band gain is 2^(energy_q9[channel][band] / 512), c2_band_gain).
constexpr int C2_BAND_SHIFT = 11, C2_POS_MASK = (1 << C2_BAND_SHIFT) - 1;
struct Celt2Part {
    uint16_t base;
    uint8_t n, k;
    uint32_t index;
    float gain;
};

// Band energies (the unquant_coarse/fine/finalise steps of RFC 6716 4.3.2, without inter-frame prediction and mean
// removal -- those tables are not in the reference): everything in Q9 (1/512 of a doubling), exact integers.
//   coarse value q           -> clamp(q, -6, 2) * 512
//   fine value v of fq bits  -> ((v + 1/2) 2^-fq - 1/2) * 512 = ((2v + 1) << (8 - fq)) - 256          (fq <= 8)
//   final bit v after fq bits-> ((v - 1/2) 2^-(fq+1)) * 512   = (2v - 1) << (7 - fq)                   (fq <= 7)
// 2^(e / 512) for a band energy e in Q9: table value of the fraction times an exact power of two (|e >> 9| <= 8)
OPN_HD float c2_band_gain(const float *exp2_q9, int e)
{
    union {
        uint32_t u;
        float f;
    } two;
    two.u = (uint32_t)(127 + (e >> 9)) << 23;
    return exp2_q9[e & 511] * two.f;
}
OPN_HD int c2_energy_coarse_q9(int32_t q) { return (q < -6 ? -6 : q > 2 ? 2 : q) * 512; }
OPN_HD int c2_energy_fine_q9(uint32_t v, int fq) { return (int)(((2u * v + 1u) << (8 - fq))) - 256; }
OPN_HD int c2_energy_final_q9(uint32_t v, int fq) { return (2 * (int)v - 1) * (1 << (7 - fq)); }

OPN_HD int c2_min(int a, int b) { return a < b ? a : b; }
OPN_HD int c2_max(int a, int b) { return a > b ? a : b; }
OPN_HD uint32_t c2_get_pulses(uint32_t i) { return i < 8 ? i : (8 + (i & 7)) << ((i >> 3) - 1); }
OPN_HD const uint8_t *c2_pulse_cache(const Celt2Tabs &T, int band, int lm) { return T.cache_bits + T.cache_index[(lm + 1) * C2_NBANDS + band]; }
OPN_HD int c2_bits2pulses(const Celt2Tabs &T, int band, int lm, int bits)
{
    const uint8_t *cache = c2_pulse_cache(T, band, lm);
    int lo = 0, hi = cache[0];
    bits--;
    for (int i = 0; i < C2_LOG_MAX_PSEUDO; i++) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)cache[mid] >= bits) hi = mid;
        else lo = mid;
    }
    return bits - (lo == 0 ? -1 : (int)cache[lo]) <= (int)cache[hi] - bits ? lo : hi;
}
OPN_HD int c2_pulses2bits(const Celt2Tabs &T, int band, int lm, int pulses) { return pulses == 0 ? 0 : c2_pulse_cache(T, band, lm)[pulses] + 1; }
OPN_HD uint32_t c2_pvq_v(const Celt2Tabs &T, uint32_t n, uint32_t k)  // pvc.rs:289-298
{
    const uint32_t a = n < k ? n : k, b = n < k ? k : n;           // U(n,k)
    const uint32_t a1 = n < k + 1 ? n : k + 1, b1 = n < k + 1 ? k + 1 : n;  // U(n,k+1)
    return T.pvq_u[T.pvq_row[a] + b] + T.pvq_u[T.pvq_row[a1] + b1];
}
OPN_HD uint32_t c2_isqrt32(uint32_t v)
{
    uint32_t g = 0, b = 1u << 15;
    for (int i = 0; i < 16; i++, b >>= 1)
        if ((uint64_t)(g + b) * (g + b) <= v) g += b;
    return g;
}

// What one frame reports besides its parts (tests compare all of it with the oracle).
struct Celt2Side {
    int32_t silence, postfilter, octave, period, gain_idx, tapset, transient, intra;
    int32_t spread, alloc_trim, coded_bands, intensity, dual_stereo, anti_collapse, balance;
    int32_t offsets[21], pulses[21], ebits[21], fine_priority[21];
    int32_t coarse[2][21], fine[2][21], fine_final[2][21];
    int32_t energy_q9[2][21];  // band energy (log2 of the band's gain) in 1/512: coarse + fine + final refinement
    uint32_t n_parts, n_pulses, n_splits, theta_sum;
    uint32_t final_rng, tell_frac;
};

// ---- compute_allocation / interp_bits2pulses (RFC 6716 4.3.3; libopus rate.c) --------------------------------
template <class Coder>
OPN_HD int c2_interp_bits2pulses(Coder &ec, const Celt2Tabs &T, int end, int skip_start, const int *bits1, const int *bits2, const int *thresh,
                                 const int *cap, int32_t total, int32_t *balance_out, int skip_rsv, int *intensity, int intensity_rsv,
                                 int *dual_stereo, int dual_stereo_rsv, int *bits, int *ebits, int *fine_priority, int C, int LM)
{
    const int start = 0, stereo = C > 1, alloc_floor = C << C2_BITRES, logM = LM << C2_BITRES;
    const uint8_t *eb = T.e_bands;
    int32_t psum;
    int lo = 0, hi = 1 << C2_ALLOC_STEPS, done, j, codedBands;
    for (int i = 0; i < C2_ALLOC_STEPS; i++) {
        const int mid = (lo + hi) >> 1;
        psum = 0;
        done = 0;
        for (j = end; j-- > start;) {
            const int tmp = bits1[j] + (int)(((int32_t)mid * bits2[j]) >> C2_ALLOC_STEPS);
            if (tmp >= thresh[j] || done) {
                done = 1;
                psum += c2_min(tmp, cap[j]);
            } else if (tmp >= alloc_floor)
                psum += alloc_floor;
        }
        if (psum > total) hi = mid;
        else lo = mid;
    }
    psum = 0;
    done = 0;
    for (j = end; j-- > start;) {
        int tmp = bits1[j] + (int)(((int32_t)lo * bits2[j]) >> C2_ALLOC_STEPS);
        if (tmp < thresh[j] && !done) tmp = tmp >= alloc_floor ? alloc_floor : 0;
        else done = 1;
        tmp = c2_min(tmp, cap[j]);
        bits[j] = tmp;
        psum += tmp;
    }
    for (codedBands = end;; codedBands--) {  // band skipping, from the top
        j = codedBands - 1;
        if (j <= skip_start) {
            total += skip_rsv;
            break;
        }
        int32_t left = total - psum;
        const int32_t percoeff = left / (eb[codedBands] - eb[start]);
        left -= (eb[codedBands] - eb[start]) * percoeff;
        const int rem = c2_max((int)left - (eb[j] - eb[start]), 0);
        const int band_width = eb[codedBands] - eb[j];
        int band_bits = (int)(bits[j] + percoeff * band_width + rem);
        if (band_bits >= c2_max(thresh[j], alloc_floor + (1 << C2_BITRES))) {
            if (ec.bit_logp(1, 850)) break;
            psum += 1 << C2_BITRES;
            band_bits -= 1 << C2_BITRES;
        }
        psum -= bits[j] + intensity_rsv;
        if (intensity_rsv > 0) intensity_rsv = T.log2_frac[j - start];
        psum += intensity_rsv;
        if (band_bits >= alloc_floor) {
            psum += alloc_floor;
            bits[j] = alloc_floor;
        } else
            bits[j] = 0;
    }
    if (intensity_rsv > 0) *intensity = start + (int)ec.uint_((uint32_t)(codedBands + 1 - start));
    else *intensity = 0;
    if (*intensity <= start) {
        total += dual_stereo_rsv;
        dual_stereo_rsv = 0;
    }
    if (dual_stereo_rsv > 0) *dual_stereo = (int)ec.bit_logp(1, 500);
    else *dual_stereo = 0;

    int32_t left = total - psum;
    const int32_t percoeff = left / (eb[codedBands] - eb[start]);
    left -= (eb[codedBands] - eb[start]) * percoeff;
    for (j = start; j < codedBands; j++) bits[j] += (int)percoeff * (eb[j + 1] - eb[j]);
    for (j = start; j < codedBands; j++) {
        const int tmp = (int)(left < eb[j + 1] - eb[j] ? left : eb[j + 1] - eb[j]);
        bits[j] += tmp;
        left -= tmp;
    }
    int32_t balance = 0;
    for (j = start; j < codedBands; j++) {
        const int N0 = eb[j + 1] - eb[j], N = N0 << LM;
        const int32_t bit = (int32_t)bits[j] + balance;
        int32_t excess;
        if (N > 1) {
            excess = bit - cap[j] > 0 ? bit - cap[j] : 0;
            bits[j] = bit - excess;
            const int den = C * N + ((C == 2 && N > 2 && !*dual_stereo && j < *intensity) ? 1 : 0);
            const int NClogN = den * (T.log_n[j] + logM);
            int offset = (NClogN >> 1) - den * C2_FINE_OFFSET;
            if (N == 2) offset += den << C2_BITRES >> 2;
            if (bits[j] + offset < den * 2 << C2_BITRES) offset += NClogN >> 2;
            else if (bits[j] + offset < den * 3 << C2_BITRES) offset += NClogN >> 3;
            ebits[j] = c2_max(0, bits[j] + offset + (den << (C2_BITRES - 1)));
            ebits[j] = (ebits[j] / den) >> C2_BITRES;
            if (C * ebits[j] > (bits[j] >> C2_BITRES)) ebits[j] = bits[j] >> stereo >> C2_BITRES;
            ebits[j] = c2_min(ebits[j], C2_MAX_FINE_BITS);
            fine_priority[j] = ebits[j] * (den << C2_BITRES) >= bits[j] + offset;
            bits[j] -= C * ebits[j] << C2_BITRES;
        } else {
            excess = bit - (C << C2_BITRES) > 0 ? bit - (C << C2_BITRES) : 0;
            bits[j] = bit - excess;
            ebits[j] = 0;
            fine_priority[j] = 1;
        }
        if (excess > 0) {
            const int extra_fine = c2_min((int)(excess >> (stereo + C2_BITRES)), C2_MAX_FINE_BITS - ebits[j]);
            ebits[j] += extra_fine;
            const int extra_bits = extra_fine * C << C2_BITRES;
            fine_priority[j] = extra_bits >= excess - balance;
            excess -= extra_bits;
        }
        balance = excess;
    }
    *balance_out = balance;
    for (; j < end; j++) {
        ebits[j] = bits[j] >> stereo >> C2_BITRES;
        bits[j] = 0;
        fine_priority[j] = ebits[j] < 1;
    }
    return codedBands;
}

template <class Coder>
OPN_HD int c2_compute_allocation(Coder &ec, const Celt2Tabs &T, int end, const int *offsets, const int *cap, int alloc_trim, int *intensity,
                                 int *dual_stereo, int32_t total, int32_t *balance, int *pulses, int *ebits, int *fine_priority, int C, int LM)
{
    const int start = 0, len = C2_NBANDS;
    const uint8_t *eb = T.e_bands;
    int thresh[C2_NBANDS], trim_offset[C2_NBANDS], bits1[C2_NBANDS], bits2[C2_NBANDS];
    int skip_start = start, j;
    if (total < 0) total = 0;
    const int skip_rsv = total >= 1 << C2_BITRES ? 1 << C2_BITRES : 0;
    total -= skip_rsv;
    int intensity_rsv = 0, dual_stereo_rsv = 0;
    if (C == 2) {
        intensity_rsv = T.log2_frac[end - start];
        if (intensity_rsv > total) intensity_rsv = 0;
        else {
            total -= intensity_rsv;
            dual_stereo_rsv = total >= 1 << C2_BITRES ? 1 << C2_BITRES : 0;
            total -= dual_stereo_rsv;
        }
    }
    for (j = start; j < end; j++) {
        const int w = eb[j + 1] - eb[j];
        thresh[j] = c2_max(C << C2_BITRES, (3 * w << LM << C2_BITRES) >> 4);
        trim_offset[j] = C * w * (alloc_trim - 5 - LM) * (end - j - 1) * (1 << (LM + C2_BITRES)) >> 6;
        if (w << LM == 1) trim_offset[j] -= C << C2_BITRES;
    }
    int lo = 1, hi = C2_NB_ALLOC_VECTORS - 1;
    do {
        int done = 0, psum = 0;
        const int mid = (lo + hi) >> 1;
        for (j = end; j-- > start;) {
            const int N = eb[j + 1] - eb[j];
            int bitsj = C * N * T.alloc_vectors[mid * len + j] << LM >> 2;
            if (bitsj > 0) bitsj = c2_max(0, bitsj + trim_offset[j]);
            bitsj += offsets[j];
            if (bitsj >= thresh[j] || done) {
                done = 1;
                psum += c2_min(bitsj, cap[j]);
            } else if (bitsj >= C << C2_BITRES)
                psum += C << C2_BITRES;
        }
        if (psum > total) hi = mid - 1;
        else lo = mid + 1;
    } while (lo <= hi);
    hi = lo--;
    for (j = start; j < end; j++) {
        const int N = eb[j + 1] - eb[j];
        int bits1j = C * N * T.alloc_vectors[lo * len + j] << LM >> 2;
        int bits2j = hi >= C2_NB_ALLOC_VECTORS ? cap[j] : C * N * T.alloc_vectors[hi * len + j] << LM >> 2;
        if (bits1j > 0) bits1j = c2_max(0, bits1j + trim_offset[j]);
        if (bits2j > 0) bits2j = c2_max(0, bits2j + trim_offset[j]);
        if (lo > 0) bits1j += offsets[j];
        bits2j += offsets[j];
        if (offsets[j] > 0) skip_start = j;
        bits2j = c2_max(0, bits2j - bits1j);
        bits1[j] = bits1j;
        bits2[j] = bits2j;
    }
    return c2_interp_bits2pulses(ec, T, end, skip_start, bits1, bits2, thresh, cap, total, balance, skip_rsv, intensity, intensity_rsv,
                                 dual_stereo, dual_stereo_rsv, pulses, ebits, fine_priority, C, LM);
}

// ---- one band of one channel: split in halves down to PVQ leaves (libopus bands.c quant_partition, mono) -------
OPN_HD int c2_exp2_table8(int i)  // 2^(i/8) in Q14
{
    switch (i & 7) {
    case 0: return 16384;
    case 1: return 17866;
    case 2: return 19483;
    case 3: return 21247;
    case 4: return 23170;
    case 5: return 25267;
    case 6: return 27554;
    default: return 30048;
    }
}
OPN_HD int c2_compute_qn(int N, int b, int offset, int pulse_cap)
{
    const int N2 = 2 * N - 1;
    int qb = (b + N2 * offset) / N2;
    qb = c2_min(b - pulse_cap - (4 << C2_BITRES), qb);
    qb = c2_min(8 << C2_BITRES, qb);
    if (qb < (1 << C2_BITRES >> 1)) return 1;
    const int qn = c2_exp2_table8(qb) >> (14 - (qb >> C2_BITRES));
    return (qn + 1) >> 1 << 1;
}

// The recursion of quant_partition is at most LM+1 splits deep; it runs on an explicit stack of frames so that the same code
// serves the device lane and the host.  phase 0: entering, 1: first half done, 2: both done.
struct C2Frame {
    int base, N, b, B, LM, phase, mbits, sbits, itheta;
    int32_t reb;
    float gain, gmid, gside;
};
constexpr int C2_MAX_DEPTH = 6;

// Sink: put_part(base, n, k, index, gain) for every leaf that holds pulses.  `st`: C2_MAX_DEPTH frames of scratch owned by the
// caller and declared next to its allocation arrays.  (Declared in here, nvcc 12.9 gave the array the local-memory slot of the
// caller's still-live fine_priority[] once everything was inlined into the kernel: the final fine-energy pass then read frame
// images instead of priorities.  Found by the GPU parity test; the host build of the same source was right.)
template <class Coder, class Sink>
OPN_HD void c2_quant_band(Coder &ec, const Celt2Tabs &T, Sink &out, C2Frame *st, int band, int base0, int N0, int b0, int B0frame, int LM0,
                          int32_t &remaining_bits, uint32_t &n_splits, uint32_t &theta_sum)
{
    typedef C2Frame Frame;
    int sp = 0;
    st[0] = Frame{base0, N0, b0, B0frame, LM0, 0, 0, 0, 0, 0, 0.03125f, 0.f, 0.f};
    sp = 1;
    while (sp > 0) {
        Frame &f = st[sp - 1];
        if (f.phase == 0) {
            const uint8_t *cache = c2_pulse_cache(T, band, f.LM);
            if (f.LM != -1 && f.b > cache[cache[0]] + 12 && f.N > 2) {
                const int Bin = f.B;
                f.N >>= 1;
                f.LM -= 1;
                f.B = (f.B + 1) >> 1;
                const int pulse_cap = T.log_n[band] + f.LM * (1 << C2_BITRES);
                const int offset = (pulse_cap >> 1) - C2_QTHETA_OFFSET;
                const int qn = c2_compute_qn(f.N, f.b, offset, pulse_cap);
                const int32_t tell = (int32_t)ec.tell_frac();
                int itheta = 0;
                if (qn != 1) {
                    itheta = Bin > 1 ? (int)ec.uint_((uint32_t)qn + 1) : (int)ec.theta_tri((uint32_t)qn);
                    itheta = (int)(((int32_t)itheta * 16384) / qn);
                }
                const int32_t qalloc = (int32_t)ec.tell_frac() - tell;
                f.b -= qalloc;
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
                    imid = bitexact_cos((int16_t)itheta);                // src/math.rs:51-55
                    iside = bitexact_cos((int16_t)(16384 - itheta));
                    delta = frac_mul16((int16_t)((f.N - 1) << 7), (int16_t)bitexact_log2tan(iside, imid));  // math.rs:59-75
                }
                n_splits += 1;
                theta_sum += (uint32_t)itheta;
                f.gmid = f.gain * ((1.0f / 32768.0f) * (float)imid);
                f.gside = f.gain * ((1.0f / 32768.0f) * (float)iside);
                if (Bin > 1 && (itheta & 0x3fff)) {
                    if (itheta > 8192) delta -= delta >> (4 - f.LM);
                    else delta = c2_min(0, delta + (f.N << C2_BITRES >> (5 - f.LM)));
                }
                f.mbits = c2_max(0, c2_min(f.b, (f.b - delta) / 2));
                f.sbits = f.b - f.mbits;
                f.itheta = itheta;
                remaining_bits -= qalloc;
                f.reb = remaining_bits;
                f.phase = 1;
                // the half with more bits goes first
                if (f.mbits >= f.sbits) st[sp] = Frame{f.base, f.N, f.mbits, f.B, f.LM, 0, 0, 0, 0, 0, f.gmid, 0.f, 0.f};
                else st[sp] = Frame{f.base + f.N, f.N, f.sbits, f.B, f.LM, 0, 0, 0, 0, 0, f.gside, 0.f, 0.f};
                sp++;
            } else {  // leaf
                int q = c2_bits2pulses(T, band, f.LM, f.b);
                int curr_bits = c2_pulses2bits(T, band, f.LM, q);
                remaining_bits -= curr_bits;
                while (remaining_bits < 0 && q > 0) {
                    remaining_bits += curr_bits;
                    q--;
                    curr_bits = c2_pulses2bits(T, band, f.LM, q);
                    remaining_bits -= curr_bits;
                }
                if (q != 0) {
                    const uint32_t K = c2_get_pulses((uint32_t)q);
                    const uint32_t index = ec.pulses_index(c2_pvq_v(T, (uint32_t)f.N, K));  // decode_pulses' decode_uint (pvc.rs:156-160)
                    out.put_part(f.base | band << C2_BAND_SHIFT, f.N, (int)K, index, f.gain);
                }
                sp--;
            }
        } else if (f.phase == 1) {
            const int32_t used = f.reb - remaining_bits;
            if (f.mbits >= f.sbits) {
                const int32_t rebalance = f.mbits - used;
                if (rebalance > 3 << C2_BITRES && f.itheta != 0) f.sbits += rebalance - (3 << C2_BITRES);
                f.phase = 2;
                st[sp] = Frame{f.base + f.N, f.N, f.sbits, f.B, f.LM, 0, 0, 0, 0, 0, f.gside, 0.f, 0.f};
            } else {
                const int32_t rebalance = f.sbits - used;
                if (rebalance > 3 << C2_BITRES && f.itheta != 16384) f.mbits += rebalance - (3 << C2_BITRES);
                f.phase = 2;
                st[sp] = Frame{f.base, f.N, f.mbits, f.B, f.LM, 0, 0, 0, 0, 0, f.gmid, 0.f, 0.f};
            }
            sp++;
        } else {
            sp--;
        }
    }
}

// One frame.  `len` = payload bytes.  Coder: tell(), tell_frac(), final_rng(), bit_logp(logp, p1_permille),
// icdf(tab, ftb, n_sym), uint_(ft), bits(n), laplace(band), theta_tri(qn), pulses_index(ft), transient_permille().
// Sink: put_part(base | band << C2_BAND_SHIFT, n, k, index, gain), put_sign(base | band << C2_BAND_SHIFT, sign),
// energy_set / energy_add / energy (c, band, Q9).  sd may be null (device batch path).
template <class Coder, class Sink> OPN_HD void celt2_frame(Coder &ec, const Celt2Tabs &T, uint32_t len, int LM, int C, Celt2Side *sd, Sink &out,
                                                           uint32_t &hdr_flags, uint32_t &n_pulses_out)
{
    const int nf = 120 << LM, end = C2_NBANDS, M = 1 << LM;
    const uint8_t *eb = T.e_bands;
    const uint8_t tapset_icdf[3] = {2, 1, 0}, spread_icdf[4] = {25, 23, 2, 0};
    const uint8_t trim_icdf[11] = {126, 124, 119, 109, 87, 41, 19, 9, 4, 2, 0};
    const int32_t total_bits = (int32_t)len * 8;
    uint32_t n_pulses = 0, n_splits = 0, theta_sum = 0;
    hdr_flags = 0;
    n_pulses_out = 0;
    const uint32_t silence = ec.bit_logp(15, 0);
    if (sd) sd->silence = (int32_t)silence;
    if (silence) {
        hdr_flags = 1u;
        return;
    }
    const uint32_t postfilter = ec.bit_logp(1, 500);
    uint32_t octave = 0, period = 0, gain_idx = 0, tapset = 0;
    if (postfilter) {
        octave = ec.uint_(6);
        period = (16u << octave) + ec.bits(4 + octave) - 1u;
        gain_idx = ec.bits(3);
        tapset = ec.icdf(tapset_icdf, 2, 3);
    }
    const uint32_t transient = ec.bit_logp(3, ec.transient_permille());
    const uint32_t intra = ec.bit_logp(3, 125);
    hdr_flags = postfilter << 1 | transient << 2 | intra << 3 | tapset << 4 | gain_idx << 8 | octave << 12 | period << 16;
    if (sd) {
        sd->postfilter = (int32_t)postfilter;
        sd->octave = (int32_t)octave;
        sd->period = (int32_t)period;
        sd->gain_idx = (int32_t)gain_idx;
        sd->tapset = (int32_t)tapset;
        sd->transient = (int32_t)transient;
        sd->intra = (int32_t)intra;
    }
    for (int b = 0; b < C2_NBANDS; b++)
        for (int c = 0; c < C; c++) {
            const int32_t v = ec.laplace(b);
            if (sd) sd->coarse[c][b] = v;
            out.energy_set(c, b, c2_energy_coarse_q9(v));
        }
    const uint32_t spread = ec.icdf(spread_icdf, 5, 4);
    if (sd) sd->spread = (int32_t)spread;
    // band boosts (RFC 6716 4.3.3)
    int cap[C2_NBANDS], offsets[C2_NBANDS];
    for (int i = 0; i < C2_NBANDS; i++) {
        const int N = (eb[i + 1] - eb[i]) << LM;
        cap[i] = (T.cache_caps[C2_NBANDS * (2 * LM + C - 1) + i] + 64) * C * N >> 2;
    }
    int alloc_trim;
    {
        int dynalloc_logp = 6;
        int32_t total_frac = total_bits << C2_BITRES;
        int32_t tell = (int32_t)ec.tell_frac();
        for (int i = 0; i < end; i++) {
            const int width = C * (eb[i + 1] - eb[i]) << LM;
            const int quanta = c2_min(width << C2_BITRES, c2_max(6 << C2_BITRES, width));
            int loop_logp = dynalloc_logp, boost = 0;
            while (tell + (loop_logp << C2_BITRES) < total_frac && boost < cap[i]) {
                const uint32_t flag = ec.bit_logp((uint32_t)loop_logp, 30);
                tell = (int32_t)ec.tell_frac();
                if (!flag) break;
                boost += quanta;
                total_frac -= quanta;
                loop_logp = 1;
            }
            offsets[i] = boost;
            if (sd) sd->offsets[i] = boost;
            if (boost > 0) dynalloc_logp = c2_max(2, dynalloc_logp - 1);
        }
        alloc_trim = tell + (6 << C2_BITRES) <= total_frac ? (int)ec.icdf(trim_icdf, 7, 11) : 5;
    }
    int32_t bits = (total_bits << C2_BITRES) - (int32_t)ec.tell_frac() - 1;
    const int anti_collapse_rsv = transient && LM >= 2 && bits >= ((LM + 2) << C2_BITRES) ? (1 << C2_BITRES) : 0;
    bits -= anti_collapse_rsv;
    int pulses[C2_NBANDS], ebits[C2_NBANDS], fine_priority[C2_NBANDS], intensity = 0, dual_stereo = 0;
    int32_t balance = 0;
    const int codedBands = c2_compute_allocation(ec, T, end, offsets, cap, alloc_trim, &intensity, &dual_stereo, bits, &balance, pulses, ebits,
                                                 fine_priority, C, LM);
    if (sd) {
        sd->alloc_trim = alloc_trim;
        sd->coded_bands = codedBands;
        sd->intensity = intensity;
        sd->dual_stereo = dual_stereo;
        sd->balance = balance;
        for (int i = 0; i < C2_NBANDS; i++) {
            sd->pulses[i] = pulses[i];
            sd->ebits[i] = ebits[i];
            sd->fine_priority[i] = fine_priority[i];
        }
    }
    for (int i = 0; i < end; i++)  // fine energy
        if (ebits[i] > 0)
            for (int c = 0; c < C; c++) {
                const uint32_t v = ec.bits((uint32_t)ebits[i]);
                if (sd) sd->fine[c][i] = (int32_t)v;
                out.energy_add(c, i, c2_energy_fine_q9(v, ebits[i]));
            }
    C2Frame frames[C2_MAX_DEPTH];
    {
        const int32_t band_total = (total_bits << C2_BITRES) - anti_collapse_rsv;
        const int Bframe = transient ? M : 1;
        for (int i = 0; i < end; i++) {
            const int32_t tell = (int32_t)ec.tell_frac();
            if (i != 0) balance -= tell;
            int32_t remaining_bits = band_total - tell - 1;
            int b = 0;
            if (i <= codedBands - 1) {
                const int32_t curr_balance = balance / c2_min(3, codedBands - i);
                b = c2_max(0, c2_min(16383, c2_min((int)remaining_bits + 1, pulses[i] + (int)curr_balance)));
            }
            const int N = (eb[i + 1] - eb[i]) << LM, base = eb[i] << LM;
            for (int c = 0; c < C; c++) {
                if (N == 1) {  // one sign bit if it fits
                    uint32_t sign = 0;
                    if (remaining_bits >= 1 << C2_BITRES) {
                        sign = ec.bits(1);
                        remaining_bits -= 1 << C2_BITRES;
                    }
                    out.put_sign((c * nf + base) | i << C2_BAND_SHIFT, sign);
                    n_pulses += 1;
                } else {
                    c2_quant_band(ec, T, out, frames, i, c * nf + base, N, b / C, Bframe, LM, remaining_bits, n_splits, theta_sum);
                }
            }
            balance += pulses[i] + tell;
        }
    }
    uint32_t anti_collapse = 0;
    if (anti_collapse_rsv > 0) anti_collapse = ec.bits(1);
    if (sd) sd->anti_collapse = (int32_t)anti_collapse;
    {  // left-over whole bits refine the energies, priority 0 bands first
        int bits_left = (int)total_bits - (int)ec.tell();
        for (int prio = 0; prio < 2; prio++)
            for (int i = 0; i < end && bits_left >= C; i++) {
                if (ebits[i] >= C2_MAX_FINE_BITS || fine_priority[i] != prio) continue;
                for (int c = 0; c < C; c++) {
                    const uint32_t v = ec.bits(1);
                    if (sd) sd->fine_final[c][i] = 1 + (int32_t)v;
                    out.energy_add(c, i, c2_energy_final_q9(v, ebits[i]));
                    bits_left--;
                }
            }
    }
    n_pulses_out = n_pulses + out.pulses();
    if (sd) {
        for (int c = 0; c < C; c++)
            for (int i = 0; i < C2_NBANDS; i++) sd->energy_q9[c][i] = out.energy(c, i);
        sd->n_splits = n_splits;
        sd->theta_sum = theta_sum;
#ifdef OPN_C2_DEBUG
        for (int i = 0; i < C2_NBANDS; i++) sd->offsets[i] = ebits[i] | fine_priority[i] << 8 | pulses[i] << 16;
        sd->balance = (int32_t)total_bits - (int32_t)ec.tell();
#endif
    }
}

}  // namespace opn
