// rangedec.cuh -- warp-per-stream entropy decoder for sm_100a.
//
// Device mirror of the reference's RangeDecoder (src/range_coder/decoder.rs:10-355, constants
// src/range_coder/mod.rs:48-117) and of decode_pulses/cwrsi (src/celt/pvc.rs:156-298).
//
// Execution model: ONE WARP owns one packet.  The coder state (rng, val, ...) is warp-uniform:
// all 32 lanes hold the same registers and take the same branches, so there is no divergence
// and no shuffle traffic for the serial part.  The packet bytes live in shared memory (staged
// once, coalesced), so the front/back byte reads are bank-broadcast loads.  Lanes only differ
// where there is data parallelism: staging the packet, the PVQ zero-run scan, writing pulse
// vectors and coefficients.
#pragma once
#include <stdint.h>

namespace opn {

// src/range_coder/mod.rs:48-68
constexpr uint32_t RC_UINT_BITS = 8, RC_BITRES = 3, RC_WINDOW_SIZE = 32, RC_SYM_BITS = 8, RC_CODE_BITS = 32;
constexpr uint32_t RC_SYM_MAX = (1u << RC_SYM_BITS) - 1u;
constexpr uint32_t RC_CODE_TOP = 1u << (RC_CODE_BITS - 1u);
constexpr uint32_t RC_CODE_BOT = RC_CODE_TOP >> RC_SYM_BITS;
constexpr uint32_t RC_CODE_EXTRA = (RC_CODE_BITS - 2u) % RC_SYM_BITS + 1u;

__device__ __forceinline__ uint32_t rc_ilog(uint32_t x) { return 32u - (uint32_t)__clz((int)x); }  // math.rs:5-7

// src/range_coder/mod.rs:114-117
__device__ __forceinline__ uint32_t laplace_freq1(uint32_t fs0, uint32_t decay)
{
    uint32_t ft = 32768u - 32u - fs0;
    return (ft * (16384u - decay)) >> 15;
}

// floor(a / b) for quotients below 2^16 (the range decoder's `val / ext`: val < rng always holds, so
// the quotient never exceeds ft + ft/ext).  One MUFU.RCP estimate, deliberately scaled down by
// 2^-20 so that it never overshoots, then a single upward correction -- exact, ~9 instructions
// instead of the ~22 of the generic 32-bit division.
__device__ __forceinline__ uint32_t div_small_quotient(uint32_t a, uint32_t b)
{
    float r;
#ifdef OPN_HOST_SHIM  // host build for tests/host_shim: any estimate within the correction step's reach gives the same quotient
    r = 1.0f / __uint2float_rn(b);
#else
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__uint2float_rn(b)));
#endif
    uint32_t q = __float2uint_rz(__uint2float_rn(a) * r * 0.999999f);
    if (a - q * b >= b) q += 1u;
    return q;
}

// floor(a / d) for an invariant divisor d given its round-up reciprocal (Granlund-Montgomery):
// magic = floor(2^32 * (2^sh - d) / d) + 1, sh = ceil(log2 d).  Exact for every 32-bit a.
__device__ __forceinline__ uint32_t div_magic(uint32_t a, uint32_t magic, uint32_t sh)
{
    uint32_t t = __umulhi(magic, a);
    return (t + ((a - t) >> 1)) >> (sh - 1u);
}

struct RangeDec {
    const uint8_t *buf;  // shared memory
    uint32_t storage, end_offs, end_window, end_bits, bits_total, offs, rng, val, ext, rem;

    // decoder.rs:86-94
    __device__ __forceinline__ uint32_t read_byte()
    {
        if (offs < storage) return buf[offs++];
        return 0u;
    }
    // decoder.rs:97-104
    __device__ __forceinline__ uint32_t read_byte_from_end()
    {
        if (end_offs < storage) {
            end_offs += 1u;
            return buf[storage - end_offs];
        }
        return 0u;
    }
    // decoder.rs:108-122
    __device__ __forceinline__ void normalize()
    {
        while (rng <= RC_CODE_BOT) {
            bits_total += RC_SYM_BITS;
            rng <<= RC_SYM_BITS;
            uint32_t symbol = rem;
            rem = read_byte();
            symbol = ((symbol << RC_SYM_BITS) | rem) >> (RC_SYM_BITS - RC_CODE_EXTRA);
            val = ((val << RC_SYM_BITS) + (RC_SYM_MAX & ~symbol)) & (RC_CODE_TOP - 1u);
        }
    }
    // decoder.rs:50-78
    __device__ __forceinline__ void init(const uint8_t *b, uint32_t len)
    {
        buf = b;
        storage = len;
        end_offs = 0u;
        end_window = 0u;
        end_bits = 0u;
        bits_total = RC_CODE_BITS + 1u - ((RC_CODE_BITS - RC_CODE_EXTRA) / RC_SYM_BITS) * RC_SYM_BITS;
        offs = 0u;
        rng = 1u << RC_CODE_EXTRA;
        ext = 0u;
        rem = read_byte();
        val = rng - 1u - (rem >> (RC_SYM_BITS - RC_CODE_EXTRA));
        normalize();
    }
    // decoder.rs:81-83
    __device__ __forceinline__ void shrink_storage(uint32_t by) { storage -= by; }
    // decoder.rs:143-147
    __device__ __forceinline__ uint32_t decode(uint32_t ft)
    {
        ext = rng / ft;
        uint32_t s = val / ext;
        return ft - min(s + 1u, ft);
    }
    // decoder.rs:150-154
    __device__ __forceinline__ uint32_t decode_bin(uint32_t bits)
    {
        ext = rng >> bits;
        uint32_t s = val / ext;
        return (1u << bits) - min(s + 1u, 1u << bits);
    }
    // decode() for ft <= 2^15 (decoder.rs:143-147) with the cheap exact quotient
    __device__ __forceinline__ uint32_t decode_small(uint32_t ft)
    {
        ext = rng / ft;
        uint32_t s = div_small_quotient(val, ext);
        return ft - min(s + 1u, ft);
    }
    // decode_bin (decoder.rs:150-154) with the cheap exact quotient (bits <= 15)
    __device__ __forceinline__ uint32_t decode_bin_small(uint32_t bits)
    {
        ext = rng >> bits;
        uint32_t s = div_small_quotient(val, ext);
        return (1u << bits) - min(s + 1u, 1u << bits);
    }
    // decode_uint (decoder.rs:245-266) for an alphabet whose split (ft1, ftb) and reciprocal were
    // precomputed on the host: same arithmetic, no runtime ilog and no generic division.
    __device__ __forceinline__ uint32_t uint_precomputed(uint32_t ft_minus1, uint32_t ft1, uint32_t ftb, uint32_t magic,
                                                         uint32_t sh)
    {
        ext = div_magic(rng, magic, sh);                 // rng / ft1
        uint32_t q = div_small_quotient(val, ext);       // val / ext
        uint32_t s = ft1 - min(q + 1u, ft1);             // decode(ft1)
        update(s, s + 1u, ft1);
        if (ftb == 0u) return s;
        uint32_t t = (s << ftb) | bits(ftb);
        return t <= ft_minus1 ? t : ft_minus1;
    }
    // decoder.rs:172-181
    __device__ __forceinline__ void update(uint32_t fl, uint32_t fh, uint32_t ft)
    {
        uint32_t s = ext * (ft - fh);
        val -= s;
        rng = fl > 0u ? ext * (fh - fl) : rng - s;
        normalize();
    }
    // decoder.rs:184-195
    __device__ __forceinline__ uint32_t bit_logp(uint32_t logp)
    {
        uint32_t r = rng, d = val, s = r >> logp;
        uint32_t ret = d < s ? 1u : 0u;
        if (!ret) val = d - s;
        rng = ret ? s : r - s;
        normalize();
        return ret;
    }
    // decoder.rs:210-232.  icdf may point to shared, global or constant memory.
    __device__ __forceinline__ uint32_t icdf(const uint8_t *tab, uint32_t ftb)
    {
        uint32_t s = rng, d = val, r = s >> ftb, t, ret = 0u;
        for (;;) {
            t = s;
            s = r * (uint32_t)tab[ret];
            if (d >= s) break;
            ret += 1u;
        }
        val = d - s;
        rng = t - s;
        normalize();
        return ret;
    }
    // decoder.rs:279-303
    __device__ __forceinline__ uint32_t bits(uint32_t nbits)
    {
        uint32_t window = end_window, available = end_bits;
        if (available < nbits) {
            do {
                window |= read_byte_from_end() << available;
                available += RC_SYM_BITS;
            } while (available <= RC_WINDOW_SIZE - RC_SYM_BITS);
        }
        uint32_t ret = window & ((1u << nbits) - 1u);
        window >>= nbits;
        available -= nbits;
        end_window = window;
        end_bits = available;
        bits_total += nbits;
        return ret;
    }
    // decoder.rs:245-266
    __device__ __forceinline__ uint32_t uint(uint32_t ft)
    {
        ft -= 1u;
        uint32_t ftb = rc_ilog(ft);
        if (ftb > RC_UINT_BITS) {
            ftb -= RC_UINT_BITS;
            uint32_t ft1 = (ft >> ftb) + 1u;
            uint32_t s = decode(ft1);
            update(s, s + 1u, ft1);
            uint32_t t = (s << ftb) | bits(ftb);
            return t <= ft ? t : ft;  // corrupt frame saturates (decoder.rs:255-259)
        }
        ft += 1u;
        uint32_t s = decode(ft);
        update(s, s + 1u, ft);
        return s;
    }
    // decoder.rs:314-355
    __device__ __forceinline__ int32_t laplace(uint32_t fs, uint32_t decay)
    {
        int32_t v = 0;
        uint32_t fm = decode_bin_small(15u);
        uint32_t fl = 0u;
        if (fm >= fs) {
            v += 1;
            fl = fs;
            fs = laplace_freq1(fs, decay) + 1u;
            while (fs != 0u && fm >= fl + 2u * fs) {
                fs *= 2u;
                fl += fs;
                fs = ((fs - 2u) * decay) >> 15;
                fs += 1u;
                v += 1;
            }
            if (fs <= 1u) {
                uint32_t di = (fm - fl) >> 1;
                v += (int32_t)di;
                fl += 2u * di;
            }
            if (fm < fl + fs) v = -v;
            else fl += fs;
        }
        update(fl, min(fl + fs, 32768u), 32768u);
        return v;
    }
    // src/range_coder/mod.rs:84-86
    __device__ __forceinline__ uint32_t tell() const { return bits_total - rc_ilog(rng); }
    // src/range_coder/mod.rs:96-111
    __device__ __forceinline__ uint32_t tell_frac() const
    {
        uint32_t nbits = bits_total << RC_BITRES;
        uint32_t l = rc_ilog(rng);
        uint32_t r = rng >> (l - 16u);
        uint32_t b = (r >> 12) - 8u;
        // correction = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535}
        uint32_t c = b == 0u ? 35733u : b == 1u ? 38967u : b == 2u ? 42495u : b == 3u ? 46340u
                   : b == 4u ? 50535u : b == 5u ? 55109u : b == 6u ? 60097u : 65535u;
        if (r > c) b += 1u;
        l = (l << 3) + b;
        return nbits - l;
    }
};

// ---------------------------------------------------------------------------------------------
// LaneDec: the same decoder for the lane-per-packet kernel (k_synth_rangedec), where the serial dependency chain of one
// packet IS the kernel's run time.  Same arithmetic, value for value, as RangeDec / the reference; what changes is how
// the bytes get there and that nothing on the chain branches:
//   * normalize (decoder.rs:108-122) runs its 0..3 iterations at once: the number of byte shifts follows from rng alone,
//     the k input bytes come out of a register (a 64-bit LSB-first shift register refilled one aligned 32-bit word at a
//     time, the word fetched long before it is needed), and the k chained updates
//     val = ((val << 8) + (255 & ~sym)) & (2^31 - 1) collapse into one shift-and-mask because the symbols of consecutive
//     iterations are consecutive 8-bit windows, one bit apart, of the string [rem | b1 .. bk];
//   * decode_bits (decoder.rs:279-303) reads from a 64-bit window refilled 32 bits at a time from the end of the packet;
//   * bytes past `storage` read as zero on both ends, as read_byte / read_byte_from_end define it;
//   * decode_laplace (decoder.rs:314-355) finds the magnitude by counting the thresholds it reaches in a table of the
//     loop's (fl, fs) states (laplace_table below) instead of walking the loop.
// The device words that hold packet bytes are read with aligned 32-bit loads: up to 3 bytes before the packet start and
// after its end are touched (and masked off), never a word that holds no packet byte.
constexpr int LAP_N = 16;      // tabulated magnitudes; larger ones continue the reference's loop
constexpr int LAP_FIRST = 9;   // thresholds every call compares against before it looks at the rest (magnitudes up to 8 stop here)

// (fl, fs) of decode_laplace after the magnitude loop stopped at magnitude v, v = 0..LAP_N (decoder.rs:319-337):
// fl[0] = 0, fs[0] = fs0; fl[1] = fs0, fs[1] = freq1 + 1; fl[v+1] = fl[v] + 2 fs[v], fs[v+1] = (((2 fs[v] - 2) decay) >> 15) + 1.
// The loop runs while fm >= fl[v] + 2 fs[v] = fl[v+1], so the magnitude is the number of fl[1..] that fm reaches.
__device__ __forceinline__ void laplace_table(uint32_t fs0, uint32_t decay, uint32_t *fl_tab, uint32_t *fs_tab)
{
    uint32_t fl = fs0, fs = laplace_freq1(fs0, decay) + 1u;
    fl_tab[0] = 0u;
    fs_tab[0] = fs0;
    for (int v = 1; v <= LAP_N; v++) {
        fl_tab[v] = fl;
        fs_tab[v] = fs;
        fs *= 2u;
        fl += fs;
        fs = (((fs - 2u) * decay) >> 15) + 1u;
    }
}

struct LaneDec {
    const uint8_t *src;
    uint32_t storage;
    uint64_t fbuf;        // front reader: the next fn bytes of the packet, first byte in bits 0..7
    uint32_t fn;
    const uint32_t *fw;   // next aligned word to fetch for the front reader
    int32_t foff;         // packet offset of that word's first byte
    uint32_t fpend;       // the word before it, already loaded (consumed by the next refill: its load latency is off the chain)
    uint64_t bbuf;        // back reader: the next bn raw bits, first bit in bit 0
    uint32_t bn;
    int32_t bhi;          // packet bytes [0, bhi) have not been handed to the back reader yet
    uint32_t bpend;       // the next 32 raw bits, already loaded and put in order
    uint32_t bits_total, rng, val, ext, rem;

    // One aligned word of the packet, bytes past `storage` zeroed (packet offset of its first byte: off >= 0).
    __device__ __forceinline__ uint32_t front_load(const uint32_t *p, int32_t off) const
    {
        uint32_t w = 0u;
        if (off < (int32_t)storage) {  // the word holds at least one packet byte
            w = __ldg(p);
            const int32_t keep = (int32_t)storage - off;
            if (keep < 4) w &= (1u << (8 * keep)) - 1u;
        }
        return w;
    }
    __device__ __forceinline__ void front_fetch()
    {
        fbuf |= (uint64_t)fpend << (8u * fn);  // loaded one refill ago
        fn += 4u;
        fpend = front_load(fw, foff);
        fw += 1;
        foff += 4;
    }
    __device__ __forceinline__ void back_fetch()
    {
        bbuf |= (uint64_t)bpend << bn;  // loaded one refill ago
        bn += 32u;
        bpend = back_load();
    }
    // the next four bytes from the end, first one in bits 0..7; zeros once the packet start is passed
    __device__ __forceinline__ uint32_t back_load()
    {
        uint32_t r = 0u;
        if (bhi >= 4) {
            const uint8_t *q = src + (bhi - 4);
            const uint32_t *base = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
            const uint32_t shb = ((uint32_t)reinterpret_cast<uintptr_t>(q) & 3u) * 8u;
            const uint32_t lo = __ldg(base);
            const uint32_t hi = shb ? __ldg(base + 1) : 0u;
            r = __byte_perm(__funnelshift_r(lo, hi, shb), 0u, 0x0123);  // q[3] is read first (decoder.rs:97-104)
        } else if (bhi > 0) {  // the first 1..3 bytes of the packet, then zeros
            for (int j = 0; j < bhi; j++) r |= (uint32_t)src[bhi - 1 - j] << (8 * j);
        }
        bhi -= 4;
        return r;
    }
    // decoder.rs:108-122, all iterations at once
    __device__ __forceinline__ void normalize()
    {
        const uint32_t k = (rng <= (1u << 23) ? 1u : 0u) + (rng <= (1u << 15) ? 1u : 0u) + (rng <= (1u << 7) ? 1u : 0u);
        const uint32_t t = __byte_perm((uint32_t)fbuf, 0u, 0x4012);    // b1 << 16 | b2 << 8 | b3
        const uint32_t w = ((rem << 24) | t) >> (8u * (3u - k));       // [rem | b1 .. bk]
        const uint32_t sh = 8u * k;
        val = ((val << sh) + (~(w >> 1) & ((1u << sh) - 1u))) & (RC_CODE_TOP - 1u);
        rng <<= sh;
        bits_total += sh;
        rem = w & 255u;
        fbuf >>= sh;
        fn -= k;
        if (fn <= 4u) front_fetch();
    }
    // decoder.rs:50-78
    __device__ __forceinline__ void init(const uint8_t *b, uint32_t len)
    {
        src = b;
        storage = len;
        const uint32_t a = (uint32_t)reinterpret_cast<uintptr_t>(b) & 3u;
        fw = reinterpret_cast<const uint32_t *>(b - a);
        foff = -(int32_t)a;
        fbuf = 0ull;
        fn = 0u;
        {  // first word: the a bytes before the packet are dropped, bytes past its end are zero
            uint32_t w = len ? __ldg(fw) : 0u;
            if (a + len < 4u) w &= (1u << (8u * (a + len))) - 1u;
            fbuf = (uint64_t)(w >> (8u * a));
            fn = 4u - a;
            fw += 1;
            foff += 4;
        }
        fpend = front_load(fw, foff);
        fw += 1;
        foff += 4;
        front_fetch();
        bbuf = 0ull;
        bn = 0u;
        bhi = (int32_t)len;
        bpend = back_load();
        back_fetch();
        back_fetch();
        bits_total = RC_CODE_BITS + 1u - ((RC_CODE_BITS - RC_CODE_EXTRA) / RC_SYM_BITS) * RC_SYM_BITS;
        rng = 1u << RC_CODE_EXTRA;
        ext = 0u;
        rem = (uint32_t)fbuf & 255u;  // read_byte
        fbuf >>= 8;
        fn -= 1u;
        val = rng - 1u - (rem >> (RC_SYM_BITS - RC_CODE_EXTRA));
        normalize();
    }
    // decoder.rs:172-181
    __device__ __forceinline__ void update(uint32_t fl, uint32_t fh, uint32_t ft)
    {
        const uint32_t s = ext * (ft - fh);
        val -= s;
        rng = fl > 0u ? ext * (fh - fl) : rng - s;
        normalize();
    }
    // decode_bin (decoder.rs:150-154), bits <= 15
    __device__ __forceinline__ uint32_t decode_bin_small(uint32_t bits)
    {
        ext = rng >> bits;
        const uint32_t s = div_small_quotient(val, ext);
        return (1u << bits) - min(s + 1u, 1u << bits);
    }
    // decode_uint (decoder.rs:245-266) for ft <= 2^8: decode + update, no raw bits
    __device__ __forceinline__ uint32_t uint_small(uint32_t ft)
    {
        ext = rng / ft;
        const uint32_t q = div_small_quotient(val, ext);
        const uint32_t s = ft - min(q + 1u, ft);
        update(s, s + 1u, ft);
        return s;
    }
    // decode (decoder.rs:143-147), any ft
    __device__ __forceinline__ uint32_t decode(uint32_t ft)
    {
        ext = rng / ft;
        const uint32_t s = val / ext;
        return ft - min(s + 1u, ft);
    }
    // decode_uint (decoder.rs:245-266), any ft >= 2
    __device__ __forceinline__ uint32_t uint_any(uint32_t ft)
    {
        ft -= 1u;
        uint32_t ftb = rc_ilog(ft);
        if (ftb > RC_UINT_BITS) {
            ftb -= RC_UINT_BITS;
            const uint32_t ft1 = (ft >> ftb) + 1u;
            const uint32_t s = decode(ft1);
            update(s, s + 1u, ft1);
            const uint32_t t = (s << ftb) | bits(ftb);
            return t <= ft ? t : ft;  // corrupt frame saturates (decoder.rs:255-259)
        }
        ft += 1u;
        const uint32_t s = decode(ft);
        update(s, s + 1u, ft);
        return s;
    }
    // src/range_coder/mod.rs:84-86
    __device__ __forceinline__ uint32_t tell() const { return bits_total - rc_ilog(rng); }
    // decode_uint with the alphabet split and reciprocal precomputed on the host (see RangeDec::uint_precomputed)
    __device__ __forceinline__ uint32_t uint_precomputed(uint32_t ft_minus1, uint32_t ft1, uint32_t ftb, uint32_t magic, uint32_t sh)
    {
        ext = div_magic(rng, magic, sh);
        const uint32_t q = div_small_quotient(val, ext);
        const uint32_t s = ft1 - min(q + 1u, ft1);
        update(s, s + 1u, ft1);
        if (ftb == 0u) return s;
        const uint32_t t = (s << ftb) | bits(ftb);
        return t <= ft_minus1 ? t : ft_minus1;  // corrupt frame saturates (decoder.rs:255-259)
    }
    // decoder.rs:184-195
    __device__ __forceinline__ uint32_t bit_logp(uint32_t logp)
    {
        const uint32_t r = rng, d = val, s = r >> logp;
        const uint32_t ret = d < s ? 1u : 0u;
        val = ret ? d : d - s;
        rng = ret ? s : r - s;
        normalize();
        return ret;
    }
    // decoder.rs:210-232
    __device__ __forceinline__ uint32_t icdf(const uint8_t *tab, uint32_t ftb)
    {
        uint32_t s = rng, d = val, r = s >> ftb, t, ret = 0u;
        for (;;) {
            t = s;
            s = r * (uint32_t)tab[ret];
            if (d >= s) break;
            ret += 1u;
        }
        val = d - s;
        rng = t - s;
        normalize();
        return ret;
    }
    // decoder.rs:279-303, nbits <= 25
    __device__ __forceinline__ uint32_t bits(uint32_t nbits)
    {
        const uint32_t ret = (uint32_t)bbuf & ((1u << nbits) - 1u);
        bbuf >>= nbits;
        bn -= nbits;
        bits_total += nbits;
        if (bn <= 32u) back_fetch();
        return ret;
    }
    // decoder.rs:314-355 on the tables of laplace_table (same fs0, decay)
    __device__ __forceinline__ int32_t laplace(const uint32_t *fl_tab, const uint32_t *fs_tab, uint32_t decay)
    {
        const uint32_t fm = decode_bin_small(15u);
        uint32_t v = 0u;
#pragma unroll
        for (int j = 1; j <= LAP_FIRST; j++) v += fm >= fl_tab[j] ? 1u : 0u;
        if (v == (uint32_t)LAP_FIRST) {
#pragma unroll
            for (int j = LAP_FIRST + 1; j <= LAP_N; j++) v += fm >= fl_tab[j] ? 1u : 0u;
        }
        uint32_t fl = fl_tab[v], fs = fs_tab[v];
        if (v == (uint32_t)LAP_N) {  // beyond the table: the reference's loop, continued
            while (fm >= fl + 2u * fs) {
                fs *= 2u;
                fl += fs;
                fs = (((fs - 2u) * decay) >> 15) + 1u;
                v += 1u;
            }
        }
        int32_t ret = (int32_t)v;
        if (v) {
            if (fm < fl + fs) ret = -ret;
            else fl += fs;
        }
        update(fl, min(fl + fs, 32768u), 32768u);
        return ret;
    }
    // src/range_coder/mod.rs:96-111
    __device__ __forceinline__ uint32_t tell_frac() const
    {
        const uint32_t nbits = bits_total << RC_BITRES;
        uint32_t l = rc_ilog(rng);
        const uint32_t r = rng >> (l - 16u);
        uint32_t b = (r >> 12) - 8u;
        const uint32_t c = b == 0u ? 35733u : b == 1u ? 38967u : b == 2u ? 42495u : b == 3u ? 46340u
                         : b == 4u ? 50535u : b == 5u ? 55109u : b == 6u ? 60097u : 65535u;
        if (r > c) b += 1u;
        l = (l << 3) + b;
        return nbits - l;
    }
};

// ---------------------------------------------------------------------------------------------
// PVQ codeword expansion.  U(n,k) table in shared memory: rowoff[15] + data[1272]
// (src/celt/pvc.rs:301-429).
struct PvqTable {
    const uint32_t *data;
    const uint16_t *row;
    __device__ __forceinline__ uint32_t u(uint32_t n, uint32_t k) const  // pvc.rs:295-298
    {
        uint32_t lo = min(n, k), hi = max(n, k);
        return data[row[lo] + hi];
    }
    __device__ __forceinline__ uint32_t v(uint32_t n, uint32_t k) const { return u(n, k) + u(n, k + 1u); }  // pvc.rs:289-291
};

#ifndef OPN_HOST_SHIM  // the warp-cooperative cwrsi below needs a warp
__device__ __forceinline__ uint32_t sat_add_u32(uint32_t a, uint32_t b)
{
    uint32_t s = a + b;
    return s < a ? 0xFFFFFFFFu : s;
}

// cwrsi (pvc.rs:182-284) executed by a full warp.  y (shared memory, n ints) receives the pulse
// vector; returns yy = sum y^2 (exact in f32: < 2^24).
//
// The "lots of dimensions" branch (k < n, pvc.rs:232-258) walks runs of empty dimensions one at
// a time in the reference: while U(k,n) <= i < U(k+1,n) { i -= U(k,n); y = 0; n -= 1 }.  Here the
// 32 lanes test 32 consecutive dimensions at once: lane t looks at dimension n-t, an exclusive
// prefix sum of U(k,n-s), s<t (saturating, so it can never wrap below i) gives the value `i`
// would have on reaching that dimension, and a ballot finds the first dimension that is not
// empty.  The result is identical to the serial walk because every test uses exactly the serial
// value of i.
__device__ __forceinline__ float cwrsi_warp(const PvqTable &T, int32_t *y, uint32_t n, uint32_t k, uint32_t i,
                                            uint32_t lane)
{
    uint32_t yp = 0u;
    int32_t yy = 0;
    while (n > 2u) {
        uint32_t p, k0;
        int32_t s, val;
        if (k >= n) {  // lots of pulses (pvc.rs:196-231), warp-uniform serial
            uint32_t row = T.row[n];
            p = T.data[row + k + 1u];
            s = i >= p ? -1 : 0;
            i -= (uint32_t)((int32_t)p & s);
            k0 = k;
            uint32_t q = T.data[row + n];
            if (q > i) {
                k = n;
                do {
                    k -= 1u;
                    p = T.data[T.row[k] + n];
                } while (p > i);
            } else {
                p = T.data[row + k];
                while (p > i) {
                    k -= 1u;
                    p = T.data[row + k];
                }
            }
            i -= p;
            val = ((int32_t)k0 - (int32_t)k + s) ^ s;
            if (lane == 0u) y[yp] = val;
            yp += 1u;
            yy += val * val;
            n -= 1u;
        } else {  // lots of dimensions (pvc.rs:232-258)
            // ---- parallel zero-run scan over dimensions n, n-1, ..., n-31
            uint32_t rk = T.row[k], rk1 = T.row[k + 1u];
            uint32_t nt = n - lane;                       // dimension this lane inspects
            // the run may only cover dimensions that stay in this branch: nt > 2 and nt > k
            bool live = lane < n - max(2u, k);
            uint32_t pt = live ? T.data[rk + nt] : 0u;
            uint32_t qt = live ? T.data[rk1 + nt] : 0u;
            uint32_t incl = pt;                           // inclusive saturating prefix sum
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (uint32_t)d) incl = sat_add_u32(incl, o);
            }
            uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
            if (lane == 0u) excl = 0u;
            bool zero = live && excl <= i && pt <= i - excl && (i - excl) < qt;
            uint32_t nz = __ballot_sync(0xFFFFFFFFu, !zero);
            uint32_t run = nz ? (uint32_t)(__ffs((int)nz) - 1) : 32u;  // number of leading empty dims
            if (lane < run) y[yp + lane] = 0;
            // i after the run = i - excl[run]  (excl of lane `run`, or incl of lane 31 if run == 32)
            uint32_t sub = __shfl_sync(0xFFFFFFFFu, run < 32u ? excl : incl, run < 32u ? run : 31u);
            i -= sub;
            yp += run;
            n -= run;
            // If the run stopped because the next dimension leaves this branch (n <= 2 or
            // k >= n) or because all 32 lanes were empty, re-dispatch from the loop head.
            if (run < 32u && n > 2u && k < n) {
                // ---- the dimension that holds pulses (pvc.rs:240-257), warp-uniform
                uint32_t q = T.data[rk1 + n];
                s = i >= q ? -1 : 0;
                i -= (uint32_t)((int32_t)q & s);
                k0 = k;
                do {
                    k -= 1u;
                    p = T.data[T.row[k] + n];
                } while (p > i);
                i -= p;
                val = ((int32_t)k0 - (int32_t)k + s) ^ s;
                if (lane == 0u) y[yp] = val;
                yp += 1u;
                yy += val * val;
                n -= 1u;
            }
        }
    }
    {
        // n == 2 (pvc.rs:262-275)
        uint32_t p = 2u * k + 1u;
        int32_t s = i >= p ? -1 : 0;
        i -= (uint32_t)((int32_t)p & s);
        uint32_t k0 = k;
        k = (i + 1u) >> 1;
        if (k != 0u) i -= 2u * k - 1u;
        int32_t val = ((int32_t)k0 - (int32_t)k + s) ^ s;
        if (lane == 0u) y[yp] = val;
        yp += 1u;
        yy += val * val;
        // n == 1 (pvc.rs:277-281)
        s = -(int32_t)i;
        val = ((int32_t)k + s) ^ s;
        if (lane == 0u) y[yp] = val;
        yy += val * val;
    }
    __syncwarp();
    return (float)yy;
}

// decode_pulses (pvc.rs:156-160)
__device__ __forceinline__ float decode_pulses_warp(RangeDec &d, const PvqTable &T, int32_t *y, uint32_t n, uint32_t k,
                                                    uint32_t lane)
{
    uint32_t ft = T.v(n, k);
    uint32_t i = d.uint(ft);
    return cwrsi_warp(T, y, n, k, i, lane);
}
#endif  // OPN_HOST_SHIM

}  // namespace opn
