// host_rangeenc.cpp -- host-side range ENCODER and the SYNTH-CELT/1 packet generator.
//
// RangeEncoder follows src/range_coder/encoder.rs; icwrs/encode_pulses follow
// src/celt/pvc.rs:143-180.  The encoder never runs on the GPU: north_star is decode-only and the
// encoder exists to synthesise packets for tests and benchmarks (SURVEY.md section 2, row
// "Range encoder").
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "opn_internal.h"
#include "opn_tables.h"
#include "celt2.cuh"

namespace opn {

namespace {

constexpr uint32_t kSymBits = 8, kCodeBits = 32, kUintBits = 8, kWindowSize = 32;
constexpr uint32_t kSymMax = (1u << kSymBits) - 1u;
constexpr uint32_t kCodeTop = 1u << (kCodeBits - 1u);
constexpr uint32_t kCodeBot = kCodeTop >> kSymBits;
constexpr uint32_t kCodeShift = kCodeBits - kSymBits - 1u;

inline uint32_t ilog(uint32_t x) { return x ? 32u - (uint32_t)__builtin_clz(x) : 0u; }

class RangeEncoder {
public:
    RangeEncoder(uint8_t *buf, uint32_t len) : buf_(buf), storage_(len) {}

    int error() const { return err_; }
    uint32_t rng() const { return rng_; }
    uint32_t range_bytes() const { return offs_; }
    uint32_t tell() const { return bits_total_ - ilog(rng_); }  // mod.rs:84-86
    uint32_t tell_frac() const                                   // mod.rs:96-111
    {
        static const uint32_t corr[8] = {35733, 38967, 42495, 46340, 50535, 55109, 60097, 65535};
        uint32_t l = ilog(rng_);
        uint32_t r = rng_ >> (l - 16);
        uint32_t b = (r >> 12) - 8;
        if (r > corr[b]) b++;
        return (bits_total_ << 3) - ((l << 3) + b);
    }

    void encode(uint32_t fl, uint32_t fh, uint32_t ft)  // encoder.rs:187-198
    {
        uint32_t r = rng_ / ft;
        if (fl > 0) {
            val_ += rng_ - r * (ft - fl);
            rng_ = r * (fh - fl);
        } else {
            rng_ -= r * (ft - fh);
        }
        normalize();
    }
    void encode_bin(uint32_t fl, uint32_t fh, uint32_t bits)  // encoder.rs:201-212
    {
        uint32_t r = rng_ >> bits;
        if (fl > 0) {
            val_ += rng_ - r * ((1u << bits) - fl);
            rng_ = r * (fh - fl);
        } else {
            rng_ -= r * ((1u << bits) - fh);
        }
        normalize();
    }
    void bit_logp(uint32_t v, uint32_t logp)  // encoder.rs:215-227
    {
        uint32_t s = rng_ >> logp, r = rng_ - s;
        if (v) val_ += r;
        rng_ = v ? s : r;
        normalize();
    }
    void icdf(uint32_t s, const uint8_t *tab, uint32_t ftb)  // encoder.rs:239-250
    {
        uint32_t r = rng_ >> ftb;
        if (s > 0) {
            val_ += rng_ - r * tab[s - 1];
            rng_ = r * (uint32_t)(tab[s - 1] - tab[s]);
        } else {
            rng_ -= r * tab[s];
        }
        normalize();
    }
    void bits(uint32_t v, uint32_t nbits)  // encoder.rs:282-305
    {
        uint32_t window = end_window_, used = end_bits_;
        if (used + nbits > kWindowSize) {
            do {
                put_back((uint8_t)(window & kSymMax));
                window >>= kSymBits;
                used -= kSymBits;
            } while (used >= kSymBits);
        }
        window |= v << used;
        used += nbits;
        end_window_ = window;
        end_bits_ = used;
        bits_total_ += nbits;
    }
    void uint(uint32_t v, uint32_t ft)  // encoder.rs:258-274
    {
        ft -= 1;
        uint32_t ftb = ilog(ft);
        if (ftb > kUintBits) {
            ftb -= kUintBits;
            uint32_t ft1 = (ft >> ftb) + 1;
            uint32_t hi = v >> ftb;
            encode(hi, hi + 1, ft1);
            bits(v & ((1u << ftb) - 1u), ftb);
        } else {
            encode(v, v + 1, ft + 1);
        }
    }
    // encoder.rs:437-482; returns the value actually encoded (large magnitudes are clamped)
    int32_t laplace(int32_t value, uint32_t fs, uint32_t decay)
    {
        uint32_t fl = 0;
        int32_t val = value;
        if (val != 0) {
            const int32_t s = val < 0 ? -1 : 0;
            val = (val + s) ^ s;
            fl = fs;
            fs = ((32768u - 32u - fs) * (16384u - decay)) >> 15;  // mod.rs:114-117
            int32_t i = 1;
            for (; fs > 0 && i < val; i++) {
                fs *= 2;
                fl += fs + 2;
                fs = (fs * decay) >> 15;
            }
            if (fs == 0) {
                int32_t ndi_max = ((int32_t)(32768u - fl) - s) >> 1;
                int32_t di = std::min(val - i, ndi_max - 1);
                fl += (uint32_t)(2 * di + 1 + s);
                fs = std::min(1u, 32768u - fl);
                value = (i + di + s) ^ s;
            } else {
                fs += 1;
                fl += (uint32_t)((int32_t)fs & ~s);
            }
        }
        encode_bin(fl, fl + fs, 15);
        return value;
    }
    void patch_initial_bits(uint32_t v, uint32_t nbits)  // encoder.rs:327-347
    {
        const uint32_t shift = kSymBits - nbits, mask = ((1u << nbits) - 1u) << shift;
        if (offs_ > 0) buf_[0] = (uint8_t)((buf_[0] & ~mask) | (v << shift));
        else if (rem_ >= 0) rem_ = (int32_t)(((uint32_t)rem_ & ~mask) | (v << shift));
        else if (rng_ <= (kCodeTop >> nbits)) val_ = (val_ & ~(mask << kCodeShift)) | (v << (kCodeShift + shift));
        else err_ = OPN_ERR_INTERNAL;
    }
    void done()  // encoder.rs:376-425
    {
        int32_t l = (int32_t)(kCodeBits - ilog(rng_));
        uint32_t mask = (kCodeTop - 1u) >> l;
        uint32_t end = (val_ + mask) & ~mask;
        if ((end | mask) >= val_ + rng_) {
            l += 1;
            mask >>= 1;
            end = (val_ + mask) & ~mask;
        }
        for (; l > 0; l -= (int32_t)kSymBits) {
            carry_out(end >> kCodeShift);
            end = (end << kSymBits) & (kCodeTop - 1u);
        }
        if (rem_ >= 0 || ext_ > 0) carry_out(0);
        uint32_t window = end_window_, used = end_bits_;
        for (; used >= kSymBits; used -= kSymBits) {
            put_back((uint8_t)(window & kSymMax));
            window >>= kSymBits;
        }
        if (err_) return;
        std::memset(buf_ + offs_, 0, storage_ - end_offs_ - offs_);
        if (used > 0) {
            if (end_offs_ >= storage_) {
                err_ = OPN_ERR_INTERNAL;
                return;
            }
            l = -l;
            if (offs_ + end_offs_ >= storage_ && l < (int32_t)used) window &= (1u << l) - 1u;
            buf_[storage_ - end_offs_ - 1] |= (uint8_t)window;
        }
    }

private:
    void put_front(uint8_t b)  // encoder.rs:91-99
    {
        if (offs_ + end_offs_ >= storage_) {
            err_ = OPN_ERR_BUFFER_TOO_SMALL;
            return;
        }
        buf_[offs_++] = b;
    }
    void put_back(uint8_t b)  // encoder.rs:102-109
    {
        if (offs_ + end_offs_ >= storage_) {
            err_ = OPN_ERR_BUFFER_TOO_SMALL;
            return;
        }
        buf_[storage_ - ++end_offs_] = b;
    }
    void carry_out(uint32_t c)  // encoder.rs:124-153
    {
        if (c == kSymMax) {
            ext_++;
            return;
        }
        const uint32_t carry = c >> kSymBits;
        if (rem_ >= 0) put_front((uint8_t)((uint32_t)rem_ + carry));
        for (; ext_ > 0; ext_--) put_front((uint8_t)((kSymMax + carry) & kSymMax));
        rem_ = (int32_t)(c & kSymMax);
    }
    void normalize()  // encoder.rs:157-168
    {
        while (rng_ <= kCodeBot) {
            carry_out(val_ >> kCodeShift);
            val_ = (val_ << kSymBits) & (kCodeTop - 1u);
            rng_ <<= kSymBits;
            bits_total_ += kSymBits;
        }
    }

    uint8_t *buf_;
    uint32_t storage_;
    uint32_t end_offs_ = 0, end_window_ = 0, end_bits_ = 0, bits_total_ = kCodeBits + 1, offs_ = 0;
    uint32_t rng_ = kCodeTop, val_ = 0, ext_ = 0;
    int32_t rem_ = -1;
    int err_ = 0;
};

inline uint32_t pvq_u(uint32_t n, uint32_t k)  // pvc.rs:295-298
{
    return OPN_PVQ_U_DATA[OPN_PVQ_U_ROW[std::min(n, k)] + std::max(n, k)];
}
inline uint32_t pvq_v(uint32_t n, uint32_t k) { return pvq_u(n, k) + pvq_u(n, k + 1); }

uint32_t icwrs(const int32_t *y, uint32_t n)  // pvc.rs:162-180
{
    uint32_t j = n - 1;
    uint32_t i = y[j] < 0 ? 1u : 0u;
    uint32_t k = (uint32_t)std::abs(y[j]);
    while (j-- > 0) {
        i += pvq_u(n - j, k);
        k += (uint32_t)std::abs(y[j]);
        if (y[j] < 0) i += pvq_u(n - j, k + 1);
    }
    return i;
}

// splitmix64 (public-domain constants), the PRNG SURVEY.md 8d names for the synthetic streams
struct SplitMix64 {
    uint64_t s;
    uint64_t next()
    {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

const uint8_t kTapsetIcdf[3] = {2, 1, 0};

}  // namespace

float host_gain_from_q8(int16_t gain_q8)
{
    if (gain_q8 == 0) return 1.0f;
    // fast_exp2(6.48814081e-4 * gain) with fast_exp2(x) = exp(x * LN_2), src/math.rs:17-19, decoder.rs:790-791
    const float x = 6.48814081e-4f * (float)gain_q8;
    return std::exp(x * 0.693147180559945309417232121458f);
}

}  // namespace opn

using namespace opn;

extern "C" {

static int synth_packet_attempt(uint64_t stream_id, uint64_t frame_idx, uint32_t attempt, int lm, int channels,
                                uint32_t pkt_bytes, uint32_t transient_permille, uint8_t *out, opn_synth_side &t)
{
    std::memset(&t, 0, sizeof(t));
    // TOC: CELT-only fullband (config 28..31 by frame size), stereo flag, code 0 (lib.rs:271-289,317-325)
    out[0] = (uint8_t)(0x80 | 0x60 | (lm << 3) | (channels == 2 ? 0x4 : 0) | 0x0);
    SplitMix64 rng{42ull + 1000003ull * stream_id + 0xD1B54A32D192ED03ull * frame_idx + 0x2545F4914F6CDD1Dull * attempt};
    RangeEncoder enc(out + 1, pkt_bytes - 1);

    enc.bit_logp(0, 15);  // silence
    t.postfilter = (int32_t)(rng.next() & 1);
    enc.bit_logp((uint32_t)t.postfilter, 1);
    if (t.postfilter) {
        t.octave = (int32_t)rng.below(6);
        const uint32_t fine_period = rng.below(1u << (4 + t.octave));
        t.period = (16 << t.octave) + (int32_t)fine_period - 1;
        t.gain_idx = (int32_t)rng.below(8);
        t.tapset = (int32_t)rng.below(3);
        enc.uint((uint32_t)t.octave, 6);
        enc.bits(fine_period, 4 + (uint32_t)t.octave);
        enc.bits((uint32_t)t.gain_idx, 3);
        enc.icdf((uint32_t)t.tapset, kTapsetIcdf, 2);
    }
    t.transient = rng.below(1000) < transient_permille ? 1 : 0;
    enc.bit_logp((uint32_t)t.transient, 3);
    t.intra = rng.below(8) == 0 ? 1 : 0;
    enc.bit_logp((uint32_t)t.intra, 3);
    for (int b = 0; b < 21; b++)
        for (int c = 0; c < channels; c++) {
            const uint32_t decay = 6000u + 400u * (uint32_t)b;
            const uint32_t fs0 = ((32768u - 33u) * (16384u - decay)) / (16384u + decay) + 1u;  // mod.rs:530-534
            const int32_t v = (int32_t)rng.below(16) - 7;
            t.coarse[c][b] = enc.laplace(v, fs0, decay);
        }
    for (int b = 0; b < 21; b++)
        for (int c = 0; c < channels; c++) {
            t.fine[c][b] = (int32_t)rng.below(4);
            enc.bits((uint32_t)t.fine[c][b], 2);
        }
    for (int b = 0; b < 21; b++)
        for (int c = 0; c < channels; c++) {
            const uint32_t n = OPN_SYNTH_SCHED[lm][b][0], parts = OPN_SYNTH_SCHED[lm][b][1], k = OPN_SYNTH_SCHED[lm][b][2];
            if (n == 1) {
                enc.bits((uint32_t)(rng.next() & 1), 1);
                t.n_pulses += 1;
                continue;
            }
            for (uint32_t p = 0; p < parts; p++) {
                const uint32_t v = pvq_v(n, k);
                enc.uint(rng.below(v), v);  // a uniform codeword index == encode_pulses(cwrsi(index))
                t.n_pulses += k;
            }
        }
    if (enc.error()) return enc.error();
    if (enc.tell() > 8u * (pkt_bytes - 1u)) return OPN_ERR_BUFFER_TOO_SMALL;
    t.tell_frac = enc.tell_frac();
    enc.done();
    if (enc.error()) return enc.error();
    return (int)pkt_bytes;
}

int opn_synth_packet(uint64_t stream_id, uint64_t frame_idx, int lm, int channels, uint32_t pkt_bytes,
                     uint32_t transient_permille, uint8_t *out, opn_synth_side *truth)
{
    if (!out || lm < 0 || lm > 3 || channels < 1 || channels > 2 || pkt_bytes < 3 || pkt_bytes > 1276) return OPN_ERR_BAD_ARG;
    opn_synth_side local;
    opn_synth_side &t = truth ? *truth : local;
    // The symbol values are random, so a few percent of draws run a little over a tight byte budget
    // (160 B stereo 20 ms: mean 1232 of 1272 bits).  Such a draw is rejected and the frame is redrawn
    // from the next sub-seed; the accepted packet is still a pure function of (stream, frame).
    int rc = OPN_ERR_BUFFER_TOO_SMALL;
    for (uint32_t attempt = 0; attempt < 8 && rc == OPN_ERR_BUFFER_TOO_SMALL; attempt++)
        rc = synth_packet_attempt(stream_id, frame_idx, attempt, lm, channels, pkt_bytes, transient_permille, out, t);
    return rc;
}

// ---- SYNTH-CELT/2 generator: celt2_frame (celt2.cuh) run by an ENCODING coder that draws every symbol value from
// splitmix64 and writes it with the host range encoder.  The frame logic is the very code the decode kernel runs; only the
// coder differs.  The oracle's generator (oracle/celt2.c, an independent restatement) must produce the same bytes.
namespace {
struct GenCoder {
    RangeEncoder &e;
    SplitMix64 rng;
    uint32_t tp;
    uint32_t tell() const { return e.tell(); }
    uint32_t tell_frac() const { return e.tell_frac(); }
    uint32_t transient_permille() const { return tp; }
    uint32_t bit_logp(uint32_t logp, uint32_t p1_permille)
    {
        const uint32_t v = rng.below(1000) < p1_permille ? 1u : 0u;
        e.bit_logp(v, logp);
        return v;
    }
    uint32_t icdf(const uint8_t *tab, uint32_t ftb, uint32_t n_sym)
    {
        const uint32_t v = rng.below(n_sym);
        e.icdf(v, tab, ftb);
        return v;
    }
    uint32_t uint_(uint32_t ft)
    {
        const uint32_t v = rng.below(ft);
        e.uint(v, ft);
        return v;
    }
    uint32_t pulses_index(uint32_t ft) { return uint_(ft); }  // a uniform codeword index == encode_pulses(cwrsi(index))
    uint32_t bits(uint32_t n)
    {
        const uint32_t v = rng.below(1u << n);
        e.bits(v, n);
        return v;
    }
    int32_t laplace(int band)
    {
        const uint32_t decay = 6000u + 400u * (uint32_t)band;
        const uint32_t fs0 = ((32768u - 33u) * (16384u - decay)) / (16384u + decay) + 1u;
        return e.laplace((int32_t)rng.below(16) - 7, fs0, decay);
    }
    uint32_t theta_tri(uint32_t qn)
    {
        const uint32_t h = qn >> 1, ft = (h + 1) * (h + 1);
        const uint32_t itheta = rng.below(qn + 1);
        uint32_t fl, fs;
        if (itheta <= h) {
            fs = itheta + 1;
            fl = (itheta * (itheta + 1)) >> 1;
        } else {
            fs = qn + 1 - itheta;
            fl = ft - (((qn + 1 - itheta) * (qn + 2 - itheta)) >> 1);
        }
        e.encode(fl, fl + fs, ft);
        return itheta;
    }
};
struct CountSink {
    uint32_t n = 0, np = 0;
    int e9[2][21] = {};
    void energy_set(int c, int band, int v) { e9[c][band] = v; }
    void energy_add(int c, int band, int v) { e9[c][band] += v; }
    int energy(int c, int band) const { return e9[c][band]; }
    void put_part(int, int, int k, uint32_t, float) { n++; np += (uint32_t)k; }
    uint32_t nsign = 0;
    void put_sign(int, uint32_t) { nsign++; }  // one-bin bands are not PVQ leaves, but the decoder's list holds them too
    uint32_t pulses() const { return np; }
};
const Celt2Tabs kHostCelt2Tabs{OPN_E_BANDS, OPN_LOG_N, OPN_ALLOC_VECTORS, OPN_CACHE_BITS, OPN_CACHE_CAPS, OPN_LOG2_FRAC_TABLE, OPN_CACHE_INDEX,
                               OPN_PVQ_U_DATA, OPN_PVQ_U_ROW};
}  // namespace

int opn_celt2_packet(uint64_t stream_id, uint64_t frame_idx, int lm, int channels, uint32_t pkt_bytes, uint32_t transient_permille,
                     uint8_t *out, opn_celt2_side *truth)
{
    if (!out || lm < 0 || lm > 3 || channels < 1 || channels > 2 || pkt_bytes < 8 || pkt_bytes > 1276) return OPN_ERR_BAD_ARG;
    static_assert(sizeof(opn_celt2_side) == sizeof(Celt2Side), "opusb200.h and celt2.cuh describe the same record");
    Celt2Side local;
    Celt2Side *sd = truth ? reinterpret_cast<Celt2Side *>(truth) : &local;
    std::memset(sd, 0, sizeof(*sd));
    out[0] = (uint8_t)(0x80 | 0x60 | (lm << 3) | (channels == 2 ? 0x4 : 0) | 0x0);
    RangeEncoder enc(out + 1, pkt_bytes - 1);
    GenCoder gc{enc, SplitMix64{4242ull + 1000003ull * stream_id + 0xD1B54A32D192ED03ull * frame_idx}, transient_permille};
    CountSink sink;
    uint32_t flags = 0, n_pulses = 0;
    celt2_frame(gc, kHostCelt2Tabs, pkt_bytes - 1, lm, channels, sd, sink, flags, n_pulses);
    sd->n_parts = sink.n;
    sd->n_pulses = n_pulses;
    sd->tell_frac = enc.tell_frac();
    sd->final_rng = enc.rng();
    if (sink.n + sink.nsign > (uint32_t)CELT2_MAX_PARTS) return OPN_ERR_INTERNAL;  // more leaves than the decoder's list holds
    if (enc.error()) return enc.error();
    if (enc.tell() > 8u * (pkt_bytes - 1u)) return OPN_ERR_BUFFER_TOO_SMALL;
    enc.done();
    if (enc.error()) return enc.error();
    return (int)pkt_bytes;
}

int opn_celt2_fill(uint64_t first_stream, uint32_t n_streams, uint64_t first_frame, uint32_t n_frames, int lm, int channels,
                   uint32_t pkt_bytes, uint32_t transient_permille, int n_threads, uint8_t *out)
{
    if (!out || n_streams == 0 || n_frames == 0) return OPN_ERR_BAD_ARG;
    if (n_threads < 1) n_threads = 1;
    std::vector<int> rc((size_t)n_threads, 0);
    std::vector<std::thread> pool;
    const uint64_t total = (uint64_t)n_streams * n_frames;
    for (int th = 0; th < n_threads; th++)
        pool.emplace_back([&, th]() {
            for (uint64_t w = total * th / n_threads; w < total * (th + 1) / n_threads; w++) {
                const uint64_t f = w / n_streams, s = w % n_streams;
                int r = opn_celt2_packet(first_stream + s, first_frame + f, lm, channels, pkt_bytes, transient_permille, out + w * pkt_bytes,
                                         nullptr);
                if (r < 0) rc[th] = r;
            }
        });
    for (auto &t : pool) t.join();
    for (int r : rc)
        if (r < 0) return r;
    return OPN_OK;
}

// SYNTH-SILK/1 packet generator: same seeded draws as the oracle's independent orc_silk_packet.
// One block of symbols = all coded channels of one frame (the regular frame, or the LBRR copy of the previous one).
static void silk_encode_block(RangeEncoder &enc, SplitMix64 &rng, int fs_khz, int nb_subfr, int channels)
{
    const int order = fs_khz == 16 ? 16 : 10, nblk = (nb_subfr * 5 * fs_khz + 15) / 16;
    for (int c = 0; c < channels; c++) {
        const uint32_t t8 = rng.below(8), type = t8 == 0 ? 0u : t8 < 3 ? 1u : 2u;
        enc.icdf(type, OPN_SILK_TYPE_ICDF, 8);
        enc.uint(16 + rng.below(36), 64);
        for (int f = 1; f < nb_subfr; f++) enc.icdf(3 + rng.below(3), OPN_SILK_DELTA_GAIN_ICDF, 8);
        for (int k = 0; k < order; k++) {
            const uint32_t half = k < 2 ? 16 : 8;
            const uint32_t a = rng.below(half + 1), b = rng.below(half);
            enc.bits(a + b, k < 2 ? 5 : 4);
        }
        if (type == 2) {
            enc.uint(rng.below((uint32_t)(16 * fs_khz + 1)), (uint32_t)(16 * fs_khz + 1));
            for (int f = 0; f < nb_subfr; f++) enc.icdf(rng.below(4), OPN_SILK_CONTOUR_ICDF, 8);
            for (int f = 0; f < nb_subfr; f++) enc.icdf(rng.below(8), OPN_SILK_LTP_ICDF, 8);
        }
        enc.bits(rng.below(4), 2);
        for (int b = 0; b < nblk; b++) {
            const uint32_t k1 = rng.below(9), k2 = rng.below(9), k = k1 < k2 ? k1 : k2;
            enc.icdf(k, OPN_SILK_PULSES_ICDF + (type != 0 ? 9 : 0), 8);
            if (k) {
                const uint32_t v = pvq_v(16, k);
                enc.uint(rng.below(v), v);
            }
        }
    }
}

static int silk_packet_attempt(uint64_t stream_id, uint64_t frame_idx, uint32_t attempt, int bandwidth, int frame_ms, int channels,
                               uint32_t pkt_bytes, uint32_t lbrr_permille, uint8_t *out)
{
    // TOC: SILK-only, config = 4 * bandwidth + (10 ms: 0, 20 ms: 1), stereo flag, code 0 (lib.rs:219-325)
    out[0] = (uint8_t)(((bandwidth * 4 + (frame_ms == 20 ? 1 : 0)) << 3) | (channels == 2 ? 0x4 : 0));
    SplitMix64 rng{77ull + 1000003ull * stream_id + 0xD1B54A32D192ED03ull * frame_idx + 0x2545F4914F6CDD1Dull * attempt};
    const int fs_khz = bandwidth == 0 ? 8 : bandwidth == 1 ? 12 : 16, nb_subfr = frame_ms / 5;
    RangeEncoder enc(out + 1, pkt_bytes - 1);
    const uint32_t lbrr = rng.below(1000) < lbrr_permille ? 1u : 0u;
    enc.bit_logp(lbrr, 1);
    if (lbrr) silk_encode_block(enc, rng, fs_khz, nb_subfr, channels);  // the redundant copy of the previous frame: its own draws
    silk_encode_block(enc, rng, fs_khz, nb_subfr, channels);
    if (enc.error()) return enc.error();
    if (enc.tell() > 8u * (pkt_bytes - 1u)) return OPN_ERR_BUFFER_TOO_SMALL;
    enc.done();
    return enc.error() ? enc.error() : (int)pkt_bytes;
}

int opn_silk_packet(uint64_t stream_id, uint64_t frame_idx, int bandwidth, int frame_ms, int channels, uint32_t pkt_bytes,
                    uint32_t lbrr_permille, uint8_t *out)
{
    if (!out || bandwidth < 0 || bandwidth > 2 || (frame_ms != 10 && frame_ms != 20) || channels < 1 || channels > 2 || pkt_bytes < 3 ||
        pkt_bytes > 1276)
        return OPN_ERR_BAD_ARG;
    int rc = OPN_ERR_BUFFER_TOO_SMALL;
    for (uint32_t attempt = 0; attempt < 16 && rc == OPN_ERR_BUFFER_TOO_SMALL; attempt++)
        rc = silk_packet_attempt(stream_id, frame_idx, attempt, bandwidth, frame_ms, channels, pkt_bytes, lbrr_permille, out);
    return rc;
}

int opn_silk_fill(uint64_t first_stream, uint32_t n_streams, uint64_t first_frame, uint32_t n_frames, int bandwidth, int frame_ms,
                  int channels, uint32_t pkt_bytes, uint32_t lbrr_permille, int n_threads, uint8_t *out)
{
    if (!out || n_streams == 0 || n_frames == 0) return OPN_ERR_BAD_ARG;
    if (n_threads < 1) n_threads = 1;
    std::vector<int> rc((size_t)n_threads, 0);
    std::vector<std::thread> pool;
    const uint64_t total = (uint64_t)n_streams * n_frames;
    for (int th = 0; th < n_threads; th++)
        pool.emplace_back([&, th]() {
            for (uint64_t w = total * th / n_threads; w < total * (th + 1) / n_threads; w++) {
                int r = opn_silk_packet(first_stream + w % n_streams, first_frame + w / n_streams, bandwidth, frame_ms, channels, pkt_bytes,
                                        lbrr_permille, out + w * pkt_bytes);
                if (r < 0) rc[th] = r;
            }
        });
    for (auto &t : pool) t.join();
    for (int r : rc)
        if (r < 0) return r;
    return OPN_OK;
}

int opn_synth_fill(uint64_t first_stream, uint32_t n_streams, uint64_t first_frame, uint32_t n_frames, int lm, int channels,
                   uint32_t pkt_bytes, uint32_t transient_permille, int n_threads, uint8_t *out)
{
    if (!out || n_streams == 0 || n_frames == 0) return OPN_ERR_BAD_ARG;
    if (n_threads < 1) n_threads = 1;
    std::vector<int> rc((size_t)n_threads, 0);
    std::vector<std::thread> pool;
    const uint64_t total = (uint64_t)n_streams * n_frames;
    for (int th = 0; th < n_threads; th++)
        pool.emplace_back([&, th]() {
            for (uint64_t w = total * th / n_threads; w < total * (th + 1) / n_threads; w++) {
                const uint64_t f = w / n_streams, s = w % n_streams;
                int r = opn_synth_packet(first_stream + s, first_frame + f, lm, channels, pkt_bytes, transient_permille,
                                         out + w * pkt_bytes, nullptr);
                if (r < 0) rc[th] = r;
            }
        });
    for (auto &t : pool) t.join();
    for (int r : rc)
        if (r < 0) return r;
    return OPN_OK;
}

int opn_enc_run_script(uint8_t *buf, uint32_t len, const opn_op *ops, const uint32_t *values, uint32_t n_ops,
                       const uint8_t *icdf_pool, const int32_t *y_in, uint32_t *tell_frac_out, uint32_t *range_bytes,
                       uint32_t *final_tell_frac)
{
    if (!buf || (!ops && n_ops)) return OPN_ERR_BAD_ARG;
    RangeEncoder enc(buf, len);
    uint32_t ny = 0;
    for (uint32_t i = 0; i < n_ops && !enc.error(); i++) {
        const uint32_t a = ops[i].a, b = ops[i].b, v = values ? values[i] : 0;
        switch (ops[i].op) {
        case OPN_OP_UINT: enc.uint(v, a); break;
        case OPN_OP_BITS: enc.bits(v, a); break;
        case OPN_OP_BIT_LOGP: enc.bit_logp(v, a); break;
        case OPN_OP_ICDF: enc.icdf(v, icdf_pool + a, b); break;
        case OPN_OP_LAPLACE: enc.laplace((int32_t)v, a, b); break;
        case OPN_OP_BIT_VIA_DECODE: enc.encode(v ? (1u << a) - 1u : 0u, (1u << a) - (v ? 0u : 1u), 1u << a); break;
        case OPN_OP_BIT_VIA_DECODE_BIN: enc.encode_bin(v ? (1u << a) - 1u : 0u, (1u << a) - (v ? 0u : 1u), a); break;
        case OPN_OP_PULSES:  // encode_pulses, pvc.rs:143-153
            enc.uint(icwrs(y_in + ny, a), pvq_v(a, b));
            ny += a;
            break;
        default: break;
        }
        if (tell_frac_out) tell_frac_out[i] = enc.tell_frac();
    }
    if (final_tell_frac) *final_tell_frac = enc.tell_frac();
    if (!enc.error()) enc.done();
    if (range_bytes) *range_bytes = enc.range_bytes();
    return enc.error();
}

}  // extern "C"
