/* CPU ORACLE (test infrastructure) -- SYNTH-SILK/1 frame decode, PARITY UNPINNED.
 *
 * The reference's SilkDecoder::decode is `unimplemented!()` (src/silk/decoder.rs:71-80); its caller is
 * src/decoder.rs:552-624 and the merge with the CELT output is src/decoder.rs:722-729
 * (`samples[i] += (1.0 / 32768.0) * silk_buffer[i]`).  No SILK table, codebook or filter exists anywhere in the crate, so no
 * reference frame format can be followed.  SYNTH-SILK/1 is a frame layout built only from range-coder
 * operations the reference implements (decode_icdf, decode_uint, decode_bits, decode_pulses) around the three arithmetic
 * stages north_star names for src/silk -- long-term prediction, the INTEGER short-term (LPC) synthesis recursion
 * (smulwb/sat32/sat16 fixed point, order 10 for NB/MB and 16 for WB, 5 ms subframes) and a polyphase
 * resampler to 48 kHz.  NOT interoperable with Opus.  This file defines the truth the CUDA path is compared with: every
 * integer (symbols, excitation, internal-rate samples) bit for bit, the float PCM within north_star's 1e-5.
 *
 * lbrr = bit_logp(1) opens the packet: 1 = a redundant copy of the previous frame (same syntax) precedes the regular frame.
 * Per coded channel (the channels of a stereo packet follow each other in the payload):
 *   type      = icdf(TYPE, 8)                          0 inactive, 1 unvoiced, 2 voiced
 *   gidx[0]   = uint(64); gidx[s] = clamp(gidx[s-1] + icdf(DELTA, 8) - 4, 0, 63)      gain_Q10 = GAIN_Q10[gidx]
 *   rc[k]     = bits(5) - 16 times 3600 (k < 2); bits(4) - 8 times 4800 (k < 6) or 2400      reflection coefficients, Q16
 *   voiced:     lag0 = 2 fs_khz + uint(16 fs_khz + 1); lag[s] = clamp(lag0 + icdf(CONTOUR, 8) - 1, 2 fs_khz, 18 fs_khz);
 *               ltp[s] = icdf(LTP, 8)
 *   seed      = bits(2)
 *   per 16-sample shell block: K = icdf(PULSES[type != 0], 8), K <= 8; K > 0: decode_pulses(y, 16, K)  (src/celt/pvc.rs:156)
 */
#include "oracle.h"
#include "oracle_tables.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static inline int32_t smulwb(int32_t a, int32_t b16) { return (int32_t)(((int64_t)a * (int64_t)(int16_t)b16) >> 16); }
static inline int32_t smulww(int32_t a, int32_t b) { return (int32_t)(((int64_t)a * (int64_t)b) >> 16); }
static inline int32_t sat16(int32_t x) { return x > 32767 ? 32767 : x < -32768 ? -32768 : x; }
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

void orc_silk_state_init(orc_silk_state *st) { memset(st, 0, sizeof(*st)); }

static int silk_fs_khz(int bandwidth) { return bandwidth == 0 ? 8 : bandwidth == 1 ? 12 : 16; }

/* one block of symbols: all coded channels of one frame (the regular frame of a packet, or its LBRR copy of the previous one) */
static void silk_decode_block(orc_dec *d, int fs_khz, int nb_subfr, int stream_channels, orc_silk_side *side)
{
    const int order = fs_khz == 16 ? 16 : 10, L = nb_subfr * 5 * fs_khz, nblk = (L + 15) / 16;
    const int min_lag = 2 * fs_khz, max_lag = 18 * fs_khz;
    for (int c = 0; c < stream_channels; c++) {
        orc_silk_chan_side *s = &side->ch[c];
        s->type = (int32_t)orc_dec_icdf(d, ORC_SILK_TYPE_ICDF, 8);
        s->gidx[0] = (int32_t)orc_dec_uint(d, 64);
        for (int f = 1; f < nb_subfr; f++) {
            int g = s->gidx[f - 1] + (int32_t)orc_dec_icdf(d, ORC_SILK_DELTA_GAIN_ICDF, 8) - 4;
            s->gidx[f] = g < 0 ? 0 : g > 63 ? 63 : g;
        }
        for (int k = 0; k < order; k++) s->rc_idx[k] = (int32_t)orc_dec_bits(d, k < 2 ? 5 : 4);
        if (s->type == 2) {
            int lag0 = min_lag + (int32_t)orc_dec_uint(d, (uint32_t)(max_lag - min_lag + 1));
            for (int f = 0; f < nb_subfr; f++) {
                int l = lag0 + (int32_t)orc_dec_icdf(d, ORC_SILK_CONTOUR_ICDF, 8) - 1;
                s->lag[f] = l < min_lag ? min_lag : l > max_lag ? max_lag : l;
            }
            for (int f = 0; f < nb_subfr; f++) s->ltp_idx[f] = (int32_t)orc_dec_icdf(d, ORC_SILK_LTP_ICDF, 8);
        }
        s->seed = (int32_t)orc_dec_bits(d, 2);
        for (int b = 0; b < nblk; b++) {
            uint32_t k = orc_dec_icdf(d, ORC_SILK_PULSES_ICDF + (s->type != 0 ? 9 : 0), 8);
            s->pulses[b] = (int32_t)k;
            s->index[b] = k ? orc_dec_uint(d, orc_pvq_v(16, k)) : 0u;
        }
    }
}

/* A packet starts with lbrr = bit_logp(1); when set, a low-bit-rate redundant copy of the PREVIOUS frame (same syntax) comes before
 * the regular frame -- the in-band FEC that Decoder::decode(.., decode_fec = true) asks for (src/decoder.rs:343-386, LostFlag::DecodeFec
 * in src/silk/decoder.rs:6-14).  fec: decode the redundant block instead of the regular frame; returns 0 when the packet has none. */
static int silk_decode_symbols(orc_dec *d, int fs_khz, int nb_subfr, int stream_channels, int fec, orc_silk_side *side)
{
    const int lbrr = orc_dec_bit_logp(d, 1);
    int have = 1;
    if (fec) {
        if (lbrr) silk_decode_block(d, fs_khz, nb_subfr, stream_channels, side);
        else have = 0;
    } else {
        if (lbrr) {
            orc_silk_side skip;
            silk_decode_block(d, fs_khz, nb_subfr, stream_channels, &skip);
        }
        silk_decode_block(d, fs_khz, nb_subfr, stream_channels, side);
    }
    side->lbrr = lbrr;
    side->final_rng = d->rng;
    side->tell_frac = orc_dec_tell_frac(d);
    return have;
}

/* reflection coefficients -> prediction coefficients: the step-up recursion A_k(z) = A_{k-1}(z) + rc_k z^-k A_{k-1}(1/z) on
 * Q24 polynomial coefficients with Q16 reflection coefficients (the shape of libopus' silk_k2a_Q16), then a_Q12 = -c */
static void silk_k2a(const int32_t *rc_idx, int order, int16_t *a_q12)
{
    int32_t c[16], t[16];
    memset(c, 0, sizeof(c));
    for (int k = 0; k < order; k++) {
        const int32_t rc = k < 2 ? (rc_idx[k] - 16) * 3600 : (rc_idx[k] - 8) * (k < 6 ? 4800 : 2400);
        for (int n = 0; n < k; n++) t[n] = wadd(c[n], smulww(c[k - 1 - n], rc));
        for (int n = 0; n < k; n++) c[n] = t[n];
        c[k] = rc * 256;
    }
    for (int k = 0; k < 16; k++) {
        int32_t x = k < order ? (int32_t)(0u - (uint32_t)c[k]) : 0;
        a_q12[k] = (int16_t)sat16(((x >> 11) + 1) >> 1);
    }
}

/* one coded channel: excitation -> long-term prediction -> LPC synthesis -> internal-rate samples out[0..L) */
static void silk_channel(orc_silk_chan *st, const orc_silk_chan_side *s, int fs_khz, int nb_subfr, int lost, int32_t *exc_out, int32_t *out)
{
    const int order = fs_khz == 16 ? 16 : 10, sub = 5 * fs_khz, L = nb_subfr * sub, nblk = (L + 15) / 16;
    int32_t pres[ORC_SILK_MAX_FRAME + 16];
    int32_t gain[4];
    int16_t a_q12[16];
    if (lost) { /* concealment stand-in: no excitation, the previous frame's filter and gain ring out */
        memset(pres, 0, sizeof(pres));
        memcpy(a_q12, st->a_q12, sizeof(a_q12));
        for (int f = 0; f < 4; f++) gain[f] = st->gain_q10;
    } else {
        static const int32_t offset_q14[3] = {32 << 4, 100 << 4, 32 << 4};
        for (int b = 0; b < nblk; b++) {
            int32_t y[16];
            memset(y, 0, sizeof(y));
            if (s->pulses[b]) orc_cwrsi(y, 16, (uint32_t)s->pulses[b], s->index[b]);
            uint32_t r = (uint32_t)(s->seed + 1) * 2654435761u + (uint32_t)b * 2246822519u;
            for (int j = 0; j < 16; j++) {
                int32_t e = y[j] * 16384;
                if (y[j] > 0) e -= 80 << 4;
                else if (y[j] < 0) e += 80 << 4;
                e += offset_q14[s->type];
                r = r * 196314165u + 907633515u;
                if (r & 0x80000000u) e = -e;
                r += (uint32_t)y[j];
                pres[16 * b + j] = e;
            }
        }
        if (s->type == 2) {
            for (int f = 0; f < nb_subfr; f++) {
                const int16_t *B = ORC_SILK_LTP_Q14 + 5 * s->ltp_idx[f];
                for (int i = f * sub; i < (f + 1) * sub; i++) {
                    int32_t pred = 2;
                    for (int k = 0; k < 5; k++) {
                        int idx = i - s->lag[f] + 2 - k;
                        pred = wadd(pred, smulwb(idx >= 0 ? pres[idx] : st->hist[ORC_SILK_HIST + idx], B[k]));
                    }
                    pres[i] = wadd(pres[i], (int32_t)((uint32_t)pred << 2));
                }
            }
        }
        silk_k2a(s->rc_idx, order, a_q12);
        for (int f = 0; f < nb_subfr; f++) gain[f] = ORC_SILK_GAIN_Q10[s->gidx[f]];
    }
    if (exc_out) memcpy(exc_out, pres, sizeof(int32_t) * (size_t)L);
    /* LPC synthesis, SURVEY.md appendix B: pred_Q10 = order/2 + sum smulwb(sLPC_Q14[i-1-k], A_Q12[k]);
     * sLPC_Q14[i] = sat32(res_Q14[i] + (pred_Q10 << 4)); out = sat16(rshift_round(smulww(sLPC_Q14[i], gain_Q10), 8)) */
    int32_t slpc[ORC_SILK_MAX_FRAME + 16];
    memcpy(slpc, st->slpc, sizeof(st->slpc)); /* slpc[15] = the newest sample of the previous frame */
    for (int i = 0; i < L; i++) {
        int32_t pred = order / 2;
        for (int k = 0; k < order; k++) pred = wadd(pred, smulwb(slpc[16 + i - 1 - k], a_q12[k]));
        int64_t v = (int64_t)pres[i] + (int64_t)pred * 16;
        int32_t v32 = v > 2147483647ll ? 2147483647 : v < -2147483648ll ? (int32_t)(-2147483647 - 1) : (int32_t)v;
        slpc[16 + i] = v32;
        int32_t w = smulww(v32, gain[i / sub]);
        out[i] = sat16(((w >> 7) + 1) >> 1);
    }
    memcpy(st->slpc, &slpc[(size_t)(unsigned)L], sizeof(st->slpc));
    /* excitation history of the long-term predictor: the last ORC_SILK_HIST samples */
    if (L >= ORC_SILK_HIST) memcpy(st->hist, pres + L - ORC_SILK_HIST, sizeof(st->hist));
    else {
        memmove(st->hist, st->hist + L, sizeof(int32_t) * (size_t)(ORC_SILK_HIST - L));
        memcpy(st->hist + ORC_SILK_HIST - L, pres, sizeof(int32_t) * (size_t)L);
    }
    memcpy(st->a_q12, a_q12, sizeof(a_q12));
    st->gain_q10 = gain[nb_subfr - 1];
}

static const float *silk_up_table(int up) { return up == 6 ? ORC_SILK_UP6 : up == 4 ? ORC_SILK_UP4 : ORC_SILK_UP3; }

int orc_silk_decode_frame(orc_silk_state *st, const uint8_t *payload, uint32_t len, int bandwidth, int frame_ms, int stream_channels,
                          int channels, int lost, orc_silk_side *side, int32_t *exc_out, int16_t *out16, float *pcm_out)
{
    if (!st || !pcm_out || channels < 1 || channels > 2 || stream_channels < 1 || stream_channels > 2) return ORC_ERR_BAD_ARG;
    if (frame_ms != 10 && frame_ms != 20) return ORC_ERR_BAD_ARG;
    const int fec = lost == 2; /* LostFlag::DecodeFec: the packet's redundant copy of the previous frame */
    if (fec) lost = 0;
    if (!lost && (bandwidth < 0 || bandwidth > 2)) return ORC_ERR_INVALID_PACKET; /* decoder.rs:566-585: SILK stops at wideband */
    if (!lost && (!payload || len <= 1)) lost = 1;                                /* decoder.rs:467 */
    const int n48 = frame_ms * 48;
    orc_silk_side local;
    if (!side) side = &local;
    memset(side, 0, sizeof(*side));
    if (!lost) { /* the symbols need nothing of the decoder's state */
        orc_dec d;
        orc_dec_init(&d, payload, len);
        if (!silk_decode_symbols(&d, silk_fs_khz(bandwidth), frame_ms / 5, stream_channels, fec, side)) lost = 1; /* FEC asked of a
                                                                                 packet without a redundant copy: conceal */
    }
    if (lost && st->fs_khz == 0) { /* nothing decoded yet: silence, state untouched */
        memset(pcm_out, 0, sizeof(float) * (size_t)n48 * (size_t)channels);
        return n48;
    }
    const int fs_khz = lost ? st->fs_khz : silk_fs_khz(bandwidth);
    const int nb_subfr = frame_ms / 5, L = nb_subfr * 5 * fs_khz, up = 48 / fs_khz;
    if (lost) stream_channels = st->stream_channels;
    if (fs_khz != st->fs_khz) { /* first frame or a change of the internal rate: every filter starts from rest */
        memset(st, 0, sizeof(*st));
        st->fs_khz = fs_khz;
    }
    st->stream_channels = stream_channels;
    int32_t out[2][ORC_SILK_MAX_FRAME];
    for (int c = 0; c < stream_channels; c++)
        silk_channel(&st->ch[c], &side->ch[c], fs_khz, nb_subfr, lost, exc_out ? exc_out + c * ORC_SILK_MAX_FRAME : NULL, out[c]);
    if (stream_channels == 2 && channels == 2) { /* mid/side -> left/right */
        for (int i = 0; i < L; i++) {
            int32_t m = out[0][i], s = out[1][i];
            out[0][i] = sat16(m + s);
            out[1][i] = sat16(m - s);
        }
    } else if (stream_channels == 1 && channels == 2) {
        memcpy(out[1], out[0], sizeof(int32_t) * (size_t)L);
    } /* stereo packet, mono decoder: the mid channel */
    if (out16)
        for (int c = 0; c < channels; c++)
            for (int i = 0; i < L; i++) out16[c * ORC_SILK_MAX_FRAME + i] = (int16_t)out[c][i];
    /* polyphase interpolation by up = 48 / fs_khz: y[up i + p] = sum_j h[p][j] x[i - j], summed in tap order; then the merge of
     * decoder.rs:722-729 onto a zero CELT contribution */
    const float *h = silk_up_table(up);
    for (int c = 0; c < channels; c++) {
        float x[ORC_SILK_MAX_FRAME + 8];
        for (int j = 0; j < 7; j++) x[6 - j] = st->rs[c][j]; /* rs[c][0] = x[-1] */
        for (int i = 0; i < L; i++) x[7 + i] = (float)out[c][i];
        for (int i = 0; i < L; i++)
            for (int p = 0; p < up; p++) {
                float acc = h[8 * p] * x[7 + i];
                for (int j = 1; j < 8; j++) acc = acc + h[8 * p + j] * x[7 + i - j];
                pcm_out[(size_t)(up * i + p) * (size_t)channels + (size_t)c] = 0.0f + (1.0f / 32768.0f) * acc;
            }
        for (int j = 0; j < 7; j++) st->rs[c][j] = x[7 + L - 1 - j];
    }
    return n48;
}

/* ------------------------------------------------------------------ packet generator (oracle's own range encoder) */
typedef struct { uint64_t s; } smix;
static uint64_t sm_next(smix *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t sm_below(smix *r, uint32_t n) { return (uint32_t)(((sm_next(r) >> 32) * (uint64_t)n) >> 32); }

static void silk_encode_block(orc_enc *pe, smix *prng, int fs_khz, int nb_subfr, int channels);

static int silk_packet_attempt(uint64_t stream_id, uint64_t frame_idx, uint32_t attempt, int bandwidth, int frame_ms, int channels,
                               uint32_t pkt_bytes, uint32_t lbrr_permille, uint8_t *out)
{
    /* TOC: SILK-only, config = 4 bandwidth + (10 ms: 0, 20 ms: 1), stereo flag, code 0 (src/lib.rs:219-325) */
    out[0] = (uint8_t)(((bandwidth * 4 + (frame_ms == 20 ? 1 : 0)) << 3) | (channels == 2 ? 0x4 : 0));
    smix rng = {77ull + 1000003ull * stream_id + 0xD1B54A32D192ED03ull * frame_idx + 0x2545F4914F6CDD1Dull * attempt};
    const int fs_khz = silk_fs_khz(bandwidth), nb_subfr = frame_ms / 5;
    orc_enc e;
    orc_enc_init(&e, out + 1, pkt_bytes - 1);
    const uint32_t lbrr = sm_below(&rng, 1000) < lbrr_permille ? 1u : 0u;
    orc_enc_bit_logp(&e, lbrr, 1);
    if (lbrr) silk_encode_block(&e, &rng, fs_khz, nb_subfr, channels); /* the redundant copy of the previous frame: its own draws */
    silk_encode_block(&e, &rng, fs_khz, nb_subfr, channels);
    if (e.error) return e.error;
    if (orc_enc_tell(&e) > 8u * (pkt_bytes - 1u)) return ORC_ERR_BUFFER_TOO_SMALL;
    orc_enc_done(&e);
    return e.error ? e.error : (int)pkt_bytes;
}

static void silk_encode_block(orc_enc *pe, smix *prng, int fs_khz, int nb_subfr, int channels)
{
    const int order = fs_khz == 16 ? 16 : 10, L = nb_subfr * 5 * fs_khz, nblk = (L + 15) / 16;
    for (int c = 0; c < channels; c++) {
        const uint32_t t8 = sm_below(prng, 8), type = t8 == 0 ? 0u : t8 < 3 ? 1u : 2u;
        orc_enc_icdf(pe, type, ORC_SILK_TYPE_ICDF, 8);
        orc_enc_uint(pe, 16 + sm_below(prng, 36), 64);
        for (int f = 1; f < nb_subfr; f++) orc_enc_icdf(pe, 3 + sm_below(prng, 3), ORC_SILK_DELTA_GAIN_ICDF, 8);
        for (int k = 0; k < order; k++) {
            const uint32_t half = k < 2 ? 16 : 8;
            const uint32_t a = sm_below(prng, half + 1), b = sm_below(prng, half); /* triangular around the middle */
            orc_enc_bits(pe, a + b, k < 2 ? 5 : 4);
        }
        if (type == 2) {
            orc_enc_uint(pe, sm_below(prng, (uint32_t)(16 * fs_khz + 1)), (uint32_t)(16 * fs_khz + 1));
            for (int f = 0; f < nb_subfr; f++) orc_enc_icdf(pe, sm_below(prng, 4), ORC_SILK_CONTOUR_ICDF, 8);
            for (int f = 0; f < nb_subfr; f++) orc_enc_icdf(pe, sm_below(prng, 8), ORC_SILK_LTP_ICDF, 8);
        }
        orc_enc_bits(pe, sm_below(prng, 4), 2);
        for (int b = 0; b < nblk; b++) {
            uint32_t k1 = sm_below(prng, 9), k2 = sm_below(prng, 9), k = k1 < k2 ? k1 : k2;
            orc_enc_icdf(pe, k, ORC_SILK_PULSES_ICDF + (type != 0 ? 9 : 0), 8);
            if (k) {
                uint32_t v = orc_pvq_v(16, k);
                orc_enc_uint(pe, sm_below(prng, v), v); /* a uniform codeword index == encode_pulses(cwrsi(index)) */
            }
        }
    }
}

int orc_silk_packet(uint64_t stream_id, uint64_t frame_idx, int bandwidth, int frame_ms, int channels, uint32_t pkt_bytes,
                    uint32_t lbrr_permille, uint8_t *out)
{
    if (!out || bandwidth < 0 || bandwidth > 2 || (frame_ms != 10 && frame_ms != 20) || channels < 1 || channels > 2 || pkt_bytes < 3 ||
        pkt_bytes > 1276)
        return ORC_ERR_BAD_ARG;
    int rc = ORC_ERR_BUFFER_TOO_SMALL; /* a draw that overruns the byte budget is redrawn from the next sub-seed */
    for (uint32_t attempt = 0; attempt < 16 && rc == ORC_ERR_BUFFER_TOO_SMALL; attempt++)
        rc = silk_packet_attempt(stream_id, frame_idx, attempt, bandwidth, frame_ms, channels, pkt_bytes, lbrr_permille, out);
    return rc;
}

typedef struct {
    uint64_t first_stream, first_frame, w0, w1;
    uint32_t n_streams, pkt_bytes, lbrr_permille;
    int bandwidth, frame_ms, channels, rc;
    uint8_t *out;
} sfill_job;

static void *sfill_thread(void *arg)
{
    sfill_job *j = (sfill_job *)arg;
    for (uint64_t w = j->w0; w < j->w1; w++) {
        int r = orc_silk_packet(j->first_stream + w % j->n_streams, j->first_frame + w / j->n_streams, j->bandwidth, j->frame_ms,
                                j->channels, j->pkt_bytes, j->lbrr_permille, j->out + w * j->pkt_bytes);
        if (r < 0) j->rc = r;
    }
    return NULL;
}

/* layout [frame][stream][pkt_bytes] */
int orc_silk_fill(uint64_t first_stream, uint32_t n_streams, uint64_t first_frame, uint32_t n_frames, int bandwidth, int frame_ms,
                  int channels, uint32_t pkt_bytes, uint32_t lbrr_permille, int n_threads, uint8_t *out)
{
    if (!out || n_streams == 0 || n_frames == 0) return ORC_ERR_BAD_ARG;
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    sfill_job *jobs = (sfill_job *)calloc((size_t)n_threads, sizeof(sfill_job));
    const uint64_t total = (uint64_t)n_streams * n_frames;
    for (int t = 0; t < n_threads; t++) {
        sfill_job *j = &jobs[t];
        j->first_stream = first_stream; j->first_frame = first_frame;
        j->w0 = total * (uint64_t)t / (uint64_t)n_threads; j->w1 = total * (uint64_t)(t + 1) / (uint64_t)n_threads;
        j->n_streams = n_streams; j->pkt_bytes = pkt_bytes; j->lbrr_permille = lbrr_permille;
        j->bandwidth = bandwidth; j->frame_ms = frame_ms; j->channels = channels; j->out = out;
        pthread_create(&th[t], NULL, sfill_thread, j);
    }
    int rc = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc < 0) rc = jobs[t].rc;
    }
    free(th);
    free(jobs);
    return rc;
}

/* ------------------------------------------------------------------ CPU baseline */
typedef struct {
    const uint8_t *packets;
    uint32_t n_streams, n_frames, pkt_bytes, s0, s1;
    int bandwidth, frame_ms, channels;
    float *pcm_last;
    uint32_t rng_xor;
} sbench_job;

static void *sbench_thread(void *arg)
{
    sbench_job *j = (sbench_job *)arg;
    const int n48 = j->frame_ms * 48;
    orc_silk_state st;
    orc_silk_side side;
    float pcm[2 * 960];
    uint32_t x = 0;
    for (uint32_t s = j->s0; s < j->s1; s++) {
        orc_silk_state_init(&st);
        for (uint32_t f = 0; f < j->n_frames; f++) {
            const uint8_t *pkt = j->packets + ((size_t)f * j->n_streams + s) * j->pkt_bytes;
            orc_silk_decode_frame(&st, pkt + 1, j->pkt_bytes - 1, j->bandwidth, j->frame_ms, j->channels, j->channels, 0, &side, NULL, NULL, pcm);
            x ^= side.final_rng;
        }
        if (j->pcm_last) memcpy(j->pcm_last + (size_t)s * (size_t)n48 * (size_t)j->channels, pcm, sizeof(float) * (size_t)n48 * (size_t)j->channels);
    }
    j->rng_xor = x;
    return NULL;
}

double orc_silk_bench(const uint8_t *packets, uint32_t n_streams, uint32_t n_frames, uint32_t pkt_bytes, int bandwidth, int frame_ms,
                      int channels, int n_threads, float *pcm_last, uint32_t *final_rng_xor)
{
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > n_streams) n_threads = (int)n_streams;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    sbench_job *jobs = (sbench_job *)calloc((size_t)n_threads, sizeof(sbench_job));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) {
        sbench_job *j = &jobs[t];
        j->packets = packets; j->n_streams = n_streams; j->n_frames = n_frames; j->pkt_bytes = pkt_bytes;
        j->s0 = (uint32_t)((uint64_t)n_streams * (uint64_t)t / (uint64_t)n_threads);
        j->s1 = (uint32_t)((uint64_t)n_streams * (uint64_t)(t + 1) / (uint64_t)n_threads);
        j->bandwidth = bandwidth; j->frame_ms = frame_ms; j->channels = channels; j->pcm_last = pcm_last;
        pthread_create(&th[t], NULL, sbench_thread, j);
    }
    uint32_t x = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        x ^= jobs[t].rng_xor;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (final_rng_xor) *final_rng_xor = x;
    free(th);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
