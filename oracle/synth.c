/* CPU ORACLE (test infrastructure) -- SYNTH-CELT/1 frame decode and the CPU baseline loop.
 *
 * The reference's CeltDecoder::decode is a stub (src/celt/decoder.rs:47-56 is `todo!()`), so
 * there is no reference frame format to follow.  SYNTH-CELT/1
 * is a frame layout built only from operations the reference implements, chained the way
 * src/decoder.rs:700-711 would chain them:
 *   RangeDecoder::new -> flags / post-filter parameters / Laplace energies / raw fine bits ->
 *   decode_pulses per band part -> (synthetic unit-norm "denormalise") -> Mdct::backward with the
 *   60-sample carry -> comb_filter_inplace on the rolling history -> interleaved f32 PCM.
 * This file is the straight-line CPU composition of the oracle primitives; the CUDA path must
 * reproduce it bit-exactly for every integer and within 1e-5 max-abs for PCM. */
#include "oracle.h"
#include "oracle_tables.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define HIST ORC_SYNTH_HIST
static const uint8_t TAPSET_ICDF[3] = {2, 1, 0};

void orc_synth_state_init(orc_synth_state *s)
{
    memset(s, 0, sizeof(*s));
    s->pos = HIST;
}

static void decode_symbols(orc_dec *d, int lm, int channels, orc_synth_side *side, int32_t *y_out,
                           float *coef)
{
    int nf = 120 << lm;
    memset(side, 0, sizeof(*side));
    memset(coef, 0, sizeof(float) * (size_t)nf * (size_t)channels);
    if (y_out) memset(y_out, 0, sizeof(int32_t) * (size_t)nf * (size_t)channels);

    side->silence = orc_dec_bit_logp(d, 15);
    if (!side->silence) {
        side->postfilter = orc_dec_bit_logp(d, 1);
        if (side->postfilter) {
            side->octave = (int32_t)orc_dec_uint(d, 6);
            side->period = (16 << side->octave) + (int32_t)orc_dec_bits(d, 4 + (uint32_t)side->octave) - 1;
            side->gain_idx = (int32_t)orc_dec_bits(d, 3);
            side->tapset = (int32_t)orc_dec_icdf(d, TAPSET_ICDF, 2);
        }
        side->transient = orc_dec_bit_logp(d, 3);
        side->intra = orc_dec_bit_logp(d, 3);
        for (int b = 0; b < 21; b++)
            for (int c = 0; c < channels; c++) {
                uint32_t decay = 6000u + 400u * (uint32_t)b;
                side->coarse[c][b] = orc_dec_laplace(d, orc_laplace_start_freq(decay), decay);
            }
        for (int b = 0; b < 21; b++)
            for (int c = 0; c < channels; c++) side->fine[c][b] = (int32_t)orc_dec_bits(d, 2);
        int32_t y[176];
        for (int b = 0; b < 21; b++)
            for (int c = 0; c < channels; c++) {
                int n = ORC_SYNTH_SCHED[lm][b][0], parts = ORC_SYNTH_SCHED[lm][b][1],
                    k = ORC_SYNTH_SCHED[lm][b][2];
                int base = c * nf + ((int)ORC_E_BANDS[b] << lm);
                if (n == 1) {
                    uint32_t sign = orc_dec_bits(d, 1);
                    coef[base] = sign ? -0.03125f : 0.03125f;
                    if (y_out) y_out[base] = sign ? -1 : 1;
                    side->n_pulses += 1;
                    continue;
                }
                for (int p = 0; p < parts; p++) {
                    float yy = orc_decode_pulses(d, y, (uint32_t)n, (uint32_t)k);
                    float g = 0.03125f / sqrtf(yy);
                    for (int j = 0; j < n; j++) {
                        coef[base + p * n + j] = (float)y[j] * g;
                        if (y_out) y_out[base + p * n + j] = y[j];
                    }
                    side->n_pulses += (uint32_t)k;
                }
            }
    }
    side->final_rng = d->rng;
    side->tell_frac = orc_dec_tell_frac(d);
}

/* Coefficients -> PCM, shared by SYNTH-CELT/1 and /2: Mdct::backward per channel onto the 60-sample carry (one long block
 * or 2^LM interleaved short blocks), comb_filter_inplace from the previous frame's post-filter parameters to this
 * frame's over the first 120 samples, interleaved output.  A lost frame keeps the previous parameters. */
int orc_synth_finish_frame(orc_synth_state *st, const float *coef, int lm, int channels, int apply_comb, int lost, int postfilter,
                           int period, int gain_idx, int tapset, int transient, float *pcm_out)
{
    int nf = 120 << lm;
    int t1 = st->pf_period, tap1 = st->pf_tapset;
    float g1 = st->pf_gain;
    if (!lost) {
        t1 = postfilter ? period : 0;
        g1 = postfilter ? 0.09375f * (float)(gain_idx + 1) : 0.0f;
        tap1 = postfilter ? tapset : 0;
    }

    int blocks = transient ? (1 << lm) : 1;
    int shift = transient ? 3 : 3 - lm;
    if (st->pos + (uint32_t)nf + 60 > ORC_SYNTH_BUF) { /* buffer full: history + tail to the front */
        for (int c = 0; c < channels; c++)
            memmove(st->buf[c], st->buf[c] + st->pos - HIST, sizeof(float) * (HIST + 60));
        st->pos = HIST;
    }
    for (int c = 0; c < channels; c++) {
        float *work = st->buf[c] + st->pos - HIST; /* work[HIST .. HIST+60) already holds the previous tail */
        memset(work + HIST + 60, 0, sizeof(float) * (size_t)nf);
        for (int b = 0; b < blocks; b++)
            orc_mdct_backward(coef + c * nf + b, work + HIST + 120 * b * (blocks > 1), ORC_WINDOW,
                              ORC_OVERLAP, shift, blocks);
        if (apply_comb)
            orc_comb_filter_inplace(work, HIST, (size_t)st->pf_period, (size_t)t1, (size_t)nf,
                                    st->pf_gain, g1, (size_t)st->pf_tapset, (size_t)tap1, ORC_OVERLAP);
        for (int i = 0; i < nf; i++) pcm_out[i * channels + c] = work[HIST + i];
    }
    st->pos += (uint32_t)nf;
    st->pf_period = t1;
    st->pf_gain = g1;
    st->pf_tapset = tap1;
    return nf;
}


int orc_synth_decode_frame(orc_synth_state *st, const uint8_t *payload, uint32_t len, int lm,
                           int channels, int apply_comb, orc_synth_side *side, int32_t *y_out,
                           float *coef_out, float *pcm_out)
{
    if (lm < 0 || lm > 3 || channels < 1 || channels > 2) return ORC_ERR_BAD_ARG;
    int nf = 120 << lm;
    float coef_local[2 * 960];
    float *coef = coef_out ? coef_out : coef_local;
    orc_synth_side side_local;
    if (!side) side = &side_local;

    int lost = len <= 1; /* src/decoder.rs:467: payloads of 0 or 1 byte trigger PLC/DTX */
    if (lost) {
        memset(side, 0, sizeof(*side));
        memset(coef, 0, sizeof(float) * (size_t)nf * (size_t)channels);
        if (y_out) memset(y_out, 0, sizeof(int32_t) * (size_t)nf * (size_t)channels);
    } else {
        orc_dec d;
        orc_dec_init(&d, payload, len);
        decode_symbols(&d, lm, channels, side, y_out, coef);
    }

    return orc_synth_finish_frame(st, coef, lm, channels, apply_comb, lost, side->postfilter, side->period, side->gain_idx, side->tapset,
                                  side->transient, pcm_out);
}

/* A packet of `stream_channels` channels in a decoder of `channels` channels (stream_channels, src/decoder.rs:332,376,395; the
 * mapping itself belongs to the stubbed CeltDecoder and is restated from libopus' celt_synthesis): mono -> stereo transforms
 * the same spectrum into both channels (each keeps its own overlap and post-filter history), stereo -> mono transforms the
 * average 0.5 * (l + r).  coef: [max(stream_channels, channels)][nf], channel-major, mapped in place. */
void orc_map_channels(float *coef, int nf, int stream_channels, int channels)
{
    if (stream_channels == 1 && channels == 2) {
        memcpy(coef + nf, coef, sizeof(float) * (size_t)nf);
    } else if (stream_channels == 2 && channels == 1) {
        for (int i = 0; i < nf; i++) coef[i] = 0.5f * (coef[i] + coef[nf + i]);
    }
}

int orc_synth_decode_frame_mapped(orc_synth_state *st, const uint8_t *payload, uint32_t len, int lm, int stream_channels, int channels,
                                  int apply_comb, orc_synth_side *side, float *pcm_out)
{
    if (lm < 0 || lm > 3 || channels < 1 || channels > 2 || stream_channels < 1 || stream_channels > 2) return ORC_ERR_BAD_ARG;
    int nf = 120 << lm;
    float coef[2 * 960];
    orc_synth_side side_local;
    if (!side) side = &side_local;
    int lost = len <= 1;
    memset(coef, 0, sizeof(coef));
    if (lost) {
        memset(side, 0, sizeof(*side));
    } else {
        orc_dec d;
        orc_dec_init(&d, payload, len);
        decode_symbols(&d, lm, stream_channels, side, NULL, coef);
        orc_map_channels(coef, nf, stream_channels, channels);
    }
    return orc_synth_finish_frame(st, coef, lm, channels, apply_comb, lost, side->postfilter, side->period, side->gain_idx, side->tapset,
                                  side->transient, pcm_out);
}

/* ------------------------------------------------------------------ packet generator
 * SYNTH-CELT/1 packets for the reference arm of bench.py (so that it never loads the product library) and for
 * cross-checking the product's generator: the same splitmix64 stream and draw order as opn_synth_packet
 * (seed 42 + 1000003*stream, per frame), written with the oracle's own range encoder. */
typedef struct { uint64_t s; } splitmix64;
static uint64_t sm_next(splitmix64 *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static uint32_t sm_below(splitmix64 *r, uint32_t n) { return (uint32_t)(((sm_next(r) >> 32) * (uint64_t)n) >> 32); }

static int synth_packet_attempt(uint64_t stream_id, uint64_t frame_idx, uint32_t attempt, int lm, int channels,
                                uint32_t pkt_bytes, uint32_t transient_permille, uint8_t *out)
{
    /* TOC: CELT-only fullband, frame size by LM, stereo flag, code 0 (src/lib.rs:271-289, 317-325) */
    out[0] = (uint8_t)(0x80 | 0x60 | (lm << 3) | (channels == 2 ? 0x4 : 0));
    splitmix64 rng = {42ull + 1000003ull * stream_id + 0xD1B54A32D192ED03ull * frame_idx + 0x2545F4914F6CDD1Dull * attempt};
    orc_enc e;
    orc_enc_init(&e, out + 1, pkt_bytes - 1);
    orc_enc_bit_logp(&e, 0, 15);
    uint32_t postfilter = (uint32_t)(sm_next(&rng) & 1);
    orc_enc_bit_logp(&e, postfilter, 1);
    if (postfilter) {
        uint32_t octave = sm_below(&rng, 6);
        uint32_t fine_period = sm_below(&rng, 1u << (4 + octave));
        uint32_t gain_idx = sm_below(&rng, 8);
        uint32_t tapset = sm_below(&rng, 3);
        orc_enc_uint(&e, octave, 6);
        orc_enc_bits(&e, fine_period, 4 + octave);
        orc_enc_bits(&e, gain_idx, 3);
        orc_enc_icdf(&e, tapset, TAPSET_ICDF, 2);
    }
    orc_enc_bit_logp(&e, sm_below(&rng, 1000) < transient_permille ? 1u : 0u, 3);
    orc_enc_bit_logp(&e, sm_below(&rng, 8) == 0 ? 1u : 0u, 3);
    for (int b = 0; b < 21; b++)
        for (int c = 0; c < channels; c++) {
            uint32_t decay = 6000u + 400u * (uint32_t)b;
            int32_t v = (int32_t)sm_below(&rng, 16) - 7;
            orc_enc_laplace(&e, &v, orc_laplace_start_freq(decay), decay);
        }
    for (int b = 0; b < 21; b++)
        for (int c = 0; c < channels; c++) orc_enc_bits(&e, sm_below(&rng, 4), 2);
    for (int b = 0; b < 21; b++)
        for (int c = 0; c < channels; c++) {
            uint32_t n = ORC_SYNTH_SCHED[lm][b][0], parts = ORC_SYNTH_SCHED[lm][b][1], k = ORC_SYNTH_SCHED[lm][b][2];
            if (n == 1) {
                orc_enc_bits(&e, (uint32_t)(sm_next(&rng) & 1), 1);
                continue;
            }
            for (uint32_t p = 0; p < parts; p++) {
                uint32_t v = orc_pvq_v(n, k);
                orc_enc_uint(&e, sm_below(&rng, v), v); /* a uniform codeword index == encode_pulses(cwrsi(index)) */
            }
        }
    if (e.error) return e.error;
    if (orc_enc_tell(&e) > 8u * (pkt_bytes - 1u)) return ORC_ERR_BUFFER_TOO_SMALL;
    orc_enc_done(&e);
    return e.error ? e.error : (int)pkt_bytes;
}

int orc_synth_packet(uint64_t stream_id, uint64_t frame_idx, int lm, int channels, uint32_t pkt_bytes,
                     uint32_t transient_permille, uint8_t *out)
{
    if (!out || lm < 0 || lm > 3 || channels < 1 || channels > 2 || pkt_bytes < 3 || pkt_bytes > 1276) return ORC_ERR_BAD_ARG;
    int rc = ORC_ERR_BUFFER_TOO_SMALL; /* a draw that overruns the byte budget is redrawn from the next sub-seed */
    for (uint32_t attempt = 0; attempt < 8 && rc == ORC_ERR_BUFFER_TOO_SMALL; attempt++)
        rc = synth_packet_attempt(stream_id, frame_idx, attempt, lm, channels, pkt_bytes, transient_permille, out);
    return rc;
}

typedef struct {
    uint64_t first_stream, first_frame, w0, w1;
    uint32_t n_streams, pkt_bytes, transient_permille;
    int lm, channels, rc;
    uint8_t *out;
} fill_job;

static void *fill_thread(void *arg)
{
    fill_job *j = (fill_job *)arg;
    for (uint64_t w = j->w0; w < j->w1; w++) {
        int r = orc_synth_packet(j->first_stream + w % j->n_streams, j->first_frame + w / j->n_streams, j->lm, j->channels,
                                 j->pkt_bytes, j->transient_permille, j->out + w * j->pkt_bytes);
        if (r < 0) j->rc = r;
    }
    return NULL;
}

/* layout [frame][stream][pkt_bytes], like opn_synth_fill */
int orc_synth_fill(uint64_t first_stream, uint32_t n_streams, uint64_t first_frame, uint32_t n_frames, int lm, int channels,
                   uint32_t pkt_bytes, uint32_t transient_permille, int n_threads, uint8_t *out)
{
    if (!out || n_streams == 0 || n_frames == 0) return ORC_ERR_BAD_ARG;
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    fill_job *jobs = (fill_job *)calloc((size_t)n_threads, sizeof(fill_job));
    const uint64_t total = (uint64_t)n_streams * n_frames;
    for (int t = 0; t < n_threads; t++) {
        fill_job *j = &jobs[t];
        j->first_stream = first_stream; j->first_frame = first_frame;
        j->w0 = total * (uint64_t)t / (uint64_t)n_threads; j->w1 = total * (uint64_t)(t + 1) / (uint64_t)n_threads;
        j->n_streams = n_streams; j->pkt_bytes = pkt_bytes; j->transient_permille = transient_permille;
        j->lm = lm; j->channels = channels; j->out = out;
        pthread_create(&th[t], NULL, fill_thread, j);
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
    int lm, channels, apply_comb;
    float *pcm_last;
    uint32_t rng_xor;
} bench_job;

static void *bench_thread(void *arg)
{
    bench_job *j = (bench_job *)arg;
    int nf = 120 << j->lm;
    orc_synth_state *st = (orc_synth_state *)malloc(sizeof(orc_synth_state));
    float pcm[2 * 960];
    orc_synth_side side;
    uint32_t x = 0;
    for (uint32_t s = j->s0; s < j->s1; s++) {
        orc_synth_state_init(st);
        for (uint32_t f = 0; f < j->n_frames; f++) {
            const uint8_t *pkt = j->packets + ((size_t)f * j->n_streams + s) * j->pkt_bytes;
            /* byte 0 is the TOC */
            orc_synth_decode_frame(st, pkt + 1, j->pkt_bytes - 1, j->lm, j->channels, j->apply_comb,
                                   &side, NULL, NULL, pcm);
            x ^= side.final_rng;
        }
        if (j->pcm_last)
            memcpy(j->pcm_last + (size_t)s * (size_t)nf * (size_t)j->channels, pcm,
                   sizeof(float) * (size_t)nf * (size_t)j->channels);
    }
    j->rng_xor = x;
    free(st);
    return NULL;
}

double orc_synth_bench(const uint8_t *packets, uint32_t n_streams, uint32_t n_frames,
                       uint32_t pkt_bytes, int lm, int channels, int apply_comb, int n_threads,
                       float *pcm_last, uint32_t *final_rng_xor)
{
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > n_streams) n_threads = (int)n_streams;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    bench_job *jobs = (bench_job *)calloc((size_t)n_threads, sizeof(bench_job));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < n_threads; t++) {
        bench_job *j = &jobs[t];
        j->packets = packets;
        j->n_streams = n_streams;
        j->n_frames = n_frames;
        j->pkt_bytes = pkt_bytes;
        j->s0 = (uint32_t)((uint64_t)n_streams * (uint64_t)t / (uint64_t)n_threads);
        j->s1 = (uint32_t)((uint64_t)n_streams * (uint64_t)(t + 1) / (uint64_t)n_threads);
        j->lm = lm;
        j->channels = channels;
        j->apply_comb = apply_comb;
        j->pcm_last = pcm_last;
        pthread_create(&th[t], NULL, bench_thread, j);
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
