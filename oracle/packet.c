/* CPU ORACLE (test infrastructure) -- TOC queries, packet parsing, soft clip, cross-fade.
 * Restates /root/reference/src/lib.rs:219-632 and src/decoder.rs:833-865. */
#include "oracle.h"
#include "oracle_tables.h"

#include <math.h>

/* lib.rs:150-190, 219-224 */
int orc_packet_bandwidth(const uint8_t *p)
{
    static const uint8_t tab[32] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 4, 4,
                                    0, 0, 0, 0, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
    return tab[(p[0] & 0xF8) >> 3];
}

/* lib.rs:233-241 */
int orc_packet_channels(const uint8_t *p) { return (p[0] & 0x4) ? 2 : 1; }

/* lib.rs:250-263 */
int orc_packet_frame_count(const uint8_t *p, size_t len)
{
    int count = p[0] & 0x3;
    if (count == 0) return 1;
    if (count != 3) return 2;
    if (len < 2) return ORC_ERR_INVALID_PACKET;
    return p[1] & 0x3F;
}

/* lib.rs:271-289 */
int orc_packet_samples_per_frame(const uint8_t *p, int fs)
{
    if (p[0] & 0x80) {
        int audio_size = (p[0] >> 3) & 0x3;
        return (fs << audio_size) / 400;
    } else if ((p[0] & 0x60) == 0x60) {
        return (p[0] & 0x08) ? fs / 50 : fs / 100;
    } else {
        int audio_size = (p[0] >> 3) & 0x3;
        if (audio_size == 3) return fs * 60 / 1000;
        return (fs << audio_size) / 100;
    }
}

/* lib.rs:299-310 */
int orc_packet_sample_count(const uint8_t *p, size_t len, int fs)
{
    int count = orc_packet_frame_count(p, len);
    if (count < 0) return count;
    int samples = count * orc_packet_samples_per_frame(p, fs);
    if (samples * 25 > fs * 3) return ORC_ERR_INVALID_PACKET;
    return samples;
}

/* lib.rs:317-325 */
int orc_packet_mode(const uint8_t *p)
{
    if ((p[0] & 0x80) == 0x80) return 2;
    if ((p[0] & 0x60) == 0x60) return 1;
    return 0;
}

/* lib.rs:500-512 */
static int parse_size(const uint8_t *d, size_t len, uint32_t *size)
{
    if (len == 0) return ORC_ERR_INVALID_PACKET;
    if (d[0] < 252) {
        *size = d[0];
        return 1;
    }
    if (len < 2) return ORC_ERR_INVALID_PACKET;
    *size = 4u * d[1] + d[0];
    return 2;
}

/* lib.rs:345-498.  The Rust code indexes slices and would panic where `len` underflows
 * (e.g. padding running past the packet); those cases return INVALID_PACKET here. */
int orc_parse_packet(const uint8_t *p, size_t plen, int self_delimited, uint32_t frames[48],
                     uint32_t sizes[48], uint32_t *payload_offset, uint32_t *packet_offset)
{
    int framesize = orc_packet_samples_per_frame(p, 48000);
    size_t offset = 1;
    long len = (long)plen - 1;
    long last_size = len;
    int cbr = 0;
    size_t pad = 0;
    int count;

    switch (p[0] & 0x3) {
    case 0: count = 1; break;
    case 1:
        count = 2;
        cbr = 1;
        if (!self_delimited) {
            if (len & 0x1) return ORC_ERR_INVALID_PACKET;
            last_size = len / 2;
            sizes[0] = (uint32_t)last_size;
        }
        break;
    case 2: {
        count = 2;
        int bytes = parse_size(p + offset, (size_t)len, &sizes[0]);
        if (bytes < 0) return bytes;
        len -= bytes;
        if ((long)sizes[0] > len) return ORC_ERR_INVALID_PACKET;
        offset += (size_t)bytes;
        last_size = len - (long)sizes[0];
        break;
    }
    default: {
        if (len < 1) return ORC_ERR_INVALID_PACKET;
        int ch = p[offset++];
        count = ch & 0x3F;
        if (framesize * count > 5760) return ORC_ERR_INVALID_PACKET;
        if (count == 0) return ORC_ERR_INVALID_PACKET; /* Rust: (0..count-1) underflows */
        len -= 1;
        if (ch & 0x40) {
            int pp = 255;
            while (pp == 255) {
                if (len <= 0) return ORC_ERR_INVALID_PACKET;
                pp = p[offset++];
                len -= 1;
                int tmp = pp == 255 ? 254 : pp;
                len -= tmp;
                pad += (size_t)tmp;
            }
        }
        if (len < 0) return ORC_ERR_INVALID_PACKET;
        cbr = (ch & 0x80) == 0;
        if (!cbr) {
            last_size = len;
            for (int i = 0; i < count - 1; i++) {
                int bytes = parse_size(p + offset, (size_t)len, &sizes[i]);
                if (bytes < 0) return bytes;
                len -= bytes;
                if ((long)sizes[i] > len) return ORC_ERR_INVALID_PACKET;
                offset += (size_t)bytes;
                last_size -= bytes + (long)sizes[i];
            }
            if (last_size < 0) return ORC_ERR_INVALID_PACKET;
        } else if (!self_delimited) {
            last_size = len / count;
            if (last_size * count != len) return ORC_ERR_INVALID_PACKET;
            for (int i = 0; i < count - 1; i++) sizes[i] = (uint32_t)last_size;
        }
        break;
    }
    }

    if (self_delimited) {
        int bytes = parse_size(p + offset, (size_t)len, &sizes[count - 1]);
        if (bytes < 0) return bytes;
        len -= bytes;
        if ((long)sizes[count - 1] > len) return ORC_ERR_INVALID_PACKET;
        offset += (size_t)bytes;
        if (cbr) {
            if ((long)sizes[count - 1] * count > len) return ORC_ERR_INVALID_PACKET;
            for (int i = 0; i < count - 1; i++) sizes[i] = sizes[count - 1];
        } else if (bytes + (long)sizes[count - 1] > last_size) {
            return ORC_ERR_INVALID_PACKET;
        }
    } else {
        if (last_size > 1275) return ORC_ERR_INVALID_PACKET;
        sizes[count - 1] = (uint32_t)last_size;
    }
    if (payload_offset) *payload_offset = (uint32_t)offset;
    for (int i = 0; i < count; i++) {
        if (frames) frames[i] = (uint32_t)offset;
        offset += sizes[i];
    }
    if (packet_offset) *packet_offset = (uint32_t)(pad + offset);
    return count;
}

static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* lib.rs:526-632, restated as written.  NOTE (reference quirk, kept): the search loop at
 * lib.rs:556-562 leaves `pos == frame_size - 1` when no sample exceeds +-1, so the
 * `pos == frame_size` exit at :564 only triggers for an empty range; the last region of every
 * channel is therefore always run through the non-linearity. */
void orc_pcm_soft_clip(float *pcm, size_t total_len, size_t channels, float *mem, size_t mem_len)
{
    if (total_len == 0 || channels == 0 || mem_len < channels) return;
    size_t frame_size = total_len / channels;
    for (size_t i = 0; i < total_len; i++) pcm[i] = clampf(pcm[i], -2.0f, 2.0f);
    for (size_t c = 0; c < channels; c++) {
        float a = mem[c];
        for (size_t i = 0; i < frame_size; i++) {
            size_t off = c + i * channels;
            if (pcm[off] * a >= 0.0f) break;
            pcm[off] += a * pcm[off] * pcm[off];
        }
        size_t curr = 0;
        float x0 = pcm[c];
        for (;;) {
            size_t pos = 0;
            for (size_t i = curr; i < frame_size; i++) {
                pos = i;
                if (pcm[c + pos * channels] > 1.0f || pcm[c + pos * channels] < -1.0f) break;
            }
            if (pos == frame_size) {
                a = 0.0f;
                break;
            }
            size_t peak_pos = pos, start = pos, end = pos;
            float maxval = fabsf(pcm[c + pos * channels]);
            while (start > 0 && pcm[c + pos * channels] * pcm[c + (start - 1) * channels] >= 0.0f) start -= 1;
            while (end < frame_size && pcm[c + pos * channels] * pcm[c + end * channels] >= 0.0f) {
                if (fabsf(pcm[c + end * channels]) > maxval) {
                    maxval = fabsf(pcm[c + end * channels]);
                    peak_pos = end;
                }
                end += 1;
            }
            int special = start == 0 && (pcm[c + pos * channels] * pcm[c]) >= 0.0f;
            a = (maxval - 1.0f) / (maxval * maxval);
            a += a * 2.4e-7f;
            if (pcm[c + pos * channels] > 0.0f) a = -a;
            for (size_t i = start; i < end; i++) {
                size_t off = c + i * channels;
                pcm[off] += a * pcm[off] * pcm[off];
            }
            if (special && peak_pos >= 2) {
                float offset = x0 - pcm[c];
                float delta = offset / (float)peak_pos;
                for (size_t i = curr; i < peak_pos; i++) {
                    size_t off = c + i * channels;
                    offset -= delta;
                    pcm[off] += offset;
                    pcm[off] = clampf(pcm[off], -1.0f, 1.0f);
                }
            }
            curr = end;
            if (curr == frame_size) break;
        }
        mem[c] = a;
    }
}

/* Sample::from_f32 (lib.rs:63-107), restated with Rust's cast semantics spelled out: `x as iN/uN` truncates toward
 * zero, saturates at the type's bounds and maps NaN to 0; f32::clamp returns NaN for NaN.  The literals are the
 * crate's, evaluated as f32: 2_147_483_647.0 is 2^31, so the i32 clamp lets 2^31 through and the cast saturates
 * it to i32::MAX; the unsigned upper bounds are midpoint + full scale (32768, 2^31), as written in the crate.
 * format: 0 f32, 1 i16, 2 i32, 3 u16, 4 u32, 5 f64 (include/opusb200.h OPN_SAMPLE_*). */
static int64_t rust_f32_as_int(float x, int64_t lo, int64_t hi)
{
    if (x != x) return 0;
    if (x <= (float)lo) return lo;
    if (x >= (float)hi) return hi;
    return (int64_t)x; /* in range: C truncates toward zero like Rust */
}
int orc_sample_from_f32(int format, const float *in, void *out, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        float f = in[i];
        switch (format) {
        case 0: ((float *)out)[i] = f; break;                                               /* lib.rs:63-68 */
        case 5: ((double *)out)[i] = (double)f; break;                                      /* lib.rs:70-75 */
        case 1: {                                                                           /* lib.rs:77-83 */
            float x = f * 32768.0f;
            x = x != x ? x : clampf(x, -32768.0f, 32767.0f);
            ((int16_t *)out)[i] = (int16_t)rust_f32_as_int(x, -32768, 32767);
            break;
        }
        case 2: {                                                                           /* lib.rs:85-91 */
            float x = f * 2147483648.0f;
            x = x != x ? x : clampf(x, -2147483648.0f, 2147483648.0f /* 2_147_483_647.0 as f32 */);
            ((int32_t *)out)[i] = (int32_t)rust_f32_as_int(x, -2147483648LL, 2147483647LL);
            break;
        }
        case 3: {                                                                           /* lib.rs:93-99 */
            float x = f * 32768.0f + 32768.0f;
            x = x != x ? x : clampf(x, 0.0f, 32768.0f);
            ((uint16_t *)out)[i] = (uint16_t)rust_f32_as_int(x, 0, 65535);
            break;
        }
        case 4: {                                                                           /* lib.rs:101-107 */
            float x = f * 2147483648.0f + 2147483648.0f;
            x = x != x ? x : clampf(x, 0.0f, 2147483648.0f);
            ((uint32_t *)out)[i] = (uint32_t)rust_f32_as_int(x, 0, 4294967295LL);
            break;
        }
        default: return -1;
        }
    }
    return 0;
}

/* decoder.rs:833-865: out = w^2*in2 + (1-w^2)*in1 */
void orc_smooth_fade(const float *in1, const float *in2, float *out, int overlap, int channels, int fs)
{
    int inc = 48000 / fs;
    for (int c = 0; c < channels; c++)
        for (int i = 0; i < overlap; i++) {
            float w = ORC_WINDOW[i * inc] * ORC_WINDOW[i * inc];
            out[c + i * channels] = (w * in2[i * channels + c]) + ((1.0f - w) * in1[i * channels + c]);
        }
}
