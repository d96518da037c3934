// host_packet.cpp -- Opus TOC / framing inspection on the host (src/lib.rs:219-512).
// These run on the CPU by design: they read 1-3 bytes per packet and decide how the batch is
// bucketed before anything is copied to the GPU.
#include <cstddef>
#include <cstdint>

#include "../../include/opusb200.h"

namespace {

// A bounds-checked forward reader over the packet; every framing field is pulled through it so
// that a truncated packet turns into OPN_ERR_INVALID_PACKET instead of an out-of-bounds read
// (the Rust code relies on slice panics for the same cases).
struct ByteReader {
    const uint8_t *p;
    size_t len, pos;
    bool ok;
    ByteReader(const uint8_t *p_, size_t len_) : p(p_), len(len_), pos(0), ok(true) {}
    size_t remaining() const { return len - pos; }
    int byte()
    {
        if (pos >= len) {
            ok = false;
            return 0;
        }
        return p[pos++];
    }
    // parse_size, lib.rs:500-512: one or two byte frame length
    bool frame_length(uint32_t &out)
    {
        int b0 = byte();
        if (!ok) return false;
        if (b0 < 252) {
            out = (uint32_t)b0;
            return true;
        }
        int b1 = byte();
        if (!ok) return false;
        out = 4u * (uint32_t)b1 + (uint32_t)b0;
        return true;
    }
};

const uint8_t kBandwidthOfConfig[32] = {
    OPN_BW_NARROW, OPN_BW_NARROW, OPN_BW_NARROW, OPN_BW_NARROW, OPN_BW_MEDIUM, OPN_BW_MEDIUM, OPN_BW_MEDIUM, OPN_BW_MEDIUM,
    OPN_BW_WIDE, OPN_BW_WIDE, OPN_BW_WIDE, OPN_BW_WIDE, OPN_BW_SUPERWIDE, OPN_BW_SUPERWIDE, OPN_BW_FULL, OPN_BW_FULL,
    OPN_BW_NARROW, OPN_BW_NARROW, OPN_BW_NARROW, OPN_BW_NARROW, OPN_BW_WIDE, OPN_BW_WIDE, OPN_BW_WIDE, OPN_BW_WIDE,
    OPN_BW_SUPERWIDE, OPN_BW_SUPERWIDE, OPN_BW_SUPERWIDE, OPN_BW_SUPERWIDE, OPN_BW_FULL, OPN_BW_FULL, OPN_BW_FULL, OPN_BW_FULL};

}  // namespace

extern "C" {

int opn_packet_bandwidth(const uint8_t *packet) { return kBandwidthOfConfig[packet[0] >> 3]; }  // lib.rs:150-190,219-224

int opn_packet_channels(const uint8_t *packet) { return (packet[0] & 0x4) ? 2 : 1; }  // lib.rs:233-241

int opn_packet_frame_count(const uint8_t *packet, size_t len)  // lib.rs:250-263
{
    switch (packet[0] & 0x3) {
    case 0: return 1;
    case 1:
    case 2: return 2;
    default: return len < 2 ? OPN_ERR_INVALID_PACKET : (packet[1] & 0x3F);
    }
}

int opn_packet_samples_per_frame(const uint8_t *packet, int32_t fs)  // lib.rs:271-289
{
    const int toc = packet[0];
    const int size_code = (toc >> 3) & 0x3;
    if (toc & 0x80) return (fs << size_code) / 400;           // CELT: 2.5 ms << code
    if ((toc & 0x60) == 0x60) return (toc & 0x08) ? fs / 50 : fs / 100;  // hybrid: 10 or 20 ms
    return size_code == 3 ? fs * 60 / 1000 : (fs << size_code) / 100;     // SILK: 10/20/40/60 ms
}

int opn_packet_sample_count(const uint8_t *packet, size_t len, int32_t fs)  // lib.rs:299-310
{
    const int frames = opn_packet_frame_count(packet, len);
    if (frames < 0) return frames;
    const int samples = frames * opn_packet_samples_per_frame(packet, fs);
    return samples * 25 > fs * 3 ? OPN_ERR_INVALID_PACKET : samples;  // more than 120 ms
}

int opn_packet_mode(const uint8_t *packet)  // lib.rs:317-325
{
    if (packet[0] & 0x80) return OPN_MODE_CELT;
    return (packet[0] & 0x60) == 0x60 ? OPN_MODE_HYBRID : OPN_MODE_SILK;
}

// parse_packet, lib.rs:345-498
int opn_parse_packet(const uint8_t *packet, size_t len, int self_delimited, uint32_t frames[48], uint32_t sizes[48],
                     uint32_t *payload_offset, uint32_t *packet_offset)
{
    if (packet == nullptr || len == 0 || sizes == nullptr) return OPN_ERR_BAD_ARG;
    ByteReader rd(packet, len);
    const int toc = rd.byte();
    const int samples_per_frame = opn_packet_samples_per_frame(packet, 48000);
    int count = 0;
    bool cbr = false;
    size_t padding = 0;
    long body = (long)rd.remaining();  // bytes that belong to frames (shrinks as headers are consumed)
    long last = body;                  // size of the last frame

    switch (toc & 0x3) {
    case 0: count = 1; break;
    case 1:
        count = 2;
        cbr = true;
        if (!self_delimited) {
            if (body & 1) return OPN_ERR_INVALID_PACKET;
            last = body / 2;
            sizes[0] = (uint32_t)last;
        }
        break;
    case 2: {
        count = 2;
        const size_t before = rd.pos;
        if (!rd.frame_length(sizes[0])) return OPN_ERR_INVALID_PACKET;
        body -= (long)(rd.pos - before);
        if ((long)sizes[0] > body) return OPN_ERR_INVALID_PACKET;
        last = body - (long)sizes[0];
        break;
    }
    default: {
        if (body < 1) return OPN_ERR_INVALID_PACKET;
        const int ch = rd.byte();
        body -= 1;
        count = ch & 0x3F;
        if (count == 0 || samples_per_frame * count > 5760) return OPN_ERR_INVALID_PACKET;
        if (ch & 0x40) {  // padding
            int p;
            do {
                if (body <= 0) return OPN_ERR_INVALID_PACKET;
                p = rd.byte();
                body -= 1;
                const int chunk = p == 255 ? 254 : p;
                body -= chunk;
                padding += (size_t)chunk;
            } while (p == 255);
            if (body < 0) return OPN_ERR_INVALID_PACKET;
        }
        cbr = (ch & 0x80) == 0;
        if (!cbr) {
            last = body;
            for (int i = 0; i < count - 1; i++) {
                const size_t before = rd.pos;
                if (!rd.frame_length(sizes[i])) return OPN_ERR_INVALID_PACKET;
                const long hdr = (long)(rd.pos - before);
                body -= hdr;
                if ((long)sizes[i] > body) return OPN_ERR_INVALID_PACKET;
                last -= hdr + (long)sizes[i];
            }
            if (last < 0) return OPN_ERR_INVALID_PACKET;
        } else if (!self_delimited) {
            last = body / count;
            if (last * count != body) return OPN_ERR_INVALID_PACKET;
            for (int i = 0; i < count - 1; i++) sizes[i] = (uint32_t)last;
        }
        break;
    }
    }

    if (self_delimited) {
        const size_t before = rd.pos;
        if (!rd.frame_length(sizes[count - 1])) return OPN_ERR_INVALID_PACKET;
        const long hdr = (long)(rd.pos - before);
        body -= hdr;
        if ((long)sizes[count - 1] > body) return OPN_ERR_INVALID_PACKET;
        if (cbr) {
            if ((long)sizes[count - 1] * count > body) return OPN_ERR_INVALID_PACKET;
            for (int i = 0; i < count - 1; i++) sizes[i] = sizes[count - 1];
        } else if (hdr + (long)sizes[count - 1] > last) {
            return OPN_ERR_INVALID_PACKET;
        }
    } else {
        if (last > 1275) return OPN_ERR_INVALID_PACKET;
        sizes[count - 1] = (uint32_t)last;
    }

    size_t at = rd.pos;
    if (payload_offset) *payload_offset = (uint32_t)at;
    for (int i = 0; i < count; i++) {
        if (frames) frames[i] = (uint32_t)at;
        at += sizes[i];
    }
    if (packet_offset) *packet_offset = (uint32_t)(padding + at);
    return count;
}

}  // extern "C"
