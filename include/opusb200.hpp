// opusb200.hpp -- header-only C++ mirror of the decode-path surface of the Rust crate
// hasenbanck/opus-native over the C ABI in opusb200.h.  Names, argument meaning and error behaviour
// follow the crate (paths relative to the crate root):
//
//   DecoderConfiguration            src/decoder.rs:27-44   (default: 48 kHz, stereo, gain 0)
//   OpusError                       src/error.rs:5-16
//   Decoder::{new, reset, decode_float, decode::<S>, getters}   src/decoder.rs:54-232
//   BatchDecoder                    the batch-of-streams entry point this engine adds
//
// Rust's Result<T, OpusError> becomes a C++ exception of type opus_native::OpusError; Option<&[u8]>
// becomes a (pointer, length) pair where a null pointer is a lost packet.
#ifndef OPUSB200_HPP
#define OPUSB200_HPP
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "opusb200.h"

namespace opus_native {

enum class SamplingRate : int32_t { Hz8000 = 8000, Hz12000 = 12000, Hz16000 = 16000, Hz24000 = 24000, Hz48000 = 48000 };  // lib.rs:124-135
enum class Channels : int32_t { Mono = 1, Stereo = 2 };                                                                   // lib.rs:111-118
enum class Bandwidth : int32_t { Narrowband = 0, Mediumband = 1, Wideband = 2, Superwideband = 3, Fullband = 4 };           // lib.rs:168-185

// The `Sample` trait (lib.rs:58-107): the output types decode::<S> is implemented for.
template <typename S> struct Sample;
template <> struct Sample<float> { static constexpr int format = OPN_SAMPLE_F32; };
template <> struct Sample<double> { static constexpr int format = OPN_SAMPLE_F64; };
template <> struct Sample<int16_t> { static constexpr int format = OPN_SAMPLE_I16; };
template <> struct Sample<int32_t> { static constexpr int format = OPN_SAMPLE_I32; };
template <> struct Sample<uint16_t> { static constexpr int format = OPN_SAMPLE_U16; };
template <> struct Sample<uint32_t> { static constexpr int format = OPN_SAMPLE_U32; };

struct DecoderConfiguration {  // decoder.rs:27-44
    SamplingRate sampling_rate = SamplingRate::Hz48000;
    Channels channels = Channels::Stereo;
    int16_t gain = 0;
};

class OpusError : public std::runtime_error {  // error.rs:5-16
public:
    enum Kind { BadArguments, BufferToSmall, InternalError, InvalidPacket, FrameSizeTooSmall, Unimplemented, Cuda };
    OpusError(int code)
        : std::runtime_error(std::string(opn_strerror(code)) + (code == OPN_ERR_CUDA ? std::string(": ") + opn_last_cuda_error() : "")),
          code_(code) {}
    int code() const { return code_; }
    Kind kind() const
    {
        switch (code_) {
        case OPN_ERR_BAD_ARG: return BadArguments;
        case OPN_ERR_BUFFER_TOO_SMALL: return BufferToSmall;
        case OPN_ERR_INVALID_PACKET: return InvalidPacket;
        case OPN_ERR_FRAME_SIZE_TOO_SMALL: return FrameSizeTooSmall;
        case OPN_ERR_UNIMPLEMENTED: return Unimplemented;
        case OPN_ERR_CUDA: return Cuda;
        default: return InternalError;
        }
    }

private:
    int code_;
};

inline int check(int rc)
{
    if (rc < 0) throw OpusError(rc);
    return rc;
}

// Decoder, src/decoder.rs:54-232.  One call at a time per object (Rust: &mut self).
class Decoder {
public:
    // Decoder::new, :61.  `bitstream`: OPN_BITSTREAM_OPUS (CELT frames are Unimplemented, as in the crate) or the explicit
    // opt-in OPN_BITSTREAM_SYNTH_CELT_1 (synthetic frame layout, not Opus-interoperable; see opusb200.h).
    explicit Decoder(const DecoderConfiguration &cfg = DecoderConfiguration(), int device = 0, int32_t bitstream = OPN_BITSTREAM_OPUS)
        : cfg_(cfg)
    {
        check(opn_decoder_create(device, (int32_t)cfg.sampling_rate, (int32_t)cfg.channels, cfg.gain, bitstream, &raw_));
    }
    ~Decoder() { opn_decoder_destroy(raw_); }
    Decoder(const Decoder &) = delete;  // Clone (decoder.rs:53) would be a device-to-device copy of the stream slot
    Decoder &operator=(const Decoder &) = delete;
    void reset() { check(opn_decoder_reset(raw_)); }  // :74

    // decode_float (:216-232): returns samples per channel; `samples` holds frame_size * channels floats.
    // packet == nullptr means the packet was lost.
    size_t decode_float(const uint8_t *packet, size_t len, float *samples, size_t samples_len, size_t frame_size, bool decode_fec)
    {
        if (samples_len < frame_size * (size_t)cfg_.channels) throw OpusError(OPN_ERR_BUFFER_TOO_SMALL);
        return (size_t)check(opn_decode_float(raw_, packet, len, samples, frame_size, decode_fec ? 1 : 0));
    }
    // decode::<S> (:148-193): soft clip, then Sample::from_f32; S = int16_t, int32_t, uint16_t, uint32_t, float, double.
    template <typename S>
    size_t decode(const uint8_t *packet, size_t len, S *samples, size_t samples_len, size_t frame_size, bool decode_fec)
    {
        return (size_t)check(opn_decode_pcm(raw_, packet, len, samples, samples_len, Sample<S>::format, frame_size, decode_fec ? 1 : 0));
    }
    SamplingRate sampling_rate() const { return cfg_.sampling_rate; }                        // :80
    Channels channels() const { return cfg_.channels; }                                      // :85
    int16_t gain() const { return cfg_.gain; }                                               // :90
    int32_t bandwidth() const { return opn_decoder_bandwidth(raw_); }                        // :95, -1 = None
    int32_t pitch() const { return opn_decoder_pitch(raw_); }                                // :100, -1 = None
    int32_t last_packet_duration() const { return opn_decoder_last_packet_duration(raw_); }  // :112, -1 = None
    uint32_t final_range() const { return opn_decoder_final_range(raw_); }                   // :121

private:
    opn_decoder *raw_ = nullptr;
    DecoderConfiguration cfg_;
};

// n independent streams on one GPU; per stream the semantics of Decoder::decode_float.
class BatchDecoder {
public:
    BatchDecoder(uint32_t n_streams, const DecoderConfiguration &cfg = DecoderConfiguration(), int device = 0, bool postfilter = true,
                 int32_t bitstream = OPN_BITSTREAM_OPUS)
        : n_(n_streams), channels_((int)cfg.channels)
    {
        opn_config c{(int32_t)cfg.sampling_rate, (int32_t)cfg.channels, cfg.gain, (int16_t)(postfilter ? 1 : 0), bitstream};
        check(opn_batch_create(device, n_streams, &c, &raw_));
    }
    ~BatchDecoder() { opn_batch_destroy(raw_); }
    BatchDecoder(const BatchDecoder &) = delete;
    BatchDecoder &operator=(const BatchDecoder &) = delete;
    uint32_t streams() const { return n_; }
    int channels() const { return channels_; }
    void reset() { check(opn_batch_reset(raw_)); }

    // One packet per stream, host buffers; lens[i] == 0 marks a lost packet.  results[i] = samples per channel
    // or a negative OPN_ERR_* for that stream only.
    void decode_float(const uint8_t *arena, const uint32_t *offsets, const uint32_t *lens, float *pcm, size_t pcm_stride, size_t frame_size,
                      int32_t *results)
    {
        check(opn_batch_decode_float(raw_, arena, offsets, lens, pcm, pcm_stride, frame_size, results, 0));
    }
    // The same, but returns at once with a ticket; at most two calls in flight.  wait(ticket) completes it.
    int submit_float(const uint8_t *arena, const uint32_t *offsets, const uint32_t *lens, float *pcm, size_t pcm_stride, size_t frame_size,
                     int32_t *results)
    {
        return check(opn_batch_decode_float(raw_, arena, offsets, lens, pcm, pcm_stride, frame_size, results, OPN_FLAG_SUBMIT_ONLY));
    }
    // Decoder::decode::<S> for every stream (soft clip + Sample::from_f32 on the device).
    template <typename S>
    void decode(const uint8_t *arena, const uint32_t *offsets, const uint32_t *lens, S *pcm, size_t pcm_stride, size_t frame_size,
                int32_t *results)
    {
        check(opn_batch_decode_pcm(raw_, arena, offsets, lens, pcm, pcm_stride, Sample<S>::format, frame_size, results, 0));
    }
    template <typename S>
    int submit(const uint8_t *arena, const uint32_t *offsets, const uint32_t *lens, S *pcm, size_t pcm_stride, size_t frame_size,
               int32_t *results)
    {
        return check(opn_batch_decode_pcm(raw_, arena, offsets, lens, pcm, pcm_stride, Sample<S>::format, frame_size, results,
                                          OPN_FLAG_SUBMIT_ONLY));
    }
    void wait(int ticket) { check(opn_batch_wait(raw_, ticket)); }
    void synchronize() { check(opn_batch_synchronize(raw_)); }
    std::vector<uint32_t> final_ranges()
    {
        std::vector<uint32_t> v(n_);
        check(opn_batch_final_ranges(raw_, v.data()));
        return v;
    }
    opn_batch *raw() { return raw_; }

private:
    opn_batch *raw_ = nullptr;
    uint32_t n_;
    int channels_;
};

}  // namespace opus_native
#endif
