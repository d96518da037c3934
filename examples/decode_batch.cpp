// decode_batch.cpp -- the decode path from compiled host code, through include/opusb200.hpp.
//
//   decode_batch [n_streams] [n_frames]
//
// Generates SYNTH-CELT/1 packets (20 ms, stereo, 160 bytes) for n_streams streams, decodes them with a
// BatchDecoder (two calls in flight), decodes stream 0 again with a single-stream Decoder, checks that both
// agree bit for bit, and prints one line with a PCM checksum and the last final_range of stream 0.
// Exit status: 0 ok, 1 mismatch, 2 library error (e.g. no sm_100 device: the library has no CPU path).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "opusb200.hpp"

int main(int argc, char **argv)
{
    using namespace opus_native;
    const uint32_t ns = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 256u;
    const uint32_t nfr = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 6u;
    const int lm = 3, channels = 2;
    const uint32_t pkt = 160, nf = 960;
    std::vector<uint8_t> packets((size_t)nfr * ns * pkt);
    if (opn_synth_fill(0, ns, 0, nfr, lm, channels, pkt, 100, 4, packets.data()) < 0) return 2;
    std::vector<uint32_t> offsets(ns), lens(ns, pkt);
    for (uint32_t s = 0; s < ns; s++) offsets[s] = s * pkt;
    try {
        BatchDecoder batch(ns, DecoderConfiguration(), 0, true, OPN_BITSTREAM_SYNTH_CELT_1);  // the synthetic frame layout is an explicit opt-in
        Decoder single(DecoderConfiguration(), 0, OPN_BITSTREAM_SYNTH_CELT_1);
        std::vector<float> pcm[2] = {std::vector<float>((size_t)ns * nf * channels), std::vector<float>((size_t)ns * nf * channels)};
        std::vector<int32_t> res[2] = {std::vector<int32_t>(ns), std::vector<int32_t>(ns)};
        std::vector<float> one(nf * channels);
        double checksum = 0.0;
        int ticket = -1;
        bool same = true;
        auto finish = [&](uint32_t f) {  // frame f has landed in pcm[f & 1]
            batch.wait(ticket);
            const std::vector<float> &p = pcm[f & 1];
            for (float v : p) checksum += std::fabs(v);
            for (int32_t r : res[f & 1]) same = same && r == (int32_t)nf;
            const size_t got = single.decode_float(&packets[(size_t)f * ns * pkt], pkt, one.data(), one.size(), nf, false);
            same = same && got == nf && std::memcmp(one.data(), p.data(), one.size() * sizeof(float)) == 0;
        };
        for (uint32_t f = 0; f < nfr; f++) {
            const int t = batch.submit_float(&packets[(size_t)f * ns * pkt], offsets.data(), lens.data(), pcm[f & 1].data(), nf * channels, nf,
                                             res[f & 1].data());
            if (f > 0) finish(f - 1);
            ticket = t;
        }
        finish(nfr - 1);
        const uint32_t fr = batch.final_ranges()[0];
        std::printf("streams=%u frames=%u checksum=%.6f final_range[0]=%08x single==batch:%s\n", ns, nfr, checksum, fr,
                    same && fr == single.final_range() ? "yes" : "NO");
        return same && fr == single.final_range() ? 0 : 1;
    } catch (const OpusError &e) {
        std::fprintf(stderr, "OpusError(%d): %s\n", e.code(), e.what());
        return 2;
    }
}
