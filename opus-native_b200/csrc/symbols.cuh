// symbols.cuh -- kernel 0: warp-per-packet symbol decode (range decoder + PVQ expansion).
//
//   k_rangedec_script : replays an arbitrary list of RangeDecoder calls (operator-level parity
//                       with src/range_coder/decoder.rs and src/celt/pvc.rs).
//   k_synth_symbols   : decodes one SYNTH-CELT/1 frame per warp (DESIGN.md "frame layout"):
//                       flags, post-filter parameters, Laplace coarse energies, raw fine bits,
//                       PVQ pulse vectors -> unit-norm coefficients for the IMDCT kernel.
#pragma once
#include "opn_device.cuh"
#include "opn_internal.h"
#include "rangedec.cuh"

namespace opn {

constexpr int PVQ_TABLE_WORDS = 1272;
constexpr int Y_STAGE = 192;  // >= 176, the largest PVQ part (pvc.rs:309-313)

__device__ __forceinline__ void stage_bytes(uint8_t *dst, const uint8_t *src, uint32_t len, uint32_t lane)
{
    if ((reinterpret_cast<uintptr_t>(src) & 3u) == 0u) {
        uint32_t words = len >> 2;
        const uint32_t *s4 = reinterpret_cast<const uint32_t *>(src);
        uint32_t *d4 = reinterpret_cast<uint32_t *>(dst);
        for (uint32_t w = lane; w < words; w += 32u) d4[w] = __ldg(s4 + w);
        for (uint32_t i = (words << 2) + lane; i < len; i += 32u) dst[i] = __ldg(src + i);
    } else {
        for (uint32_t i = lane; i < len; i += 32u) dst[i] = __ldg(src + i);
    }
    __syncwarp();
}

__device__ __forceinline__ void load_pvq_table(uint32_t *s_data, uint16_t *s_row)
{
    for (int i = threadIdx.x; i < PVQ_TABLE_WORDS; i += blockDim.x) s_data[i] = g_tab.pvq_u_data[i];
    if (threadIdx.x < 15) s_row[threadIdx.x] = g_tab.pvq_u_row[threadIdx.x];
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SYM_WARPS_PER_CTA * 32)
k_rangedec_script(const uint8_t *__restrict__ arena, const uint32_t *__restrict__ offsets,
                  const uint32_t *__restrict__ lens, uint32_t n_packets, const opn_op *__restrict__ ops,
                  uint32_t n_ops, const uint8_t *__restrict__ icdf_pool, opn_op_out *__restrict__ out,
                  int32_t *__restrict__ y_out, uint32_t y_stride, uint32_t pkt_cap)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *s_pvq = reinterpret_cast<uint32_t *>(smem);
    uint16_t *s_row = reinterpret_cast<uint16_t *>(s_pvq + PVQ_TABLE_WORDS);
    int32_t *s_y = reinterpret_cast<int32_t *>(s_row + 16) + (threadIdx.x >> 5) * Y_STAGE;
    uint8_t *s_pkt = reinterpret_cast<uint8_t *>(reinterpret_cast<int32_t *>(s_row + 16) + SYM_WARPS_PER_CTA * Y_STAGE) +
                     (threadIdx.x >> 5) * pkt_cap;
    load_pvq_table(s_pvq, s_row);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t pkt = blockIdx.x * SYM_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (pkt >= n_packets) return;
    const uint32_t len = lens[pkt];
    // Packets that fit are staged in shared memory; oversized test streams (the reference's
    // test_simple_uint_bits stream is ~0.5 MB) are read in place from global memory.
    const uint8_t *bytes = arena + offsets[pkt];
    if (len <= pkt_cap) {
        stage_bytes(s_pkt, bytes, len, lane);
        bytes = s_pkt;
    }
    PvqTable T{s_pvq, s_row};
    RangeDec d;
    d.init(bytes, len);
    uint32_t ny = 0u;
    for (uint32_t i = 0; i < n_ops; i++) {
        const uint32_t op = ops[i].op, a = ops[i].a, b = ops[i].b;
        uint32_t v = 0u;
        switch (op) {
        case OPN_OP_UINT: v = d.uint(a); break;
        case OPN_OP_BITS: v = d.bits(a); break;
        case OPN_OP_BIT_LOGP: v = d.bit_logp(a); break;
        case OPN_OP_ICDF: v = d.icdf(icdf_pool + a, b); break;
        case OPN_OP_LAPLACE: v = (uint32_t)d.laplace(a, b); break;
        case OPN_OP_BIT_VIA_DECODE: {  // src/range_coder/mod.rs:446-454
            uint32_t fs = d.decode(1u << a);
            uint32_t s = fs >= (1u << a) - 1u ? 1u : 0u;
            d.update(s ? (1u << a) - 1u : 0u, (1u << a) - (s ? 0u : 1u), 1u << a);
            v = s;
            break;
        }
        case OPN_OP_BIT_VIA_DECODE_BIN: {  // src/range_coder/mod.rs:455-463
            uint32_t fs = d.decode_bin(a);
            uint32_t s = fs >= (1u << a) - 1u ? 1u : 0u;
            d.update(s ? (1u << a) - 1u : 0u, (1u << a) - (s ? 0u : 1u), 1u << a);
            v = s;
            break;
        }
        case OPN_OP_PULSES: {
            float yy = decode_pulses_warp(d, T, s_y, a, b, lane);
            v = __float_as_uint(yy);
            if (y_out)
                for (uint32_t j = lane; j < a; j += 32u) y_out[(size_t)pkt * y_stride + ny + j] = s_y[j];
            ny += a;
            __syncwarp();
            break;
        }
        case OPN_OP_SHRINK: d.shrink_storage(a); break;
        case OPN_OP_TELL: v = d.tell(); break;
        default: break;
        }
        if (lane == 0u) {
            opn_op_out o;
            o.value = v;
            o.tell_frac = d.tell_frac();
            o.rng = d.rng;
            out[(size_t)pkt * n_ops + i] = o;
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SYM_WARPS_PER_CTA * 32) k_synth_symbols(SymbolArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *s_pvq = reinterpret_cast<uint32_t *>(smem);
    uint16_t *s_row = reinterpret_cast<uint16_t *>(s_pvq + PVQ_TABLE_WORDS);
    int32_t *s_y = reinterpret_cast<int32_t *>(s_row + 16) + (threadIdx.x >> 5) * Y_STAGE;
    uint8_t *s_pkt = reinterpret_cast<uint8_t *>(reinterpret_cast<int32_t *>(s_row + 16) + SYM_WARPS_PER_CTA * Y_STAGE) +
                     (threadIdx.x >> 5) * A.pkt_cap;
    load_pvq_table(s_pvq, s_row);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t item = blockIdx.x * SYM_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (item >= A.n_items) return;
    const uint32_t stream = A.stream_idx ? A.stream_idx[item] : item;
    const int lm = A.lm, C = A.channels, nf = 120 << lm;
    uint32_t len = A.lens[item];
    const uint8_t *src = A.arena + A.offsets[item];
    opn_synth_side *side = A.side + stream;
    float *coef = A.coef ? A.coef + (size_t)stream * C * nf : nullptr;
    int32_t *yo = A.y_out ? A.y_out + (size_t)stream * C * nf : nullptr;

    int32_t status = ITEM_OK;
    if (A.has_toc && len > 0u) {
        // TOC checks a host caller does with query_packet_* (src/lib.rs:219-325) before decode_frame
        const uint32_t toc = __ldg(src);
        if ((toc & 0x80u) == 0u) status = OPN_ERR_UNIMPLEMENTED;                       // SILK / hybrid
        else if ((toc & 0x3u) != 0u) status = OPN_ERR_UNIMPLEMENTED;                   // multi-frame: host path only
        else if ((int)((toc >> 3) & 0x3u) != lm) status = OPN_ERR_FRAME_SIZE_TOO_SMALL;  // frame size != call's
        else if (((toc & 0x4u) ? 2 : 1) != C) status = OPN_ERR_UNIMPLEMENTED;          // mono<->stereo mapping
        src += 1;
        len -= 1u;
    }
    if (status == ITEM_OK && len <= 1u) status = ITEM_LOST;  // src/decoder.rs:467
    if (status == ITEM_OK && len > A.pkt_cap) status = OPN_ERR_INVALID_PACKET;
    if (lane == 0u) A.status[stream] = status;
    if (status < 0) return;

    // zero side info (lanes cooperate: the struct is 95 words)
    {
        uint32_t *sw = reinterpret_cast<uint32_t *>(side);
        for (uint32_t i = lane; i < sizeof(opn_synth_side) / 4u; i += 32u) sw[i] = 0u;
        __syncwarp();
    }
    if (status == ITEM_LOST) {
        if (coef) for (int i = lane; i < C * nf; i += 32) coef[i] = 0.0f;
        if (yo) for (int i = lane; i < C * nf; i += 32) yo[i] = 0;
        return;
    }
    stage_bytes(s_pkt, src, len, lane);
    PvqTable T{s_pvq, s_row};
    RangeDec d;
    d.init(s_pkt, len);

    const uint32_t silence = d.bit_logp(15u);
    uint32_t n_pulses = 0u;
    if (silence) {
        if (coef) for (int i = lane; i < C * nf; i += 32) coef[i] = 0.0f;
        if (yo) for (int i = lane; i < C * nf; i += 32) yo[i] = 0;
        if (lane == 0u) side->silence = 1;
    } else {
        const uint32_t postfilter = d.bit_logp(1u);
        uint32_t octave = 0u, period = 0u, gain_idx = 0u, tapset = 0u;
        if (postfilter) {
            octave = d.uint(6u);
            period = (16u << octave) + d.bits(4u + octave) - 1u;
            gain_idx = d.bits(3u);
            tapset = d.icdf(g_tab.tapset_icdf, 2u);
        }
        const uint32_t transient = d.bit_logp(3u);
        const uint32_t intra = d.bit_logp(3u);
        if (lane == 0u) {
            side->postfilter = (int32_t)postfilter;
            side->octave = (int32_t)octave;
            side->period = (int32_t)period;
            side->gain_idx = (int32_t)gain_idx;
            side->tapset = (int32_t)tapset;
            side->transient = (int32_t)transient;
            side->intra = (int32_t)intra;
        }
        // coarse energy: 21 bands x C Laplace symbols; lane (b*C+c)%32 keeps the value to store
        for (int b = 0; b < 21; b++)
            for (int c = 0; c < C; c++) {
                const uint32_t decay = 6000u + 400u * (uint32_t)b;
                // get_start_freq (src/range_coder/mod.rs:530-534)
                const uint32_t fs0 = ((32768u - 33u) * (16384u - decay)) / (16384u + decay) + 1u;
                int32_t v = d.laplace(fs0, decay);
                if (lane == 0u) side->coarse[c][b] = v;
            }
        for (int b = 0; b < 21; b++)
            for (int c = 0; c < C; c++) {
                uint32_t v = d.bits(2u);
                if (lane == 0u) side->fine[c][b] = (int32_t)v;
            }
        // bins above the last band stay zero
        {
            const int top = 100 << lm;
            for (int c = 0; c < C; c++)
                for (int i = top + (int)lane; i < nf; i += 32) {
                    if (coef) coef[c * nf + i] = 0.0f;
                    if (yo) yo[c * nf + i] = 0;
                }
        }
        for (int b = 0; b < 21; b++)
            for (int c = 0; c < C; c++) {
                const int n = g_tab.synth_sched[lm][b][0], parts = g_tab.synth_sched[lm][b][1],
                          k = g_tab.synth_sched[lm][b][2];
                const int base = c * nf + ((int)g_tab.e_bands[b] << lm);
                if (n == 1) {
                    const uint32_t sign = d.bits(1u);
                    if (lane == 0u) {
                        if (coef) coef[base] = sign ? -0.03125f : 0.03125f;
                        if (yo) yo[base] = sign ? -1 : 1;
                    }
                    n_pulses += 1u;
                    continue;
                }
                for (int p = 0; p < parts; p++) {
                    const float yy = decode_pulses_warp(d, T, s_y, (uint32_t)n, (uint32_t)k, lane);
                    const float g = 0.03125f / sqrtf(yy);
                    for (int j = lane; j < n; j += 32) {
                        const int32_t yv = s_y[j];
                        if (coef) coef[base + p * n + j] = (float)yv * g;
                        if (yo) yo[base + p * n + j] = yv;
                    }
                    __syncwarp();
                    n_pulses += (uint32_t)k;
                }
            }
    }
    if (lane == 0u) {
        side->final_rng = d.rng;
        side->tell_frac = d.tell_frac();
        side->n_pulses = n_pulses;
    }
}

}  // namespace opn
