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
// SYNTH-CELT/1 symbol decode, two kernels:
//
//   k_synth_rangedec  ONE LANE decodes one packet.  The entropy decoder is a strictly serial
//      dependency chain per packet (two divisions and a byte refill per symbol); a warp that
//      shares one chain between its 32 lanes issues ~60 warp-instructions per symbol for one
//      packet, a warp whose lanes each own a chain issues about the same for 32 packets.  The chain
//      only produces what depends on it: flags, post-filter parameters, energies and, for every PVQ
//      part, the codeword INDEX (decode_uint with the host-precomputed alphabet split and
//      reciprocal).  Packet bytes are read in place through L1 (each lane walks its own packet
//      front-to-back and back-to-front, so every 128-byte line is fetched once).
//   k_synth_expand    ONE WARP expands one packet's indices into pulse vectors (cwrsi never feeds
//      back into the range decoder): the parts are dealt to the 32 lanes by a host-computed
//      longest-first schedule, one part per lane and slot; a lane walks its part in events (one run of
//      empty dimensions, found by bisection, plus the occupied dimension after it), then the warp writes
//      pulses x gain as coalesced float4 coefficient rows.
//
// Shared memory of k_synth_expand (per CTA of EXPAND_WARPS_PER_CTA warps):
//   PVQ U(n,k) table 5088 B + bisection table 10176 B (both by TMA) + row offsets 32 B + mbarrier 16 B |
//   entry table 72 x 16 B | per warp: codeword indices 72 x 4 B, gains 76 x 4 B (slot 72 = 0 for bins without a part), 16-bit pulses 2 x 960 x 2 B.
constexpr int SYM_Y16 = 2 * 960;
constexpr int SYNTH_GAIN_SLOTS = SYNTH_MAX_ENTRIES + 4;
constexpr size_t SYM_EXPAND_WARP_BYTES = SYNTH_MAX_ENTRIES * 4 + SYNTH_GAIN_SLOTS * 4 + SYM_Y16 * 2;
__host__ __device__ constexpr size_t synth_expand_smem()
{
    return 3 * PVQ_TABLE_WORDS * 4 + 32 + 16 + SYNTH_MAX_ENTRIES * sizeof(SynthEntry) + (size_t)EXPAND_WARPS_PER_CTA * SYM_EXPAND_WARP_BYTES;
}

__global__ void __launch_bounds__(RANGEDEC_WARPS_PER_CTA * 32) k_synth_rangedec(SymbolArgs A)
{
    __shared__ SynthEntry s_ent[SYNTH_MAX_ENTRIES];
    __shared__ uint32_t s_fs0[21];  // get_start_freq(decay) per band (src/range_coder/mod.rs:530-534)
    const uint32_t lane = threadIdx.x;  // index inside the CTA: one packet per thread
    const int lm = A.lm, C = A.channels;
    if (lane < 21u) {
        const uint32_t decay = 6000u + 400u * lane;
        s_fs0[lane] = ((32768u - 33u) * (16384u - decay)) / (16384u + decay) + 1u;
    }
    const int ne = g_tab.synth_n_entries[lm][C - 1];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_tab.synth_entries[lm][C - 1]);
        uint4 *dst = reinterpret_cast<uint4 *>(s_ent);
        for (int i = lane; i < ne; i += RANGEDEC_WARPS_PER_CTA * 32) dst[i] = src[i];
    }
    __syncthreads();
    const uint32_t item = blockIdx.x * (RANGEDEC_WARPS_PER_CTA * 32u) + lane;
    if (item >= A.n_items) return;
    const uint32_t stream = A.stream_idx ? A.stream_idx[item] : item;
    uint32_t len = A.lens[item];
    const uint8_t *src = A.arena + A.offsets[item];

    int32_t status = len == 0u ? ITEM_LOST : ITEM_OK;
    if (A.has_toc && len > 0u) {
        // TOC checks a host caller does with query_packet_* (src/lib.rs:219-325) before decode_frame
        const uint32_t toc = __ldg(src);
        if ((toc & 0x80u) == 0u) status = OPN_ERR_UNIMPLEMENTED;                         // SILK / hybrid
        else if ((toc & 0x3u) != 0u) status = OPN_ERR_UNIMPLEMENTED;                     // multi-frame: host path only
        else if ((int)((toc >> 3) & 0x3u) != lm) status = OPN_ERR_FRAME_SIZE_TOO_SMALL;  // frame size != call's
        else if (((toc & 0x4u) ? 2 : 1) != C) status = OPN_ERR_UNIMPLEMENTED;            // mono<->stereo mapping
        src += 1;
        len -= 1u;
    }
    if (status == ITEM_OK && len <= 1u) status = ITEM_LOST;  // decoder.rs:467: len <= 1 means PLC/DTX
    A.status[stream] = status;
    if (status < 0) return;

    opn_synth_side *sd = A.side + stream;
    uint32_t *sw = reinterpret_cast<uint32_t *>(sd);
    constexpr uint32_t SIDE_WORDS = sizeof(opn_synth_side) / 4u;
    if (status == ITEM_LOST) {
        for (uint32_t i = 0; i < SIDE_WORDS; i++) sw[i] = 0u;
        return;
    }
    RangeDec d;
    d.init(src, len);
    const uint32_t silence = d.bit_logp(15u);
    if (silence) {
        for (uint32_t i = 0; i < SIDE_WORDS; i++) sw[i] = 0u;
        sd->silence = 1;
        sd->final_rng = d.rng;
        sd->tell_frac = d.tell_frac();
        return;
    }
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
    sd->silence = 0;
    sd->postfilter = (int32_t)postfilter;
    sd->octave = (int32_t)octave;
    sd->period = (int32_t)period;
    sd->gain_idx = (int32_t)gain_idx;
    sd->tapset = (int32_t)tapset;
    sd->transient = (int32_t)transient;
    sd->intra = (int32_t)intra;
    for (int b = 0; b < 21; b++) {
        const uint32_t decay = 6000u + 400u * (uint32_t)b;
        const uint32_t fs0 = s_fs0[b];
        for (int c = 0; c < C; c++) sd->coarse[c][b] = d.laplace(fs0, decay);
        if (C == 1) sd->coarse[1][b] = 0;
    }
    for (int b = 0; b < 21; b++) {
        for (int c = 0; c < C; c++) sd->fine[c][b] = (int32_t)d.bits(2u);
        if (C == 1) sd->fine[1][b] = 0;
    }
    uint32_t *idx = A.idx + (size_t)stream * SYNTH_MAX_ENTRIES;
    uint32_t n_pulses = 0u;
    for (int e = 0; e < ne; e++) {
        const SynthEntry E = s_ent[e];
        uint32_t v;
        if (E.n == 1) {
            v = d.bits(1u);
            n_pulses += 1u;
        } else {
            v = d.uint_precomputed(E.ft_minus1, E.ft1, E.ftb, E.magic, E.sh);
            n_pulses += E.k;
        }
        idx[e] = v;
    }
    sd->final_rng = d.rng;
    sd->tell_frac = d.tell_frac();
    sd->n_pulses = n_pulses;
}

__global__ void __launch_bounds__(EXPAND_WARPS_PER_CTA * 32) k_synth_expand(SymbolArgs A)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t *s_pvq = reinterpret_cast<uint32_t *>(smem);
    uint2 *s_cw = reinterpret_cast<uint2 *>(s_pvq + PVQ_TABLE_WORDS);
    uint16_t *s_row = reinterpret_cast<uint16_t *>(s_cw + PVQ_TABLE_WORDS);
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_row + 16);
    SynthEntry *s_ent = reinterpret_cast<SynthEntry *>(bar + 2);
    uint8_t *wbase = reinterpret_cast<uint8_t *>(s_ent + SYNTH_MAX_ENTRIES) + (size_t)warp * SYM_EXPAND_WARP_BYTES;
    int16_t *s_y = reinterpret_cast<int16_t *>(wbase);  // 16-byte aligned: zeroed as uint4, read back as 4 x int16
    uint32_t *s_idx = reinterpret_cast<uint32_t *>(s_y + SYM_Y16);
    float *s_gain = reinterpret_cast<float *>(s_idx + SYNTH_MAX_ENTRIES);

    const int lm = A.lm, C = A.channels, nf = 120 << lm;
    const int ne = g_tab.synth_n_entries[lm][C - 1];
    // the two PVQ tables (15 KB) arrive by TMA while the warps fetch their indices and clear their pulse rows
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, 3 * PVQ_TABLE_WORDS * 4);
        bulk_g2s(s_pvq, g_tab.pvq_u_data, PVQ_TABLE_WORDS * 4, bar);
        bulk_g2s(s_cw, g_tab.pvq_cw_data, 2 * PVQ_TABLE_WORDS * 4, bar);
    }
    if (threadIdx.x < 15) s_row[threadIdx.x] = g_tab.pvq_u_row[threadIdx.x];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_tab.synth_entries[lm][C - 1]);
        uint4 *dst = reinterpret_cast<uint4 *>(s_ent);
        for (int i = threadIdx.x; i < ne; i += blockDim.x) dst[i] = src[i];
    }
    const uint32_t item = blockIdx.x * EXPAND_WARPS_PER_CTA + warp;
    const bool in_range = item < A.n_items;
    const uint32_t stream = in_range ? (A.stream_idx ? A.stream_idx[item] : item) : 0u;
    // status, silence flag and codeword indices are requested together (one memory round trip, not three): the
    // index rows exist for every stream, whatever the status turns out to be
    const int32_t status = in_range ? A.status[stream] : -1;
    const int32_t silence = in_range ? A.side[stream].silence : 0;
    if (in_range)
        for (int e = lane; e < ne; e += 32) s_idx[e] = A.idx[(size_t)stream * SYNTH_MAX_ENTRIES + e];
    const bool zero_frame = status == ITEM_LOST || (status >= 0 && silence != 0);
    for (int i = lane; i < C * nf / 8; i += 32) reinterpret_cast<uint4 *>(s_y)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();  // entry table, row offsets and the mbarrier are set up
    if (status < 0) {
        mbar_wait(bar, 0);  // do not let the CTA retire under the table transfer
        return;
    }

    float4 *coef4 = A.coef ? reinterpret_cast<float4 *>(A.coef + (size_t)stream * C * nf) : nullptr;
    int4 *yo4 = A.y_out ? reinterpret_cast<int4 *>(A.y_out + (size_t)stream * C * nf) : nullptr;
    const int nvec = C * nf / 4;
    if (zero_frame) {
        for (int i = lane; i < nvec; i += 32) {
            if (coef4) coef4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yo4) yo4[i] = make_int4(0, 0, 0, 0);
        }
        mbar_wait(bar, 0);
        return;
    }
    // bin -> part map of the coefficient write at the end: fetched now (15 words per lane at most), so the loads
    // are long complete when the pulses are (the map sits in global memory; read in the loop, each L1 round trip
    // was exposed: 28 % of the kernel's stall samples)
    constexpr int MAX_VEC_PER_LANE = 2 * 960 / 4 / 32;
    uint32_t ids[MAX_VEC_PER_LANE];
    {
        const uint32_t *ent4 = reinterpret_cast<const uint32_t *>(g_tab.synth_entry_of[lm][C - 1]);
#pragma unroll
        for (int j = 0; j < MAX_VEC_PER_LANE; j++) ids[j] = __ldg(ent4 + lane + 32 * j);  // always inside the 1920-byte map; unconditional so nothing waits on it here
    }
    // index -> pulse vector (cwrsi, pvc.rs:182-284; only nonzero pulses are stored).  The parts are sorted by
    // size and dealt 32 at a time ("slots", tables: opn_kernels.cu), one part per lane; every lane walks its part
    // in events (see below), the warp leaves a slot when its slowest lane is done.
    mbar_wait(bar, 0);
    {
        const uint32_t *U = s_pvq;
        const uint2 *CW = s_cw;
        const uint16_t *row = s_row;
        const int nslots = g_tab.synth_n_slots[lm][C - 1];
#pragma unroll 1
        for (int slot = 0; slot < nslots; slot++) {
            const uint32_t e = g_tab.synth_slot_entries[lm][C - 1][slot][lane];
            const bool has = e != 0xFFu;
            const SynthEntry E = s_ent[has ? e : 0u];
            uint32_t n = has ? E.n : 0u, k = E.k, i = has ? s_idx[e] : 0u;
            int16_t *y = s_y + E.base;
            int32_t yy = 0;
            if (has && n == 1u) {  // sign-only band
                *y = i ? (int16_t)-1 : (int16_t)1;
                yy = 1;
            } else if (has && k == 1u) {
                // one pulse: cwrsi reduces to a closed form, y[i] = +1 for i < n, y[2n-1-i] = -1 otherwise
                // (checked against the oracle for every band size in tests/test_oracle_kat.py::test_cwrsi_single_pulse_closed_form)
                if (i < n) y[i] = (int16_t)1;
                else y[2u * n - 1u - i] = (int16_t)-1;
                yy = 1;
                n = 0u;  // done: takes no part in the walk below
            }
            // A lane walks its part in EVENTS, not dimensions.  While k < n (pvc.rs:232-258) a dimension is
            // empty iff U(k,n) <= i < U(k+1,n), and stepping over it subtracts U(k,n); after t empty dimensions
            // i has lost A(t) = C(k,n) - C(k,n-t) (C = running row sum of U), so "dimension n-t is empty given
            // all before it were" reads A(t) + U(k,n-t) <= i < A(t) + U(k+1,n-t), or as one unsigned comparison
            // i - C(k,n) + C(k,n-t-1) < U(k+1,n-t) - U(k,n-t) (no sum wraps for the parts of the schedule: checked
            // when the tables are built).  That predicate is monotone in t
            // (once a dimension is occupied the sequential loop stops there), so the length of the run of empty
            // dimensions is found by bisection over t in [0, T], T = n - max(k,2) being where the regime ends
            // (k >= n, or the closed-form tail n == 2).  One event = one run + the occupied dimension after it:
            // a part with k pulses takes at most k+1 events instead of n-2 lockstep steps.
            uint32_t rk = row[min(k, 14u)], rk1 = row[min(k + 1u, 14u)];  // row offsets of U(k,.) and U(k+1,.): only read when k < n
#pragma unroll 1
            while (n > 2u) {
                if (k >= n) {  // lots of pulses, pvc.rs:196-231: one dimension per event
                    const uint32_t rn = row[n];
                    uint32_t p = U[rn + k + 1u];
                    const int32_t sg = i >= p ? -1 : 0;
                    i -= (uint32_t)((int32_t)p & sg);
                    const uint32_t k0 = k;
                    const uint32_t q = U[rn + n];
                    if (q > i) {
                        k = n;
                        do {
                            k -= 1u;
                            p = U[row[k] + n];
                        } while (p > i);
                    } else {
                        p = U[rn + k];
                        while (p > i) {
                            k -= 1u;
                            p = U[rn + k];
                        }
                    }
                    i -= p;
                    const int32_t val = ((int32_t)k0 - (int32_t)k + sg) ^ sg;
                    *y = (int16_t)val;
                    yy += val * val;
                    rk = row[min(k, 14u)];
                    rk1 = row[min(k + 1u, 14u)];
                    y++;
                    n -= 1u;
                } else {  // lots of dimensions, pvc.rs:232-258
                    const uint32_t T = n - max(k, 2u);
                    const uint2 *pw = CW + rk + n;              // pw[-t] = (C(k,n-t-1), V(n-t-1,k))
                    const uint32_t ic = i - (pw[0].x + U[rk + n]);  // i - C(k,n)
                    uint32_t lo = 0u, hi = T;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        const uint2 cw = pw[-(int32_t)mid];
                        // i - A(mid) - U(k,n-mid) as one unsigned number: below V(n-mid-1,k) iff dimension n-mid is empty
                        const bool empty = ic + cw.x < cw.y;
                        lo = empty ? mid + 1u : lo;
                        hi = empty ? hi : mid;
                    }
                    if (lo) i = ic + pw[1 - (int32_t)lo].x;  // i - A(lo)
                    y += lo;
                    n -= lo;
                    if (lo < T) {  // dimension n holds pulses
                        const uint32_t q = U[rk1 + n];
                        const int32_t sg = i >= q ? -1 : 0;
                        i -= (uint32_t)((int32_t)q & sg);
                        const uint32_t k0 = k;
                        uint32_t p;
                        do {
                            k -= 1u;
                            p = U[row[k] + n];
                        } while (p > i);
                        i -= p;
                        const int32_t val = ((int32_t)k0 - (int32_t)k + sg) ^ sg;
                        *y = (int16_t)val;
                        yy += val * val;
                        rk = row[k];
                        rk1 = row[k + 1u];
                        y++;
                        n -= 1u;
                    }
                }
            }
            if (n == 2u) {
                // n == 2 (pvc.rs:262-275)
                uint32_t p = 2u * k + 1u;
                int32_t sg = i >= p ? -1 : 0;
                i -= (uint32_t)((int32_t)p & sg);
                const uint32_t k0 = k;
                k = (i + 1u) >> 1;
                if (k != 0u) i -= 2u * k - 1u;
                int32_t val = ((int32_t)k0 - (int32_t)k + sg) ^ sg;
                y[0] = (int16_t)val;
                yy += val * val;
                // n == 1 (pvc.rs:277-281)
                sg = -(int32_t)i;
                val = ((int32_t)k + sg) ^ sg;
                y[1] = (int16_t)val;
                yy += val * val;
            }
            if (has) s_gain[e] = 0.03125f / sqrtf((float)yy);
        }
        if (lane == 0) s_gain[SYNTH_MAX_ENTRIES] = 0.0f;
    }
    __syncwarp();
    // pulses x gain -> coefficient rows, 4 bins per lane and step (entry_of: bin -> part)
    {
        const uint2 *y4 = reinterpret_cast<const uint2 *>(s_y);
#pragma unroll
        for (int j = 0; j < MAX_VEC_PER_LANE; j++) {
            const int i = (int)lane + 32 * j;
            if (i >= nvec) break;
            const uint32_t idw = ids[j];
            const uint2 yr = y4[i];
            int4 yv;
            float4 cv;
            const uint32_t e0 = idw & 0xFFu, e1 = (idw >> 8) & 0xFFu, e2 = (idw >> 16) & 0xFFu, e3 = idw >> 24;
            yv.x = (int32_t)(int16_t)(yr.x & 0xFFFFu);  // bins without a part were zeroed above
            yv.y = (int32_t)(int16_t)(yr.x >> 16);
            yv.z = (int32_t)(int16_t)(yr.y & 0xFFFFu);
            yv.w = (int32_t)(int16_t)(yr.y >> 16);
            cv.x = (float)yv.x * s_gain[e0];  // bins without a part: pulse 0 x gain slot 72 (= 0)
            cv.y = (float)yv.y * s_gain[e1];
            cv.z = (float)yv.z * s_gain[e2];
            cv.w = (float)yv.w * s_gain[e3];
            if (coef4) coef4[i] = cv;
            if (yo4) yo4[i] = yv;
        }
    }
}

}  // namespace opn
