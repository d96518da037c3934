// celt2_lane.cuh -- the decode-side Coder and part sink of celt2_frame (celt2.cuh) on top of LaneDec (rangedec.cuh):
// what one lane of k_celt2_rangedec runs.  Also compiled for the host by tests/host_shim (CUDA qualifiers defined away), so
// that the CPU test suite decodes SYNTH-CELT/2 packets through this very code.
#pragma once
#include "celt2.cuh"
#include "rangedec.cuh"

namespace opn {

struct LaneCoder {  // the Coder of celt2_frame on top of LaneDec
    LaneDec d;
    const uint32_t (*lfl)[LAP_N + 1];
    const uint32_t (*lfs)[LAP_N + 1];
    __device__ __forceinline__ uint32_t tell() const { return d.tell(); }
    __device__ __forceinline__ uint32_t tell_frac() const { return d.tell_frac(); }
    __device__ __forceinline__ uint32_t transient_permille() const { return 0u; }
    __device__ __forceinline__ uint32_t bit_logp(uint32_t logp, uint32_t) { return d.bit_logp(logp); }
    __device__ __forceinline__ uint32_t icdf(const uint8_t *tab, uint32_t ftb, uint32_t) { return d.icdf(tab, ftb); }
    __device__ __forceinline__ uint32_t uint_(uint32_t ft) { return d.uint_any(ft); }
    __device__ __forceinline__ uint32_t pulses_index(uint32_t ft) { return d.uint_any(ft); }
    __device__ __forceinline__ uint32_t bits(uint32_t n) { return d.bits(n); }
    __device__ __forceinline__ int32_t laplace(int band) { return d.laplace(lfl[band], lfs[band], 6000u + 400u * (uint32_t)band); }
    // the split angle, triangular pdf over 0..qn (RFC 6716 4.3.4.2; decode + update, decoder.rs:143-181)
    __device__ __forceinline__ uint32_t theta_tri(uint32_t qn)
    {
        const uint32_t h = qn >> 1, ft = (h + 1u) * (h + 1u);
        const uint32_t fm = d.decode(ft);
        uint32_t itheta, fl, fs;
        if (fm < ((h * (h + 1u)) >> 1)) {
            itheta = (c2_isqrt32(8u * fm + 1u) - 1u) >> 1;
            fs = itheta + 1u;
            fl = (itheta * (itheta + 1u)) >> 1;
        } else {
            itheta = (2u * (qn + 1u) - c2_isqrt32(8u * (ft - fm - 1u) + 1u)) >> 1;
            fs = qn + 1u - itheta;
            fl = ft - (((qn + 1u - itheta) * (qn + 2u - itheta)) >> 1);
        }
        d.update(fl, fl + fs, ft);
        return itheta;
    }
};
struct LanePartSink {
    Celt2Part *parts;
    uint32_t n, np, nsign;
    int16_t *e9;         // band energies of this packet, Q9: e9[(c * 21 + band) * e9_stride] (shared memory, one column per lane)
    uint32_t e9_stride;
    __device__ __forceinline__ void energy_set(int c, int band, int v) { e9[(uint32_t)(c * 21 + band) * e9_stride] = (int16_t)v; }
    __device__ __forceinline__ void energy_add(int c, int band, int v) { e9[(uint32_t)(c * 21 + band) * e9_stride] += (int16_t)v; }
    __device__ __forceinline__ int energy(int c, int band) const { return e9[(uint32_t)(c * 21 + band) * e9_stride]; }
    __device__ __forceinline__ void put_part(int base, int nn, int k, uint32_t index, float gain)
    {
        if (n < (uint32_t)CELT2_MAX_PARTS) parts[n] = Celt2Part{(uint16_t)base, (uint8_t)nn, (uint8_t)k, index, gain};
        n += 1u;
        np += (uint32_t)k;
    }
    __device__ __forceinline__ void put_sign(int base, uint32_t sign)  // a one-bin band: its pulse is counted by celt2_frame itself
    {
        put_part(base, 1, 1, sign, 0.03125f);
        np -= 1u;
        nsign += 1u;
    }
    __device__ __forceinline__ uint32_t pulses() const { return np; }
};

}  // namespace opn
