// Host build of the DEVICE source csrc/rangedec.cuh (LaneDec: the branch-free range decoder of k_synth_rangedec), so that
// the CPU test suite can replay range-coder scripts through the very code the kernel runs and compare every symbol,
// tell_frac and rng with the oracle.  Test infrastructure: the CUDA intrinsics the header uses are given host bodies.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#define __device__
#define __forceinline__ inline
#define OPN_HOST_SHIM 1
using std::max;
using std::min;
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __uint2float_rn(uint32_t a) { return (float)a; }
static inline uint32_t __float2uint_rz(float f) { return f <= 0.f ? 0u : (uint32_t)f; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 255u) << (8 * i);
    return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
#include "../../opus-native_b200/csrc/rangedec.cuh"
#include "../../opus-native_b200/csrc/celt2_lane.cuh"
#include "../../opus-native_b200/csrc/opn_tables.h"

using namespace opn;
struct Op { uint32_t op, a, b; };
struct Out { uint32_t value, tell_frac, rng; };
enum { OP_UINT = 0, OP_BITS = 1, OP_BIT_LOGP = 2, OP_ICDF = 3, OP_LAPLACE = 4 };

// `buf` must be readable from the aligned word below buf to the aligned word above buf+len (the kernel's contract).
extern "C" int lanedec_run_script(const uint8_t *buf, uint32_t len, const Op *ops, uint32_t n_ops, const uint8_t *icdf_pool, Out *out)
{
    LaneDec d;
    d.init(buf, len);
    for (uint32_t i = 0; i < n_ops; i++) {
        const uint32_t a = ops[i].a, b = ops[i].b;
        uint32_t v = 0;
        switch (ops[i].op) {
        case OP_UINT: {
            if (b == 1u) { v = d.uint_any(a); break; }  // the generic form (SYNTH-CELT/2: alphabets known only at run time)
            if (a <= 256u) { v = d.uint_small(a); break; }
            // the alphabet split and reciprocal of upload_tables (opn_kernels.cu)
            const uint32_t ftm1 = a - 1u;
            uint32_t ftb = 32u - (uint32_t)__builtin_clz(ftm1);
            uint32_t ft1;
            if (ftb > 8) { ftb -= 8; ft1 = (ftm1 >> ftb) + 1; } else { ftb = 0; ft1 = a; }
            uint32_t sh = 0;
            while ((1ull << sh) < ft1) sh++;
            const uint32_t magic = (uint32_t)((((1ull << sh) - ft1) << 32) / ft1 + 1);
            v = d.uint_precomputed(ftm1, ft1, ftb, magic, sh);
            break;
        }
        case OP_BITS: v = d.bits(a); break;
        case OP_BIT_LOGP: v = d.bit_logp(a); break;
        case OP_ICDF: v = d.icdf(icdf_pool + a, b); break;
        case OP_LAPLACE: {
            uint32_t fl[LAP_N + 1], fs[LAP_N + 1];
            laplace_table(a, b, fl, fs);
            v = (uint32_t)d.laplace(fl, fs, b);
            break;
        }
        default: return -1;
        }
        out[i].value = v;
        out[i].tell_frac = d.tell_frac();
        out[i].rng = d.rng;
    }
    return 0;
}

// SYNTH-CELT/2 frame decode through the device's LaneCoder / LanePartSink (celt2_lane.cuh) and celt2_frame (celt2.cuh).
extern "C" int lanedec_celt2_decode(const uint8_t *payload, uint32_t len, int lm, int channels, Celt2Side *side, Celt2Part *parts)
{
    static uint32_t lfl[21][LAP_N + 1], lfs[21][LAP_N + 1];
    for (uint32_t b = 0; b < 21; b++) {
        const uint32_t decay = 6000u + 400u * b;
        laplace_table(((32768u - 33u) * (16384u - decay)) / (16384u + decay) + 1u, decay, lfl[b], lfs[b]);
    }
    const Celt2Tabs T{OPN_E_BANDS, OPN_LOG_N, OPN_ALLOC_VECTORS, OPN_CACHE_BITS, OPN_CACHE_CAPS, OPN_LOG2_FRAC_TABLE, OPN_CACHE_INDEX, OPN_PVQ_U_DATA,
                      OPN_PVQ_U_ROW};
    std::memset(side, 0, sizeof(*side));
    LaneCoder ec;
    ec.lfl = lfl;
    ec.lfs = lfs;
    ec.d.init(payload, len);
    int16_t e9[2 * 21] = {};
    LanePartSink sink{parts, 0u, 0u, 0u, e9, 1u};
    uint32_t flags = 0, n_pulses = 0;
    celt2_frame(ec, T, len, lm, channels, side, sink, flags, n_pulses);
    side->n_parts = sink.n - sink.nsign;
    side->n_pulses = n_pulses;
    side->final_rng = ec.d.rng;
    side->tell_frac = ec.d.tell_frac();
    return (int)sink.n;
}
