// ptxas_contraction_probe.cu -- what ptxas 12.9 does with packed FP32 (sm_100a), checked in SASS, no GPU needed:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -c ptxas_contraction_probe.cu -o /tmp/p.o && cuobjdump -sass /tmp/p.o | grep -E "Function|FADD|FMUL|FFMA"
// Findings ("Packed sums"):
//   k_packed_mul_add : mul.rn.f32x2 + add.rn.f32x2  ->  ONE FFMA2 (contracted although both carry .rn and --fmad=false is passed)
//   k_scalar_mul_add : mul.rn.f32   + add.rn.f32    ->  FMUL, FADD (never contracted)
//   k_scalar_mul_packed_add : scalar products feeding a packed sum -> FMUL, FMUL, FADD2 (not contracted): what the kernels use
//   k_swizzle : p_add(a, (s.y, -s.x)) -> ONE FADD2 with operand modifiers R.F32x2.LO_HI.NP (swap + per-half negation are free)
#include <cuda_runtime.h>
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float adds(float a, float b) { float r; asm("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float muls(float a, float b) { float r; asm("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__global__ void k_packed_mul_add(float2 *p) { float2 a = p[threadIdx.x], w = p[64]; p[threadIdx.x] = add2(mul2(a, w), w); }
__global__ void k_scalar_mul_add(float2 *p) { float2 a = p[threadIdx.x], w = p[64]; p[threadIdx.x] = make_float2(adds(muls(a.x, w.x), w.x), adds(muls(a.y, w.y), w.y)); }
__global__ void k_scalar_mul_packed_add(float2 *p, float g) { float2 y = p[threadIdx.x], x = p[threadIdx.x + 32]; p[threadIdx.x] = add2(y, make_float2(g * x.x, g * x.y)); }
__global__ void k_swizzle(float2 *p) { float2 a = p[threadIdx.x], s = p[threadIdx.x + 32]; p[threadIdx.x] = add2(a, make_float2(s.y, -s.x)); }
