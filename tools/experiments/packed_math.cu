// packed_math.cu -- does sm_100a's FADD2/FMUL2 (add/mul.rn.f32x2) raise FP32 throughput per issue slot?
// Each lane runs CH independent chains of float2 add/mul; scalar version = 2 instructions per float2 op,
// packed version = 1.  Results are bit-identical (both round each element to nearest once).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o packed_math packed_math.cu
#include <cstdio>
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
template <int PACKED, int CH> __global__ void __launch_bounds__(32) k(float2 *p, int iters)
{
    float2 a[CH], w = p[1000 + threadIdx.x];
#pragma unroll
    for (int c = 0; c < CH; c++) a[c] = p[blockIdx.x * 32 * CH + c * 32 + threadIdx.x];
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (PACKED) {
                a[c] = add2(mul2(a[c], w), w);
            } else {
                a[c].x = a[c].x * w.x + w.x;  // -fmad=false: FMUL + FADD
                a[c].y = a[c].y * w.y + w.y;
            }
        }
    }
    float2 s = a[0];
#pragma unroll
    for (int c = 1; c < CH; c++) { s.x += a[c].x; s.y += a[c].y; }
    p[blockIdx.x * 32 + threadIdx.x] = s;
}
template <int PACKED, int CH> float run(float2 *d, int grid, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PACKED, CH><<<grid, 32>>>(d, iters);
    cudaEventRecord(e0);
    k<PACKED, CH><<<grid, 32>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main()
{
    float2 *d; cudaMalloc(&d, 64 << 20); cudaMemset(d, 0, 64 << 20);
    const int iters = 4096;
    for (int wps : {4, 8, 16, 20, 32}) {
        const int grid = 148 * wps;
        const float s = run<0, 8>(d, grid, iters), q = run<1, 8>(d, grid, iters);
        const double ops = (double)grid * 32 * 8 * iters * 4;  // flops
        printf("warps/SM %2d: scalar %.3f ms (%.1f TFLOP/s)  packed %.3f ms (%.1f TFLOP/s)  ratio %.2f\n", wps, s, ops / s * 1e-9,
               q, ops / q * 1e-9, s / q);
    }
    float2 h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
