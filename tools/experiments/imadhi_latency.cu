// Latency and throughput of IMAD.HI (32x32 -> high word, with addend) against IMAD on sm_100a: one warp, dependent chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/imadhi tools/experiments/imadhi_latency.cu && /tmp/imadhi
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out, long long *clk, int a0, int b0)
{
    int x = a0 + threadIdx.x, b = b0, c = 1;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = __mulhi(x, b) + c;  // dependent through the multiplicand
    }
    long long t1 = clock64();
    int y = a0 + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 1000; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) y = __mulhi(b, c + j) + y;  // dependent through the addend only
    }
    long long t2 = clock64();
    int z = a0 + threadIdx.x;
#pragma unroll 1
    for (int i = 0; i < 1000; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) z = z * b + c;  // IMAD chain
    }
    long long t3 = clock64();
    int w0 = a0, w1 = a0 + 1, w2 = a0 + 2, w3 = a0 + 3, w4 = a0 + 4, w5 = a0 + 5, w6 = a0 + 6, w7 = a0 + 7;
#pragma unroll 1
    for (int i = 0; i < 1000; i++) {  // eight independent multiplicand chains: throughput
#pragma unroll
        for (int j = 0; j < 2; j++) {
            w0 = __mulhi(w0, b) + c; w1 = __mulhi(w1, b) + c; w2 = __mulhi(w2, b) + c; w3 = __mulhi(w3, b) + c;
            w4 = __mulhi(w4, b) + c; w5 = __mulhi(w5, b) + c; w6 = __mulhi(w6, b) + c; w7 = __mulhi(w7, b) + c;
        }
    }
    long long t4 = clock64();
    if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
    out[threadIdx.x] = x + y + z + w0 + w1 + w2 + w3 + w4 + w5 + w6 + w7;
}
int main()
{
    int *o; long long *c, h[4];
    cudaMalloc(&o, 128); cudaMalloc(&c, 32);
    k<<<1, 32>>>(o, c, 123456789, 0x7f00ff01);
    k<<<1, 32>>>(o, c, 123456789, 0x7f00ff01);
    cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    printf("IMAD.HI chain through multiplicand: %.1f cycles/op\n", h[0] / 16000.0);
    printf("IMAD.HI chain through addend:       %.1f cycles/op\n", h[1] / 16000.0);
    printf("IMAD chain:                         %.1f cycles/op\n", h[2] / 16000.0);
    printf("IMAD.HI, 8 independent chains:      %.1f cycles/op\n", h[3] / 16000.0);
    return 0;
}
