#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float2 *o, int iters, unsigned long long *cyc) {
    unsigned long long a[8]; unsigned long long b, c;
    float2 fb = make_float2(1.0001f, 0.9999f), fc = make_float2(0.5f, 0.25f);
    b = *reinterpret_cast<unsigned long long*>(&fb); c = *reinterpret_cast<unsigned long long*>(&fc);
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x * 0.001f + i, i + 1.f); a[i] = *reinterpret_cast<unsigned long long*>(&v); }
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    }
    unsigned long long t1 = clock64();
    float2 s = make_float2(0, 0);
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&a[i]); s.x += v.x; s.y += v.y; }
    o[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float2 *o; unsigned long long *c, h; cudaMalloc(&o, 4096 * 8); cudaMalloc(&c, 8);
    for (int w = 1; w <= 2; ++w) {
        k<<<1, 128 * w>>>(o, 100, c); k<<<1, 128 * w>>>(o, 4000, c);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("FFMA2 warps/SMSP=%d cycles per instr = %.2f\n", w, h / (4000.0 * 32 * w));
    }
}
