// Micro-benchmark: issue cost of instruction mixes on one SMSP (1 or 2 warps), sm_100a.
// Each variant runs an unrolled body of independent chains; reports cycles per body instruction.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
template <int MODE>
__global__ void kern(float *out, double *outd, int iters, unsigned long long *cyc) {
    float f[CHAINS]; double d[CHAINS]; unsigned u[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { f[i] = threadIdx.x * 0.001f + i; d[i] = threadIdx.x * 0.001 + i; u[i] = threadIdx.x + i; }
    const float a = 1.0001f, b = 0.5f; const double da = 1.0001, db = 0.5;
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (MODE == 0) { f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); }
                if (MODE == 1) { f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); d[i] = fma(d[i], da, db); }
                if (MODE == 2) { f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); f[i] = fmaf(f[i], a, b); u[i] = (u[i] ^ 0x5bd1e995u) + (u[i] >> 3); }  // 3 FFMA + ~3 ALU
                if (MODE == 3) { f[i] = __fadd_rn(f[i], b); f[i] = __fmul_rn(f[i], a); f[i] = __fadd_rn(f[i], b); f[i] = __fmul_rn(f[i], a); f[i] = __fadd_rn(f[i], b); }
                if (MODE == 4) { d[i] = fma(d[i], da, db); d[i] = fma(d[i], da, db); }
                if (MODE == 5) { f[i] = fmaxf(f[i], b) ; f[i] = fminf(f[i], 1e30f); u[i] = (u[i] & 0xffff) | (u[i] << 1); }  // ALU only
            }
        }
    }
    unsigned long long t1 = clock64();
    float s = 0; double sd = 0; unsigned su = 0;
    for (int i = 0; i < CHAINS; ++i) { s += f[i]; sd += d[i]; su += u[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + su; outd[blockIdx.x * blockDim.x + threadIdx.x] = sd;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, int per_iter, int threads) {
    float *o; double *od; unsigned long long *c, h;
    cudaMalloc(&o, 4096 * 4); cudaMalloc(&od, 4096 * 8); cudaMalloc(&c, 8);
    const int iters = 20000;
    kern<MODE><<<1, threads>>>(o, od, 100, c);
    kern<MODE><<<1, threads>>>(o, od, iters, c);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    // threads = 128*w: w warps on each of the 4 SMSPs
    const double instr_per_warp = (double)iters * per_iter;
    printf("%-28s warps/SMSP=%d  cycles/instr(per warp)=%.3f  SMSP IPC=%.3f\n", name, threads / 128,
           h / instr_per_warp, instr_per_warp * (threads / 128) / h);
    cudaFree(o); cudaFree(od); cudaFree(c);
}

int main() {
    for (int w = 1; w <= 2; ++w) {
        run<0>("FFMA only", 4 * CHAINS * 5, 128 * w);
        run<3>("FADD/FMUL only", 4 * CHAINS * 5, 128 * w);
        run<1>("4 FFMA : 1 DFMA", 4 * CHAINS * 5, 128 * w);
        run<4>("DFMA only", 4 * CHAINS * 2, 128 * w);
        run<2>("3 FFMA : ~3 ALU", 4 * CHAINS * 6, 128 * w);
        run<5>("ALU only (~4)", 4 * CHAINS * 4, 128 * w);
    }
    return 0;
}
