// Micro-benchmark: reciprocal issue cost of single SASS opcodes and of two-opcode mixes on one SMSP
// (2 warps per SMSP, 8 independent chains per thread), sm_100a.  Drives the K1 cost model in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o op_rates op_rates.cu && ./op_rates
#include <cstdio>
#include <cuda_runtime.h>

#define CH 8
#define REP 4

enum Op { FADD, FFMA, FMNMX, FSETSEL, LOP, SHF, IADD, IMAD, DFMA, DADD, DMUL, F2D, D2F, I2D, D2I, LDS32, LDS64, LDS128,
          STS32, EX2, LG2, MIX_DFMA_FFMA, MIX_DFMA_ALU, MIX_FFMA_ALU, MIX_D2F_FFMA, MIX_F2D_DFMA, MIX_LDS_FFMA, CVT_FDF, CVT_IDI, FSETP_ONLY, FSEL_ONLY, ISETP_SEL, IMADMOV, LDS_DEP, STS_NC, SHFL, VOTE, FMNMX3, FMUL_IMM, NOPS };

template <int OP>
__global__ void kern(float *out, int iters, unsigned long long *cyc) {
    __shared__ double sh[1024];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = i * 0.25;
    __syncthreads();
    float f[CH], g[CH];
    double d[CH];
    unsigned u[CH];
    for (int i = 0; i < CH; ++i) {
        f[i] = threadIdx.x * 0.001f + i + 1.0f; g[i] = 0.5f + i; d[i] = threadIdx.x * 0.001 + i + 1.0; u[i] = threadIdx.x * 8 + i * 264;
    }
    const unsigned pr = threadIdx.x & 1;
    const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(sh));
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
                if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(g[i]));
                if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));
                if (OP == FSETSEL) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %2, %1, p;}" : "+f"(f[i]) : "f"(g[i]), "f"(g[(i + 1) % CH]));
                if (OP == LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
                if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
                if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
                if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
                if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));
                if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));
                if (OP == DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));
                if (OP == F2D) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i]));
                if (OP == D2F) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i]));
                if (OP == I2D) asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[i]) : "r"(u[i]));
                if (OP == D2I) asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(u[i]) : "d"(d[i]));
                if (OP == LDS32) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(f[i]) : "r"(sbase + (u[i] & 0xffc)));
                if (OP == LDS64) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(d[i]) : "r"(sbase + (u[i] & 0xff8)));
                if (OP == LDS128) asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(d[i]), "=d"(d[(i + 1) % CH]) : "r"(sbase + (u[i] & 0xff0)));
                if (OP == STS32) asm volatile("st.shared.f32 [%1], %0;" ::"f"(f[i]), "r"(sbase + (u[i] & 0xffc)));
                if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (OP == LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                if (OP == MIX_DFMA_FFMA) {
                    asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(g[i]));
                }
                if (OP == MIX_DFMA_ALU) {
                    asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
                }
                if (OP == MIX_FFMA_ALU) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(g[i]));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
                }
                if (OP == MIX_D2F_FFMA) {
                    asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(g[i]) : "d"(d[i]));
                    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(g[(i + 1) % CH]));
                }
                if (OP == MIX_F2D_DFMA) {
                    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i]));
                    asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[(i + 3) % CH]) : "d"(d[(i + 1) % CH]));
                }
                if (OP == CVT_FDF) {  // dependent pair: F2F.F64.F32 then F2F.F32.F64
                    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i]));
                    asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i]));
                }
                if (OP == CVT_IDI) {  // dependent pair: I2F.F64.S32 then F2I.S32.F64
                    asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[i]) : "r"(u[i]));
                    asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(u[i]) : "d"(d[i]));
                }
                if (OP == FSETP_ONLY) asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; @p add.u32 %0, %0, 1;}" : "+r"(u[i]) : "f"(f[i]), "f"(g[i]));
                if (OP == FSEL_ONLY) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.f32 %0, %0, %1, p;}" : "+f"(f[i]) : "f"(g[i]), "r"(pr));
                if (OP == ISETP_SEL) asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.b32 %0, %2, %1, p;}" : "+r"(u[i]) : "r"(u[(i + 1) % CH]), "r"(u[(i + 2) % CH]));
                if (OP == IMADMOV) asm volatile("mov.b32 %0, %1;" : "=r"(u[i]) : "r"(u[(i + 1) % CH]));
                if (OP == LDS_DEP) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(u[i]) : "r"(sbase + (u[i] & 0xffc)));
                if (OP == STS_NC) asm volatile("st.shared.f32 [%1], %0;" ::"f"(f[i]), "r"(sbase + 4 * threadIdx.x + 512 * i));
                if (OP == SHFL) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 31, 0xffffffff;" : "+r"(u[i]) : "r"(u[(i + 1) % CH] & 31));
                if (OP == VOTE) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(u[i]));
                if (OP == FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(g[(i + 1) % CH]));
                if (OP == FMUL_IMM) asm volatile("mul.rn.f32 %0, %0, 0f3F800347;" : "+f"(f[i]));
                if (OP == MIX_LDS_FFMA) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(g[i]) : "r"(sbase + (u[i] & 0xffc)));
                    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i]) : "f"(g[(i + 1) % CH]));
                }
            }
        }
    }
    unsigned long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CH; ++i) s += f[i] + g[i] + static_cast<float>(d[i]) + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP> void run(const char *name, int ops_per_slot) {
    float *o; unsigned long long *c, h;
    cudaMalloc(&o, 4096 * 4); cudaMalloc(&c, 8);
    const int iters = 4000;
    for (int w = 1; w <= 2; ++w) {
        kern<OP><<<1, 128 * w>>>(o, 100, c);
        kern<OP><<<1, 128 * w>>>(o, iters, c);
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double slots = (double)iters * REP * CH;  // per warp
        // SMSP cycles per warp-slot (a slot = ops_per_slot instructions): elapsed / (slots * warps on the SMSP)
        printf("%-16s warps/SMSP=%d  SMSP cycles per %d instr = %.2f\n", name, w, ops_per_slot, h / (slots * w));
    }
    cudaFree(o); cudaFree(c);
}

int main() {
    run<FADD>("FADD", 1); run<FFMA>("FFMA", 1); run<FMNMX>("FMNMX", 1); run<FSETSEL>("FSETP+FSEL", 2);
    run<LOP>("LOP3", 1); run<SHF>("SHF", 1); run<IADD>("IADD3", 1); run<IMAD>("IMAD", 1);
    run<DFMA>("DFMA", 1); run<DADD>("DADD", 1); run<DMUL>("DMUL", 1);
    run<F2D>("F2F.F64.F32", 1); run<D2F>("F2F.F32.F64", 1); run<I2D>("I2F.F64.S32", 1); run<D2I>("F2I.S32.F64", 1);
    run<LDS32>("LDS.32", 1); run<LDS64>("LDS.64", 1); run<LDS128>("LDS.128", 1); run<STS32>("STS.32", 1);
    run<EX2>("MUFU.EX2", 1); run<LG2>("MUFU.LG2", 1);
    run<MIX_DFMA_FFMA>("DFMA+FFMA", 2); run<MIX_DFMA_ALU>("DFMA+LOP3", 2); run<MIX_FFMA_ALU>("FFMA+LOP3", 2);
    run<CVT_FDF>("F2D->D2F chain", 2); run<CVT_IDI>("I2D->D2I chain", 2); run<FSETP_ONLY>("FSETP+@P IADD", 2);
    run<FSEL_ONLY>("FSEL", 1); run<ISETP_SEL>("ISETP+SEL", 2); run<IMADMOV>("MOV", 1); run<LDS_DEP>("LDS dep", 1);
    run<STS_NC>("STS no conflict", 1); run<SHFL>("SHFL.IDX", 1); run<VOTE>("ISETP+VOTE", 2); run<FMNMX3>("FMNMX3", 1);
    run<FMUL_IMM>("FMUL imm", 1);
    run<MIX_D2F_FFMA>("D2F+FFMA", 2); run<MIX_F2D_DFMA>("F2D+DFMA", 2); run<MIX_LDS_FFMA>("LDS+FFMA", 2);
    return 0;
}
