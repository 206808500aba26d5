// K6, tensor-core path for the reference's default network shape (model.py:52-75 defaults): two Conv1d
// layers (k = 3, padding 1) C0 -> 8 -> 16 with SiLU, Linear(16 W -> out <= 4).
//
// conv2 carries 78 % of the MACs and is a dense contraction per window: D[t][oc] = sum_k A[t][k] B[k][oc]
// with t = 256 positions, k = ic * 3 + tap = 24, oc = 16.  It runs as warp-level mma.sync m16n8k8 TF32
// instructions with the **3xTF32 split** (a = a_hi + a_lo, b = b_hi + b_lo; a_lo b_hi + a_hi b_lo + a_hi b_hi
// accumulated in FP32), which keeps the result at float32 accuracy (the dropped a_lo b_lo term is 2^-22
// relative) -- the parity test against the torch float32 modules keeps its 2e-4 bound.  A fragments are
// gathered straight from the conv1 activation rows in shared memory (im2col on the fly: 4 LDS per fragment),
// the B fragments and biases of a lane live in registers for the whole kernel, and the D fragments go
// through SiLU into the Linear layer without touching memory (fc weights in shared memory with a 260-word
// channel stride, conflict-free for the D-fragment layout).  conv1 (C0 * 3 = 9-deep, 19 % of the MACs) stays
// on the FP32 pipe with the register tile of the generic kernel.
//
// Why mma.sync and not tcgen05: one window's contraction is 256 x 16 x 24 -- a single tcgen05.mma
// (M = 128, N = 16, K = 8) would be fed by descriptors over an im2col tile that does not exist in memory,
// and the TMEM round trip per 100 kFLOP costs more than the math; the warp-level instruction reads the
// activations where conv1 left them.
#pragma once

namespace ofp {

#ifndef OFP_K6T_WARPS
#define OFP_K6T_WARPS 16
#endif
constexpr int K6T_C1 = 8, K6T_C2 = 16, K6T_FCS = 260;  // channels, fc channel stride in shared memory
#ifndef OFP_K6T_MINCTA
#define OFP_K6T_MINCTA 1
#endif
constexpr int K6T_WARPS = OFP_K6T_WARPS;            // warps per CTA of the tensor-core kernel

__device__ __forceinline__ uint32_t tf32_hi(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void tf32_split(float v, uint32_t &hi, uint32_t &lo) {
    hi = tf32_hi(v);
    lo = tf32_hi(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// shared memory: [conv1 wT + bias | fc weights [out][16][260] | per warp: input rows (C0 x RS), h1 rows (8 x RS)]
template <int P>
__global__ void __launch_bounds__(K6T_WARPS * 32, OFP_K6T_MINCTA) k6_cnn_tc(const K6Args a) {
    extern __shared__ __align__(16) float k6_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int RS = a.row_stride, W = a.W, C0 = a.C0;
    float *w1 = k6_smem;                                   // conv1: wT[ic][k][8] then bias[8]
    const int n_w1 = C0 * 3 * 8 + 8;
    float *fcs = w1 + ((n_w1 + 3) & ~3);                    // [out][16][260]
    float *bufs = fcs + a.out_size * K6T_C2 * K6T_FCS;
    float *inb = bufs + static_cast<size_t>(warp) * (C0 + K6T_C1) * RS;
    float *h1 = inb + C0 * RS;
    for (int i = tid; i < n_w1; i += NW * 32) w1[i] = a.params[i];
    for (int i = tid; i < a.out_size * K6T_C2 * K6T_FCS; i += NW * 32) {
        const int q = i / (K6T_C2 * K6T_FCS), r = i - q * (K6T_C2 * K6T_FCS), oc = r / K6T_FCS, t = r - oc * K6T_FCS;
        fcs[i] = t < W ? a.params[a.fc_w_off + (q * K6T_C2 + oc) * W + t] : 0.f;
    }
    for (int i = lane; i < (C0 + K6T_C1) * RS; i += 32) inb[i] = 0.f;  // halos stay zero
    // ---- per-lane constants of conv2: B fragments (hi / lo), biases, A gather offsets ----
    const int g = lane >> 2, tg = lane & 3;
    const float *w2 = a.params + a.w_off[1], *b2 = a.params + a.b_off[1];
    uint32_t bh[3][2][2], bl[3][2][2];
    int offA[3][2];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int k = 8 * ks + tg + 4 * j;      // k = ic * 3 + tap
            offA[ks][j] = (k / 3) * RS + (k % 3);   // + position: h1 row ic, padded index t + tap
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) tf32_split(__ldg(w2 + k * K6T_C2 + 8 * nt + g), bh[ks][nt][j], bl[ks][nt][j]);
        }
    }
    float bias2[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) { bias2[nt][0] = __ldg(b2 + 8 * nt + 2 * tg); bias2[nt][1] = __ldg(b2 + 8 * nt + 2 * tg + 1); }
    const float *fcb = a.params + a.fc_b_off;
    __syncthreads();

    for (int64_t wi = static_cast<int64_t>(blockIdx.x) * NW + warp; wi < a.n; wi += static_cast<int64_t>(gridDim.x) * NW) {
        const float *xw = a.x + wi * a.win_stride;
        for (int c = 0; c < C0; ++c)
            for (int t = lane; t < W; t += 32) inb[c * RS + 1 + t] = __ldg(xw + c * W + t);
        __syncwarp();
        // ---- conv1 + SiLU on the FP32 pipe: two passes of 4 channels x P positions per lane (32 accumulators:
        // the kernel then fits 168 registers = 12 warps per SM instead of 8) ----
#pragma unroll 1
        for (int ob = 0; ob < 8; ob += 4) {
            float acc[4][P];
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float b = w1[C0 * 24 + ob + o];
#pragma unroll
                for (int p = 0; p < P; ++p) acc[o][p] = b;
            }
            for (int ic = 0; ic < C0; ++ic) {
                const float *row = inb + ic * RS + lane;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float4 wa = *reinterpret_cast<const float4 *>(w1 + (ic * 3 + k) * 8 + ob);
                    const float wv[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const float xin = row[32 * p + k];
#pragma unroll
                        for (int o = 0; o < 4; ++o) acc[o][p] = fmaf(wv[o], xin, acc[o][p]);
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (lane + 32 * p < W) h1[(ob + o) * RS + 1 + lane + 32 * p] = k6_act<0>(acc[o][p]);
        }
        __syncwarp();
        // ---- conv2 on the tensor cores, SiLU, Linear ----
        float fcacc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int mt = 0; mt < W / 16; ++mt) {
            const int t0 = 16 * mt + g;
            // three independent accumulator sets (one per split term) keep the HMMA chains 3 deep
            float d[2][4], dl[2][4], dm[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                d[nt][0] = d[nt][2] = bias2[nt][0]; d[nt][1] = d[nt][3] = bias2[nt][1];
#pragma unroll
                for (int i = 0; i < 4; ++i) dl[nt][i] = dm[nt][i] = 0.f;
            }
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                uint32_t ah[4], al[4];
                tf32_split(h1[offA[ks][0] + t0], ah[0], al[0]);
                tf32_split(h1[offA[ks][0] + t0 + 8], ah[1], al[1]);
                tf32_split(h1[offA[ks][1] + t0], ah[2], al[2]);
                tf32_split(h1[offA[ks][1] + t0 + 8], ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    mma_tf32(dl[nt], al, bh[ks][nt]);
                    mma_tf32(dm[nt], ah, bl[ks][nt]);
                    mma_tf32(d[nt], ah, bh[ks][nt]);
                }
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) d[nt][i] += dl[nt][i] + dm[nt][i];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int oc = 8 * nt + 2 * tg + (i & 1), t = t0 + 8 * (i >> 1);
                    const float h = k6_act<0>(d[nt][i]);
                    const float *fw = fcs + oc * K6T_FCS + t;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (q < a.out_size) fcacc[q] = fmaf(fw[q * (K6T_C2 * K6T_FCS)], h, fcacc[q]);
                }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = fcacc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && q < a.out_size) a.out[wi * a.out_size + q] = v + __ldg(fcb + q);
        }
        __syncwarp();
    }
}

}  // namespace ofp
