// K6b: model.CCCNN.forward in eval mode (reference model.py:443-538, group = False): the conv stack is applied
// to every sensor channel separately (shared weights, in_channels = 1), each of the K feature maps of a channel is
// correlated with ITSELF over all 2V-1 lags (F.conv1d(inputs, filters, groups = B C K, padding = V-1)), the K
// auto-correlations are summed, soft-maxed over the lags, and the C x (2V-1) probabilities feed a Linear layer.
//
// The auto-correlation sum is a dense contraction in disguise: with F [K, V] the channel's feature maps,
// G = F^T F (V x V x K) holds every product sum_k F[k][i] F[k][j], and cc[lag] is the sum of G's lag-th diagonal.
// G is formed on the tensor cores (mma.sync m16n8k8, 3xTF32 as in cnn_infer_tc.cuh) -- and the diagonal sums
// come for free: all 16 x 8 tiles with the same offset j0 - i0 accumulate into ONE fragment, in which a lane's
// four elements keep a fixed lag, so the fragment is scattered into the lag bins once per tile diagonal
// (shared-memory atomics), not once per tile.  Only tile diagonals with j0 >= i0 are visited (they contain every
// pair with lag >= 0; elements with a negative lag are masked) and cc[-lag] = cc[lag].
// The feature maps are split into their TF32 hi / lo parts once (every element is an operand of ~16 tiles), and
// a diagonal's fragment goes through a 16 x 8 staging tile so that each lag bin has exactly one owner lane (no
// atomics).  One warp per window, channels one after the other; conv layers on the FP32 pipe as in k6_cnn.
#pragma once

namespace ofp {

template <int KS, int P, int ND>
__global__ void __launch_bounds__(3 * 32, 2) k6_cccnn(const K6Args a, const int n_ch, const int rows_a, const int rows_b) {
    extern __shared__ __align__(16) float k6_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int RS = a.row_stride, pad = a.pad;
    const int Kp = a.coutp[a.n_layers - 1], V = a.lout[a.n_layers - 1], nb = 2 * V - 1;
    float *prm = k6_smem;                                     // conv params
    const float *fcs = a.params + a.fc_w_off;                 // fc weights [out][n_ch * nb], read through L1
    float *bufs = prm + ((a.conv_params + 3) & ~3);
    const int per_warp = (rows_a + rows_b) * RS + ((V + 3) & ~3) + 128;
    float *bufA = bufs + static_cast<size_t>(warp) * per_warp;
    float *bufB = bufA + rows_a * RS;
    float *cc = bufB + rows_b * RS;                           // [V] lag bins (lag >= 0)
    float *tile = cc + ((V + 3) & ~3);                        // [16][8] staging of one diagonal's fragment
    for (int i = tid; i < a.conv_params; i += NW * 32) prm[i] = a.params[i];
    for (int i = lane; i < per_warp; i += 32) bufA[i] = 0.f;
    __syncthreads();
    const float *fcb = a.params + a.fc_b_off;
    const int g = lane >> 2, tg = lane & 3;

    for (int64_t wi = static_cast<int64_t>(blockIdx.x) * NW + warp; wi < a.n; wi += static_cast<int64_t>(gridDim.x) * NW) {
        float fcacc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < n_ch; ++c) {
            // ---- stage this channel's window as the single input row ----
            const float *xw = a.x + wi * a.win_stride + static_cast<int64_t>(c) * a.W;
            for (int t = lane; t < a.W; t += 32) bufA[pad + t] = __ldg(xw + t);
            for (int t = pad + a.W + lane; t < RS; t += 32) bufA[t] = 0.f;
            __syncwarp();
            float *in = bufA, *outb = bufB;
            // ---- conv stack, every layer stored ----
            for (int l = 0; l < a.n_layers; ++l) {
                const int Cin = a.cin[l], Cout = a.cout[l], CP = a.coutp[l], Lout = a.lout[l];
                // group = True (model.py:512-538: groups = channels): channel c runs its OWN copy of the stack
                const float *wT = prm + c * a.group_stride + a.w_off[l], *bias = prm + c * a.group_stride + a.b_off[l];
                for (int ob = 0; ob < CP; ob += 8) {
                    float acc[8][P];
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const float b = bias[ob + o];
#pragma unroll
                        for (int p = 0; p < P; ++p) acc[o][p] = b;
                    }
                    for (int ic = 0; ic < Cin; ++ic) {
                        const float *row = in + ic * RS + lane;
#pragma unroll
                        for (int k = 0; k < KS; ++k) {
                            const float4 w0 = *reinterpret_cast<const float4 *>(wT + (ic * KS + k) * CP + ob);
                            const float4 w1 = *reinterpret_cast<const float4 *>(wT + (ic * KS + k) * CP + ob + 4);
                            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                const float xin = row[32 * p + k];
#pragma unroll
                                for (int o = 0; o < 8; ++o) acc[o][p] = fmaf(wv[o], xin, acc[o][p]);
                            }
                        }
                    }
                    switch (a.act) {  // uniform
                        case 0: k6_activate<0, P>(acc); break;
                        case 1: k6_activate<1, P>(acc); break;
                        case 2: k6_activate<2, P>(acc); break;
                        default: break;
                    }
                    float dummy[4];
                    k6_epilogue<P, 0>(acc, ob, Cout, Lout, lane, outb, RS, pad, prm, dummy);
                }
                for (int oc = 0; oc < Cout; ++oc)
                    for (int t = pad + Lout + lane; t < RS; t += 32) outb[oc * RS + t] = 0.f;
                __syncwarp();
                float *t2 = in; in = outb; outb = t2;
            }
            // split the maps into TF32 hi (in place) and lo (the other buffer, free now): rows_a, rows_b >= Kp
            float *Fh = in + pad, *Fl = outb + pad;
            for (int k = 0; k < Kp; ++k)
                for (int i = lane; i < V; i += 32) {
                    uint32_t hi, lo;
                    tf32_split(Fh[k * RS + i], hi, lo);
                    Fh[k * RS + i] = __uint_as_float(hi);
                    Fl[k * RS + i] = __uint_as_float(lo);
                }
            for (int i = lane; i < V; i += 32) cc[i] = 0.f;
            __syncwarp();
            // ---- cc[lag] = sum over the lag-th diagonal of F^T F, tile diagonal by tile diagonal ----
            // ND neighbouring tile diagonals per pass: they share the A fragments and give 3 ND independent HMMA
            // chains (the kernel is latency bound at 6 warps per SM)
            for (int d8 = 0; d8 < V / 8; d8 += ND) {
                float acc[ND][3][4];
#pragma unroll
                for (int n = 0; n < ND; ++n)
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[n][t][e] = 0.f;
                for (int mt = 0; 8 * (d8 + 2 * mt) < V; ++mt) {
                    const int i0 = 16 * mt + g, j0 = 8 * (d8 + 2 * mt) + g;
                    for (int ks = 0; ks < Kp; ks += 8) {
                        const int r0 = (ks + tg) * RS, r1 = r0 + 4 * RS;
                        const uint32_t ah[4] = {__float_as_uint(Fh[r0 + i0]), __float_as_uint(Fh[r0 + i0 + 8]),
                                                __float_as_uint(Fh[r1 + i0]), __float_as_uint(Fh[r1 + i0 + 8])};
                        const uint32_t al[4] = {__float_as_uint(Fl[r0 + i0]), __float_as_uint(Fl[r0 + i0 + 8]),
                                                __float_as_uint(Fl[r1 + i0]), __float_as_uint(Fl[r1 + i0 + 8])};
#pragma unroll
                        for (int n = 0; n < ND; ++n) {
                            if (j0 - g + 8 * n < V) {  // diagonal d8 + n still has a tile in this tile row (uniform)
                                const uint32_t bh[2] = {__float_as_uint(Fh[r0 + j0 + 8 * n]), __float_as_uint(Fh[r1 + j0 + 8 * n])};
                                const uint32_t bl[2] = {__float_as_uint(Fl[r0 + j0 + 8 * n]), __float_as_uint(Fl[r1 + j0 + 8 * n])};
                                mma_tf32(acc[n][1], al, bh);
                                mma_tf32(acc[n][2], ah, bl);
                                mma_tf32(acc[n][0], ah, bh);
                            }
                        }
                    }
                }
                // fragment -> 16 x 8 tile (rows g, g + 8; columns 2 tg, 2 tg + 1); lane l then owns the diagonal
                // col - row = l - 15 of the tile, i.e. lag 8 (d8 + n) + l - 15: one owner per bin, no atomics
#pragma unroll
                for (int n = 0; n < ND; ++n) {
                    tile[g * 8 + 2 * tg] = acc[n][0][0] + (acc[n][1][0] + acc[n][2][0]);
                    tile[g * 8 + 2 * tg + 1] = acc[n][0][1] + (acc[n][1][1] + acc[n][2][1]);
                    tile[(g + 8) * 8 + 2 * tg] = acc[n][0][2] + (acc[n][1][2] + acc[n][2][2]);
                    tile[(g + 8) * 8 + 2 * tg + 1] = acc[n][0][3] + (acc[n][1][3] + acc[n][2][3]);
                    __syncwarp();
                    const int off = lane - 15, lagv = 8 * (d8 + n) + off;
                    if (lane < 23 && lagv >= 0 && lagv < V) {
                        float sdiag = 0.f;
                        for (int row = max(0, -off); row < 16 && row + off < 8; ++row) sdiag += tile[row * 8 + row + off];
                        cc[lagv] += sdiag;
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            // ---- softmax over the 2V-1 lags (cc[-lag] = cc[lag]) and this channel's share of the Linear layer ----
            float mx = -INFINITY;
            for (int i = lane; i < V; i += 32) mx = fmaxf(mx, cc[i]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
            for (int i = lane; i < V; i += 32) {
                const float e = __expf(cc[i] - mx);
                cc[i] = e;
                sum += i == 0 ? e : 2.0f * e;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.0f / sum;
            for (int i = lane; i < V; i += 32) {
                const float pr = cc[i] * inv;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < a.out_size) {
                        const float *wq = fcs + (q * n_ch + c) * nb + (V - 1);
                        fcacc[q] = fmaf(pr, i == 0 ? wq[0] : wq[i] + wq[-i], fcacc[q]);
                    }
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v = fcacc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && q < a.out_size) a.out[wi * a.out_size + q] = v + __ldg(fcb + q);
        }
    }
}

}  // namespace ofp

namespace ofp {

// Same network, one CTA (4 warps) per window: the warp-per-window kernel above keeps 33 KB of hi / lo planes per
// WARP, which caps an SM at 6 warps and leaves it latency bound (issue 28 % busy).  Here the four warps of a CTA
// share one set of planes: every thread owns positions t = tid + 128 p of the conv layers, the tile diagonals of the
// Gram matrix are dealt round-robin to the warps (lag bins through shared-memory atomics: neighbouring diagonals of
// different warps overlap in 15 bins), warp 0 does the softmax and the Linear layer.  6 CTAs = 24 warps per SM.
// sum over the 128 threads of a CTA, returned to all of them (scratch: 8 floats of shared memory)
__device__ __forceinline__ float block_sum_128(float v, float *scratch, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // scratch free again
    if ((tid & 31) == 0) scratch[tid >> 5] = v;
    __syncthreads();
    return (scratch[0] + scratch[1]) + (scratch[2] + scratch[3]);
}

#ifndef OFP_K6CC_ND
#define OFP_K6CC_ND 4
#endif
#ifndef OFP_K6CC_MINCTA
#define OFP_K6CC_MINCTA 5
#endif
// GEN = true: the constructor options the reference leaves off by default (model.py:451-457, 485-503) -- per-layer
// kernel sizes and strides, dilation, GroupNorm(1, K) behind every activation (`batch_norm=True` builds a GroupNorm,
// model.py:494-498: statistics over all K x L values of ONE sensor channel's feature maps, biased variance, eps 1e-5)
// and MaxPool1d(2, 2) -- with run-time loop bounds; KS is ignored.  Parameters per layer: wT[cin][ks_l][coutp],
// bias[coutp], then gamma[coutp], beta[coutp] when a.bn.
template <int KS, int PP, int ND, bool GEN>
__global__ void __launch_bounds__(128, OFP_K6CC_MINCTA) k6_cccnn_cta(const K6Args a, const int n_ch, const int rows) {
    extern __shared__ __align__(16) float k6_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RS = a.row_stride, pad = a.pad;
    const int Kp = a.coutp[a.n_layers - 1], V = a.lout[a.n_layers - 1], nb = 2 * V - 1;
    float *prm = k6_smem;
    const float *fcs = a.params + a.fc_w_off;
    float *bufA = prm + ((a.conv_params + 3) & ~3);
    float *bufB = bufA + rows * RS;
    float *cc = bufB + rows * RS;
    float *tile = cc + ((V + 3) & ~3) + warp * 128;
    float *red = cc + ((V + 3) & ~3);  // the staging tiles double as reduction scratch of the norm (GEN)
    for (int i = tid; i < a.conv_params; i += 128) prm[i] = a.params[i];
    for (int i = tid; i < 2 * rows * RS; i += 128) bufA[i] = 0.f;
    __syncthreads();
    const float *fcb = a.params + a.fc_b_off;
    const int g = lane >> 2, tg = lane & 3;

    for (int64_t wi = blockIdx.x; wi < a.n; wi += gridDim.x) {
        float fcacc[4] = {0.f, 0.f, 0.f, 0.f};  // warp 0
        for (int c = 0; c < n_ch; ++c) {
            const float *xw = a.x + wi * a.win_stride + static_cast<int64_t>(c) * a.W;
            for (int t = tid; t < a.W; t += 128) bufA[pad + t] = __ldg(xw + t);
            for (int t = pad + a.W + tid; t < RS; t += 128) bufA[t] = 0.f;
            __syncthreads();
            float *in = bufA, *outb = bufB;
            for (int l = 0; l < a.n_layers; ++l) {
                const int Cin = a.cin[l], Cout = a.cout[l], CP = a.coutp[l], Lout = a.lout[l];
                const int Lc = GEN ? a.lconv[l] : Lout;  // length before pooling
                // group = True (model.py:512-538: groups = channels): channel c runs its OWN copy of the stack
                const float *wT = prm + c * a.group_stride + a.w_off[l], *bias = prm + c * a.group_stride + a.b_off[l];
                for (int ob = 0; ob < CP; ob += 8) {
                    float acc[8][PP];
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const float b = bias[ob + o];
#pragma unroll
                        for (int p = 0; p < PP; ++p) acc[o][p] = b;
                    }
                    if (GEN) {
                        const int ksl = a.ksl[l], strl = a.strl[l], dil = a.dil;
                        for (int ic = 0; ic < Cin; ++ic) {
                            const float *row = in + ic * RS;
                            for (int k = 0; k < ksl; ++k) {
                                const float4 w0 = *reinterpret_cast<const float4 *>(wT + (ic * ksl + k) * CP + ob);
                                const float4 w1 = *reinterpret_cast<const float4 *>(wT + (ic * ksl + k) * CP + ob + 4);
                                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                                for (int p = 0; p < PP; ++p) {
                                    // positions past the layer's output would read past the row: clamp the index (the
                                    // value is never stored)
                                    const float xin = row[min((tid + 128 * p) * strl + k * dil, RS - 1)];
#pragma unroll
                                    for (int o = 0; o < 8; ++o) acc[o][p] = fmaf(wv[o], xin, acc[o][p]);
                                }
                            }
                        }
                    } else
                    for (int ic = 0; ic < Cin; ++ic) {
                        const float *row = in + ic * RS + tid;
#pragma unroll
                        for (int k = 0; k < KS; ++k) {
                            const float4 w0 = *reinterpret_cast<const float4 *>(wT + (ic * KS + k) * CP + ob);
                            const float4 w1 = *reinterpret_cast<const float4 *>(wT + (ic * KS + k) * CP + ob + 4);
                            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                            for (int p = 0; p < PP; ++p) {
                                const float xin = row[128 * p + k];
#pragma unroll
                                for (int o = 0; o < 8; ++o) acc[o][p] = fmaf(wv[o], xin, acc[o][p]);
                            }
                        }
                    }
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const int oc = ob + o;
                        if (oc < Cout) {
#pragma unroll
                            for (int p = 0; p < PP; ++p) {
                                float h = acc[o][p];
                                switch (a.act) {  // uniform
                                    case 0: h = k6_act<0>(h); break;
                                    case 1: h = k6_act<1>(h); break;
                                    case 2: h = k6_act<2>(h); break;
                                    default: break;
                                }
                                if (tid + 128 * p < Lc) outb[oc * RS + pad + tid + 128 * p] = h;
                            }
                        }
                    }
                }
                if (GEN && a.bn) {  // GroupNorm(1, Cout) over the Cout x Lc values just stored
                    __syncthreads();
                    const int nel = Cout * Lc;
                    float s1 = 0.f;
                    for (int e = tid; e < nel; e += 128) s1 += outb[(e / Lc) * RS + pad + e % Lc];
                    const float mean = block_sum_128(s1, red, tid) / nel;
                    float s2 = 0.f;
                    for (int e = tid; e < nel; e += 128) {
                        const float dv = outb[(e / Lc) * RS + pad + e % Lc] - mean;
                        s2 = fmaf(dv, dv, s2);
                    }
                    const float rstd = rsqrtf(block_sum_128(s2, red, tid) / nel + 1e-5f);
                    for (int e = tid; e < nel; e += 128) {
                        const int oc = e / Lc;
                        float *q = outb + oc * RS + pad + e % Lc;
                        *q = fmaf((*q - mean) * rstd, bias[CP + oc], bias[2 * CP + oc]);
                    }
                }
                if (GEN && a.pool) {  // MaxPool1d(2, 2) in place: all reads of a pass before its writes
                    __syncthreads();
                    for (int base = 0; base < Cout * Lout; base += 128) {
                        const int e = base + tid, oc = e / Lout, j = e - oc * Lout;
                        float m = 0.f;
                        if (e < Cout * Lout) m = fmaxf(outb[oc * RS + pad + 2 * j], outb[oc * RS + pad + 2 * j + 1]);
                        __syncthreads();
                        if (e < Cout * Lout) outb[oc * RS + pad + j] = m;
                        __syncthreads();
                    }
                }
                for (int oc = 0; oc < Cout; ++oc)
                    for (int t = pad + Lout + tid; t < RS; t += 128) outb[oc * RS + t] = 0.f;
                __syncthreads();
                float *t2 = in; in = outb; outb = t2;
            }
            float *Fh = in + pad, *Fl = outb + pad;
            for (int e = tid; e < Kp * V; e += 128) {
                const int k = e / V, i = e - k * V;
                uint32_t hi, lo;
                tf32_split(Fh[k * RS + i], hi, lo);
                Fh[k * RS + i] = __uint_as_float(hi);
                Fl[k * RS + i] = __uint_as_float(lo);
            }
            for (int i = tid; i < V; i += 128) cc[i] = 0.f;
            __syncthreads();
            for (int d8 = ND * warp; d8 < V / 8; d8 += 4 * ND) {
                float acc[ND][3][4];
#pragma unroll
                for (int n = 0; n < ND; ++n)
#pragma unroll
                    for (int t = 0; t < 3; ++t)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[n][t][e] = 0.f;
                for (int mt = 0; 8 * (d8 + 2 * mt) < V; ++mt) {
                    const int i0 = 16 * mt + g, j0 = 8 * (d8 + 2 * mt) + g;
                    for (int ks = 0; ks < Kp; ks += 8) {
                        const int r0 = (ks + tg) * RS, r1 = r0 + 4 * RS;
                        const uint32_t ah[4] = {__float_as_uint(Fh[r0 + i0]), __float_as_uint(Fh[r0 + i0 + 8]),
                                                __float_as_uint(Fh[r1 + i0]), __float_as_uint(Fh[r1 + i0 + 8])};
                        const uint32_t al[4] = {__float_as_uint(Fl[r0 + i0]), __float_as_uint(Fl[r0 + i0 + 8]),
                                                __float_as_uint(Fl[r1 + i0]), __float_as_uint(Fl[r1 + i0 + 8])};
#pragma unroll
                        for (int n = 0; n < ND; ++n) {
                            if (j0 - g + 8 * n < V) {
                                const uint32_t bh[2] = {__float_as_uint(Fh[r0 + j0 + 8 * n]), __float_as_uint(Fh[r1 + j0 + 8 * n])};
                                const uint32_t bl[2] = {__float_as_uint(Fl[r0 + j0 + 8 * n]), __float_as_uint(Fl[r1 + j0 + 8 * n])};
                                mma_tf32(acc[n][1], al, bh);
                                mma_tf32(acc[n][2], ah, bl);
                                mma_tf32(acc[n][0], ah, bh);
                            }
                        }
                    }
                }
#pragma unroll
                for (int n = 0; n < ND; ++n) {
                    tile[g * 8 + 2 * tg] = acc[n][0][0] + (acc[n][1][0] + acc[n][2][0]);
                    tile[g * 8 + 2 * tg + 1] = acc[n][0][1] + (acc[n][1][1] + acc[n][2][1]);
                    tile[(g + 8) * 8 + 2 * tg] = acc[n][0][2] + (acc[n][1][2] + acc[n][2][2]);
                    tile[(g + 8) * 8 + 2 * tg + 1] = acc[n][0][3] + (acc[n][1][3] + acc[n][2][3]);
                    __syncwarp();
                    const int off = lane - 15, lagv = 8 * (d8 + n) + off;
                    if (lane < 23 && lagv >= 0 && lagv < V) {
                        float sdiag = 0.f;
                        for (int row = max(0, -off); row < 16 && row + off < 8; ++row) sdiag += tile[row * 8 + row + off];
                        atomicAdd(&cc[lagv], sdiag);
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            if (warp == 0) {
                float mx = -INFINITY;
                for (int i = lane; i < V; i += 32) mx = fmaxf(mx, cc[i]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float sum = 0.f;
                for (int i = lane; i < V; i += 32) {
                    const float e = __expf(cc[i] - mx);
                    cc[i] = e;
                    sum += i == 0 ? e : 2.0f * e;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float inv = 1.0f / sum;
                for (int i = lane; i < V; i += 32) {
                    const float pr = cc[i] * inv;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q < a.out_size) {
                            const float *wq = fcs + (q * n_ch + c) * nb + (V - 1);
                            fcacc[q] = fmaf(pr, i == 0 ? wq[0] : wq[i] + wq[-i], fcacc[q]);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v = fcacc[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && q < a.out_size) a.out[wi * a.out_size + q] = v + __ldg(fcb + q);
            }
        }
    }
}

}  // namespace ofp
