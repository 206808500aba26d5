// K6: onset-window classifier / regressor inference (configs[4] tail).
//
// Replaces model.CNN.forward in eval mode (reference model.py:52-120): a stack of
// Conv1d(kernel_size, padding, stride 1, dilation 1, groups 1) + activation layers over a window
// [channels, input_size], flatten (channel-major), Dropout (identity at inference), Linear.
//
// One WARP per window, persistent over the batch; nothing but the window (3 KB for 3 x 256) is read
// from HBM and nothing but the output_size results is written.  A lane owns the positions
// t = lane + 32 p (p < P), so every shared-memory access of a warp is to 32 consecutive words, and keeps
// an 8-output-channel x P accumulator tile in registers; the weights of a layer sit in shared memory
// transposed to [ic][k][oc] so that one broadcast LDS.128 feeds 4 output channels.  The last layer's
// activations never leave registers: the Linear layer is accumulated from them directly against
// fc weights in shared memory and finished with a warp shuffle reduction.  FP32 FMAs in PyTorch's
// accumulation order per output (ic outer, tap inner); tolerance vs torch fp32 is set in the test.
#include "ofp_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace ofp {

constexpr int K6_WARPS = 4;  // warps per CTA (fewer when a network's activations are large)
constexpr int K6_MAX_LAYERS = 6;

struct K6Args {
    const float *x;          // [n, C0, W]
    int64_t n, win_stride;
    int32_t C0, W, n_layers, ks, pad, act, out_size;
    int32_t cin[K6_MAX_LAYERS], cout[K6_MAX_LAYERS], coutp[K6_MAX_LAYERS], lin[K6_MAX_LAYERS], lout[K6_MAX_LAYERS];
    int32_t ksl[K6_MAX_LAYERS], strl[K6_MAX_LAYERS], lconv[K6_MAX_LAYERS];  // k6_cccnn_cta<GEN>: per-layer kernel size,
                                                                            // stride, output length before pooling
    int32_t dil, pool, bn;   // k6_cnn / k6_cccnn_cta<GEN>: dilation (>= 1), MaxPool1d(2, 2) after every layer, eval-mode BatchNorm1d as
                             // scale[coutp] + shift[coutp] behind every layer's bias; lout is the length AFTER pooling
    int32_t w_off[K6_MAX_LAYERS], b_off[K6_MAX_LAYERS];  // float offsets into params
    int32_t fc_w_off, fc_b_off, n_params, conv_params;   // conv_params: floats staged in shared memory always
    int32_t fc_in_smem;
    int32_t group_stride;    // CCCNN(group=True): floats between the conv parameter blocks of two channels (0 = shared stack)
    int32_t row_stride;      // floats per activation row (>= 32 P + ks, multiple of 4... plus halo)
    int32_t buf_rows;        // rows per activation buffer
    const float *params;     // packed: per layer wT[ic][k][coutp] + bias[coutp]; fc w[out][flat] + bias[out]
    float *out;              // [n, out_size]
};

template <int ACT>
__device__ __forceinline__ float k6_act(float v) {
    if (ACT == 0) {  // SiLU = v / (1 + 2^(-v log2 e)): ex2.approx + rcp.approx, 5 instructions (rel. error ~2e-7)
        float e, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
        return v * r;
    }
    if (ACT == 1) return fmaxf(v, 0.0f);                    // ReLU
    if (ACT == 2) return tanhf(v);
    return v;
}

template <int ACT, int P>
__device__ __forceinline__ void k6_activate(float (&acc)[8][P]) {
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
        for (int p = 0; p < P; ++p) acc[o][p] = k6_act<ACT>(acc[o][p]);
}

// Epilogue of one activated 8-channel x P tile: the next layer's input rows (OUT = 0) or the Linear layer
// with OUT <= 4 outputs accumulated straight from the registers against fc weights in shared memory.
// MaxPool1d(2, 2) (model.py:104-107) of one activated tile into the next layer's input rows: neighbouring positions
// sit in neighbouring lanes, the even lane of a pair writes max(y[2j], y[2j+1]) to position j < Lout = floor(L / 2).
template <int P>
__device__ __forceinline__ void k6_epilogue_pool(const float (&acc)[8][P], int ob, int Cout, int Lout, int lane,
                                                 float *outb, int RS, int pad) {
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        if (ob + o < Cout) {  // uniform
            float *orow = outb + (ob + o) * RS + pad;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float m = fmaxf(acc[o][p], __shfl_xor_sync(0xffffffffu, acc[o][p], 1));
                const int j = (lane + 32 * p) >> 1;
                if (!(lane & 1) && j < Lout) orow[j] = m;
            }
        }
    }
}

template <int P, int OUT>
__device__ __forceinline__ void k6_epilogue(const float (&acc)[8][P], int ob, int Cout, int Lout, int lane,
                                            float *outb, int RS, int pad, const float *fcw, float (&fcacc)[4]) {
    const int flat = Cout * Lout;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        const int oc = ob + o;
        if (oc < Cout) {  // uniform
            float *orow = outb + oc * RS + pad + lane;
            const float *fw = fcw + oc * Lout + lane;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float h = acc[o][p];
                const bool valid = lane + 32 * p < Lout;
                if (OUT == 0) {
                    if (valid) orow[32 * p] = h;
                } else {
#pragma unroll
                    for (int q = 0; q < OUT; ++q) {
                        const float w = valid ? fw[q * flat + 32 * p] : 0.0f;
                        fcacc[q] = fmaf(w, h, fcacc[q]);
                    }
                }
            }
        }
    }
}

template <int KS, int P>
__global__ void __launch_bounds__(K6_WARPS * 32, 2) k6_cnn(const K6Args a) {
    extern __shared__ __align__(16) float k6_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *prm = k6_smem;                                   // conv params (+ fc when it fits)
    const int n_stage = a.fc_in_smem ? a.n_params : a.conv_params;
    float *bufs = prm + ((n_stage + 3) & ~3);
    float *bufA = bufs + static_cast<size_t>(warp) * 2 * a.buf_rows * a.row_stride;
    float *bufB = bufA + static_cast<size_t>(a.buf_rows) * a.row_stride;
    const int NW = blockDim.x >> 5;
    for (int i = tid; i < n_stage; i += NW * 32) prm[i] = a.params[i];
    // zero both activation buffers once: the halos stay zero (layers only write [pad, pad + lout))
    for (int i = lane; i < 2 * a.buf_rows * a.row_stride; i += 32) bufA[i] = 0.f;
    __syncthreads();
    // the Linear layer is fused into the last conv layer's epilogue when it has <= 4 outputs and its
    // weights fit in shared memory (the reference shape: 2 x 4096); otherwise it reads them from global
    const bool fuse_fc = a.fc_in_smem != 0;
    const float *fcw = prm + a.fc_w_off;                  // shared (only dereferenced when fuse_fc)
    const float *fcw_g = a.params + a.fc_w_off, *fcb = a.params + a.fc_b_off;
    const int RS = a.row_stride, pad = a.pad, dil = a.dil;

    for (int64_t wi = static_cast<int64_t>(blockIdx.x) * NW + warp; wi < a.n;
         wi += static_cast<int64_t>(gridDim.x) * NW) {
        // ---- stage the window: row c = [pad zeros | W samples | zeros] ----
        const float *xw = a.x + wi * a.win_stride;
        for (int c = 0; c < a.C0; ++c) {
            for (int t = lane; t < a.W; t += 32) bufA[c * RS + pad + t] = __ldg(xw + c * a.W + t);
            for (int t = pad + a.W + lane; t < RS; t += 32) bufA[c * RS + t] = 0.f;  // a longer row may have lived here
        }
        __syncwarp();
        float *in = bufA, *outb = bufB;
        float fcacc[4] = {0.f, 0.f, 0.f, 0.f};  // out_size <= 4 in registers, else generic path below
        for (int l = 0; l < a.n_layers; ++l) {
            const int Cin = a.cin[l], Cout = a.cout[l], CP = a.coutp[l], Lout = a.lout[l];
            const float *wT = prm + a.w_off[l], *bias = prm + a.b_off[l];
            const bool last = l == a.n_layers - 1;
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
                            const float xin = row[32 * p + k * dil];
#pragma unroll
                            for (int o = 0; o < 8; ++o) acc[o][p] = fmaf(wv[o], xin, acc[o][p]);
                        }
                    }
                }
                // ---- activation; store for the next layer or feed the Linear layer ----
                switch (a.act) {  // uniform
                    case 0: k6_activate<0, P>(acc); break;
                    case 1: k6_activate<1, P>(acc); break;
                    case 2: k6_activate<2, P>(acc); break;
                    default: break;
                }
                if (a.bn) {  // eval-mode BatchNorm1d behind the activation (model.py:100-103): y * scale + shift
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        const float sc = bias[CP + ob + o], sh = bias[2 * CP + ob + o];
#pragma unroll
                        for (int p = 0; p < P; ++p) acc[o][p] = fmaf(acc[o][p], sc, sh);
                    }
                }
                if (a.pool) {  // uniform; never together with the fused Linear layer
                    k6_epilogue_pool<P>(acc, ob, Cout, Lout, lane, outb, RS, pad);
                    continue;
                }
                switch ((last && fuse_fc) ? a.out_size : 0) {  // uniform
                    case 0: k6_epilogue<P, 0>(acc, ob, Cout, Lout, lane, outb, RS, pad, fcw, fcacc); break;
                    case 1: k6_epilogue<P, 1>(acc, ob, Cout, Lout, lane, outb, RS, pad, fcw, fcacc); break;
                    case 2: k6_epilogue<P, 2>(acc, ob, Cout, Lout, lane, outb, RS, pad, fcw, fcacc); break;
                    case 3: k6_epilogue<P, 3>(acc, ob, Cout, Lout, lane, outb, RS, pad, fcw, fcacc); break;
                    default: k6_epilogue<P, 4>(acc, ob, Cout, Lout, lane, outb, RS, pad, fcw, fcacc); break;
                }
            }
            if (!last || !fuse_fc) {
                // right halo of a shorter output row must read as zero for the next layer
                if (!last) {
                    for (int oc = 0; oc < Cout; ++oc)
                        for (int t = pad + Lout + lane; t < RS; t += 32) outb[oc * RS + t] = 0.f;
                }
                __syncwarp();
                float *t2 = in; in = outb; outb = t2;
            }
        }
        // ---- Linear ----
        if (fuse_fc) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float v = fcacc[q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && q < a.out_size) a.out[wi * a.out_size + q] = v + fcb[q];
            }
        } else {
            const int Cl = a.cout[a.n_layers - 1], Ll = a.lout[a.n_layers - 1], flat = Cl * Ll;
            for (int q = 0; q < a.out_size; ++q) {
                float v = 0.f;
                for (int f = lane; f < flat; f += 32) {
                    const int oc = f / Ll, t = f - oc * Ll;
                    v = fmaf(__ldg(fcw_g + q * flat + f), in[oc * RS + pad + t], v);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) a.out[wi * a.out_size + q] = v + fcb[q];
            }
        }
        __syncwarp();
        // the buffers' data regions are fully rewritten by the next window; halos were never touched,
        // except the right-halo zeros written above, which only ever hold zeros
    }
}

}  // namespace ofp
#include "cnn_infer_tc.cuh"
#include "cnn_cc_infer.cuh"

using namespace ofp;

extern "C" {

int ofp_cnn_param_count(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                        int32_t kernel_size, int32_t padding, int32_t out_size, int64_t *n_params_out,
                        int32_t *flat_out) {
    return ofp_cnn_param_count_ex(channels, input_size, n_layers, layer_sizes_host, kernel_size, padding, 1, 0, 0,
                                  out_size, n_params_out, flat_out);
}

int ofp_cnn_param_count_ex(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                           int32_t kernel_size, int32_t padding, int32_t dilation, int32_t pool, int32_t batch_norm,
                           int32_t out_size, int64_t *n_params_out, int32_t *flat_out) {
    OFP_REQUIRE(n_layers >= 1 && n_layers <= K6_MAX_LAYERS && layer_sizes_host && n_params_out, "bad argument");
    OFP_REQUIRE(dilation >= 1, "dilation must be >= 1");
    int64_t n = 0;
    int cin = channels, len = input_size;
    for (int l = 0; l < n_layers; ++l) {
        const int cout = layer_sizes_host[l], cp = (cout + 7) & ~7;
        n += static_cast<int64_t>(cin) * kernel_size * cp + cp + (batch_norm ? 2 * cp : 0);
        len = len + 2 * padding - dilation * (kernel_size - 1);
        OFP_REQUIRE(len >= 1, "layer %d has no output positions", l);
        if (pool) len /= 2;
        OFP_REQUIRE(len >= 1, "layer %d has no output positions after pooling", l);
        cin = cout;
    }
    n += static_cast<int64_t>(out_size) * cin * len + out_size;
    *n_params_out = n;
    if (flat_out) *flat_out = cin * len;
    return OFP_OK;
}

int ofp_cnn_forward(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                    int32_t n_layers, const int32_t *layer_sizes_host, int32_t kernel_size, int32_t padding,
                    int32_t activation, const float *params_dev, int32_t out_size, float *out_dev, void *stream) {
    return ofp_cnn_forward_ex(x_dev, n_windows, win_stride, channels, input_size, n_layers, layer_sizes_host, kernel_size,
                              padding, 1, 0, 0, activation, params_dev, out_size, out_dev, stream);
}

int ofp_cnn_forward_ex(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                       int32_t n_layers, const int32_t *layer_sizes_host, int32_t kernel_size, int32_t padding,
                       int32_t dilation, int32_t pool, int32_t batch_norm, int32_t activation, const float *params_dev,
                       int32_t out_size, float *out_dev, void *stream) {
    if (n_windows == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(x_dev && params_dev && out_dev && layer_sizes_host, "null argument");
    OFP_REQUIRE(dilation >= 1 && dilation <= 16, "dilation 1..16 supported");
    pool = pool != 0; batch_norm = batch_norm != 0;
    const bool plain = dilation == 1 && !pool && !batch_norm;
    OFP_REQUIRE(n_layers >= 1 && n_layers <= K6_MAX_LAYERS, "1..%d conv layers supported", K6_MAX_LAYERS);
    OFP_REQUIRE(kernel_size == 1 || kernel_size == 3 || kernel_size == 5 || kernel_size == 7,
                "kernel_size must be 1, 3, 5 or 7");
    OFP_REQUIRE(padding >= 0 && padding <= 8 && channels >= 1 && channels <= 64 && out_size >= 1 && out_size <= 64,
                "bad argument");
    OFP_REQUIRE(input_size >= 1 && input_size <= 256, "input_size up to 256 supported");
    OFP_REQUIRE(activation >= 0 && activation <= 3, "activation: 0 SiLU, 1 ReLU, 2 tanh, 3 identity");
    K6Args a{};
    a.x = x_dev; a.n = n_windows; a.win_stride = win_stride; a.C0 = channels; a.W = input_size; a.n_layers = n_layers;
    a.ks = kernel_size; a.pad = padding; a.act = activation; a.out_size = out_size; a.params = params_dev;
    a.out = out_dev; a.dil = dilation; a.pool = pool; a.bn = batch_norm;
    int cin = channels, len = input_size, off = 0, max_len = input_size, max_rows = channels, max_rows_all = channels;
    for (int l = 0; l < n_layers; ++l) {
        const int cout = layer_sizes_host[l], cp = (cout + 7) & ~7;
        OFP_REQUIRE(cout >= 1 && cout <= 64, "layer sizes 1..64 supported");
        a.cin[l] = cin; a.cout[l] = cout; a.coutp[l] = cp; a.lin[l] = len;
        a.w_off[l] = off; off += cin * kernel_size * cp;
        a.b_off[l] = off; off += cp + (batch_norm ? 2 * cp : 0);
        len = len + 2 * padding - dilation * (kernel_size - 1);
        OFP_REQUIRE(len >= 1 && len <= 256, "layer %d output length %d outside 1..256", l, len);
        max_len = std::max(max_len, len);  // the register tile covers the positions BEFORE pooling
        if (pool) len /= 2;
        OFP_REQUIRE(len >= 1, "layer %d has no output positions after pooling", l);
        a.lout[l] = len;
        max_rows_all = std::max(max_rows_all, cout);
        if (l + 1 < n_layers) max_rows = std::max(max_rows, cout);  // a fused last layer stays in registers
        cin = cout;
    }
    a.conv_params = off;
    const int flat = cin * len;
    a.fc_w_off = off; off += out_size * flat;
    a.fc_b_off = off; off += out_size;
    a.n_params = off;
    const int P = max_len <= 64 ? 2 : (max_len <= 128 ? 4 : 8);
    // odd multiple-of-nothing row stride is fine (all accesses are 32 consecutive words); room for the halo
    a.row_stride = 32 * P + 2 * padding + dilation * (kernel_size - 1) + 2;
    // Tensor-core path for the reference's default shape: [C0 -> 8 -> 16], k = 3, padding 1, SiLU, <= 4 outputs
    const bool no_tc = getenv("OFP_K6_NO_TC") != nullptr;  // read per call: tests compare both kernels in one process
    if (!no_tc && plain && n_layers == 2 && kernel_size == 3 && padding == 1 && activation == 0 && out_size <= 4 &&
        layer_sizes_host[0] == K6T_C1 && layer_sizes_host[1] == K6T_C2 && input_size % 16 == 0 && channels <= 8) {
        const int n_w1 = channels * 24 + 8;
        const size_t smem_tc = sizeof(float) * (((n_w1 + 3) & ~3) + static_cast<size_t>(out_size) * K6T_C2 * K6T_FCS +
                                                static_cast<size_t>(K6T_WARPS) * (channels + K6T_C1) * a.row_stride);
        if (smem_tc <= (226 * 1024) / OFP_K6T_MINCTA) {
            auto kt = P == 2 ? k6_cnn_tc<2> : (P == 4 ? k6_cnn_tc<4> : k6_cnn_tc<8>);
            OFP_CUDA_CHECK(cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_tc)));
            int per_sm = 0;
            OFP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kt, K6T_WARPS * 32, smem_tc));
            per_sm = std::max(per_sm, 1);
            const int64_t want = (n_windows + K6T_WARPS - 1) / K6T_WARPS;
            const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(sm_count()) * per_sm));
            kt<<<grid, K6T_WARPS * 32, smem_tc, static_cast<cudaStream_t>(stream)>>>(a);
            OFP_CUDA_CHECK(cudaGetLastError());
            return OFP_OK;
        }
    }
    // shared memory: staged parameters + two activation buffers per warp.  4 warps with the Linear weights
    // staged and fused (2 CTAs per SM) when that fits, else 4 / 2 / 1 warps and the Linear read through L1.
    int warps = K6_WARPS;
    auto act_bytes_of = [&](int w, int rows) { return sizeof(float) * w * 2 * static_cast<size_t>(rows) * a.row_stride; };
    a.fc_in_smem = !pool && out_size <= 4 && (((a.n_params + 3) & ~3) * sizeof(float) + act_bytes_of(warps, max_rows)) <= 112 * 1024;
    a.buf_rows = a.fc_in_smem ? max_rows : max_rows_all;  // a fused last layer stays in registers
    if (!a.fc_in_smem)
        while (warps > 1 && ((a.conv_params + 3) & ~3) * sizeof(float) + act_bytes_of(warps, a.buf_rows) > 112 * 1024)
            warps /= 2;
    const int staged = a.fc_in_smem ? a.n_params : a.conv_params;
    const size_t smem = ((staged + 3) & ~3) * sizeof(float) + act_bytes_of(warps, a.buf_rows);
    OFP_REQUIRE(smem <= 227 * 1024, "network needs %zu bytes of shared memory per CTA", smem);
    void (*kern)(const K6Args) = nullptr;
#define K6_PICK(KS_)                                                                              \
    kern = P == 2 ? k6_cnn<KS_, 2> : (P == 4 ? k6_cnn<KS_, 4> : k6_cnn<KS_, 8>)
    switch (kernel_size) {
        case 1: K6_PICK(1); break;
        case 3: K6_PICK(3); break;
        case 5: K6_PICK(5); break;
        default: K6_PICK(7); break;
    }
#undef K6_PICK
    OFP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    OFP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    per_sm = std::max(per_sm, 1);
    const int64_t want = (n_windows + warps - 1) / warps;
    const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(sm_count()) * per_sm));
    kern<<<grid, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

// Output length of one CCCNN layer with the options of model.py:485-503 (Conv1d, then MaxPool1d(2, 2) when pool).
static int cccnn_layer_len(int len, int ks, int stride, int padding, int dilation, int pool, int *conv_len) {
    const int span = len + 2 * padding - dilation * (ks - 1) - 1;
    const int lc = span < 0 ? 0 : span / stride + 1;
    if (conv_len) *conv_len = lc;
    return pool ? lc / 2 : lc;
}

int ofp_cccnn_param_count_ex(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                             const int32_t *kernel_sizes_host, const int32_t *strides_host, int32_t padding,
                             int32_t dilation, int32_t pool, int32_t group_norm, int32_t out_size, int32_t group,
                             int64_t *n_params_out, int32_t *n_lags_out) {
    OFP_REQUIRE(n_layers >= 1 && n_layers <= K6_MAX_LAYERS && layer_sizes_host && kernel_sizes_host && strides_host &&
                    n_params_out, "bad argument");
    OFP_REQUIRE(dilation >= 1, "dilation must be >= 1");
    int64_t n = 0;
    int cin = 1, len = input_size;
    for (int l = 0; l < n_layers; ++l) {
        const int cout = layer_sizes_host[l], cp = (cout + 7) & ~7;
        OFP_REQUIRE(kernel_sizes_host[l] >= 1 && strides_host[l] >= 1, "layer %d: kernel size and stride must be >= 1", l);
        n += static_cast<int64_t>(cin) * kernel_sizes_host[l] * cp + cp + (group_norm ? 2 * cp : 0);
        len = cccnn_layer_len(len, kernel_sizes_host[l], strides_host[l], padding, dilation, pool, nullptr);
        OFP_REQUIRE(len >= 1, "layer %d has no output positions", l);
        cin = cout;
    }
    if (group) n *= channels;
    n += static_cast<int64_t>(out_size) * channels * (2 * len - 1) + out_size;
    *n_params_out = n;
    if (n_lags_out) *n_lags_out = 2 * len - 1;
    return OFP_OK;
}

int ofp_cccnn_forward_ex(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                         int32_t n_layers, const int32_t *layer_sizes_host, const int32_t *kernel_sizes_host,
                         const int32_t *strides_host, int32_t padding, int32_t dilation, int32_t pool, int32_t group_norm,
                         int32_t activation, int32_t group, const float *params_dev, int32_t out_size, float *out_dev,
                         void *stream) {
    if (n_windows == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(x_dev && params_dev && out_dev && layer_sizes_host && kernel_sizes_host && strides_host, "null argument");
    OFP_REQUIRE(n_layers >= 1 && n_layers <= K6_MAX_LAYERS, "1..%d conv layers supported", K6_MAX_LAYERS);
    OFP_REQUIRE(padding >= 0 && padding <= 8 && channels >= 1 && channels <= 64 && out_size >= 1 && out_size <= 4,
                "bad argument (out_size up to 4)");
    OFP_REQUIRE(input_size >= 1 && input_size <= 256, "input_size up to 256 supported");
    OFP_REQUIRE(dilation >= 1 && dilation <= 16, "dilation 1..16 supported");
    OFP_REQUIRE(activation >= 0 && activation <= 3, "activation: 0 SiLU, 1 ReLU, 2 tanh, 3 identity");
    // GroupNorm(1, K * channels) of the grouped stack (model.py:494-498 with group = True) takes its statistics over
    // ALL sensor channels of a window at once; the kernel runs the channels one after the other
    OFP_REQUIRE(!(group && group_norm), "group = True together with the norm layer is not supported");
    pool = pool != 0; group_norm = group_norm != 0;
    K6Args a{};
    a.x = x_dev; a.n = n_windows; a.win_stride = win_stride; a.C0 = 1; a.W = input_size; a.n_layers = n_layers;
    a.ks = 0; a.pad = padding; a.act = activation; a.out_size = out_size; a.params = params_dev; a.out = out_dev;
    a.dil = dilation; a.pool = pool; a.bn = group_norm;
    int cin = 1, len = input_size, off = 0, max_conv = 1, max_in = input_size, rows = 1;
    for (int l = 0; l < n_layers; ++l) {
        const int cout = layer_sizes_host[l], cp = (cout + 7) & ~7, ks = kernel_sizes_host[l], st = strides_host[l];
        OFP_REQUIRE(cout >= 1 && cout <= 64, "layer sizes 1..64 supported");
        OFP_REQUIRE(ks >= 1 && ks <= 15 && st >= 1 && st <= 8, "layer %d: kernel size 1..15, stride 1..8 supported", l);
        a.cin[l] = cin; a.cout[l] = cout; a.coutp[l] = cp; a.lin[l] = len; a.ksl[l] = ks; a.strl[l] = st;
        a.w_off[l] = off; off += cin * ks * cp;
        a.b_off[l] = off; off += cp + (group_norm ? 2 * cp : 0);
        int lc = 0;
        len = cccnn_layer_len(len, ks, st, padding, dilation, pool, &lc);
        OFP_REQUIRE(len >= 1 && lc <= 256, "layer %d output length %d outside 1..256", l, lc);
        a.lconv[l] = lc; a.lout[l] = len;
        max_conv = std::max(max_conv, lc);
        max_in = std::max(max_in, len);
        rows = std::max(rows, cout);
        cin = cout;
    }
    OFP_REQUIRE(cin % 8 == 0, "the last layer size must be a multiple of 8 (tensor-core K dimension), got %d", cin);
    OFP_REQUIRE(len % 16 == 0, "the feature-map length must be a multiple of 16, got %d", len);
    a.group_stride = group ? off : 0;
    if (group) off *= channels;
    a.conv_params = off;
    OFP_REQUIRE(off <= 16384, "conv stack too large for shared memory (%d floats)", off);
    const int nb = 2 * len - 1;
    a.fc_w_off = off; off += out_size * channels * nb;
    a.fc_b_off = off; off += out_size;
    a.n_params = off;
    a.row_stride = ((std::max(max_in, max_conv) + 2 * padding + 4) | 1);  // odd: rows start in different banks
    const size_t smem_c = sizeof(float) * (((a.conv_params + 3) & ~3) + 2 * static_cast<size_t>(rows) * a.row_stride +
                                           ((len + 3) & ~3) + 4 * 128);
    OFP_REQUIRE(smem_c <= 227 * 1024, "network needs %zu bytes of shared memory per CTA", smem_c);
    auto kc = max_conv <= 128 ? k6_cccnn_cta<0, 1, OFP_K6CC_ND, true> : k6_cccnn_cta<0, 2, OFP_K6CC_ND, true>;
    OFP_CUDA_CHECK(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_c)));
    int per_sm_c = 0;
    OFP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, kc, 128, smem_c));
    per_sm_c = std::max(per_sm_c, 1);
    const int grid_c = static_cast<int>(std::min<int64_t>(n_windows, static_cast<int64_t>(sm_count()) * per_sm_c));
    kc<<<grid_c, 128, smem_c, static_cast<cudaStream_t>(stream)>>>(a, channels, rows);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

int ofp_cccnn_param_count(int32_t channels, int32_t input_size, int32_t n_layers, const int32_t *layer_sizes_host,
                          int32_t kernel_size, int32_t padding, int32_t out_size, int32_t group,
                          int64_t *n_params_out, int32_t *n_lags_out) {
    OFP_REQUIRE(n_layers >= 1 && n_layers <= K6_MAX_LAYERS && layer_sizes_host && n_params_out, "bad argument");
    int64_t n = 0;
    int cin = 1, len = input_size;
    for (int l = 0; l < n_layers; ++l) {
        const int cout = layer_sizes_host[l], cp = (cout + 7) & ~7;
        n += static_cast<int64_t>(cin) * kernel_size * cp + cp;
        len = len + 2 * padding - (kernel_size - 1);
        OFP_REQUIRE(len >= 1, "layer %d has no output positions", l);
        cin = cout;
    }
    if (group) n *= channels;  // one copy of the stack per sensor channel (groups = channels)
    n += static_cast<int64_t>(out_size) * channels * (2 * len - 1) + out_size;
    *n_params_out = n;
    if (n_lags_out) *n_lags_out = 2 * len - 1;
    return OFP_OK;
}

int ofp_cccnn_forward(const float *x_dev, int64_t n_windows, int64_t win_stride, int32_t channels, int32_t input_size,
                      int32_t n_layers, const int32_t *layer_sizes_host, int32_t kernel_size, int32_t padding,
                      int32_t activation, int32_t group, const float *params_dev, int32_t out_size, float *out_dev,
                      void *stream) {
    if (n_windows == 0) return OFP_OK;  // an empty batch may come with null buffers
    OFP_REQUIRE(x_dev && params_dev && out_dev && layer_sizes_host, "null argument");
    OFP_REQUIRE(n_layers >= 1 && n_layers <= K6_MAX_LAYERS, "1..%d conv layers supported", K6_MAX_LAYERS);
    OFP_REQUIRE(kernel_size == 1 || kernel_size == 3 || kernel_size == 5 || kernel_size == 7,
                "kernel_size must be 1, 3, 5 or 7");
    OFP_REQUIRE(padding >= 0 && padding <= 8 && channels >= 1 && channels <= 64 && out_size >= 1 && out_size <= 4,
                "bad argument (out_size up to 4)");
    OFP_REQUIRE(input_size >= 1 && input_size <= 256, "input_size up to 256 supported");
    OFP_REQUIRE(activation >= 0 && activation <= 3, "activation: 0 SiLU, 1 ReLU, 2 tanh, 3 identity");
    K6Args a{};
    a.x = x_dev; a.n = n_windows; a.win_stride = win_stride; a.C0 = 1; a.W = input_size; a.n_layers = n_layers;
    a.ks = kernel_size; a.pad = padding; a.act = activation; a.out_size = out_size; a.params = params_dev;
    a.out = out_dev;
    int cin = 1, len = input_size, off = 0, max_len = input_size, rows_a = 1, rows_b = 1;
    for (int l = 0; l < n_layers; ++l) {
        const int cout = layer_sizes_host[l], cp = (cout + 7) & ~7;
        OFP_REQUIRE(cout >= 1 && cout <= 64, "layer sizes 1..64 supported");
        a.cin[l] = cin; a.cout[l] = cout; a.coutp[l] = cp; a.lin[l] = len;
        a.w_off[l] = off; off += cin * kernel_size * cp;
        a.b_off[l] = off; off += cp;
        len = len + 2 * padding - (kernel_size - 1);
        OFP_REQUIRE(len >= 1 && len <= 256, "layer %d output length %d outside 1..256", l, len);
        a.lout[l] = len;
        max_len = std::max(max_len, len);
        (l % 2 == 0 ? rows_b : rows_a) = std::max(l % 2 == 0 ? rows_b : rows_a, cout);  // input in A, layer 0 -> B, ...
        cin = cout;
    }
    OFP_REQUIRE(cin % 8 == 0, "the last layer size must be a multiple of 8 (tensor-core K dimension), got %d", cin);
    rows_a = std::max(rows_a, cin); rows_b = std::max(rows_b, cin);  // the hi / lo planes of the feature maps
    OFP_REQUIRE(len % 16 == 0, "the feature-map length must be a multiple of 16, got %d", len);
    // group = True: `channels` copies of the stack one after the other, channel c at c * group_stride
    a.group_stride = group ? off : 0;
    if (group) off *= channels;
    a.conv_params = off;
    OFP_REQUIRE(off <= 16384, "conv stack too large for shared memory (%d floats)", off);
    const int nb = 2 * len - 1;
    a.fc_w_off = off; off += out_size * channels * nb;
    a.fc_b_off = off; off += out_size;
    a.n_params = off;
    const int P = max_len <= 64 ? 2 : (max_len <= 128 ? 4 : 8);
    a.row_stride = 32 * P + 2 * padding + kernel_size + 1;
    // CTA-per-window kernel (4 warps share one set of planes, 24 warps per SM) unless OFP_K6CC_WARP=1 (A/B)
    if (getenv("OFP_K6CC_WARP") == nullptr && len <= 256) {
        const int rows = std::max(rows_a, rows_b);
        const size_t smem_c = sizeof(float) * (((a.conv_params + 3) & ~3) + 2 * static_cast<size_t>(rows) * a.row_stride +
                                               ((len + 3) & ~3) + 4 * 128);
        const int PP = max_len <= 128 ? 1 : 2;
        void (*kc)(const K6Args, int, int) = nullptr;
#define K6CC_PICK(KS_) kc = PP == 1 ? k6_cccnn_cta<KS_, 1, OFP_K6CC_ND, false> : k6_cccnn_cta<KS_, 2, OFP_K6CC_ND, false>
        switch (kernel_size) {
            case 1: K6CC_PICK(1); break;
            case 3: K6CC_PICK(3); break;
            case 5: K6CC_PICK(5); break;
            default: K6CC_PICK(7); break;
        }
#undef K6CC_PICK
        OFP_CUDA_CHECK(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_c)));
        int per_sm_c = 0;
        OFP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, kc, 128, smem_c));
        per_sm_c = std::max(per_sm_c, 1);
        const int grid_c = static_cast<int>(std::min<int64_t>(n_windows, static_cast<int64_t>(sm_count()) * per_sm_c));
        kc<<<grid_c, 128, smem_c, static_cast<cudaStream_t>(stream)>>>(a, channels, rows);
        OFP_CUDA_CHECK(cudaGetLastError());
        return OFP_OK;
    }
    const int warps = 3;
    const size_t smem = sizeof(float) * (((a.conv_params + 3) & ~3) +
                                         static_cast<size_t>(warps) * ((rows_a + rows_b) * a.row_stride + ((len + 3) & ~3) + 128));
    OFP_REQUIRE(smem <= 227 * 1024, "network needs %zu bytes of shared memory per CTA", smem);
    void (*kern)(const K6Args, int, int, int) = nullptr;
    // diagonals per pass: 4 when the number of tile diagonals (V / 8) allows, else 2 (V is a multiple of 16)
#define K6C_PICK(KS_)                                                                                       \
    kern = (len % 32 == 0) ? (P == 2 ? k6_cccnn<KS_, 2, 4> : (P == 4 ? k6_cccnn<KS_, 4, 4> : k6_cccnn<KS_, 8, 4>)) \
                           : (P == 2 ? k6_cccnn<KS_, 2, 2> : (P == 4 ? k6_cccnn<KS_, 4, 2> : k6_cccnn<KS_, 8, 2>))
    switch (kernel_size) {
        case 1: K6C_PICK(1); break;
        case 3: K6C_PICK(3); break;
        case 5: K6C_PICK(5); break;
        default: K6C_PICK(7); break;
    }
#undef K6C_PICK
    OFP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int per_sm = 0;
    OFP_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    per_sm = std::max(per_sm, 1);
    const int64_t want = (n_windows + warps - 1) / warps;
    const int grid = static_cast<int>(std::min<int64_t>(want, static_cast<int64_t>(sm_count()) * per_sm));
    kern<<<grid, warps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a, channels, rows_a, rows_b);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

}  // extern "C"
