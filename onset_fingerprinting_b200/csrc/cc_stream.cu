// Streaming cross-correlation (sm_100a): twin of the reference's CPython extension
// online_cc.CrossCorrelation(n, block_size).update(a, b) (c/cross_corr.c:106-193, 257-273), which
// returns the full 2n-1-lag cross-correlation of the last n samples of two streams after every block.
//
// The reference updates running per-lag sums incrementally (Kahan-compensated, with a periodic exact
// recompute of one row per call) because a CPU cannot afford n^2 MACs per block.  On the GPU the
// exact recompute of ALL lags is cheap (n = 256: 65 k MACs per pair, one CTA), has no drift, and
// batches over thousands of stream pairs per launch; accumulation is in double.  The reference's own
// acceptance threshold is |err| < 1e-3 against np.correlate (c/test.py:42).
#include "ofp_common.cuh"

struct ofp_ccstream {
    int32_t n_pairs, n, block;
    float *ring;  // [P, 2, n] last n samples of each stream, time ordered
};

namespace ofp {

__global__ void __launch_bounds__(256) k_ccstream(float *ring, const float *a, const float *b, float *out, int n,
                                                  int block) {
    extern __shared__ double sh[];
    double *xa = sh, *xb = sh + n;
    const int p = blockIdx.x, tid = threadIdx.x;
    float *ra = ring + static_cast<int64_t>(p) * 2 * n, *rb = ra + n;
    // shift in the new block (circular_array.h:49-60 + rearrange, 129-141)
    for (int i = tid; i < n; i += 256) {
        const float va = i < n - block ? ra[i + block] : a[static_cast<int64_t>(p) * block + (i - (n - block))];
        const float vb = i < n - block ? rb[i + block] : b[static_cast<int64_t>(p) * block + (i - (n - block))];
        xa[i] = va; xb[i] = vb;
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) { ra[i] = static_cast<float>(xa[i]); rb[i] = static_cast<float>(xb[i]); }
    // np.correlate(a, b, "full")[k] = sum_i a[i + m] * b[i], m = k - (n - 1)
    for (int k = tid; k < 2 * n - 1; k += 256) {
        const int m = k - (n - 1);
        const int i0 = m < 0 ? -m : 0, i1 = m > 0 ? n - m : n;
        double acc = 0.0;
        for (int i = i0; i < i1; ++i) acc = __fma_rn(xa[i + m], xb[i], acc);
        out[static_cast<int64_t>(p) * (2 * n - 1) + k] = static_cast<float>(acc);
    }
}

}  // namespace ofp

using namespace ofp;

extern "C" {

int ofp_ccstream_create(ofp_ccstream **out, int32_t n_pairs, int32_t n, int32_t block_size) {
    OFP_REQUIRE(out && n_pairs > 0 && n > 0 && block_size > 0 && block_size <= n, "bad argument");
    OFP_REQUIRE(n <= 8192, "n must be <= 8192");
    ofp_ccstream *h = new ofp_ccstream{n_pairs, n, block_size, nullptr};
    const size_t bytes = sizeof(float) * 2 * static_cast<size_t>(n) * n_pairs;
    if (cudaMalloc(&h->ring, bytes) != cudaSuccess || cudaMemset(h->ring, 0, bytes) != cudaSuccess) {
        set_error("cudaMalloc/cudaMemset ccstream ring: %s", cudaGetErrorString(cudaGetLastError()));
        delete h;
        return OFP_ECUDA;
    }
    *out = h;
    return OFP_OK;
}

int ofp_ccstream_destroy(ofp_ccstream *h) {
    if (!h) return OFP_OK;
    cudaFree(h->ring);
    delete h;
    return OFP_OK;
}

int ofp_ccstream_update(ofp_ccstream *h, const float *a_dev, const float *b_dev, float *out_dev, void *stream) {
    OFP_REQUIRE(h && a_dev && b_dev && out_dev, "null argument");
    const size_t smem = sizeof(double) * 2 * h->n;
    OFP_CUDA_CHECK(cudaFuncSetAttribute(k_ccstream, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    k_ccstream<<<h->n_pairs, 256, smem, static_cast<cudaStream_t>(stream)>>>(h->ring, a_dev, b_dev, out_dev, h->n,
                                                                             h->block);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}

}  // extern "C"
