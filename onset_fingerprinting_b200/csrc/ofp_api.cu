// libofp.so: error reporting, version, driver entry points shared by the kernels.
#include "ofp_common.cuh"

#include <cudaTypedefs.h>
#include <stdarg.h>

namespace ofp {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;  // B200
    }
    return n;
}

int encode_tmap_2d_f32(CUtensorMap *map, const void *base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                       uint32_t box0, uint32_t box1) {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
            return OFP_ECUDA;
        }
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d): dims %llu x %llu stride %llu box %u x %u", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)stride1_bytes, box0, box1);
        return OFP_ECUDA;
    }
    return OFP_OK;
}

}  // namespace ofp

extern "C" {
const char *ofp_last_error(void) { return ofp::g_err; }
const char *ofp_version(void) { return "libofp 0.2 sm_100a"; }

int ofp_copy2d_async(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width_bytes, size_t rows,
                     int to_host, void *stream) {
    OFP_REQUIRE(dst && src && dst_pitch >= width_bytes && src_pitch >= width_bytes, "bad argument");
    if (width_bytes == 0 || rows == 0) return OFP_OK;
    OFP_CUDA_CHECK(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows,
                                     to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice,
                                     static_cast<cudaStream_t>(stream)));
    return OFP_OK;
}
}
