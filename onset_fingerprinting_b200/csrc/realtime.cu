// Realtime session: PlayRec's per-block callback (reference realtime/audio.py:62-122: write block ->
// detector -> locate) for S concurrent streams as ONE replayed CUDA graph per block.
//
// ofp_detect_block + ofp_stream_locate driven from Python cost ~110 us per 128-sample block, almost all of it
// host work (argument marshalling, tensor allocation, tensor-map encoding, two launches, a blocking read of
// the result).  The session owns every buffer, captures [detector kernel -> locate kernel -> index advance ->
// copy of (xy, found) to pinned host memory] once, and each step is: one async copy of the block into the
// fixed staging buffer, one cudaGraphLaunch, one stream synchronisation.  Kernel arguments are baked into a
// captured graph, so the only per-block scalar -- the block's start index -- lives in device memory
// (ofp_stream_locate_dev).
#include "ofp_common.cuh"

#include <cstring>

extern "C" int ofp_stream_locate_ring_dev(const double *, int32_t, const float *, int32_t, const float *, const float *,
                                          const float *, double, double, double, double, int32_t, int32_t,
                                          const int32_t *, const int32_t *, const int32_t *, int64_t *, int32_t,
                                          const float *, int32_t, int32_t, int32_t, int32_t, int32_t *, int32_t *,
                                          int32_t *, int64_t *, double *, int32_t *, void *);
extern "C" int ofp_ring_write(float *, int32_t, const float *, int64_t, int32_t, int32_t, int32_t, const int64_t *, void *);
extern "C" int ofp_stream_locate_dev(const double *, int32_t, const float *, int32_t, const float *, const float *,
                                     const float *, double, double, double, double, int32_t, int32_t, const int32_t *,
                                     const int32_t *, const int32_t *, int64_t *, int32_t, int32_t *, int32_t *,
                                     int32_t *, int64_t *, double *, int32_t *, void *);

struct ofp_rt {
    int32_t S = 0, C = 0, B = 0;
    ofp_detector *det = nullptr;
    float *in = nullptr;                        // [S, B, C] staging
    float *ring = nullptr;                      // [S, ring_rows, C] most recent audio rows (ring_rows = 0: no refinement)
    int32_t ring_rows = 0, ring_tol = 50, ring_cutoff = 10;
    int32_t *ch = nullptr, *dl = nullptr, *cnt = nullptr, *found = nullptr;
    double *xy = nullptr;
    int64_t *cur = nullptr;                     // device: start index of the next block
    void *state[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t state_bytes[4] = {0, 0, 0, 0};
    double *xy_h = nullptr;                     // pinned mirrors
    int32_t *found_h = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t producer_ev = nullptr;          // ofp_rt_wait_stream
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    // geometry (device pointers owned by the caller, must outlive the session)
    const double *locs; const float *maps, *mx, *mn, *mm;
    int32_t n_sensors, map_size;
    double radius, spcm, sr, c;
};

namespace {
int rt_enqueue(ofp_rt *r) {
    int rc = ofp_detect_block(r->det, r->in, static_cast<int64_t>(r->B) * r->C, nullptr, r->ch, r->dl, r->cnt, r->stream);
    if (rc != OFP_OK) return rc;
    if (r->ring) {  // PlayRec.callback: rec_audio.write(block) -> detect_hits -> locate(..., rec_audio)
        rc = ofp_ring_write(r->ring, r->ring_rows, r->in, 0, r->S, r->B, r->C, r->cur, r->stream);
        if (rc != OFP_OK) return rc;
        rc = ofp_stream_locate_ring_dev(r->locs, r->n_sensors, r->maps, r->map_size, r->mx, r->mn, r->mm, r->radius,
                                        r->spcm, r->sr, r->c, r->S, r->C, r->ch, r->dl, r->cnt, r->cur, r->B, r->ring,
                                        r->ring_rows, r->B, r->ring_tol, r->ring_cutoff,
                                        static_cast<int32_t *>(r->state[0]), static_cast<int32_t *>(r->state[1]),
                                        static_cast<int32_t *>(r->state[2]), static_cast<int64_t *>(r->state[3]), r->xy,
                                        r->found, r->stream);
    } else {
        rc = ofp_stream_locate_dev(r->locs, r->n_sensors, r->maps, r->map_size, r->mx, r->mn, r->mm, r->radius, r->spcm,
                                   r->sr, r->c, r->S, r->C, r->ch, r->dl, r->cnt, r->cur, r->B,
                                   static_cast<int32_t *>(r->state[0]), static_cast<int32_t *>(r->state[1]),
                                   static_cast<int32_t *>(r->state[2]), static_cast<int64_t *>(r->state[3]), r->xy,
                                   r->found, r->stream);
    }
    if (rc != OFP_OK) return rc;
    OFP_CUDA_CHECK(cudaMemcpyAsync(r->xy_h, r->xy, sizeof(double) * 2 * r->S, cudaMemcpyDeviceToHost, r->stream));
    OFP_CUDA_CHECK(cudaMemcpyAsync(r->found_h, r->found, sizeof(int32_t) * r->S, cudaMemcpyDeviceToHost, r->stream));
    return OFP_OK;
}
}  // namespace

extern "C" {

int ofp_rt_destroy(ofp_rt *r) {
    if (!r) return OFP_OK;
    if (r->exec) cudaGraphExecDestroy(r->exec);
    if (r->producer_ev) cudaEventDestroy(r->producer_ev);
    if (r->graph) cudaGraphDestroy(r->graph);
    if (r->stream) cudaStreamDestroy(r->stream);
    cudaFree(r->in); cudaFree(r->ch); cudaFree(r->dl); cudaFree(r->cnt); cudaFree(r->found); cudaFree(r->xy);
    cudaFree(r->cur); cudaFree(r->ring);
    for (void *p : r->state) cudaFree(p);
    cudaFreeHost(r->xy_h); cudaFreeHost(r->found_h);
    ofp_detector_destroy(r->det);
    delete r;
    return OFP_OK;
}

int ofp_rt_reset(ofp_rt *r) {
    OFP_REQUIRE(r, "null session");
    int rc = ofp_detector_reset(r->det, r->stream);
    if (rc != OFP_OK) return rc;
    for (int i = 0; i < 4; ++i) OFP_CUDA_CHECK(cudaMemsetAsync(r->state[i], 0, r->state_bytes[i], r->stream));
    OFP_CUDA_CHECK(cudaMemsetAsync(r->cur, 0, sizeof(int64_t), r->stream));
    if (r->ring)
        OFP_CUDA_CHECK(cudaMemsetAsync(r->ring, 0, sizeof(float) * static_cast<size_t>(r->S) * r->C * r->ring_rows, r->stream));
    OFP_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    return OFP_OK;
}

int ofp_rt_create(ofp_rt **out, int32_t n_streams, const ofp_detector_params *p, const double *sensor_xyz_dev,
                  int32_t n_sensors, const float *lag_maps_dev, int32_t map_size, const float *max_lags_dev,
                  const float *min_lags_dev, const float *max_max_dev, double radius_cm, double samples_per_cm,
                  double sr, double c_cm_s, int32_t use_graph, int32_t ring_rows) {
    OFP_REQUIRE(out && p && sensor_xyz_dev && lag_maps_dev && max_lags_dev && min_lags_dev && max_max_dev, "null argument");
    OFP_REQUIRE(n_streams > 0, "n_streams must be positive");
    OFP_REQUIRE(ring_rows == 0 || ring_rows >= p->block_size, "ring_rows must be 0 (no refinement) or >= block_size");
    ofp_rt *r = new ofp_rt;
    r->S = n_streams; r->C = p->n_channels; r->B = p->block_size; r->ring_rows = ring_rows;
    r->locs = sensor_xyz_dev; r->maps = lag_maps_dev; r->mx = max_lags_dev; r->mn = min_lags_dev; r->mm = max_max_dev;
    r->n_sensors = n_sensors; r->map_size = map_size; r->radius = radius_cm; r->spcm = samples_per_cm; r->sr = sr;
    r->c = c_cm_s;
#define RT_CHECK(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::ofp::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            ofp_rt_destroy(r);                                                                    \
            return OFP_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)
    int rc = ofp_detector_create(&r->det, n_streams, p);
    if (rc != OFP_OK) { ofp_rt_destroy(r); return rc; }
    rc = ofp_stream_locate_state_bytes(n_streams, &r->state_bytes[0], &r->state_bytes[1], &r->state_bytes[2],
                                       &r->state_bytes[3]);
    if (rc != OFP_OK) { ofp_rt_destroy(r); return rc; }
    const size_t SC = static_cast<size_t>(r->S) * r->C;
    RT_CHECK(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    RT_CHECK(cudaMalloc(&r->in, sizeof(float) * SC * r->B));
    if (ring_rows > 0) {
        RT_CHECK(cudaMalloc(&r->ring, sizeof(float) * SC * ring_rows));
        RT_CHECK(cudaMemset(r->ring, 0, sizeof(float) * SC * ring_rows));
    }
    RT_CHECK(cudaMalloc(&r->ch, sizeof(int32_t) * SC));
    RT_CHECK(cudaMalloc(&r->dl, sizeof(int32_t) * SC));
    RT_CHECK(cudaMalloc(&r->cnt, sizeof(int32_t) * r->S));
    RT_CHECK(cudaMalloc(&r->found, sizeof(int32_t) * r->S));
    RT_CHECK(cudaMalloc(&r->xy, sizeof(double) * 2 * r->S));
    RT_CHECK(cudaMalloc(&r->cur, sizeof(int64_t)));
    for (int i = 0; i < 4; ++i) RT_CHECK(cudaMalloc(&r->state[i], static_cast<size_t>(r->state_bytes[i])));
    RT_CHECK(cudaMallocHost(&r->xy_h, sizeof(double) * 2 * r->S));
    RT_CHECK(cudaMallocHost(&r->found_h, sizeof(int32_t) * r->S));
    RT_CHECK(cudaMemset(r->in, 0, sizeof(float) * SC * r->B));
    RT_CHECK(cudaDeviceSynchronize());
    // one eager pass uploads the detector's tables and sets the kernel attributes (not capturable), then the
    // state is reset and the same sequence is captured.  The pass must already see defined state: the group
    // tables and counters come from cudaMalloc and the locate kernel indexes with them.
    RT_CHECK(cudaMemset(r->ch, 0, sizeof(int32_t) * SC));
    RT_CHECK(cudaMemset(r->dl, 0, sizeof(int32_t) * SC));
    RT_CHECK(cudaMemset(r->cnt, 0, sizeof(int32_t) * r->S));
    rc = ofp_rt_reset(r);
    if (rc == OFP_OK) rc = rt_enqueue(r);
    if (rc == OFP_OK) rc = ofp_rt_reset(r);
    if (rc != OFP_OK) { ofp_rt_destroy(r); return rc; }
    if (use_graph) {
        RT_CHECK(cudaStreamBeginCapture(r->stream, cudaStreamCaptureModeThreadLocal));
        rc = rt_enqueue(r);
        cudaError_t e = cudaStreamEndCapture(r->stream, &r->graph);
        if (rc != OFP_OK || e != cudaSuccess) {
            if (rc == OFP_OK) ::ofp::set_error("cudaStreamEndCapture -> %s", cudaGetErrorString(e));
            ofp_rt_destroy(r);
            return rc != OFP_OK ? rc : OFP_ECUDA;
        }
        RT_CHECK(cudaGraphInstantiate(&r->exec, r->graph, 0));
    }
#undef RT_CHECK
    *out = r;
    return OFP_OK;
}

int ofp_rt_step(ofp_rt *r, const float *blocks, int32_t blocks_on_host, int64_t stream_stride, double *xy_host,
                int32_t *found_host) {
    OFP_REQUIRE(r && blocks, "null argument");
    const size_t row = sizeof(float) * r->B * r->C;
    const cudaMemcpyKind kind = blocks_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (stream_stride <= 0) stream_stride = static_cast<int64_t>(r->B) * r->C;
    OFP_CUDA_CHECK(cudaMemcpy2DAsync(r->in, row, blocks, sizeof(float) * stream_stride, row, r->S, kind, r->stream));
    if (r->exec) {
        OFP_CUDA_CHECK(cudaGraphLaunch(r->exec, r->stream));
    } else {
        int rc = rt_enqueue(r);
        if (rc != OFP_OK) return rc;
    }
    OFP_CUDA_CHECK(cudaStreamSynchronize(r->stream));
    if (xy_host) memcpy(xy_host, r->xy_h, sizeof(double) * 2 * r->S);
    if (found_host) memcpy(found_host, r->found_h, sizeof(int32_t) * r->S);
    return OFP_OK;
}

/* Order the session's private stream after everything queued so far on `producer_stream` (the stream that
 * writes the device-resident blocks handed to the next ofp_rt_step): an event recorded there, waited on here. */
int ofp_rt_wait_stream(ofp_rt *r, void *producer_stream) {
    OFP_REQUIRE(r, "null session");
    if (!r->producer_ev) OFP_CUDA_CHECK(cudaEventCreateWithFlags(&r->producer_ev, cudaEventDisableTiming));
    OFP_CUDA_CHECK(cudaEventRecord(r->producer_ev, static_cast<cudaStream_t>(producer_stream)));
    OFP_CUDA_CHECK(cudaStreamWaitEvent(r->stream, r->producer_ev, 0));
    return OFP_OK;
}

/* pinned result buffers of the last step (valid until the next step): avoids the copy into caller memory */
int ofp_rt_results(ofp_rt *r, const double **xy_host, const int32_t **found_host) {
    OFP_REQUIRE(r, "null session");
    if (xy_host) *xy_host = r->xy_h;
    if (found_host) *found_host = r->found_h;
    return OFP_OK;
}

}  // extern "C"
