// Device-side synthetic multi-mic drum audio for the benchmark (SURVEY.md section 8d signal model:
// one decaying 900 Hz burst every `hit_period` samples from a random position on the drumhead,
// per-channel arrival delay round(dist / c * sr), amplitude 0.5 * 10 / dist, white noise).
// Not part of the reference; it only feeds bench.py (the CPU baseline gets a D2H copy of the same
// samples) and the large-size property tests.
#include "ofp_common.cuh"

namespace ofp {

struct SynthArgs {
    float *x;
    int64_t R, N;
    int32_t C;
    float sensors[32 * 3];
    float c_cm_s, sr, noise, radius;
    int64_t first_hit, hit_period;
    int32_t burst_len, tail_guard;
    uint64_t seed;
    int64_t rec_offset;  // global index of recording 0 (multi-GPU shards draw different hits)
};

__device__ __forceinline__ uint64_t mix64(uint64_t z) {  // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void k_synth(const SynthArgs a) {
    const int64_t total = a.R * a.N * a.C;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % a.C);
        const int64_t n = (i / a.C) % a.N;
        const int64_t r = i / (a.C * a.N) + a.rec_offset;
        // white noise, Box-Muller on two hashed uniforms
        const uint64_t h1 = mix64(a.seed ^ mix64(static_cast<uint64_t>(r) * 0x100000001B3ull + n * 64 + c));
        const uint64_t h2 = mix64(h1);
        float v = a.noise * sqrtf(-2.0f * __logf(u01(h1))) * __cosf(6.2831853f * u01(h2));
        const int64_t rel = n - a.first_hit;
        if (rel >= 0) {
            const int64_t hit = rel / a.hit_period;
            const uint64_t g1 = mix64(a.seed * 31 + mix64(static_cast<uint64_t>(r) * 1000003ull + hit));
            const uint64_t g2 = mix64(g1);
            const float rr = 0.85f * a.radius * sqrtf(u01(g1));
            const float ang = 6.2831853f * u01(g2);
            const float px = rr * cosf(ang), py = rr * sinf(ang);
            const float dx = a.sensors[3 * c] - px, dy = a.sensors[3 * c + 1] - py, dz = a.sensors[3 * c + 2];
            const float dist = sqrtf(dx * dx + dy * dy + dz * dz);
            const int64_t delay = static_cast<int64_t>(rintf(dist / a.c_cm_s * a.sr));
            const int64_t k = rel - hit * a.hit_period - delay;
            if (k >= 0 && k < a.burst_len && a.first_hit + hit * a.hit_period + delay + a.tail_guard < a.N) {
                const float t = static_cast<float>(k) / a.sr;
                v += 0.5f * 10.0f / dist * expf(-400.0f * t) * sinf(6.2831853f * 900.0f * t);
            }
        }
        a.x[i] = v;
    }
}

}  // namespace ofp

using namespace ofp;

extern "C" int ofp_synth_drum(float *x_dev, int64_t n_rec, int64_t n_samples, int32_t n_channels,
                              const float *sensors_xyz_host, float c_cm_s, float sr, float noise, float radius_cm,
                              int64_t first_hit, int64_t hit_period, int32_t tail_guard, uint64_t seed,
                              int64_t rec_offset, void *stream) {
    OFP_REQUIRE(x_dev && sensors_xyz_host && n_channels >= 1 && n_channels <= 32, "bad argument");
    SynthArgs a;
    a.x = x_dev; a.R = n_rec; a.N = n_samples; a.C = n_channels;
    for (int i = 0; i < 3 * n_channels; ++i) a.sensors[i] = sensors_xyz_host[i];
    a.c_cm_s = c_cm_s; a.sr = sr; a.noise = noise; a.radius = radius_cm;
    a.first_hit = first_hit; a.hit_period = hit_period; a.burst_len = 4096; a.tail_guard = tail_guard; a.seed = seed;
    a.rec_offset = rec_offset;
    const int blocks = sm_count() * 16;
    k_synth<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    OFP_CUDA_CHECK(cudaGetLastError());
    return OFP_OK;
}
