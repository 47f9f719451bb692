// pskd_synth.cu -- counter-based synthetic PSK channel-bank generator (benchmark / test input).
// Not part of the reference; SURVEY.md section 8d fixes the value distribution: unit-amplitude
// M-PSK x pulse envelope, per-channel carrier offset / initial phase / timing shift, slow phase
// wander, complex AWGN.  Sample n of channel c depends only on (seed, c, n), so any sub-range of
// any channel can be regenerated independently (e.g. on another GPU or for the CPU baseline).
#include "../../include/pskd.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(256)
k_synth(float2* __restrict__ iq, size_t iq_stride, int ch0, size_t n_complex, pskd_synth cfg)
{
    const int cl = blockIdx.y;
    const uint64_t c = (uint64_t)(ch0 + cl);
    const uint64_t ckey = mix64(cfg.seed ^ (c * 0xD1B54A32D192ED03ull));
    // per-channel constants
    const uint64_t h0 = mix64(ckey ^ 0x1111), h1 = mix64(ckey ^ 0x2222), h2 = mix64(ckey ^ 0x3333);
    const float phase0 = 6.2831853f * u01((uint32_t)h0);
    const int S = cfg.samplesPerBaud, M = cfg.constelationSize;
    double freq = (double)cfg.freq_max * (2.0 * (double)u01((uint32_t)(h0 >> 32)) - 1.0);
    if (cfg.period > 0) {                          // replayable buffer: M * freq * period is a whole number of cycles
        const double q = (double)M * (double)cfg.period;
        freq = rint(freq * q) / q;
    }
    const int shift = (int)((uint32_t)h1 % (uint32_t)S);
    const double wf0 = 2e-6 + 2e-5 * (double)u01((uint32_t)(h1 >> 32));
    const double wf1 = 2e-6 + 2e-5 * (double)u01((uint32_t)h2);
    const float wp0 = 6.2831853f * u01((uint32_t)(h2 >> 32)), wp1 = 6.2831853f * u01((uint32_t)mix64(h2));
    float2* out = iq + (size_t)cl * iq_stride;
    for (size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x; n < n_complex; n += (size_t)gridDim.x * blockDim.x) {
        const uint64_t pos = n + (uint64_t)shift;
        const uint64_t j = pos / (uint64_t)S;
        const int p = (int)(pos - j * (uint64_t)S);
        const uint32_t sym = (uint32_t)(mix64(ckey ^ (j * 0xA24BAED4963EE407ull)) >> 33) % (uint32_t)M;
        double cyc = freq * (double)n;
        cyc -= floor(cyc);
        float ph = 6.2831853f * ((float)sym / (float)M) + 6.2831853f * (float)cyc + phase0;
        if (cfg.pn_sigma > 0.0f) {
            double a0 = wf0 * (double)n; a0 -= floor(a0);
            double a1 = wf1 * (double)n; a1 -= floor(a1);
            ph += cfg.pn_sigma * (__sinf(6.2831853f * (float)a0 + wp0) + __sinf(6.2831853f * (float)a1 + wp1));
        }
        float env = 1.0f;
        if (cfg.shaped != 0.0f) env = 0.6f + 0.4f * __sinf(3.14159265f * ((float)p + 0.5f) / (float)S);
        float sn, cs;
        __sincosf(ph, &sn, &cs);
        const uint64_t hn = mix64(ckey ^ ((uint64_t)n * 0x9FB21C651E98DF25ull) ^ 0x5555);
        const float u1 = u01((uint32_t)hn), u2 = u01((uint32_t)(hn >> 32));
        const float rad = cfg.sigma * sqrtf(-2.0f * __logf(u1));
        float ns, nc;
        __sincosf(6.2831853f * u2, &ns, &nc);
        out[n] = make_float2(env * cs + rad * nc, env * sn + rad * ns);
    }
}

}  // namespace

extern "C" int pskd_synth_fill(int device, float* iq_dev, size_t iq_stride, int ch0, int n_channels,
                               size_t n_complex, const pskd_synth* cfg, void* stream)
{
    if (!iq_dev || !cfg || n_channels < 1 || cfg->samplesPerBaud < 1 || cfg->constelationSize < 1) return PSKD_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return PSKD_ERR_CUDA;
    if (n_complex == 0) return PSKD_OK;
    unsigned gx = (unsigned)((n_complex + 255) / 256);
    if (gx > 2048) gx = 2048;
    for (int c0 = 0; c0 < n_channels; c0 += 32768) {
        int nc = n_channels - c0 < 32768 ? n_channels - c0 : 32768;
        dim3 grid(gx, (unsigned)nc);
        k_synth<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(iq_dev) + (size_t)c0 * iq_stride,
                                                         iq_stride, ch0 + c0, n_complex, *cfg);
    }
    return cudaGetLastError() == cudaSuccess ? PSKD_OK : PSKD_ERR_CUDA;
}
