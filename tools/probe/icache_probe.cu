// icache_probe.cu -- how fast can an SM issue when the loop body does not fit the L0 instruction cache?
// Body = N independent FFMAs (8 accumulator chains), straight-line, looped.  Reports issue rate per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void __launch_bounds__(1024) k_body(float* out, int iters, float a, float b, int desync) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 0.001f + i;
    if (desync) {                                        // spread the warps of an SMSP evenly over the loop body
        const long long t0 = clock64(), dl = (long long)((threadIdx.x >> 5) >> 2) * N;
        while (clock64() - t0 < dl) {}
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) r[i & 7] = fmaf(r[i & 7], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i];
    if (s == 12345.678f) out[0] = s;
}
template <int N>
void run(int warps_per_sm, int nsm, float* d, int desync) {
    const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
    const int ctas_per_sm = (warps_per_sm * 32 + threads - 1) / threads;
    const long long target = 200000000LL;                 // warp-instructions per SMSP-ish
    int iters = (int)(target / ((long long)N * (warps_per_sm / 4 > 0 ? warps_per_sm / 4 : 1)));
    if (iters < 4) iters = 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_body<N><<<nsm * ctas_per_sm, threads>>>(d, 4, 1.0001f, 0.5f, desync);
    cudaEventRecord(e0);
    k_body<N><<<nsm * ctas_per_sm, threads>>>(d, iters, 1.0001f, 0.5f, desync);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double instr_per_smsp = (double)N * iters * warps_per_sm / 4.0;
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%s body %5d instrs (%6.1f KB)  warps/SM %2d : %.3f warp-instr / cycle / SMSP (at %d MHz nominal)\n",
           desync ? "desync" : "sync  ", N, N * 16 / 1024.0, warps_per_sm, instr_per_smsp / cycles, clk_khz / 1000);
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 4);
    for (int ds : {0, 1})
        for (int w : {8, 16, 24}) {
            run<256>(w, nsm, d, ds); run<384>(w, nsm, d, ds); run<512>(w, nsm, d, ds); run<768>(w, nsm, d, ds); run<1024>(w, nsm, d, ds); run<1536>(w, nsm, d, ds); run<1792>(w, nsm, d, ds); run<2048>(w, nsm, d, ds); run<2560>(w, nsm, d, ds);
        }
    return 0;
}
