// icache_probe2.cu -- do the warps of one scheduler share instruction fetches?
// Every warp loops over a straight-line body of N FFMAs.  K = 1: all warps run the same copy of the
// body; K = 8: warp w runs copy (w / 4) % 8, i.e. the warps of one SMSP (w % 4 equal) all sit in
// DIFFERENT code, as the warps of the fused kernel do when they are in different stages.
#include <cstdio>
#include <cuda_runtime.h>
template <int N, int COPY>
__device__ __noinline__ void body(float (&r)[8], float a, float b, int iters) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < N; i++) r[i & 7] = fmaf(r[i & 7], a + COPY, b);
    }
}
template <int N, int K>
__global__ void __launch_bounds__(1024) k_body(float* out, int iters, float a, float b) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = threadIdx.x * 0.001f + i;
    const int copy = (K == 1) ? 0 : ((threadIdx.x >> 5) >> 2) % K;
    switch (copy) {
        case 0: body<N, 0>(r, a, b, iters); break;
        case 1: body<N, 1>(r, a, b, iters); break;
        case 2: body<N, 2>(r, a, b, iters); break;
        case 3: body<N, 3>(r, a, b, iters); break;
        case 4: body<N, 4>(r, a, b, iters); break;
        case 5: body<N, 5>(r, a, b, iters); break;
        case 6: body<N, 6>(r, a, b, iters); break;
        default: body<N, 7>(r, a, b, iters); break;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += r[i];
    if (s == 12345.678f) out[0] = s;
}
template <int N, int K>
void run(int warps_per_sm, int nsm, float* d) {
    const int threads = warps_per_sm * 32;
    int iters = (int)(100000000LL / ((long long)N * (warps_per_sm / 4)));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_body<N, K><<<nsm, threads>>>(d, 4, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k_body<N, K><<<nsm, threads>>>(d, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double instr_per_smsp = (double)N * iters * warps_per_sm / 4.0;
    printf("copies %d  body %5d instrs (%5.1f KB each)  warps/SM %2d : %.3f warp-instr / cycle / SMSP\n",
           K, N, N * 16 / 1024.0, warps_per_sm, instr_per_smsp / (ms * 1e-3 * clk_khz * 1e3));
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float* d; cudaMalloc(&d, 4);
    for (int w : {8, 20, 32}) {
        run<96, 1>(w, nsm, d);  run<96, 8>(w, nsm, d);
        run<192, 1>(w, nsm, d); run<192, 8>(w, nsm, d);
        run<320, 1>(w, nsm, d); run<320, 8>(w, nsm, d);
        run<640, 1>(w, nsm, d); run<640, 8>(w, nsm, d);
    }
    return 0;
}
