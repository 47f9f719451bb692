// power_probe.cu -- what does one warp instruction / one byte cost in ENERGY on this board?
//
// The demod kernels run into the B200's 1000 W power cap (tools/probe/step_trace.py): once they do, a step's time is
// its energy divided by the cap, so the budget that matters is joules, not issue slots.  This probe runs one
// micro-kernel at a time for ~1.5 s on every SM (full occupancy, every scheduler issuing the named instruction in
// independent chains), reads NVML's energy counter and SM clock around it and prints
//     watts, SM clock, instructions (or bytes) per second  ->  marginal energy above the "resident but idle" kernel.
// Build + run on the GPU box:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o power_probe power_probe.cu -lnvidia-ml && ./power_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include <nvml.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;     // loop trips per launch
constexpr int UNR = 16;         // independent instructions per trip and thread

__global__ void k_idle(unsigned long long* out, int ns) {          // resident, sleeping
    for (int i = 0; i < ns; i++) __nanosleep(1000);
    if (threadIdx.x == 1025) out[0] = 1;
}
__global__ void k_ffma(float* out, float a, float b) {
    float x[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) x[j] = fmaf(x[j], a, b);
    }
    float s = 0; for (int j = 0; j < UNR; j++) s += x[j];
    if (s == 12345.678f) out[0] = s;
}
__global__ void k_fmul2(unsigned long long* out, unsigned long long a) {      // packed f32x2 multiply
    unsigned long long x[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) x[j] = a + threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[j]) : "l"(a));
    }
    unsigned long long s = 0; for (int j = 0; j < UNR; j++) s ^= x[j];
    if (s == 0x1234567ull) out[0] = s;
}
__global__ void k_dadd(double* out, double b) {
    double x[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) x[j] = __dadd_rn(x[j], b);
    }
    double s = 0; for (int j = 0; j < UNR; j++) s += x[j];
    if (s == 12345.678) out[0] = s;
}
__global__ void k_f2f(float* out) {                                 // float -> double -> float: two conversions per count
    float x[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS / 2; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(x[j])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x[j]) : "d"(d)); }
    }
    float s = 0; for (int j = 0; j < UNR; j++) s += x[j];
    if (s == 12345.678f) out[0] = s;
}
__global__ void k_imad(int* out, int a, int b) {
    int x[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) x[j] = x[j] * a + b;
    }
    int s = 0; for (int j = 0; j < UNR; j++) s ^= x[j];
    if (s == 0x12345677) out[0] = s;
}
__global__ void k_lds64(double* out) {                              // conflict-free 64-bit shared-memory reads
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double acc[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) acc[j] = 0;
    int idx = threadIdx.x & 31;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) { double v; asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + ((idx + 32 * j) & 1023))) : "memory"); acc[j] += v; }
        idx = (idx + 1) & 31;
    }
    double s = 0; for (int j = 0; j < UNR; j++) s += acc[j];
    if (s == 12345.678) out[0] = s;
}
__global__ void k_shfl(int* out) {
    int x[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNR; j++) x[j] = __shfl_up_sync(0xffffffffu, x[j], 1);
    }
    int s = 0; for (int j = 0; j < UNR; j++) s ^= x[j];
    if (s == 0x12345677) out[0] = s;
}
__global__ void k_copy(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_read(const float4* __restrict__ src, float* out, size_t n) {
    float s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = __ldcs(src + i); s += v.x + v.y + v.z + v.w; }
    if (s == 12345.678f) out[0] = s;
}

static nvmlDevice_t g_dev;
static double energy_j() { unsigned long long mj = 0; nvmlDeviceGetTotalEnergyConsumption(g_dev, &mj); return mj / 1e3; }

struct Result { double secs, watts, mhz; long long launches; };
template <class F> static Result run_for(double seconds, F launch) {
    std::vector<unsigned> clocks;
    bool go = true;
    std::thread mon([&] { while (go) { unsigned c = 0; if (nvmlDeviceGetClockInfo(g_dev, NVML_CLOCK_SM, &c) == NVML_SUCCESS) clocks.push_back(c); std::this_thread::sleep_for(std::chrono::milliseconds(5)); } });
    for (int i = 0; i < 3; i++) launch();
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    // let the power reading settle into the load, then measure
    auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 0.6) { for (int i = 0; i < 4; i++) launch(); CK(cudaDeviceSynchronize()); }
    clocks.clear();
    const double e0 = energy_j();
    t0 = std::chrono::steady_clock::now();
    long long n = 0;
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) { for (int i = 0; i < 4; i++) launch(); n += 4; CK(cudaDeviceSynchronize()); }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double e1 = energy_j();
    go = false; mon.join();
    std::sort(clocks.begin(), clocks.end());
    return {secs, (e1 - e0) / secs, clocks.empty() ? 0.0 : (double)clocks[clocks.size() / 2], n};
}

int main() {
    if (nvmlInit_v2() != NVML_SUCCESS || nvmlDeviceGetHandleByIndex_v2(0, &g_dev) != NVML_SUCCESS) { printf("no NVML\n"); return 1; }
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int grid = sms * 8, block = 256;                          // 64 warps per SM
    void* out; CK(cudaMalloc(&out, 4096));
    const size_t big = (size_t)4 << 30, small = (size_t)48 << 20;   // 4 GiB (HBM stream), 48 MiB (L2 resident)
    float4 *src, *dst; CK(cudaMalloc(&src, big)); CK(cudaMalloc(&dst, big)); CK(cudaMemset(src, 1, big)); CK(cudaMemset(dst, 0, big));
    const double per_launch_instr = (double)grid * (block / 32) * ITERS * UNR;     // warp instructions of the named kind
    std::this_thread::sleep_for(std::chrono::milliseconds(1500));
    const double e0 = energy_j(); std::this_thread::sleep_for(std::chrono::milliseconds(1000));
    const double idle_w = energy_j() - e0;
    printf("{\"probe\": \"idle (no kernel)\", \"watts\": %.1f}\n", idle_w);
    Result base = run_for(1.5, [&] { k_idle<<<grid, block>>>((unsigned long long*)out, 2000); });
    printf("{\"probe\": \"resident warps asleep (nanosleep)\", \"watts\": %.1f, \"sm_mhz\": %.0f}\n", base.watts, base.mhz);
    auto report = [&](const char* name, Result r, double units_per_launch, const char* unit) {
        const double rate = units_per_launch * r.launches / r.secs;
        printf("{\"probe\": \"%s\", \"watts\": %.1f, \"sm_mhz\": %.0f, \"%s_per_s\": %.4g, \"pJ_per_%s_above_asleep\": %.2f}\n",
               name, r.watts, r.mhz, unit, rate, unit, (r.watts - base.watts) / rate * 1e12);
        fflush(stdout);
    };
    report("FFMA", run_for(1.5, [&] { k_ffma<<<grid, block>>>((float*)out, 1.0001f, 0.5f); }), per_launch_instr, "warp_instr");
    report("FMUL2 (f32x2)", run_for(1.5, [&] { k_fmul2<<<grid, block>>>((unsigned long long*)out, 0x3f8000013f800001ull); }), per_launch_instr, "warp_instr");
    report("IMAD", run_for(1.5, [&] { k_imad<<<grid, block>>>((int*)out, 3, 7); }), per_launch_instr, "warp_instr");
    report("DADD", run_for(1.5, [&] { k_dadd<<<grid, block>>>((double*)out, 0.5); }), per_launch_instr, "warp_instr");
    report("F2F (f32 -> f64 -> f32, per conversion)", run_for(1.5, [&] { k_f2f<<<grid, block>>>((float*)out); }), per_launch_instr, "warp_instr");
    report("LDS.64", run_for(1.5, [&] { k_lds64<<<grid, block>>>((double*)out); }), per_launch_instr, "warp_instr");
    report("SHFL", run_for(1.5, [&] { k_shfl<<<grid, block>>>((int*)out); }), per_launch_instr, "warp_instr");
    report("HBM copy 4 GiB (read + write bytes)", run_for(1.5, [&] { k_copy<<<sms * 16, 512>>>(src, dst, big / 16); }), 2.0 * big, "byte");
    report("HBM read 4 GiB", run_for(1.5, [&] { k_read<<<sms * 16, 512>>>(src, (float*)out, big / 16); }), 1.0 * big, "byte");
    report("L2-resident read 48 MiB", run_for(1.5, [&] { k_read<<<sms * 16, 512>>>(src, (float*)out, small / 16); }), 1.0 * small, "byte");
    return 0;
}
