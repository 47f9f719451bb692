// pskd_kernels.cu -- sm_100a kernels of the PSK soft-demod path.
//
//   k_front      ingest + symbol timing + M-th power angle     (cpp/psk_soft.cpp:442-474, 568-584, 619-636)
//   k_chain_seq  unwrap + LinearFit recursion, packet prologue/epilogue, one thread per channel
//                                                               (cpp/psk_soft.cpp:380-426, 476-482, 592-603, 35-185)
//   k_back       differential decode / derotate / slice / bits  (cpp/psk_soft.cpp:484-566)
//   k_finish     carry the window tail and `last` into the next call
//
// Data layout: see DESIGN.md.  All arithmetic that decides an integer output follows the
// reference's rounding order through the intrinsics of pskd_exact.cuh.
#include "pskd_internal.h"

namespace pskd {

const char* kernel_name(int kid) {
    static const char* names[KID_COUNT] = {"k_front", "k_chain_seq", "k_chain_spec", "k_chain_scan", "k_chain_exact", "k_back", "k_finish"};
    return (kid >= 0 && kid < KID_COUNT) ? names[kid] : "?";
}
cudaEvent_t Profiler::get() {
    if (n_pool > 0) return pool[--n_pool];
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void Profiler::begin(int kid, cudaStream_t s) {
    if (!enabled) return;
    if (n_pending == cap_pending) {
        int nc = cap_pending ? 2 * cap_pending : 64;
        Pair* np = new Pair[nc];
        for (int i = 0; i < n_pending; i++) np[i] = pending[i];
        delete[] pending; pending = np; cap_pending = nc;
    }
    Pair& p = pending[n_pending];
    p.a = get(); p.b = get(); p.kid = kid;
    cudaEventRecord(p.a, s);
}
void Profiler::end(cudaStream_t s) {
    if (!enabled) return;
    cudaEventRecord(pending[n_pending].b, s);
    n_pending++;
}
void Profiler::drain() {
    for (int i = 0; i < n_pending; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pending[i].a, pending[i].b) == cudaSuccess) { ms[pending[i].kid] += t; launches[pending[i].kid]++; }
        for (cudaEvent_t e : {pending[i].a, pending[i].b}) {
            if (n_pool == cap_pool) {
                int nc = cap_pool ? 2 * cap_pool : 128;
                cudaEvent_t* np = new cudaEvent_t[nc];
                for (int j = 0; j < n_pool; j++) np[j] = pool[j];
                delete[] pool; pool = np; cap_pool = nc;
            }
            pool[n_pool++] = e;
        }
    }
    n_pending = 0;
}
void Profiler::destroy() {
    drain();
    for (int i = 0; i < n_pool; i++) cudaEventDestroy(pool[i]);
    delete[] pool; delete[] pending; pool = nullptr; pending = nullptr; n_pool = cap_pool = cap_pending = 0;
}

// virtual stream = tail ++ in
struct VStream {
    const float2* tail; const float2* in; long long tail_len;
    __device__ __forceinline__ float2 at(long long i) const {
        return (i < tail_len) ? tail[i] : __ldg(in + (i - tail_len));
    }
};

// ---------------------------------------------------------------------------------------------
// k_front: one CTA = one tile of FT output symbols of one channel.
//   pass 1: per (phase p, run r) thread: energies e = f32(re^2+im^2) -> double inclusive prefix
//           along the symbol axis, stored in shared memory pre[row][p]  (row stride SP = S|1
//           doubles so both passes are bank-conflict free for 64-bit accesses)
//   pass 2: per output symbol: E_p[k] = pre[k+A-1] - pre[k-1] for every p, FIRST maximum
//           (std::max_element, cpp/psk_soft.cpp:462), gather the chosen sample of the OLDEST
//           symbol in the window (:465), M-th power angle (:474).
// The window sums are differences of double prefix sums of float-exact energies: identical to
// the reference's add/subtract sliding sums whenever those are exact (DESIGN.md "exactness").
// ---------------------------------------------------------------------------------------------
constexpr int FT = 512;          // output symbols per tile
constexpr int FRONT_THREADS = 256;

__global__ void __launch_bounds__(FRONT_THREADS)
k_front(const ChanDesc* __restrict__ desc, int16_t* __restrict__ out_sidx,
        float2* __restrict__ sel, float* __restrict__ theta)
{
    const ChanDesc& d = desc[blockIdx.y];
    const long long k0 = (long long)blockIdx.x * FT;
    if (k0 >= d.K) return;
    const int S = d.S, A = d.A, M = d.M;
    const int nk = (int)min((long long)FT, d.K - k0);
    const int nsym = nk + A - 1;                 // input symbols [k0, k0+nsym)
    const int SP = S | 1;
    extern __shared__ double smem[];
    double* pre = smem;                          // [(nsym+1)][SP], row 0 = 0
    const int nruns = FRONT_THREADS / S;
    double* tot = pre + (size_t)(FT + A) * SP;   // [nruns][SP] run totals -> exclusive offsets

    VStream vs{d.tail, d.in, d.tail_len};
    const int tid = threadIdx.x;
    const int r = tid / S, p = tid - r * S;
    const int R = (nsym + nruns - 1) / nruns;
    if (tid < SP) pre[tid] = 0.0;
    if (r < nruns) {
        int j0 = r * R, j1 = min(j0 + R, nsym);
        double acc = 0.0;
        long long base = (k0 + j0) * S + p;
        int j = j0;
        for (; j + 4 <= j1; j += 4) {            // 4 independent loads in flight
            float2 x0 = vs.at(base), x1 = vs.at(base + S), x2 = vs.at(base + 2 * S), x3 = vs.at(base + 3 * S);
            base += 4 * S;
            acc = daddr(acc, (double)energy_f32(x0.x, x0.y)); pre[(size_t)(j + 1) * SP + p] = acc;
            acc = daddr(acc, (double)energy_f32(x1.x, x1.y)); pre[(size_t)(j + 2) * SP + p] = acc;
            acc = daddr(acc, (double)energy_f32(x2.x, x2.y)); pre[(size_t)(j + 3) * SP + p] = acc;
            acc = daddr(acc, (double)energy_f32(x3.x, x3.y)); pre[(size_t)(j + 4) * SP + p] = acc;
        }
        for (; j < j1; j++) {
            float2 x = vs.at(base); base += S;
            acc = daddr(acc, (double)energy_f32(x.x, x.y)); pre[(size_t)(j + 1) * SP + p] = acc;
        }
        tot[r * SP + p] = acc;
    }
    __syncthreads();
    if (tid < S) {                               // exclusive scan of the run totals, per phase
        double run = 0.0;
        for (int rr = 0; rr < nruns; rr++) {
            double t = tot[rr * SP + tid];
            tot[rr * SP + tid] = run;
            run = daddr(run, t);
        }
    }
    __syncthreads();
    for (int t = tid; t < nk; t += FRONT_THREADS) {
        const int ja = t + A - 1;                // newest symbol of the window (0-based row ja+1)
        const int ra = ja / R;
        const int rb = (t > 0) ? (t - 1) / R : 0;
        const double* pa = pre + (size_t)(ja + 1) * SP;
        const double* pb = pre + (size_t)t * SP;
        const double* oa = tot + ra * SP;
        const double* ob = tot + rb * SP;
        double best = 0.0; int idx = 0;
        for (int q = 0; q < S; q++) {
            double a = daddr(pa[q], oa[q]);
            double b = (t > 0) ? daddr(pb[q], ob[q]) : 0.0;
            double E = dsubr(a, b);
            if (q == 0) best = E;
            else if (best < E) { best = E; idx = q; }
        }
        const long long k = k0 + t;
        float2 s = vs.at(k * S + idx);
        out_sidx[d.sym_off + k] = (int16_t)idx;
        sel[d.scr_off + k] = s;
        theta[d.scr_off + k] = mth_power_angle(s, (unsigned)M);
    }
}

cudaError_t launch_front(const LaunchCtx& c) {
    if (c.Kmax <= 0) return cudaSuccess;
    int SPmax = c.Smax | 1;
    size_t smem = ((size_t)(FT + c.Amax) * SPmax + (size_t)(FRONT_THREADS / 2) * SPmax) * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_front, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid((unsigned)((c.Kmax + FT - 1) / FT), (unsigned)c.n_channels);
    c.prof->begin(KID_FRONT, c.stream);
    k_front<<<grid, FRONT_THREADS, smem, c.stream>>>(c.d_desc, c.out_sidx, c.d_sel, c.d_theta);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// y-history ring of one channel in global memory
// ---------------------------------------------------------------------------------------------
struct GlobalRing {
    float* base;
    __device__ __forceinline__ float get(int i) const { return base[i]; }
    __device__ __forceinline__ void set(int i, float v) const { base[i] = v; }
    // keep the newest `keep` of the `pts` values that start at `head` (modulo oldn), re-packed at 0..keep
    __device__ void repack(int head, int oldn, int drop, int keep, int newn) const {
        // in place, front to back is safe only when the data does not wrap; use a two-step copy via
        // the spare half of the ring allocation (the bank allocates 2*ring_cap floats per channel)
        float* spare = base + (oldn > newn ? oldn : newn);
        int idx = head + drop; idx %= (oldn > 0 ? oldn : 1);
        for (int j = 0; j < keep; j++) { spare[j] = base[idx]; if (++idx == oldn) idx = 0; }
        for (int j = 0; j < keep; j++) base[j] = spare[j];
    }
};

// packet prologue for the phase estimator (cpp/psk_soft.cpp:393-426).  The SRI block runs on
// every packet because `numDataPts > samples.size()` holds at every packet start (:380-383).
__device__ __forceinline__ void chain_packet_prologue(ChanState& st, GlobalRing ring, const ChanDesc& d,
                                                       double sri_xdelta, int& flags) {
    if (sri_xdelta != (double)st.sampleRate) {                                         // :394-398
        st.sampleRate = __double2float_rn(__ddiv_rn(1.0, sri_xdelta));
        fit_reset(st.fit, ring, nullptr, &st.sampleRate, false);
    }
    if (flags & CH_RESET_NUMSYMS) {                                                    // :416-420
        fit_reset(st.fit, ring, nullptr, nullptr, true);
        flags &= ~CH_RESET_NUMSYMS;
    }
    if (flags & CH_RESET_PHASEAVG) {                                                   // :421-426
        int numPts = d.P;
        fit_reset(st.fit, ring, &numPts, nullptr, false);
        flags &= ~CH_RESET_PHASEAVG;
    }
}

// packet epilogue: wrap the estimate back (cpp/psk_soft.cpp:592-603)
__device__ __forceinline__ void chain_packet_epilogue(ChanState& st, GlobalRing ring, int M) {
    float wrapValue = __double2float_rn(dmulr(PSKD_M_2PI, (double)M));
    if (wrap_needed(st.est, wrapValue)) {
        float q = __fdiv_rn(st.est, wrapValue);
        float nw = roundf(q);                                                          // :598
        long long numWraps = (long long)nw;
        st.est = fit_subtract_const(st.fit, ring, fmulr((float)numWraps, wrapValue));  // :601-602
        st.wraps++;
    }
}

// one thread per channel: the reference's recursion, literally, over every emulated packet.
// This is the generic / fallback chain; the speculative parallel chain lives in pskd_chain.cuh.
__global__ void k_chain_seq(const ChanDesc* __restrict__ desc, ChanState* __restrict__ state,
                            float* __restrict__ ring_base, const float* __restrict__ theta,
                            float* __restrict__ out_phase, double sri_xdelta, int n_channels,
                            DevCounters* counters)
{
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= n_channels) return;
    const ChanDesc d = desc[ch];
    ChanState st = state[ch];
    GlobalRing ring{ring_base + d.ring_off};
    int flags = d.flags;
    const float* th = theta + d.scr_off;
    float* ph = out_phase + d.sym_off;
    unsigned long long wraps0 = st.wraps;
    for (int j = 0; j < d.n_pkts; j++) {
        chain_packet_prologue(st, ring, d, sri_xdelta, flags);
        long long klo = first_symbol_at((long long)j * d.pkt_len, d.tail_len, d.S, d.A, d.K);
        long long khi = (j + 1 == d.n_pkts) ? d.K
                        : first_symbol_at((long long)(j + 1) * d.pkt_len, d.tail_len, d.S, d.A, d.K);
        for (long long k = klo; k < khi; k++) {
            float y = unwrap_against(st.est, th[k], nullptr);                          // :474-478
            st.est = fit_next(st.fit, ring, y);                                        // :481
            ph[k] = st.est;                                                            // :482
        }
        chain_packet_epilogue(st, ring, d.M);
    }
    state[ch] = st;
    if (st.wraps != wraps0) atomicAdd(&counters->wraps, st.wraps - wraps0);
    atomicAdd(&counters->seq_channels, 1ULL);
}

cudaError_t launch_chain_seq(const LaunchCtx& c) {
    int threads = 32;
    int blocks = (c.n_channels + threads - 1) / threads;
    float* phase = c.out_phase ? c.out_phase : c.d_phase_tmp;
    c.prof->begin(KID_CHAIN_SEQ, c.stream);
    k_chain_seq<<<blocks, threads, 0, c.stream>>>(c.d_desc, c.d_state, c.d_ring, c.d_theta, phase,
                                                  c.sri_xdelta, c.n_channels, c.d_counters);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_back: per symbol, differential decode or derotate, then slice (cpp/psk_soft.cpp:484-566)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_back(const ChanDesc* __restrict__ desc, const ChanState* __restrict__ state,
       const float2* __restrict__ sel, const float* __restrict__ phase,
       float2* __restrict__ out_soft, int16_t* __restrict__ out_bits)
{
    const ChanDesc& d = desc[blockIdx.y];
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d.K) return;
    float2 s = sel[d.scr_off + k];
    float2 sample = s;
    if (d.D) {
        float2 last = (k > 0) ? sel[d.scr_off + k - 1] : state[blockIdx.y].last;
        sample = cdiv_f32(s, last);                                                    // :488
    }
    float est = d.D ? 0.0f : phase[d.sym_off + k];
    float pc = phase_correction(est, d.M, d.D != 0);
    float2 c = derotate(sample, pc);
    if (out_soft) out_soft[d.sym_off + k] = c;
    if (out_bits && d.bpb) {
        unsigned b = slice_bits(c, d.bpb);
        int16_t* o = out_bits + d.bits_off + k * d.bpb;
        for (int j = 0; j < d.bpb; j++) o[j] = (int16_t)((b >> j) & 1u);
    }
}

cudaError_t launch_back(const LaunchCtx& c) {
    if (c.Kmax <= 0) return cudaSuccess;
    if (!c.out_soft && !c.out_bits) return cudaSuccess;
    dim3 grid((unsigned)((c.Kmax + 255) / 256), (unsigned)c.n_channels);
    const float* phase = c.out_phase ? c.out_phase : c.d_phase_tmp;
    c.prof->begin(KID_BACK, c.stream);
    k_back<<<grid, 256, 0, c.stream>>>(c.d_desc, c.d_state, c.d_sel, phase, (float2*)c.out_soft, c.out_bits);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// k_finish: one CTA per channel. Carry `last` and the un-consumed tail of the virtual stream.
// ---------------------------------------------------------------------------------------------
__global__ void k_finish(const ChanDesc* __restrict__ desc, ChanState* __restrict__ state,
                         const float2* __restrict__ sel)
{
    const ChanDesc& d = desc[blockIdx.x];
    if (threadIdx.x == 0 && d.K > 0 && d.D) state[blockIdx.x].last = sel[d.scr_off + d.K - 1];   // :489
    VStream vs{d.tail, d.in, d.tail_len};
    const long long start = d.K * d.S;
    for (long long i = threadIdx.x; i < d.next_tail_len; i += blockDim.x)
        d.tail_next[i] = vs.at(start + i);
}

cudaError_t launch_finish(const LaunchCtx& c) {
    c.prof->begin(KID_FINISH, c.stream);
    k_finish<<<c.n_channels, 128, 0, c.stream>>>(c.d_desc, c.d_state, c.d_sel);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

}  // namespace pskd
