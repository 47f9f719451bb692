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
    static const char* names[KID_COUNT] = {"k_front", "k_chain_seq", "k_chain_par", "k_chain_scan", "k_chain_exact", "k_back", "k_finish"};
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
template <class Ring>
__device__ __forceinline__ void chain_packet_prologue(ChanState& st, Ring ring, const ChanDesc& d,
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
template <class Ring>
__device__ __forceinline__ void chain_packet_epilogue(ChanState& st, Ring ring, int M) {
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
    if (d.flags & CH_FAST) return;              // handled by k_chain_par
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
    if (c.n_seq_channels == 0) return cudaSuccess;
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
    if (d.flags & CH_FAST) return;              // k_chain_par derotates/slices its own channels
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
    if (c.n_seq_channels == 0) return cudaSuccess;
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

// =============================================================================================
// k_chain_par: scan-based phase chain + derotate/slice, one CTA per channel.
//
// The reference's per-symbol recursion (cpp/psk_soft.cpp:474-482 + LinearFit::next :48-87)
//     n_k = round((est_{k-1} - theta_k)/2pi);  y_k = f32(theta_k + 2pi n_k);  est_k = fit(y_{k-P+1..k})
// is sequential only through the INTEGER n_k.  Per sub-block of CP_B symbols this kernel
//   1. predicts n_k by classic sample-to-sample unwrapping (an integer prefix sum), anchored on
//      the exact rule for the first symbol of the block,
//   2. evaluates y_k, the window sums ySum (differences of double prefix sums) and the xySum
//      recurrence X_k = X_{k-1} - xdelta*ySum'_k + T_k (another double prefix sum, T_k rounded in
//      float exactly as :78) and from them est_k with the reference's own rounding (:157-162),
//   3. VERIFIES every n_k against the reference's rule using est_{k-1}; on the first mismatch it
//      shifts the remaining predictions by the observed difference and repeats (each pass
//      proves at least one more symbol), falling back to the literal sequential recursion for
//      that sub-block after CP_MAX_ITERS passes.
// So the emitted n_k are exactly those of the sequential recursion run on the same theta_k; the
// double sums are formed in scan order instead of symbol order (<= 1e-15 relative apart, far
// below the float rounding of m,b; see DESIGN.md "phase chain").  Fill-up (fewer than P points),
// the 2^20-call re-sum (:51-52) and packet prologue/epilogue run on one thread, literally.
// =============================================================================================
constexpr int CP_THREADS = 256;
constexpr int CP_V = 4;
constexpr int CP_B = CP_THREADS * CP_V;
constexpr int CP_MAX_ITERS = 24;

struct SmemRing {
    float* base;
    __device__ __forceinline__ float get(int i) const { return base[i]; }
    __device__ __forceinline__ void set(int i, float v) const { base[i] = v; }
    __device__ void repack(int, int, int, int, int) const {}   // never called: P changes take the sequential chain
};

struct CpShared {
    ChanState st;
    int flags;
    int mis;          // first mismatching symbol of the pass
    int delta;        // correction to add to n_i for i >= mis
    int mode;         // 0 parallel, 1 sequential
    unsigned int passes, seq_blocks;
};

__device__ __forceinline__ int warp_scan_int(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    return v;
}
__device__ __forceinline__ double warp_scan_dbl(double v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { double u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = daddr(v, u); }
    return v;
}
// exclusive offset of this thread's partial `tot` over the CTA (CP_THREADS threads); wtot: smem[8]
__device__ __forceinline__ int block_excl_int(int tot, int* wtot, int tid) {
    int lane = tid & 31, w = tid >> 5;
    int inc = warp_scan_int(tot, lane);
    if (lane == 31) wtot[w] = inc;
    __syncthreads();
    int off = 0;
#pragma unroll
    for (int i = 0; i < CP_THREADS / 32; i++) if (i < w) off += wtot[i];
    __syncthreads();
    return off + inc - tot;
}
__device__ __forceinline__ double block_excl_dbl(double tot, double* wtot, int tid) {
    int lane = tid & 31, w = tid >> 5;
    double inc = warp_scan_dbl(tot, lane);
    if (lane == 31) wtot[w] = inc;
    __syncthreads();
    double off = 0.0;
#pragma unroll
    for (int i = 0; i < CP_THREADS / 32; i++) if (i < w) off = daddr(off, wtot[i]);
    __syncthreads();
    return daddr(off, dsubr(inc, tot));
}

// the reference's unwrap count for one symbol (cpp/psk_soft.cpp:477)
__device__ __forceinline__ int unwrap_count(float est_prev, float theta) {
    double q = __ddiv_rn(dsubr((double)est_prev, (double)theta), PSKD_M_2PI);
    return (int)(long long)round(q);
}

__global__ void __launch_bounds__(CP_THREADS)
k_chain_par(const ChanDesc* __restrict__ desc, ChanState* __restrict__ state, float* __restrict__ ring_base,
            const float* __restrict__ theta, const float2* __restrict__ sel,
            float* __restrict__ out_phase, float2* __restrict__ out_soft, int16_t* __restrict__ out_bits,
            double sri_xdelta, int Pcap, DevCounters* counters)
{
    const ChanDesc& d = desc[blockIdx.x];
    if (!(d.flags & CH_FAST)) return;
    const int tid = threadIdx.x;
    const int P = d.P;

    extern __shared__ double smem_d[];
    double* ps   = smem_d;                    // [CP_B]      inclusive prefix of the block's y
    double* hs   = ps + CP_B;                 // [Pcap + 1]  suffix sums of the history
    double* wtot = hs + Pcap + 1;             // [8]
    float*  yb   = (float*)(wtot + 8);        // [Pcap + CP_B] history (ring, then logical order) ++ block y
    float*  th   = yb + Pcap + CP_B;          // [CP_B]
    float*  estv = th + CP_B;                 // [CP_B + 1]  estv[0] = estimate before the block
    int*    wtoti = (int*)(estv + CP_B + 1);  // [8]
    __shared__ CpShared sh;

    float* gring = ring_base + d.ring_off;
    if (tid == 0) { sh.st = state[blockIdx.x]; sh.flags = d.flags; sh.passes = 0; sh.seq_blocks = 0; }
    for (int j = tid; j < P; j += CP_THREADS) yb[j] = gring[j];
    __syncthreads();
    SmemRing ring{yb};
    const unsigned long long wraps0 = sh.st.wraps;
    const float* thg = theta + d.scr_off;
    const float2* selg = sel + d.scr_off;
    const int M = d.M, bpb = d.bpb;
    const bool diff = d.D != 0;

    for (int pkt = 0; pkt < d.n_pkts; pkt++) {
        if (tid == 0) chain_packet_prologue(sh.st, ring, d, sri_xdelta, sh.flags);
        const long long klo = first_symbol_at((long long)pkt * d.pkt_len, d.tail_len, d.S, d.A, d.K);
        const long long khi = (pkt + 1 == d.n_pkts) ? d.K
                              : first_symbol_at((long long)(pkt + 1) * d.pkt_len, d.tail_len, d.S, d.A, d.K);
        __syncthreads();
        long long k = klo;
        while (k < khi) {
            // ---- choose the sub-block and its mode (uniform: every thread reads the same shared state)
            if (tid == 0) {
                FitState& f = sh.st.fit;
                if (f.count == 1048576 && f.pts == f.n) fit_resum(f, ring);            // :51-52 at a block edge
                sh.mode = (f.pts < f.n) ? 1 : 0;
                sh.mis = 0x7fffffff;
            }
            __syncthreads();
            int nb = (int)min((long long)CP_B, khi - k);
            int mode = sh.mode;
            if (mode == 0) nb = min(nb, 1048576 - sh.st.fit.count);
            else nb = min(nb, max(1, sh.st.fit.n - sh.st.fit.pts));                     // only the fill-up runs sequentially
            // ---- stage theta
            for (int i = tid; i < nb; i += CP_THREADS) th[i] = thg[k + i];
            if (tid == 0) estv[0] = sh.st.est;
            // ---- make the history logical (head == 0) for the scan path
            if (mode == 0 && sh.st.fit.head != 0) {
                const int head = sh.st.fit.head;
                __syncthreads();
                for (int j = tid; j < P; j += CP_THREADS) { int s2 = head + j; if (s2 >= P) s2 -= P; estv[1 + j] = yb[s2]; }  // P <= CP_B
                __syncthreads();
                for (int j = tid; j < P; j += CP_THREADS) yb[j] = estv[1 + j];
                if (tid == 0) sh.st.fit.head = 0;
            }
            __syncthreads();

            if (mode == 0) {
                // history suffix sums hs[j] = sum_{m=j}^{P-1} yb[m]  (scan over the reversed history)
                {
                    double loc[CP_V]; double run = 0.0;
#pragma unroll
                    for (int v = 0; v < CP_V; v++) {
                        int r = tid * CP_V + v;                   // reversed index: element P-1-r
                        double x = (r < P) ? (double)yb[P - 1 - r] : 0.0;
                        run = daddr(run, x); loc[v] = run;
                    }
                    double off = block_excl_dbl(run, wtot, tid);
#pragma unroll
                    for (int v = 0; v < CP_V; v++) { int r = tid * CP_V + v; if (r < P) hs[P - 1 - r] = daddr(off, loc[v]); }
                    if (tid == 0) hs[P] = 0.0;
                }
                // classic-unwrap prediction of n (integer prefix sum), first symbol by the exact rule
                int nloc[CP_V];
                {
                    int run = 0;
#pragma unroll
                    for (int v = 0; v < CP_V; v++) {
                        int i = tid * CP_V + v; int dn = 0;
                        if (i < nb) {
                            if (i == 0) dn = unwrap_count(estv[0], th[0]);
                            else dn = -__float2int_rn((th[i] - th[i - 1]) * 0.15915494309189535f);
                        }
                        run += dn; nloc[v] = run;
                    }
                    int off = block_excl_int(run, wtoti, tid);
#pragma unroll
                    for (int v = 0; v < CP_V; v++) nloc[v] += off;
                }
                const float xdelta = sh.st.fit.xdelta;
                const double xd = (double)xdelta;
                const float fP1 = (float)(P - 1);
                const double X0 = sh.st.fit.xySum;
                FitState fc = sh.st.fit;                          // constants: n, pts, denominator, xAvg, xdelta
                int iter = 0;
                bool ok = false;
                float yl[CP_V], el[CP_V]; double Yl[CP_V], Xl[CP_V];
                while (true) {
                    // y_k and the block prefix
                    double run = 0.0;
#pragma unroll
                    for (int v = 0; v < CP_V; v++) {
                        int i = tid * CP_V + v;
                        float y = 0.0f;
                        if (i < nb) {
                            y = __double2float_rn(daddr((double)th[i], dmulr((double)nloc[v], PSKD_M_2PI)));   // :478,481
                            yb[P + i] = y;
                        }
                        yl[v] = y; run = daddr(run, (double)y); Yl[v] = run;
                    }
                    double off = block_excl_dbl(run, wtot, tid);
#pragma unroll
                    for (int v = 0; v < CP_V; v++) { int i = tid * CP_V + v; if (i < nb) ps[i] = daddr(off, Yl[v]); }
                    __syncthreads();
                    // window sums, xySum increments
                    double trun = 0.0;
#pragma unroll
                    for (int v = 0; v < CP_V; v++) {
                        int i = tid * CP_V + v;
                        double t = 0.0;
                        if (i < nb) {
                            double W = (i + 1 <= P - 1) ? hs[i + 1] : 0.0;                 // history part of y_{k-P+1..k-1}
                            double blk = (i >= 1) ? ps[i - 1] : 0.0;
                            if (i - P >= 0) blk = dsubr(blk, ps[i - P]);
                            W = daddr(W, blk);                                             // ySum after :70
                            double a = dmulr(xd, W);                                       // :72
                            double T = (double)fmulr(fmulr(yl[v], fP1), xdelta);           // :78
                            t = dsubr(T, a);
                            Yl[v] = daddr(W, (double)yl[v]);                               // :75
                        }
                        trun = daddr(trun, t); Xl[v] = trun;
                    }
                    double xoff = block_excl_dbl(trun, wtot, tid);
#pragma unroll
                    for (int v = 0; v < CP_V; v++) {
                        int i = tid * CP_V + v;
                        if (i < nb) {
                            fc.ySum = Yl[v];
                            fc.xySum = daddr(X0, daddr(xoff, Xl[v]));
                            Xl[v] = fc.xySum;
                            el[v] = fit_calc_fit(fc, yl[v]);                               // :135-162
                            estv[i + 1] = el[v];
                        }
                    }
                    __syncthreads();
                    // verify every predicted n against the reference's rule (:477)
#pragma unroll
                    for (int v = 0; v < CP_V; v++) {
                        int i = tid * CP_V + v;
                        if (i >= 1 && i < nb) {
                            int nt = unwrap_count(estv[i], th[i]);
                            if (nt != nloc[v]) atomicMin(&sh.mis, i);
                        }
                    }
                    __syncthreads();
                    const int mis = sh.mis;
                    if (mis == 0x7fffffff) { ok = true; break; }
                    if (++iter > CP_MAX_ITERS) break;
                    if (mis / CP_V == tid) {
                        int nv = 0;
#pragma unroll
                        for (int v = 0; v < CP_V; v++) if (v == mis % CP_V) nv = nloc[v];
                        sh.delta = unwrap_count(estv[mis], th[mis]) - nv;
                    }
                    __syncthreads();
                    const int delta = sh.delta;
#pragma unroll
                    for (int v = 0; v < CP_V; v++) if (tid * CP_V + v >= mis) nloc[v] += delta;
                    if (tid == 0) sh.mis = 0x7fffffff;
                    __syncthreads();
                }
                if (tid == 0) sh.passes += (unsigned)iter;
                if (ok) {
                    // commit: sums/estimate after the last symbol, history = last P of (history ++ block)
                    const int last = nb - 1;
                    if (last / CP_V == tid) {
                        double Yv = 0.0, Xv = 0.0; float yv = 0.0f, ev = 0.0f;
#pragma unroll
                        for (int v = 0; v < CP_V; v++) if (v == last % CP_V) { Yv = Yl[v]; Xv = Xl[v]; yv = yl[v]; ev = el[v]; }
                        FitState& f = sh.st.fit;
                        fc.ySum = Yv; fc.xySum = Xv;
                        (void)fit_calc_fit(fc, yv);
                        f.ySum = Yv; f.xySum = Xv; f.m = fc.m; f.b = fc.b; f.count += nb;
                        sh.st.est = ev;
                    }
                    float keep[(CHAIN_PAR_PMAX + CP_THREADS - 1) / CP_THREADS];
#pragma unroll
                    for (int q = 0; q < (CHAIN_PAR_PMAX + CP_THREADS - 1) / CP_THREADS; q++) {
                        int j = tid + q * CP_THREADS;
                        keep[q] = (j < P) ? yb[nb + j] : 0.0f;
                    }
                    __syncthreads();
#pragma unroll
                    for (int q = 0; q < (CHAIN_PAR_PMAX + CP_THREADS - 1) / CP_THREADS; q++) {
                        int j = tid + q * CP_THREADS;
                        if (j < P) yb[j] = keep[q];
                    }
                    __syncthreads();
                } else {
                    mode = 1;                                   // too many passes: literal recursion for this block
                    if (tid == 0) sh.seq_blocks++;
                    __syncthreads();
                }
            }
            if (mode == 1) {
                if (tid == 0) {
                    ChanState& st = sh.st;
                    for (int i = 0; i < nb; i++) {
                        float y = unwrap_against(st.est, th[i], nullptr);
                        st.est = fit_next(st.fit, ring, y);
                        estv[i + 1] = st.est;
                    }
                }
                __syncthreads();
            }
            // ---- outputs for symbols k .. k+nb: phase, then derotate / slice (cpp/psk_soft.cpp:482-566)
            for (int i = tid; i < nb; i += CP_THREADS) {
                const long long kk = k + i;
                const float est = estv[i + 1];
                if (out_phase) out_phase[d.sym_off + kk] = est;
                if (out_soft || (out_bits && bpb)) {
                    float2 s = selg[kk];
                    if (diff) {
                        float2 last = (kk > 0) ? selg[kk - 1] : sh.st.last;
                        s = cdiv_f32(s, last);
                    }
                    float pc = phase_correction(est, M, diff);
                    float2 c = derotate(s, pc);
                    if (out_soft) out_soft[d.sym_off + kk] = c;
                    if (out_bits && bpb) {
                        unsigned b = slice_bits(c, bpb);
                        int16_t* o = out_bits + d.bits_off + kk * bpb;
                        for (int j = 0; j < bpb; j++) o[j] = (int16_t)((b >> j) & 1u);
                    }
                }
            }
            __syncthreads();
            k += nb;
        }
        if (tid == 0) chain_packet_epilogue(sh.st, ring, M);
        __syncthreads();
    }
    // ---- store the carried state (ring in whatever rotation it has; head says where it starts)
    for (int j = tid; j < P; j += CP_THREADS) gring[j] = yb[j];
    if (tid == 0) {
        state[blockIdx.x] = sh.st;      // `last` is carried by k_finish
        if (sh.st.wraps != wraps0) atomicAdd(&counters->wraps, sh.st.wraps - wraps0);
        atomicAdd(&counters->spec_chunks, (unsigned long long)((d.K + CP_B - 1) / CP_B));
        if (sh.passes) atomicAdd(&counters->spec_misses, (unsigned long long)sh.passes);
        if (sh.seq_blocks) atomicAdd(&counters->seq_channels, (unsigned long long)sh.seq_blocks);
    }
}

cudaError_t launch_chain_par(const LaunchCtx& c) {
    if (c.n_fast_channels == 0) return cudaSuccess;
    int Pcap = c.Pmax_fast < 1 ? 1 : c.Pmax_fast;
    size_t smem = (size_t)(CP_B + Pcap + 1 + 8) * sizeof(double)
                + (size_t)(Pcap + CP_B + CP_B + CP_B + 1) * sizeof(float) + 8 * sizeof(int) + 16;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_chain_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    float* phase = c.out_phase ? c.out_phase : c.d_phase_tmp;
    c.prof->begin(KID_CHAIN_PAR, c.stream);
    k_chain_par<<<c.n_channels, CP_THREADS, smem, c.stream>>>(c.d_desc, c.d_state, c.d_ring, c.d_theta, c.d_sel,
                                                            phase, (float2*)c.out_soft, c.out_bits, c.sri_xdelta, Pcap, c.d_counters);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

}  // namespace pskd
