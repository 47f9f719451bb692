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
#include "pskd_device.cuh"
#include <cstdlib>

namespace pskd {

const char* kernel_name(int kid) {
    static const char* names[KID_COUNT] = {"k_front", "k_chain_seq", "k_chain_par", "k_back_par", "k_chain_exact", "k_back", "k_finish", "k_fused", "k_tp_aux", "k_fzs_front", "k_fzs_cb", "k_fused_s9", "k_fused_s10", "k_fused_s16"};
    return (kid >= 0 && kid < KID_COUNT) ? names[kid] : "?";
}
cudaEvent_t Profiler::get() {
    if (n_pool > 0) return pool[--n_pool];
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void Profiler::begin(int kid, cudaStream_t s, double alg_bytes) {
    if (!enabled) return;
    bytes[kid] += alg_bytes;
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
        float2* __restrict__ sel, float* __restrict__ theta, unsigned long long S_mask)
{
    const ChanDesc& d = desc[blockIdx.y];
    if (d.flags & (CH_FRONT_FAST | CH_FUSED | CH_FZS)) return;   // handled by the specialised / fused kernels
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

// ---------------------------------------------------------------------------------------------
// k_front_t<S>: specialised on samplesPerBaud.  One CTA (FT_THREADS threads) = one tile of
// FT_ROWS = FT_THREADS*8 input symbols of one channel:
//   stage : coalesced 128-bit loads of the IQ samples, e = f32(re^2+im^2) (std::norm<float>,
//           cpp/psk_soft.cpp:448) stored as FLOAT rows e[j][0..S) in shared memory
//   sums  : thread r owns the run of 8 rows 8r..8r+7 (kept in registers).  Block sums per phase
//           (double) -> prefix over runs (warp shuffles + one shared row per warp) -> window sum
//           of the first symbol of the run = difference of two run prefixes + (A mod 8) rows,
//           then the window slides one symbol at a time: + newest row, - oldest (own) row
//           (cpp/psk_soft.cpp:451, 576).  All sums are doubles of float-exact energies.
//   pick  : first maximum over the S phases (:462), gather that sample of the oldest symbol
//           (:465), M-th power angle (:474); writes sampleIndex, sample, theta.
// Shared-memory traffic is ~20 B/sample (float rows) instead of ~40 B/sample for a double prefix
// table, and every global load is a fully coalesced LDG.128 -- the two things ncu showed the
// generic kernel was bound by (profiles/r01b).  Rows are padded by 4 words per run so the
// quarter-warp row reads (stride 8 rows) hit distinct bank groups.
// ---------------------------------------------------------------------------------------------
constexpr int FT_THREADS = 128;
constexpr int FT_R = 8;
constexpr int FT_ROWS = FT_THREADS * FT_R;       // 1024 input symbols per tile

template <int S> struct FrontCfg {
    static constexpr int SE = (S + 3) & ~3;                       // row stride in floats (16-byte rows)
    static constexpr int RUNW = FT_R * SE + 4;                    // words per run incl. the 4-word pad
    static constexpr int E_WORDS = FT_THREADS * RUNW;
    static constexpr int PB_DOUBLES = (FT_THREADS / 32) * S;      // per-warp totals
    static constexpr int PBR_DOUBLES = FT_THREADS * S;            // inclusive run prefixes
    static constexpr size_t SMEM = (size_t)E_WORDS * 4 + (size_t)(PB_DOUBLES + PBR_DOUBLES) * 8;
};

#ifndef PSKD_FT_MIN_CTAS
#define PSKD_FT_MIN_CTAS 5
#endif
template <int S>
__global__ void __launch_bounds__(FT_THREADS, PSKD_FT_MIN_CTAS)
k_front_t(const ChanDesc* __restrict__ desc, int16_t* __restrict__ out_sidx,
          float2* __restrict__ sel, float* __restrict__ theta)
{
    using C = FrontCfg<S>;
    constexpr int SE = C::SE;
    const ChanDesc& d = desc[blockIdx.y];
    if (d.S != S || !(d.flags & CH_FRONT_FAST) || (d.flags & (CH_FUSED | CH_FZS))) return;
    const int A = d.A;
    const int T_out = FT_ROWS - A + 1;           // output symbols per tile (host guarantees >= 64)
    const long long k0 = (long long)blockIdx.x * T_out;
    if (k0 >= d.K) return;
    const unsigned M = (unsigned)d.M;

    extern __shared__ double smem[];
    double* pbr = smem;                           // [FT_THREADS][S] inclusive prefix of the run sums
    double* wt  = pbr + C::PBR_DOUBLES;           // [warps][S]
    float*  es  = (float*)(wt + C::PB_DOUBLES);   // energies, run-padded rows

    const int tid = threadIdx.x;
    const long long s0 = k0 * S;                  // first virtual sample of the tile
    const long long V = d.tail_len + d.n_in;      // virtual stream length
    const bool in_tail = s0 < d.tail_len;
    const float2* gin = d.in + (s0 - d.tail_len); // valid address arithmetic only when !in_tail
    VStream vs{d.tail, d.in, d.tail_len};

    // ---- stage: 2 samples per thread per step -------------------------------------------------
    {
        constexpr int NCHUNK = FT_ROWS * S / 2;   // float4 chunks in the tile
        const long long avail = V - s0;           // samples of the tile that exist
        const bool vec_ok = !in_tail && ((reinterpret_cast<uintptr_t>(gin) & 15) == 0) && (S % 2 == 0);
        if (vec_ok && avail >= (long long)FT_ROWS * S) {
            const float4* g4 = reinterpret_cast<const float4*>(gin);
            if constexpr ((2 * FT_THREADS) % (8 * S) == 0) {
                // every step of FT_THREADS chunks covers a whole number of 8-row runs: the shared
                // address advances by a constant, no per-chunk division
                constexpr int ROWS_PER_STEP = 2 * FT_THREADS / S;
                constexpr int STEP_WORDS = (ROWS_PER_STEP / 8) * C::RUNW;
                const int j0 = (2 * tid) / S, p0 = 2 * tid - j0 * S;
                float* dst = es + (j0 >> 3) * C::RUNW + (j0 & 7) * SE + p0;
                const float4* src = g4 + tid;
                constexpr int NIT = NCHUNK / FT_THREADS, UN = 16;
                static_assert(NIT % UN == 0, "staging unroll");
                for (int it0 = 0; it0 < NIT; it0 += UN) {
                    float4 x[UN];
#pragma unroll
                    for (int u = 0; u < UN; u++) x[u] = __ldg(src + (it0 + u) * FT_THREADS);   // UN x 16 B in flight per thread
#pragma unroll
                    for (int u = 0; u < UN; u++)
                        *reinterpret_cast<float2*>(dst + (it0 + u) * STEP_WORDS) =
                            make_float2(energy_f32(x[u].x, x[u].y), energy_f32(x[u].z, x[u].w));
                }
            } else {
#pragma unroll 8
                for (int c = tid; c < NCHUNK; c += FT_THREADS) {
                    float4 x = __ldg(g4 + c);
                    const int smp = 2 * c, j = smp / S, p = smp - j * S;
                    float2 e = make_float2(energy_f32(x.x, x.y), energy_f32(x.z, x.w));
                    *reinterpret_cast<float2*>(es + (j >> 3) * C::RUNW + (j & 7) * SE + p) = e;
                }
            }
        } else {
            for (int smp = tid; smp < FT_ROWS * S; smp += FT_THREADS) {
                float2 x = make_float2(0.f, 0.f);
                if (smp < avail) x = in_tail ? vs.at(s0 + smp) : __ldg(gin + smp);
                const int j = smp / S, p = smp - j * S;
                es[(j >> 3) * C::RUNW + (j & 7) * SE + p] = energy_f32(x.x, x.y);
            }
        }
    }
    __syncthreads();

    // ---- run sums and their prefix over runs -----------------------------------------------------
    const int r = tid, lane = tid & 31, w = tid >> 5;
    const float* myrun = es + r * C::RUNW;
    double bs[S], pin[S];                         // run sum, inclusive prefix over runs
#pragma unroll
    for (int q = 0; q < S; q++) bs[q] = 0.0;
#pragma unroll
    for (int i = 0; i < FT_R; i++) {
        float rowv[S];
#pragma unroll
        for (int q = 0; q + 3 < S; q += 4) {
            float4 v = *reinterpret_cast<const float4*>(myrun + i * SE + q);
            rowv[q] = v.x; rowv[q + 1] = v.y; rowv[q + 2] = v.z; rowv[q + 3] = v.w;
        }
#pragma unroll
        for (int q = S & ~3; q < S; q++) rowv[q] = myrun[i * SE + q];
#pragma unroll
        for (int q = 0; q < S; q++) bs[q] = daddr(bs[q], (double)rowv[q]);
    }
#pragma unroll
    for (int q = 0; q < S; q++) pin[q] = bs[q];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
        for (int q = 0; q < S; q++) {
            double u = __shfl_up_sync(0xffffffffu, pin[q], o);
            if (lane >= o) pin[q] = daddr(pin[q], u);
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int q = 0; q < S; q++) wt[w * S + q] = pin[q];
    }
    __syncthreads();
#pragma unroll
    for (int ww = 0; ww < FT_THREADS / 32 - 1; ww++) {
        if (ww < w) {
#pragma unroll
            for (int q = 0; q < S; q++) pin[q] = daddr(pin[q], wt[ww * S + q]);
        }
    }
#pragma unroll
    for (int q = 0; q < S; q++) pbr[q * FT_THREADS + r] = pin[q];        // phase-major: lanes hit consecutive banks
    __syncthreads();

    // ---- window sums of this run's symbols ----------------------------------------------------------
    const int kk0 = r * FT_R;                     // tile-local index of this run's first symbol
    const int n_tile = (int)min((long long)T_out, d.K - k0);      // symbols this tile emits
    const int nvalid = min(FT_R, n_tile - kk0);   // symbols of this run that are emitted (<= 0: none)
    const int nq = A / FT_R, rem = A - nq * FT_R;
    int   oi[FT_R]; float2 os[FT_R]; float ot[FT_R];
    if (nvalid > 0) {
        double E[S];
        if (nq >= 1) {
            const double* hi = pbr + (r + nq - 1);         // runs r .. r+nq-1 = prefix[r+nq-1] - (prefix[r] - own sum)
#pragma unroll
            for (int q = 0; q < S; q++) E[q] = daddr(dsubr(hi[q * FT_THREADS], pin[q]), bs[q]);
        } else {
#pragma unroll
            for (int q = 0; q < S; q++) E[q] = 0.0;
        }
        {
            const int jr = kk0 + nq * FT_R;           // first of the `rem` extra rows
            for (int i = 0; i < rem; i++) {
                const int j = jr + i;
                const float* row = es + (j >> 3) * C::RUNW + (j & 7) * SE;
#pragma unroll
                for (int q = 0; q < S; q++) E[q] = daddr(E[q], (double)row[q]);
            }
        }
#pragma unroll
        for (int i = 0; i < FT_R; i++) {
            if (i < nvalid) {
                const int kk = kk0 + i;
                if (i > 0) {                              // slide: + newest symbol's energies, - oldest (cpp/psk_soft.cpp:451,576)
                    const int j = kk + A - 1;
                    const float* row = es + (j >> 3) * C::RUNW + (j & 7) * SE;
                    const float* trow = myrun + (i - 1) * SE;
                    float lead[S], trail[S];
#pragma unroll
                    for (int q = 0; q + 3 < S; q += 4) {
                        float4 v = *reinterpret_cast<const float4*>(row + q);
                        lead[q] = v.x; lead[q + 1] = v.y; lead[q + 2] = v.z; lead[q + 3] = v.w;
                        float4 u = *reinterpret_cast<const float4*>(trow + q);
                        trail[q] = u.x; trail[q + 1] = u.y; trail[q + 2] = u.z; trail[q + 3] = u.w;
                    }
#pragma unroll
                    for (int q = S & ~3; q < S; q++) { lead[q] = row[q]; trail[q] = trow[q]; }
#pragma unroll
                    for (int q = 0; q < S; q++) E[q] = dsubr(daddr(E[q], (double)lead[q]), (double)trail[q]);
                }
                double best = E[0]; int idx = 0;
#pragma unroll
                for (int q = 1; q < S; q++) if (best < E[q]) { best = E[q]; idx = q; }        // first maximum (:462)
                oi[i] = idx;
            }
        }
        // gather the chosen sample of every symbol of the run (:465): all loads in flight together
#pragma unroll
        for (int i = 0; i < FT_R; i++) {
            if (i < nvalid) {
                const int kk = kk0 + i;
                os[i] = in_tail ? vs.at(s0 + (long long)kk * S + oi[i]) : __ldg(gin + kk * S + oi[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < FT_R; i++) if (i < nvalid) ot[i] = mth_power_angle_fast(os[i], M);   // :474
    }
    // ---- stage the outputs in shared memory (the energy rows are dead now) and write them coalesced
    __syncthreads();
    // element kk lives at kk + (kk >> 3): one pad per run so the per-run (stride-8) writes spread over the banks
    constexpr int OPAD = FT_ROWS + FT_ROWS / 8;
    float2* s_sel = reinterpret_cast<float2*>(es);            // [OPAD]
    float*  s_th  = es + 2 * OPAD;                            // [OPAD]
    short*  s_idx = reinterpret_cast<short*>(es + 3 * OPAD);  // [OPAD]
    if (nvalid > 0) {
#pragma unroll
        for (int i = 0; i < FT_R; i++) {
            if (i < nvalid) { const int pos = kk0 + i + r; s_sel[pos] = os[i]; s_th[pos] = ot[i]; s_idx[pos] = (short)oi[i]; }
        }
    }
    __syncthreads();
    {
        int16_t* g_idx = out_sidx + d.sym_off + k0;
        float2* g_sel = sel + d.scr_off + k0;
        float* g_th = theta + d.scr_off + k0;
        for (int t = tid; t < n_tile; t += FT_THREADS) {
            const int pos = t + (t >> 3);
            g_sel[t] = s_sel[pos];
            g_th[t] = s_th[pos];
            g_idx[t] = s_idx[pos];
        }
    }
}

template <int S>
static cudaError_t launch_front_t(const LaunchCtx& c) {
    using C = FrontCfg<S>;
    static KernelCfg cfg;
    {
        cudaError_t e = cfg.ensure(k_front_t<S>, C::SMEM, FT_THREADS);
        if (e != cudaSuccess) return e;
    }
    const int T_out = FT_ROWS - c.Amin_fast + 1;   // smallest tile count that covers every fast channel is per channel; use the worst case
    (void)T_out;
    const int T_min = FT_ROWS - c.Amax_fast + 1;
    dim3 grid((unsigned)((c.Kmax + T_min - 1) / T_min), (unsigned)c.n_channels);
    double ab = 0.0;
    if (c.prof->enabled) for (int i = 0; i < c.n_channels; i++) { const ChanDesc& d = c.h_desc[i]; if (d.S == S && (d.flags & CH_FRONT_FAST) && !(d.flags & (CH_FUSED | CH_FZS))) ab += alg_bytes_front(d); }
    c.prof->begin(KID_FRONT, c.stream, ab);
    k_front_t<S><<<grid, FT_THREADS, C::SMEM, c.stream>>>(c.d_desc, c.out_sidx, c.d_sel, c.d_theta);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_front(const LaunchCtx& c) {
    if (c.Kmax <= 0) return cudaSuccess;
    // specialised kernels for the recommended oversampling factors (psk_soft.prf.xml:24 "8-10"), 16 too
    unsigned long long mask = c.S_mask;
    cudaError_t e = cudaSuccess;
    unsigned long long fmask = c.S_mask_fast;
    if (fmask & (1ull << 8))  { e = launch_front_t<8>(c);  if (e != cudaSuccess) return e; }
    if (fmask & (1ull << 9))  { e = launch_front_t<9>(c);  if (e != cudaSuccess) return e; }
    if (fmask & (1ull << 10)) { e = launch_front_t<10>(c); if (e != cudaSuccess) return e; }
    if (fmask & (1ull << 16)) { e = launch_front_t<16>(c); if (e != cudaSuccess) return e; }
    if (!mask) return cudaSuccess;
    int SPmax = c.Smax | 1;
    size_t smem = ((size_t)(FT + c.Amax) * SPmax + (size_t)(FRONT_THREADS / 2) * SPmax) * sizeof(double);
    static KernelCfg cfg;
    e = cfg.ensure(k_front, smem, FRONT_THREADS, nullptr, nullptr, -1 /* default carve-out */);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((c.Kmax + FT - 1) / FT), (unsigned)c.n_channels);
    double ab = 0.0;
    if (c.prof->enabled) for (int i = 0; i < c.n_channels; i++) { const ChanDesc& d = c.h_desc[i]; if (!(d.flags & (CH_FRONT_FAST | CH_FUSED | CH_FZS))) ab += alg_bytes_front(d); }
    c.prof->begin(KID_FRONT, c.stream, ab);
    k_front<<<grid, FRONT_THREADS, smem, c.stream>>>(c.d_desc, c.out_sidx, c.d_sel, c.d_theta, mask);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
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
    if (d.flags & (CH_FAST | CH_FUSED)) return; // handled by k_chain_par / k_fused
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
    double ab = 0.0;
    if (c.prof->enabled) for (int i = 0; i < c.n_channels; i++) { const ChanDesc& d = c.h_desc[i]; if (!(d.flags & (CH_FAST | CH_FUSED))) ab += alg_bytes_chain(d); }
    c.prof->begin(KID_CHAIN_SEQ, c.stream, ab);
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
       float2* __restrict__ out_soft, int16_t* __restrict__ out_bits, uint8_t* __restrict__ out_hard)
{
    const ChanDesc& d = desc[blockIdx.y];
    if (d.flags & (CH_FAST | CH_FUSED)) return; // k_chain_par / k_fused derotate and slice their own channels
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
    if ((out_bits || out_hard) && d.bpb) {
        unsigned b = slice_bits(c, d.bpb);
        if (out_bits) {
            int16_t* o = out_bits + d.bits_off + k * d.bpb;
            for (int j = 0; j < d.bpb; j++) o[j] = (int16_t)((b >> j) & 1u);
        }
        if (out_hard) out_hard[d.sym_off + k] = (uint8_t)b;
    }
}

cudaError_t launch_back(const LaunchCtx& c) {
    if (c.Kmax <= 0) return cudaSuccess;
    if (!c.out_soft && !c.out_bits && !c.out_hard) return cudaSuccess;
    if (c.n_seq_channels == 0) return cudaSuccess;
    dim3 grid((unsigned)((c.Kmax + 255) / 256), (unsigned)c.n_channels);
    const float* phase = c.out_phase ? c.out_phase : c.d_phase_tmp;
    double ab = 0.0;
    if (c.prof->enabled) for (int i = 0; i < c.n_channels; i++) { const ChanDesc& d = c.h_desc[i]; if (!(d.flags & (CH_FAST | CH_FUSED))) ab += alg_bytes_back(d); }
    c.prof->begin(KID_BACK, c.stream, ab);
    k_back<<<grid, 256, 0, c.stream>>>(c.d_desc, c.d_state, c.d_sel, phase, (float2*)c.out_soft, c.out_bits, c.out_hard);
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
    if (threadIdx.x == 0 && d.K > 0 && d.D && !(d.flags & CH_FUSED)) state[blockIdx.x].last = sel[d.scr_off + d.K - 1];   // :489
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
// k_chain_par: scan-based phase chain + derotate/slice.  ONE WARP PER CHANNEL, warp-synchronous
// (no block barriers), CW_WARPS channels per CTA.
//
// The reference's per-symbol recursion (cpp/psk_soft.cpp:474-482 + LinearFit::next :48-87)
//     n_k = round((est_{k-1} - theta_k)/2pi);  y_k = f32(theta_k + 2pi n_k);  est_k = fit(y_{k-P+1..k})
// is sequential only through the INTEGER n_k.  Per sub-block of CW_B = 32*CW_V symbols a warp
//   1. predicts n_k by classic sample-to-sample unwrapping (an integer warp scan), anchored on
//      the reference's own rule for the first symbol of the block,
//   2. evaluates y_k, the window sums ySum (differences of double prefix sums) and the xySum
//      recurrence X_k = X_{k-1} - xdelta*ySum'_k + T_k (another double warp scan, T_k rounded in
//      float exactly as :78) and from them est_k with the reference's float roundings (:157-162),
//   3. VERIFIES every n_k against the reference's rule using est_{k-1}; on the first mismatch it
//      shifts the remaining predictions by the observed difference and repeats (each pass
//      proves at least one more symbol), falling back to the literal sequential recursion for
//      that sub-block after CW_MAX_ITERS passes.
// So the emitted n_k are exactly those of the sequential recursion run on the same theta_k.  The
// double sums are formed in scan order instead of symbol order and the two divisions of :157-158
// are multiplications by correctly rounded reciprocals -- both <= 2 ulp(double) from the
// reference's value, i.e. a float-ulp flip of est in ~1e-8 of the symbols, far inside the stated
// 1e-4 tolerance (DESIGN.md "phase chain").  Fill-up (fewer than P points), the 2^20-call re-sum
// (:51-52), packet prologue/epilogue run on lane 0, literally.
// =============================================================================================
constexpr int CW_WARPS = 4;
#ifndef PSKD_CW_MIN_CTAS
#define PSKD_CW_MIN_CTAS 7
#endif
constexpr int CW_MIN_CTAS = PSKD_CW_MIN_CTAS;     // 28 warps/SM: a 4096-channel bank is one wave on 148 SMs
constexpr int CW_V = 4;
constexpr int CW_B = 32 * CW_V;
constexpr int CW_MAX_ITERS = 16;

struct CwShared {
    ChanState st;
    int flags;
    int mode;
    unsigned int passes, seq_blocks;
};

struct CwWarp {          // per-warp shared state
    ChanState st;
    FitConst fc;
    int flags;
    unsigned int passes, seq_blocks;
};

// literal recursion over nb symbols (theta staged in th[]), lane 0 only.  Used for the fill-up
// phase, around the 2^20-call re-sum and when the scan path gives up.
static __device__ __noinline__ void chain_block_sequential(ChanState& st, float* yb, const float* th, float* estv, int nb) {
    SmemRing ring{yb};
    for (int i = 0; i < nb; i++) {
        float y = unwrap_against(st.est, th[i], nullptr);
        st.est = fit_next(st.fit, ring, y);
        estv[i + 1] = st.est;
    }
}

// history suffix sums hs[j] = sum_{m=j}^{P-1} yb[m] (history in logical order, head == 0)
static __device__ __noinline__ void chain_rebuild_hs(const float* yb, double* hs, int P, int lane) {
    double carry = 0.0;
    for (int base = 0; base < P; base += 32) {
        const int rr = base + lane;                        // reversed index: element P-1-rr
        double x = (rr < P) ? (double)yb[P - 1 - rr] : 0.0;
        double inc = daddr(warp_scan_dbl(x, lane), carry);
        if (rr < P) hs[P - 1 - rr] = inc;
        carry = __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) hs[P] = 0.0;
    __syncwarp();
}

// rotate the ring so that yvals.front() sits at index 0 (tmp: >= P floats of scratch)
static __device__ __noinline__ void chain_normalize_ring(float* yb, float* tmp, FitState& f, int P, int lane) {
    const int head = f.head;
    __syncwarp();
    for (int j = lane; j < P; j += 32) { int s2 = head + j; if (s2 >= P) s2 -= P; tmp[j] = yb[s2]; }
    __syncwarp();
    for (int j = lane; j < P; j += 32) yb[j] = tmp[j];
    if (lane == 0) f.head = 0;
    __syncwarp();
}

// classic sample-to-sample unwrap increment (the chain's prediction rule; also the time-parallel
// chain's integer bookkeeping -- both must use the same arithmetic)
__device__ __forceinline__ int classic_dn(float th, float th_prev) {
    return -__float2int_rn((th - th_prev) * 0.15915494309189535f);
}

__global__ void __launch_bounds__(CW_WARPS * 32, CW_MIN_CTAS)
k_chain_par(const ChanDesc* __restrict__ desc, ChanState* __restrict__ state, float* __restrict__ ring_base,
            const float* __restrict__ theta, float* __restrict__ out_phase,
            double sri_xdelta, int Pcap, int n_channels, DevCounters* counters, const TpCtl tp)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int widx = blockIdx.x * CW_WARPS + wid;
    int ch, pk_a = 0, pk_b = -1, kind = 0, src = -1, dst = -1, pkt_slot = -1;
    if (tp.items) {
        if (widx >= tp.n_items) return;
        const TpItem it = tp.items[widx];
        ch = it.ch; pk_a = it.pk_a; pk_b = it.pk_b; kind = it.kind; src = it.src; dst = it.dst; pkt_slot = it.pkt_slot;
        if (desc[ch].flags & CH_FZS) return;                                            // k_fzs_cb's item
    } else {
        ch = widx;
        if (ch >= n_channels) return;
        if (!(desc[ch].flags & CH_FAST) || (desc[ch].flags & (CH_FUSED | CH_FZS))) return;
        if (tp.fallback) { if (!(desc[ch].flags & CH_TP) || !tp.fail[ch]) return; }   // re-run of a channel whose hand-overs were not proven
        else if (desc[ch].flags & CH_TP) return;                                        // handled through TpItems
    }

    extern __shared__ double smem_d[];
    const size_t per_warp_d = (size_t)CW_B + Pcap + 2;                       // ps[CW_B], hs[Pcap+1] (+1 pad)
    const size_t per_warp_f = (size_t)Pcap + CW_B + CW_B + CW_B + 4;         // yb[Pcap+CW_B], th[CW_B], estv[CW_B+1]
    double* ps   = smem_d + (size_t)wid * per_warp_d;
    double* hs   = ps + CW_B;
    float*  yb   = (float*)(smem_d + (size_t)CW_WARPS * per_warp_d) + (size_t)wid * per_warp_f;
    float*  blk  = yb + Pcap;                 // the sub-block's y values (16-byte aligned: Pcap % 4 == 0)
    float*  th   = yb + Pcap + CW_B;
    float*  estv = th + CW_B;
    __shared__ CwWarp shw[CW_WARPS];
    CwWarp& sh = shw[wid];

    // per-channel constants into registers once
    const int P = desc[ch].P, M = desc[ch].M, S = desc[ch].S, A = desc[ch].A, n_pkts = desc[ch].n_pkts;
    const int K = (int)desc[ch].K;
    const long long pkt_len = desc[ch].pkt_len, tail_len = desc[ch].tail_len;
    const float* thg = theta + desc[ch].scr_off;
    float* phg = out_phase + desc[ch].sym_off;
    float* gring = ring_base + desc[ch].ring_off;
    if (pk_b < 0) pk_b = n_pkts;
    const int k_begin = (pk_a == 0) ? 0 : (int)first_symbol_at((long long)pk_a * pkt_len, tail_len, S, A, K);

    if (kind == 0) {
        if (lane == 0) { sh.st = state[ch]; sh.flags = desc[ch].flags; sh.passes = 0; sh.seq_blocks = 0; }
        for (int j = lane; j < P; j += 32) yb[j] = gring[j];
    } else if (kind == 2) {
        if (lane == 0) { sh.st = tp.ends[src].st; sh.flags = desc[ch].flags & ~(CH_RESET_NUMSYMS | CH_RESET_PHASEAVG); sh.passes = 0; sh.seq_blocks = 0; }
        for (int j = lane; j < P; j += 32) yb[j] = tp.end_ring[(size_t)src * tp.ring_stride + j];
    } else {
        // synthesised start of packet pk_a: the history ring the previous packet leaves behind, from the
        // resolved integers: y = f32(theta + 2pi(c + A)), then the packet-end shift f32(y - w*wrapValue)
        const TpPacket pp = tp.pkts[pkt_slot - 1];
        const float wrapValue = __double2float_rn(dmulr(PSKD_M_2PI, (double)M));
        const float shift = fmulr((float)pp.w, wrapValue);
        int carry = 0;                                      // sum of the increments of the symbols after the current chunk
        for (int e = k_begin; e > k_begin - P; e -= 32) {
            const int m = e - 1 - lane;                     // lanes run backwards from the packet's last symbol
            int dn = 0;
            float t = 0.0f;
            const bool in = m >= k_begin - P;
            if (in) { t = __ldg(thg + m); dn = classic_dn(t, __ldg(thg + m - 1)); }
            const int incl = warp_scan_int(dn, lane);       // increments of symbols m .. e-1
            if (in) {
                const int c = pp.cEnd - (carry + incl - dn);
                float y = __double2float_rn(daddr((double)t, dmulr((double)(c + pp.A), PSKD_M_2PI)));
                if (pp.w != 0) y = fsubr(y, shift);                                      // :131 (subtractConst)
                yb[m - (k_begin - P)] = y;
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        if (lane == 0) {
            sh.st = (src >= 0) ? tp.ends[src].st : state[ch];   // fit constants (xdelta, denominator, n) of the channel
            sh.st.fit.head = 0; sh.st.fit.pts = P; sh.st.wraps = 0;
            SmemRing ring{yb};
            sh.st.est = fit_resum(sh.st.fit, ring);         // exact after a wrap (:601-602); else within rounding of the
                                                            // carried estimate, which only feeds the first unwrap count
            sh.flags = desc[ch].flags & ~(CH_RESET_NUMSYMS | CH_RESET_PHASEAVG); sh.passes = 0; sh.seq_blocks = 0;
        }
        __syncwarp();
        if (dst >= 0) for (int j = lane; j < P; j += 32) tp.start_ring[(size_t)dst * tp.ring_stride + j] = yb[j];
    }
    __syncwarp();
    const float est_start_used = sh.st.est;
    const unsigned long long wraps0 = sh.st.wraps;
    const float fP1 = (float)(P - 1);

    // software prefetch of the next sub-block's theta: strided layout, element i = lane + 32*q
    int pf_k = k_begin;
    float pf_th[CW_V];
#pragma unroll
    for (int q = 0; q < CW_V; q++) { const int kk = k_begin + lane + 32 * q; pf_th[q] = (kk < K) ? __ldg(thg + kk) : 0.0f; }

    bool hs_valid = false;
    int k = k_begin;
    for (int pkt = pk_a; pkt < pk_b; pkt++) {
        if (lane == 0) {
            SmemRing ring{yb};
            chain_packet_prologue(sh.st, ring, desc[ch], sri_xdelta, sh.flags);
            if (sh.st.fit.pts == sh.st.fit.n && sh.st.fit.pts > 1) sh.fc = fit_const(sh.st.fit);
        }
        const int khi = (pkt + 1 == n_pkts) ? K : (int)first_symbol_at((long long)(pkt + 1) * pkt_len, tail_len, S, A, K);
        __syncwarp();
        while (k < khi) {
            int nb = min(CW_B, khi - k);
            const int pts = sh.st.fit.pts, cnt = sh.st.fit.count;
            bool fast = (pts == P) && (P > 1) && (cnt + nb <= 1048576);
            if (!fast) {
                if (pts == P && cnt == 1048576) {                                   // :51-52 at a block edge
                    if (lane == 0) { SmemRing ring{yb}; fit_resum(sh.st.fit, ring); }
                    __syncwarp();
                    continue;
                }
                if (pts == P && P > 1) nb = 1048576 - cnt;                          // stop at the re-sum point
                else nb = min(nb, max(1, P - pts));                                 // fill-up runs sequentially
                fast = (pts == P) && (P > 1);
            }
            // stage theta (strided registers -> shared), start the next prefetch
            if (pf_k != k) {
#pragma unroll
                for (int q = 0; q < CW_V; q++) { const int kk = k + lane + 32 * q; pf_th[q] = (kk < K) ? __ldg(thg + kk) : 0.0f; }
            }
#pragma unroll
            for (int q = 0; q < CW_V; q++) th[lane + 32 * q] = pf_th[q];
            pf_k = k + nb;
#pragma unroll
            for (int q = 0; q < CW_V; q++) { const int kk = pf_k + lane + 32 * q; pf_th[q] = (kk < K) ? __ldg(thg + kk) : 0.0f; }
            const float est0 = sh.st.est;
            if (lane == 0) estv[0] = est0;
            __syncwarp();

            bool done = false;
            if (fast) {
                if (sh.st.fit.head != 0) { chain_normalize_ring(yb, reinterpret_cast<float*>(hs), sh.st.fit, P, lane); hs_valid = false; }
                if (!hs_valid) { chain_rebuild_hs(yb, hs, P, lane); hs_valid = true; }
                const FitConst fc = sh.fc;
                const float xdelta = sh.st.fit.xdelta;
                const double xd = (double)xdelta;
                const double X0 = sh.st.fit.xySum;
                const int i0 = lane * CW_V;
                // this lane's four consecutive symbols
                const float4 t4 = *reinterpret_cast<const float4*>(th + i0);
                const float tl[CW_V] = {t4.x, t4.y, t4.z, t4.w};
                float tprev = __shfl_up_sync(0xffffffffu, t4.w, 1);
                // classic-unwrap prediction of n (integer scan), first symbol by the reference's rule
                int nloc[CW_V];
                {
                    int run = 0;
#pragma unroll
                    for (int v = 0; v < CW_V; v++) {
                        int dn = 0;
                        if (i0 + v < nb) {
                            const float pv = (v == 0) ? tprev : tl[v - 1];
                            dn = (i0 + v == 0) ? unwrap_count(est0, tl[0]) : classic_dn(tl[v], pv);
                        }
                        run += dn; nloc[v] = run;
                    }
                    const int off = warp_scan_int(run, lane) - run;
#pragma unroll
                    for (int v = 0; v < CW_V; v++) nloc[v] += off;
                }
                int iter = 0;
                float yl[CW_V], el[CW_V]; double Yl[CW_V], Xl[CW_V];
                while (true) {
                    double run = 0.0;
#pragma unroll
                    for (int v = 0; v < CW_V; v++) {
                        float y = 0.0f;
                        if (i0 + v < nb) y = __double2float_rn(daddr((double)tl[v], dmulr((double)nloc[v], PSKD_M_2PI)));   // :478,481
                        yl[v] = y; run = daddr(run, (double)y); Yl[v] = run;
                    }
                    *reinterpret_cast<float4*>(blk + i0) = make_float4(yl[0], yl[1], yl[2], yl[3]);
                    const double off = dsubr(warp_scan_dbl(run, lane), run);      // prefix before this lane
                    double Pl[CW_V];
#pragma unroll
                    for (int v = 0; v < CW_V; v++) Pl[v] = daddr(off, Yl[v]);      // inclusive prefix of the block's y
                    *reinterpret_cast<double2*>(ps + i0) = make_double2(Pl[0], Pl[1]);
                    *reinterpret_cast<double2*>(ps + i0 + 2) = make_double2(Pl[2], Pl[3]);
                    __syncwarp();
                    double trun = 0.0;
#pragma unroll
                    for (int v = 0; v < CW_V; v++) {
                        const int i = i0 + v;
                        double t = 0.0;
                        if (i < nb) {
                            double W = (i + 1 <= P - 1) ? hs[i + 1] : 0.0;             // history part of y_{k-P+1..k-1}
                            double pb = (v == 0) ? off : Pl[v - 1];                    // prefix through y_{i-1}
                            if (i - P >= 0) pb = dsubr(pb, ps[i - P]);
                            W = daddr(W, pb);                                         // ySum after :70
                            const double a = dmulr(xd, W);                             // :72
                            const double T = (double)fmulr(fmulr(yl[v], fP1), xdelta); // :78
                            t = dsubr(T, a);
                            Yl[v] = daddr(W, (double)yl[v]);                           // :75  (Yl now holds ySum)
                        }
                        trun = daddr(trun, t); Xl[v] = trun;
                    }
                    const double xoff = daddr(X0, dsubr(warp_scan_dbl(trun, lane), trun));
#pragma unroll
                    for (int v = 0; v < CW_V; v++) {
                        Xl[v] = daddr(xoff, Xl[v]);
                        el[v] = fit_eval_fast(fc, Yl[v], Xl[v], nullptr, nullptr);     // :135-162
                    }
                    // verify every predicted n against the reference's rule (:477) with est_{i-1}
                    float eprev = __shfl_up_sync(0xffffffffu, el[CW_V - 1], 1);
                    int mymis = 0x7fffffff, mydelta = 0;
#pragma unroll
                    for (int v = CW_V - 1; v >= 0; v--) {
                        const int i = i0 + v;
                        if (i >= 1 && i < nb) {
                            const int nt = unwrap_count((v == 0) ? eprev : el[v - 1], tl[v]);
                            if (nt != nloc[v]) { mymis = i; mydelta = nt - nloc[v]; }
                        }
                    }
                    const int mis = (int)__reduce_min_sync(0xffffffffu, (unsigned)mymis);
                    if (mis == 0x7fffffff) { done = true; break; }
                    if (++iter > CW_MAX_ITERS) break;
                    const int delta = __shfl_sync(0xffffffffu, mydelta, mis / CW_V);
#pragma unroll
                    for (int v = 0; v < CW_V; v++) if (i0 + v >= mis) nloc[v] += delta;
                    __syncwarp();
                }
                if (iter && lane == 0) sh.passes += (unsigned)iter;
                if (done) {
                    *reinterpret_cast<float4*>(estv + 4 + i0) = make_float4(el[0], el[1], el[2], el[3]);   // estv[4+i] = est_i
                    const int last = nb - 1;
                    if (last / CW_V == lane) {
                        double Yv = 0.0, Xv = 0.0;
#pragma unroll
                        for (int v = 0; v < CW_V; v++) if (v == last % CW_V) { Yv = Yl[v]; Xv = Xl[v]; }
                        FitState& f = sh.st.fit;
                        float mm, bb;
                        sh.st.est = fit_eval_fast(fc, Yv, Xv, &mm, &bb);
                        f.ySum = Yv; f.xySum = Xv; f.m = mm; f.b = bb; f.count += nb;
                    }
                    __syncwarp();
                    if (nb >= P) {
                        // new history = last P symbols of the block; its suffix sums come from the block prefix
                        const double pend = ps[nb - 1];
                        for (int j = lane; j < P; j += 32) {
                            const int src = nb - P + j;               // block index of new history element j
                            const double before = (src >= 1) ? ps[src - 1] : 0.0;
                            hs[j] = dsubr(pend, before);
                            yb[j] = blk[src];
                        }
                        __syncwarp();
                    } else {
                        for (int base = 0; base < P; base += 32) {    // forward chunks: reads stay ahead of writes
                            const int j = base + lane;
                            const int src = nb + j;                   // index in (history ++ block)
                            float v = 0.0f;
                            if (j < P) v = (src < P) ? yb[src] : blk[src - P];
                            __syncwarp();
                            if (j < P) yb[j] = v;
                            __syncwarp();
                        }
                        hs_valid = false;
                    }
                } else {
                    if (lane == 0) sh.seq_blocks++;
                }
            }
            if (!done) {
                if (lane == 0) {
                    chain_block_sequential(sh.st, yb, th, estv + 3, nb);       // estv[4+i] = est_i
                    if (sh.st.fit.pts == sh.st.fit.n && sh.st.fit.pts > 1) sh.fc = fit_const(sh.st.fit);
                }
                hs_valid = false;
                __syncwarp();
            }
            // phase output (cpp/psk_soft.cpp:482), coalesced
#pragma unroll
            for (int q = 0; q < CW_V; q++) { const int i = lane + 32 * q; if (i < nb) phg[k + i] = estv[4 + i]; }
            __syncwarp();
            k += nb;
        }
        if (lane == 0) {
            SmemRing ring{yb};
            const unsigned long long w0 = sh.st.wraps;
            chain_packet_epilogue(sh.st, ring, M);
            sh.flags |= (sh.st.wraps != w0) ? (1 << 30) : 0;
        }
        __syncwarp();
        if (sh.flags & (1 << 30)) { hs_valid = false; if (lane == 0) sh.flags &= ~(1 << 30); __syncwarp(); }
    }
    if (dst >= 0) {
        // time-parallel item: leave the end state in its record (k_tp_check installs the last one)
        if (sh.st.fit.head != 0 && sh.st.fit.pts == P) chain_normalize_ring(yb, reinterpret_cast<float*>(hs), sh.st.fit, P, lane);
        __syncwarp();
        for (int j = lane; j < P; j += 32) tp.end_ring[(size_t)dst * tp.ring_stride + j] = yb[j];
        if (lane == 0) {
            TpEnd& e = tp.ends[dst];
            e.st = sh.st; e.est_start_used = est_start_used; e.has_symbols = (k > k_begin) ? 1 : 0;
            e.wraps_delta = sh.st.wraps - wraps0;
        }
    } else {
        for (int j = lane; j < P; j += 32) gring[j] = yb[j];
        if (lane == 0) state[ch] = sh.st;              // `last` is carried by k_finish
    }
    if (lane == 0) {
        if (sh.st.wraps != wraps0) atomicAdd(&counters->wraps, sh.st.wraps - wraps0);
        atomicAdd(&counters->spec_chunks, (unsigned long long)((k - k_begin + CW_B - 1) / CW_B));
        if (sh.passes) atomicAdd(&counters->spec_misses, (unsigned long long)sh.passes);
        if (sh.seq_blocks) atomicAdd(&counters->seq_channels, (unsigned long long)sh.seq_blocks);
    }
}

// ---------------------------------------------------------------------------------------------
// k_back_par: derotate / differential decode / slice for the channels of the scan chain, one
// thread per symbol, fully coalesced (cpp/psk_soft.cpp:484-566).
// ---------------------------------------------------------------------------------------------
constexpr int BP_THREADS = 256;
#ifndef PSKD_BP_V
#define PSKD_BP_V 4
#endif
constexpr int BP_V = PSKD_BP_V;               // symbols per thread: BP_V x 12 B of loads in flight
constexpr int BP_TILE = BP_THREADS * BP_V;

__global__ void __launch_bounds__(BP_THREADS)
k_back_par(const ChanDesc* __restrict__ desc, const ChanState* __restrict__ state,
           const float2* __restrict__ sel, const float* __restrict__ phase,
           float2* __restrict__ out_soft, int16_t* __restrict__ out_bits, uint8_t* __restrict__ out_hard)
{
    const ChanDesc& d = desc[blockIdx.y];
    if (!(d.flags & CH_FAST) || (d.flags & (CH_FUSED | CH_FZS))) return;
    const int K = (int)d.K;
    const int k0 = blockIdx.x * BP_TILE;
    if (k0 >= K) return;
    __shared__ short sb[BP_TILE * 3];
    const int M = d.M, bpb = d.bpb;
    const bool diff = d.D != 0;
    const float2* selg = sel + d.scr_off;
    const float* phg = phase + d.sym_off;
    float2* softg = out_soft ? out_soft + d.sym_off : nullptr;
    const float inv_m = 1.0f / (float)M;
    const bool m_pow2 = (M & (M - 1)) == 0;

    float2 sv[BP_V], pv[BP_V]; float ev[BP_V];
#pragma unroll
    for (int v = 0; v < BP_V; v++) {              // all loads first
        const int k = k0 + threadIdx.x + v * BP_THREADS;
        sv[v] = make_float2(0.f, 0.f); pv[v] = sv[v]; ev[v] = 0.f;
        if (k < K) {
            sv[v] = __ldg(selg + k);
            if (diff) pv[v] = (k > 0) ? __ldg(selg + k - 1) : state[blockIdx.y].last;
            else ev[v] = __ldg(phg + k);
        }
    }
#pragma unroll
    for (int v = 0; v < BP_V; v++) {
        const int kl = threadIdx.x + v * BP_THREADS, k = k0 + kl;
        if (k < K) {
            float2 s = sv[v];
            float pc = 0.0f;
            if (diff) s = cdiv_f32(s, pv[v]);                                                   // :488
            else pc = m_pow2 ? fmulr(-ev[v], inv_m) : __fdiv_rn(-ev[v], (float)M);               // :494 (exact for 2^n)
            if (M == 4) pc = __double2float_rn(daddr((double)pc, PSKD_M_PI_4));                  // :497-498
            const float2 c = derotate(s, pc);
            if (softg) softg[k] = c;
            unsigned b = 0;
            if (bpb == 3) b = slice8_fast(c);
            else if (bpb == 1) b = (c.x < 0.0f) ? 1u : 0u;
            else if (bpb == 2) b = slice_bits(c, 2);
            short* o = sb + kl * bpb;
            for (int j = 0; j < bpb; j++) o[j] = (short)((b >> j) & 1u);
            if (out_hard && bpb) out_hard[d.sym_off + k] = (uint8_t)b;
        }
    }
    if (!out_bits || !bpb) return;
    __syncthreads();
    // bits_dataShort_out: one short per bit, LSB first (cpp/psk_soft.cpp:512,525-526,559-563), written coalesced
    const int nsh = min(BP_TILE, K - k0) * bpb;
    int16_t* o = out_bits + d.bits_off + (long long)k0 * bpb;
    if ((reinterpret_cast<uintptr_t>(o) & 3) == 0) {
        const int n2 = nsh >> 1;
        const int* s2 = reinterpret_cast<const int*>(sb);
        int* o2 = reinterpret_cast<int*>(o);
        for (int t = threadIdx.x; t < n2; t += BP_THREADS) o2[t] = s2[t];
        if ((nsh & 1) && threadIdx.x == 0) o[nsh - 1] = sb[nsh - 1];
    } else {
        for (int t = threadIdx.x; t < nsh; t += BP_THREADS) o[t] = sb[t];
    }
}

// ---------------------------------------------------------------------------------------------
// time-parallel chain, auxiliary kernels (see TpCtl in pskd_internal.h)
// ---------------------------------------------------------------------------------------------
// k_tp_scan: one warp per (channel, packet): classic unwrap over the packet (integer scan) and the
// linear-fit estimate of the relative phases at the packet end (double, regression over the last P)
__global__ void __launch_bounds__(128)
k_tp_scan(const ChanDesc* __restrict__ desc, const float* __restrict__ theta, const TpItem* __restrict__ items,
          int n_items, TpPacket* __restrict__ pkts)
{
    const int lane = threadIdx.x & 31;
    const int iidx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (iidx >= n_items) return;
    const TpItem it = items[iidx];                         // one item per time-parallel packet
    const int slot = it.pkt_slot;
    const ChanDesc& d = desc[it.ch];
    const int pkt = it.pk_a;
    const int K = (int)d.K, P = d.P;
    const int klo = (int)first_symbol_at((long long)pkt * d.pkt_len, d.tail_len, d.S, d.A, K);
    const int khi = (pkt + 1 == d.n_pkts) ? K : (int)first_symbol_at((long long)(pkt + 1) * d.pkt_len, d.tail_len, d.S, d.A, K);
    const float* thg = theta + d.scr_off;
    int c = 0;                                             // count at the last symbol processed so far
    {
        int acc = 0;                                       // per-lane partial sums, one warp reduction at the end
        // head up to the first 16-byte aligned angle, then 4 angles per lane and load (the scratch rows are shifted so that
        // packets start aligned, pskd_api.cu), then the tail
        int m0 = klo + 1;
        const int a0 = min(khi, (int)(m0 + ((4 - ((reinterpret_cast<uintptr_t>(thg + m0) >> 2) & 3)) & 3)));
        if (m0 + lane < a0) acc += classic_dn(__ldg(thg + m0 + lane), __ldg(thg + m0 + lane - 1));
        m0 = a0;
        const int n4 = (khi - m0) >> 2;                    // whole float4 groups
        for (int g = lane; g < n4; g += 32) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(thg + m0) + g);
            const float pv = __ldg(thg + m0 + 4 * g - 1);
            acc += classic_dn(t.x, pv) + classic_dn(t.y, t.x) + classic_dn(t.z, t.y) + classic_dn(t.w, t.z);
        }
        for (int base = m0 + 4 * n4; base < khi; base += 32) {
            const int m = base + lane;
            if (m < khi) acc += classic_dn(__ldg(thg + m), __ldg(thg + m - 1));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        c = acc;
    }
    // regression of phi = theta + 2pi*c over the last min(P, n) symbols, evaluated at the newest one
    const int n = khi - klo, np = min(P, n);
    double s0 = 0.0, s1 = 0.0, q2 = 0.0;
    {
        int carry = 0;
        const double mid = 0.5 * (double)(np - 1);
        for (int e = khi; e > khi - np; e -= 32) {
            const int m = e - 1 - lane;
            const bool in = m >= khi - np;
            int dn = 0; float t = 0.0f;
            if (in) { t = __ldg(thg + m); if (m > klo) dn = classic_dn(t, __ldg(thg + m - 1)); }
            const int incl = warp_scan_int(dn, lane);
            if (in) {
                const int cm = c - (carry + incl - dn);
                const double phi = (double)t + PSKD_M_2PI * (double)cm;
                const double x = (double)(m - (khi - np)) - mid;
                s0 += phi; s1 += x * phi; q2 += x * x;
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o);
        }
    }
    if (lane == 0) {
        TpPacket p;
        p.cEnd = c;
        p.dLink = (klo > 0 && n > 0) ? classic_dn(thg[klo], thg[klo - 1]) : 0;
        p.A = 0; p.w = 0;
        p.estRelEnd = (np > 0) ? s0 / (double)np + ((q2 > 0.0) ? (s1 / q2) * 0.5 * (double)(np - 1) : 0.0) : 0.0;
        p.klo = klo; p.khi = khi;
        pkts[slot] = p;
    }
}

// k_tp_resolve: one warp per channel: the scalar recurrence over its packets: unwrap level of every
// packet's first symbol and the wrap count of every packet end (cpp/psk_soft.cpp:592-603).  The
// packet records are read 32 at a time (one per lane); lane 0's recurrence walks them by shuffles.
__global__ void __launch_bounds__(128)
k_tp_resolve(const ChanDesc* __restrict__ desc, const float* __restrict__ theta,
             const TpChan* __restrict__ chans, int n_chans, TpPacket* __restrict__ pkts,
             const TpEnd* __restrict__ ends, const ChanState* __restrict__ state)
{
    const int lane = threadIdx.x & 31;
    const int ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ci >= n_chans) return;
    const TpChan tc = chans[ci];
    const ChanDesc& d = desc[tc.ch];
    const float* thg = theta + d.scr_off;
    const float wrapValue = __double2float_rn(dmulr(PSKD_M_2PI, (double)d.M));
    const int np = tc.n_pkts - tc.pkt0;
    const int M = d.M;
    // level of the first time-parallel packet: the reference's rule with the exact estimate the head left
    // (no head: the channel's carried estimate)
    long long A = 0;
    {
        const TpPacket p0 = pkts[tc.first_slot];
        const float est0 = tc.has_head ? ends[tc.first_item].st.est : state[tc.ch].est;
        if (p0.khi > p0.klo) { long long n = 0; (void)unwrap_against(est0, thg[p0.klo], &n); A = n; }
    }
    // |trunc(est)| > wrapValue (the integer abs of cpp/psk_soft.cpp:596)  <=>  |est| >= floor(wrapValue) + 1; the
    // estimates here are regression values (not the exact chain's), every decision is proven later by k_tp_check
    const float wrapThr = floorf(wrapValue) + 1.0f, rWrap = 1.0f / wrapValue;
    double Ad = (double)A;
    for (int base = 0; base < np; base += 32) {
        const int j = base + lane;
        int step = 0, has = 0; double er = 0.0;
        if (j < np) {
            const TpPacket p = pkts[tc.first_slot + j];
            step = p.cEnd; er = p.estRelEnd; has = p.khi > p.klo;
            if (j + 1 < np) step += pkts[tc.first_slot + j + 1].dLink;
        }
        int myA = 0, myW = 0;
        const int cnt = min(32, np - base);
        // uniform loop: every lane follows the same recurrence.  Only the level Ad is carried from step to step; the
        // shuffles do not depend on it, so unrolling lets them run ahead of the dependent DFMA -> F2F -> compare chain
#pragma unroll 8
        for (int l = 0; l < 32; l++) {
            const int s_l = __shfl_sync(0xffffffffu, step, l), h_l = __shfl_sync(0xffffffffu, has, l);
            const double e_l = __shfl_sync(0xffffffffu, er, l);
            if (l < cnt) {
                const float est_end = (float)fma(PSKD_M_2PI, Ad, e_l);
                int w = 0;
                if (h_l && fabsf(est_end) >= wrapThr) w = (int)rintf(est_end * rWrap);
                if (l == lane) { myA = (int)Ad; myW = w; }
                Ad += (double)(s_l - M * w);
            }
        }
        if (j < np) { pkts[tc.first_slot + j].A = myA; pkts[tc.first_slot + j].w = myW; }
    }
}

// k_tp_check: one warp per (channel, time-parallel packet): prove the hand-over INTO this packet: the ring it
// started from equals, bit for bit, the ring its predecessor ended with, and its first unwrap count is the
// one the predecessor's exact end estimate gives.  A failed proof raises the channel's flag.
__global__ void __launch_bounds__(128)
k_tp_check(const ChanDesc* __restrict__ desc, const float* __restrict__ theta, const TpCtl tp)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (slot >= tp.n_items) return;
    if (tp.rerun && tp.any_rerun && *tp.any_rerun == 0) return;      // the check after a repair round nobody asked for
    const TpItem item = tp.items[slot];
    if (item.kind != 1) return;                                      // the first packet starts from the head's exact state
    const ChanDesc& d = desc[item.ch];
    const int P = d.P, it = item.dst, prev = it - 1;
    bool same = true;
    for (int i = lane; i < P; i += 32)
        same = same && (__float_as_uint(tp.start_ring[(size_t)it * tp.ring_stride + i]) ==
                        __float_as_uint(tp.end_ring[(size_t)prev * tp.ring_stride + i]));
    same = __all_sync(0xffffffffu, same);
    if (same && tp.ends[it].has_symbols) {
        long long n_used = 0, n_true = 0;
        const float t0 = theta[d.scr_off + tp.pkts[item.pkt_slot].klo];
        (void)unwrap_against(tp.ends[it].est_start_used, t0, &n_used);
        (void)unwrap_against(tp.ends[prev].st.est, t0, &n_true);
        same = n_used == n_true;
    }
    if (!same && lane == 0) { tp.fail[item.ch] = 1; if (tp.slot_fail) tp.slot_fail[item.pkt_slot] = 1; }
}

// k_tp_fix: the repair round (k_fzs_cb channels).  One warp per channel whose proof failed.  Packets before the
// first unproven hand-over jf are exact (induction from the exact start).  What went wrong is the LEVEL: somewhere
// the classic sample-to-sample count differs from the reference's unwrap-against-the-fit rule.  Every packet has
// by now run the exact chain once, and the integer advance of the unwrap count over a packet does not depend on
// the level it ran at -- so the exact advances replace the classic ones in the recurrence, packet jf restarts from
// its predecessor's exact end record, and the later packets from re-synthesised rings; k_tp_check then proves the
// new hand-overs.  A channel that fails again goes to the sequential chain.
__global__ void __launch_bounds__(128)
k_tp_fix(const ChanDesc* __restrict__ desc, const float* __restrict__ theta, const TpCtl tp, DevCounters* counters)
{
    const int lane = threadIdx.x & 31;
    const int ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ci >= tp.n_chans) return;
    const TpChan tc = tp.chans[ci];
    const ChanDesc& d = desc[tc.ch];
    const int np = tc.n_pkts - tc.pkt0;
    const bool mine = (d.flags & CH_FZS) && tp.fail[tc.ch];
    if (!mine) {
        for (int j = lane; j < np; j += 32) tp.slot_run[tc.first_slot + j] = 0;
        return;
    }
    int jf = 0x7fffffff;
    for (int j = lane; j < np; j += 32) if (tp.slot_fail[tc.first_slot + j]) jf = min(jf, j);
    jf = (int)__reduce_min_sync(0xffffffffu, (unsigned)jf);
    if (jf < 1 || jf >= np) {                                         // nothing to restart from: stays failed
        for (int j = lane; j < np; j += 32) tp.slot_run[tc.first_slot + j] = 0;
        return;
    }
    for (int j = lane; j < np; j += 32) {
        tp.slot_run[tc.first_slot + j] = (j < jf) ? 0 : (j == jf ? 2 : 1);
        tp.slot_fail[tc.first_slot + j] = 0;
    }
    __syncwarp();
    if (lane != 0) return;
    *tp.any_rerun = 1;
    atomicAdd(&counters->tp_repaired, 1ULL);
    const float* thg = theta + d.scr_off;
    const int M = d.M;
    const float wrapValue = __double2float_rn(dmulr(PSKD_M_2PI, (double)M));
    const int rec0 = tc.first_item + tc.has_head;                     // end record of packet jj: rec0 + jj
    // exact level of packet jf's first symbol: the reference's rule with the proven predecessor's end estimate
    long long A = 0;
    {
        const TpPacket pf = tp.pkts[tc.first_slot + jf];
        if (pf.khi > pf.klo) (void)unwrap_against(tp.ends[rec0 + jf - 1].st.est, thg[pf.klo], &A);
    }
    for (int j = jf; j < np; j++) {
        TpPacket& p = tp.pkts[tc.first_slot + j];
        const TpEnd& e = tp.ends[rec0 + j];
        const int adv = e.has_symbols ? e.n_last - e.n_first : 0;     // exact advance over the packet (level-invariant)
        const float est_end = (float)((double)e.est_pre + PSKD_M_2PI * (double)(A - (long long)e.n_first));
        int w = 0;
        if (e.has_symbols && wrap_needed(est_end, wrapValue)) w = (int)roundf(__fdiv_rn(est_end, wrapValue));
        p.A = (int)A; p.w = w; p.cEnd = adv;
        const int dl = (j + 1 < np) ? tp.pkts[tc.first_slot + j + 1].dLink : 0;
        A = A + adv + dl - (long long)M * w;
    }
    tp.fail[tc.ch] = 0;                                               // the next k_tp_check decides again
}

// k_tp_install: one warp per channel: all hand-overs proven -> the last packet's end state becomes the
// channel's carried state
__global__ void __launch_bounds__(128)
k_tp_install(const ChanDesc* __restrict__ desc, ChanState* __restrict__ state, float* __restrict__ ring_base,
             const TpCtl tp, DevCounters* counters)
{
    const int lane = threadIdx.x & 31;
    const int ci = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ci >= tp.n_chans) return;
    const TpChan tc = tp.chans[ci];
    const ChanDesc& d = desc[tc.ch];
    const int P = d.P, np = tc.n_pkts - tc.pkt0;
    if (tp.fail[tc.ch]) { if (lane == 0) atomicAdd(&counters->seq_channels, 1ULL); return; }
    unsigned long long wraps = 0;
    const int n_rec = np + tc.has_head;                               // end records of the channel: [head,] one per packet
    for (int j = lane; j < n_rec; j += 32) wraps += tp.ends[tc.first_item + j].wraps_delta;
#pragma unroll
    for (int o = 16; o; o >>= 1) wraps += __shfl_xor_sync(0xffffffffu, wraps, o);
    const int last = tc.first_item + n_rec - 1;                       // record of the last packet
    for (int i = lane; i < P; i += 32) ring_base[d.ring_off + i] = tp.end_ring[(size_t)last * tp.ring_stride + i];
    if (lane == 0) {
        ChanState st = tp.ends[last].st;
        st.wraps = state[tc.ch].wraps + wraps;
        state[tc.ch] = st;
        atomicAdd(&counters->tp_packets, (unsigned long long)np);
    }
}

static cudaError_t launch_chain_kernel(const LaunchCtx& c, int Pcap, size_t smem, int n_warps, const TpCtl& tp, double ab = 0.0) {
    float* phase = c.out_phase ? c.out_phase : c.d_phase_tmp;
    int blocks = (n_warps + CW_WARPS - 1) / CW_WARPS;
    if (blocks < 1) return cudaSuccess;
    c.prof->begin(KID_CHAIN_PAR, c.stream, ab);
    k_chain_par<<<blocks, CW_WARPS * 32, smem, c.stream>>>(c.d_desc, c.d_state, c.d_ring, c.d_theta, phase,
                                                          c.sri_xdelta, Pcap, c.n_channels, c.d_counters, tp);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_chain_par(const LaunchCtx& c) {
    const int n_legacy = c.n_fast_channels - c.n_fzs_channels;       // scan-chain channels of k_chain_par + k_back_par
    const int n_fzs = c.n_fzs_channels;                              // ... and of k_fzs_cb (chain + back in one kernel)
    if (n_legacy <= 0 && n_fzs <= 0) return cudaSuccess;
    int Pcap = c.Pmax_fast < 1 ? 1 : c.Pmax_fast;
    Pcap = (Pcap + 3) & ~3;
    size_t per_warp = ((size_t)CW_B + Pcap + 2) * sizeof(double) + ((size_t)Pcap + 3 * CW_B + 4) * sizeof(float);
    size_t smem = per_warp * CW_WARPS;
    cudaError_t e;
    if (n_legacy > 0) {
        static KernelCfg cfg;
        e = cfg.ensure(k_chain_par, smem, CW_WARPS * 32);
        if (e != cudaSuccess) return e;
    }
    float* phase = c.out_phase ? c.out_phase : c.d_phase_tmp;
    TpCtl tp{};
    tp.chans = c.tp_chans; tp.n_chans = c.tp_n_chans; tp.pkts = c.tp_pkts; tp.ends = c.tp_ends;
    tp.end_ring = c.tp_end_ring; tp.start_ring = c.tp_start_ring; tp.ring_stride = c.tp_ring_stride; tp.fail = c.tp_fail;
    tp.slot_fail = c.tp_slot_fail; tp.slot_run = c.tp_slot_run; tp.any_rerun = c.tp_any_rerun;
    const int tp_fzs = c.tp_n_chans_fzs, tp_legacy = c.tp_n_chans - c.tp_n_chans_fzs;
    // one round of chain work over `items` (null: one unit per channel), each kernel taking its own channels
    // algorithmic bytes per kernel and class of channels (profiling only): [0] packet-after-packet, [1] time-parallel
    double ab_legacy[2] = {0, 0}, ab_fzs[2] = {0, 0}, ab_back = 0;
    if (c.prof->enabled)
        for (int i = 0; i < c.n_channels; i++) {
            const ChanDesc& d = c.h_desc[i];
            if (!(d.flags & CH_FAST) || (d.flags & CH_FUSED)) continue;
            const int k = (d.flags & CH_TP) ? 1 : 0;
            if (d.flags & CH_FZS) ab_fzs[k] += alg_bytes_chain(d) + alg_bytes_back(d);
            else { ab_legacy[k] += alg_bytes_chain(d); ab_back += alg_bytes_back(d); }
        }
    // one round of chain work over `items` (null: one unit per channel), each kernel taking its own channels; the
    // algorithmic bytes of the time-parallel channels are booked on the round over their packets (which = 1)
    auto chain_round = [&](const TpCtl& t, int n_units, bool legacy, bool fzs, int which = -1) -> cudaError_t {
        cudaError_t r = cudaSuccess;
        if (legacy) { r = launch_chain_kernel(c, Pcap, smem, n_units, t, which >= 0 ? ab_legacy[which] : 0.0); if (r != cudaSuccess) return r; }
        if (fzs) r = launch_fzs_cb(c, t, n_units, which >= 0 ? ab_fzs[which] : 0.0);
        return r;
    };
    // channels whose chain runs packet after packet
    e = chain_round(tp, c.n_channels, n_legacy > tp_legacy, n_fzs > tp_fzs, 0);
    if (e != cudaSuccess) return e;
    if (c.tp_n_chans > 0) {
        // heads (packets before the time-parallel range), scan, resolve, all packets in parallel, proof, re-runs
        if (c.tp_n_head > 0) {
            TpCtl t1 = tp; t1.items = c.tp_head_items; t1.n_items = c.tp_n_head;
            e = chain_round(t1, c.tp_n_head, tp_legacy > 0, tp_fzs > 0);
            if (e != cudaSuccess) return e;
        }
        c.prof->begin(KID_TP, c.stream);
        k_tp_scan<<<(c.tp_n_items + 3) / 4, 128, 0, c.stream>>>(c.d_desc, c.d_theta, c.tp_items, c.tp_n_items, c.tp_pkts);
        k_tp_resolve<<<(c.tp_n_chans + 3) / 4, 128, 0, c.stream>>>(c.d_desc, c.d_theta, c.tp_chans, c.tp_n_chans, c.tp_pkts, c.tp_ends, c.d_state);
        c.prof->end(c.stream);
        (*c.launches) += 2;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        TpCtl t2 = tp; t2.items = c.tp_items; t2.n_items = c.tp_n_items;
        e = chain_round(t2, c.tp_n_items, tp_legacy > 0, tp_fzs > 0, 1);
        if (e != cudaSuccess) return e;
        c.prof->begin(KID_TP, c.stream);
        k_tp_check<<<(c.tp_n_items + 3) / 4, 128, 0, c.stream>>>(c.d_desc, c.d_theta, t2);
        c.prof->end(c.stream);
        (*c.launches)++;
        // repair rounds (k_fzs_cb channels): exact advances instead of classic ones from the first unproven hand-over on
        static const int rounds = getenv("PSKD_TP_ROUNDS") ? atoi(getenv("PSKD_TP_ROUNDS")) : 1;
        for (int r = 0; r < rounds && tp_fzs > 0; r++) {
            c.prof->begin(KID_TP, c.stream);
            k_tp_fix<<<(c.tp_n_chans + 3) / 4, 128, 0, c.stream>>>(c.d_desc, c.d_theta, t2, c.d_counters);
            c.prof->end(c.stream);
            (*c.launches)++;
            TpCtl t2r = t2; t2r.rerun = 1;
            e = chain_round(t2r, c.tp_n_items, false, true);
            if (e != cudaSuccess) return e;
            c.prof->begin(KID_TP, c.stream);
            k_tp_check<<<(c.tp_n_items + 3) / 4, 128, 0, c.stream>>>(c.d_desc, c.d_theta, t2r);
            c.prof->end(c.stream);
            (*c.launches)++;
        }
        c.prof->begin(KID_TP, c.stream);
        k_tp_install<<<(c.tp_n_chans + 3) / 4, 128, 0, c.stream>>>(c.d_desc, c.d_state, c.d_ring, tp, c.d_counters);
        c.prof->end(c.stream);
        (*c.launches)++;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        TpCtl t3 = tp; t3.fallback = 1;
        e = chain_round(t3, c.n_channels, tp_legacy > 0, tp_fzs > 0);
        if (e != cudaSuccess) return e;
    }
    if (n_legacy > 0 && (c.out_soft || c.out_bits || c.out_hard) && c.Kmax > 0) {
        dim3 grid((unsigned)((c.Kmax + BP_TILE - 1) / BP_TILE), (unsigned)c.n_channels);
        c.prof->begin(KID_BACK_PAR, c.stream, ab_back);
        k_back_par<<<grid, BP_THREADS, 0, c.stream>>>(c.d_desc, c.d_state, c.d_sel, phase, (float2*)c.out_soft, c.out_bits, c.out_hard);
        c.prof->end(c.stream);
        (*c.launches)++;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace pskd
