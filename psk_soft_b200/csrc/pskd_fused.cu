// pskd_fused.cu -- the whole demod path of one channel in ONE kernel pass (sm_100a).
//
//   k_fused<S>   ingest + symbol timing + M-th power angle + unwrap/LinearFit chain + derotate /
//                differential decode / slice, per channel, with every intermediate (energies,
//                window sums, selected samples, angles, unwrapped phases) living in shared memory
//                or registers.  HBM sees only the algorithmic bytes: each IQ sample read once,
//                each output written once.
//                reference rows: cpp/psk_soft.cpp:380-603, 619-636 and LinearFit :35-185.
//
// Work decomposition: ONE WARP PER UNIT, warp-synchronous (no block barriers).  A unit is a run
// of consecutive emulated BULKIO packets of one channel (>= ~4096 symbols).  A persistent grid
// pulls units from a ticket counter in packet-major order; unit (ch, j) waits for (ch, j-1)
// (its phase-chain state travels through global memory), which was ticketed n_channels earlier.
//
// Per unit the warp streams over its symbols in chunks of 32 input rows (1 row = 1 symbol = S
// samples):
//   ingest : coalesced 128-bit loads (prefetched one chunk ahead in registers), e = f32(re^2+im^2)
//            (std::norm<float>, :448) into a ring of the last numAvg-1+32 rows' energies
//   timing : lane (phase p, row group g) forms the exact (double) sliding window sums of its rows
//            (:451, :576), an exclusive scan over the row groups adds the carried window sum; the
//            sums are transposed through shared memory so that lane = row takes the FIRST maximum
//            over the phases (:462), gathers that sample of the window's oldest symbol (:465,
//            an L2 hit: the row was streamed numAvg-1 rows ago) and forms atan2f(s^M) (:474)
//   chain  : every 128 symbols (or at a packet end): classic-unwrap prediction of the integer
//            unwrap counts, double prefix sums for LinearFit's ySum / xySum, point-wise verification
//            of every count against the reference's rule round((est_{k-1}-theta_k)/2pi) (:477) and
//            repair -- the emitted integers are exactly those of the sequential recursion
//   back   : derotate by -est/M (+pi/4) or divide by the previous sample (:484-501), slice
//            (:503-566); phase, soft and bits are staged in shared memory and written coalesced.
#include "pskd_internal.h"
#include "pskd_device.cuh"

namespace pskd {

constexpr int FZ_WARPS = 4;            // warps (= concurrent units) per CTA
constexpr int FZ_CH = 32;              // input rows per ingest chunk
constexpr int FZ_B = 128;              // symbols per chain block (4 per lane)
constexpr int FZ_BUF = 160;            // capacity of the (theta, sample) block buffer
constexpr int FZ_MAX_ITERS = 16;
#ifndef PSKD_FZ_MIN_CTAS
#define PSKD_FZ_MIN_CTAS 5
#endif

template <int S> struct FzCfg {
    static constexpr int G = 32 / S;                         // row groups (lane = g*S + p)
    static constexpr int R = (32 + G - 1) / G;               // rows per group
    static constexpr bool PADDED = ((R * S) % 32) == 0;      // groups would collide on the banks: pad
    static constexpr int PAD = PADDED ? 32 / G : 0;          // floats of padding after every R rows
    static constexpr int ES = (S & 1) ? S : S + 1;           // row stride of the transposition buffer (doubles)
    static constexpr int NQ = (S * 16 + 31) / 32;            // float4 loads per lane per chunk
    static_assert(!PADDED || (32 % R) == 0, "padded groups must tile a chunk");
    __host__ __device__ static constexpr int fpos(int pos) { return pos * S + (PADDED ? (pos / R) * PAD : 0); }
    __host__ __device__ static constexpr int ring_rows(int A) { return ((A - 1 + 32 + 31) / 32) * 32; }
    __host__ __device__ static constexpr int ring_floats(int A) { return (fpos(ring_rows(A) + R) + 3) & ~3; }
};

struct FzLayout {          // byte offsets inside one warp's shared-memory region
    int off_th, off_sel, off_yh, off_st, off_cz, off_alias, bytes;
};

struct FzWarp {
    ChanState st;
    FitConst fc;
    int flags;
    unsigned int passes, seq_blocks, blocks;
};

struct FusedParams {
    const ChanDesc* desc; ChanState* state; float* ring_base;
    const int* list; int n_list;       // channels served by this launch
    int n_units;                       // n_list * max units per channel
    int pkts_per_unit;
    int* ticket;                       // unit ticket counter (zeroed before the launch)
    int* done;                         // [n_channels] units completed per channel (zeroed before the launch)
    float2* out_soft; int16_t* out_bits; float* out_phase; int16_t* out_sidx;
    double sri_xdelta;
    int Pcap;
    FzLayout lay;
    DevCounters* counters;
};

template <int S>
static FzLayout fz_layout(int Amax, int Pcap) {
    using C = FzCfg<S>;
    FzLayout L;
    int o = C::ring_floats(Amax) * 4;
    L.off_th = o;                      o += FZ_BUF * 4;
    L.off_sel = o;                     o += (FZ_BUF + 2) * 8;
    o = (o + 15) & ~15;
    L.off_yh = o;                      o += Pcap * 4;
    o = (o + 15) & ~15;
    L.off_st = o;                      o += (int)((sizeof(FzWarp) + 15) & ~15);
    L.off_cz = o;                      o += ((Pcap + 1) * 8 + 15) & ~15;
    L.off_alias = o;
    int e = 32 * C::ES * 8;                          // window-sum transposition buffer
    int c = FZ_B * 8 + FZ_B * 4 + (FZ_B + 4) * 4;    // chain: prefix block, y block, est block
    o += ((e > c ? e : c) + 15) & ~15;
    L.bytes = o;
    return L;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// history prefix cz[j] = sum_{i<j} yh[i], j = 0..P (history in logical order)
static __device__ __noinline__ void fz_rebuild_cz(const float* yh, double* cz, int P, int lane) {
    double carry = 0.0;
    for (int base = 0; base < P; base += 32) {
        const int j = base + lane;
        const double x = (j < P) ? (double)yh[j] : 0.0;
        const double inc = daddr(warp_scan_dbl(x, lane), carry);
        if (j < P) cz[j + 1] = inc;
        carry = __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) cz[0] = 0.0;
    __syncwarp();
}

// rotate the ring so that yvals.front() sits at index 0 (tmp: >= P floats of scratch)
static __device__ __noinline__ void fz_normalize_ring(float* yh, float* tmp, FitState& f, int P, int lane) {
    const int head = f.head;
    __syncwarp();
    for (int j = lane; j < P; j += 32) { int s2 = head + j; if (s2 >= P) s2 -= P; tmp[j] = yh[s2]; }
    __syncwarp();
    for (int j = lane; j < P; j += 32) yh[j] = tmp[j];
    if (lane == 0) f.head = 0;
    __syncwarp();
}

static __device__ __noinline__ void fz_block_sequential(ChanState& st, float* yh, const float* th, float* estv, int nb) {
    SmemRing ring{yh};
    for (int i = 0; i < nb; i++) {
        float y = unwrap_against(st.est, th[i], nullptr);
        st.est = fit_next(st.fit, ring, y);
        estv[i] = st.est;
    }
}

template <int S>
__global__ void __launch_bounds__(FZ_WARPS * 32, PSKD_FZ_MIN_CTAS)
k_fused(const FusedParams prm)
{
    using C = FzCfg<S>;
    constexpr int G = C::G, R = C::R, ES = C::ES, NQ = C::NQ;
    extern __shared__ __align__(16) unsigned char fz_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wb = fz_smem + (size_t)wid * prm.lay.bytes;
    float*  ring = reinterpret_cast<float*>(wb);
    float*  th   = reinterpret_cast<float*>(wb + prm.lay.off_th);
    float2* selb = reinterpret_cast<float2*>(wb + prm.lay.off_sel);     // [1] = previous symbol's sample, [2..] = block
    float*  yh   = reinterpret_cast<float*>(wb + prm.lay.off_yh);
    FzWarp& sh   = *reinterpret_cast<FzWarp*>(wb + prm.lay.off_st);
    unsigned char* alias = wb + prm.lay.off_alias;
    double* ebuf = reinterpret_cast<double*>(alias);                    // [32][ES] window sums (ingest phase)
    float*  yblk = reinterpret_cast<float*>(alias + FZ_B * 8);          // chain phase: the block's y values
    float*  estv = yblk + FZ_B;                                         // chain phase: est per symbol (sequential path)

    // lane roles of the timing step
    const int wg = lane / S, wp = lane - wg * S;       // row group, phase
    const bool wact = lane < G * S;

    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(prm.ticket, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= prm.n_units) break;
        const int ug = u / prm.n_list;
        const int ch = __ldg(prm.list + (u - ug * prm.n_list));
        const ChanDesc& dg = prm.desc[ch];
        const int n_pkts = dg.n_pkts;
        const int pk0 = ug * prm.pkts_per_unit;
        if (pk0 >= n_pkts) continue;
        const int pk1 = min(pk0 + prm.pkts_per_unit, n_pkts);
        const int A = dg.A, M = dg.M, P = dg.P, bpb = dg.bpb;
        const bool diff = dg.D != 0;
        const long long tail_len = dg.tail_len, pkt_len = dg.pkt_len;
        const long long V = tail_len + dg.n_in;
        const int K = (int)dg.K;
        const int lag = A - 1;
        const int RR = C::ring_rows(A), NS = RR / 32;
        VStream vs{dg.tail, dg.in, tail_len};
        const int kA = (int)first_symbol_at((long long)pk0 * pkt_len, tail_len, S, A, K);
        const int kB = (pk1 == n_pkts) ? K : (int)first_symbol_at((long long)pk1 * pkt_len, tail_len, S, A, K);
        double* cz = reinterpret_cast<double*>(alias) - (P + 1);        // cz[0..P] history prefix, cz[P+1..] block prefix

        // ---- ingest prefetch of chunk 0 (in flight while we wait for the predecessor unit) -------
        const int nchunks = (kB - kA + FZ_CH - 1) / FZ_CH;
        float4 xq[NQ];
        bool xq_valid = false;
        auto chunk_fast = [&](int c) -> bool {          // whole chunk inside `in`, 16-byte aligned
            const long long sA = (long long)(kA + c * FZ_CH + lag) * S;
            if (sA < tail_len || sA + 32 * S > V) return false;
            return (reinterpret_cast<uintptr_t>(dg.in + (sA - tail_len)) & 15) == 0;
        };
        auto chunk_issue = [&](int c) {
            const long long sA = (long long)(kA + c * FZ_CH + lag) * S;
            const float4* g4 = reinterpret_cast<const float4*>(dg.in + (sA - tail_len));
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const int f = lane + 32 * q;
                if ((S * 16) % 32 == 0 || f < S * 16) xq[q] = __ldg(g4 + f);
            }
        };
        if (nchunks > 0 && chunk_fast(0)) { chunk_issue(0); xq_valid = true; }

        // ---- wait for the previous unit of this channel, then load its carried state --------------
        if (ug > 0) {
            if (lane == 0) { while (ld_acquire(prm.done + ch) < ug) __nanosleep(200); }
            __syncwarp();
        }
        float* gring = prm.ring_base + dg.ring_off;
        if (lane == 0) {
            const int4* src = reinterpret_cast<const int4*>(prm.state + ch);
            int4* dst = reinterpret_cast<int4*>(&sh.st);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(ChanState) / 16); i++) dst[i] = __ldcg(src + i);
            sh.flags = (ug == 0) ? dg.flags : (dg.flags & ~(CH_RESET_NUMSYMS | CH_RESET_PHASEAVG));
            sh.passes = 0; sh.seq_blocks = 0; sh.blocks = 0;
        }
        for (int j = lane; j < P; j += 32) yh[j] = __ldcg(gring + j);
        __syncwarp();
        const unsigned long long wraps0 = sh.st.wraps;
        const float fP1 = (float)(P - 1);
        if (lane == 0) selb[1] = sh.st.last;

        // ---- prime the energy ring with rows [kA, kA+lag) and the carried window sum ---------------
        double Cw = 0.0;               // lane (p, g): sum over the window of output k0 WITHOUT its newest row
        if (nchunks > 0) {
            const long long s0 = (long long)kA * S;
            for (int s = lane; s < lag * S; s += 32) {
                const long long v = s0 + s;
                const float2 x = (v < V) ? vs.at(v) : make_float2(0.f, 0.f);
                const int row = s / S, p = s - row * S;
                ring[C::fpos(RR - lag + row) + p] = energy_f32(x.x, x.y);
            }
            __syncwarp();
            if (wact) {
                double acc = 0.0;
                for (int i = wg; i < lag; i += G) acc = daddr(acc, (double)ring[C::fpos(RR - lag + i) + wp]);
                ebuf[wg * ES + wp] = acc;
            }
            __syncwarp();
            if (wact) {
#pragma unroll
                for (int g2 = 0; g2 < G; g2++) Cw = daddr(Cw, ebuf[g2 * ES + wp]);
            }
            __syncwarp();
        }

        // ---- chain bookkeeping -------------------------------------------------------------------
        int pkt = pk0;                 // packet whose prologue has run
        int kchain = kA;               // next symbol the chain consumes
        int nbuf = 0;                  // symbols waiting in (th, selb)
        bool cz_valid = false;
        if (lane == 0) {
            SmemRing r{yh};
            chain_packet_prologue(sh.st, r, dg, prm.sri_xdelta, sh.flags);
            if (sh.st.fit.pts == sh.st.fit.n && sh.st.fit.pts > 1) sh.fc = fit_const(sh.st.fit);
        }
        __syncwarp();
        int pk_hi = (pkt + 1 == n_pkts) ? K : (int)first_symbol_at((long long)(pkt + 1) * pkt_len, tail_len, S, A, K);
        bool unit_done = false;

        // consume buffered symbols: blocks of FZ_B (shorter at a packet end); run the packet
        // epilogue / next prologue whenever a packet is exhausted
        auto drain = [&]() {
            while (!unit_done) {
                const int rem = pk_hi - kchain;
                if (rem == 0) {
                    if (lane == 0) {
                        SmemRing r{yh};
                        const unsigned long long w0 = sh.st.wraps;
                        chain_packet_epilogue(sh.st, r, M);                                   // :592-603
                        sh.flags = (sh.flags & ~(1 << 30)) | ((sh.st.wraps != w0) ? (1 << 30) : 0);
                    }
                    __syncwarp();
                    if (sh.flags & (1 << 30)) cz_valid = false;
                    pkt++;
                    if (pkt == pk1) { unit_done = true; break; }
                    if (lane == 0) {
                        SmemRing r{yh};
                        const int c0 = sh.st.fit.count, h0 = sh.st.fit.head;
                        chain_packet_prologue(sh.st, r, dg, prm.sri_xdelta, sh.flags);       // :393-426
                        if (sh.st.fit.pts == sh.st.fit.n && sh.st.fit.pts > 1) sh.fc = fit_const(sh.st.fit);
                        (void)c0; (void)h0;
                    }
                    __syncwarp();
                    pk_hi = (pkt + 1 == n_pkts) ? K : (int)first_symbol_at((long long)(pkt + 1) * pkt_len, tail_len, S, A, K);
                    continue;
                }
                const int want = min(FZ_B, rem);
                if (nbuf < want) break;

                // ---- one sub-block of m symbols at buffer offset 0 ---------------------------------
                int m = want;
                const int pts = sh.st.fit.pts, cnt = sh.st.fit.count;
                bool fast = (pts == P) && (P > 1);
                if (fast && cnt + m > 1048576) {
                    if (cnt == 1048576) {                                                     // :51-52 at a block edge
                        if (lane == 0) { SmemRing r{yh}; fit_resum(sh.st.fit, r); }
                        __syncwarp();
                        continue;
                    }
                    m = 1048576 - cnt;                                                        // stop at the re-sum point
                }
                if (!fast) m = min(m, max(1, P - pts));                                       // fill-up runs sequentially
                const int i0 = lane * 4;
                const float4 t4 = *reinterpret_cast<const float4*>(th + i0);
                const float tl[4] = {t4.x, t4.y, t4.z, t4.w};
                const float est0 = sh.st.est;
                float el[4];
                bool done = false;
                if (fast) {
                    if (sh.st.fit.head != 0) { fz_normalize_ring(yh, estv, sh.st.fit, P, lane); cz_valid = false; }
                    if (!cz_valid) { fz_rebuild_cz(yh, cz, P, lane); cz_valid = true; }
                    const FitConst fc = sh.fc;
                    const float xdelta = sh.st.fit.xdelta;
                    const double xd = (double)xdelta;
                    const double X0 = sh.st.fit.xySum;
                    const double HPP = cz[P];
                    // classic-unwrap prediction of n (integer scan), first symbol by the reference's rule
                    int nloc[4];
                    {
                        const float tprev = __shfl_up_sync(0xffffffffu, t4.w, 1);
                        int run = 0;
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            const float pv = (v == 0) ? tprev : tl[v - 1];
                            int dn = -__float2int_rn((tl[v] - pv) * 0.15915494309189535f);
                            if (v == 0 && lane == 0) dn = unwrap_count(est0, tl[0]);
                            if (i0 + v >= m) dn = 0;
                            run += dn; nloc[v] = run;
                        }
                        const int off = warp_scan_int(run, lane) - run;
#pragma unroll
                        for (int v = 0; v < 4; v++) nloc[v] += off;
                    }
                    int iter = 0;
                    float yl[4]; double Ys[4], Xl[4];
                    while (true) {
                        double Cl[4];
                        {
                            double run = 0.0;
#pragma unroll
                            for (int v = 0; v < 4; v++) {
                                float y = 0.0f;
                                if (i0 + v < m) y = __double2float_rn(daddr((double)tl[v], dmulr((double)nloc[v], PSKD_M_2PI)));   // :478,481
                                yl[v] = y; run = daddr(run, (double)y); Cl[v] = run;
                            }
                            const double off = daddr(HPP, dsubr(warp_scan_dbl(run, lane), run));
#pragma unroll
                            for (int v = 0; v < 4; v++) Cl[v] = daddr(off, Cl[v]);             // cz[P+1+i]
                            Ys[0] = off;                                                      // cz[P+i0]
                        }
                        *reinterpret_cast<float4*>(yblk + i0) = make_float4(yl[0], yl[1], yl[2], yl[3]);
                        *reinterpret_cast<double2*>(cz + P + 1 + i0) = make_double2(Cl[0], Cl[1]);
                        *reinterpret_cast<double2*>(cz + P + 3 + i0) = make_double2(Cl[2], Cl[3]);
                        __syncwarp();
                        double trun = 0.0;
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            const int i = i0 + v;
                            const double hi = (v == 0) ? Ys[0] : Cl[v - 1];                   // cz[P+i]
                            const double W = dsubr(hi, cz[i + 1]);                            // ySum after :70
                            const double a = dmulr(xd, W);                                    // :72
                            const double T = (double)fmulr(fmulr(yl[v], fP1), xdelta);        // :78
                            trun = daddr(trun, dsubr(T, a)); Xl[v] = trun;
                            Ys[v] = daddr(W, (double)yl[v]);                                  // :75
                        }
                        const double xoff = daddr(X0, dsubr(warp_scan_dbl(trun, lane), trun));
#pragma unroll
                        for (int v = 0; v < 4; v++) {
                            Xl[v] = daddr(xoff, Xl[v]);
                            el[v] = fit_eval_fast(fc, Ys[v], Xl[v], nullptr, nullptr);        // :135-162
                        }
                        // verify every predicted n against the reference's rule (:477) with est_{i-1}
                        const float eprev = __shfl_up_sync(0xffffffffu, el[3], 1);
                        int mymis = 0x7fffffff, mydelta = 0;
#pragma unroll
                        for (int v = 3; v >= 0; v--) {
                            const int i = i0 + v;
                            if (i >= 1 && i < m) {
                                const int nt = unwrap_count((v == 0) ? eprev : el[v - 1], tl[v]);
                                if (nt != nloc[v]) { mymis = i; mydelta = nt - nloc[v]; }
                            }
                        }
                        const int mis = (int)__reduce_min_sync(0xffffffffu, (unsigned)mymis);
                        if (mis == 0x7fffffff) { done = true; break; }
                        if (++iter > FZ_MAX_ITERS) break;
                        const int delta = __shfl_sync(0xffffffffu, mydelta, mis >> 2);
#pragma unroll
                        for (int v = 0; v < 4; v++) if (i0 + v >= mis) nloc[v] += delta;
                        __syncwarp();
                    }
                    if (iter && lane == 0) sh.passes += (unsigned)iter;
                    if (done) {
                        const int last = m - 1;
                        if ((last >> 2) == lane) {
                            double Yv = Ys[0], Xv = Xl[0];
#pragma unroll
                            for (int v = 1; v < 4; v++) if (v == (last & 3)) { Yv = Ys[v]; Xv = Xl[v]; }
                            FitState& f = sh.st.fit;
                            float mm, bb;
                            sh.st.est = fit_eval_fast(fc, Yv, Xv, &mm, &bb);
                            f.ySum = Yv; f.xySum = Xv; f.m = mm; f.b = bb; f.count += m;
                        }
                        __syncwarp();
                        // new history = last P of (history ++ block): shift the prefix and the values by m
                        const double czm = cz[m];
                        for (int base = 0; base <= P; base += 32) {
                            const int j = base + lane;
                            double pv = 0.0; float yv = 0.0f;
                            if (j <= P) pv = dsubr(cz[m + j], czm);
                            if (j < P) yv = (m + j < P) ? yh[m + j] : yblk[m + j - P];
                            __syncwarp();
                            if (j <= P) cz[j] = pv;
                            if (j < P) yh[j] = yv;
                            __syncwarp();
                        }
                    } else {
                        if (lane == 0) sh.seq_blocks++;
                        cz_valid = false;
                    }
                }
                if (!done) {
                    __syncwarp();
                    if (lane == 0) {
                        fz_block_sequential(sh.st, yh, th, estv, m);
                        if (sh.st.fit.pts == sh.st.fit.n && sh.st.fit.pts > 1) sh.fc = fit_const(sh.st.fit);
                    }
                    cz_valid = false;
                    __syncwarp();
#pragma unroll
                    for (int v = 0; v < 4; v++) el[v] = (i0 + v < m) ? estv[i0 + v] : 0.0f;
                    __syncwarp();
                }
                if (lane == 0) sh.blocks++;

                // ---- back: derotate / differential decode / slice (cpp/psk_soft.cpp:484-566) ----------
                const float2 prev_new = selb[2 + m - 1];
                float2 sv[4];
                {
                    const float4 a = *reinterpret_cast<const float4*>(selb + 2 + i0);
                    const float4 b = *reinterpret_cast<const float4*>(selb + 4 + i0);
                    sv[0] = make_float2(a.x, a.y); sv[1] = make_float2(a.z, a.w);
                    sv[2] = make_float2(b.x, b.y); sv[3] = make_float2(b.z, b.w);
                }
                const float2 sprev = selb[1 + i0];
                __syncwarp();
                // phase_dataFloat_out (:482): stage in th[0..m), consumed entries only
                *reinterpret_cast<float4*>(th + i0) = (i0 + 3 < m) ? make_float4(el[0], el[1], el[2], el[3]) : t4;
                if (i0 < m && i0 + 3 >= m) {
#pragma unroll
                    for (int v = 0; v < 4; v++) if (i0 + v < m) th[i0 + v] = el[v];
                }
                unsigned bw[6] = {0, 0, 0, 0, 0, 0};        // this lane's 4*bpb bits as shorts, packed in words
                {
                    const float inv_m = 1.0f / (float)M;
                    const bool m_pow2 = (M & (M - 1)) == 0;
#pragma unroll
                    for (int v = 0; v < 4; v++) {
                        float2 s = sv[v];
                        float pc = 0.0f;
                        if (diff) s = cdiv_f32(s, (v == 0) ? sprev : sv[v - 1]);                               // :488
                        else pc = m_pow2 ? fmulr(-el[v], inv_m) : __fdiv_rn(-el[v], (float)M);                  // :494 (exact for 2^n)
                        if (M == 4) pc = __double2float_rn(daddr((double)pc, PSKD_M_PI_4));                     // :497-498
                        const float2 c = derotate(s, pc);                                                      // :499-501
                        if (i0 + v < m) selb[2 + i0 + v] = c;
                        unsigned b = 0;
                        if (bpb == 3) b = slice8_fast(c);
                        else if (bpb == 1) b = (c.x < 0.0f) ? 1u : 0u;
                        else if (bpb == 2) b = slice_bits(c, 2);
                        // shorts: symbol v occupies shorts [v*bpb, (v+1)*bpb)
#pragma unroll
                        for (int j = 0; j < 3; j++) {
                            if (j < bpb) {
                                const int sidx = v * bpb + j;
                                bw[sidx >> 1] |= ((b >> j) & 1u) << ((sidx & 1) * 16);
                            }
                        }
                    }
                }
                short* bstage = reinterpret_cast<short*>(alias);           // chain buffers are dead now
                if (bpb > 0 && prm.out_bits) {
                    unsigned* bs32 = reinterpret_cast<unsigned*>(bstage) + lane * 2 * bpb;
#pragma unroll
                    for (int w = 0; w < 6; w++) if (w < 2 * bpb) bs32[w] = bw[w];
                }
                __syncwarp();
                {
                    const long long so = dg.sym_off + kchain;
                    if (prm.out_phase) {
                        float* o = prm.out_phase + so;
#pragma unroll
                        for (int q = 0; q < 4; q++) { const int i = lane + 32 * q; if (i < m) o[i] = th[i]; }
                    }
                    if (prm.out_soft) {
                        float2* o = prm.out_soft + so;
#pragma unroll
                        for (int q = 0; q < 4; q++) { const int i = lane + 32 * q; if (i < m) o[i] = selb[2 + i]; }
                    }
                    if (bpb > 0 && prm.out_bits) {
                        // bits_dataShort_out: one short per bit, LSB first (:512, 525-526, 559-563)
                        int16_t* o = prm.out_bits + dg.bits_off + (long long)kchain * bpb;
                        const int nsh = m * bpb;
                        if ((reinterpret_cast<uintptr_t>(o) & 3) == 0) {
                            const unsigned* s32 = reinterpret_cast<const unsigned*>(bstage);
                            unsigned* o32 = reinterpret_cast<unsigned*>(o);
                            for (int t = lane; t < (nsh >> 1); t += 32) o32[t] = s32[t];
                            if ((nsh & 1) && lane == 0) o[nsh - 1] = bstage[nsh - 1];
                        } else {
                            for (int t = lane; t < nsh; t += 32) o[t] = bstage[t];
                        }
                    }
                }
                __syncwarp();
                // ---- drop the consumed symbols from the buffer ----------------------------------------
                const int left = nbuf - m;
                for (int base = 0; base < left; base += 32) {
                    const int i = base + lane;
                    float tv = 0.f; float2 sv2 = make_float2(0.f, 0.f);
                    if (i < left) { tv = th[m + i]; sv2 = selb[2 + m + i]; }
                    __syncwarp();
                    if (i < left) { th[i] = tv; selb[2 + i] = sv2; }
                    __syncwarp();
                }
                if (lane == 0) selb[1] = prev_new;
                __syncwarp();
                nbuf = left; kchain += m;
            }
        };

        drain();     // packets without symbols before the first chunk (and units without any symbol)

        // ---- the chunk loop ------------------------------------------------------------------------
        for (int c = 0; c < nchunks && !unit_done; c++) {
            const int k0 = kA + c * FZ_CH;
            const int nrows = min(FZ_CH, kB - k0);
            const int slot = c % NS;
            float* slotp = ring + C::fpos(32 * slot);
            // ingest: energies of the chunk's 32 newest rows (the windows' leading edge, :448-451)
            if (xq_valid) {
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const int f = lane + 32 * q;
                    if ((S * 16) % 32 == 0 || f < S * 16) {
                        const int s = 2 * f;
                        const int off = s + (C::PADDED ? ((s / S) / R) * C::PAD : 0);
                        const float2 e = make_float2(energy_f32(xq[q].x, xq[q].y), energy_f32(xq[q].z, xq[q].w));
                        if (S % 2 == 0) {
                            *reinterpret_cast<float2*>(slotp + off) = e;
                            if (slot == 0 && s < R * S) *reinterpret_cast<float2*>(ring + C::fpos(RR) + s) = e;
                        } else {
                            slotp[s] = e.x; slotp[s + 1] = e.y;
                            if (slot == 0) {
                                if (s < R * S) ring[C::fpos(RR) + s] = e.x;
                                if (s + 1 < R * S) ring[C::fpos(RR) + s + 1] = e.y;
                            }
                        }
                    }
                }
            } else {
                const long long sA = (long long)(k0 + lag) * S;
                for (int s = lane; s < 32 * S; s += 32) {
                    const long long v = sA + s;
                    const float2 x = (v < V) ? vs.at(v) : make_float2(0.f, 0.f);
                    const int row = s / S, p = s - row * S;
                    const float e = energy_f32(x.x, x.y);
                    ring[C::fpos(32 * slot + row) + p] = e;
                    if (slot == 0 && row < R) ring[C::fpos(RR + row) + p] = e;
                }
            }
            xq_valid = false;
            if (c + 1 < nchunks && chunk_fast(c + 1)) { chunk_issue(c + 1); xq_valid = true; }
            __syncwarp();

            // timing, part 1: exact sliding window sums, lane = (phase wp, row group wg).  Lanes
            // beyond G*S (S = 9, 10) run along on group 0 and store nothing.
            {
                const int g = wact ? wg : 0, p = wact ? wp : 0;
                const float* addp = slotp + (C::PADDED ? g * (R * S + C::PAD) : g * R * S) + p;
                int P0 = 32 * slot + R * g - lag;
                if (P0 < 0) P0 += RR;
                const float* subp = ring + C::fpos(P0) + p;
                const int tcar = C::PADDED ? (R - (P0 % R)) : 2 * R;   // rows i >= tcar sit behind one more pad
                double Eloc[R];
                double x = 0.0;
#pragma unroll
                for (int i = 0; i < R; i++) {
                    if (G * R == 32 || R * g + i < 32) {
                        const float a = addp[i * S];
                        const float sb = subp[i * S + ((C::PADDED && i >= tcar) ? C::PAD : 0)];
                        x = daddr(x, (double)a);                                          // :451
                        Eloc[i] = x;
                        x = dsubr(x, (double)sb);                                         // :576
                    } else Eloc[i] = 0.0;
                }
                // exclusive scan of the group totals over the row groups
                double off = Cw, tot = 0.0;
#pragma unroll
                for (int g2 = 0; g2 < G; g2++) {
                    const double tg = __shfl_sync(0xffffffffu, x, g2 * S + p, 32);
                    if (g2 < g) off = daddr(off, tg);
                    tot = daddr(tot, tg);
                }
                Cw = daddr(Cw, tot);
                if (wact) {
#pragma unroll
                    for (int i = 0; i < R; i++) {
                        const int mrow = R * g + i;
                        if (G * R == 32 || mrow < 32) ebuf[mrow * ES + p] = daddr(off, Eloc[i]);
                    }
                }
            }
            __syncwarp();

            // timing, part 2: lane = row: first maximum (:462), gather (:465), M-th power angle (:474)
            if (lane < nrows) {
                const double* er = ebuf + lane * ES;
                double best = er[0]; int idx = 0;
#pragma unroll
                for (int q = 1; q < S; q++) { const double e = er[q]; if (best < e) { best = e; idx = q; } }
                const int k = k0 + lane;
                if (prm.out_sidx) prm.out_sidx[dg.sym_off + k] = (int16_t)idx;            // :466
                const float2 x = vs.at((long long)k * S + idx);
                th[nbuf + lane] = mth_power_angle_fast(x, (unsigned)M);
                selb[2 + nbuf + lane] = x;
            }
            __syncwarp();
            nbuf += nrows;
            drain();
        }

        // ---- hand the channel's state to the next unit / the next call -----------------------------
        if (sh.st.fit.head != 0 && sh.st.fit.pts == P) fz_normalize_ring(yh, estv, sh.st.fit, P, lane);
        for (int j = lane; j < P; j += 32) __stcg(gring + j, yh[j]);
        if (lane == 0) {
            if (diff && kB > kA) sh.st.last = selb[1];                                    // :489
            const int4* src = reinterpret_cast<const int4*>(&sh.st);
            int4* dst = reinterpret_cast<int4*>(prm.state + ch);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(ChanState) / 16); i++) __stcg(dst + i, src[i]);
            if (sh.st.wraps != wraps0) atomicAdd(&prm.counters->wraps, sh.st.wraps - wraps0);
            if (sh.blocks) atomicAdd(&prm.counters->spec_chunks, (unsigned long long)sh.blocks);
            if (sh.passes) atomicAdd(&prm.counters->spec_misses, (unsigned long long)sh.passes);
            if (sh.seq_blocks) atomicAdd(&prm.counters->seq_channels, (unsigned long long)sh.seq_blocks);
        }
        __syncwarp();
        __threadfence();
        if (lane == 0) st_release(prm.done + ch, ug + 1);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int S>
static cudaError_t launch_fused_t(const LaunchCtx& c, const FusedLaunch& f) {
    using C = FzCfg<S>;
    (void)sizeof(C);
    FusedParams p{};
    p.desc = c.d_desc; p.state = c.d_state; p.ring_base = c.d_ring;
    p.list = f.d_list; p.n_list = f.n_list;
    p.pkts_per_unit = f.pkts_per_unit;
    p.n_units = f.n_list * f.units_per_channel;
    p.ticket = f.d_ticket; p.done = f.d_done;
    p.out_soft = (float2*)c.out_soft; p.out_bits = c.out_bits; p.out_phase = c.out_phase; p.out_sidx = c.out_sidx;
    p.sri_xdelta = c.sri_xdelta;
    p.Pcap = (f.Pmax + 3) & ~3;
    p.lay = fz_layout<S>(f.Amax, p.Pcap);
    p.counters = c.d_counters;
    const size_t smem = (size_t)p.lay.bytes * FZ_WARPS;
    static size_t configured = 0;
    static int ctas_per_sm = 0, n_sm = 0;
    cudaError_t e;
    if (smem > configured || ctas_per_sm == 0) {
        e = cudaFuncSetAttribute(k_fused<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_fused<S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_fused<S>, FZ_WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm < 1) return cudaErrorInvalidConfiguration;
    if (n_sm == 0) {
        int dev = 0;
        e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
    }
    int grid = n_sm * ctas_per_sm;
    const int need = (p.n_units + FZ_WARPS - 1) / FZ_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    c.prof->begin(KID_FUSED, c.stream);
    k_fused<S><<<grid, FZ_WARPS * 32, smem, c.stream>>>(p);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

bool fused_supports(int S, int A, int P) {
    if (!(S == 8 || S == 9 || S == 10 || S == 16)) return false;
    if (A < 1 || A > FUSED_AMAX) return false;
    if (P < 1 || P > FUSED_PMAX) return false;
    return true;
}

cudaError_t launch_fused(const LaunchCtx& c, const FusedLaunch& f) {
    if (f.n_list == 0) return cudaSuccess;
    switch (f.S) {
        case 8:  return launch_fused_t<8>(c, f);
        case 9:  return launch_fused_t<9>(c, f);
        case 10: return launch_fused_t<10>(c, f);
        case 16: return launch_fused_t<16>(c, f);
    }
    return cudaErrorInvalidValue;
}

}  // namespace pskd
