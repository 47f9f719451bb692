// pskd_fused.cu -- the whole demod path of one channel in ONE kernel pass (sm_100a).
//
//   k_fused<S,..> ingest + symbol timing + M-th power angle + unwrap/LinearFit chain + derotate /
//                differential decode / slice, per channel, with every intermediate (energies,
//                window sums, selected samples, angles, unwrapped phases) living in shared memory
//                or registers.  HBM sees only the algorithmic bytes: each IQ sample is read from
//                DRAM once (a second time from L2), each output written once.
//                reference rows: cpp/psk_soft.cpp:380-603, 619-636 and LinearFit :35-185.
//
// Work decomposition: ONE WARP PER UNIT, warp-synchronous (no block barriers).  A unit is a run
// of consecutive emulated BULKIO packets of one channel (>= ~4096 symbols).  A persistent grid
// pulls units from a ticket counter in packet-major order; unit (ch, j) waits for (ch, j-1)
// (its phase-chain state travels through global memory), which was ticketed n_channels earlier.
//
// Per unit the warp streams over its symbols in chunks of 32 rows (1 row = 1 symbol = S samples).
// A chunk needs two blocks of 32 rows of raw samples in shared memory, both copied by cp.async one
// chunk ahead: the LEAD block (the rows entering the 32 energy windows: the stream's new samples,
// from HBM, prefetched into L2 two chunks ahead) and the TRAIL block (the rows leaving the windows,
// numAvg-1 rows behind: an L2 hit, they were the lead block three chunks ago).  There is no energy
// ring: a warp's shared-memory footprint does not depend on numAvg and stays below 9.3 KB, which
// is what lets 24 warps share an SM.
//   timing : lane (phase p, row group g) forms e = f32(re^2+im^2) (std::norm<float>, :448) of its
//            rows' lead and trail samples and from them the exact (double) sliding window sums
//            (:451, :576); a scan over the row groups adds the carried window sum; the sums are
//            transposed through shared memory so that lane = row takes the FIRST maximum over the
//            phases (:462), picks that sample of its row out of the trail block (:465) and forms
//            atan2f(s^M) (:474)
//   chain  : every 128 symbols (or at a packet end): classic-unwrap prediction of the integer
//            unwrap counts, double prefix sums for LinearFit's ySum / xySum, point-wise verification
//            of every count against the reference's rule round((est_{k-1}-theta_k)/2pi) (:477) and
//            repair -- the emitted integers are exactly those of the sequential recursion
//   back   : derotate by -est/M (+pi/4) or divide by the previous sample (:484-501), slice
//            (:503-566); phase, soft and bits are staged in shared memory and written coalesced.
// The front stage (fz_chunk) and the chain + back stage (fz_drain) are non-inlined functions so
// that each gets its own register allocation; what they share lives in the warp's FzCtx in
// shared memory.
#include "pskd_internal.h"
#include "pskd_device.cuh"
#include <cstdlib>

namespace pskd {

bool fused_supports(int S, int A, int P);

#ifndef PSKD_FZ_WARPS
#define PSKD_FZ_WARPS 4
#endif
constexpr int FZ_WARPS = PSKD_FZ_WARPS; // warps (= concurrent units) per CTA
constexpr int FZ_CH = 32;              // rows per chunk
constexpr int FZ_B = 128;              // symbols per chain block (4 per lane)
#ifndef PSKD_FZ_BLOCKS
#define PSKD_FZ_BLOCKS 1
#endif
constexpr int FZ_BLOCKS = PSKD_FZ_BLOCKS;               // chain blocks the front stage collects before the chain/back stage runs
constexpr int FZ_BUF = FZ_BLOCKS * FZ_B + 32;          // capacity of the (theta, sample) block buffer
constexpr int FZ_MAX_ITERS = 16;
#ifndef PSKD_FZ_PF
#define PSKD_FZ_PF 2
#endif
constexpr int FZ_PF = PSKD_FZ_PF;      // lead blocks are prefetched into L2 this many chunks ahead of their cp.async
// The hot stage functions are inlined into ONE loop body, every rare path is a separate non-inlined
// function: the hot code (~19 KB) then sits in a few contiguous runs, which is what the 32 KB L1.5
// instruction cache needs (measured: stage functions as separate calls 14.1 ms per launch, inlined 13.4;
// tools/probe/icache_probe2.cu shows the cliff when the code of an SM's warps outgrows that cache).
#ifdef PSKD_FZ_OUTLINE_HOT
#define FZ_HOT __noinline__
#else
#define FZ_HOT __forceinline__
#endif
#ifndef PSKD_FZ_MIN_CTAS
#define PSKD_FZ_MIN_CTAS 5             // 5 CTAs x 4 warps per SM at 96 registers.  6 CTAs at 80 registers (the shared memory allows it): 4096
                                       // channels 12.3 vs 12.6 ms per launch, but 3072 channels 10.3 vs 9.6 ms (fewer channels than the
                                       // 3552 resident warps: every warp is slower and the channels' chains set the pace)
#endif

struct FzCtx {                         // one per warp, shared memory
    ChanState st;
    FitConst fc;
    const float2* in_mt;               // in - tail_len: virtual sample v >= tail_len lives at in_mt[v]
    const float2* tail;
    const ChanDesc* desc;
    float2* o_soft; float* o_phase; int16_t* o_bits;       // this channel's output rows (or null)
    int16_t* o_sidx;
    uint8_t* o_hard;                   // optional packed hard symbols
    double sri_xdelta;
    long long tail_len, pkt_len, V;
    unsigned long long wraps0;
    int n_pkts, pk1, K, A, M, P, bpb, diff;
    int pkt, pk_hi, kchain, nbuf, cz_valid, unit_done, flags;
    unsigned int passes, seq_blocks, blocks;
    float fP1;
    // front-stage constants of the unit and its loop state (fz_chunk)
    int kA, kB, lag, c_lo, c_hi, nchunks, ch, ug;
    int a16;                           // staged blocks are copied in 16-byte pieces (else 8-byte)
    int end_mid;                       // the unit ends inside its packet: no epilogue, the next unit carries on
    int c;                             // next chunk
    int inflight;                      // chunk c's blocks are already on their way (cp.async)
    // chain + back kernel of the staged path (k_fzs_cb): run-time samplesPerBaud, the unit's end record
    int S_rt, dst, k_begin;
    float est_start_used;
    int n_first, n_last, have_first;   // exact unwrap counts of the unit's first / last symbol (TpEnd)
    float est_pre;                     // estimate before the last packet-end wrap
};

constexpr int fz_align16(int x) { return (x + 15) & ~15; }
// A warp's region is a whole number of 128-byte lines: its staged blocks then start on a line, which cp.async's
// shared-memory writes need to coalesce (a region of 9424 instead of 9472 bytes cost 53 -> 93 LDGSTS wavefronts per chunk
// and 12.5 -> 13.7 ms per launch).
constexpr int fz_align128(int x) { return (x + 127) & ~127; }

// The block buffer keeps the timing-selected samples split into real parts and imaginary parts, selx[FZ_SELN] and
// sely[FZ_SELN] (entry 2 + i = buffered symbol i, entry 1 = the sample before them, for the differential decoder):
// two adjacent symbols' real (imaginary) parts are then ONE aligned 64-bit read = one packed f32x2 operand of the
// per-block stages, with no register shuffling.
constexpr int FZ_SELN = FZ_BUF + 2;
__device__ __forceinline__ float2 fz_sel_get(const float* selx, int e) { return make_float2(selx[e], selx[FZ_SELN + e]); }
__device__ __forceinline__ void fz_sel_put(float* selx, int e, float2 v) { selx[e] = v.x; selx[FZ_SELN + e] = v.y; }

template <int S> struct FzCfg {
    static constexpr int G = 32 / S;                         // row groups (lane = g*S + p)
    static constexpr int R = (32 + G - 1) / G;               // rows per group
    static constexpr int ES = (S & 1) ? S : S + 1;           // row stride of the transposition buffer (doubles)
    static constexpr int CHS = FZ_CH * S;                    // samples per staged block
    static constexpr int NQ16 = (S * 16 + 31) / 32;          // 16-byte pieces per lane per block
    // physical position of logical sample n (= row*S + phase) inside a staged block.  S = 8: rows
    // 8..15 and 24..31 swap places in pairs, which puts the four row groups of one 64-bit read
    // (lane = (phase, group), same row of every group) on disjoint banks.
    // Measured and rejected (r02, profiles/r02_summary.md): additionally permuting the four 16-byte pieces of a row by bits
    // 1-2 of the row (the 128-byte XOR swizzle) turns the 16-way bank conflict of the gather of the timing-selected sample
    // (lane = row, 13 of the kernel's 185 shared-memory wavefronts per chunk) into a 4-way one: -6 % wavefronts, but the
    // eight lane-dependent block positions it needs cost 20 more integer instructions per chunk at 96 registers: 13.1 vs
    // 13.3 ms with them, against 12.5 ms for this layout with its two base pointers and immediate offsets.
    __host__ __device__ static constexpr int phys(int n) { return S == 8 ? (n ^ (((n >> 6) & 1) << 3)) : n; }
};

// compile-time layout of one warp's shared-memory region.  PC = phaseAvg capacity.
template <int S, int PC> struct FzL {
    using C = FzCfg<S>;
    static constexpr int S_STATIC = S;                                     // samplesPerBaud known at compile time
    static constexpr bool TRACK_N = false;                                 // no end records (TpEnd) in the fused kernel
    static constexpr int BLK = C::CHS * 8;                                 // one staged block of raw samples
    static constexpr int OFF_L = 0;                                        // float2 lead[32*S]
    static constexpr int OFF_T = BLK;                                      // float2 trail[32*S]
    static constexpr int OFF_CW = 2 * BLK;                                 // double cw[16]     carried window sum per phase
    static constexpr int OFF_TH = OFF_CW + fz_align16(S * 8);              // float  th[FZ_BUF]
    static constexpr int OFF_SEL = OFF_TH + FZ_BUF * 4;                    // float selx[FZ_SELN], sely[FZ_SELN]; [1] = previous sample
    static constexpr int OFF_YH = fz_align16(OFF_SEL + (FZ_BUF + 2) * 8);  // float  yh[PC]   y history, logical order
    static constexpr int OFF_CTX = fz_align16(OFF_YH + PC * 4);
    static constexpr int OFF_CZ = fz_align16(OFF_CTX + (int)sizeof(FzCtx));// double cz[PC + 1], ends where ALIAS starts
    static constexpr int OFF_ALIAS = OFF_CZ + fz_align16((PC + 1) * 8);
    static constexpr int E_BYTES = 32 * C::ES * 8;                         // front stage: window sums [32][ES]
    static constexpr int C_BYTES = FZ_B * 8 + FZ_B * 4 + (FZ_B + 4) * 4;   // chain stage: prefix block, y block, est block
    static constexpr int BYTES = fz_align128(OFF_ALIAS + fz_align16(E_BYTES > C_BYTES ? E_BYTES : C_BYTES));
    static_assert(BYTES % 128 == 0 && OFF_T % 128 == 0, "staged blocks must start on a 128-byte line");
};

struct FusedParams {
    const ChanDesc* desc; ChanState* state; float* ring_base;
    const int* list; int n_list;       // channels served by this launch
    int n_units;                       // n_list * max units per channel
    int pkts_per_unit;
    int parts_per_pkt;                 // > 1: a unit is one of these parts of ONE packet (pkts_per_unit == 1)
    int* ticket;                       // unit ticket counter (zeroed before the launch)
    int* done;                         // [n_channels] units completed per channel (zeroed before the launch)
    float2* out_soft; int16_t* out_bits; float* out_phase; int16_t* out_sidx; uint8_t* out_hard;
    double sri_xdelta;
    DevCounters* counters;
};

extern __shared__ __align__(128) unsigned char fz_smem[];

__device__ __forceinline__ void fz_prefetch_line(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ int fz_lane() {
#ifdef PSKD_FZ_LANE_TID
    return threadIdx.x & 31;
#endif
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));     // volatile: kept in a register, not re-derived from tid at every use
    return l;
}
__device__ __forceinline__ void fz_cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
#ifdef PSKD_FZ_CPASYNC_CA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");     // through L1 (experiment: 14.1 vs 12.8 ms on the bench bank, off)
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void fz_cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fz_cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fz_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// e = f32(f32(re*re) + f32(im*im)) (std::norm<float>, cpp/psk_soft.cpp:448) with the two products in one
// packed instruction (FMUL2: both halves rounded to nearest like the scalar multiply)
__device__ __forceinline__ float fz_energy(float2 v) {
    float px, py;
    asm("{\n\t.reg .b64 a, r;\n\tmov.b64 a, {%2, %3};\n\tmul.rn.f32x2 r, a, a;\n\tmov.b64 {%0, %1}, r;\n\t}"
        : "=f"(px), "=f"(py) : "f"(v.x), "f"(v.y));
    return faddr(px, py);
}

// ---------------------------------------------------------------------------------------------
// Packed single precision (Blackwell f32x2: one FMUL2 / FADD2 / FFMA2 does the operation on both
// halves of a 64-bit register pair, each half rounded to nearest exactly like the scalar
// instruction).  The per-symbol stages that run once per block put TWO symbols in a lane with
// these: half the issue slots for the same arithmetic.
// ptxas contracts mul.rn.f32x2 + add/sub.rn.f32x2 into one FFMA2 (it does not for the scalar forms) and folds an
// fma by +1 back into that, so sums and differences of products that the reference rounds separately are written
// as a subtraction through an fma by -1 (a sum a + b as a - (-b)).
// ---------------------------------------------------------------------------------------------
typedef unsigned long long fz_p2;
__device__ __forceinline__ fz_p2 p2_make(float lo, float hi) { return ((fz_p2)__float_as_uint(hi) << 32) | (fz_p2)__float_as_uint(lo); }
__device__ __forceinline__ fz_p2 p2_dup(float c) { return p2_make(c, c); }
__device__ __forceinline__ void p2_get(fz_p2 v, float& lo, float& hi) { lo = __uint_as_float((unsigned)v); hi = __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ fz_p2 p2_mul(fz_p2 a, fz_p2 b) { fz_p2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ fz_p2 p2_fma(fz_p2 a, fz_p2 b, fz_p2 c) { fz_p2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ fz_p2 p2_sub(fz_p2 a, fz_p2 b) { return p2_fma(b, p2_dup(-1.0f), a); }    // f32(a - b)

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

// literal recursion (lane 0): fill-up, around the 2^20-call re-sum, and when the scan path gives up
static __device__ __noinline__ void fz_block_sequential(FzCtx& cx, float* yh, const float* th, float* estv, int nb) {
    SmemRing ring{yh};
    ChanState st = cx.st;
    long long n = 0;
    for (int i = 0; i < nb; i++) {
        float y = unwrap_against(st.est, th[i], &n);
        if (i == 0 && !cx.have_first) { cx.n_first = (int)n; cx.have_first = 1; }
        st.est = fit_next(st.fit, ring, y);
        estv[i] = st.est;
    }
    if (nb > 0) cx.n_last = (int)n;
    cx.st = st;
    if (st.fit.pts == st.fit.n && st.fit.pts > 1) cx.fc = fit_const(st.fit);
}

// packet prologue of the phase estimator (cpp/psk_soft.cpp:393-426), lane 0
static __device__ __noinline__ void fz_prologue(FzCtx& cx, float* yh) {
    SmemRing r{yh};
    ChanState st = cx.st;
    int flags = cx.flags;
    chain_packet_prologue(st, r, *cx.desc, cx.sri_xdelta, flags);
    cx.st = st; cx.flags = flags;
    if (st.fit.pts == st.fit.n && st.fit.pts > 1) cx.fc = fit_const(st.fit);
}

// packet epilogue (cpp/psk_soft.cpp:592-603), lane 0; returns through cx.cz_valid whether the history moved
static __device__ __noinline__ void fz_epilogue(FzCtx& cx, float* yh) {
    SmemRing r{yh};
    ChanState st = cx.st;
    const unsigned long long w0 = st.wraps;
    cx.est_pre = st.est;
    chain_packet_epilogue(st, r, cx.M);
    if (st.wraps != w0) { cx.st = st; cx.cz_valid = 0; }
}

static __device__ __noinline__ void fz_resum(FzCtx& cx, float* yh) { SmemRing r{yh}; fit_resum(cx.st.fit, r); }   // :51-52
// The rare paths below are reached from the hot, inlined stage code.  They take the warp's shared-memory OFFSET, not
// pointers: a pointer parameter into shared memory makes the caller form generic addresses (S2R + 64-bit adds) on the hot
// path, whether or not the rare path is taken (measured: ~9 of 320 instructions per 32 symbols).
template <class L> static __device__ __noinline__ void fz_normalize_ring_w(const unsigned wofs, const int lane) {
    unsigned char* wb = fz_smem + wofs;
    FzCtx& cx = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    fz_normalize_ring(reinterpret_cast<float*>(wb + L::OFF_YH), reinterpret_cast<float*>(wb + L::OFF_ALIAS + FZ_B * 8) + FZ_B,
                      cx.st.fit, cx.P, lane);
}
template <class L> static __device__ __noinline__ void fz_rebuild_cz_w(const unsigned wofs, const int lane) {
    unsigned char* wb = fz_smem + wofs;
    FzCtx& cx = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    fz_rebuild_cz(reinterpret_cast<const float*>(wb + L::OFF_YH), reinterpret_cast<double*>(wb + L::OFF_ALIAS) - (cx.P + 1), cx.P, lane);
}
template <class L> static __device__ __noinline__ void fz_resum_w(const unsigned wofs) {
    unsigned char* wb = fz_smem + wofs;
    fz_resum(*reinterpret_cast<FzCtx*>(wb + L::OFF_CTX), reinterpret_cast<float*>(wb + L::OFF_YH));
}

static __device__ __noinline__ int fz_unwrap_count_slow(float est_prev, float theta) {
    const double dlt = dsubr((double)est_prev, (double)theta);
    return (int)(long long)round(__ddiv_rn(dlt, PSKD_M_2PI));
}
// the reference's unwrap count (cpp/psk_soft.cpp:477), fast form: multiply by 1/2pi and round to
// nearest; `bad` is raised whenever the quotient is within 1e-7 of a half-integer (where the
// division's last bit or the tie rule could matter) -- the caller then takes the literal
// division + round().
__device__ __forceinline__ int fz_unwrap_count(float est_prev, float theta, bool& bad) {
    const double dlt = dsubr((double)est_prev, (double)theta);       // exact
    const double q = dmulr(dlt, 0.15915494309189535);
    const double t = daddr(q, 6755399441055744.0);                   // 1.5 * 2^52: round to nearest integer
    const double qr = dsubr(t, 6755399441055744.0);
    const double fr = fabs(dsubr(q, qr));
    bad = bad || fr > 0.4999999 || !(fabs(q) < 1.0e9);
    return __double2loint(t);
}

// ---------------------------------------------------------------------------------------------
// The derotation phasor (cpp/psk_soft.cpp:499, std::polar(1.0f, pc)) of two symbols at once: Cody-Waite
// reduction by 2 pi in packed arithmetic, then the hardware's sine / cosine (MUFU, |error| <= 2^-21.4 on
// [-pi, pi]): |error| <= 6e-7 on the phasor, i.e. on the soft decisions relative to their magnitude
// (tolerance 1e-4; glibc's and CUDA's sincosf are "about an ulp" functions); a decision that close to a
// slicer boundary is inside the slicer's guard band (1e-5) and is redone literally.  |pc| >= 1e4 or
// non-finite arguments take the library call.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fz_sincos2(fz_p2 X, fz_p2& SN, fz_p2& CS, bool& badA, bool& badB) {
    float xa, xb;
    p2_get(X, xa, xb);
    badA = badA || !(fabsf(xa) < 1.0e4f);
    badB = badB || !(fabsf(xb) < 1.0e4f);
    const fz_p2 T = p2_fma(X, p2_dup(0.15915494309189535f), p2_dup(12582912.0f));      // x / 2pi, rounded to an integer in the low bits
    const fz_p2 K = p2_fma(T, p2_dup(1.0f), p2_dup(-12582912.0f));
    fz_p2 R = p2_fma(K, p2_dup(-6.2831854820251465f), X);                               // 2 pi = 6.2831854820251465 - 1.7484555e-7
    R = p2_fma(K, p2_dup(1.7484555e-7f), R);
    float ra, rb;
    p2_get(R, ra, rb);
    SN = p2_make(__sinf(ra), __sinf(rb));
    CS = p2_make(__cosf(ra), __cosf(rb));
}

// 8-PSK slicer (cpp/psk_soft.cpp:547-563) as a sector test against the rays at odd multiples of
// pi/8; inside a guard band of the rays (or for zero / non-finite input) `bad` asks for the literal
// atan2f -> /pi*4 -> roundf path.  ONE guard comparison: zero input (d1 = d2 = g = 0), NaN and infinity
// (g or both d are NaN / inf) all fail min(|d1|, |d2|) > g.
__device__ __forceinline__ unsigned fz_slice8_guard(float x, float y, bool& bad) {
    const float a = fabsf(x), b = fabsf(y);
    const float T = 0.41421356237309503f;       // tan(pi/8)
    const float d1 = b - T * a, d2 = a - T * b;
    const float g = 1.0e-5f * (a + b);
    bad = bad || !(fminf(fabsf(d1), fabsf(d2)) > g);
    const unsigned quad = (x > 0.0f) ? ((y > 0.0f) ? 1u : 7u) : ((y > 0.0f) ? 3u : 5u);
    const unsigned ax = (x > 0.0f) ? 0u : 4u, ay = (y > 0.0f) ? 2u : 6u;
    return (d1 < 0.0f) ? ax : ((d2 < 0.0f) ? ay : quad);
}

// sample / last (libgcc __divsc3: evaluated in double, narrowed to float; cpp/psk_soft.cpp:488) with
// one reciprocal: the double quotients are within 2 ulp(double) of the divided ones, i.e. the
// narrowed floats differ in ~1e-8 of the cases by one float ulp.  Zero / non-finite operands or
// results ask for the literal path.
__device__ __forceinline__ float2 fz_cdiv_fast(float2 n, float2 dn, bool& bad) {
    const double a = n.x, b = n.y, c = dn.x, d = dn.y;
    const double denom = fma(c, c, d * d);
    const double r = 1.0 / denom;
    const float x = __double2float_rn((a * c + b * d) * r);
    const float y = __double2float_rn((b * c - a * d) * r);
    bad = bad || !(denom > 1.0e-300 && denom < 1.0e300) || !(fabsf(x) < 3.0e38f) || !(fabsf(y) < 3.0e38f);
    return make_float2(x, y);
}

// the literal path of one symbol: libgcc-style division, library sincosf, checked complex
// multiply, atan2f slicer -- for the symbols whose fast evaluation raised a flag
template <class L>
static __device__ __noinline__ void fz_back_literal(const unsigned wofs, const int i, int M, int bpb, int diff) {
    unsigned char* wb = fz_smem + wofs;                  // (offsets, not pointers: see the note at fz_normalize_ring_w)
    const float*  th   = reinterpret_cast<const float*>(wb + L::OFF_TH);
    const float* selx = reinterpret_cast<const float*>(wb + L::OFF_SEL);
    float2* c_stage = reinterpret_cast<float2*>(wb + L::OFF_ALIAS) + i;
    short*  b_stage = reinterpret_cast<short*>(wb + L::OFF_ALIAS + FZ_B * 8) + i * bpb;
    unsigned char* h_stage = wb + L::OFF_ALIAS + FZ_B * 8 + FZ_B * 6 + i;
    const float2 s = fz_sel_get(selx, 2 + i), prev = fz_sel_get(selx, 1 + i);
    const float est = th[i];
    float2 x = s;
    if (diff) x = cdiv_f32(s, prev);
    const float pc = phase_correction(est, M, diff != 0);
    const float2 c = derotate(x, pc);
    *c_stage = c;
    const unsigned b = slice_bits(c, bpb);
    for (int j = 0; j < bpb; j++) b_stage[j] = (short)((b >> j) & 1u);
    *h_stage = (unsigned char)b;
}

// derotate / differential decode / slice (cpp/psk_soft.cpp:484-566), specialised on bits per symbol and on
// differential decoding.  QPSK <=> BPB == 2 (constelationSize 4).  Flagged symbols take fz_back_literal.
// TWO adjacent symbols per lane (symbols 2j and 2j + 1, j = 32 * trip + lane): the loads are 64-bit reads of the two
// samples' real parts, imaginary parts and estimates, the phasor / derotation arithmetic is packed (FMUL2 / FFMA2), the
// phasor comes from the hardware sine / cosine after a Cody-Waite reduction (fz_sincos2), and the results leave in one
// 128-bit store (soft) and 32-bit stores (bits): 90 instructions per 64 symbols where the one-symbol-per-lane form took 168.
// A ROLLED loop, two trips per 128-symbol block: the code runs once per block, so its cost is dominated by
// instruction fetch (unrolling the two trips: 13.1 vs 12.6 ms per launch).
// The buffers are addressed from the warp's shared-memory offset (not through generic pointer parameters: those cost
// ~15 instructions per 32 symbols of address conversion and turn every access into a generic LD / ST).
template <class L, int BPB, bool DIFF, bool HARD>
static __device__ __forceinline__ void fz_back_pairs_body(const unsigned wofs, const int lane, const int M, const int m) {
    unsigned char* wb = fz_smem + wofs;
    const float2* __restrict__ th2  = reinterpret_cast<const float2*>(wb + L::OFF_TH);
    const float*  __restrict__ selx = reinterpret_cast<const float*>(wb + L::OFF_SEL);
    const float2* __restrict__ sx2  = reinterpret_cast<const float2*>(selx + 2);                   // real parts of symbols 2j, 2j + 1
    const float2* __restrict__ sy2  = reinterpret_cast<const float2*>(selx + FZ_SELN + 2);         // imaginary parts
    float4*   __restrict__ cst4 = reinterpret_cast<float4*>(wb + L::OFF_ALIAS);                    // soft staging [FZ_B]
    unsigned* __restrict__ bst  = reinterpret_cast<unsigned*>(wb + L::OFF_ALIAS + FZ_B * 8);       // bits staging [FZ_B * 3] int16
    unsigned short* __restrict__ hst2 = reinterpret_cast<unsigned short*>(wb + L::OFF_ALIAS + FZ_B * 8 + FZ_B * 6);   // packed hard symbols
    // -1 / M (:494): BPB > 0 <=> constelationSize == 1 << BPB (bpb_of, pskd_api.cu): a compile-time constant for BPSK / QPSK / 8-PSK
    const float ninv_m = (BPB > 0) ? (-1.0f / (float)(1 << (BPB > 0 ? BPB : 1))) : (-1.0f / (float)M);
    const bool inexact = (BPB == 0 && (M & (M - 1)) != 0);                 // -est/M not an exact multiply
    const int npairs = (m + 1) >> 1;
#pragma unroll 1
    for (int j = lane; j < ((npairs + 31) & ~31); j += 32) {
        bool badA = inexact, badB = inexact;
        const float2 vx = sx2[j], vy = sy2[j];
        float2 sA = make_float2(vx.x, vy.x), sB = make_float2(vx.y, vy.y);
        fz_p2 SN, CS;
        if (DIFF) {
            const float2 pA = fz_sel_get(selx, 1 + 2 * j);
            const float2 rawA = sA;
            sA = fz_cdiv_fast(rawA, pA, badA);                                                     // :488
            sB = fz_cdiv_fast(sB, rawA, badB);
            // pc = 0 (+ pi/4 for QPSK, :497-498): the phasor is a constant
            SN = (BPB == 2) ? p2_dup(0.7071067966408575f) : p2_dup(0.0f);
            CS = (BPB == 2) ? p2_dup(0.7071067657322372f) : p2_dup(1.0f);
        } else {
            const float2 e2 = th2[j];
            fz_p2 PCV = p2_mul(p2_make(e2.x, e2.y), p2_dup(ninv_m));                               // :494 (exact for M = 2^n)
            if (BPB == 2) {                                                                        // :497-498
                float pa, pb;
                p2_get(PCV, pa, pb);
                PCV = p2_make(__double2float_rn(daddr((double)pa, PSKD_M_PI_4)), __double2float_rn(daddr((double)pb, PSKD_M_PI_4)));
            }
            fz_sincos2(PCV, SN, CS, badA, badB);                                                   // :499
        }
        const fz_p2 X = p2_make(sA.x, sB.x), Y = p2_make(sA.y, sB.y);
        const fz_p2 XO = p2_sub(p2_mul(X, CS), p2_mul(Y, SN));                                     // :500-501, unfused
        const fz_p2 YO = p2_sub(p2_mul(X, SN), p2_mul(p2_mul(Y, p2_dup(-1.0f)), CS));              // xs - (-(y cs)): stays unfused
        float xA, xB, yA, yB;
        p2_get(XO, xA, xB);
        p2_get(YO, yA, yB);
        if (BPB != 3) {                                                                            // NaN / inf: __mulsc3 recovery decides
            badA = badA || !(fabsf(xA) + fabsf(yA) < 3.0e38f);                                     // (8-PSK: the slicer's guard test covers it)
            badB = badB || !(fabsf(xB) + fabsf(yB) < 3.0e38f);
        }
        cst4[j] = make_float4(xA, yA, xB, yB);
        if (BPB == 3) {
            const unsigned bA = fz_slice8_guard(xA, yA, badA), bB = fz_slice8_guard(xB, yB, badB);
            // one int16 per bit, LSB first (:559-563): b * 0x40008001 puts bit 0 at 0, bit 1 at 16, bit 2 at 32
            const unsigned long long wA = (unsigned long long)bA * 0x40008001ull, wB = (unsigned long long)bB * 0x40008001ull;
            const unsigned loA = (unsigned)wA & 0x00010001u, hiA = (unsigned)(wA >> 32) & 1u;
            const unsigned loB = (unsigned)wB & 0x00010001u, hiB = (unsigned)(wB >> 32) & 1u;
            bst[3 * j] = loA;
            bst[3 * j + 1] = hiA | (loB << 16);
            bst[3 * j + 2] = (loB >> 16) | (hiB << 16);
            if (HARD) hst2[j] = (unsigned short)(bA | (bB << 8));
        } else if (BPB == 1) {
            const unsigned bA = (xA < 0.0f) ? 1u : 0u, bB = (xB < 0.0f) ? 1u : 0u;                 // :512
            bst[j] = bA | (bB << 16);
            if (HARD) hst2[j] = (unsigned short)(bA | (bB << 8));
        } else if (BPB == 2) {                                                                     // :523-526 (float -> bool, sic)
            const unsigned a0 = ((xA != 0.0f) != (yA != 0.0f)) ? 1u : 0u, a1 = (yA != 0.0f) ? 0u : 1u;
            const unsigned b0 = ((xB != 0.0f) != (yB != 0.0f)) ? 1u : 0u, b1 = (yB != 0.0f) ? 0u : 1u;
            reinterpret_cast<uint2*>(bst)[j] = make_uint2(a0 | (a1 << 16), b0 | (b1 << 16));
            if (HARD) hst2[j] = (unsigned short)((a0 | (a1 << 1)) | ((b0 | (b1 << 1)) << 8));
        }
        const int iA = 2 * j;
        badA = badA && iA < m;
        badB = badB && iA + 1 < m;
        if (__any_sync(0xffffffffu, badA || badB)) {    // rare: literal evaluation of the flagged symbols
            if (badA) fz_back_literal<L>(wofs, iA, M, BPB, DIFF ? 1 : 0);
            if (badB) fz_back_literal<L>(wofs, iA + 1, M, BPB, DIFF ? 1 : 0);
        }
    }
}

// out of line (its own register allocation, compact code): what the stage code calls when the variant is only known at run time
template <class L, int BPB, bool DIFF, bool HARD>
static __device__ __noinline__ void fz_back_pairs(const unsigned wofs, const int lane, const int M, const int m) {
    fz_back_pairs_body<L, BPB, DIFF, HARD>(wofs, lane, M, m);
}

// ---------------------------------------------------------------------------------------------
// The chain + back stage runs as three non-inlined functions (own register allocations, compact
// code): fz_drain (packet bookkeeping, block sizing, buffer compaction) calls, per block of up to
// FZ_B buffered symbols, fz_chain_fast (scan / verify / repair; the estimates land in th[], which
// doubles as the staging of phase_dataFloat_out) and fz_back_block (derotate / slice / stores).
// ---------------------------------------------------------------------------------------------

// phase chain of one block of m symbols at buffer offset 0 (full history window, cx.st.fit.pts == P):
// cpp/psk_soft.cpp:476-482 with LinearFit::next (:48-87).  Returns false if the counts did not settle
// within FZ_MAX_ITERS passes (the caller then runs the literal recursion); on success the state, the
// history (yh, cz) and th[0..m) = est are updated.
template <class L>
static __device__ FZ_HOT bool fz_chain_fast(const unsigned wofs, const int m)
{
    unsigned char* wb = fz_smem + wofs;
    float*  th   = reinterpret_cast<float*>(wb + L::OFF_TH);
    float*  yh   = reinterpret_cast<float*>(wb + L::OFF_YH);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    double* czblk = reinterpret_cast<double*>(wb + L::OFF_ALIAS);                    // cz[P+1 ...]
    float*  yblk = reinterpret_cast<float*>(wb + L::OFF_ALIAS + FZ_B * 8);
    float*  estv = yblk + FZ_B;
    const int lane = fz_lane();
    const int P = cx.P;
    double* cz = czblk - (P + 1);
    const int i0 = lane * 4;

    if (cx.st.fit.head != 0) { fz_normalize_ring_w<L>(wofs, lane); if (lane == 0) cx.cz_valid = 0; __syncwarp(); }
    if (!cx.cz_valid) { fz_rebuild_cz_w<L>(wofs, lane); if (lane == 0) cx.cz_valid = 1; __syncwarp(); }

    float4 t4 = *reinterpret_cast<const float4*>(th + i0);
    if (m < FZ_B) {                    // a short block: what lies behind it in the buffer may be anything (never written,
        if (i0 + 0 >= m) t4.x = 0.0f;  // NaN, huge) and would poison the scans below (exclusive = inclusive - own)
        if (i0 + 1 >= m) t4.y = 0.0f;
        if (i0 + 2 >= m) t4.z = 0.0f;
        if (i0 + 3 >= m) t4.w = 0.0f;
    }
    const float tl[4] = {t4.x, t4.y, t4.z, t4.w};
    const float xdelta = cx.st.fit.xdelta;
    const float fP1 = cx.fP1;
    const double xd = (double)xdelta;
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
            if (v == 0 && lane == 0) {                                         // :477 against the carried estimate
                bool knife = false;
                dn = fz_unwrap_count(cx.st.est, tl[0], knife);
                if (knife) dn = fz_unwrap_count_slow(cx.st.est, tl[0]);
            }
            run += dn; nloc[v] = run;
        }
        const int off = warp_scan_int(run, lane) - run;
#pragma unroll
        for (int v = 0; v < 4; v++) nloc[v] += off;
    }
    const int last = m - 1;
    int iter = 0;
    bool done = false;
    float yl[4], el[4];
    double Yv = 0.0, Xv = 0.0;                       // sums after the block's last symbol (owner lane)
    while (true) {
        double Cl[4], hi0;
        {
            double run = 0.0;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const float y = __double2float_rn(daddr((double)tl[v], dmulr((double)nloc[v], PSKD_M_2PI)));   // :478,481
                yl[v] = y; run = daddr(run, (double)y); Cl[v] = run;
            }
            hi0 = daddr(HPP, dsubr(warp_scan_dbl(run, lane), run));          // cz[P+i0]
#pragma unroll
            for (int v = 0; v < 4; v++) Cl[v] = daddr(hi0, Cl[v]);           // cz[P+1+i0+v]
        }
        *reinterpret_cast<float4*>(yblk + i0) = make_float4(yl[0], yl[1], yl[2], yl[3]);
        *reinterpret_cast<double2*>(czblk + i0) = make_double2(Cl[0], Cl[1]);
        *reinterpret_cast<double2*>(czblk + i0 + 2) = make_double2(Cl[2], Cl[3]);
        __syncwarp();
        double Ys[4], Xl[4];
        double trun = 0.0;
#pragma unroll
        for (int v = 0; v < 4; v++) {
            const double hi = (v == 0) ? hi0 : Cl[v - 1];                    // cz[P+i]
            const double W = dsubr(hi, cz[i0 + v + 1]);                      // ySum after :70
            const double a = dmulr(xd, W);                                   // :72
            const double T = (double)fmulr(fmulr(yl[v], fP1), xdelta);       // :78
            trun = daddr(trun, dsubr(T, a)); Xl[v] = trun;
            Ys[v] = daddr(W, (double)yl[v]);                                 // :75
        }
        const double xoff = daddr(cx.st.fit.xySum, dsubr(warp_scan_dbl(trun, lane), trun));
        {
            const FitConst fc = cx.fc;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                Xl[v] = daddr(xoff, Xl[v]);
                el[v] = fit_eval_fast(fc, Ys[v], Xl[v], nullptr, nullptr);   // :135-162
            }
        }
        Yv = Ys[3]; Xv = Xl[3];                               // a full block ends on the owner lane's last symbol
        if (m != FZ_B) {                                      // (overriding assignments, not Ys[last & 3]: an index the compiler
            const int lv = last & 3;                          // cannot resolve puts the arrays into local memory on the hot path)
            if (lv < 3) { Yv = Ys[2]; Xv = Xl[2]; }
            if (lv < 2) { Yv = Ys[1]; Xv = Xl[1]; }
            if (lv < 1) { Yv = Ys[0]; Xv = Xl[0]; }
        }
        // verify every predicted n against the reference's rule (:477) with est_{i-1}
        const float eprev = __shfl_up_sync(0xffffffffu, el[3], 1);
        int mymis = 0x7fffffff, mydelta = 0;
        {
            // the usual outcome -- every predicted count is the reference's -- is established without forming the counts:
            // |(est_{i-1} - theta_i) / 2pi - n_i| < 0.4999999 means round() of that quotient is n_i, and not by a hair (the
            // quotient by multiplication is within 1e-15 of the divided one).  One vote; only a count that is off, or next
            // to a half-integer, pays for the exact counts and for locating the first symbol that disagrees.
            bool differs = false;
#pragma unroll
            for (int v = 0; v < 4; v++) {
                const double dlt = dsubr((double)((v == 0) ? eprev : el[v - 1]), (double)tl[v]);       // exact
                const double dq = dsubr(dmulr(dlt, 0.15915494309189535), (double)nloc[v]);
                differs = differs || (!(fabs(dq) < 0.4999999) && i0 + v < m && (v > 0 || lane > 0));
            }
            if (!__any_sync(0xffffffffu, differs)) { done = true; break; }
            int nt[4];
            bool knife = false;
#pragma unroll
            for (int v = 0; v < 4; v++) nt[v] = fz_unwrap_count((v == 0) ? eprev : el[v - 1], tl[v], knife);
            if (__any_sync(0xffffffffu, knife)) {           // rare: a quotient next to a half-integer
#pragma unroll
                for (int v = 0; v < 4; v++) nt[v] = fz_unwrap_count_slow((v == 0) ? eprev : el[v - 1], tl[v]);
            }
#pragma unroll
            for (int v = 3; v >= 0; v--) {
                const int i = i0 + v;
                if (i >= 1 && i < m && nt[v] != nloc[v]) { mymis = i; mydelta = nt[v] - nloc[v]; }
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
    if (iter && lane == 0) cx.passes += (unsigned)iter;
    if (!done) {
        if (lane == 0) { cx.seq_blocks++; cx.cz_valid = 0; }
        __syncwarp();
        return false;
    }
    if ((last >> 2) == lane) {
        const FitConst fc = cx.fc;
        FitState& f = cx.st.fit;
        float mm, bb;
        cx.st.est = fit_eval_fast(fc, Yv, Xv, &mm, &bb);
        f.ySum = Yv; f.xySum = Xv; f.m = mm; f.b = bb; f.count += m;
        if (L::TRACK_N) {                      // the verified (= the reference's) count of the block's last symbol
            int nl = nloc[3];
            if (m != FZ_B) {
                const int lv = last & 3;
                if (lv < 3) nl = nloc[2];
                if (lv < 2) nl = nloc[1];
                if (lv < 1) nl = nloc[0];
            }
            cx.n_last = nl;
        }
    }
    if (L::TRACK_N && lane == 0 && !cx.have_first) { cx.n_first = nloc[0]; cx.have_first = 1; }
    // phase_dataFloat_out (:482) is staged in th[0..m): the angles are consumed
    if (i0 + 3 < m) *reinterpret_cast<float4*>(th + i0) = make_float4(el[0], el[1], el[2], el[3]);
    else if (i0 < m) {
#pragma unroll
        for (int v = 0; v < 3; v++) if (i0 + v < m) th[i0 + v] = el[v];
    }
    __syncwarp();
    // new history = last P of (history ++ block): shift the prefix and the values by m
    if (m > P) {
        // every full block: the new history comes from the block alone, and what is read (the block's prefix sums and
        // values) does not overlap what is written (the history areas): no read-all-then-write passes
        const double* src = czblk + (m - P - 1);             // src[j] = cz[m + j]
        const float* ysrc = yblk + (m - P);
        const double czm_b = src[0];
        for (int j = lane; j <= P; j += 32) {
            cz[j] = dsubr(src[j], czm_b);
            if (j < P) yh[j] = ysrc[j];
        }
        __syncwarp();
        return true;
    }
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
    return true;
}

// back stage of one block of m symbols at buffer offset 0: derotate / differential decode / slice
// (cpp/psk_soft.cpp:484-566) from selb[] (samples) and th[] (estimates) into the staging buffers
// (soft, bits: the chain's buffers are dead by now), then the coalesced stores of phase / soft /
// bits for symbols [kchain, kchain + m).
// BK >= 0: every channel of the launch has bits-per-symbol BK / 2 and differentialDecoding BK & 1, and no packed hard symbols are
// asked for (launch_fused_t checks): the back loop is inlined into the kernel's loop body -- no dispatch, no call, and the hot code is
// ONE contiguous run (see the placement note in profiles/r02_summary.md).  BK < 0: dispatch at run time.
template <class L, int BK = -1>
static __device__ FZ_HOT void fz_back_block(const unsigned wofs, const int m)
{
    unsigned char* wb = fz_smem + wofs;
    float*  th   = reinterpret_cast<float*>(wb + L::OFF_TH);
    float*  selx = reinterpret_cast<float*>(wb + L::OFF_SEL);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    float2* cst  = reinterpret_cast<float2*>(wb + L::OFF_ALIAS);              // soft staging [FZ_B]
    short* bstage = reinterpret_cast<short*>(wb + L::OFF_ALIAS + FZ_B * 8);   // bits staging [FZ_B * 3]
    const int lane = fz_lane();
    const int M = cx.M, bpb = cx.bpb, kchain = cx.kchain;
    const bool diff = cx.diff != 0;
    uint8_t* o_hard = cx.o_hard;
    const bool hard = (BK < 0) && o_hard != nullptr && bpb > 0;
    if (BK >= 0) {
        fz_back_pairs_body<L, (BK >= 0 ? BK / 2 : 0), (BK >= 0 && (BK & 1) != 0), false>(wofs, lane, M, m);
    } else if (!hard) {
        switch (bpb * 2 + (diff ? 1 : 0)) {
            case 6: fz_back_pairs<L, 3, false, false>(wofs, lane, M, m); break;
            case 7: fz_back_pairs<L, 3, true, false>(wofs, lane, M, m); break;
            case 4: fz_back_pairs<L, 2, false, false>(wofs, lane, M, m); break;
            case 5: fz_back_pairs<L, 2, true, false>(wofs, lane, M, m); break;
            case 2: fz_back_pairs<L, 1, false, false>(wofs, lane, M, m); break;
            case 3: fz_back_pairs<L, 1, true, false>(wofs, lane, M, m); break;
            case 0: fz_back_pairs<L, 0, false, false>(wofs, lane, M, m); break;
            default: fz_back_pairs<L, 0, true, false>(wofs, lane, M, m); break;
        }
    } else {
        switch (bpb * 2 + (diff ? 1 : 0)) {
            case 6: fz_back_pairs<L, 3, false, true>(wofs, lane, M, m); break;
            case 7: fz_back_pairs<L, 3, true, true>(wofs, lane, M, m); break;
            case 4: fz_back_pairs<L, 2, false, true>(wofs, lane, M, m); break;
            case 5: fz_back_pairs<L, 2, true, true>(wofs, lane, M, m); break;
            case 2: fz_back_pairs<L, 1, false, true>(wofs, lane, M, m); break;
            default: fz_back_pairs<L, 1, true, true>(wofs, lane, M, m); break;
        }
    }
    int16_t* o_bits = cx.o_bits;
    __syncwarp();
    // coalesced stores; a full block at a 16-byte aligned position goes out in 128-bit pieces
    {
        const bool full = (m == FZ_B);
        float* o_phase = cx.o_phase;
        if (o_phase) {
            float* o = o_phase + kchain;
            if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                __stcs(reinterpret_cast<float4*>(o) + lane, reinterpret_cast<const float4*>(th)[lane]);
            } else {
                o += lane;
                const float* si = th + lane;
#pragma unroll 1
                for (int i = lane; i < m; i += 32, o += 32, si += 32) __stcs(o, *si);
            }
        }
        float2* o_soft = cx.o_soft;
        if (o_soft) {
            float2* o = o_soft + kchain;
            if (full && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                const float4* s4 = reinterpret_cast<const float4*>(cst) + lane;
                float4* o4 = reinterpret_cast<float4*>(o) + lane;
                const float4 v0 = s4[0], v1 = s4[32];
                __stcs(o4, v0); __stcs(o4 + 32, v1);
            } else {
                o += lane;
                const float2* si = cst + lane;
#pragma unroll 1
                for (int i = lane; i < m; i += 32, o += 32, si += 32) __stcs(o, *si);
            }
        }
        if (bpb > 0 && o_bits) {
            int16_t* o = o_bits + (long long)kchain * bpb;
            const int nsh = m * bpb;
            if (full && (reinterpret_cast<uintptr_t>(o) & 7) == 0) {
                const uint2* s2 = reinterpret_cast<const uint2*>(bstage) + lane;
                uint2* o2 = reinterpret_cast<uint2*>(o) + lane;
#pragma unroll
                for (int q = 0; q < 3; q++) if (q < bpb) __stcs(o2 + 32 * q, s2[32 * q]);
            } else if ((reinterpret_cast<uintptr_t>(o) & 3) == 0) {
                const unsigned* s32 = reinterpret_cast<const unsigned*>(bstage) + lane;
                unsigned* o32 = reinterpret_cast<unsigned*>(o) + lane;
                const int nw = nsh >> 1;                  // <= 192 words
#pragma unroll 1
                for (int i = lane; i < nw; i += 32, o32 += 32, s32 += 32) __stcs(o32, *s32);
                if ((nsh & 1) && lane == 0) o[nsh - 1] = bstage[nsh - 1];
            } else {
                for (int t = lane; t < nsh; t += 32) o[t] = bstage[t];
            }
        }
        if (hard) {                                       // packed hard symbols, one byte each
            const unsigned char* hst = wb + L::OFF_ALIAS + FZ_B * 8 + FZ_B * 6;
            uint8_t* o = o_hard + kchain;
            if (full && (reinterpret_cast<uintptr_t>(o) & 3) == 0) reinterpret_cast<unsigned*>(o)[lane] = reinterpret_cast<const unsigned*>(hst)[lane];
            else for (int t = lane; t < m; t += 32) o[t] = hst[t];
        }
    }
    __syncwarp();
}

// literal recursion for one block (lane 0) and its estimates into th[0..m)
template <class L>
static __device__ __noinline__ void fz_chain_slow(const unsigned wofs, const int m)
{
    unsigned char* wb = fz_smem + wofs;
    float*  th   = reinterpret_cast<float*>(wb + L::OFF_TH);
    float*  yh   = reinterpret_cast<float*>(wb + L::OFF_YH);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    float*  estv = reinterpret_cast<float*>(wb + L::OFF_ALIAS + FZ_B * 8) + FZ_B;
    const int lane = fz_lane();
    __syncwarp();
    if (lane == 0) { fz_block_sequential(cx, yh, th, estv, m); cx.cz_valid = 0; }
    __syncwarp();
    for (int i = lane; i < m; i += 32) th[i] = estv[i];
    __syncwarp();
}

// a packet is exhausted: its epilogue (cpp/psk_soft.cpp:592-603) and, unless it was the unit's last, the
// next packet's prologue (:393-426).  Returns true when the unit is done.
template <class L>
static __device__ __noinline__ bool fz_next_packet(const unsigned wofs, const int lane) {
    unsigned char* wb = fz_smem + wofs;
    FzCtx& cx = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    float* yh = reinterpret_cast<float*>(wb + L::OFF_YH);
    const int S = L::S_STATIC ? L::S_STATIC : cx.S_rt;
    if (cx.end_mid) { if (lane == 0) cx.unit_done = 1; __syncwarp(); return true; }    // the packet goes on in the next unit
    if (lane == 0) fz_epilogue(cx, yh);
    __syncwarp();
    const int pkt = cx.pkt + 1;
    if (pkt == cx.pk1) { if (lane == 0) cx.unit_done = 1; __syncwarp(); return true; }
    if (lane == 0) {
        cx.pkt = pkt;
        fz_prologue(cx, yh);
        cx.pk_hi = (pkt + 1 == cx.n_pkts) ? cx.K
                   : (int)first_symbol_at((long long)(pkt + 1) * cx.pkt_len, cx.tail_len, S, cx.A, cx.K);
    }
    __syncwarp();
    return false;
}

// fz_drain: consume buffered symbols: chain blocks of FZ_B (shorter at a packet end) followed by
// the output stage; runs the packet epilogue / next prologue whenever a packet is exhausted.
template <class L, int BK = -1>
static __device__ FZ_HOT void fz_drain(const unsigned wofs)
{
    unsigned char* wb = fz_smem + wofs;
    float*  th   = reinterpret_cast<float*>(wb + L::OFF_TH);
    float*  selx = reinterpret_cast<float*>(wb + L::OFF_SEL);
    float*  yh   = reinterpret_cast<float*>(wb + L::OFF_YH);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    const int lane = fz_lane();

    while (!cx.unit_done) {
        const int kchain = cx.kchain;
        const int rem = cx.pk_hi - kchain;
        if (rem == 0) {
            if (fz_next_packet<L>(wofs, lane)) break;
            continue;
        }
        const int nbuf = cx.nbuf;
        const int want = min(FZ_B, rem);      // blocks are counted from the packet's first symbol: the same packets give the
        if (nbuf < want) break;               // same blocks (and bit-identical results) however the stream is cut into calls

        // ---- one sub-block of m symbols at buffer offset 0 ----------------------------------------
        int m = want;
        const int P = cx.P;
        const int pts = cx.st.fit.pts, cnt = cx.st.fit.count;
        const bool fast = (pts == P) && (P > 1);
        if (fast && cnt + m > 1048576) {
            if (cnt == 1048576) {                                                   // :51-52 at a block edge
                if (lane == 0) fz_resum_w<L>(wofs);
                __syncwarp();
                continue;
            }
            m = 1048576 - cnt;                                                      // stop at the re-sum point
        }
        if (!fast) m = min(m, max(1, P - pts));                                     // fill-up runs sequentially
        bool done = false;
        if (fast) done = fz_chain_fast<L>(wofs, m);
        if (!done) fz_chain_slow<L>(wofs, m);
        if (lane == 0) cx.blocks++;
        fz_back_block<L, BK>(wofs, m);
        const float2 prev_new = fz_sel_get(selx, 2 + m - 1);       // (the stages above leave the samples alone)
        // ---- drop the consumed symbols from the buffer ------------------------------------------------
        const int left = nbuf - m;
        for (int base = 0; base < left; base += 32) {
            const int i = base + lane;
            float tv = 0.f; float2 sv2 = make_float2(0.f, 0.f);
            if (i < left) { tv = th[m + i]; sv2 = fz_sel_get(selx, 2 + m + i); }
            __syncwarp();
            if (i < left) { th[i] = tv; fz_sel_put(selx, 2 + i, sv2); }
            __syncwarp();
        }
        if (lane == 0) { fz_sel_put(selx, 1, prev_new); cx.nbuf = left; cx.kchain = kchain + m; }
        __syncwarp();
        if (left == 0 && m < rem) break;      // buffer empty, packet not exhausted: the next block has to be collected first
    }
}

// ---------------------------------------------------------------------------------------------
// atan2f for the M-th power angle (cpp/psk_soft.cpp:474).  glibc's atan2f (the reference) and
// CUDA's are both "a few ulp" functions; this one is too (|error| <= 3e-7 rad: 2-ulp quotient,
// degree-8 minimax polynomial in r^2, pi split in two floats), in ~25 instructions instead of ~60.
// The angle only feeds the unwrap count (an integer; every count is verified) and the fitted
// phase (tolerance 1e-4).  Zero, non-finite and extreme-magnitude inputs take the library call.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fz_atan2(float y, float x, bool& bad) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    bad = !(mx > 1.0e-18f && mx < 1.0e18f);
    const float r = __fdividef(mn, mx);
    const float t = r * r;
    float q = 0.0029035548213869333f;
    q = fmaf(q, t, -0.01628301665186882f);
    q = fmaf(q, t, 0.04303938150405884f);
    q = fmaf(q, t, -0.0753367692232132f);
    q = fmaf(q, t, 0.1065467819571495f);
    q = fmaf(q, t, -0.14207133650779724f);
    q = fmaf(q, t, 0.19993054866790771f);
    q = fmaf(q, t, -0.3333309292793274f);
    float a = fmaf(r * t, q, r);
    if (ay > ax) a = (1.57079637050628662109375f - a) + -4.37113900018624283e-8f;
    if (x < 0.0f) a = (3.1415927410125732421875f - a) + -8.74227800037248566e-8f;
    return copysignf(a, y);
}

// atan2f(Im s^M, Re s^M) with s^M by repeated squaring, unfused (libstdc++ __complex_pow_unsigned,
// cpp/psk_soft.cpp:474); straight-line for M = 2, 4, 8.  `bad` asks for the literal path
// (fz_theta_fixup): other M, zero / non-finite / extreme powers (a (NaN, NaN) product can also
// mean the reference's __mulsc3 recovery ran).
__device__ __forceinline__ float2 fz_csq(float2 x) {
    return make_float2(fsubr(fmulr(x.x, x.x), fmulr(x.y, x.y)), faddr(fmulr(x.x, x.y), fmulr(x.y, x.x)));
}
__device__ __forceinline__ float fz_theta(float2 s, int M, bool& bad) {
#ifdef PSKD_FZ_THETA_BRANCHY
    float2 y = s;
    if (M == 8) y = fz_csq(fz_csq(fz_csq(s)));
    else if (M == 4) y = fz_csq(fz_csq(s));
    else if (M == 2) y = fz_csq(s);
    return fz_atan2(y.y, y.x, bad);
#else
    // M = 2, 4, 8: one, two or three squarings, selected without branches (M is warp-uniform)
    const float2 y1 = fz_csq(s);
    const float2 y2 = fz_csq(y1);
    const float2 y3 = fz_csq(y2);
    float2 y = (M >= 8) ? y3 : y2;
    y = (M >= 4) ? y : y1;
    return fz_atan2(y.y, y.x, bad);
#endif
}
// literal path for the lanes that asked for it: th[i] = atan2f(pow(sel[i], M)) with the library call
static __device__ __noinline__ void fz_theta_fixup(float* th, float2 sel, int i, unsigned M) {
    const float2 z = cpow_unsigned(sel, M);
    th[i] = atan2f(z.y, z.x);
}

template <class L> static __device__ __noinline__ void fz_theta_fixup_w(const unsigned wofs, int i, unsigned M) {
    unsigned char* wb = fz_smem + wofs;
    fz_theta_fixup(reinterpret_cast<float*>(wb + L::OFF_TH), fz_sel_get(reinterpret_cast<const float*>(wb + L::OFF_SEL), 2 + i), i, M);
}

__device__ __forceinline__ void fz_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------
// staged blocks: 32 rows of raw samples in shared memory, sample n (= row*S + phase) at
// FzCfg<S>::phys(n).  fz_issue copies a block that lies wholly inside this call's input with
// cp.async (16-byte pieces when the block's global address allows it, else 8-byte pieces);
// fz_fill_slow assembles a block that touches the carried tail or the end of the stream.
// ---------------------------------------------------------------------------------------------
template <int S, bool A16>
__device__ __forceinline__ void fz_issue(float2* st, const float2* src, int lane) {
    using C = FzCfg<S>;
    if (A16) {
        const float4* g4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(st);
#pragma unroll
        for (int q = 0; q < C::NQ16; q++) {
            const int f = lane + 32 * q;
            if ((S * 16) % 32 == 0 || f < S * 16) {
                const int fp = C::phys(2 * f) >> 1;                                    // phys() on 16-byte pieces (keeps pairs together)
#ifdef PSKD_FZ_ISSUE_SRC
                // the permutation is applied on the SOURCE side (phys() is an involution): the lanes of one LDGSTS write
                // consecutive 16-byte pieces, which is what its shared-memory writes coalesce on; the global side still
                // touches the same sectors.  Experiment, off: 185 -> 180 shared-memory wavefronts per chunk, 12.81 vs 12.77 ms
                fz_cp_async16(d4 + f, g4 + fp);
#else
                fz_cp_async16(d4 + fp, g4 + f);
#endif
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < S; q++) {
            const int n = lane + 32 * q;
            fz_cp_async8(st + C::phys(n), src + n);
        }
    }
}
template <int S, int PC>
static __device__ __noinline__ void fz_fill_slow(const unsigned wofs, const int trail, long long s0, int lane) {
    using C = FzCfg<S>;
    using L = FzL<S, PC>;
    unsigned char* wb = fz_smem + wofs;
    float2* st = reinterpret_cast<float2*>(wb + (trail ? L::OFF_T : L::OFF_L));
    const FzCtx& cx = *reinterpret_cast<const FzCtx*>(wb + L::OFF_CTX);
    const long long V = cx.V, tail_len = cx.tail_len;
    const float2* tailp = cx.tail;
    const float2* in_mt = cx.in_mt;
#pragma unroll 2
    for (int q = 0; q < S; q++) {
        const int n = lane + 32 * q;
        const long long v = s0 + n;
        float2 x = make_float2(0.f, 0.f);
        if (v < V) x = (v < tail_len) ? tailp[v] : __ldg(in_mt + v);
        st[C::phys(n)] = x;
    }
}

// ---------------------------------------------------------------------------------------------
// fz_unit_begin: set-up of one unit: geometry, wait for the predecessor unit, carried state into
// shared memory, first packet prologue, the carried window sums.  Returns the number of chunks
// (< 0: nothing to do for this ticket).
// ---------------------------------------------------------------------------------------------
template <int S, int PC>
static __device__ __noinline__ int fz_unit_begin(const FusedParams& prm, const unsigned wofs, const int u)
{
    using C = FzCfg<S>;
    using L = FzL<S, PC>;
    constexpr int G = C::G, ES = C::ES;
    constexpr int CHS = C::CHS;
    unsigned char* wb = fz_smem + wofs;
    float2* Lst  = reinterpret_cast<float2*>(wb + L::OFF_L);
    float2* Tst  = reinterpret_cast<float2*>(wb + L::OFF_T);
    float*  selx = reinterpret_cast<float*>(wb + L::OFF_SEL);
    float*  yh   = reinterpret_cast<float*>(wb + L::OFF_YH);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    double* cwp  = reinterpret_cast<double*>(wb + L::OFF_CW);
    double* ebuf = reinterpret_cast<double*>(wb + L::OFF_ALIAS);
    const int lane = threadIdx.x & 31;
    const bool wact = lane < G * S;
    const int wg = wact ? lane / S : 0, wp = wact ? lane - (lane / S) * S : 0;

    const int ug = u / prm.n_list;
    const int ch = __ldg(prm.list + (u - ug * prm.n_list));
    const ChanDesc* dgp = prm.desc + ch;
    const int n_pkts = dgp->n_pkts;
    const int parts = prm.parts_per_pkt;
    const int part = (parts > 1) ? ug % parts : 0;
    const int pk0 = (parts > 1) ? ug / parts : ug * prm.pkts_per_unit;
    if (pk0 >= n_pkts) return -1;
    const int pk1 = (parts > 1) ? pk0 + 1 : min(pk0 + prm.pkts_per_unit, n_pkts);
    const int A = dgp->A, P = dgp->P;
    const long long tail_len = dgp->tail_len, pkt_len = dgp->pkt_len;
    const int K = (int)dgp->K;
    const long long V = tail_len + dgp->n_in;
    const int lag = A - 1;
    int kA = (int)first_symbol_at((long long)pk0 * pkt_len, tail_len, S, A, K);
    int kB = (pk1 == n_pkts) ? K : (int)first_symbol_at((long long)pk1 * pkt_len, tail_len, S, A, K);
    if (parts > 1) {                   // this unit: part `part` of the packet's symbols (cuts at multiples of 32 symbols)
        const int per = (((kB - kA + parts - 1) / parts) + 31) & ~31;
        const int a = min(kB, kA + part * per);
        kB = (part == parts - 1) ? kB : min(kB, a + per);
        kA = a;
    }
    const bool start_mid = part > 0;                  // the packet's prologue ran in part 0
    const bool end_mid = part < parts - 1;            // the packet's epilogue runs in its last part
    const int nchunks = (kB - kA + FZ_CH - 1) / FZ_CH;
    const float2* in_mt = dgp->in - tail_len;
    const float2* tailp = dgp->tail;
    // chunks [c_lo, c_hi): trail block (rows kA+32c ..) and lead block (rows kA+32c+lag ..) both lie
    // wholly inside `in` -> cp.async copies; 16-byte pieces if both blocks start 16-byte aligned
    int c_lo = 0, c_hi = 0, a16 = 0;
    {
        const long long sT0 = (long long)kA * S, sL0 = (long long)(kA + lag) * S;
        long long lo = (tail_len - sT0 + CHS - 1) / CHS;
        if (lo < 0) lo = 0;
        long long hi = (V - sL0) / CHS;                                   // lead blocks fully below V
        if (hi > nchunks) hi = nchunks;
        if (lo < hi) { c_lo = (int)lo; c_hi = (int)hi; }
        a16 = (((reinterpret_cast<uintptr_t>(in_mt + sT0) | reinterpret_cast<uintptr_t>(in_mt + sL0)) & 15) == 0) ? 1 : 0;
        // start the stream: the first two lead blocks towards L2
        if (lane == 0 && c_lo < c_hi && a16)
            fz_prefetch_l2(in_mt + sL0 + (long long)c_lo * CHS, (unsigned)min(FZ_PF, c_hi - c_lo) * CHS * 8);
    }
    const bool pre0 = nchunks > 0 && c_lo == 0 && c_hi > 0;
    if (pre0) {                                                           // chunk 0's blocks: input only, no dependence on the predecessor
        if (a16) {
            fz_issue<S, true>(Tst, in_mt + (long long)kA * S, lane);
            fz_issue<S, true>(Lst, in_mt + (long long)(kA + lag) * S, lane);
        } else {
            fz_issue<S, false>(Tst, in_mt + (long long)kA * S, lane);
            fz_issue<S, false>(Lst, in_mt + (long long)(kA + lag) * S, lane);
        }
    }
    // ---- wait for the previous unit of this channel, then load its carried state ------------------
    if (ug > 0) {
        if (lane == 0) { while (ld_acquire(prm.done + ch) < ug) __nanosleep(200); }
        __syncwarp();
    }
    const float* gring = prm.ring_base + dgp->ring_off;
    if (lane == 0) {
        const int4* src = reinterpret_cast<const int4*>(prm.state + ch);
        int4* dst = reinterpret_cast<int4*>(&cx.st);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(ChanState) / 16); i++) dst[i] = __ldcg(src + i);
        const long long sym_off = dgp->sym_off;
        cx.flags = (ug == 0) ? dgp->flags : (dgp->flags & ~(CH_RESET_NUMSYMS | CH_RESET_PHASEAVG));
        cx.passes = 0; cx.seq_blocks = 0; cx.blocks = 0;
        cx.in_mt = in_mt; cx.tail = tailp; cx.desc = dgp;
        cx.o_soft = prm.out_soft ? prm.out_soft + sym_off : nullptr;
        cx.o_phase = prm.out_phase ? prm.out_phase + sym_off : nullptr;
        cx.o_bits = prm.out_bits ? prm.out_bits + dgp->bits_off : nullptr;
        cx.o_sidx = prm.out_sidx ? prm.out_sidx + sym_off : nullptr;
        cx.o_hard = prm.out_hard ? prm.out_hard + sym_off : nullptr;
        cx.sri_xdelta = prm.sri_xdelta;
        cx.tail_len = tail_len; cx.pkt_len = pkt_len;
        cx.n_pkts = n_pkts; cx.pk1 = pk1; cx.K = K; cx.A = A; cx.M = dgp->M; cx.P = P; cx.bpb = dgp->bpb; cx.diff = dgp->D;
        cx.pkt = pk0; cx.kchain = kA; cx.nbuf = 0; cx.cz_valid = 0; cx.unit_done = 0;
        cx.fP1 = (float)(P - 1);
        cx.pk_hi = (parts > 1) ? kB
                   : (pk0 + 1 == n_pkts) ? K : (int)first_symbol_at((long long)(pk0 + 1) * pkt_len, tail_len, S, A, K);
        cx.end_mid = end_mid ? 1 : 0;
        cx.kA = kA; cx.kB = kB; cx.lag = lag;
        cx.c_lo = c_lo; cx.c_hi = c_hi; cx.nchunks = nchunks; cx.a16 = a16;
        cx.ch = ch; cx.ug = ug; cx.V = V;
        cx.c = 0; cx.inflight = pre0 ? 1 : 0;
    }
    for (int j = lane; j < P; j += 32) yh[j] = __ldcg(gring + j);
    __syncwarp();
    if (lane == 0) {
        cx.wraps0 = cx.st.wraps; fz_sel_put(selx, 1, cx.st.last);
        if (!start_mid) fz_prologue(cx, yh);
        else if (cx.st.fit.pts == cx.st.fit.n && cx.st.fit.pts > 1) cx.fc = fit_const(cx.st.fit);
    }
    __syncwarp();

    // ---- carried window sums: rows [kA, kA+lag) per phase, exact double sums ------------------------
    if (nchunks > 0) {
        const long long s0 = (long long)kA * S;
        double acc = 0.0;
        if (wact) {
            for (int i = wg; i < lag; i += G) {
                const long long v = s0 + (long long)i * S + wp;
                float2 x = make_float2(0.f, 0.f);
                if (v < V) x = (v < tail_len) ? tailp[v] : __ldg(in_mt + v);
                acc = daddr(acc, (double)energy_f32(x.x, x.y));
            }
            ebuf[wg * ES + wp] = acc;
        }
        __syncwarp();
        if (lane < S) {
            double Cw = 0.0;
#pragma unroll
            for (int g2 = 0; g2 < G; g2++) Cw = daddr(Cw, ebuf[g2 * ES + lane]);
            cwp[lane] = Cw;                    // sum over the window of output kA WITHOUT its newest row
        }
        __syncwarp();
    }
    return nchunks;
}

// fz_unit_end: hand the channel's state to the next unit / the next call
template <int S, int PC>
static __device__ __noinline__ void fz_unit_end(const FusedParams& prm, const unsigned wofs)
{
    using L = FzL<S, PC>;
    unsigned char* wb = fz_smem + wofs;
    float*  selx = reinterpret_cast<float*>(wb + L::OFF_SEL);
    float*  yh   = reinterpret_cast<float*>(wb + L::OFF_YH);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    const int lane = threadIdx.x & 31;
    const int P = cx.P, ch = cx.ch;
    float* gring = prm.ring_base + cx.desc->ring_off;
    if (cx.st.fit.head != 0 && cx.st.fit.pts == P)
        fz_normalize_ring(yh, reinterpret_cast<float*>(wb + L::OFF_ALIAS), cx.st.fit, P, lane);
    for (int j = lane; j < P; j += 32) __stcg(gring + j, yh[j]);
    if (lane == 0) {
        if (cx.diff && cx.kB > cx.kA) cx.st.last = fz_sel_get(selx, 1);                   // :489
        const int4* src = reinterpret_cast<const int4*>(&cx.st);
        int4* dst = reinterpret_cast<int4*>(prm.state + ch);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(ChanState) / 16); i++) __stcg(dst + i, src[i]);
        if (cx.st.wraps != cx.wraps0) atomicAdd(&prm.counters->wraps, cx.st.wraps - cx.wraps0);
        if (cx.blocks) atomicAdd(&prm.counters->spec_chunks, (unsigned long long)cx.blocks);
        if (cx.passes) atomicAdd(&prm.counters->spec_misses, (unsigned long long)cx.passes);
        if (cx.seq_blocks) atomicAdd(&prm.counters->seq_channels, (unsigned long long)cx.seq_blocks);
    }
    __syncwarp();
    __threadfence();
    if (lane == 0) st_release(prm.done + ch, cx.ug + 1);
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// fz_chunk: the front stage (its own register allocation; state in FzCtx between calls).  Runs
// chunks until the block buffer holds what the chain stage wants (or the unit's chunks are
// exhausted).  One chunk:
//   wait for the two staged blocks; window sums (lane = phase x row group) from the energies of
//   the lead and trail samples; scan over the row groups; transpose; issue the next lead block;
//   first maximum per row (lane = row); pick the selected sample out of the trail block; issue the
//   next trail block; M-th power angle; append (theta, sample) to the block buffer.
// ---------------------------------------------------------------------------------------------
template <int S, int PC, bool A16>
static __device__ FZ_HOT void fz_chunk(const unsigned wofs)
{
    using C = FzCfg<S>;
    using L = FzL<S, PC>;
    constexpr int G = C::G, R = C::R, ES = C::ES;
    constexpr int CHS = C::CHS;
    unsigned char* wb = fz_smem + wofs;
    float2* Lst  = reinterpret_cast<float2*>(wb + L::OFF_L);
    float2* Tst  = reinterpret_cast<float2*>(wb + L::OFF_T);
    float*  th   = reinterpret_cast<float*>(wb + L::OFF_TH);
    float*  selx = reinterpret_cast<float*>(wb + L::OFF_SEL);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    double* cwp  = reinterpret_cast<double*>(wb + L::OFF_CW);
    double* ebuf = reinterpret_cast<double*>(wb + L::OFF_ALIAS);       // [32][ES] window sums
    const int lane = fz_lane();
    const bool wact = lane < G * S;
    const int wg = wact ? lane / S : 0, wp = wact ? lane - (lane / S) * S : 0;

    int c = cx.c, nbuf = cx.nbuf;
    bool inflight = cx.inflight != 0;
    const int nchunks = cx.nchunks, c_lo = cx.c_lo, c_hi = cx.c_hi, lag = cx.lag;
    const int kA = cx.kA, kB = cx.kB, M = cx.M;
    const int want = min(FZ_BLOCKS * FZ_B, cx.pk_hi - cx.kchain);
    const bool m_ok = (M == 2 || M == 4 || M == 8);
    const float2* in_mt = cx.in_mt;
    int16_t* o_sidx = cx.o_sidx;
    double Cw = cwp[wp];
    // lane = (phase, group): where this lane's rows sit in a staged block (even / odd rows, see phys())
    const int od = (S == 8) ? (wg & 1) : 0;
    const int ofs_e = (R * wg + od) * S + wp, ofs_o = (R * wg - od) * S + wp;
    // lane = row: where this lane's row sits in the trail block
    const int rowp = (S == 8) ? (lane ^ ((lane >> 3) & 1)) * S : lane * S;

    do {
        const int krow = kA + FZ_CH * c;
        if (!inflight) {
            if (c >= c_lo && c < c_hi) {
                fz_issue<S, A16>(Tst, in_mt + (long long)krow * S, lane);
                fz_issue<S, A16>(Lst, in_mt + (long long)(krow + lag) * S, lane);
            } else {
                fz_fill_slow<S, PC>(wofs, 1, (long long)krow * S, lane);
                fz_fill_slow<S, PC>(wofs, 0, (long long)(krow + lag) * S, lane);
            }
        }
        fz_cp_async_wait_all();
        __syncwarp();

        // ---- timing, part 1: exact sliding window sums (:451, :576), lane = (phase wp, row group wg);
        // lanes beyond G*S (S = 9, 10) run along on (0, 0) and store nothing
        double Eloc[R];
        double x = 0.0;
        {
            const float2* le = Lst + ofs_e; const float2* lo = Lst + ofs_o;
            const float2* te = Tst + ofs_e; const float2* to = Tst + ofs_o;
#pragma unroll
            for (int i = 0; i < R; i++) {
                if (G * R == 32 || R * wg + i < 32) {
                    const float2 a = (i & 1) ? lo[i * S] : le[i * S];
                    const float2 b = (i & 1) ? to[i * S] : te[i * S];
                    x = daddr(x, (double)fz_energy(a));                       // :448-451
                    Eloc[i] = x;
                    x = dsubr(x, (double)fz_energy(b));                       // :576
                } else Eloc[i] = 0.0;
            }
        }
        {
            // scan of the group totals over the row groups (all partial sums exact)
            double incl = x;
#pragma unroll
            for (int d = 1; d < G; d <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, incl, d * S, 32);
                if (wg >= d) incl = daddr(incl, t);
            }
            const double ex = __shfl_up_sync(0xffffffffu, incl, S, 32);
            const double tot = __shfl_sync(0xffffffffu, incl, (G - 1) * S + wp, 32);
            const double off = (wg >= 1) ? daddr(Cw, ex) : Cw;
            Cw = daddr(Cw, tot);
            if (wact) {
                double* eo = ebuf + (R * wg) * ES + wp;
#pragma unroll
                for (int i = 0; i < R; i++)
                    if (G * R == 32 || R * wg + i < 32) eo[i * ES] = daddr(off, Eloc[i]);
            }
        }
        __syncwarp();
        const bool nfast = (c + 1 >= c_lo) && (c + 1 < c_hi);
        if (nfast) fz_issue<S, A16>(Lst, in_mt + (long long)(krow + FZ_CH + lag) * S, lane);

        // ---- timing, part 2: lane = row: first maximum (:462), the selected sample (:465) ------------
        const int nrows = min(FZ_CH, kB - krow);
        float2 gx;
        {
            const double* er = ebuf + lane * ES;
            double e[S];
#pragma unroll
            for (int q = 0; q < S; q++) e[q] = er[q];
            // tournament over contiguous ranges keeps std::max_element's FIRST maximum
            int ix[S];
#pragma unroll
            for (int q = 0; q < S; q++) ix[q] = q;
#pragma unroll
            for (int lv = 0; lv < 5; lv++) {
                const int w = 1 << lv;
#pragma unroll
                for (int q = 0; q < S; q++) {
                    if (w < S && (q % (2 * w)) == 0 && q + w < S) {
                        if (e[q] < e[q + w]) { e[q] = e[q + w]; ix[q] = ix[q + w]; }
                    }
                }
            }
            const int idx = ix[0];
            if (o_sidx && lane < nrows) __stcs(o_sidx + krow + lane, (int16_t)idx);        // :466
            gx = Tst[rowp + idx];
        }
        __syncwarp();
        if (nfast) {
            fz_issue<S, A16>(Tst, in_mt + (long long)(krow + FZ_CH) * S, lane);
            // and the lead block after the next towards L2, one 128-byte line per lane
            if (c + FZ_PF < c_hi) {
                const float2* nx = in_mt + (long long)(krow + FZ_PF * FZ_CH + lag) * S;
                if (A16) { if (lane == 0) fz_prefetch_l2(nx, (unsigned)CHS * 8u); }
                else if (lane * 16 < CHS) fz_prefetch_line(nx + lane * 16);
            }
        }
        inflight = nfast;

        // ---- M-th power angle (:474), append to the block buffer ---------------------------------------
        // (forming the angles of a whole block in a packed two-symbols-per-lane stage saves 22 instructions per chunk
        // and costs time: 12.95 vs 12.75 ms -- here the angle's dependent chain overlaps the rest of the loop body)
        bool bad = false;
        const float thv = fz_theta(gx, M, bad);
        bad = (bad || !m_ok) && lane < nrows;
        if (lane < nrows) {
            th[nbuf + lane] = thv;
            fz_sel_put(selx, 2 + nbuf + lane, gx);
        }
        if (__any_sync(0xffffffffu, bad)) {            // rare: literal angle for odd inputs / other M
            __syncwarp();
            if (bad) fz_theta_fixup_w<L>(wofs, nbuf + lane, (unsigned)M);
        }
        nbuf += nrows;
        c++;
    } while (c < nchunks && nbuf < want);
    __syncwarp();
    if (lane == 0) { cx.c = c; cx.nbuf = nbuf; cx.inflight = inflight ? 1 : 0; }
    if (lane < S) cwp[lane] = Cw;
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// CT = resident CTAs per SM the register allocation aims at: 5 (96 registers) is the faster code per warp, 6 (80 registers,
// 24 warps per SM) hides more latency and wins once a launch has enough channels to keep all of those warps busy
// (launch_fused_t decides).
template <int S, int PC, int CT, int BK = -1>
__global__ void __launch_bounds__(FZ_WARPS * 32, CT)
k_fused(const FusedParams prm)
{
    using L = FzL<S, PC>;
    const int lane = threadIdx.x & 31;
    const unsigned wofs = (threadIdx.x >> 5) * (unsigned)L::BYTES;
    FzCtx& cx = *reinterpret_cast<FzCtx*>(fz_smem + wofs + L::OFF_CTX);
#ifdef PSKD_FZ_FILL_SMEM            // debugging aid: start from a known pattern instead of whatever the last kernel left
    for (int i = lane; i < L::BYTES / 4; i += 32) reinterpret_cast<unsigned*>(fz_smem + wofs)[i] = PSKD_FZ_FILL_SMEM;
    __syncwarp();
#endif

    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(prm.ticket, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= prm.n_units) break;
        const int nchunks = fz_unit_begin<S, PC>(prm, wofs, u);
        if (nchunks < 0) continue;
        fz_drain<L, BK>(wofs);         // packets without symbols before the first chunk (and units without any symbol)
        if (cx.a16) {
            while (cx.c < nchunks) { fz_chunk<S, PC, true>(wofs); fz_drain<L, BK>(wofs); }
        } else {
            while (cx.c < nchunks) { fz_chunk<S, PC, false>(wofs); fz_drain<L, BK>(wofs); }
        }
        fz_unit_end<S, PC>(prm, wofs);
    }
#ifdef PSKD_FZ_PAD_CODE
    // layout experiment (tools/probe): never-executed instructions between the kernel's hot code and its out-of-line stage functions
    if (prm.n_units == -12345) {
        unsigned acc = 0, t;
#define PSKD_PAD1 asm volatile("mov.u32 %0, %%clock;" : "=r"(t)); acc ^= t;
#define PSKD_PAD8 PSKD_PAD1 PSKD_PAD1 PSKD_PAD1 PSKD_PAD1 PSKD_PAD1 PSKD_PAD1 PSKD_PAD1 PSKD_PAD1
#define PSKD_PAD64 PSKD_PAD8 PSKD_PAD8 PSKD_PAD8 PSKD_PAD8 PSKD_PAD8 PSKD_PAD8 PSKD_PAD8 PSKD_PAD8
#pragma unroll
        for (int i = 0; i < PSKD_FZ_PAD_CODE; i++) { PSKD_PAD8 }
        prm.ticket[1] = (int)acc;
    }
#endif
}

// =============================================================================================
// Staged path built from the fused kernel's stages (few channels / long channels: the regime where
// one warp per channel leaves the GPU empty).  The two data-dependent halves of the path are split at
// the only sequential quantity, the integer level of the unwrapped phase:
//   k_fzs_front<S>  timing + M-th power angle of EVERY symbol, fully time-parallel: a unit is a
//                   segment of one channel (the carried window sums of a segment are rebuilt exactly
//                   from the numAvg-1 rows in front of it, as fz_unit_begin does).  Writes sampleIndex
//                   and 12 B/symbol of scratch (selected sample, angle).  cpp/psk_soft.cpp:442-474.
//   k_fzs_cb<PC>    unwrap / LinearFit chain + derotate / slice of one unit = consecutive packets of
//                   one channel, from the scratch.  With the time-parallel plan (TpCtl, pskd_internal.h)
//                   every packet is its own unit, started from a history ring synthesised from the
//                   resolved integer levels and proven afterwards by k_tp_check.  :476-603.
// =============================================================================================
#ifndef PSKD_FZS_FRONT_MIN_CTAS
#define PSKD_FZS_FRONT_MIN_CTAS 6
#endif
#ifndef PSKD_FZS_CB_MIN_CTAS
#define PSKD_FZS_CB_MIN_CTAS 6         // 24 warps per SM at 80 registers.  Measured, 512 x 1M coherent 8-PSK: 7 CTAs / 72 registers 0.948 ms,
                                       // 6 / 80: 0.988, 5 / 96: 1.008 -- but 7 CTAs spill: differential 256 x 4M 1.87 vs 1.78 ms, one
                                       // channel x 64M (latency-bound) 0.27 vs 0.22 ms
#endif

template <int S> struct FzsFL {          // per-warp shared memory of the front kernel
    using C = FzCfg<S>;
    static constexpr int BLK = C::CHS * 8;
    static constexpr int OFF_L = 0;                    // float2 lead[32*S]
    static constexpr int OFF_T = BLK;                  // float2 trail[32*S]
    static constexpr int OFF_E = 2 * BLK;              // double e[32][ES] window sums
    static constexpr int OFF_CW = OFF_E + 32 * C::ES * 8;   // double cw[16]
    static constexpr int BYTES = fz_align128(OFF_CW + 16 * 8);
    static_assert(BYTES % 128 == 0 && OFF_T % 128 == 0, "staged blocks must start on a 128-byte line");
};

struct FzsFrontParams {
    const ChanDesc* desc; int n_channels;
    int seg_syms;                      // symbols per unit (multiple of 32)
    int n_units;                       // n_channels * ceil(Kmax / seg_syms), segment-major
    int* ticket;
    float2* sel; float* theta; int16_t* out_sidx;
};

template <int S>
static __device__ __noinline__ void fzs_fill_slow(float2* st, long long s0, long long V, long long tail_len,
                                                  const float2* tailp, const float2* in_mt, int lane) {
    using C = FzCfg<S>;
#pragma unroll 2
    for (int q = 0; q < S; q++) {
        const int n = lane + 32 * q;
        const long long v = s0 + n;
        float2 x = make_float2(0.f, 0.f);
        if (v < V) x = (v < tail_len) ? tailp[v] : __ldg(in_mt + v);
        st[C::phys(n)] = x;
    }
}

// all chunks of one front unit: rows [kA, kB) of one channel (same arithmetic as fz_chunk; the angle and
// the selected sample go to the scratch arrays instead of the block buffer)
template <int S, bool A16>
static __device__ __forceinline__ void fzs_front_chunks(unsigned char* wb, const int lane, const int kA, const int kB,
                                                       const int lag, const int c_lo, const int c_hi, const int nchunks,
                                                       bool inflight, const int M, const float2* in_mt, const float2* tailp,
                                                       const long long tail_len, const long long V,
                                                       int16_t* o_sidx, float* o_th, float2* o_sel)
{
    using C = FzCfg<S>;
    using L = FzsFL<S>;
    constexpr int G = C::G, R = C::R, ES = C::ES;
    constexpr int CHS = C::CHS;
    float2* Lst  = reinterpret_cast<float2*>(wb + L::OFF_L);
    float2* Tst  = reinterpret_cast<float2*>(wb + L::OFF_T);
    double* cwp  = reinterpret_cast<double*>(wb + L::OFF_CW);
    double* ebuf = reinterpret_cast<double*>(wb + L::OFF_E);
    const bool wact = lane < G * S;
    const int wg = wact ? lane / S : 0, wp = wact ? lane - (lane / S) * S : 0;
    const bool m_ok = (M == 2 || M == 4 || M == 8);
    double Cw = cwp[wp];
    const int od = (S == 8) ? (wg & 1) : 0;
    const int ofs_e = (R * wg + od) * S + wp, ofs_o = (R * wg - od) * S + wp;
    const int rowp = (S == 8) ? (lane ^ ((lane >> 3) & 1)) * S : lane * S;

#pragma unroll 1
    for (int c = 0; c < nchunks; c++) {
        const int krow = kA + FZ_CH * c;
        if (!inflight) {
            if (c >= c_lo && c < c_hi) {
                fz_issue<S, A16>(Tst, in_mt + (long long)krow * S, lane);
                fz_issue<S, A16>(Lst, in_mt + (long long)(krow + lag) * S, lane);
            } else {
                fzs_fill_slow<S>(Tst, (long long)krow * S, V, tail_len, tailp, in_mt, lane);
                fzs_fill_slow<S>(Lst, (long long)(krow + lag) * S, V, tail_len, tailp, in_mt, lane);
            }
        }
        fz_cp_async_wait_all();
        __syncwarp();
        // ---- timing, part 1: exact sliding window sums (:451, :576), lane = (phase wp, row group wg)
        double Eloc[R];
        double x = 0.0;
        {
            const float2* le = Lst + ofs_e; const float2* lo = Lst + ofs_o;
            const float2* te = Tst + ofs_e; const float2* to = Tst + ofs_o;
#pragma unroll
            for (int i = 0; i < R; i++) {
                if (G * R == 32 || R * wg + i < 32) {
                    const float2 a = (i & 1) ? lo[i * S] : le[i * S];
                    const float2 b = (i & 1) ? to[i * S] : te[i * S];
                    x = daddr(x, (double)fz_energy(a));                       // :448-451
                    Eloc[i] = x;
                    x = dsubr(x, (double)fz_energy(b));                       // :576
                } else Eloc[i] = 0.0;
            }
        }
        {
            double incl = x;
#pragma unroll
            for (int d = 1; d < G; d <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, incl, d * S, 32);
                if (wg >= d) incl = daddr(incl, t);
            }
            const double ex = __shfl_up_sync(0xffffffffu, incl, S, 32);
            const double tot = __shfl_sync(0xffffffffu, incl, (G - 1) * S + wp, 32);
            const double off = (wg >= 1) ? daddr(Cw, ex) : Cw;
            Cw = daddr(Cw, tot);
            if (wact) {
                double* eo = ebuf + (R * wg) * ES + wp;
#pragma unroll
                for (int i = 0; i < R; i++)
                    if (G * R == 32 || R * wg + i < 32) eo[i * ES] = daddr(off, Eloc[i]);
            }
        }
        __syncwarp();
        const bool nfast = (c + 1 >= c_lo) && (c + 1 < c_hi);
        if (nfast) fz_issue<S, A16>(Lst, in_mt + (long long)(krow + FZ_CH + lag) * S, lane);
        // ---- timing, part 2: lane = row: first maximum (:462), the selected sample (:465)
        const int nrows = min(FZ_CH, kB - krow);
        float2 gx;
        {
            const double* er = ebuf + lane * ES;
            double e[S];
#pragma unroll
            for (int q = 0; q < S; q++) e[q] = er[q];
            int ix[S];
#pragma unroll
            for (int q = 0; q < S; q++) ix[q] = q;
#pragma unroll
            for (int lv = 0; lv < 5; lv++) {
                const int w = 1 << lv;
#pragma unroll
                for (int q = 0; q < S; q++) {
                    if (w < S && (q % (2 * w)) == 0 && q + w < S) {
                        if (e[q] < e[q + w]) { e[q] = e[q + w]; ix[q] = ix[q + w]; }
                    }
                }
            }
            const int idx = ix[0];
            if (lane < nrows) __stcs(o_sidx + krow + lane, (int16_t)idx);                  // :466
            gx = Tst[rowp + idx];
        }
        __syncwarp();
        if (nfast) {
            fz_issue<S, A16>(Tst, in_mt + (long long)(krow + FZ_CH) * S, lane);
            if (c + FZ_PF < c_hi) {
                const float2* nx = in_mt + (long long)(krow + FZ_PF * FZ_CH + lag) * S;
                if (A16) { if (lane == 0) fz_prefetch_l2(nx, (unsigned)CHS * 8u); }
                else if (lane * 16 < CHS) fz_prefetch_line(nx + lane * 16);
            }
        }
        inflight = nfast;
        // ---- M-th power angle (:474) -> scratch
        bool bad = false;
        const float thv = fz_theta(gx, M, bad);
        bad = (bad || !m_ok) && lane < nrows;
        if (lane < nrows) {
            o_th[krow + lane] = thv;
            o_sel[krow + lane] = gx;
        }
        if (__any_sync(0xffffffffu, bad)) {            // rare: literal angle for odd inputs / other M
            __syncwarp();
            if (bad) fz_theta_fixup(o_th, gx, krow + lane, (unsigned)M);
        }
    }
    __syncwarp();
}

template <int S>
__global__ void __launch_bounds__(FZ_WARPS * 32, PSKD_FZS_FRONT_MIN_CTAS)
k_fzs_front(const FzsFrontParams prm)
{
    using C = FzCfg<S>;
    using L = FzsFL<S>;
    constexpr int G = C::G, ES = C::ES, CHS = C::CHS;
    const int lane = fz_lane();
    unsigned char* wb = fz_smem + (threadIdx.x >> 5) * (unsigned)L::BYTES;
    float2* Lst  = reinterpret_cast<float2*>(wb + L::OFF_L);
    float2* Tst  = reinterpret_cast<float2*>(wb + L::OFF_T);
    double* cwp  = reinterpret_cast<double*>(wb + L::OFF_CW);
    double* ebuf = reinterpret_cast<double*>(wb + L::OFF_E);
    const bool wact = lane < G * S;
    const int wg = wact ? lane / S : 0, wp = wact ? lane - (lane / S) * S : 0;

    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(prm.ticket, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= prm.n_units) break;
        const int seg = u / prm.n_channels;
        const int ch = u - seg * prm.n_channels;
        const ChanDesc* dgp = prm.desc + ch;
        if (!(dgp->flags & CH_FZS) || dgp->S != S) continue;
        const int K = (int)dgp->K;
        const int kA = seg * prm.seg_syms;
        if (kA >= K) continue;
        const int kB = min(K, kA + prm.seg_syms);
        const int A = dgp->A, lag = A - 1;
        const long long tail_len = dgp->tail_len;
        const long long V = tail_len + dgp->n_in;
        const int nchunks = (kB - kA + FZ_CH - 1) / FZ_CH;
        const float2* in_mt = dgp->in - tail_len;
        const float2* tailp = dgp->tail;
        int c_lo = 0, c_hi = 0, a16 = 0;
        {
            const long long sT0 = (long long)kA * S, sL0 = (long long)(kA + lag) * S;
            long long lo = (tail_len - sT0 + CHS - 1) / CHS;
            if (lo < 0) lo = 0;
            long long hi = (V - sL0) / CHS;
            if (hi > nchunks) hi = nchunks;
            if (lo < hi) { c_lo = (int)lo; c_hi = (int)hi; }
            a16 = (((reinterpret_cast<uintptr_t>(in_mt + sT0) | reinterpret_cast<uintptr_t>(in_mt + sL0)) & 15) == 0) ? 1 : 0;
            if (lane == 0 && c_lo < c_hi && a16)
                fz_prefetch_l2(in_mt + sL0 + (long long)c_lo * CHS, (unsigned)min(FZ_PF, c_hi - c_lo) * CHS * 8);
        }
        const bool pre0 = c_lo == 0 && c_hi > 0;
        __syncwarp();                                                 // the previous unit's last reads of the staged blocks
        if (pre0) {
            if (a16) {
                fz_issue<S, true>(Tst, in_mt + (long long)kA * S, lane);
                fz_issue<S, true>(Lst, in_mt + (long long)(kA + lag) * S, lane);
            } else {
                fz_issue<S, false>(Tst, in_mt + (long long)kA * S, lane);
                fz_issue<S, false>(Lst, in_mt + (long long)(kA + lag) * S, lane);
            }
        }
        // carried window sums: rows [kA, kA+lag) per phase, exact double sums
        {
            const long long s0 = (long long)kA * S;
            double acc = 0.0;
            if (wact) {
                for (int i = wg; i < lag; i += G) {
                    const long long v = s0 + (long long)i * S + wp;
                    float2 x = make_float2(0.f, 0.f);
                    if (v < V) x = (v < tail_len) ? tailp[v] : __ldg(in_mt + v);
                    acc = daddr(acc, (double)energy_f32(x.x, x.y));
                }
                ebuf[wg * ES + wp] = acc;
            }
            __syncwarp();
            if (lane < S) {
                double Cw = 0.0;
#pragma unroll
                for (int g2 = 0; g2 < G; g2++) Cw = daddr(Cw, ebuf[g2 * ES + lane]);
                cwp[lane] = Cw;
            }
            __syncwarp();
        }
        int16_t* o_sidx = prm.out_sidx + dgp->sym_off;
        float* o_th = prm.theta + dgp->scr_off;
        float2* o_sel = prm.sel + dgp->scr_off;
        if (a16) fzs_front_chunks<S, true>(wb, lane, kA, kB, lag, c_lo, c_hi, nchunks, pre0, dgp->M, in_mt, tailp, tail_len, V, o_sidx, o_th, o_sel);
        else     fzs_front_chunks<S, false>(wb, lane, kA, kB, lag, c_lo, c_hi, nchunks, pre0, dgp->M, in_mt, tailp, tail_len, V, o_sidx, o_th, o_sel);
    }
}

// ---- chain + back kernel ------------------------------------------------------------------------
template <int PC> struct FzsCbL {        // per-warp shared memory of the chain + back kernel (no staged blocks)
    static constexpr int S_STATIC = 0;
    static constexpr bool TRACK_N = true;                                   // exact first / last unwrap counts into the end record
    static constexpr int OFF_TH = 0;                                        // float  th[FZ_BUF]
    static constexpr int OFF_SEL = OFF_TH + FZ_BUF * 4;                     // float selx[FZ_SELN], sely[FZ_SELN]; [1] = previous sample
    static constexpr int OFF_YH = fz_align16(OFF_SEL + (FZ_BUF + 2) * 8);   // float  yh[PC]
    static constexpr int OFF_CTX = fz_align16(OFF_YH + PC * 4);
    static constexpr int OFF_CZ = fz_align16(OFF_CTX + (int)sizeof(FzCtx)); // double cz[PC + 1], ends where ALIAS starts
    static constexpr int OFF_ALIAS = OFF_CZ + fz_align16((PC + 1) * 8);
    static constexpr int C_BYTES = FZ_B * 8 + FZ_B * 4 + (FZ_B + 4) * 4;    // prefix block, y block, est block; back: soft + bits staging
    static constexpr int BYTES = fz_align128(OFF_ALIAS + fz_align16(C_BYTES));
};

struct FzsCbParams {
    const ChanDesc* desc; ChanState* state; float* ring_base; int n_channels;
    const float2* sel; const float* theta;
    float2* out_soft; int16_t* out_bits; float* out_phase; uint8_t* out_hard;
    double sri_xdelta; DevCounters* counters;
    TpCtl tp;                          // items: one unit per item; else one unit per channel (all its packets, from state[ch])
    int n_units; int* ticket;
};

// set-up of one chain unit; false: nothing to do for this ticket
template <int PC>
static __device__ __noinline__ bool fzs_cb_begin(const FzsCbParams& prm, const unsigned wofs, const int u)
{
    using L = FzsCbL<PC>;
    unsigned char* wb = fz_smem + wofs;
    float*  selx = reinterpret_cast<float*>(wb + L::OFF_SEL);
    float*  yh   = reinterpret_cast<float*>(wb + L::OFF_YH);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    const int lane = threadIdx.x & 31;
    const TpCtl& tp = prm.tp;
    int ch, pk_a = 0, pk_b = -1, kind = 0, src = -1, dst = -1, pkt_slot = -1;
    if (tp.items) {
        const TpItem it = tp.items[u];
        ch = it.ch; pk_a = it.pk_a; pk_b = it.pk_b; kind = it.kind; src = it.src; dst = it.dst; pkt_slot = it.pkt_slot;
        if (!(prm.desc[ch].flags & CH_FZS)) return false;
        if (tp.rerun) {                                   // repair round: only what k_tp_fix asked for
            const int r = tp.slot_run[pkt_slot];
            if (r == 0) return false;
            kind = r;
            if (r == 2) src = dst - 1;                    // from the proven predecessor's exact end record
        }
    } else {
        ch = u;
        const int fl = prm.desc[ch].flags;
        if (!(fl & CH_FZS)) return false;
        if (tp.fallback) { if (!(fl & CH_TP) || !tp.fail[ch]) return false; }   // re-run of a channel whose hand-overs were not proven
        else if (fl & CH_TP) return false;                                       // handled through TpItems
    }
    const ChanDesc* dgp = prm.desc + ch;
    const int P = dgp->P, M = dgp->M, S = dgp->S, A = dgp->A, n_pkts = dgp->n_pkts;
    const int K = (int)dgp->K;
    const long long pkt_len = dgp->pkt_len, tail_len = dgp->tail_len;
    const float* thg = prm.theta + dgp->scr_off;
    if (pk_b < 0) pk_b = n_pkts;
    const int k_begin = (pk_a == 0) ? 0 : (int)first_symbol_at((long long)pk_a * pkt_len, tail_len, S, A, K);
    __syncwarp();
    if (kind == 0) {
        if (lane == 0) { cx.st = prm.state[ch]; cx.flags = dgp->flags; }
        const float* gring = prm.ring_base + dgp->ring_off;
        for (int j = lane; j < P; j += 32) yh[j] = gring[j];
    } else if (kind == 2) {
        if (lane == 0) { cx.st = tp.ends[src].st; cx.flags = dgp->flags & ~(CH_RESET_NUMSYMS | CH_RESET_PHASEAVG); }
        for (int j = lane; j < P; j += 32) {
            const float v = tp.end_ring[(size_t)src * tp.ring_stride + j];
            yh[j] = v;
            if (dst >= 0) tp.start_ring[(size_t)dst * tp.ring_stride + j] = v;       // what k_tp_check compares
        }
    } else {
        // synthesised start of packet pk_a (see k_chain_par): the history ring the previous packet leaves behind,
        // from the resolved integers: y = f32(theta + 2pi(c + A)), then the packet-end shift f32(y - w*wrapValue)
        const TpPacket pp = tp.pkts[pkt_slot - 1];
        const float wrapValue = __double2float_rn(dmulr(PSKD_M_2PI, (double)M));
        const float shift = fmulr((float)pp.w, wrapValue);
        int carry = 0;
        for (int e = k_begin; e > k_begin - P; e -= 32) {
            const int m = e - 1 - lane;
            int dn = 0;
            float t = 0.0f;
            const bool in = m >= k_begin - P;
            if (in) { t = __ldg(thg + m); dn = -__float2int_rn((t - __ldg(thg + m - 1)) * 0.15915494309189535f); }
            const int incl = warp_scan_int(dn, lane);
            if (in) {
                const int c = pp.cEnd - (carry + incl - dn);
                float y = __double2float_rn(daddr((double)t, dmulr((double)(c + pp.A), PSKD_M_2PI)));
                if (pp.w != 0) y = fsubr(y, shift);                                      // :131 (subtractConst)
                yh[m - (k_begin - P)] = y;
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        if (lane == 0) {
            cx.st = (src >= 0) ? tp.ends[src].st : prm.state[ch];   // fit constants (xdelta, denominator, n) of the channel
            cx.st.fit.head = 0; cx.st.fit.pts = P; cx.st.wraps = 0;
            SmemRing ring{yh};
            cx.st.est = fit_resum(cx.st.fit, ring);
            cx.flags = dgp->flags & ~(CH_RESET_NUMSYMS | CH_RESET_PHASEAVG);
        }
        __syncwarp();
        if (dst >= 0) for (int j = lane; j < P; j += 32) tp.start_ring[(size_t)dst * tp.ring_stride + j] = yh[j];
    }
    __syncwarp();
    if (lane == 0) {
        const long long sym_off = dgp->sym_off;
        cx.passes = 0; cx.seq_blocks = 0; cx.blocks = 0;
        cx.desc = dgp;
        cx.o_soft = prm.out_soft ? prm.out_soft + sym_off : nullptr;
        cx.o_phase = prm.out_phase ? prm.out_phase + sym_off : nullptr;
        cx.o_bits = prm.out_bits ? prm.out_bits + dgp->bits_off : nullptr;
        cx.o_sidx = nullptr;
        cx.o_hard = prm.out_hard ? prm.out_hard + sym_off : nullptr;
        cx.sri_xdelta = prm.sri_xdelta;
        cx.tail_len = tail_len; cx.pkt_len = pkt_len;
        cx.n_pkts = n_pkts; cx.pk1 = pk_b; cx.K = K; cx.A = A; cx.M = M; cx.P = P; cx.bpb = dgp->bpb; cx.diff = dgp->D;
        cx.pkt = pk_a; cx.kchain = k_begin; cx.nbuf = 0; cx.cz_valid = 0; cx.unit_done = (pk_a >= pk_b) ? 1 : 0;
        cx.fP1 = (float)(P - 1);
        cx.pk_hi = (pk_a + 1 >= n_pkts) ? K : (int)first_symbol_at((long long)(pk_a + 1) * pkt_len, tail_len, S, A, K);
        cx.end_mid = 0;
        cx.ch = ch; cx.S_rt = S; cx.dst = dst; cx.k_begin = k_begin;
        cx.est_start_used = cx.st.est;
        cx.n_first = 0; cx.n_last = 0; cx.have_first = 0; cx.est_pre = cx.st.est;
        cx.wraps0 = cx.st.wraps;
        fz_sel_put(selx, 1, (k_begin > 0) ? prm.sel[dgp->scr_off + k_begin - 1] : cx.st.last);     // :486-489
        if (pk_a < pk_b) fz_prologue(cx, yh);
    }
    __syncwarp();
    return true;
}

template <int PC>
static __device__ __noinline__ void fzs_cb_end(const FzsCbParams& prm, const unsigned wofs)
{
    using L = FzsCbL<PC>;
    unsigned char* wb = fz_smem + wofs;
    float*  yh   = reinterpret_cast<float*>(wb + L::OFF_YH);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    const int lane = threadIdx.x & 31;
    const int P = cx.P, ch = cx.ch, dst = cx.dst;
    const TpCtl& tp = prm.tp;
    if (cx.st.fit.head != 0 && cx.st.fit.pts == P)
        fz_normalize_ring(yh, reinterpret_cast<float*>(wb + L::OFF_ALIAS), cx.st.fit, P, lane);
    __syncwarp();
    if (dst >= 0) {
        for (int j = lane; j < P; j += 32) tp.end_ring[(size_t)dst * tp.ring_stride + j] = yh[j];
        if (lane == 0) {
            TpEnd& e = tp.ends[dst];
            e.st = cx.st; e.est_start_used = cx.est_start_used; e.has_symbols = (cx.kchain > cx.k_begin) ? 1 : 0;
            e.wraps_delta = cx.st.wraps - cx.wraps0;
            e.n_first = cx.n_first; e.n_last = cx.n_last; e.est_pre = cx.est_pre; e.pad = 0;
        }
    } else {
        float* gring = prm.ring_base + cx.desc->ring_off;
        for (int j = lane; j < P; j += 32) gring[j] = yh[j];
        if (lane == 0) prm.state[ch] = cx.st;              // `last` is carried by k_finish
    }
    if (lane == 0) {
        if (cx.st.wraps != cx.wraps0) atomicAdd(&prm.counters->wraps, cx.st.wraps - cx.wraps0);
        if (cx.blocks) atomicAdd(&prm.counters->spec_chunks, (unsigned long long)cx.blocks);
        if (cx.passes) atomicAdd(&prm.counters->spec_misses, (unsigned long long)cx.passes);
        if (cx.seq_blocks) atomicAdd(&prm.counters->seq_channels, (unsigned long long)cx.seq_blocks);
    }
    __syncwarp();
}

template <int PC>
__global__ void __launch_bounds__(FZ_WARPS * 32, PSKD_FZS_CB_MIN_CTAS)
k_fzs_cb(const FzsCbParams prm)
{
    using L = FzsCbL<PC>;
    const int lane = fz_lane();
    const unsigned wofs = (threadIdx.x >> 5) * (unsigned)L::BYTES;
    unsigned char* wb = fz_smem + wofs;
    float*  th   = reinterpret_cast<float*>(wb + L::OFF_TH);
    float*  selx = reinterpret_cast<float*>(wb + L::OFF_SEL);
    FzCtx&  cx   = *reinterpret_cast<FzCtx*>(wb + L::OFF_CTX);
    if (prm.tp.rerun && *prm.tp.any_rerun == 0) return;       // a repair round nobody asked for

    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(prm.ticket, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= prm.n_units) break;
        if (!fzs_cb_begin<PC>(prm, wofs, u)) continue;
        const float* thg = prm.theta + cx.desc->scr_off;
        const float2* selg = prm.sel + cx.desc->scr_off;
        for (;;) {
            fz_drain<L>(wofs);
            if (cx.unit_done) break;
            // the next block of (angle, sample) pairs from the scratch, appended to the block buffer
            const int nbuf = cx.nbuf, kchain = cx.kchain, k0 = kchain + nbuf;
            const int want = min(FZ_B, cx.pk_hi - kchain);
            const int need = want - nbuf;
            __syncwarp();
            if (nbuf == 0 && need == FZ_B && ((reinterpret_cast<uintptr_t>(thg + k0) | reinterpret_cast<uintptr_t>(selg + k0)) & 15) == 0) {
                // a full block at an aligned position (every block of a packet but its first and last)
#ifndef PSKD_FZS_CB_NO_PF
                // the block after it towards L2 (the scratch of a large call has left L2 by the time the chain reads it: an
                // HBM round trip per block otherwise, a fifth of this kernel's stall cycles)
                if (lane == 0 && k0 + 2 * FZ_B <= cx.pk_hi) {
                    fz_prefetch_l2(thg + k0 + FZ_B, FZ_B * 4u);
                    fz_prefetch_l2(selg + k0 + FZ_B, FZ_B * 8u);
                }
#endif
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(thg + k0) + lane);
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(selg + k0) + lane);
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(selg + k0) + 32 + lane);
                reinterpret_cast<float4*>(th)[lane] = t4;
                float2* sx2 = reinterpret_cast<float2*>(selx + 2);
                float2* sy2 = reinterpret_cast<float2*>(selx + FZ_SELN + 2);
                sx2[lane] = make_float2(s0.x, s0.z);      sy2[lane] = make_float2(s0.y, s0.w);
                sx2[32 + lane] = make_float2(s1.x, s1.z); sy2[32 + lane] = make_float2(s1.y, s1.w);
            } else {
#pragma unroll
                for (int q = 0; q < FZ_B / 32; q++) {
                    const int i = lane + 32 * q;
                    if (i < need) { th[nbuf + i] = __ldg(thg + k0 + i); fz_sel_put(selx, 2 + nbuf + i, __ldg(selg + k0 + i)); }
                }
            }
            if (lane == 0) cx.nbuf = want;
            __syncwarp();
        }
        fzs_cb_end<PC>(prm, wofs);
    }
#ifdef PSKD_FZS_CB_PAD_CODE
    // layout experiment (tools/probe/ab_layout.sh): never-executed instructions between the loop body and the out-of-line stage functions
    if (prm.n_units == -12345) {
        unsigned acc = 0, t;
#define PSKD_CBPAD1 asm volatile("mov.u32 %0, %%clock;" : "=r"(t)); acc ^= t;
#define PSKD_CBPAD8 PSKD_CBPAD1 PSKD_CBPAD1 PSKD_CBPAD1 PSKD_CBPAD1 PSKD_CBPAD1 PSKD_CBPAD1 PSKD_CBPAD1 PSKD_CBPAD1
#pragma unroll
        for (int i = 0; i < PSKD_FZS_CB_PAD_CODE; i++) { PSKD_CBPAD8 }
        prm.ticket[1] = (int)acc;
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int S, int PC, int CT, int BK = -1>
static cudaError_t launch_fused_ct(const LaunchCtx& c, const FusedLaunch& f, const FusedParams& p, double alg_bytes) {
    using L = FzL<S, PC>;
    static const size_t pad = getenv("PSKD_FZ_PAD_SMEM") ? (size_t)atoi(getenv("PSKD_FZ_PAD_SMEM")) : 0;   // tuning: lowers occupancy
    static const int carve = getenv("PSKD_FZ_CARVEOUT") ? atoi(getenv("PSKD_FZ_CARVEOUT")) : (int)cudaSharedmemCarveoutMaxShared;   // tuning: % of the maximum
    const size_t smem = (size_t)L::BYTES * FZ_WARPS + pad;
    static KernelCfg cfg;                       // per device (function attributes and occupancy are per device)
    int ctas_per_sm = 0, n_sm = 0;
    cudaError_t e = cfg.ensure(k_fused<S, PC, CT, BK>, smem, FZ_WARPS * 32, &ctas_per_sm, &n_sm, carve);
    if (e != cudaSuccess) return e;
    int grid = n_sm * ctas_per_sm;
    if (f.grid_share > 0.0 && f.grid_share < 1.0) grid = (int)(grid * f.grid_share + 0.999);   // co-resident launches share the SMs
    const int need = (p.n_units + FZ_WARPS - 1) / FZ_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    c.prof->begin(S == 8 ? KID_FUSED : S == 9 ? KID_FUSED_S9 : S == 10 ? KID_FUSED_S10 : KID_FUSED_S16, c.stream, alg_bytes);
    k_fused<S, PC, CT, BK><<<grid, FZ_WARPS * 32, smem, c.stream>>>(p);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}
template <int S, int PC>
static cudaError_t launch_fused_t(const LaunchCtx& c, const FusedLaunch& f) {
    using L = FzL<S, PC>;
    FusedParams p{};
    p.desc = c.d_desc; p.state = c.d_state; p.ring_base = c.d_ring;
    p.list = f.d_list; p.n_list = f.n_list;
    p.pkts_per_unit = f.pkts_per_unit; p.parts_per_pkt = f.parts_per_pkt;
    p.n_units = f.n_list * f.units_per_channel;
    p.ticket = f.d_ticket; p.done = f.d_done;
    p.out_soft = (float2*)c.out_soft; p.out_bits = c.out_bits; p.out_phase = c.out_phase; p.out_sidx = c.out_sidx; p.out_hard = c.out_hard;
    p.sri_xdelta = c.sri_xdelta;
    p.counters = c.d_counters;
    double ab = 0.0;
    if (c.prof->enabled && f.h_list)
        for (int i = 0; i < f.n_list; i++) { const ChanDesc& d = c.h_desc[f.h_list[i]]; ab += alg_bytes_front(d) + alg_bytes_chain(d) + alg_bytes_back(d); }
    // 6 CTAs per SM when the launch's channels fill them (every channel is one sequential chain of units: with fewer channels
    // than resident warps the extra warps idle and the 80-register code is the slower one per warp).  Measured on the bench bank:
    // 4096 channels 12.2 vs 12.5 ms with 6, 3600 channels (48 more than the 3552 resident warps) 11.2 vs 11.1, 3072 channels 10.5 vs
    // 9.5 ms: six from 1.1 x the resident warps on.
    static const int force_ct = getenv("PSKD_FZ_CTAS") ? atoi(getenv("PSKD_FZ_CTAS")) : 0;          // tuning: 5 or 6
    int n_sm = 0, dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    constexpr bool fits6 = ((size_t)L::BYTES * FZ_WARPS + 1024) * 6 <= 228 * 1024;               // shared memory of six CTAs (S = 8 only)
    bool six = false;
    if constexpr (fits6) six = force_ct ? (force_ct >= 6) : (f.n_list * 10 >= n_sm * 6 * FZ_WARPS * 11 && !(f.grid_share > 0.0 && f.grid_share < 1.0));
#ifndef PSKD_FZ_NO_BACK_SPEC
    // the launch's channels share one back-stage variant (a uniform bank: the usual case) and no packed hard symbols are asked for:
    // the kernel with that variant inlined (coherent BPSK / QPSK / 8-PSK at samplesPerBaud 8, phaseAvg <= 52)
    if constexpr (S == 8 && PC == 52) {
        int bk = -1;
        if (!c.out_hard && f.h_list && f.n_list > 0) {
            const ChanDesc& d0 = c.h_desc[f.h_list[0]];
            bk = d0.bpb * 2 + (d0.D ? 1 : 0);
            for (int i = 1; i < f.n_list && bk >= 0; i++) { const ChanDesc& d = c.h_desc[f.h_list[i]]; if (d.bpb * 2 + (d.D ? 1 : 0) != bk) bk = -1; }
        }
        if constexpr (fits6) {
            if (six) {
                if (bk == 6) return launch_fused_ct<S, PC, 6, 6>(c, f, p, ab);
                if (bk == 4) return launch_fused_ct<S, PC, 6, 4>(c, f, p, ab);
                if (bk == 2) return launch_fused_ct<S, PC, 6, 2>(c, f, p, ab);
            }
        }
        if (!six) {
            if (bk == 6) return launch_fused_ct<S, PC, PSKD_FZ_MIN_CTAS, 6>(c, f, p, ab);
            if (bk == 4) return launch_fused_ct<S, PC, PSKD_FZ_MIN_CTAS, 4>(c, f, p, ab);
            if (bk == 2) return launch_fused_ct<S, PC, PSKD_FZ_MIN_CTAS, 2>(c, f, p, ab);
        }
    }
#endif
    if constexpr (fits6) { if (six) return launch_fused_ct<S, PC, 6>(c, f, p, ab); }
    return launch_fused_ct<S, PC, PSKD_FZ_MIN_CTAS>(c, f, p, ab);
}

// ---- staged path through the fused kernel's stages ---------------------------------------------
static int* fzs_take_ticket(const LaunchCtx& c) {
    if (!c.d_fzs_ticket || !c.fzs_ticket_next || *c.fzs_ticket_next >= c.fzs_ticket_cap) return nullptr;
    return c.d_fzs_ticket + (*c.fzs_ticket_next)++;
}

template <int S>
static cudaError_t launch_fzs_front_t(const LaunchCtx& c) {
    using L = FzsFL<S>;
    const size_t smem = (size_t)L::BYTES * FZ_WARPS;
    static KernelCfg cfg;
    int ctas_per_sm = 0, n_sm = 0;
    cudaError_t e = cfg.ensure(k_fzs_front<S>, smem, FZ_WARPS * 32, &ctas_per_sm, &n_sm);
    if (e != cudaSuccess) return e;
    FzsFrontParams p{};
    p.desc = c.d_desc; p.n_channels = c.n_channels;
    // unit = a segment of one channel: >= ~8 units per resident warp when the call is large enough, 1024..8192 symbols
    static const int seg_env = getenv("PSKD_FZS_SEG") ? atoi(getenv("PSKD_FZS_SEG")) : 0;
    const long long slots = (long long)n_sm * ctas_per_sm * FZ_WARPS;
    long long seg = seg_env > 0 ? seg_env : (c.Kmax_fzs * c.n_fzs_channels) / (8 * slots);
    seg = (seg + 31) & ~31LL;
    if (seg < 1024) seg = 1024;
    if (seg > 8192 && seg_env <= 0) seg = 8192;
    p.seg_syms = (int)seg;
    const long long nseg = (c.Kmax_fzs + seg - 1) / seg;
    if (nseg * c.n_channels > 0x7fffffffLL) return cudaErrorInvalidValue;
    p.n_units = (int)(nseg * c.n_channels);
    p.ticket = fzs_take_ticket(c);
    if (!p.ticket) return cudaErrorInvalidValue;
    p.sel = c.d_sel; p.theta = c.d_theta; p.out_sidx = c.out_sidx;
    static const int cta_cap = getenv("PSKD_FZS_FRONT_CTAS") ? atoi(getenv("PSKD_FZS_FRONT_CTAS")) : 0;   // tuning: resident CTAs per SM
    int grid = n_sm * ((cta_cap > 0 && cta_cap < ctas_per_sm) ? cta_cap : ctas_per_sm);
    const int need = (p.n_units + FZ_WARPS - 1) / FZ_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    double ab = 0.0;
    if (c.prof->enabled) for (int i = 0; i < c.n_channels; i++) { const ChanDesc& d = c.h_desc[i]; if ((d.flags & CH_FZS) && d.S == S) ab += alg_bytes_front(d); }
    c.prof->begin(KID_FZS_FRONT, c.stream, ab);
    k_fzs_front<S><<<grid, FZ_WARPS * 32, smem, c.stream>>>(p);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

cudaError_t launch_fzs_front(const LaunchCtx& c) {
    if (c.n_fzs_channels == 0 || c.Kmax_fzs <= 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (c.S_mask_fzs & (1ull << 8))  { e = launch_fzs_front_t<8>(c);  if (e != cudaSuccess) return e; }
    if (c.S_mask_fzs & (1ull << 9))  { e = launch_fzs_front_t<9>(c);  if (e != cudaSuccess) return e; }
    if (c.S_mask_fzs & (1ull << 10)) { e = launch_fzs_front_t<10>(c); if (e != cudaSuccess) return e; }
    if (c.S_mask_fzs & (1ull << 16)) { e = launch_fzs_front_t<16>(c); if (e != cudaSuccess) return e; }
    return e;
}

template <int PC>
static cudaError_t launch_fzs_cb_t(const LaunchCtx& c, const TpCtl& tp, int n_units, double alg_bytes) {
    using L = FzsCbL<PC>;
    const size_t smem = (size_t)L::BYTES * FZ_WARPS;
    static KernelCfg cfg;
    int ctas_per_sm = 0, n_sm = 0;
    cudaError_t e = cfg.ensure(k_fzs_cb<PC>, smem, FZ_WARPS * 32, &ctas_per_sm, &n_sm);
    if (e != cudaSuccess) return e;
    FzsCbParams p{};
    p.desc = c.d_desc; p.state = c.d_state; p.ring_base = c.d_ring; p.n_channels = c.n_channels;
    p.sel = c.d_sel; p.theta = c.d_theta;
    p.out_soft = (float2*)c.out_soft; p.out_bits = c.out_bits; p.out_phase = c.out_phase; p.out_hard = c.out_hard;
    p.sri_xdelta = c.sri_xdelta; p.counters = c.d_counters;
    p.tp = tp;
    p.n_units = n_units;
    p.ticket = fzs_take_ticket(c);
    if (!p.ticket) return cudaErrorInvalidValue;
    static const int cta_cap = getenv("PSKD_FZS_CB_CTAS") ? atoi(getenv("PSKD_FZS_CB_CTAS")) : 0;   // tuning: resident CTAs per SM
    int grid = n_sm * ((cta_cap > 0 && cta_cap < ctas_per_sm) ? cta_cap : ctas_per_sm);
    const int need = (n_units + FZ_WARPS - 1) / FZ_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) return cudaSuccess;
    c.prof->begin(KID_FZS_CB, c.stream, alg_bytes);
    k_fzs_cb<PC><<<grid, FZ_WARPS * 32, smem, c.stream>>>(p);
    c.prof->end(c.stream);
    (*c.launches)++;
    return cudaGetLastError();
}

// one launch of the chain + back kernel: tp.items != null -> one unit per item, else one unit per channel
cudaError_t launch_fzs_cb(const LaunchCtx& c, const TpCtl& tp, int n_units, double alg_bytes) {
    if (n_units <= 0) return cudaSuccess;
    // (the chain + back kernel with the back-stage variant inlined, as k_fused has it, measured no gain: 1.90 vs 1.89 ms on the 512-channel shard)
    return c.Pmax_fzs <= 52 ? launch_fzs_cb_t<52>(c, tp, n_units, alg_bytes) : launch_fzs_cb_t<128>(c, tp, n_units, alg_bytes);
}

bool fzs_supports(int S, int A, int P) { return fused_supports(S, A, P); }

bool fused_supports(int S, int A, int P) {
    if (!(S == 8 || S == 9 || S == 10 || S == 16)) return false;
    if (A < 1 || A > FUSED_AMAX) return false;
    if (P < 1 || P > FUSED_PMAX) return false;
    return true;
}

// two shared-memory size classes: phaseAvg <= 52 (the component's default and below: 9.3 KB per warp at
// S = 8, six CTAs of four warps per SM) and everything up to FUSED_PMAX.  numAvg does not enter: the
// kernel keeps no energy ring (FUSED_AMAX only bounds how far behind the trail block is re-read).
cudaError_t launch_fused(const LaunchCtx& c, const FusedLaunch& f) {
    if (f.n_list == 0) return cudaSuccess;
    const bool small = f.Pmax <= 52;
    switch (f.S) {
        case 8:  return small ? launch_fused_t<8, 52>(c, f) : launch_fused_t<8, 128>(c, f);
        case 9:  return small ? launch_fused_t<9, 52>(c, f) : launch_fused_t<9, 128>(c, f);
        case 10: return small ? launch_fused_t<10, 52>(c, f) : launch_fused_t<10, 128>(c, f);
        case 16: return launch_fused_t<16, 128>(c, f);
    }
    return cudaErrorInvalidValue;
}

}  // namespace pskd
