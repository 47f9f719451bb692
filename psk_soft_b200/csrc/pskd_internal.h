// pskd_internal.h -- structures shared by the host runtime (pskd_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "pskd_exact.cuh"

namespace pskd {

// Per-channel, per-call descriptor.  Filled on the host, read by every kernel of the call.
// The channel's logical input is the VIRTUAL stream  tail ++ in  (tail = the samples the
// reference would still hold in its `samples` deque, cpp/psk_soft.h:66).
struct ChanDesc {
    const float2* in;        // this call's new samples
    const float2* tail;      // carried samples (device, owned by the bank)
    float2*       tail_next; // where the samples to carry into the next call are written
    long long n_in;          // new complex samples
    long long tail_len;      // carried complex samples (< S*A in steady state)
    long long K;             // symbols emitted by this call
    long long next_tail_len; // tail_len + n_in - K*S
    long long sym_off;       // element offset of this channel in soft/phase/sample_index
    long long bits_off;      // element offset of this channel in bits
    long long scr_off;       // element offset of this channel in the scratch arrays
    long long pkt_len;       // emulated BULKIO packet length in complex samples
    int n_pkts;              // emulated packets in this call
    int S, A, M, P, D, bpb;
    int ring_off;            // element offset of this channel's y-history ring
    int flags;               // CH_* below
};
enum { CH_RESET_NUMSYMS = 1, CH_RESET_PHASEAVG = 2, CH_SRI_CHANGED = 4,
       CH_FAST = 8 /* phase chain + back run in k_chain_par (scan-based), else k_chain_seq + k_back */,
       CH_FRONT_FAST = 16 /* timing runs in k_front_t<S>, else in the generic k_front */,
       CH_FUSED = 32 /* the whole path of this channel runs in k_fused<S> (pskd_fused.cu); the staged kernels skip it */,
       CH_TP = 64 /* staged path: the phase chain of this channel runs time-parallel over its packets (TpCtl) */,
       CH_FZS = 128 /* staged path through the fused kernel's stages: k_fzs_front<S> + k_fzs_cb (pskd_fused.cu) instead of
                       k_front_t + k_chain_par + k_back_par */,
       CH_STALLED = 256 /* the timing window is over-full (numAvg*samplesPerBaud shrank): the channel consumes its input and emits
                           nothing (cpp/psk_soft.cpp:457); of the packet prologue only what a pending reset asks for runs */ };
constexpr int FRONT_FAST_AMAX = 768;   // largest numAvg the specialised front tile (1024 symbols) still uses efficiently
constexpr int CHAIN_PAR_PMAX = 1024;
constexpr int FUSED_AMAX = 256;        // largest numAvg whose energy ring the fused kernel keeps in shared memory
constexpr int FUSED_PMAX = 128;        // largest phaseAvg of the fused kernel's chain blocks   // largest phaseAvg the scan-based chain stages in shared memory

// Carried phase-tracking state of one channel (cpp/psk_soft.h:70-85 minus the timing deques).
struct ChanState {
    FitState fit;
    float  est;          // phaseEstimate
    float  sampleRate;   // psk_soft_i::sampleRate
    float2 last;         // psk_soft_i::last
    unsigned long long wraps;
};

struct DevCounters {
    unsigned long long wraps, spec_chunks, spec_misses, seq_channels, tp_packets, tp_repaired;
};

// first call-local output symbol whose emission sample lies at or after input position x
// (symbol k is emitted when virtual sample (k+A)*S-1 arrives, cpp/psk_soft.cpp:454-457)
__host__ __device__ inline long long first_symbol_at(long long x, long long tail_len, int S, int A, long long K) {
    long long v = x + tail_len + 1;
    long long k = (v + S - 1) / S - A;
    if (k < 0) k = 0;
    if (k > K) k = K;
    return k;
}

// ---- time-parallel phase chain (staged path, few channels with many packets) -------------------
// The chain of a channel is sequential only through (a) the history ring handed from packet to
// packet and (b) the integer level of the unwrapped phase, which the packet-end wrap
// (cpp/psk_soft.cpp:592-603) turns into a scalar recurrence over packets.  So: every packet's classic
// (sample-to-sample) unwrap and an estimate of its end phase are formed in parallel (k_tp_scan), a
// scan over the packets of a channel resolves the integer levels and the wrap counts (k_tp_resolve),
// every packet then runs the exact chain from a history ring SYNTHESISED from those integers
// (k_chain_par over TpItems), and k_tp_check proves the hand-overs: the ring a packet started from
// must equal, bit for bit, the ring its predecessor ended with, and its first unwrap count must be
// the one the predecessor's exact end estimate gives.  Channels that fail the proof are re-run by
// the sequential chain from their untouched carried state.
struct TpItem {            // one warp of work for k_chain_par
    int ch;                // channel (index into d_desc)
    int pk_a, pk_b;        // emulated packets [pk_a, pk_b)
    int kind;              // 0: start from state[ch]; 1: synthesised start of packet pk_a; 2: start from end record `src`
    int src;               // kind 2: end record to start from; kind 1: end record holding the channel's fit constants (< 0: state[ch])
    int dst;               // end record this item writes
    int pkt_slot;          // index of packet pk_a in the TpPacket array
    int pad;
};
struct TpPacket {          // per (channel, packet) of the time-parallel range
    int cEnd;              // classic unwrap count at the packet's last symbol, relative to its first symbol
    int dLink;             // classic increment from the previous packet's last symbol to this packet's first
    int A;                 // resolved: unwrap count of the packet's first symbol
    int w;                 // resolved: numWraps the packet-end wrap applies (0: none)
    double estRelEnd;      // estimate at the packet end for the relative phases (A = 0)
    long long klo, khi;    // symbols of the packet
};
struct TpEnd {             // what a chain item leaves behind
    ChanState st;          // state after the item's last packet epilogue
    float est_start_used;  // the estimate the item's first symbol was unwrapped against
    int   has_symbols;
    unsigned long long wraps_delta;
    // k_fzs_cb items only (the repair round of k_tp_fix): the EXACT unwrap counts of the item's first and last
    // symbol (in the frame of the level it ran at) and the estimate before the last packet-end wrap
    int   n_first, n_last;
    float est_pre;
    int   pad;
};
struct TpChan {            // per time-parallel channel
    int ch, pkt0, n_pkts;  // first time-parallel packet, packets in the call
    int first_item;        // index of the channel's first end record (the sequential head's, if any); records follow in packet order
    int first_slot;        // index of packet pkt0 in the TpPacket array
    int has_head;          // 1: packets [0, pkt0) ran as a sequential head item (record first_item); 0: pkt0 == 0, the first
                           //    packet starts from the channel's carried state (its history is known to be full)
};
struct TpCtl {
    const TpItem* items; int n_items;
    const TpChan* chans; int n_chans;
    TpPacket* pkts;
    TpEnd* ends; float* end_ring; float* start_ring; int ring_stride;   // per item: ring after the last epilogue / ring the item started from
    int* fail;             // [n_channels] set by k_tp_check when a hand-over could not be proven
    int fallback;          // launch flag: process only channels with fail[ch] != 0, from state[ch]
    // repair round (k_tp_fix, k_fzs_cb channels): per packet slot
    int* slot_fail;        // set by k_tp_check: the hand-over INTO this packet could not be proven
    int* slot_run;         // set by k_tp_fix: 0 leave, 1 re-run from a re-synthesised ring, 2 re-run from the predecessor's end record
    int rerun;             // launch flag: process only items with slot_run != 0
    int* any_rerun;        // set by k_tp_fix when some channel asked for a repair round (else the round's launches return at once)
};

// ---- optional per-kernel event timing --------------------------------------------------------
enum KernelId { KID_FRONT = 0, KID_CHAIN_SEQ, KID_CHAIN_PAR, KID_BACK_PAR, KID_CHAIN_EXACT, KID_BACK, KID_FINISH, KID_FUSED, KID_TP, KID_FZS_FRONT, KID_FZS_CB, KID_FUSED_S9, KID_FUSED_S10, KID_FUSED_S16, KID_COUNT };
struct Profiler {
    bool enabled = false;
    struct Pair { cudaEvent_t a, b; int kid; };
    Pair* pending = nullptr; int n_pending = 0, cap_pending = 0;
    cudaEvent_t* pool = nullptr; int n_pool = 0, cap_pool = 0;
    double ms[KID_COUNT] = {0}; unsigned long long launches[KID_COUNT] = {0};
    double bytes[KID_COUNT] = {0};   // algorithmic bytes of the channels the timed launches served
    cudaEvent_t get();
    void begin(int kid, cudaStream_t s, double alg_bytes = 0.0);
    void end(cudaStream_t s);
    void drain();          // caller has synchronised the stream
    void destroy();
};
const char* kernel_name(int kid);

// ---- per-device launch configuration of one kernel --------------------------------------------
// cudaFuncSetAttribute (dynamic shared memory opt-in, carve-out) and the occupancy it yields are PER DEVICE:
// a process that drives several GPUs (one bank per GPU, one host thread each -- include/pskd.h) must
// configure every kernel once on every device it launches on.
constexpr int PSKD_MAX_DEVICES = 64;
struct KernelCfg {
    std::mutex mu;
    size_t smem[PSKD_MAX_DEVICES] = {};
    int ctas[PSKD_MAX_DEVICES] = {}, nsm[PSKD_MAX_DEVICES] = {};
    bool done[PSKD_MAX_DEVICES] = {};
    // make `kernel` launchable with `need` bytes of dynamic shared memory on the current device; optionally
    // returns resident CTAs per SM (for `threads` per CTA) and the SM count
    template <class K>
    cudaError_t ensure(K kernel, size_t need, int threads, int* ctas_per_sm = nullptr, int* n_sm = nullptr, int carveout = -2) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= PSKD_MAX_DEVICES) return cudaErrorInvalidDevice;
        std::lock_guard<std::mutex> lk(mu);
        if (!done[dev] || need > smem[dev]) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     carveout == -2 ? (int)cudaSharedmemCarveoutMaxShared : carveout);
            if (e != cudaSuccess) return e;
            int c = 0, n = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, kernel, threads, need);
            if (e != cudaSuccess) return e;
            if (c < 1) return cudaErrorInvalidConfiguration;
            e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            smem[dev] = need; ctas[dev] = c; nsm[dev] = n; done[dev] = true;
        }
        if (ctas_per_sm) *ctas_per_sm = ctas[dev];
        if (n_sm) *n_sm = nsm[dev];
        return cudaSuccess;
    }
};

// ---- kernel launchers (pskd_kernels.cu) ----
struct LaunchCtx {
    cudaStream_t stream;
    int n_channels;
    long long Kmax;          // max K over channels
    int Smax, Amax, Pmax_fast;
    unsigned long long S_mask;   // bit s set: some channel with samplesPerBaud == s takes the generic front kernel
    unsigned long long S_mask_fast;   // same for the specialised kernels
    int Amax_fast, Amin_fast;
    int n_seq_channels, n_fast_channels;
    const ChanDesc* d_desc;
    const ChanDesc* h_desc;  // the same descriptors on the host (valid during the call)
    ChanState* d_state;
    float* d_ring;
    float2* d_sel;           // scratch: timing-selected sample per symbol
    float* d_theta;          // scratch: M-th power angle per symbol
    float* d_phase_tmp;      // scratch phase when the caller passes no phase buffer
    float* out_soft; int16_t* out_bits; float* out_phase; int16_t* out_sidx;
    uint8_t* out_hard;       // optional packed hard symbols (one byte per symbol), or null
    double sri_xdelta;
    DevCounters* d_counters;
    unsigned long long* launches;   // host counter
    Profiler* prof;
    // time-parallel chain (all null / 0 when unused)
    const TpItem* tp_head_items; int tp_n_head;      // sequential heads: packets [0, pkt0) of every TP channel
    const TpItem* tp_items; int tp_n_items;          // one item per time-parallel packet
    const TpChan* tp_chans; int tp_n_chans; int tp_n_slots;
    TpPacket* tp_pkts; TpEnd* tp_ends; float* tp_end_ring; float* tp_start_ring; int tp_ring_stride;
    int* tp_fail; int* tp_slot_fail; int* tp_slot_run; int* tp_any_rerun;
    int tp_n_chans_fzs;      // how many of the time-parallel channels run through k_fzs_cb (the rest through k_chain_par)
    // staged path through the fused kernel's stages (CH_FZS channels)
    int n_fzs_channels, Pmax_fzs;
    unsigned long long S_mask_fzs;
    long long Kmax_fzs;
    int* d_fzs_ticket;       // zeroed ticket counters for this call's k_fzs_* launches
    int* fzs_ticket_next;    // host-side index of the next unused counter
    int fzs_ticket_cap;
};

// one launch of the fused kernel: the CH_FUSED channels of one samplesPerBaud value
struct FusedLaunch {
    int S;
    const int* d_list; int n_list;     // channel indices (relative to LaunchCtx::d_desc)
    const int* h_list;                 // the same list on the host
    int units_per_channel;             // ceil(max n_pkts / pkts_per_unit)
    int pkts_per_unit;
    int parts_per_pkt;                 // > 1 (only with pkts_per_unit == 1): every packet is cut into this many units
    int Amax, Pmax;                    // over the listed channels (sizes the shared-memory regions)
    int* d_ticket;                     // one int, zero before the launch
    int* d_done;                       // [n_channels] zero before the launch (indexed like d_desc)
    double grid_share;                 // 0 / 1: the whole GPU; else the fraction of the resident CTAs this launch takes (the
                                       // launches of a mixed bank's samples-per-symbol classes run side by side)
};
// algorithmic bytes of a channel's stages (8 N in; per symbol 8 soft + 4 phase + 2 sampleIndex + 2 b bits)
inline double alg_bytes_front(const ChanDesc& d) { return 8.0 * (double)d.n_in + 2.0 * (double)d.K; }
inline double alg_bytes_chain(const ChanDesc& d) { return 4.0 * (double)d.K; }
inline double alg_bytes_back(const ChanDesc& d) { return (8.0 + 2.0 * d.bpb) * (double)d.K; }
bool fused_supports(int S, int A, int P);
cudaError_t launch_fused(const LaunchCtx& c, const FusedLaunch& f);

cudaError_t launch_fzs_front(const LaunchCtx& c);
cudaError_t launch_fzs_cb(const LaunchCtx& c, const TpCtl& tp, int n_units, double alg_bytes = 0.0);
bool fzs_supports(int S, int A, int P);

cudaError_t launch_front(const LaunchCtx& c);
cudaError_t launch_chain_seq(const LaunchCtx& c);
cudaError_t launch_chain_par(const LaunchCtx& c);
cudaError_t launch_back(const LaunchCtx& c);
cudaError_t launch_finish(const LaunchCtx& c);

}  // namespace pskd
