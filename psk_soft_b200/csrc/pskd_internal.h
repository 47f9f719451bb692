// pskd_internal.h -- structures shared by the host runtime (pskd_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
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
       CH_TP = 64 /* staged path: the phase chain of this channel runs time-parallel over its packets (TpCtl) */ };
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
    unsigned long long wraps, spec_chunks, spec_misses, seq_channels, tp_packets;
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
    int src;               // kind 2: end record to start from; kind 1: end record holding the channel's fit constants
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
};
struct TpChan {            // per time-parallel channel
    int ch, pkt0, n_pkts;  // first time-parallel packet, packets in the call
    int first_item;        // index of the channel's first TpItem (the sequential head); items follow in packet order
    int first_slot;        // index of packet pkt0 in the TpPacket array
    int pad;
};
struct TpCtl {
    const TpItem* items; int n_items;
    const TpChan* chans; int n_chans;
    TpPacket* pkts;
    TpEnd* ends; float* end_ring; float* start_ring; int ring_stride;   // per item: ring after the last epilogue / ring the item started from
    int* fail;             // [n_channels] set by k_tp_check when a hand-over could not be proven
    int fallback;          // launch flag: process only channels with fail[ch] != 0, from state[ch]
};

// ---- optional per-kernel event timing --------------------------------------------------------
enum KernelId { KID_FRONT = 0, KID_CHAIN_SEQ, KID_CHAIN_PAR, KID_BACK_PAR, KID_CHAIN_EXACT, KID_BACK, KID_FINISH, KID_FUSED, KID_TP, KID_COUNT };
struct Profiler {
    bool enabled = false;
    struct Pair { cudaEvent_t a, b; int kid; };
    Pair* pending = nullptr; int n_pending = 0, cap_pending = 0;
    cudaEvent_t* pool = nullptr; int n_pool = 0, cap_pool = 0;
    double ms[KID_COUNT] = {0}; unsigned long long launches[KID_COUNT] = {0};
    cudaEvent_t get();
    void begin(int kid, cudaStream_t s);
    void end(cudaStream_t s);
    void drain();          // caller has synchronised the stream
    void destroy();
};
const char* kernel_name(int kid);

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
    ChanState* d_state;
    float* d_ring;
    float2* d_sel;           // scratch: timing-selected sample per symbol
    float* d_theta;          // scratch: M-th power angle per symbol
    float* d_phase_tmp;      // scratch phase when the caller passes no phase buffer
    float* out_soft; int16_t* out_bits; float* out_phase; int16_t* out_sidx;
    double sri_xdelta;
    DevCounters* d_counters;
    unsigned long long* launches;   // host counter
    Profiler* prof;
    // time-parallel chain (all null / 0 when unused)
    const TpItem* tp_head_items; int tp_n_head;      // sequential heads: packets [0, pkt0) of every TP channel
    const TpItem* tp_items; int tp_n_items;          // one item per time-parallel packet
    const TpChan* tp_chans; int tp_n_chans; int tp_n_slots;
    TpPacket* tp_pkts; TpEnd* tp_ends; float* tp_end_ring; float* tp_start_ring; int tp_ring_stride;
    int* tp_fail;
};

// one launch of the fused kernel: the CH_FUSED channels of one samplesPerBaud value
struct FusedLaunch {
    int S;
    const int* d_list; int n_list;     // channel indices (relative to LaunchCtx::d_desc)
    int units_per_channel;             // ceil(max n_pkts / pkts_per_unit)
    int pkts_per_unit;
    int parts_per_pkt;                 // > 1 (only with pkts_per_unit == 1): every packet is cut into this many units
    int Amax, Pmax;                    // over the listed channels (sizes the shared-memory regions)
    int* d_ticket;                     // one int, zero before the launch
    int* d_done;                       // [n_channels] zero before the launch (indexed like d_desc)
};
bool fused_supports(int S, int A, int P);
cudaError_t launch_fused(const LaunchCtx& c, const FusedLaunch& f);

cudaError_t launch_front(const LaunchCtx& c);
cudaError_t launch_chain_seq(const LaunchCtx& c);
cudaError_t launch_chain_par(const LaunchCtx& c);
cudaError_t launch_back(const LaunchCtx& c);
cudaError_t launch_finish(const LaunchCtx& c);

}  // namespace pskd
