// pskd_api.cu -- host runtime behind the C ABI of include/pskd.h.
//
// Plays the role of the reference's packet prologue/epilogue (cpp/psk_soft.cpp:346-427,
// 605-618) for a whole bank of channels: latches properties per call, applies the reset /
// listener logic (:353-372, :638-651), works out how many symbols every channel emits and where
// the emulated BULKIO packet boundaries fall, then launches the kernels.  No CPU demodulation
// exists here: without a CUDA device every entry point returns PSKD_ERR_CUDA.
#include "../../include/pskd.h"
#include "pskd_internal.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace pskd;

static constexpr int MAX_SLABS = 64, RING = 3;
static constexpr int FUSED_TICKETS = 4 * MAX_SLABS, FZS_TICKETS = 16 * MAX_SLABS;
static constexpr long long STALL_CAP = 32768;    // samples kept of a stalled (over-full) timing window; > every admissible numAvg*samplesPerBaud

static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return fail(_e == cudaErrorMemoryAllocation ? PSKD_ERR_NOMEM : PSKD_ERR_CUDA,     \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = n + n / 8 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc((void**)&p, n * sizeof(T)); want = n; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct ChanHost {
    pskd_props props;          // as configured (live)
    pskd_props latched;        // as used by the last process call
    long long tail_len = 0;    // samples of the window held on the device between calls (= samples.size() unless stalled)
    long long win_size = 0;    // samples.size() between calls (grows without bound while the component is stalled)
    bool resetNumSymbols = true, resetPhaseAvg = true, resetSamplesPerBaud = true;   // cpp/psk_soft.cpp:191-193
    size_t symbolEnergySize = 10;   // symbolEnergy.size() (cpp/psk_soft.cpp:189), for the listener at :640
    bool first_packet = true;
    int fit_n = 0;             // LinearFit::n currently held in the device state
    long long tail_off = 0;    // element offset of this channel's tail region
    long long tail_cap = 0;
    pskd_sri_out sri{};
    // what the host knows about the device-side LinearFit history (time-parallel plan without a sequential head)
    long long hist_syms = 0;   // symbols pushed into the history since it was last cleared
    double last_xdelta = 0.0;  // SRI.xdelta of the previous call (a change clears the history, cpp/psk_soft.cpp:91-102)
};

struct pskd_bank {
    int device = 0;
    int n_channels = 0;
    cudaStream_t stream = nullptr;
    std::vector<ChanHost> ch;
    // device
    ChanDesc* d_desc = nullptr;
    ChanDesc* h_desc_slot[2] = {nullptr, nullptr};   // pinned, double-buffered so NO_SYNC calls can overlap
    cudaEvent_t desc_ev[2] = {nullptr, nullptr};     // upload of slot i consumed
    int desc_slot = 0;
    ChanDesc* h_desc = nullptr;       // slot in use by the current call
    ChanState* d_state = nullptr;
    DevCounters* d_counters = nullptr;
    float* d_ring = nullptr;  int ring_cap = 0;     // floats per channel (2x the max phaseAvg)
    float2* d_tail[2] = {nullptr, nullptr}; long long tail_total = 0; int tail_cur = 0;
    DevBuf<float2> sel; DevBuf<float> theta; DevBuf<float> phase_tmp; DevBuf<int16_t> sidx_tmp;
    // host-buffer staging
    DevBuf<float> st_in; DevBuf<float> st_soft; DevBuf<float> st_phase; DevBuf<int16_t> st_bits; DevBuf<int16_t> st_sidx; DevBuf<uint8_t> st_hard;
    unsigned long long launches = 0;
    pskd_stats stats{};
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host-buffer mode: H2D / D2H overlap the kernels slab by slab
    // ring of staging slots (host-buffer mode): slot r was filled (ev_h2d), consumed by the kernels (ev_kern), emptied (ev_d2h)
    cudaEvent_t ev_h2d[RING] = {nullptr}, ev_kern[RING] = {nullptr}, ev_d2h[RING] = {nullptr};
    unsigned long long slab_seq = 0;     // slabs issued so far (ring position continues across calls)
    // mixed banks: the fused launches of the samples-per-symbol classes run on separate streams, so that one
    // class's tail (few warps left) overlaps the next class's start; ev_call orders them after this call's uploads
    cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_aux[3] = {nullptr, nullptr, nullptr}, ev_call = nullptr;
    long long slab_bytes = 256LL << 20;  // host-buffer mode: staging per slab (PSKD_SLAB_MB)
    int tp_max_channels = 16384;         // auto: time-parallel chain for launches of at most this many scan-chain channels (PSKD_TP_MAX)
    int chain_mode = 0;        // 0 auto (scan-based where possible), 1 force the sequential chain (PSKD_CHAIN=seq)
    int fused_mode = -1;       // -1 auto (large banks), 0 never, 1 whenever a channel qualifies (PSKD_FUSED)
    int host_slabs = MAX_SLABS; // host-buffer mode: most channel slabs one call is cut into (PSKD_SLABS, 1..MAX_SLABS)
    int fused_min_channels = 2816;   // auto: fusable channels per launch from which the fused kernel beats the time-parallel staged kernels
                                     // (measured, 1M-sample 8-PSK calls: 2048 channels 8.4 vs 7.9 ms, 3072 channels 9.8 vs 11.1 ms; PSKD_FUSED_MIN)
    int* d_list = nullptr;     // fused launch lists (channel indices), one segment per (slab, samplesPerBaud)
    int* h_list_slot[2] = {nullptr, nullptr};
    int* d_done = nullptr;     // [n_channels] units completed per channel in the current call
    int* d_ticket = nullptr;   // [16 slabs x 4 samplesPerBaud values] unit ticket counters of k_fused, then FZS_TICKETS for the k_fzs_* launches
    int fzs_ticket_next = 0;
    int fzs_mode = 1;          // staged path through the fused kernel's stages (k_fzs_front + k_fzs_cb) where a channel qualifies (PSKD_FZS=0: legacy staged kernels)
    int tp_mode = -1;          // time-parallel chain of the staged path: -1 auto (few channels, many packets), 0 never, 1 whenever possible (PSKD_TP)
    DevBuf<TpItem> tp_items; DevBuf<TpChan> tp_chans; DevBuf<TpPacket> tp_pkts; DevBuf<TpEnd> tp_ends;
    DevBuf<float> tp_end_ring, tp_start_ring; DevBuf<int> tp_fail, tp_slot_flags;
    Profiler prof;
};

static void default_props(pskd_props* p) {
    p->samplesPerBaud = 10; p->numAvg = 100; p->constelationSize = 4; p->phaseAvg = 50;
    p->differentialDecoding = 0; p->resetState = 0;
}

static int check_props(const pskd_props& p) {
    if (p.samplesPerBaud < 2) return fail(PSKD_ERR_UNSUPPORTED, "samplesPerBaud=%u: the GPU path needs >= 2 (the reference's sps==1 branch only emits with numAvg==0)", p.samplesPerBaud);
    if (p.samplesPerBaud > 63) return fail(PSKD_ERR_UNSUPPORTED, "samplesPerBaud=%u > 63", p.samplesPerBaud);
    if ((size_t)(512 + p.numAvg) * (p.samplesPerBaud | 1) * 8 + 16384 > 220 * 1024)
        return fail(PSKD_ERR_UNSUPPORTED, "numAvg=%u x samplesPerBaud=%u: the timing window does not fit the shared-memory tile", p.numAvg, p.samplesPerBaud);
    if (p.numAvg < 1) return fail(PSKD_ERR_UNSUPPORTED, "numAvg=0 never emits a symbol in the reference");
    if ((unsigned long long)p.numAvg * p.samplesPerBaud > (1ull << 22)) return fail(PSKD_ERR_UNSUPPORTED, "numAvg*samplesPerBaud too large");
    if (p.phaseAvg < 1) return fail(PSKD_ERR_UNSUPPORTED, "phaseAvg=0 is undefined behaviour in the reference (front() of an empty deque)");
    if (p.constelationSize < 1) return fail(PSKD_ERR_UNSUPPORTED, "constelationSize=0");
    return PSKD_OK;
}

static int bpb_of(int M) { return M == 2 ? 1 : M == 4 ? 2 : M == 8 ? 3 : 0; }

// (re)allocate the carried-tail regions so every channel can hold S*A samples (+ one symbol of slack)
static int ensure_tails(pskd_bank* b, const std::vector<long long>* min_cap = nullptr) {
    bool grow = false;
    for (int i = 0; i < b->n_channels; i++) {
        auto& c = b->ch[i];
        long long need = (long long)c.props.samplesPerBaud * c.props.numAvg + c.props.samplesPerBaud;
        if (min_cap) need = std::max(need, (*min_cap)[i]);
        if (need > c.tail_cap) grow = true;
    }
    if (!grow) return PSKD_OK;
    std::vector<long long> new_off(b->n_channels), new_cap(b->n_channels);
    long long total = 0;
    for (int i = 0; i < b->n_channels; i++) {
        auto& c = b->ch[i];
        long long need = (long long)c.props.samplesPerBaud * c.props.numAvg + c.props.samplesPerBaud;
        if (min_cap) need = std::max(need, (*min_cap)[i]);
        new_cap[i] = std::max(need, c.tail_cap);
        new_off[i] = total;
        total += (new_cap[i] + 1) & ~1LL;     // keep 16-byte alignment of every region
    }
    float2* nt[2] = {nullptr, nullptr};
    CUDA_TRY(cudaMalloc((void**)&nt[0], std::max<long long>(total, 1) * sizeof(float2)));
    {
        cudaError_t e1 = cudaMalloc((void**)&nt[1], std::max<long long>(total, 1) * sizeof(float2));
        if (e1 != cudaSuccess) { cudaFree(nt[0]); CUDA_TRY(e1); }
    }
    if (b->d_tail[0]) {
        cudaError_t e1 = cudaSuccess;
        for (int i = 0; i < b->n_channels && e1 == cudaSuccess; i++) {
            auto& c = b->ch[i];
            if (c.tail_len > 0)
                e1 = cudaMemcpyAsync(nt[b->tail_cur] + new_off[i], b->d_tail[b->tail_cur] + c.tail_off,
                                     c.tail_len * sizeof(float2), cudaMemcpyDeviceToDevice, b->stream);
        }
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(b->stream);
        if (e1 != cudaSuccess) { cudaFree(nt[0]); cudaFree(nt[1]); CUDA_TRY(e1); }
        cudaFree(b->d_tail[0]); cudaFree(b->d_tail[1]);
    }
    b->d_tail[0] = nt[0]; b->d_tail[1] = nt[1]; b->tail_total = total;
    for (int i = 0; i < b->n_channels; i++) { b->ch[i].tail_off = new_off[i]; b->ch[i].tail_cap = new_cap[i]; }
    return PSKD_OK;
}

static int ensure_rings(pskd_bank* b) {
    int maxP = 1;
    for (auto& c : b->ch) maxP = std::max<int>(maxP, c.props.phaseAvg);
    int need = 2 * maxP;                       // second half is the repack spare (GlobalRing::repack)
    if (need <= b->ring_cap) return PSKD_OK;
    if ((long long)need * b->n_channels > 0x7fffffffLL)     // ChanDesc::ring_off and the kernels index the rings with 32-bit offsets
        return fail(PSKD_ERR_UNSUPPORTED, "phaseAvg=%d x %d channels: the history rings exceed 2^31 floats", maxP, b->n_channels);
    float* nr = nullptr;
    CUDA_TRY(cudaMalloc((void**)&nr, (size_t)need * b->n_channels * sizeof(float)));
    CUDA_TRY(cudaMemsetAsync(nr, 0, (size_t)need * b->n_channels * sizeof(float), b->stream));
    if (b->d_ring) {
        CUDA_TRY(cudaMemcpy2DAsync(nr, (size_t)need * sizeof(float), b->d_ring, (size_t)b->ring_cap * sizeof(float),
                                   (size_t)b->ring_cap * sizeof(float), b->n_channels, cudaMemcpyDeviceToDevice, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        cudaFree(b->d_ring);
    }
    b->d_ring = nr; b->ring_cap = need;
    return PSKD_OK;
}

extern "C" {

int pskd_abi_version(void) { return PSKD_ABI_VERSION; }
int pskd_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}
const char* pskd_last_error(void) { return g_last_error.c_str(); }
void pskd_default_props(pskd_props* p) { if (p) default_props(p); }

int pskd_create(pskd_handle* out, int device, int n_channels, const pskd_props* props) {
    if (!out || n_channels < 1) return fail(PSKD_ERR_ARG, "pskd_create: bad arguments");
    *out = nullptr;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(PSKD_ERR_CUDA, "pskd_create: CUDA device %d not present (%d devices); there is no CPU path", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    pskd_bank* b = new (std::nothrow) pskd_bank();
    if (!b) return fail(PSKD_ERR_NOMEM, "out of host memory");
    b->device = device; b->n_channels = n_channels;
    b->ch.resize(n_channels);
    for (int i = 0; i < n_channels; i++) {
        if (props) b->ch[i].props = props[i]; else default_props(&b->ch[i].props);
        int rc = check_props(b->ch[i].props);
        if (rc != PSKD_OK) { delete b; return rc; }
        b->ch[i].latched = b->ch[i].props;
        b->ch[i].fit_n = b->ch[i].props.phaseAvg;
    }
    if (const char* e = getenv("PSKD_CHAIN")) b->chain_mode = (strcmp(e, "seq") == 0) ? 1 : 0;
    if (const char* e = getenv("PSKD_FUSED")) b->fused_mode = (strcmp(e, "auto") == 0) ? -1 : atoi(e) != 0;
    if (const char* e = getenv("PSKD_FUSED_MIN")) b->fused_min_channels = std::max(1, atoi(e));
    if (const char* e = getenv("PSKD_SLABS")) b->host_slabs = std::min(MAX_SLABS, std::max(1, atoi(e)));
    if (const char* e = getenv("PSKD_SLAB_MB")) b->slab_bytes = (long long)std::max(1, atoi(e)) << 20;
    if (const char* e = getenv("PSKD_TP_MAX")) b->tp_max_channels = std::max(1, atoi(e));
    if (const char* e = getenv("PSKD_TP")) b->tp_mode = (strcmp(e, "auto") == 0) ? -1 : atoi(e) != 0;
    if (const char* e = getenv("PSKD_FZS")) b->fzs_mode = atoi(e) != 0;
    cudaError_t e;
#define CT(expr) do { e = (expr); if (e != cudaSuccess) { int rc = fail(e == cudaErrorMemoryAllocation ? PSKD_ERR_NOMEM : PSKD_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e)); pskd_destroy(b); return rc; } } while (0)
    CT(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CT(cudaStreamCreateWithFlags(&b->copy_in, cudaStreamNonBlocking));
    CT(cudaStreamCreateWithFlags(&b->copy_out, cudaStreamNonBlocking));
    CT(cudaEventCreateWithFlags(&b->ev_call, cudaEventDisableTiming));
    for (int i = 0; i < 3; i++) {
        CT(cudaStreamCreateWithFlags(&b->aux_stream[i], cudaStreamNonBlocking));
        CT(cudaEventCreateWithFlags(&b->ev_aux[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < RING; i++) {
        CT(cudaEventCreateWithFlags(&b->ev_h2d[i], cudaEventDisableTiming));
        CT(cudaEventCreateWithFlags(&b->ev_kern[i], cudaEventDisableTiming));
        CT(cudaEventCreateWithFlags(&b->ev_d2h[i], cudaEventDisableTiming));
    }
    CT(cudaMalloc((void**)&b->d_desc, sizeof(ChanDesc) * n_channels));
    for (int i = 0; i < 2; i++) {
        CT(cudaMallocHost((void**)&b->h_desc_slot[i], sizeof(ChanDesc) * n_channels));
        CT(cudaMallocHost((void**)&b->h_list_slot[i], sizeof(int) * n_channels));
        CT(cudaEventCreateWithFlags(&b->desc_ev[i], cudaEventDisableTiming));
    }
    CT(cudaMalloc((void**)&b->d_list, sizeof(int) * n_channels));
    CT(cudaMalloc((void**)&b->d_done, sizeof(int) * n_channels));
    CT(cudaMalloc((void**)&b->d_ticket, sizeof(int) * (FUSED_TICKETS + FZS_TICKETS)));
    b->h_desc = b->h_desc_slot[0];
    CT(cudaMalloc((void**)&b->d_state, sizeof(ChanState) * n_channels));
    CT(cudaMalloc((void**)&b->d_counters, sizeof(DevCounters)));
    CT(cudaMemsetAsync(b->d_counters, 0, sizeof(DevCounters), b->stream));
    {   // initial state: LinearFit(phaseAvg, 1.0), phaseEstimate 0, sampleRate 1.0 (cpp/psk_soft.cpp:35-46,187-199)
        std::vector<ChanState> init(n_channels);
        for (int i = 0; i < n_channels; i++) {
            ChanState& s = init[i];
            memset(&s, 0, sizeof(s));
            s.fit.n = b->ch[i].props.phaseAvg;
            s.fit.xdelta = 1.0f; s.fit.denominator = 1.0f;
            s.est = 0.0f; s.sampleRate = 1.0f; s.last = make_float2(0.f, 0.f);
        }
        CT(cudaMemcpyAsync(b->d_state, init.data(), sizeof(ChanState) * n_channels, cudaMemcpyHostToDevice, b->stream));
        CT(cudaStreamSynchronize(b->stream));
    }
#undef CT
    int rc = ensure_tails(b);
    if (rc == PSKD_OK) rc = ensure_rings(b);
    if (rc != PSKD_OK) { pskd_destroy(b); return rc; }
    *out = b;
    return PSKD_OK;
}

int pskd_destroy(pskd_handle b) {
    if (!b) return PSKD_OK;
    cudaSetDevice(b->device);
    if (b->copy_in) cudaStreamSynchronize(b->copy_in);
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->copy_out) cudaStreamSynchronize(b->copy_out);
    cudaFree(b->d_desc);
    for (int i = 0; i < 2; i++) { if (b->h_desc_slot[i]) cudaFreeHost(b->h_desc_slot[i]); if (b->h_list_slot[i]) cudaFreeHost(b->h_list_slot[i]); if (b->desc_ev[i]) cudaEventDestroy(b->desc_ev[i]); }
    cudaFree(b->d_list); cudaFree(b->d_done); cudaFree(b->d_ticket);
    cudaFree(b->d_state); cudaFree(b->d_counters); cudaFree(b->d_ring);
    cudaFree(b->d_tail[0]); cudaFree(b->d_tail[1]);
    b->sel.release(); b->theta.release(); b->phase_tmp.release(); b->sidx_tmp.release();
    b->st_in.release(); b->st_soft.release(); b->st_phase.release(); b->st_bits.release(); b->st_sidx.release(); b->st_hard.release();
    b->tp_items.release(); b->tp_chans.release(); b->tp_pkts.release(); b->tp_ends.release();
    b->tp_end_ring.release(); b->tp_start_ring.release(); b->tp_fail.release(); b->tp_slot_flags.release();
    b->prof.destroy();
    for (int i = 0; i < RING; i++) for (cudaEvent_t e : {b->ev_h2d[i], b->ev_kern[i], b->ev_d2h[i]}) if (e) cudaEventDestroy(e);
    for (int i = 0; i < 3; i++) {
        if (b->aux_stream[i]) { cudaStreamSynchronize(b->aux_stream[i]); cudaStreamDestroy(b->aux_stream[i]); }
        if (b->ev_aux[i]) cudaEventDestroy(b->ev_aux[i]);
    }
    if (b->ev_call) cudaEventDestroy(b->ev_call);
    if (b->copy_in) cudaStreamDestroy(b->copy_in);
    if (b->copy_out) cudaStreamDestroy(b->copy_out);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
    return PSKD_OK;
}

int pskd_set_props(pskd_handle b, int ch, const pskd_props* p) {
    if (!b || !p || ch < -1 || ch >= b->n_channels) return fail(PSKD_ERR_ARG, "pskd_set_props: bad arguments");
    int rc = check_props(*p);
    if (rc != PSKD_OK) return rc;
    int lo = ch < 0 ? 0 : ch, hi = ch < 0 ? b->n_channels : ch + 1;
    for (int i = lo; i < hi; i++) {
        ChanHost& c = b->ch[i];
        // change listeners fire only when the value changed (cpp/psk_soft.cpp:638-651)
        if (p->samplesPerBaud != c.props.samplesPerBaud)
            c.resetSamplesPerBaud = (p->samplesPerBaud != c.symbolEnergySize);         // :640
        if (p->constelationSize != c.props.constelationSize) c.resetNumSymbols = true; // :645
        if (p->phaseAvg != c.props.phaseAvg) c.resetPhaseAvg = true;                   // :650
        c.props = *p;
    }
    return PSKD_OK;
}

int pskd_get_props(pskd_handle b, int ch, pskd_props* p) {
    if (!b || !p || ch < 0 || ch >= b->n_channels) return fail(PSKD_ERR_ARG, "pskd_get_props: bad arguments");
    *p = b->ch[ch].props;
    return PSKD_OK;
}

size_t pskd_max_symbols(pskd_handle b, int ch, size_t n_complex) {
    if (!b || ch < 0 || ch >= b->n_channels) return 0;
    const ChanHost& c = b->ch[ch];
    return (size_t)((n_complex + (size_t)c.tail_len) / std::max<int>(1, c.props.samplesPerBaud)) + 1;
}

void* pskd_stream(pskd_handle b) { return b ? (void*)b->stream : nullptr; }
uint64_t pskd_launch_count(pskd_handle b) { return b ? b->launches : 0; }

int pskd_sync(pskd_handle b) {
    if (!b) return fail(PSKD_ERR_ARG, "null handle");
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaStreamSynchronize(b->copy_in));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->copy_out));       // host-buffer calls with PSKD_FLAG_NO_SYNC: the outputs are on the host now
    return PSKD_OK;
}

int pskd_profile_enable(pskd_handle b, int on) {
    if (!b) return fail(PSKD_ERR_ARG, "null handle");
    b->prof.enabled = on != 0;
    return PSKD_OK;
}

int pskd_profile_read(pskd_handle b, pskd_kernel_time* out, int cap, int* n, int reset) {
    if (!b || !n) return fail(PSKD_ERR_ARG, "pskd_profile_read: bad arguments");
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->prof.drain();
    int k = 0;
    for (int i = 0; i < KID_COUNT; i++) {
        if (!b->prof.launches[i]) continue;
        if (out && k < cap) {
            memset(&out[k], 0, sizeof(out[k]));
            strncpy(out[k].name, kernel_name(i), sizeof(out[k].name) - 1);
            out[k].ms_total = b->prof.ms[i]; out[k].launches = b->prof.launches[i]; out[k].alg_bytes = b->prof.bytes[i];
        }
        k++;
    }
    *n = k;
    if (reset) for (int i = 0; i < KID_COUNT; i++) { b->prof.ms[i] = 0; b->prof.launches[i] = 0; b->prof.bytes[i] = 0; }
    return PSKD_OK;
}

int pskd_get_sri(pskd_handle b, int ch, pskd_sri_out* sri) {
    if (!b || !sri || ch < 0 || ch >= b->n_channels) return fail(PSKD_ERR_ARG, "pskd_get_sri: bad arguments");
    *sri = b->ch[ch].sri;
    return PSKD_OK;
}

int pskd_get_stats(pskd_handle b, pskd_stats* st) {
    if (!b || !st) return fail(PSKD_ERR_ARG, "pskd_get_stats: bad arguments");
    CUDA_TRY(cudaSetDevice(b->device));
    DevCounters c;
    CUDA_TRY(cudaMemcpyAsync(&c, b->d_counters, sizeof(c), cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    *st = b->stats;
    st->wraps = c.wraps; st->spec_chunks = c.spec_chunks; st->spec_misses = c.spec_misses; st->seq_channels = c.seq_channels;
    st->tp_packets = c.tp_packets; st->tp_repaired = c.tp_repaired;
    return PSKD_OK;
}

// ---- checkpoint / resume ----------------------------------------------------------------------
namespace {
struct StateHeader { uint32_t magic, version; int32_t n_channels, ring_cap; uint64_t total_bytes; };
struct StateChan {           // host-side view of one channel (ChanHost without the device offsets)
    pskd_props props, latched;
    long long tail_len;
    int32_t resetNumSymbols, resetPhaseAvg, resetSamplesPerBaud, first_packet, fit_n;
    uint64_t symbolEnergySize;
    pskd_sri_out sri;
    ChanState dev;
};
constexpr uint32_t STATE_MAGIC = 0x444b5350u /* "PSKD" */, STATE_VERSION = 1;
}

size_t pskd_state_size(pskd_handle b) {
    if (!b) return 0;
    size_t n = sizeof(StateHeader) + (size_t)b->n_channels * (sizeof(StateChan) + (size_t)b->ring_cap * sizeof(float));
    for (auto& c : b->ch) n += (size_t)c.tail_len * sizeof(float2);
    return n;
}

int pskd_state_export(pskd_handle b, void* host_buf, size_t cap, size_t* written) {
    if (!b || !host_buf) return fail(PSKD_ERR_ARG, "pskd_state_export: bad arguments");
    const size_t need = pskd_state_size(b);
    if (cap < need) return fail(PSKD_ERR_CAPACITY, "pskd_state_export: %zu bytes needed, %zu given", need, cap);
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    const int nch = b->n_channels;
    std::vector<ChanState> dev(nch);
    CUDA_TRY(cudaMemcpy(dev.data(), b->d_state, sizeof(ChanState) * nch, cudaMemcpyDeviceToHost));
    unsigned char* p = static_cast<unsigned char*>(host_buf);
    StateHeader h{STATE_MAGIC, STATE_VERSION, nch, b->ring_cap, (uint64_t)need};
    memcpy(p, &h, sizeof(h)); p += sizeof(h);
    for (int i = 0; i < nch; i++) {
        const ChanHost& c = b->ch[i];
        StateChan sc;
        memset(&sc, 0, sizeof(sc));
        sc.props = c.props; sc.latched = c.latched; sc.tail_len = c.tail_len;
        sc.resetNumSymbols = c.resetNumSymbols; sc.resetPhaseAvg = c.resetPhaseAvg; sc.resetSamplesPerBaud = c.resetSamplesPerBaud;
        sc.first_packet = c.first_packet; sc.fit_n = c.fit_n; sc.symbolEnergySize = c.symbolEnergySize; sc.sri = c.sri;
        sc.dev = dev[i];
        memcpy(p, &sc, sizeof(sc)); p += sizeof(sc);
    }
    CUDA_TRY(cudaMemcpy(p, b->d_ring, (size_t)nch * b->ring_cap * sizeof(float), cudaMemcpyDeviceToHost));
    p += (size_t)nch * b->ring_cap * sizeof(float);
    for (int i = 0; i < nch; i++) {
        const ChanHost& c = b->ch[i];
        if (c.tail_len > 0) {
            CUDA_TRY(cudaMemcpy(p, b->d_tail[b->tail_cur] + c.tail_off, (size_t)c.tail_len * sizeof(float2), cudaMemcpyDeviceToHost));
            p += (size_t)c.tail_len * sizeof(float2);
        }
    }
    if (written) *written = (size_t)(p - static_cast<unsigned char*>(host_buf));
    return PSKD_OK;
}

int pskd_state_import(pskd_handle b, const void* host_buf, size_t n_bytes) {
    if (!b || !host_buf || n_bytes < sizeof(StateHeader)) return fail(PSKD_ERR_ARG, "pskd_state_import: bad arguments");
    const unsigned char* p0 = static_cast<const unsigned char*>(host_buf);
    StateHeader h;
    memcpy(&h, p0, sizeof(h));
    if (h.magic != STATE_MAGIC || h.version != STATE_VERSION) return fail(PSKD_ERR_ARG, "pskd_state_import: not a pskd state blob (or another version)");
    if (h.n_channels != b->n_channels) return fail(PSKD_ERR_ARG, "pskd_state_import: blob has %d channels, bank has %d", h.n_channels, b->n_channels);
    const int nch = b->n_channels;
    // ---- validate EVERYTHING against the blob's own size before a single byte reaches the bank ----
    if (h.ring_cap < 2 || h.ring_cap > 2 * 65535 || (h.ring_cap & 1)) return fail(PSKD_ERR_ARG, "pskd_state_import: corrupt blob (history ring of %d floats)", h.ring_cap);
    const size_t fixed = sizeof(StateHeader) + (size_t)nch * (sizeof(StateChan) + (size_t)h.ring_cap * sizeof(float));
    if (h.total_bytes > n_bytes || h.total_bytes < fixed) return fail(PSKD_ERR_ARG, "pskd_state_import: truncated blob");
    std::vector<StateChan> sc(nch);
    memcpy(sc.data(), p0 + sizeof(StateHeader), sizeof(StateChan) * nch);
    size_t expect = fixed;
    int maxP = 1;
    for (int i = 0; i < nch; i++) {
        const StateChan& c = sc[i];
        int rc = check_props(c.props);
        if (rc != PSKD_OK) return rc;
        const long long cap = std::max((long long)c.props.samplesPerBaud * c.props.numAvg + c.props.samplesPerBaud, STALL_CAP);
        if (c.tail_len < 0 || c.tail_len > cap) return fail(PSKD_ERR_ARG, "pskd_state_import: channel %d: carried window of %lld samples does not fit its properties", i, c.tail_len);
        const FitState& f = c.dev.fit;
        if (c.fit_n < 1 || c.fit_n > 65535 || f.n != c.fit_n || 2 * f.n > h.ring_cap || f.pts < 0 || f.pts > f.n || f.head < 0 || f.head >= f.n ||
            f.count < 0 || f.count > 1048576)
            return fail(PSKD_ERR_ARG, "pskd_state_import: channel %d: corrupt LinearFit state (n %d pts %d head %d)", i, f.n, f.pts, f.head);
        maxP = std::max(maxP, std::max<int>(c.props.phaseAvg, f.n));
        expect += (size_t)c.tail_len * sizeof(float2);
    }
    if (expect != h.total_bytes) return fail(PSKD_ERR_ARG, "pskd_state_import: blob is %llu bytes, its contents need %zu", (unsigned long long)h.total_bytes, expect);
    if ((long long)nch * std::max(2 * maxP, h.ring_cap) > 0x7fffffffLL) return fail(PSKD_ERR_UNSUPPORTED, "pskd_state_import: history rings exceed 2^31 floats");
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    // ---- make room (capacities only grow; the bank's own state is still untouched and consistent) ----
    const std::vector<ChanHost> saved = b->ch;
    for (int i = 0; i < nch; i++) b->ch[i].props = sc[i].props;       // sizes the tail regions and the rings
    std::vector<long long> tail_need(nch);
    for (int i = 0; i < nch; i++) tail_need[i] = sc[i].tail_len;
    int rc = ensure_tails(b, &tail_need);
    if (rc == PSKD_OK) rc = ensure_rings(b);
    if (rc == PSKD_OK && h.ring_cap > b->ring_cap) {                  // blob written by a bank with a longer history ring
        // grow to the blob's ring: ensure_rings sizes from phaseAvg, so ask for it explicitly
        float* nr = nullptr;
        if (cudaMalloc((void**)&nr, (size_t)h.ring_cap * nch * sizeof(float)) != cudaSuccess) { (void)cudaGetLastError(); rc = fail(PSKD_ERR_NOMEM, "pskd_state_import: out of device memory"); }
        else { cudaFree(b->d_ring); b->d_ring = nr; b->ring_cap = h.ring_cap; }
    }
    if (rc != PSKD_OK) { for (int i = 0; i < nch; i++) b->ch[i].props = saved[i].props; return rc; }
    // ---- commit: device state first (stream-ordered copies from the caller's buffer), then the host view ----
    std::vector<ChanState> dev(nch);
    for (int i = 0; i < nch; i++) dev[i] = sc[i].dev;
    const unsigned char* p = p0 + sizeof(StateHeader) + sizeof(StateChan) * nch;
    cudaError_t e = cudaMemcpy(b->d_state, dev.data(), sizeof(ChanState) * nch, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(b->d_ring, 0, (size_t)b->ring_cap * nch * sizeof(float));
    if (e == cudaSuccess)
        e = cudaMemcpy2D(b->d_ring, (size_t)b->ring_cap * sizeof(float), p, (size_t)h.ring_cap * sizeof(float),
                         (size_t)h.ring_cap * sizeof(float), nch, cudaMemcpyHostToDevice);
    p += (size_t)nch * h.ring_cap * sizeof(float);
    for (int i = 0; i < nch && e == cudaSuccess; i++) {
        const long long tl = sc[i].tail_len;
        if (tl > 0) {
            e = cudaMemcpy(b->d_tail[b->tail_cur] + b->ch[i].tail_off, p, (size_t)tl * sizeof(float2), cudaMemcpyHostToDevice);
            p += (size_t)tl * sizeof(float2);
        }
    }
    if (e != cudaSuccess) {
        // a CUDA failure half-way leaves the device state undefined: make the bank say so on its next use
        for (int i = 0; i < nch; i++) { b->ch[i] = saved[i]; b->ch[i].props.resetState = 1; b->ch[i].tail_len = 0; }
        return fail(PSKD_ERR_CUDA, "pskd_state_import: %s; the bank was reset", cudaGetErrorString(e));
    }
    for (int i = 0; i < nch; i++) {
        ChanHost& c = b->ch[i];
        c.props = sc[i].props; c.latched = sc[i].latched;
        c.resetNumSymbols = sc[i].resetNumSymbols != 0; c.resetPhaseAvg = sc[i].resetPhaseAvg != 0;
        c.resetSamplesPerBaud = sc[i].resetSamplesPerBaud != 0; c.first_packet = sc[i].first_packet != 0;
        c.fit_n = sc[i].fit_n; c.symbolEnergySize = (size_t)sc[i].symbolEnergySize; c.sri = sc[i].sri;
        c.tail_len = sc[i].tail_len; c.win_size = sc[i].tail_len;
        c.hist_syms = 0; c.last_xdelta = 0.0;       // not part of the blob: the next call plans conservatively
    }
    return PSKD_OK;
}

int pskd_process(pskd_handle b, const pskd_input* in, pskd_output* out) {
    if (!b || !in || !out) return fail(PSKD_ERR_ARG, "pskd_process: null argument");
    if (!in->iq && (in->n_complex || in->n_complex_all)) return fail(PSKD_ERR_ARG, "pskd_process: iq is NULL");
    const int nch = b->n_channels;
    const bool host_bufs = (in->flags & PSKD_FLAG_HOST_BUFFERS) != 0;
    CUDA_TRY(cudaSetDevice(b->device));

    // ---- packet prologue, bank-wide (cpp/psk_soft.cpp:353-372) -------------------------------
    if (in->flags & PSKD_FLAG_QUEUE_FLUSHED)
        for (auto& c : b->ch) c.props.resetState = 1;                                  // :353-357
    if (in->sri_mode != 1) {                                                           // :359-363
        if (out->n_symbols) std::fill(out->n_symbols, out->n_symbols + nch, (size_t)0);
        if (out->n_bits) std::fill(out->n_bits, out->n_bits + nch, (size_t)0);
        return PSKD_IGNORED_REAL_DATA;
    }
    if (!(in->sri_xdelta > 0.0)) return fail(PSKD_ERR_ARG, "pskd_process: sri_xdelta must be > 0");
    for (auto& c : b->ch) {
        if (c.props.resetState) {                                                      // :365-372
            c.resetSamplesPerBaud = c.resetNumSymbols = c.resetPhaseAvg = true;
            c.props.resetState = 0;
        }
    }
    // ---- the timing window at the packet start: resyncEnergy and the stall (cpp/psk_soft.cpp:380-383, 408-412,
    // 457, 619-636).  resyncEnergy runs when asked for, or whenever the window is not full (every packet in
    // steady state); its observable effects are the truncation of an over-full window to its OLDEST numDataPts
    // samples and symbolEnergy.size().  A window that is full or over-full at the packet start (after numAvg *
    // samplesPerBaud shrank) never emits again -- `samples.size() == numDataPts` (:457) cannot become true while the
    // deque only grows -- until the window length grows past the deque size: the component keeps consuming packets
    // and emits nothing.  That stall is emulated: the deque size is tracked, its first STALL_CAP samples are kept
    // (they are what a later, longer window would contain; a deque beyond every admissible window length can never
    // recover, so nothing more needs to be stored).
    std::vector<unsigned char> stalled(nch, 0), resync_asked(nch, 0);
    {
        std::vector<long long> cap_need(nch, 0);
        bool any_stall = false;
        for (int i = 0; i < nch; i++) {
            ChanHost& c = b->ch[i];
            const long long n_in = (long long)(in->n_complex ? in->n_complex[i] : in->n_complex_all);
            const long long numDataPts = (long long)c.props.samplesPerBaud * c.props.numAvg;
            resync_asked[i] = c.resetSamplesPerBaud;
            if (c.resetSamplesPerBaud || numDataPts > c.win_size) {
                if (c.win_size > numDataPts) { c.win_size = numDataPts; c.tail_len = std::min(c.tail_len, numDataPts); }
                c.symbolEnergySize = c.props.samplesPerBaud;
                c.resetSamplesPerBaud = false;
            }
            if (c.win_size >= numDataPts && n_in > 0) {
                stalled[i] = 1; any_stall = true;
                if (c.tail_len == c.win_size) cap_need[i] = std::min(c.tail_len + n_in, STALL_CAP);
            }
        }
        int rc0 = ensure_tails(b, any_stall ? &cap_need : nullptr);
        if (rc0 != PSKD_OK) return rc0;
    }
    int rc = ensure_rings(b);
    if (rc != PSKD_OK) return rc;

    // ---- per-channel geometry ---------------------------------------------------------------
    b->desc_slot ^= 1;
    b->h_desc = b->h_desc_slot[b->desc_slot];
    CUDA_TRY(cudaEventSynchronize(b->desc_ev[b->desc_slot]));   // previous upload from this slot is done
    long long Kmax = 0, scr_total = 0, nmax = 0;
    bool any_nobits = false;
    const int next_tail = b->tail_cur ^ 1;
    std::vector<unsigned char> fusable(nch, 0), fzsable(nch, 0);
    for (int i = 0; i < nch; i++) {
        ChanHost& c = b->ch[i];
        const long long n_in = (long long)(in->n_complex ? in->n_complex[i] : in->n_complex_all);
        if ((size_t)n_in > in->iq_stride && nch > 1) return fail(PSKD_ERR_ARG, "channel %d: n_complex > iq_stride", i);
        const int S = c.props.samplesPerBaud, A = (int)c.props.numAvg, M = c.props.constelationSize, P = c.props.phaseAvg;
        ChanDesc& d = b->h_desc[i];
        d.in = host_bufs ? nullptr : reinterpret_cast<const float2*>(in->iq) + (size_t)i * in->iq_stride;
        d.tail = b->d_tail[b->tail_cur] + c.tail_off;
        d.tail_next = b->d_tail[next_tail] + c.tail_off;
        d.n_in = n_in; d.tail_len = c.tail_len;
        d.S = S; d.A = A; d.M = M; d.P = P; d.D = c.props.differentialDecoding ? 1 : 0; d.bpb = bpb_of(M);
        d.ring_off = i * b->ring_cap;
        d.flags = (c.resetNumSymbols ? CH_RESET_NUMSYMS : 0) | (c.resetPhaseAvg ? CH_RESET_PHASEAVG : 0);
        d.sym_off = (long long)i * (long long)out->sym_stride;
        d.bits_off = (long long)i * (long long)out->bits_stride;
        long long K = 0;
        if (stalled[i]) {
            // consumes the packet(s), emits nothing.  The phase estimator's prologue / epilogue still run, but only
            // they: one emulated packet when a reset is pending (its SRI block only with sriChanged / resetNumSymbols,
            // :393), none otherwise (nothing of the carried state changes)
            d.flags |= CH_STALLED | (((in->flags & PSKD_FLAG_SRI_CHANGED) || resync_asked[i]) ? CH_SRI_CHANGED : 0);
            const bool pending = (d.flags & (CH_RESET_NUMSYMS | CH_RESET_PHASEAVG | CH_SRI_CHANGED)) != 0;
            d.K = 0; d.pkt_len = std::max<long long>(n_in, 1); d.n_pkts = pending ? 1 : 0;
            d.next_tail_len = (c.tail_len == c.win_size) ? std::min(c.tail_len + n_in, STALL_CAP) : c.tail_len;
        } else {
            const long long total = c.tail_len + n_in;
            K = total / S - A + 1;
            if (K < 0) K = 0;
            const long long pkt = in->packet_len ? (long long)in->packet_len : std::max<long long>(n_in, 1);
            const long long npk = n_in > 0 ? (n_in + pkt - 1) / pkt : 0;
            if (npk > 0x7fffffffLL) return fail(PSKD_ERR_ARG, "too many packets");
            d.K = K; d.pkt_len = pkt; d.n_pkts = (int)npk;
            d.next_tail_len = total - K * S;
        }
        if (d.next_tail_len > c.tail_cap) return fail(PSKD_ERR_UNSUPPORTED, "channel %d: carried window exceeds its capacity", i);
        // the scan-based chain needs the history staged in shared memory and an unchanged window length
        const bool fast = b->chain_mode == 0 && P <= CHAIN_PAR_PMAX && c.fit_n == P && !stalled[i];
        if (fast) d.flags |= CH_FAST;
        fusable[i] = fast && b->fused_mode != 0 && fused_supports(S, A, P) && K < (1LL << 30);
        fzsable[i] = fast && b->fzs_mode != 0 && fzs_supports(S, A, P) && K < (1LL << 30);
        if (d.bpb == 0) any_nobits = true;
        if ((size_t)K > out->sym_stride && (out->soft || out->phase || out->sample_index || out->hard))
            return fail(PSKD_ERR_CAPACITY, "channel %d emits %lld symbols > sym_stride %zu", i, K, out->sym_stride);
        if (out->bits && (size_t)(K * d.bpb) > out->bits_stride)
            return fail(PSKD_ERR_CAPACITY, "channel %d emits %lld bits > bits_stride %zu", i, K * d.bpb, out->bits_stride);
        Kmax = std::max(Kmax, K); nmax = std::max(nmax, n_in);
    }

    // ---- slabs.  Device buffers: one slab = the whole bank.  Host buffers: slabs of consecutive channels flow
    // H2D (copy_in) -> kernels (stream) -> D2H (copy_out) through a RING of staging buffers, so the device memory
    // the call needs is bounded by the slab size, not the bank size, and the transfers of neighbouring slabs (and of
    // the next call, with PSKD_FLAG_NO_SYNC) overlap the kernels.
    int n_slabs = 1;
    size_t sym_stride = out->sym_stride, bits_stride = out->bits_stride, in_stride = 0;
    if (host_bufs) {
        sym_stride = (size_t)((Kmax + 7) & ~7LL); bits_stride = sym_stride * 3;
        in_stride = (size_t)((nmax + 1) & ~1LL);
        const long long per_ch = std::max<long long>(1, (long long)in_stride * 8 + (long long)sym_stride * 20);
        long long ch_per_slab = std::max<long long>(1, b->slab_bytes / per_ch);
        long long ns = (nch + ch_per_slab - 1) / ch_per_slab;
        ns = std::min<long long>(std::min<long long>(ns, b->host_slabs), nch);
        n_slabs = (int)std::max<long long>(ns, 1);
    }
    auto slab_lo = [&](int s) { return (int)((long long)s * nch / n_slabs); };
    int slab_ch_max = 0;
    for (int s = 0; s < n_slabs; s++) slab_ch_max = std::max(slab_ch_max, slab_lo(s + 1) - slab_lo(s));

    // ---- which channels take the fused kernel (per slab), which the staged kernels --------------
    struct FusedSeg { int slab, S, first, count, Amax, Pmax, max_pkts; long long pkt_len_min; };
    std::vector<FusedSeg> segs;
    int* h_list = b->h_list_slot[b->desc_slot];
    int n_listed = 0;
    static const int fusedS[4] = {8, 9, 10, 16};
    for (int s = 0; s < n_slabs; s++) {
        const int lo = slab_lo(s), hi = slab_lo(s + 1);
        for (int si = 0; si < 4; si++) {
            // the fused kernel needs enough channels of ONE samples-per-symbol class to fill the resident warps (each channel's
            // units run one after the other); smaller classes take the time-parallel staged kernels
            int nf = 0;
            for (int i = lo; i < hi; i++) nf += (fusable[i] && b->h_desc[i].S == fusedS[si]);
            const bool use = nf > 0 && (b->fused_mode == 1 || nf >= b->fused_min_channels);
            if (!use) continue;
            FusedSeg g{s, fusedS[si], n_listed, 0, 1, 1, 0, 1LL << 62};
            for (int i = lo; i < hi; i++) {
                ChanDesc& d = b->h_desc[i];
                if (!fusable[i] || d.S != g.S) continue;
                d.flags |= CH_FUSED;
                h_list[n_listed++] = i - lo;
                g.count++; g.Amax = std::max(g.Amax, d.A); g.Pmax = std::max(g.Pmax, d.P);
                g.max_pkts = std::max(g.max_pkts, d.n_pkts);
                if (d.n_pkts > 0) g.pkt_len_min = std::min(g.pkt_len_min, d.pkt_len);
            }
            if (g.count) segs.push_back(g);
        }
    }
    // ---- per-slab summary of the staged channels + the time-parallel chain plan ---------------------
    // Time-parallel plan (few channels with many packets, whose chain would otherwise run packet after packet on one
    // warp): heads = packets before the first packet that is certain to start with a full history; one item per
    // later packet.  Channel indices in the plan are slab-relative (the kernels of a slab see its channels only);
    // packet slots and end records are numbered across the whole call.
    struct SlabInfo {
        int n_fast = 0, n_seq = 0, n_fzs = 0, Pmax_fast = 1, Pmax_fzs = 1, Smax = 2, Amax = 1, Amax_fast = 1, Amin_fast = 1 << 30;
        unsigned long long S_mask = 0, S_mask_fast = 0, S_mask_fzs = 0;
        long long Kmax = 0, Kmax_fzs = 0;
        int head0 = 0, n_head = 0, item0 = 0, n_item = 0, chan0 = 0, n_chan = 0, n_slot = 0, tp_fzs = 0;
    };
    std::vector<SlabInfo> slabs(n_slabs);
    std::vector<TpItem> tp_heads, tp_items;
    std::vector<TpChan> tp_chans;
    int tp_slots = 0, tp_records = 0, tp_Pmax = 1;
    static const bool headless_ok = !(getenv("PSKD_TP_HEADLESS") && atoi(getenv("PSKD_TP_HEADLESS")) == 0);
    bool any_staged = false;
    for (int s = 0; s < n_slabs; s++) {
        const int lo = slab_lo(s), hi = slab_lo(s + 1);
        SlabInfo& si = slabs[s];
        long long scr_slab = 0;            // scratch is reused slab after slab (their kernels are ordered on one stream)
        for (int i = lo; i < hi; i++) {
            ChanDesc& d = b->h_desc[i];
            if (host_bufs) { d.sym_off = (long long)(i - lo) * (long long)sym_stride; d.bits_off = (long long)(i - lo) * (long long)bits_stride; }
            if (d.flags & CH_FUSED) { d.scr_off = 0; continue; }
            any_staged = true;
            if (d.flags & CH_FAST) si.n_fast++; else si.n_seq++;
            // the chain kernels read the scratch packet by packet: shift the row so that packet 1 (and with it every packet
            // whose symbol count is a multiple of 4) starts 16-byte aligned
            const long long k1 = first_symbol_at(d.pkt_len, d.tail_len, d.S, d.A, d.K);
            d.scr_off = scr_slab + ((4 - (k1 & 3)) & 3);
            scr_slab += ((d.K + 3) & ~3LL) + 4;
            si.Kmax = std::max(si.Kmax, d.K);
            if (fzsable[i]) {              // staged, through the fused kernel's stages
                d.flags |= CH_FZS;
                si.n_fzs++; si.Pmax_fzs = std::max(si.Pmax_fzs, d.P); si.Kmax_fzs = std::max(si.Kmax_fzs, d.K);
                if (d.K > 0) si.S_mask_fzs |= 1ull << d.S;
                continue;
            }
            if (d.flags & CH_FAST) si.Pmax_fast = std::max(si.Pmax_fast, d.P);
            si.Smax = std::max(si.Smax, d.S); si.Amax = std::max(si.Amax, d.A);
            if (d.K > 0) {
                const bool front_fast = (d.S == 8 || d.S == 9 || d.S == 10 || d.S == 16) && d.A <= FRONT_FAST_AMAX;
                if (front_fast) { d.flags |= CH_FRONT_FAST; si.S_mask_fast |= 1ull << d.S; si.Amax_fast = std::max(si.Amax_fast, d.A); si.Amin_fast = std::min(si.Amin_fast, d.A); }
                else si.S_mask |= 1ull << d.S;
            }
        }
        scr_total = std::max(scr_total, scr_slab);
        si.head0 = (int)tp_heads.size(); si.item0 = (int)tp_items.size(); si.chan0 = (int)tp_chans.size();
        if (b->tp_mode != 0 && si.n_fast > 0 && (b->tp_mode == 1 || si.n_fast <= b->tp_max_channels) && in->sri_xdelta != 1.0) {
            for (int i = lo; i < hi; i++) {
                ChanDesc& d = b->h_desc[i];
                const ChanHost& c = b->ch[i];
                if (!(d.flags & CH_FAST) || (d.flags & CH_FUSED) || d.P < 2 || d.K <= 0) continue;
                if (d.pkt_len / d.S - 1 < d.P + 2) continue;                   // every full packet must hold the whole history
                // No sequential head when the carried history is known to be full and nothing clears it at this call's
                // first packet: packet 0 starts from the carried state, every later packet from a synthesised ring.
                const bool headless = headless_ok && c.hist_syms >= d.P && c.last_xdelta == in->sri_xdelta &&
                                      !(d.flags & (CH_RESET_NUMSYMS | CH_RESET_PHASEAVG)) && d.n_pkts >= 2 &&
                                      first_symbol_at(d.pkt_len, d.tail_len, d.S, d.A, d.K) >= d.P + 1;
                int j0 = -1;
                if (headless) j0 = 0;
                else
                    for (int j = 1; j < d.n_pkts; j++)
                        if (first_symbol_at((long long)j * d.pkt_len, d.tail_len, d.S, d.A, d.K) >= d.P) { j0 = j; break; }
                if (j0 < 0 || d.n_pkts - j0 < (b->tp_mode == 1 ? 2 : 4)) continue;
                d.flags |= CH_TP;
                if (d.flags & CH_FZS) si.tp_fzs++;
                const int has_head = headless ? 0 : 1;
                const int cr = i - lo;                                         // slab-relative channel
                tp_chans.push_back(TpChan{cr, j0, d.n_pkts, tp_records, tp_slots, has_head});
                if (has_head) tp_heads.push_back(TpItem{cr, 0, j0, 0, -1, tp_records, -1, 0});
                for (int j = j0; j < d.n_pkts; j++) {
                    const int rec = tp_records + has_head + (j - j0);
                    const int kind = (j == j0) ? (has_head ? 2 : 0) : 1;
                    tp_items.push_back(TpItem{cr, j, j + 1, kind, has_head ? tp_records : -1, rec, tp_slots + (j - j0), 0});
                }
                tp_records += has_head + (d.n_pkts - j0);
                tp_slots += d.n_pkts - j0;
                si.n_slot += d.n_pkts - j0;
                tp_Pmax = std::max(tp_Pmax, d.P);
            }
        }
        si.n_head = (int)tp_heads.size() - si.head0; si.n_item = (int)tp_items.size() - si.item0; si.n_chan = (int)tp_chans.size() - si.chan0;
    }

    // ---- buffers ------------------------------------------------------------------------------
    CUDA_TRY(b->sel.reserve((size_t)scr_total + 4));
    CUDA_TRY(b->theta.reserve((size_t)scr_total + 4));
    float* dev_soft = out->soft; float* dev_phase = out->phase; int16_t* dev_bits = out->bits; int16_t* dev_sidx = out->sample_index;
    uint8_t* dev_hard = out->hard;
    size_t slot_in = 0, slot_sym = 0, slot_bits = 0;          // elements per ring slot
    if (host_bufs) {
        slot_in = 2 * in_stride * slab_ch_max; slot_sym = sym_stride * slab_ch_max; slot_bits = bits_stride * slab_ch_max;
        // growing a staging buffer frees the old one: let every slab in flight (an earlier NO_SYNC call) drain first
        const bool grow = RING * slot_in + 4 > b->st_in.cap || (out->soft && RING * 2 * slot_sym + 4 > b->st_soft.cap) ||
                          (out->phase && RING * slot_sym + 4 > b->st_phase.cap) || (out->bits && RING * slot_bits + 4 > b->st_bits.cap) ||
                          (out->sample_index && RING * slot_sym + 4 > b->st_sidx.cap) || (out->hard && RING * slot_sym + 4 > b->st_hard.cap);
        if (grow) { CUDA_TRY(cudaStreamSynchronize(b->copy_in)); CUDA_TRY(cudaStreamSynchronize(b->stream)); CUDA_TRY(cudaStreamSynchronize(b->copy_out)); }
        CUDA_TRY(b->st_in.reserve(RING * slot_in + 4));
        if (out->soft) CUDA_TRY(b->st_soft.reserve(RING * 2 * slot_sym + 4));
        if (out->phase) CUDA_TRY(b->st_phase.reserve(RING * slot_sym + 4));
        if (out->bits) CUDA_TRY(b->st_bits.reserve(RING * slot_bits + 4));
        if (out->sample_index) CUDA_TRY(b->st_sidx.reserve(RING * slot_sym + 4));
        if (out->hard) CUDA_TRY(b->st_hard.reserve(RING * slot_sym + 4));
    }
    const size_t ostride_all = host_bufs ? slot_sym : sym_stride * (size_t)nch;     // elements of a whole-call temporary
    if (!out->phase && any_staged) { CUDA_TRY(b->phase_tmp.reserve(ostride_all + 4)); }
    if (!out->sample_index) { CUDA_TRY(b->sidx_tmp.reserve(ostride_all + 4)); if (!host_bufs) dev_sidx = b->sidx_tmp.p; }

    // channel descriptors: host mode fills in the ring slot addresses first
    if (host_bufs) {
        for (int s = 0; s < n_slabs; s++) {
            const int lo = slab_lo(s), hi = slab_lo(s + 1);
            const int r = (int)((b->slab_seq + (unsigned long long)s) % RING);
            for (int i = lo; i < hi; i++)
                b->h_desc[i].in = reinterpret_cast<const float2*>(b->st_in.p + (size_t)r * slot_in) + (size_t)(i - lo) * in_stride;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_desc, b->h_desc, sizeof(ChanDesc) * nch, cudaMemcpyHostToDevice, b->stream));
    if (n_listed > 0) {
        CUDA_TRY(cudaMemcpyAsync(b->d_list, h_list, sizeof(int) * n_listed, cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemsetAsync(b->d_done, 0, sizeof(int) * nch, b->stream));
    }
    CUDA_TRY(cudaMemsetAsync(b->d_ticket, 0, sizeof(int) * (FUSED_TICKETS + FZS_TICKETS), b->stream));
    b->fzs_ticket_next = 0;
    CUDA_TRY(cudaEventRecord(b->desc_ev[b->desc_slot], b->stream));
    const int tp_stride = (tp_Pmax + 3) & ~3;
    if (!tp_chans.empty()) {
        CUDA_TRY(b->tp_items.reserve(tp_heads.size() + tp_items.size()));
        CUDA_TRY(b->tp_chans.reserve(tp_chans.size()));
        CUDA_TRY(b->tp_pkts.reserve((size_t)tp_slots));
        CUDA_TRY(b->tp_ends.reserve((size_t)tp_records));
        CUDA_TRY(b->tp_end_ring.reserve((size_t)tp_records * tp_stride));
        CUDA_TRY(b->tp_start_ring.reserve((size_t)tp_records * tp_stride));
        CUDA_TRY(b->tp_fail.reserve((size_t)nch));
        CUDA_TRY(b->tp_slot_flags.reserve(2 * (size_t)tp_slots + 4));
        // pageable sources: the copies are staged before the calls return
        CUDA_TRY(cudaMemcpyAsync(b->tp_items.p, tp_heads.data(), sizeof(TpItem) * tp_heads.size(), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemcpyAsync(b->tp_items.p + tp_heads.size(), tp_items.data(), sizeof(TpItem) * tp_items.size(), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemcpyAsync(b->tp_chans.p, tp_chans.data(), sizeof(TpChan) * tp_chans.size(), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemsetAsync(b->tp_fail.p, 0, sizeof(int) * nch, b->stream));
        CUDA_TRY(cudaMemsetAsync(b->tp_slot_flags.p, 0, sizeof(int) * (2 * (size_t)tp_slots + 4), b->stream));
    }

    // the kernels of one slab of channels [lo, hi)
    auto run_slab = [&](int slab, int lo, int hi) -> int {
        const SlabInfo& si = slabs[slab];
        LaunchCtx Ls{};
        Ls.stream = b->stream; Ls.n_channels = hi - lo;
        Ls.Kmax = si.Kmax; Ls.Smax = si.Smax; Ls.Amax = si.Amax;
        Ls.S_mask = si.S_mask; Ls.S_mask_fast = si.S_mask_fast; Ls.Amax_fast = si.Amax_fast; Ls.Amin_fast = si.Amin_fast;
        Ls.Pmax_fast = si.Pmax_fast; Ls.n_fast_channels = si.n_fast; Ls.n_seq_channels = si.n_seq;
        Ls.d_desc = b->d_desc + lo; Ls.h_desc = b->h_desc + lo; Ls.d_state = b->d_state + lo; Ls.d_ring = b->d_ring;
        Ls.d_sel = b->sel.p; Ls.d_theta = b->theta.p; Ls.d_phase_tmp = b->phase_tmp.p;
        Ls.out_soft = dev_soft; Ls.out_bits = dev_bits; Ls.out_phase = dev_phase; Ls.out_sidx = dev_sidx; Ls.out_hard = dev_hard;
        Ls.sri_xdelta = in->sri_xdelta; Ls.d_counters = b->d_counters; Ls.launches = &b->launches; Ls.prof = &b->prof;
        if (si.n_chan > 0) {
            Ls.tp_head_items = b->tp_items.p + si.head0; Ls.tp_n_head = si.n_head;
            Ls.tp_items = b->tp_items.p + tp_heads.size() + si.item0; Ls.tp_n_items = si.n_item;
            Ls.tp_chans = b->tp_chans.p + si.chan0; Ls.tp_n_chans = si.n_chan; Ls.tp_n_slots = si.n_slot;
            Ls.tp_pkts = b->tp_pkts.p; Ls.tp_ends = b->tp_ends.p; Ls.tp_end_ring = b->tp_end_ring.p; Ls.tp_start_ring = b->tp_start_ring.p;
            Ls.tp_ring_stride = tp_stride; Ls.tp_fail = b->tp_fail.p + lo;
            Ls.tp_slot_fail = b->tp_slot_flags.p; Ls.tp_slot_run = b->tp_slot_flags.p + tp_slots;
            Ls.tp_any_rerun = b->tp_slot_flags.p + 2 * (size_t)tp_slots;
            Ls.tp_n_chans_fzs = si.tp_fzs;
        }
        Ls.n_fzs_channels = si.n_fzs; Ls.Pmax_fzs = si.Pmax_fzs; Ls.S_mask_fzs = si.S_mask_fzs; Ls.Kmax_fzs = si.Kmax_fzs;
        Ls.d_fzs_ticket = b->d_ticket + FUSED_TICKETS; Ls.fzs_ticket_next = &b->fzs_ticket_next; Ls.fzs_ticket_cap = FZS_TICKETS;
        // A mixed bank has one fused launch per samples-per-symbol class that is large enough: the first on the bank's stream,
        // the others on aux streams, so that one class's tail (few warps left) overlaps the next class's start.
        int n_seg = 0, seg_total = 0, seg_count = 0;
        static const bool multi_stream = !(getenv("PSKD_FUSED_STREAMS") && atoi(getenv("PSKD_FUSED_STREAMS")) == 0);
        for (const FusedSeg& g : segs) if (g.slab == slab) { seg_total += g.count; seg_count++; }
        for (const FusedSeg& g : segs) {
            if (g.slab != slab) continue;
            FusedLaunch f{};
            f.S = g.S; f.d_list = b->d_list + g.first; f.h_list = h_list + g.first; f.n_list = g.count;
            // a unit = enough consecutive packets for >= ~4096 symbols
            static const long long unit_syms = getenv("PSKD_FUSED_UNIT") ? std::max(256, atoi(getenv("PSKD_FUSED_UNIT"))) : 4096;
            const long long ppu = std::max<long long>(1, (unit_syms * g.S + g.pkt_len_min - 1) / g.pkt_len_min);
            f.pkts_per_unit = (int)std::min<long long>(ppu, 1 << 20);
            // ONE long packet per channel and call (the packet-by-packet use behind serviceFunction): the packet is
            // cut into parts of >= ~2048 symbols, so that a 4096-channel bank is not 4096 units over 2960 resident
            // warps (measured: 64000-sample calls 1.05 -> 1.00 ms; with two packets per call the extra unit
            // start-ups already cost 5 %, on 1M-sample calls 8 %, hence the limit).  PSKD_FUSED_PARTS=1 switches it off.
            f.parts_per_pkt = 1;
            if (f.pkts_per_unit == 1 && g.max_pkts <= 1) {
                static const int max_parts = getenv("PSKD_FUSED_PARTS") ? std::max(1, atoi(getenv("PSKD_FUSED_PARTS"))) : 8;
                f.parts_per_pkt = (int)std::min<long long>(max_parts, std::max<long long>(1, g.pkt_len_min / (2048LL * g.S)));
            }
            f.units_per_channel = f.parts_per_pkt * ((g.max_pkts + f.pkts_per_unit - 1) / f.pkts_per_unit);
            if ((long long)f.units_per_channel * g.count > 0x7fffffffLL) return fail(PSKD_ERR_ARG, "too many packets");
            f.Amax = g.Amax; f.Pmax = g.Pmax;
            f.d_ticket = b->d_ticket + (slab * 4 + (g.S == 8 ? 0 : g.S == 9 ? 1 : g.S == 10 ? 2 : 3));
            f.d_done = b->d_done + lo;
            static const bool share = getenv("PSKD_FUSED_SHARE") && atoi(getenv("PSKD_FUSED_SHARE")) != 0;   // measured slower (config5: 49 vs 38 ms)
            f.grid_share = (share && multi_stream && seg_count > 1) ? (double)g.count / (double)seg_total : 0.0;
            if (n_seg == 0 || !multi_stream) {
                CUDA_TRY(launch_fused(Ls, f));
            } else {
                if (n_seg == 1) CUDA_TRY(cudaEventRecord(b->ev_call, b->stream));       // everything this call queued before the launches
                LaunchCtx La = Ls;
                La.stream = b->aux_stream[n_seg - 1];
                CUDA_TRY(cudaStreamWaitEvent(La.stream, b->ev_call, 0));
                CUDA_TRY(launch_fused(La, f));
                CUDA_TRY(cudaEventRecord(b->ev_aux[n_seg - 1], La.stream));
            }
            n_seg++;
        }
        for (int k = 1; k < n_seg && multi_stream; k++) CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_aux[k - 1], 0));
        if (si.n_fast + si.n_seq > 0) {
            CUDA_TRY(launch_front(Ls));
            CUDA_TRY(launch_fzs_front(Ls));
            CUDA_TRY(launch_chain_par(Ls));
            CUDA_TRY(launch_chain_seq(Ls));
            CUDA_TRY(launch_back(Ls));
        }
        CUDA_TRY(launch_finish(Ls));
        return PSKD_OK;
    };
    if (!host_bufs) {
        rc = run_slab(0, 0, nch);
        if (rc != PSKD_OK) return rc;
    } else {
        for (int s = 0; s < n_slabs; s++) {
            const int lo = slab_lo(s), hi = slab_lo(s + 1);
            const int r = (int)(b->slab_seq % RING);
            b->slab_seq++;
            const size_t rows = (size_t)(hi - lo);
            // H2D into ring slot r once the kernels of the slab that used it last have read their input
            CUDA_TRY(cudaStreamWaitEvent(b->copy_in, b->ev_kern[r], 0));
            if (nmax > 0)
                CUDA_TRY(cudaMemcpy2DAsync(b->st_in.p + (size_t)r * slot_in, in_stride * 8, in->iq + 2 * in->iq_stride * lo,
                                           in->iq_stride * 8, (size_t)nmax * 8, rows, cudaMemcpyHostToDevice, b->copy_in));
            CUDA_TRY(cudaEventRecord(b->ev_h2d[r], b->copy_in));
            // kernels once the input is there and the slot's previous outputs have left for the host
            CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_h2d[r], 0));
            CUDA_TRY(cudaStreamWaitEvent(b->stream, b->ev_d2h[r], 0));
            dev_soft = out->soft ? b->st_soft.p + (size_t)r * 2 * slot_sym : nullptr;
            dev_phase = out->phase ? b->st_phase.p + (size_t)r * slot_sym : nullptr;
            dev_bits = out->bits ? b->st_bits.p + (size_t)r * slot_bits : nullptr;
            dev_sidx = out->sample_index ? b->st_sidx.p + (size_t)r * slot_sym : b->sidx_tmp.p;
            dev_hard = out->hard ? b->st_hard.p + (size_t)r * slot_sym : nullptr;
            rc = run_slab(s, lo, hi);
            if (rc != PSKD_OK) return rc;
            CUDA_TRY(cudaEventRecord(b->ev_kern[r], b->stream));
            CUDA_TRY(cudaStreamWaitEvent(b->copy_out, b->ev_kern[r], 0));
            if (Kmax > 0) {
                if (out->soft) CUDA_TRY(cudaMemcpy2DAsync(out->soft + 2 * out->sym_stride * lo, out->sym_stride * 8, dev_soft, sym_stride * 8, (size_t)Kmax * 8, rows, cudaMemcpyDeviceToHost, b->copy_out));
                if (out->phase) CUDA_TRY(cudaMemcpy2DAsync(out->phase + out->sym_stride * lo, out->sym_stride * 4, dev_phase, sym_stride * 4, (size_t)Kmax * 4, rows, cudaMemcpyDeviceToHost, b->copy_out));
                if (out->sample_index) CUDA_TRY(cudaMemcpy2DAsync(out->sample_index + out->sym_stride * lo, out->sym_stride * 2, dev_sidx, sym_stride * 2, (size_t)Kmax * 2, rows, cudaMemcpyDeviceToHost, b->copy_out));
                if (out->hard) CUDA_TRY(cudaMemcpy2DAsync(out->hard + out->sym_stride * lo, out->sym_stride, dev_hard, sym_stride, (size_t)Kmax, rows, cudaMemcpyDeviceToHost, b->copy_out));
                if (out->bits) {
                    size_t w = std::min((size_t)Kmax * 3, out->bits_stride);
                    CUDA_TRY(cudaMemcpy2DAsync(out->bits + out->bits_stride * lo, out->bits_stride * 2, dev_bits, bits_stride * 2, w * 2, rows, cudaMemcpyDeviceToHost, b->copy_out));
                }
            }
            CUDA_TRY(cudaEventRecord(b->ev_d2h[r], b->copy_out));
        }
    }

    // ---- host-side bookkeeping (what the reference's members hold after the packets) ------------
    for (int i = 0; i < nch; i++) {
        ChanHost& c = b->ch[i];
        const ChanDesc& d = b->h_desc[i];
        c.tail_len = d.next_tail_len;
        c.win_size = stalled[i] ? c.win_size + d.n_in : d.next_tail_len;
        c.latched = c.props;
        if (d.n_pkts > 0) {
            // the LinearFit history is cleared by resetNumSymbols (:416-420) and by a change of the SRI rate (:91-102);
            // a phaseAvg change keeps at most the newest values (:103-109)
            if (c.resetNumSymbols || c.resetPhaseAvg || c.last_xdelta != in->sri_xdelta) c.hist_syms = 0;
            c.hist_syms += d.K;
            const bool sri_block = !stalled[i] || (d.flags & (CH_SRI_CHANGED | CH_RESET_NUMSYMS));
            if (sri_block) c.last_xdelta = in->sri_xdelta; else c.last_xdelta = 0.0;
            c.resetNumSymbols = false; c.resetPhaseAvg = false;
            c.fit_n = d.P;
            if (sri_block) {
                // out-port SRIs (cpp/psk_soft.cpp:399-404), pushed on every packet (a stalled one: only with sriChanged / resetNumSymbols)
                double xd = in->sri_xdelta * (double)d.S;
                c.sri.soft_xdelta = xd; c.sri.soft_mode = 1;
                c.sri.phase_xdelta = xd; c.sri.phase_mode = 0;
                c.sri.bits_xdelta = xd / (double)d.bpb; c.sri.bits_mode = 0;
                c.sri.sri_pushes += d.n_pkts;
            }
        }
        if (out->n_symbols) out->n_symbols[i] = (size_t)d.K;
        if (out->n_bits) out->n_bits[i] = (size_t)(d.K * d.bpb);
        b->stats.symbols_out += (uint64_t)d.K;
        b->stats.samples_in += (uint64_t)d.n_in;
        b->stats.packets += (uint64_t)d.n_pkts;
    }
    b->tail_cur = next_tail;

    if (!(in->flags & PSKD_FLAG_NO_SYNC)) {
        if (host_bufs) CUDA_TRY(cudaStreamSynchronize(b->copy_out));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return any_nobits ? PSKD_NO_BITS : PSKD_OK;
}

}  // extern "C"
