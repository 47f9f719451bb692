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
    long long tail_len = 0;    // samples.size() between calls
    bool resetNumSymbols = true, resetPhaseAvg = true, resetSamplesPerBaud = true;   // cpp/psk_soft.cpp:191-193
    size_t symbolEnergySize = 10;   // symbolEnergy.size() (cpp/psk_soft.cpp:189), for the listener at :640
    bool first_packet = true;
    int fit_n = 0;             // LinearFit::n currently held in the device state
    long long tail_off = 0;    // element offset of this channel's tail region
    long long tail_cap = 0;
    pskd_sri_out sri{};
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
    DevBuf<float> st_in; DevBuf<float> st_soft; DevBuf<float> st_phase; DevBuf<int16_t> st_bits; DevBuf<int16_t> st_sidx;
    unsigned long long launches = 0;
    pskd_stats stats{};
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // host-buffer mode: H2D / D2H overlap the kernels slab by slab
    cudaEvent_t slab_in[16] = {nullptr}, slab_done[16] = {nullptr};
    int chain_mode = 0;        // 0 auto (scan-based where possible), 1 force the sequential chain (PSKD_CHAIN=seq)
    int fused_mode = -1;       // -1 auto (large banks), 0 never, 1 whenever a channel qualifies (PSKD_FUSED)
    int host_slabs = 16;       // host-buffer mode: channel slabs the H2D / kernels / D2H pipeline works through (PSKD_SLABS, 1..16)
    int fused_min_channels = 1152;   // auto: channels per launch from which the fused kernel beats the staged ones (measured crossover ~1120 for 1M-sample 8-PSK calls; PSKD_FUSED_MIN)
    int* d_list = nullptr;     // fused launch lists (channel indices), one segment per (slab, samplesPerBaud)
    int* h_list_slot[2] = {nullptr, nullptr};
    int* d_done = nullptr;     // [n_channels] units completed per channel in the current call
    int* d_ticket = nullptr;   // [16 slabs x 4 samplesPerBaud values] unit ticket counters
    int tp_mode = -1;          // time-parallel chain of the staged path: -1 auto (few channels, many packets), 0 never, 1 whenever possible (PSKD_TP)
    DevBuf<TpItem> tp_items; DevBuf<TpChan> tp_chans; DevBuf<TpPacket> tp_pkts; DevBuf<TpEnd> tp_ends;
    DevBuf<float> tp_end_ring, tp_start_ring; DevBuf<int> tp_fail;
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
static int ensure_tails(pskd_bank* b) {
    bool grow = false;
    for (auto& c : b->ch) {
        long long need = (long long)c.props.samplesPerBaud * c.props.numAvg + c.props.samplesPerBaud;
        if (need > c.tail_cap) grow = true;
    }
    if (!grow) return PSKD_OK;
    std::vector<long long> new_off(b->n_channels), new_cap(b->n_channels);
    long long total = 0;
    for (int i = 0; i < b->n_channels; i++) {
        auto& c = b->ch[i];
        long long need = (long long)c.props.samplesPerBaud * c.props.numAvg + c.props.samplesPerBaud;
        new_cap[i] = std::max(need, c.tail_cap);
        new_off[i] = total;
        total += (new_cap[i] + 1) & ~1LL;     // keep 16-byte alignment of every region
    }
    float2* nt[2] = {nullptr, nullptr};
    CUDA_TRY(cudaMalloc((void**)&nt[0], std::max<long long>(total, 1) * sizeof(float2)));
    CUDA_TRY(cudaMalloc((void**)&nt[1], std::max<long long>(total, 1) * sizeof(float2)));
    if (b->d_tail[0]) {
        for (int i = 0; i < b->n_channels; i++) {
            auto& c = b->ch[i];
            if (c.tail_len > 0)
                CUDA_TRY(cudaMemcpyAsync(nt[b->tail_cur] + new_off[i], b->d_tail[b->tail_cur] + c.tail_off,
                                         c.tail_len * sizeof(float2), cudaMemcpyDeviceToDevice, b->stream));
        }
        CUDA_TRY(cudaStreamSynchronize(b->stream));
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
    if (const char* e = getenv("PSKD_SLABS")) b->host_slabs = std::min(16, std::max(1, atoi(e)));
    if (const char* e = getenv("PSKD_TP")) b->tp_mode = (strcmp(e, "auto") == 0) ? -1 : atoi(e) != 0;
    cudaError_t e;
#define CT(expr) do { e = (expr); if (e != cudaSuccess) { int rc = fail(e == cudaErrorMemoryAllocation ? PSKD_ERR_NOMEM : PSKD_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e)); pskd_destroy(b); return rc; } } while (0)
    CT(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CT(cudaStreamCreateWithFlags(&b->copy_in, cudaStreamNonBlocking));
    CT(cudaStreamCreateWithFlags(&b->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 16; i++) {
        CT(cudaEventCreateWithFlags(&b->slab_in[i], cudaEventDisableTiming));
        CT(cudaEventCreateWithFlags(&b->slab_done[i], cudaEventDisableTiming));
    }
    CT(cudaMalloc((void**)&b->d_desc, sizeof(ChanDesc) * n_channels));
    for (int i = 0; i < 2; i++) {
        CT(cudaMallocHost((void**)&b->h_desc_slot[i], sizeof(ChanDesc) * n_channels));
        CT(cudaMallocHost((void**)&b->h_list_slot[i], sizeof(int) * n_channels));
        CT(cudaEventCreateWithFlags(&b->desc_ev[i], cudaEventDisableTiming));
    }
    CT(cudaMalloc((void**)&b->d_list, sizeof(int) * n_channels));
    CT(cudaMalloc((void**)&b->d_done, sizeof(int) * n_channels));
    CT(cudaMalloc((void**)&b->d_ticket, sizeof(int) * 64));
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
    if (b->stream) cudaStreamSynchronize(b->stream);
    cudaFree(b->d_desc);
    for (int i = 0; i < 2; i++) { if (b->h_desc_slot[i]) cudaFreeHost(b->h_desc_slot[i]); if (b->h_list_slot[i]) cudaFreeHost(b->h_list_slot[i]); if (b->desc_ev[i]) cudaEventDestroy(b->desc_ev[i]); }
    cudaFree(b->d_list); cudaFree(b->d_done); cudaFree(b->d_ticket);
    cudaFree(b->d_state); cudaFree(b->d_counters); cudaFree(b->d_ring);
    cudaFree(b->d_tail[0]); cudaFree(b->d_tail[1]);
    b->sel.release(); b->theta.release(); b->phase_tmp.release(); b->sidx_tmp.release();
    b->st_in.release(); b->st_soft.release(); b->st_phase.release(); b->st_bits.release(); b->st_sidx.release();
    b->tp_items.release(); b->tp_chans.release(); b->tp_pkts.release(); b->tp_ends.release();
    b->tp_end_ring.release(); b->tp_start_ring.release(); b->tp_fail.release();
    b->prof.destroy();
    for (int i = 0; i < 16; i++) { if (b->slab_in[i]) cudaEventDestroy(b->slab_in[i]); if (b->slab_done[i]) cudaEventDestroy(b->slab_done[i]); }
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
    CUDA_TRY(cudaStreamSynchronize(b->stream));
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
            out[k].ms_total = b->prof.ms[i]; out[k].launches = b->prof.launches[i];
        }
        k++;
    }
    *n = k;
    if (reset) for (int i = 0; i < KID_COUNT; i++) { b->prof.ms[i] = 0; b->prof.launches[i] = 0; }
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
    st->tp_packets = c.tp_packets;
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
    const unsigned char* p = static_cast<const unsigned char*>(host_buf);
    StateHeader h;
    memcpy(&h, p, sizeof(h)); p += sizeof(h);
    if (h.magic != STATE_MAGIC || h.version != STATE_VERSION) return fail(PSKD_ERR_ARG, "pskd_state_import: not a pskd state blob (or another version)");
    if (h.n_channels != b->n_channels) return fail(PSKD_ERR_ARG, "pskd_state_import: blob has %d channels, bank has %d", h.n_channels, b->n_channels);
    if (h.total_bytes > n_bytes || h.ring_cap < 0) return fail(PSKD_ERR_ARG, "pskd_state_import: truncated blob");
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    const int nch = b->n_channels;
    std::vector<StateChan> sc(nch);
    memcpy(sc.data(), p, sizeof(StateChan) * nch); p += sizeof(StateChan) * nch;
    for (int i = 0; i < nch; i++) {
        int rc = check_props(sc[i].props);
        if (rc != PSKD_OK) return rc;
        if (sc[i].tail_len < 0) return fail(PSKD_ERR_ARG, "pskd_state_import: corrupt blob");
    }
    for (int i = 0; i < nch; i++) {
        ChanHost& c = b->ch[i];
        c.props = sc[i].props; c.latched = sc[i].latched;
        c.resetNumSymbols = sc[i].resetNumSymbols != 0; c.resetPhaseAvg = sc[i].resetPhaseAvg != 0;
        c.resetSamplesPerBaud = sc[i].resetSamplesPerBaud != 0; c.first_packet = sc[i].first_packet != 0;
        c.fit_n = sc[i].fit_n; c.symbolEnergySize = (size_t)sc[i].symbolEnergySize; c.sri = sc[i].sri;
        c.tail_len = 0;          // set below once the regions are large enough
    }
    int rc = ensure_tails(b);
    if (rc == PSKD_OK) rc = ensure_rings(b);
    if (rc != PSKD_OK) return rc;
    if (h.ring_cap > b->ring_cap) return fail(PSKD_ERR_ARG, "pskd_state_import: history ring of the blob does not fit the bank");
    std::vector<ChanState> dev(nch);
    for (int i = 0; i < nch; i++) dev[i] = sc[i].dev;
    CUDA_TRY(cudaMemcpy(b->d_state, dev.data(), sizeof(ChanState) * nch, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy2D(b->d_ring, (size_t)b->ring_cap * sizeof(float), p, (size_t)h.ring_cap * sizeof(float),
                          (size_t)h.ring_cap * sizeof(float), nch, cudaMemcpyHostToDevice));
    p += (size_t)nch * h.ring_cap * sizeof(float);
    for (int i = 0; i < nch; i++) {
        ChanHost& c = b->ch[i];
        const long long tl = sc[i].tail_len;
        if (tl > c.tail_cap) return fail(PSKD_ERR_ARG, "pskd_state_import: carried window of channel %d does not fit", i);
        if (tl > 0) {
            CUDA_TRY(cudaMemcpy(b->d_tail[b->tail_cur] + c.tail_off, p, (size_t)tl * sizeof(float2), cudaMemcpyHostToDevice));
            p += (size_t)tl * sizeof(float2);
        }
        c.tail_len = tl;
    }
    return PSKD_OK;
}

int pskd_process(pskd_handle b, const pskd_input* in, pskd_output* out) {
    if (!b || !in || !out) return fail(PSKD_ERR_ARG, "pskd_process: null argument");
    if (!in->iq && (in->n_complex || in->n_complex_all)) return fail(PSKD_ERR_ARG, "pskd_process: iq is NULL");
    const int nch = b->n_channels;
    const bool host_bufs = (in->flags & PSKD_FLAG_HOST_BUFFERS) != 0;
    if (host_bufs && (in->flags & PSKD_FLAG_NO_SYNC)) return fail(PSKD_ERR_ARG, "PSKD_FLAG_NO_SYNC needs device buffers");
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
    int rc = ensure_tails(b);
    if (rc != PSKD_OK) return rc;
    rc = ensure_rings(b);
    if (rc != PSKD_OK) return rc;

    // ---- per-channel geometry ---------------------------------------------------------------
    b->desc_slot ^= 1;
    b->h_desc = b->h_desc_slot[b->desc_slot];
    CUDA_TRY(cudaEventSynchronize(b->desc_ev[b->desc_slot]));   // previous upload from this slot is done
    long long Kmax = 0, scr_total = 0, nmax = 0;
    int Smax = 2, Amax = 1, Pmax_fast = 1, n_fast = 0, n_seq = 0;
    unsigned long long S_mask = 0, S_mask_fast = 0;
    int Amax_fast = 1, Amin_fast = 1 << 30;
    bool any_nobits = false;
    const int next_tail = b->tail_cur ^ 1;
    // host-buffer mode works through slabs of channels (H2D / kernels / D2H overlap); the fused
    // kernel is chosen per launch, i.e. per slab
    int n_slabs = 1;
    if (host_bufs) {
        long long nm = 0;
        for (int i = 0; i < nch; i++) nm = std::max(nm, (long long)(in->n_complex ? in->n_complex[i] : in->n_complex_all));
        n_slabs = std::min(nch, nm * (long long)nch >= (1 << 22) ? b->host_slabs : 1);
    }
    std::vector<unsigned char> fusable(nch, 0);
    for (int i = 0; i < nch; i++) {
        ChanHost& c = b->ch[i];
        const long long n_in = (long long)(in->n_complex ? in->n_complex[i] : in->n_complex_all);
        if ((size_t)n_in > in->iq_stride && nch > 1) return fail(PSKD_ERR_ARG, "channel %d: n_complex > iq_stride", i);
        const int S = c.props.samplesPerBaud, A = (int)c.props.numAvg, M = c.props.constelationSize, P = c.props.phaseAvg;
        const long long numDataPts = (long long)S * A;
        // resyncEnergy (cpp/psk_soft.cpp:619-636): runs when asked for, or whenever the window is not
        // full (:380-383, i.e. every packet in steady state).  Its only observable effects here are
        // the truncation of an over-full window and symbolEnergy.size().
        if (c.resetSamplesPerBaud || numDataPts > c.tail_len) {
            if (c.tail_len > numDataPts) c.tail_len = numDataPts;
            c.symbolEnergySize = S;
            c.resetSamplesPerBaud = false;
        }
        if (c.tail_len >= numDataPts && n_in > 0)
            return fail(PSKD_ERR_UNSUPPORTED, "channel %d: window over-full after a numAvg/samplesPerBaud decrease; the reference stalls "
                        "forever here (cpp/psk_soft.cpp:457 never true again) -- not carried on the GPU path", i);
        const long long total = c.tail_len + n_in;
        long long K = total / S - A + 1;
        if (K < 0) K = 0;
        const long long pkt = in->packet_len ? (long long)in->packet_len : std::max<long long>(n_in, 1);
        const long long npk = n_in > 0 ? (n_in + pkt - 1) / pkt : 0;
        if (npk > 0x7fffffffLL) return fail(PSKD_ERR_ARG, "too many packets");
        ChanDesc& d = b->h_desc[i];
        d.in = host_bufs ? nullptr : reinterpret_cast<const float2*>(in->iq) + (size_t)i * in->iq_stride;
        d.tail = b->d_tail[b->tail_cur] + c.tail_off;
        d.tail_next = b->d_tail[next_tail] + c.tail_off;
        d.n_in = n_in; d.tail_len = c.tail_len; d.K = K;
        d.next_tail_len = total - K * S;
        if (d.next_tail_len > c.tail_cap) return fail(PSKD_ERR_UNSUPPORTED, "channel %d: carried window exceeds its capacity", i);
        d.sym_off = (long long)i * (long long)out->sym_stride;
        d.bits_off = (long long)i * (long long)out->bits_stride;
        d.pkt_len = pkt; d.n_pkts = (int)npk;
        d.S = S; d.A = A; d.M = M; d.P = P; d.D = c.props.differentialDecoding ? 1 : 0; d.bpb = bpb_of(M);
        d.ring_off = i * b->ring_cap;
        d.flags = (c.resetNumSymbols ? CH_RESET_NUMSYMS : 0) | (c.resetPhaseAvg ? CH_RESET_PHASEAVG : 0);
        // the scan-based chain needs the history staged in shared memory and an unchanged window length
        const bool fast = b->chain_mode == 0 && P <= CHAIN_PAR_PMAX && c.fit_n == P;
        if (fast) d.flags |= CH_FAST;
        fusable[i] = fast && b->fused_mode != 0 && fused_supports(S, A, P) && K < (1LL << 30);
        if (d.bpb == 0) any_nobits = true;
        if ((size_t)K > out->sym_stride && (out->soft || out->phase || out->sample_index))
            return fail(PSKD_ERR_CAPACITY, "channel %d emits %lld symbols > sym_stride %zu", i, K, out->sym_stride);
        if (out->bits && (size_t)(K * d.bpb) > out->bits_stride)
            return fail(PSKD_ERR_CAPACITY, "channel %d emits %lld bits > bits_stride %zu", i, K * d.bpb, out->bits_stride);
        Kmax = std::max(Kmax, K); nmax = std::max(nmax, n_in);
    }
    // ---- which channels take the fused kernel (per slab), which the staged kernels --------------
    struct FusedSeg { int slab, S, first, count, Amax, Pmax, max_pkts; long long pkt_len_min; };
    std::vector<FusedSeg> segs;
    int* h_list = b->h_list_slot[b->desc_slot];
    int n_listed = 0;
    static const int fusedS[4] = {8, 9, 10, 16};
    for (int s = 0; s < n_slabs; s++) {
        const int lo = (int)((long long)s * nch / n_slabs), hi = (int)((long long)(s + 1) * nch / n_slabs);
        int nf = 0;
        for (int i = lo; i < hi; i++) nf += fusable[i];
        const bool use = nf > 0 && (b->fused_mode == 1 || nf >= b->fused_min_channels);
        if (!use) continue;
        for (int si = 0; si < 4; si++) {
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
    for (int i = 0; i < nch; i++) {
        ChanDesc& d = b->h_desc[i];
        if (d.flags & CH_FUSED) { d.scr_off = 0; continue; }
        if (d.flags & CH_FAST) { n_fast++; Pmax_fast = std::max(Pmax_fast, d.P); } else n_seq++;
        d.scr_off = scr_total;
        scr_total += (d.K + 3) & ~3LL;
        Smax = std::max(Smax, d.S); Amax = std::max(Amax, d.A);
        if (d.K > 0) {
            const bool front_fast = (d.S == 8 || d.S == 9 || d.S == 10 || d.S == 16) && d.A <= FRONT_FAST_AMAX;
            if (front_fast) { d.flags |= CH_FRONT_FAST; S_mask_fast |= 1ull << d.S; Amax_fast = std::max(Amax_fast, d.A); Amin_fast = std::min(Amin_fast, d.A); }
            else S_mask |= 1ull << d.S;
        }
    }
    const bool any_staged = (n_fast + n_seq) > 0;

    // ---- time-parallel chain plan (staged path): channels with many packets whose chain would otherwise
    // run packet after packet on one warp.  Heads = packets before the first packet that is certain to start
    // with a full history; one item per later packet.
    std::vector<TpItem> tp_heads, tp_items;
    std::vector<TpChan> tp_chans;
    int tp_slots = 0, tp_records = 0, tp_Pmax = 1;
    if (b->tp_mode != 0 && (!host_bufs || n_slabs == 1) && n_fast > 0 && (b->tp_mode == 1 || n_fast <= 2048) && in->sri_xdelta != 1.0) {
        for (int i = 0; i < nch; i++) {
            ChanDesc& d = b->h_desc[i];
            if (!(d.flags & CH_FAST) || (d.flags & CH_FUSED) || d.P < 2 || d.K <= 0) continue;
            if (d.pkt_len / d.S - 1 < d.P + 2) continue;                   // every full packet must hold the whole history
            int j0 = -1;
            for (int j = 1; j < d.n_pkts; j++)
                if (first_symbol_at((long long)j * d.pkt_len, d.tail_len, d.S, d.A, d.K) >= d.P) { j0 = j; break; }
            if (j0 < 0 || d.n_pkts - j0 < (b->tp_mode == 1 ? 2 : 4)) continue;
            d.flags |= CH_TP;
            TpChan tc{i, j0, d.n_pkts, tp_records, tp_slots, 0};
            tp_chans.push_back(tc);
            tp_heads.push_back(TpItem{i, 0, j0, 0, -1, tp_records, -1, 0});
            for (int j = j0; j < d.n_pkts; j++) {
                const int rec = tp_records + 1 + (j - j0);
                tp_items.push_back(TpItem{i, j, j + 1, j == j0 ? 2 : 1, tp_records, rec, tp_slots + (j - j0), 0});
            }
            tp_records += 1 + (d.n_pkts - j0);
            tp_slots += d.n_pkts - j0;
            tp_Pmax = std::max(tp_Pmax, d.P);
        }
    }

    // ---- buffers ------------------------------------------------------------------------------
    CUDA_TRY(b->sel.reserve((size_t)scr_total + 4));
    CUDA_TRY(b->theta.reserve((size_t)scr_total + 4));
    float* dev_soft = out->soft; float* dev_phase = out->phase; int16_t* dev_bits = out->bits; int16_t* dev_sidx = out->sample_index;
    size_t sym_stride = out->sym_stride, bits_stride = out->bits_stride, in_stride = 0;
    if (host_bufs) {
        // stage through device buffers with the caller's strides squeezed to what this call needs
        sym_stride = (size_t)((Kmax + 7) & ~7LL); bits_stride = sym_stride * 3;
        in_stride = (size_t)((nmax + 1) & ~1LL);
        CUDA_TRY(b->st_in.reserve(2 * in_stride * nch + 4));
        if (out->soft) { CUDA_TRY(b->st_soft.reserve(2 * sym_stride * nch + 4)); dev_soft = b->st_soft.p; }
        if (out->phase) { CUDA_TRY(b->st_phase.reserve(sym_stride * nch + 4)); dev_phase = b->st_phase.p; }
        if (out->bits) { CUDA_TRY(b->st_bits.reserve(bits_stride * nch + 4)); dev_bits = b->st_bits.p; }
        if (out->sample_index) { CUDA_TRY(b->st_sidx.reserve(sym_stride * nch + 4)); dev_sidx = b->st_sidx.p; }
        for (int i = 0; i < nch; i++) {
            ChanDesc& d = b->h_desc[i];
            d.in = reinterpret_cast<const float2*>(b->st_in.p) + (size_t)i * in_stride;
            d.sym_off = (long long)i * (long long)sym_stride;
            d.bits_off = (long long)i * (long long)bits_stride;
        }
    }
    if (!dev_phase && any_staged) { CUDA_TRY(b->phase_tmp.reserve(sym_stride * nch + 4)); }
    if (!dev_sidx) { CUDA_TRY(b->sidx_tmp.reserve(sym_stride * nch + 4)); dev_sidx = b->sidx_tmp.p; }

    CUDA_TRY(cudaMemcpyAsync(b->d_desc, b->h_desc, sizeof(ChanDesc) * nch, cudaMemcpyHostToDevice, b->stream));
    if (n_listed > 0) {
        CUDA_TRY(cudaMemcpyAsync(b->d_list, h_list, sizeof(int) * n_listed, cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemsetAsync(b->d_done, 0, sizeof(int) * nch, b->stream));
        CUDA_TRY(cudaMemsetAsync(b->d_ticket, 0, sizeof(int) * 64, b->stream));
    }
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
        // pageable sources: the copies are staged before the calls return
        CUDA_TRY(cudaMemcpyAsync(b->tp_items.p, tp_heads.data(), sizeof(TpItem) * tp_heads.size(), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemcpyAsync(b->tp_items.p + tp_heads.size(), tp_items.data(), sizeof(TpItem) * tp_items.size(), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemcpyAsync(b->tp_chans.p, tp_chans.data(), sizeof(TpChan) * tp_chans.size(), cudaMemcpyHostToDevice, b->stream));
        CUDA_TRY(cudaMemsetAsync(b->tp_fail.p, 0, sizeof(int) * nch, b->stream));
    }

    LaunchCtx L{};
    L.stream = b->stream; L.n_channels = nch; L.Kmax = Kmax; L.Smax = Smax; L.Amax = Amax;
    L.S_mask = S_mask; L.S_mask_fast = S_mask_fast; L.Amax_fast = Amax_fast; L.Amin_fast = Amin_fast; L.Pmax_fast = Pmax_fast; L.n_fast_channels = n_fast; L.n_seq_channels = n_seq;
    L.d_desc = b->d_desc; L.d_state = b->d_state; L.d_ring = b->d_ring;
    L.d_sel = b->sel.p; L.d_theta = b->theta.p; L.d_phase_tmp = b->phase_tmp.p;
    L.out_soft = dev_soft; L.out_bits = dev_bits; L.out_phase = dev_phase; L.out_sidx = dev_sidx;
    L.sri_xdelta = in->sri_xdelta; L.d_counters = b->d_counters; L.launches = &b->launches; L.prof = &b->prof;
    if (!tp_chans.empty()) {
        L.tp_head_items = b->tp_items.p; L.tp_n_head = (int)tp_heads.size();
        L.tp_items = b->tp_items.p + tp_heads.size(); L.tp_n_items = (int)tp_items.size();
        L.tp_chans = b->tp_chans.p; L.tp_n_chans = (int)tp_chans.size(); L.tp_n_slots = tp_slots;
        L.tp_pkts = b->tp_pkts.p; L.tp_ends = b->tp_ends.p; L.tp_end_ring = b->tp_end_ring.p; L.tp_start_ring = b->tp_start_ring.p;
        L.tp_ring_stride = tp_stride; L.tp_fail = b->tp_fail.p;
    }

    // the kernels of one slab of channels [lo, hi)
    auto run_slab = [&](const LaunchCtx& Ls, int slab, int lo) -> int {
        for (const FusedSeg& g : segs) {
            if (g.slab != slab) continue;
            FusedLaunch f{};
            f.S = g.S; f.d_list = b->d_list + g.first; f.n_list = g.count;
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
            CUDA_TRY(launch_fused(Ls, f));
        }
        if (any_staged) {
            CUDA_TRY(launch_front(Ls));
            CUDA_TRY(launch_chain_par(Ls));
            CUDA_TRY(launch_chain_seq(Ls));
            CUDA_TRY(launch_back(Ls));
        }
        CUDA_TRY(launch_finish(Ls));
        return PSKD_OK;
    };
    if (!host_bufs) {
        rc = run_slab(L, 0, 0);
        if (rc != PSKD_OK) return rc;
    } else {
        // host buffers: channel slabs flow H2D (copy_in) -> kernels (stream) -> D2H (copy_out), so the
        // PCIe transfers of neighbouring slabs overlap the kernels (channels are independent)
        for (int s = 0; s < n_slabs; s++) {
            const int lo = (int)((long long)s * nch / n_slabs), hi = (int)((long long)(s + 1) * nch / n_slabs);
            if (nmax > 0)
                CUDA_TRY(cudaMemcpy2DAsync(b->st_in.p + 2 * in_stride * lo, in_stride * 8, in->iq + 2 * in->iq_stride * lo,
                                           in->iq_stride * 8, (size_t)nmax * 8, hi - lo, cudaMemcpyHostToDevice, b->copy_in));
            CUDA_TRY(cudaEventRecord(b->slab_in[s], b->copy_in));
        }
        for (int s = 0; s < n_slabs; s++) {
            const int lo = (int)((long long)s * nch / n_slabs), hi = (int)((long long)(s + 1) * nch / n_slabs);
            LaunchCtx Ls = L;
            Ls.d_desc = b->d_desc + lo; Ls.d_state = b->d_state + lo; Ls.n_channels = hi - lo;
            CUDA_TRY(cudaStreamWaitEvent(b->stream, b->slab_in[s], 0));
            rc = run_slab(Ls, s, lo);
            if (rc != PSKD_OK) return rc;
            CUDA_TRY(cudaEventRecord(b->slab_done[s], b->stream));
            CUDA_TRY(cudaStreamWaitEvent(b->copy_out, b->slab_done[s], 0));
            if (Kmax > 0) {
                const size_t rows = (size_t)(hi - lo);
                if (out->soft) CUDA_TRY(cudaMemcpy2DAsync(out->soft + 2 * out->sym_stride * lo, out->sym_stride * 8, dev_soft + 2 * sym_stride * lo, sym_stride * 8, (size_t)Kmax * 8, rows, cudaMemcpyDeviceToHost, b->copy_out));
                if (out->phase) CUDA_TRY(cudaMemcpy2DAsync(out->phase + out->sym_stride * lo, out->sym_stride * 4, dev_phase + sym_stride * lo, sym_stride * 4, (size_t)Kmax * 4, rows, cudaMemcpyDeviceToHost, b->copy_out));
                if (out->sample_index) CUDA_TRY(cudaMemcpy2DAsync(out->sample_index + out->sym_stride * lo, out->sym_stride * 2, dev_sidx + sym_stride * lo, sym_stride * 2, (size_t)Kmax * 2, rows, cudaMemcpyDeviceToHost, b->copy_out));
                if (out->bits) {
                    size_t w = std::min((size_t)Kmax * 3, out->bits_stride);
                    CUDA_TRY(cudaMemcpy2DAsync(out->bits + out->bits_stride * lo, out->bits_stride * 2, dev_bits + bits_stride * lo, bits_stride * 2, w * 2, rows, cudaMemcpyDeviceToHost, b->copy_out));
                }
            }
        }
    }

    // ---- host-side bookkeeping (what the reference's members hold after the packets) ------------
    for (int i = 0; i < nch; i++) {
        ChanHost& c = b->ch[i];
        const ChanDesc& d = b->h_desc[i];
        c.tail_len = d.next_tail_len;
        c.latched = c.props;
        if (d.n_pkts > 0) {
            c.resetNumSymbols = false; c.resetPhaseAvg = false;
            c.fit_n = d.P;
            // out-port SRIs (cpp/psk_soft.cpp:399-404), pushed on every packet
            double xd = in->sri_xdelta * (double)d.S;
            c.sri.soft_xdelta = xd; c.sri.soft_mode = 1;
            c.sri.phase_xdelta = xd; c.sri.phase_mode = 0;
            c.sri.bits_xdelta = xd / (double)d.bpb; c.sri.bits_mode = 0;
            c.sri.sri_pushes += d.n_pkts;
        }
        if (out->n_symbols) out->n_symbols[i] = (size_t)d.K;
        if (out->n_bits) out->n_bits[i] = (size_t)(d.K * d.bpb);
        b->stats.symbols_out += (uint64_t)d.K;
        b->stats.samples_in += (uint64_t)d.n_in;
        b->stats.packets += (uint64_t)d.n_pkts;
    }
    b->tail_cur = next_tail;

    if (host_bufs) {
        CUDA_TRY(cudaStreamSynchronize(b->copy_out));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    } else if (!(in->flags & PSKD_FLAG_NO_SYNC)) {
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    return any_nobits ? PSKD_NO_BITS : PSKD_OK;
}

}  // extern "C"
