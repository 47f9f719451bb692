// psk_soft_gpu.hpp -- C++ host-side mirror of the reference component's interface for the demod
// path, over the C ABI of include/pskd.h (libpskd.so).  Header-only, C++11, no CUDA headers.
//
//   psk_soft_gpu   one component: the reference's property members under their own names
//                  (cpp/psk_soft_base.h:44-56), a serviceFunction()-shaped call that takes what
//                  getPacket() delivers (cpp/psk_soft.cpp:349: dataBuffer, SRI.xdelta, SRI.mode,
//                  inputQueueFlushed) and fills what the four pushPacket calls send (:605-615).
//   psk_bank_gpu   a bank of independent channels on one GPU (device- or host-resident buffers).
//   psk_box_gpu    a bank sharded by contiguous channel ranges over the GPUs of one box, one host
//                  thread per GPU, no collective.
//
// Return conventions follow the reference: NORMAL (1) for every consumed packet, including an
// ignored real-data packet (cpp/psk_soft.cpp:359-363); NOOP (-1) when there is nothing to do
// (:350-352).  Device / argument problems throw psk_gpu_error (the reference has no such failure
// mode; there is NO CPU fallback).
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/pskd.h"

enum { PSK_NOOP = -1, PSK_NORMAL = 1 };

struct psk_gpu_error : std::runtime_error {
    int code;
    psk_gpu_error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void psk_check(int rc) {
    if (rc < 0) throw psk_gpu_error(rc, std::string("pskd: ") + pskd_last_error());
}

// the packet a BULKIO in-port hands over (bulkio::InFloatPort::dataTransfer, fields used at
// cpp/psk_soft.cpp:353,359,394,428)
struct psk_packet {
    std::vector<float> dataBuffer;      // interleaved re,im
    double xdelta = 1.0;                // SRI.xdelta
    short mode = 1;                     // SRI.mode (1 = complex)
    bool sriChanged = false;
    bool inputQueueFlushed = false;
    bool EOS = false;
};

// what the component pushes (cpp/psk_soft.cpp:605-615); empty vectors are "not pushed"
struct psk_outputs {
    std::vector<std::complex<float> > softDecision;   // softDecision_dataFloat_out
    std::vector<short> bits;                          // bits_dataShort_out
    std::vector<float> phase;                         // phase_dataFloat_out
    std::vector<short> sampleIndex;                   // sampleIndex_dataShort_out
    pskd_sri_out sri;                                 // out-port SRIs (:393-405)
};

class psk_soft_gpu {
public:
    // properties: same names, types and defaults as the generated base class
    unsigned short samplesPerBaud;
    uint32_t numAvg;
    unsigned short constelationSize;
    unsigned short phaseAvg;
    bool differentialDecoding;
    bool resetState;

    explicit psk_soft_gpu(int device = 0) : h_(nullptr) {
        pskd_props p; pskd_default_props(&p);
        samplesPerBaud = p.samplesPerBaud; numAvg = p.numAvg; constelationSize = p.constelationSize;
        phaseAvg = p.phaseAvg; differentialDecoding = p.differentialDecoding != 0; resetState = false;
        psk_check(pskd_create(&h_, device, 1, &p));
    }
    ~psk_soft_gpu() { if (h_) pskd_destroy(h_); }
    psk_soft_gpu(const psk_soft_gpu&) = delete;
    psk_soft_gpu& operator=(const psk_soft_gpu&) = delete;

    // serviceFunction() with the packet passed in instead of pulled from the port.
    // pkt == nullptr plays getPacket() returning no data.
    int serviceFunction(const psk_packet* pkt, psk_outputs& out) {
        if (!pkt) return PSK_NOOP;                                                  // :350-352
        pskd_props p;
        p.samplesPerBaud = samplesPerBaud; p.numAvg = numAvg; p.constelationSize = constelationSize;
        p.phaseAvg = phaseAvg; p.differentialDecoding = differentialDecoding; p.resetState = resetState;
        psk_check(pskd_set_props(h_, 0, &p));       // property snapshot + change listeners (:374-378, :638-651)
        const size_t n = pkt->dataBuffer.size() / 2;                                 // :428 (odd trailing float dropped)
        const size_t cap = pskd_max_symbols(h_, 0, n) + 1;
        out.softDecision.resize(cap); out.phase.resize(cap); out.sampleIndex.resize(cap); out.bits.resize(3 * cap);
        pskd_input in{};
        in.iq = pkt->dataBuffer.data(); in.iq_stride = n; in.n_complex = nullptr; in.n_complex_all = n;
        in.sri_xdelta = pkt->xdelta; in.sri_mode = pkt->mode; in.packet_len = 0;
        in.flags = PSKD_FLAG_HOST_BUFFERS | (pkt->inputQueueFlushed ? PSKD_FLAG_QUEUE_FLUSHED : 0) |
                   (pkt->sriChanged ? PSKD_FLAG_SRI_CHANGED : 0);
        pskd_output o{};
        size_t ns = 0, nb = 0;
        o.soft = reinterpret_cast<float*>(out.softDecision.data()); o.bits = out.bits.data();
        o.phase = out.phase.data(); o.sample_index = out.sampleIndex.data();
        o.sym_stride = cap; o.bits_stride = 3 * cap; o.n_symbols = &ns; o.n_bits = &nb;
        psk_check(pskd_process(h_, &in, &o));
        // resetState stays pending until a COMPLEX packet consumes it (:353-372): read it back, so that a flag
        // raised by a queue flush on an ignored real-data packet (:359-363) survives the next property snapshot
        psk_check(pskd_get_props(h_, 0, &p));
        resetState = p.resetState != 0;
        out.softDecision.resize(ns); out.phase.resize(ns); out.sampleIndex.resize(ns); out.bits.resize(nb);
        pskd_get_sri(h_, 0, &out.sri);
        return PSK_NORMAL;                                                           // :362, :617
    }
    pskd_handle handle() const { return h_; }

private:
    pskd_handle h_;
};

// n_channels independent components on one GPU; buffers are caller-owned (device or pinned host)
class psk_bank_gpu {
public:
    psk_bank_gpu(int device, const std::vector<pskd_props>& props) : h_(nullptr), n_((int)props.size()) {
        psk_check(pskd_create(&h_, device, n_, props.data()));
    }
    ~psk_bank_gpu() { if (h_) pskd_destroy(h_); }
    psk_bank_gpu(const psk_bank_gpu&) = delete;
    psk_bank_gpu& operator=(const psk_bank_gpu&) = delete;
    int channels() const { return n_; }
    int process(const pskd_input& in, pskd_output& out) { int rc = pskd_process(h_, &in, &out); psk_check(rc); return rc; }
    void sync() { psk_check(pskd_sync(h_)); }
    void set_props(int ch, const pskd_props& p) { psk_check(pskd_set_props(h_, ch, &p)); }
    pskd_stats stats() { pskd_stats s; psk_check(pskd_get_stats(h_, &s)); return s; }
    pskd_handle handle() const { return h_; }

private:
    pskd_handle h_;
    int n_;
};

// contiguous channel ranges, sizes differing by at most one (no collective: channels are independent)
inline std::vector<std::pair<int, int> > psk_channel_ranges(int n_channels, int n_gpus) {
    std::vector<std::pair<int, int> > r;
    int base = n_channels / n_gpus, extra = n_channels % n_gpus, lo = 0;
    for (int g = 0; g < n_gpus; g++) { int hi = lo + base + (g < extra ? 1 : 0); r.push_back(std::make_pair(lo, hi)); lo = hi; }
    return r;
}

// a channel bank sharded over the GPUs of one box.  run() calls fn(gpu, bank, lo, hi) on one host
// thread per GPU; each thread drives only its own bank / device.
class psk_box_gpu {
public:
    psk_box_gpu(int n_gpus, const std::vector<pskd_props>& props) : ranges_(psk_channel_ranges((int)props.size(), n_gpus)) {
        for (int g = 0; g < n_gpus; g++) {
            std::vector<pskd_props> sub(props.begin() + ranges_[g].first, props.begin() + ranges_[g].second);
            banks_.push_back(sub.empty() ? nullptr : new psk_bank_gpu(g, sub));
        }
    }
    ~psk_box_gpu() { for (auto* b : banks_) delete b; }
    template <class F> void run(F fn) {
        std::vector<std::thread> th;
        std::vector<std::string> err(banks_.size());
        for (size_t g = 0; g < banks_.size(); g++) {
            if (!banks_[g]) continue;
            th.emplace_back([&, g] {
                try { fn((int)g, *banks_[g], ranges_[g].first, ranges_[g].second); }
                catch (const std::exception& e) { err[g] = e.what(); }
            });
        }
        for (auto& t : th) t.join();
        for (auto& e : err) if (!e.empty()) throw psk_gpu_error(PSKD_ERR_CUDA, e);
    }
    const std::vector<std::pair<int, int> >& ranges() const { return ranges_; }

private:
    std::vector<std::pair<int, int> > ranges_;
    std::vector<psk_bank_gpu*> banks_;
};
