/* TEST INFRASTRUCTURE ONLY -- C-callable harness around the UNMODIFIED reference
 * component class psk_soft_i (compiled in place from /root/reference/cpp by
 * oracle/Makefile against oracle/stubs/).  It plays the role of the REDHAWK
 * framework for exactly one code path: configure properties, queue one BULKIO
 * packet on dataFloat_in, call psk_soft_i::serviceFunction() (psk_soft.cpp:346),
 * and read back what the component pushed on its four out-ports
 * (psk_soft.cpp:605-615).  Nothing under psk_soft_b200/ may link or load this.
 */
#include "psk_soft.h"

#include <cstring>
#include <cstddef>

namespace {

class ref_component : public psk_soft_i {
public:
    ref_component() : psk_soft_i("oracle", "oracle") { constructor(); }

    /* what CORBA configure() + the property-change machinery do: assign, then
     * call the registered listener if the value changed */
    int configure(const char* id, double v) {
        std::string s(id);
        if (s == "samplesPerBaud")        { unsigned short n = (unsigned short)v; bool ch = n != samplesPerBaud; samplesPerBaud = n; if (ch) oracle_fire_listener(s); }
        else if (s == "numAvg")           { numAvg = (CORBA::ULong)v; }
        else if (s == "constelationSize") { unsigned short n = (unsigned short)v; bool ch = n != constelationSize; constelationSize = n; if (ch) oracle_fire_listener(s); }
        else if (s == "phaseAvg")         { unsigned short n = (unsigned short)v; bool ch = n != phaseAvg; phaseAvg = n; if (ch) oracle_fire_listener(s); }
        else if (s == "differentialDecoding") { differentialDecoding = (v != 0); }
        else if (s == "resetState")       { resetState = (v != 0); }
        else return -1;
        return 0;
    }
    double query(const char* id) {
        std::string s(id);
        if (s == "samplesPerBaud") return samplesPerBaud;
        if (s == "numAvg") return numAvg;
        if (s == "constelationSize") return constelationSize;
        if (s == "phaseAvg") return phaseAvg;
        if (s == "differentialDecoding") return differentialDecoding;
        if (s == "resetState") return resetState;
        return -1;
    }
    int push(const float* iq, size_t n_floats, double xdelta, int mode, int flushed, int sri_changed) {
        bulkio::InFloatPort::dataTransfer* p = new bulkio::InFloatPort::dataTransfer();
        p->dataBuffer.assign(iq, iq + n_floats);
        p->SRI.xdelta = xdelta;
        p->SRI.mode = (short)mode;
        p->sriChanged = sri_changed != 0;
        p->inputQueueFlushed = flushed != 0;
        dataFloat_in->oracle_enqueue(p);
        int rc = serviceFunction();
        if (mode != 1) delete p;   /* psk_soft.cpp:359-363 returns without releasing the packet */
        return rc;
    }
    bulkio::OutFloatPort* soft()  { return softDecision_dataFloat_out; }
    bulkio::OutShortPort* bits()  { return bits_dataShort_out; }
    bulkio::OutFloatPort* phase() { return phase_dataFloat_out; }
    bulkio::OutShortPort* sidx()  { return sampleIndex_dataShort_out; }
};

template <class V> size_t drain(V& v, void* dst, size_t cap_elems) {
    size_t n = v.size();
    if (dst) {
        size_t m = n < cap_elems ? n : cap_elems;
        if (m) std::memcpy(dst, &v[0], m * sizeof(v[0]));
        v.clear();
    }
    return n;
}

}  // namespace

extern "C" {

void* ref_create(void) { return new ref_component(); }
void  ref_destroy(void* h) { delete static_cast<ref_component*>(h); }
int   ref_configure(void* h, const char* id, double v) { return static_cast<ref_component*>(h)->configure(id, v); }
double ref_query(void* h, const char* id) { return static_cast<ref_component*>(h)->query(id); }

/* one packet: n_floats interleaved re,im floats. returns serviceFunction()'s value */
int ref_push(void* h, const float* iq, size_t n_floats, double xdelta, int mode, int flushed, int sri_changed) {
    return static_cast<ref_component*>(h)->push(iq, n_floats, xdelta, mode, flushed, sri_changed);
}

/* port: 0 soft (floats, 2 per symbol), 1 bits (short), 2 phase (float), 3 sampleIndex (short).
 * dst==NULL: return number of elements waiting. else copy up to cap elements, clear, return count. */
size_t ref_read(void* h, int port, void* dst, size_t cap_elems) {
    ref_component* c = static_cast<ref_component*>(h);
    switch (port) {
        case 0: return drain(c->soft()->data, dst, cap_elems);
        case 1: return drain(c->bits()->data, dst, cap_elems);
        case 2: return drain(c->phase()->data, dst, cap_elems);
        case 3: return drain(c->sidx()->data, dst, cap_elems);
    }
    return 0;
}
/* out-SRI bookkeeping (psk_soft.cpp:399-404): count of pushSRI calls and last xdelta/mode per port */
long   ref_sri_count(void* h, int port) {
    ref_component* c = static_cast<ref_component*>(h);
    switch (port) { case 0: return c->soft()->n_sri; case 1: return c->bits()->n_sri; case 2: return c->phase()->n_sri; case 3: return c->sidx()->n_sri; }
    return -1;
}
double ref_sri_xdelta(void* h, int port) {
    ref_component* c = static_cast<ref_component*>(h);
    switch (port) { case 0: return c->soft()->last_sri.xdelta; case 1: return c->bits()->last_sri.xdelta; case 2: return c->phase()->last_sri.xdelta; case 3: return c->sidx()->last_sri.xdelta; }
    return -1;
}
int ref_sri_mode(void* h, int port) {
    ref_component* c = static_cast<ref_component*>(h);
    switch (port) { case 0: return c->soft()->last_sri.mode; case 1: return c->bits()->last_sri.mode; case 2: return c->phase()->last_sri.mode; case 3: return c->sidx()->last_sri.mode; }
    return -1;
}
long ref_packet_count(void* h, int port) {
    ref_component* c = static_cast<ref_component*>(h);
    switch (port) { case 0: return c->soft()->n_packets; case 1: return c->bits()->n_packets; case 2: return c->phase()->n_packets; case 3: return c->sidx()->n_packets; }
    return -1;
}

/* Whole-stream convenience used by the timing harness: push `n_complex` samples in packets of
 * `packet_len` complex samples and count outputs (outputs are drained and discarded unless
 * buffers are given).  Returns number of symbols produced. */
size_t ref_demod(void* h, const float* iq, size_t n_complex, size_t packet_len, double xdelta,
                 float* soft, short* bits, float* phase, short* sidx,
                 size_t cap_syms, size_t cap_bits, size_t* n_bits_out) {
    ref_component* c = static_cast<ref_component*>(h);
    size_t ns = 0, nb = 0;
    for (size_t off = 0; off < n_complex; off += packet_len) {
        size_t m = n_complex - off < packet_len ? n_complex - off : packet_len;
        c->push(iq + 2 * off, 2 * m, xdelta, 1, 0, off == 0);
        size_t k = c->phase()->data.size();
        size_t b = c->bits()->data.size();
        if (soft && ns + k <= cap_syms)  { if (k) std::memcpy(soft + 2 * ns, &c->soft()->data[0], k * 2 * sizeof(float)); }
        if (phase && ns + k <= cap_syms) { if (k) std::memcpy(phase + ns, &c->phase()->data[0], k * sizeof(float)); }
        if (sidx && ns + k <= cap_syms)  { size_t ki = c->sidx()->data.size(); if (ki) std::memcpy(sidx + ns, &c->sidx()->data[0], ki * sizeof(short)); }
        if (bits && nb + b <= cap_bits)  { if (b) std::memcpy(bits + nb, &c->bits()->data[0], b * sizeof(short)); }
        c->soft()->data.clear(); c->phase()->data.clear(); c->sidx()->data.clear(); c->bits()->data.clear();
        ns += k; nb += b;
    }
    if (n_bits_out) *n_bits_out = nb;
    return ns;
}

}  // extern "C"
