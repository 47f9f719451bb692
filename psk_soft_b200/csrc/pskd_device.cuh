// pskd_device.cuh -- device helpers shared by the staged kernels (pskd_kernels.cu) and the fused
// kernel (pskd_fused.cu): the virtual input stream, the packet prologue/epilogue of the phase
// estimator, the fast (guarded) forms of the unwrap count, the fit evaluation and the 8-PSK slicer.
#pragma once
#include "pskd_internal.h"

namespace pskd {

// virtual stream = tail ++ in
struct VStream {
    const float2* tail; const float2* in; long long tail_len;
    __device__ __forceinline__ float2 at(long long i) const {
        return (i < tail_len) ? tail[i] : __ldg(in + (i - tail_len));
    }
};


__device__ __forceinline__ float mth_power_angle_fast(float2 s, unsigned M) {
    // unchecked squarings; a (NaN,NaN) anywhere propagates to the end, and only then can the
    // reference have gone through __mulsc3 -> redo with the checked multiply.
    float2 x = s, y = (M & 1u) ? s : make_float2(1.0f, 0.0f);
    unsigned n = M;
    while (n >>= 1) {
        x = make_float2(fsubr(fmulr(x.x, x.x), fmulr(x.y, x.y)), faddr(fmulr(x.x, x.y), fmulr(x.y, x.x)));
        if (n & 1u) y = make_float2(fsubr(fmulr(y.x, x.x), fmulr(y.y, x.y)), faddr(fmulr(y.x, x.y), fmulr(y.y, x.x)));
    }
    if (isnan(y.x) && isnan(y.y)) y = cpow_unsigned(s, M);
    return atan2f(y.y, y.x);
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
    // :393 -- the SRI block runs when sriChanged, resetNumSymbols or resetSamplesPerBaud; the last one holds at every
    // packet start except while the component is stalled on an over-full window (:380-383)
    const bool sri_block = !(flags & CH_STALLED) || (flags & (CH_SRI_CHANGED | CH_RESET_NUMSYMS));
    if (sri_block && sri_xdelta != (double)st.sampleRate) {                            // :394-398
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


struct SmemRing {
    float* base;
    __device__ __forceinline__ float get(int i) const { return base[i]; }
    __device__ __forceinline__ void set(int i, float v) const { base[i] = v; }
    __device__ void repack(int, int, int, int, int) const {}   // never called: P changes take the sequential chain
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

// the reference's unwrap count for one symbol (cpp/psk_soft.cpp:477): round((est - theta)/2pi),
// C round().  Fast form: multiply by 1/2pi and round to nearest; whenever the quotient is within
// 1e-7 of a half-integer (where the division's last bit or the tie rule could matter) the
// literal division + round() decides.
__device__ __forceinline__ int unwrap_count(float est_prev, float theta) {
    const double dlt = dsubr((double)est_prev, (double)theta);       // exact
    const double q = dmulr(dlt, 0.15915494309189535);
    const double t = daddr(q, 6755399441055744.0);                   // 1.5 * 2^52: round to nearest integer
    const double qr = dsubr(t, 6755399441055744.0);
    const double fr = fabs(dsubr(q, qr));
    if (fr > 0.4999999 || !(fabs(q) < 1.0e9))
        return (int)(long long)round(__ddiv_rn(dlt, PSKD_M_2PI));
    return __double2loint(t);
}

// constants of calculateFit for a full window (cpp/psk_soft.cpp:153-162)
struct FitConst {
    double half_span_d, rden, rpts;
    float span, xAvg;
};
__device__ __forceinline__ FitConst fit_const(const FitState& f) {
    FitConst c;
    c.span = fmulr(f.xdelta, (float)(f.pts - 1));
    c.half_span_d = (double)fmulr(c.span, 0.5f);
    c.rden = __ddiv_rn(1.0, (double)f.denominator);
    c.rpts = __ddiv_rn(1.0, (double)f.pts);
    c.xAvg = f.xAvg;
    return c;
}
__device__ __forceinline__ float fit_eval_fast(const FitConst& c, double ySum, double xySum, float* m_out, float* b_out) {
    const double num = dsubr(xySum, dmulr(c.half_span_d, ySum));
    const float m = __double2float_rn(dmulr(num, c.rden));                                        // :157
    const float b = __double2float_rn(dsubr(dmulr(ySum, c.rpts), (double)fmulr(m, c.xAvg)));      // :158
    if (m_out) { *m_out = m; *b_out = b; }
    return faddr(fmulr(m, c.span), b);                                                            // :161-162
}

// 8-PSK slicer (cpp/psk_soft.cpp:547-563) without atan2f: sym = round(angle/(pi/4)) mod 8 is a
// sector test against the rays at odd multiples of pi/8.  Within a guard band of the rays (or
// for non-finite / zero input) the literal atan2f path decides.
__device__ __forceinline__ unsigned slice8_fast(float2 c) {
    const float a = fabsf(c.x), b = fabsf(c.y);
    const float T = 0.41421356237309503f;       // tan(pi/8)
    const float sum = a + b;
    const float d1 = b - T * a, d2 = a - T * b;
    const float g = 1.0e-5f * sum;
    if (!(sum > 0.0f) || !(sum < 3.0e38f) || fabsf(d1) <= g || fabsf(d2) <= g) return slice_bits(c, 3);
    if (d1 < 0.0f) return (c.x > 0.0f) ? 0u : 4u;
    if (d2 < 0.0f) return (c.y > 0.0f) ? 2u : 6u;
    return (c.x > 0.0f) ? ((c.y > 0.0f) ? 1u : 7u) : ((c.y > 0.0f) ? 3u : 5u);
}


}  // namespace pskd
