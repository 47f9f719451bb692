// pskd_exact.cuh -- rounding-exact device arithmetic for the phase chain and the slicer.
//
// The reference is scalar x86-64 C++ built with -O2, no FMA contraction, gnu++98
// (SURVEY.md section 3.2).  Every float/double operation below is an explicit round-to-nearest
// intrinsic so nvcc can never contract a*b+c into an FMA; the operation ORDER follows the
// reference line cited beside each function.  Only atan2f / sincosf are library calls (CUDA
// libdevice vs glibc: <= 2 ulp apart, inside the stated 1e-4 tolerance).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PSKD_M_PI    3.14159265358979323846
#define PSKD_M_2PI   (2.0 * PSKD_M_PI)            // cpp/psk_soft.h:65
#define PSKD_M_PI_4  0.78539816339744830962

namespace pskd {

__device__ __forceinline__ float fmulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float faddr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsubr(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double dmulr(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double daddr(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsubr(double a, double b) { return __dsub_rn(a, b); }

// std::norm(complex<float>) = re*re + im*im in float (libstdc++ _Norm_helper), widened to double
// at cpp/psk_soft.cpp:448.
__device__ __forceinline__ float energy_f32(float re, float im) {
    return faddr(fmulr(re, re), fmulr(im, im));
}

// builtin complex<float> multiply (cpp/psk_soft.cpp:474 via pow, :500): unfused
// (a*c - b*d, a*d + b*c); libgcc __mulsc3 only when both parts come out NaN.
static __device__ __noinline__ float2 mulsc3_recover(float a, float b, float c, float d, float x, float y) {
    // C99 Annex G recovery of infinities, as in libgcc2.c __mulsc3
    bool recalc = false;
    if (isinf(a) || isinf(b)) {
        a = copysignf(isinf(a) ? 1.f : 0.f, a);
        b = copysignf(isinf(b) ? 1.f : 0.f, b);
        if (isnan(c)) c = copysignf(0.f, c);
        if (isnan(d)) d = copysignf(0.f, d);
        recalc = true;
    }
    if (isinf(c) || isinf(d)) {
        c = copysignf(isinf(c) ? 1.f : 0.f, c);
        d = copysignf(isinf(d) ? 1.f : 0.f, d);
        if (isnan(a)) a = copysignf(0.f, a);
        if (isnan(b)) b = copysignf(0.f, b);
        recalc = true;
    }
    if (!recalc) {
        float ac = fmulr(a, c), bd = fmulr(b, d), ad = fmulr(a, d), bc = fmulr(b, c);
        if (isinf(ac) || isinf(bd) || isinf(ad) || isinf(bc)) {
            if (isnan(a)) a = copysignf(0.f, a);
            if (isnan(b)) b = copysignf(0.f, b);
            if (isnan(c)) c = copysignf(0.f, c);
            if (isnan(d)) d = copysignf(0.f, d);
            recalc = true;
        }
    }
    if (recalc) {
        x = fmulr(__int_as_float(0x7f800000), fsubr(fmulr(a, c), fmulr(b, d)));
        y = fmulr(__int_as_float(0x7f800000), faddr(fmulr(a, d), fmulr(b, c)));
    }
    return make_float2(x, y);
}

__device__ __forceinline__ float2 cmul_f32(float2 p, float2 q) {
    float x = fsubr(fmulr(p.x, q.x), fmulr(p.y, q.y));
    float y = faddr(fmulr(p.x, q.y), fmulr(p.y, q.x));
    if (isnan(x) && isnan(y)) return mulsc3_recover(p.x, p.y, q.x, q.y, x, y);
    return make_float2(x, y);
}

// std::pow(complex<float>, size_t) under gnu++98 -> libstdc++ __complex_pow_unsigned
// (cpp/psk_soft.cpp:474).
__device__ __forceinline__ float2 cpow_unsigned(float2 x, unsigned n) {
    float2 y = (n & 1u) ? x : make_float2(1.0f, 0.0f);
    while (n >>= 1) {
        x = cmul_f32(x, x);
        if (n & 1u) y = cmul_f32(y, x);
    }
    return y;
}

// libgcc __divsc3 (GCC >= 11: evaluated in double, results narrowed to float), cpp/psk_soft.cpp:488.
__device__ __forceinline__ float2 cdiv_f32(float2 n, float2 dnm) {
    double a = n.x, b = n.y, c = dnm.x, d = dnm.y;
    double denom = daddr(dmulr(c, c), dmulr(d, d));
    float x = __double2float_rn(__ddiv_rn(daddr(dmulr(a, c), dmulr(b, d)), denom));
    float y = __double2float_rn(__ddiv_rn(dsubr(dmulr(b, c), dmulr(a, d)), denom));
    if (isnan(x) && isnan(y)) {
        float af = n.x, bf = n.y, cf = dnm.x, df = dnm.y;
        if (cf == 0.0f && df == 0.0f && (!isnan(af) || !isnan(bf))) {
            float inf = copysignf(__int_as_float(0x7f800000), cf);
            x = fmulr(inf, af);
            y = fmulr(inf, bf);
        } else if ((isinf(af) || isinf(bf)) && isfinite(cf) && isfinite(df)) {
            af = copysignf(isinf(af) ? 1.f : 0.f, af);
            bf = copysignf(isinf(bf) ? 1.f : 0.f, bf);
            x = fmulr(__int_as_float(0x7f800000), faddr(fmulr(af, cf), fmulr(bf, df)));
            y = fmulr(__int_as_float(0x7f800000), fsubr(fmulr(bf, cf), fmulr(af, df)));
        } else if ((isinf(cf) || isinf(df)) && isfinite(af) && isfinite(bf)) {
            cf = copysignf(isinf(cf) ? 1.f : 0.f, cf);
            df = copysignf(isinf(df) ? 1.f : 0.f, df);
            x = fmulr(0.0f, faddr(fmulr(af, cf), fmulr(bf, df)));
            y = fmulr(0.0f, fsubr(fmulr(bf, cf), fmulr(af, df)));
        }
    }
    return make_float2(x, y);
}

// arg(pow(sample, numSyms)) widened to double (cpp/psk_soft.cpp:474)
__device__ __forceinline__ float mth_power_angle(float2 s, unsigned M) {
    float2 z = cpow_unsigned(s, M);
    return atan2f(z.y, z.x);
}

// ---------------------------------------------------------------------------------------------
// LinearFit (cpp/psk_soft.h:33-53, cpp/psk_soft.cpp:35-185): sliding least-squares line over the
// last n floats, evaluated at the newest point.  The y history lives in a ring supplied by the
// caller: element i of the reference's deque is ring[((head + i) % n) * ring_stride].
// ---------------------------------------------------------------------------------------------
struct FitState {
    double ySum, xySum;
    float  xdelta, denominator, xAvg;
    float  m, b;
    int    n;       // window length (phaseAvg)
    int    pts;     // yvals.size()
    int    head;    // ring index of yvals.front()
    int    count;   // next() calls since the last reset
};

// cpp/psk_soft.cpp:176-185 (operation order read off the compiled reference, oracle/README.md)
__device__ __forceinline__ void fit_calc_denominator(FitState& f) {
    int pts = f.pts;
    if (pts <= 1) return;
    double p = (double)(pts - 1);
    double p2 = dmulr(p, p);
    double p3 = dmulr(p2, p);                       // pow(p,3.0): exact for p < 2^17, as is libm's
    double acc = daddr(__ddiv_rn(p3, 3.0), dmulr(p2, 0.5));
    acc = daddr(acc, __ddiv_rn(p, 6.0));
    acc = dsubr(acc, dmulr(dmulr((double)pts, p2), 0.25));
    double xd = (double)f.xdelta;
    f.denominator = __double2float_rn(dmulr(dmulr(xd, xd), acc));
    f.xAvg = fmulr(fmulr(f.xdelta, (float)(pts - 1)), 0.5f);
}

// cpp/psk_soft.cpp:135-174.  `newest` = yvals.back() (only used when pts == 1).
__device__ __forceinline__ float fit_calc_fit(FitState& f, float newest) {
    int pts = f.pts;
    if (pts > 1) {
        float span = fmulr(f.xdelta, (float)(pts - 1));
        float half_span = fmulr(span, 0.5f);                                           // :157
        double num = dsubr(f.xySum, dmulr((double)half_span, f.ySum));
        f.m = __double2float_rn(__ddiv_rn(num, (double)f.denominator));
        f.b = __double2float_rn(dsubr(__ddiv_rn(f.ySum, (double)pts), (double)fmulr(f.m, f.xAvg)));  // :158
        return faddr(fmulr(f.m, span), f.b);                                           // :161-162
    }
    f.m = 0.0f;
    f.b = (pts == 0) ? 0.0f : newest;
    return f.b;
}

// cpp/psk_soft.cpp:110-122: direct re-sum of both sums from the history, count = 0.
template <class Ring>
__device__ __forceinline__ float fit_resum(FitState& f, Ring ring) {
    double ySum = 0.0, xySum = 0.0;
    float newest = 0.0f;
    int idx = f.head;
    for (int j = 0; j < f.pts; j++) {
        float y = ring.get(idx);
        ySum = daddr(ySum, (double)y);
        xySum = daddr(xySum, (double)fmulr(fmulr((float)j, f.xdelta), y));             // :118
        newest = y;
        if (++idx == f.n) idx = 0;
    }
    f.ySum = ySum;
    f.xySum = xySum;
    fit_calc_denominator(f);
    f.count = 0;
    return fit_calc_fit(f, newest);
}

// cpp/psk_soft.cpp:89-109 : optional new rate / forced clear / new window length, then re-sum.
template <class Ring>
__device__ __forceinline__ float fit_reset(FitState& f, Ring ring, const int* numPts, const float* sampleRate,
                                           bool forceHistoryClear) {
    if (sampleRate) {
        float newXdelta = __double2float_rn(__ddiv_rn(1.0, (double)(*sampleRate)));
        if (f.xdelta != newXdelta) {
            f.xdelta = newXdelta;
            forceHistoryClear = true;
        }
    }
    if (forceHistoryClear) { f.pts = 0; f.head = 0; }
    if (numPts && *numPts != f.n) {
        // the ring is addressed modulo n: re-pack the surviving (newest) values for the new n.
        // Only the sequential chain calls this with a changed n and its ring has room for both.
        int newn = *numPts;
        int keep = f.pts < newn ? f.pts : newn;
        int drop = f.pts - keep;
        ring.repack(f.head, f.n, drop, keep, newn);
        f.n = newn; f.pts = keep; f.head = 0;
    }
    return fit_resum(f, ring);
}

// cpp/psk_soft.cpp:48-87
template <class Ring>
__device__ __forceinline__ float fit_next(FitState& f, Ring ring, float yval) {
    if (f.count == 1048576) fit_resum(f, ring);                                        // :51-52
    bool steady = (f.pts == f.n);
    int size = f.pts;
    if (steady) {
        f.ySum = dsubr(f.ySum, (double)ring.get(f.head));                              // :70
        if (++f.head == f.n) f.head = 0;                                               // :71
        size -= 1;
        f.xySum = dsubr(f.xySum, dmulr((double)f.xdelta, f.ySum));                     // :72
    }
    f.ySum = daddr(f.ySum, (double)yval);                                              // :75
    f.xySum = daddr(f.xySum, (double)fmulr(fmulr(yval, (float)size), f.xdelta));       // :78
    int slot = f.head + size;
    if (slot >= f.n) slot -= f.n;
    ring.set(slot, yval);                                                              // :79
    f.pts = size + 1;
    if (!steady) fit_calc_denominator(f);                                              // :81-83
    f.count++;
    return fit_calc_fit(f, yval);
}

// cpp/psk_soft.cpp:126-133
template <class Ring>
__device__ __forceinline__ float fit_subtract_const(FitState& f, Ring ring, float c) {
    int idx = f.head;
    for (int j = 0; j < f.pts; j++) {
        ring.set(idx, fsubr(ring.get(idx), c));
        if (++idx == f.n) idx = 0;
    }
    return fit_resum(f, ring);
}

// cpp/psk_soft.cpp:476-478: unwrap against the previous estimate. Returns y = float(theta + 2*pi*n).
__device__ __forceinline__ float unwrap_against(float est_prev, float theta, long long* n_out) {
    double th = (double)theta;
    double q = __ddiv_rn(dsubr((double)est_prev, th), PSKD_M_2PI);
    double r = round(q);                       // C round(): half away from zero
    long long n = (long long)r;
    if (n_out) *n_out = n;
    return __double2float_rn(daddr(th, dmulr((double)n, PSKD_M_2PI)));
}

// cpp/psk_soft.cpp:595-596: `abs(phaseEstimate) > wrapValue` with ::abs(int) (cvttss2si; abs; cvtsi2ss)
__device__ __forceinline__ bool wrap_needed(float est, float wrapValue) {
    int i;
    if (est >= -2147483648.0f && est < 2147483648.0f) i = __float2int_rz(est);
    else i = (int)0x80000000;                  // what cvttss2si yields for NaN / out of range
    int a = (i == (int)0x80000000) ? i : (i < 0 ? -i : i);
    return (float)a > wrapValue;
}

// ---------------------------------------------------------------------------------------------
// derotate + slice (cpp/psk_soft.cpp:484-566).  `sample` is the timing-selected sample, or the
// differentially decoded one.  Returns the soft decision; bits[0..bpb) get 0/1.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float phase_correction(float est, int M, bool differential) {
    float pc = 0.0f;
    if (!differential) pc = __fdiv_rn(-est, (float)M);                                 // :494
    if (M == 4) pc = __double2float_rn(daddr((double)pc, PSKD_M_PI_4));                // :497-498
    return pc;
}

__device__ __forceinline__ float2 derotate(float2 sample, float pc) {
    float sn, cs;
    sincosf(pc, &sn, &cs);                                                             // :499 std::polar(1.0f, pc)
    return cmul_f32(sample, make_float2(cs, sn));                                      // :500
}

// returns bit j in bit j of the result; nbits via bpb
__device__ __forceinline__ unsigned slice_bits(float2 c, int bpb) {
    if (bpb == 1) return (c.x < 0.0f) ? 1u : 0u;                                       // :512
    if (bpb == 2) {                                                                    // :523-526 (float -> bool, sic)
        unsigned real = (c.x != 0.0f) ? 1u : 0u, imag = (c.y != 0.0f) ? 1u : 0u;
        return (real ^ imag) | ((imag ^ 1u) << 1);
    }
    if (bpb == 3) {                                                                    // :547-563
        float theta = atan2f(c.y, c.x);
        float softsym = __double2float_rn(dmulr(__ddiv_rn((double)theta, PSKD_M_PI), 4.0));
        if (softsym < -0.5f) softsym = faddr(softsym, 8.0f);
        float r = roundf(softsym);
        // (unsigned short) of cvttss2si: NaN / out of range -> 0x80000000 -> low 16 bits 0
        int iv = (r >= -2147483648.0f && r < 2147483648.0f) ? __float2int_rz(r) : (int)0x80000000;
        return (unsigned)iv & 7u;
    }
    return 0u;
}

__device__ __forceinline__ int bits_per_baud(int M) { return M == 2 ? 1 : M == 4 ? 2 : M == 8 ? 3 : 0; }

}  // namespace pskd
