// Time-parallel chain plan, device side (see TpCtl in pskd_internal.h): the per-packet scan and the per-channel
// resolve as device functions, shared by the stand-alone kernels (k_tp_scan, k_tp_resolve; pskd_kernels.cu) and by
// the task kernel k_fzs_uni (pskd_fused.cu), which runs them between a channel's front units and its chain units
// INSIDE one launch.  Everything another warp of the same launch may have produced (the angle scratch, the packet
// records) is therefore read with ld.global.cg: L2 is the point of coherence, a line another SM wrote after this SM
// cached its neighbour must not be served from L1.
#pragma once
#include "pskd_internal.h"
#include "pskd_device.cuh"

namespace pskd {

// classic sample-to-sample unwrap increment (the chain's prediction rule; also the time-parallel
// chain's integer bookkeeping -- both must use the same arithmetic)
__device__ __forceinline__ int classic_dn(float th, float th_prev) {
    return -__float2int_rn((th - th_prev) * 0.15915494309189535f);
}

__device__ __forceinline__ TpPacket tp_ld_pkt(const TpPacket* p) {
    static_assert(sizeof(TpPacket) == 40, "TpPacket is read as five 64-bit words");
    union { TpPacket t; long long v[5]; } u;
    const long long* q = reinterpret_cast<const long long*>(p);
#pragma unroll
    for (int i = 0; i < 5; i++) u.v[i] = __ldcg(q + i);
    return u.t;
}

// one warp, one (channel, packet): classic unwrap over the packet (integer scan) and the linear-fit estimate of
// the relative phases at the packet end (double, regression over the last P)
__device__ __forceinline__ void tp_scan_item(const ChanDesc* __restrict__ desc, const float* __restrict__ theta, const TpItem it,
                                             TpPacket* __restrict__ pkts, const int lane)
{
    const int slot = it.pkt_slot;
    const ChanDesc& d = desc[it.ch];
    const int pkt = it.pk_a;
    const int K = (int)d.K, P = d.P;
    const int klo = (int)first_symbol_at((long long)pkt * d.pkt_len, d.tail_len, d.S, d.A, K);
    const int khi = (pkt + 1 == d.n_pkts) ? K : (int)first_symbol_at((long long)(pkt + 1) * d.pkt_len, d.tail_len, d.S, d.A, K);
    const float* thg = theta + d.scr_off;
    int c = 0;                                             // count at the last symbol processed so far
    {
        int acc = 0;                                       // per-lane partial sums, one warp reduction at the end
        // head up to the first 16-byte aligned angle, then 4 angles per lane and load (the scratch rows are shifted so that
        // packets start aligned, pskd_api.cu), then the tail
        int m0 = klo + 1;
        const int a0 = min(khi, (int)(m0 + ((4 - ((reinterpret_cast<uintptr_t>(thg + m0) >> 2) & 3)) & 3)));
        if (m0 + lane < a0) acc += classic_dn(__ldcg(thg + m0 + lane), __ldcg(thg + m0 + lane - 1));
        m0 = a0;
        const int n4 = (khi - m0) >> 2;                    // whole float4 groups
        for (int g = lane; g < n4; g += 32) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(thg + m0) + g);
            const float pv = __ldcg(thg + m0 + 4 * g - 1);
            acc += classic_dn(t.x, pv) + classic_dn(t.y, t.x) + classic_dn(t.z, t.y) + classic_dn(t.w, t.z);
        }
        for (int base = m0 + 4 * n4; base < khi; base += 32) {
            const int m = base + lane;
            if (m < khi) acc += classic_dn(__ldcg(thg + m), __ldcg(thg + m - 1));
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        c = acc;
    }
    // regression of phi = theta + 2pi*c over the last min(P, n) symbols, evaluated at the newest one
    const int n = khi - klo, np = min(P, n);
    double s0 = 0.0, s1 = 0.0, q2 = 0.0;
    {
        int carry = 0;
        const double mid = 0.5 * (double)(np - 1);
        for (int e = khi; e > khi - np; e -= 32) {
            const int m = e - 1 - lane;
            const bool in = m >= khi - np;
            int dn = 0; float t = 0.0f;
            if (in) { t = __ldcg(thg + m); if (m > klo) dn = classic_dn(t, __ldcg(thg + m - 1)); }
            const int incl = warp_scan_int(dn, lane);
            if (in) {
                const int cm = c - (carry + incl - dn);
                const double phi = (double)t + PSKD_M_2PI * (double)cm;
                const double x = (double)(m - (khi - np)) - mid;
                s0 += phi; s1 += x * phi; q2 += x * x;
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o);
        }
    }
    if (lane == 0) {
        TpPacket p;
        p.cEnd = c;
        p.dLink = (klo > 0 && n > 0) ? classic_dn(__ldcg(thg + klo), __ldcg(thg + klo - 1)) : 0;
        p.A = 0; p.w = 0;
        p.estRelEnd = (np > 0) ? s0 / (double)np + ((q2 > 0.0) ? (s1 / q2) * 0.5 * (double)(np - 1) : 0.0) : 0.0;
        p.klo = klo; p.khi = khi;
        pkts[slot] = p;
    }
}

// one warp, one channel: the scalar recurrence over its packets: unwrap level of every packet's first symbol and
// the wrap count of every packet end (cpp/psk_soft.cpp:592-603).  The packet records are read 32 at a time (one
// per lane); the recurrence walks them by shuffles.
__device__ __forceinline__ void tp_resolve_chan(const ChanDesc* __restrict__ desc, const float* __restrict__ theta, const TpChan tc,
                                                TpPacket* __restrict__ pkts, const TpEnd* __restrict__ ends,
                                                const ChanState* __restrict__ state, const int lane)
{
    const ChanDesc& d = desc[tc.ch];
    const float* thg = theta + d.scr_off;
    const float wrapValue = __double2float_rn(dmulr(PSKD_M_2PI, (double)d.M));
    const int np = tc.n_pkts - tc.pkt0;
    const int M = d.M;
    // level of the first time-parallel packet: the reference's rule with the exact estimate the head left
    // (no head: the channel's carried estimate)
    long long A = 0;
    {
        const TpPacket p0 = tp_ld_pkt(pkts + tc.first_slot);
        const float est0 = tc.has_head ? ends[tc.first_item].st.est : state[tc.ch].est;
        if (p0.khi > p0.klo) { long long n = 0; (void)unwrap_against(est0, __ldcg(thg + p0.klo), &n); A = n; }
    }
    // |trunc(est)| > wrapValue (the integer abs of cpp/psk_soft.cpp:596)  <=>  |est| >= floor(wrapValue) + 1; the
    // estimates here are regression values (not the exact chain's), every decision is proven later by k_tp_check
    const float wrapThr = floorf(wrapValue) + 1.0f, rWrap = 1.0f / wrapValue;
    double Ad = (double)A;
    for (int base = 0; base < np; base += 32) {
        const int j = base + lane;
        int step = 0, has = 0; double er = 0.0;
        if (j < np) {
            const TpPacket p = tp_ld_pkt(pkts + tc.first_slot + j);
            step = p.cEnd; er = p.estRelEnd; has = p.khi > p.klo;
            if (j + 1 < np) step += tp_ld_pkt(pkts + tc.first_slot + j + 1).dLink;
        }
        int myA = 0, myW = 0;
        const int cnt = min(32, np - base);
        // uniform loop: every lane follows the same recurrence.  Only the level Ad is carried from step to step; the
        // shuffles do not depend on it, so unrolling lets them run ahead of the dependent DFMA -> F2F -> compare chain
#pragma unroll 8
        for (int l = 0; l < 32; l++) {
            const int s_l = __shfl_sync(0xffffffffu, step, l), h_l = __shfl_sync(0xffffffffu, has, l);
            const double e_l = __shfl_sync(0xffffffffu, er, l);
            if (l < cnt) {
                const float est_end = (float)fma(PSKD_M_2PI, Ad, e_l);
                int w = 0;
                if (h_l && fabsf(est_end) >= wrapThr) w = (int)rintf(est_end * rWrap);
                if (l == lane) { myA = (int)Ad; myW = w; }
                Ad += (double)(s_l - M * w);
            }
        }
        if (j < np) { pkts[tc.first_slot + j].A = myA; pkts[tc.first_slot + j].w = myW; }
    }
}

}  // namespace pskd
