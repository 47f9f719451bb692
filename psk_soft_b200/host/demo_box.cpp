// demo_box.cpp -- one channel bank sharded over ALL visible GPUs of the box by the C++ host mirror
// (psk_box_gpu: contiguous channel ranges, one host thread + one pskd bank per GPU, no collective; the
// reference equivalent is one psk_soft_i instance per channel, cpp/psk_soft.h:66-86).  Prints the symbol
// count of every GPU's shard and checks sampled channels against a single-GPU, single-channel run.
// Built by `python -m psk_soft_b200._build` into psk_soft_b200/lib/demo_box; run on a GPU box:
//     demo_box [n_channels=1024] [n_samples=40000] [n_gpus=all]
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "psk_soft_gpu.hpp"

static void synth_channel(float* iq, size_t n, int ch, int S, int M) {
    unsigned lcg = 12345u + 7919u * (unsigned)ch;
    const double f = 2e-5 * ((ch % 11) / 5.0 - 1.0), ph0 = 0.37 * ch;
    unsigned sym = 0;
    for (size_t i = 0; i < n; i++) {
        const size_t pos = i + (size_t)(ch % S);
        if (pos % S == 0 || i == 0) { lcg = lcg * 1664525u + 1013904223u; sym = (lcg >> 24) % (unsigned)M; }
        lcg = lcg * 1664525u + 1013904223u;
        const double nz_r = ((lcg >> 8) & 0xffff) / 65536.0 - 0.5;
        lcg = lcg * 1664525u + 1013904223u;
        const double nz_i = ((lcg >> 8) & 0xffff) / 65536.0 - 0.5;
        const double ph = 6.283185307179586 * ((double)sym / M + f * (double)i) + ph0;
        const double env = 0.6 + 0.4 * std::sin(3.141592653589793 * ((pos % S) + 0.5) / S);
        iq[2 * i] = (float)(env * std::cos(ph) + 0.04 * nz_r);
        iq[2 * i + 1] = (float)(env * std::sin(ph) + 0.04 * nz_i);
    }
}

int main(int argc, char** argv) {
    try {
        const int nch = argc > 1 ? std::atoi(argv[1]) : 1024;
        const size_t n = argc > 2 ? (size_t)std::atol(argv[2]) : 40000;
        int n_gpus = pskd_device_count();
        if (argc > 3) n_gpus = std::min(n_gpus, std::atoi(argv[3]));
        if (n_gpus < 1) { std::fprintf(stderr, "no CUDA device (there is no CPU path)\n"); return 3; }
        // a mixed bank: three samples-per-symbol classes, both staged and (with enough channels per GPU) fused kernels
        std::vector<pskd_props> props((size_t)nch);
        for (int c = 0; c < nch; c++) {
            pskd_default_props(&props[c]);
            props[c].samplesPerBaud = (c % 3 == 0) ? 8 : (c % 3 == 1) ? 10 : 16;
            props[c].constelationSize = (c % 2) ? 8 : 4;
            props[c].numAvg = 100; props[c].phaseAvg = (c % 5 == 0) ? 100 : 50;
            props[c].differentialDecoding = (c % 7 == 0);
        }
        const size_t cap = n / 8 + 8;
        std::vector<float> iq((size_t)nch * n * 2), soft((size_t)nch * cap * 2), phase((size_t)nch * cap);
        std::vector<short> sidx((size_t)nch * cap), bits((size_t)nch * cap * 3);
        std::vector<size_t> nsym((size_t)nch), nbits((size_t)nch);
        for (int c = 0; c < nch; c++) synth_channel(&iq[(size_t)c * n * 2], n, c, props[c].samplesPerBaud, props[c].constelationSize);

        psk_box_gpu box(n_gpus, props);
        std::vector<unsigned long long> per_gpu((size_t)n_gpus, 0);
        box.run([&](int g, psk_bank_gpu& bank, int lo, int hi) {
            pskd_input in; std::memset(&in, 0, sizeof(in));
            in.iq = &iq[(size_t)lo * n * 2]; in.iq_stride = n; in.n_complex_all = n;
            in.sri_xdelta = 0.01; in.sri_mode = 1; in.packet_len = 8000; in.flags = PSKD_FLAG_HOST_BUFFERS;
            pskd_output out; std::memset(&out, 0, sizeof(out));
            out.soft = &soft[(size_t)lo * cap * 2]; out.phase = &phase[(size_t)lo * cap];
            out.sample_index = &sidx[(size_t)lo * cap]; out.bits = &bits[(size_t)lo * cap * 3];
            out.sym_stride = cap; out.bits_stride = cap * 3; out.n_symbols = &nsym[lo]; out.n_bits = &nbits[lo];
            bank.process(in, out);
            for (int c = lo; c < hi; c++) per_gpu[g] += nsym[c];
        });
        unsigned long long total = 0, expect = 0;
        for (int g = 0; g < n_gpus; g++) {
            std::printf("gpu %d: channels [%d, %d) -> %llu symbols\n", g, box.ranges()[g].first, box.ranges()[g].second, per_gpu[g]);
            total += per_gpu[g];
        }
        for (int c = 0; c < nch; c++) expect += n / props[c].samplesPerBaud - props[c].numAvg + 1;
        std::printf("total symbols %llu (expected %llu) on %d GPU(s)\n", total, expect, n_gpus);
        if (total != expect) return 1;

        // sampled channels (first / last of every shard) against a single-channel bank on GPU 0
        int checked = 0;
        for (int g = 0; g < n_gpus; g++) {
            const int picks[2] = {box.ranges()[g].first, box.ranges()[g].second - 1};
            for (int k = 0; k < 2; k++) {
                const int c = picks[k];
                if (c < box.ranges()[g].first || c >= box.ranges()[g].second) continue;
                std::vector<pskd_props> one(1, props[c]);
                psk_bank_gpu single(0, one);
                std::vector<float> s1(cap * 2), p1(cap);
                std::vector<short> i1(cap), b1(cap * 3);
                size_t ns1 = 0, nb1 = 0;
                pskd_input in; std::memset(&in, 0, sizeof(in));
                in.iq = &iq[(size_t)c * n * 2]; in.iq_stride = n; in.n_complex_all = n;
                in.sri_xdelta = 0.01; in.sri_mode = 1; in.packet_len = 8000; in.flags = PSKD_FLAG_HOST_BUFFERS;
                pskd_output out; std::memset(&out, 0, sizeof(out));
                out.soft = s1.data(); out.phase = p1.data(); out.sample_index = i1.data(); out.bits = b1.data();
                out.sym_stride = cap; out.bits_stride = cap * 3; out.n_symbols = &ns1; out.n_bits = &nb1;
                single.process(in, out);
                if (ns1 != nsym[c] || nb1 != nbits[c]) { std::printf("channel %d: counts differ\n", c); return 1; }
                if (std::memcmp(i1.data(), &sidx[(size_t)c * cap], ns1 * sizeof(short)) || std::memcmp(b1.data(), &bits[(size_t)c * cap * 3], nb1 * sizeof(short))) {
                    std::printf("channel %d: sampleIndex / bits differ between the shard and the single-channel run\n", c); return 1;
                }
                const size_t k0 = props[c].differentialDecoding ? 1 : 0;      // first differential symbol: inf/NaN
                for (size_t i = k0; i < ns1; i++) {
                    const float a = phase[(size_t)c * cap + i], b = p1[i];
                    if (!(std::fabs(a - b) <= 1e-4f * std::fmax(1.0f, std::fabs(b)))) { std::printf("channel %d: phase differs at %zu\n", c, i); return 1; }
                    for (int q = 0; q < 2; q++) {
                        const float x = soft[((size_t)c * cap + i) * 2 + q], y = s1[i * 2 + q];
                        if (!(std::fabs(x - y) <= 1e-4f * std::fmax(1.0f, std::fabs(y)))) { std::printf("channel %d: soft differs at %zu\n", c, i); return 1; }
                    }
                }
                checked++;
            }
        }
        std::printf("box ok: %d sampled channels identical (bits, sampleIndex) / within 1e-4 (phase, soft) to single-channel runs\n", checked);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
