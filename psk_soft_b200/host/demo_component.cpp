// demo_component.cpp -- smallest end-to-end use of the C++ host mirror: one component, packets of
// a synthetic QPSK stream pushed through serviceFunction(), symbol counts printed.
// Built by `python -m psk_soft_b200._build` into psk_soft_b200/lib/demo_component; run on a GPU box.
#include <cmath>
#include <cstdio>
#include "psk_soft_gpu.hpp"

int main() {
    try {
        psk_soft_gpu comp(0);
        comp.samplesPerBaud = 8; comp.constelationSize = 4; comp.numAvg = 100;
        unsigned lcg = 12345u;
        size_t total = 0;
        psk_outputs out;
        for (int pk = 0; pk < 4; pk++) {
            psk_packet p;
            p.xdelta = 0.01; p.mode = 1; p.sriChanged = (pk == 0);
            p.dataBuffer.resize(2 * 16000);
            for (int i = 0; i < 16000; i++) {
                if (i % 8 == 0) lcg = lcg * 1664525u + 1013904223u;
                double ph = 1.5707963267948966 * ((lcg >> 30) & 3) + 0.3;
                double env = 0.6 + 0.4 * std::sin(3.141592653589793 * ((i % 8) + 0.5) / 8);
                p.dataBuffer[2 * i] = (float)(env * std::cos(ph));
                p.dataBuffer[2 * i + 1] = (float)(env * std::sin(ph));
            }
            int rc = comp.serviceFunction(&p, out);
            total += out.softDecision.size();
            std::printf("packet %d: rc=%d symbols=%zu bits=%zu soft_xdelta=%g\n", pk, rc, out.softDecision.size(), out.bits.size(), out.sri.soft_xdelta);
        }
        std::printf("total symbols %zu (expected %d)\n", total, 64000 / 8 - 99);
        return total == 64000 / 8 - 99 ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
