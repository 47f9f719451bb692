/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the PSK soft-demod hot path.
 *
 * A plain-C, scalar, sequential RESTATEMENT of the algorithm of the REDHAWK component
 * rh.psk_soft (reference: cpp/psk_soft.cpp:35-185 LinearFit, :346-618 serviceFunction,
 * :619-651 resyncEnergy + property listeners; state: cpp/psk_soft.h:33-86).  Written from the
 * reference's behaviour, one rounding at a time; every function cites the lines it follows.
 * It is the checker the CUDA path is compared against -- only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (psk_soft_b200/) never links, loads or falls back to this file.
 *
 * PARITY PINNING: this restatement is pinned bit-for-bit (soft, bits, phase, sampleIndex)
 * against the UNMODIFIED reference compiled in place (oracle/_ref/libpsk_ref.so, see
 * oracle/Makefile) by tests/test_oracle_cpu.py in the build container and against the
 * committed golden vectors tests/golden/ (npz files) (generated from that reference build by
 * tests/golden/make_golden.py) everywhere else.  The reference itself ships no golden
 * vectors; its own test (tests/test_psk_soft.py:178-238) only bounds the soft-symbol error
 * by 1e-3, which tests/test_oracle_cpu.py::test_reference_own_assertions_hold re-creates.
 *
 * Arithmetic notes (all verified against the reference build, see oracle/README.md):
 *  - float expressions are evaluated in float (x86-64 SSE, FLT_EVAL_METHOD 0), no FMA
 *    contraction (compile with -ffp-contract=off, no -march);
 *  - complex<float> pow(z, size_t) under gnu++98 is libstdc++'s __complex_pow_unsigned
 *    (binary exponentiation with builtin complex multiplies), complex division is libgcc's
 *    __divsc3, complex multiply is the inline formula with the __mulsc3 NaN fallback --
 *    C99 `float _Complex` arithmetic compiles to exactly the same sequences;
 *  - `abs(phaseEstimate)` at psk_soft.cpp:596 is ::abs(int).
 */
#define _GNU_SOURCE 1 /* sincosf */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

#define ORC_M_2PI (2 * M_PI) /* cpp/psk_soft.h:65 */

/* ---------------------------------------------------------------------------------------
 * tiny growable containers standing in for std::deque / std::vector
 * ------------------------------------------------------------------------------------- */
typedef struct { float* v; size_t head, size, cap; } fdeque;
typedef struct { double* v; size_t head, size, cap; } ddeque;
typedef struct { float complex* v; size_t head, size, cap; } cdeque;

#define DEQ_PUSH(T, d, x)                                                        \
    do {                                                                         \
        if ((d)->head + (d)->size == (d)->cap) {                                 \
            if ((d)->head > (d)->cap / 2) {                                      \
                memmove((d)->v, (d)->v + (d)->head, (d)->size * sizeof(T));      \
                (d)->head = 0;                                                   \
            } else {                                                             \
                (d)->cap = (d)->cap ? 2 * (d)->cap : 1024;                       \
                (d)->v = (T*)realloc((d)->v, (d)->cap * sizeof(T));              \
            }                                                                    \
        }                                                                        \
        (d)->v[(d)->head + (d)->size++] = (x);                                   \
    } while (0)
#define DEQ_AT(d, i) ((d)->v[(d)->head + (i)])
#define DEQ_POP_FRONT(d, n) do { (d)->head += (n); (d)->size -= (n); if ((d)->size == 0) (d)->head = 0; } while (0)
#define DEQ_TRUNC(d, n) do { if ((d)->size > (n)) (d)->size = (n); } while (0)

typedef struct { void* v; size_t size, cap, esz; } vec;
static void vec_push(vec* a, const void* x) {
    if (a->size == a->cap) {
        a->cap = a->cap ? 2 * a->cap : 4096;
        a->v = realloc(a->v, a->cap * a->esz);
    }
    memcpy((char*)a->v + a->size * a->esz, x, a->esz);
    a->size++;
}

/* ---------------------------------------------------------------------------------------
 * LinearFit (cpp/psk_soft.h:33-53, cpp/psk_soft.cpp:35-185)
 * ------------------------------------------------------------------------------------- */
typedef struct {
    fdeque yvals;
    float m, b;
    double ySum, xySum;
    size_t n;
    float xdelta, denominator, xAvg;
    size_t count;
} linfit;

/* cpp/psk_soft.cpp:35-46 */
static void linfit_init(linfit* f, size_t numPts, float sampleRate) {
    memset(f, 0, sizeof(*f));
    f->n = numPts;
    f->xdelta = (float)(1.0 / (double)sampleRate);
    f->denominator = 1.0f;
}

/* cpp/psk_soft.cpp:176-185.  pow(float,2) and pow(size_t,k) promote to double; the
 * compiled reference evaluates p*p in double, calls pow(p,3.0), turns /2.0 and /4.0 into
 * exact *0.5 and *0.25, keeps /3.0 and /6.0 as divisions. */
static void linfit_calc_denominator(linfit* f) {
    size_t pts = f->yvals.size;
    if (pts <= 1) return;
    size_t p1 = pts - 1;
    double p = (double)p1;
    double xd = (double)f->xdelta;
    double acc = pow(p, 3.0) / 3.0 + (p * p) / 2.0 + p / 6.0 - (p * p) * (double)pts / 4.0;
    f->denominator = (float)((xd * xd) * acc);
    f->xAvg = f->xdelta * (float)p1 / 2; /* float*float, then float/2 */
}

/* cpp/psk_soft.cpp:135-174 */
static float linfit_calc_fit(linfit* f) {
    size_t pts = f->yvals.size;
    if (pts > 1) {
        size_t p1 = pts - 1;
        float half_span = f->xdelta * (float)p1 / 2;                                   /* :157 float */
        f->m = (float)((f->xySum - (double)half_span * f->ySum) / (double)f->denominator);
        f->b = (float)(f->ySum / (double)pts - (double)(f->m * f->xAvg));              /* :158 */
        float xVal = f->xdelta * (float)p1;                                            /* :161 */
        return f->m * xVal + f->b;                                                     /* :162 */
    }
    f->m = 0;
    if (pts == 0) f->b = 0;
    else f->b = DEQ_AT(&f->yvals, pts - 1);
    return f->b;
}

/* cpp/psk_soft.cpp:89-124 */
static float linfit_reset(linfit* f, const size_t* numPts, const float* sampleRate, int forceHistoryClear) {
    if (sampleRate) {
        float newXdelta = (float)(1.0 / (double)(*sampleRate));
        if (f->xdelta != newXdelta) {
            f->xdelta = newXdelta;
            forceHistoryClear = 1;
        }
    }
    if (forceHistoryClear) { f->yvals.size = 0; f->yvals.head = 0; }
    if (numPts && *numPts != f->n) {
        f->n = *numPts;
        while (f->yvals.size > f->n) DEQ_POP_FRONT(&f->yvals, 1);
    }
    f->ySum = 0;
    f->xySum = 0;
    for (size_t j = 0; j < f->yvals.size; j++) {
        float y = DEQ_AT(&f->yvals, j);
        f->ySum += (double)y;
        f->xySum += (double)((float)(unsigned int)j * f->xdelta * y); /* :118 all-float product */
    }
    linfit_calc_denominator(f);
    f->count = 0;
    return linfit_calc_fit(f);
}

/* cpp/psk_soft.cpp:48-87 */
static float linfit_next(linfit* f, float yval) {
    if (f->count == 1048576) linfit_reset(f, NULL, NULL, 0);
    int steady = f->yvals.size == f->n;
    if (steady) {
        f->ySum -= (double)DEQ_AT(&f->yvals, 0);                 /* :70 */
        DEQ_POP_FRONT(&f->yvals, 1);                             /* :71 */
        f->xySum -= (double)f->xdelta * f->ySum;                 /* :72 */
    }
    f->ySum += (double)yval;                                     /* :75 */
    f->xySum += (double)(yval * (float)f->yvals.size * f->xdelta); /* :78 float product */
    DEQ_PUSH(float, &f->yvals, yval);                            /* :79 */
    if (!steady) linfit_calc_denominator(f);                     /* :81-83 */
    f->count++;
    return linfit_calc_fit(f);
}

/* cpp/psk_soft.cpp:126-133 */
static float linfit_subtract_const(linfit* f, float c) {
    for (size_t j = 0; j < f->yvals.size; j++) DEQ_AT(&f->yvals, j) -= c;
    return linfit_reset(f, NULL, NULL, 0);
}

/* ---------------------------------------------------------------------------------------
 * component state (cpp/psk_soft.h:65-86) + properties (cpp/psk_soft_base.cpp:94-150)
 * ------------------------------------------------------------------------------------- */
typedef struct {
    /* properties, same names/types/defaults as psk_soft.prf.xml:23-60 */
    unsigned short samplesPerBaud;
    unsigned int numAvg;
    unsigned short constelationSize;
    unsigned short phaseAvg;
    int differentialDecoding;
    int resetState;
    /* members */
    cdeque samples;
    ddeque energy;
    double* symbolEnergy;
    size_t symbolEnergySize;
    size_t index;
    float complex last;
    int resetSamplesPerBaud, resetNumSymbols, resetPhaseAvg;
    float phaseEstimate;
    float sampleRate;
    size_t count;
    linfit phaseEstimator;
    /* what the four out-ports received */
    vec out_soft, out_bits, out_phase, out_sidx;
    long n_sri[4], n_packets[4];
    double sri_xdelta[4];
    int sri_mode[4];
    long warn_count;
} orc;

static void symE_assign(orc* c, size_t n) {
    c->symbolEnergy = (double*)realloc(c->symbolEnergy, (n ? n : 1) * sizeof(double));
    for (size_t i = 0; i < n; i++) c->symbolEnergy[i] = 0.0;
    c->symbolEnergySize = n;
}

/* cpp/psk_soft.cpp:187-199 + psk_soft_base.cpp:94-150 */
orc* orc_create(void) {
    orc* c = (orc*)calloc(1, sizeof(orc));
    c->samplesPerBaud = 10;
    c->numAvg = 100;
    c->constelationSize = 4;
    c->phaseAvg = 50;
    c->differentialDecoding = 0;
    c->resetState = 0;
    symE_assign(c, c->samplesPerBaud);
    c->index = 0;
    c->last = 0;
    c->resetSamplesPerBaud = 1;
    c->resetNumSymbols = 1;
    c->resetPhaseAvg = 1;
    c->phaseEstimate = 0.0f;
    c->sampleRate = 1.0f;
    c->count = 0;
    linfit_init(&c->phaseEstimator, c->phaseAvg, c->sampleRate);
    c->out_soft.esz = 2 * sizeof(float);
    c->out_bits.esz = sizeof(short);
    c->out_phase.esz = sizeof(float);
    c->out_sidx.esz = sizeof(short);
    return c;
}

void orc_destroy(orc* c) {
    if (!c) return;
    free(c->samples.v); free(c->energy.v); free(c->symbolEnergy); free(c->phaseEstimator.yvals.v);
    free(c->out_soft.v); free(c->out_bits.v); free(c->out_phase.v); free(c->out_sidx.v);
    free(c);
}

/* cpp/psk_soft.cpp:619-636 */
static void resync_energy(orc* c, size_t samplesPerSymbol, size_t numDataPts) {
    symE_assign(c, samplesPerSymbol);
    if (c->samples.size > numDataPts) {
        DEQ_TRUNC(&c->samples, numDataPts);
        DEQ_TRUNC(&c->energy, numDataPts);
    }
    c->index = 0;
    for (size_t i = 0; i < c->energy.size; i++) {
        c->symbolEnergy[c->index] += DEQ_AT(&c->energy, i);
        c->index++;
        if (c->index == samplesPerSymbol) c->index = 0;
    }
    c->count = 0;
}

/* configure() + change listeners, cpp/psk_soft.cpp:205-213, 638-651.  A listener runs only
 * when the value actually changed (same convention as oracle/ref_driver.cpp). */
int orc_configure(orc* c, const char* id, double v) {
    if (!strcmp(id, "samplesPerBaud")) {
        unsigned short n = (unsigned short)v; int ch = n != c->samplesPerBaud; c->samplesPerBaud = n;
        if (ch) c->resetSamplesPerBaud = (c->samplesPerBaud != c->symbolEnergySize);   /* :640 */
    } else if (!strcmp(id, "numAvg")) {
        c->numAvg = (unsigned int)v;
    } else if (!strcmp(id, "constelationSize")) {
        unsigned short n = (unsigned short)v; int ch = n != c->constelationSize; c->constelationSize = n;
        if (ch) c->resetNumSymbols = 1;                                                /* :645 */
    } else if (!strcmp(id, "phaseAvg")) {
        unsigned short n = (unsigned short)v; int ch = n != c->phaseAvg; c->phaseAvg = n;
        if (ch) c->resetPhaseAvg = 1;                                                  /* :650 */
    } else if (!strcmp(id, "differentialDecoding")) {
        c->differentialDecoding = (v != 0);
    } else if (!strcmp(id, "resetState")) {
        c->resetState = (v != 0);
    } else return -1;
    return 0;
}

double orc_query(orc* c, const char* id) {
    if (!strcmp(id, "samplesPerBaud")) return c->samplesPerBaud;
    if (!strcmp(id, "numAvg")) return c->numAvg;
    if (!strcmp(id, "constelationSize")) return c->constelationSize;
    if (!strcmp(id, "phaseAvg")) return c->phaseAvg;
    if (!strcmp(id, "differentialDecoding")) return c->differentialDecoding;
    if (!strcmp(id, "resetState")) return c->resetState;
    return -1;
}

/* libstdc++ std::pow(complex<float>, int) -> __complex_pow_unsigned (n >= 0 here) */
static float complex cpow_unsigned(float complex x, unsigned n) {
    float complex y = (n % 2) ? x : (float complex)1.0f;
    while (n >>= 1) {
        x *= x;
        if (n % 2) y *= x;
    }
    return y;
}

static void push_sri(orc* c, int port, double xdelta, int mode) {
    c->n_sri[port]++;
    c->sri_xdelta[port] = xdelta;
    c->sri_mode[port] = mode;
}

/* One BULKIO packet through serviceFunction (cpp/psk_soft.cpp:346-618).
 * n_floats interleaved re,im; returns NORMAL (1). */
int orc_push(orc* c, const float* iq, size_t n_floats, double sri_xdelta, int sri_mode, int flushed, int sri_changed) {
    if (flushed) { c->warn_count++; c->resetState = 1; }                                /* :353-357 */
    if (sri_mode != 1) { c->warn_count++; return 1; }                                    /* :359-363 */
    if (c->resetState) {                                                                 /* :365-372 */
        c->resetSamplesPerBaud = 1; c->resetNumSymbols = 1; c->resetPhaseAvg = 1; c->resetState = 0;
    }
    const size_t samplesPerSymbol = c->samplesPerBaud;                                   /* :376 */
    const size_t numDataPts = samplesPerSymbol * c->numAvg;                              /* :377 */
    const size_t numSyms = c->constelationSize;                                          /* :378 */
    if (numDataPts > c->samples.size) c->resetSamplesPerBaud = 1;                        /* :380-383 */
    size_t bitsPerBaud = 0;                                                              /* :384-390 */
    if (numSyms == 2) bitsPerBaud = 1;
    else if (numSyms == 4) bitsPerBaud = 2;
    else if (numSyms == 8) bitsPerBaud = 3;

    if (sri_changed || c->resetNumSymbols || c->resetSamplesPerBaud) {                   /* :393-405 */
        if (sri_xdelta != (double)c->sampleRate) {
            c->sampleRate = (float)(1.0 / sri_xdelta);
            linfit_reset(&c->phaseEstimator, NULL, &c->sampleRate, 0);
        }
        double xd = sri_xdelta;
        xd *= (double)samplesPerSymbol;
        push_sri(c, 0, xd, sri_mode);
        push_sri(c, 2, xd, 0);
        xd /= (double)bitsPerBaud;
        push_sri(c, 1, xd, 0);
    }
    if (c->resetSamplesPerBaud) {                                                        /* :408-412 */
        resync_energy(c, samplesPerSymbol, numDataPts);
        c->resetSamplesPerBaud = 0;
    }
    if (c->resetNumSymbols) {                                                            /* :416-420 */
        linfit_reset(&c->phaseEstimator, NULL, NULL, 1);
        c->resetNumSymbols = 0;
    }
    if (c->resetPhaseAvg) {                                                              /* :421-426 */
        size_t numPts = c->phaseAvg;
        linfit_reset(&c->phaseEstimator, &numPts, NULL, 0);
        c->resetPhaseAvg = 0;
    }

    const size_t n_complex = n_floats / 2;                                               /* :428 */
    size_t n_out = 0, n_bits = 0, n_sidx = 0;
    float complex sample = 0;
    const size_t lastSample = samplesPerSymbol - 1;
    for (size_t i = 0; i < n_complex; i++) {                                             /* :442 */
        float complex x = CMPLXF(iq[2 * i], iq[2 * i + 1]);
        if (samplesPerSymbol > 1) {                                                      /* :445-452 */
            DEQ_PUSH(float complex, &c->samples, x);
            float re = crealf(x), im = cimagf(x);
            double sampleEnergy = (double)(re * re + im * im);   /* std::norm<float>: float arithmetic */
            DEQ_PUSH(double, &c->energy, sampleEnergy);
            c->symbolEnergy[c->index] += sampleEnergy;
        }
        if (c->index == lastSample) {                                                    /* :454 */
            if (c->samples.size == numDataPts) {                                         /* :457 */
                if (samplesPerSymbol > 1) {                                              /* :459-467 */
                    size_t sampleIndex = 0;                      /* std::max_element: first maximum */
                    for (size_t p = 1; p < c->symbolEnergySize; p++)
                        if (c->symbolEnergy[sampleIndex] < c->symbolEnergy[p]) sampleIndex = p;
                    sample = DEQ_AT(&c->samples, sampleIndex);
                    short si = (short)sampleIndex;
                    vec_push(&c->out_sidx, &si); n_sidx++;
                } else
                    sample = x;                                                          /* :469 */

                double thisPhase = (double)cargf(cpow_unsigned(sample, (unsigned)(int)numSyms)); /* :474 */
                long numWraps = (long)round(((double)c->phaseEstimate - thisPhase) / ORC_M_2PI); /* :477 */
                thisPhase += (double)numWraps * ORC_M_2PI;                               /* :478 */
                c->phaseEstimate = linfit_next(&c->phaseEstimator, (float)thisPhase);    /* :481 */
                vec_push(&c->out_phase, &c->phaseEstimate);                              /* :482 */

                float phaseCorrection = 0;
                if (c->differentialDecoding) {                                           /* :486-491 */
                    float complex decoded = sample / c->last;    /* libgcc __divsc3 */
                    c->last = sample;
                    sample = decoded;
                } else {
                    phaseCorrection = -c->phaseEstimate / (float)numSyms;                /* :494 */
                }
                if (numSyms == 4) phaseCorrection = (float)((double)phaseCorrection + M_PI_4); /* :497-498 */
                float sn, cs;
                sincosf(phaseCorrection, &sn, &cs);                                      /* :499 std::polar(1.0f, pc) */
                float complex phasor = CMPLXF(1.0f * cs, 1.0f * sn);
                float complex corrected = sample * phasor;                               /* :500 */
                float o[2] = { crealf(corrected), cimagf(corrected) };
                vec_push(&c->out_soft, o); n_out++;                                      /* :501 */

                if (bitsPerBaud == 1) {                                                  /* :503-513 */
                    short bit = (o[0] < 0);
                    vec_push(&c->out_bits, &bit); n_bits++;
                } else if (bitsPerBaud == 2) {                                           /* :514-527 (float -> bool, sic) */
                    int real = (o[0] != 0), imag = (o[1] != 0);
                    short b0 = (short)(real ^ imag), b1 = (short)(!imag);
                    vec_push(&c->out_bits, &b0); vec_push(&c->out_bits, &b1); n_bits += 2;
                } else if (bitsPerBaud == 3) {                                           /* :528-564 */
                    float theta = atan2f(o[1], o[0]);
                    float softsym = (float)((double)theta / M_PI * 4);
                    if (softsym < -.5) softsym += 8;
                    unsigned short sym = (unsigned short)(int)roundf(softsym);
                    for (size_t j = 0; j != 3; j++) {
                        short bit = sym & 1;
                        vec_push(&c->out_bits, &bit); n_bits++;
                        sym = sym >> 1;
                    }
                } else
                    c->warn_count++;                                                     /* :565-566 */

                if (samplesPerSymbol > 1) {                                              /* :568-584 */
                    for (size_t p = 0; p < samplesPerSymbol; p++)
                        c->symbolEnergy[p] -= DEQ_AT(&c->energy, p);
                    DEQ_POP_FRONT(&c->energy, samplesPerSymbol);
                    DEQ_POP_FRONT(&c->samples, samplesPerSymbol);
                    c->count++;
                    if (c->count == 1048576) resync_energy(c, samplesPerSymbol, numDataPts);
                }
            }
            c->index = 0;                                                                /* :587 */
        } else
            c->index++;                                                                  /* :590 */
    }
    /* packet-end wrap, :592-603.  abs() is ::abs(int): cvttss2si, integer abs, back to float */
    float wrapValue = (float)(ORC_M_2PI * (double)numSyms);
    float pe = c->phaseEstimate;
    int pe_i = (pe >= -2147483648.0f && pe < 2147483648.0f) ? (int)pe : (int)0x80000000; /* cvttss2si */
    int pe_abs = (pe_i == (int)0x80000000) ? pe_i : (pe_i < 0 ? -pe_i : pe_i);
    if ((float)pe_abs > wrapValue) {
        long numWraps = (long)roundf(c->phaseEstimate / wrapValue);                      /* :598 */
        c->phaseEstimate = linfit_subtract_const(&c->phaseEstimator, (float)numWraps * wrapValue); /* :601-602 */
    }
    if (n_out) c->n_packets[0]++;                                                        /* :605-615 */
    if (n_bits) c->n_packets[1]++;
    if (n_out) c->n_packets[2]++;
    if (n_sidx) c->n_packets[3]++;
    return 1;
}

/* port: 0 soft, 1 bits, 2 phase, 3 sampleIndex.  dst NULL -> elements waiting (soft counted in floats) */
size_t orc_read(orc* c, int port, void* dst, size_t cap_elems) {
    vec* a = port == 0 ? &c->out_soft : port == 1 ? &c->out_bits : port == 2 ? &c->out_phase : &c->out_sidx;
    size_t mult = port == 0 ? 2 : 1;
    size_t n = a->size * mult;
    if (dst) {
        size_t m = n < cap_elems ? n : cap_elems;
        if (m) memcpy(dst, a->v, m * (a->esz / mult));
        a->size = 0;
    }
    return n;
}
long orc_sri_count(orc* c, int port) { return c->n_sri[port]; }
double orc_sri_xdelta(orc* c, int port) { return c->sri_xdelta[port]; }
int orc_sri_mode(orc* c, int port) { return c->sri_mode[port]; }
long orc_packet_count(orc* c, int port) { return c->n_packets[port]; }

/* whole stream in packets of packet_len complex samples; same contract as ref_demod */
size_t orc_demod(orc* c, const float* iq, size_t n_complex, size_t packet_len, double xdelta,
                 float* soft, short* bits, float* phase, short* sidx,
                 size_t cap_syms, size_t cap_bits, size_t* n_bits_out) {
    size_t ns = 0, nb = 0;
    for (size_t off = 0; off < n_complex; off += packet_len) {
        size_t m = n_complex - off < packet_len ? n_complex - off : packet_len;
        orc_push(c, iq + 2 * off, 2 * m, xdelta, 1, 0, off == 0);
        size_t k = c->out_phase.size, b = c->out_bits.size;
        if (soft && k && ns + k <= cap_syms) memcpy(soft + 2 * ns, c->out_soft.v, k * 2 * sizeof(float));
        if (phase && k && ns + k <= cap_syms) memcpy(phase + ns, c->out_phase.v, k * sizeof(float));
        if (sidx && c->out_sidx.size && ns + k <= cap_syms) memcpy(sidx + ns, c->out_sidx.v, c->out_sidx.size * sizeof(short));
        if (bits && b && nb + b <= cap_bits) memcpy(bits + nb, c->out_bits.v, b * sizeof(short));
        c->out_soft.size = c->out_phase.size = c->out_sidx.size = c->out_bits.size = 0;
        ns += k; nb += b;
    }
    if (n_bits_out) *n_bits_out = nb;
    return ns;
}
