/* pskd.h -- C ABI of the B200-native PSK soft demodulator (libpskd.so).
 *
 * Drop-in boundary for ONE path of the REDHAWK component rh.psk_soft: the demod core that
 * runs inside psk_soft_i::serviceFunction() (reference: cpp/psk_soft.cpp:346-618) together
 * with its LinearFit helper (cpp/psk_soft.cpp:35-185) and resyncEnergy (:619-636).
 * Everything else of the component (CORBA ports, property plumbing, service thread) stays in
 * the caller; INTEGRATION.md shows the psk_soft_i-side binding.
 *
 * One handle ("bank") = n_channels independent demodulators = n_channels independent
 * psk_soft_i instances, all resident on one GPU.  The library owns only the carried state of
 * each channel (reference: the members at cpp/psk_soft.h:66-86); the caller owns every data
 * buffer.  Calls on one handle must be serial; different handles may be driven from different
 * host threads (one handle per GPU is how a channel bank is sharded, no collective).
 *
 * There is NO CPU fallback: every entry point fails with PSKD_ERR_CUDA when no sm_100
 * device / driver is present.
 *
 * Plain C types only -- no torch, no C++ types cross this boundary.
 */
#ifndef PSKD_H
#define PSKD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSKD_ABI_VERSION 2

/* ---- return codes.  Mirrors the reference's conventions: serviceFunction returns NORMAL for
 * every packet it consumed, including ones it ignores (cpp/psk_soft.cpp:359-363,617); problems
 * it can live with are warnings (:355,361,566). */
#define PSKD_OK                 0
#define PSKD_IGNORED_REAL_DATA  1   /* sri_mode != 1: packet dropped, no state change (cpp/psk_soft.cpp:359-363) */
#define PSKD_NO_BITS            2   /* some channel has constelationSize not in {2,4,8}: soft/phase/index still produced, no bits (:565-566) */
#define PSKD_ERR_ARG           -1
#define PSKD_ERR_CUDA          -2
#define PSKD_ERR_NOMEM         -3
#define PSKD_ERR_UNSUPPORTED   -4   /* property combination outside the GPU path (see DESIGN.md "limits") */
#define PSKD_ERR_CAPACITY      -5   /* an output buffer is too small for what this call produces */

/* ---- properties: same ids, types and defaults as psk_soft.prf.xml:23-60 /
 * cpp/psk_soft_base.cpp:94-150 / cpp/psk_soft_base.h:44-56. */
typedef struct pskd_props {
    uint16_t samplesPerBaud;        /* ushort, default 10 */
    uint32_t numAvg;                /* ulong,  default 100 */
    uint16_t constelationSize;      /* ushort, default 4 (2 BPSK / 4 QPSK / 8 8-PSK) */
    uint16_t phaseAvg;              /* ushort, default 50 */
    uint8_t  differentialDecoding;  /* bool,   default false */
    uint8_t  resetState;            /* bool,   default false; consumed (cleared) by the next pskd_process */
} pskd_props;

/* ---- flags for pskd_process */
#define PSKD_FLAG_HOST_BUFFERS   0x1u  /* iq and every output pointer are HOST memory (pinned preferred); the
                                          call stages H2D / D2H itself and returns when the outputs are on the host */
#define PSKD_FLAG_QUEUE_FLUSHED  0x2u  /* dataTransfer::inputQueueFlushed (cpp/psk_soft.cpp:353-357) */
#define PSKD_FLAG_SRI_CHANGED    0x4u  /* dataTransfer::sriChanged (cpp/psk_soft.cpp:393) */
#define PSKD_FLAG_NO_SYNC        0x8u  /* device buffers only: enqueue on the bank's stream and return; pair with pskd_sync */

/* ---- one call's input: what getPacket() hands to serviceFunction (cpp/psk_soft.cpp:349,428),
 * for every channel of the bank at once.  Channel c's samples start at iq + 2*c*iq_stride. */
typedef struct pskd_input {
    const float*  iq;          /* interleaved re,im float32 (BULKIO dataFloat, SRI.mode==1) */
    size_t        iq_stride;   /* complex samples between consecutive channels' first samples */
    const size_t* n_complex;   /* [n_channels] complex samples per channel (HOST array), or NULL */
    size_t        n_complex_all; /* used for every channel when n_complex == NULL */
    double        sri_xdelta;  /* SRI.xdelta of the input stream (cpp/psk_soft.cpp:394-399); a numeric input */
    int           sri_mode;    /* SRI.mode; must be 1 (complex) */
    size_t        packet_len;  /* the call is processed AS IF delivered in BULKIO packets of this many complex
                                  samples (last one shorter); 0 = the whole call is one packet.  Packet
                                  boundaries are numerically visible in the reference (cpp/psk_soft.cpp:380-426,
                                  592-603), so they are part of the contract. */
    uint32_t      flags;
} pskd_input;

/* ---- one call's outputs: the four out-ports (psk_soft.scd.xml:32-73, cpp/psk_soft.cpp:605-615).
 * Channel c writes at soft + 2*c*sym_stride, phase + c*sym_stride, sample_index + c*sym_stride,
 * bits + c*bits_stride.  A NULL pointer skips that port.  Capacity needed per channel:
 * pskd_max_symbols() symbols and 3x that many bits. */
typedef struct pskd_output {
    float*   soft;          /* softDecision_dataFloat_out: complex float per symbol (re,im) */
    int16_t* bits;          /* bits_dataShort_out: one short (0/1) per bit, bitsPerBaud per symbol, LSB first */
    float*   phase;         /* phase_dataFloat_out: phase estimate per symbol */
    int16_t* sample_index;  /* sampleIndex_dataShort_out: chosen sample phase per symbol */
    size_t   sym_stride;    /* symbols between consecutive channels in soft/phase/sample_index */
    size_t   bits_stride;   /* shorts between consecutive channels in bits */
    size_t*  n_symbols;     /* [n_channels] HOST array, filled by the call (may be NULL) */
    size_t*  n_bits;        /* [n_channels] HOST array, filled by the call (may be NULL) */
    uint8_t* hard;          /* ADDITIONAL output, not a port of the reference: the decided symbol as one byte per symbol, its
                               bitsPerBaud bits packed LSB first (bit j = bits[k*bitsPerBaud + j]); row stride sym_stride;
                               NULL skips it.  Never replaces `bits`. */
} pskd_output;

/* ---- out-port stream metadata the component pushes beside the data (cpp/psk_soft.cpp:393-405):
 * soft xdelta*S mode 1; phase same xdelta mode 0; bits xdelta*S/bitsPerBaud mode 0;
 * sampleIndex never gets an SRI. */
typedef struct pskd_sri_out {
    double soft_xdelta;  int soft_mode;
    double phase_xdelta; int phase_mode;
    double bits_xdelta;  int bits_mode;
    long   sri_pushes;   /* how many times the component would have called pushSRI on each of those ports */
} pskd_sri_out;

typedef struct pskd_stats {
    uint64_t symbols_out;        /* total symbols produced since create */
    uint64_t samples_in;         /* total complex samples consumed since create */
    uint64_t packets;            /* emulated BULKIO packets since create (summed over channels) */
    uint64_t wraps;              /* packet-end phase wraps applied (cpp/psk_soft.cpp:596-603) */
    uint64_t spec_chunks;        /* phase-chain chunks run speculatively */
    uint64_t spec_misses;        /* chunks whose speculation failed verification and were re-run exactly */
    uint64_t seq_channels;       /* channel-calls that took the sequential (non-speculative) chain */
    uint64_t tp_packets;         /* emulated packets whose phase chain ran time-parallel and whose hand-over was proven */
    uint64_t tp_repaired;        /* channel-calls whose time-parallel plan needed the repair round (a hand-over was not proven:
                                    the classic unwrap count differs from the reference's rule somewhere) */
} pskd_stats;

typedef struct pskd_bank* pskd_handle;

/* fill *p with the defaults above */
void pskd_default_props(pskd_props* p);

/* create a bank of n_channels demodulators on CUDA device `device`.
 * props: [n_channels] per-channel properties, or NULL for defaults everywhere.
 * replaces: psk_soft_i::psk_soft_i + constructor() (cpp/psk_soft.cpp:187-213) x n_channels. */
int pskd_create(pskd_handle* out, int device, int n_channels, const pskd_props* props);

/* replaces: psk_soft_i::~psk_soft_i (cpp/psk_soft.cpp:201-203) */
int pskd_destroy(pskd_handle h);

/* configure() of one channel (ch >= 0) or of all channels (ch == -1); takes effect at the next
 * pskd_process (properties are latched per packet, cpp/psk_soft.cpp:374-378) and runs the change
 * listeners' logic (cpp/psk_soft.cpp:638-651). */
int pskd_set_props(pskd_handle h, int ch, const pskd_props* p);
int pskd_get_props(pskd_handle h, int ch, pskd_props* p);

/* upper bound of symbols one channel can emit for a call with n_complex new samples
 * (the reserve() bound of cpp/psk_soft.cpp:434 plus the carried window) */
size_t pskd_max_symbols(pskd_handle h, int ch, size_t n_complex);

/* THE hot path.  replaces: the body of psk_soft_i::serviceFunction() between getPacket and the
 * four pushPacket calls (cpp/psk_soft.cpp:353-603), for every channel, for every emulated packet. */
int pskd_process(pskd_handle h, const pskd_input* in, pskd_output* out);

/* wait for everything enqueued with PSKD_FLAG_NO_SYNC; returns PSKD_ERR_CUDA on a device fault */
int pskd_sync(pskd_handle h);

/* the CUDA stream (cudaStream_t) the bank launches on, for event timing by the caller */
void* pskd_stream(pskd_handle h);

int pskd_get_sri(pskd_handle h, int ch, pskd_sri_out* sri);
int pskd_get_stats(pskd_handle h, pskd_stats* st);

/* ---- checkpoint / resume of a bank's carried state (reference: the members at cpp/psk_soft.h:66-86 of
 * every channel: the timing window's samples, the LinearFit history and sums, phaseEstimate, `last`,
 * sampleRate, the three reset flags; plus the live properties and the out-SRI bookkeeping).  The
 * reference has no checkpointing -- its state dies with the process; this lets a
 * caller migrate a live bank to another process / GPU or survive a restart without re-acquisition.
 * The blob is self-describing (magic, version, n_channels) and position-independent. */
size_t pskd_state_size(pskd_handle h);                         /* bytes pskd_state_export writes right now */
int pskd_state_export(pskd_handle h, void* host_buf, size_t cap, size_t* written);
int pskd_state_import(pskd_handle h, const void* host_buf, size_t n_bytes);   /* bank must have the same n_channels */

/* number of kernel launches this bank has issued since create (bench.py's gpu_launches) */
uint64_t pskd_launch_count(pskd_handle h);

/* per-kernel device timing, measured with CUDA events on the bank's stream around every launch
 * (bench.py's live roofline figure).  Off by default. */
typedef struct pskd_kernel_time {
    char     name[32];
    double   ms_total;     /* summed event-to-event time of this kernel's launches */
    uint64_t launches;
    double   alg_bytes;    /* algorithmic bytes (8 N in + per symbol 8 soft + 4 phase + 2 sampleIndex + 2 b bits, split per
                              stage) of the channels those launches served: alg_bytes / ms_total is the kernel's roofline rate */
} pskd_kernel_time;
int pskd_profile_enable(pskd_handle h, int on);
/* waits for the stream, then copies up to cap entries; *n = entries available; reset != 0 zeroes them */
int pskd_profile_read(pskd_handle h, pskd_kernel_time* out, int cap, int* n, int reset);

/* last CUDA / argument error text for this thread */
const char* pskd_last_error(void);

int pskd_abi_version(void);

/* CUDA devices visible to this process (0 when there is no driver / device: every other entry point then fails with
 * PSKD_ERR_CUDA -- there is no CPU path) */
int pskd_device_count(void);

/* ---- synthetic channel-bank generator (benchmark / test input, generated in HBM so that
 * 30+ GB banks never cross PCIe).  Counter-based: sample n of channel c depends only on
 * (seed, c, n).  Not part of the reference; SURVEY.md section 8d fixes the value distribution. */
typedef struct pskd_synth {
    uint64_t seed;
    uint16_t samplesPerBaud;
    uint16_t constelationSize;
    float    sigma;        /* AWGN std-dev per dimension */
    float    freq_max;     /* per-channel carrier offset drawn uniformly in +-freq_max cycles/sample */
    float    pn_sigma;     /* phase-noise amplitude (rad), slow pseudo-random wander */
    float    shaped;       /* 1: envelope 0.6+0.4*sin(pi*(p+.5)/S), 0: rectangular pulses */
    uint32_t period;       /* > 0: the carrier offset of every channel is rounded to a multiple of 1/(constelationSize*period)
                              cycles/sample, so that the M-th power phase of a buffer of `period` samples continues seamlessly
                              when the buffer is replayed (bench.py replays one resident buffer with carried state) */
} pskd_synth;

/* fills iq_dev[c][0..n_complex) (device memory, stride in complex samples) for channels
 * [ch0, ch0+n_channels) of a notional global bank */
int pskd_synth_fill(int device, float* iq_dev, size_t iq_stride, int ch0, int n_channels,
                    size_t n_complex, const pskd_synth* cfg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSKD_H */
