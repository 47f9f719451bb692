/* TEST INFRASTRUCTURE ONLY -- stand-in for <bulkio/bulkio.h>: just enough of the
 * BULKIO port classes for psk_soft.cpp:349-363,393-405,428,605-616 and
 * psk_soft_base.cpp:38-47.  The in-port hands out whatever packet the harness
 * queued; the out-ports record what was pushed so the harness can read it back. */
#ifndef ORACLE_STUB_BULKIO_H
#define ORACLE_STUB_BULKIO_H

#include <string>
#include <vector>
#include <deque>

namespace BULKIO {
    struct StreamSRI {
        double xdelta;
        short  mode;
        StreamSRI() : xdelta(1.0), mode(0) {}
    };
    struct PrecisionUTCTime {
        double twsec, tfsec;
        PrecisionUTCTime() : twsec(0), tfsec(0) {}
    };
}

namespace bulkio {
    namespace Const {
        const float BLOCKING = -1.0f;
        const float NON_BLOCKING = 0.0f;
    }

    template <class E>
    class InPortStub {
    public:
        struct dataTransfer {
            std::vector<E> dataBuffer;
            BULKIO::StreamSRI SRI;
            bool sriChanged;
            bool inputQueueFlushed;
            BULKIO::PrecisionUTCTime T_;
            BULKIO::PrecisionUTCTime& T;
            bool EOS;
            std::string streamID;
            dataTransfer() : sriChanged(false), inputQueueFlushed(false), T(T_), EOS(false) {}
        };
        explicit InPortStub(const std::string&) {}
        ~InPortStub() { while (!q.empty()) { delete q.front(); q.pop_front(); } }
        dataTransfer* getPacket(float) {
            if (q.empty()) return 0;
            dataTransfer* p = q.front(); q.pop_front(); return p;
        }
        void oracle_enqueue(dataTransfer* p) { q.push_back(p); }
    private:
        std::deque<dataTransfer*> q;
    };

    template <class E>
    class OutPortStub {
    public:
        explicit OutPortStub(const std::string&) : n_sri(0), n_packets(0) {}
        void pushSRI(const BULKIO::StreamSRI& s) { last_sri = s; n_sri++; }
        void pushPacket(const std::vector<E>& d, const BULKIO::PrecisionUTCTime&, bool, const std::string&) {
            data.insert(data.end(), d.begin(), d.end());
            n_packets++;
        }
        std::vector<E> data;          /* everything pushed since the harness last cleared it */
        BULKIO::StreamSRI last_sri;
        long n_sri, n_packets;
    };

    typedef InPortStub<float>  InFloatPort;
    typedef OutPortStub<float> OutFloatPort;
    typedef OutPortStub<short> OutShortPort;
}

#endif
