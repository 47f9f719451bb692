/* TEST INFRASTRUCTURE ONLY -- stand-in for <ossie/ThreadedComponent.h>
 * (psk_soft_base.h:29, psk_soft_base.cpp:68-80, psk_soft.cpp:351,362,617).
 * The oracle harness calls serviceFunction() itself; no thread is started. */
#ifndef ORACLE_STUB_OSSIE_THREADEDCOMPONENT_H
#define ORACLE_STUB_OSSIE_THREADEDCOMPONENT_H

enum { NOOP = -1, FINISH = 0, NORMAL = 1 };

class ThreadedComponent {
public:
    ThreadedComponent() {}
    virtual ~ThreadedComponent() {}
    virtual int serviceFunction() = 0;
    void startThread() {}
    bool stopThread() { return true; }
};

#endif
