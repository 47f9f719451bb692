/* TEST INFRASTRUCTURE ONLY -- minimal stand-in for the REDHAWK core-framework
 * header <ossie/Component.h>, written from scratch so that the UNMODIFIED
 * reference sources (/root/reference/cpp/psk_soft.cpp, psk_soft_base.cpp) compile
 * in place without REDHAWK/omniORB/boost.  It only provides what those two files
 * name (psk_soft_base.h:24-30, psk_soft_base.cpp:32-150, psk_soft.cpp:33,205-213,
 * 355-367,566,639-649).  Nothing here is product code.
 *
 * The real framework headers transitively provide <complex> and <cmath> (psk_soft.h:65-66
 * uses M_PI and std::complex before psk_soft.cpp:30-31 includes them), so this stub includes
 * exactly those two -- and deliberately NOT the C++ wrapper <math.h>/<stdlib.h> (which on
 * modern libstdc++ add `using std::abs;`): on the reference's supported toolchains (el6 gcc 4.4 /
 * el7 gcc 4.8, .gitlab-ci.yml) the unqualified `abs(phaseEstimate)` at psk_soft.cpp:596
 * resolves to ::abs(int), and that is what this build must keep (see oracle/README.md).
 */
#ifndef ORACLE_STUB_OSSIE_COMPONENT_H
#define ORACLE_STUB_OSSIE_COMPONENT_H

#include <complex>
#include <cmath>
#include <string>
#include <vector>
#include <deque>
#include <map>
#include <algorithm>   /* psk_soft.cpp:462 uses std::max_element without including it */

namespace CORBA {
    typedef unsigned int ULong;
    struct SystemException {};
}

namespace CF {
    enum ErrorNumberType { CF_NOTSET };
    namespace Resource {
        struct StartError { };
        struct StopError {
            StopError(ErrorNumberType, const char* m) : msg(m) {}
            std::string msg;
        };
    }
    namespace LifeCycle {
        struct ReleaseError { };
    }
}

/* logging macros: the stream expression is dropped unevaluated (no iostream in the oracle) */
struct oracle_stub_logsink {
    static int& warn_count() { static int n = 0; return n; }
};
#define ENABLE_LOGGING
#define PREPARE_LOGGING(cls)
#define LOG_WARN(cls, expr)  { oracle_stub_logsink::warn_count()++; }
#define LOG_DEBUG(cls, expr) { }

/* type-erased property-change listener */
struct oracle_stub_listener {
    virtual ~oracle_stub_listener() {}
    virtual void fire(const std::string& id) = 0;
};
template <class T>
struct oracle_stub_member_listener : public oracle_stub_listener {
    typedef void (T::*fn_t)(const std::string&);
    T* obj; fn_t fn;
    oracle_stub_member_listener(T* o, fn_t f) : obj(o), fn(f) {}
    void fire(const std::string& id) { (obj->*fn)(id); }
};

class Component {
public:
    Component(const char* uuid, const char* label) : _uuid(uuid ? uuid : ""), _label(label ? label : "") {}
    virtual ~Component() {
        for (std::map<std::string, oracle_stub_listener*>::iterator i = _listeners.begin(); i != _listeners.end(); ++i)
            delete i->second;
    }
    virtual void start() {}
    virtual void stop() {}
    virtual void releaseObject() {}
    virtual void constructor() {}

    /* addProperty(member, default, id, name, mode, units, action, kinds): assigns the default */
    template <class T, class D>
    void addProperty(T& ref, const D& dflt, const std::string&, const std::string&, const std::string&,
                     const std::string&, const std::string&, const std::string&) {
        ref = static_cast<T>(dflt);
    }
    template <class P>
    void addPort(const std::string&, const std::string&, P*) {}
    template <class P>
    void addPort(const std::string&, P*) {}

    template <class T>
    void setPropertyChangeListener(const std::string& id, T* obj, void (T::*fn)(const std::string&)) {
        std::map<std::string, oracle_stub_listener*>::iterator i = _listeners.find(id);
        if (i != _listeners.end()) { delete i->second; }
        _listeners[id] = new oracle_stub_member_listener<T>(obj, fn);
    }
    /* harness hook: what the framework does after a configure() of property `id` */
    void oracle_fire_listener(const std::string& id) {
        std::map<std::string, oracle_stub_listener*>::iterator i = _listeners.find(id);
        if (i != _listeners.end()) i->second->fire(id);
    }
private:
    std::string _uuid, _label;
    std::map<std::string, oracle_stub_listener*> _listeners;
};

#endif
