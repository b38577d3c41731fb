// LBFGSpp/PhaseClock.h -- optional wall-clock accounting of named phases of the host-driven loops (bench / profiles only; off by
// default and then costs one branch per scope).  A scope synchronises the device on entry and exit when the clock is enabled, so the
// seconds it records are device + host time of exactly the calls it encloses.  Not part of the reference's API.
#ifndef LBFGSPP_B200_PHASE_CLOCK_H
#define LBFGSPP_B200_PHASE_CLOCK_H

#include <chrono>
#include <map>
#include <string>

#include "DeviceVector.h"

namespace LBFGSpp {

class PhaseClock
{
public:
    struct Entry { double seconds; long calls; };
    static PhaseClock& get()
    {
        static thread_local PhaseClock c;
        return c;
    }
    bool enabled;
    std::map<std::string, Entry> acc;
    PhaseClock() : enabled(false) {}
    void reset() { acc.clear(); }
    static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    class Scope
    {
        Device* m_dev;
        const char* m_name;
        double m_t0;
    public:
        Scope(Device& dev, const char* name) : m_dev(nullptr), m_name(name), m_t0(0)
        {
            if (!PhaseClock::get().enabled) return;
            m_dev = &dev;
            dev.synchronize();
            m_t0 = now();
        }
        ~Scope()
        {
            if (!m_dev) return;
            try { m_dev->synchronize(); } catch (...) {}
            Entry& e = PhaseClock::get().acc[m_name];
            e.seconds += now() - m_t0;
            e.calls += 1;
        }
    };
};

}  // namespace LBFGSpp

#endif  // LBFGSPP_B200_PHASE_CLOCK_H
