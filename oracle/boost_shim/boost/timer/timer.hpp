// Test-infrastructure only: stand-in for <boost/timer/timer.hpp> (oracle build).
#pragma once
#include <chrono>
#include <string>
#include <iostream>
#include <iomanip>
#include <sstream>
namespace boost { namespace timer {
class auto_cpu_timer {
public:
    auto_cpu_timer(short places, const std::string& fmt) : places_(places), fmt_(fmt), t0_(std::chrono::steady_clock::now()) {}
    ~auto_cpu_timer() {
        double w = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
        std::ostringstream ss; ss << std::fixed << std::setprecision(places_) << w;
        std::string out(fmt_); size_t p;
        while ((p = out.find("%w")) != std::string::npos) out.replace(p, 2, ss.str());
        std::cout << out; std::cout.flush();
    }
private:
    short places_; std::string fmt_; std::chrono::steady_clock::time_point t0_;
};
}}
