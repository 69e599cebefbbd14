// Test-infrastructure only: stand-in for <boost/lexical_cast.hpp> (oracle build, see oracle/README.md).
#pragma once
#include <sstream>
#include <string>
#include <stdexcept>
#include <type_traits>
namespace boost {
class bad_lexical_cast : public std::bad_cast {
public: const char* what() const noexcept override { return "bad lexical cast"; }
};
namespace shim_detail {
template <class T, class S> struct caster {
    static T cast(const S& s) {
        std::stringstream ss; ss.precision(17); ss << s;
        T t; if (!(ss >> t)) throw bad_lexical_cast();
        return t;
    }
};
template <class S> struct caster<std::string, S> {
    static std::string cast(const S& s) { std::ostringstream ss; ss.precision(17); ss << s; return ss.str(); }
};
template <> struct caster<std::string, std::string> {
    static std::string cast(const std::string& s) { return s; }
};
template <> struct caster<std::string, double> {
    static std::string cast(const double& s) { std::ostringstream ss; ss.precision(17); ss << s; return ss.str(); }
};
}
template <class T, class S> inline T lexical_cast(const S& s) { return shim_detail::caster<T, S>::cast(s); }
template <class T> inline T lexical_cast(const char* s) { return shim_detail::caster<T, std::string>::cast(std::string(s)); }
}
