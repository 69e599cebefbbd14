// Test-infrastructure only: minimal stand-in for <boost/exception/all.hpp> so that the
// UNMODIFIED reference sources under /root/reference compile without Boost (see oracle/README.md).
#pragma once
#include <exception>
#include <string>
#include <sstream>
#include <typeinfo>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <climits>
#include <unistd.h>
namespace boost {
class exception {
public:
    virtual ~exception() noexcept {}
    mutable std::string shim_msg;   // thrown temporaries are const
};
template <class Tag, class T> class error_info {
public:
    typedef T value_type;
    error_info(const value_type& v) : v_(v) {}
    const value_type& value() const { return v_; }
private:
    value_type v_;
};
template <class E, class Tag, class T>
inline const E& operator<<(const E& x, const error_info<Tag, T>& v) {
    std::ostringstream ss; ss << v.value();
    if (!x.shim_msg.empty()) x.shim_msg += "\n";
    x.shim_msg += ss.str();
    return x;
}
template <class ErrorInfo, class E>
inline const typename ErrorInfo::value_type* get_error_info(const E&) { return nullptr; }
inline std::string diagnostic_information(const exception& e) { return e.shim_msg; }
inline std::string diagnostic_information(const std::exception& e) {
    const exception* be = dynamic_cast<const exception*>(&e);
    return be ? be->shim_msg : std::string(e.what());
}
}
#define BOOST_THROW_EXCEPTION(x) throw (x)
