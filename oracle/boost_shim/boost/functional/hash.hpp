// Test-infrastructure only: stand-in for <boost/functional/hash.hpp> (oracle build).
#pragma once
#include <functional>
#include <string>
namespace boost {
template <class T> inline std::size_t hash_value(const T& v) { return std::hash<T>()(v); }
template <class T> inline void hash_combine(std::size_t& seed, const T& v) {
    seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
}
