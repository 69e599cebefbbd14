// Test-infrastructure only: stand-in for <boost/algorithm/string.hpp> (oracle build).
#pragma once
#include <string>
#include <vector>
#include <cctype>
#include <algorithm>
#include <sstream>
namespace boost {
namespace algorithm {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct shim_any_of { std::string set; bool operator()(char c) const { return set.find(c) != std::string::npos; } };
inline shim_any_of is_any_of(const std::string& s) { return shim_any_of{s}; }
inline void to_upper(std::string& s) { for (auto& c : s) c = (char)std::toupper((unsigned char)c); }
inline std::string to_upper_copy(const std::string& s) { std::string r(s); to_upper(r); return r; }
inline void to_lower(std::string& s) { for (auto& c : s) c = (char)std::tolower((unsigned char)c); }
inline std::string to_lower_copy(const std::string& s) { std::string r(s); to_lower(r); return r; }
inline bool iequals(const std::string& a, const std::string& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
inline void trim(std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e && std::isspace((unsigned char)s[b])) b++;
    while (e > b && std::isspace((unsigned char)s[e - 1])) e--;
    s = s.substr(b, e - b);
}
inline std::string trim_copy(const std::string& s) { std::string r(s); trim(r); return r; }
template <class Seq, class Pred>
inline Seq& split(Seq& out, const std::string& in, Pred pred, token_compress_mode_type mode = token_compress_off) {
    out.clear();
    std::string cur;
    bool lastSep = false;
    for (size_t i = 0; i < in.size(); i++) {
        if (pred(in[i])) {
            if (mode == token_compress_on && lastSep) continue;
            out.push_back(typename Seq::value_type(cur)); cur.clear(); lastSep = true;
        } else { cur.push_back(in[i]); lastSep = false; }
    }
    out.push_back(typename Seq::value_type(cur));
    return out;
}
template <class Seq>
inline std::string join(const Seq& parts, const std::string& sep) {
    std::ostringstream ss; bool first = true;
    for (const auto& p : parts) { if (!first) ss << sep; ss << p; first = false; }
    return ss.str();
}
}
using algorithm::token_compress_on; using algorithm::token_compress_off; using algorithm::is_any_of;
using algorithm::to_upper; using algorithm::to_upper_copy; using algorithm::to_lower; using algorithm::to_lower_copy;
using algorithm::iequals; using algorithm::trim; using algorithm::trim_copy; using algorithm::split; using algorithm::join;
}
