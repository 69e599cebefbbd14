// Test-infrastructure only: stand-in for <boost/filesystem.hpp> on top of std::filesystem
// (oracle build, see oracle/README.md).
#pragma once
#include <boost/exception/all.hpp>
#include <filesystem>
#include <string>
#include <ctime>
#include <sys/stat.h>
namespace boost { namespace filesystem {
class path : public std::filesystem::path {
public:
    typedef std::filesystem::path base;
    path() {}
    path(const base& p) : base(p) {}
    path(const std::string& s) : base(s) {}
    path(const char* s) : base(s) {}
    path leaf() const { return path(filename()); }
    path parent_path() const { return path(base::parent_path()); }
    path filename() const { return path(base::filename()); }
    path stem() const { return path(base::stem()); }
    path extension() const { return path(base::extension()); }
    path& operator/=(const path& o) { base::operator/=(static_cast<const base&>(o)); return *this; }
    path& operator/=(const std::string& o) { base::operator/=(o); return *this; }
    path& operator/=(const char* o) { base::operator/=(o); return *this; }
};
inline path operator/(const path& a, const path& b) { path r(a); r /= b; return r; }
inline path operator/(const path& a, const std::string& b) { path r(a); r /= b; return r; }
inline path operator/(const path& a, const char* b) { path r(a); r /= b; return r; }
inline bool exists(const path& p) { std::error_code ec; return std::filesystem::exists(p, ec); }
inline bool is_directory(const path& p) { std::error_code ec; return std::filesystem::is_directory(p, ec); }
inline bool create_directories(const path& p) { std::error_code ec; std::filesystem::create_directories(p, ec); return is_directory(p); }
inline bool create_directory(const path& p) { std::error_code ec; std::filesystem::create_directory(p, ec); return is_directory(p); }
inline bool remove(const path& p) { std::error_code ec; return std::filesystem::remove(p, ec); }
inline std::uintmax_t remove_all(const path& p) { std::error_code ec; return std::filesystem::remove_all(p, ec); }
inline void rename(const path& a, const path& b) { std::filesystem::rename(a, b); }
inline void copy_file(const path& a, const path& b) { std::filesystem::copy_file(a, b); }
inline void create_symlink(const path& to, const path& link) { std::filesystem::create_symlink(to, link); }
inline path canonical(const path& p) { return path(std::filesystem::canonical(p)); }
inline path absolute(const path& p) { return path(std::filesystem::absolute(p)); }
inline path system_complete(const path& p) { return path(std::filesystem::absolute(p)); }
inline path current_path() { return path(std::filesystem::current_path()); }
inline bool symbolic_link_exists(const path& p) { std::error_code ec; return std::filesystem::is_symlink(std::filesystem::symlink_status(p, ec)); }
inline bool is_symlink(const path& p) { return symbolic_link_exists(p); }
inline std::time_t last_write_time(const path& p) { struct stat st; if (::stat(p.c_str(), &st) != 0) return 0; return st.st_mtime; }
inline std::uintmax_t file_size(const path& p) { return std::filesystem::file_size(p); }
}}
