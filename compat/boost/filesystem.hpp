/* compat/boost/filesystem.hpp -- boost::filesystem::path as src/kitti.cpp and src/viso.h:139 use it: construction
 * from strings, operator/, string(), create_directories().  Used only when Boost is not installed. */
#ifndef VISO_COMPAT_BOOST_FILESYSTEM_HPP_
#define VISO_COMPAT_BOOST_FILESYSTEM_HPP_
#include "optional.hpp"
#include <ostream>
#include <string>
#include <sys/stat.h>
#include <sys/types.h>
namespace boost {
namespace filesystem {
class path {
public:
    path() {}
    path(const char* s) : s_(s ? s : "") {}
    path(const std::string& s) : s_(s) {}
    const std::string& string() const { return s_; }
    const char* c_str() const { return s_.c_str(); }
    bool empty() const { return s_.empty(); }
    path& operator/=(const path& o)
    {
        if (!s_.empty() && s_[s_.size() - 1] != '/' && !o.s_.empty() && o.s_[0] != '/') s_ += '/';
        s_ += o.s_;
        return *this;
    }
    path filename() const { const size_t p = s_.find_last_of('/'); return p == std::string::npos ? *this : path(s_.substr(p + 1)); }
    path parent_path() const { const size_t p = s_.find_last_of('/'); return p == std::string::npos ? path() : path(s_.substr(0, p)); }
private:
    std::string s_;
};
inline path operator/(const path& a, const path& b) { path r(a); r /= b; return r; }
inline std::ostream& operator<<(std::ostream& os, const path& p) { return os << '"' << p.string() << '"'; }
inline bool exists(const path& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool is_directory(const path& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool create_directories(const path& p)
{
    const std::string& s = p.string();
    bool made = false;
    for (size_t i = 1; i <= s.size(); ++i)
        if (i == s.size() || s[i] == '/') {
            const std::string sub = s.substr(0, i);
            if (!sub.empty() && ::mkdir(sub.c_str(), 0777) == 0) made = true;
        }
    return made;
}
inline bool create_directory(const path& p) { return ::mkdir(p.c_str(), 0777) == 0; }
} // namespace filesystem
} // namespace boost
#endif
