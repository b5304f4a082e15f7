/*
 * compat/boost/log/trivial.hpp -- BOOST_LOG_TRIVIAL(severity) << ... and the few Boost.Log set-up calls src/kitti.cpp
 * names in init_log() (never called: kitti.cpp:95).  A record is written to std::clog when it is complete; the
 * threshold comes from the environment variable VISO_LOG_LEVEL (trace, debug, info, warning, error, fatal; default
 * info, Boost's trivial logger prints everything).  Used only when Boost is not installed.
 */
#ifndef VISO_COMPAT_BOOST_LOG_TRIVIAL_HPP_
#define VISO_COMPAT_BOOST_LOG_TRIVIAL_HPP_
#include "../assert.hpp"
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
namespace boost {
namespace log {
namespace trivial {
enum severity_level { trace, debug, info, warning, error, fatal };
struct filter_expr { severity_level min; };
struct severity_keyword {};
const severity_keyword severity = severity_keyword();
inline filter_expr operator>=(const severity_keyword&, severity_level l) { filter_expr f = {l}; return f; }
inline severity_level& threshold()
{
    static severity_level t = [] {
        const char* e = std::getenv("VISO_LOG_LEVEL");
        static const char* names[] = {"trace", "debug", "info", "warning", "error", "fatal"};
        if (e)
            for (int i = 0; i < 6; ++i)
                if (!std::strcmp(e, names[i])) return (severity_level)i;
        return info;
    }();
    return t;
}
class record_pump {
public:
    explicit record_pump(severity_level l) : lvl_(l), on_(l >= threshold()) {}
    ~record_pump()
    {
        static const char* names[] = {"trace", "debug", "info", "warning", "error", "fatal"};
        if (on_) std::clog << "[" << names[lvl_] << "] " << ss_.str() << std::endl;
    }
    template <class T> record_pump& operator<<(const T& v) { if (on_) ss_ << v; return *this; }
    record_pump& operator<<(std::ostream& (*m)(std::ostream&)) { if (on_) ss_ << m; return *this; }
private:
    severity_level lvl_;
    bool on_;
    std::ostringstream ss_;
};
} // namespace trivial
namespace sources {}
namespace sinks {
namespace file {
struct rotation_at_time_point { rotation_at_time_point(int, int, int) {} };
}
}
namespace keywords {
struct keyword { template <class T> const keyword& operator=(const T&) const { return *this; } };
const keyword file_name = keyword(), rotation_size = keyword(), time_based_rotation = keyword(), format = keyword(), auto_flush = keyword();
}
template <class... A> inline void add_file_log(const A&...) {}
class core {
public:
    static core* get() { static core c; return &c; }
    void set_filter(const trivial::filter_expr& f) { trivial::threshold() = f.min; }
};
} // namespace log
} // namespace boost
#define BOOST_LOG_TRIVIAL(lvl) ::boost::log::trivial::record_pump(::boost::log::trivial::lvl)
#endif
