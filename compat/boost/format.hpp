/*
 * compat/boost/format.hpp -- stand-in for the parts of Boost.Format / Boost.Optional the reference's headers and
 * callers use (src/viso.h:23, 81-119; src/kitti.cpp; src/estimation.cpp:32-39): printf-style directives fed with
 * operator%, str(), stream output.  Used only when Boost is not installed.
 */
#ifndef VISO_COMPAT_BOOST_FORMAT_HPP_
#define VISO_COMPAT_BOOST_FORMAT_HPP_

#include "optional.hpp"

#include <cstdio>
#include <ostream>
#include <sstream>
#include <string>
#include <type_traits>

namespace boost {

class format {
public:
    format() : pos_(0) {}
    format(const char* f) : fmt_(f ? f : ""), pos_(0) { flush_literals(); }
    format(const std::string& f) : fmt_(f), pos_(0) { flush_literals(); }

    template <class T> format& operator%(const T& v)
    {
        feed(v);
        flush_literals();
        return *this;
    }
    std::string str() const { return out_ + fmt_.substr(pos_); }

private:
    /* copy literal text (and %% escapes) up to the next directive */
    void flush_literals()
    {
        while (pos_ < fmt_.size()) {
            if (fmt_[pos_] != '%') { out_ += fmt_[pos_++]; continue; }
            if (pos_ + 1 < fmt_.size() && fmt_[pos_ + 1] == '%') { out_ += '%'; pos_ += 2; continue; }
            break;
        }
    }
    /* the directive at pos_: flags / width / precision text and the conversion character */
    bool directive(std::string& spec, char& conv)
    {
        if (pos_ >= fmt_.size() || fmt_[pos_] != '%') return false;
        size_t p = pos_ + 1;
        while (p < fmt_.size() && std::string("-+ #0123456789.").find(fmt_[p]) != std::string::npos) ++p;
        spec = fmt_.substr(pos_ + 1, p - pos_ - 1);
        while (p < fmt_.size() && (fmt_[p] == 'l' || fmt_[p] == 'h' || fmt_[p] == 'z')) ++p; /* length modifiers */
        conv = p < fmt_.size() ? fmt_[p] : 's';
        if (conv == '%') { /* positional %N% */
            spec.clear();
            conv = 's';
        }
        pos_ = p < fmt_.size() ? p + 1 : p;
        return true;
    }
    template <class T> typename std::enable_if<std::is_integral<T>::value>::type feed(const T& v)
    {
        std::string spec; char conv;
        if (!directive(spec, conv)) return;
        char buf[128];
        if (conv == 'x' || conv == 'X' || conv == 'o' || conv == 'u' || std::is_unsigned<T>::value) {
            const std::string f = "%" + spec + "ll" + (conv == 'x' || conv == 'X' || conv == 'o' ? std::string(1, conv) : std::string("u"));
            snprintf(buf, sizeof(buf), f.c_str(), (unsigned long long)v);
        } else if (conv == 'c') {
            snprintf(buf, sizeof(buf), ("%" + spec + "c").c_str(), (int)v);
        } else {
            snprintf(buf, sizeof(buf), ("%" + spec + "lld").c_str(), (long long)v);
        }
        out_ += buf;
    }
    template <class T> typename std::enable_if<std::is_floating_point<T>::value>::type feed(const T& v)
    {
        std::string spec; char conv;
        if (!directive(spec, conv)) return;
        if (std::string("eEfFgG").find(conv) == std::string::npos) conv = 'g';
        char buf[512];
        snprintf(buf, sizeof(buf), ("%" + spec + conv).c_str(), (double)v);
        out_ += buf;
    }
    template <class T> typename std::enable_if<!std::is_arithmetic<T>::value>::type feed(const T& v)
    {
        std::string spec; char conv;
        if (!directive(spec, conv)) return;
        std::ostringstream ss;
        ss << v;
        const std::string s = ss.str();
        size_t width = 0;
        bool left = false;
        for (size_t i = 0; i < spec.size(); ++i) {
            if (spec[i] == '-') left = true;
            else if (spec[i] >= '0' && spec[i] <= '9') { width = std::strtoul(spec.c_str() + i, 0, 10); break; }
        }
        if (s.size() < width && !left) out_ += std::string(width - s.size(), ' ');
        out_ += s;
        if (s.size() < width && left) out_ += std::string(width - s.size(), ' ');
    }

    std::string fmt_, out_;
    size_t pos_;
};

inline std::string str(const format& f) { return f.str(); }
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }

} // namespace boost
#endif
