/* compat/boost/optional.hpp -- boost::optional<T> as src/viso.h:84, 104 uses it (value or nothing, tested in a
 * condition, dereferenced).  Used only when Boost is not installed. */
#ifndef VISO_COMPAT_BOOST_OPTIONAL_HPP_
#define VISO_COMPAT_BOOST_OPTIONAL_HPP_
#include <cassert>
namespace boost {
struct none_t {};
const none_t none = none_t();
template <class T> class optional {
public:
    optional() : has_(false), v_() {}
    optional(none_t) : has_(false), v_() {}
    optional(const T& v) : has_(true), v_(v) {}
    optional& operator=(const T& v) { v_ = v; has_ = true; return *this; }
    optional& operator=(none_t) { has_ = false; v_ = T(); return *this; }
    explicit operator bool() const { return has_; }
    bool operator!() const { return !has_; }
    bool is_initialized() const { return has_; }
    T& operator*() { assert(has_); return v_; }
    const T& operator*() const { assert(has_); return v_; }
    T* operator->() { assert(has_); return &v_; }
    const T* operator->() const { assert(has_); return &v_; }
    T& get() { assert(has_); return v_; }
    const T& get() const { assert(has_); return v_; }
private:
    bool has_;
    T v_;
};
} // namespace boost
#endif
