/*
 * compat/boost/test/unit_test.hpp -- the part of Boost.Test the reference's test/test.cpp uses: BOOST_TEST_MODULE
 * (defines main), BOOST_AUTO_TEST_CASE, BOOST_REQUIRE_EQUAL, BOOST_CHECK_SMALL and their closest relatives.  Test cases
 * register themselves and run in definition order; a failed REQUIRE ends its case; the exit status is the number
 * of failed cases (0 = all passed), like Boost's runner.  Used only when Boost is not installed.
 */
#ifndef VISO_COMPAT_BOOST_TEST_UNIT_TEST_HPP_
#define VISO_COMPAT_BOOST_TEST_UNIT_TEST_HPP_
#include <cmath>
#include <cstdio>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost {
namespace unit_test {
struct test_case { const char* name; void (*fn)(); };
inline std::vector<test_case>& registry() { static std::vector<test_case> r; return r; }
struct registrar { registrar(const char* n, void (*f)()) { test_case t = {n, f}; registry().push_back(t); } };
struct require_failed : std::runtime_error { require_failed() : std::runtime_error("BOOST_REQUIRE failed") {} };
inline int& failures() { static int f = 0; return f; }
inline void report(bool ok, bool fatal, const char* expr, const char* file, int line)
{
    if (ok) return;
    ++failures();
    std::cerr << file << "(" << line << "): " << (fatal ? "fatal error" : "error") << ": check " << expr << " has failed" << std::endl;
    if (fatal) throw require_failed();
}
inline int run_all(const char* module)
{
    int failed_cases = 0;
    std::cout << "Running " << registry().size() << " test case" << (registry().size() == 1 ? "" : "s") << "..." << std::endl;
    for (size_t i = 0; i < registry().size(); ++i) {
        const int before = failures();
        try { registry()[i].fn(); }
        catch (const require_failed&) {}
        catch (const std::exception& e) { ++failures(); std::cerr << "exception in \"" << registry()[i].name << "\": " << e.what() << std::endl; }
        if (failures() != before) ++failed_cases;
    }
    if (failed_cases) std::cerr << "\n*** " << failed_cases << " failure" << (failed_cases == 1 ? "" : "s") << " detected in test suite \"" << module << "\"" << std::endl;
    else std::cout << "\n*** No errors detected" << std::endl;
    return failed_cases;
}
} // namespace unit_test
} // namespace boost
#define BOOST_AUTO_TEST_CASE(name)                                                        \
    static void name();                                                                   \
    static ::boost::unit_test::registrar name##_registrar(#name, &name);                  \
    static void name()
#define BOOST_CHECK(e) ::boost::unit_test::report(!!(e), false, #e, __FILE__, __LINE__)
#define BOOST_REQUIRE(e) ::boost::unit_test::report(!!(e), true, #e, __FILE__, __LINE__)
#define BOOST_CHECK_EQUAL(a, b) ::boost::unit_test::report((a) == (b), false, #a " == " #b, __FILE__, __LINE__)
#define BOOST_REQUIRE_EQUAL(a, b) ::boost::unit_test::report((a) == (b), true, #a " == " #b, __FILE__, __LINE__)
#define BOOST_CHECK_SMALL(v, tol) ::boost::unit_test::report(std::fabs(v) < (tol), false, "|" #v "| < " #tol, __FILE__, __LINE__)
#define BOOST_REQUIRE_SMALL(v, tol) ::boost::unit_test::report(std::fabs(v) < (tol), true, "|" #v "| < " #tol, __FILE__, __LINE__)
#define BOOST_CHECK_CLOSE(a, b, pct) ::boost::unit_test::report(std::fabs((a) - (b)) <= std::fabs(b) * (pct) / 100.0, false, #a " ~ " #b, __FILE__, __LINE__)
#define BOOST_CHECK_MESSAGE(e, m) BOOST_CHECK(e)
#define BOOST_REQUIRE_MESSAGE(e, m) BOOST_REQUIRE(e)
#define BOOST_TEST_MESSAGE(m) (std::cout << m << std::endl)
#ifdef BOOST_TEST_MODULE
#define VISO_COMPAT_STR2(x) #x
#define VISO_COMPAT_STR(x) VISO_COMPAT_STR2(x)
int main(int, char**) { return ::boost::unit_test::run_all(VISO_COMPAT_STR(BOOST_TEST_MODULE)); }
#endif
#endif
