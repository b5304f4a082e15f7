// Host harness for tests/test_oracle_golden.py::test_sort_matches_is_libstdcxx_std_sort: runs the restated introsort (libviso_b200/csrc/introsort.h, the code
// the device executes) and libstdc++'s std::sort (what the reference calls at viso.cpp:724) on the same input.
#include <algorithm>
#include <cstring>
#include <vector>
#include "../libviso_b200/csrc/introsort.h"

extern "C" int introsort_matches_std(const int32_t* data, int n, int32_t* out_restated)
{
    std::vector<viso_sort::M3> a(n), b(n);
    if (n) { std::memcpy(a.data(), data, (size_t)n * 12); std::memcpy(b.data(), data, (size_t)n * 12); }
    viso_sort::sort(a.data(), n);
    std::sort(b.begin(), b.end(), [](const viso_sort::M3& x, const viso_sort::M3& y) { return x.d < y.d; });
    if (out_restated && n) std::memcpy(out_restated, a.data(), (size_t)n * 12);
    return n == 0 || std::memcmp(a.data(), b.data(), (size_t)n * 12) == 0;
}
