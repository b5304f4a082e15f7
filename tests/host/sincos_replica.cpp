// Host build of libviso_b200/csrc/glibc_sincos.h (the very text the device code compiles) for tests/test_sincos.py:
// g++ -O2 -ffp-contract=off -shared -fPIC.  fma() is the correctly rounded libm / hardware fused multiply-add.
#include "../../libviso_b200/csrc/glibc_sincos.h"

extern "C" void replica_sincos(const double* x, int n, double* s, double* c)
{
    for (int i = 0; i < n; ++i) { s[i] = viso_sc::sin_glibc(x[i]); c[i] = viso_sc::cos_glibc(x[i]); }
}

// libm's sin / cos called separately (no sincos merging: -fno-builtin-sin -fno-builtin-cos), like the reference's -O0 build
extern "C" void libm_sincos(const double* x, int n, double* s, double* c)
{
    for (int i = 0; i < n; ++i) { s[i] = sin(x[i]); c[i] = cos(x[i]); }
}
