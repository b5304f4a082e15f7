/* compat/boost/test/floating_point_comparison.hpp -- see unit_test.hpp */
#include "unit_test.hpp"
