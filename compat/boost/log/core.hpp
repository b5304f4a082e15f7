/* compat/boost/log/core.hpp -- see trivial.hpp */
#include "trivial.hpp"
