/* compat/boost/log/expressions.hpp -- see trivial.hpp */
#include "trivial.hpp"
