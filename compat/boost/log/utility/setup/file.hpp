/* compat/boost/log/utility/setup/file.hpp -- see ../../trivial.hpp */
#include "../../trivial.hpp"
