/* compat/boost/log/sinks/text_file_backend.hpp -- see ../trivial.hpp */
#include "../trivial.hpp"
