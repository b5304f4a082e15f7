/* compat/boost/make_shared.hpp -- included by src/viso.cpp:25; nothing of it is used */
#ifndef VISO_COMPAT_BOOST_MAKE_SHARED_HPP_
#define VISO_COMPAT_BOOST_MAKE_SHARED_HPP_
#endif
