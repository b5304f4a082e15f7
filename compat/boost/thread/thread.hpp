/* compat/boost/thread/thread.hpp -- included by src/viso.cpp:24; nothing of it is used */
#ifndef VISO_COMPAT_BOOST_THREAD_HPP_
#define VISO_COMPAT_BOOST_THREAD_HPP_
#endif
