/* compat/boost/assert.hpp -- BOOST_ASSERT / BOOST_ASSERT_MSG (abort with the message unless NDEBUG) */
#ifndef VISO_COMPAT_BOOST_ASSERT_HPP_
#define VISO_COMPAT_BOOST_ASSERT_HPP_
#include <cstdio>
#include <cstdlib>
#ifdef NDEBUG
#define BOOST_ASSERT(expr) ((void)0)
#define BOOST_ASSERT_MSG(expr, msg) ((void)0)
#else
#define BOOST_ASSERT(expr) ((expr) ? (void)0 : (fprintf(stderr, "%s:%d: assertion failed: %s\n", __FILE__, __LINE__, #expr), abort()))
#define BOOST_ASSERT_MSG(expr, msg) ((expr) ? (void)0 : (fprintf(stderr, "%s:%d: assertion failed: %s: %s\n", __FILE__, __LINE__, #expr, (msg)), abort()))
#endif
#endif
