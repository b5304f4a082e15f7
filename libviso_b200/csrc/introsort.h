/*
 * introsort.h -- restatement of libstdc++'s std::sort (bits/stl_algo.h: __introsort_loop,
 * __unguarded_partition_pivot, __move_median_to_first, __partial_sort/heap ops, __final_insertion_sort)
 * for arrays of Match = (i1, i2, dist) compared on dist only.
 *
 * Why: the reference orders matches with an UNSTABLE std::sort (viso.cpp:724); the order of equal-distance
 * matches decides the column order of the RANSAC inputs, which the reference's weight indexing quirk
 * (viso.cpp:1449) and the sample table make observable.  Bit-exact parity therefore needs the exact permutation
 * libstdc++ produces.  This file restates the algorithm (no recursion, explicit stack) so the SAME code runs in a
 * CUDA thread on the device and under g++ in tests/test_oracle_golden.py (tests/introsort_check.cpp), where it is compared with std::sort itself; the
 * device additionally runs a warp-parallel equivalent (sort_circle.cu) checked against this one.
 */
#ifndef VISO_INTROSORT_H_
#define VISO_INTROSORT_H_

#include <stdint.h>

#ifdef __CUDACC__
#define VISO_HD __host__ __device__ __forceinline__
#else
#define VISO_HD inline
#endif

namespace viso_sort {

struct M3 { int32_t a, b, d; };
/* (dist, payload): sorts exactly like M3 because the algorithm only looks at d */
struct KV { int32_t d, pos; };

template <class T> VISO_HD bool less(const T& x, const T& y) { return x.d < y.d; }
template <class T> VISO_HD void swp(T* p, int i, int j) { T t = p[i]; p[i] = p[j]; p[j] = t; }

/* std::__lg */
VISO_HD int lg(int n) { int k = 0; while (n > 1) { n >>= 1; ++k; } return k; }

/* std::__push_heap */
template <class T> VISO_HD void push_heap_(T* first, int hole, int top, T value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && less(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

/* std::__adjust_heap */
template <class T> VISO_HD void adjust_heap_(T* first, int hole, int len, T value)
{
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (less(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value);
}

/* std::__partial_sort(first, last, last) == __heap_select (make_heap only, middle==last) + __sort_heap */
template <class T> VISO_HD void heap_sort_(T* first, int len)
{
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            T v = first[parent];
            adjust_heap_(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        T v = first[last];          /* __pop_heap(first, last, last) */
        first[last] = first[0];
        adjust_heap_(first, 0, last, v);
    }
}

/* std::__move_median_to_first(result, a, b, c) */
template <class T> VISO_HD void median_to_first_(T* p, int result, int a, int b, int c)
{
    if (less(p[a], p[b])) {
        if (less(p[b], p[c])) swp(p, result, b);
        else if (less(p[a], p[c])) swp(p, result, c);
        else swp(p, result, a);
    } else if (less(p[a], p[c])) swp(p, result, a);
    else if (less(p[b], p[c])) swp(p, result, c);
    else swp(p, result, b);
}

/* std::__unguarded_partition(first, last, pivot) */
template <class T> VISO_HD int unguarded_partition_(T* p, int first, int last, int pivot)
{
    while (true) {
        while (less(p[first], p[pivot])) ++first;
        --last;
        while (less(p[pivot], p[last])) --last;
        if (!(first < last)) return first;
        swp(p, first, last);
        ++first;
    }
}

/* std::__unguarded_linear_insert */
template <class T> VISO_HD void unguarded_linear_insert_(T* p, int last)
{
    T val = p[last];
    int next = last - 1;
    while (less(val, p[next])) {
        p[last] = p[next];
        last = next;
        --next;
    }
    p[last] = val;
}

/* std::__insertion_sort */
template <class T> VISO_HD void insertion_sort_(T* p, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (less(p[i], p[first])) {
            T val = p[i];
            for (int k = i; k > first; --k) p[k] = p[k - 1]; /* std::move_backward(first, i, i+1) */
            p[first] = val;
        } else
            unguarded_linear_insert_(p, i);
    }
}

/* std::sort(p, p+n, less) */
template <class T> VISO_HD void sort(T* p, int n)
{
    const int S_threshold = 16;
    if (n <= 0) return;
    /* __introsort_loop(first, last, 2*lg(n)) with the recursion on [cut,last) turned into a stack */
    int stk_first[64], stk_last[64], stk_depth[64];
    int sp = 0;
    stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = lg(n) * 2; sp = 1;
    while (sp > 0) {
        --sp;
        int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
        while (last - first > S_threshold) {
            if (depth == 0) {
                heap_sort_(p + first, last - first);
                break;
            }
            --depth;
            int mid = first + (last - first) / 2;
            median_to_first_(p, first, first + 1, mid, last - 1);
            int cut = unguarded_partition_(p, first + 1, last, first);
            /* recursive call on [cut,last) with the decremented depth; loop continues on [first,cut) */
            stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; ++sp;
            last = cut;
        }
    }
    /* __final_insertion_sort */
    if (n > S_threshold) {
        insertion_sort_(p, 0, S_threshold);
        for (int i = S_threshold; i != n; ++i) unguarded_linear_insert_(p, i);
    } else
        insertion_sort_(p, 0, n);
}

} // namespace viso_sort
#endif
