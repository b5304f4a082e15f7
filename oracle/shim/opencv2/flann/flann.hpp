/*
 * oracle/shim/opencv2/flann/flann.hpp -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * cvflann::Index<L1<float>> with LinearIndexParams as match_desc uses it (viso.cpp:682-685, :171-186): brute-force
 * radius search.  FLANN is third-party code absent from /root/reference; restated from its published algorithm
 * (flann/nn_index.h radiusSearch + result_set.h RadiusUniqueResultSet): every point with L1 distance <= radius enters
 * a std::set ordered by (distance, index); the first `max_nn` are copied out in that order; the return value is the
 * size of the set, which may exceed what was copied.  Pinned against OpenCV 4.13's flann (tests/golden/
 * flann_radius.npz, tests/test_ref_pin.py).
 */
#ifndef VISO_ORACLE_SHIM_OPENCV2_FLANN_FLANN_HPP_
#define VISO_ORACLE_SHIM_OPENCV2_FLANN_FLANN_HPP_

#include <opencv2/core/core.hpp>
#include <set>

namespace cvflann {

template <class T> class Matrix {
public:
    size_t rows, cols, stride;
    T* data;
    Matrix() : rows(0), cols(0), stride(0), data(0) {}
    Matrix(T* d, size_t r, size_t c, size_t s = 0) : rows(r), cols(c), stride(s ? s : c), data(d) {}
    T* operator[](size_t i) const { return data + i * stride; }
};

template <class T> struct L1 {
    typedef T ElementType;
    typedef float ResultType;
    /* flann/dist.h L1::operator(): four differences per step, summed left to right, then the tail */
    ResultType operator()(const T* a, const T* b, size_t size) const
    {
        ResultType result = ResultType();
        const T* last = a + size;
        const T* lastgroup = last - 3;
        while (a < lastgroup) {
            const ResultType d0 = (ResultType)std::abs(a[0] - b[0]), d1 = (ResultType)std::abs(a[1] - b[1]);
            const ResultType d2 = (ResultType)std::abs(a[2] - b[2]), d3 = (ResultType)std::abs(a[3] - b[3]);
            result += d0 + d1 + d2 + d3;
            a += 4; b += 4;
        }
        while (a < last) { result += (ResultType)std::abs(*a++ - *b++); }
        return result;
    }
};

struct IndexParams {};
struct LinearIndexParams : IndexParams {};
struct KDTreeIndexParams : IndexParams { explicit KDTreeIndexParams(int = 4) {} };
struct SearchParams {
    int checks; float eps; bool sorted;
    SearchParams(int c = 32, float e = 0, bool s = true) : checks(c), eps(e), sorted(s) {}
};

template <class Distance> class Index {
public:
    typedef typename Distance::ElementType ElementType;
    typedef typename Distance::ResultType DistanceType;
    Index(const Matrix<ElementType>& dataset, const IndexParams&, Distance d = Distance()) : data_(dataset), dist_(d) {}
    void buildIndex() {}
    int radiusSearch(const Matrix<ElementType>& query, Matrix<int>& indices, Matrix<DistanceType>& dists, float radius,
                     const SearchParams& params)
    {
        if (query.rows != 1) return -1;
        std::set<std::pair<DistanceType, int> > found;
        for (size_t i = 0; i < data_.rows; ++i) {
            const DistanceType d = dist_(data_[i], query[0], data_.cols);
            if (d <= (DistanceType)radius) found.insert(std::make_pair(d, (int)i));
        }
        if (indices.cols > 0) {
            size_t k = 0;
            for (typename std::set<std::pair<DistanceType, int> >::const_iterator it = found.begin();
                 it != found.end() && k < indices.cols; ++it, ++k) {
                indices[0][k] = it->second;
                dists[0][k] = it->first;
            }
            (void)params; /* sorted copy either way: the set is ordered */
        }
        return (int)found.size();
    }

private:
    Matrix<ElementType> data_;
    Distance dist_;
};

} // namespace cvflann

namespace cv { namespace flann {
/* named by dead helpers only (viso.cpp:147-169, 728-796) */
class Index;
struct SearchParams { explicit SearchParams(int = 32) {} };
} }
#endif
