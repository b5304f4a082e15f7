/*
 * mvg.h (B200) -- the reference's src/mvg.h restated with the same names and signatures:
 *   slice<T>               mvg.h:17-31      rows x cols sub-matrix
 *   arange                 mvg.cpp:33-42    0 .. range-1
 *   F_from_P<T>            mvg.h:41-66      fundamental matrix of two camera matrices: F(r,c) = det[X_c; Y_r], X_j / Y_j =
 *                                           P1 / P2 with row j left out (cyclic), via cv::determinant like the reference
 *   P_from_KRt<T>          mvg.cpp:92-107   K [R t] (defined in viso.cpp for T = float, double)
 *   triangulate_dlt        mvg.cpp:124-169  on the device (viso_triangulate_dlt)
 *   triangulate_rectified  mvg.cpp:172-192  on the device (viso_triangulate_rectified_f32); the pipeline's double version
 *                                           is the template declared in viso.h
 *   Camera, StereoCam      mvg.h:88-118
 */
#ifndef VISO_B200_HOST_MVG_H_
#define VISO_B200_HOST_MVG_H_

#include <iostream>
#include <vector>
#include <opencv2/core/core.hpp>
#include <Eigen/Dense>

#include "misc.h"

using namespace std;
using cv::Mat;
using cv::Mat_;
using cv::Vec4f;

using Eigen::MatrixXd;

template <class T> Mat slice(const Mat& x, const vector<int>& rows, const vector<int>& cols)
{
    Mat res((int)rows.size(), (int)cols.size(), x.type());
    for (size_t i = 0; i < rows.size(); ++i)
        for (size_t j = 0; j < cols.size(); ++j) res.at<T>((int)i, (int)j) = x.at<T>(rows[i], cols[j]);
    return res;
}

vector<int> arange(int range);

template <class T> Mat F_from_P(Mat P1, Mat P2)
{
    static const int pick[3][2] = {{1, 2}, {2, 0}, {0, 1}}; /* P with row j omitted, in the reference's cyclic order */
    const vector<int> all = arange(P1.cols);
    Mat F(3, 3, cv::DataType<T>::type);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            const Mat Xc = slice<T>(P1, vector<int>(pick[c], pick[c] + 2), all), Yr = slice<T>(P2, vector<int>(pick[r], pick[r] + 2), all);
            F.at<T>(r, c) = (T)determinant(vcat<T>(Xc, Yr));
        }
    return F;
}

/*** P = K*[R t] */
template <class T> Mat P_from_KRt(const Mat& K, const Mat& R, const Mat& t);

/* linear triangulation */
Mat triangulate_dlt(const Mat& x1, const Mat& x2, const Mat& P1, const Mat& P2);

Mat triangulate_rectified(const Mat& x1, /* pixel coordinates in the 1st image */
                          const Mat& x2, /* pixel coordinates in the 2nd image */
                          double f,      /* focal distance */
                          double base,   /* camera base line distance */
                          double c1u,    /* principal point, u */
                          double c1v     /* principal point, v */);

/* central projection */
class Camera {
public:
    Mat K;   // intrinsics
    Vec4f D; // distortion parameters
};

class StereoCam {
public:
    Camera c1, c2;
    Mat R, t; /* rotation / translation c1 -> c2 */
    Mat p1() const { return P_from_KRt<float>(c1.K, Mat::eye(3, 3, CV_32FC1), Mat::zeros(1, 3, CV_32FC1)); }
    Mat p2() const { return P_from_KRt<float>(c2.K, R, t); }
    Mat F() const { return F_from_P<float>(p1(), p2()); }

    Mat R1, R2; /* rotations that rectify the pair */
    Mat P1, P2; /* rectified camera matrices */
    Mat Q;
};

#endif // VISO_B200_HOST_MVG_H_
