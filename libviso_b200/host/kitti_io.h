/*
 * kitti_io.h (B200) -- the two file formats of the reference's driver, same names and behaviour:
 *   loadCalib   reference src/kitti.cpp:23-46  KITTI calib.txt: "P0: <12 doubles>\nP1: <12 doubles>" -> P1, P2 (3x4 CV_64F)
 *   savePoses   reference src/kitti.cpp:49-64  one line per pose: the top 3x4 of the 4x4 matrix, "%lf" x 12
 * Host-side text I/O; the poses come from sequence_odometry() / viso_chain_poses().
 */
#ifndef VISO_B200_HOST_KITTI_IO_H_
#define VISO_B200_HOST_KITTI_IO_H_

#include <opencv2/core/core.hpp>

#include <cstdio>
#include <string>
#include <vector>

inline bool loadCalib(const std::string& file_name, cv::Mat& p1, cv::Mat& p2)
{
    FILE* fp = std::fopen(file_name.c_str(), "r");
    if (!fp) return false;
    if (p1.rows != 3 || p1.cols != 4 || p1.type() != CV_64F) p1.create(3, 4, CV_64F);
    if (p2.rows != 3 || p2.cols != 4 || p2.type() != CV_64F) p2.create(3, 4, CV_64F);
    bool ok = true;
    cv::Mat* ps[2] = {&p1, &p2};
    for (int c = 0; c < 2 && ok; ++c) {
        int n = 0;
        ok = std::fscanf(fp, " P%d:", &n) == 1; /* kitti.cpp:29,37 */
        double* d = ps[c]->ptr<double>(0);
        for (int i = 0; i < 12 && ok; ++i) ok = std::fscanf(fp, "%lf", d + i) == 1;
    }
    std::fclose(fp);
    return ok;
}

inline bool savePoses(const std::string& file_name, const std::vector<cv::Mat>& poses)
{
    FILE* fp = std::fopen(file_name.c_str(), "w+");
    if (!fp) return false; /* the reference evaluates `false;` and carries on (kitti.cpp:52); a null FILE* would crash */
    for (const cv::Mat& pose : poses) {
        const double* d = pose.ptr<double>(0);
        if (std::fprintf(fp, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf\n", d[0], d[1], d[2], d[3], d[4], d[5],
                         d[6], d[7], d[8], d[9], d[10], d[11]) <= 0) {
            std::fclose(fp);
            return false;
        }
    }
    std::fclose(fp);
    return true;
}

#endif
