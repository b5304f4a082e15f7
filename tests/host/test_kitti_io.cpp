// CPU-only test of libviso_b200/host/kitti_io.h (reference src/kitti.cpp:23-64 formats): round trip + KITTI sample.
#include "../../libviso_b200/host/kitti_io.h"
#include <cmath>
#include <cstdlib>
#include <cstring>

#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

int main(int argc, char** argv)
{
    REQUIRE(argc == 2);
    const std::string dir = argv[1];
    // KITTI sequence 00 calib.txt (the constants of test/test.cpp:56-65)
    {
        FILE* f = std::fopen((dir + "/calib.txt").c_str(), "w");
        std::fprintf(f, "P0: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 0.000000000000e+00 0.000000000000e+00 "
                        "7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 0.000000000000e+00 0.000000000000e+00 "
                        "1.000000000000e+00 0.000000000000e+00\n"
                        "P1: 7.188560000000e+02 0.000000000000e+00 6.071928000000e+02 -3.861448000000e+02 0.000000000000e+00 "
                        "7.188560000000e+02 1.852157000000e+02 0.000000000000e+00 0.000000000000e+00 0.000000000000e+00 "
                        "1.000000000000e+00 0.000000000000e+00\nP2: 1 2 3\n");
        std::fclose(f);
    }
    cv::Mat P1, P2;
    REQUIRE(loadCalib(dir + "/calib.txt", P1, P2));
    REQUIRE(P1.at<double>(0, 0) == 718.856 && P1.at<double>(0, 2) == 607.1928 && P1.at<double>(1, 2) == 185.2157);
    REQUIRE(P2.at<double>(0, 3) == -386.1448 && P1.at<double>(0, 3) == 0 && P2.at<double>(2, 2) == 1);
    REQUIRE(!loadCalib(dir + "/missing.txt", P1, P2));
    std::vector<cv::Mat> poses;
    for (int i = 0; i < 3; ++i) {
        cv::Mat T = cv::Mat::eye(4, 4, CV_64F);
        T.at<double>(0, 3) = 0.25 * i; T.at<double>(2, 3) = 1.5 * i; T.at<double>(0, 1) = -1e-3 * i;
        poses.push_back(T);
    }
    REQUIRE(savePoses(dir + "/poses.txt", poses));
    FILE* f = std::fopen((dir + "/poses.txt").c_str(), "r");
    char line[512];
    int n = 0;
    while (std::fgets(line, sizeof(line), f)) {
        double v[12];
        REQUIRE(std::sscanf(line, "%lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf %lf", v, v + 1, v + 2, v + 3, v + 4, v + 5, v + 6,
                            v + 7, v + 8, v + 9, v + 10, v + 11) == 12);
        for (int k = 0; k < 12; ++k) REQUIRE(std::fabs(v[k] - poses[n].ptr<double>(0)[k]) < 1e-6); // "%lf" keeps 6 decimals
        ++n;
    }
    std::fclose(f);
    REQUIRE(n == 3);
    std::printf("kitti_io OK\n");
    return 0;
}
