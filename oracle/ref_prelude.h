/*
 * oracle/ref_prelude.h -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * Declarations the extracted reference functions need from parts of /root/reference/src/viso.cpp that are NOT
 * extracted: the debug-image writers (viso.cpp:287-388, 516-649: stubbed as no-ops in ref_abi.inc, their output is
 * JPEG files nobody reads) and randomsample (viso.cpp:87-107), whose std::random_device-seeded sampler is replaced by
 * a host-supplied sample table exactly as BASELINE.json's north_star prescribes.  Signatures as in the reference.
 */
#ifndef VISO_ORACLE_REF_PRELUDE_H_
#define VISO_ORACLE_REF_PRELUDE_H_

void randomsample(int n, int N, std::vector<int>& samples); /* viso.cpp:87-88 */

void save1(const Mat& im, const KeyPoints& kp, const string& file_name, int lim = INT_MAX, Scalar color = Scalar(255, 0, 0),
           int thickness = 1, int linetype = -1);                                                            /* viso.cpp:310-312 */
void save1reproj(const Mat& im, const Mat& X, const Mat& x, const Mat& P, const string& file_name);          /* viso.cpp:380-381 */
void save2blend(const cv::Mat& im1, const cv::Mat& im2, const KeyPoints& kp1, const KeyPoints& kp2, const Matches& match,
                const string& file_name, int lim = INT_MAX);                                                 /* viso.cpp:545-548 */
void save2blend(const cv::Mat& im1, const cv::Mat& im2, const Mat& x, const string& file_name, int lim = INT_MAX); /* :572-574 */
void save4(const Mat& im1, const Mat& im1_prev, const Mat& im2, const Mat& im2_prev, const KeyPoints& kp1,
           const KeyPoints& kp1_prev, const KeyPoints& kp2, const KeyPoints& kp2_prev, const std::vector<cv::Vec4i>& circ_match,
           const string& file_name, int lim = INT_MAX);                                                      /* viso.cpp:616-622 */

#endif
