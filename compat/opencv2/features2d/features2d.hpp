/* compat/opencv2/features2d/features2d.hpp -- the two abstract bases the reference's front-end classes derive from
 * (HarrisBinnedFeatureDetector viso.cpp:911, MyFeatureExtractor viso.cpp:981), with OpenCV 3.0-alpha's dispatch:
 * detect() -> detectImpl(), compute() -> computeImpl() (KeyPointsFilter::runByImageBorder with a zero border and
 * runByKeypointSize(epsilon) keep every keypoint the reference produces: size = 11). */
#ifndef VISO_COMPAT_OPENCV2_FEATURES2D_HPP_
#define VISO_COMPAT_OPENCV2_FEATURES2D_HPP_
#include "../core/core.hpp"
namespace cv {

class FeatureDetector {
public:
    virtual ~FeatureDetector() {}
    void detect(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = Mat()) const
    {
        keypoints.clear();
        if (image.empty()) return;
        detectImpl(image, keypoints, mask);
    }
    virtual bool empty() const { return false; }

protected:
    virtual void detectImpl(InputArray image, std::vector<KeyPoint>& keypoints, InputArray mask = Mat()) const = 0;
};

class DescriptorExtractor {
public:
    virtual ~DescriptorExtractor() {}
    void compute(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) const
    {
        if (image.empty() || keypoints.empty()) { descriptors.release(); return; }
        std::vector<KeyPoint> kept;
        for (size_t i = 0; i < keypoints.size(); ++i)
            if (keypoints[i].size >= FLT_EPSILON && keypoints[i].size <= FLT_MAX) kept.push_back(keypoints[i]);
        keypoints.swap(kept);
        computeImpl(image, keypoints, descriptors);
    }
    virtual int descriptorSize() const = 0;
    virtual int descriptorType() const = 0;
    virtual int defaultNorm() const = 0;
    virtual bool empty() const { return false; }

protected:
    virtual void computeImpl(InputArray image, std::vector<KeyPoint>& keypoints, OutputArray descriptors) const = 0;
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(FLT_MAX) {}
};

} // namespace cv
#endif
