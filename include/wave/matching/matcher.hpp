// wave::Matcher<T> - the abstract scan-matcher interface, source compatible with the reference
// (wave_matching/include/wave/matching/matcher.hpp:23-99): same constructors, accessors, virtuals
// and protected members.  Convention established by the reference tests (tests/icp_tests.cpp:31-32,
// 59): with target = P * ref, match() yields result ~= P, i.e. result maps ref into target.
#ifndef WAVE_MATCHING_MATCHER_HPP
#define WAVE_MATCHING_MATCHER_HPP

#include "wave/utils/utils.hpp"

namespace wave {

template <typename T>
class Matcher {
 public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW

    /// res: edge length of the voxel filter applied before matching (<= 0: full resolution)
    Matcher(float res) : resolution(res) {}  // implicit, as matcher.hpp:32
    Matcher() : resolution(-1) {}
    virtual ~Matcher() {}

    const Eigen::Affine3d getResult() { return this->result; }
    const Mat6 &getInfo() { return this->information; }
    float getRes() { return this->resolution; }

    /// The initial transform is always identity: pre-transform the target if a guess exists.
    virtual void setRef(const T &ref) = 0;
    virtual void setTarget(const T &target) = 0;
    void setup(const T &ref, const T &target) {
        this->setRef(ref);
        this->setTarget(target);
    }

    /// true if the match succeeded
    virtual bool match() { return false; }

    virtual void estimateInfo() { this->information = Mat6::Identity(6, 6); }

 protected:
    float resolution;     ///< voxel edge length, -1 when no down-sampling happens
    Affine3 result;       ///< estimated transform
    Mat6 information;     ///< 6x6 information matrix (x, y, z, then rotations)
};

}  // namespace wave

#endif  // WAVE_MATCHING_MATCHER_HPP
