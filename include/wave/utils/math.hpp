// wave::Mat6 / Vec6 / Affine3 as wave_utils/include/wave/utils/math.hpp:31-41 names them.  With
// Eigen installed these are the Eigen typedefs the reference uses; without it (this image) a
// small value-type stand-in with the handful of operations the matcher interface and its tests
// need: element access, Identity(), matrix(), translation() with comma initialisation,
// subtraction, Frobenius norm().
#ifndef WAVE_UTILS_MATH_HPP
#define WAVE_UTILS_MATH_HPP

#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && __has_include(<Eigen/Geometry>)
#define WAVE_HAVE_EIGEN 1
#endif
#endif

#ifdef WAVE_HAVE_EIGEN
#include <Eigen/Dense>
#include <Eigen/Geometry>
namespace wave {
typedef Eigen::Matrix<double, 6, 1> Vec6;
typedef Eigen::Matrix<double, 6, 6> Mat6;
typedef Eigen::Affine3d Affine3;
}  // namespace wave
#else

#include <array>
#include <cmath>
#include <cstddef>

#ifndef EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#endif

namespace Eigen {

template <int R, int C>
class SmallMatrix {
 public:
    SmallMatrix() { v_.fill(0.0); }
    static SmallMatrix Zero() { return SmallMatrix(); }
    static SmallMatrix Identity() {
        SmallMatrix m;
        for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = 1.0;
        return m;
    }
    static SmallMatrix Identity(int, int) { return Identity(); }
    double &operator()(int r, int c) { return v_[static_cast<std::size_t>(r * C + c)]; }
    double operator()(int r, int c) const { return v_[static_cast<std::size_t>(r * C + c)]; }
    double &operator()(int i) { return v_[static_cast<std::size_t>(i)]; }
    double operator()(int i) const { return v_[static_cast<std::size_t>(i)]; }
    int rows() const { return R; }
    int cols() const { return C; }
    double *data() { return v_.data(); }               // row major
    const double *data() const { return v_.data(); }
    SmallMatrix operator-(const SmallMatrix &o) const {
        SmallMatrix m;
        for (std::size_t i = 0; i < v_.size(); ++i) m.v_[i] = v_[i] - o.v_[i];
        return m;
    }
    SmallMatrix operator+(const SmallMatrix &o) const {
        SmallMatrix m;
        for (std::size_t i = 0; i < v_.size(); ++i) m.v_[i] = v_[i] + o.v_[i];
        return m;
    }
    template <int K>
    SmallMatrix<R, K> operator*(const SmallMatrix<C, K> &o) const {
        SmallMatrix<R, K> m;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < K; ++j) {
                double s = 0;
                for (int k = 0; k < C; ++k) s += (*this)(i, k) * o(k, j);
                m(i, j) = s;
            }
        return m;
    }
    double norm() const {  // Frobenius
        double s = 0;
        for (double x : v_) s += x * x;
        return std::sqrt(s);
    }

 private:
    std::array<double, static_cast<std::size_t>(R * C)> v_;
};

typedef SmallMatrix<4, 4> Matrix4d;
typedef SmallMatrix<3, 3> Matrix3d;
typedef SmallMatrix<3, 1> Vector3d;

// `m.translation() << x, y, z;`
class CommaFill3 {
 public:
    CommaFill3(double *a, double *b, double *c) : p_{a, b, c} {}
    CommaFill3 &operator<<(double v) { return put(v); }
    CommaFill3 &operator,(double v) { return put(v); }

 private:
    CommaFill3 &put(double v) {
        if (i_ < 3) *p_[i_++] = v;
        return *this;
    }
    double *p_[3];
    int i_ = 0;
};

class TranslationRef {
 public:
    explicit TranslationRef(Matrix4d &m) : m_(m) {}
    CommaFill3 operator<<(double v) {
        CommaFill3 f(&m_(0, 3), &m_(1, 3), &m_(2, 3));
        f << v;
        return f;
    }
    double x() const { return m_(0, 3); }
    double y() const { return m_(1, 3); }
    double z() const { return m_(2, 3); }
    double &operator()(int i) { return m_(i, 3); }

 private:
    Matrix4d &m_;
};

class Affine3d {
 public:
    Affine3d() : m_(Matrix4d::Identity()) {}
    Affine3d(const Matrix4d &m) : m_(m) {}  // NOLINT: the reference assigns matrices to transforms
    static Affine3d Identity() { return Affine3d(); }
    Matrix4d &matrix() { return m_; }
    const Matrix4d &matrix() const { return m_; }
    TranslationRef translation() { return TranslationRef(m_); }
    Matrix3d rotation() const {
        Matrix3d r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r(i, j) = m_(i, j);
        return r;
    }
    Affine3d &operator=(const Matrix4d &m) {
        m_ = m;
        return *this;
    }
    Affine3d operator*(const Affine3d &o) const { return Affine3d(m_ * o.m_); }

 private:
    Matrix4d m_;
};

}  // namespace Eigen

namespace wave {
typedef Eigen::SmallMatrix<6, 1> Vec6;
typedef Eigen::SmallMatrix<6, 6> Mat6;
typedef Eigen::Affine3d Affine3;
}  // namespace wave
#endif  // WAVE_HAVE_EIGEN

#endif  // WAVE_UTILS_MATH_HPP
