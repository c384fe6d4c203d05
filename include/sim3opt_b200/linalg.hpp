// linalg.hpp -- small fixed-size matrix / quaternion types for the host-side facade.
//
// The reference builds its graphs with Eigen (Eigen::Matrix<double,7,7>, Eigen::Quaterniond, ...;
// kitti_surf.cpp:592-594, :606-609).  Eigen is not a dependency of this library, so the facade
// ships the handful of operations those call sites use, with Eigen's spelling.  Everything here is
// host code for building and reading graphs; no optimisation arithmetic runs through it.
#pragma once
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <memory>
#include <ostream>

namespace s3o {

template <class T, int R, int C>
class Matrix;

// writable view of N consecutive coefficients of a vector (v.head<3>(), v.tail<3>())
template <class T, int N>
class Segment {
public:
    explicit Segment(T *p) : p_(p) {}
    Segment &operator=(const Matrix<T, N, 1> &v) {
        for (int i = 0; i < N; ++i) p_[i] = v[i];
        return *this;
    }
    operator Matrix<T, N, 1>() const {
        Matrix<T, N, 1> out;
        for (int i = 0; i < N; ++i) out[i] = p_[i];
        return out;
    }
    T &operator[](int i) { return p_[i]; }
    T operator[](int i) const { return p_[i]; }

private:
    T *p_;
};

template <class T, int R, int C>
class Matrix {
public:
    static constexpr int Rows = R, Cols = C;
    Matrix() { for (int i = 0; i < R * C; ++i) a_[i] = T(0); }
    Matrix(std::initializer_list<T> v) {
        int i = 0;
        for (T x : v) if (i < R * C) a_[i++] = x;
        for (; i < R * C; ++i) a_[i] = T(0);
    }
    // Vector3d(x, y, z) / Vector4d(a, b, c, d)
    Matrix(T x, T y, T z) { static_assert(R * C == 3, "3-vector"); a_[0] = x; a_[1] = y; a_[2] = z; }
    Matrix(T x, T y, T z, T w) { static_assert(R * C == 4, "4-vector"); a_[0] = x; a_[1] = y; a_[2] = z; a_[3] = w; }
    template <int N>
    Matrix(const Segment<T, N> &s) { static_assert(N == R && C == 1, "segment size"); for (int i = 0; i < N; ++i) a_[i] = s[i]; }

    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() {
        Matrix m;
        for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1);
        return m;
    }
    void setZero() { *this = Zero(); }
    void setIdentity() { *this = Identity(); }

    T &operator()(int r, int c) { return a_[r * C + c]; }
    T operator()(int r, int c) const { return a_[r * C + c]; }
    T &operator()(int i) { return a_[i]; }
    T operator()(int i) const { return a_[i]; }
    T &operator[](int i) { return a_[i]; }
    T operator[](int i) const { return a_[i]; }
    T *data() { return a_; }              // row-major (unlike Eigen's default) -- the C ABI's layout
    const T *data() const { return a_; }
    T x() const { return a_[0]; }
    T y() const { return a_[1]; }
    T z() const { return a_[2]; }

    Matrix<T, C, R> transpose() const {
        Matrix<T, C, R> t;
        for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c);
        return t;
    }
    template <int K>
    Matrix<T, R, K> operator*(const Matrix<T, C, K> &o) const {
        Matrix<T, R, K> m;
        for (int r = 0; r < R; ++r)
            for (int k = 0; k < K; ++k) {
                T acc = T(0);
                for (int c = 0; c < C; ++c) acc += (*this)(r, c) * o(c, k);
                m(r, k) = acc;
            }
        return m;
    }
    Matrix operator+(const Matrix &o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.a_[i] = a_[i] + o.a_[i]; return m; }
    Matrix operator-(const Matrix &o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.a_[i] = a_[i] - o.a_[i]; return m; }
    Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.a_[i] = -a_[i]; return m; }
    Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.a_[i] = a_[i] * s; return m; }
    Matrix operator/(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.a_[i] = a_[i] / s; return m; }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; ++i) a_[i] += o.a_[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; ++i) a_[i] -= o.a_[i]; return *this; }
    Matrix &operator*=(T s) { for (int i = 0; i < R * C; ++i) a_[i] *= s; return *this; }
    bool operator==(const Matrix &o) const { for (int i = 0; i < R * C; ++i) if (a_[i] != o.a_[i]) return false; return true; }
    bool operator!=(const Matrix &o) const { return !(*this == o); }

    T squaredNorm() const { T s = T(0); for (int i = 0; i < R * C; ++i) s += a_[i] * a_[i]; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    T dot(const Matrix &o) const { T s = T(0); for (int i = 0; i < R * C; ++i) s += a_[i] * o.a_[i]; return s; }
    T trace() const { T s = T(0); for (int i = 0; i < (R < C ? R : C); ++i) s += (*this)(i, i); return s; }
    Matrix cross(const Matrix &o) const {
        static_assert(R * C == 3, "cross needs 3-vectors");
        return Matrix(a_[1] * o.a_[2] - a_[2] * o.a_[1], a_[2] * o.a_[0] - a_[0] * o.a_[2], a_[0] * o.a_[1] - a_[1] * o.a_[0]);
    }
    bool isIdentity(T tol = T(0)) const {
        for (int r = 0; r < R; ++r)
            for (int c = 0; c < C; ++c)
                if (std::fabs((*this)(r, c) - (r == c ? T(1) : T(0))) > tol) return false;
        return true;
    }
    template <int N> Segment<T, N> head() { return Segment<T, N>(a_); }
    template <int N> Segment<T, N> tail() { return Segment<T, N>(a_ + R * C - N); }
    template <int N> Matrix<T, N, 1> head() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v[i] = a_[i]; return v; }
    template <int N> Matrix<T, N, 1> tail() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v[i] = a_[R * C - N + i]; return v; }

private:
    T a_[R * C];
};

template <class T, int R, int C>
Matrix<T, R, C> operator*(T s, const Matrix<T, R, C> &m) { return m * s; }

template <class T, int R, int C>
std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C> &m) {
    for (int r = 0; r < R; ++r) {
        for (int c = 0; c < C; ++c) os << (c ? " " : "") << m(r, c);
        if (r + 1 < R) os << "\n";
    }
    return os;
}

// Unit quaternion, Eigen conventions: constructor (w, x, y, z), coeffs() = (x, y, z, w).
template <class T>
class Quaternion {
public:
    Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
    Quaternion(T w, T x, T y, T z) : x_(x), y_(y), z_(z), w_(w) {}
    explicit Quaternion(const Matrix<T, 3, 3> &R) {
        // the branch structure matches the device-side rot_to_quat so host-built and device-built
        // quaternions agree bit for bit
        T t = R(0, 0) + R(1, 1) + R(2, 2);
        T q[4];
        if (t > T(0)) {
            t = std::sqrt(t + T(1));
            q[3] = T(0.5) * t;
            t = T(0.5) / t;
            q[0] = (R(2, 1) - R(1, 2)) * t;
            q[1] = (R(0, 2) - R(2, 0)) * t;
            q[2] = (R(1, 0) - R(0, 1)) * t;
        } else {
            int i = 0;
            if (R(1, 1) > R(0, 0)) i = 1;
            if (R(2, 2) > R(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + T(1));
            q[i] = T(0.5) * t;
            t = T(0.5) / t;
            q[3] = (R(k, j) - R(j, k)) * t;
            q[j] = (R(j, i) + R(i, j)) * t;
            q[k] = (R(k, i) + R(i, k)) * t;
        }
        x_ = q[0]; y_ = q[1]; z_ = q[2]; w_ = q[3];
    }
    static Quaternion Identity() { return Quaternion(); }
    T x() const { return x_; }
    T y() const { return y_; }
    T z() const { return z_; }
    T w() const { return w_; }
    T &x() { return x_; }
    T &y() { return y_; }
    T &z() { return z_; }
    T &w() { return w_; }
    Matrix<T, 4, 1> coeffs() const { return Matrix<T, 4, 1>(x_, y_, z_, w_); }
    Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(x_, y_, z_); }
    T norm() const { return std::sqrt(x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_); }
    void normalize() { const T n = norm(); x_ /= n; y_ /= n; z_ /= n; w_ /= n; }
    Quaternion normalized() const { Quaternion q = *this; q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
    Quaternion inverse() const { return conjugate(); }
    Quaternion operator*(const Quaternion &b) const {
        return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_,
                          w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                          w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_,
                          w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
    }
    Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1> &v) const { return toRotationMatrix() * v; }
    Matrix<T, 3, 3> toRotationMatrix() const {
        const T tx = 2 * x_, ty = 2 * y_, tz = 2 * z_;
        const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
        const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
        const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        Matrix<T, 3, 3> R;
        R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
        R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
        R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
        return R;
    }

private:
    T x_, y_, z_, w_;
};

}  // namespace s3o

// Optional Eigen spellings for code written against the reference's call sites.  Only define this
// when the real Eigen is NOT in the translation unit.
#ifdef S3O_FACADE_EIGEN_NAMES
namespace Eigen {
template <class T, int R, int C> using Matrix = s3o::Matrix<T, R, C>;
template <class T> using Quaternion = s3o::Quaternion<T>;
using Matrix3d = s3o::Matrix<double, 3, 3>;
using Matrix4d = s3o::Matrix<double, 4, 4>;
using Vector3d = s3o::Matrix<double, 3, 1>;
using Vector4d = s3o::Matrix<double, 4, 1>;
using Vector2d = s3o::Matrix<double, 2, 1>;
using Quaterniond = s3o::Quaternion<double>;
template <class T> using aligned_allocator = std::allocator<T>;
}  // namespace Eigen
#endif
