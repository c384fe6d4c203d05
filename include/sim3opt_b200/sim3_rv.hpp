// sim3_rv.hpp -- RobotVision::Sim3 with the reference's spelling and conventions.
//
// The reference carries Strasdat's similarity-transform class (sim3_rv.h:71-226, ln at :241-320) for
// its map re-projection helpers (drawPTAMPoints.h:42,93,98; kitti_surf.cpp:1337,1375).  This is the
// same host-side value type over the facade's small matrices (TooN is not a dependency here):
//   x -> s (R x) + t,   tangent order [upsilon(0..2), omega(3..5), sigma(6)]   (sim3_rv.h:130-132,308-310)
//   exp / ln branch four ways on |sigma| < eps and theta < eps (exp) or d > 1 - eps (ln), eps = 1e-5,
//   with the coefficients A, B, C of W = A Om + B Om^2 + C I as written there -- including the
//   small-angle R = I + Om + Om^2 and B = ((sigma^2/2 - sigma + 1) s)/sigma^3 -- unless
//   Sim3<>::corrected_limits() is switched on (the consistent Taylor limits).
// The optimiser path does not run through this class: the device code (csrc/sim3_math.cuh) and the
// CPU oracle (oracle/lie.c) hold the same math in g2o's tangent order [omega, upsilon, sigma].
#pragma once
#include <cmath>
#include <ostream>
#include <utility>

#include "linalg.hpp"

namespace RobotVision {

template <typename Precision = double>
class Sim3 {
public:
    using Mat3 = s3o::Matrix<Precision, 3, 3>;
    using Vec3 = s3o::Matrix<Precision, 3, 1>;
    using Vec7 = s3o::Matrix<Precision, 7, 1>;

    Sim3() : R_(Mat3::Identity()), s_(1) {}
    Sim3(const Mat3 &R, const Vec3 &t, Precision s) : R_(R), t_(t), s_(s) {}

    Mat3 &get_rotation() { return R_; }
    const Mat3 &get_rotation() const { return R_; }
    Vec3 &get_translation() { return t_; }
    const Vec3 &get_translation() const { return t_; }
    Precision &get_scale() { return s_; }
    const Precision &get_scale() const { return s_; }

    // false (default): sim3_rv.h as written; true: consistent small-angle limits
    static bool &corrected_limits() { static bool flag = false; return flag; }

    static Sim3 exp(const Vec7 &v) {
        const Vec3 upsilon(v[0], v[1], v[2]), omega(v[3], v[4], v[5]);
        const Precision sigma = v[6], theta = omega.norm(), s = std::exp(sigma);
        const Mat3 Om = hat(omega), Om2 = Om * Om;
        const bool small = theta < eps();
        Precision A, B, C;
        coefficients(sigma, s, theta, small, A, B, C);
        Mat3 R = Mat3::Identity();
        if (small) R = R + Om + Om2 * (corrected_limits() ? Precision(0.5) : Precision(1));
        else R = R + Om * (std::sin(theta) / theta) + Om2 * ((1 - std::cos(theta)) / (theta * theta));
        const Mat3 W = Om * A + Om2 * B + Mat3::Identity() * C;
        return Sim3(R, W * upsilon, s);
    }

    static Vec7 ln(const Sim3 &S) {
        const Mat3 &R = S.R_;
        const Precision s = S.s_, sigma = std::log(s);
        const Precision d = Precision(0.5) * (R(0, 0) + R(1, 1) + R(2, 2) - 1);
        const bool small = d > 1 - eps();
        const Vec3 dr = deltaR(R);
        Vec3 omega;
        Precision theta = 0;
        if (small) omega = dr * Precision(0.5);
        else {
            theta = std::acos(d);
            omega = dr * (theta / (2 * std::sqrt(1 - d * d)));
        }
        const Mat3 Om = hat(omega), Om2 = Om * Om;
        Precision A, B, C;
        coefficients(sigma, s, theta, small, A, B, C);
        const Mat3 W = Om * A + Om2 * B + Mat3::Identity() * C;
        const Vec3 upsilon = solve(W, S.t_);
        Vec7 out;
        for (int i = 0; i < 3; ++i) { out[i] = upsilon[i]; out[3 + i] = omega[i]; }
        out[6] = sigma;
        return out;
    }
    Vec7 ln() const { return ln(*this); }

    Sim3 inverse() const {
        const Mat3 Ri = R_.transpose();
        return Sim3(Ri, (Ri * t_) * (Precision(-1) / s_), Precision(1) / s_);
    }
    Sim3 operator*(const Sim3 &o) const { return Sim3(R_ * o.R_, (R_ * o.t_) * s_ + t_, s_ * o.s_); }
    Sim3 &operator*=(const Sim3 &o) { *this = *this * o; return *this; }
    Vec3 operator*(const Vec3 &x) const { return (R_ * x) * s_ + t_; }

private:
    static Precision eps() { return Precision(0.00001); }
    static Mat3 hat(const Vec3 &w) {
        Mat3 m;
        m(0, 1) = -w[2]; m(0, 2) = w[1];
        m(1, 0) = w[2];  m(1, 2) = -w[0];
        m(2, 0) = -w[1]; m(2, 1) = w[0];
        return m;
    }
    static Vec3 deltaR(const Mat3 &R) { return Vec3(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1)); }
    static void coefficients(Precision sigma, Precision s, Precision theta, bool small, Precision &A, Precision &B,
                             Precision &C) {
        if (std::fabs(sigma) < eps()) {
            C = 1;
            if (small) { A = Precision(1) / 2; B = Precision(1) / 6; }
            else {
                const Precision t2 = theta * theta;
                A = (1 - std::cos(theta)) / t2;
                B = (theta - std::sin(theta)) / (t2 * theta);
            }
        } else {
            C = (s - 1) / sigma;
            const Precision g2 = sigma * sigma;
            if (small) {
                A = ((sigma - 1) * s + 1) / g2;
                B = ((g2 / 2 - sigma + 1) * s - (corrected_limits() ? 1 : 0)) / (g2 * sigma);
            } else {
                const Precision a = s * std::sin(theta), b = s * std::cos(theta), c = theta * theta + g2;
                A = (a * sigma + (1 - b) * theta) / (theta * c);
                B = (C - ((b - 1) * sigma + a * theta) / c) / (theta * theta);
            }
        }
    }
    static Vec3 solve(const Mat3 &M, const Vec3 &b) {     // 3x3 Gaussian elimination with partial pivoting
        Precision a[3][4];
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) a[r][c] = M(r, c); a[r][3] = b[r]; }
        for (int k = 0; k < 3; ++k) {
            int piv = k;
            for (int r = k + 1; r < 3; ++r) if (std::fabs(a[r][k]) > std::fabs(a[piv][k])) piv = r;
            for (int c = 0; c < 4; ++c) std::swap(a[k][c], a[piv][c]);
            for (int r = k + 1; r < 3; ++r) {
                const Precision f = a[r][k] / a[k][k];
                for (int c = k; c < 4; ++c) a[r][c] -= f * a[k][c];
            }
        }
        Vec3 x;
        for (int r = 2; r >= 0; --r) {
            Precision acc = a[r][3];
            for (int c = r + 1; c < 3; ++c) acc -= a[r][c] * x[c];
            x[r] = acc / a[r][r];
        }
        return x;
    }

    Mat3 R_;
    Vec3 t_;
    Precision s_;
};

// drawPTAMPoints.h:41-46
template <typename A>
inline s3o::Matrix<A, 3, 1> transform(const Sim3<A> &T, const s3o::Matrix<A, 3, 1> &x) { return T * x; }

template <typename Precision>
inline std::ostream &operator<<(std::ostream &os, const Sim3<Precision> &S) {
    for (int i = 0; i < 3; ++i)
        os << S.get_rotation()(i, 0) << " " << S.get_rotation()(i, 1) << " " << S.get_rotation()(i, 2) << " "
           << S.get_translation()[i] << std::endl;
    return os << S.get_scale() << std::endl;
}

}  // namespace RobotVision
