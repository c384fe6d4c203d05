// sim3_math.cuh -- device-side Sim3 Lie-group math (fp64) for the sm_100a kernels.
//
// Convention (reference: sim3_rv.h:125-190 exp, :241-320 ln, :199-220 inverse/compose;
// g2o::Sim3 storage and tangent order):
//   state   = unit quaternion (x,y,z,w) + translation + scale, x -> s*(R x) + t
//   tangent = [omega(3), upsilon(3), sigma]
//   eps     = 1e-5 four-way branch on |sigma| and theta / trace, small-angle R = I + Om + Om^2
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#ifndef S3O_JL_INLINE
#define S3O_JL_INLINE __forceinline__
#endif
#ifndef S3O_ERR_INLINE
#define S3O_ERR_INLINE __forceinline__
#endif

namespace s3o {

struct Sim3 {
    double qx, qy, qz, qw;
    double tx, ty, tz;
    double s;
};

#define S3O_EPS 0.00001

__device__ __forceinline__ void quat_to_rot(double x, double y, double z, double w, double R[9]) {
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ void rot_to_quat(const double R[9], double &x, double &y, double &z, double &w) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        x = (R[7] - R[5]) * t;
        y = (R[2] - R[6]) * t;
        z = (R[3] - R[1]) * t;
    } else if (R[0] >= R[4] && R[0] >= R[8]) {       // i = 0
        t = sqrt(R[0] - R[4] - R[8] + 1.0);
        x = 0.5 * t; t = 0.5 / t;
        w = (R[7] - R[5]) * t; y = (R[3] + R[1]) * t; z = (R[6] + R[2]) * t;
    } else if (R[4] > R[0] && R[4] >= R[8]) {        // i = 1
        t = sqrt(R[4] - R[8] - R[0] + 1.0);
        y = 0.5 * t; t = 0.5 / t;
        w = (R[2] - R[6]) * t; z = (R[7] + R[5]) * t; x = (R[1] + R[3]) * t;
    } else {                                          // i = 2
        t = sqrt(R[8] - R[0] - R[4] + 1.0);
        z = 0.5 * t; t = 0.5 / t;
        w = (R[3] - R[1]) * t; x = (R[2] + R[6]) * t; y = (R[5] + R[7]) * t;
    }
}

__device__ __forceinline__ void quat_rotate(double qx, double qy, double qz, double qw,
                                            double vx, double vy, double vz,
                                            double &ox, double &oy, double &oz) {
    double ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;
    ux += ux; uy += uy; uz += uz;
    ox = vx + qw * ux + (qy * uz - qz * uy);
    oy = vy + qw * uy + (qz * ux - qx * uz);
    oz = vz + qw * uz + (qx * uy - qy * ux);
}

__device__ __forceinline__ Sim3 sim3_mul(const Sim3 &a, const Sim3 &b) {
    Sim3 c;
    c.qw = a.qw * b.qw - a.qx * b.qx - a.qy * b.qy - a.qz * b.qz;
    c.qx = a.qw * b.qx + a.qx * b.qw + a.qy * b.qz - a.qz * b.qy;
    c.qy = a.qw * b.qy + a.qy * b.qw + a.qz * b.qx - a.qx * b.qz;
    c.qz = a.qw * b.qz + a.qz * b.qw + a.qx * b.qy - a.qy * b.qx;
    double rx, ry, rz;
    quat_rotate(a.qx, a.qy, a.qz, a.qw, b.tx, b.ty, b.tz, rx, ry, rz);
    c.tx = a.s * rx + a.tx;
    c.ty = a.s * ry + a.ty;
    c.tz = a.s * rz + a.tz;
    c.s = a.s * b.s;
    return c;
}

__device__ __forceinline__ Sim3 sim3_inv(const Sim3 &a) {
    Sim3 c;
    c.qx = -a.qx; c.qy = -a.qy; c.qz = -a.qz; c.qw = a.qw;
    const double k = -1. / a.s;
    quat_rotate(c.qx, c.qy, c.qz, c.qw, k * a.tx, k * a.ty, k * a.tz, c.tx, c.ty, c.tz);
    c.s = 1. / a.s;
    return c;
}

__device__ __forceinline__ void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

__device__ __forceinline__ void skew3(double x, double y, double z, double S[9]) {
    S[0] = 0;  S[1] = -z; S[2] = y;
    S[3] = z;  S[4] = 0;  S[5] = -x;
    S[6] = -y; S[7] = x;  S[8] = 0;
}

// coefficients of W = A*Om + B*Om^2 + C*I  (sim3_rv.h:143-181 / :261-303)
// corrected=false: exactly as written in the reference (sim3_rv.h:165,:291: small-angle B without
// the "-1", R = I + Om + Om^2).  corrected=true: the consistent Taylor limits (B with "-1",
// R = I + Om + Om^2/2) -- see s3o_set_math_mode.
__device__ __forceinline__ void sim3_abc(double sigma, double s, double theta, bool small_angle, bool corrected,
                                         double &A, double &B, double &C) {
    if (fabs(sigma) < S3O_EPS) {
        C = 1;
        if (small_angle) {
            A = 1. / 2.;
            B = 1. / 6.;
        } else {
            const double theta2 = theta * theta;
            double sn, cs;
            sincos(theta, &sn, &cs);
            A = (1 - cs) / theta2;
            B = (theta - sn) / (theta2 * theta);
        }
    } else {
        C = (s - 1) / sigma;
        if (small_angle) {
            const double sigma2 = sigma * sigma;
            A = ((sigma - 1) * s + 1) / sigma2;
            B = corrected ? ((0.5 * sigma2 - sigma + 1) * s - 1) / (sigma2 * sigma)
                          : ((0.5 * sigma2 - sigma + 1) * s) / (sigma2 * sigma);   // as written at sim3_rv.h:165
        } else {
            double sn, cs;
            sincos(theta, &sn, &cs);
            const double a = s * sn;
            const double b = s * cs;
            const double theta2 = theta * theta;
            const double c = theta2 + sigma * sigma;
            A = (a * sigma + (1 - b) * theta) / (theta * c);
            B = (C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
        }
    }
}

__device__ __forceinline__ Sim3 sim3_exp(const double v[7], bool corrected) {
    const double sigma = v[6];
    const double theta = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double Om[9], Om2[9], R[9];
    skew3(v[0], v[1], v[2], Om);
    mat3_mul(Om, Om, Om2);
    const double s = exp(sigma);
    const bool small_angle = theta < S3O_EPS;
    double A, B, C;
    sim3_abc(sigma, s, theta, small_angle, corrected, A, B, C);
    if (small_angle) {
        const double k2 = corrected ? 0.5 : 1.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = Om[i] + k2 * Om2[i];
    } else {
        double sn, cs;
        sincos(theta, &sn, &cs);
        const double k1 = sn / theta, k2 = (1 - cs) / (theta * theta);
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = k1 * Om[i] + k2 * Om2[i];
    }
    R[0] += 1; R[4] += 1; R[8] += 1;
    Sim3 S;
    rot_to_quat(R, S.qx, S.qy, S.qz, S.qw);
    double W[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) W[i] = A * Om[i] + B * Om2[i];
    W[0] += C; W[4] += C; W[8] += C;
    S.tx = W[0] * v[3] + W[1] * v[4] + W[2] * v[5];
    S.ty = W[3] * v[3] + W[4] * v[4] + W[5] * v[5];
    S.tz = W[6] * v[3] + W[7] * v[4] + W[8] * v[5];
    S.s = s;
    return S;
}

// 3x3 solve with partial pivoting (the reference back-substitutes an LU of W, sim3_rv.h:305-307)
__device__ __forceinline__ void solve3(const double Ain[9], double b0, double b1, double b2,
                                       double &x0, double &x1, double &x2) {
    double a[3][4] = { { Ain[0], Ain[1], Ain[2], b0 }, { Ain[3], Ain[4], Ain[5], b1 }, { Ain[6], Ain[7], Ain[8], b2 } };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int piv = k;
#pragma unroll
        for (int i = k + 1; i < 3; ++i)
            if (fabs(a[i][k]) > fabs(a[piv][k])) piv = i;
#pragma unroll
        for (int i = k + 1; i < 3; ++i)
            if (i == piv) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { double t = a[k][j]; a[k][j] = a[i][j]; a[i][j] = t; }
            }
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            const double f = a[i][k] / a[k][k];
#pragma unroll
            for (int j = k; j < 4; ++j) a[i][j] -= f * a[k][j];
        }
    }
    x2 = a[2][3] / a[2][2];
    x1 = (a[1][3] - a[1][2] * x2) / a[1][1];
    x0 = (a[0][3] - a[0][1] * x1 - a[0][2] * x2) / a[0][0];
}

__device__ __forceinline__ void sim3_log(const Sim3 &S, double v[7], bool corrected) {
    const double s = S.s;
    const double sigma = log(s);
    double R[9];
    quat_to_rot(S.qx, S.qy, S.qz, S.qw, R);
    const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
    const double d0 = R[7] - R[5], d1 = R[2] - R[6], d2 = R[3] - R[1];
    const bool small_angle = d > 1 - S3O_EPS;
    double theta = 0, k = 0.5;
    if (!small_angle) {
        theta = acos(d);
        k = theta / (2 * sqrt(1 - d * d));
    }
    const double w0 = k * d0, w1 = k * d1, w2 = k * d2;
    double A, B, C;
    sim3_abc(sigma, s, theta, small_angle, corrected, A, B, C);
    double Om[9], Om2[9], W[9];
    skew3(w0, w1, w2, Om);
    mat3_mul(Om, Om, Om2);
#pragma unroll
    for (int i = 0; i < 9; ++i) W[i] = A * Om[i] + B * Om2[i];
    W[0] += C; W[4] += C; W[8] += C;
    v[0] = w0; v[1] = w1; v[2] = w2;
    solve3(W, S.tx, S.ty, S.tz, v[3], v[4], v[5]);
    v[6] = sigma;
}

// EdgeSim3::computeError: e = log(C * Si * Sj^-1)   
__device__ __forceinline__ void sim3_edge_error(const Sim3 &C, const Sim3 &Si, const Sim3 &Sj, double e[7],
                                                bool corrected) {
    sim3_log(sim3_mul(sim3_mul(C, Si), sim3_inv(Sj)), e, corrected);
}

__device__ __forceinline__ bool inv3(const double A[9], double I[9]) {
    const double c0 = A[4] * A[8] - A[5] * A[7];
    const double c1 = A[5] * A[6] - A[3] * A[8];
    const double c2 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
    const double id = 1.0 / det;
    I[0] = c0 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = c1 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = c2 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return det != 0.0;
}

// Inverse left Jacobian of Sim3 in block form.  With ad_e = [[Om,0,0],[Up,M,-ups],[0,0,0]],
// M = Om + sigma*I, the series Jl(e) = sum_n ad_e^n/(n+1)! has the block structure
//   Jl = [[Jw,0,0],[Q,W,w],[0,0,1]],  ad^n = [[Om^n,0,0],[P_n,M^n,-M^(n-1) ups],[0,0,0]],
//   P_n = M P_(n-1) + Up Om^(n-1).
// Jl^-1 = [[Jw^-1,0,0],[X,W^-1,y],[0,0,1]],  X = -W^-1 Q Jw^-1,  y = -W^-1 w.
struct JlInv {
    double Jw[9];  // Jw^-1
    double X[9];
    double Wi[9];  // W^-1
    double y[3];
};

// 1/(n+1)!, n = 0..79: the series coefficients (a table instead of one fp64 division per term; constant bank: the
// loop counter is the index, so the access is uniform over the warp)
static __constant__ double kInvFact1[80] = {
    1.0, 0.5, 0.16666666666666666, 0.041666666666666664, 0.008333333333333333, 0.001388888888888889,
    0.0001984126984126984, 2.48015873015873e-05, 2.7557319223985893e-06, 2.755731922398589e-07,
    2.505210838544172e-08, 2.08767569878681e-09, 1.6059043836821613e-10, 1.1470745597729725e-11,
    7.647163731819816e-13, 4.779477332387385e-14, 2.8114572543455206e-15, 1.5619206968586225e-16,
    8.22063524662433e-18, 4.110317623312165e-19, 1.9572941063391263e-20, 8.896791392450574e-22,
    3.868170170630684e-23, 1.6117375710961184e-24, 6.446950284384474e-26, 2.4795962632247976e-27,
    9.183689863795546e-29, 3.279889237069838e-30, 1.1309962886447716e-31, 3.7699876288159054e-33,
    1.216125041553518e-34, 3.8003907548547434e-36, 1.151633562077195e-37, 3.387157535521162e-39,
    9.67759295863189e-41, 2.6882202662866363e-42, 7.265460179153071e-44, 1.911963205040282e-45,
    4.902469756513544e-47, 1.2256174391283858e-48, 2.9893108271424046e-50, 7.117406731291439e-52,
    1.6552108677421951e-53, 3.7618428812322616e-55, 8.359650847182804e-57, 1.817315401561479e-58,
    3.866628513960594e-60, 8.055476070751236e-62, 1.643974708316579e-63, 3.287949416633158e-65,
    6.446959640457172e-67, 1.2397999308571486e-68, 2.3392451525606576e-70, 4.331935467704922e-72,
    7.876246304918039e-74, 1.4064725544496498e-75, 2.4674957095607893e-77, 4.254302947518602e-79,
    7.2106829618959365e-81, 1.2017804936493226e-82, 1.9701319568021682e-84, 3.1776321883905942e-86,
    5.043860616493007e-88, 7.881032213270323e-90, 1.2124664943492804e-91, 1.8370704459837581e-93,
    2.74189618803546e-95, 4.0322002765227353e-97, 5.843768516699616e-99, 8.34824073814231e-101,
    1.1758085546679308e-102, 1.633067437038793e-104, 2.2370786808750587e-106, 3.023079298479809e-108,
    4.030772397973079e-110, 5.30364789206984e-112, 6.887854405285506e-114, 8.830582570878855e-116,
    1.117795262136564e-117, 1.397244077670705e-119 };

// out = w x (columns of A)  (= Om A, Om = skew(w))
__device__ __forceinline__ void cross_cols(double x, double y, double z, const double A[9], double out[9]) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        out[j] = y * A[6 + j] - z * A[3 + j];
        out[3 + j] = z * A[j] - x * A[6 + j];
        out[6 + j] = x * A[3 + j] - y * A[j];
    }
}
// out = (rows of A) x w  (= A Om)
__device__ __forceinline__ void cross_rows(const double A[9], double x, double y, double z, double out[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        out[i * 3] = A[i * 3 + 1] * z - A[i * 3 + 2] * y;
        out[i * 3 + 1] = A[i * 3 + 2] * x - A[i * 3] * z;
        out[i * 3 + 2] = A[i * 3] * y - A[i * 3 + 1] * x;
    }
}

// The series runs on scalars: Om^3 = -theta^2 Om, so every block is a polynomial in Om --
//   Om^n, M^n in span{I, Om, Om^2},   P_n = sum_ab p_ab(n) Om^a Up Om^b,  a, b in {0,1,2}
// (left multiplication by M acts on the index a: (x0,x1,x2) -> (s x0, s x1 + x0 - th2 x2, s x2 + x1)).  One term
// costs ~45 fp64 operations instead of three 3x3 products; the nine products Om^a Up Om^b are formed once.
static __device__ S3O_JL_INLINE void sim3_jl_inv(const double e[7], JlInv &out) {
    const double sg = e[6];
    // Number of terms from a bound on the series tail instead of a per-term maximum over 30 entries:
    // |ad_e^n| <= a^n with a = |omega| + |upsilon| + |sigma| (row sums of the blocks), term n carries 1/(n+1)!.
    const double theta2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    const double a = sqrt(theta2) + sqrt(e[3] * e[3] + e[4] * e[4] + e[5] * e[5]) + fabs(sg);
    int nterms = 2;
    {   // (fp32 is plenty for a stopping bound; 1e-18 keeps the fp64 series exact to the last bit)
        const float af = (float)a, lim = 1e-18f / (1.0f + af);
        float bound = 0.5f * af;         // a^(nterms-1) / nterms!
        while (bound >= lim && nterms < 80) { bound *= __fdividef(af, (float)(nterms + 1)); ++nterms; }
    }
    double o0 = 1, o1 = 0, o2 = 0;       // Om^(n-1)
    double m0 = 1, m1 = 0, m2 = 0;       // M^(n-1)
    double p[3][3], q[3][3];             // P_(n-1), Q
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { p[i][j] = 0; q[i][j] = 0; }
    double j1 = 0, j2 = 0;               // Jw = I + j1 Om + j2 Om^2
    double w0 = 1, w1 = 0, w2 = 0;       // W
    double u0 = 0, u1 = 0, u2 = 0;       // sum_n c_n M^(n-1):  w = -(that) ups
    for (int n = 1; n < nterms; ++n) {
        const double c = kInvFact1[n];
        const double ob[3] = { o0, o1, o2 };
#pragma unroll
        for (int b = 0; b < 3; ++b) {    // P_n = M P_(n-1) + Up Om^(n-1)
            const double t0 = sg * p[0][b] + ob[b];
            const double t1 = sg * p[1][b] + (p[0][b] - theta2 * p[2][b]);
            const double t2 = sg * p[2][b] + p[1][b];
            p[0][b] = t0; p[1][b] = t1; p[2][b] = t2;
            q[0][b] += c * t0; q[1][b] += c * t1; q[2][b] += c * t2;
        }
        u0 += c * m0; u1 += c * m1; u2 += c * m2;
        {   // M^n
            const double t0 = sg * m0, t1 = sg * m1 + (m0 - theta2 * m2), t2 = sg * m2 + m1;
            m0 = t0; m1 = t1; m2 = t2;
        }
        {   // Om^n
            const double t1 = o0 - theta2 * o2, t2 = o1;
            o0 = 0; o1 = t1; o2 = t2;
        }
        j1 += c * o1; j2 += c * o2;
        w0 += c * m0; w1 += c * m1; w2 += c * m2;
    }
    const double wx = e[0], wy = e[1], wz = e[2];
    double Om[9], Om2[9], Up[9];
    skew3(wx, wy, wz, Om);
    skew3(e[3], e[4], e[5], Up);
    {   // Om^2 = w w^T - theta^2 I
        Om2[0] = wx * wx - theta2; Om2[1] = wx * wy;          Om2[2] = wx * wz;
        Om2[3] = Om2[1];           Om2[4] = wy * wy - theta2; Om2[5] = wy * wz;
        Om2[6] = Om2[2];           Om2[7] = Om2[5];           Om2[8] = wz * wz - theta2;
    }
    double Jw[9], W[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        Jw[i] = j1 * Om[i] + j2 * Om2[i];
        W[i] = w1 * Om[i] + w2 * Om2[i];
    }
    Jw[0] += 1; Jw[4] += 1; Jw[8] += 1;
    W[0] += w0; W[4] += w0; W[8] += w0;
    // Q = G0 + G1 Om + G2 Om^2,  G_b = sum_a q_ab Om^a Up
    double U1[9], U2[9], G[9], T[9], T2[9], Q[9];
    cross_cols(wx, wy, wz, Up, U1);
    cross_cols(wx, wy, wz, U1, U2);
#pragma unroll
    for (int i = 0; i < 9; ++i) Q[i] = q[0][0] * Up[i] + q[1][0] * U1[i] + q[2][0] * U2[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) G[i] = q[0][1] * Up[i] + q[1][1] * U1[i] + q[2][1] * U2[i];
    cross_rows(G, wx, wy, wz, T);
#pragma unroll
    for (int i = 0; i < 9; ++i) { Q[i] += T[i]; G[i] = q[0][2] * Up[i] + q[1][2] * U1[i] + q[2][2] * U2[i]; }
    cross_rows(G, wx, wy, wz, T);
    cross_rows(T, wx, wy, wz, T2);
#pragma unroll
    for (int i = 0; i < 9; ++i) Q[i] += T2[i];
    // w = -(u0 I + u1 Om + u2 Om^2) ups
    double w[3];
    {
        const double ux = e[3], uy = e[4], uz = e[5];
        const double cx = wy * uz - wz * uy, cy = wz * ux - wx * uz, cz = wx * uy - wy * ux;     // Om ups
        const double dx = wy * cz - wz * cy, dy = wz * cx - wx * cz, dz = wx * cy - wy * cx;     // Om^2 ups
        w[0] = -(u0 * ux + u1 * cx + u2 * dx);
        w[1] = -(u0 * uy + u1 * cy + u2 * dy);
        w[2] = -(u0 * uz + u1 * cz + u2 * dz);
    }
    inv3(Jw, out.Jw);
    inv3(W, out.Wi);
    mat3_mul(out.Wi, Q, T);
    mat3_mul(T, out.Jw, out.X);
#pragma unroll
    for (int i = 0; i < 9; ++i) out.X[i] = -out.X[i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
        out.y[i] = -(out.Wi[i * 3] * w[0] + out.Wi[i * 3 + 1] * w[1] + out.Wi[i * 3 + 2] * w[2]);
}

// Analytic Jacobians of e = log(C Si Sj^-1) w.r.t. left perturbations S <- exp(delta) S:
//   Ji = Jl^-1(e) * Ad_C,   Jj = -Jl^-1(-e)           (below row a18)
// J is row-major 7x7.
__device__ __forceinline__ void sim3_edge_jacobians(const Sim3 &C, const double e[7], double Ji[49], double Jj[49]) {
    JlInv L;
    sim3_jl_inv(e, L);
    double R[9], Tx[9], TR[9];
    quat_to_rot(C.qx, C.qy, C.qz, C.qw, R);
    skew3(C.tx, C.ty, C.tz, Tx);
    mat3_mul(Tx, R, TR);
    double A11[9], A21[9], T1[9], T2[9], A22[9];
    mat3_mul(L.Jw, R, A11);
    mat3_mul(L.X, R, T1);
    mat3_mul(L.Wi, TR, T2);
#pragma unroll
    for (int i = 0; i < 9; ++i) A21[i] = T1[i] + T2[i];
    mat3_mul(L.Wi, R, A22);
#pragma unroll
    for (int i = 0; i < 49; ++i) Ji[i] = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Ji[r * 7 + c] = A11[r * 3 + c];
            Ji[(3 + r) * 7 + c] = A21[r * 3 + c];
            Ji[(3 + r) * 7 + 3 + c] = C.s * A22[r * 3 + c];
        }
        Ji[(3 + r) * 7 + 6] = L.y[r] - (L.Wi[r * 3] * C.tx + L.Wi[r * 3 + 1] * C.ty + L.Wi[r * 3 + 2] * C.tz);
    }
    Ji[48] = 1;
    // Jj = -Jl^-1(-e).  The odd Bernoulli numbers beyond B_1 vanish, so Jl^-1(-e) = Jl^-1(e) + ad_e exactly:
    // the second series is not needed.  ad_e = [[Om,0,0],[Up,Om+sigma I,-ups],[0,0,0]].
    double Om[9], Up[9];
    skew3(e[0], e[1], e[2], Om);
    skew3(e[3], e[4], e[5], Up);
#pragma unroll
    for (int i = 0; i < 49; ++i) Jj[i] = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Jj[r * 7 + c] = -(L.Jw[r * 3 + c] + Om[r * 3 + c]);
            Jj[(3 + r) * 7 + c] = -(L.X[r * 3 + c] + Up[r * 3 + c]);
            Jj[(3 + r) * 7 + 3 + c] = -(L.Wi[r * 3 + c] + Om[r * 3 + c] + (r == c ? e[6] : 0.0));
        }
        Jj[(3 + r) * 7 + 6] = -(L.y[r] - e[3 + r]);
    }
    Jj[48] = -1;
}

}  // namespace s3o
