/*
 * lie.c -- Sim3 / SE3 Lie-group math of the oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Follows the written convention of the reference:
 *   exp      sim3_rv.h:125-190   (four-way branch on |sigma|<eps x theta<eps, eps=1e-5,
 *                                 small-angle R = I + Omega + Omega^2 "sic")
 *   ln       sim3_rv.h:241-320   (branch on |sigma|<eps x d>1-eps; upsilon = W^-1 t by 3x3 LU)
 *   inverse  sim3_rv.h:199-203
 *   compose  sim3_rv.h:214-220
 * with g2o::Sim3's storage (unit quaternion + t + s) and tangent order
 * [omega, upsilon, sigma].  Quaternion <-> matrix
 * conversions follow Eigen's published formulas, which g2o::Sim3 relies on.
 */
#include "oracle.h"
#include <math.h>
#include <string.h>

static const double SIM3_EPS = 0.00001; /* sim3_rv.h:133, :258 */

/* ORC_MATH_REFERENCE (default): coefficients exactly as written in the reference, including the
 * small-angle / non-zero-sigma B of sim3_rv.h:165,:291 (no "-1") and R = I + Om + Om^2.
 * ORC_MATH_CORRECTED: B = ((sigma^2/2 - sigma + 1) s - 1)/sigma^3 and R = I + Om + Om^2/2, i.e. the
 * consistent Taylor limits (what later g2o releases ship); needed wherever LM must converge on
 * graphs whose residuals enter the small-angle branch with sigma != 0. */
static int g_math_mode = ORC_MATH_REFERENCE;
void orc_set_math_mode(int mode) { g_math_mode = mode; }
int orc_get_math_mode(void) { return g_math_mode; }

void orc_quat_to_rot(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

void orc_rot_to_quat(const double R[9], double q[4]) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
        q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
        q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
    }
}

static void quat_mul(const double a[4], const double b[4], double c[4]) {
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    c[3] = aw * bw - ax * bx - ay * by - az * bz;
    c[0] = aw * bx + ax * bw + ay * bz - az * by;
    c[1] = aw * by + ay * bw + az * bx - ax * bz;
    c[2] = aw * bz + az * bw + ax * by - ay * bx;
}

static void quat_rotate(const double q[4], const double v[3], double out[3]) {
    /* v + w*(2 q_v x v) + q_v x (2 q_v x v) */
    double uv[3] = { q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0] };
    uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
    const double c0 = q[1] * uv[2] - q[2] * uv[1];
    const double c1 = q[2] * uv[0] - q[0] * uv[2];
    const double c2 = q[0] * uv[1] - q[1] * uv[0];
    const double r0 = v[0] + q[3] * uv[0] + c0;
    const double r1 = v[1] + q[3] * uv[1] + c1;
    const double r2 = v[2] + q[3] * uv[2] + c2;
    out[0] = r0; out[1] = r1; out[2] = r2;
}

static void skew3(const double w[3], double S[9]) {
    S[0] = 0;     S[1] = -w[2]; S[2] = w[1];
    S[3] = w[2];  S[4] = 0;     S[5] = -w[0];
    S[6] = -w[1]; S[7] = w[0];  S[8] = 0;
}

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    memcpy(C, T, sizeof T);
}

/* coefficients A,B,C of W = A*Omega + B*Omega^2 + C*I  (sim3_rv.h:143-181 / :261-303) */
static void sim3_abc(double sigma, double s, double theta, int small_angle, double *A, double *B, double *C) {
    if (fabs(sigma) < SIM3_EPS) {
        *C = 1;
        if (small_angle) {
            *A = 1. / 2.;
            *B = 1. / 6.;
        } else {
            const double theta2 = theta * theta;
            *A = (1 - cos(theta)) / theta2;
            *B = (theta - sin(theta)) / (theta2 * theta);
        }
    } else {
        *C = (s - 1) / sigma;
        if (small_angle) {
            const double sigma2 = sigma * sigma;
            *A = ((sigma - 1) * s + 1) / sigma2;
            if (g_math_mode == ORC_MATH_CORRECTED)
                *B = ((0.5 * sigma2 - sigma + 1) * s - 1) / (sigma2 * sigma);
            else
                *B = ((0.5 * sigma2 - sigma + 1) * s) / (sigma2 * sigma); /* as written at sim3_rv.h:165 */
        } else {
            const double a = s * sin(theta);
            const double b = s * cos(theta);
            const double theta2 = theta * theta;
            const double c = theta2 + sigma * sigma;
            *A = (a * sigma + (1 - b) * theta) / (theta * c);
            *B = (*C - ((b - 1) * sigma + a * theta) / c) * 1. / theta2;
        }
    }
}

void orc_sim3_exp(const double v[7], double S[8]) {
    const double omega[3] = { v[0], v[1], v[2] };
    const double upsilon[3] = { v[3], v[4], v[5] };
    const double sigma = v[6];
    const double theta = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
    double Omega[9], Omega2[9], R[9];
    skew3(omega, Omega);
    mat3_mul(Omega, Omega, Omega2);
    const double s = exp(sigma);
    double A, B, C;
    const int small_angle = theta < SIM3_EPS;
    sim3_abc(sigma, s, theta, small_angle, &A, &B, &C);
    if (small_angle) {
        const double k2 = g_math_mode == ORC_MATH_CORRECTED ? 0.5 : 1.0;
        for (int i = 0; i < 9; ++i) R[i] = Omega[i] + k2 * Omega2[i];
        R[0] += 1; R[4] += 1; R[8] += 1;
    } else {
        const double k1 = sin(theta) / theta, k2 = (1 - cos(theta)) / (theta * theta);
        for (int i = 0; i < 9; ++i) R[i] = k1 * Omega[i] + k2 * Omega2[i];
        R[0] += 1; R[4] += 1; R[8] += 1;
    }
    orc_rot_to_quat(R, S);
    double W[9];
    for (int i = 0; i < 9; ++i) W[i] = A * Omega[i] + B * Omega2[i];
    W[0] += C; W[4] += C; W[8] += C;
    for (int i = 0; i < 3; ++i)
        S[4 + i] = W[i * 3] * upsilon[0] + W[i * 3 + 1] * upsilon[1] + W[i * 3 + 2] * upsilon[2];
    S[7] = s;
}

/* 3x3 solve with partial pivoting (the reference uses an LU back-substitution, sim3_rv.h:305-307) */
static void solve3(const double Ain[9], const double b[3], double x[3]) {
    double A[9], y[3] = { b[0], b[1], b[2] };
    memcpy(A, Ain, sizeof A);
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        for (int i = k + 1; i < 3; ++i)
            if (fabs(A[i * 3 + k]) > fabs(A[piv * 3 + k])) piv = i;
        if (piv != k) {
            for (int j = 0; j < 3; ++j) { double t = A[k * 3 + j]; A[k * 3 + j] = A[piv * 3 + j]; A[piv * 3 + j] = t; }
            double t = y[k]; y[k] = y[piv]; y[piv] = t;
        }
        for (int i = k + 1; i < 3; ++i) {
            const double f = A[i * 3 + k] / A[k * 3 + k];
            for (int j = k; j < 3; ++j) A[i * 3 + j] -= f * A[k * 3 + j];
            y[i] -= f * y[k];
        }
    }
    for (int i = 2; i >= 0; --i) {
        double acc = y[i];
        for (int j = i + 1; j < 3; ++j) acc -= A[i * 3 + j] * x[j];
        x[i] = acc / A[i * 3 + i];
    }
}

void orc_sim3_log(const double S[8], double v[7]) {
    const double s = S[7];
    const double sigma = log(s);
    double R[9];
    orc_quat_to_rot(S, R);
    const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
    const double dR[3] = { R[7] - R[5], R[2] - R[6], R[3] - R[1] };
    double omega[3], Omega[9], Omega2[9];
    const int small_angle = d > 1 - SIM3_EPS;
    double theta = 0;
    if (small_angle) {
        for (int i = 0; i < 3; ++i) omega[i] = 0.5 * dR[i];
    } else {
        theta = acos(d);
        const double k = theta / (2 * sqrt(1 - d * d));
        for (int i = 0; i < 3; ++i) omega[i] = k * dR[i];
    }
    double A, B, C;
    sim3_abc(sigma, s, theta, small_angle, &A, &B, &C);
    skew3(omega, Omega);
    mat3_mul(Omega, Omega, Omega2);
    double W[9];
    for (int i = 0; i < 9; ++i) W[i] = A * Omega[i] + B * Omega2[i];
    W[0] += C; W[4] += C; W[8] += C;
    double upsilon[3];
    solve3(W, S + 4, upsilon);
    v[0] = omega[0]; v[1] = omega[1]; v[2] = omega[2];
    v[3] = upsilon[0]; v[4] = upsilon[1]; v[5] = upsilon[2];
    v[6] = sigma;
}

void orc_sim3_mul(const double A[8], const double B[8], double C[8]) {
    double q[4], rt[3];
    quat_mul(A, B, q);
    quat_rotate(A, B + 4, rt);
    const double s = A[7];
    C[4] = s * rt[0] + A[4];
    C[5] = s * rt[1] + A[5];
    C[6] = s * rt[2] + A[6];
    C[7] = A[7] * B[7];
    C[0] = q[0]; C[1] = q[1]; C[2] = q[2]; C[3] = q[3];
}

void orc_sim3_inv(const double A[8], double C[8]) {
    const double qc[4] = { -A[0], -A[1], -A[2], A[3] };
    const double k = -1. / A[7];
    const double tt[3] = { k * A[4], k * A[5], k * A[6] };
    double rt[3];
    quat_rotate(qc, tt, rt);
    C[0] = qc[0]; C[1] = qc[1]; C[2] = qc[2]; C[3] = qc[3];
    C[4] = rt[0]; C[5] = rt[1]; C[6] = rt[2];
    C[7] = 1. / A[7];
}

/* Ad_S = [[R,0,0],[[t]x R, sR, -t],[0,0,1]]  (tangent [omega,upsilon,sigma]) */
void orc_sim3_adjoint(const double S[8], double Ad[49]) {
    double R[9], T[9], TR[9];
    orc_quat_to_rot(S, R);
    skew3(S + 4, T);
    mat3_mul(T, R, TR);
    memset(Ad, 0, 49 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            Ad[i * 7 + j] = R[i * 3 + j];
            Ad[(3 + i) * 7 + j] = TR[i * 3 + j];
            Ad[(3 + i) * 7 + 3 + j] = S[7] * R[i * 3 + j];
        }
    for (int i = 0; i < 3; ++i) Ad[(3 + i) * 7 + 6] = -S[4 + i];
    Ad[48] = 1;
}

/* ad_e = [[Om,0,0],[Up, Om + sigma I, -upsilon],[0,0,0]] */
void orc_sim3_ad(const double e[7], double ad[49]) {
    double Om[9], Up[9];
    skew3(e, Om);
    skew3(e + 3, Up);
    memset(ad, 0, 49 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            ad[i * 7 + j] = Om[i * 3 + j];
            ad[(3 + i) * 7 + j] = Up[i * 3 + j];
            ad[(3 + i) * 7 + 3 + j] = Om[i * 3 + j] + (i == j ? e[6] : 0.0);
        }
    for (int i = 0; i < 3; ++i) ad[(3 + i) * 7 + 6] = -e[3 + i];
}

static int inv7(const double Ain[49], double Inv[49]) {
    double A[49];
    memcpy(A, Ain, sizeof A);
    for (int i = 0; i < 49; ++i) Inv[i] = 0;
    for (int i = 0; i < 7; ++i) Inv[i * 8] = 1;
    for (int k = 0; k < 7; ++k) {
        int piv = k;
        for (int i = k + 1; i < 7; ++i)
            if (fabs(A[i * 7 + k]) > fabs(A[piv * 7 + k])) piv = i;
        if (A[piv * 7 + k] == 0) return -1;
        if (piv != k)
            for (int j = 0; j < 7; ++j) {
                double t = A[k * 7 + j]; A[k * 7 + j] = A[piv * 7 + j]; A[piv * 7 + j] = t;
                t = Inv[k * 7 + j]; Inv[k * 7 + j] = Inv[piv * 7 + j]; Inv[piv * 7 + j] = t;
            }
        const double ip = 1.0 / A[k * 7 + k];
        for (int j = 0; j < 7; ++j) { A[k * 7 + j] *= ip; Inv[k * 7 + j] *= ip; }
        for (int i = 0; i < 7; ++i) {
            if (i == k) continue;
            const double f = A[i * 7 + k];
            if (f == 0) continue;
            for (int j = 0; j < 7; ++j) { A[i * 7 + j] -= f * A[k * 7 + j]; Inv[i * 7 + j] -= f * Inv[k * 7 + j]; }
        }
    }
    return 0;
}

/* Jl(e) = sum_n ad_e^n/(n+1)!  (entire series, SURVEY.md "Hard parts"), then 7x7 inverse */
void orc_sim3_jl_inv(const double e[7], double Jinv[49]) {
    double ad[49], term[49], J[49], tmp[49];
    orc_sim3_ad(e, ad);
    memset(J, 0, sizeof J);
    memset(term, 0, sizeof term);
    for (int i = 0; i < 7; ++i) { J[i * 8] = 1; term[i * 8] = 1; }
    for (int n = 1; n < 80; ++n) {
        /* term <- term * ad / (n+1) */
        double mx = 0;
        for (int i = 0; i < 7; ++i)
            for (int j = 0; j < 7; ++j) {
                double acc = 0;
                for (int k = 0; k < 7; ++k) acc += term[i * 7 + k] * ad[k * 7 + j];
                acc /= (double)(n + 1);
                tmp[i * 7 + j] = acc;
                if (fabs(acc) > mx) mx = fabs(acc);
            }
        memcpy(term, tmp, sizeof term);
        for (int i = 0; i < 49; ++i) J[i] += term[i];
        if (mx < 1e-18) break;
    }
    inv7(J, Jinv);
}

void orc_roteu2ro(const double eul[3], double R[9]) { /* kittiDetector.h:225-243 */
    const double cr = cos(eul[0]), sr = sin(eul[0]);
    const double cp = cos(eul[1]), sp = sin(eul[1]);
    const double ch = cos(eul[2]), sh = sin(eul[2]);
    R[0] = cp * ch; R[1] = (sp * sr * ch) - (cr * sh); R[2] = (cr * sp * ch) + (sh * sr);
    R[3] = cp * sh; R[4] = (sr * sp * sh) + (cr * ch); R[5] = (cr * sp * sh) - (sr * ch);
    R[6] = -sp;     R[7] = sr * cp;                    R[8] = cr * cp;
}

/* ---- SE3Quat (g2o/types/slam3d/se3quat.h) ------------- */
void orc_se3_exp(const double v[6], double T[7]) {
    const double omega[3] = { v[0], v[1], v[2] };
    const double upsilon[3] = { v[3], v[4], v[5] };
    const double theta = sqrt(omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2]);
    double Omega[9], Omega2[9], R[9], V[9];
    skew3(omega, Omega);
    mat3_mul(Omega, Omega, Omega2);
    if (theta < 0.00001) {
        for (int i = 0; i < 9; ++i) R[i] = Omega[i] + Omega2[i];
        R[0] += 1; R[4] += 1; R[8] += 1;
        memcpy(V, R, sizeof V);
    } else {
        const double k1 = sin(theta) / theta, k2 = (1 - cos(theta)) / (theta * theta);
        const double k3 = (theta - sin(theta)) / (theta * theta * theta);
        for (int i = 0; i < 9; ++i) { R[i] = k1 * Omega[i] + k2 * Omega2[i]; V[i] = k2 * Omega[i] + k3 * Omega2[i]; }
        R[0] += 1; R[4] += 1; R[8] += 1;
        V[0] += 1; V[4] += 1; V[8] += 1;
    }
    orc_rot_to_quat(R, T);
    for (int i = 0; i < 3; ++i)
        T[4 + i] = V[i * 3] * upsilon[0] + V[i * 3 + 1] * upsilon[1] + V[i * 3 + 2] * upsilon[2];
}

void orc_se3_mul(const double A[7], const double B[7], double C[7]) {
    /* SE3Quat::operator*: t = t1 + r1*t2; r = r1*r2; then normalizeRotation() */
    double q[4], rt[3];
    quat_mul(A, B, q);
    quat_rotate(A, B + 4, rt);
    C[4] = A[4] + rt[0]; C[5] = A[5] + rt[1]; C[6] = A[6] + rt[2];
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    C[0] = q[0] / n; C[1] = q[1] / n; C[2] = q[2] / n; C[3] = q[3] / n;
}

/* ---- per-edge ------------------------------------------------------------ */
void orc_sim3_edge_error(const double C[8], const double Si[8], const double Sj[8], double e[7]) {
    double Sjinv[8], T1[8], E[8];
    orc_sim3_inv(Sj, Sjinv);
    orc_sim3_mul(C, Si, T1);
    orc_sim3_mul(T1, Sjinv, E);
    orc_sim3_log(E, e);
}

void orc_sim3_edge_jac_numeric(const double C[8], const double Si[8], const double Sj[8],
                               double h, double Ji[49], double Jj[49]) {
    /* g2o BaseBinaryEdge::linearizeOplus: perturb each vertex through oplus, +h then -h */
    const double scalar = 1.0 / (2 * h);
    for (int side = 0; side < 2; ++side) {
        double *J = side == 0 ? Ji : Jj;
        const double *S = side == 0 ? Si : Sj;
        for (int d = 0; d < 7; ++d) {
            double add[7] = { 0, 0, 0, 0, 0, 0, 0 }, U[8], Sp[8], e1[7], e2[7];
            add[d] = h;
            orc_sim3_exp(add, U);
            orc_sim3_mul(U, S, Sp);
            if (side == 0) orc_sim3_edge_error(C, Sp, Sj, e1); else orc_sim3_edge_error(C, Si, Sp, e1);
            add[d] = -h;
            orc_sim3_exp(add, U);
            orc_sim3_mul(U, S, Sp);
            if (side == 0) orc_sim3_edge_error(C, Sp, Sj, e2); else orc_sim3_edge_error(C, Si, Sp, e2);
            for (int r = 0; r < 7; ++r) J[r * 7 + d] = scalar * (e1[r] - e2[r]);
        }
    }
}

void orc_sim3_edge_jac_analytic(const double C[8], const double Si[8], const double Sj[8],
                                double Ji[49], double Jj[49]) {
    double e[7], me[7], Jl[49], Ad[49], Jr[49];
    orc_sim3_edge_error(C, Si, Sj, e);
    for (int i = 0; i < 7; ++i) me[i] = -e[i];
    orc_sim3_jl_inv(e, Jl);
    orc_sim3_adjoint(C, Ad);
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            double acc = 0;
            for (int k = 0; k < 7; ++k) acc += Jl[i * 7 + k] * Ad[k * 7 + j];
            Ji[i * 7 + j] = acc;
        }
    orc_sim3_jl_inv(me, Jr);
    for (int i = 0; i < 49; ++i) Jj[i] = -Jr[i];
}

/* ---- robust kernels ------------------------------------------------------ */
void orc_robustify(int kind, double param, double e2, double rho[3]) {
    switch (kind) {
    case ORC_ROBUST_HUBER: { /* g2o RobustKernelHuber::robustify */
        const double dsqr = param * param;
        if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1; rho[2] = 0; }
        else {
            const double sqrte = sqrt(e2);
            rho[0] = 2 * sqrte * param - dsqr;
            rho[1] = param / sqrte;
            rho[2] = -0.5 * rho[1] / e2;
        }
        break;
    }
    case ORC_ROBUST_PTAM_TUKEY: { /* MEstimator.h:54-76, param = sigma^2 */
        if (e2 > param) { rho[0] = 1.0; rho[1] = 0.0; }
        else {
            const double d = 1.0 - e2 / param;
            rho[0] = 1.0 - d * d * d;
            rho[1] = d * d;
        }
        rho[2] = 0;
        break;
    }
    case ORC_ROBUST_PTAM_CAUCHY: /* MEstimator.h:97-110 */
        rho[0] = log(1.0 + e2 / param);
        rho[1] = 1.0 / (1.0 + e2 / param);
        rho[2] = 0;
        break;
    case ORC_ROBUST_PTAM_HUBER: /* MEstimator.h:131-154 */
        if (e2 < param) { rho[0] = 0.5 * e2; rho[1] = 1; }
        else {
            const double ds = sqrt(param), de = sqrt(e2);
            rho[0] = ds * (de - 0.5 * ds);
            rho[1] = sqrt(param / e2);
        }
        rho[2] = 0;
        break;
    case ORC_ROBUST_PTAM_LS: /* MEstimator.h:174-187 */
    case ORC_ROBUST_NONE:
    default:
        rho[0] = e2; rho[1] = 1; rho[2] = 0;
        break;
    }
}

static void sort_doubles(double *a, int n) {
    /* heap sort: deterministic, no libc qsort comparator overhead */
    for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) ++child;
            if (a[root] < a[child]) { double t = a[root]; a[root] = a[child]; a[child] = t; root = child; } else break;
        }
    }
    for (int end = n - 1; end > 0; --end) {
        double t = a[0]; a[0] = a[end]; a[end] = t;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) ++child;
            if (a[root] < a[child]) { double u = a[root]; a[root] = a[child]; a[child] = u; root = child; } else break;
        }
    }
}

double orc_ptam_find_sigma_squared(int kind, double *err_sq, int n) {
    /* MEstimator.h:79-89, :113-123, :157-167, :190-198 -- sorts err_sq in place */
    if (kind == ORC_ROBUST_PTAM_LS) {
        if (n == 0) return 0.0;
        double sum = 0;
        for (int i = 0; i < n; ++i) sum += err_sq[i];
        return sum / n;
    }
    sort_doubles(err_sq, n);
    const double med = err_sq[n / 2];
    double sigma = 1.4826 * (1 + 5.0 / (n * 2 - 6)) * sqrt(med);
    sigma = (kind == ORC_ROBUST_PTAM_HUBER ? 1.345 : 4.6851) * sigma;
    return sigma * sigma;
}
