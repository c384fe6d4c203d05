// Host build of the DEVICE math header (sim3opt_b200/csrc/sim3_math.cuh) for the CPU test suite: the same source the
// kernels compile, through g++ (the CUDA qualifiers expand to nothing outside nvcc).  tests/test_device_math_host.py
// compares it with the oracle's restatement -- the two were written independently of each other.
#include <cmath>
static inline float __fdividef(float a, float b) { return a / b; }
#include "../../sim3opt_b200/csrc/sim3_math.cuh"

using namespace s3o;

static Sim3 load(const double *x) {
    Sim3 S;
    S.qx = x[0]; S.qy = x[1]; S.qz = x[2]; S.qw = x[3];
    S.tx = x[4]; S.ty = x[5]; S.tz = x[6]; S.s = x[7];
    return S;
}
static void store(const Sim3 &S, double *x) {
    x[0] = S.qx; x[1] = S.qy; x[2] = S.qz; x[3] = S.qw;
    x[4] = S.tx; x[5] = S.ty; x[6] = S.tz; x[7] = S.s;
}

extern "C" {
void dm_exp(const double *v, int corrected, double *S) { store(sim3_exp(v, corrected != 0), S); }
void dm_log(const double *S, int corrected, double *v) { sim3_log(load(S), v, corrected != 0); }
void dm_edge_error(const double *C, const double *Si, const double *Sj, int corrected, double *e) {
    sim3_edge_error(load(C), load(Si), load(Sj), e, corrected != 0);
}
void dm_edge_jacobians(const double *C, const double *e, double *Ji, double *Jj) { sim3_edge_jacobians(load(C), e, Ji, Jj); }
// Jl^-1(e) as a full row-major 7x7 on the tangent [omega, upsilon, sigma]
void dm_jl_inv(const double *e, double *J) {
    JlInv L;
    sim3_jl_inv(e, L);
    for (int i = 0; i < 49; ++i) J[i] = 0;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) {
            J[r * 7 + c] = L.Jw[r * 3 + c];
            J[(3 + r) * 7 + c] = L.X[r * 3 + c];
            J[(3 + r) * 7 + 3 + c] = L.Wi[r * 3 + c];
        }
        J[(3 + r) * 7 + 6] = L.y[r];
    }
    J[48] = 1;
}
}
