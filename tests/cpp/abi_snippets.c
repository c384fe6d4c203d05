/* Compiled (never run) by tests/test_abi.py with a plain C compiler: include/sim3opt_b200.h must be valid C,
 * and the call sequences shown in INTEGRATION.md (sections 2, 6, 7) must type-check against it. */
#include <stddef.h>
#include <stdint.h>

#include "sim3opt_b200.h"

int integration_section_2(int n, const double *est, const uint8_t *fixed, int ne, const int32_t *v0, const int32_t *v1,
                          const double *meas, double *out) {
    s3o_problem *prob = NULL;
    int iters = 0;
    double chi2 = 0, lambda = 0;
    if (s3o_create(S3O_KIND_SIM3, 0, &prob) != S3O_OK) return -1;
    s3o_set_vertices(prob, n, est, fixed, NULL);
    s3o_set_edges(prob, ne, v0, v1, meas, NULL);
    s3o_set_jacobian_mode(prob, S3O_JAC_NUMERIC, 1e-9);
    s3o_set_lm(prob, 1e-5, 0.0, 10);
    s3o_set_pcg(prob, 1e-8, 1000);
    s3o_set_preconditioner(prob, S3O_PRECOND_AUTO);
    s3o_set_math_mode(prob, S3O_MATH_REFERENCE);
    s3o_build_structure(prob, NULL, NULL);
    s3o_optimize(prob, 100, 0.0, &iters, &chi2, &lambda, NULL, 0);
    s3o_get_vertices(prob, out);
    return s3o_destroy(prob);
}

int integration_section_6(int C, double *cams, const uint8_t *cam_fixed, int P, double *xyz, int M, const int32_t *cam_idx,
                          const int32_t *point_idx, const double *uv) {
    s3o_problem *ba = NULL;
    int iters = 0;
    double chi2 = 0, lambda = 0;
    if (s3o_create(S3O_KIND_BA, 0, &ba) != S3O_OK) return -1;
    s3o_ba_set_intrinsics(ba, 718.856, 607.1928, 185.2157);
    s3o_ba_set_cameras(ba, C, cams, cam_fixed);
    s3o_ba_set_points(ba, P, xyz, NULL);
    s3o_ba_set_observations(ba, M, cam_idx, point_idx, uv, NULL);
    s3o_set_robust(ba, S3O_ROBUST_HUBER, 2.5);
    s3o_optimize(ba, 20, 0.0, &iters, &chi2, &lambda, NULL, 0);
    s3o_ba_get_cameras(ba, cams);
    s3o_ba_get_points(ba, xyz);
    return s3o_destroy(ba);
}

int integration_section_7(s3o_problem *p, int rank, int world, int n, const double *q, const double *t) {
    char id[128];
    double S221[16], rmse = 0, max_dev = 0;
    s3o_stats st;
    if (rank == 0) s3o_comm_unique_id(id);
    s3o_set_comm(p, rank, world, id);
    s3o_get_stats(p, &st);
    s3o_align_similarity(0, n, q, t, 0, S221, &rmse, &max_dev);
    return st.p2p_halo + st.multilevel_levels;
}
