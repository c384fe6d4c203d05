// Reads tangent vectors [upsilon omega sigma] (sim3_rv.h order), one per line, and prints for each line
//   R (9, row-major) t (3) s  |  ln(exp(v)) (7)  |  (exp(v) * exp(v_prev)^-1) as R t s
// tests/test_sim3_rv.py compares the numbers with the CPU oracle (oracle/lie.c, g2o tangent order).
#include <cstdio>
#include <iostream>

#include "sim3opt_b200/sim3_rv.hpp"

int main(int argc, char **argv) {
    using S = RobotVision::Sim3<>;
    if (argc > 1 && argv[1][0] == 'c') S::corrected_limits() = true;
    S prev;
    double v[7];
    while (std::scanf("%lf %lf %lf %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6]) == 7) {
        S::Vec7 x;
        for (int i = 0; i < 7; ++i) x[i] = v[i];
        const S T = S::exp(x);
        const S::Vec7 back = T.ln();
        const S rel = T * prev.inverse();
        auto dump = [](const S &A) {
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) std::printf("%.17g ", A.get_rotation()(r, c));
            for (int i = 0; i < 3; ++i) std::printf("%.17g ", A.get_translation()[i]);
            std::printf("%.17g ", A.get_scale());
        };
        dump(T);
        for (int i = 0; i < 7; ++i) std::printf("%.17g ", back[i]);
        dump(rel);
        std::printf("\n");
        prev = T;
    }
    return 0;
}
