// ba_demo.cpp -- the reference's bundle-adjustment driver (bal_example.cpp:44-243) on the B200 back-end.
//
// Reads the BAL-like text file the reference reads (header "numCameras numPoints numObservations",
// then observations "cam point u v", then 9 numbers per camera: angle-axis(3) t(3) f k1 k2 -- the
// last three ignored --, then the points; bal_example.cpp:104-194), builds the same g2o graph through
// include/sim3opt_b200/g2o_facade.hpp (VertexSE3Expmap ids 0..C-1, marginalised VertexSBAPointXYZ,
// EdgeProjectXYZ2UV with vertex(0) = point / vertex(1) = camera, Huber 2.5, one CameraParameters
// f = 718.856, pp = (607.1928, 185.2157)), runs Levenberg-Marquardt and writes
// "id t_c_in_w q_c2w(xyzw)" per camera (bal_example.cpp:216-241).
//
// usage: ba_demo [-i iterations] [-o outputFile] [-v] [-pcg] [-stats file] <graph-input>
//        (-pcg and -stats are accepted and ignored, as in the reference: bal_example.cpp:54,56)
#define S3O_FACADE_EIGEN_NAMES
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "sim3opt_b200/g2o_facade.hpp"

namespace {

// angle-axis -> unit quaternion (w, x, y, z); first-order form at the origin (bal_example.h:31-57)
void AngleAxisToQuaternion(const double aa[3], double q[4]) {
    const double th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    double k = 0.5;
    q[0] = 1.0;
    if (th2 > 0.0) {
        const double th = std::sqrt(th2);
        k = std::sin(0.5 * th) / th;
        q[0] = std::cos(0.5 * th);
    }
    q[1] = aa[0] * k; q[2] = aa[1] * k; q[3] = aa[2] * k;
}

}  // namespace

int main(int argc, char **argv) {
    int maxIterations = 5;
    bool verbose = false;
    std::string outputFilename, inputFilename;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        if (a == "-i" && k + 1 < argc) maxIterations = std::atoi(argv[++k]);
        else if (a == "-o" && k + 1 < argc) outputFilename = argv[++k];
        else if (a == "-stats" && k + 1 < argc) ++k;
        else if (a == "-v") verbose = true;
        else if (a == "-pcg") {}
        else inputFilename = a;
    }
    if (inputFilename.empty()) { std::cerr << "usage: ba_demo [-i n] [-o out] [-v] <graph-input>\n"; return 1; }

    const double PIXEL_NOISE = 1.0;
    const bool ROBUST_KERNEL = true;

    g2o::SparseOptimizer optimizer;
    optimizer.setVerbose(verbose);
    std::unique_ptr<g2o::BlockSolver_6_3::LinearSolverType> linearSolver =
        g2o::make_unique<g2o::LinearSolverEigen<g2o::BlockSolver_6_3::PoseMatrixType>>();
    optimizer.setAlgorithm(new g2o::OptimizationAlgorithmLevenberg(g2o::make_unique<g2o::BlockSolver_6_3>(std::move(linearSolver))));
    optimizer.setPcg(1e-10, 20000);

    g2o::CameraParameters *cam_params = new g2o::CameraParameters(718.856, Eigen::Vector2d{607.1928, 185.2157}, 0.);
    cam_params->setId(0);
    if (!optimizer.addParameter(cam_params)) { std::cerr << "cannot add the camera parameters\n"; return 2; }

    std::cout << "Loading BAL dataset " << inputFilename << std::endl;
    std::ifstream ifs(inputFilename);
    int numCameras = 0, numPoints = 0, numObservations = 0;
    if (!(ifs >> numCameras >> numPoints >> numObservations)) { std::cerr << "cannot read " << inputFilename << "\n"; return 2; }
    std::cerr << "numCameras=" << numCameras << " numPoints=" << numPoints << " numObservations=" << numObservations << std::endl;

    std::vector<g2o::VertexSE3Expmap *> cameras;
    std::vector<g2o::VertexSBAPointXYZ *> points;
    int id = 0;
    for (int i = 0; i < numCameras; ++i, ++id) {
        g2o::VertexSE3Expmap *cam = new g2o::VertexSE3Expmap();
        cam->setId(id);
        optimizer.addVertex(cam);
        cameras.push_back(cam);
    }
    for (int i = 0; i < numPoints; ++i, ++id) {
        g2o::VertexSBAPointXYZ *p = new g2o::VertexSBAPointXYZ();
        p->setId(id);
        p->setMarginalized(true);
        if (!optimizer.addVertex(p)) std::cerr << "failing adding vertex" << std::endl;
        points.push_back(p);
    }
    for (int i = 0; i < numObservations; ++i) {
        int camIndex, pointIndex;
        double obsX, obsY;
        ifs >> camIndex >> pointIndex >> obsX >> obsY;
        if (camIndex < 0 || camIndex >= numCameras || pointIndex < 0 || pointIndex >= numPoints) { std::cerr << "observation " << i << ": index out of bounds\n"; return 2; }
        g2o::EdgeProjectXYZ2UV *e = new g2o::EdgeProjectXYZ2UV();
        e->setVertex(0, points[pointIndex]);
        e->setVertex(1, cameras[camIndex]);
        e->setInformation(Eigen::Matrix<double, 2, 2>::Identity() / (PIXEL_NOISE * PIXEL_NOISE));
        e->setMeasurement(Eigen::Vector2d{obsX, obsY});
        if (ROBUST_KERNEL) {
            g2o::RobustKernelHuber *rk = new g2o::RobustKernelHuber;
            rk->setDelta(2.5);
            e->setRobustKernel(rk);
        }
        e->setParameterId(0, 0);
        if (!optimizer.addEdge(e)) std::cerr << "error adding edge" << std::endl;
    }
    for (int i = 0; i < numCameras; ++i) {
        double c[9];
        for (int j = 0; j < 9; ++j) ifs >> c[j];
        double q[4];
        AngleAxisToQuaternion(c, q);
        cameras[i]->setEstimate(g2o::SE3Quat(Eigen::Quaterniond(q[0], q[1], q[2], q[3]), Eigen::Vector3d(c[3], c[4], c[5])));
    }
    for (int i = 0; i < numPoints; ++i) {
        Eigen::Vector3d p;
        ifs >> p(0) >> p(1) >> p(2);
        points[i]->setEstimate(p);
    }
    if (!ifs) { std::cerr << "truncated input file\n"; return 2; }
    std::cout << "done." << std::endl;

    if (!optimizer.initializeOptimization()) { std::cerr << "initializeOptimization: " << optimizer.lastError() << "\n"; return 3; }
    optimizer.computeActiveErrors();
    double maxError = 0;
    for (const g2o::Edge *e : optimizer.edges()) maxError = std::max(maxError, static_cast<const g2o::EdgeProjectXYZ2UV *>(e)->error().norm());
    std::cout << "max edge error norm " << maxError << std::endl;
    std::cout << std::setprecision(12) << "initial chi2 " << optimizer.activeRobustChi2() << " free cameras " << optimizer.numFreeVertices()
              << " schur blocks " << optimizer.numHessianBlocks() << std::endl;

    std::cout << "\nPerforming full BA:" << std::endl;
    const auto t0 = std::chrono::steady_clock::now();
    const int iters = optimizer.optimize(maxIterations);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (iters <= 0) { std::cerr << "optimize: " << optimizer.lastError() << "\n"; return 3; }
    const std::vector<double> &h = optimizer.history();
    std::cout << "iterations " << iters << " chi2_final " << h[h.size() - 5] << " lambda_final " << h[h.size() - 4] << " optimize_ms " << ms << std::endl;

    if (!outputFilename.empty()) {
        std::ofstream fout(outputFilename);
        fout << "% SE3 optimization result: kf id, tcinw, rc2w(qxyzw):" << std::endl;
        fout << std::setprecision(17);
        int jack = 0;
        for (const g2o::VertexSE3Expmap *cam : cameras) {
            const g2o::SE3Quat est = cam->estimate();
            const Eigen::Quaterniond qc2w = est.rotation().conjugate();
            const Eigen::Vector3d tcinw = -(qc2w * est.translation());
            fout << jack++ << " " << tcinw.transpose() << " " << qc2w.coeffs().transpose() << std::endl;
        }
    }
    return 0;
}
