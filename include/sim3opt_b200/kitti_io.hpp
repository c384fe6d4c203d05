// kitti_io.hpp -- host-side readers / writer for the KITTI-00 pose-graph fixtures (row a5).
//
// Semantics follow the reference:
//   key-frame list      cc.txt               one frame id per line            kitti_surf.cpp:232-254
//   frame poses         framePoses.txt       2 header lines, then
//                       "id, time, roll, pitch, yaw, x, y, z" of T_c2w; the key frame keeps
//                       T_w2c = (roteu2ro(rpy), xyz)^-1                       kitti_surf.cpp:255-292
//   loop constraints    loopConstraints.txt  5 header lines, then groups of 4 lines; the frame ids
//                       come from line 1, the Sim3 (count, scale, rpy, t) from line 4
//                                                                             kitti_surf.cpp:145-205
//   Euler -> DCM        roteu2ro             R = R3(yaw) R2(pitch) R1(roll)   kittiDetector.h:225-243
//   result file         "kfId s_w2i t_i_in_w q_i2w(xyzw)"                     kitti_surf.cpp:678-703
// Errors are reported by return value / exception-free bool, not by exit(-1) as in the reference.
#pragma once
#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "g2o_facade.hpp"

namespace s3o {
namespace kitti {

inline g2o::Matrix3 roteu2ro(const g2o::Vector3 &eul) {
    const double cr = std::cos(eul[0]), sr = std::sin(eul[0]);
    const double cp = std::cos(eul[1]), sp = std::sin(eul[1]);
    const double ch = std::cos(eul[2]), sh = std::sin(eul[2]);
    g2o::Matrix3 R;
    R(0, 0) = cp * ch; R(0, 1) = sp * sr * ch - cr * sh; R(0, 2) = cr * sp * ch + sh * sr;
    R(1, 0) = cp * sh; R(1, 1) = sr * sp * sh + cr * ch; R(1, 2) = cr * sp * sh - sr * ch;
    R(2, 0) = -sp;     R(2, 1) = sr * cp;                R(2, 2) = cr * cp;
    return R;
}

// Rigid transform world -> camera of a key frame (the reference keeps a Sophus::SE3d)
struct SE3 {
    g2o::Matrix3 R = g2o::Matrix3::Identity();
    g2o::Vector3 t;
    SE3() {}
    SE3(const g2o::Matrix3 &R_, const g2o::Vector3 &t_) : R(R_), t(t_) {}
    SE3(const g2o::Quaternion &q, const g2o::Vector3 &t_) : R(q.toRotationMatrix()), t(t_) {}
    SE3 inverse() const { const g2o::Matrix3 Rt = R.transpose(); return SE3(Rt, -(Rt * t)); }
    const g2o::Matrix3 &rotationMatrix() const { return R; }
    const g2o::Vector3 &translation() const { return t; }
};

struct KeyFrame {                 // kittiDetector.h:425-435
    int mnId = -1;                // index of the key frame in the image sequence
    int mnFrameId = -1;           // index among the key frames
    SE3 Tw2c;
    KeyFrame(int nid = -1, int kfid = -1) : mnId(nid), mnFrameId(kfid) {}
    bool isBad() const { return false; }
    g2o::Matrix3 GetRotation() const { return Tw2c.rotationMatrix(); }
    g2o::Vector3 GetTranslation() const { return Tw2c.translation(); }
    void SetPose(const SE3 &T) { Tw2c = T; }
};

template <class Trans, int DoF>
struct Constraint {               // kittiDetector.h:404-422
    int trans_id1, trans_id2;     // first / second frame
    Trans mean;                   // S_second<-first, scale = s_second / s_first
    Matrix<double, DoF, DoF> fisher_information;
    Constraint(int id1, int id2, const Trans &m, const Matrix<double, DoF, DoF> &info)
        : trans_id1(id1), trans_id2(id2), mean(m), fisher_information(info) {}
};
typedef Constraint<g2o::Sim3, 7> Sim3Constraint;

inline bool LoadKFIndices(const std::string &ccFile, std::vector<KeyFrame> &kfs) {
    kfs.clear();
    std::ifstream in(ccFile);
    if (!in) return false;
    int frame;
    while (in >> frame) kfs.emplace_back(frame, (int)kfs.size());
    return !kfs.empty();
}

inline bool LoadKFPoses(const std::string &poseFile, std::vector<KeyFrame> &kfs) {
    std::ifstream in(poseFile);
    if (!in) return false;
    std::string line;
    for (int h = 0; h < 2; ++h) std::getline(in, line);
    size_t next = 0;
    while (std::getline(in, line) && next < kfs.size()) {
        for (char &c : line) if (c == ',') c = ' ';
        std::istringstream ss(line);
        int frame;
        double time, v[6];
        if (!(ss >> frame >> time >> v[0] >> v[1] >> v[2] >> v[3] >> v[4] >> v[5])) continue;
        if (frame != kfs[next].mnId) continue;
        const SE3 Tc2w(roteu2ro(g2o::Vector3(v[0], v[1], v[2])), g2o::Vector3(v[3], v[4], v[5]));
        kfs[next++].Tw2c = Tc2w.inverse();
    }
    return next == kfs.size();
}

inline bool GetAllKeyFrames(const std::string &dir, std::vector<KeyFrame> &kfs) {
    return LoadKFIndices(dir + "/cc.txt", kfs) && LoadKFPoses(dir + "/framePoses.txt", kfs);
}

inline bool LoadLoopConstraints(const std::string &file, std::vector<Sim3Constraint> &out) {
    out.clear();
    std::ifstream in(file);
    if (!in) return false;
    std::string l1, l2, l3, l4;
    for (int h = 0; h < 5; ++h) std::getline(in, l1);
    while (std::getline(in, l1) && std::getline(in, l2) && std::getline(in, l3) && std::getline(in, l4)) {
        std::istringstream first(l1), fourth(l4);
        unsigned id1, id2;
        int matches;
        double scale, v[6];
        if (!(first >> id1 >> id2)) break;
        if (!(fourth >> matches >> scale >> v[0] >> v[1] >> v[2] >> v[3] >> v[4] >> v[5])) return false;
        const g2o::Sim3 S(roteu2ro(g2o::Vector3(v[0], v[1], v[2])), g2o::Vector3(v[3], v[4], v[5]), scale);
        out.emplace_back((int)id1, (int)id2, S, Matrix<double, 7, 7>::Identity());
    }
    return true;
}

inline g2o::Vector4 toScaleTrans(const g2o::Sim3 &S) {     // kitti_surf.cpp:533-539
    return g2o::Vector4(S.scale(), S.translation()[0], S.translation()[1], S.translation()[2]);
}

// "kfId s_w2i t_i_in_w q_i2w(xyzw)".  precision 6 reproduces the reference's default ostream
// output; the hand-off between stages should use 17.
inline void WriteSim3Line(std::ostream &os, int frame_id, const g2o::Sim3 &Siw, int precision = 6) {
    const g2o::Sim3 Swi = Siw.inverse();
    const g2o::Vector4 q = Swi.rotation().coeffs();
    os << std::setprecision(precision) << frame_id << " " << Siw.scale() << " " << Swi.translation()[0] << " "
       << Swi.translation()[1] << " " << Swi.translation()[2] << " " << q[0] << " " << q[1] << " " << q[2] << " "
       << q[3] << "\n";
}

// KITTI odometry ground truth: every line holds the 12 numbers of the 3x4 T_c2w, row-major
// (readKITTIPoseFile, kitti_surf.cpp:1166-1190).  Keeps the camera positions (column 3).
inline bool ReadKITTIPosePositions(const std::string &poseFile, std::vector<g2o::Vector3> &positions) {
    std::ifstream in(poseFile);
    if (!in) return false;
    positions.clear();
    double m[12];
    while (in >> m[0]) {
        for (int j = 1; j < 12; ++j) if (!(in >> m[j])) return false;
        positions.push_back(g2o::Vector3(m[3], m[7], m[11]));
    }
    return true;
}

// Result file of the optimisers: a '%' comment line, then "kfId s_w2i t_i_in_w q_i2w(xyzw)" per key frame
// (readOptimizedSim3PoseFile, kitti_surf.cpp:1197-1228; the scale and the rotation are read and dropped).
inline bool ReadOptimizedSim3Positions(const std::string &poseFile, std::vector<int> &frameIds,
                                       std::vector<g2o::Vector3> &positions) {
    std::ifstream in(poseFile);
    if (!in) return false;
    std::string header;
    std::getline(in, header);
    frameIds.clear();
    positions.clear();
    int id;
    double s, x, y, z, q[4];
    while (in >> id) {
        if (!(in >> s >> x >> y >> z >> q[0] >> q[1] >> q[2] >> q[3])) return false;
        frameIds.push_back(id);
        positions.push_back(g2o::Vector3(x, y, z));
    }
    return true;
}

}  // namespace kitti
}  // namespace s3o
