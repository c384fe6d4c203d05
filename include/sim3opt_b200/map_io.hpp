// map_io.hpp -- map re-projection and BAL export: turns the PTAM key-frame
// dumps plus an optimised key-frame trajectory into the bundle-adjustment input bal_example reads.
//
// Semantics follow the reference:
//   LoadComboKeyFrame   KeyFrame%06d.bin: int32 id, int32 name length + name, 2 f64 image size, 5 f64 camera
//                       parameters, 9 f64 R_w2c (row-major), 3 f64 t_w_in_c, 1 byte fixed flag, int32 count, then
//                       count x {uint32 point id, 3 f64 p_w, f64 cos(init angle), 2 f64 pixel}
//                                                                              drawPTAMPoints.cpp:33-84
//   ReprojectMap        figureKITTIBA: merge the points of all key frames (later frames win), compact the
//                       point ids in ascending order, express every point in the frame that observed it
//                       (old pose), map it back with the optimised Sim3 of that frame (observations in file
//                       order, the last one wins), then move everything by S221 / S221^-1
//                                                                              drawPTAMPoints.cpp:285-456
//   SaveBALFile         header, "cam point u v" (%g), 9 numbers per camera (angle-axis of R_w2c, t, f, k1, k2;
//                       %.16g), points                                         drawPTAMPoints.cpp:218-283
// Host code: format conversion only, nothing here is on the optimisation path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "linalg.hpp"
#include "sim3_rv.hpp"

namespace s3o {
namespace mapio {

using Mat3 = s3o::Matrix<double, 3, 3>;
using Vec3 = s3o::Matrix<double, 3, 1>;

struct IdObs {                    // one image observation
    int frame_id = -1;            // image index on load, position in the pose list after compaction
    unsigned point_id = 0;
    double u = 0, v = 0;
};

struct ComboKeyFrame {
    unsigned id = 0;
    Mat3 Rw2c;
    Vec3 twinc;
    bool fixed = false;
    std::vector<IdObs> obs;
    std::vector<std::pair<unsigned, Vec3>> points;     // in file order
};

inline bool LoadComboKeyFrame(const std::string &file, ComboKeyFrame &kf) {
    std::ifstream in(file, std::ios::binary);
    if (!in) return false;
    auto rd = [&](void *dst, size_t n) { in.read(reinterpret_cast<char *>(dst), (std::streamsize)n); return (bool)in; };
    int32_t id = 0, len = 0, count = -1;
    if (!rd(&id, 4) || !rd(&len, 4) || len < 0 || len > 199) return false;
    char name[200];
    double size2[2], cam5[5], R[9], t[3];
    unsigned char fixed = 0;
    if (!rd(name, (size_t)len) || !rd(size2, 16) || !rd(cam5, 40) || !rd(R, 72) || !rd(t, 24) || !rd(&fixed, 1) || !rd(&count, 4) || count < 0)
        return false;
    kf.id = (unsigned)id;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) kf.Rw2c(r, c) = R[r * 3 + c]; kf.twinc[r] = t[r]; }
    kf.fixed = fixed != 0;
    kf.obs.resize((size_t)count);
    kf.points.clear();
    kf.points.reserve((size_t)count);
    for (int k = 0; k < count; ++k) {
        uint32_t pid = 0;
        double p[3], cosang, px[2];
        if (!rd(&pid, 4) || !rd(p, 24) || !rd(&cosang, 8) || !rd(px, 16)) return false;
        kf.points.emplace_back(pid, Vec3(p[0], p[1], p[2]));
        kf.obs[(size_t)k].frame_id = id;
        kf.obs[(size_t)k].point_id = pid;
        kf.obs[(size_t)k].u = px[0];
        kf.obs[(size_t)k].v = px[1];
    }
    return true;
}

struct BalProblem {
    std::vector<Mat3> Rw2c;       // per camera
    std::vector<Vec3> twinc;
    std::vector<Vec3> points;
    std::vector<IdObs> obs;       // compact camera / point indices
    double f = 718.856, k1 = 0, k2 = 0;     // drawPTAMPoints.cpp:443
    std::vector<int> image_ids;   // image index of every camera
};

// rotation matrix -> rotation vector (what rotro2qr + QuaternionToAngleAxis produce: angle in [0, pi])
inline Vec3 RotationVector(const Mat3 &R) {
    const s3o::Quaternion<double> q(R);
    double w = q.w(), x = q.x(), y = q.y(), z = q.z();
    if (w < 0) { w = -w; x = -x; y = -y; z = -z; }
    const double s2 = x * x + y * y + z * z;
    if (s2 > 0) {
        const double s = std::sqrt(s2);
        const double k = 2.0 * std::atan2(s, w) / s;
        return Vec3(x * k, y * k, z * k);
    }
    return Vec3(2 * x, 2 * y, 2 * z);
}

// figureKITTIBA with lineFormat = 1 ("kfId s_w2i t_i_in_w q_i2w(xyzw)" after one header line)
inline bool ReprojectMap(const std::string &keyFrameDir, const std::string &transFile, const RobotVision::Sim3<> &S221,
                         BalProblem &out, std::string *err = nullptr) {
    auto fail = [&](const std::string &m) { if (err) *err = m; return false; };
    namespace fs = std::filesystem;
    std::vector<std::string> files;
    std::error_code ec;
    for (const auto &e : fs::directory_iterator(keyFrameDir, ec))
        if (e.path().extension() == ".bin" && e.path().filename().string().rfind("KeyFrame", 0) == 0) files.push_back(e.path().string());
    if (ec || files.empty()) return fail("no KeyFrame*.bin in " + keyFrameDir);
    std::sort(files.begin(), files.end());
    out = BalProblem();
    std::map<unsigned, unsigned> frameid2poseid;
    std::map<unsigned, Vec3> pointsallframe;           // ascending point id
    std::vector<Mat3> oldR;
    std::vector<Vec3> oldt;
    for (size_t k = 0; k < files.size(); ++k) {
        ComboKeyFrame kf;
        if (!LoadComboKeyFrame(files[k], kf)) return fail("cannot read " + files[k]);
        frameid2poseid[kf.id] = (unsigned)k;
        out.image_ids.push_back((int)kf.id);
        oldR.push_back(kf.Rw2c);
        oldt.push_back(kf.twinc);
        for (const auto &pt : kf.points) pointsallframe[pt.first] = pt.second;      // the latest estimate wins
        out.obs.insert(out.obs.end(), kf.obs.begin(), kf.obs.end());
    }
    std::map<unsigned, unsigned> pointid2compactid;
    std::vector<Vec3> point_vec;
    for (const auto &pt : pointsallframe) { pointid2compactid[pt.first] = (unsigned)point_vec.size(); point_vec.push_back(pt.second); }
    for (IdObs &o : out.obs) {
        o.frame_id = (int)frameid2poseid.at((unsigned)o.frame_id);
        o.point_id = pointid2compactid.at(o.point_id);
    }
    // corrected poses
    std::ifstream in(transFile);
    if (!in) return fail("cannot open " + transFile);
    std::string header;
    std::getline(in, header);
    std::vector<RobotVision::Sim3<>> updated;          // corrected Sim3 world -> camera
    unsigned fid;
    double s, t[3], q[4];
    while (in >> fid >> s >> t[0] >> t[1] >> t[2] >> q[0] >> q[1] >> q[2] >> q[3]) {
        const Mat3 Rc2w = s3o::Quaternion<double>(q[3], q[0], q[1], q[2]).toRotationMatrix();
        const Mat3 Rw2c = Rc2w.transpose();
        const Vec3 twinc = (Rw2c * Vec3(t[0], t[1], t[2])) * (-s);
        updated.emplace_back(Rw2c, twinc, s);
        out.Rw2c.push_back(Rw2c);
        out.twinc.push_back(twinc / s);
    }
    if (updated.size() != files.size()) return fail("the pose file and the key-frame directory differ in length");
    // points: through the frame that observed them, observations in order (the last one wins)
    out.points = point_vec;
    for (const IdObs &o : out.obs) {
        const Vec3 rel = oldR[(size_t)o.frame_id] * point_vec[o.point_id] + oldt[(size_t)o.frame_id];
        out.points[o.point_id] = updated[(size_t)o.frame_id].inverse() * rel;
    }
    // align with the ground-truth frame
    const RobotVision::Sim3<> S122 = S221.inverse();
    for (size_t k = 0; k < out.Rw2c.size(); ++k) {
        const RobotVision::Sim3<> Sw12c = RobotVision::Sim3<>(out.Rw2c[k], out.twinc[k], 1.0) * S122;
        out.Rw2c[k] = Sw12c.get_rotation();
        out.twinc[k] = Sw12c.get_translation() / Sw12c.get_scale();
    }
    for (Vec3 &p : out.points) p = S221 * p;
    return true;
}

inline bool SaveBALFile(const BalProblem &P, const std::string &file) {
    FILE *f = std::fopen(file.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "%d %d %d\n", (int)P.Rw2c.size(), (int)P.points.size(), (int)P.obs.size());
    for (const IdObs &o : P.obs) std::fprintf(f, "%d %u %g %g\n", o.frame_id, o.point_id, o.u, o.v);
    for (size_t k = 0; k < P.Rw2c.size(); ++k) {
        const Vec3 aa = RotationVector(P.Rw2c[k]);
        const double row[9] = { aa[0], aa[1], aa[2], P.twinc[k][0], P.twinc[k][1], P.twinc[k][2], P.f, P.k1, P.k2 };
        for (double x : row) std::fprintf(f, "%.16g\n", x);
    }
    for (const Vec3 &p : P.points)
        for (int j = 0; j < 3; ++j) std::fprintf(f, "%.16g\n", p[j]);
    std::fclose(f);
    return true;
}

}  // namespace mapio
}  // namespace s3o
