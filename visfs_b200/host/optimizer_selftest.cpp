// optimizer_selftest.cpp — drives VISFS::Optimizer::Optimizer exactly as corelib/src/Estimator.cpp:254 does,
// from a window file written by tests/host_io.py, and dumps what came back.
//   optimizer_selftest marshal <in> <out>   host-only: the flat arrays handed to the C ABI (no GPU needed)
//   optimizer_selftest solve   <in> <out>   full localOptimize on the GPU
//   optimizer_selftest time    <in> <out>   the same 20 times: wall time per call and the marshalling share
//   optimizer_selftest resident <in> <out>  the same window fed to VISFS::Optimizer::ResidentLocalMap signature by signature
//                                           (LocalMap's deltas), then ONE localOptimize on the resident map: same output file
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "Optimizer.h"

using namespace VISFS;

namespace {
template <typename T> T rd(std::istream &is) { T v; is.read(reinterpret_cast<char *>(&v), sizeof v); return v; }
template <typename T> void wr(std::ostream &os, const T &v) { os.write(reinterpret_cast<const char *>(&v), sizeof v); }
template <typename T> void wrv(std::ostream &os, const std::vector<T> &v) {
    wr<int64_t>(os, (int64_t)v.size());
    if (!v.empty()) os.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
}
template <typename T> void wrv(std::ostream &os, const Optimizer::detail::HostArray<T> &v) {
    wr<int64_t>(os, (int64_t)v.size());
    if (!v.empty()) os.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
}
}  // namespace

int main(int argc, char **argv) {
    if (argc != 4) { std::fprintf(stderr, "usage: %s marshal|marshaltime|solve|time|resident <in> <out>\n", argv[0]); return 2; }
    const std::string mode = argv[1];
    std::ifstream in(argv[2], std::ios::binary);
    if (!in) { std::fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
    const int64_t P = rd<int64_t>(in), L = rd<int64_t>(in), E = rd<int64_t>(in), rootId = rd<int64_t>(in), nCam = rd<int64_t>(in);
    const int64_t iterations = rd<int64_t>(in), solver = rd<int64_t>(in), trust = rd<int64_t>(in);
    const double fx = rd<double>(in), fy = rd<double>(in), cx = rd<double>(in), cy = rd<double>(in), baseline = rd<double>(in);
    const double pixelVariance = rd<double>(in), delta = rd<double>(in);

    std::map<std::size_t, Eigen::Isometry3d> poses;
    for (int64_t i = 0; i < P; ++i) {
        const int64_t id = rd<int64_t>(in);
        Eigen::Isometry3d T;
        double M[16];
        for (double &v : M) v = rd<double>(in);
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.linear()(r, c) = M[4 * r + c]; T.translation()[r] = M[4 * r + 3]; }
        poses.emplace((std::size_t)id, T);
    }
    std::map<std::size_t, std::tuple<Eigen::Vector3d, bool>> points3D;
    for (int64_t i = 0; i < L; ++i) {
        const int64_t id = rd<int64_t>(in);
        const double x = rd<double>(in), y = rd<double>(in), z = rd<double>(in);
        const int64_t fixed = rd<int64_t>(in);
        points3D.emplace((std::size_t)id, std::make_tuple(Eigen::Vector3d(x, y, z), fixed != 0));
    }
    std::map<std::size_t, std::map<std::size_t, Optimizer::FeatureBA>> wordReferences;
    for (int64_t i = 0; i < E; ++i) {
        const int64_t fid = rd<int64_t>(in), pid = rd<int64_t>(in);
        const float x = rd<float>(in), y = rd<float>(in), depth = rd<float>(in);
        (void)rd<float>(in);
        wordReferences[(std::size_t)fid].emplace((std::size_t)pid, Optimizer::FeatureBA(cv::KeyPoint(x, y), depth));
    }
    // optional trailer: odometry links (count, odometry covariance, then from id, to id, 4x4 robot-frame transform)
    std::map<std::size_t, std::tuple<std::size_t, std::size_t, Eigen::Isometry3d>> links;
    double odometryCovariance = 0.00005;
    {
        const int64_t nLinks = rd<int64_t>(in);
        if (in && nLinks > 0) {
            odometryCovariance = rd<double>(in);
            for (int64_t i = 0; i < nLinks; ++i) {
                const int64_t from = rd<int64_t>(in), to = rd<int64_t>(in);
                Eigen::Isometry3d T;
                double M[16];
                for (double &v : M) v = rd<double>(in);
                for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T.linear()(r, c) = M[4 * r + c]; T.translation()[r] = M[4 * r + 3]; }
                links.emplace((std::size_t)i, std::make_tuple((std::size_t)from, (std::size_t)to, T));
            }
        }
    }
    std::vector<std::shared_ptr<GeometricCamera>> cameraModels;
    for (int64_t c = 0; c < nCam; ++c)
        cameraModels.push_back(std::make_shared<GeometricCamera>(std::vector<double>{fx, fy, cx, cy, baseline}));

    std::ofstream out(argv[3], std::ios::binary);
    if (mode == "marshaltime") {   // the map walk alone, no device needed: ms per call over 50 calls
        Optimizer::detail::MarshalledWindow m;
        Optimizer::Optimizer::marshal((std::size_t)rootId, poses, cameraModels, points3D, wordReferences, m);
        const auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < 50; ++r) Optimizer::Optimizer::marshal((std::size_t)rootId, poses, cameraModels, points3D, wordReferences, m);
        std::printf("{\"marshal_ms\": %.4f, \"edges\": %lld}\n",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / 50, (long long)m.edge_pose.size());
        return 0;
    }
    if (mode == "marshal") {
        Optimizer::detail::MarshalledWindow m;
        const bool ok = Optimizer::Optimizer::marshal((std::size_t)rootId, poses, cameraModels, points3D, wordReferences, m);
        wr<int64_t>(out, ok ? 1 : 0);
        wrv(out, m.pose_tq); wrv(out, m.pose_id); wrv(out, m.pose_fixed); wrv(out, m.point_xyz); wrv(out, m.point_id);
        wrv(out, m.point_fixed);
        { std::vector<double> wide(m.edge_obs.begin(), m.edge_obs.end()); wrv(out, wide); }   // (floats, dumped widened)
        wrv(out, m.edge_pose); wrv(out, m.edge_point); wrv(out, m.edge_kind);
        const std::vector<double> intr{m.fx, m.fy, m.cx, m.cy, m.bf};
        wrv(out, intr);
        return 0;
    }

    ParametersMap params;
    params["Optimizer/Iterations"] = std::to_string(iterations);
    params["Optimizer/Solver"] = std::to_string(solver);
    params["Optimizer/TrustRegion"] = std::to_string(trust);
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.17g", pixelVariance); params["Optimizer/PixelVariance"] = buf;
    std::snprintf(buf, sizeof buf, "%.17g", delta); params["Optimizer/RobustKernelDelta"] = buf;
    std::snprintf(buf, sizeof buf, "%.17g", odometryCovariance); params["Optimizer/OdometryCovariance"] = buf;
    if (mode == "resident") {
        Optimizer::ResidentLocalMap map(params, cameraModels, (int)std::max<int64_t>(P, 2), (int)std::max<int64_t>(L, 1));
        bool ok = true;
        for (auto &kv : points3D) ok = ok && map.setFeature(kv.first, std::get<0>(kv.second), std::get<1>(kv.second));
        for (auto &pk : poses) {   // ascending signature id, each with its observations (LocalMap::insertSignature)
            std::map<std::size_t, Optimizer::FeatureBA> obs;
            for (auto &fk : wordReferences) {
                auto it = fk.second.find(pk.first);
                if (it != fk.second.end() && points3D.count(fk.first)) obs.emplace(fk.first, it->second);
            }
            ok = ok && map.insertSignature(pk.first, pk.second, obs);
        }
        ok = ok && map.setLinks(links);   // LocalMap::getSignatureLinks -> the resident map's link set
        if (!ok) { std::fprintf(stderr, "ResidentLocalMap: %s\n", map.lastMessage().c_str()); return 1; }
        std::vector<std::tuple<std::size_t, std::size_t>> outliers;
        auto result = map.localOptimize((std::size_t)rootId, outliers);
        wr<int64_t>(out, (int64_t)result.size());
        for (auto &kv : result) {
            wr<int64_t>(out, (int64_t)kv.first);
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) wr<double>(out, kv.second.linear()(r, c)); wr<double>(out, kv.second.translation()[r]); }
            wr<double>(out, 0.0); wr<double>(out, 0.0); wr<double>(out, 0.0); wr<double>(out, 1.0);
        }
        wr<int64_t>(out, (int64_t)points3D.size());
        for (auto &kv : points3D) {
            Eigen::Vector3d p;
            map.getFeaturePose(kv.first, p);
            wr<int64_t>(out, (int64_t)kv.first);
            wr<double>(out, p[0]); wr<double>(out, p[1]); wr<double>(out, p[2]);
        }
        wr<int64_t>(out, (int64_t)outliers.size());
        for (auto &o : outliers) { wr<int64_t>(out, (int64_t)std::get<0>(o)); wr<int64_t>(out, (int64_t)std::get<1>(o)); }
        std::cout << "poses " << result.size() << " outliers " << outliers.size() << " message '" << map.lastMessage() << "'\n";
        return 0;
    }
    Optimizer::Optimizer optimizer(params);
    std::vector<std::tuple<std::size_t, std::size_t>> outliers;
    const std::vector<Sensor::PointCloud> pointClouds;
    const std::shared_ptr<const Map::Submap2D> submap;
    if (mode == "time") {
        // what Estimator::process pays per key frame: the whole call, maps in, maps out
        const auto pointsIn = points3D;
        double best = 1e30, sum = 0.0, marshalMs = 0.0;
        const int reps = 20;
        for (int r = 0; r < reps + 3; ++r) {
            points3D = pointsIn; outliers.clear();
            const auto t0 = std::chrono::steady_clock::now();
            auto res = optimizer.localOptimize((std::size_t)rootId, poses, links, cameraModels, points3D, wordReferences,
                                               pointClouds, submap, outliers);
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (res.size() != poses.size()) { std::fprintf(stderr, "localOptimize failed: %s\n", optimizer.lastMessage().c_str()); return 1; }
            if (r >= 3) { best = std::min(best, ms); sum += ms; }
        }
        {
            Optimizer::detail::MarshalledWindow m;
            Optimizer::Optimizer::marshal((std::size_t)rootId, poses, cameraModels, points3D, wordReferences, m);   // sizes the arrays
            const auto t0 = std::chrono::steady_clock::now();
            for (int r = 0; r < reps; ++r) Optimizer::Optimizer::marshal((std::size_t)rootId, poses, cameraModels, points3D, wordReferences, m);
            marshalMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        }
        std::printf("{\"poses\": %lld, \"points\": %lld, \"edges\": %lld, \"local_optimize_ms_mean\": %.4f, \"local_optimize_ms_best\": %.4f, \"marshal_ms\": %.4f}\n",
                    (long long)P, (long long)L, (long long)E, sum / reps, best, marshalMs);
        return 0;
    }
    auto result = optimizer.localOptimize((std::size_t)rootId, poses, links, cameraModels, points3D, wordReferences,
                                          pointClouds, submap, outliers);
    wr<int64_t>(out, (int64_t)result.size());
    for (auto &kv : result) {
        wr<int64_t>(out, (int64_t)kv.first);
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) wr<double>(out, kv.second.linear()(r, c)); wr<double>(out, kv.second.translation()[r]); }
        wr<double>(out, 0.0); wr<double>(out, 0.0); wr<double>(out, 0.0); wr<double>(out, 1.0);
    }
    wr<int64_t>(out, (int64_t)points3D.size());
    for (auto &kv : points3D) {
        wr<int64_t>(out, (int64_t)kv.first);
        const Eigen::Vector3d &p = std::get<0>(kv.second);
        wr<double>(out, p[0]); wr<double>(out, p[1]); wr<double>(out, p[2]);
    }
    wr<int64_t>(out, (int64_t)outliers.size());
    for (auto &o : outliers) { wr<int64_t>(out, (int64_t)std::get<0>(o)); wr<int64_t>(out, (int64_t)std::get<1>(o)); }
    std::cout << "poses " << result.size() << " outliers " << outliers.size() << " message '" << optimizer.lastMessage() << "'\n";
    return 0;
}
