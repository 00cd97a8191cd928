// Optimizer.cpp — the reference-facing plugin: VISFS::Optimizer::Optimizer::localOptimize on the B200.
// Follows corelib/src/Optimizer/Optimizer.cpp:58-364 (g2o branch) step by step; the numbered comments
// are the reference's line ranges.  All arithmetic that g2o did runs behind visfs_ba_solve().
#include "Optimizer.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <limits>

#include "visfs_ba.h"

namespace VISFS {
namespace Optimizer {

namespace {

int g_device = 0;

#ifndef LOG_ERROR
// stand-alone build: the reference's Boost.Log macros (utilite/include/Log.h:38-43) become stderr lines
void logLine(const char * level, const std::string & text) { std::fprintf(stderr, "[VISFS %s] %s\n", level, text.c_str()); }
#define VISFS_B200_LOG(level, text) logLine(level, text)
#else
#define VISFS_B200_LOG(level, text) do { if (level[0] == 'E') { LOG_ERROR << text; } else { LOG_WARN << text; } } while (0)
#endif

// Eigen::Quaterniond(Matrix3d) followed by CameraPose::normalizeRotation (OptimizeTypeDefine.h:30-41)
void rotationToQuaternion(const Eigen::Matrix3d & m, double q[4] /* x y z w */) {
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    double x, y, z, w;
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        w = 0.5 * t;
        t = 0.5 / t;
        x = (m(2, 1) - m(1, 2)) * t;
        y = (m(0, 2) - m(2, 0)) * t;
        z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        x = v[0]; y = v[1]; z = v[2];
    }
    if (w < 0.0) { x = -x; y = -y; z = -z; w = -w; }
    const double n = std::sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}

// CameraPose::toHomogeneousMatrix (OptimizeTypeDefine.h:74-81): Eigen toRotationMatrix of (x y z w)
Eigen::Isometry3d stateToIsometry(const double * tq) {
    const double x = tq[3], y = tq[4], z = tq[5], w = tq[6];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    // (with real Eigen Transform::linear() returns a Block by value, so it cannot bind to a Matrix3d &: build R locally)
    Eigen::Matrix3d R;
    R(0, 0) = 1.0 - (tyy + tzz); R(0, 1) = txy - twz;         R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;         R(1, 1) = 1.0 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;         R(2, 1) = tyz + twx;         R(2, 2) = 1.0 - (txx + tyy);
    Eigen::Isometry3d T = Eigen::Isometry3d::Identity();
    T.linear() = R;
    T.translation() = Eigen::Vector3d(tq[0], tq[1], tq[2]);
    return T;
}

bool isZeroTransform(const Eigen::Isometry3d & T) {
    // t.isApprox(Isometry3d(Matrix4d::Zero())) (Optimizer.cpp:330): true only for an all-zero matrix
    for (int i = 0; i < 3; ++i) {
        if (T.translation()[i] != 0.0) return false;
        for (int j = 0; j < 3; ++j) if (T.linear()(i, j) != 0.0) return false;
    }
    return true;
}

}  // namespace

void Optimizer::setDevice(int _device) { g_device = _device; }

// corelib/src/Optimizer/Optimizer.cpp:37-56
Optimizer::Optimizer(const ParametersMap & _parameters) :
    framework_(Parameters::defaultOptimizerFramework()),
    solver_(Parameters::defaultOptimizerSolver()),
    trustRegion_(Parameters::defaultOptimizerTrustRegion()),
    iterations_(Parameters::defaultOptimizerIterations()),
    pixelVariance_(Parameters::defaultOptimizerPixelVariance()),
    odometryCovariance_(Parameters::defaultOptimizerOdometryCovariance()),
    laserCovariance_(Parameters::defaultOptimizerLaserCovariance()),
    robustKernelDelta_(Parameters::defaultOptimizerRobustKernelDelta()),
    handle_(nullptr) {
    Parameters::parse(_parameters, Parameters::kOptimizerFramework(), framework_);
    Parameters::parse(_parameters, Parameters::kOptimizerSolver(), solver_);
    Parameters::parse(_parameters, Parameters::kOptimizerTrustRegion(), trustRegion_);
    Parameters::parse(_parameters, Parameters::kOptimizerIterations(), iterations_);
    Parameters::parse(_parameters, Parameters::kOptimizerPixelVariance(), pixelVariance_);
    Parameters::parse(_parameters, Parameters::kOptimizerOdometryCovariance(), odometryCovariance_);
    Parameters::parse(_parameters, Parameters::kOptimizerLaserCovariance(), laserCovariance_);
    Parameters::parse(_parameters, Parameters::kOptimizerRobustKernelDelta(), robustKernelDelta_);
}

Optimizer::~Optimizer() {
    if (handle_) visfs_ba_destroy(handle_);
}

bool Optimizer::ensureHandle() {
    if (handle_) return true;
    visfs_ba_config cfg{};
    cfg.abi_version = VISFS_BA_ABI_VERSION;
    cfg.device = g_device;
    if (visfs_ba_create(&cfg, &handle_) != VISFS_BA_OK) {
        const char * why = visfs_ba_last_error(nullptr);
        message_ = std::string("Optimizer: cannot create the CUDA bundle adjuster: ") + (why ? why : "");
        VISFS_B200_LOG("ERROR", message_);
        handle_ = nullptr;
        return false;
    }
    return true;
}

namespace detail {
void * hostAlloc(std::size_t _bytes, bool * _pinned) {
    void * p = visfs_ba_host_alloc(_bytes);      // NULL without a CUDA device
    *_pinned = p != nullptr;
    if (!p) p = std::malloc(_bytes ? _bytes : 1);
    return p;
}
void hostFree(void * _p, bool _pinned) {
    if (!_p) return;
    if (_pinned) visfs_ba_host_free(_p); else std::free(_p);
}
}   // detail

// Optimizer.cpp:100-114 (poses) and :152-223 (points, visual edges) of the reference
bool Optimizer::marshal(std::size_t _rootId,
                        const std::map<std::size_t, Eigen::Isometry3d> & _poses,
                        const std::vector<std::shared_ptr<GeometricCamera>> & _cameraModels,
                        const std::map<std::size_t, std::tuple<Eigen::Vector3d, bool>> & _points3D,
                        const std::map<std::size_t, std::map<std::size_t, FeatureBA>> & _wordReferences,
                        detail::MarshalledWindow & m) {
    if (_cameraModels.empty() || !_cameraModels.front()) return false;
    const GeometricCamera & cameraModel = *_cameraModels.front();
    const Eigen::Isometry3d Trc = cameraModel.getTansformImageToRobot();
    m.clear();

    // pose ids in ascending order with their indices: the observations of a feature are keyed by signature id in ascending
    // order too (std::map), so both lookups below are merge walks instead of tree searches
    std::vector<std::size_t> poseKey;
    for (auto iter = _poses.begin(); iter != _poses.end(); ++iter) {
        if (iter->first > 0) {                                            // :101
            Eigen::Isometry3d cameraPose = iter->second * Trc;           // :104  Twc = Twr * Trc
            cameraPose = cameraPose.inverse();                            // :108  Tcw
            double q[4];
            rotationToQuaternion(cameraPose.linear(), q);                 // :109  CameraPose(R, t)
            poseKey.push_back(iter->first);
            m.pose_id.push_back(static_cast<int64_t>(iter->first));
            m.pose_fixed.push_back(iter->first == _rootId ? 1 : 0);       // :111
            const Eigen::Vector3d & t = cameraPose.translation();
            const double rec[7] = {t[0], t[1], t[2], q[0], q[1], q[2], q[3]};
            m.pose_tq.append(rec, rec + 7);
        }
    }
    const std::size_t nPose = poseKey.size();

    const Eigen::Matrix3d K = cameraModel.eigenKdouble();                 // :176
    double baseLine = 0.0;
    if (_cameraModels.size() > 1) baseLine = cameraModel.getBaseLine();   // :181-183 (float -> double)
    m.fx = K(0, 0); m.fy = K(1, 1); m.cx = K(0, 2); m.cy = K(1, 2);
    m.bf = baseLine * m.fx;                                               // :195
    const double bfx = baseLine * K(0, 0);

    auto pit = _points3D.begin();
    for (auto iter = _wordReferences.begin(); iter != _wordReferences.end(); ++iter) {   // :156
        const std::size_t id = iter->first;
        while (pit != _points3D.end() && pit->first < id) ++pit;          // _points3D.find(id), both maps ascend
        if (pit == _points3D.end() || pit->first != id) continue;         // :158
        const int pointIndex = static_cast<int>(m.point_id.size());
        const Eigen::Vector3d & pointPose = std::get<0>(pit->second);
        m.point_id.push_back(static_cast<int64_t>(id));
        m.point_fixed.push_back(std::get<1>(pit->second) ? 1 : 0);        // :165
        m.point_xyz.push_back(pointPose[0]); m.point_xyz.push_back(pointPose[1]); m.point_xyz.push_back(pointPose[2]);
        std::size_t pk = 0;
        for (auto jter = iter->second.begin(); jter != iter->second.end(); ++jter) {     // :169
            while (pk < nPose && poseKey[pk] < jter->first) ++pk;         // uContains(_poses, jter->first)
            if (pk == nPose) break;                                       // (every later signature id is larger still)
            if (poseKey[pk] != jter->first) continue;                     // :172
            const FeatureBA & pt = jter->second;
            const double depth = pt.depth;                                // :174
            float obs[3] = {pt.kpt.pt.x, pt.kpt.pt.y, 0.0f};
            uint8_t kind = VISFS_BA_EDGE_MONO;
            if (std::isfinite(depth) && depth > 0.0 && baseLine > 0.0) {  // :184
                const float disparity = static_cast<float>(bfx / depth);  // :187
                obs[2] = pt.kpt.pt.x - disparity;                         // :188  (float - float; the device widens it)
                kind = VISFS_BA_EDGE_STEREO;
            }
            // else: the reference's mono branch is commented out (:197-208) and its live code is undefined
            // behaviour; this build defines the mono edge as rows 0-1 of EdgeStereo.
            m.edge_obs.append(obs, obs + 3);
            m.edge_pose.push_back(static_cast<int32_t>(pk));
            m.edge_point.push_back(pointIndex);
            m.edge_kind.push_back(kind);
        }
    }
    return true;
}

std::map<std::size_t, Eigen::Isometry3d> Optimizer::localOptimize(
    std::size_t _rootId,
    const std::map<std::size_t, Eigen::Isometry3d> & _poses,
    const std::map<std::size_t,std::tuple<std::size_t, std::size_t, Eigen::Isometry3d>> & _links,
    const std::vector<std::shared_ptr<GeometricCamera>> & _cameraModels,
    std::map<std::size_t, std::tuple<Eigen::Vector3d, bool>> & _points3D,
    const std::map<std::size_t, std::map<std::size_t, FeatureBA>> & _wordReferences,
    const std::vector<Sensor::PointCloud> & _pointClouds,
    const std::shared_ptr<const Map::Submap2D> & _submap,
    std::vector<std::tuple<std::size_t, std::size_t>> & _outliers) {

    std::map<std::size_t, Eigen::Isometry3d> optimizedPoses;
    message_.clear();
    if (_cameraModels.empty()) {   // the reference asserts (:69)
        message_ = "Optimizer: no camera model.";
        VISFS_B200_LOG("ERROR", message_);
        return optimizedPoses;
    }
    if (framework_ == CERES) {
        message_ = "Optimizer: Optimizer/Framework=1 (ceres) is not part of the CUDA build.";
        VISFS_B200_LOG("ERROR", message_);
        return optimizedPoses;
    }

    if (_poses.size() >= 2 && iterations_ > 0 && _poses.begin()->first > 0) {           // :74
        if (!_pointClouds.empty() && _submap != nullptr) {   // :225-258, laser edges: SURVEY.md §8 f-4
            message_ = "Optimizer: laser observations are not handled by the CUDA optimiser.";
            VISFS_B200_LOG("ERROR", message_);
            return optimizedPoses;
        }
        if (trustRegion_ != 0 && trustRegion_ != 1) {        // :93-97 leaves the algorithm unset
            message_ = "Optimizer: unknown Optimizer/TrustRegion.";
            VISFS_B200_LOG("ERROR", message_);
            return optimizedPoses;
        }
        if (!ensureHandle()) return optimizedPoses;

        detail::MarshalledWindow & m = window_;
        if (!marshal(_rootId, _poses, _cameraModels, _points3D, _wordReferences, m)) return optimizedPoses;

        visfs_ba_problem prob{};
        prob.n_poses = static_cast<int32_t>(m.pose_id.size());
        prob.n_points = static_cast<int32_t>(m.point_id.size());
        prob.n_edges = static_cast<int32_t>(m.edge_pose.size());
        prob.pose_tq = m.pose_tq.data(); prob.pose_id = m.pose_id.data(); prob.pose_fixed = m.pose_fixed.data();
        prob.point_xyz = m.point_xyz.data(); prob.point_id = m.point_id.data(); prob.point_fixed = m.point_fixed.data();
        prob.edge_obs = nullptr; prob.edge_obs_f32 = m.edge_obs.data();   // 12 instead of 24 bytes per edge over PCIe, nothing lost
        prob.edge_pose = m.edge_pose.data(); prob.edge_point = m.edge_point.data();
        prob.edge_kind = m.edge_kind.data();
        prob.fx = m.fx; prob.fy = m.fy; prob.cx = m.cx; prob.cy = m.cy; prob.bf = m.bf;
        prob.pixel_variance = pixelVariance_;                   // :153
        prob.huber_delta = robustKernelDelta_;                  // :212-216
        prob.iterations = iterations_;                          // :265, :311 (each pass runs iterations_/2)
        prob.solver = solver_;                                  // :76-91
        prob.trust_region = trustRegion_;                       // :93-97

        // :116-150 — odometry constraints (EdgePoseConstraint) between poses of the window
        std::vector<int32_t> linkFrom, linkTo;
        std::vector<double> linkTq;
        {
            const Eigen::Isometry3d Tri = _cameraModels.front()->getTansformImageToRobot();
            const Eigen::Isometry3d TriInv = Tri.inverse();
            auto indexOf = [&](std::size_t id) -> int {
                auto it = std::lower_bound(m.pose_id.begin(), m.pose_id.end(), static_cast<int64_t>(id));
                return (it != m.pose_id.end() && *it == static_cast<int64_t>(id)) ? static_cast<int>(it - m.pose_id.begin()) : -1;
            };
            for (auto iter = _links.begin(); iter != _links.end(); ++iter) {
                const std::size_t fromId = std::get<0>(iter->second), toId = std::get<1>(iter->second);
                if (!(fromId > 0 && toId > 0) || fromId == toId) continue;               // :126-129
                const int a = indexOf(fromId), b = indexOf(toId);
                if (a < 0 || b < 0) continue;                                            // uContains(_poses, ...)
                const Eigen::Isometry3d Tc1c2 = TriInv * std::get<2>(iter->second) * Tri;   // :133
                double q[4];
                rotationToQuaternion(Tc1c2.linear(), q);                                 // g2o::SE3Quat(R, t): w >= 0, unit
                const Eigen::Vector3d & t = Tc1c2.translation();
                const double rec[7] = {t[0], t[1], t[2], q[0], q[1], q[2], q[3]};
                linkFrom.push_back(a); linkTo.push_back(b);
                linkTq.insert(linkTq.end(), rec, rec + 7);
            }
        }
        prob.n_links = static_cast<int32_t>(linkFrom.size());
        prob.link_from = linkFrom.data(); prob.link_to = linkTo.data(); prob.link_tq = linkTq.data();
        prob.odometry_variance = odometryCovariance_;           // :120

        detail::HostArray<double> & poseOut = poseOut_;
        detail::HostArray<double> & pointOut = pointOut_;
        detail::HostArray<uint8_t> & levelOut = levelOut_;
        poseOut.resize(m.pose_tq.size()); pointOut.resize(m.point_xyz.size()); levelOut.resize(m.edge_pose.size());
        visfs_ba_result res{};
        res.pose_tq = poseOut.data(); res.point_xyz = pointOut.data(); res.edge_level = levelOut.data();
        const int status = visfs_ba_solve(handle_, &prob, &res);

        if (status == VISFS_BA_ERR_NUMERIC_PASS1) {             // :272-280
            message_ = std::isnan(res.chi2_pass1) ? "Optimization generated NANs, aborting optimization!"
                                                  : "g2o: Large optimization error detected in the first time optimize, aborting optimization!";
            VISFS_B200_LOG("ERROR", message_);
            return optimizedPoses;
        }
        if (status != VISFS_BA_OK && status != VISFS_BA_ERR_NUMERIC_PASS2) {   // CUDA / argument failure -> "BA failed"
            const char * why = visfs_ba_last_error(handle_);
            message_ = std::string("Optimizer: CUDA bundle adjustment failed: ") + (why ? why : "");
            VISFS_B200_LOG("ERROR", message_);
            return optimizedPoses;
        }

        // :283-309 — outliers are appended in edge insertion order, before the second-pass guard
        if (robustKernelDelta_ > 0.0) {
            for (std::size_t e = 0; e < levelOut.size(); ++e) {
                if (levelOut[e]) {
                    _outliers.emplace_back(std::make_tuple(static_cast<std::size_t>(m.point_id[m.edge_point[e]]),
                                                           static_cast<std::size_t>(m.pose_id[m.edge_pose[e]])));
                }
            }
            if (_outliers.size() > levelOut.size() / 2) {        // :305-308 (warning only)
                message_ = "Optimizer: large outliers detect, outliers size: " + std::to_string(_outliers.size()) +
                           ", total edges: " + std::to_string(levelOut.size());
                VISFS_B200_LOG("WARN", message_);
            }
        }
        if (status == VISFS_BA_ERR_NUMERIC_PASS2) {             // :315-318
            message_ = "g2o: Large optimization error detected in the second time optimize, aborting optimization!";
            VISFS_B200_LOG("ERROR", message_);
            return optimizedPoses;
        }

        // :320-340 — poses back to T_world<-robot
        const Eigen::Isometry3d TrcInv = _cameraModels.front()->getTansformImageToRobot().inverse();
        for (std::size_t i = 0; i < m.pose_id.size(); ++i) {
            Eigen::Isometry3d t = stateToIsometry(&poseOut[7 * i]);
            t = t.inverse();
            t = t * TrcInv;
            if (isZeroTransform(t)) {
                message_ = "Optimized pose " + std::to_string(m.pose_id[i]) + " is null.";
                VISFS_B200_LOG("WARN", message_);
                optimizedPoses.clear();
                return optimizedPoses;
            }
            optimizedPoses.emplace(static_cast<std::size_t>(m.pose_id[i]), t);
        }

        // :343-358 — points: accept moves shorter than 5 m, NaN for points that never became a vertex
        // (point_id ascends like _points3D's keys: one merge walk instead of a lookup table)
        std::size_t l = 0;
        for (auto iter = _points3D.begin(); iter != _points3D.end(); ++iter) {
            while (l < m.point_id.size() && static_cast<std::size_t>(m.point_id[l]) < iter->first) ++l;
            const bool hasVertex = l < m.point_id.size() && static_cast<std::size_t>(m.point_id[l]) == iter->first;
            const Eigen::Vector3d oldPose = std::get<0>(iter->second);
            const bool fixSymbol = std::get<1>(iter->second);
            if (hasVertex) {
                const double * np = &pointOut[3 * l];
                const double dx = oldPose[0] - np[0], dy = oldPose[1] - np[1], dz = oldPose[2] - np[2];
                if (std::sqrt(dx * dx + dy * dy + dz * dz) < 5.0) {
                    iter->second = std::make_tuple(Eigen::Vector3d(np[0], np[1], np[2]), fixSymbol);
                }
            } else {
                const double nan = std::numeric_limits<double>::quiet_NaN();
                iter->second = std::make_tuple(Eigen::Vector3d(nan, nan, nan), fixSymbol);
            }
        }
    } else if (_poses.size() == 1 || iterations_ <= 0) {        // :360-361
        optimizedPoses = _poses;
    } else {                                                    // :362-364
        message_ = "This method should be called at least with 1 pose!";
        VISFS_B200_LOG("ERROR", message_);
    }
    return optimizedPoses;
}

// ================================================================================================
// ResidentLocalMap: LocalMap's deltas -> visfs_ba_window_* (include/visfs_ba.h)
// ================================================================================================
ResidentLocalMap::ResidentLocalMap(const ParametersMap & _parameters, const std::vector<std::shared_ptr<GeometricCamera>> & _cameraModels,
                                   int _maxSignatures, int _maxFeatures) :
    handle_(nullptr), window_(nullptr), fx_(0.0), baseLine_(0.0), odometryCovariance_(Parameters::defaultOptimizerOdometryCovariance()),
    maxSignatures_(_maxSignatures),
    maxObservations_(4 * _maxSignatures * _maxFeatures) {
    // the same parameters, with the same defaults, as Optimizer (corelib/src/Optimizer/Optimizer.cpp:37-56)
    int solver = Parameters::defaultOptimizerSolver(), trust = Parameters::defaultOptimizerTrustRegion(), iterations = Parameters::defaultOptimizerIterations();
    double pixelVariance = Parameters::defaultOptimizerPixelVariance(), delta = Parameters::defaultOptimizerRobustKernelDelta();
    Parameters::parse(_parameters, Parameters::kOptimizerSolver(), solver);
    Parameters::parse(_parameters, Parameters::kOptimizerTrustRegion(), trust);
    Parameters::parse(_parameters, Parameters::kOptimizerIterations(), iterations);
    Parameters::parse(_parameters, Parameters::kOptimizerPixelVariance(), pixelVariance);
    Parameters::parse(_parameters, Parameters::kOptimizerRobustKernelDelta(), delta);
    Parameters::parse(_parameters, Parameters::kOptimizerOdometryCovariance(), odometryCovariance_);
    if (_cameraModels.empty() || !_cameraModels.front()) { fail("ResidentLocalMap: no camera model"); return; }
    const GeometricCamera & cameraModel = *_cameraModels.front();
    Trc_ = cameraModel.getTansformImageToRobot();
    const Eigen::Matrix3d K = cameraModel.eigenKdouble();                       // Optimizer.cpp:176
    if (_cameraModels.size() > 1) baseLine_ = cameraModel.getBaseLine();        // :181-183
    fx_ = K(0, 0);
    visfs_ba_config cfg{};
    cfg.abi_version = VISFS_BA_ABI_VERSION;
    cfg.device = g_device;
    if (visfs_ba_create(&cfg, &handle_) != VISFS_BA_OK) { handle_ = nullptr; fail("ResidentLocalMap: cannot create the CUDA bundle adjuster"); return; }
    visfs_ba_window_config wc{};
    wc.max_frames = _maxSignatures + 1; wc.max_points = _maxFeatures; wc.max_observations = maxObservations_;
    wc.fx = K(0, 0); wc.fy = K(1, 1); wc.cx = K(0, 2); wc.cy = K(1, 2); wc.bf = baseLine_ * wc.fx;   // :191-195
    wc.pixel_variance = pixelVariance; wc.huber_delta = delta;
    wc.iterations = iterations; wc.solver = solver; wc.trust_region = trust;
    if (visfs_ba_window_create(handle_, &wc, &window_) != VISFS_BA_OK) { window_ = nullptr; fail("ResidentLocalMap: cannot create the resident window"); }
}

ResidentLocalMap::~ResidentLocalMap() {
    if (window_) visfs_ba_window_destroy(window_);
    if (handle_) visfs_ba_destroy(handle_);
}

bool ResidentLocalMap::fail(const char * _what) {
    const char * why = handle_ ? visfs_ba_last_error(handle_) : visfs_ba_last_error(nullptr);
    message_ = std::string(_what) + ": " + (why ? why : "");
    VISFS_B200_LOG("ERROR", message_);
    return false;
}

namespace {
// Optimizer.cpp:102-109: T_cw = (T_wr * T_rc)^-1 as CameraPose(R, t)
void cameraState(const Eigen::Isometry3d & _Twr, const Eigen::Isometry3d & _Trc, double _tq[7]) {
    Eigen::Isometry3d cameraPose = _Twr * _Trc;
    cameraPose = cameraPose.inverse();
    double q[4];
    rotationToQuaternion(cameraPose.linear(), q);
    const Eigen::Vector3d & t = cameraPose.translation();
    _tq[0] = t[0]; _tq[1] = t[1]; _tq[2] = t[2]; _tq[3] = q[0]; _tq[4] = q[1]; _tq[5] = q[2]; _tq[6] = q[3];
}
}   // namespace

bool ResidentLocalMap::setFeature(std::size_t _featureId, const Eigen::Vector3d & _pose, bool _stable) {
    if (!window_) return false;
    const int64_t id = static_cast<int64_t>(_featureId);
    const double xyz[3] = {_pose[0], _pose[1], _pose[2]};
    const uint8_t fixed = _stable ? 1 : 0;                                      // LocalMap.cpp:278: fixed iff STABLE
    return visfs_ba_window_set_points(window_, 1, &id, xyz, &fixed) == VISFS_BA_OK || fail("ResidentLocalMap::setFeature");
}

bool ResidentLocalMap::insertSignature(std::size_t _id, const Eigen::Isometry3d & _pose, const std::map<std::size_t, FeatureBA> & _observations) {
    if (!window_) return false;
    double tq[7];
    cameraState(_pose, Trc_, tq);
    ids_.clear(); obs_.clear(); kinds_.clear();
    for (auto iter = _observations.begin(); iter != _observations.end(); ++iter) {
        const FeatureBA & pt = iter->second;
        const double depth = pt.depth;                                          // Optimizer.cpp:174
        float obs[3] = {pt.kpt.pt.x, pt.kpt.pt.y, 0.0f};
        uint8_t kind = VISFS_BA_EDGE_MONO;
        if (std::isfinite(depth) && depth > 0.0 && baseLine_ > 0.0) {          // :184
            const float disparity = static_cast<float>(baseLine_ * fx_ / depth);   // :187
            obs[2] = pt.kpt.pt.x - disparity;                                   // :188
            kind = VISFS_BA_EDGE_STEREO;
        }
        ids_.push_back(static_cast<int64_t>(iter->first));
        obs_.insert(obs_.end(), obs, obs + 3);
        kinds_.push_back(kind);
    }
    return visfs_ba_window_insert_frame(window_, static_cast<int64_t>(_id), tq, static_cast<int32_t>(ids_.size()), ids_.data(), obs_.data(),
                                        kinds_.data()) == VISFS_BA_OK || fail("ResidentLocalMap::insertSignature");
}

bool ResidentLocalMap::removeSignature(std::size_t _id) {
    return window_ && (visfs_ba_window_remove_frame(window_, static_cast<int64_t>(_id)) == VISFS_BA_OK || fail("ResidentLocalMap::removeSignature"));
}

bool ResidentLocalMap::removeFeature(std::size_t _featureId) {
    const int64_t id = static_cast<int64_t>(_featureId);
    return window_ && (visfs_ba_window_remove_points(window_, 1, &id) == VISFS_BA_OK || fail("ResidentLocalMap::removeFeature"));
}

bool ResidentLocalMap::removeObservation(std::size_t _featureId, std::size_t _signatureId) {
    const int64_t f = static_cast<int64_t>(_featureId), s = static_cast<int64_t>(_signatureId);
    return window_ && (visfs_ba_window_remove_observations(window_, 1, &f, &s) == VISFS_BA_OK || fail("ResidentLocalMap::removeObservation"));
}

bool ResidentLocalMap::setSignaturePose(std::size_t _id, const Eigen::Isometry3d & _pose) {
    if (!window_) return false;
    double tq[7];
    cameraState(_pose, Trc_, tq);
    const int64_t id = static_cast<int64_t>(_id);
    return visfs_ba_window_set_poses(window_, 1, &id, tq) == VISFS_BA_OK || fail("ResidentLocalMap::setSignaturePose");
}

bool ResidentLocalMap::setLinks(const std::map<std::size_t, std::tuple<std::size_t, std::size_t, Eigen::Isometry3d>> & _links) {
    if (!window_) return false;
    std::vector<int64_t> from, to;
    std::vector<double> tq;
    const Eigen::Isometry3d TriInv = Trc_.inverse();
    for (auto iter = _links.begin(); iter != _links.end(); ++iter) {
        const std::size_t fromId = std::get<0>(iter->second), toId = std::get<1>(iter->second);
        if (!(fromId > 0 && toId > 0) || fromId == toId) continue;                   // Optimizer.cpp:126-129
        const Eigen::Isometry3d Tc1c2 = TriInv * std::get<2>(iter->second) * Trc_;   // :133
        double q[4];
        rotationToQuaternion(Tc1c2.linear(), q);
        const Eigen::Vector3d & t = Tc1c2.translation();
        const double rec[7] = {t[0], t[1], t[2], q[0], q[1], q[2], q[3]};
        from.push_back(static_cast<int64_t>(fromId)); to.push_back(static_cast<int64_t>(toId));
        tq.insert(tq.end(), rec, rec + 7);
    }
    return visfs_ba_window_set_links(window_, static_cast<int32_t>(from.size()), from.data(), to.data(), tq.data(), odometryCovariance_) == VISFS_BA_OK ||
           fail("ResidentLocalMap::setLinks");
}

bool ResidentLocalMap::getFeaturePose(std::size_t _featureId, Eigen::Vector3d & _pose) {
    if (!window_) return false;
    const int64_t id = static_cast<int64_t>(_featureId);
    double xyz[3];
    if (visfs_ba_window_get_points(window_, 1, &id, xyz) != VISFS_BA_OK) return fail("ResidentLocalMap::getFeaturePose");
    _pose = Eigen::Vector3d(xyz[0], xyz[1], xyz[2]);
    return true;
}

std::map<std::size_t, Eigen::Isometry3d> ResidentLocalMap::localOptimize(std::size_t _rootId, std::vector<std::tuple<std::size_t, std::size_t>> & _outliers) {
    std::map<std::size_t, Eigen::Isometry3d> optimizedPoses;
    if (!window_) return optimizedPoses;
    const std::size_t F = static_cast<std::size_t>(maxSignatures_ + 1);
    ids_.assign(F, 0); poses_.assign(7 * F, 0.0);
    outlierPoint_.resize(static_cast<std::size_t>(maxObservations_)); outlierFrame_.resize(static_cast<std::size_t>(maxObservations_));
    visfs_ba_window_result res{};
    res.frame_id = ids_.data(); res.pose_tq = poses_.data();
    res.outlier_point_id = outlierPoint_.data(); res.outlier_frame_id = outlierFrame_.data(); res.outlier_capacity = maxObservations_;
    const int st = visfs_ba_window_solve(window_, static_cast<int64_t>(_rootId), &res);
    // the culled observations are reported before the second-pass guard, like the reference (Optimizer.cpp:283-318)
    if (st == VISFS_BA_OK || st == VISFS_BA_ERR_NUMERIC_PASS2)
        for (int k = 0; k < std::min(res.n_outliers, res.outlier_capacity); ++k)
            _outliers.emplace_back(static_cast<std::size_t>(outlierPoint_[static_cast<std::size_t>(k)]), static_cast<std::size_t>(outlierFrame_[static_cast<std::size_t>(k)]));
    if (st != VISFS_BA_OK) { fail("ResidentLocalMap::localOptimize"); return optimizedPoses; }
    const Eigen::Isometry3d Tcr = Trc_.inverse();
    for (int k = 0; k < res.n_frames; ++k) {                                    // Optimizer.cpp:320-340
        Eigen::Isometry3d t = stateToIsometry(poses_.data() + 7 * static_cast<std::size_t>(k));
        t = t.inverse();
        t = t * Tcr;
        if (isZeroTransform(t)) { optimizedPoses.clear(); return optimizedPoses; }
        optimizedPoses.emplace(static_cast<std::size_t>(ids_[static_cast<std::size_t>(k)]), t);
    }
    return optimizedPoses;
}

}   // Optimizer
}   // VISFS
