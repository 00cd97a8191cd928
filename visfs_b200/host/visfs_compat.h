// visfs_compat.h — the few VISFS / Eigen / OpenCV types that appear in the signature of
// VISFS::Optimizer::Optimizer::localOptimize (corelib/include/Optimizer/Optimizer.h:46-56).
//
// Inside the VISFS tree compile with -DVISFS_B200_WITH_VISFS_HEADERS: the real headers are used and
// this file adds nothing.  Stand-alone (this repository: no Eigen, OpenCV, PCL or Boost in the image)
// the stand-ins below provide exactly the members Optimizer.cpp touches, with Eigen's semantics.
#pragma once

#ifdef VISFS_B200_WITH_VISFS_HEADERS
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <opencv2/core/core.hpp>
#include "CameraModels/GeometricCamera.h"
#include "Map/2d/Submap2D.h"
#include "Parameters.h"
#include "Sensor/PointCloud.h"
#else

#include <array>
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace Eigen {

struct Vector3d {
    double v[3]{0, 0, 0};
    Vector3d() = default;
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double &operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double x() const { return v[0]; }
    double y() const { return v[1]; }
    double z() const { return v[2]; }
};

struct Matrix3d {
    double m[3][3]{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static Matrix3d Identity() { return Matrix3d(); }
    double &operator()(int r, int c) { return m[r][c]; }
    double operator()(int r, int c) const { return m[r][c]; }
    Matrix3d operator*(const Matrix3d &o) const {
        Matrix3d r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j];
        return r;
    }
    Vector3d operator*(const Vector3d &p) const {
        return Vector3d(m[0][0] * p[0] + m[0][1] * p[1] + m[0][2] * p[2], m[1][0] * p[0] + m[1][1] * p[1] + m[1][2] * p[2],
                        m[2][0] * p[0] + m[2][1] * p[1] + m[2][2] * p[2]);
    }
    Matrix3d transpose() const {
        Matrix3d r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r.m[i][j] = m[j][i];
        return r;
    }
};

// rigid transform: p -> R p + t  (Eigen::Isometry3d restricted to what localOptimize uses)
struct Isometry3d {
    Matrix3d R;
    Vector3d t;
    static Isometry3d Identity() { return Isometry3d(); }
    const Matrix3d &linear() const { return R; }
    Matrix3d &linear() { return R; }
    const Vector3d &translation() const { return t; }
    Vector3d &translation() { return t; }
    Isometry3d operator*(const Isometry3d &o) const {
        Isometry3d r;
        r.R = R * o.R;
        const Vector3d rt = R * o.t;
        r.t = Vector3d(rt[0] + t[0], rt[1] + t[1], rt[2] + t[2]);
        return r;
    }
    Isometry3d inverse() const {
        Isometry3d r;
        r.R = R.transpose();
        const Vector3d rt = r.R * t;
        r.t = Vector3d(-rt[0], -rt[1], -rt[2]);
        return r;
    }
    void prerotate(const Matrix3d &rot) {
        R = rot * R;
        t = rot * t;
    }
    bool isZero() const {
        for (int i = 0; i < 3; ++i) {
            if (t[i] != 0.0) return false;
            for (int j = 0; j < 3; ++j)
                if (R(i, j) != 0.0) return false;
        }
        return true;
    }
};

}  // namespace Eigen

namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint {
    Point2f pt;
    KeyPoint() = default;
    KeyPoint(float x, float y) { pt.x = x; pt.y = y; }
};
}  // namespace cv

namespace VISFS {

using ParametersMap = std::map<std::string, std::string>;

// corelib/include/CameraModels/GeometricCamera.h:12-58 (members localOptimize reads)
class GeometricCamera {
public:
    GeometricCamera() { init(); }
    explicit GeometricCamera(const std::vector<double> &parameters) : parameters_(parameters) { init(); }
    virtual ~GeometricCamera() {}
    virtual Eigen::Matrix3d eigenKdouble() const {
        Eigen::Matrix3d K;
        if (parameters_.size() >= 4) { K(0, 0) = parameters_[0]; K(1, 1) = parameters_[1]; K(0, 2) = parameters_[2]; K(1, 2) = parameters_[3]; }
        return K;
    }
    virtual float getBaseLine() const { return parameters_.size() > 4 ? static_cast<float>(parameters_[4]) : 0.f; }
    Eigen::Isometry3d getTansformImageToRobot() const { return tansformFromImageToRobot_; }

protected:
    void init() {
        Eigen::Matrix3d R;
        R(0, 0) = 0.0; R(0, 1) = 0.0; R(0, 2) = 1.0;
        R(1, 0) = -1.0; R(1, 1) = 0.0; R(1, 2) = 0.0;
        R(2, 0) = 0.0; R(2, 1) = -1.0; R(2, 2) = 0.0;
        tansformFromImageToRobot_.prerotate(R);
    }
    std::vector<double> parameters_;  // fx, fy, cx, cy, baseline
    Eigen::Isometry3d tansformFromImageToRobot_;
};

namespace Sensor { class PointCloud {}; }      // laser inputs: passed through, unused by the visual BA
namespace Map { class Submap2D {}; }

// the Optimizer/* keys and defaults of corelib/include/Parameters.h:184-191
struct Parameters {
    static std::string kOptimizerFramework() { return "Optimizer/Framework"; }
    static std::string kOptimizerSolver() { return "Optimizer/Solver"; }
    static std::string kOptimizerTrustRegion() { return "Optimizer/TrustRegion"; }
    static std::string kOptimizerIterations() { return "Optimizer/Iterations"; }
    static std::string kOptimizerPixelVariance() { return "Optimizer/PixelVariance"; }
    static std::string kOptimizerOdometryCovariance() { return "Optimizer/OdometryCovariance"; }
    static std::string kOptimizerLaserCovariance() { return "Optimizer/LaserCovariance"; }
    static std::string kOptimizerRobustKernelDelta() { return "Optimizer/RobustKernelDelta"; }
    static int defaultOptimizerFramework() { return 0; }
    static int defaultOptimizerSolver() { return 0; }
    static int defaultOptimizerTrustRegion() { return 0; }
    static int defaultOptimizerIterations() { return 10; }
    static double defaultOptimizerPixelVariance() { return 1.5; }
    static double defaultOptimizerOdometryCovariance() { return 0.00005; }
    static double defaultOptimizerLaserCovariance() { return 0.1; }
    static double defaultOptimizerRobustKernelDelta() { return 8.0; }
    static bool parse(const ParametersMap &p, const std::string &key, int &value) {
        auto it = p.find(key);
        if (it == p.end()) return false;
        value = std::stoi(it->second);
        return true;
    }
    static bool parse(const ParametersMap &p, const std::string &key, double &value) {
        auto it = p.find(key);
        if (it == p.end()) return false;
        value = std::stod(it->second);
        return true;
    }
};

}  // namespace VISFS
#endif  // VISFS_B200_WITH_VISFS_HEADERS
