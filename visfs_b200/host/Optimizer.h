// Optimizer.h — C++17 mirror of the reference's BA interface
// (corelib/include/Optimizer/Optimizer.h:17-77): same namespace, class, struct, constructor and
// localOptimize signature, so corelib/src/Estimator.cpp:254 links against it unchanged.  The body
// (Optimizer.cpp) marshals the window into the C ABI of include/visfs_ba.h and runs the sm_100a
// kernels; there is no g2o, no Ceres and no CPU fallback behind it.
#ifndef VISFS_OPTIMIZER_H
#define VISFS_OPTIMIZER_H

#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "visfs_compat.h"

struct visfs_ba_handle;
struct visfs_ba_window;

namespace VISFS {
namespace Optimizer {

struct FeatureBA {
    cv::KeyPoint kpt;
    float depth;
    FeatureBA(const cv::KeyPoint & _kpt, const float _depth) {
        kpt = _kpt;
        depth = _depth;
    }
};

namespace detail {
// Growable host array in page-locked memory (visfs_ba_host_alloc), so that the C ABI moves it by DMA without a staging
// copy; plain malloc when there is no CUDA device (the marshalling itself needs none).  Kept across calls by Optimizer.
void * hostAlloc(std::size_t _bytes, bool * _pinned);
void hostFree(void * _p, bool _pinned);

template <typename T>
class HostArray {
public:
    HostArray() = default;
    ~HostArray() { hostFree(p_, pinned_); }
    HostArray(const HostArray &) = delete;
    HostArray & operator=(const HostArray &) = delete;

    void clear() { n_ = 0; }
    std::size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    T * data() { return p_; }
    const T * data() const { return p_; }
    T & operator[](std::size_t _i) { return p_[_i]; }
    const T & operator[](std::size_t _i) const { return p_[_i]; }
    const T * begin() const { return p_; }
    const T * end() const { return p_ + n_; }
    void push_back(const T & _v) { if (n_ == cap_) grow(n_ + 1); p_[n_++] = _v; }
    void append(const T * _b, const T * _e) {
        const std::size_t k = static_cast<std::size_t>(_e - _b);
        if (n_ + k > cap_) grow(n_ + k);
        for (std::size_t i = 0; i < k; ++i) p_[n_ + i] = _b[i];
        n_ += k;
    }
    void resize(std::size_t _n) { if (_n > cap_) grow(_n); n_ = _n; }

private:
    void grow(std::size_t _need) {
        std::size_t cap = cap_ ? 2 * cap_ : 4096;
        if (cap < _need) cap = _need;
        bool pinned = false;
        T * q = static_cast<T *>(hostAlloc(cap * sizeof(T), &pinned));
        for (std::size_t i = 0; i < n_; ++i) q[i] = p_[i];
        hostFree(p_, pinned_);
        p_ = q; cap_ = cap; pinned_ = pinned;
    }
    T * p_ = nullptr;
    std::size_t n_ = 0, cap_ = 0;
    bool pinned_ = false;
};

// The window in the flat form the C ABI takes (host memory), plus the id tables needed to map results back.
struct MarshalledWindow {
    HostArray<double> pose_tq;          // [P][7] T_cw
    HostArray<int64_t> pose_id;
    HostArray<uint8_t> pose_fixed;
    HostArray<double> point_xyz;        // [L][3]
    HostArray<int64_t> point_id;        // feature ids that got a vertex (Optimizer.cpp:158), ascending
    HostArray<uint8_t> point_fixed;
    HostArray<float> edge_obs;          // [E][3] u, v, u_right: floats, as the reference builds them (Optimizer.cpp:187-188)
    HostArray<int32_t> edge_pose, edge_point;
    HostArray<uint8_t> edge_kind;
    double fx = 0, fy = 0, cx = 0, cy = 0, bf = 0;
    void clear() {
        pose_tq.clear(); pose_id.clear(); pose_fixed.clear(); point_xyz.clear(); point_id.clear(); point_fixed.clear();
        edge_obs.clear(); edge_pose.clear(); edge_point.clear(); edge_kind.clear();
        fx = fy = cx = cy = bf = 0;
    }
};
}  // namespace detail

class Optimizer {
public:
    Optimizer(const ParametersMap & _parameters = ParametersMap());
    ~Optimizer();
    Optimizer(const Optimizer &) = delete;
    Optimizer & operator=(const Optimizer &) = delete;

    /** \brief Optimize the local maps (visual part of the reference's local fusion).
     * \param[in] rootId Fixed pose.
     * \param[in] poses Poses to optimize (T_world<-robot).
     * \param[in] links Odometry links between poses (EdgePoseConstraint, Optimizer.cpp:116-150).
     * \param[in] cameraModels Left (and right) camera model; stereo iff size() > 1.
     * \param[in&out] points3D World points, fixed flag.
     * \param[in] wordReferences Observations: feature id -> pose id -> key point + depth.
     * \param[in] pointClouds Laser points (must be empty: SURVEY.md §8 f-4).
     * \param[in] submap Laser submap (must be null).
     * \param[out] outliers (feature id, pose id) of every edge culled after the first pass.
     * \return optimised poses; EMPTY on failure, exactly like the reference.
     */
    std::map<std::size_t, Eigen::Isometry3d> localOptimize(
        std::size_t _rootId,
        const std::map<std::size_t, Eigen::Isometry3d> & _poses,
        const std::map<std::size_t,std::tuple<std::size_t, std::size_t, Eigen::Isometry3d>> & _links,
        const std::vector<std::shared_ptr<GeometricCamera>> & _cameraModels,
        std::map<std::size_t, std::tuple<Eigen::Vector3d, bool>> & _points3D,
        const std::map<std::size_t, std::map<std::size_t, FeatureBA>> & _wordReferences,
        const std::vector<Sensor::PointCloud> & _pointClouds,
        const std::shared_ptr<const Map::Submap2D> & _submap,
        std::vector<std::tuple<std::size_t, std::size_t>> & _outliers
    );

    // Host-only half of localOptimize (Optimizer.cpp:100-223 of the reference): map-of-maps -> flat arrays.
    // Public so that it can be tested without a GPU.
    static bool marshal(std::size_t _rootId,
                        const std::map<std::size_t, Eigen::Isometry3d> & _poses,
                        const std::vector<std::shared_ptr<GeometricCamera>> & _cameraModels,
                        const std::map<std::size_t, std::tuple<Eigen::Vector3d, bool>> & _points3D,
                        const std::map<std::size_t, std::map<std::size_t, FeatureBA>> & _wordReferences,
                        detail::MarshalledWindow & _out);

    // Device ordinal used by every Optimizer created afterwards (default 0).
    static void setDevice(int _device);

    // Last message a LOG_ERROR / LOG_WARN of the reference would have carried.
    const std::string & lastMessage() const { return message_; }

private:
    enum Framework {
        G2O     =   0,      // the g2o-semantics LM, executed by the CUDA kernels
        CERES   =   1,      // not provided by this build
        CUDA    =   2       // alias of 0
    };

    bool ensureHandle();

    int framework_;
    int solver_;
    int trustRegion_;
    int iterations_;
    double pixelVariance_;
    double odometryCovariance_;
    double laserCovariance_;
    double robustKernelDelta_;

    visfs_ba_handle * handle_;
    std::string message_;

    // marshalling and result buffers, reused from call to call (results never depend on earlier calls)
    detail::MarshalledWindow window_;
    detail::HostArray<double> poseOut_, pointOut_;
    detail::HostArray<uint8_t> levelOut_;
};

/** \brief The bundle-adjustment side of VISFS::LocalMap (corelib/src/LocalMap.cpp) kept resident on the GPU.
 *
 * The reference rebuilds poses / points3D / wordReferences from LocalMap on every frame (Estimator.cpp:216-254) although the
 * map changed by one signature.  This class takes the same changes as LocalMap does, as deltas:
 *   LocalMap::insertSignature  (LocalMap.cpp:48-131)  ->  setFeature() for the features it creates, insertSignature()
 *   LocalMap::removeSignature  (LocalMap.cpp:133-168) ->  removeSignature(), removeFeature() for the features it erases
 *   LocalMap::updateLocalMap   (LocalMap.cpp:170-226) ->  nothing for poses / feature poses (the solve already wrote them into
 *                                                         the resident state), removeObservation() for every outlier,
 *                                                         setSignaturePose() where the caller overrides a pose (Estimator.cpp:393)
 * and localOptimize() replaces getSignaturePoses + getFeaturePosesAndObservations + Optimizer::localOptimize.  Results are
 * those of Optimizer::localOptimize on the maps the reference would have built.  Odometry links and laser data are not part
 * of the resident state yet (use Optimizer::localOptimize for those configurations).
 */
class ResidentLocalMap {
public:
    ResidentLocalMap(const ParametersMap & _parameters, const std::vector<std::shared_ptr<GeometricCamera>> & _cameraModels,
                     int _maxSignatures = 6, int _maxFeatures = 4096);
    ~ResidentLocalMap();
    ResidentLocalMap(const ResidentLocalMap &) = delete;
    ResidentLocalMap & operator=(const ResidentLocalMap &) = delete;

    bool setFeature(std::size_t _featureId, const Eigen::Vector3d & _pose, bool _stable);
    bool insertSignature(std::size_t _id, const Eigen::Isometry3d & _pose, const std::map<std::size_t, FeatureBA> & _observations);
    bool removeSignature(std::size_t _id);
    bool removeFeature(std::size_t _featureId);
    bool removeObservation(std::size_t _featureId, std::size_t _signatureId);
    bool setSignaturePose(std::size_t _id, const Eigen::Isometry3d & _pose);
    /** \brief LocalMap::getSignatureLinks' result (<link id, <from, to, T_r1r2>>): replaces the odometry links of the map.
      * A link takes part in a solve when both its signatures are in the map then (Optimizer.cpp:126-131). */
    bool setLinks(const std::map<std::size_t, std::tuple<std::size_t, std::size_t, Eigen::Isometry3d>> & _links);
    bool getFeaturePose(std::size_t _featureId, Eigen::Vector3d & _pose);

    /** \brief Optimizer::localOptimize on the resident map.  \return optimised poses (T_world<-robot); EMPTY on failure. */
    std::map<std::size_t, Eigen::Isometry3d> localOptimize(std::size_t _rootId, std::vector<std::tuple<std::size_t, std::size_t>> & _outliers);

    const std::string & lastMessage() const { return message_; }

private:
    bool fail(const char * _what);
    visfs_ba_handle * handle_;
    visfs_ba_window * window_;
    Eigen::Isometry3d Trc_;
    double fx_, baseLine_, odometryCovariance_;
    int maxSignatures_, maxObservations_;
    std::string message_;
    std::vector<int64_t> ids_, outlierPoint_, outlierFrame_;
    std::vector<double> poses_;
    std::vector<float> obs_;
    std::vector<uint8_t> kinds_;
};

}   // Optimizer
}   // VISFS

#endif  // VISFS_OPTIMIZER_H
