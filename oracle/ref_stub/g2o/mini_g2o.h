// mini_g2o.h — TEST INFRASTRUCTURE ONLY.  The few g2o base classes that the reference's OptimizeTypeDefine.{h,cpp}
// derive from, reduced to the data members those files read and write (`_vertices`, `_measurement`, `_error`,
// `_jacobianOplusXi/Xj`, `_estimate`).  No optimiser, no solver: g2o's LM / BlockSolver semantics stay a restatement
// (oracle/ba_oracle.cpp); what this pins is the reference's OWN per-edge and per-vertex
// arithmetic, compiled from its unmodified sources (oracle/Makefile, target _ref/libvisfs_ref.so).
#ifndef VISFS_ORACLE_MINI_G2O_H
#define VISFS_ORACLE_MINI_G2O_H

#include <iostream>
#include <vector>
#include "../mini_eigen.h"

typedef double number_t;

namespace g2o {

typedef Eigen::Matrix<double, 3, 1> Vector3;
typedef Eigen::Matrix<double, 6, 1> Vector6;
typedef Eigen::Matrix<double, 7, 1> Vector7;
typedef Eigen::Matrix<double, 3, 3> Matrix3;
typedef Eigen::Quaternion<double> Quaternion;

class OptimizableGraph {
public:
    class Vertex {
    public:
        virtual ~Vertex() {}
        virtual void oplusImpl(const double *) = 0;
        virtual void setToOriginImpl() = 0;
    };
};

template <int D, typename T> class BaseVertex : public OptimizableGraph::Vertex {
public:
    typedef T EstimateType;
    static const int Dimension = D;
    const T &estimate() const { return _estimate; }
    void setEstimate(const T &e) { _estimate = e; updateCache(); }
    void updateCache() {}
    void oplus(const double *v) { oplusImpl(v); }
protected:
    T _estimate;
};

template <int D, typename E, typename VertexXi, typename VertexXj> class BaseBinaryEdge {
public:
    typedef E Measurement;
    typedef Eigen::Matrix<double, D, 1> ErrorVector;
    typedef Eigen::Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
    typedef Eigen::Matrix<double, D, VertexXj::Dimension> JacobianXjOplusType;
    BaseBinaryEdge() : _vertices(2, nullptr) {}
    virtual ~BaseBinaryEdge() {}
    virtual void computeError() = 0;
    virtual void linearizeOplus() = 0;
    void setVertex(size_t i, OptimizableGraph::Vertex *v) { _vertices[i] = v; }
    virtual void setMeasurement(const E &m) { _measurement = m; }
    const E &measurement() const { return _measurement; }
    const ErrorVector &error() const { return _error; }
    const JacobianXiOplusType &jacobianOplusXi() const { return _jacobianOplusXi; }
    const JacobianXjOplusType &jacobianOplusXj() const { return _jacobianOplusXj; }
protected:
    std::vector<OptimizableGraph::Vertex *> _vertices;
    E _measurement;
    ErrorVector _error;
    JacobianXiOplusType _jacobianOplusXi;
    JacobianXjOplusType _jacobianOplusXj;
};

// g2o::SE3Quat: rotation quaternion + translation; vector form (tx ty tz qx qy qz qw)
class SE3Quat {
public:
    SE3Quat() { _t.setZero(); _r.setIdentity(); }
    SE3Quat(const Quaternion &q, const Vector3 &t) : _r(q), _t(t) { normalizeRotation(); }
    SE3Quat(const Matrix3 &R, const Vector3 &t) : _r(Quaternion(R)), _t(t) { normalizeRotation(); }   // Optimizer.cpp:140
    const Vector3 &translation() const { return _t; }
    const Quaternion &rotation() const { return _r; }
    void normalizeRotation() {
        if (_r.w() < 0) _r.coeffs() *= -1;
        _r.normalize();
    }
    void fromVector(const Vector7 &v) { _r = Quaternion(v[6], v[3], v[4], v[5]); _t = Vector3(v[0], v[1], v[2]); }
    Vector7 toVector() const {
        Vector7 v;
        v[0] = _t[0]; v[1] = _t[1]; v[2] = _t[2]; v[3] = _r.x(); v[4] = _r.y(); v[5] = _r.z(); v[6] = _r.w();
        return v;
    }
    Eigen::Matrix<double, 4, 4> to_homogeneous_matrix() const {
        Eigen::Matrix<double, 4, 4> m;
        m.setIdentity();
        m.block(0, 0, 3, 3) = _r.toRotationMatrix();
        m.col(3).head(3) = _t;
        return m;
    }
protected:
    Quaternion _r;
    Vector3 _t;
};

// g2o::VertexPointXYZ (types_sba / slam3d): 3-dof point, oplus adds the increment
class VertexPointXYZ : public BaseVertex<3, Vector3> {
public:
    virtual void setToOriginImpl() { _estimate.setZero(); }
    virtual void oplusImpl(const double *u) { _estimate += Vector3(u[0], u[1], u[2]); }
};

}  // namespace g2o

#endif
