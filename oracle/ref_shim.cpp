// ref_shim.cpp — TEST INFRASTRUCTURE ONLY.  extern "C" entry points over the reference's OWN classes
// (VISFS::Optimizer::{CameraPose, VertexPose, EdgeStereo, EdgePoseConstraint}), compiled from the unmodified
// /root/reference sources against oracle/ref_stub (a minimal Eigen / g2o-base-class stand-in).  Built into
// oracle/_ref/libvisfs_ref.so by oracle/Makefile when /root/reference exists; used by tests/ to pin
//   * the C++ oracle (oracle/ba_oracle.cpp) and
//   * the CUDA path (visfs_ba_linearize, visfs_ba_debug_trial)
// to code the reference itself executes: rows a2, a3, a6, a7 and f-1 of SURVEY.md §8.  g2o's optimiser (LM,
// BlockSolver, Huber) is NOT here — those semantics stay a restatement.
//
// Objects are built exactly as corelib/src/Optimizer/Optimizer.cpp builds them:
//   pose vertex   : vCam->setEstimate(CameraPose(R, t))                       (Optimizer.cpp:106-109)
//   point vertex  : vpt3d->setEstimate(p)                                     (Optimizer.cpp:161-163)
//   stereo edge   : setMeasurement(obs); fx, fy, cx, cy, bf; vertex 0 = point, vertex 1 = pose   (186-211)
//   odometry edge : setMeasurement(g2o::SE3Quat(R, t)); vertex 0 = from, vertex 1 = to            (135-141)
#include "Optimizer/g2o/OptimizeTypeDefine.h"
#include "Math.h"

using namespace VISFS::Optimizer;

namespace {

VertexPose *make_pose(const double *tq) {   // tq = tx ty tz qx qy qz qw (CameraPose::toVector order)
    VertexPose *v = new VertexPose();
    v->setEstimate(CameraPose(Eigen::Quaterniond(tq[6], tq[3], tq[4], tq[5]), Eigen::Vector3d(tq[0], tq[1], tq[2])));
    return v;
}

void pose_out(const CameraPose &p, double *tq) {
    const Eigen::Matrix<double, 7, 1> v = p.toVector();
    for (int i = 0; i < 7; ++i) tq[i] = v[i];
}

}  // namespace

extern "C" {

int visfs_ref_version(void) { return 1; }

// CameraPose(R, t) as at Optimizer.cpp:109 (rotation matrix row-major in, tq out; w >= 0 forced, normalised)
void visfs_ref_pose_from_matrix(const double *R_rowmajor, const double *t, double *tq_out) {
    Eigen::Matrix3d R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R_rowmajor[3 * i + j];
    pose_out(CameraPose(R, Eigen::Vector3d(t[0], t[1], t[2])), tq_out);
}

// CameraPose(q, t): normalizeRotation of an arbitrary quaternion
void visfs_ref_pose_normalize(const double *tq_in, double *tq_out) {
    VertexPose *v = make_pose(tq_in);
    pose_out(v->estimate(), tq_out);
    delete v;
}

// CameraPose::map (OptimizeTypeDefine.h:45-47) and toHomogeneousMatrix (74-81, row-major out)
void visfs_ref_pose_map(const double *tq, const double *pw, double *pc_out, double *T_rowmajor_out) {
    VertexPose *v = make_pose(tq);
    const Eigen::Vector3d pc = v->estimate().map(Eigen::Vector3d(pw[0], pw[1], pw[2]));
    for (int i = 0; i < 3; ++i) pc_out[i] = pc[i];
    if (T_rowmajor_out) {
        const Eigen::Matrix<double, 4, 4> T = v->estimate().toHomogeneousMatrix();
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T_rowmajor_out[4 * i + j] = T(i, j);
    }
    delete v;
}

// VertexPose::oplusImpl -> CameraPose::update (OptimizeTypeDefine.cpp:7-14) + deltaQ (Math.h:277-287)
void visfs_ref_pose_oplus(const double *tq_in, const double *delta6, double *tq_out) {
    VertexPose *v = make_pose(tq_in);
    v->oplusImpl(delta6);
    pose_out(v->estimate(), tq_out);
    delete v;
}

// EdgeStereo::computeError + linearizeOplus (OptimizeTypeDefine.h:121-187) for n edges.
// pose_tq [n][7], point [n][3], obs [n][3], intr = fx fy cx cy bf; outputs row-major, any may be NULL.
void visfs_ref_edge_stereo(int n, const double *pose_tq, const double *point, const double *obs, const double *intr,
                           double *err, double *J_point, double *J_pose, int *depth_positive) {
    for (int e = 0; e < n; ++e) {
        VertexPose *vp = make_pose(pose_tq + 7 * e);
        g2o::VertexPointXYZ *vl = new g2o::VertexPointXYZ();
        vl->setEstimate(Eigen::Vector3d(point[3 * e], point[3 * e + 1], point[3 * e + 2]));
        EdgeStereo *es = new EdgeStereo();
        es->setMeasurement(Eigen::Vector3d(obs[3 * e], obs[3 * e + 1], obs[3 * e + 2]));
        es->fx = intr[0]; es->fy = intr[1]; es->cx = intr[2]; es->cy = intr[3]; es->bf = intr[4];
        es->setVertex(0, vl);
        es->setVertex(1, vp);
        es->computeError();
        es->linearizeOplus();
        if (err) for (int i = 0; i < 3; ++i) err[3 * e + i] = es->error()[i];
        if (J_point) for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J_point[9 * e + 3 * i + j] = es->jacobianOplusXi()(i, j);
        if (J_pose) for (int i = 0; i < 3; ++i) for (int j = 0; j < 6; ++j) J_pose[18 * e + 6 * i + j] = es->jacobianOplusXj()(i, j);
        if (depth_positive) depth_positive[e] = es->isDepthPositive() ? 1 : 0;
        delete es; delete vl; delete vp;
    }
}

// EdgePoseConstraint::computeError + linearizeOplus (OptimizeTypeDefine.cpp:35-72) for n links.
// from_tq / to_tq / meas_tq [n][7]; the measurement goes through g2o::SE3Quat like Optimizer.cpp:140.
void visfs_ref_edge_pose_constraint(int n, const double *from_tq, const double *to_tq, const double *meas_tq,
                                    double *err, double *J_from, double *J_to) {
    for (int k = 0; k < n; ++k) {
        VertexPose *v1 = make_pose(from_tq + 7 * k), *v2 = make_pose(to_tq + 7 * k);
        const double *m = meas_tq + 7 * k;
        EdgePoseConstraint *ep = new EdgePoseConstraint();
        ep->setVertex(0, v1);
        ep->setVertex(1, v2);
        ep->setMeasurement(g2o::SE3Quat(Eigen::Quaterniond(m[6], m[3], m[4], m[5]), Eigen::Vector3d(m[0], m[1], m[2])));
        ep->computeError();
        ep->linearizeOplus();
        if (err) for (int i = 0; i < 6; ++i) err[6 * k + i] = ep->error()[i];
        if (J_from) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) J_from[36 * k + 6 * i + j] = ep->jacobianOplusXi()(i, j);
        if (J_to) for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) J_to[36 * k + 6 * i + j] = ep->jacobianOplusXj()(i, j);
        delete ep; delete v1; delete v2;
    }
}

// g2o::VertexPointXYZ::oplusImpl (a5)
void visfs_ref_point_oplus(const double *p_in, const double *delta3, double *p_out) {
    g2o::VertexPointXYZ v;
    v.setEstimate(Eigen::Vector3d(p_in[0], p_in[1], p_in[2]));
    v.oplusImpl(delta3);
    for (int i = 0; i < 3; ++i) p_out[i] = v.estimate()[i];
}

// uNorm (Math.h:248-251), the write-back clamp's distance (Optimizer.cpp:349-350)
double visfs_ref_unorm3(double x, double y, double z) { return uNorm(x, y, z); }

}  // extern "C"
