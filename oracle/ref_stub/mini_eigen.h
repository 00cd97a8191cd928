// mini_eigen.h — TEST INFRASTRUCTURE ONLY (never part of the product, never linked into libvisfs_ba.so).
//
// The smallest stand-in for <Eigen/Core> + <Eigen/Geometry> that lets the reference's OWN, UNMODIFIED sources
//     /root/reference/corelib/include/Optimizer/g2o/OptimizeTypeDefine.h
//     /root/reference/corelib/src/Optimizer/g2o/OptimizeTypeDefine.cpp
//     /root/reference/utilite/include/Math.h
// compile in an image that has no Eigen (see oracle/Makefile, target _ref/libvisfs_ref.so).  Fixed-size dense
// matrices and quaternions with value semantics, no expression templates.  Only the members those three files use
// exist.  The arithmetic of the members follows Eigen's published definitions (column-major storage, Hamilton
// product, toRotationMatrix, Shepperd's rotation->quaternion, q*v as v + w*t + q.vec x t with t = 2 q.vec x v);
// rounding may differ from real Eigen in the last bits, which is far inside the 1e-9 parity gate.
#ifndef VISFS_ORACLE_MINI_EIGEN_H
#define VISFS_ORACLE_MINI_EIGEN_H

#include <cmath>
#include <cstddef>
#include <iostream>
#include <type_traits>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1, Dynamic = -1 };

template <typename T, int R, int C, int Opt = ColMajor> class Matrix;
template <typename T> class Quaternion;

template <typename Derived> struct traits;
template <typename T, int R, int C, int Opt> struct traits<Matrix<T, R, C, Opt>> {
    typedef T Scalar;
    enum { Rows = R, Cols = C };
};

// ---- proxies for writable sub-blocks ---------------------------------------------------------------------------
template <typename T> class BlockRef {   // run-time sized view into a matrix (strided)
public:
    BlockRef(T *p, int r, int c, int rs, int cs) : p_(p), r_(r), c_(c), rs_(rs), cs_(cs) {}
    template <int R, int C, int O> BlockRef &operator=(const Matrix<T, R, C, O> &m) {
        for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) p_[i * rs_ + j * cs_] = m(i, j);
        return *this;
    }
    BlockRef head(int n) { return BlockRef(p_, n, 1, rs_, cs_); }
    T &operator()(int i, int j) { return p_[i * rs_ + j * cs_]; }
private:
    T *p_; int r_, c_, rs_, cs_;
};

template <typename T, int BR, int BC> class FixedBlock {   // compile-time sized view
public:
    FixedBlock(T *p, int rs, int cs) : p_(p), rs_(rs), cs_(cs) {}
    template <int O> FixedBlock &operator=(const Matrix<T, BR, BC, O> &m) {
        for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) p_[i * rs_ + j * cs_] = m(i, j);
        return *this;
    }
    operator Matrix<T, BR, BC>() const {
        Matrix<T, BR, BC> m;
        for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) m(i, j) = p_[i * rs_ + j * cs_];
        return m;
    }
private:
    T *p_; int rs_, cs_;
};

template <typename T, int R, int C, int Opt> class CommaInit {
public:
    CommaInit(Matrix<T, R, C, Opt> &m, T first) : m_(m), k_(0) { put(first); }
    CommaInit &operator,(T v) { put(v); return *this; }
private:
    void put(T v) { m_(k_ / C, k_ % C) = v; ++k_; }   // the comma initialiser fills row by row
    Matrix<T, R, C, Opt> &m_; int k_;
};

// ---- MatrixBase: what function templates of Math.h deduce against ------------------------------------------------
template <typename Derived> class MatrixBase {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    const Derived &derived() const { return *static_cast<const Derived *>(this); }
    Derived &derived() { return *static_cast<Derived *>(this); }
    Scalar operator()(int i) const { return derived().coeff(i); }
    Scalar operator[](int i) const { return derived().coeff(i); }
    Scalar operator()(int i, int j) const { return derived().coeff(i, j); }
};

template <typename T, int R, int C, int Opt> class Matrix : public MatrixBase<Matrix<T, R, C, Opt>> {
public:
    typedef T Scalar;
    enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
    Matrix() { for (int i = 0; i < R * C; ++i) d_[i] = T(0); }   // (Eigen leaves it uninitialised; zero is a valid instance of that)
    Matrix(const Matrix &) = default;
    Matrix &operator=(const Matrix &) = default;
    template <int O2> Matrix(const Matrix<T, R, C, O2> &o) { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) (*this)(i, j) = o(i, j); }
    Matrix(const MatrixBase<Matrix> &o) { *this = o.derived(); }
    Matrix(T x, T y) { static_assert(R * C == 2, "2-vector"); d_[0] = x; d_[1] = y; }
    Matrix(T x, T y, T z) { static_assert(R * C == 3, "3-vector"); d_[0] = x; d_[1] = y; d_[2] = z; }
    Matrix(T x, T y, T z, T w) { static_assert(R * C == 4, "4-vector"); d_[0] = x; d_[1] = y; d_[2] = z; d_[3] = w; }

    static int idx(int i, int j) { return Opt == RowMajor ? i * C + j : j * R + i; }
    T &operator()(int i, int j) { return d_[idx(i, j)]; }
    T operator()(int i, int j) const { return d_[idx(i, j)]; }
    T coeff(int i, int j) const { return d_[idx(i, j)]; }
    T &operator()(int i) { static_assert(R == 1 || C == 1, "vector"); return d_[i]; }
    T operator()(int i) const { static_assert(R == 1 || C == 1, "vector"); return d_[i]; }
    T &operator[](int i) { static_assert(R == 1 || C == 1, "vector"); return d_[i]; }
    T operator[](int i) const { static_assert(R == 1 || C == 1, "vector"); return d_[i]; }
    T coeff(int i) const { return d_[i]; }
    T &x() { return d_[0]; } T &y() { return d_[1]; } T &z() { return d_[2]; } T &w() { return d_[3]; }
    T x() const { return d_[0]; } T y() const { return d_[1]; } T z() const { return d_[2]; } T w() const { return d_[3]; }
    T *data() { return d_; }
    const T *data() const { return d_; }
    int rows() const { return R; }
    int cols() const { return C; }

    Matrix &setZero() { for (int i = 0; i < R * C; ++i) d_[i] = T(0); return *this; }
    Matrix &setIdentity() { setZero(); for (int i = 0; i < (R < C ? R : C); ++i) (*this)(i, i) = T(1); return *this; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }

    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; ++i) d_[i] += o.d_[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; ++i) d_[i] -= o.d_[i]; return *this; }
    Matrix &operator*=(T s) { for (int i = 0; i < R * C; ++i) d_[i] *= s; return *this; }
    Matrix &operator/=(T s) { for (int i = 0; i < R * C; ++i) d_[i] /= s; return *this; }
    Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d_[i] = -d_[i]; return m; }
    Matrix operator+(const Matrix &o) const { Matrix m(*this); m += o; return m; }
    Matrix operator-(const Matrix &o) const { Matrix m(*this); m -= o; return m; }
    Matrix operator*(T s) const { Matrix m(*this); m *= s; return m; }
    Matrix operator/(T s) const { Matrix m(*this); m /= s; return m; }
    template <int C2, int O2> Matrix<T, R, C2> operator*(const Matrix<T, C, C2, O2> &o) const {
        Matrix<T, R, C2> m;
        for (int i = 0; i < R; ++i)
            for (int j = 0; j < C2; ++j) {
                T s = (*this)(i, 0) * o(0, j);
                for (int k = 1; k < C; ++k) s += (*this)(i, k) * o(k, j);
                m(i, j) = s;
            }
        return m;
    }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> m; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) m(j, i) = (*this)(i, j); return m; }
    const Matrix &matrix() const { return *this; }
    T squaredNorm() const { T s = T(0); for (int i = 0; i < R * C; ++i) s += d_[i] * d_[i]; return s; }
    T norm() const { using std::sqrt; return sqrt(squaredNorm()); }
    T dot(const Matrix &o) const { T s = T(0); for (int i = 0; i < R * C; ++i) s += d_[i] * o.d_[i]; return s; }
    Matrix cross(const Matrix &o) const {
        static_assert(R * C == 3, "3-vector");
        return Matrix(d_[1] * o.d_[2] - d_[2] * o.d_[1], d_[2] * o.d_[0] - d_[0] * o.d_[2], d_[0] * o.d_[1] - d_[1] * o.d_[0]);
    }

    // sub-blocks
    BlockRef<T> block(int i, int j, int r, int c) { return BlockRef<T>(&(*this)(i, j), r, c, idx(1, 0) - idx(0, 0), C > 1 ? idx(0, 1) - idx(0, 0) : 0); }
    BlockRef<T> col(int j) { return BlockRef<T>(&(*this)(0, j), R, 1, R > 1 ? idx(1, 0) - idx(0, 0) : 0, 0); }
    BlockRef<T> head(int n) { static_assert(C == 1, "column vector"); return BlockRef<T>(d_, n, 1, 1, 0); }
    template <int BR, int BC> FixedBlock<T, BR, BC> block(int i, int j) {
        return FixedBlock<T, BR, BC>(&(*this)(i, j), R > 1 ? idx(1, 0) - idx(0, 0) : 0, C > 1 ? idx(0, 1) - idx(0, 0) : 0);
    }
    template <int BR, int BC> Matrix<T, BR, BC> block(int i, int j) const {
        Matrix<T, BR, BC> m;
        for (int a = 0; a < BR; ++a) for (int b = 0; b < BC; ++b) m(a, b) = (*this)(i + a, j + b);
        return m;
    }
    template <int BR, int BC> Matrix<T, BR, BC> bottomRightCorner() const { return static_cast<const Matrix *>(this)->template block<BR, BC>(R - BR, C - BC); }

private:
    T d_[R * C];
};

template <typename S, typename T, int R, int C, int O, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
Matrix<T, R, C, O> operator*(S s, const Matrix<T, R, C, O> &m) { Matrix<T, R, C, O> r(m); r *= T(s); return r; }

template <typename T, int R, int C, int O> CommaInit<T, R, C, O> operator<<(Matrix<T, R, C, O> &m, T first) { return CommaInit<T, R, C, O>(m, first); }

template <typename T, int R, int C, int O> std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C, O> &m) {
    for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < R) os << "\n"; }
    return os;
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;

// ---- Map<M>: a pointer viewed as a fixed-size matrix -----------------------------------------------------------------
template <typename M> class Map {
    typedef typename std::remove_const<M>::type Plain;
    typedef typename Plain::Scalar T;
    typedef typename std::conditional<std::is_const<M>::value, const T, T>::type Elem;
public:
    explicit Map(Elem *p) : p_(p) {}
    operator Plain() const { Plain m; for (int i = 0; i < m.rows() * m.cols(); ++i) m.data()[i] = p_[i]; return m; }
    Map &operator=(const Plain &m) { for (int i = 0; i < m.rows() * m.cols(); ++i) p_[i] = m.data()[i]; return *this; }
    Elem &operator[](int i) const { return p_[i]; }
    Elem &operator()(int i) const { return p_[i]; }
private:
    Elem *p_;
};

// ---- quaternions (coefficients stored x, y, z, w like Eigen) -----------------------------------------------------------
template <typename T> struct traits<Quaternion<T>> { typedef T Scalar; };

template <typename Derived> class QuaternionBase {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    const Derived &derived() const { return *static_cast<const Derived *>(this); }
};

template <typename T> class Quaternion : public QuaternionBase<Quaternion<T>> {
public:
    typedef T Scalar;
    typedef Matrix<T, 3, 1> Vec3;
    typedef Matrix<T, 3, 3> Mat3;
    Quaternion() : c_(T(0), T(0), T(0), T(1)) {}   // (Eigen: uninitialised)
    Quaternion(T w, T x, T y, T z) : c_(x, y, z, w) {}
    Quaternion(const Quaternion &) = default;
    Quaternion &operator=(const Quaternion &) = default;
    Quaternion(const QuaternionBase<Quaternion> &o) : c_(o.derived().c_) {}
    Quaternion &operator=(const QuaternionBase<Quaternion> &o) { c_ = o.derived().c_; return *this; }
    explicit Quaternion(const Mat3 &m) { *this = m; }

    // rotation matrix -> quaternion: Eigen's quaternionbase_assign_impl (after Shoemake / Shepperd)
    Quaternion &operator=(const Mat3 &m) {
        using std::sqrt;
        T t = m(0, 0) + m(1, 1) + m(2, 2);
        if (t > T(0)) {
            t = sqrt(t + T(1.0));
            w() = T(0.5) * t;
            t = T(0.5) / t;
            x() = (m(2, 1) - m(1, 2)) * t;
            y() = (m(0, 2) - m(2, 0)) * t;
            z() = (m(1, 0) - m(0, 1)) * t;
        } else {
            int i = 0;
            if (m(1, 1) > m(0, 0)) i = 1;
            if (m(2, 2) > m(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
            c_[i] = T(0.5) * t;
            t = T(0.5) / t;
            w() = (m(k, j) - m(j, k)) * t;
            c_[j] = (m(j, i) + m(i, j)) * t;
            c_[k] = (m(k, i) + m(i, k)) * t;
        }
        return *this;
    }

    T &x() { return c_[0]; } T &y() { return c_[1]; } T &z() { return c_[2]; } T &w() { return c_[3]; }
    T x() const { return c_[0]; } T y() const { return c_[1]; } T z() const { return c_[2]; } T w() const { return c_[3]; }
    Matrix<T, 4, 1> &coeffs() { return c_; }
    const Matrix<T, 4, 1> &coeffs() const { return c_; }
    Vec3 vec() const { return Vec3(c_[0], c_[1], c_[2]); }
    Quaternion &setIdentity() { c_ = Matrix<T, 4, 1>(T(0), T(0), T(0), T(1)); return *this; }
    static Quaternion Identity() { return Quaternion(); }
    T squaredNorm() const { return c_.squaredNorm(); }
    T norm() const { return c_.norm(); }
    void normalize() { c_ /= c_.norm(); }
    Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Quaternion inverse() const {   // Eigen: conjugate / squaredNorm (zero quaternion -> zero)
        const T n2 = squaredNorm();
        if (n2 > T(0)) { Quaternion q = conjugate(); q.c_ /= n2; return q; }
        Quaternion q; q.c_.setZero(); return q;
    }
    Quaternion operator*(const Quaternion &b) const {   // Hamilton product
        const Quaternion &a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                          a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                          a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Vec3 operator*(const Vec3 &v) const {   // Eigen's _transformVector
        Vec3 uv = vec().cross(v);
        uv += uv;
        return v + uv * w() + vec().cross(uv);
    }
    Mat3 toRotationMatrix() const {   // Eigen's QuaternionBase::toRotationMatrix
        Mat3 r;
        const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
        const T twx = tx * w(), twy = ty * w(), twz = tz * w();
        const T txx = tx * x(), txy = ty * x(), txz = tz * x();
        const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        r(0, 0) = T(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = T(1) - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = T(1) - (txx + tyy);
        return r;
    }
private:
    Matrix<T, 4, 1> c_;
};

typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

}  // namespace Eigen

#endif
