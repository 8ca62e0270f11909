// ORACLE — TEST INFRASTRUCTURE ONLY.  Not shipped, never linked into the product library.
//
// CPU restatement of GXBeam.jl's rotation algebra (reference: src/math.jl:13-331), templated on the
// scalar type so that the analytic Jacobians can be checked against complex-step derivatives
// (the reference's own test strategy: test/jacobians.jl uses ForwardDiff).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
#pragma once
#include <cmath>
#include <complex>

namespace gxo {

template <class T> struct V3 {
    T v[3];
    T& operator[](int i) { return v[i]; }
    const T& operator[](int i) const { return v[i]; }
};
template <class T> struct M3 {
    T m[3][3];  // m[row][col]
    T& operator()(int i, int j) { return m[i][j]; }
    const T& operator()(int i, int j) const { return m[i][j]; }
};

inline double realpart(double x) { return x; }
inline double realpart(const std::complex<double>& x) { return x.real(); }

template <class T> V3<T> vec(T a, T b, T c) { return V3<T>{{a, b, c}}; }
template <class T> V3<T> zero3() { return V3<T>{{T(0), T(0), T(0)}}; }
template <class T> M3<T> zero33() { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = T(0); return r; }
template <class T> M3<T> eye3() { M3<T> r = zero33<T>(); r.m[0][0] = r.m[1][1] = r.m[2][2] = T(1); return r; }

template <class T> V3<T> operator+(const V3<T>& a, const V3<T>& b) { return V3<T>{{a[0] + b[0], a[1] + b[1], a[2] + b[2]}}; }
template <class T> V3<T> operator-(const V3<T>& a, const V3<T>& b) { return V3<T>{{a[0] - b[0], a[1] - b[1], a[2] - b[2]}}; }
template <class T> V3<T> operator-(const V3<T>& a) { return V3<T>{{-a[0], -a[1], -a[2]}}; }
template <class T> V3<T> operator*(const V3<T>& a, T s) { return V3<T>{{a[0] * s, a[1] * s, a[2] * s}}; }
template <class T> V3<T> operator*(T s, const V3<T>& a) { return V3<T>{{s * a[0], s * a[1], s * a[2]}}; }
template <class T> V3<T> operator/(const V3<T>& a, T s) { return V3<T>{{a[0] / s, a[1] / s, a[2] / s}}; }
template <class T> T dot(const V3<T>& a, const V3<T>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> V3<T> cross(const V3<T>& a, const V3<T>& b) {
    return V3<T>{{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}};
}

template <class T> M3<T> operator+(const M3<T>& a, const M3<T>& b) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
template <class T> M3<T> operator-(const M3<T>& a, const M3<T>& b) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
template <class T> M3<T> operator-(const M3<T>& a) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = -a.m[i][j]; return r; }
template <class T> M3<T> operator*(const M3<T>& a, T s) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s; return r; }
template <class T> M3<T> operator*(T s, const M3<T>& a) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j]; return r; }
template <class T> M3<T> operator/(const M3<T>& a, T s) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] / s; return r; }
template <class T> M3<T> operator*(const M3<T>& a, const M3<T>& b) {
    M3<T> r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
template <class T> V3<T> operator*(const M3<T>& a, const V3<T>& b) {
    V3<T> r;
    for (int i = 0; i < 3; i++) r.v[i] = a.m[i][0] * b[0] + a.m[i][1] * b[1] + a.m[i][2] * b[2];
    return r;
}
template <class T> M3<T> tr(const M3<T>& a) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i]; return r; }
// outer product a*b'
template <class T> M3<T> outer(const V3<T>& a, const V3<T>& b) { M3<T> r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a[i] * b[j]; return r; }
// matrix whose columns are c1, c2, c3 (Julia hcat)
template <class T> M3<T> hcat(const V3<T>& c1, const V3<T>& c2, const V3<T>& c3) {
    M3<T> r; for (int i = 0; i < 3; i++) { r.m[i][0] = c1[i]; r.m[i][1] = c2[i]; r.m[i][2] = c3[i]; } return r;
}
template <class T> V3<T> row(const M3<T>& a, int i) { return V3<T>{{a.m[i][0], a.m[i][1], a.m[i][2]}}; }
template <class T> V3<T> col(const M3<T>& a, int j) { return V3<T>{{a.m[0][j], a.m[1][j], a.m[2][j]}}; }
template <class T> V3<T> unit(int i) { V3<T> e = zero3<T>(); e.v[i] = T(1); return e; }

// math.jl:13
template <class T> M3<T> tilde(const V3<T>& x) {
    M3<T> r;
    r.m[0][0] = T(0);  r.m[0][1] = -x[2]; r.m[0][2] = x[1];
    r.m[1][0] = x[2];  r.m[1][1] = T(0);  r.m[1][2] = -x[0];
    r.m[2][0] = -x[1]; r.m[2][1] = x[0];  r.m[2][2] = T(0);
    return r;
}

// math.jl:24-36
template <class T> M3<T> wiener_milenkovic(const V3<T>& c) {
    T c0 = T(2) - dot(c, c) / T(8);
    T f = T(1) / ((T(4) - c0) * (T(4) - c0));
    M3<T> C;
    C.m[0][0] = c0 * c0 + c[0] * c[0] - c[1] * c[1] - c[2] * c[2];
    C.m[0][1] = T(2) * (c[0] * c[1] + c0 * c[2]);
    C.m[0][2] = T(2) * (c[0] * c[2] - c0 * c[1]);
    C.m[1][0] = T(2) * (c[0] * c[1] - c0 * c[2]);
    C.m[1][1] = c0 * c0 - c[0] * c[0] + c[1] * c[1] - c[2] * c[2];
    C.m[1][2] = T(2) * (c[1] * c[2] + c0 * c[0]);
    C.m[2][0] = T(2) * (c[0] * c[2] + c0 * c[1]);
    C.m[2][1] = T(2) * (c[1] * c[2] - c0 * c[0]);
    C.m[2][2] = c0 * c0 - c[0] * c[0] - c[1] * c[1] + c[2] * c[2];
    return f * C;
}

// math.jl:52-67.  For T = double this is the reference verbatim.  For the complex-step scalar the norm keeps its
// analytic dependence on θ (the integer wrap count m comes from the real part), which is what ForwardDiff's Dual
// numbers do in the reference's own derivative tests (`real(::Dual)` is the identity, `div` drops the partials).
template <class T> T rotation_parameter_scaling(const V3<T>& th) {
    using std::sqrt;
    T thnorm = sqrt(dot(th, th));
    double tn = realpart(thnorm);
    // div(x, y) on floats: truncated division
    double m = std::trunc((std::trunc(tn / 4.0) + 1.0) / 2.0);
    T scaling = T(1.0);
    if (tn > 4.0) scaling -= T(8.0 * m) / thnorm;
    return scaling;
}

// math.jl:74-77
template <class T> V3<T> rotation_parameter_scaling_th(T scaling, const V3<T>& th) {
    if (realpart(scaling) == 1.0) return zero3<T>();
    return ((T(1.0) - scaling) * th) / dot(th, th);
}

// math.jl:84-87
template <class T> M3<T> rotation_parameter_scaling_th_th(T scaling, const V3<T>& th) {
    if (realpart(scaling) == 1.0) return zero33<T>();
    T tt = dot(th, th);
    return ((T(1.0) - scaling) / tt) * (eye3<T>() - (T(3) * outer(th, th)) / tt);
}

// math.jl:94-101
template <class T> M3<T> get_C(const V3<T>& th) {
    T s = rotation_parameter_scaling(th);
    return wiener_milenkovic(s * th);
}

// common pieces: c = s θ, c_θ = s_θ θ' + s I   (math.jl:113-117)
template <class T> void scaled_params(const V3<T>& th, V3<T>& c, M3<T>& c_th) {
    T s = rotation_parameter_scaling(th);
    V3<T> s_th = rotation_parameter_scaling_th(s, th);
    c = s * th;
    c_th = outer(s_th, th) + s * eye3<T>();
}

// math.jl:111-151
template <class T> void get_C_th(const M3<T>& C, const V3<T>& th, M3<T>& C_t1, M3<T>& C_t2, M3<T>& C_t3) {
    V3<T> c; M3<T> c_th;
    scaled_params(th, c, c_th);
    T c0 = T(2) - dot(c, c) / T(8);
    V3<T> c0_c = -c / T(4);
    T tmp = T(1) / ((T(4) - c0) * (T(4) - c0));
    V3<T> tmp_c = (T(2) / ((T(4) - c0) * (T(4) - c0) * (T(4) - c0))) * c0_c;

    M3<T> A1, A2, A3;
    A1.m[0][0] = T(2) * c0 * c0_c[0] + T(2) * c[0]; A1.m[0][1] = T(2) * (c[1] + c0_c[0] * c[2]);  A1.m[0][2] = T(2) * (c[2] - c0_c[0] * c[1]);
    A1.m[1][0] = T(2) * (c[1] - c0_c[0] * c[2]);    A1.m[1][1] = T(2) * c0 * c0_c[0] - T(2) * c[0]; A1.m[1][2] = T(2) * (c0_c[0] * c[0] + c0);
    A1.m[2][0] = T(2) * (c[2] + c0_c[0] * c[1]);    A1.m[2][1] = -T(2) * (c0_c[0] * c[0] + c0);   A1.m[2][2] = T(2) * c0 * c0_c[0] - T(2) * c[0];

    A2.m[0][0] = T(2) * c0 * c0_c[1] - T(2) * c[1]; A2.m[0][1] = T(2) * (c[0] + c0_c[1] * c[2]);  A2.m[0][2] = -T(2) * (c0_c[1] * c[1] + c0);
    A2.m[1][0] = T(2) * (c[0] - c0_c[1] * c[2]);    A2.m[1][1] = T(2) * c0 * c0_c[1] + T(2) * c[1]; A2.m[1][2] = T(2) * (c[2] + c0_c[1] * c[0]);
    A2.m[2][0] = T(2) * (c0_c[1] * c[1] + c0);      A2.m[2][1] = T(2) * (c[2] - c0_c[1] * c[0]);  A2.m[2][2] = T(2) * c0 * c0_c[1] - T(2) * c[1];

    A3.m[0][0] = T(2) * c0 * c0_c[2] - T(2) * c[2]; A3.m[0][1] = T(2) * (c0_c[2] * c[2] + c0);    A3.m[0][2] = T(2) * (c[0] - c0_c[2] * c[1]);
    A3.m[1][0] = -T(2) * (c0_c[2] * c[2] + c0);     A3.m[1][1] = T(2) * c0 * c0_c[2] - T(2) * c[2]; A3.m[1][2] = T(2) * (c[1] + c0_c[2] * c[0]);
    A3.m[2][0] = T(2) * (c[0] + c0_c[2] * c[1]);    A3.m[2][1] = T(2) * (c[1] - c0_c[2] * c[0]);  A3.m[2][2] = T(2) * c0 * c0_c[2] + T(2) * c[2];

    M3<T> C_c1 = (tmp_c[0] * C) / tmp + tmp * A1;
    M3<T> C_c2 = (tmp_c[1] * C) / tmp + tmp * A2;
    M3<T> C_c3 = (tmp_c[2] * C) / tmp + tmp * A3;

    C_t1 = C_c1 * c_th(0, 0) + C_c2 * c_th(1, 0) + C_c3 * c_th(2, 0);
    C_t2 = C_c1 * c_th(0, 1) + C_c2 * c_th(1, 1) + C_c3 * c_th(2, 1);
    C_t3 = C_c1 * c_th(0, 2) + C_c2 * c_th(1, 2) + C_c3 * c_th(2, 2);
}

// math.jl:160-169
template <class T> M3<T> get_Q(const V3<T>& th) {
    T s = rotation_parameter_scaling(th);
    V3<T> c = s * th;
    T c0 = T(2) - dot(c, c) / T(8);
    T f = T(1) / ((T(4) - c0) * (T(4) - c0));
    return f * ((T(4) - T(0.25) * dot(c, c)) * eye3<T>() - T(2) * tilde(c) + T(0.5) * outer(c, c));
}

// math.jl:180-205
template <class T> void get_Q_th(const M3<T>& Q, const V3<T>& th, M3<T>& Q_t1, M3<T>& Q_t2, M3<T>& Q_t3) {
    V3<T> c; M3<T> c_th;
    scaled_params(th, c, c_th);
    T c0 = T(2) - dot(c, c) / T(8);
    V3<T> c0_c = -c / T(4);
    T tmp = T(1) / ((T(4) - c0) * (T(4) - c0));
    V3<T> tmp_c = (T(2) / ((T(4) - c0) * (T(4) - c0) * (T(4) - c0))) * c0_c;
    M3<T> Q_c[3];
    for (int k = 0; k < 3; k++) {
        V3<T> e = unit<T>(k);
        Q_c[k] = (tmp_c[k] * Q) / tmp +
                 tmp * ((-c[k] / T(2)) * eye3<T>() - T(2) * tilde(e) + (outer(e, c) + outer(c, e)) / T(2));
    }
    Q_t1 = Q_c[0] * c_th(0, 0) + Q_c[1] * c_th(1, 0) + Q_c[2] * c_th(2, 0);
    Q_t2 = Q_c[0] * c_th(0, 1) + Q_c[1] * c_th(1, 1) + Q_c[2] * c_th(2, 1);
    Q_t3 = Q_c[0] * c_th(0, 2) + Q_c[1] * c_th(1, 2) + Q_c[2] * c_th(2, 2);
}

// row-vector * matrix : (v' * A)'
template <class T> V3<T> vtm(const V3<T>& v, const M3<T>& A) { return tr(A) * v; }

// math.jl:212-231
template <class T> M3<T> get_dQ(const V3<T>& th, const V3<T>& dth, const M3<T>& Q) {
    T s = rotation_parameter_scaling(th);
    V3<T> s_th = rotation_parameter_scaling_th(s, th);
    V3<T> c = s * th;
    M3<T> c_th = outer(s_th, th) + s * eye3<T>();
    T c0 = T(2) - dot(c, c) / T(8);
    V3<T> c0_th = vtm(-c / T(4), c_th);                     // row vector
    T tmp1 = T(1) / ((T(4) - c0) * (T(4) - c0));
    V3<T> tmp1_th = (T(2) / ((T(4) - c0) * (T(4) - c0) * (T(4) - c0))) * c0_th;  // row vector
    M3<T> tmp2 = -outer(dth, c) + T(4) * tilde(dth) + dot(c, dth) * eye3<T>() + outer(c, dth);
    return outer(Q * dth, tmp1_th) / tmp1 + (T(0.5) * tmp1) * (tmp2 * (s * eye3<T>() + outer(s_th, th)));
}

// math.jl:241-280
template <class T> void get_dQ_th(const V3<T>& th, const V3<T>& dth, const M3<T>& Q, const M3<T>& Q_t1, const M3<T>& Q_t2,
                                  const M3<T>& Q_t3, M3<T>& dQ_t1, M3<T>& dQ_t2, M3<T>& dQ_t3) {
    T s = rotation_parameter_scaling(th);
    V3<T> s_th = rotation_parameter_scaling_th(s, th);
    M3<T> s_th_th = rotation_parameter_scaling_th_th(s, th);
    V3<T> c = s * th;
    M3<T> c_th = outer(s_th, th) + s * eye3<T>();
    T c0 = T(2) - dot(c, c) / T(8);
    V3<T> c0_th = vtm(-c / T(4), c_th);
    // c0_θ_θ = -I/4*c_θ'*c_θ - 1/2*c*scaling_θ' - 1/4*c'*θ*scaling_θ_θ   (math.jl:255)
    M3<T> c0_th_th = (T(-0.25) * (tr(c_th) * c_th)) - T(0.5) * outer(c, s_th) - (T(0.25) * dot(c, th)) * s_th_th;
    T d = T(4) - c0;
    T tmp1 = T(1) / (d * d);
    V3<T> tmp1_th = (T(2) / (d * d * d)) * c0_th;
    M3<T> tmp1_th_th = (T(6) / (d * d * d * d)) * outer(c0_th, c0_th) + (T(2) / (d * d * d)) * c0_th_th;
    M3<T> tmp2 = -outer(dth, c) + T(4) * tilde(dth) + dot(c, dth) * eye3<T>() + outer(c, dth);
    M3<T> SI = s * eye3<T>() + outer(s_th, th);
    const M3<T>* Q_t[3] = {&Q_t1, &Q_t2, &Q_t3};
    M3<T>* out[3] = {&dQ_t1, &dQ_t2, &dQ_t3};
    V3<T> Qd = Q * dth;
    for (int k = 0; k < 3; k++) {
        V3<T> ck = row(c_th, k);  // c_θ[k,:]
        M3<T> tmp2_k = -outer(dth, ck) + dot(ck, dth) * eye3<T>() + outer(ck, dth);
        V3<T> e = unit<T>(k);
        *out[k] = outer((*Q_t[k]) * dth, tmp1_th) / tmp1 + outer(Qd, row(tmp1_th_th, k)) / tmp1 -
                  (outer(Qd, tmp1_th) / (tmp1 * tmp1)) * tmp1_th[k] + (T(0.5) * tmp1_th[k]) * (tmp2 * SI) +
                  (T(0.5) * tmp1) * (tmp2_k * SI) +
                  (T(0.5) * tmp1) * (tmp2 * (s_th[k] * eye3<T>() + outer(row(s_th_th, k), th) + outer(s_th, e)));
    }
}

// math.jl:289-298
template <class T> M3<T> get_Qinv(const V3<T>& th) {
    T s = rotation_parameter_scaling(th);
    V3<T> c = s * th;
    return (T(1) - dot(c, c) / T(16)) * eye3<T>() + T(0.5) * tilde(c) + T(0.125) * outer(c, c);
}

// math.jl:306-323
template <class T> void get_Qinv_th(const V3<T>& th, M3<T>& Qi_t1, M3<T>& Qi_t2, M3<T>& Qi_t3) {
    V3<T> c; M3<T> c_th;
    scaled_params(th, c, c_th);
    M3<T> Qc[3];
    for (int k = 0; k < 3; k++) {
        V3<T> e = unit<T>(k);
        Qc[k] = (-c[k] / T(8)) * eye3<T>() + T(0.5) * tilde(e) + T(0.125) * (outer(e, c) + outer(c, e));
    }
    Qi_t1 = Qc[0] * c_th(0, 0) + Qc[1] * c_th(1, 0) + Qc[2] * c_th(2, 0);
    Qi_t2 = Qc[0] * c_th(0, 1) + Qc[1] * c_th(1, 1) + Qc[2] * c_th(2, 1);
    Qi_t3 = Qc[0] * c_th(0, 2) + Qc[1] * c_th(1, 2) + Qc[2] * c_th(2, 2);
}

// math.jl:331
template <class T> M3<T> mul3(const M3<T>& A1, const M3<T>& A2, const M3<T>& A3, const V3<T>& b) {
    return hcat(A1 * b, A2 * b, A3 * b);
}

}  // namespace gxo
