// Device-side 3-vector / 3x3 algebra and Wiener-Milenkovic rotation kernels for the fused assembly kernels.
// Everything lives in registers (the reference uses StaticArrays for the same reason).
// Reference semantics: src/math.jl:13-331 (tilde, wiener_milenkovic, rotation_parameter_scaling(_θ), get_C(_θ),
// get_Qinv(_θ), mul3).  The formulas are re-derived in factored form (shared sub-expressions, the scaling == 1 fast
// path) rather than transcribed; parity is enforced against the CPU oracle in tests/.
#pragma once
#include <cuda_runtime.h>

#define GXD __device__ __forceinline__

struct v3 { double x, y, z; };
struct m3 { double a[9]; };  // row-major: a[3*i+j]

GXD v3 mk(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
GXD v3 operator+(v3 a, v3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
GXD v3 operator-(v3 a, v3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
GXD v3 operator-(v3 a) { return mk(-a.x, -a.y, -a.z); }
GXD v3 operator*(double s, v3 a) { return mk(s * a.x, s * a.y, s * a.z); }
GXD double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GXD v3 cross(v3 a, v3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GXD double comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

GXD m3 operator+(const m3& A, const m3& B) { m3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = A.a[i] + B.a[i]; return r; }
GXD m3 operator-(const m3& A, const m3& B) { m3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = A.a[i] - B.a[i]; return r; }
GXD m3 operator-(const m3& A) { m3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = -A.a[i]; return r; }
GXD m3 operator*(double s, const m3& A) { m3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = s * A.a[i]; return r; }
GXD m3 operator*(const m3& A, const m3& B) {
    m3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[3 * i] * B.a[j] + A.a[3 * i + 1] * B.a[3 + j] + A.a[3 * i + 2] * B.a[6 + j];
    return r;
}
GXD v3 operator*(const m3& A, v3 b) {
    return mk(A.a[0] * b.x + A.a[1] * b.y + A.a[2] * b.z, A.a[3] * b.x + A.a[4] * b.y + A.a[5] * b.z,
              A.a[6] * b.x + A.a[7] * b.y + A.a[8] * b.z);
}
// A' * b
GXD v3 tmul(const m3& A, v3 b) {
    return mk(A.a[0] * b.x + A.a[3] * b.y + A.a[6] * b.z, A.a[1] * b.x + A.a[4] * b.y + A.a[7] * b.z,
              A.a[2] * b.x + A.a[5] * b.y + A.a[8] * b.z);
}
// A' * B
GXD m3 tmul(const m3& A, const m3& B) {
    m3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[i] * B.a[j] + A.a[3 + i] * B.a[3 + j] + A.a[6 + i] * B.a[6 + j];
    return r;
}
GXD m3 transpose(const m3& A) { m3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[3 * j + i]; return r; }
GXD m3 zero33() { m3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = 0.0; return r; }
GXD m3 eye33() { m3 r = zero33(); r.a[0] = r.a[4] = r.a[8] = 1.0; return r; }
GXD m3 tilde(v3 v) { m3 r; r.a[0] = 0; r.a[1] = -v.z; r.a[2] = v.y; r.a[3] = v.z; r.a[4] = 0; r.a[5] = -v.x; r.a[6] = -v.y; r.a[7] = v.x; r.a[8] = 0; return r; }
// matrix with columns c0,c1,c2
GXD m3 cols(v3 c0, v3 c1, v3 c2) { m3 r; r.a[0] = c0.x; r.a[1] = c1.x; r.a[2] = c2.x; r.a[3] = c0.y; r.a[4] = c1.y; r.a[5] = c2.y; r.a[6] = c0.z; r.a[7] = c1.z; r.a[8] = c2.z; return r; }
// scale column j of A by m.{x,y,z}  (right-multiplication by diag(m): the θ_θ / u_u masks of point.jl:193-206)
GXD m3 colmask(const m3& A, v3 m) { m3 r;
#pragma unroll
    for (int i = 0; i < 3; i++) { r.a[3 * i] = A.a[3 * i] * m.x; r.a[3 * i + 1] = A.a[3 * i + 1] * m.y; r.a[3 * i + 2] = A.a[3 * i + 2] * m.z; } return r; }

// math.jl:52-67 (truncating `div`), returns scaling and ||θ||²
GXD double rot_scaling(v3 th) {
    double n = sqrt(dot(th, th));
    double s = 1.0;
    if (n > 4.0) {
        double m = trunc((trunc(n / 4.0) + 1.0) / 2.0);
        s -= 8.0 * m / n;
    }
    return s;
}

struct Rot {       // C(θ) and ∂C/∂θ_k
    m3 C, C1, C2, C3;
};

// math.jl:24-36 with c = s θ
GXD m3 rot_C_from_c(v3 c, double& c0, double& tmp) {
    c0 = 2.0 - dot(c, c) / 8.0;
    double d = 4.0 - c0;
    tmp = 1.0 / (d * d);
    m3 C;
    double c00 = c0 * c0, x2 = c.x * c.x, y2 = c.y * c.y, z2 = c.z * c.z;
    C.a[0] = tmp * (c00 + x2 - y2 - z2);
    C.a[1] = tmp * (2.0 * (c.x * c.y + c0 * c.z));
    C.a[2] = tmp * (2.0 * (c.x * c.z - c0 * c.y));
    C.a[3] = tmp * (2.0 * (c.x * c.y - c0 * c.z));
    C.a[4] = tmp * (c00 - x2 + y2 - z2);
    C.a[5] = tmp * (2.0 * (c.y * c.z + c0 * c.x));
    C.a[6] = tmp * (2.0 * (c.x * c.z + c0 * c.y));
    C.a[7] = tmp * (2.0 * (c.y * c.z - c0 * c.x));
    C.a[8] = tmp * (c00 - x2 - y2 + z2);
    return C;
}
GXD m3 rot_C(v3 th) {
    double s = rot_scaling(th), c0, tmp;
    return rot_C_from_c(s * th, c0, tmp);
}

// c_θ = s_θ θ' + s I  (math.jl:74-77, 117); returns false when it is the identity (s == 1)
GXD bool rot_c_theta(v3 th, double s, m3& cth) {
    if (s == 1.0) return false;
    double tt = dot(th, th);
    v3 sth = ((1.0 - s) / tt) * th;
    cth.a[0] = sth.x * th.x + s; cth.a[1] = sth.x * th.y;     cth.a[2] = sth.x * th.z;
    cth.a[3] = sth.y * th.x;     cth.a[4] = sth.y * th.y + s; cth.a[5] = sth.y * th.z;
    cth.a[6] = sth.z * th.x;     cth.a[7] = sth.z * th.y;     cth.a[8] = sth.z * th.z + s;
    return true;
}
// chain rule X_θk = Σ_m X_cm c_θ[m,k]
GXD void rot_chain(const m3& cth, m3& X1, m3& X2, m3& X3) {
    m3 Y1, Y2, Y3;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        Y1.a[i] = X1.a[i] * cth.a[0] + X2.a[i] * cth.a[3] + X3.a[i] * cth.a[6];
        Y2.a[i] = X1.a[i] * cth.a[1] + X2.a[i] * cth.a[4] + X3.a[i] * cth.a[7];
        Y3.a[i] = X1.a[i] * cth.a[2] + X2.a[i] * cth.a[5] + X3.a[i] * cth.a[8];
    }
    X1 = Y1; X2 = Y2; X3 = Y3;
}

// C and its derivatives (math.jl:94-151).  ∂C/∂c_k = f_k C + tmp A_k with f_k = tmp_c[k]/tmp = -c_k / (2 (4 - c0)).
GXD void rot_C_dC(v3 th, Rot& R) {
    double s = rot_scaling(th), c0, tmp;
    v3 c = s * th;
    R.C = rot_C_from_c(c, c0, tmp);
    double q = -0.5 / (4.0 - c0);
    // c0_c = -c/4
    double gx = -0.25 * c.x, gy = -0.25 * c.y, gz = -0.25 * c.z;
    double t2 = 2.0 * tmp;
    {
        double f = q * c.x, d = c0 * gx;
        R.C1.a[0] = f * R.C.a[0] + t2 * (d + c.x);         R.C1.a[1] = f * R.C.a[1] + t2 * (c.y + gx * c.z);   R.C1.a[2] = f * R.C.a[2] + t2 * (c.z - gx * c.y);
        R.C1.a[3] = f * R.C.a[3] + t2 * (c.y - gx * c.z);  R.C1.a[4] = f * R.C.a[4] + t2 * (d - c.x);          R.C1.a[5] = f * R.C.a[5] + t2 * (gx * c.x + c0);
        R.C1.a[6] = f * R.C.a[6] + t2 * (c.z + gx * c.y);  R.C1.a[7] = f * R.C.a[7] - t2 * (gx * c.x + c0);    R.C1.a[8] = f * R.C.a[8] + t2 * (d - c.x);
    }
    {
        double f = q * c.y, d = c0 * gy;
        R.C2.a[0] = f * R.C.a[0] + t2 * (d - c.y);         R.C2.a[1] = f * R.C.a[1] + t2 * (c.x + gy * c.z);   R.C2.a[2] = f * R.C.a[2] - t2 * (gy * c.y + c0);
        R.C2.a[3] = f * R.C.a[3] + t2 * (c.x - gy * c.z);  R.C2.a[4] = f * R.C.a[4] + t2 * (d + c.y);          R.C2.a[5] = f * R.C.a[5] + t2 * (c.z + gy * c.x);
        R.C2.a[6] = f * R.C.a[6] + t2 * (gy * c.y + c0);   R.C2.a[7] = f * R.C.a[7] + t2 * (c.z - gy * c.x);   R.C2.a[8] = f * R.C.a[8] + t2 * (d - c.y);
    }
    {
        double f = q * c.z, d = c0 * gz;
        R.C3.a[0] = f * R.C.a[0] + t2 * (d - c.z);         R.C3.a[1] = f * R.C.a[1] + t2 * (gz * c.z + c0);    R.C3.a[2] = f * R.C.a[2] + t2 * (c.x - gz * c.y);
        R.C3.a[3] = f * R.C.a[3] - t2 * (gz * c.z + c0);   R.C3.a[4] = f * R.C.a[4] + t2 * (d - c.z);          R.C3.a[5] = f * R.C.a[5] + t2 * (c.y + gz * c.x);
        R.C3.a[6] = f * R.C.a[6] + t2 * (c.x + gz * c.y);  R.C3.a[7] = f * R.C.a[7] + t2 * (c.y - gz * c.x);   R.C3.a[8] = f * R.C.a[8] + t2 * (d + c.z);
    }
    m3 cth;
    if (rot_c_theta(th, s, cth)) rot_chain(cth, R.C1, R.C2, R.C3);
}

// Qinv = (1 - c'c/16) I + tilde(c)/2 + c c'/8   (math.jl:289-298)
GXD m3 rot_Qinv_from_c(v3 c) {
    double d = 1.0 - dot(c, c) / 16.0;
    m3 Q;
    Q.a[0] = d + 0.125 * c.x * c.x;          Q.a[1] = -0.5 * c.z + 0.125 * c.x * c.y; Q.a[2] = 0.5 * c.y + 0.125 * c.x * c.z;
    Q.a[3] = 0.5 * c.z + 0.125 * c.y * c.x;  Q.a[4] = d + 0.125 * c.y * c.y;          Q.a[5] = -0.5 * c.x + 0.125 * c.y * c.z;
    Q.a[6] = -0.5 * c.y + 0.125 * c.z * c.x; Q.a[7] = 0.5 * c.x + 0.125 * c.z * c.y;  Q.a[8] = d + 0.125 * c.z * c.z;
    return Q;
}
// Qinv and (∂Qinv/∂θ_k) b for a given vector b, returned as the matrix with those three columns
// (= mul3(Qinv_θ1, Qinv_θ2, Qinv_θ3, b), math.jl:306-331).
//   ∂Qinv/∂c_k b = -(c_k/8) b + ½ e_k × b + ⅛ (e_k (c·b) + c b_k)
GXD m3 rot_Qinv_dQinv_b(v3 th, v3 b, m3& Qinv) {
    double s = rot_scaling(th);
    v3 c = s * th;
    Qinv = rot_Qinv_from_c(c);
    double cb = dot(c, b);
    v3 k0 = mk(-0.125 * c.x * b.x + 0.125 * (cb + c.x * b.x), -0.125 * c.x * b.y - 0.5 * b.z + 0.125 * c.y * b.x, -0.125 * c.x * b.z + 0.5 * b.y + 0.125 * c.z * b.x);
    v3 k1 = mk(-0.125 * c.y * b.x + 0.5 * b.z + 0.125 * c.x * b.y, -0.125 * c.y * b.y + 0.125 * (cb + c.y * b.y), -0.125 * c.y * b.z - 0.5 * b.x + 0.125 * c.z * b.y);
    v3 k2 = mk(-0.125 * c.z * b.x - 0.5 * b.y + 0.125 * c.x * b.z, -0.125 * c.z * b.y + 0.5 * b.x + 0.125 * c.y * b.z, -0.125 * c.z * b.z + 0.125 * (cb + c.z * b.z));
    m3 D = cols(k0, k1, k2);   // column k = (∂Qinv/∂c_k) b
    m3 cth;
    if (rot_c_theta(th, s, cth)) D = D * cth;   // column j = Σ_k D[:,k] c_θ[k,j]
    return D;
}

// ---- Q, ∂Q/∂θ, ΔQ, ∂ΔQ/∂θ for structural damping (math.jl:160-280) ----
GXD m3 outer(v3 a, v3 b) { m3 r; r.a[0] = a.x * b.x; r.a[1] = a.x * b.y; r.a[2] = a.x * b.z; r.a[3] = a.y * b.x; r.a[4] = a.y * b.y; r.a[5] = a.y * b.z; r.a[6] = a.z * b.x; r.a[7] = a.z * b.y; r.a[8] = a.z * b.z; return r; }
GXD v3 rowv(const m3& A, int i) { return mk(A.a[3 * i], A.a[3 * i + 1], A.a[3 * i + 2]); }
GXD m3 diagm(double d) { m3 r = zero33(); r.a[0] = r.a[4] = r.a[8] = d; return r; }

struct RotQ { m3 Q, Q1, Q2, Q3; };      // Q(θ) and ∂Q/∂θ_k

// c, c_θ (identity when the scaling is 1)
GXD void rot_c(v3 th, v3& c, m3& cth, double& s, v3& sth) {
    s = rot_scaling(th);
    c = s * th;
    if (s == 1.0) { cth = eye33(); sth = mk(0, 0, 0); }
    else { double tt = dot(th, th); sth = ((1.0 - s) / tt) * th; cth = outer(sth, th) + diagm(s); }
}
GXD m3 rot_Q_from_c(v3 c, double& c0, double& tmp) {
    c0 = 2.0 - dot(c, c) / 8.0;
    const double d = 4.0 - c0;
    tmp = 1.0 / (d * d);
    return tmp * (diagm(4.0 - 0.25 * dot(c, c)) - 2.0 * tilde(c) + 0.5 * outer(c, c));
}
GXD m3 rot_Q(v3 th) { double c0, tmp; return rot_Q_from_c(rot_scaling(th) * th, c0, tmp); }
// math.jl:180-205:  ∂Q/∂c_k = f_k Q + tmp (-c_k/2 I - 2 tilde(e_k) + (e_k c' + c e_k')/2),  f_k = -c_k / (2 (4 - c0))
GXD void rot_Q_dQ(v3 th, RotQ& R) {
    v3 c, sth; m3 cth; double s, c0, tmp;
    rot_c(th, c, cth, s, sth);
    R.Q = rot_Q_from_c(c, c0, tmp);
    const double q = -0.5 / (4.0 - c0);
    m3* out[3] = {&R.Q1, &R.Q2, &R.Q3};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const v3 e = mk(k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0);
        const double ck = comp(c, k);
        *out[k] = (q * ck) * R.Q + tmp * (diagm(-0.5 * ck) - 2.0 * tilde(e) + 0.5 * (outer(e, c) + outer(c, e)));
    }
    if (s != 1.0) rot_chain(cth, R.Q1, R.Q2, R.Q3);
}
// math.jl:212-231
GXD m3 rot_dQ(v3 th, v3 dth, const m3& Q) {
    v3 c, sth; m3 cth; double s;
    rot_c(th, c, cth, s, sth);
    const double c0 = 2.0 - dot(c, c) / 8.0, d = 4.0 - c0, tmp1 = 1.0 / (d * d);
    const v3 c0th = tmul(cth, -0.25 * c);                     // row vector -c'/4 c_θ
    const m3 tmp2 = -outer(dth, c) + 4.0 * tilde(dth) + diagm(dot(c, dth)) + outer(c, dth);
    return outer(Q * dth, (2.0 / d) * c0th) + (0.5 * tmp1) * (tmp2 * cth);
}
// math.jl:241-280
GXD void rot_dQ_dth(v3 th, v3 dth, const RotQ& R, m3& D1, m3& D2, m3& D3) {
    v3 c, sth; m3 cth; double s;
    rot_c(th, c, cth, s, sth);
    m3 sthth = zero33();
    if (s != 1.0) { const double tt = dot(th, th); sthth = ((1.0 - s) / tt) * (eye33() - (3.0 / tt) * outer(th, th)); }
    const double c0 = 2.0 - dot(c, c) / 8.0, d = 4.0 - c0;
    const v3 c0th = tmul(cth, -0.25 * c);
    const m3 c0thth = (-0.25) * tmul(cth, cth) - 0.5 * outer(c, sth) - (0.25 * dot(c, th)) * sthth;
    const double tmp1 = 1.0 / (d * d);
    const v3 tmp1th = (2.0 / (d * d * d)) * c0th;
    const m3 tmp1thth = (6.0 / (d * d * d * d)) * outer(c0th, c0th) + (2.0 / (d * d * d)) * c0thth;
    const m3 tmp2 = -outer(dth, c) + 4.0 * tilde(dth) + diagm(dot(c, dth)) + outer(c, dth);
    const m3 t2SI = tmp2 * cth;
    const v3 Qd = R.Q * dth;
    const m3* Qk[3] = {&R.Q1, &R.Q2, &R.Q3};
    m3* out[3] = {&D1, &D2, &D3};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const v3 ck = rowv(cth, k);
        const m3 tmp2k = -outer(dth, ck) + diagm(dot(ck, dth)) + outer(ck, dth);
        const v3 e = mk(k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0);
        const double t1k = comp(tmp1th, k);
        *out[k] = (1.0 / tmp1) * outer((*Qk[k]) * dth, tmp1th) + (1.0 / tmp1) * outer(Qd, rowv(tmp1thth, k)) - (t1k / (tmp1 * tmp1)) * outer(Qd, tmp1th) +
                  (0.5 * t1k) * t2SI + (0.5 * tmp1) * (tmp2k * cth) +
                  (0.5 * tmp1) * (tmp2 * (diagm(comp(sth, k)) + outer(rowv(sthth, k), th) + outer(sth, e)));
    }
}
