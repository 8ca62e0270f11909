// Fused steady-state residual + Jacobian assembly on a DynamicSystem (12 states per point: u/F, θ/M, V, Ω; 6 per element),
// one thread per station (point i + element i of a chain), one CTA per instance.  Replaces
//     steady_system_residual!  src/system.jl:427-455  -> steady_point_residual! (point.jl:2137), steady_element_residual! (element.jl:3019)
//     steady_system_jacobian!  src/system.jl:693-720  -> steady_point_jacobian! (point.jl:2346), steady_element_jacobian! (element.jl:3209)
// (properties element.jl:117-160, 649-710; resultants element.jl:1901-1919, 2219-2328; point.jl:630-672, 1438-1449, 1658-1666,
//  1728-1739; inserts element.jl:2685-2759, point.jl:1907-1935, 1809-1826).
//
// Chain layout [p0(12) e0(6) p1(12) e1(6) ... pN(12)]; in 6-blocks station i owns blocks 3i (point equilibrium rows, columns
// u/θ), 3i+1 (velocity rows rV, rΩ; columns V/Ω) and 3i+2 (element).  Block bandwidths are KLB = 3, KUB = 4, so a stored row
// holds the 8 column blocks B-3 .. B+4 (48 doubles).
//
// Structural damping (element.jl:165-217, 713-820; docs/src/theory.md): γ -= μ11 γdot, κ -= μ22 κdot, with the strain rates built
// from the *velocity states* of the two end points; evaluated by damping_terms() below.
// Point masses (point.jl:630-672, 964-1023, 1134-1200) reuse the element's inertia routines with the frame A = C(θ_p)'.
// Body-frame accelerations as state variables (a DOF with pd & pl): the residual reads them from x (body_accelerations); their dense
// Jacobian columns are assembled by k_body_columns and handled by a Woodbury correction around the band solve (gxb_lib.cu).
#pragma once
#include "gxb_static.cuh"

#define GXB_DROWW 48      // doubles per stored Jacobian row (dynamic): 8 blocks x 6
#define GXB_DPARK 48      // first parking slot of a tile row (12 slots: the four column groups both neighbours add into)
#define GXB_DTROW 60      // tile row: 48 slots + 12 parking
#define GXB_DTILE 182     // tile stride per station: 3 rows x 60 + 2 pad

struct DynArgs {
    StaticArgs s;             // shared inputs (elemS incl. mass blocks, bcS, dlS, grav, fs, flags, x, r, Jrow = [b][n][48], status)
    const double* exS;        // [b][3][N]    element midpoints Δx (Element.x)
    const double* pxS;        // [b][3][N+1]  point coordinates Δx
    const double* motion;     // [b][12]      vb, ωb, ab, αb
    int chunk;                // stations per tile flush (shared-memory tile holds `chunk` stations)
    // Newmark time marching (MODE 1): rate initialisation terms `paug` (analyses.jl:2693-2738) and 2/dt
    const double* init;       // [b][12][N+1] udot_init, θdot_init, Vdot_init, Ωdot_init
    double c2dt;              // 2/dt
    int damping;              // structural_damping (elemS fields 82..87 hold Element.mu)
    // body-frame accelerations that are state variables (a DOF with displacement AND load prescribed, system.jl:138-150, 167-182):
    // ibody[i] = 0-based column of x holding linear (i < 3) / angular (i >= 3) acceleration component i, or -1
    int ibody[6];
};
// body_accelerations (system.jl:167-182)
GXD void body_accelerations(const DynArgs& A, const double* x, const double* mo, v3& ab, v3& alb) {
    ab = mk(A.ibody[0] >= 0 ? x[A.ibody[0]] : mo[6], A.ibody[1] >= 0 ? x[A.ibody[1]] : mo[7], A.ibody[2] >= 0 ? x[A.ibody[2]] : mo[8]);
    alb = mk(A.ibody[3] >= 0 ? x[A.ibody[3]] : mo[9], A.ibody[4] >= 0 ? x[A.ibody[4]] : mo[10], A.ibody[5] >= 0 ? x[A.ibody[5]] : mo[11]);
}

GXD void dblk(double* t, int s, const m3& B) {
#pragma unroll
    for (int i = 0; i < 3; i++) { t[GXB_DTROW * i + s] = B.a[3 * i]; t[GXB_DTROW * i + s + 1] = B.a[3 * i + 1]; t[GXB_DTROW * i + s + 2] = B.a[3 * i + 2]; }
}
GXD m3 dget(const double* t, int s) { m3 B;
#pragma unroll
    for (int i = 0; i < 3; i++) { B.a[3 * i] = t[GXB_DTROW * i + s]; B.a[3 * i + 1] = t[GXB_DTROW * i + s + 1]; B.a[3 * i + 2] = t[GXB_DTROW * i + s + 2]; }
    return B; }
GXD void dtile_clear(double* t) {
#pragma unroll
    for (int j = 0; j < 3 * GXB_DTROW; j += 2) *reinterpret_cast<double2*>(t + j) = make_double2(0.0, 0.0);
}
// stream rows of the tiles of stations [c0, c0+cnt) to rows rowof(st) + roff .. +2
GXD void dtile_flush(const double* tiles, double* J, int c0, int cnt, int roff) {
    for (int c = threadIdx.x; c < cnt * 72; c += blockDim.x) {           // 3 rows x 24 16-byte chunks per station
        const int st = c / 72, q = c - 72 * st, row = q / 24, qq = q - 24 * row;
        const double2 v = *reinterpret_cast<const double2*>(tiles + (size_t)st * GXB_DTILE + row * GXB_DTROW + 2 * qq);
        *reinterpret_cast<double2*>(J + (size_t)(18 * (c0 + st) + roff + row) * GXB_DROWW + 2 * qq) = v;
    }
}
GXD void dtile_planar(double* t, int first_row_in_slot, int diag_slot0) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int k = first_row_in_slot + i;
        if (k >= 2 && k <= 4) {
#pragma unroll
            for (int c = 0; c < GXB_DROWW; c++) t[GXB_DTROW * i + c] = (c == diag_slot0 + k) ? 1.0 : 0.0;
        }
    }
}
GXD m3 mul3c(const Rot& R, v3 w) { return cols(R.C1 * w, R.C2 * w, R.C3 * w); }   // mul3(C_θ1, C_θ2, C_θ3, w)


// ---- structural damping: strain rates and their derivatives (element.jl:165-217, 713-820) ----
struct DampOut {
    v3 gamdot, kapdot;
    // ∂γ/∂(u1, θ, V1, Ω) and ∂κ/∂(θ1, θ2, Ω1, Ω2), damping coefficients already applied (γ_u2 = -γ_u1, γ_V2 = -γ_V1,
    // γ_θ1 = γ_θ2, γ_Ω1 = γ_Ω2)
    m3 g_u1, g_th, g_V1, g_Om, k_th1, k_th2, k_Om1, k_Om2;
};
GXD m3 rowscale(const m3& A, v3 m) { m3 r;
#pragma unroll
    for (int j = 0; j < 3; j++) { r.a[j] = A.a[j] * m.x; r.a[3 + j] = A.a[3 + j] * m.y; r.a[6 + j] = A.a[6 + j] * m.z; } return r; }

template <bool WANT_J>
__device__ __noinline__ void damping_terms(const m3& Cab, const m3& CtCab, const Rot& R, v3 th, v3 th1, v3 th2, v3 u1, v3 u2, v3 V1, v3 V2,
                                           v3 Om1, v3 Om2, v3 Om, v3 vb, v3 wb, v3 dx1, v3 dx2, double L, v3 mua, v3 mub, DampOut& o) {
    Rot R1, R2;
    if (WANT_J) { rot_C_dC(th1, R1); rot_C_dC(th2, R2); } else { R1.C = rot_C(th1); R2.C = rot_C(th2); }
    const v3 w1 = Om1 - wb, w2 = Om2 - wb;
    const v3 udot1 = V1 - vb - cross(wb, dx1 + u1), udot2 = V2 - vb - cross(wb, dx2 + u2);
    m3 Qi1, Qi2, tth1, tth2;             // θdot_i = Qinv_i C_i (Ω_i - ωb) and ∂θdot_i/∂θ_i
    if (WANT_J) {
        tth1 = rot_Qinv_dQinv_b(th1, R1.C * w1, Qi1) + Qi1 * mul3c(R1, w1);
        tth2 = rot_Qinv_dQinv_b(th2, R2.C * w2, Qi2) + Qi2 * mul3c(R2, w2);
    } else { Qi1 = rot_Qinv_from_c(rot_scaling(th1) * th1); Qi2 = rot_Qinv_from_c(rot_scaling(th2) * th2); }
    const v3 thd1 = Qi1 * (R1.C * w1), thd2 = Qi2 * (R2.C * w2);
    const v3 thed = 0.5 * (thd1 + thd2);
    const v3 du = u2 - u1, dth = th2 - th1, dud = udot2 - udot1, dthd = thd2 - thd1;
    RotQ RQ;
    if (WANT_J) rot_Q_dQ(th, RQ); else RQ.Q = rot_Q(th);
    const m3 dQ = rot_dQ(th, dth, RQ.Q);
    const v3 W = Om - wb;
    const v3 Cab1 = mk(Cab.a[0], Cab.a[3], Cab.a[6]);
    o.gamdot = tmul(CtCab, dud - cross(W, du) - L * cross(W, Cab1));
    o.kapdot = tmul(Cab, RQ.Q * dthd + dQ * thed);
    if (!WANT_J) return;
    const m3 At = transpose(CtCab), Wt = tilde(W);
    o.g_u1 = -rowscale(At * Wt + At * tilde(wb), mua);
    {
        const m3 t = tmul(Cab, mul3c(R, dud - Wt * du - L * (Wt * Cab1)));
        o.g_th = -0.5 * rowscale(t, mua);
    }
    o.g_V1 = rowscale(At, mua);
    o.g_Om = -0.5 * rowscale(At * tilde(du) + L * (At * tilde(Cab1)), mua);
    m3 D1, D2, D3;
    rot_dQ_dth(th, dth, RQ, D1, D2, D3);
    const m3 t1 = tmul(Cab, cols(RQ.Q1 * dthd, RQ.Q2 * dthd, RQ.Q3 * dthd));
    const m3 t2 = tmul(Cab, cols(D1 * thed, D2 * thed, D3 * thed));
    const m3 t3 = tmul(Cab, thed.x * RQ.Q1 + thed.y * RQ.Q2 + thed.z * RQ.Q3);
    const m3 CQ = tmul(Cab, RQ.Q), CdQ = tmul(Cab, dQ);
    const m3 Km = 0.5 * CdQ - CQ, Kp = 0.5 * CdQ + CQ;
    const m3 half = 0.5 * (t1 + t2);
    o.k_th1 = -rowscale(half - t3 + Km * tth1, mub);
    o.k_th2 = -rowscale(half + t3 + Kp * tth2, mub);
    o.k_Om1 = -rowscale(Km * (Qi1 * R1.C), mub);
    o.k_Om2 = -rowscale(Kp * (Qi2 * R2.C), mub);
}


// ---- inertial loads of a body with mass blocks m11..m22 carried in a frame A (element: A = CtCab, point mass: A = C') ----
// momenta P = A m11 A' V + A m12 A' Ω, H likewise; rates Pdot, Hdot for MODE 0 (Vdot, Ωdot given, A fixed) or MODE 1 (Newmark:
// Adot = tilde(Ω - ωb) A; element.jl:397-407, point.jl:779-792).  Returns tf = ωb x P + Pdot and tm = ωb x H + V x P + Hdot.
template <int MODE>
GXD void inertia_loads(const m3& A, const m3& m11, const m3& m12, const m3& m21, const m3& m22, v3 V, v3 Om, v3 Vdot, v3 Omdot, v3 wb,
                       v3& Pm, v3& H, v3& tf, v3& tm) {
    const v3 cV = tmul(A, V), cO = tmul(A, Om), cVd = tmul(A, Vdot), cOd = tmul(A, Omdot);
    Pm = A * (m11 * cV) + A * (m12 * cO);
    H = A * (m21 * cV) + A * (m22 * cO);
    v3 Pdot = A * (m11 * cVd) + A * (m12 * cOd);
    v3 Hdot = A * (m21 * cVd) + A * (m22 * cOd);
    if (MODE == 1) {      // A m Bᵀ V = -A m Aᵀ (W x V),  B m Aᵀ V = W x (A m Aᵀ V)  with B = tilde(W) A, W = Ω - ωb
        const v3 W = Om - wb;
        const v3 cWV = tmul(A, cross(W, V)), cWO = tmul(A, cross(W, Om));
        Pdot = Pdot - (A * (m11 * cWV) + A * (m12 * cWO)) + cross(W, Pm);
        Hdot = Hdot - (A * (m21 * cWV) + A * (m22 * cWO)) + cross(W, H);
    }
    tf = cross(wb, Pm) + Pdot;
    tm = cross(wb, H) + cross(V, Pm) + Hdot;
}
// Derivatives of tf, tm with respect to (u, θ, V, Ω) of the carrying frame (element.jl:649-710, 1049-1112, 2258-2320;
// point.jl:964-1023, 1134-1200, 1517-1540):  T_u = Pdot_u, T_θ = ω~b P_θ + Pdot_θ, T_V = ω~b P_V + Pdot_V, T_Ω = ω~b P_Ω + Pdot_Ω and the
// moment twins (with the V~ P_x and -P~ terms).  `Cab` is the fixed part of A = C' Cab (identity for a point mass).
struct InertiaJac { m3 Fu, Fth, FV, FO, Mu, Mth, MV, MO; };
template <int MODE>
GXD void inertia_jacobians(const m3& A, const m3& Cab, const Rot& R, const m3& m11, const m3& m12, const m3& m21, const m3& m22, v3 V, v3 Om,
                           v3 Vdot, v3 Omdot, v3 wb, v3 alb, v3 Pm, double c2, InertiaJac& o) {
    const m3 At = transpose(A), wbt = tilde(wb);
    const m3 A11 = (A * m11) * At, A12 = (A * m12) * At, A21 = (A * m21) * At, A22 = (A * m22) * At;
    const v3 cV = tmul(A, V), cO = tmul(A, Om), cVd = tmul(A, Vdot), cOd = tmul(A, Omdot);
    const m3 cthV = tmul(Cab, mul3c(R, V)), cthO = tmul(Cab, mul3c(R, Om)), cthVd = tmul(Cab, mul3c(R, Vdot)), cthOd = tmul(Cab, mul3c(R, Omdot));
    const m3 P_th = mul3t(R, Cab * (m11 * cV + m12 * cO)) + A * (m11 * cthV) + A * (m12 * cthO);
    const m3 H_th = mul3t(R, Cab * (m21 * cV + m22 * cO)) + A * (m21 * cthV) + A * (m22 * cthO);
    m3 Pd_th = mul3t(R, Cab * (m11 * cVd)) + mul3t(R, Cab * (m12 * cOd)) + A * (m11 * cthVd) + A * (m12 * cthOd);
    m3 Hd_th = mul3t(R, Cab * (m21 * cVd)) + mul3t(R, Cab * (m22 * cOd)) + A * (m21 * cthVd) + A * (m22 * cthOd);
    const m3 Vt = tilde(V);
    m3 PdV = zero33(), PdO = zero33(), HdV = zero33(), HdO = zero33();
    if (MODE == 1) {
        const v3 Wv = Om - wb;
        const m3 Wt = tilde(Wv);
        const m3 Bm = Wt * A;                           // Adot
        const v3 bV = tmul(Bm, V), bO = tmul(Bm, Om);   // Bᵀ V, Bᵀ Ω
        const m3 cthWV = tmul(Cab, mul3c(R, Wt * V)), cthWO = tmul(Cab, mul3c(R, Wt * Om));
        Pd_th = Pd_th + Wt * mul3t(R, Cab * (m11 * cV + m12 * cO)) + Bm * (m11 * cthV) + Bm * (m12 * cthO) +
                mul3t(R, Cab * (m11 * bV + m12 * bO)) - A * (m11 * cthWV) - A * (m12 * cthWO);
        Hd_th = Hd_th + Wt * mul3t(R, Cab * (m21 * cV + m22 * cO)) + Bm * (m21 * cthV) + Bm * (m22 * cthO) +
                mul3t(R, Cab * (m21 * bV + m22 * bO)) - A * (m21 * cthWV) - A * (m22 * cthWO);
        const m3 BmT = transpose(Bm), Ot = tilde(Om);
        PdV = (A * m11) * BmT + (Bm * m11) * At + c2 * A11;
        PdO = (A * m12) * BmT + (Bm * m12) * At + A11 * Vt + A12 * Ot - tilde(A11 * V) - tilde(A12 * Om) + c2 * A12;
        HdV = (A * m21) * BmT + (Bm * m21) * At + c2 * A21;
        HdO = (A * m22) * BmT + (Bm * m22) * At + A21 * Vt + A22 * Ot - tilde(A21 * V) - tilde(A22 * Om) + c2 * A22;
    }
    const m3 albt = tilde(alb);
    o.Fu = (MODE == 0) ? A11 * albt : zero33();
    o.Fth = wbt * P_th + Pd_th;
    o.FV = wbt * A11 + PdV;
    o.FO = wbt * A12 + PdO;
    o.Mu = (MODE == 0) ? A21 * albt : zero33();
    o.Mth = wbt * H_th + Vt * P_th + Hd_th;
    o.MV = wbt * A21 + Vt * A11 - tilde(Pm) + HdV;
    o.MO = wbt * A22 + Vt * A12 + HdO;
}

// One tile pass: every thread may deposit into the tile of station `i+1` (nb) before the barrier and into its own afterwards.
// The CTA walks the instance in chunks of A.chunk stations; deposits happen when the target station is in the current chunk.
#define GXB_PASS_BEGIN(c0)  for (int c0 = 0; c0 < NP; c0 += A.chunk) { const int cnt_ = min(A.chunk, NP - c0);
#define GXB_PASS_END        }

// MODE 0: steady state (udot = θdot = 0, Vdot = ab + αb x (Δx + u), Ωdot = αb)
// MODE 1: Newmark step  (xdot = 2/dt x - xdot_init for u, θ, V, Ω; point.jl:766-795, element.jl:381-412)
// PASSES selects which row passes of the Jacobian this call produces (bit 0: point force rows, 1: point moment rows, 2: the two velocity
// row passes, 3: the two element compatibility passes) and whether it stores the residual (bit 4).  The default does everything in one
// thread per station; k_assemble_dynamic_split gives the three groups {residual + force rows}, {moment rows}, {velocity + element rows}
// to three CTAs per instance, each repeating the shared core evaluation, so that a SMALL batch sees a third of the serial chain
// (the rounds of a time-domain run are latency-bound: 152 us per Jacobian launch with one thread per station).
#define GXB_DPASS_ALL 31
template <bool WANT_J, int MODE, int PASSES = GXB_DPASS_ALL>
__device__ void dynamic_station(const DynArgs& A, int b, int i, double* smem) {
    const StaticArgs& S = A.s;
    const int N = S.N, NP = N + 1;
    double* xch = smem;                                             // [blockDim][6]
    double* tiles = smem + (size_t)blockDim.x * GXB_XCH;            // [chunk][GXB_DTILE]   (WANT_J only)
    const double fs = S.fs[b];
    const double rfs = 1.0 / fs;
    const double* x = S.x + (size_t)b * S.n;
    double* r = S.r + (size_t)b * S.n;
    double* J = WANT_J ? S.Jrow + (size_t)b * S.n * GXB_DROWW : nullptr;
    const bool has_elem = i < N;
    const bool valid = i <= N;

    v3 g = mk(0, 0, 0);
    if (S.has_grav && S.grav) g = mk(S.grav[3 * b], S.grav[3 * b + 1], S.grav[3 * b + 2]);
    const double* mo = A.motion + 12 * (size_t)b;
    const v3 vb = mk(mo[0], mo[1], mo[2]), wb = mk(mo[3], mo[4], mo[5]);
    v3 ab = mk(0, 0, 0), alb = ab;
    if (MODE == 0) body_accelerations(A, x, mo, ab, alb);
    const m3 wbt = tilde(wb);
    const double c2 = MODE == 1 ? A.c2dt : 0.0;
    const double* ini = (MODE == 1) ? A.init + ((size_t)b * 12) * NP : nullptr;   // field f of point p at ini[f*NP + p]

    // ---- point i (point.jl:630-672): displacements, loads, velocities ----
    v3 u1 = mk(0, 0, 0), th1 = u1, Fp = u1, Mp = u1, V1 = u1, Om1 = u1, dxp = u1;
    v3 mu1 = mk(1, 1, 1), mth1 = mu1;
    unsigned fl1 = 0;
    bool follower = false;
    if (valid) {
        fl1 = S.flags[i];
        const double* bc = S.bcS + ((size_t)b * 18) * NP + i;
        const double* xp = x + 18 * i;
        u1 = mk((fl1 & 1) ? bc[0] : xp[0], (fl1 & 2) ? bc[NP] : xp[1], (fl1 & 4) ? bc[2 * NP] : xp[2]);
        th1 = mk((fl1 & 8) ? bc[3 * NP] : xp[3], (fl1 & 16) ? bc[4 * NP] : xp[4], (fl1 & 32) ? bc[5 * NP] : xp[5]);
        mu1 = mk((fl1 & 1) ? 0.0 : 1.0, (fl1 & 2) ? 0.0 : 1.0, (fl1 & 4) ? 0.0 : 1.0);
        mth1 = mk((fl1 & 8) ? 0.0 : 1.0, (fl1 & 16) ? 0.0 : 1.0, (fl1 & 32) ? 0.0 : 1.0);
        V1 = mk(xp[6], xp[7], xp[8]); Om1 = mk(xp[9], xp[10], xp[11]);
        const double* px = A.pxS + ((size_t)b * 3) * NP + i;
        dxp = ldv(px, NP);
        v3 F = ldv(bc + 6 * NP, NP), M = ldv(bc + 9 * NP, NP), Ff = ldv(bc + 12 * NP, NP), Mf = ldv(bc + 15 * NP, NP);
        follower = (Ff.x != 0.0) | (Ff.y != 0.0) | (Ff.z != 0.0) | (Mf.x != 0.0) | (Mf.y != 0.0) | (Mf.z != 0.0);
        if (follower) { m3 C = rot_C(th1); F = F + tmul(C, Ff); M = M + tmul(C, Mf); }
        Fp = mk((fl1 & 64) ? F.x : xp[0] * fs, (fl1 & 128) ? F.y : xp[1] * fs, (fl1 & 256) ? F.z : xp[2] * fs);
        Mp = mk((fl1 & 512) ? M.x : xp[3] * fs, (fl1 & 1024) ? M.y : xp[4] * fs, (fl1 & 2048) ? M.z : xp[5] * fs);
    }

    // ---- element i core (element.jl:62-107, 117-160) ----
    const double* es = S.elemS + ((size_t)b * S.elem_nf) * N + (has_elem ? i : 0);
    double L = 0.0;
    m3 Cab = zero33(), CtCab = zero33();
    Rot R;
    v3 th = mk(0, 0, 0), F = th, M = th, Le1g = th, kap = th, u2 = th, th2 = th, V = th, Om = th, Vdot = th, Omdot = th, Pm = th, H = th;
    v3 mth2 = mk(1, 1, 1), mu2 = mth2;
    v3 ru = th, rth = th, F1 = th, M1 = th, F2 = th, M2 = th;
    v3 f1f = th, f2f = th, m1f = th, m2f = th, CtCabTg = th;
    DampOut dmp;
    if (has_elem) {
        L = es[0];
        Cab = ldm(es + N, N);
        const unsigned fl2 = S.flags[i + 1];
        const double* bc2 = S.bcS + ((size_t)b * 18) * NP + i + 1;
        const double* xq = x + 18 * (i + 1);
        u2 = mk((fl2 & 1) ? bc2[0] : xq[0], (fl2 & 2) ? bc2[NP] : xq[1], (fl2 & 4) ? bc2[2 * NP] : xq[2]);
        th2 = mk((fl2 & 8) ? bc2[3 * NP] : xq[3], (fl2 & 16) ? bc2[4 * NP] : xq[4], (fl2 & 32) ? bc2[5 * NP] : xq[5]);
        mu2 = mk((fl2 & 1) ? 0.0 : 1.0, (fl2 & 2) ? 0.0 : 1.0, (fl2 & 4) ? 0.0 : 1.0);
        mth2 = mk((fl2 & 8) ? 0.0 : 1.0, (fl2 & 16) ? 0.0 : 1.0, (fl2 & 32) ? 0.0 : 1.0);
        const v3 V2 = mk(xq[6], xq[7], xq[8]), Om2 = mk(xq[9], xq[10], xq[11]);
        const double* xe = x + 18 * i + 12;
        F = mk(xe[0] * fs, xe[1] * fs, xe[2] * fs); M = mk(xe[3] * fs, xe[4] * fs, xe[5] * fs);
        th = 0.5 * (th1 + th2);
        const v3 u = 0.5 * (u1 + u2);
        V = 0.5 * (V1 + V2); Om = 0.5 * (Om1 + Om2);
        if (WANT_J) rot_C_dC(th, R); else R.C = rot_C(th);
        CtCab = tmul(R.C, Cab);
        v3 gam;
        {
            m3 S11 = ldm(es + 10 * N, N), S12 = ldm(es + 19 * N, N);
            gam = S11 * F + S12 * M;
        }
        {
            m3 S21 = ldm(es + 28 * N, N), S22 = ldm(es + 37 * N, N);
            kap = S21 * F + S22 * M;
        }
        if (A.damping) {      // γ -= μ11 γdot, κ -= μ22 κdot   (element.jl:207-209)
            const v3 mua = ldv(es + 82 * N, N), mub = ldv(es + 85 * N, N);
            const v3 dx2 = ldv(A.pxS + ((size_t)b * 3) * NP + i + 1, NP);
            damping_terms<WANT_J>(Cab, CtCab, R, th, th1, th2, u1, u2, V1, V2, Om1, Om2, Om, vb, wb, dxp, dx2, L, mua, mub, dmp);
            gam = gam - mk(mua.x * dmp.gamdot.x, mua.y * dmp.gamdot.y, mua.z * dmp.gamdot.z);
            kap = kap - mk(mub.x * dmp.kapdot.x, mub.y * dmp.kapdot.y, mub.z * dmp.kapdot.z);
        }
        Le1g = mk(L + gam.x, gam.y, gam.z);
        v3 Cab1 = mk(Cab.a[0], Cab.a[3], Cab.a[6]);
        ru = (u2 - u1) - (CtCab * Le1g - L * Cab1);
        rth = (th2 - th1) - rot_Qinv_from_c(rot_scaling(th) * th) * (Cab * kap);
        v3 CF = CtCab * F;
        v3 t1 = CtCab * M, t2 = 0.5 * (CtCab * cross(Le1g, F));
        F1 = CF; F2 = CF; M1 = t1 + t2; M2 = t1 - t2;
        if (S.has_grav) {
            CtCabTg = tmul(CtCab, g);
            v3 tg = -(CtCab * (ldm(es + 46 * N, N) * CtCabTg));
            F1 = F1 - 0.5 * tg; F2 = F2 + 0.5 * tg;
            tg = -(CtCab * (ldm(es + 64 * N, N) * CtCabTg));
            M1 = M1 - 0.5 * tg; M2 = M2 + 0.5 * tg;
        }
        if (S.dlS) {
            const double* d = S.dlS + ((size_t)b * 24) * N + i;
            f1f = ldv(d + 12 * N, N); f2f = ldv(d + 15 * N, N); m1f = ldv(d + 18 * N, N); m2f = ldv(d + 21 * N, N);
            F1 = F1 + (ldv(d, N) + tmul(R.C, f1f));
            F2 = F2 - (ldv(d + 3 * N, N) + tmul(R.C, f2f));
            M1 = M1 + (ldv(d + 6 * N, N) + tmul(R.C, m1f));
            M2 = M2 - (ldv(d + 9 * N, N) + tmul(R.C, m2f));
        }
        // inertial loads (element.jl:138-155, 1901-1919)
        const v3 dxe = ldv(A.exS + ((size_t)b * 3) * N + i, N);
        if (MODE == 0) { Vdot = ab + cross(alb, dxe + u); Omdot = alb; }
        else {
            const v3 V1d = c2 * V1 - ldv(ini + 6 * NP + i, NP), O1d = c2 * Om1 - ldv(ini + 9 * NP + i, NP);
            const v3 V2d = c2 * V2 - ldv(ini + 6 * NP + i + 1, NP), O2d = c2 * Om2 - ldv(ini + 9 * NP + i + 1, NP);
            Vdot = 0.5 * (V1d + V2d); Omdot = 0.5 * (O1d + O2d);
        }
        {
            const m3 m11 = ldm(es + 46 * N, N), m12 = ldm(es + 55 * N, N), m21 = ldm(es + 64 * N, N), m22 = ldm(es + 73 * N, N);
            v3 tf, tm;
            inertia_loads<MODE>(CtCab, m11, m12, m21, m22, V, Om, Vdot, Omdot, wb, Pm, H, tf, tm);
            F1 = F1 - 0.5 * tf; F2 = F2 + 0.5 * tf;
            M1 = M1 - 0.5 * tm; M2 = M2 + 0.5 * tm;
        }
        double* mine = xch + (size_t)threadIdx.x * GXB_XCH;
        mine[0] = F2.x * rfs; mine[1] = F2.y * rfs; mine[2] = F2.z * rfs;
        mine[3] = M2.x * rfs; mine[4] = M2.y * rfs; mine[5] = M2.z * rfs;
    }
    __syncthreads();

    // ---- residual ----
    m3 Cp = zero33(), Qp = zero33();       // C(θ_p), Qinv(θ_p) of the point (velocity rows)
    const bool pmass = S.pmS != nullptr;
    v3 Vdp = mk(0, 0, 0), Odp = Vdp, Pp = Vdp;      // point-mass Vdot, Ωdot, linear momentum
    if (valid) {
        const double* prev = xch + (size_t)(i > 0 ? threadIdx.x - 1 : 0) * GXB_XCH;
        double* rp = r + 18 * i;
        Cp = rot_C(th1); Qp = rot_Qinv_from_c(rot_scaling(th1) * th1);
        if (pmass) {       // point.jl:630-672, 766-795, 1422-1449
            const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
            const m3 m11 = ldm(pm, NP), m12 = ldm(pm + 9 * NP, NP), m21 = ldm(pm + 18 * NP, NP), m22 = ldm(pm + 27 * NP, NP);
            if (MODE == 0) { Vdp = ab + cross(alb, dxp + u1); Odp = alb; }
            else { Vdp = c2 * V1 - ldv(ini + 6 * NP + i, NP); Odp = c2 * Om1 - ldv(ini + 9 * NP + i, NP); }
            const m3 Ap = transpose(Cp);
            v3 Hp, tf, tm;
            inertia_loads<MODE>(Ap, m11, m12, m21, m22, V1, Om1, Vdp, Odp, wb, Pp, Hp, tf, tm);
            Fp = Fp - tf; Mp = Mp - tm;
            if (S.has_grav) { const v3 Cg = Cp * g; Fp = Fp + tmul(Cp, m11 * Cg); Mp = Mp + tmul(Cp, m21 * Cg); }
        }
        double own[6] = {-Fp.x * rfs, -Fp.y * rfs, -Fp.z * rfs, -Mp.x * rfs, -Mp.y * rfs, -Mp.z * rfs};
        double sub[6] = {F1.x * rfs, F1.y * rfs, F1.z * rfs, M1.x * rfs, M1.y * rfs, M1.z * rfs};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            double v = own[k];
            if (i > 0) v += prev[k];
            if (has_elem) v -= sub[k];
            if (S.two_d && k >= 2 && k <= 4) v = x[18 * i + k];
            if (PASSES & 16) rp[k] = v;
        }
        // velocity rows (point.jl:1658-1666): rV = V - vb - ωb x (Δx + u),  rΩ = Qinv C (Ω - ωb)
        v3 rV = V1 - vb - cross(wb, dxp) - cross(wb, u1);
        v3 rO = Qp * (Cp * (Om1 - wb));
        if (MODE == 1) {     // - udot, - θdot with xdot = 2/dt x - xdot_init   (point.jl:776-777, 1663-1664)
            rV = rV - (c2 * u1 - ldv(ini + i, NP));
            rO = rO - (c2 * th1 - ldv(ini + 3 * NP + i, NP));
        }
        double rv[6] = {rV.x, rV.y, rV.z, rO.x, rO.y, rO.z};
#pragma unroll
        for (int k = 0; k < 6; k++) if (PASSES & 16) rp[6 + k] = (S.two_d && k >= 2 && k <= 4) ? x[18 * i + 6 + k] : rv[k];
        if (has_elem && (PASSES & 16)) {
            double re[6] = {ru.x, ru.y, ru.z, rth.x, rth.y, rth.z};
#pragma unroll
            for (int k = 0; k < 6; k++) rp[12 + k] = (S.two_d && k >= 2 && k <= 4) ? x[18 * i + 12 + k] : re[k];
        }
    }
    if (!WANT_J) return;

    // ---- element Jacobian pieces shared by the two equilibrium passes (element.jl:649-710, 2219-2328) ----
    // "h" quantities are the ½·½ inertial terms that enter BOTH end points with the same sign:
    //   rows of the start point:  -= F1_x / fs = + hF_x / fs ;  rows of the stop point: += F2_x / fs = + hF_x / fs
    m3 hFu, hFth, hFV, hFO, hMu, hMth, hMV, hMO;
    if (has_elem && (PASSES & 3)) {
        const m3 m11 = ldm(es + 46 * N, N), m12 = ldm(es + 55 * N, N), m21 = ldm(es + 64 * N, N), m22 = ldm(es + 73 * N, N);
        InertiaJac ij;
        inertia_jacobians<MODE>(CtCab, Cab, R, m11, m12, m21, m22, V, Om, Vdot, Omdot, wb, alb, Pm, c2, ij);
        hFu = 0.25 * ij.Fu; hFth = 0.25 * ij.Fth; hFV = 0.25 * ij.FV; hFO = 0.25 * ij.FO;
        hMu = 0.25 * ij.Mu; hMth = 0.25 * ij.Mth; hMV = 0.25 * ij.MV; hMO = 0.25 * ij.MO;
    }

    // =================== pass P1: force rows of point i (rows 18i+0..2) ===================
    if constexpr ((PASSES & 1) != 0) {
    m3 JFa = zero33(), JFb = zero33();
    if (has_elem) {
        m3 Fth = 0.5 * mul3t(R, Cab * F);
        JFa = Fth; JFb = Fth;
        if (S.has_grav) {
            m3 m11 = ldm(es + 46 * N, N);
            m3 Cthg = mul3c(R, g);
            m3 tg = -0.5 * (mul3t(R, Cab * (m11 * CtCabTg)) + CtCab * (m11 * tmul(Cab, Cthg)));
            JFa = JFa - 0.5 * tg; JFb = JFb + 0.5 * tg;
        }
        if (S.dlS) { JFa = JFa + 0.5 * mul3t(R, f1f); JFb = JFb - 0.5 * mul3t(R, f2f); }
        JFa = JFa - hFth; JFb = JFb + hFth;
    }
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        const bool next_in = has_elem && i + 1 >= c0 && i + 1 < c0 + cnt_;
        if (mine_in) dtile_clear(mytile);
        __syncthreads();
        if (next_in) {      // rows of the stop point (station i+1): columns of point i (slots 0..11), of this element (12..17); shared groups parked
            double* nt = tiles + (size_t)(i + 1 - c0) * GXB_DTILE;
            dblk(nt, 0, rfs * colmask(hFu, mu1));
            dblk(nt, 3, rfs * colmask(JFb, mth1));
            dblk(nt, 6, rfs * hFV);
            dblk(nt, 9, rfs * hFO);
            dblk(nt, 12, CtCab);                                  // F2_F
            dblk(nt, GXB_DPARK + 0, rfs * colmask(hFu, mu2));
            dblk(nt, GXB_DPARK + 3, rfs * colmask(JFb, mth2));
            dblk(nt, GXB_DPARK + 6, rfs * hFV);
            dblk(nt, GXB_DPARK + 9, rfs * hFO);
        }
        __syncthreads();
        if (mine_in) {
            m3 own = zero33();
            m3 FF = zero33();
            FF.a[0] = (fl1 & 64) ? 0.0 : -1.0; FF.a[4] = (fl1 & 128) ? 0.0 : -1.0; FF.a[8] = (fl1 & 256) ? 0.0 : -1.0;
            m3 bu = FF + dget(mytile, GXB_DPARK + 0);
            m3 bV = dget(mytile, GXB_DPARK + 6), bO = dget(mytile, GXB_DPARK + 9);
            if (follower || pmass) {
                Rot Rp; rot_C_dC(th1, Rp);
                const double* bc = S.bcS + ((size_t)b * 18) * NP + i;
                m3 Fth = rowmask(mul3t(Rp, ldv(bc + 12 * NP, NP)), fl1 & 64, fl1 & 128, fl1 & 256);     // follower loads (point.jl:75-105)
                if (pmass) {
                    const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
                    const m3 m11 = ldm(pm, NP), m12 = ldm(pm + 9 * NP, NP), m21 = ldm(pm + 18 * NP, NP), m22 = ldm(pm + 27 * NP, NP);
                    if (S.has_grav) Fth = Fth + mul3t(Rp, m11 * (Rp.C * g)) + tmul(Rp.C, m11 * mul3c(Rp, g));   // point.jl:1484
                    InertiaJac ij;
                    inertia_jacobians<MODE>(transpose(Rp.C), eye33(), Rp, m11, m12, m21, m22, V1, Om1, Vdp, Odp, wb, alb, Pp, c2, ij);
                    Fth = Fth - ij.Fth;                           // F_θ -= ω~b P_θ + Pdot_θ   (point.jl:1527)
                    bu = bu + rfs * colmask(ij.Fu, mu1);          // -F_u u_u / fs, F_u = -Pdot_u
                    bV = bV + rfs * ij.FV; bO = bO + rfs * ij.FO; // -F_V / fs, -F_Ω / fs
                }
                own = (-rfs) * colmask(Fth, mth1);
            }
            m3 bt = own + dget(mytile, GXB_DPARK + 3);
            if (has_elem) {
                bu = bu + rfs * colmask(hFu, mu1);                // -= F1_u1 u1_u1 / fs
                bt = bt - rfs * colmask(JFa, mth1);
                bV = bV + rfs * hFV;
                bO = bO + rfs * hFO;
                dblk(mytile, 30, -CtCab);                         // -F1_F
                dblk(mytile, 36, rfs * colmask(hFu, mu2));
                dblk(mytile, 39, (-rfs) * colmask(JFa, mth2));
                dblk(mytile, 42, rfs * hFV);
                dblk(mytile, 45, rfs * hFO);
            }
            dblk(mytile, 18, bu); dblk(mytile, 21, bt); dblk(mytile, 24, bV); dblk(mytile, 27, bO);
            if (S.two_d) dtile_planar(mytile, 0, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 0);
        __syncthreads();
    GXB_PASS_END
    }

    // =================== pass P2: moment rows of point i (rows 18i+3..5) ===================
    if constexpr ((PASSES & 2) != 0) {
    m3 JMa = zero33(), JMb = zero33(), M1F = zero33(), hM = zero33();
    m3 dMu1 = zero33(), dMV1 = zero33(), dMO = zero33();      // structural-damping parts (u2, V2 carry the opposite sign)
    if (has_elem) {
        m3 tm1 = mul3t(R, Cab * M);
        m3 tm2 = 0.5 * mul3t(R, Cab * cross(Le1g, F));
        JMa = 0.5 * (tm1 + tm2);
        JMb = 0.5 * (tm1 - tm2);
        if (S.has_grav) {
            m3 m21 = ldm(es + 64 * N, N);
            m3 Cthg = mul3c(R, g);
            m3 tg = -0.5 * (mul3t(R, Cab * (m21 * CtCabTg)) + CtCab * (m21 * tmul(Cab, Cthg)));
            JMa = JMa - 0.5 * tg; JMb = JMb + 0.5 * tg;
        }
        if (S.dlS) { JMa = JMa + 0.5 * mul3t(R, m1f); JMb = JMb - 0.5 * mul3t(R, m2f); }
        JMa = JMa - hMth; JMb = JMb + hMth;
        m3 CtF = CtCab * tilde(F);
        if (A.damping) {      // M1_x -= ½ CtCab F~ γ_x,  M2_x += ½ CtCab F~ γ_x   (element.jl:2232-2256)
            const m3 tF = 0.5 * CtF;
            const m3 dth_ = tF * dmp.g_th;
            JMa = JMa - dth_; JMb = JMb + dth_;
            dMu1 = tF * dmp.g_u1; dMV1 = tF * dmp.g_V1; dMO = tF * dmp.g_Om;
        }
        M1F = 0.5 * (CtCab * tilde(Le1g) - CtF * ldm(es + 10 * N, N));
        hM = 0.5 * (CtF * ldm(es + 19 * N, N));
    }
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        const bool next_in = has_elem && i + 1 >= c0 && i + 1 < c0 + cnt_;
        if (mine_in) dtile_clear(mytile);
        __syncthreads();
        if (next_in) {
            double* nt = tiles + (size_t)(i + 1 - c0) * GXB_DTILE;
            dblk(nt, 0, rfs * colmask(hMu + dMu1, mu1));
            dblk(nt, 3, rfs * colmask(JMb, mth1));
            dblk(nt, 6, rfs * (hMV + dMV1));
            dblk(nt, 9, rfs * (hMO + dMO));
            dblk(nt, 12, -M1F);                                   // M2_F
            dblk(nt, 15, CtCab + hM);                             // M2_M
            dblk(nt, GXB_DPARK + 0, rfs * colmask(hMu - dMu1, mu2));
            dblk(nt, GXB_DPARK + 3, rfs * colmask(JMb, mth2));
            dblk(nt, GXB_DPARK + 6, rfs * (hMV - dMV1));
            dblk(nt, GXB_DPARK + 9, rfs * (hMO + dMO));
        }
        __syncthreads();
        if (mine_in) {
            m3 MM = zero33();
            MM.a[0] = (fl1 & 512) ? 0.0 : -1.0; MM.a[4] = (fl1 & 1024) ? 0.0 : -1.0; MM.a[8] = (fl1 & 2048) ? 0.0 : -1.0;
            m3 own = MM;
            m3 bu = dget(mytile, GXB_DPARK + 0);
            m3 bV = dget(mytile, GXB_DPARK + 6), bO = dget(mytile, GXB_DPARK + 9);
            if (follower || pmass) {
                Rot Rp; rot_C_dC(th1, Rp);
                const double* bc = S.bcS + ((size_t)b * 18) * NP + i;
                m3 Mth = rowmask(mul3t(Rp, ldv(bc + 15 * NP, NP)), fl1 & 512, fl1 & 1024, fl1 & 2048);
                if (pmass) {
                    const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
                    const m3 m11 = ldm(pm, NP), m12 = ldm(pm + 9 * NP, NP), m21 = ldm(pm + 18 * NP, NP), m22 = ldm(pm + 27 * NP, NP);
                    if (S.has_grav) Mth = Mth + mul3t(Rp, m21 * (Rp.C * g)) + tmul(Rp.C, m21 * mul3c(Rp, g));
                    InertiaJac ij;
                    inertia_jacobians<MODE>(transpose(Rp.C), eye33(), Rp, m11, m12, m21, m22, V1, Om1, Vdp, Odp, wb, alb, Pp, c2, ij);
                    Mth = Mth - ij.Mth;
                    bu = bu + rfs * colmask(ij.Mu, mu1);
                    bV = bV + rfs * ij.MV; bO = bO + rfs * ij.MO;
                }
                own = MM - rfs * colmask(Mth, mth1);
            }
            m3 bt = own + dget(mytile, GXB_DPARK + 3);
            if (has_elem) {
                bu = bu + rfs * colmask(hMu + dMu1, mu1);
                bt = bt - rfs * colmask(JMa, mth1);
                bV = bV + rfs * (hMV + dMV1);
                bO = bO + rfs * (hMO + dMO);
                dblk(mytile, 30, -M1F);                           // -M1_F
                dblk(mytile, 33, hM - CtCab);                     // -M1_M
                dblk(mytile, 36, rfs * colmask(hMu - dMu1, mu2));
                dblk(mytile, 39, (-rfs) * colmask(JMa, mth2));
                dblk(mytile, 42, rfs * (hMV - dMV1));
                dblk(mytile, 45, rfs * (hMO + dMO));
            }
            dblk(mytile, 18, bu); dblk(mytile, 21, bt); dblk(mytile, 24, bV); dblk(mytile, 27, bO);
            if (S.two_d) dtile_planar(mytile, 3, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 3);
        __syncthreads();
    GXB_PASS_END
    }

    if constexpr ((PASSES & 4) != 0) {
    // =================== passes V1/V2: velocity rows of point i (rows 18i+6..8, 18i+9..11; block 3i+1) ===================
    // row block 3i+1 stores column blocks 3i-2 ..: point i u/θ at slots 12..17, V/Ω at 18..23   (point.jl:1728-1739, 1928-1932)
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            m3 rVu = -wbt;
            if (MODE == 1) { rVu.a[0] -= c2; rVu.a[4] -= c2; rVu.a[8] -= c2; }   // - udot_u = -(2/dt) I   (point.jl:1708-1720)
            dblk(mytile, 12, colmask(rVu, mu1));                  // rV_u u_u
            dblk(mytile, 18, eye33());                            // rV_V
            if (S.two_d) dtile_planar(mytile, 0, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 6);
        __syncthreads();
    GXB_PASS_END
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = valid && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            Rot Rp; rot_C_dC(th1, Rp);
            const v3 dO = Om1 - wb;
            m3 Qi;
            m3 rOth = rot_Qinv_dQinv_b(th1, Rp.C * dO, Qi) + Qi * mul3c(Rp, dO);
            if (MODE == 1) { rOth.a[0] -= c2; rOth.a[4] -= c2; rOth.a[8] -= c2; }
            dblk(mytile, 15, colmask(rOth, mth1));                // rΩ_θ θ_θ
            dblk(mytile, 21, Qi * Rp.C);                          // rΩ_Ω = Qinv C
            if (S.two_d) dtile_planar(mytile, 3, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, cnt_, 9);
        __syncthreads();
    GXB_PASS_END
    }

    if constexpr ((PASSES & 8) != 0) {
    // =================== passes E1/E2: compatibility rows of element i (rows 18i+12..14, 18i+15..17; block 3i+2) ===================
    // row block 3i+2 stores column blocks 3i-1 ..: point i u/θ at slots 6..11, V/Ω 12..17, element 18..23, point i+1 u/θ 24..29
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = has_elem && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            m3 ruth = -0.5 * mul3t(R, Cab * Le1g);
            m3 ruu1 = -eye33(), ruu2 = eye33();
            if (A.damping) {      // dynamic_compatibility_jacobians (element.jl:1550-1565)
                const m3 Agu = CtCab * dmp.g_u1;
                ruu1 = ruu1 - Agu; ruu2 = ruu2 + Agu;             // γ_u2 = -γ_u1
                ruth = ruth - CtCab * dmp.g_th;
                const m3 AgV = CtCab * dmp.g_V1, AgO = CtCab * dmp.g_Om;
                dblk(mytile, 12, -AgV); dblk(mytile, 15, -AgO);   // ru_V1, ru_Ω1
                dblk(mytile, 30, AgV); dblk(mytile, 33, -AgO);    // ru_V2 = +CtCab γ_V1, ru_Ω2
            }
            dblk(mytile, 6, colmask(ruu1, mu1));
            dblk(mytile, 9, colmask(ruth, mth1));
            dblk(mytile, 18, (-fs) * (CtCab * ldm(es + 10 * N, N)));
            dblk(mytile, 21, (-fs) * (CtCab * ldm(es + 19 * N, N)));
            dblk(mytile, 24, colmask(ruu2, mu2));
            dblk(mytile, 27, colmask(ruth, mth2));
            if (S.two_d) dtile_planar(mytile, 0, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, min(cnt_, N - c0), 12);
        __syncthreads();
    GXB_PASS_END
    GXB_PASS_BEGIN(c0)
        double* mytile = tiles + (size_t)(i - c0) * GXB_DTILE;
        const bool mine_in = has_elem && i >= c0 && i < c0 + cnt_;
        if (mine_in) {
            dtile_clear(mytile);
            m3 Qinv;
            m3 hd = -0.5 * rot_Qinv_dQinv_b(th, Cab * kap, Qinv);
            m3 QC = Qinv * Cab;
            m3 rt1 = hd - eye33(), rt2 = hd + eye33();
            if (A.damping) {      // rθ_θ -= Qinv Cab κ_θ, rθ_Ω = -Qinv Cab κ_Ω   (element.jl:1567-1572)
                rt1 = rt1 - QC * dmp.k_th1; rt2 = rt2 - QC * dmp.k_th2;
                dblk(mytile, 15, -(QC * dmp.k_Om1)); dblk(mytile, 33, -(QC * dmp.k_Om2));
            }
            dblk(mytile, 9, colmask(rt1, mth1));
            dblk(mytile, 18, (-fs) * (QC * ldm(es + 28 * N, N)));
            dblk(mytile, 21, (-fs) * (QC * ldm(es + 37 * N, N)));
            dblk(mytile, 27, colmask(rt2, mth2));
            if (S.two_d) dtile_planar(mytile, 3, 18);
        }
        __syncthreads();
        dtile_flush(tiles, J, c0, min(cnt_, N - c0), 15);
        __syncthreads();
    GXB_PASS_END
    }
}
