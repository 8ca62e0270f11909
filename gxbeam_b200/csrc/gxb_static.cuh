// Fused static residual + Jacobian assembly, one thread per "station" (point i and element i of a chain), one CTA per
// instance.  Replaces the per-iteration calls
//     static_system_residual!  src/system.jl:400-418   -> static_point_residual! (point.jl:2113), static_element_residual! (element.jl:2995)
//     static_system_jacobian!  src/system.jl:663-683   -> static_point_jacobian! (point.jl:2321), static_element_jacobian! (element.jl:3182)
// The reference evaluates the element/point properties twice (once per call) and inserts 3x3 blocks into a CSC matrix;
// here both are produced in one pass straight into the row-block storage the batched LU consumes (see gxb_lu.cuh):
//
//   Jrow[b][r][30] : row r (natural SystemIndices order) holds the 5 column 6-blocks B-2..B+2 around its own 6-block
//   B = r / 6.  Chain layout [p0 e0 p1 e1 ... pN]: point i is block 2i, element i is block 2i+1.
//
// Summation order in point rows is the reference's (point term assigned first, then element i-1 `+=`, then element i
// `-=`; system.jl:668-676, element.jl:2526-2549): the i-1 contribution to the shared θ-block travels through shared
// memory instead of atomics, so results are deterministic.
#pragma once
#include "gxb_math.cuh"

#define GXB_ROWW 30  // doubles per stored Jacobian row (static): 5 blocks x 6

struct StaticArgs {
    int N;                 // elements in the chain (points = N+1)
    int n;                 // 12 N + 6
    int two_d;
    int has_grav;          // gravity != 0 for some instance (element mass / point mass terms evaluated)
    const double* elemS;   // [b][47 or 83][N]  : L, Cab[9], S11,S12,S21,S22 (=L*compliance, 3x3 row-major each) [, m11,m12,m21,m22 (=L*mass)]
    int elem_nf;           // number of element fields stored
    const double* bcS;     // [b][18][N+1]
    const double* dlS;     // [b][24][N] or null
    const double* pmS;     // [b][36][N+1] (4 3x3 row-major blocks m11,m12,m21,m22) or null
    const double* grav;    // [b][3] or null
    const double* fs;      // [b]
    const unsigned short* flags;  // [N+1] : bits 0-5 pd, bits 6-11 pl
    const double* x;       // [b][n] state to evaluate at
    double* r;             // [b][n] residual out
    double* Jrow;          // [b][n][30] or null (residual only)
    const int* status;     // [b] or null; when given only instances with status == want_status are evaluated
    int want_status;
};

GXD v3 ldv(const double* p, int stride) { return mk(p[0], p[stride], p[2 * stride]); }
GXD m3 ldm(const double* p, int stride) { m3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = p[i * stride]; return r; }

GXD void st3(double* row, int slot, double a, double b, double c) { row[slot] = a; row[slot + 1] = b; row[slot + 2] = c; }
// write 3 rows x 3 cols block B at rows (r0..r0+2), slot s
GXD void stblk(double* J, int r0, int s, const m3& B) {
#pragma unroll
    for (int i = 0; i < 3; i++) st3(J + (size_t)(r0 + i) * GXB_ROWW, s, B.a[3 * i], B.a[3 * i + 1], B.a[3 * i + 2]);
}

// Per-thread exchange record: element i's contribution to the rows of its stop point (i+1) that overlaps with what
// thread i+1 writes: the (F,M)x θ block (18) and the residual (6).
#define GXB_XCH 24

template <bool WANT_J>
__device__ void static_station(const StaticArgs& A, int b, int i, double* xch /* smem [blockDim][24] */) {
    const int N = A.N, NP = N + 1;
    const double fs = A.fs[b];
    const double rfs = 1.0 / fs;   // NOTE: reference divides by fs; fs is a power of two by default so this is exact
    const double* x = A.x + (size_t)b * A.n;
    double* r = A.r + (size_t)b * A.n;
    double* J = WANT_J ? A.Jrow + (size_t)b * A.n * GXB_ROWW : nullptr;
    const bool has_elem = i < N;
    const bool valid = i <= N;

    v3 g = mk(0, 0, 0);
    if (A.has_grav && A.grav) g = mk(A.grav[3 * b], A.grav[3 * b + 1], A.grav[3 * b + 2]);

    // ---- point i: displacement / load selection (point.jl:20-39, 147-167) ----
    v3 u1 = mk(0, 0, 0), th1 = u1, Fp = u1, Mp = u1;       // point i
    v3 mu1 = mk(1, 1, 1), mth1 = mu1;                       // u_u, θ_θ diagonal masks (1 = free)
    unsigned fl1 = 0;
    m3 pJFth = zero33(), pJMth = zero33();                  // point-own ∂F/∂θ, ∂M/∂θ (follower loads, point-mass gravity)
    if (valid) {
        fl1 = A.flags[i];
        const double* bc = A.bcS + ((size_t)b * 18) * NP + i;
        const double* xp = x + 12 * i;
        u1 = mk((fl1 & 1) ? bc[0] : xp[0], (fl1 & 2) ? bc[NP] : xp[1], (fl1 & 4) ? bc[2 * NP] : xp[2]);
        th1 = mk((fl1 & 8) ? bc[3 * NP] : xp[3], (fl1 & 16) ? bc[4 * NP] : xp[4], (fl1 & 32) ? bc[5 * NP] : xp[5]);
        mu1 = mk((fl1 & 1) ? 0.0 : 1.0, (fl1 & 2) ? 0.0 : 1.0, (fl1 & 4) ? 0.0 : 1.0);
        mth1 = mk((fl1 & 8) ? 0.0 : 1.0, (fl1 & 16) ? 0.0 : 1.0, (fl1 & 32) ? 0.0 : 1.0);
        v3 F = ldv(bc + 6 * NP, NP), M = ldv(bc + 9 * NP, NP), Ff = ldv(bc + 12 * NP, NP), Mf = ldv(bc + 15 * NP, NP);
        const bool follower = (Ff.x != 0.0) | (Ff.y != 0.0) | (Ff.z != 0.0) | (Mf.x != 0.0) | (Mf.y != 0.0) | (Mf.z != 0.0);
        const bool pmass = A.has_grav && A.pmS != nullptr;
        v3 Fg = mk(0, 0, 0), Mg = Fg;
        if (follower || pmass) {
            Rot R;
            if (WANT_J) rot_C_dC(th1, R); else R.C = rot_C(th1);
            F = F + tmul(R.C, Ff);                 // F + C' Ff
            M = M + tmul(R.C, Mf);
            if (WANT_J) {                          // F_θ = mul3(C_θ1', C_θ2', C_θ3', Ff), rows of non-prescribed loads zeroed (point.jl:75-105)
                pJFth = cols(tmul(R.C1, Ff), tmul(R.C2, Ff), tmul(R.C3, Ff));
                pJMth = cols(tmul(R.C1, Mf), tmul(R.C2, Mf), tmul(R.C3, Mf));
            }
            if (pmass) {                           // point.jl:1422-1431, 1476-1487
                const double* pm = A.pmS + ((size_t)b * 36) * NP + i;
                m3 m11 = ldm(pm, NP), m21 = ldm(pm + 18 * NP, NP);
                v3 Cg = R.C * g;
                v3 a1 = m11 * Cg, a2 = m21 * Cg;
                Fg = tmul(R.C, a1); Mg = tmul(R.C, a2);
                if (WANT_J) {
                    m3 Cthg = cols(R.C1 * g, R.C2 * g, R.C3 * g);
                    m3 t1 = cols(tmul(R.C1, a1), tmul(R.C2, a1), tmul(R.C3, a1)) + tmul(R.C, m11 * Cthg);
                    m3 t2 = cols(tmul(R.C1, a2), tmul(R.C2, a2), tmul(R.C3, a2)) + tmul(R.C, m21 * Cthg);
                    // rows of the *load* jacobian are masked by pl, the gravity part is not (point.jl:86-105 vs 1484-1485)
                    m3 ml = pJFth;
#pragma unroll
                    for (int k = 0; k < 3; k++) { if (!(fl1 & (64 << 0))) ml.a[k] = 0.0; if (!(fl1 & (64 << 1))) ml.a[3 + k] = 0.0; if (!(fl1 & (64 << 2))) ml.a[6 + k] = 0.0; }
                    pJFth = ml + t1;
                    ml = pJMth;
#pragma unroll
                    for (int k = 0; k < 3; k++) { if (!(fl1 & (64 << 3))) ml.a[k] = 0.0; if (!(fl1 & (64 << 4))) ml.a[3 + k] = 0.0; if (!(fl1 & (64 << 5))) ml.a[6 + k] = 0.0; }
                    pJMth = ml + t2;
                }
            } else if (WANT_J) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (!(fl1 & (64 << 0))) pJFth.a[k] = 0.0; if (!(fl1 & (64 << 1))) pJFth.a[3 + k] = 0.0; if (!(fl1 & (64 << 2))) pJFth.a[6 + k] = 0.0;
                    if (!(fl1 & (64 << 3))) pJMth.a[k] = 0.0; if (!(fl1 & (64 << 4))) pJMth.a[3 + k] = 0.0; if (!(fl1 & (64 << 5))) pJMth.a[6 + k] = 0.0;
                }
            }
        }
        // loads: prescribed value where pl, else the state variable times fs  (NaN-safe: selected by flag)
        Fp = mk((fl1 & 64) ? F.x : xp[0] * fs, (fl1 & 128) ? F.y : xp[1] * fs, (fl1 & 256) ? F.z : xp[2] * fs);
        Mp = mk((fl1 & 512) ? M.x : xp[3] * fs, (fl1 & 1024) ? M.y : xp[4] * fs, (fl1 & 2048) ? M.z : xp[5] * fs);
        Fp = Fp + Fg; Mp = Mp + Mg;
    }

    // ---- element i (element.jl:62-107, 1478-1490, 1860-1893) ----
    v3 ru = mk(0, 0, 0), rth = ru, F1 = ru, M1 = ru, F2 = ru, M2 = ru;
    m3 JF1th, JM1th, JF2th, JM2th;   // ∂(F1,M1,F2,M2)/∂θ (same for θ1 and θ2 before masking)
    v3 mth2 = mk(1, 1, 1), mu2 = mth2;
    if (has_elem) {
        const double* es = A.elemS + ((size_t)b * A.elem_nf) * N + i;
        const double L = es[0];
        m3 Cab = ldm(es + N, N);
        m3 S11 = ldm(es + 10 * N, N), S12 = ldm(es + 19 * N, N), S21 = ldm(es + 28 * N, N), S22 = ldm(es + 37 * N, N);
        const unsigned fl2 = A.flags[i + 1];
        const double* bc2 = A.bcS + ((size_t)b * 18) * NP + i + 1;
        const double* xq = x + 12 * (i + 1);
        v3 u2 = mk((fl2 & 1) ? bc2[0] : xq[0], (fl2 & 2) ? bc2[NP] : xq[1], (fl2 & 4) ? bc2[2 * NP] : xq[2]);
        v3 th2 = mk((fl2 & 8) ? bc2[3 * NP] : xq[3], (fl2 & 16) ? bc2[4 * NP] : xq[4], (fl2 & 32) ? bc2[5 * NP] : xq[5]);
        mu2 = mk((fl2 & 1) ? 0.0 : 1.0, (fl2 & 2) ? 0.0 : 1.0, (fl2 & 4) ? 0.0 : 1.0);
        mth2 = mk((fl2 & 8) ? 0.0 : 1.0, (fl2 & 16) ? 0.0 : 1.0, (fl2 & 32) ? 0.0 : 1.0);
        const double* xe = x + 12 * i + 6;
        v3 F = mk(xe[0] * fs, xe[1] * fs, xe[2] * fs), M = mk(xe[3] * fs, xe[4] * fs, xe[5] * fs);
        v3 th = 0.5 * (th1 + th2);
        Rot R;
        if (WANT_J) rot_C_dC(th, R); else R.C = rot_C(th);
        m3 CtCab = tmul(R.C, Cab);
        v3 gam = S11 * F + S12 * M, kap = S21 * F + S22 * M;
        v3 Le1g = mk(L + gam.x, gam.y, gam.z);                 // L e1 + γ
        v3 Cab1 = mk(Cab.a[0], Cab.a[3], Cab.a[6]);            // Cab e1
        v3 du = CtCab * Le1g - L * Cab1;
        v3 Cabk = Cab * kap;
        m3 Qinv, dQb;
        if (WANT_J) dQb = rot_Qinv_dQinv_b(th, Cabk, Qinv); else Qinv = rot_Qinv_from_c(rot_scaling(th) * th);
        ru = (u2 - u1) - du;
        rth = (th2 - th1) - Qinv * Cabk;
        v3 CF = CtCab * F;
        v3 LgxF = cross(Le1g, F);
        v3 t1 = CtCab * M, t2 = 0.5 * (CtCab * LgxF);
        F1 = CF; F2 = CF; M1 = t1 + t2; M2 = t1 - t2;
        m3 m11, m21;
        v3 CtCabTg = mk(0, 0, 0);
        if (A.has_grav) {
            m11 = ldm(es + 46 * N, N); m21 = ldm(es + 64 * N, N);
            CtCabTg = tmul(CtCab, g);
            v3 tg = -(CtCab * (m11 * CtCabTg));
            F1 = F1 - 0.5 * tg; F2 = F2 + 0.5 * tg;
            tg = -(CtCab * (m21 * CtCabTg));
            M1 = M1 - 0.5 * tg; M2 = M2 + 0.5 * tg;
        }
        v3 f1f, f2f, m1f, m2f;
        if (A.dlS) {
            const double* d = A.dlS + ((size_t)b * 24) * N + i;
            f1f = ldv(d + 12 * N, N); f2f = ldv(d + 15 * N, N); m1f = ldv(d + 18 * N, N); m2f = ldv(d + 21 * N, N);
            F1 = F1 + (ldv(d, N) + tmul(R.C, f1f));
            F2 = F2 - (ldv(d + 3 * N, N) + tmul(R.C, f2f));
            M1 = M1 + (ldv(d + 6 * N, N) + tmul(R.C, m1f));
            M2 = M2 - (ldv(d + 9 * N, N) + tmul(R.C, m2f));
        }
        if (WANT_J) {
            double* Je = J + (size_t)(12 * i + 6) * GXB_ROWW;
            // compatibility rows (element.jl:1492-1521, 2507-2524)
            v3 w = Cab * Le1g;
            m3 duth = cols(tmul(R.C1, w), tmul(R.C2, w), tmul(R.C3, w));
            m3 ruth = -0.5 * duth;
            stblk(Je, 0, 6, colmask(-eye33(), mu1));
            stblk(Je, 0, 9, colmask(ruth, mth1));
            stblk(Je, 0, 12, (-fs) * (CtCab * S11));
            stblk(Je, 0, 15, (-fs) * (CtCab * S12));
            stblk(Je, 0, 18, colmask(eye33(), mu2));
            stblk(Je, 0, 21, colmask(ruth, mth2));
            m3 hd = -0.5 * dQb;
            m3 QC = Qinv * Cab;
            stblk(Je, 3, 6, zero33());
            stblk(Je, 3, 9, colmask(hd - eye33(), mth1));
            stblk(Je, 3, 12, (-fs) * (QC * S21));
            stblk(Je, 3, 15, (-fs) * (QC * S22));
            stblk(Je, 3, 18, zero33());
            stblk(Je, 3, 21, colmask(hd + eye33(), mth2));
            // resultant jacobians (element.jl:1945-2027)
            v3 wF = Cab * F, wM = Cab * M, wX = Cab * LgxF;
            m3 Fth = 0.5 * cols(tmul(R.C1, wF), tmul(R.C2, wF), tmul(R.C3, wF));
            m3 tm1 = cols(tmul(R.C1, wM), tmul(R.C2, wM), tmul(R.C3, wM));
            m3 tm2 = 0.5 * cols(tmul(R.C1, wX), tmul(R.C2, wX), tmul(R.C3, wX));
            JF1th = Fth; JF2th = Fth;
            JM1th = 0.5 * (tm1 + tm2); JM2th = 0.5 * (tm1 - tm2);
            if (A.has_grav) {
                m3 Cthg = cols(R.C1 * g, R.C2 * g, R.C3 * g);
                v3 a1 = Cab * (m11 * CtCabTg);
                m3 tg = -0.5 * (cols(tmul(R.C1, a1), tmul(R.C2, a1), tmul(R.C3, a1)) + CtCab * (m11 * tmul(Cab, Cthg)));
                JF1th = JF1th - 0.5 * tg; JF2th = JF2th + 0.5 * tg;
                v3 a2 = Cab * (m21 * CtCabTg);
                tg = -0.5 * (cols(tmul(R.C1, a2), tmul(R.C2, a2), tmul(R.C3, a2)) + CtCab * (m21 * tmul(Cab, Cthg)));
                JM1th = JM1th - 0.5 * tg; JM2th = JM2th + 0.5 * tg;
            }
            if (A.dlS) {
                JF1th = JF1th + 0.5 * cols(tmul(R.C1, f1f), tmul(R.C2, f1f), tmul(R.C3, f1f));
                JF2th = JF2th - 0.5 * cols(tmul(R.C1, f2f), tmul(R.C2, f2f), tmul(R.C3, f2f));
                JM1th = JM1th + 0.5 * cols(tmul(R.C1, m1f), tmul(R.C2, m1f), tmul(R.C3, m1f));
                JM2th = JM2th - 0.5 * cols(tmul(R.C1, m2f), tmul(R.C2, m2f), tmul(R.C3, m2f));
            }
            m3 M1F = 0.5 * (CtCab * (tilde(Le1g) - tilde(F) * S11));
            m3 hM = 0.5 * ((CtCab * tilde(F)) * S12);
            // rows of point i (start): element columns assigned, θ columns of point i+1 own to this element alone
            double* Ja = J + (size_t)(12 * i) * GXB_ROWW;
            stblk(Ja, 0, 18, -CtCab); stblk(Ja, 0, 21, zero33());
            stblk(Ja, 0, 27, (-rfs) * colmask(JF1th, mth2));
            stblk(Ja, 3, 18, -M1F); stblk(Ja, 3, 21, hM - CtCab);
            stblk(Ja, 3, 27, (-rfs) * colmask(JM1th, mth2));
            // rows of point i+1 (stop): columns of point i and of this element
            double* Jb = J + (size_t)(12 * (i + 1)) * GXB_ROWW;
            stblk(Jb, 0, 3, rfs * colmask(JF2th, mth1));
            stblk(Jb, 0, 6, CtCab); stblk(Jb, 0, 9, zero33());
            stblk(Jb, 3, 3, rfs * colmask(JM2th, mth1));
            stblk(Jb, 3, 6, -M1F); stblk(Jb, 3, 9, CtCab + hM);
        }
    }

    // ---- exchange: element i -> thread i+1 (its stop point's shared θ block and residual) ----
    double* mine = xch + (size_t)threadIdx.x * GXB_XCH;
    if (has_elem) {
        if (WANT_J) {
            m3 bF = rfs * colmask(JF2th, mth2), bM = rfs * colmask(JM2th, mth2);
#pragma unroll
            for (int k = 0; k < 9; k++) { mine[k] = bF.a[k]; mine[9 + k] = bM.a[k]; }
        }
        mine[18] = F2.x * rfs; mine[19] = F2.y * rfs; mine[20] = F2.z * rfs;
        mine[21] = M2.x * rfs; mine[22] = M2.y * rfs; mine[23] = M2.z * rfs;
    }
    __syncthreads();
    if (!valid) return;
    const double* prev = xch + (size_t)(i > 0 ? threadIdx.x - 1 : 0) * GXB_XCH;   // element i-1's record, used when i > 0
    double* rp = r + 12 * i;
    // ---- point i rows: own term, += element i-1, -= element i ----
    {
        double own[6] = {-Fp.x * rfs, -Fp.y * rfs, -Fp.z * rfs, -Mp.x * rfs, -Mp.y * rfs, -Mp.z * rfs};
        double sub[6] = {F1.x * rfs, F1.y * rfs, F1.z * rfs, M1.x * rfs, M1.y * rfs, M1.z * rfs};
#pragma unroll
        for (int k = 0; k < 6; k++) {
            double v = own[k];
            if (i > 0) v += prev[18 + k];
            if (has_elem) v -= sub[k];
            rp[k] = v;
        }
    }
    if (has_elem) { double* re = rp + 6; re[0] = ru.x; re[1] = ru.y; re[2] = ru.z; re[3] = rth.x; re[4] = rth.y; re[5] = rth.z; }
    if (WANT_J) {
        double* Ja = J + (size_t)(12 * i) * GXB_ROWW;
        // -F_F : identity on DOFs whose load is the unknown (point.jl:107-115, 1795)
        m3 FF = zero33(), MM = zero33();
        FF.a[0] = (fl1 & 64) ? 0.0 : -1.0; FF.a[4] = (fl1 & 128) ? 0.0 : -1.0; FF.a[8] = (fl1 & 256) ? 0.0 : -1.0;
        MM.a[0] = (fl1 & 512) ? 0.0 : -1.0; MM.a[4] = (fl1 & 1024) ? 0.0 : -1.0; MM.a[8] = (fl1 & 2048) ? 0.0 : -1.0;
        m3 JFth = (-rfs) * colmask(pJFth, mth1);                 // -F_θ θ_θ / fs
        m3 JMth = MM - rfs * colmask(pJMth, mth1);               // -M_M - M_θ θ_θ / fs
        if (i > 0) {
#pragma unroll
            for (int k = 0; k < 9; k++) { JFth.a[k] += prev[k]; JMth.a[k] += prev[9 + k]; }
        }
        if (has_elem) {
            JFth = JFth - rfs * colmask(JF1th, mth1);
            JMth = JMth - rfs * colmask(JM1th, mth1);
        }
        stblk(Ja, 0, 12, FF); stblk(Ja, 0, 15, JFth);
        stblk(Ja, 3, 12, zero33()); stblk(Ja, 3, 15, JMth);
    }
    // ---- planar constraint rows (system.jl:1072-1103): rows 2,3,4 of every 6-slot become x = 0 ----
    if (A.two_d) {
        const int nslots = has_elem ? 2 : 1;
        for (int s = 0; s < nslots; s++) {
            const double* xs = x + 12 * i + 6 * s;
            double* rs = rp + 6 * s;
            for (int k = 2; k <= 4; k++) {
                rs[k] = xs[k];
                if (WANT_J) {
                    double* row = J + (size_t)(12 * i + 6 * s + k) * GXB_ROWW;
                    for (int c = 0; c < GXB_ROWW; c++) row[c] = (c == 12 + k) ? 1.0 : 0.0;
                }
            }
        }
    }
}
