// Device side of the solver: packing kernels, the kernel wrappers of the fused assembly, body-acceleration columns, initial-condition
// helpers, the Newton step (factorise + solve, Woodbury correction, Armijo / convergence update), mass matrix and Arnoldi kernels.
// Host code and the C-ABI live in gxb_lib.cu.
#pragma once
#include "gxb_lu.cuh"
#include "gxb_lu_mma.cuh"
#include "gxb_lu_part.cuh"
#include "gxb_lu_two.cuh"
#include "gxb_static.cuh"
#include "gxb_dynamic.cuh"
#include "gxb_initial.cuh"

#ifndef GXB_LU_MINBLOCKS
#define GXB_LU_MINBLOCKS 4
#endif
#ifndef GXB_LU_MINBLOCKS_DYN
#define GXB_LU_MINBLOCKS_DYN 2      // CTAs per SM of the tensor-core factorisation for dynamic systems (218 registers)
#endif

// ST_SINGULAR: the factorisation hit an exactly-zero / non-finite pivot; the instance waits for k_singular_fallback (NLsolve's
// (J'J + λI) step), which moves it on to ST_TRIAL
enum { ST_SOLVE = 0, ST_TRIAL = 1, ST_DONE = 2, ST_FAILED = 3, ST_SINGULAR = 4 };
enum { FAIL_NONE = 0, FAIL_NONFINITE = 1, FAIL_LINESEARCH = 2, FAIL_MAXITER = 3, FAIL_SINGULAR = 4 };

// per-instance Newton / line-search scalars (structure of arrays)
struct NewtonState {
    int* status; int* failcode; int* iters; int* fcalls; int* ls_k; int* ls_finite;
    double* a1; double* a2; double* phi0; double* dphi0; double* phx0; double* finf;
};


// ------------------------------------------------------------------------------------------------------------------
// packing kernels: caller's array-of-structs (the reference's memory layout) -> structure-of-arrays in HBM
// elemS[b][f][e]: f=0 L; 1..9 Cab (row-major); 10..45 S11,S12,S21,S22 = L*compliance; 46..81 m11,m12,m21,m22 = L*mass;
// 82..87 mu (structural damping coefficients)
#define GXB_ELEM_NF 88
// compact variant for static contexts without gravity: raw holds 49 doubles per element (L, x[3], compliance[36], Cab[9]) fetched with
// two strided copies; the mass and damping fields, which that path never reads, are zero-filled
__global__ void k_pack_elements_compact(const double* __restrict__ raw, double* __restrict__ out, int N, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * N) return;
    int64_t b = t / N; int e = (int)(t % N);
    const double* s = raw + t * 49;
    double* o = out + (size_t)b * GXB_ELEM_NF * N + e;
    const double L = s[0];
    o[0] = L;
    for (int k = 0; k < 6; k++) o[(size_t)(82 + k) * N] = 0.0;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[(size_t)(1 + 3 * i + j) * N] = s[40 + i + 3 * j];
    for (int bi = 0; bi < 2; bi++) for (int bj = 0; bj < 2; bj++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        int f = 9 * (2 * bi + bj) + 3 * i + j;
        int src = (3 * bi + i) + 6 * (3 * bj + j);
        o[(size_t)(10 + f) * N] = s[4 + src] * L;
        o[(size_t)(46 + f) * N] = 0.0;
    }
}
__global__ void k_pack_elements(const double* __restrict__ raw, double* __restrict__ out, double* __restrict__ exS, int N, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * N) return;
    int64_t b = t / N; int e = (int)(t % N);
    const double* s = raw + t * GXB_ELEM_STRIDE;
    double* o = out + (size_t)b * GXB_ELEM_NF * N + e;
    const double L = s[0];
    o[0] = L;
    if (exS) for (int k = 0; k < 3; k++) exS[((size_t)b * 3 + k) * N + e] = s[1 + k];    // Element.x (midpoint)
    for (int k = 0; k < 6; k++) o[(size_t)(82 + k) * N] = s[85 + k];                       // Element.mu
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[(size_t)(1 + 3 * i + j) * N] = s[76 + i + 3 * j];
    for (int bi = 0; bi < 2; bi++) for (int bj = 0; bj < 2; bj++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        int f = 9 * (2 * bi + bj) + 3 * i + j;
        int src = (3 * bi + i) + 6 * (3 * bj + j);
        o[(size_t)(10 + f) * N] = s[4 + src] * L;     // compliance *= L   (element.jl:69)
        o[(size_t)(46 + f) * N] = s[40 + src] * L;    // mass *= L         (element.jl:70)
    }
}
// out[b][f][p] = raw[b][p][f]
__global__ void k_pack_transpose(const double* __restrict__ raw, double* __restrict__ out, int P, int F, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * P * F) return;
    int f = (int)(t % F); int64_t q = t / F; int p = (int)(q % P); int64_t b = q / P;
    out[((size_t)b * F + f) * P + p] = raw[t];
}
// point masses: raw[b][p][36 col-major] -> out[b][9*(2bi+bj)+3i+j][p]
__global__ void k_pack_pmass(const double* __restrict__ raw, double* __restrict__ out, int P, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * P * 36) return;
    int f = (int)(t % 36); int64_t q = t / 36; int p = (int)(q % P); int64_t b = q / P;
    int blk = f / 9, i = (f % 9) / 3, j = f % 3, bi = blk / 2, bj = blk % 2;
    out[((size_t)b * 36 + f) * P + p] = raw[q * 36 + (3 * bi + i) + 6 * (3 * bj + j)];
}

// ------------------------------------------------------------------------------------------------------------------
template <bool WANT_J> __global__ void __launch_bounds__(256) k_assemble(StaticArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.status && A.status[b] != A.want_status) return;
    static_station<WANT_J>(A, b, threadIdx.x, asm_smem);
}
// residual-only evaluation of chains of up to 127 elements: 128-thread CTAs, three of them per SM (168 registers instead of 182)
__global__ void __launch_bounds__(128, 3) k_assemble_residual_128(StaticArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.status && A.status[b] != A.want_status) return;
    static_station<false>(A, b, threadIdx.x, asm_smem);
}

template <bool WANT_J, int MODE> __global__ void __launch_bounds__(256) k_assemble_dynamic(DynArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.s.status && A.s.status[b] != A.s.want_status) return;
    dynamic_station<WANT_J, MODE>(A, b, threadIdx.x, asm_smem);
}

// pass-parallel Jacobian assembly for small batches: three CTAs per instance (blockIdx.y), see dynamic_station
// (gridDim.y = 3: {residual + force rows}, {moment rows}, {velocity + element rows};  gridDim.y = 2: {force + moment rows},
//  {residual + velocity + element rows} -- for batches whose 3 nbatch CTAs would not be resident in one wave)
template <int MODE> __global__ void __launch_bounds__(256) k_assemble_dynamic_split(DynArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.s.status && A.s.status[b] != A.s.want_status) return;
    if (gridDim.y == 3) {
        if (blockIdx.y == 0) dynamic_station<true, MODE, 1 | 16>(A, b, threadIdx.x, asm_smem);
        else if (blockIdx.y == 1) dynamic_station<true, MODE, 2>(A, b, threadIdx.x, asm_smem);
        else dynamic_station<true, MODE, 4 | 8>(A, b, threadIdx.x, asm_smem);
    } else {
        if (blockIdx.y == 0) dynamic_station<true, MODE, 1 | 2>(A, b, threadIdx.x, asm_smem);
        else dynamic_station<true, MODE, 4 | 8 | 16>(A, b, threadIdx.x, asm_smem);
    }
}

template <bool WANT_J> __global__ void __launch_bounds__(256) k_assemble_initial(InitArgs A) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.d.s.status && A.d.s.status[b] != A.d.s.want_status) return;
    initial_station<WANT_J>(A, b, threadIdx.x, asm_smem);
}

// Dense Jacobian columns of the body-frame accelerations that are state variables (insert_steady_point_jacobians! point.jl:1829-1841,
// insert_steady_element_jacobians! element.jl:2567-2583; the same terms for the initial-condition analysis, point.jl:1879-1895).
// Column i (acceleration component i, stored at column c_i = ibody[i]) has entries in the force/moment rows of EVERY point:
//   + (tf_x, tm_x)[:, i] / fs of the point's own mass and + 1/2 of each adjacent element's, with
//   tf_ab = A m11 A', tf_alb = -A m11 A' tilde(dx + u) + A m12 A', tm likewise with m21, m22   (Vdot_ab = I, Vdot_alb = -tilde(dx + u), Ωdot_alb = I)
// which does not fit the band.  The band matrix B keeps a unit entry at (c_i, c_i) instead (the force-free, displacement-free DOF
// owns that row), and Dn[b][i][:] = column_i - e_{c_i}, so that J = B + sum_i Dn_i e_{c_i}' exactly; the solve undoes the split
// with a Woodbury correction (k_woodbury).  One thread per station, both adjacent elements evaluated by the thread itself.
struct BodyArgs { DynArgs d; const double* icS; const unsigned* rvm; int mode; double* Dn; };   // Dn: [6][nbatch][n]
__global__ void __launch_bounds__(256) k_body_columns(BodyArgs B) {
    const DynArgs& A = B.d; const StaticArgs& S = A.s;
    const int b = blockIdx.x, p = threadIdx.x, N = S.N, NP = N + 1;
    if (S.status && S.status[b] != S.want_status) return;
    if (p > N) return;
    const double rfs = 1.0 / S.fs[b];
    const double* x = S.x + (size_t)b * S.n;
    auto disp = [&](int q, v3& u, v3& th) {
        const unsigned fl = S.flags[q];
        const double* bc = S.bcS + ((size_t)b * 18) * NP + q;
        const double* xp = x + 18 * q;
        double t[6];
        for (int k = 0; k < 6; k++) {
            const bool pd = (fl >> k) & 1u;
            if (B.mode == 2) {
                const bool r = (B.rvm[(size_t)b * NP + q] >> k) & 1u;
                t[k] = pd ? bc[(size_t)k * NP] : (r ? B.icS[((size_t)b * 18 + k) * NP + q] : xp[k]);
            } else t[k] = pd ? bc[(size_t)k * NP] : xp[k];
        }
        u = mk(t[0], t[1], t[2]); th = mk(t[3], t[4], t[5]);
    };
    double G[6][6];                      // rows: F(3), M(3) of this point; columns: ab(3), alb(3)
    for (int r = 0; r < 6; r++) for (int cc = 0; cc < 6; cc++) G[r][cc] = 0.0;
    auto add = [&](const m3& Am, const m3& m11, const m3& m12, const m3& m21, const m3& m22, v3 rr, double w) {
        const m3 At = transpose(Am), rt = tilde(rr);
        const m3 A11 = (Am * m11) * At, A12 = (Am * m12) * At, A21 = (Am * m21) * At, A22 = (Am * m22) * At;
        const m3 Fal = A12 - A11 * rt, Mal = A22 - A21 * rt;
        for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) {
            G[r][cc] += w * A11.a[3 * r + cc]; G[r][3 + cc] += w * Fal.a[3 * r + cc];
            G[3 + r][cc] += w * A21.a[3 * r + cc]; G[3 + r][3 + cc] += w * Mal.a[3 * r + cc];
        }
    };
    v3 up, thp; disp(p, up, thp);
    if (S.pmS) {
        const double* pm = S.pmS + ((size_t)b * 36) * NP + p;
        const v3 dxp = ldv(A.pxS + ((size_t)b * 3) * NP + p, NP);
        add(transpose(rot_C(thp)), ldm(pm, NP), ldm(pm + 9 * NP, NP), ldm(pm + 18 * NP, NP), ldm(pm + 27 * NP, NP), dxp + up, rfs);
    }
    for (int side = 0; side < 2; side++) {
        const int e = p - 1 + side;
        if (e < 0 || e >= N) continue;
        v3 u1, t1, u2, t2; disp(e, u1, t1); disp(e + 1, u2, t2);
        const double* es = S.elemS + ((size_t)b * S.elem_nf) * N + e;
        const m3 Am = tmul(rot_C(0.5 * (t1 + t2)), ldm(es + N, N));
        const v3 dxe = ldv(A.exS + ((size_t)b * 3) * N + e, N);
        add(Am, ldm(es + 46 * N, N), ldm(es + 55 * N, N), ldm(es + 64 * N, N), ldm(es + 73 * N, N), dxe + 0.5 * (u1 + u2), 0.5 * rfs);
    }
    for (int i = 0; i < 6; i++) {
        const int c = A.ibody[i];
        if (c < 0) continue;
        double* Di = B.Dn + ((size_t)i * gridDim.x + b) * S.n + 18 * p;
        for (int k = 0; k < 6; k++) {
            double v = G[k][i];
            if (B.mode == 2 && ((B.rvm[(size_t)b * NP + p] >> (8 + k)) & 1u)) v = 0.0;    // equilibrium row replaced by Vdot_k = Vdot0_k (system.jl:756-813)
            if (S.two_d && k >= 2 && k <= 4) v = 0.0;                                      // planar constraint rows (system.jl:1090-1100)
            if (18 * p + k == c) v -= 1.0;
            Di[k] = v;
        }
        if (c / 18 == p && S.Jrow) S.Jrow[((size_t)b * S.n + c) * GXB_DROWW + 18 + (c % 18)] = 1.0;    // the unit entry B keeps at (c_i, c_i)
    }
}

// rate_vars of the initial-condition analysis (src/analyses.jl:1835-1873): column sums of the mass matrix at x = 0 over the V/Ω
// columns of each point (rate_vars1), and the same with the element mass blocks replaced by the compliance blocks and the point
// masses ignored (rate_vars2).  One thread per (instance, point).  rvm bit k: rate_vars2, bit 8+k: equilibrium row replaced
// (!pd & !rate_vars1 & rate_vars2, system.jl:491-513), bit 16+k: rate_vars1.
__global__ void k_initial_rate_vars(const double* __restrict__ elemS, int elem_nf, const double* __restrict__ bcS, const double* __restrict__ pmS,
                                    const unsigned short* __restrict__ flags, int N, int two_d, unsigned* __restrict__ rvm, int64_t nbatch) {
    const int NP = N + 1;
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * NP) return;
    const int64_t b = t / NP; const int p = (int)(t % NP);
    auto theta = [&](int q) {
        const unsigned fl = flags[q];
        const double* bc = bcS + ((size_t)b * 18) * NP + q;
        return mk((fl & 8) ? bc[3 * NP] : 0.0, (fl & 16) ? bc[4 * NP] : 0.0, (fl & 32) ? bc[5 * NP] : 0.0);
    };
    double s1[6] = {0, 0, 0, 0, 0, 0}, s2[6] = {0, 0, 0, 0, 0, 0};
    auto addcols = [&](double* s, const m3& PV, const m3& PO, const m3& HV, const m3& HO, double w) {
        for (int r = 0; r < 3; r++) {
            const bool lin = !(two_d && r == 2), ang = !(two_d && r < 2);       // lmask = (1,1,0), amask = (0,0,1) (point.jl:2043-2049)
            for (int cc = 0; cc < 3; cc++) {
                if (lin) { s[cc] += w * PV.a[3 * r + cc]; s[3 + cc] += w * PO.a[3 * r + cc]; }
                if (ang) { s[cc] += w * HV.a[3 * r + cc]; s[3 + cc] += w * HO.a[3 * r + cc]; }
            }
        }
    };
    if (pmS) {
        const double* pm = pmS + ((size_t)b * 36) * NP + p;
        const m3 C = rot_C(theta(p));
        addcols(s1, tmul(C, ldm(pm, NP) * C), tmul(C, ldm(pm + 9 * NP, NP) * C), tmul(C, ldm(pm + 18 * NP, NP) * C), tmul(C, ldm(pm + 27 * NP, NP) * C), 1.0);
    }
    for (int side = 0; side < 2; side++) {
        const int e = p - 1 + side;
        if (e < 0 || e >= N) continue;
        const double* es = elemS + ((size_t)b * elem_nf) * N + e;
        const m3 Cab = ldm(es + N, N);
        const m3 Am = tmul(rot_C(0.5 * (theta(e) + theta(e + 1))), Cab), At = transpose(Am);
        // ¼ A m A' enters the rows of both end points of the element
        addcols(s1, (Am * ldm(es + 46 * N, N)) * At, (Am * ldm(es + 55 * N, N)) * At, (Am * ldm(es + 64 * N, N)) * At, (Am * ldm(es + 73 * N, N)) * At, 0.5);
        addcols(s2, (Am * ldm(es + 10 * N, N)) * At, (Am * ldm(es + 19 * N, N)) * At, (Am * ldm(es + 28 * N, N)) * At, (Am * ldm(es + 37 * N, N)) * At, 0.5);
    }
    const unsigned fl = flags[p];
    unsigned m = 0;
    for (int k = 0; k < 6; k++) {
        const bool r1 = s1[k] != 0.0, r2 = s2[k] != 0.0, pd = (fl >> k) & 1u;
        if (r2) m |= 1u << k;
        if (!pd && !r1 && r2) m |= 1u << (8 + k);
        if (r1) m |= 1u << (16 + k);
    }
    rvm[t] = m;
}

// initial_output (src/analyses.jl:2332-2390): the solved initial-condition vector -> DynamicSystem states (in place) and the point
// rates udot, θdot, Vdot, Ωdot ([b][NP][12]) that seed the Newmark march.  One thread per (instance, point).
__global__ void k_initial_output(double* __restrict__ x, const double* __restrict__ bcS, const double* __restrict__ icS, const unsigned* __restrict__ rvm,
                                 const unsigned short* __restrict__ flags, const double* __restrict__ pxS, const double* __restrict__ motion,
                                 double* __restrict__ rates, int NP, int n, int64_t nbatch, DynArgs body) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * NP) return;
    const int64_t b = t / NP; const int p = (int)(t % NP);
    const unsigned fl = flags[p], rv = rvm[t];
    const double* bc = bcS + ((size_t)b * 18) * NP + p;
    const double* ic = icS + ((size_t)b * 18) * NP + p;
    double* xp = x + b * n + 18 * p;
    const double* mo = motion + 12 * b;
    const v3 vb = mk(mo[0], mo[1], mo[2]), wb = mk(mo[3], mo[4], mo[5]);
    v3 ab, alb;
    body_accelerations(body, x + b * n, mo, ab, alb);          // system.jl:167-182 (analyses.jl:2348)
    double ut[6], rt[6];
    for (int k = 0; k < 6; k++) {
        const bool pd = (fl >> k) & 1u, r = (rv >> k) & 1u;
        ut[k] = pd ? bc[(size_t)k * NP] : (r ? ic[(size_t)k * NP] : xp[k]);
        rt[k] = pd ? 0.0 : (r ? xp[k] : ic[(size_t)(12 + k) * NP]);
    }
    const v3 u = mk(ut[0], ut[1], ut[2]);
    const v3 udot = mk(xp[6], xp[7], xp[8]), thdot = mk(xp[9], xp[10], xp[11]);
    const v3 dx = ldv(pxS + ((size_t)b * 3) * NP + p, NP);
    const v3 V = ldv(ic + (size_t)6 * NP, NP) + vb + cross(wb, dx + u), Om = ldv(ic + (size_t)9 * NP, NP) + wb;
    const v3 Vd = mk(rt[0], rt[1], rt[2]) + ab + cross(alb, dx + u) + cross(wb, udot), Od = mk(rt[3], rt[4], rt[5]) + alb;
    double* ra = rates + ((size_t)b * NP + p) * 12;
    for (int k = 0; k < 6; k++) if (!((fl >> k) & 1u)) xp[k] = ut[k];          // displacement where it is a state; loads stay
    ra[0] = (fl & 1) ? 0.0 : udot.x; ra[1] = (fl & 2) ? 0.0 : udot.y; ra[2] = (fl & 4) ? 0.0 : udot.z;
    ra[3] = (fl & 8) ? 0.0 : thdot.x; ra[4] = (fl & 16) ? 0.0 : thdot.y; ra[5] = (fl & 32) ? 0.0 : thdot.z;
    ra[6] = Vd.x; ra[7] = Vd.y; ra[8] = Vd.z; ra[9] = Od.x; ra[10] = Od.y; ra[11] = Od.z;
    xp[6] = V.x; xp[7] = V.y; xp[8] = V.z; xp[9] = Om.x; xp[10] = Om.y; xp[11] = Om.z;
}

// AssemblyState extraction on the device (src/output.jl:146-203 extract_point_state, :317-403 extract_element_state): one thread per
// (instance, station) writes the PointState of point i and the ElementState of element i as 30 doubles each, in the field order of the
// reference's structs (u, udot, theta, thetadot, V, Vdot, Omega, Omegadot, F, M), so that a Julia caller can reinterpret the arrays as
// Vector{PointState{Float64}} / Vector{ElementState{Float64}}.  theta is returned as Wiener-Milenkovic parameters (θ · scaling,
// output.jl:166-167).  rates ([b][NP][12]: udot, θdot, Vdot, Ωdot per point, the dx of newmark_output) may be null: zeros for a
// StaticSystem, NaN for a DynamicSystem (the reference's FillArrays.Fill(NaN) for `AssemblyState(x, system, ...)`).
struct StateArgs {
    int kind, N, n; int64_t nbatch;
    const double* x; const double* bcS; const unsigned short* flags; const double* fs; const double* rates;
    double* points; double* elems;          // [b][NP][30], [b][N][30]
};
__global__ void __launch_bounds__(128) k_assembly_state(StateArgs A) {
    const int N = A.N, NP = N + 1, stride = A.kind == 0 ? 12 : 18;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= A.nbatch * NP) return;
    const int64_t b = t / NP; const int i = (int)(t % NP);
    const double* x = A.x + b * A.n;
    const double fs = A.fs[b];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    struct P { v3 u, th, udot, thdot, V, Om, Vd, Od; };
    auto point = [&](int p) {
        P r;
        const unsigned fl = A.flags[p];
        const double* bc = A.bcS + ((size_t)b * 18) * NP + p;
        const double* xp = x + (size_t)stride * p;
        r.u = mk((fl & 1) ? bc[0] : xp[0], (fl & 2) ? bc[NP] : xp[1], (fl & 4) ? bc[2 * NP] : xp[2]);
        r.th = mk((fl & 8) ? bc[3 * NP] : xp[3], (fl & 16) ? bc[4 * NP] : xp[4], (fl & 32) ? bc[5 * NP] : xp[5]);
        r.V = r.Om = mk(0, 0, 0);
        if (A.kind == 1) { r.V = mk(xp[6], xp[7], xp[8]); r.Om = mk(xp[9], xp[10], xp[11]); }
        const double dflt = A.kind == 0 ? 0.0 : qnan;
        double d[12];
        for (int k = 0; k < 12; k++) d[k] = A.rates ? A.rates[((size_t)b * NP + p) * 12 + k] : dflt;
        // point_displacement_rates (point.jl:265-281): zero where the displacement is prescribed
        for (int k = 0; k < 6; k++) if ((fl >> k) & 1u) d[k] = 0.0;
        r.udot = mk(d[0], d[1], d[2]); r.thdot = mk(d[3], d[4], d[5]); r.Vd = mk(d[6], d[7], d[8]); r.Od = mk(d[9], d[10], d[11]);
        return r;
    };
    auto put = [&](double* o, v3 u, v3 ud, v3 th, v3 thd, v3 V, v3 Vd, v3 Om, v3 Od, v3 F, v3 M) {
        const double s = rot_scaling(th);
        const v3 a[10] = {u, ud, s * th, thd, V, Vd, Om, Od, F, M};
        for (int k = 0; k < 10; k++) { o[3 * k] = a[k].x; o[3 * k + 1] = a[k].y; o[3 * k + 2] = a[k].z; }
    };
    const P p1 = point(i);
    {
        // point_loads (point.jl:20-39): prescribed load + follower part rotated by C(θ)', or the state variable times force_scaling
        const unsigned fl = A.flags[i];
        const double* bc = A.bcS + ((size_t)b * 18) * NP + i;
        const double* xp = x + (size_t)stride * i;
        const m3 C = rot_C(p1.th);
        const v3 F = ldv(bc + 6 * NP, NP) + tmul(C, ldv(bc + 12 * NP, NP)), M = ldv(bc + 9 * NP, NP) + tmul(C, ldv(bc + 15 * NP, NP));
        const v3 Fo = mk((fl & 64) ? F.x : xp[0] * fs, (fl & 128) ? F.y : xp[1] * fs, (fl & 256) ? F.z : xp[2] * fs);
        const v3 Mo = mk((fl & 512) ? M.x : xp[3] * fs, (fl & 1024) ? M.y : xp[4] * fs, (fl & 2048) ? M.z : xp[5] * fs);
        put(A.points + ((size_t)b * NP + i) * 30, p1.u, p1.udot, p1.th, p1.thdot, p1.V, p1.Vd, p1.Om, p1.Od, Fo, Mo);
    }
    if (i < N) {
        const P p2 = point(i + 1);
        const double* xe = x + (size_t)stride * i + (stride - 6);
        const v3 F = mk(xe[0] * fs, xe[1] * fs, xe[2] * fs), M = mk(xe[3] * fs, xe[4] * fs, xe[5] * fs);
        put(A.elems + ((size_t)b * N + i) * 30, 0.5 * (p1.u + p2.u), 0.5 * (p1.udot + p2.udot), 0.5 * (p1.th + p2.th), 0.5 * (p1.thdot + p2.thdot),
            0.5 * (p1.V + p2.V), 0.5 * (p1.Vd + p2.Vd), 0.5 * (p1.Om + p2.Om), 0.5 * (p1.Od + p2.Od), F, M);
    }
}

// xhist[b][slot][:] = x[b][:]
__global__ void k_save_state(const double* __restrict__ x, double* __restrict__ xhist, int n, int64_t nsave, int64_t slot, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * n) return;
    int64_t b = t / n; int c = (int)(t % n);
    xhist[((size_t)b * nsave + slot) * n + c] = x[t];
}

struct SolveArgs {
    int nb, n; int64_t nbatch;
    const double* J; const double* F; const double* x; double* U; double* p; double* xt;
    NewtonState ns;
    double maxstep; int linear;
    int direction_only;   // parity entry point: just p, no state change
    double sign;          // sol = sign * (J \ F): -1 for the Newton direction, +1 for the eigen operator
    // Woodbury-corrected steps (body-acceleration states): F / p above are the right-hand side / solution of ONE band solve,
    // these are the Newton residual and direction the line search works with
    const double* Fnewton; double* pnewton;
    // singular Jacobian: with `fallback` the instance is parked in ST_SINGULAR for k_singular_fallback (counter: unfinished instances of
    // the round, singular_seen: sticky flag the host reads with the counter); without it the instance fails (FAIL_SINGULAR)
    int fallback; int* counter; int* singular_seen;
    // optional record of the factorisation for further right-hand sides (tensor-core kernel only): [nbatch][l_stride] doubles
    double* Lst; size_t l_stride;
};
__device__ __forceinline__ void mark_singular(const SolveArgs& A, int64_t b) {
    if (A.fallback) {
        A.ns.status[b] = ST_SINGULAR;
        if (A.counter) atomicAdd(A.counter, 1);
        if (A.singular_seen) *A.singular_seen = 1;
    } else { A.ns.status[b] = ST_FAILED; A.ns.failcode[b] = FAIL_SINGULAR; }
}
__device__ void linesearch_start(const SolveArgs& A, int64_t b);

// one warp: factorise + solve instance b, then start the line search (|p|_inf, φ0, dφ0, first trial point)
template <int KLB, int KUB, int MMA = 0> __device__ void factor_instance(const SolveArgs& A, int64_t b, double* sm) {
    using S = LuShape<KLB, KUB>;
    const int lane = threadIdx.x & 31;
    const size_t n = A.n;
    const double* F = A.F + b * n;
    double* p = A.p + b * n;
    int bad;
    if constexpr (MMA) {       // window in DMMA fragments, 8-column elimination blocks (gxb_lu_mma.cuh); U rows are wider, n padded to 8
        using M = MmaShape<KLB, KUB>;
        bad = warp_band_solve_mma<KLB, KUB>(A.J + b * n * S::W, F, A.U + b * ((n + 7) & ~(size_t)7) * M::UW, p, A.nb, A.sign, sm,
                                            A.Lst ? A.Lst + b * A.l_stride : nullptr);
    } else {
        bad = warp_band_solve<KLB, KUB>(A.J + b * n * S::W, F, A.U + b * n * S::UW, p, A.nb, A.sign, sm);
    }
    __syncwarp();
    if (A.direction_only == 1) { if (bad && lane == 0) A.ns.failcode[b] = FAIL_SINGULAR; return; }     // the caller reads the flag (eigen path)
    if (bad) { if (lane == 0) { if (A.direction_only) { A.ns.status[b] = ST_FAILED; A.ns.failcode[b] = FAIL_SINGULAR; } else mark_singular(A, b); } return; }
    if (A.direction_only) return;        // 2: one of several solves of a Woodbury-corrected step (k_woodbury finishes it)
    linesearch_start(A, b);
}
// |p|_inf, φ0 = ½ F'F, dφ0 = g'p with g = J'F and J p = -F  =>  dφ0 = -F'F; first trial point.  One warp; p is complete.
__device__ void linesearch_start(const SolveArgs& A, int64_t b) {
    const int lane = threadIdx.x & 31;
    const size_t n = A.n;
    const double* F = A.Fnewton ? A.Fnewton + b * n : A.F + b * n;
    double* p = A.pnewton ? A.pnewton + b * n : A.p + b * n;
    {
    double pm = 0.0, ff = 0.0;
    for (int c = lane; c < (int)n; c += 32) { pm = fmax(pm, fabs(p[c])); double f = F[c]; ff = fma(f, f, ff); }
    for (int o = 16; o; o >>= 1) { pm = fmax(pm, __shfl_xor_sync(GXB_FULL, pm, o)); ff += __shfl_xor_sync(GXB_FULL, ff, o); }
    const double a0 = A.linear ? 1.0 : fmin(1.0, A.maxstep / pm);
    const double* x = A.x + b * n; double* xt = A.xt + b * n;
    for (int c = lane; c < (int)n; c += 32) xt[c] = x[c] + a0 * p[c];
    if (lane == 0) {
        A.ns.a1[b] = a0; A.ns.a2[b] = a0; A.ns.phi0[b] = 0.5 * ff; A.ns.dphi0[b] = -ff; A.ns.phx0[b] = 0.5 * ff;
        A.ns.ls_k[b] = 0; A.ns.ls_finite[b] = 0; A.ns.status[b] = ST_TRIAL;
    }
    }
}
template <int KLB, int KUB> __global__ void __launch_bounds__(128, (KLB == 2 ? GXB_LU_MINBLOCKS : 2)) k_factor_solve(SolveArgs A) {
    using S = LuShape<KLB, KUB>;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.nbatch) return;
    if (A.direction_only != 1 && A.ns.status[b] != ST_SOLVE) return;
    factor_instance<KLB, KUB>(A, b, smem + (size_t)warp * S::SM_DOUBLES);
}
// the same, with the tensor-core factorisation of gxb_lu_mma.cuh
template <int KLB, int KUB> __global__ void __launch_bounds__(128, (KLB == 2 ? GXB_LU_MINBLOCKS : GXB_LU_MINBLOCKS_DYN)) k_factor_solve_mma(SolveArgs A) {
    using S = MmaShape<KLB, KUB>;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.nbatch) return;
    if (A.direction_only != 1 && A.ns.status[b] != ST_SOLVE) return;
    factor_instance<KLB, KUB, 1>(A, b, smem + (size_t)warp * S::SM_DOUBLES);
}

// |p|_inf, φ0, dφ0 and the first trial point by the whole CTA (the partitioned factorisation runs P warps per instance)
__device__ void linesearch_start_cta(const SolveArgs& A, int64_t b) {
    const size_t n = A.n;
    const double* F = A.F + b * n; const double* p = A.p + b * n;
    __shared__ double s_pm[8], s_ff[8];
    double pm = 0.0, ff = 0.0;
    for (int c = threadIdx.x; c < (int)n; c += blockDim.x) { pm = fmax(pm, fabs(p[c])); const double f = F[c]; ff = fma(f, f, ff); }
    for (int o = 16; o; o >>= 1) { pm = fmax(pm, __shfl_xor_sync(GXB_FULL, pm, o)); ff += __shfl_xor_sync(GXB_FULL, ff, o); }
    if ((threadIdx.x & 31) == 0) { s_pm[threadIdx.x >> 5] = pm; s_ff[threadIdx.x >> 5] = ff; }
    __syncthreads();
    pm = 0.0; ff = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); k++) { pm = fmax(pm, s_pm[k]); ff += s_ff[k]; }
    const double a0 = A.linear ? 1.0 : fmin(1.0, A.maxstep / pm);
    const double* x = A.x + b * n; double* xt = A.xt + b * n;
    for (int c = threadIdx.x; c < (int)n; c += blockDim.x) xt[c] = x[c] + a0 * p[c];
    if (threadIdx.x == 0) {
        A.ns.a1[b] = a0; A.ns.a2[b] = a0; A.ns.phi0[b] = 0.5 * ff; A.ns.dphi0[b] = -ff; A.ns.phx0[b] = 0.5 * ff;
        A.ns.ls_k[b] = 0; A.ns.ls_finite[b] = 0; A.ns.status[b] = ST_TRIAL;
    }
}
// two-sided factorisation (gxb_lu_two.cuh): one CTA of two warps per instance, top-down + bottom-up sweeps meeting at row block R
template <int KLB, int KUB, int AB, int BB> __global__ void __launch_bounds__(64) k_factor_solve_two(SolveArgs A, int R) {
    using S = LuShape<KLB, KUB>;
    using T2 = TwoShape<KLB, KUB, AB, BB>;
    extern __shared__ __align__(16) double smem[];
    const int64_t b = blockIdx.x;
    if (A.direction_only != 1 && A.ns.status[b] != ST_SOLVE) return;
    const size_t n = A.n;
    const int bad = cta_two_sided_solve<KLB, KUB, AB, BB>(A.J + b * n * S::W, A.F + b * n, A.U + b * n * T2::UW, A.p + b * n, A.nb, R, A.sign, smem);
    if (A.direction_only == 1) { if (bad && threadIdx.x == 0) A.ns.failcode[b] = FAIL_SINGULAR; return; }
    if (bad) { if (threadIdx.x == 0) { if (A.direction_only) { A.ns.status[b] = ST_FAILED; A.ns.failcode[b] = FAIL_SINGULAR; } else mark_singular(A, b); } return; }
    if (A.direction_only) return;
    linesearch_start_cta(A, b);
}
// further right-hand sides with the factorisation recorded by k_factor_solve_mma (A.Lst): p = sign * (J \ F)
template <int KLB, int KUB> __global__ void __launch_bounds__(128) k_resolve_mma(SolveArgs A) {
    using S = MmaShape<KLB, KUB>;
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.nbatch) return;
    const size_t n = A.n;
    warp_band_resolve_mma<KLB, KUB>(A.Lst + b * A.l_stride, A.F + b * n, A.U + b * ((n + 7) & ~(size_t)7) * S::UW, A.p + b * n, A.nb, A.sign,
                                    smem + (size_t)warp * S::SM_DOUBLES);
}
// partitioned factorisation (gxb_lu_part.cuh): one CTA of PT warps per instance, static systems; for batches too small to fill the GPU
// with one warp per instance
// register budget: 12 warps per SM -> 170 registers per thread (the lane-per-row window + spike are 96 registers of state; capped at
// 128 the block step ran at half speed: 6400 instead of 3250 cycles, profiles/r2_part_timing.txt)
#ifndef GXB_PART_MINB
#define GXB_PART_MINB(PT) ((PT) >= 8 ? 1 : 12 / (PT))
#endif
template <int PT> __global__ void __launch_bounds__(32 * PT, GXB_PART_MINB(PT)) k_factor_solve_part(SolveArgs A, PartPlan plan) {
    extern __shared__ __align__(16) double smem[];
    const int64_t b = blockIdx.x;
    if (A.direction_only != 1 && A.ns.status[b] != ST_SOLVE) return;
    const size_t n = A.n;
    const int bad = cta_part_solve(A.J + b * n * PartShape::W, A.F + b * n, A.U + b * n * PartShape::UW, A.p + b * n, plan, A.sign, smem);
    if (A.direction_only == 1) return;
    if (bad) { if (threadIdx.x == 0) { if (A.direction_only) { A.ns.status[b] = ST_FAILED; A.ns.failcode[b] = FAIL_SINGULAR; } else mark_singular(A, b); } return; }
    if (A.direction_only) return;
    linesearch_start_cta(A, b);
}

// NLsolve's singular-Jacobian fallback (newton.jl, restated in SURVEY App. C): when the factorisation
// of J fails,   p = -(J'J + λ I) \ (J'F),   λ = 1e6 sqrt(n eps) ||J'J||_1.   A rare path -- clarity over speed: one CTA per parked
// instance; M = J'J + λI (symmetric positive definite, cond <= ~3 because λ ~ 0.5 ||J'J||) is built in lower-band storage inside the
// instance's U scratch and factorised as L D L' without pivoting, column by column.  Then the line search starts with the TRUE
// dφ0 = g'p, g = J'F (J p != -F here).   Row r of the row-block storage holds columns 6 (r/6 - KLB) .. + W.
struct FallbackArgs { SolveArgs S; int W, KLB; size_t u_stride; };      // u_stride: doubles of U scratch per instance
__global__ void __launch_bounds__(256) k_singular_fallback(FallbackArgs Fa) {
    const SolveArgs& A = Fa.S;
    const int64_t b = blockIdx.x;
    if (A.ns.status[b] != ST_SINGULAR) return;
    const int n = A.n, W = Fa.W, KLB = Fa.KLB, nb = n / 6;
    const int KUB = W / 6 - 1 - KLB;
    const int m = 6 * (KLB + KUB) + 5;                       // half bandwidth of J'J
    const double* J = A.J + (size_t)b * n * W;
    const double* F = A.F + (size_t)b * n;
    double* Mb = A.U + (size_t)b * Fa.u_stride;              // [n][m+1]: Mb[j][d] = M[j+d][j]
    double* g = A.xt + (size_t)b * n;                        // J'F (xt is rewritten by the line-search start below)
    double* p = A.p + (size_t)b * n;
    __shared__ double red[8]; __shared__ double s_bc;
    auto block_reduce = [&](double v, bool is_max) {
        for (int o = 16; o; o >>= 1) { const double w = __shfl_xor_sync(GXB_FULL, v, o); v = is_max ? fmax(v, w) : v + w; }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) { double t = red[0]; for (int k = 1; k < (int)(blockDim.x >> 5); k++) t = is_max ? fmax(t, red[k]) : t + red[k]; s_bc = t; }
        __syncthreads();
        return s_bc;
    };
    auto Jat = [&](int r, int c) -> double {                 // J[r][c] or 0 outside the stored window
        const int j = c - 6 * (r / 6 - KLB);
        return (j >= 0 && j < W) ? J[(size_t)r * W + j] : 0.0;
    };
    // M = J'J (lower band) and g = J'F: column j is touched by the rows of blocks j/6 - KUB .. j/6 + KLB
    for (int t = threadIdx.x; t < n * (m + 1); t += blockDim.x) {
        const int j = t / (m + 1), d = t - j * (m + 1), i = j + d;
        double acc = 0.0;
        if (i < n) {
            const int B0 = max(0, i / 6 - KUB), B1 = min(nb - 1, j / 6 + KLB);
            for (int r = 6 * B0; r < 6 * (B1 + 1); r++) acc = fma(Jat(r, i), Jat(r, j), acc);
        }
        Mb[t] = acc;
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int B0 = max(0, j / 6 - KUB), B1 = min(nb - 1, j / 6 + KLB);
        double acc = 0.0;
        for (int r = 6 * B0; r < 6 * (B1 + 1); r++) acc = fma(Jat(r, j), F[r], acc);
        g[j] = acc;
    }
    __syncthreads();
    // ||J'J||_1 = max column sum (symmetric: lower band + mirrored upper band)
    double cmax = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double cs = 0.0;
        for (int d = 0; d <= m && j + d < n; d++) cs += fabs(Mb[(size_t)j * (m + 1) + d]);
        for (int d = 1; d <= m && j - d >= 0; d++) cs += fabs(Mb[(size_t)(j - d) * (m + 1) + d]);
        cmax = fmax(cmax, cs);
    }
    const double lambda = 1e6 * sqrt((double)n * 2.220446049250313e-16) * block_reduce(cmax, true);
    for (int j = threadIdx.x; j < n; j += blockDim.x) Mb[(size_t)j * (m + 1)] += lambda;
    __syncthreads();
    // L D L' (right-looking): after step k, Mb[k][0] = D_k and Mb[k][d] = L[k+d][k]
    for (int k = 0; k < n; k++) {
        const double* ck = Mb + (size_t)k * (m + 1);
        const double dk = ck[0];
        const int mk = min(m, n - 1 - k);
        for (int t = threadIdx.x; t < mk * (mk + 1) / 2; t += blockDim.x) {
            // pairs 1 <= dj <= di <= mk, enumerated row by row
            int di = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
            while ((di + 1) * (di + 2) / 2 <= t) di++;
            while (di * (di + 1) / 2 > t) di--;
            const int dj = t - di * (di + 1) / 2 + 1;
            const int dI = di + 1;
            Mb[(size_t)(k + dj) * (m + 1) + (dI - dj)] -= ck[dI] * ck[dj] / dk;
        }
        __syncthreads();
        for (int d = 1 + threadIdx.x; d <= mk; d += blockDim.x) Mb[(size_t)k * (m + 1) + d] = ck[d] / dk;
        __syncthreads();
    }
    // solve M y = g in p: forward (L), diagonal, backward (L')
    for (int j = threadIdx.x; j < n; j += blockDim.x) p[j] = g[j];
    __syncthreads();
    for (int k = 0; k < n; k++) {
        const double yk = p[k];
        const int mk = min(m, n - 1 - k);
        for (int d = 1 + threadIdx.x; d <= mk; d += blockDim.x) p[k + d] -= Mb[(size_t)k * (m + 1) + d] * yk;
        __syncthreads();
    }
    for (int j = threadIdx.x; j < n; j += blockDim.x) p[j] /= Mb[(size_t)j * (m + 1)];
    __syncthreads();
    for (int k = n - 1; k >= 0; k--) {
        const int mk = min(m, n - 1 - k);
        double acc = 0.0;
        for (int d = 1 + threadIdx.x; d <= mk; d += blockDim.x) acc += Mb[(size_t)k * (m + 1) + d] * p[k + d];
        acc = block_reduce(acc, false);
        if (threadIdx.x == 0) p[k] -= acc;
        __syncthreads();
    }
    // p = -(M \ g); line-search start with dφ0 = g'p
    double pm = 0.0, ff = 0.0, gp = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const double pc = -p[c];
        p[c] = pc;
        pm = fmax(pm, fabs(pc)); ff = fma(F[c], F[c], ff); gp = fma(g[c], pc, gp);
    }
    pm = block_reduce(pm, true); ff = block_reduce(ff, false); gp = block_reduce(gp, false);
    const double a0 = A.linear ? 1.0 : fmin(1.0, A.maxstep / pm);
    const double* x = A.x + (size_t)b * n; double* xt = A.xt + (size_t)b * n;
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += blockDim.x) xt[c] = x[c] + a0 * p[c];
    if (threadIdx.x == 0) {
        A.ns.a1[b] = a0; A.ns.a2[b] = a0; A.ns.phi0[b] = 0.5 * ff; A.ns.dphi0[b] = gp; A.ns.phx0[b] = 0.5 * ff;
        A.ns.ls_k[b] = 0; A.ns.ls_finite[b] = 0; A.ns.status[b] = ST_TRIAL;
    }
}

// opts.exact_dphi0: dφ0 = g'p with g = J'F as NLsolve evaluates it (newton.jl), i.e. F'(J p), instead of -F'F (equal in
// exact arithmetic because J p = -F; they differ by the residual of the linear solve).  One CTA per instance that has just started a
// line search (status TRIAL, ls_k = 0, no evaluation yet); J is still intact in the row-block storage.
struct DphiArgs { SolveArgs S; int W, KLB; };
__global__ void __launch_bounds__(128) k_exact_dphi0(DphiArgs D) {
    const SolveArgs& A = D.S;
    const int64_t b = blockIdx.x;
    if (A.ns.status[b] != ST_TRIAL || A.ns.ls_k[b] != 0 || A.ns.ls_finite[b] != 0) return;
    const int n = A.n, W = D.W, KLB = D.KLB;
    const double* J = A.J + (size_t)b * n * W; const double* F = A.F + (size_t)b * n; const double* p = A.p + (size_t)b * n;
    double acc = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int c0 = 6 * (r / 6 - KLB);
        double jp = 0.0;
        for (int j = 0; j < W; j++) { const int c = c0 + j; if (c >= 0 && c < n) jp = fma(J[(size_t)r * W + j], p[c], jp); }
        acc = fma(F[r], jp, acc);
    }
    __shared__ double red[4];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(GXB_FULL, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) A.ns.dphi0[b] = (red[0] + red[1]) + (red[2] + red[3]);
}

// Woodbury correction of a Newton step whose Jacobian is J = B + sum_i D_i e_{c_i}' (body-acceleration columns, see k_body_columns):
// with y = B^-1 F (in p) and z_i = B^-1 D_i,   J^-1 F = y - Z (I + E'Z)^-1 E'y.   One warp per instance; the k x k system (k <= 6) is
// solved by lane 0 with partial pivoting.  Then the line search starts exactly as after an ordinary factorisation.
struct WoodArgs { SolveArgs S; const double* Z; int ibody[6]; };
__global__ void __launch_bounds__(128) k_woodbury(WoodArgs W) {
    const SolveArgs& A = W.S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (b >= A.nbatch) return;
    if (A.ns.status[b] != ST_SOLVE) return;
    const size_t n = A.n;
    double* p = A.p + b * n;
    int idx[6], k = 0;
    for (int i = 0; i < 6; i++) if (W.ibody[i] >= 0) idx[k++] = i;
    double a[6] = {0, 0, 0, 0, 0, 0};
    int bad = 0;
    if (lane == 0) {
        double M[6][7];
        for (int j = 0; j < k; j++) {
            const int cj = W.ibody[idx[j]];
            for (int i = 0; i < k; i++) M[j][i] = (i == j ? 1.0 : 0.0) + W.Z[((size_t)idx[i] * A.nbatch + b) * n + cj];
            M[j][k] = p[cj];
        }
        for (int col = 0; col < k && !bad; col++) {
            int piv = col; double best = fabs(M[col][col]);
            for (int r = col + 1; r < k; r++) if (fabs(M[r][col]) > best) { best = fabs(M[r][col]); piv = r; }
            if (!(best > 0.0) || !isfinite(best)) { bad = 1; break; }
            if (piv != col) for (int q = col; q <= k; q++) { double t = M[col][q]; M[col][q] = M[piv][q]; M[piv][q] = t; }
            for (int r = col + 1; r < k; r++) { const double f = M[r][col] / M[col][col]; for (int q = col; q <= k; q++) M[r][q] -= f * M[col][q]; }
        }
        if (!bad) for (int r = k - 1; r >= 0; r--) { double v = M[r][k]; for (int q = r + 1; q < k; q++) v -= M[r][q] * a[q]; a[r] = v / M[r][r]; }
    }
    bad = __shfl_sync(GXB_FULL, bad, 0);
    if (bad) { if (lane == 0) { A.ns.status[b] = ST_FAILED; A.ns.failcode[b] = FAIL_SINGULAR; } return; }
    for (int j = 0; j < 6; j++) a[j] = __shfl_sync(GXB_FULL, a[j], 0);
    for (int c = lane; c < (int)n; c += 32) {
        double v = p[c];
        for (int j = 0; j < k; j++) v = fma(-W.Z[((size_t)idx[j] * A.nbatch + b) * n + c], a[j], v);
        p[c] = -v;                       // Newton direction p = -J^-1 F
    }
    __syncwarp();
    linesearch_start(A, b);
}

struct UpdateArgs {
    int n; int mode;   // mode 0: initial evaluation, 1: line-search step, 2: linear (accept unconditionally, converged)
    double* x; double* xt; double* F; const double* Ft; const double* p;
    NewtonState ns;
    double ftol, c1, rho_hi, rho_lo; int64_t iterations; int ls_iterations;
    int* counter;
};

// Armijo test / accept / convergence for instance b by the whole CTA (any block size up to 256); returns the new status
__device__ int update_instance(const UpdateArgs& A, int b) {
    const size_t n = A.n;
    const double* Ft = A.Ft + b * n;
    // all per-instance scalars are read before the reduction barrier; thread 0 rewrites them only after it
    const double a1 = A.ns.a1[b], a2 = A.ns.a2[b], phi0 = A.ns.phi0[b], dphi0 = A.ns.dphi0[b], phx0 = A.ns.phx0[b];
    int k = A.ns.ls_k[b], nf = A.ns.ls_finite[b];
    int iters = A.ns.iters[b];
    __shared__ double s_sum[8], s_max[8]; __shared__ int s_nan[8];
    double ss = 0.0, mx = 0.0; int nan = 0;
    for (int c = threadIdx.x; c < (int)n; c += blockDim.x) { double f = Ft[c]; ss = fma(f, f, ss); mx = fmax(mx, fabs(f)); nan |= (f != f); }
    for (int o = 16; o; o >>= 1) { ss += __shfl_xor_sync(GXB_FULL, ss, o); mx = fmax(mx, __shfl_xor_sync(GXB_FULL, mx, o)); nan |= __shfl_xor_sync(GXB_FULL, nan, o); }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = ss; s_max[threadIdx.x >> 5] = mx; s_nan[threadIdx.x >> 5] = nan; }
    if (threadIdx.x < 8 && threadIdx.x >= (blockDim.x >> 5)) { s_sum[threadIdx.x] = 0.0; s_max[threadIdx.x] = 0.0; s_nan[threadIdx.x] = 0; }
    __syncthreads();
    ss = (s_sum[0] + s_sum[1]) + (s_sum[2] + s_sum[3]);
    mx = fmax(fmax(s_max[0], s_max[1]), fmax(s_max[2], s_max[3]));
    nan = s_nan[0] | s_nan[1] | s_nan[2] | s_nan[3];
    if (blockDim.x > 128) {      // persistent kernels run with up to 8 warps; the 4-warp order above is kept for bit-compatibility
        ss += (s_sum[4] + s_sum[5]) + (s_sum[6] + s_sum[7]);
        mx = fmax(mx, fmax(fmax(s_max[4], s_max[5]), fmax(s_max[6], s_max[7])));
        nan |= s_nan[4] | s_nan[5] | s_nan[6] | s_nan[7];
    }
    const double phx1 = 0.5 * ss;
    const bool finite = isfinite(ss) && !nan;
    double* x = A.x + b * n; double* xt = A.xt + b * n; double* F = A.F + b * n; const double* p = A.p + b * n;
    // every thread takes the same decision from the same scalars
    bool accept = false; int newstatus = ST_TRIAL; int failcode = FAIL_NONE; double a_new = 0.0;
    if (A.mode == 0) {
        accept = true;
        if (!finite) { newstatus = ST_FAILED; failcode = FAIL_NONFINITE; }     // NLsolve check_isfinite
        else newstatus = (mx <= A.ftol) ? ST_DONE : ST_SOLVE;
        if (newstatus == ST_SOLVE && A.iterations <= 0) { newstatus = ST_FAILED; failcode = FAIL_MAXITER; }
    } else if (A.mode == 2) {
        accept = true; newstatus = ST_DONE; iters = 1;
    } else {
        double n_a1 = a1, n_a2 = a2, n_phx0 = phx0;
        if (k == 0 && !isfinite(phx1) && nf < 52) {                 // back off until finite (LineSearches iterfinite loop)
            nf++; n_a1 = a2; n_a2 = a2 / 2;
        } else if (phx1 > phi0 + A.c1 * a2 * dphi0) {               // Armijo fails -> interpolate
            k++;
            if (k > A.ls_iterations) { newstatus = ST_FAILED; failcode = FAIL_LINESEARCH; }
            double at;
            if (k == 1) at = -(dphi0 * a2 * a2) / (2 * (phx1 - phi0 - dphi0 * a2));
            else {
                double dv = 1.0 / (a1 * a1 * a2 * a2 * (a2 - a1));
                double ca = (a1 * a1 * (phx1 - phi0 - dphi0 * a2) - a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * dv;
                double cb = (-a1 * a1 * a1 * (phx1 - phi0 - dphi0 * a2) + a2 * a2 * a2 * (phx0 - phi0 - dphi0 * a1)) * dv;
                if (fabs(ca) <= 2.220446049250313e-16) at = dphi0 / (2 * cb);
                else { double d = fmax(cb * cb - 3 * ca * dphi0, 0.0); at = (-cb + sqrt(d)) / (3 * ca); }
            }
            n_a1 = a2;
            double hi = a2 * A.rho_hi, lo = a2 * A.rho_lo;
            at = (at != at) ? hi : fmin(at, hi);
            n_a2 = fmax(at, lo);
            n_phx0 = phx1;
        } else {
            accept = true;
        }
        if (!accept && newstatus == ST_TRIAL) {
            a_new = n_a2;
            if (threadIdx.x == 0) { A.ns.a1[b] = n_a1; A.ns.a2[b] = n_a2; A.ns.phx0[b] = n_phx0; A.ns.ls_k[b] = k; A.ns.ls_finite[b] = nf; }
        }
        if (accept) {
            iters++;
            const bool fconv = mx <= A.ftol;                         // maximum(abs, F) <= ftol ; NaN compares false
            bool stopped = nan != 0;
            bool xzero = true;                                       // x-convergence: ||x - xold||_inf <= xtol = 0
            for (int c = threadIdx.x; c < (int)n; c += blockDim.x) { if (xt[c] != x[c]) xzero = false; if (xt[c] != xt[c]) stopped = true; }
            xzero = __syncthreads_and(xzero); stopped = __syncthreads_or(stopped);
            if (fconv) newstatus = ST_DONE;
            else if (stopped) { newstatus = ST_FAILED; failcode = FAIL_NONFINITE; }
            else if (xzero) { newstatus = ST_FAILED; failcode = FAIL_LINESEARCH; }
            else if (iters >= A.iterations) { newstatus = ST_FAILED; failcode = FAIL_MAXITER; }
            else newstatus = ST_SOLVE;
        }
    }
    if (accept) {
        for (int c = threadIdx.x; c < (int)n; c += blockDim.x) { x[c] = xt[c]; F[c] = Ft[c]; }
        if (threadIdx.x == 0) { A.ns.finf[b] = mx; A.ns.iters[b] = iters; }
    } else if (newstatus == ST_TRIAL) {
        for (int c = threadIdx.x; c < (int)n; c += blockDim.x) xt[c] = x[c] + a_new * p[c];
    }
    if (threadIdx.x == 0) {
        A.ns.fcalls[b] += 1;
        A.ns.status[b] = newstatus;
        if (failcode) A.ns.failcode[b] = failcode;
        if (A.counter && (newstatus == ST_SOLVE || newstatus == ST_TRIAL)) atomicAdd(A.counter, 1);
    }
    return newstatus;
}
__global__ void __launch_bounds__(128) k_update(UpdateArgs A) {
    const int b = blockIdx.x;
    if (A.ns.status[b] != ST_TRIAL) return;
    update_instance(A, b);
}
// Residual at the trial point AND the Armijo / convergence update in one launch (static chains of up to 127 elements): the CTA that
// evaluated Ft = r(xt) of its instance applies update_instance to it right away -- one launch per Newton round less (the rounds of a
// small batch are latency-bound: 25 us per launch at 512 instances).
__global__ void __launch_bounds__(128, 3) k_residual_update_128(StaticArgs A, UpdateArgs Up) {
    extern __shared__ __align__(16) double asm_smem[];
    const int b = blockIdx.x;
    if (A.status[b] != ST_TRIAL) return;
    static_station<false>(A, b, threadIdx.x, asm_smem);
    __threadfence_block();
    __syncthreads();                 // Ft of this instance is complete and visible to the whole CTA
    update_instance(Up, b);
}

__global__ void k_sum_counts(const int* __restrict__ iters, const int* __restrict__ fcalls, long long* __restrict__ tot, int64_t n) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    long long a = t < n ? iters[t] : 0, b = t < n ? fcalls[t] : 0;
    for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(GXB_FULL, a, o); b += __shfl_xor_sync(GXB_FULL, b, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(reinterpret_cast<unsigned long long*>(tot), (unsigned long long)a); atomicAdd(reinterpret_cast<unsigned long long*>(tot) + 1, (unsigned long long)b); }
}
__global__ void k_fill_int(int* p, int v, int64_t n) { int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; if (t < n) p[t] = v; }

// Jblk[b][k][3i+j] = J[brow[k]+i, bcol[k]+j] read back from the row-block storage
__global__ void k_extract_blocks(const double* __restrict__ Jrow, const int64_t* __restrict__ brow, const int64_t* __restrict__ bcol,
                                 double* __restrict__ out, int64_t nblk, int n, int W, int KLB, int64_t nbatch) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * nblk * 9) return;
    int e = (int)(t % 9); int64_t q = t / 9; int64_t k = q % nblk; int64_t b = q / nblk;
    int64_t r = brow[k] + e / 3, c = bcol[k] + e % 3;
    int64_t slot = c - 6 * (r / 6 - KLB);
    out[t] = (slot >= 0 && slot < W) ? Jrow[((size_t)b * n + r) * W + slot] : 0.0;
}


// ------------------------------------------------------------------------------------------------------------------
// Mass matrix M = ∂r/∂ẋ in the same row-block storage as the Jacobian (system_mass_matrix!, src/system.jl:961-995;
// point.jl:1356-1384, 2033-2061; element.jl:1408-1441, 2370-2411, 2903-2953).  One thread per station; a thread evaluates the
// frames of element i-1 and element i itself (only C(θ) is needed), so no exchange between threads is required.
struct MassArgs { StaticArgs s; double* Mrow; };
__global__ void k_mass_matrix(MassArgs A) {
    const StaticArgs& S = A.s;
    const int b = blockIdx.x, i = threadIdx.x, N = S.N, NP = N + 1;
    if (i > N) return;
    const double rfs = 1.0 / S.fs[b];
    const double* x = S.x + (size_t)b * S.n;
    double* Mr = A.Mrow + ((size_t)b * S.n + 18 * i) * GXB_DROWW;
    auto theta = [&](int p) {
        const unsigned fl = S.flags[p];
        const double* bc = S.bcS + ((size_t)b * 18) * NP + p;
        const double* xp = x + 18 * p;
        return mk((fl & 8) ? bc[3 * NP] : xp[3], (fl & 16) ? bc[4 * NP] : xp[4], (fl & 32) ? bc[5 * NP] : xp[5]);
    };
    const v3 thi = theta(i);
    m3 FV[3], FO[3], MV[3], MO[3];       // column points i-1, i, i+1
    for (int k = 0; k < 3; k++) { FV[k] = zero33(); FO[k] = zero33(); MV[k] = zero33(); MO[k] = zero33(); }
    if (S.pmS) {
        const double* pm = S.pmS + ((size_t)b * 36) * NP + i;
        const m3 C = rot_C(thi);
        FV[1] = rfs * tmul(C, ldm(pm, NP) * C); FO[1] = rfs * tmul(C, ldm(pm + 9 * NP, NP) * C);
        MV[1] = rfs * tmul(C, ldm(pm + 18 * NP, NP) * C); MO[1] = rfs * tmul(C, ldm(pm + 27 * NP, NP) * C);
    }
    for (int side = 0; side < 2; side++) {       // element i-1 (this point is its stop), then element i (this point is its start)
        const int e = i - 1 + side;
        if (e < 0 || e >= N) continue;
        const double* es = S.elemS + ((size_t)b * S.elem_nf) * N + e;
        const m3 Cab = ldm(es + N, N);
        const v3 th = 0.5 * (theta(e) + theta(e + 1));
        const m3 Am = tmul(rot_C(th), Cab), At = transpose(Am);
        const double q = 0.25 * rfs;
        const m3 a11 = q * ((Am * ldm(es + 46 * N, N)) * At), a12 = q * ((Am * ldm(es + 55 * N, N)) * At);
        const m3 a21 = q * ((Am * ldm(es + 64 * N, N)) * At), a22 = q * ((Am * ldm(es + 73 * N, N)) * At);
        const int k0 = side;                     // column points (e, e+1) = (i-1, i) or (i, i+1)
        FV[k0] = FV[k0] + a11; FV[k0 + 1] = FV[k0 + 1] + a11; FO[k0] = FO[k0] + a12; FO[k0 + 1] = FO[k0 + 1] + a12;
        MV[k0] = MV[k0] + a21; MV[k0 + 1] = MV[k0 + 1] + a21; MO[k0] = MO[k0] + a22; MO[k0 + 1] = MO[k0 + 1] + a22;
    }
    const bool two_d = S.two_d != 0;
    auto put = [&](int r0, int slot, const m3& B, bool angular) {
        for (int r = 0; r < 3; r++) {
            const bool zero = two_d && (angular ? r < 2 : r == 2);       // lmask = (1,1,0), amask = (0,0,1)
            for (int cc = 0; cc < 3; cc++) Mr[(size_t)(r0 + r) * GXB_DROWW + slot + cc] = zero ? 0.0 : B.a[3 * r + cc];
        }
    };
    for (int k = 0; k < 3; k++) {
        const int p = i - 1 + k;
        if (p < 0 || p > N) continue;
        const int sV = 6 + 18 * k, sO = 9 + 18 * k;     // slots of (V, Ω) of points i-1, i, i+1 in an equilibrium row
        put(0, sV, FV[k], false); put(0, sO, FO[k], false);
        put(3, sV, MV[k], true); put(3, sO, MO[k], true);
    }
    const unsigned fl = S.flags[i];
    m3 uu = zero33(), tt = zero33();
    uu.a[0] = (fl & 1) ? 0.0 : -1.0; uu.a[4] = (fl & 2) ? 0.0 : -1.0; uu.a[8] = (fl & 4) ? 0.0 : -1.0;
    tt.a[0] = (fl & 8) ? 0.0 : -1.0; tt.a[4] = (fl & 16) ? 0.0 : -1.0; tt.a[8] = (fl & 32) ? 0.0 : -1.0;
    put(6, 12, uu, false);       // rV rows: -u_u
    put(9, 15, tt, true);        // rΩ rows: -θ_θ
}

// y = M v in row-block storage: 8 threads per row (W = 48: six doubles each, W = 30: four), so that a warp reads 4 consecutive rows as
// one contiguous 1.5 KB piece; partial sums are combined by shuffles.  skip_stride > 0: rows r with r % skip_stride >= skip_from are
// structurally zero and are not read (the element rows of the mass matrix: 6 of every 18, system.jl:961-995).
__global__ void __launch_bounds__(256) k_rowblock_spmv(const double* __restrict__ Mrow, const double* __restrict__ v, double* __restrict__ y, int n, int W, int KLB,
                                                       int64_t nbatch, int64_t vstride, int skip_stride, int skip_from) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t rg = t >> 3;                       // global row
    const int part = (int)(t & 7);
    const bool valid = rg < nbatch * n;
    double acc = 0.0;
    int r = 0;
    if (valid) {
        const int64_t b = rg / n; r = (int)(rg % n);
        if (!(skip_stride > 0 && r % skip_stride >= skip_from)) {
            const int per = (W + 7) / 8;
            const double* row = Mrow + (size_t)rg * W;
            const double* vb = v + b * vstride;
            const int c0 = 6 * (r / 6 - KLB);
            for (int j = part * per; j < min(W, (part + 1) * per); j++) { const int c = c0 + j; if (c >= 0 && c < n) acc = fma(row[j], vb[c], acc); }
        }
    }
    acc += __shfl_xor_sync(GXB_FULL, acc, 1);
    acc += __shfl_xor_sync(GXB_FULL, acc, 2);
    acc += __shfl_xor_sync(GXB_FULL, acc, 4);
    if (valid && part == 0) y[rg] = acc;
}

// Transpose in row-block storage: in has block bandwidths KLB / KUB (row r: columns 6 (r/6 - KLB) ..), out has KUB / KLB.
// One thread per stored entry of the output.
__global__ void k_rowblock_transpose(const double* __restrict__ in, double* __restrict__ out, int n, int W, int KLB, int KUB, int64_t nbatch) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nbatch * n * W) return;
    const int j = (int)(t % W); const int64_t q = t / W; const int r = (int)(q % n); const int64_t b = q / n;
    const int c = 6 * (r / 6 - KUB) + j;                       // out[r][c] = in[c][r]
    double v = 0.0;
    if (c >= 0 && c < n) {
        const int jj = r - 6 * (c / 6 - KLB);
        if (jj >= 0 && jj < W) v = in[((size_t)b * n + c) * W + jj];
    }
    out[t] = v;
}

// out[b][c][j] = sum_k Vk[b][k][c] * Y[b][k][j]  (Y complex): thread per state c, four eigenvectors per pass
__global__ void __launch_bounds__(128) k_ritz_vectors(const double* __restrict__ Vk, const double* __restrict__ Y, double* __restrict__ out, int n, int m, int nev) {
    const int b = blockIdx.x, c = blockIdx.y * blockDim.x + threadIdx.x;      // batch on grid.x (no 65535 limit)
    const double* V = Vk + (size_t)b * (m + 1) * n;
    const double* Yb = Y + (size_t)b * m * nev * 2;
    for (int j0 = 0; j0 < nev; j0 += 4) {
        double re[4] = {0, 0, 0, 0}, im[4] = {0, 0, 0, 0};
        if (c < n)
            for (int k = 0; k < m; k++) {
                const double v = V[(size_t)k * n + c];
#pragma unroll
                for (int q = 0; q < 4; q++)
                    if (j0 + q < nev) { re[q] = fma(v, Yb[((size_t)k * nev + j0 + q) * 2], re[q]); im[q] = fma(v, Yb[((size_t)k * nev + j0 + q) * 2 + 1], im[q]); }
            }
        if (c < n)
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (j0 + q < nev) { double* o = out + (((size_t)b * n + c) * nev + j0 + q) * 2; o[0] = re[q]; o[1] = im[q]; }
    }
}

// Arnoldi bookkeeping, one CTA per instance.  Vk: [b][m+1][n] Krylov basis, H: [b][m+1][m] (row-major), w: [b][n] = A v_j.
__global__ void __launch_bounds__(256) k_arnoldi_start(double* __restrict__ Vk, int n, int m) {
    const int b = blockIdx.x;
    double* v0 = Vk + (size_t)b * (m + 1) * n;
    __shared__ double red[8];
    double ss = 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        // deterministic pseudo-random start vector (the reference's ArnoldiMethod draws a random one)
        unsigned h = (unsigned)c * 2654435761u + 12345u; h ^= h >> 13; h *= 2246822519u; h ^= h >> 16;
        const double val = 0.5 + (double)(h & 0xffffff) / 16777216.0;
        v0[c] = val; ss = fma(val, val, ss);
    }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(GXB_FULL, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    double tot = 0.0; for (int k = 0; k < (int)(blockDim.x >> 5); k++) tot += red[k];
    const double inv = 1.0 / sqrt(tot);
    for (int c = threadIdx.x; c < n; c += blockDim.x) v0[c] *= inv;
}
// Orthogonalisation of w = A v_j against v_0 .. v_j: classical Gram-Schmidt, h = V'w then w -= V h, repeated once when the DGKS criterion
// ||w'|| < ||w|| / sqrt(2) detects cancellation -- what ArnoldiMethod.jl's `orthogonalize!` does (third party; partialschur is the
// reference's eigen-solver, analyses.jl:951-954).  One CTA per instance; w is staged in shared memory, the basis is read twice per pass
// (dots in chunks of 8 vectors with ONE block reduction per chunk, then the update), instead of four times by a two-sweep modified
// Gram-Schmidt: the kernel is bound by the HBM traffic of the basis ((j + 1) n doubles per read).
__global__ void __launch_bounds__(256) k_arnoldi_orth(double* __restrict__ Vk, double* __restrict__ H, const double* __restrict__ w_in, int n, int m, int j) {
    extern __shared__ __align__(16) double orth_sm[];            // [n] w, then [m + 1] h
    const int b = blockIdx.x;
    double* V = Vk + (size_t)b * (m + 1) * n;
    double* Hb = H + (size_t)b * (m + 1) * m;
    double* vn = V + (size_t)(j + 1) * n;
    const double* w = w_in + (size_t)b * n;
    double* ws = orth_sm;
    double* hs = orth_sm + n;
    __shared__ double red[8][8]; __shared__ double bc8[8];
    const int nw = blockDim.x >> 5, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // sum of up to 8 values per thread over the CTA -> bc8[0..7]
    auto block_sum8 = [&](double (&a)[8]) {
#pragma unroll
        for (int q = 0; q < 8; q++) for (int o = 16; o; o >>= 1) a[q] += __shfl_xor_sync(GXB_FULL, a[q], o);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) red[wid][q] = a[q];
        }
        __syncthreads();
        if (threadIdx.x < 8) { double t = 0.0; for (int k = 0; k < nw; k++) t += red[k][threadIdx.x]; bc8[threadIdx.x] = t; }
        __syncthreads();
    };
    double nrm0;
    {
        double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = threadIdx.x; c < n; c += blockDim.x) { const double x = w[c]; ws[c] = x; a[0] = fma(x, x, a[0]); }
        block_sum8(a);
        nrm0 = sqrt(bc8[0]);
    }
    for (int i = threadIdx.x; i <= j; i += blockDim.x) hs[i] = 0.0;
    double nrm = nrm0;
    for (int pass = 0; pass < 2; pass++) {
        // h = V' w, eight basis vectors per block reduction
        for (int i0 = 0; i0 <= j; i0 += 8) {
            double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int c = threadIdx.x; c < n; c += blockDim.x) {
                const double x = ws[c];
#pragma unroll
                for (int q = 0; q < 8; q++) if (i0 + q <= j) a[q] = fma(V[(size_t)(i0 + q) * n + c], x, a[q]);
            }
            block_sum8(a);
            if (threadIdx.x < 8 && i0 + threadIdx.x <= j) red[0][threadIdx.x] = bc8[threadIdx.x];      // keep this chunk's coefficients
            __syncthreads();
            // w -= V h for this chunk right away is NOT classical Gram-Schmidt (later dots would see the updated w): stash h, update below
            if (threadIdx.x < 8 && i0 + threadIdx.x <= j) { hs[m + 1 + i0 + threadIdx.x] = red[0][threadIdx.x]; }
            __syncthreads();
        }
        double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            double x = ws[c];
            for (int i = 0; i <= j; i++) x = fma(-hs[m + 1 + i], V[(size_t)i * n + c], x);
            ws[c] = x;
            a[0] = fma(x, x, a[0]);
        }
        for (int i = threadIdx.x; i <= j; i += blockDim.x) hs[i] += hs[m + 1 + i];
        block_sum8(a);
        const double nrm_new = sqrt(bc8[0]);
        const bool again = pass == 0 && nrm_new < 0.7071067811865476 * nrm;       // DGKS: uniform over the CTA
        nrm = nrm_new;
        if (!again) break;
    }
    for (int i = threadIdx.x; i <= j; i += blockDim.x) Hb[(size_t)i * m + j] = hs[i];
    if (threadIdx.x == 0) Hb[(size_t)(j + 1) * m + j] = nrm;
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int c = threadIdx.x; c < n; c += blockDim.x) vn[c] = ws[c] * inv;
}
