// Batched pivoted block-banded LU + solve, one warp per instance, FP64.
//
// Replaces the `linsolve` of the Newton iteration: `x .= ImplicitAD.implicit_linear(A, b)` = SparseArrays `A \ b`
// (UMFPACK, symbolic + numeric factorisation on every call), src/analyses.jl:3518, 3585.
//
// The Jacobian of a chain is block-banded in the natural SystemIndices ordering: with 6x6 blocks the static system has
// block bandwidths KLB = KUB = 2.  Its diagonal blocks are structurally singular (free points have no
// ∂F/∂u), so partial pivoting is mandatory; row pivoting inside the band lets U grow to KLB+KUB block super-diagonals.
//
// Mapping: lane = one row of the active window (rows of block-rows J..J+KLB, at most 6(KLB+1) <= 32); the lane keeps its
// row segment (window columns = blocks J..J+KLB+KUB, W = 6(KLB+KUB+1) doubles) plus its right-hand side in registers.
// Right-looking blocked elimination with block size 6, per block column J:
//   A. panel: 6 pivots on window columns 0..5 — warp arg-max (two REDUX + one VOTE), pivot-row panel entries by shuffle,
//      every live row records its multipliers m[0..5];
//   B. the six pivot lanes publish their multipliers (= L11) through shared memory; every lane turns its multipliers into
//      effective ones e = m L11^{-1}, so that the trailing update can use the *unmodified* pivot rows;
//   C. the six pivot lanes publish their trailing entries + rhs in ONE set of vector stores (they are different lanes of
//      the same warp), then every lane does a rank-6 update of its 24 trailing entries, written straight into the
//      shifted register positions (the window slide costs no moves).  Pivot lanes end up holding their final U row, which
//      they stream to global memory (again one set of stores for all six), and then pick up the rows of block-row
//      J+KLB+1 that were prefetched with cp.async during the elimination.
// L is never stored: each Jacobian is used for exactly one right-hand side, which is eliminated on the fly.
// Back substitution walks the blocks backwards, lane k < 6 owns row k of the block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef GXB_LU_TIMING
#include <cstdio>
#endif

#define GXB_FULL 0xffffffffu

template <int KLB, int KUB> struct LuShape {
    static constexpr int W = 6 * (KLB + KUB + 1);   // window / stored-row width
    static constexpr int ROWS = 6 * (KLB + 1);      // live rows in the window
    static constexpr int T = W - 6;                 // trailing columns
    static constexpr int UW = ((W + 2 + 31) / 32) * 32;  // stored U row: W entries, rhs, 1/pivot, pad
    // per-warp shared memory (doubles)
    static constexpr int PSTRIDE = T + 2;           // one published pivot row: T trailing entries, rhs, pad (16B multiple)
    static constexpr int SM_P = 6 * PSTRIDE;        // published pivot rows
    static constexpr int SM_L = 6 * 6;              // published multipliers (L11)
    static constexpr int USTRIDE = W + 4;           // padded U row in shared memory (bank-conflict-free for W = 30)
    static constexpr int NS = 4;                    // prefetch ring depth (block-rows in flight: NS-1)
    static constexpr int STG = 6 * W + 6;           // one staged block-row: 6 rows + 6 rhs
    static constexpr int SM_STAGE = NS * ((STG > 6 * USTRIDE) ? STG : 6 * USTRIDE);
    static constexpr int SM_X = W + 2;              // x window for the back substitution
    static constexpr int SM_BAR = 2 * NS;           // mbarriers of the two prefetch rings (factorisation, back substitution)
    static constexpr int SM_DOUBLES = SM_P + SM_L + SM_STAGE + SM_X + SM_BAR;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: ONE instruction moves a whole block-row (6 rows of the
// row-block storage are contiguous in HBM) instead of ~5 predicated LDGSTS rounds with their address arithmetic ----
#ifndef GXB_LU_TMA
#define GXB_LU_TMA 1
#endif
// Back substitution: 1 = column-oriented (each lane owns a pending row, one 6-term axpy per block), 0 = row-oriented dot products
#ifndef GXB_LU_AXPY
#define GXB_LU_AXPY 1
#endif
#ifndef GXB_LU_NSET
#define GXB_LU_NSET 2
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Wait for the phase with the given parity.  try_wait suspends the thread in hardware up to a time limit; the retry count is bounded so
// that a lost transaction surfaces as a kernel error (trap) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    for (int it = 0; it < (1 << 22); it++) {
        unsigned ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}

// 1/x without the slow-path branch of the IEEE division: hardware seed (MUFU.RCP64H) + three Newton steps (< 2 ulp),
// so that it can be scheduled in the shadow of the warp arg-max.  x = 0 / denormal / inf give inf or NaN, which the
// caller flags (zero pivot) or which surfaces as a non-finite Newton step.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
}

// Solve J y = rhs for one instance; writes sol = sign * y.  Returns 0, or 1 if an exactly-zero / non-finite pivot was hit.
//   Jrow : [n][W]  row-block storage (see gxb_static.cuh)      rhs : [n]
//   U    : [n][UW] scratch for the U rows                       sol : [n]
template <int KLB, int KUB>
__device__ int warp_band_solve(const double* __restrict__ Jrow, const double* __restrict__ rhs, double* __restrict__ U,
                               double* __restrict__ sol, int nb, double sign, double* sm) {
    using S = LuShape<KLB, KUB>;
    constexpr int W = S::W, ROWS = S::ROWS, UW = S::UW, T = S::T, PS = S::PSTRIDE;
    static_assert(T % 2 == 0 && W % 2 == 0, "vector access needs even widths");
    const int lane = threadIdx.x & 31;
    const int n = 6 * nb;
    // stored U row: AXPY layout [rhs, 1/pivot, panel(6), trailing(T)], else [panel(6), trailing(T), rhs, 1/pivot]
    constexpr int UOFF_PANEL = GXB_LU_AXPY ? 2 : 0, UOFF_HEAD = GXB_LU_AXPY ? 0 : W;
    double* Psm = sm;                      // [6][PS]
    double* Lsm = sm + S::SM_P;            // [6][6]
    double* stage = Lsm + S::SM_L;         // 6*W row entries, then 6 rhs
    double* xs = stage + S::SM_STAGE;      // x window
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + S::SM_X);   // [2][NS]
#if GXB_LU_TMA
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 2 * S::NS; q++) mbar_init(bars + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncwarp();
#endif

    double a[W];
    double b = 0.0;
    bool live = false;
    int bad = 0;

    // ---- initial window: block-rows 0..KLB; row-block B stores column blocks B-KLB.., window starts at block 0 ----
    {
        const int row = lane;
        const int off = 6 * (KLB - lane / 6);
        if (lane < ROWS && row < n) {
            live = true;
            const double* src = Jrow + (size_t)row * W;
#pragma unroll
            for (int j = 0; j < W; j++) a[j] = (j + off < W) ? src[j + off] : 0.0;
            b = rhs[row];
        } else {
#pragma unroll
            for (int j = 0; j < W; j++) a[j] = 0.0;
        }
    }

    // block-rows are prefetched NS-1 steps ahead into a ring of shared-memory buffers (cp.async groups, one per step)
    constexpr int NS = S::NS, STG = S::STG;
    auto prefetch_rows = [&](int step) {               // block-row needed at the end of elimination step `step`
        const int r0 = 6 * (step + KLB + 1);
#if GXB_LU_TMA
        if (r0 < n && lane == 0) {
            double* dst = stage + (step % NS) * STG;
            uint64_t* bar = bars + (step % NS);
            mbar_expect_tx(bar, 6 * W * 8 + 48);
            bulk_g2s(dst, Jrow + (size_t)r0 * W, 6 * W * 8, bar);
            bulk_g2s(dst + 6 * W, rhs + r0, 48, bar);
        }
#else
        if (r0 < n) {
            double* dst = stage + (step % NS) * STG;
            const char* g = reinterpret_cast<const char*>(Jrow + (size_t)r0 * W);
            char* sdst = reinterpret_cast<char*>(dst);
            for (int c = lane; c < (6 * W * 8) / 16; c += 32) cp_async16(sdst + 16 * c, g + 16 * c);
            if (lane < 3) cp_async16(reinterpret_cast<char*>(dst + 6 * W) + 16 * lane, reinterpret_cast<const char*>(rhs + r0) + 16 * lane);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
    };
#pragma unroll
    for (int q = 0; q < NS - 1; q++) prefetch_rows(q);

#ifdef GXB_LU_TIMING
    long long tph[5] = {0, 0, 0, 0, 0}, tq = clock64(), tn;
#define GXB_TICK(i) do { tn = clock64(); tph[i] += tn - tq; tq = tn; } while (0)
#else
#define GXB_TICK(i) do { } while (0)
#endif
    for (int J = 0; J < nb; J++) {
        prefetch_rows(J + NS - 1);
        const bool have_next = 6 * (J + KLB + 1) < n;
        GXB_TICK(4);
        // ---- A. panel factorisation on window columns 0..5 ----
        double m[6];
        int kp = 6;            // which pivot this lane became (6 = none)
        double myrinv = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            // arg-max of |a[k]| over live lanes with ONE warp reduction: key = float(|a|) with the 5 low mantissa bits
            // replaced by (31 - lane).  Partial pivoting only needs a near-maximal pivot (relative slack 2^-18); exact ties
            // go to the lowest lane, so the choice is deterministic.
            const unsigned fk = __float_as_uint(__double2float_ru(fabs(a[k])));
            const unsigned key = live ? ((fk & 0xffffffe0u) | (31u - (unsigned)lane)) : 0u;
            const double rc = fast_rcp(a[k]);                  // every lane inverts its own candidate meanwhile
            const unsigned mk = __reduce_max_sync(GXB_FULL, key);
            const int p = 31 - (int)(mk & 31u);
            const double rinv = __shfl_sync(GXB_FULL, rc, p);
            const bool is_piv = (lane == p) && live;
            if (!(mk >> 5) || !isfinite(rinv)) bad = 1;        // zero (or sub-float-denormal) / NaN pivot column
            if (is_piv) { live = false; kp = k; myrinv = rinv; }
            m[k] = live ? a[k] * rinv : 0.0;
#pragma unroll
            for (int j = k + 1; j < 6; j++) {
                const double pj = __shfl_sync(GXB_FULL, a[j], p);
                a[j] = fma(-m[k], pj, a[j]);
            }
        }
        GXB_TICK(0);
        // ---- B/C. publish L11 rows, original trailing parts and rhs of the six pivot rows (one store set for all six) ----
        if (kp < 6) {
            double* Lr = Lsm + kp * 6;
#pragma unroll
            for (int j = 0; j < 6; j += 2) *reinterpret_cast<double2*>(Lr + j) = make_double2(m[j], m[j + 1]);
            double* Pr = Psm + kp * PS;
#pragma unroll
            for (int j = 0; j < T; j += 2) *reinterpret_cast<double2*>(Pr + j) = make_double2(a[6 + j], a[6 + j + 1]);
            Pr[T] = b;
            // panel part of the U row goes straight to global memory (entries left of the diagonal are never read back)
            double* urow = U + (size_t)(6 * J + kp) * UW + UOFF_PANEL;
#pragma unroll
            for (int j = 0; j < 6; j += 2) *reinterpret_cast<double2*>(urow + j) = make_double2(a[j], a[j + 1]);
        }
        __syncwarp();
        // effective multipliers: e L11 = m, L11 unit lower triangular with L11[k][k'] = m_{pivot k}[k']
        double e[6];
        e[5] = m[5];
#pragma unroll
        for (int kk = 4; kk >= 0; kk--) {
            double acc = m[kk];
#pragma unroll
            for (int k = kk + 1; k < 6; k++) acc = fma(-e[k], Lsm[k * 6 + kk], acc);
            e[kk] = acc;
        }
        GXB_TICK(1);
        // rank-6 trailing update, written into the shifted positions (window slides by one block)
#pragma unroll
        for (int j = 0; j < T; j += 2) {
            double x0 = a[6 + j], x1 = a[6 + j + 1];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const double2 pv = *reinterpret_cast<const double2*>(Psm + k * PS + j);
                x0 = fma(-e[k], pv.x, x0);
                x1 = fma(-e[k], pv.y, x1);
            }
            a[j] = x0; a[j + 1] = x1;
        }
#pragma unroll
        for (int k = 0; k < 6; k++) b = fma(-e[k], Psm[k * PS + T], b);
#pragma unroll
        for (int j = T; j < W; j++) a[j] = 0.0;
        const unsigned rmask = __ballot_sync(GXB_FULL, kp < 6);
        GXB_TICK(2);
#if GXB_LU_TMA
        if (have_next) mbar_wait(bars + (J % NS), (unsigned)(J / NS) & 1u);        // the block-row of this step has landed
#else
        asm volatile("cp.async.wait_group %0;\n" ::"n"(S::NS - 1) : "memory");   // the group of this step has landed
        __syncwarp();
#endif
        if (kp < 6) {
            // this lane now holds the final U row of pivot kp (trailing part + rhs): stream it out, then take a new row
            double* urow = U + (size_t)(6 * J + kp) * UW;
#pragma unroll
            for (int j = 0; j < T; j += 2) *reinterpret_cast<double2*>(urow + UOFF_PANEL + 6 + j) = make_double2(a[j], a[j + 1]);
            *reinterpret_cast<double2*>(urow + UOFF_HEAD) = make_double2(b, myrinv);
            const int rank = __popc(rmask & ((1u << lane) - 1u));
            if (have_next) {
                const double* sbuf = stage + (J % NS) * STG;
                const double* src = sbuf + rank * W;
#pragma unroll
                for (int j = 0; j < W; j += 2) { double2 t = *reinterpret_cast<const double2*>(src + j); a[j] = t.x; a[j + 1] = t.y; }
                b = sbuf[6 * W + rank];
                live = true;
            } else {
#pragma unroll
                for (int j = 0; j < W; j++) a[j] = 0.0;
                b = 0.0;
            }
        }
        __syncwarp();
        GXB_TICK(3);
    }
#ifdef GXB_LU_TIMING
    const long long tfac_end = clock64();
#endif

#if GXB_LU_AXPY
    // ---- back substitution, column-oriented.  A row of block B needs the solutions of blocks B+1 .. B+NBLK (NBLK = KLB + KUB block
    // super-diagonals of U), so NBLK + 1 blocks of rows are "pending" at any time: each pending row lives in ONE lane (slot) with
    // its running sum.  Per block J, descending: gather the six finished sums of block J, solve the 6x6 triangle redundantly in
    // every lane (no shuffle chain), then every pending row subtracts its six entries of column block J times x_J -- one 6-term
    // axpy per lane, entries prefetched one step ahead straight from global memory into registers.  Only the 64-byte heads of the
    // rows (rhs, 1/pivot, the triangle) go through shared memory.
    {
        constexpr int NBLK = KLB + KUB, M = NBLK + 1, R = 6 * M, PEND = (R + 31) / 32, HS = 4;
        static_assert(HS * 48 <= S::SM_STAGE, "head ring must fit the staging area");
#if !GXB_LU_TMA
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
        __syncwarp();                                    // the U rows written by other lanes of this warp are visible from here on
        double* hb = stage;                              // [HS][6][8] heads
        auto prefetch_head = [&](int Jb) {
            if (Jb >= 0 && lane < 24) cp_async16(hb + (Jb % HS) * 48 + (lane >> 2) * 8 + 2 * (lane & 3), U + (size_t)(6 * Jb + (lane >> 2)) * UW + 2 * (lane & 3));
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
        double ps[PEND];                                 // running sums of the pending rows of this lane
        int psb[PEND], pk[PEND];                         // slot block and row within the block
        bool pv[PEND];
        // the six entries (and, for a row that enters the pending set, its rhs) for the current step (A) and the next one (B):
        // loaded two steps ahead so that the L2 latency is off the per-block dependency chain
        constexpr int NSET = GXB_LU_NSET;                // prefetch distance in block steps
        double2 f0_[NSET][PEND], f1_[NSET][PEND], f2_[NSET][PEND];
        double rin_[NSET][PEND];
#pragma unroll
        for (int q = 0; q < PEND; q++) { const int pr = lane + 32 * q; pv[q] = pr < R; psb[q] = pr / 6; pk[q] = pr % 6; ps[q] = 0.0; }
        auto load_step = [&](int Jt, int q, double2& f0, double2& f1, double2& f2, double& rin) {
            f0 = f1 = f2 = make_double2(0.0, 0.0); rin = 0.0;
            if (Jt < 0 || !pv[q]) return;
            const int d = ((Jt - psb[q]) % M + M) % M;   // distance of the slot's row block from block Jt
            const int B = Jt - d;
            if (d < 1 || B < 0) return;
            const double* rp = U + (size_t)(6 * B + pk[q]) * UW;
            const double* ep = rp + 8 + 6 * (d - 1);
            f0 = *reinterpret_cast<const double2*>(ep); f1 = *reinterpret_cast<const double2*>(ep + 2); f2 = *reinterpret_cast<const double2*>(ep + 4);
            if (d == NBLK) rin = rp[0];
        };
#pragma unroll
        for (int q = 0; q < PEND; q++) {
            const int d = ((nb - 1 - psb[q]) % M + M) % M;
            const int B = nb - 1 - d;
            ps[q] = (pv[q] && B >= 0) ? U[(size_t)(6 * B + pk[q]) * UW] : 0.0;      // every row starts from its right-hand side
#pragma unroll
            for (int t = 0; t < NSET; t++) load_step(nb - 1 - t, q, f0_[t][q], f1_[t][q], f2_[t][q], rin_[t][q]);
        }
#pragma unroll
        for (int q = 0; q < HS - 1; q++) prefetch_head(nb - 1 - q);
        // The U rows of a Newton round do not fit the L2 (1.3 GB for C2): by the time the back substitution walks back to a row it
        // has been evicted to HBM, and a DRAM round trip (~2 us under load) is longer than several block steps.  One TMA L2-prefetch
        // per step (a block-row of U is contiguous) brings the rows back PD steps before the register prefetch touches them.
        constexpr int PD = 10;
        auto prefetch_l2 = [&](int Jb) {
            if (Jb >= 0 && lane == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(U + (size_t)(6 * Jb) * UW), "r"(6 * UW * 8) : "memory");
        };
        for (int q = 3; q < PD; q++) prefetch_l2(nb - 1 - q);
        // one block step; (f0, f1, f2, rin) is the register set that was loaded for this step two steps ago and is reloaded for
        // step J-NSET at the end.  The loop below cycles through the NSET sets (unrolled): rotating them with moves would
        // make every iteration wait for the load issued by the previous one.
        auto step = [&](int J, double2 (&f0)[PEND], double2 (&f1)[PEND], double2 (&f2)[PEND], double (&rin)[PEND]) {
            prefetch_l2(J - PD);
            prefetch_head(J - (HS - 1));
            asm volatile("cp.async.wait_group %0;\n" ::"n"(HS - 1) : "memory");
            __syncwarp();
            // the six finished sums of block J
            const int base = 6 * (J % M);
            double x[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const int idx = base + k;
                double v = __shfl_sync(GXB_FULL, ps[0], idx & 31);
                if (PEND > 1) { const double w2 = __shfl_sync(GXB_FULL, ps[PEND - 1], idx & 31); v = idx >= 32 ? w2 : v; }
                x[k] = v;
            }
            // 6x6 upper triangle, the same arithmetic in every lane: x_k = (s_k - sum_{j>k} u_kj x_j) / u_kk
            const double* h = hb + (J % HS) * 48;
#pragma unroll
            for (int k = 5; k >= 0; k--) {
                x[k] *= h[8 * k + 1];
#pragma unroll
                for (int i = 0; i < k; i++) x[i] = fma(-h[8 * i + 2 + k], x[k], x[i]);
            }
            if (lane < 6) {
                double xv = x[0];
#pragma unroll
                for (int k = 1; k < 6; k++) xv = lane == k ? x[k] : xv;
                sol[6 * J + lane] = sign * xv;
            }
            // axpy into the pending rows (two interleaved 3-term chains), then reload this register set for step J-NSET
#pragma unroll
            for (int q = 0; q < PEND; q++) {
                const int d = ((J - psb[q]) % M + M) % M;
                if (pv[q] && d >= 1 && J - d >= 0) {
                    double acc = (d == NBLK && J != nb - 1) ? rin[q] : ps[q];        // a row entering at this step starts from its rhs
                    double acc2 = -f0[q].y * x[1];
                    acc = fma(-f0[q].x, x[0], acc); acc2 = fma(-f1[q].y, x[3], acc2);
                    acc = fma(-f1[q].x, x[2], acc); acc2 = fma(-f2[q].y, x[5], acc2);
                    acc = fma(-f2[q].x, x[4], acc);
                    ps[q] = acc + acc2;
                }
                load_step(J - NSET, q, f0[q], f1[q], f2[q], rin[q]);
            }
            __syncwarp();
        };
        for (int J = nb - 1; J >= 0; J -= NSET) {
#pragma unroll
            for (int t = 0; t < NSET; t++)
                if (J - t >= 0) step(J - t, f0_[t], f1_[t], f2_[t], rin_[t]);
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
#else
    // ---- back substitution, block by block from the end; lane k < 6 owns row 6J+k ----
    // U rows of block J-1 are prefetched (cp.async, double buffered) while block J is being solved.
    constexpr int US = S::USTRIDE;
#if !GXB_LU_TMA
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
    __syncwarp();
    double* ub = stage;                                  // [NS][6][US], reuses the factorisation's staging area
    // q-th prefetch of the back substitution fetches block nb-1-q into slot q % NS (phase parity (q / NS) & 1)
    auto prefetch_u = [&](int q) {
        const int Jb = nb - 1 - q;
#if GXB_LU_TMA
        if (Jb >= 0) {
            uint64_t* bar = bars + NS + (q % NS);
            if (lane == 0) mbar_expect_tx(bar, 6 * (W + 2) * 8);
            __syncwarp();
            if (lane < 6) bulk_g2s(ub + (q % NS) * 6 * US + lane * US, U + (size_t)(6 * Jb + lane) * UW, (W + 2) * 8, bar);
        }
#else
        if (Jb >= 0) {
            double* dst = ub + (q % NS) * 6 * US;
            const double* src = U + (size_t)(6 * Jb) * UW;
            constexpr int CH = (W + 2) / 2;                  // 16-byte chunks per row
            for (int c = lane; c < 6 * CH; c += 32) {
                const int r = c / CH, qq = c % CH;
                cp_async16(dst + r * US + 2 * qq, src + (size_t)r * UW + 2 * qq);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
#endif
    };
    for (int c = lane; c < W + 2; c += 32) xs[c] = 0.0;
    __syncwarp();
#if GXB_LU_TMA
    // the U rows were written with ordinary stores by this warp: order them before the async-proxy reads of the bulk copies
    __threadfence_block();
    asm volatile("fence.proxy.async;\n" ::: "memory");
    __syncwarp();
#endif
#pragma unroll
    for (int q = 0; q < NS - 1; q++) prefetch_u(q);
    for (int J = nb - 1; J >= 0; J--) {
        const int qi = nb - 1 - J;                       // index of this block in the prefetch sequence
        prefetch_u(qi + NS - 1);
#if GXB_LU_TMA
        mbar_wait(bars + NS + (qi % NS), (unsigned)(qi / NS) & 1u);
#else
        asm volatile("cp.async.wait_group %0;\n" ::"n"(S::NS - 1) : "memory");
        __syncwarp();
#endif
        const double* row = ub + (qi % NS) * 6 * US + (lane < 6 ? lane : 0) * US;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        double up[6];
        {
            const double2 t = *reinterpret_cast<const double2*>(row + W);   // rhs, 1/pivot
            s0 = t.x;
            const double myr_ = t.y;
#pragma unroll
            for (int j = 0; j < 6; j += 2) { const double2 q = *reinterpret_cast<const double2*>(row + j); up[j] = q.x; up[j + 1] = q.y; }
#pragma unroll
            for (int j = 6; j < W; j += 4) {
                const double2 u0 = *reinterpret_cast<const double2*>(row + j);
                const double2 x0 = *reinterpret_cast<const double2*>(xs + j);
                s0 = fma(-u0.x, x0.x, s0);
                s1 = fma(-u0.y, x0.y, s1);
                if (j + 2 < W) {
                    const double2 u1 = *reinterpret_cast<const double2*>(row + j + 2);
                    const double2 x1 = *reinterpret_cast<const double2*>(xs + j + 2);
                    s2 = fma(-u1.x, x1.x, s2);
                    s3 = fma(-u1.y, x1.y, s3);
                }
            }
            double s = (s0 + s1) + (s2 + s3);
#pragma unroll
            for (int k = 5; k >= 0; k--) {
                const double xk = __shfl_sync(GXB_FULL, s * myr_, k);
                if (lane < k) s = fma(-up[k], xk, s);
            }
            s *= myr_;
            // slide the x window for block J-1: xs'[6..11] = this block's solution, xs'[j+6] = xs[j] for the older
            // entries (xs[0..5] is never read)
            double keep0 = (lane < 6) ? s : xs[lane];
            double keep1 = (lane + 32 < W - 6) ? xs[lane + 32] : 0.0;
            __syncwarp();
            if (lane < W - 6) xs[lane + 6] = keep0;
            if (lane + 32 < W - 6) xs[lane + 38] = keep1;
            if (lane < 6) sol[6 * J + lane] = sign * s;
        }
        __syncwarp();
    }
#endif
#if GXB_LU_TMA
    // the shared memory is reused by the caller (persistent kernels run assembly tiles through it): retire the barrier objects
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 2 * S::NS; q++) asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bars + q)) : "memory");
    }
    __syncwarp();
#endif
#ifdef GXB_LU_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 0)
        printf("[lu timing KLB=%d nb=%d] cycles/block-step: panel %lld, publish+e %lld, trailing %lld, tail %lld, prefetch-issue %lld | back-subst/block %lld\n",
               KLB, nb, tph[0] / nb, tph[1] / nb, tph[2] / nb, tph[3] / nb, tph[4] / nb, (clock64() - tfac_end) / nb);
#endif
    return __any_sync(GXB_FULL, bad) ? 1 : 0;
}
