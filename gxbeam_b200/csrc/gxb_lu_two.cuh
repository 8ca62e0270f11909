// Two-sided ("burn at both ends") batched pivoted band LU + solve: two warps per instance, no spikes.
//
// Same job as the other factorisation kernels (the `A \ b` of the Newton iteration, src/analyses.jl:3518, 3585).  gxb_lu_part.cuh cuts a
// static chain into P partitions whose interior ones carry an 18-wide spike; for DYNAMIC systems (6-block bandwidths 3 / 4) an interior
// partition would start with 36 live rows, more than the 32 lanes of the lane-per-row window.  Two partitions need no spike at all:
//   warp 0 eliminates the column blocks [0, R - A) top-down with the rows of blocks [0, R)             (exactly warp_band_solve's sweep);
//   warp 1 eliminates the column blocks [R + B, nb) BOTTOM-UP with the rows of blocks [R, nb): the same sweep on the matrix with rows and
//          columns reversed (block bandwidths swap, a stored row is read back to front), so no second code path exists;
//   A / B = blocks below / above the cut R that the respective side must leave alone because rows of the other side touch them
//          (static chain, cut in front of a point: 2 / 1; dynamic chain: 3 / 2 -- SURVEY App. A.3).  As in gxb_lu_part.cuh the covered
//          columns restricted to a side's rows are complete columns of J, so restricted partial pivoting always finds a pivot.
//   The 6 A + 6 B rows never chosen as a pivot meet in a dense 6 (A + B) system (18 x 18 static, 30 x 30 dynamic: at most 30 rows, one
//   lane each) that warp 0 eliminates with the same 6-column block step; both warps then back-substitute their side.
// Critical path: half the columns per warp.  Measured on config 4 (512 x n = 444, 2000 Newmark steps): see DESIGN.md 3.2f.
#pragma once
#include "gxb_lu_part.cuh"

template <int KLB, int KUB, int AB, int BB> struct TwoShape {
    static constexpr int W = 6 * (KLB + KUB + 1), T = W - 6, NBLK = KLB + KUB;
    static constexpr int NI = 6 * (AB + BB);               // interface unknowns = left-over rows
    static constexpr int UW = 2 + 6 + T;                    // stored U row: rhs, 1/pivot, panel, trailing
    static constexpr int NS = 3, STG = 6 * W + 6;
    static constexpr int NB = 4;                            // back substitution: ring of staged U block rows
    static constexpr int SM_SWEEP = 6 * (T + 2) + 36 + NS * STG;
    static constexpr int SM_BACK = NB * 6 * UW + (NBLK + 1) * 6;
    static constexpr int SM_WORK = ((SM_SWEEP > SM_BACK ? SM_SWEEP : SM_BACK) + 1) & ~1;
    static constexpr int SM_BAR = NS + NB + 1;
    static constexpr int WARP_DOUBLES = SM_WORK + SM_BAR + (SM_BAR & 1);
    static constexpr int RW = NI + 2;                       // a left-over row: NI interface entries, rhs, pad
    static constexpr int RUW = 2 + 6 + (NI - 6);            // reduced U row
    static constexpr int CTA_DOUBLES = 2 * WARP_DOUBLES + NI * RW + NI * RUW + NI + 6 * (NI - 6 + 2) + 36 + 4;
    static_assert(NI <= 32 && NI >= 12, "one lane per left-over row");
    static_assert(6 * (KLB + 1) <= 32 && 6 * (KUB + 1) <= 32, "one lane per window row on both sides");
    static_assert(UW % 2 == 0 && RW % 2 == 0 && RUW % 2 == 0 && STG % 2 == 0, "16-byte alignment");
};

// Solve J y = rhs with the two warps of the CTA; sol = sign * y.  Every thread returns the same flag (1: zero / non-finite pivot).
//   Jrow : [n][W] row-block storage     rhs : [n]     U : [n][UW] scratch     sol : [n]     R : cut (first row block of the lower side)
template <int KLB, int KUB, int AB, int BB>
__device__ int cta_two_sided_solve(const double* __restrict__ Jrow, const double* __restrict__ rhs, double* __restrict__ U,
                                   double* __restrict__ sol, int nb, int R, double sign, double* sm) {
    using S = TwoShape<KLB, KUB, AB, BB>;
    constexpr int W = S::W, T = S::T, NBLK = S::NBLK, NI = S::NI, UW = S::UW, NS = S::NS, STG = S::STG, RW = S::RW, RUW = S::RUW;
    const int lane = threadIdx.x & 31, dir = threadIdx.x >> 5;          // dir 0: top-down, dir 1: bottom-up (mirrored)
    const int n = 6 * nb;
    const int klb = dir ? KUB : KLB;                     // lower block bandwidth of this side's (possibly mirrored) matrix
    const int Rm = dir ? nb - R : R;                     // row blocks of this side
    const int C1 = dir ? nb - R - BB : R - AB;           // column blocks it eliminates: [0, C1) in its own numbering
    // mirrored block Bm <-> original block nb-1-Bm; mirrored row 6 Bm + k <-> original row 6 (nb-1-Bm) + 5 - k; a stored row back to front
    auto orig_row = [&](int rm) { return dir ? n - 1 - rm : rm; };

    double* wsm = sm + (size_t)dir * S::WARP_DOUBLES;
    double* Psm = wsm;
    double* Lsm = Psm + 6 * (T + 2);
    double* stage = Lsm + 36;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + S::SM_WORK);
    double* left = sm + 2 * S::WARP_DOUBLES;             // [NI][RW] left-over rows: top side first
    double* Ured = left + NI * RW;                       // [NI][RUW]
    double* xint = Ured + NI * RUW;                      // [NI]
    double* rPsm = xint + NI;                            // [6][NI - 6 + 2]
    double* rLsm = rPsm + 6 * (NI - 6 + 2);              // [36]
    int* flags = reinterpret_cast<int*>(rLsm + 36);      // [0..1] bad per side, [2..3] left-over counts

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NS + S::NB; i++) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncwarp();

#ifdef GXB_TWO_TIMING
    long long t0 = clock64(), t1 = 0, t2 = 0, t3 = 0;
#endif
    // ------------------------------------------------------------------ sweep of this side
    double a[W], dummy[1] = {0.0};
    double b = 0.0;
    bool live = false;
    int bad = 0;
    {
        const int nrows0 = 6 * min(klb + 1, Rm);
        if (lane < nrows0) {
            live = true;
            const int B = lane / 6, off = 6 * (klb - B);
            const double* src = Jrow + (size_t)orig_row(lane) * W;
#pragma unroll
            for (int j = 0; j < W; j++) { const int q = j + off; a[j] = q < W ? (dir ? src[W - 1 - q] : src[q]) : 0.0; }
            b = rhs[orig_row(lane)];
        } else {
#pragma unroll
            for (int j = 0; j < W; j++) a[j] = 0.0;
        }
    }
    auto prefetch_rows = [&](int step) {                 // row block entering at the end of elimination step `step`
        const int Bn = step + klb + 1;
        if (Bn < Rm && lane == 0) {
            const int Bo = dir ? nb - 1 - Bn : Bn;       // the original block: 6 contiguous rows either way
            double* dst = stage + (step % NS) * STG;
            uint64_t* bar = bars + (step % NS);
            mbar_expect_tx(bar, 6 * W * 8 + 48);
            bulk_g2s(dst, Jrow + (size_t)(6 * Bo) * W, 6 * W * 8, bar);
            bulk_g2s(dst + 6 * W, rhs + 6 * Bo, 48, bar);
        }
    };
#pragma unroll
    for (int i = 0; i < NS - 1; i++) prefetch_rows(i);

    for (int J = 0; J < C1; J++) {
        prefetch_rows(J + NS - 1);
        const bool have_next = J + klb + 1 < Rm;
        int kp; double myrinv;
        // U rows of mirrored block J live at the original rows n-1-6J-k: base of pivot 0 and a negative stride
        double* ubase = dir ? U + (size_t)(n - 1 - 6 * J) * UW : U + (size_t)(6 * J) * UW;
        const int ustride = dir ? -UW : UW;
        lu_block_step<W, 0>(a, dummy, b, live, bad, Psm, Lsm, ubase, ustride, kp, myrinv, false);
        const unsigned rmask = __ballot_sync(GXB_FULL, kp < 6);
        if (have_next) mbar_wait(bars + (J % NS), (unsigned)(J / NS) & 1u);
        if (kp < 6) {
            double* urow = ubase + (ptrdiff_t)kp * ustride;
            *reinterpret_cast<double2*>(urow) = make_double2(b, myrinv);
#pragma unroll
            for (int j = 0; j < T; j += 2) *reinterpret_cast<double2*>(urow + 8 + j) = make_double2(a[j], a[j + 1]);
            const int rank = __popc(rmask & ((1u << lane) - 1u));
            if (have_next) {
                const double* sbuf = stage + (J % NS) * STG;
                if (dir) {
                    const double* src = sbuf + (5 - rank) * W;
#pragma unroll
                    for (int j = 0; j < W; j++) a[j] = src[W - 1 - j];
                    b = sbuf[6 * W + 5 - rank];
                } else {
                    const double* src = sbuf + rank * W;
#pragma unroll
                    for (int j = 0; j < W; j += 2) { const double2 t = *reinterpret_cast<const double2*>(src + j); a[j] = t.x; a[j + 1] = t.y; }
                    b = sbuf[6 * W + rank];
                }
                live = true;
            } else {
#pragma unroll
                for (int j = 0; j < W; j++) a[j] = 0.0;
                b = 0.0;
            }
        }
        __syncwarp();
    }
    // the U rows were written with ordinary stores and come back through the async proxy in the back substitution
    __threadfence_block();
    asm volatile("fence.proxy.async;\n" ::: "memory");
    // left-over rows -> the shared interface system, in ORIGINAL column order (the mirrored side's window runs backwards)
    {
        __syncwarp();
        const unsigned lm = __ballot_sync(GXB_FULL, live);
        const int rank = __popc(lm & ((1u << lane) - 1u));
        const int cnt = __popc(lm), expect = dir ? 6 * BB : 6 * AB;
        if (live && rank < expect) {
            double* o = left + (size_t)((dir ? 6 * AB : 0) + rank) * RW;
#pragma unroll
            for (int j = 0; j < NI; j++) o[dir ? NI - 1 - j : j] = a[j];
            o[NI] = b;
        }
        if (cnt != expect) bad = 1;
        bad = __any_sync(GXB_FULL, bad) ? 1 : 0;
        if (lane == 0) flags[dir] = bad;
    }
#ifdef GXB_TWO_TIMING
    t1 = clock64();
#endif
    __syncthreads();

    // ------------------------------------------------------------------ interface system (warp 0): NI rows, NI columns, dense
    if (dir == 0) {
        double w[NI];
        double rb = 0.0;
        bool rlive = false;
        int rbad = 0;
#pragma unroll
        for (int j = 0; j < NI; j++) w[j] = 0.0;
        if (lane < NI) {
            const double* o = left + (size_t)lane * RW;
#pragma unroll
            for (int j = 0; j < NI; j++) w[j] = o[j];
            rb = o[NI];
            rlive = true;
        }
#pragma unroll 1
        for (int j = 0; j < AB + BB; j++) {
            int kp; double myrinv;
            double* ubase = Ured + (size_t)(6 * j) * RUW;
            lu_block_step<NI, 0>(w, dummy, rb, rlive, rbad, rPsm, rLsm, ubase, RUW, kp, myrinv, false);
            if (kp < 6) {
                double* urow = ubase + (size_t)kp * RUW;
                *reinterpret_cast<double2*>(urow) = make_double2(rb, myrinv);
#pragma unroll
                for (int c = 0; c < NI - 6; c += 2) *reinterpret_cast<double2*>(urow + 8 + c) = make_double2(w[c], w[c + 1]);
#pragma unroll
                for (int c = 0; c < NI; c++) w[c] = 0.0;
                rb = 0.0;
            }
            __syncwarp();
        }
        if (__any_sync(GXB_FULL, rlive)) rbad = 1;
        // back substitution block by block: lane k < 6 owns row k of the block; its trailing entries multiply the already solved x
        for (int blk = AB + BB - 1; blk >= 0; blk--) {
            const double* urow = Ured + (size_t)(6 * blk + (lane < 6 ? lane : 0)) * RUW;
            double s = urow[0];
            const double rinv = urow[1];
            double pan[6];
#pragma unroll
            for (int k = 0; k < 6; k++) pan[k] = urow[2 + k];
            for (int c = 0; c < NI - 6 * (blk + 1); c++) s = fma(-urow[8 + c], xint[6 * (blk + 1) + c], s);
#pragma unroll
            for (int k = 5; k >= 0; k--) {
                const double xk = __shfl_sync(GXB_FULL, s * rinv, k);
                if (lane < k) s = fma(-pan[k], xk, s);
                if (lane == k) { xint[6 * blk + k] = xk; sol[6 * (R - AB) + 6 * blk + k] = sign * xk; }
            }
            __syncwarp();
        }
        rbad = __any_sync(GXB_FULL, rbad) ? 1 : 0;
        if (lane == 0 && rbad) flags[0] = 1;
    }
    __syncthreads();
#ifdef GXB_TWO_TIMING
    t2 = clock64();
#endif

    // ------------------------------------------------------------------ back substitution of this side
    {
        constexpr int NB = S::NB;
        const int r = lane >> 2, part = lane & 3;
        double* uring = wsm;                                   // [NB][6][UW]
        double* xring = wsm + NB * 6 * UW;                     // [NBLK + 1][6]: x of (mirrored) column block Bk in slot Bk % (NBLK + 1)
        uint64_t* bbars = bars + NS;
        auto prefetch_u = [&](int i) {                         // i-th block of the descent: J = C1 - 1 - i; its 6 U rows are contiguous
            if (i < C1 && lane == 0) {
                const int J = C1 - 1 - i;
                const double* src = dir ? U + (size_t)(n - 6 * J - 6) * UW : U + (size_t)(6 * J) * UW;
                uint64_t* bar = bbars + (i % NB);
                mbar_expect_tx(bar, 6 * UW * 8);
                bulk_g2s(uring + (i % NB) * 6 * UW, src, 6 * UW * 8, bar);
            }
        };
        if (lane == 0) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NB - 1; i++) prefetch_u(i);
        // blocks C1 .. C1 + NBLK - 1 of this side's numbering: the interface (A + B blocks, reversed for the mirrored side), then zeros
        for (int t = lane; t < 6 * NBLK; t += 32) {
            const int Bk = C1 + t / 6;
            double v = 0.0;
            if (t < NI) v = dir ? xint[NI - 1 - t] : xint[t];
            xring[(Bk % (NBLK + 1)) * 6 + t % 6] = v;
        }
        __syncwarp();
        for (int i = 0; i < C1; i++) {
            const int J = C1 - 1 - i;
            prefetch_u(i + NB - 1);
            mbar_wait(bbars + (i % NB), (unsigned)(i / NB) & 1u);
            const int rr = r < 6 ? r : 0;
            const double* urow = uring + (i % NB) * 6 * UW + (dir ? 5 - rr : rr) * UW;
            const double2 head = *reinterpret_cast<const double2*>(urow);
            double pan[6];
#pragma unroll
            for (int k = 0; k < 6; k += 2) { const double2 t = *reinterpret_cast<const double2*>(urow + 2 + k); pan[k] = t.x; pan[k + 1] = t.y; }
            double s = part == 0 ? head.x : 0.0, s2 = 0.0;
#pragma unroll
            for (int q = 0; q < (NBLK + 3) / 4; q++) {
                const int tb = part + 4 * q;                  // trailing block = column block J + 1 + tb
                if (tb < NBLK) {
                    const double* xs = xring + ((J + 1 + tb) % (NBLK + 1)) * 6;
#pragma unroll
                    for (int c = 0; c < 6; c += 2) {
                        const double2 t = *reinterpret_cast<const double2*>(urow + 8 + 6 * tb + c);
                        s = fma(-t.x, xs[c], s); s2 = fma(-t.y, xs[c + 1], s2);
                    }
                }
            }
            s += s2;
            s += __shfl_xor_sync(GXB_FULL, s, 1);
            s += __shfl_xor_sync(GXB_FULL, s, 2);
            double x[6];
            tri6_rows(s, head.y, pan, r, x);
            __syncwarp();
            if (lane < 6) {
                double xv = x[0];
#pragma unroll
                for (int k = 1; k < 6; k++) xv = lane == k ? x[k] : xv;
                xring[(J % (NBLK + 1)) * 6 + lane] = xv;
                sol[orig_row(6 * J + lane)] = sign * xv;
            }
            __syncwarp();
        }
    }
#ifdef GXB_TWO_TIMING
    t3 = clock64();
    if (blockIdx.x == 0 && lane == 0)
        printf("[two-sided timing dir=%d steps=%d] sweep %lld (%lld / step), wait+interface %lld, back substitution %lld (%lld / block)\n", dir, C1,
               t1 - t0, (t1 - t0) / max(C1, 1), t2 - t1, t3 - t2, (t3 - t2) / max(C1, 1));
#endif
    __syncthreads();
    return flags[0] | flags[1];
}
