// Partitioned ("substructured") batched pivoted band LU + solve for small per-GPU batches: P warps per instance.
//
// Same job as warp_band_solve / warp_band_solve_mma: the `A \ b` of the Newton iteration, src/analyses.jl:3518, 3585 (UMFPACK in the
// reference).  With one warp per instance a factorisation is a chain of n dependent pivots (~0.4 ms for n = 1206 however small the batch
// is), so a GPU that holds only a few hundred instances (BASELINE's target: 4096 cantilevers over 8 GPUs = 512 per GPU) idles.  Here the
// chain [p0 e0 p1 e1 ... pN] is cut into P partitions of consecutive stations; warp q owns the ROWS of partition q.
//
//   * A column can be eliminated inside a partition iff every row that touches it belongs to the partition ("covered" column).  The
//     rows of a partition restricted to its covered columns have full column rank whenever J is non-singular (they are complete columns
//     of J), so partial pivoting restricted to the partition's rows always finds a pivot, and the smallest singular value of that
//     sub-block is bounded below by J's: the partition sweep is as stable as the global band elimination.
//   * For the static chain (6-blocks: point i = block 2i, element i = block 2i+1, bandwidths 2/2) a cut in front of point m leaves the
//     18 columns of blocks p_{m-1}, e_{m-1}, p_m uncovered: the INTERFACE of the cut (15 would do -- p_{m-1}.u is covered by the upper
//     partition -- but three whole blocks keep every index block-aligned).  Partition q (row blocks [R0, R1)) eliminates the column
//     blocks [R0 + 1, R1 - 2) left to right exactly like warp_band_solve (lane = row of the window, 6-column panels, effective
//     multipliers, rank-6 trailing update), carrying its entries in the LEFT interface columns along as an 18-wide "spike" that is
//     updated like a right-hand side.  Rows that were never chosen as a pivot are left over: 12 (first partition), 18 (interior), 6 (last).
//   * The left-over rows of all partitions form the reduced system in the 18 (P - 1) interface unknowns (block bidiagonal); warp 0
//     eliminates it interface by interface with the same 6-column block step (at most 30 rows live) and back-substitutes.
//   * Every warp then back-substitutes its own partition from the two interface solutions next to it.
//
// U rows (50 doubles: rhs, 1/pivot, 6 panel, 24 trailing, 18 spike entries) are streamed to global memory by the sweep and read back by
// the back substitution, like the other two kernels; L is not stored.  Block rows are prefetched by TMA (cp.async.bulk + mbarrier).
#pragma once
#include "gxb_lu.cuh"

#define GXB_PART_MAXP 8

// partition q owns the row blocks [R0[q], R0[q+1]); R0[0] = 0, R0[P] = nb (number of 6-blocks)
struct PartPlan { int P; int R0[GXB_PART_MAXP + 1]; };

struct PartShape {
    static constexpr int W = 30, T = 24, SPK = 18, ROWS = 24;
    static constexpr int PS = T + 2 + SPK;                 // published pivot row: trailing, rhs, pad, spike
    static constexpr int UW = 2 + 6 + T + SPK;             // stored U row: rhs, 1/pivot, panel, trailing, spike
    static constexpr int NS = 3;                           // TMA ring depth (block rows)
    static constexpr int STG = 6 * W + 6;                  // a staged block row: 6 rows + 6 right-hand sides
    static constexpr int SM_P = 6 * PS, SM_L = 36, SM_STAGE = NS * STG;
    static constexpr int NB = 4;                           // back substitution: ring of staged U block rows (6 rows x UW)
    static constexpr int SM_URING = NB * 6 * UW, SM_XRING = 30, SM_BAR = NS + NB + 1;
    // the sweep's areas (published rows, L11, staged block rows) are reused by the left-over rows and then by the U ring
    static constexpr int SM_WORK = (SM_P + SM_L + SM_STAGE > SM_URING + SM_XRING) ? SM_P + SM_L + SM_STAGE : SM_URING + SM_XRING;
    static constexpr int WARP_DOUBLES = SM_WORK + SM_BAR;
    static constexpr int RW = 38;                          // a left-over row: 18 spike, 18 right-interface entries, rhs, pad
    static constexpr int LEFT_MAX = 18;
    // reduced system (warp 0): window = [interface t | interface t+1] = 36 columns
    static constexpr int RWIN = 36, RT = 30, RPS = RT + 2, RUW = 2 + 6 + RT;  // 38 doubles per reduced U row
    static_assert(LEFT_MAX * RW <= SM_WORK, "the left-over rows are dumped over the sweep's staging area");
    static_assert(PS % 2 == 0 && UW % 2 == 0 && RPS % 2 == 0 && RUW % 2 == 0 && RW % 2 == 0 && WARP_DOUBLES % 2 == 0 && SM_WORK % 2 == 0, "16-byte alignment");
    // doubles of shared memory per CTA of P warps
    __host__ __device__ static constexpr int cta_doubles(int P) {
        return P * WARP_DOUBLES                      // per-warp sweep areas
             + (P - 1) * 18 * RUW                    // reduced U rows
             + (P - 1) * 18 + 36                     // interface solutions (+ zero tail)
             + 6 * RPS + 36                          // published pivot rows / L11 of the reduced block step
             + 2 * P;                                // left-over row counts, bad flags (ints, two per double)
    }
};

// One 6-column block step of the lane-per-row elimination (phases A-C of warp_band_solve), window width W, SPK spike entries per row.
// On return: kp < 6 in the lanes whose row became pivot kp of this block (rinv = 1 / pivot); those lanes hold their final U row
// (a[0..W-7] = trailing entries, spk, b) and have written its panel part to ubase[kp * ustride + 2 .. + 7]; every other live lane holds
// its updated row; the window has moved one block to the right.
template <int W, int SPK>
__device__ __forceinline__ void lu_block_step(double (&a)[W], double (&spk)[SPK > 0 ? SPK : 1], double& b, bool& live, int& bad,
                                              double* __restrict__ Psm, double* __restrict__ Lsm, double* __restrict__ ubase, int ustride,
                                              int& kp, double& myrinv, const bool has_spk = true) {
    constexpr int T = W - 6, PS = T + 2 + SPK;
    const int lane = threadIdx.x & 31;
    double m[6];
    kp = 6; myrinv = 0.0;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        // arg-max of |a[k]| over the live rows with one warp reduction (see gxb_lu.cuh): near-maximal pivot, deterministic ties
        const unsigned fk = __float_as_uint(__double2float_ru(fabs(a[k])));
        const unsigned key = live ? ((fk & 0xffffffe0u) | (31u - (unsigned)lane)) : 0u;
        const double rc = fast_rcp(a[k]);
        const unsigned mk = __reduce_max_sync(GXB_FULL, key);
        const int p = 31 - (int)(mk & 31u);
        const double rinv = __shfl_sync(GXB_FULL, rc, p);
        const bool is_piv = (lane == p) && live;
        if (!(mk >> 5) || !isfinite(rinv)) bad = 1;
        if (is_piv) { live = false; kp = k; myrinv = rinv; }
        m[k] = live ? a[k] * rinv : 0.0;
#pragma unroll
        for (int j = k + 1; j < 6; j++) {
            const double pj = __shfl_sync(GXB_FULL, a[j], p);
            a[j] = fma(-m[k], pj, a[j]);
        }
    }
    if (kp < 6) {
        double* Lr = Lsm + kp * 6;
#pragma unroll
        for (int j = 0; j < 6; j += 2) *reinterpret_cast<double2*>(Lr + j) = make_double2(m[j], m[j + 1]);
        double* Pr = Psm + kp * PS;
#pragma unroll
        for (int j = 0; j < T; j += 2) *reinterpret_cast<double2*>(Pr + j) = make_double2(a[6 + j], a[6 + j + 1]);
        Pr[T] = b;
        if constexpr (SPK > 0) {
            if (has_spk) {
#pragma unroll
                for (int j = 0; j < SPK; j += 2) *reinterpret_cast<double2*>(Pr + T + 2 + j) = make_double2(spk[j], spk[j + 1]);
            }
        }
        double* urow = ubase + (ptrdiff_t)kp * ustride + 2;      // ustride may be negative (mirrored sweep of gxb_lu_two.cuh)
#pragma unroll
        for (int j = 0; j < 6; j += 2) *reinterpret_cast<double2*>(urow + j) = make_double2(a[j], a[j + 1]);
    }
    __syncwarp();
    // effective multipliers e L11 = m (the trailing update then uses the unmodified pivot rows)
    double e[6];
    e[5] = m[5];
#pragma unroll
    for (int kk = 4; kk >= 0; kk--) {
        double acc = m[kk];
#pragma unroll
        for (int k = kk + 1; k < 6; k++) acc = fma(-e[k], Lsm[k * 6 + kk], acc);
        e[kk] = acc;
    }
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
    if constexpr (SPK > 0) {
        if (has_spk)
#pragma unroll
        for (int j = 0; j < SPK; j += 2) {
            double x0 = spk[j], x1 = spk[j + 1];
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const double2 pv = *reinterpret_cast<const double2*>(Psm + k * PS + T + 2 + j);
                x0 = fma(-e[k], pv.x, x0);
                x1 = fma(-e[k], pv.y, x1);
            }
            spk[j] = x0; spk[j + 1] = x1;
        }
    }
#pragma unroll
    for (int j = T; j < W; j++) a[j] = 0.0;
    __syncwarp();          // Psm / Lsm may be rewritten by the next step
}

// 6 x 6 triangular solve of a block whose rows live in the lanes 4k (k = 0..5): s = reduced right-hand side of the lane's row, rinv and
// pan[] (panel part of the U row) likewise.  Every lane receives the six solutions.
__device__ __forceinline__ void tri6_rows(double s, double rinv, const double (&pan)[6], int r, double (&x)[6]) {
#pragma unroll
    for (int k = 5; k >= 0; k--) {
        const double xk = __shfl_sync(GXB_FULL, s * rinv, 4 * k);
        x[k] = xk;
        if (r < k) s = fma(-pan[k], xk, s);
    }
}

// Solve J y = rhs for one instance with the P warps of the CTA; sol = sign * y.  Every thread returns the same flag: 0, or 1 if an
// exactly-zero / non-finite pivot was hit.  Contains CTA-wide barriers: all P warps must call it.
//   Jrow : [n][30] row-block storage     rhs : [n]     U : [n][PartShape::UW] scratch     sol : [n]     sm : cta_doubles(P) doubles
__device__ int cta_part_solve(const double* __restrict__ Jrow, const double* __restrict__ rhs, double* __restrict__ U,
                              double* __restrict__ sol, const PartPlan& plan, double sign, double* sm) {
    using S = PartShape;
    constexpr int W = S::W, T = S::T, SPK = S::SPK, UW = S::UW, NS = S::NS, STG = S::STG, RW = S::RW;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5, P = plan.P;
    const int nb = plan.R0[P];
    const int R0 = plan.R0[q], R1 = plan.R0[q + 1];
    const int C0 = q == 0 ? 0 : R0 + 1, C1 = q == P - 1 ? nb : R1 - 2;

    double* wsm = sm + (size_t)q * S::WARP_DOUBLES;
    double* Psm = wsm;
    double* Lsm = Psm + S::SM_P;
    double* stage = Lsm + S::SM_L;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + S::SM_WORK);       // [NS] sweep ring, then [NB] back-substitution ring
    double* Ured = sm + (size_t)P * S::WARP_DOUBLES;                     // [(P-1) 18][RUW]
    double* xint = Ured + (size_t)(P - 1) * 18 * S::RUW;                 // [(P-1) 18 + 36]
    double* rPsm = xint + (P - 1) * 18 + 36;                             // [6][RPS]
    double* rLsm = rPsm + 6 * S::RPS;                                    // [6][6]
    int* nleft_s = reinterpret_cast<int*>(rLsm + 36);                    // [P]
    int* bad_s = nleft_s + P;                                            // [P]

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NS + S::NB; i++) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncwarp();

#ifdef GXB_PART_TIMING
    long long t_start = clock64(), t_fwd = 0, t_red = 0, t_bwd = 0;
#endif
    // ------------------------------------------------------------------ forward sweep of partition q
    double a[W], spk[SPK];
    double b = 0.0;
    bool live = false;
    int bad = 0;
    {
        const int nrows0 = 6 * (min(C0 + 3, R1) - R0);
        const int B = R0 + lane / 6;
        const int offw = 6 * (C0 - B + 2);               // window entry w <- stored entry w + offw
        const int offs = 6 * (R0 - B);                   // spike entry s  <- stored entry s + offs (left interface = blocks R0-2 .. R0)
        if (lane < nrows0) {
            live = true;
            const int row = 6 * R0 + lane;
            const double* src = Jrow + (size_t)row * W;
#pragma unroll
            for (int j = 0; j < W; j++) a[j] = (j + offw >= 0 && j + offw < W) ? src[j + offw] : 0.0;
#pragma unroll
            for (int j = 0; j < SPK; j++) spk[j] = (q > 0 && j + offs >= 0 && j + offs < W) ? src[j + offs] : 0.0;
            b = rhs[row];
        } else {
#pragma unroll
            for (int j = 0; j < W; j++) a[j] = 0.0;
#pragma unroll
            for (int j = 0; j < SPK; j++) spk[j] = 0.0;
        }
    }
    auto prefetch_rows = [&](int step) {                 // block row entering at the end of elimination step `step`
        const int Bn = step + 3;
        if (Bn < R1 && lane == 0) {
            const int slot = (step - C0) % NS;
            double* dst = stage + slot * STG;
            uint64_t* bar = bars + slot;
            mbar_expect_tx(bar, 6 * W * 8 + 48);
            bulk_g2s(dst, Jrow + (size_t)(6 * Bn) * W, 6 * W * 8, bar);
            bulk_g2s(dst + 6 * W, rhs + 6 * Bn, 48, bar);
        }
    };
#pragma unroll
    for (int i = 0; i < NS - 1; i++) prefetch_rows(C0 + i);

    for (int J = C0; J < C1; J++) {
        prefetch_rows(J + NS - 1);
        const bool have_next = J + 3 < R1;
        int kp; double myrinv;
        double* ubase = U + (size_t)(6 * J) * UW;
        lu_block_step<W, SPK>(a, spk, b, live, bad, Psm, Lsm, ubase, UW, kp, myrinv, q > 0);
        const unsigned rmask = __ballot_sync(GXB_FULL, kp < 6);
        if (have_next) mbar_wait(bars + ((J - C0) % NS), (unsigned)((J - C0) / NS) & 1u);
        if (kp < 6) {
            double* urow = ubase + (size_t)kp * UW;
            *reinterpret_cast<double2*>(urow) = make_double2(b, myrinv);
#pragma unroll
            for (int j = 0; j < T; j += 2) *reinterpret_cast<double2*>(urow + 8 + j) = make_double2(a[j], a[j + 1]);
            if (q > 0) {
#pragma unroll
                for (int j = 0; j < SPK; j += 2) *reinterpret_cast<double2*>(urow + 8 + T + j) = make_double2(spk[j], spk[j + 1]);
            }
            const int rank = __popc(rmask & ((1u << lane) - 1u));
#pragma unroll
            for (int j = 0; j < SPK; j++) spk[j] = 0.0;
            if (have_next) {
                const double* sbuf = stage + ((J - C0) % NS) * STG;
                const double* src = sbuf + rank * W;
#pragma unroll
                for (int j = 0; j < W; j += 2) { const double2 t = *reinterpret_cast<const double2*>(src + j); a[j] = t.x; a[j + 1] = t.y; }
                b = sbuf[6 * W + rank];
                live = true;
            } else {
#pragma unroll
                for (int j = 0; j < W; j++) a[j] = 0.0;
                b = 0.0;
            }
        }
        __syncwarp();
    }
    // the U rows were written with ordinary stores and come back through the async proxy (bulk copies) in the back substitution
    __threadfence_block();
    asm volatile("fence.proxy.async;\n" ::: "memory");
    // left-over rows -> the warp's (now idle) staging area: [18 spike][18 right-interface entries][rhs]
    {
        __syncwarp();
        const unsigned lm = __ballot_sync(GXB_FULL, live);
        const int rank = __popc(lm & ((1u << lane) - 1u));
        const int cnt = __popc(lm);
        if (live && rank < S::LEFT_MAX) {
            double* o = wsm + rank * RW;
#pragma unroll
            for (int j = 0; j < SPK; j += 2) *reinterpret_cast<double2*>(o + j) = make_double2(spk[j], spk[j + 1]);
#pragma unroll
            for (int j = 0; j < 18; j += 2) *reinterpret_cast<double2*>(o + 18 + j) = make_double2(a[j], a[j + 1]);
            o[36] = b;
        }
        const int expect = P == 1 ? 0 : (q == 0 ? 12 : (q == P - 1 ? 6 : 18));
        if (cnt != expect) bad = 1;
        if (lane == 0) nleft_s[q] = cnt < S::LEFT_MAX ? cnt : S::LEFT_MAX;
        bad = __any_sync(GXB_FULL, bad) ? 1 : 0;
        if (lane == 0) bad_s[q] = bad;
    }
#ifdef GXB_PART_TIMING
    t_fwd = clock64();
#endif
    __syncthreads();

    // ------------------------------------------------------------------ reduced system in the interface unknowns (warp 0)
    if (q == 0 && P > 1) {
        constexpr int RWIN = S::RWIN, RUW = S::RUW;
        double w[RWIN], dummy[1] = {0.0};
        double rb = 0.0;
        bool rlive = false;
        int rbad = 0;
#pragma unroll
        for (int j = 0; j < RWIN; j++) w[j] = 0.0;
        for (int c = lane; c < 36; c += 32) xint[(P - 1) * 18 + c] = 0.0;
        // rows of partition 0: entries on interface 0 only
        if (lane < nleft_s[0]) {
            const double* o = sm + lane * RW;
#pragma unroll
            for (int j = 0; j < 18; j += 2) { const double2 t = *reinterpret_cast<const double2*>(o + 18 + j); w[j] = t.x; w[j + 1] = t.y; }
            rb = o[36];
            rlive = true;
        }
        for (int t = 0; t < P - 1; t++) {
            // rows of partition t+1 go to the free lanes: [spike = interface t | right part = interface t+1]
            {
                const unsigned freem = __ballot_sync(GXB_FULL, !rlive);
                const int rank = __popc(freem & ((1u << lane) - 1u));
                if (!rlive && rank < nleft_s[t + 1]) {
                    const double* o = sm + (size_t)(t + 1) * S::WARP_DOUBLES + rank * RW;
#pragma unroll
                    for (int j = 0; j < RWIN; j += 2) { const double2 v = *reinterpret_cast<const double2*>(o + j); w[j] = v.x; w[j + 1] = v.y; }
                    rb = o[36];
                    rlive = true;
                }
            }
#pragma unroll 1
            for (int j = 0; j < 3; j++) {
                int kp; double myrinv;
                double* ubase = Ured + (size_t)(t * 18 + 6 * j) * RUW;
                lu_block_step<RWIN, 0>(w, dummy, rb, rlive, rbad, rPsm, rLsm, ubase, RUW, kp, myrinv);
                if (kp < 6) {
                    double* urow = ubase + (size_t)kp * RUW;
                    *reinterpret_cast<double2*>(urow) = make_double2(rb, myrinv);
#pragma unroll
                    for (int c = 0; c < S::RT; c += 2) *reinterpret_cast<double2*>(urow + 8 + c) = make_double2(w[c], w[c + 1]);
#pragma unroll
                    for (int c = 0; c < RWIN; c++) w[c] = 0.0;
                    rb = 0.0;
                }
                __syncwarp();
            }
        }
        if (__any_sync(GXB_FULL, rlive)) rbad = 1;            // every left-over row must have become a pivot
        // back substitution, interface by interface: lanes (r = lane / 4, part = lane % 4), 6 rows x 4 slices of the 30 trailing entries
        {
            const int r = lane >> 2, part = lane & 3;
            for (int blk = 3 * (P - 1) - 1; blk >= 0; blk--) {
                const double* urow = Ured + (size_t)(6 * blk + (r < 6 ? r : 0)) * RUW;
                const double* xw = xint + 6 * blk + 6;
                double s = part == 0 ? urow[0] : 0.0;
                const double rinv = urow[1];
                double pan[6];
#pragma unroll
                for (int k = 0; k < 6; k++) pan[k] = urow[2 + k];
#pragma unroll
                for (int c = 0; c < 8; c++) { const int cc = 8 * part + c; if (cc < S::RT) s = fma(-urow[8 + cc], xw[cc], s); }
                s += __shfl_xor_sync(GXB_FULL, s, 1);
                s += __shfl_xor_sync(GXB_FULL, s, 2);
                double x[6];
                tri6_rows(s, rinv, pan, r, x);
                __syncwarp();
                if (lane < 6) {
                    double xv = x[0];
#pragma unroll
                    for (int k = 1; k < 6; k++) xv = lane == k ? x[k] : xv;
                    xint[6 * blk + lane] = xv;
                    // interface t = blk / 3 holds the column blocks R0[t+1] - 2 .. R0[t+1]
                    const int t = blk / 3;
                    sol[6 * (plan.R0[t + 1] - 2 + (blk - 3 * t)) + lane] = sign * xv;
                }
                __syncwarp();
            }
        }
        rbad = __any_sync(GXB_FULL, rbad) ? 1 : 0;
        if (lane == 0 && rbad) bad_s[0] = 1;
    }
    __syncthreads();
#ifdef GXB_PART_TIMING
    t_red = clock64();
#endif

    // ------------------------------------------------------------------ back substitution of partition q
    // U block rows (6 rows x 50 doubles, contiguous) come back through a TMA ring in shared memory; lanes (r = lane / 4, part = lane % 4):
    // row r of the block, slice `part` of its 24 trailing entries (= column block J+1+part) and of its 18 spike entries.
    {
        constexpr int NB = S::NB;
        const int r = lane >> 2, part = lane & 3;
        double* uring = wsm;                                   // [NB][6][UW]
        double* xring = wsm + S::SM_URING;                     // [5][6]: x of column block Bk in slot Bk % 5
        uint64_t* bbars = bars + NS;
        const int nblk = C1 - C0;
        auto prefetch_u = [&](int i) {                         // i-th block of the descent: J = C1 - 1 - i
            if (i < nblk && lane == 0) {
                uint64_t* bar = bbars + (i % NB);
                mbar_expect_tx(bar, 6 * UW * 8);
                bulk_g2s(uring + (i % NB) * 6 * UW, U + (size_t)(6 * (C1 - 1 - i)) * UW, 6 * UW * 8, bar);
            }
        };
        // the staging area was last written with ordinary stores (left-over rows): order them before the async-proxy writes
        if (lane == 0) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NB - 1; i++) prefetch_u(i);
        // blocks C1 .. C1+3: the right interface (three blocks), then structural zeros
        if (lane < 24) {
            const int Bk = C1 + lane / 6;
            double v = 0.0;
            if (q < P - 1 && lane < 18) v = xint[q * 18 + lane];
            xring[(Bk % 5) * 6 + lane % 6] = v;
        }
        double xl[6];
#pragma unroll
        for (int j = 0; j < 6; j++) xl[j] = (q > 0 && part < 3) ? xint[(q - 1) * 18 + 6 * part + j] : 0.0;
        __syncwarp();
        for (int i = 0; i < nblk; i++) {
            const int J = C1 - 1 - i;
            prefetch_u(i + NB - 1);
            mbar_wait(bbars + (i % NB), (unsigned)(i / NB) & 1u);
            const double* urow = uring + (i % NB) * 6 * UW + (r < 6 ? r : 0) * UW;
            const double2 head = *reinterpret_cast<const double2*>(urow);
            double pan[6];
#pragma unroll
            for (int k = 0; k < 6; k += 2) { const double2 t = *reinterpret_cast<const double2*>(urow + 2 + k); pan[k] = t.x; pan[k + 1] = t.y; }
            const double* xs = xring + ((J + 1 + part) % 5) * 6;
            double s = part == 0 ? head.x : 0.0, s2 = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c += 2) {
                const double2 t = *reinterpret_cast<const double2*>(urow + 8 + 6 * part + c);
                s = fma(-t.x, xs[c], s); s2 = fma(-t.y, xs[c + 1], s2);
            }
            if (q > 0 && part < 3) {
#pragma unroll
                for (int c = 0; c < 6; c += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(urow + 8 + T + 6 * part + c);
                    s = fma(-t.x, xl[c], s); s2 = fma(-t.y, xl[c + 1], s2);
                }
            }
            s += s2;
            s += __shfl_xor_sync(GXB_FULL, s, 1);
            s += __shfl_xor_sync(GXB_FULL, s, 2);
            double x[6];
            tri6_rows(s, head.y, pan, r, x);
            __syncwarp();                                      // everybody has read the x slot that is overwritten now (block J+5 = slot J % 5)
            if (lane < 6) {
                double xv = x[0];
#pragma unroll
                for (int k = 1; k < 6; k++) xv = lane == k ? x[k] : xv;
                xring[(J % 5) * 6 + lane] = xv;
                sol[6 * J + lane] = sign * xv;
            }
            __syncwarp();
        }
    }
#ifdef GXB_PART_TIMING
    t_bwd = clock64();
    if (blockIdx.x == 0 && lane == 0)
        printf("[part timing P=%d q=%d steps=%d] forward %lld (%lld / step), wait+reduced %lld, backward %lld (%lld / block)\n", P, q, C1 - C0,
               t_fwd - t_start, (t_fwd - t_start) / max(C1 - C0, 1), t_red - t_fwd, t_bwd - t_red, (t_bwd - t_red) / max(C1 - C0, 1));
#endif
    __syncthreads();
    int anybad = 0;
    for (int i = 0; i < P; i++) anybad |= bad_s[i];
    return anybad;
}
