// Batched pivoted block-banded LU + solve, one warp per instance, FP64.
//
// Replaces the `linsolve` of the Newton iteration: `x .= ImplicitAD.implicit_linear(A, b)` = SparseArrays `A \ b`
// (UMFPACK, symbolic + numeric factorisation on every call), src/analyses.jl:3518, 3585.
//
// The Jacobian of a chain is block-banded in the natural SystemIndices ordering: with 6x6 blocks the static system has
// block bandwidths KLB = KUB = 2.  Its diagonal blocks are structurally singular (free points have no
// ∂F/∂u), so partial pivoting is mandatory; row pivoting inside the band lets U grow to KLB+KUB block super-diagonals.
//
// Mapping: lane = one row of the active window (rows of block-rows J..J+KLB, at most 6(KLB+1) <= 32), the lane keeps its
// row segment (window columns = blocks J..J+KLB+KUB, W = 6(KLB+KUB+1) doubles) plus its right-hand side in registers.
// Per pivot: warp arg-max on the pivot column (two REDUX + one VOTE), the pivot lane stages its row through shared
// memory, every other live lane does W-k FMAs.  The pivot row, scaled by 1/pivot, is the U row and is streamed to global
// memory with one coalesced 32-double store; L is never stored (each Jacobian is used for exactly one right-hand side,
// which is eliminated on the fly).  After 6 pivots the window slides by one block: survivors shift their registers,
// the 6 retired lanes pick up the rows of block-row J+KLB+1, which were prefetched with cp.async during the elimination.
// Back substitution walks the blocks backwards, lane k < 6 owns row k of the block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GXB_FULL 0xffffffffu

template <int KLB, int KUB> struct LuShape {
    static constexpr int W = 6 * (KLB + KUB + 1);   // window / stored-row width
    static constexpr int ROWS = 6 * (KLB + 1);      // live rows in the window
    static constexpr int UW = ((W + 2 + 31) / 32) * 32;  // stored U row: W entries (scaled), rhs (scaled), pad
    // per-warp shared memory (doubles): 2 pivot-row buffers, staged next block-row (6 rows + 6 rhs), x window
    static constexpr int PROW = W + 2;
    static constexpr int SM_DOUBLES = 2 * PROW + 6 * W + 6 + W + 2;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Solve J y = rhs for one instance; writes sol = sign * y.  Returns 0, or 1 if an exactly-zero / non-finite pivot was hit.
//   Jrow : [n][W]  row-block storage (see gxb_static.cuh)      rhs : [n]
//   U    : [n][UW] scratch for the scaled U rows                sol : [n]
template <int KLB, int KUB>
__device__ int warp_band_solve(const double* __restrict__ Jrow, const double* __restrict__ rhs, double* __restrict__ U,
                               double* __restrict__ sol, int nb, double sign, double* sm) {
    using S = LuShape<KLB, KUB>;
    constexpr int W = S::W, ROWS = S::ROWS, UW = S::UW, PROW = S::PROW;
    const int lane = threadIdx.x & 31;
    const int n = 6 * nb;
    double* prow0 = sm;
    double* prow1 = sm + PROW;
    double* stage = sm + 2 * PROW;        // 6*W row entries, then 6 rhs
    double* xs = stage + 6 * W + 6;       // x window for the back substitution

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

    for (int J = 0; J < nb; J++) {
        // prefetch block-row J+KLB+1 (6 rows, contiguous 6*W doubles) and its rhs into shared memory
        const int nrow0 = 6 * (J + KLB + 1);
        const bool have_next = nrow0 < n;
        if (have_next) {
            const char* g = reinterpret_cast<const char*>(Jrow + (size_t)nrow0 * W);
            char* s = reinterpret_cast<char*>(stage);
            for (int c = lane; c < (6 * W * 8) / 16; c += 32) cp_async16(s + 16 * c, g + 16 * c);
            if (lane < 3) cp_async16(reinterpret_cast<char*>(stage + 6 * W) + 16 * lane, reinterpret_cast<const char*>(rhs + nrow0) + 16 * lane);
        }
        unsigned retired = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            double* prow = (k & 1) ? prow1 : prow0;
            // --- pivot search: max |a[k]| over live lanes, lowest lane wins ties ---
            const double v = fabs(a[k]);
            const unsigned hi = live ? (unsigned)__double2hiint(v) : 0u;
            const unsigned mhi = __reduce_max_sync(GXB_FULL, hi);
            const bool c1 = live && hi == mhi;
            const unsigned lo = c1 ? (unsigned)__double2loint(v) : 0u;
            const unsigned mlo = __reduce_max_sync(GXB_FULL, lo);
            const unsigned cand = __ballot_sync(GXB_FULL, c1 && lo == mlo);
            const int p = cand ? (__ffs(cand) - 1) : 0;
            const bool is_piv = (lane == p) && cand;
            if (is_piv) {
#pragma unroll
                for (int j = 0; j < W; j += 2) *reinterpret_cast<double2*>(prow + j) = make_double2(a[j], a[j + 1]);
                prow[W] = b;
                live = false;
                retired |= 1u;   // marker, the lane mask is rebuilt by ballot below
            }
            __syncwarp();
            const double piv = prow[k];
            if (!(fabs(piv) > 0.0) || !isfinite(piv) || !cand) bad = 1;
            const double rinv = 1.0 / piv;
            if (live) {
                const double m = a[k] * rinv;
                a[k] = 0.0;
#pragma unroll
                for (int j = k + 1; j < W; j++) a[j] = fma(-m, prow[j], a[j]);
                b = fma(-m, prow[W], b);
            }
            // scaled U row: entries 0..W-1, rhs at W; the diagonal slot k holds 1 and is never read back
            {
                double* urow = U + (size_t)(6 * J + k) * UW;
                for (int c = lane; c < UW; c += 32) urow[c] = (c <= W) ? prow[c] * rinv : 0.0;
            }
        }
        const unsigned rmask = __ballot_sync(GXB_FULL, retired != 0);
        // --- slide the window by one block ---
        if (have_next) cp_async_wait_all();
        __syncwarp();
        if (retired) {
            const int rank = __popc(rmask & ((1u << lane) - 1u));
            if (have_next && rank < 6) {
                const double* src = stage + rank * W;
#pragma unroll
                for (int j = 0; j < W; j += 2) { double2 t = *reinterpret_cast<const double2*>(src + j); a[j] = t.x; a[j + 1] = t.y; }
                b = stage[6 * W + rank];
                live = true;
            } else {
#pragma unroll
                for (int j = 0; j < W; j++) a[j] = 0.0;
                b = 0.0;
            }
        } else {
#pragma unroll
            for (int j = 0; j < W - 6; j++) a[j] = a[j + 6];
#pragma unroll
            for (int j = W - 6; j < W; j++) a[j] = 0.0;
        }
        __syncwarp();
    }

    // ---- back substitution, block by block from the end; lane k < 6 owns row 6J+k ----
    for (int c = lane; c < W + 2; c += 32) xs[c] = 0.0;
    __syncwarp();
    double u[W + 2];
    {
        const int J = nb - 1;
        if (lane < 6) {
            const double* urow = U + (size_t)(6 * J + lane) * UW;
#pragma unroll
            for (int j = 0; j < W + 2; j += 2) { double2 t = *reinterpret_cast<const double2*>(urow + j); u[j] = t.x; u[j + 1] = t.y; }
        }
    }
    for (int J = nb - 1; J >= 0; J--) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (lane < 6) {
            s0 = u[W];
#pragma unroll
            for (int j = 6; j < W; j += 4) {
                s0 = fma(-u[j], xs[j], s0);
                s1 = fma(-u[j + 1], xs[j + 1], s1);
                if (j + 2 < W) s2 = fma(-u[j + 2], xs[j + 2], s2);
                if (j + 3 < W) s3 = fma(-u[j + 3], xs[j + 3], s3);
            }
        }
        double s = (s0 + s1) + (s2 + s3);
        // prefetch the previous block's U rows while the 6x6 triangle is solved
        double un[W + 2];
        if (J > 0 && lane < 6) {
            const double* urow = U + (size_t)(6 * (J - 1) + lane) * UW;
#pragma unroll
            for (int j = 0; j < W + 2; j += 2) { double2 t = *reinterpret_cast<const double2*>(urow + j); un[j] = t.x; un[j + 1] = t.y; }
        }
#pragma unroll
        for (int k = 5; k >= 0; k--) {
            const double xk = __shfl_sync(GXB_FULL, s, k);
            if (lane < k) s = fma(-u[k], xk, s);
        }
        // slide the x window for block J-1: xs'[6..11] = this block's solution, xs'[j+6] = xs[j] for the older entries
        // (xs[0..5] is never read)
        double keep0 = (lane < 6) ? s : xs[lane];
        double keep1 = (lane + 32 < W - 6) ? xs[lane + 32] : 0.0;
        __syncwarp();
        if (lane < W - 6) xs[lane + 6] = keep0;
        if (lane + 32 < W - 6) xs[lane + 38] = keep1;
        if (lane < 6) sol[6 * J + lane] = sign * s;
        __syncwarp();
        if (J > 0) {
#pragma unroll
            for (int j = 0; j < W + 2; j++) u[j] = un[j];
        }
    }
    return __any_sync(GXB_FULL, bad) ? 1 : 0;
}
