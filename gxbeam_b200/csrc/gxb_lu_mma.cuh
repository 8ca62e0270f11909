// Batched pivoted band LU + solve, one warp per instance, FP64 — trailing update on the FP64 tensor cores (DMMA.8x8x4).
//
// Same job as warp_band_solve (gxb_lu.cuh): the `A \ b` of the Newton iteration, src/analyses.jl:3518, 3585 (UMFPACK in the reference).
// ncu on the row-per-lane kernel (profiles/r1_v5_prof_factor_raw.csv) showed an issue-bound kernel whose rank-6 trailing update is a real
// dense contraction (18 x 6 times 6 x 24 per block step: 144 DFMA + 72 LDS.128 warp instructions with 18 of 32 lanes live).  Here the
// active window lives in DMMA accumulator fragments instead, the elimination block is 8 columns (one fragment), and the rank-8 update of
// a 24 x 32 trailing window is 24 DMMA.8x8x4 instructions (12 tiles x 2 k-steps) — the same FP64-pipe time, a fifth of the issue slots.
//
// Window: SLOTS = 8 RT row slots x 8 CT columns (static: 24 x 40; window column 0 = global column 8 s at step s).
//   lane (g = lane / 4, c2 = lane % 4) holds c[rt][ct][h] = window(slot 8 rt + g, column 8 ct + 2 c2 + h)        (DMMA C/D fragment layout)
// Rows are stored in the row-block format of gxb_static.cuh (row r of 6-block B: W entries starting at global column 6 (B - KLB)); a row
// enters the window when the elimination reaches its first stored column and takes any free slot; pivot rows leave.  n is padded to a
// multiple of 8 with virtual identity rows.  Per step:
//   1. panel tile (ct = 0) -> shared memory, transposed to one row per lane (lane r = slot r);
//   2. panel factorisation, 8 pivots: single-REDUX arg-max, multipliers m[0..7] per row (as in gxb_lu.cuh);
//   3. pivot lanes publish L11; every row turns m into effective multipliers e = m L11^-1 (so the trailing update can use the UNMODIFIED
//      pivot rows), updates its right-hand side, publishes -e (A fragments), its pivot index and the new row its slot receives;
//   4. fragment layout: slots that became pivots publish their unmodified trailing part (B fragments);
//   5. 24 DMMA: c[rt][ct-1] = (-E) P + c[rt][ct]   — the window slides by one tile for free (D written to the shifted registers);
//   6. pivot slots now hold final U rows: stream them out;  7. free slots pick up the rows prefetched by TMA into a shared-memory ring.
// Back substitution: column-oriented over 8-blocks, 32 lanes = the 32 pending rows of the four blocks below the one being solved.
#pragma once
#include "gxb_lu.cuh"

#ifndef GXB_LU_MMA
#define GXB_LU_MMA 3        // default mask for the multi-kernel rounds: bit 0 static systems, bit 1 dynamic systems
#endif
// Shared memory per warp decides how much of the 256 KB per SM is left as L1 for the U rows the back substitution reads back: with a
// six-block ring (58 KB per CTA) a launch over 4096 instances took 1.69 ms, with a four-block ring (45 KB) 1.50 ms.
#ifndef GXB_MMA_NSA
#define GXB_MMA_NSA 1       // factorisation: TMA prefetch distance (steps); the ring needs GXB_MMA_RB >= blocks of NSA + 1 steps
#define GXB_MMA_RB 4
#endif
#ifndef GXB_MMA_ES
#define GXB_MMA_ES 10
#endif
#ifndef GXB_MMA_NSET
#define GXB_MMA_NSET 2      // back substitution: register prefetch distance (block steps)
#endif
#ifndef GXB_MMA_PD
#define GXB_MMA_PD 10       // back substitution: L2 prefetch distance (block steps)
#endif

template <int KLB, int KUB> struct MmaShape {
    static constexpr int W = 6 * (KLB + KUB + 1);                  // stored row width (row-block storage)
    static constexpr int RT = (6 * (KLB + 1) + 6 + 7) / 8;         // row tiles: live rows <= 6 (KLB + 1) + 6
    static constexpr int SLOTS = 8 * RT;
    static constexpr int CT = (6 * (KLB + KUB) + 13 + 7) / 8;      // column tiles: window extent <= 6 (KLB + KUB) + 13 columns
    static constexpr int TW = 8 * (CT - 1);                        // trailing width
    // U is stored per block-row of 8 pivot rows, tile-major: [CT-1 trailing tiles: 8 rows x 8][heads: 8 rows x (rhs, 1/pivot, panel 8)],
    // so that everything the back substitution reads is contiguous and sector-aligned
    static constexpr int BRW = 8 * TW + 80;                        // doubles per block-row
    static constexpr int HOFF = 8 * TW;                            // heads
    static constexpr int UW = BRW / 8;                             // doubles per pivot row (allocation)
    static constexpr int PS = TW + 4;                              // published pivot row stride  (== 4 mod 16: conflict-free B-fragment loads)
    static constexpr int AS = 10;                                  // transposed panel row stride (conflict-free 16-byte row reads)
    static constexpr int ES = GXB_MMA_ES;                          // effective-multiplier row stride (12: conflict-free A-fragment loads; 10: smaller)
    static constexpr int RB = GXB_MMA_RB;                          // ring of staged 6-row blocks
    static constexpr int NSA = GXB_MMA_NSA;                        // prefetch distance in steps
    static constexpr int NSB = NSA + 1;                            // mbarriers
    static constexpr int HS = 4;                                   // head ring of the back substitution
    // per-warp shared memory (doubles)
    static constexpr int SM_P = (8 * PS > SLOTS * AS) ? 8 * PS : SLOTS * AS;   // published pivot rows; aliased by the transposed panel
    static constexpr int SM_E = SLOTS * ES;
    static constexpr int SM_L = 0;                                 // L11 and the pivot right-hand sides (72 doubles) alias the pivot rows too
    static constexpr int SM_SLOT = SLOTS;                          // two int arrays: pivot index per slot, incoming row per slot
    static constexpr int SM_RING = RB * 6 * W;                     // staged rows (back substitution: head ring)
    static constexpr int SM_BAR = NSB + (NSB & 1);
    static constexpr int SM_DOUBLES = SM_P + SM_E + SM_L + SM_SLOT + SM_RING + SM_BAR;
    static_assert(SLOTS <= 32, "one lane per row slot");
    static_assert(PS % 16 == 4, "B-fragment bank layout");
    static_assert(HS * 80 + 8 <= SM_RING, "head ring + x block must fit the row ring");
    static_assert(SM_DOUBLES % 2 == 0 && SM_P % 2 == 0 && SM_E % 2 == 0 && SM_SLOT % 2 == 0 && ES % 2 == 0, "16-byte alignment of the carve-up");
    static_assert(SM_P >= 72, "L11 + pivot right-hand sides alias the pivot rows");
};

// D = A B + C, A 8x4 (row), B 4x8 (col), C/D 8x8:  a = A[g][c2], b = B[c2][g], c/d = {C[g][2 c2], C[g][2 c2 + 1]}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int KLB, int KUB>
__device__ void mma_back_substitute(const double* __restrict__ U, double* __restrict__ sol, int n, int nsteps, double sign, double* ring);

// doubles of L storage per elimination step (optional, see Lst below): 8 negated effective multipliers + the pivot index for each of 32 lanes
#define GXB_MMA_LSTEP (9 * 32)

// Solve J y = rhs for one instance; writes sol = sign * y.  Returns 0, or 1 if an exactly-zero / non-finite pivot was hit.
//   Jrow : [n][W] row-block storage      rhs : [n]      U : [n8][UW] scratch, n8 = n rounded up to 8      sol : [n]
//   Lst  : NULL, or [n8 / 8][GXB_MMA_LSTEP]: the factorisation records, per step and row slot, the negated effective multipliers and
//          which pivot the slot became, so that warp_band_resolve_mma can eliminate further right-hand sides without factorising again
//          (the eigen operator K^-1 M applies one factorisation ~70 times, src/analyses.jl:948-951)
template <int KLB, int KUB>
__device__ int warp_band_solve_mma(const double* __restrict__ Jrow, const double* __restrict__ rhs, double* __restrict__ U,
                                   double* __restrict__ sol, int nb, double sign, double* sm, double* __restrict__ Lst = nullptr) {
    using S = MmaShape<KLB, KUB>;
    constexpr int W = S::W, RT = S::RT, CT = S::CT, SLOTS = S::SLOTS, BRW = S::BRW, HOFF = S::HOFF, PS = S::PS, AS = S::AS, ES = S::ES, RB = S::RB;
    constexpr int NSA = S::NSA, NSB = S::NSB;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, c2 = lane & 3;
    const int n = 6 * nb;
    const int n8 = (n + 7) & ~7;
    const int nsteps = n8 >> 3;

    double* Psm = sm;                                  // [8][PS] published pivot rows (phases 4-5); phases 1-2: transposed panel [SLOTS][AS];
    double* Lsm = Psm;                                 // phase 3: L11 [8][8], then pb[8]
    double* pbs = Lsm + 64;
    double* Esm = Psm + S::SM_P;                       // [SLOTS][ES]
    int* slotkp = reinterpret_cast<int*>(Esm + S::SM_E);   // [SLOTS]
    int* slotnew = slotkp + SLOTS;                     // [SLOTS]
    double* ring = Esm + S::SM_E + S::SM_SLOT;         // [RB][6 W] staged rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S::SM_RING);

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NSB; q++) mbar_init(bars + q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncwarp();

    // Bookkeeping of the rows entering the window.  Real rows come in 6-row blocks: bb(s) = min(nb, KLB + 2 + s + s / 3) blocks have
    // entered before step s  (= (8 s + 7) / 6 + KLB + 1: the block of the last panel column plus KLB); the virtual identity rows that pad
    // n to a multiple of 8 come with the last block.
    auto blocks_before = [&](int s) { const int r = KLB + 2 + s + s / 3; return r < nb ? r : nb; };
    auto rows_of = [&](int bb) { return bb >= nb ? n8 : 6 * bb; };
    // TMA prefetch of the real rows entering after step t (blocks b0 .. b1-1): a 6-row block is contiguous in the row-block storage
    auto prefetch_rows = [&](int t, int b0, int b1) {
        if (lane != 0 || b1 <= b0) return;
        uint64_t* bar = bars + (t % NSB);
        mbar_expect_tx(bar, (unsigned)(b1 - b0) * (6 * W * 8));
        for (int B = b0; B < b1; B++) bulk_g2s(ring + (B % RB) * 6 * W, Jrow + (size_t)B * 6 * W, 6 * W * 8, bar);
    };
    // right-hand sides of the rows entering after a step: lane i fetches the one of the i-th incoming row, it is handed to the slot that
    // receives the row by a shuffle at the end of the step
    auto prefetch_rhs = [&](int r0, int r1) { return (r0 + lane < r1 && r0 + lane < n) ? rhs[r0 + lane] : 0.0; };

    double c[RT][CT][2];
    double b = 0.0;            // row-lane state: right-hand side of slot `lane`
    bool live = false;         //                 slot `lane` holds a row that has not been a pivot yet
    unsigned minkey = 0xffffffffu;

    // generic (slow) slot fill, used for the initial window and for the steps that bring in virtual rows:
    // slot <- row nr for the window starting at global column c0n; from_ring: the staged copy, else straight from global memory
    auto load_slot = [&](double (&cr)[CT][2], int nr, int c0n, bool from_ring) {
        if (nr < 0) return;
        if (nr < n) {
            const int B = nr / 6;
            const double* rowp = from_ring ? ring + (B % RB) * 6 * W + (nr - 6 * B) * W : Jrow + (size_t)nr * W;
            const int colbase = 6 * (B - KLB);
#pragma unroll
            for (int ct = 0; ct < CT; ct++) {
                const int j = c0n + 8 * ct + 2 * c2 - colbase;
                double2 v = make_double2(0.0, 0.0);
                if (j >= 0 && j < W) v = *reinterpret_cast<const double2*>(rowp + j);
                cr[ct][0] = v.x; cr[ct][1] = v.y;
            }
        } else {                                       // virtual identity row
#pragma unroll
            for (int ct = 0; ct < CT; ct++) {
                const int gc = c0n + 8 * ct + 2 * c2;
                cr[ct][0] = gc == nr ? 1.0 : 0.0; cr[ct][1] = gc + 1 == nr ? 1.0 : 0.0;
            }
        }
    };

    // ---- initial window: rows 0 .. rows_of(blocks_before(0)) - 1 in slots of the same index ----
    {
        const int r0 = rows_of(blocks_before(0));
#pragma unroll
        for (int rt = 0; rt < RT; rt++) {
#pragma unroll
            for (int ct = 0; ct < CT; ct++) c[rt][ct][0] = c[rt][ct][1] = 0.0;
            const int slot = 8 * rt + g;
            load_slot(c[rt], slot < r0 ? slot : -1, 0, false);
        }
        if (lane < r0) { live = true; b = lane < n ? rhs[lane] : 0.0; }
    }
    // The rows entering after step s + 1 are requested in the middle of step s, in the shadow of the tensor-core update (the DMMA burst
    // keeps the FP64 pipe busy for ~400 cycles during which the warp has nothing else to issue): one step of prefetch distance,
    // two steps' worth of blocks in the ring.
    static_assert(NSA == 1 && NSB == 2, "prefetch schedule");
    int bb0 = blocks_before(0), bb1 = blocks_before(1);
    prefetch_rows(0, bb0, bb1);
    double rpre = prefetch_rhs(rows_of(bb0), rows_of(bb1));

#ifdef GXB_LU_TIMING
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tq = clock64(), tn;
#define GXB_MTICK(i) do { tn = clock64(); tph[i] += tn - tq; tq = tn; } while (0)
#else
#define GXB_MTICK(i) do { } while (0)
#endif
    for (int s = 0; s < nsteps; s++) {
        const int c0 = 8 * s;
        const int rin0 = rows_of(bb0), rin1 = rows_of(bb1);      // rows entering after this step
        GXB_MTICK(0);
        // ---- 1. panel tile -> one row per lane ----
#pragma unroll
        for (int rt = 0; rt < RT; rt++)
            *reinterpret_cast<double2*>(Psm + (8 * rt + g) * AS + 2 * c2) = make_double2(c[rt][0][0], c[rt][0][1]);
        __syncwarp();
        double a[8];
        {
            const double* src = Psm + (lane < SLOTS ? lane : 0) * AS;
#pragma unroll
            for (int j = 0; j < 8; j += 2) { const double2 t = *reinterpret_cast<const double2*>(src + j); a[j] = t.x; a[j + 1] = t.y; }
        }
        __syncwarp();                                  // Psm is reused for the published pivot rows in phase 4
        GXB_MTICK(1);
        // ---- 2. panel factorisation ----
        // arg-max of |a[k]| over the live rows with ONE warp reduction: key = high word of |a[k]| (exponent + 20 mantissa bits: monotonic
        // in |a|) with the 5 low bits replaced by (31 - lane); rows that are not live have key 0.  Partial pivoting only needs a
        // near-maximal pivot (relative slack 2^-15); ties go to the lowest lane.  A zero pivot column shows up as a key < 32 (minkey).
        double m[8];
        int kp = 8;
        double myrinv = 1.0;
        unsigned kmask = live ? 0x7fffffe0u : 0u, lbits = live ? 31u - (unsigned)lane : 0u;
        double lf = live ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned key = ((unsigned)__double2hiint(a[k]) & kmask) | lbits;
            const double rc = fast_rcp(a[k]);
            const unsigned mk = __reduce_max_sync(GXB_FULL, key);
            minkey = min(minkey, mk);
            const int p = 31 - (int)(mk & 31u);
            const double rinv = __shfl_sync(GXB_FULL, rc, p);
            double alf = a[k] * lf;
            if (lane == p) { kmask = 0u; lbits = 0u; lf = 0.0; alf = 0.0; kp = k; myrinv = rinv; }
            m[k] = alf * rinv;
#pragma unroll
            for (int j = k + 1; j < 8; j++) {
                const double pj = __shfl_sync(GXB_FULL, a[j], p);
                a[j] = fma(-m[k], pj, a[j]);
            }
        }
        live = kmask != 0u;
        GXB_MTICK(2);
        if (kp < 8 && !isfinite(myrinv)) minkey = 0u;        // NaN / Inf pivot
        // ---- 3. publish L11 rows and pivot right-hand sides; panel part of the U rows goes straight to global memory ----
        if (kp < 8) {
            double* Lr = Lsm + kp * 8;
#pragma unroll
            for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(Lr + j) = make_double2(m[j], m[j + 1]);
            pbs[kp] = b;
            // the 8 x 8 triangle is stored scaled by 1/pivot of its row (the back substitution then has one FMA per level on its chain)
            double* urow = U + (size_t)s * BRW + HOFF + kp * 10 + 2;
#pragma unroll
            for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(urow + j) = make_double2(a[j] * myrinv, a[j + 1] * myrinv);
        }
        __syncwarp();
        // effective multipliers: e L11 = m, L11 unit lower triangular with L11[k][k'] = m_{pivot k}[k'];  kept negated (en = -e), which is
        // what the A fragments of D = A B + C need
        // Row-oriented: once en[k] is final, row k of L11 (contiguous: 16-byte loads) updates every en[kk < k]; the critical path is
        // one FMA per level.
        double en[8];
#pragma unroll
        for (int k = 0; k < 8; k++) en[k] = -m[k];
#pragma unroll
        for (int k = 7; k >= 1; k--) {
#pragma unroll
            for (int kk = 0; kk < k; kk += 2) {
                const double2 l = *reinterpret_cast<const double2*>(Lsm + k * 8 + kk);
                en[kk] = fma(-en[k], l.x, en[kk]);
                if (kk + 1 < k) en[kk + 1] = fma(-en[k], l.y, en[kk + 1]);
            }
        }
        {
            double b2 = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const double2 pb = *reinterpret_cast<const double2*>(pbs + k);
                b = fma(en[k], pb.x, b); b2 = fma(en[k + 1], pb.y, b2);
            }
            b += b2;
        }
        // which row does this slot receive?  free slots (never filled, or pivoted) take the incoming rows in order.  The slot's owner
        // publishes where the fragment lanes find the row: ring index of the entry of window column 0, and the shift of its first entry
        const bool virt = rin1 > n && rin1 > rin0;         // this step brings in virtual rows: generic path
        int newrow = -1, info = -1;
        const unsigned freem = __ballot_sync(GXB_FULL, !live) & ((SLOTS == 32) ? 0xffffffffu : ((1u << SLOTS) - 1u));
        const int rank = __popc(freem & ((1u << lane) - 1u));
        {
            if (!live && lane < SLOTS && rin0 + rank < rin1) {
                newrow = rin0 + rank;
                if (virt) info = newrow;
                else {
                    const int rb = (rank >= 6) + (rank >= 12);                // rin0 is a multiple of 6, at most 18 rows enter
                    const int B = bb0 + rb, slotB = B % RB, rr = rank - 6 * rb;
                    const int shift = 6 * (B - KLB) - (c0 + 8);            // 0, 2, 4 or 6
                    info = ((slotB * 6 + rr) * W - shift + 8) | (shift << 16);
                }
            }
        }
        if (lane < SLOTS) {
            double* Er = Esm + lane * ES;
#pragma unroll
            for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(Er + j) = make_double2(en[j], en[j + 1]);
            slotkp[lane] = kp;
            slotnew[lane] = info;
        }
        if (Lst) {
            double* Ls = Lst + (size_t)s * GXB_MMA_LSTEP + lane;
#pragma unroll
            for (int k = 0; k < 8; k++) Ls[32 * k] = en[k];
            Ls[32 * 8] = (double)kp;
        }
        if (kp < 8) *reinterpret_cast<double2*>(U + (size_t)s * BRW + HOFF + kp * 10) = make_double2(b, myrinv);   // final rhs of the U row, 1/pivot
        __syncwarp();
        GXB_MTICK(3);
        // ---- 4. pivot slots publish their unmodified trailing part ----
        int kps[RT];
#pragma unroll
        for (int rt = 0; rt < RT; rt++) {
            kps[rt] = slotkp[8 * rt + g];
            if (kps[rt] < 8) {
                double* Pr = Psm + kps[rt] * PS + 2 * c2;
#pragma unroll
                for (int ct = 1; ct < CT; ct++) *reinterpret_cast<double2*>(Pr + 8 * (ct - 1)) = make_double2(c[rt][ct][0], c[rt][ct][1]);
            }
        }
        __syncwarp();
        GXB_MTICK(4);
        // ---- 5. rank-8 update on the tensor cores, window slides one tile to the left ----
        {
            double af[RT][2];
#pragma unroll
            for (int rt = 0; rt < RT; rt++) { af[rt][0] = Esm[(8 * rt + g) * ES + c2]; af[rt][1] = Esm[(8 * rt + g) * ES + 4 + c2]; }
#pragma unroll
            for (int ct = 1; ct < CT; ct++) {
                const double b0 = Psm[c2 * PS + 8 * (ct - 1) + g], b1 = Psm[(4 + c2) * PS + 8 * (ct - 1) + g];
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    double d0, d1;
                    dmma884(d0, d1, af[rt][0], b0, c[rt][ct][0], c[rt][ct][1]);
                    dmma884(c[rt][ct - 1][0], c[rt][ct - 1][1], af[rt][1], b1, d0, d1);
                }
            }
#pragma unroll
            for (int rt = 0; rt < RT; rt++) c[rt][CT - 1][0] = c[rt][CT - 1][1] = 0.0;
        }
        // in the shadow of the DMMA burst: request the rows (and right-hand sides) of the next step
        const int bb2 = blocks_before(s + 2);
        prefetch_rows(s + 1, bb1, bb2);
        const double rpre_next = prefetch_rhs(rin1, rows_of(bb2));
        GXB_MTICK(5);
        // ---- 6. final U rows of the pivot slots ----
#pragma unroll
        for (int rt = 0; rt < RT; rt++) {
            if (kps[rt] < 8) {
                double* urow = U + (size_t)s * BRW + kps[rt] * 8 + 2 * c2;
#pragma unroll
                for (int ct = 0; ct < CT - 1; ct++) *reinterpret_cast<double2*>(urow + 64 * ct) = make_double2(c[rt][ct][0], c[rt][ct][1]);
            }
        }
        GXB_MTICK(6);
        // ---- 7. incoming rows ----
        if (rin1 > rin0) {
            if (bb1 > bb0) mbar_wait(bars + (s % NSB), (unsigned)(s / NSB) & 1u);
            if (virt) {
#pragma unroll
                for (int rt = 0; rt < RT; rt++) load_slot(c[rt], slotnew[8 * rt + g], c0 + 8, true);
            } else {
                // a real row entering now has its first stored entry at window column shift in {0,2,4,6} and W entries from there:
                // tiles 1 .. (W - 2) / 8 - 1 are always inside, the first and the last ones are range-checked
                const int w0 = 2 * c2;
#pragma unroll
                for (int rt = 0; rt < RT; rt++) {
                    const int inf = slotnew[8 * rt + g];
                    if (inf >= 0) {
                        const int shift = inf >> 16;
                        const double* rp = ring + ((inf & 0xffff) - 8) + w0;
#pragma unroll
                        for (int ct = 0; ct < CT; ct++) {
                            const int w = 8 * ct + w0;
                            double2 v;
                            if (8 * ct >= 6 && 8 * ct + 8 <= W) v = *reinterpret_cast<const double2*>(rp + 8 * ct);
                            else { v = make_double2(0.0, 0.0); if (w >= shift && w < shift + W) v = *reinterpret_cast<const double2*>(rp + 8 * ct); }
                            c[rt][ct][0] = v.x; c[rt][ct][1] = v.y;
                        }
                    }
                }
            }
            const double nbv = __shfl_sync(GXB_FULL, rpre, rank & 31);
            if (newrow >= 0) { live = true; b = nbv; }
        }
        bb0 = bb1; bb1 = bb2; rpre = rpre_next;
        __syncwarp();
        GXB_MTICK(7);
    }
#ifdef GXB_LU_TIMING
    const long long tfac_end = clock64();
#endif
    // an exactly-zero / sub-float-denormal pivot column (key without magnitude bits) or a non-finite 1/pivot
    const int bad = __any_sync(GXB_FULL, minkey < 32u) ? 1 : 0;

    __syncwarp();
    mma_back_substitute<KLB, KUB>(U, sol, n, nsteps, sign, ring);
#ifdef GXB_LU_TIMING
    if ((blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && threadIdx.x == 0)
        printf("[mma lu timing blk=%d steps=%d] cycles/step: prefetch %lld, transpose %lld, panel %lld, publish+e %lld, pubP %lld, dmma %lld, ustore %lld, newrows %lld | back-subst/block %lld\n",
               blockIdx.x, nsteps, tph[0] / nsteps, tph[1] / nsteps, tph[2] / nsteps, tph[3] / nsteps, tph[4] / nsteps, tph[5] / nsteps, tph[6] / nsteps, tph[7] / nsteps,
               (clock64() - tfac_end) / nsteps);
#endif
    // the shared memory is reused by the caller: retire the barrier objects
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NSB; q++) asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bars + q)) : "memory");
    }
    __syncwarp();
    return bad;
}

    // ---- back substitution, column-oriented over 8-blocks.  Row i needs x of the columns of its own block (the 8x8 triangle, solved
    // redundantly in every lane from the 80-byte heads staged in shared memory) and of the NT = CT-1 following blocks.  The lanes hold the
    // running sums of the 8 NT rows of the NT blocks below the one being solved: slot p of lane (q = lane / 8, k8 = lane % 8) owns row k8 of
    // the pending block congruent to q + 4 p mod NT; entries are prefetched NSET steps ahead straight from global memory into registers. ----
template <int KLB, int KUB>
__device__ void mma_back_substitute(const double* __restrict__ U, double* __restrict__ sol, int n, int nsteps, double sign, double* ring) {
    using S = MmaShape<KLB, KUB>;
    constexpr int CT = S::CT, BRW = S::BRW, HOFF = S::HOFF;
    const int lane = threadIdx.x & 31;
    {
        constexpr int HS = S::HS, NSET = GXB_MMA_NSET, PD = GXB_MMA_PD, NT = CT - 1, PEND = (8 * NT + 31) / 32;
        __syncwarp();
        double* hb = ring;                               // [HS][8][10] heads: rhs, 1/pivot, panel(8)
        double* xs = ring + HS * 80;                     // scaled sums of the block being solved
        const int q = lane >> 3, k8 = lane & 7;
        auto prefetch_head = [&](int Jb) {
            if (Jb >= 0) {
                const double* hsrc = U + (size_t)Jb * BRW + HOFF;
                cp_async16(hb + (Jb % HS) * 80 + 2 * lane, hsrc + 2 * lane);
                if (lane < 8) cp_async16(hb + (Jb % HS) * 80 + 64 + 2 * lane, hsrc + 64 + 2 * lane);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
        auto prefetch_l2 = [&](int Jb) {
            if (PD > 0 && Jb >= 0 && lane == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(U + (size_t)Jb * BRW), "r"(BRW * 8) : "memory");
        };
        // distance of slot p's block from block Jt (jm = Jt mod NT): 1 .. NT
        auto dist = [&](int jm, int p) { int d = jm - (q + 4 * p); if (d <= 0) d += NT; return d; };
        auto wrap = [&](int jm) { return jm < 0 ? jm + NT : jm; };
        auto load_step = [&](int Jt, int jm, double2 (&f)[PEND][4], double (&rin)[PEND]) {
#pragma unroll
            for (int p = 0; p < PEND; p++) {
                f[p][0] = f[p][1] = f[p][2] = f[p][3] = make_double2(0.0, 0.0); rin[p] = 0.0;
                if (Jt < 0 || q + 4 * p >= NT) continue;
                const int d = dist(jm, p), B = Jt - d;
                if (B < 0) continue;
                const double* rp = U + (size_t)B * BRW;
                const double* ep = rp + 64 * (d - 1) + 8 * k8;
#pragma unroll
                for (int i = 0; i < 4; i++) f[p][i] = *reinterpret_cast<const double2*>(ep + 2 * i);
                if (d == NT) rin[p] = rp[HOFF + 10 * k8];
            }
        };
        const int Jl = nsteps - 1;
        const int jml = Jl % NT;
        double ps[PEND];
#pragma unroll
        for (int p = 0; p < PEND; p++) {
            // the last block congruent to the slot's residue: its rows start from their right-hand sides
            const int Bq = Jl - (dist(jml, p) % NT);
            ps[p] = (q + 4 * p < NT && Bq >= 0) ? U[(size_t)Bq * BRW + HOFF + 10 * k8] : 0.0;
        }
        double2 f_[NSET][PEND][4]; double rin_[NSET][PEND];
#pragma unroll
        for (int t = 0; t < NSET; t++) load_step(Jl - t, wrap(jml - t), f_[t], rin_[t]);
#pragma unroll
        for (int t = 0; t < HS - 1; t++) prefetch_head(Jl - t);
        for (int t = NSET + 1; t < PD; t++) prefetch_l2(Jl - t);
        auto step = [&](int J, int jm, double2 (&f)[PEND][4], double (&rin)[PEND]) {
            prefetch_l2(J - PD);
            prefetch_head(J - (HS - 1));
            asm volatile("cp.async.wait_group %0;\n" ::"n"(HS - 1) : "memory");
            __syncwarp();
            // the owners of block J scale their finished sums by 1/pivot and hand them to everybody through shared memory
            const double* h = hb + (J % HS) * 80;
#pragma unroll
            for (int p = 0; p < PEND; p++)
                if (q + 4 * p == jm) xs[k8] = ps[p] * h[10 * k8 + 1];
            __syncwarp();
            double x[8];
#pragma unroll
            for (int k = 0; k < 8; k += 2) { const double2 t = *reinterpret_cast<const double2*>(xs + k); x[k] = t.x; x[k + 1] = t.y; }
            // 8 x 8 unit upper triangle (stored scaled), the same arithmetic in every lane
#pragma unroll
            for (int k = 7; k >= 1; k--) {
#pragma unroll
                for (int i = 0; i < k; i++) x[i] = fma(-h[10 * i + 2 + k], x[k], x[i]);
            }
            if (lane < 8 && 8 * J + lane < n) {
                double xv = x[0];
#pragma unroll
                for (int k = 1; k < 8; k++) xv = lane == k ? x[k] : xv;
                sol[8 * J + lane] = sign * xv;
            }
#pragma unroll
            for (int p = 0; p < PEND; p++) {
                if (q + 4 * p >= NT) continue;
                const int d = dist(jm, p);
                if (J - d >= 0) {
                    double acc = d == NT ? rin[p] : ps[p];
                    double acc2 = -f[p][0].y * x[1];
                    acc = fma(-f[p][0].x, x[0], acc); acc2 = fma(-f[p][1].y, x[3], acc2);
                    acc = fma(-f[p][1].x, x[2], acc); acc2 = fma(-f[p][2].y, x[5], acc2);
                    acc = fma(-f[p][2].x, x[4], acc); acc2 = fma(-f[p][3].y, x[7], acc2);
                    acc = fma(-f[p][3].x, x[6], acc);
                    ps[p] = acc + acc2;
                }
            }
            int jn = jm - NSET; if (jn < 0) jn += NT;
            load_step(J - NSET, jn, f, rin);
            __syncwarp();
        };
        static_assert(NSET <= NT, "mod-NT bookkeeping");
        for (int J = Jl, jm = jml; J >= 0; J -= NSET, jm = wrap(jm - NSET)) {
#pragma unroll
            for (int t = 0; t < NSET; t++)
                if (J - t >= 0) step(J - t, wrap(jm - t), f_[t], rin_[t]);
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
}

// Further right-hand sides with a recorded factorisation (Lst written by warp_band_solve_mma, U rows still in place): replays the
// elimination of the right-hand side step by step -- pivot slots publish their b, every slot adds its recorded -e . b_pivot, pivot slots
// write the final right-hand side of their U row, free slots take the incoming rows -- then back-substitutes.  sol = sign * (J \ rhs).
template <int KLB, int KUB>
__device__ void warp_band_resolve_mma(const double* __restrict__ Lst, const double* __restrict__ rhs, double* __restrict__ U,
                                      double* __restrict__ sol, int nb, double sign, double* sm) {
    using S = MmaShape<KLB, KUB>;
    constexpr int SLOTS = S::SLOTS, BRW = S::BRW, HOFF = S::HOFF;
    const int lane = threadIdx.x & 31;
    const int n = 6 * nb, n8 = (n + 7) & ~7, nsteps = n8 >> 3;
    double* pbs = sm;                                  // [8] right-hand sides of the step's pivot rows
    double* ring = sm + S::SM_P + S::SM_E + S::SM_SLOT;
    auto blocks_before = [&](int s) { const int r = KLB + 2 + s + s / 3; return r < nb ? r : nb; };
    auto rows_of = [&](int bb) { return bb >= nb ? n8 : 6 * bb; };
    double b = 0.0;
    bool live = false;
    {
        const int r0 = rows_of(blocks_before(0));
        if (lane < r0) { live = true; b = lane < n ? rhs[lane] : 0.0; }
    }
    int bb0 = blocks_before(0), bb1 = blocks_before(1);
    double en[8], kpd;
    auto load_l = [&](int s, double (&e)[8], double& k) {
        const double* Ls = Lst + (size_t)s * GXB_MMA_LSTEP + lane;
#pragma unroll
        for (int q = 0; q < 8; q++) e[q] = Ls[32 * q];
        k = Ls[32 * 8];
    };
    load_l(0, en, kpd);
    for (int s = 0; s < nsteps; s++) {
        const int rin0 = rows_of(bb0), rin1 = rows_of(bb1);
        // the incoming rows' right-hand sides and the next step's record are requested before this step's dependent chain
        const double rnew = (rin0 + lane < rin1 && rin0 + lane < n) ? rhs[rin0 + lane] : 0.0;
        double en_n[8], kpd_n = 8.0;
        if (s + 1 < nsteps) load_l(s + 1, en_n, kpd_n);
        const int kp = (int)kpd;
        if (kp < 8) { pbs[kp] = b; live = false; }
        __syncwarp();
        {
            double b2 = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k += 2) {
                const double2 pb = *reinterpret_cast<const double2*>(pbs + k);
                b = fma(en[k], pb.x, b); b2 = fma(en[k + 1], pb.y, b2);
            }
            b += b2;
        }
        if (kp < 8) U[(size_t)s * BRW + HOFF + kp * 10] = b;
        const unsigned freem = __ballot_sync(GXB_FULL, !live) & ((SLOTS == 32) ? 0xffffffffu : ((1u << SLOTS) - 1u));
        const int rank = __popc(freem & ((1u << lane) - 1u));
        const double nbv = __shfl_sync(GXB_FULL, rnew, rank & 31);
        if (!live && lane < SLOTS && rin0 + rank < rin1) { live = true; b = nbv; }
        else if (!live) b = 0.0;
        bb0 = bb1; bb1 = blocks_before(s + 2);
#pragma unroll
        for (int q = 0; q < 8; q++) en[q] = en_n[q];
        kpd = kpd_n;
        __syncwarp();
    }
    __syncwarp();
    mma_back_substitute<KLB, KUB>(U, sol, n, nsteps, sign, ring);
}
