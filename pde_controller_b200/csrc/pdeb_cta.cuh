/*
 * pdeb_cta.cuh — "CTA family": one BLOCK of 128 threads owns one trajectory of a refined mesh (up to 1150 elements;
 * BASELINE.json configs[3]: 1,024 elements x 10,000 steps).  North star: "warp- or CTA-per-trajectory ... staged in
 * shared memory".
 *
 * Same mathematics as the chunk family (pdeb_chunk.cuh): heat, generalised trapezoidal rule in increment form on the
 * deviation w = d - g0,  A (x' - x) = dt (f - K x)  with the right-hand side in element-flux form, restating what
 * femformal's heatlinfem_mix system and APCont / cspec robustness do per sample behind eval_robustness_wrapper.py:16
 * of the reference.  What differs is the mapping:
 *
 *   rows per lane R = 8 instead of 32, so EVERYTHING a lane needs per step lives in registers: its 8 state values,
 *   its own twisted L D L^T factors, its two spike vectors and its flux coefficients.  The time loop has no table
 *   loads at all (the chunk family at R = 32 read three 16-byte pairs per row and step from shared memory and spilt:
 *   553 M local loads, 31 % LSU pipe, 36 % FP64 pipe).
 *
 * Layout.  The N free rows are end-aligned onto 128 x 9 - 1 = 1151 positions: position p = row + shift.  Thread t owns
 * the interior positions 9 t + 1 .. 9 t + 8 and, for t < 127, the separator 9 t + 9 behind them.
 *
 * Per step the constant SPD tridiagonal system is solved by TWO-LEVEL substructuring:
 *   1. every lane solves its interior block (registers, two elimination chains meeting in the middle);
 *   2. the 127 separator unknowns satisfy a tridiagonal Schur system.  The separators owned by lane 31 of warps 0..2
 *      are LEVEL-2 unknowns B0..B2; the 31 others of each warp are solved INSIDE the warp by parallel cyclic reduction
 *      with precomputed coefficients (5 shuffle rounds) as  sigma = tau - B_left P - B_right Q,  where P and Q are the
 *      warp's precomputed responses to its two level-2 neighbours;
 *   3. B0..B2 satisfy a 3 x 3 system whose inverse is precomputed; its right-hand side needs three numbers from every
 *      warp, exchanged through shared memory behind THE ONE block barrier of the step (double-buffered slots);
 *   4. interiors are corrected with the two spike vectors (2 FMA per row).
 * Between barriers a lane needs two things from the neighbouring WARP: the separator in front of lane 0 (kept as a ghost
 * copy that is advanced with the owner's own arithmetic, so it agrees bit for bit) and, in the level-2 row of lane 31, the
 * first row behind it (that term is contributed by its owner through the same exchange).
 *
 * STL: G clauses accumulate a per-lane running maximum over steps AND nodes (max is associative), reduced across the
 * block once, after the time loop; F clauses need the block-wide maximum of every step, which rides on the step's
 * barrier.  All control flow is block-uniform (one trajectory per block).
 */
#ifndef PDEB_CTA_CUH
#define PDEB_CTA_CUH

#include <cuda_runtime.h>

#include "pdeb_chunk.cuh"

#ifndef PDEB_CTA_MINBLOCKS
#define PDEB_CTA_MINBLOCKS 2      /* 4 x 8 shape: no register cap */
#endif
#ifndef PDEB_CTA2_MINBLOCKS
#define PDEB_CTA2_MINBLOCKS 4     /* 2 x 16 shape, 64 threads: no register cap (a 168-register build spills and is 1.6x slower) */
#endif

/*
 * Shared memory of one block.  WARPS x R is the shape of the mapping: 4 x 8 (128 threads, every table in registers) or
 * 2 x 16 (64 threads; the spike vectors and flux coefficients then live in per-lane columns of shared memory, TAB = true:
 * 16 rows of them do not fit the register file next to the state and the factors; class-shared tables for ALL per-row
 * constants, as in the chunk family, were measured slower: 87 against 69 ms, the factors sit on the dependent chain).
 */
template <int WARPS, int R, bool TAB>
struct PdebCtaShared {
    static constexpr int NT = 32 * WARPS;
    double ex[2][WARPS][4];                /* per step, double-buffered: [warp]{ y[0] of lane 0, tau of lane 0, level-2 piece of lane 31 } */
    double fx[2][WARPS][PDEB_MAX_PRED];    /* per step: warp maxima of the open F clauses */
    double G[3][3];                        /* inverse of the level-2 matrix (WARPS - 1 unknowns) */
    double afn[3], supB[3];                /* a_first_next and Sup of the level-2 separators */
    double rsr[WARPS];                     /* flux coefficient right of every warp's last separator */
    double scr[16];                        /* set-up and rare-path scratch */
    double set[WARPS][8];                  /* set-up exchange: per warp { P0, Q0, P30, Q30, SloB, SdiB, SupB, afn } */
    double nb[NT + 1];                     /* set-up: neighbour exchange */
    double gacc[PDEB_MAX_PRED][NT];        /* per-lane running maxima of the G clauses */
    PdebD2 vw[TAB ? R : 1][NT];            /* TAB: { v_l, w_l } spikes, one column per lane (conflict-free) */
    double kqs[TAB ? R + 1 : 1][NT];       /* TAB: flux coefficients, one column per lane */
    double cval[PDEB_MAX_PRED];            /* clause values for the formula tree */
    /* clause data (block-uniform, written by thread j at set-up) */
    double c_mh[PDEB_MAX_PRED], c_c0[PDEB_MAX_PRED], c_acc[PDEB_MAX_PRED];
    int32_t c_klo[PDEB_MAX_PRED], c_khi[PDEB_MAX_PRED], c_ilo[PDEB_MAX_PRED], c_ihi[PDEB_MAX_PRED];
    uint32_t c_info[PDEB_MAX_PRED];
    int32_t flag;
};

template <int KIND, int WARPS, int R, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB)
pdeb_cta_kernel(const __grid_constant__ PdebDevBatch b, const __grid_constant__ PdebConventions cv,
                const __grid_constant__ PdebDevOut out, const int32_t* __restrict__ order, int64_t first, int64_t count,
                  int gather) {
    static_assert(KIND == PDEB_KIND_HEAT, "the CTA family is instantiated for the heat equation");
    constexpr int NT = 32 * WARPS, CAP = NT * (R + 1) - 1, M = R / 2, NB = WARPS - 1;   /* NB: level-2 unknowns */
    constexpr bool TAB = R > 8;
    static_assert(WARPS >= 2 && WARPS <= 4 && R <= 32, "two to four warps per trajectory");
    __shared__ PdebCtaShared<WARPS, R, TAB> sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t = blockIdx.x;
    if (t >= count) return;
    const int64_t so = order[first + t];           /* the sample's index in the caller's order: outputs */
    const int64_t s = gather ? so : first + t;     /* inputs: the image as uploaded, or the plan-ordered image */
    const int64_t st = b.stride;

    /* ---- parameters (block-uniform) ---- */
    const int N = b.n_elem[s];
    const double L = b.L[s], dt = b.dt[s], T = b.T[s], g0 = b.g0[s];
    const double x_a = b.x_a[s], x_b = b.x_b[s];
    const double rho_in = b.rho_in[s], rho_out = b.rho_out[s], E_in = b.E_in[s], E_out = b.E_out[s];
    int np = b.n_pred[s];
    const int nk = b.n_knots[s], nt = b.n_tree[s];
    int status = PDEB_OK;
    int64_t S = 0;
    if (!pdeb_params_ok(KIND, N, L, dt, T, g0, rho_in, rho_out, E_in, E_out) || N > CAP - 1) status = PDEB_ST_BADPARAM;
    if (np < 1 || np > b.pmax || nk < 0 || nk > b.kmax || nt < 1 || nt > b.tmax) status = PDEB_ST_BADPARAM;
    const double* kt = b.knot_t ? b.knot_t + s : nullptr;      /* NULL: uniform knots */
    const double* ky = b.knot_y + s;
    if (status == PDEB_OK) {
        for (int j = 1; j < nk && kt; ++j)
            if (!(kt[(int64_t)j * st] > kt[(int64_t)(j - 1) * st])) status = PDEB_ST_BADPARAM;
        for (int j = 0; j < np; ++j)
            if (b.pred_rel[(int64_t)j * st + s] > 2 || b.pred_top[(int64_t)j * st + s] > 1) status = PDEB_ST_BADPARAM;
    }
    if (status == PDEB_OK) {
        S = pdeb_step_count(T, dt, cv.steps_rule);
        if (S > PDEB_MAX_STEPS) status = PDEB_ST_BADPARAM;
    }
    const int shift = CAP - N;                    /* position = row + shift */
    const int base = tid * (R + 1);               /* interior positions base+1 .. base+R, separator base+R+1 */
    const double h = L / N;
    const int lumped = cv.mass_heat;
    PdebRowCoef rc;
    {
        const double cA = cv.theta * dt, cR = (1.0 - cv.theta) * dt;
        const double m_o = rho_out * h, k_o = E_out / h, m_i = rho_in * h, k_i = E_in / h;
        const double mo_o = lumped ? 0.0 : m_o / 6.0, mo_i = lumped ? 0.0 : m_i / 6.0;
        const double md_o = lumped ? 0.5 * m_o : m_o / 3.0, md_i = lumped ? 0.5 * m_i : m_i / 3.0;
        rc.ad_o = md_o + cA * k_o; rc.ad_i = md_i + cA * k_i;
        rc.ao_o = mo_o - cA * k_o; rc.ao_i = mo_i - cA * k_i;
        rc.rd_o = md_o - cR * k_o; rc.rd_i = md_i - cR * k_i;
        rc.ro_o = mo_o + cR * k_o; rc.ro_i = mo_i + cR * k_i;
        rc.kq_o = dt * k_o; rc.kq_i = dt * k_i;
        rc.h = h; rc.x_a = x_a; rc.x_b = x_b; rc.N = (status == PDEB_OK) ? N : 0; rc.sampling = cv.material_sampling;
    }
    if (status != PDEB_OK) { S = 0; np = 0; }
    const int g_first = base + 1 - shift;         /* row of the lane's first interior position */
    const int g_sep = g_first + R;                /* row of its separator */

    /* ---- per-lane factors, all in registers ---- */
    double x[R], y[R];                            /* state (deviation) and work vector of the interior rows */
    double e[R], dinv[R], v[R], w[R];             /* multipliers, pivot inverses, the two spike vectors */
    double kq[R + 1];                             /* flux coefficients: kq[0] element left of row 0, kq[l+1] element right of row l */
    double a_last, a_first, a_first_next, rs_r;   /* couplings of the separator; flux coefficient right of the separator */
    bool ok = true;
    {
        double Ad[R], Ao[R];
#pragma unroll
        for (int l = 0; l < R; ++l) {
            double Rd, Ro;
            rc.row(g_first + l, Ad[l], Ao[l], Rd, Ro);
            kq[l + 1] = rc.kq(g_first + l);
        }
        kq[0] = rc.kq(g_first - 1);
        a_last = Ao[R - 1];
        {
            double Adp, Rdp, rf;
            rc.row(g_first - 1, Adp, a_first, Rdp, rf);
            if (tid == 0) a_first = 0.0;
        }
        /* twisted L D L^T of the interior block: top sweep 0..M-1, bottom sweep R-1..M+1, twist row M */
        double dprev = 0.0;
#pragma unroll
        for (int l = 0; l < M; ++l) {
            double piv = Ad[l];
            if (l > 0) piv -= (Ao[l - 1] * dprev) * Ao[l - 1];
            if (!(piv > 0.0)) ok = false;
            dprev = 1.0 / piv;
            e[l] = Ao[l] * dprev;                 /* edge (l, l+1) eliminated downwards */
            dinv[l] = dprev;
        }
        const double dtop = dprev;
        dprev = 0.0;
#pragma unroll
        for (int l = R - 1; l > M; --l) {
            double piv = Ad[l];
            if (l < R - 1) piv -= (Ao[l] * dprev) * Ao[l];            /* edge (l, l+1) eliminated upwards */
            if (!(piv > 0.0)) ok = false;
            const double e_up = (l < R - 1) ? Ao[l] * dprev : 0.0;
            dprev = 1.0 / piv;
            e[l] = e_up;
            dinv[l] = dprev;
        }
        {
            const double e_dn = (M > 0) ? Ao[M - 1] * dtop : 0.0;     /* equals e[M-1] */
            const double e_up = (M < R - 1) ? Ao[M] * dprev : 0.0;
            const double piv = Ad[M] - ((M > 0) ? e_dn * Ao[M - 1] : 0.0) - ((M < R - 1) ? e_up * Ao[M] : 0.0);
            if (!(piv > 0.0)) ok = false;
            const double dm = 1.0 / piv;
            if (!isfinite(dm) || !isfinite(dtop) || !isfinite(dprev)) ok = false;
            e[M] = e_up;
            dinv[M] = dm;
        }
    }
    /* twisted solve of the interior block IN PLACE: z holds the right-hand side on entry, the solution on return */
    auto solve_with = [&](double (&z)[R], auto E, auto DI) {
        double zt = 0.0, zb = 0.0;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int it = j, ib = R - 1 - j;
            if (it < M) {
                const double q = (it > 0) ? fma(-E(it - 1), zt, z[it]) : z[it];
                zt = q;
                z[it] = q * DI(it);
            }
            if (ib > M) {
                const double q = (ib < R - 1) ? fma(-E(ib), zb, z[ib]) : z[ib];
                zb = q;
                z[ib] = q * DI(ib);
            }
        }
        {
            double q = z[M];
            if (M > 0) q = fma(-E(M - 1), zt, q);
            if (M < R - 1) q = fma(-E(M), zb, q);
            z[M] = q * DI(M);
        }
#pragma unroll
        for (int l = M - 1; l >= 0; --l) z[l] = fma(-E(l), z[l + 1], z[l]);
#pragma unroll
        for (int l = M + 1; l < R; ++l) z[l] = fma(-E(l - 1), z[l - 1], z[l]);
    };
    auto local_solve = [&](double (&z)[R]) { solve_with(z, [&](int l) { return e[l]; }, [&](int l) { return dinv[l]; }); };
    {
#pragma unroll
        for (int l = 0; l < R; ++l) v[l] = (l == 0) ? a_first : 0.0;
        local_solve(v);
#pragma unroll
        for (int l = 0; l < R; ++l) w[l] = (l == R - 1) ? a_last : 0.0;
        local_solve(w);
    }
    /* neighbour exchange through shared memory (set-up only): value of thread tid + d (0 outside the block) */
    auto from_tid = [&](double val, int d) {
        __syncthreads();
        sm.nb[tid] = val;
        __syncthreads();
        const int src = tid + d;
        return (src >= 0 && src < NT) ? sm.nb[src] : 0.0;
    };
    /* separator row and Schur complement */
    double As, Ao_s, rs_d, rs_l;
    rc.row(g_sep, As, Ao_s, rs_d, rs_l);
    rs_r = rc.kq(g_sep);                                                 /* element between the separator and the next lane's row 0 */
    if (tid == NT - 1) { As = 1.0; Ao_s = 0.0; rs_r = 0.0; a_last = 0.0; }
    a_first_next = Ao_s;
    const double v0n = from_tid(v[0], 1), w0n = from_tid(w[0], 1);       /* spikes of the next lane at its row 0 */
    double Sdi = As - a_last * w[R - 1] - a_first_next * v0n;
    double Slo = (tid == 0) ? 0.0 : -a_last * v[R - 1];
    double Sup = -a_first_next * w0n;
    if (tid == NT - 1) { Sdi = 1.0; Slo = 0.0; Sup = 0.0; }
    if (tid == NT - 2) Sup = 0.0;                                        /* separator 127 does not exist */
    /* the Dirichlet node (row 0) is an identity row that receives the flux of element 0: its increment must stay 0 */
#pragma unroll
    for (int l = 0; l < R; ++l) if (g_first + l == 0) dinv[l] = 0.0;
    if (TAB) {
        /* 16 rows per lane: spikes and flux coefficients go to per-lane columns of shared memory (conflict-free: consecutive
         * lanes, consecutive addresses); from here on the register copies are dead */
#pragma unroll
        for (int l = 0; l < R; ++l) { PdebD2 p2; p2.x = v[l]; p2.y = w[l]; sm.vw[TAB ? l : 0][tid] = p2; }
#pragma unroll
        for (int l = 0; l <= R; ++l) sm.kqs[TAB ? l : 0][tid] = kq[l];
    }
    auto KQ = [&](int l) -> double { if (TAB) return sm.kqs[TAB ? l : 0][tid]; return kq[l]; };
    const bool sep_is_dirichlet = (g_sep == 0);
    /* level 1 / level 2 split: lane 31 of every warp is a level-2 unknown (an identity row inside the warp's system) */
    const double Slo0 = Slo, Sup0 = Sup, Sdi0 = Sdi;
    if (lane == 0) Slo = 0.0;
    if (lane == 30) Sup = 0.0;
    if (lane == 31) { Sdi = 1.0; Slo = 0.0; Sup = 0.0; }
    double pcr_a[5], pcr_g[5], dinvS;
#pragma unroll
    for (int lev = 0; lev < 5; ++lev) {
        const int sft = 1 << lev;
        const double di_m = __shfl_up_sync(PDEB_FULL, Sdi, sft), di_p = __shfl_down_sync(PDEB_FULL, Sdi, sft);
        const double lo_m = __shfl_up_sync(PDEB_FULL, Slo, sft), lo_p = __shfl_down_sync(PDEB_FULL, Slo, sft);
        const double up_m = __shfl_up_sync(PDEB_FULL, Sup, sft), up_p = __shfl_down_sync(PDEB_FULL, Sup, sft);
        const bool hm = lane - sft >= 0, hp = lane + sft < 32;
        const double al = hm ? Slo / di_m : 0.0, ga = hp ? Sup / di_p : 0.0;
        pcr_a[lev] = al;
        pcr_g[lev] = ga;
        Sdi = Sdi - (hm ? al * up_m : 0.0) - (hp ? ga * lo_p : 0.0);
        Slo = hm ? -al * lo_m : 0.0;
        Sup = hp ? -ga * up_p : 0.0;
    }
    if (!(Sdi > 0.0)) ok = false;
    dinvS = 1.0 / Sdi;
    if (!isfinite(dinvS)) ok = false;
    if (sep_is_dirichlet && lane != 31) dinvS = 0.0;
    auto pcr_apply = [&](double sh) {
#pragma unroll
        for (int lev = 0; lev < 5; ++lev) {
            const int sft = 1 << lev;
            const double sm_ = __shfl_up_sync(PDEB_FULL, sh, sft), sp = __shfl_down_sync(PDEB_FULL, sh, sft);
            sh = fma(-pcr_a[lev], sm_, fma(-pcr_g[lev], sp, sh));       /* coefficients are 0 where no neighbour exists */
        }
        return sh * dinvS;
    };
    const double P = pcr_apply(lane == 0 ? Slo0 : 0.0);      /* response of the warp's separators to B_left */
    const double Q = pcr_apply(lane == 30 ? Sup0 : 0.0);     /* ... and to B_right */
    /* level-2 matrix and its inverse */
    __syncthreads();
    if (lane == 0) { sm.set[warp][0] = P; sm.set[warp][1] = Q; }
    if (lane == 30) { sm.set[warp][2] = P; sm.set[warp][3] = Q; }
    if (lane == 31) { sm.set[warp][4] = Slo0; sm.set[warp][5] = Sdi0; sm.set[warp][6] = Sup0; sm.set[warp][7] = a_first_next; sm.rsr[warp] = rs_r; }
    {
        const bool bad = !ok;
        if (tid == 0) sm.flag = 0;
        __syncthreads();
        if (bad) atomicOr(&sm.flag, 1);
    }
    __syncthreads();
    if (tid == 0) {
        double m2[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, inv[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int j = 0; j < NB; ++j) {
            const double sl = sm.set[j][4], sd = sm.set[j][5], su = sm.set[j][6];
            /* sigma(lane 30 of warp j) = tau - B_{j-1} P_j[30] - B_j Q_j[30];  sigma(lane 0 of warp j+1) = tau - B_j P_{j+1}[0] - B_{j+1} Q_{j+1}[0] */
            m2[j][0] = m2[j][1] = m2[j][2] = 0.0;
            if (j > 0) m2[j][j - 1] = -sl * sm.set[j][2];
            m2[j][j] = sd - sl * sm.set[j][3] - su * sm.set[j + 1][0];
            if (j + 1 < NB) m2[j][j + 1] = -su * sm.set[j + 1][1];
            sm.afn[j] = sm.set[j][7];
            sm.supB[j] = su;
        }
        for (int j = NB; j < 3; ++j) { sm.afn[j] = 0.0; sm.supB[j] = 0.0; }
        /* inverse of the (at most 3 x 3, tridiagonal, diagonally dominant) level-2 matrix by Gauss-Jordan without pivoting */
        bool bad = false;
        for (int c = 0; c < NB; ++c) {
            const double piv = m2[c][c];
            if (!(piv != 0.0) || !isfinite(piv)) { bad = true; break; }
            const double ip = 1.0 / piv;
            for (int q = 0; q < 3; ++q) { m2[c][q] *= ip; inv[c][q] *= ip; }
            for (int r2 = 0; r2 < NB; ++r2) {
                if (r2 == c) continue;
                const double f = m2[r2][c];
                for (int q = 0; q < 3; ++q) { m2[r2][q] -= f * m2[c][q]; inv[r2][q] -= f * inv[c][q]; }
            }
        }
        for (int r2 = 0; r2 < 3; ++r2) for (int q = 0; q < 3; ++q) { sm.G[r2][q] = (r2 < NB && q < NB) ? inv[r2][q] : 0.0; if (!isfinite(sm.G[r2][q])) bad = true; }
        if (bad) sm.flag = 1;
    }
    __syncthreads();
    if (sm.flag && status == PDEB_OK) { status = PDEB_ST_NONFINITE; S = 0; np = 0; }
    /* level-2 constants of this warp in registers: rows of G for the separator in front of the warp (gp) and its own (go),
     * and, for lane 0, the couplings of the PREVIOUS warp's level-2 row to this lane (a_first_next, Sup, flux coefficient) */
    double gp[3] = {0.0, 0.0, 0.0}, go[3] = {0.0, 0.0, 0.0};
    /* (a level-2 separator that is the Dirichlet node keeps its increment at 0: its row of G is taken as zero) */
    const int dirichlet_b0 = (shift % (R + 1) == 0 && (shift / (R + 1) - 1) % 32 == 31) ? (shift / (R + 1) - 1) / 32 : -1;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        if (warp >= 1 && warp - 1 < NB && warp - 1 != dirichlet_b0) gp[j] = sm.G[warp - 1][j];
        if (warp < NB && warp != dirichlet_b0) go[j] = sm.G[warp][j];
    }
    const double pv_afn = (warp >= 1) ? sm.afn[warp - 1] : 0.0, pv_supB = (warp >= 1) ? sm.supB[warp - 1] : 0.0;
    const double pv_rsr = (warp >= 1) ? sm.rsr[warp - 1] : 0.0;

    PdebPwl<double> in;
    in.init(kt, ky, (status == PDEB_OK) ? nk : 0, st, T);
    double f_prev = in.eval(0.0, kt, ky, st);

    /* ---- clause data: thread j sets clause j up in shared memory ---- */
    if (tid < PDEB_MAX_PRED) {
        sm.c_klo[tid] = 0; sm.c_khi[tid] = -1; sm.c_ilo[tid] = 1; sm.c_ihi[tid] = 0; sm.c_info[tid] = 0;
        sm.c_mh[tid] = 0.0; sm.c_c0[tid] = 0.0; sm.c_acc[tid] = 0.0;
    }
    __syncthreads();
    if (tid < np) {
        const int j = tid;
        const double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
        const double xlo = b.pred_xlo[(int64_t)j * st + s], xhi = b.pred_xhi[(int64_t)j * st + s];
        const double tlo = b.pred_tlo[(int64_t)j * st + s], thi = b.pred_thi[(int64_t)j * st + s];
        const double qlo = tlo / dt, qhi = thi / dt;
        double a0, a1;
        if (cv.window_rule == 1) { a0 = trunc(qlo); a1 = trunc(qhi); }
        else { a0 = ceil(qlo - PDEB_EPS); a1 = floor(qhi + PDEB_EPS); }
        if (!(a0 >= 0.0)) a0 = 0.0;
        if (!(a1 <= (double)S)) a1 = (double)S;
        if (a0 > a1) {
            double k = floor(0.5 * (tlo + thi) / dt + 0.5);
            if (!(k >= 0.0)) k = 0.0;
            if (k > (double)S) k = (double)S;
            a0 = a1 = k;
        }
        double ilo = ceil(xlo / h - PDEB_EPS), ihi = floor(xhi / h + PDEB_EPS);
        if (!(ilo >= 0.0)) ilo = 0.0;
        if (!(ihi <= (double)N)) ihi = (double)N;
        const uint32_t top = b.pred_top[(int64_t)j * st + s], rel = b.pred_rel[(int64_t)j * st + s];
        const bool none = !(ilo <= ihi);
        const uint32_t endp = (none || cv.region_rule == 1) ? 1u : 0u;
        sm.c_mh[j] = m * h;
        sm.c_c0[j] = bb - g0;
        sm.c_acc[j] = (top == PDEB_TOP_G) ? INFINITY : -INFINITY;     /* end points (G) / whole clause (F) */
        sm.c_klo[j] = (int)a0; sm.c_khi[j] = (int)a1;
        sm.c_ilo[j] = none ? 1 : (int)ilo; sm.c_ihi[j] = none ? 0 : (int)ihi;
        sm.c_info[j] = rel | (top << 2) | (endp << 3);
    }
#pragma unroll
    for (int j = 0; j < PDEB_MAX_PRED; ++j) sm.gacc[j][tid] = -INFINITY;
    __syncthreads();
    int32_t S32 = 0;                              /* simulate only up to the last step any clause observes */
    for (int j = 0; j < np; ++j) S32 = sm.c_khi[j] > S32 ? sm.c_khi[j] : S32;
    const double eq_eps = (cv.eq_rule == 1) ? cv.eq_eps : 0.0;

#pragma unroll
    for (int l = 0; l < R; ++l) { x[l] = 0.0; y[l] = 0.0; }
    double xs = 0.0;                              /* own separator */
    double xsp = 0.0;                             /* ghost of the neighbouring warp's separator in front of lane 0 */
    bool bad_state = false;                       /* sticky: state not finite at an observed step */
    uint32_t open = 0;
    int32_t next_ev = 0;
    const double th_dt = cv.theta * dt;
    const double rh = 1.0 / h;

    for (int32_t k = 0;; ++k) {
        const int buf = k & 1;
        /* ---- window events (block-uniform) ---- */
        if (k == next_ev) {
            open = 0;
            next_ev = 0x7fffffff;
            for (int j = 0; j < np; ++j) {
                const int32_t a0 = sm.c_klo[j], a1 = sm.c_khi[j];
                if (k >= a0 && k <= a1) { open |= 1u << j; if (a1 + 1 < next_ev) next_ev = a1 + 1; }
                else if (k < a0 && a0 < next_ev) next_ev = a0;
            }
        }
        /* ---- (1) clauses whose window contains step k ---- */
        uint32_t fopen = 0;                       /* open F clauses: their block maximum is formed behind the step's barrier */
        if (open) {
            bad_state = bad_state || !(xs - xs == 0.0) || !(x[0] - x[0] == 0.0) || !(x[R - 1] - x[R - 1] == 0.0);
            uint32_t todo = open;
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1u;
                const double mh = sm.c_mh[j], c0 = sm.c_c0[j];
                const int ilo = sm.c_ilo[j], ihi = sm.c_ihi[j];
                const uint32_t info = sm.c_info[j];
                const uint32_t rel = info & 3u;
                const uint32_t and_mask = (rel == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
                const uint32_t xor_mask = (rel == PDEB_REL_GT) ? 0x80000000u : 0u;
                /* node g sits at local index l = g - g_first; nodes ilo..ihi */
                const int l_lo = ilo - g_first, l_hi = ihi - g_first;
                const double off = fma(mh, (double)g_first, c0);
                double hi4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
                const int b_lo = l_lo < 0 ? 0 : l_lo, b_hi = l_hi > R - 1 ? R - 1 : l_hi;
                const uint32_t rowmask = (b_lo <= b_hi) ? (((1u << (b_hi + 1)) - 1u) & ~((1u << b_lo) - 1u)) : 0u;
                PdebChunkMargin<0, R>::run(x, mh, off, and_mask, xor_mask, rowmask, hi4);
                {
                    const double d = fma(-mh, (double)R, xs - off);
                    const double g = pdeb_signop(d, and_mask, xor_mask);
                    const bool tk = (tid < NT - 1) & (R >= l_lo) & (R <= l_hi) & (g > hi4[R & 3]);
                    hi4[R & 3] = tk ? g : hi4[R & 3];
                }
                const double a0 = hi4[0] > hi4[1] ? hi4[0] : hi4[1], a1 = hi4[2] > hi4[3] ? hi4[2] : hi4[3];
                double gmax = a0 > a1 ? a0 : a1;
                if (((info >> 2) & 1u) == PDEB_TOP_G) {
                    const double prev = sm.gacc[j][tid];
                    if (gmax > prev) sm.gacc[j][tid] = gmax;
                } else {
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const double other = __shfl_xor_sync(PDEB_FULL, gmax, o);
                        gmax = other > gmax ? other : gmax;
                    }
                    if (lane == 0) sm.fx[buf][warp][j] = gmax;
                    fopen |= 1u << j;
                }
                if (info & 8u) {
                    /* uncommon: interpolated end points; the two nodes of an element are published by their owner lanes */
                    const double xe[2] = { b.pred_xlo[(int64_t)j * st + s], b.pred_xhi[(int64_t)j * st + s] };
                    const double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
                    double xi[2], xx2[2];
                    __syncthreads();
                    for (int q = 0; q < 2; ++q) {
                        const double xx = xe[q] < 0 ? 0 : (xe[q] > L ? L : xe[q]);
                        int el = (int)(xx * rh);
                        if (el < 0) el = 0;
                        if (el > N - 1) el = N - 1;
                        xi[q] = (xx - el * h) * rh;
                        xx2[q] = xx;
                        for (int d2 = 0; d2 < 2; ++d2) {
                            const int p = el + d2 + shift;                 /* position of node el + d2 */
                            const int ow = (p - 1) / (R + 1), lo = (p - 1) % (R + 1);
                            if (ow == tid) {
                                double mine = xs;
#pragma unroll
                                for (int l = 0; l < R; ++l) mine = (lo == l) ? x[l] : mine;
                                sm.scr[2 * q + d2] = mine;
                            }
                        }
                    }
                    __syncthreads();
                    double r = INFINITY;
                    for (int q = 0; q < 2; ++q) {
                        const double u0 = sm.scr[2 * q], u1 = sm.scr[2 * q + 1];
                        const double u = g0 + (u0 + xi[q] * (u1 - u0));
                        const double pr = m * xx2[q] + bb;
                        double mg;
                        if (rel == PDEB_REL_GT) mg = u - pr;
                        else if (rel == PDEB_REL_LT) mg = pr - u;
                        else { const double lo2 = u - (pr - eq_eps), hi2 = (pr + eq_eps) - u; mg = lo2 < hi2 ? lo2 : hi2; }
                        if (mg < r || mg != mg) r = mg;
                    }
                    if (tid == 0) {
                        /* G: running minimum of the end-point margins; F: kept for the combination behind the barrier */
                        if (((info >> 2) & 1u) == PDEB_TOP_G) { if (r < sm.c_acc[j] || r != r) sm.c_acc[j] = r; }
                        else sm.scr[8 + j] = r;
                    }
                }
            }
        }
        if (k >= S32) {
            /* last observed step: the F clauses of this step still need their block maximum */
            __syncthreads();
            if (fopen && tid == 0) {
                uint32_t todo = fopen;
                while (todo) {
                    const int j = __ffs(todo) - 1;
                    todo &= todo - 1u;
                    const uint32_t info = sm.c_info[j];
                    double gm = sm.fx[buf][0][j];
                    for (int q = 1; q < WARPS; ++q) gm = sm.fx[buf][q][j] > gm ? sm.fx[buf][q][j] : gm;
                    double r = ((info & 3u) == PDEB_REL_EQ) ? eq_eps - gm : -gm;
                    if (sm.c_ilo[j] > sm.c_ihi[j]) r = INFINITY;
                    if (info & 8u) { const double re = sm.scr[8 + j]; if (re < r || re != re) r = re; }
                    if (r > sm.c_acc[j] || r != r) sm.c_acc[j] = r;
                }
            }
            break;
        }
        /* ---- (2) advance to step k+1 ---- */
        const double f_next = in.eval((double)(k + 1) * dt, kt, ky, st);
        const double fN = fma(dt - th_dt, f_prev, th_dt * f_next);
        f_prev = f_next;
        {
            const double xs_up = __shfl_up_sync(PDEB_FULL, xs, 1), x0_dn = __shfl_down_sync(PDEB_FULL, x[0], 1);
            const double xs_prev = (lane == 0) ? xsp : xs_up;             /* warp 0: the ghost stays 0 */
            /* right-hand side in ELEMENT-FLUX form (see pdeb_chunk.cuh): Q_e = kq_e (x_{e+1} - x_e), row i gets Q_i - Q_{i-1};
             * formed straight into y, which the solve then overwrites with the increment */
            {
                double q_prev = KQ(0) * (x[0] - xs_prev);
#pragma unroll
                for (int l = 0; l < R - 1; ++l) {
                    const double q = KQ(l + 1) * (x[l + 1] - x[l]);
                    y[l] = q - q_prev;
                    q_prev = q;
                }
                const double q_sep = KQ(R) * (xs - x[R - 1]);
                y[R - 1] = q_sep - q_prev + ((tid == NT - 1) ? fN : 0.0);
                /* separator row: rs_r (x_nf - xs) - q_sep.  For the level-2 separator (lane 31) x_nf belongs to the next warp:
                 * that warp's lane 0 carries the term rs_r x_nf in what it publishes (c2 below), so no ghost copy is needed */
                const double rsep = (lane == 31) ? fma(-rs_r, xs, -q_sep) : fma(rs_r, x0_dn - xs, -q_sep);
                const double x0_own = x[0];
                local_solve(y);
                const double y0_dn = __shfl_down_sync(PDEB_FULL, y[0], 1);
                /* level-1 Schur right-hand side (lanes 0..30) and the warp's PCR: tau */
                const double sh = (lane == 31) ? 0.0 : fma(-a_last, y[R - 1], fma(-a_first_next, y0_dn, rsep));
                const double tau = pcr_apply(sh);
                const double tau30 = __shfl_sync(PDEB_FULL, tau, 30);
                /* what the level-2 right-hand side needs from this warp: lane 31 its own row's piece, lane 0 the terms of the
                 * PREVIOUS warp's level-2 row that live here (constants of that row were exchanged at set-up) */
                if (lane == 31) sm.ex[buf][warp][2] = fma(-a_last, y[R - 1], rsep) - Slo0 * tau30;
                if (lane == 0) sm.ex[buf][warp][0] = fma(pv_afn, y[0], fma(pv_supB, tau, -pv_rsr * x0_own));
                __syncthreads();                                                              /* THE barrier of the step */
                /* level 2: beta_j = piece_j - c2_{j+1};  B = G beta (G rows of this warp's two neighbours are in registers) */
                double beta[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int j = 0; j < NB; ++j) beta[j] = sm.ex[buf][j][2] - sm.ex[buf][j + 1][0];
                double Bp = 0.0, Bo = 0.0;                 /* B of the separator in front of this warp, and of its own */
#pragma unroll
                for (int j = 0; j < NB; ++j) { Bp = fma(gp[j], beta[j], Bp); Bo = fma(go[j], beta[j], Bo); }
                const double xs_new = (lane == 31) ? Bo : fma(-Bo, Q, fma(-Bp, P, tau));
                const double xn_up = __shfl_up_sync(PDEB_FULL, xs_new, 1);
                const double xs_new_prev = (lane == 0) ? Bp : xn_up;
                xsp += Bp;                                 /* ghost of the separator in front of lane 0 (used by lane 0 only) */
#pragma unroll
                for (int l = 0; l < R; ++l) {
                    double vl, wl;
                    if (TAB) { const PdebD2 p2 = sm.vw[TAB ? l : 0][tid]; vl = p2.x; wl = p2.y; } else { vl = v[l]; wl = w[l]; }
                    x[l] += fma(-vl, xs_new_prev, fma(-wl, xs_new, y[l]));                       /* the solve returned the increment */
                }
                xs += xs_new;
            }
        }
        /* ---- F clauses of step k: block maximum from the slots published before the barrier ---- */
        if (fopen && tid == 0) {
            uint32_t todo = fopen;
            while (todo) {
                const int j = __ffs(todo) - 1;
                todo &= todo - 1u;
                const uint32_t info = sm.c_info[j];
                double gm = sm.fx[buf][0][j];
                for (int q = 1; q < WARPS; ++q) gm = sm.fx[buf][q][j] > gm ? sm.fx[buf][q][j] : gm;
                double r = ((info & 3u) == PDEB_REL_EQ) ? eq_eps - gm : -gm;
                if (sm.c_ilo[j] > sm.c_ihi[j]) r = INFINITY;
                if (info & 8u) { const double re = sm.scr[8 + j]; if (re < r || re != re) r = re; }
                if (r > sm.c_acc[j] || r != r) sm.c_acc[j] = r;
            }
        }
    }

    /* ---- G clauses: block maximum of the per-lane running maxima; clause values; formula tree (thread 0) ---- */
    __syncthreads();
    for (int j = 0; j < np; ++j) {
        const uint32_t info = sm.c_info[j];
        if (((info >> 2) & 1u) != PDEB_TOP_G) continue;
        double gm = sm.gacc[j][tid];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const double other = __shfl_xor_sync(PDEB_FULL, gm, o);
            gm = other > gm ? other : gm;
        }
        if (lane == 0) sm.fx[0][warp][j] = gm;
    }
    if (bad_state) atomicOr(&sm.flag, 2);
    __syncthreads();
    if (tid == 0) {
        for (int j = 0; j < np; ++j) {
            const uint32_t info = sm.c_info[j];
            if (((info >> 2) & 1u) == PDEB_TOP_G) {
                double gm = sm.fx[0][0][j];
                for (int q = 1; q < WARPS; ++q) gm = sm.fx[0][q][j] > gm ? sm.fx[0][q][j] : gm;
                double r = ((info & 3u) == PDEB_REL_EQ) ? eq_eps - gm : -gm;
                if (sm.c_ilo[j] > sm.c_ihi[j]) r = INFINITY;
                const double re = sm.c_acc[j];                 /* end-point margins (INFINITY when there are none) */
                sm.cval[j] = (re < r || re != re) ? re : r;
            } else sm.cval[j] = sm.c_acc[j];
        }
        double val[PDEB_MAX_TREE];
        double rob = PDEB_FAILED_ROBUSTNESS;
        if (status == PDEB_OK) {
            for (int j = 0; j < nt; ++j) {
                const int op = b.tree_op[(int64_t)j * st + s], lh = b.tree_lhs[(int64_t)j * st + s], rhh = b.tree_rhs[(int64_t)j * st + s];
                double xv = NAN;
                if (op == PDEB_TREE_LEAF) {
                    if (lh >= 0 && lh < np) xv = sm.cval[lh]; else status = PDEB_ST_BADPARAM;
                } else if ((op == PDEB_TREE_AND || op == PDEB_TREE_OR) && lh >= 0 && lh < j && rhh >= 0 && rhh < j) {
                    const double a0 = val[lh], a1 = val[rhh];
                    if (a0 != a0 || a1 != a1) xv = NAN;
                    else if (op == PDEB_TREE_AND) xv = a0 < a1 ? a0 : a1;
                    else xv = a0 > a1 ? a0 : a1;
                } else status = PDEB_ST_BADPARAM;
                val[j] = xv;
            }
            if (status == PDEB_OK) {
                rob = val[nt - 1];
                if (!isfinite(rob) || (sm.flag & 2)) status = PDEB_ST_NONFINITE;
                for (int j = 0; j < np; ++j) if (!isfinite(sm.cval[j])) status = PDEB_ST_NONFINITE;
            }
        }
        out.robustness[so] = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
        if (out.status) out.status[so] = status;
        if (out.clause)
            for (int j = 0; j < b.pmax; ++j)
                out.clause[(int64_t)j * st + so] = (status == PDEB_OK && j < np) ? sm.cval[j] : PDEB_FAILED_ROBUSTNESS;
    }
}

#endif /* PDEB_CTA_CUH */
