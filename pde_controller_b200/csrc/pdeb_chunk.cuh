/*
 * pdeb_chunk.cuh — "chunk family": one WARP owns one trajectory of arbitrary mesh size
 * (up to 32 R + 30 elements; R = rows per lane, a template parameter).  This is the kernel
 * behind the refined-mesh configuration (1,024 elements x 10,000 steps) and behind every mesh
 * the register family is not instantiated for.
 *
 * Same mathematics as pdeb_core.h (heat: generalised trapezoidal d-form on the deviation
 * w = d - g0; wave: Newmark a-form; STL clauses reduced on the fly), restating what femformal's
 * heatlinfem_mix / mechlinfem systems and APCont / cspec robustness do per sample behind
 * eval_robustness_wrapper.py:16 of the reference.
 *
 * Layout.  The N free rows are END-ALIGNED onto 32 R + 31 positions: position p = row + shift,
 * shift = 32 R + 31 - N >= 1.  Lane j owns the R consecutive "interior" positions
 * j (R+1) + 1 .. j (R+1) + R and, for j < 31, the "separator" position (j+1)(R+1) behind them.
 * Positions <= shift are padding (identity rows, value 0); position shift is the Dirichlet node
 * (deviation 0), and the loaded free end (row N) is always lane 31's last interior row.
 *
 * Per time step the constant SPD tridiagonal step matrix is solved by substructuring:
 *   1. every lane solves its interior block with its own prefactored L D L^T      (serial, R rows)
 *   2. the 31 separator unknowns satisfy a tridiagonal Schur system, solved across the lanes by
 *      parallel cyclic reduction with PRECOMPUTED per-level coefficients       (5 shuffle rounds)
 *   3. the interior values are corrected with the two precomputed "spike" vectors (2 FMA per row)
 * Interior blocks of lanes that lie in the same material are identical, so their factor tables
 * are stored once per CLASS of lane in shared memory (<= PDEB_CHUNK_NCLS classes per trajectory)
 * and read as broadcasts.  The state stays in registers; nothing per step touches HBM.
 */
#ifndef PDEB_CHUNK_CUH
#define PDEB_CHUNK_CUH

#include <cuda_runtime.h>

#include "pdeb_core.h"

#define PDEB_CHUNK_NCLS 8      /* lane classes per trajectory: outer, inner, padding, <= 5 special */
#define PDEB_CHUNK_NVEC 6      /* table values per row and class, as three 16-byte pairs: {e, dinv} {v, w} {rd, ror} */
#define PDEB_CHUNK_WARPS 4     /* warps (trajectories) per block */
#ifndef PDEB_CHUNK_MINBLOCKS
#define PDEB_CHUNK_MINBLOCKS 1  /* resident blocks per SM the kernels are compiled for */
#endif
#define PDEB_FULL 0xffffffffu

__device__ __forceinline__ double pdeb_shfl(double x, int src) { return __shfl_sync(PDEB_FULL, x, src); }
__device__ __forceinline__ double pdeb_shfl_up1(double x, int lane) { double y = __shfl_up_sync(PDEB_FULL, x, 1); return lane == 0 ? 0.0 : y; }
__device__ __forceinline__ double pdeb_shfl_dn1(double x, int lane) { double y = __shfl_down_sync(PDEB_FULL, x, 1); return lane == 31 ? 0.0 : y; }

/* per-row coefficients of the step matrix A and of the right-hand-side operator R, as functions of the row */
struct PdebRowCoef {
    double ad_o, ad_i, ao_o, ao_i, rd_o, rd_i, ro_o, ro_i;   /* per element, o = outer material, i = inner */
    double kq_o, kq_i;       /* flux coefficient of an element: heat dt k, wave k (right-hand side in element-flux form) */
    double h, x_a, x_b;
    int N, sampling;

    __device__ __forceinline__ bool outer(int e) const {
        const double xs = (sampling == 1) ? e * h : (sampling == 2) ? (e + 1) * h : (e + 0.5) * h;
        return xs < x_a || xs > x_b;
    }
    /* element e joins nodes e and e+1, e = 0..N-1 (node 0 is the Dirichlet node, deviation 0); none outside */
    __device__ __forceinline__ double kq(int e) const {
        if (e < 0 || e > N - 1) return 0.0;
        return outer(e) ? kq_o : kq_i;
    }
    /* row g in 1..N: diagonal and coupling to row g+1; rows outside are identity / empty */
    __device__ __forceinline__ void row(int g, double& Ad, double& Ao, double& Rd, double& Ro) const {
        if (g < 1 || g > N) { Ad = 1.0; Ao = 0.0; Rd = 0.0; Ro = 0.0; return; }
        const bool ol = outer(g - 1);
        Ad = ol ? ad_o : ad_i;
        Rd = ol ? rd_o : rd_i;
        Ao = 0.0; Ro = 0.0;
        if (g < N) {
            const bool orr = outer(g);
            Ad += orr ? ad_o : ad_i;
            Rd += orr ? rd_o : rd_i;
            Ao = orr ? ao_o : ao_i;
            Ro = orr ? ro_o : ro_i;
        }
    }
};

/* masked max of the relation-transformed margins of a lane's R interior rows (PTX form of pdeb_core.h) */
template <int L, int R>
struct PdebChunkMargin {
    static __device__ __forceinline__ void run(const double* x, double mh, double off, uint32_t and_mask, uint32_t xor_mask,
                                               uint32_t rowmask, double* hi4) {
        const double d = fma(-mh, (double)L, x[L] - off);
        pdeb_masked_max<L>(hi4[L & 3], pdeb_signop(d, and_mask, xor_mask), rowmask);
        PdebChunkMargin<L + 1, R>::run(x, mh, off, and_mask, xor_mask, rowmask, hi4);
    }
};
template <int R>
struct PdebChunkMargin<R, R> {
    static __device__ __forceinline__ void run(const double*, double, double, uint32_t, uint32_t, uint32_t, double*) {}
};

template <int R>
struct PdebChunkGeom {
    static constexpr int CAP = 32 * R + 31;                   /* positions */
    static constexpr int CLS_STRIDE = PDEB_CHUNK_NVEC * R + 2;  /* reals; the +2 staggers the banks of the classes, keeps 16-byte alignment */
    static constexpr int M = R / 2;                           /* twist row of the interior block */
    static constexpr int TAB_REALS = PDEB_CHUNK_NCLS * CLS_STRIDE;
    static constexpr int SCR_REALS = 2 * (CAP + 1);               /* two position-indexed scratch vectors (set-up) */
    static constexpr int WARP_REALS = TAB_REALS > SCR_REALS ? TAB_REALS : SCR_REALS;
};

template <int KIND, int R>
__global__ void __launch_bounds__(32 * PDEB_CHUNK_WARPS, PDEB_CHUNK_MINBLOCKS)
pdeb_chunk_kernel(const __grid_constant__ PdebDevBatch b, const __grid_constant__ PdebConventions cv,
                  const __grid_constant__ PdebDevOut out, const int32_t* __restrict__ order, int64_t first, int64_t count,
                  int gather) {
    typedef PdebChunkGeom<R> G;
    extern __shared__ __align__(16) unsigned char pdeb_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * PDEB_CHUNK_WARPS + warp;
    if (t >= count) return;                                   /* whole warp */
    const int64_t so = order[first + t];                      /* the sample's index in the caller's order: outputs */
    const int64_t s = gather ? so : first + t;                /* inputs: the image as uploaded, or the plan-ordered image */
    const int64_t st = b.stride;
    double* tab = reinterpret_cast<double*>(pdeb_smem) + (size_t)warp * G::WARP_REALS;

    /* ---- parameters (warp-uniform) ---- */
    const int N = b.n_elem[s];
    const double L = b.L[s], dt = b.dt[s], T = b.T[s], g0 = b.g0[s];
    const double x_a = b.x_a[s], x_b = b.x_b[s];
    const double rho_in = b.rho_in[s], rho_out = b.rho_out[s], E_in = b.E_in[s], E_out = b.E_out[s];
    int np = b.n_pred[s];
    const int nk = b.n_knots[s], nt = b.n_tree[s];
    int status = PDEB_OK;
    int64_t S = 0;
    if (!pdeb_params_ok(KIND, N, L, dt, T, g0, rho_in, rho_out, E_in, E_out) || N > G::CAP - 1) status = PDEB_ST_BADPARAM;
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
    const int shift = G::CAP - N;                 /* position = row + shift */
    const int base = lane * (R + 1);              /* interior positions base+1 .. base+R, separator base+R+1 */
    const double h = L / N;

    /* ---- assembly, factorisation, Schur complement, PCR coefficients ---- */
    double x[R], y[R];                            /* state (deviation) and work vector of the interior rows */
    double vv[KIND == PDEB_KIND_WAVE ? R : 1], aa[KIND == PDEB_KIND_WAVE ? R : 1];
    double xs = 0.0, vs = 0.0, as = 0.0;          /* separator state */
    double a_last = 0.0, a_first_next = 0.0;      /* A couplings of the own separator to both sides */
    double rs_d = 0.0, rs_l = 0.0, rs_r = 0.0;    /* rs_r: flux coefficient of the element right of the separator (rs_d, rs_l: set-up scratch) */
    double rol0 = 0.0;                            /* flux coefficient of the element between the previous separator and row 0 */
    double pcr_a[5], pcr_g[5], dinvS = 1.0;
    int cls = 2;
    const int lumped = (KIND == PDEB_KIND_HEAT) ? cv.mass_heat : cv.mass_wave;
    PdebRowCoef rc;
    {
        const double cA = (KIND == PDEB_KIND_HEAT) ? cv.theta * dt : cv.newmark_beta * dt * dt;
        const double cR = (1.0 - cv.theta) * dt;
        const double m_o = rho_out * h, k_o = E_out / h, m_i = rho_in * h, k_i = E_in / h;
        const double mo_o = lumped ? 0.0 : m_o / 6.0, mo_i = lumped ? 0.0 : m_i / 6.0;
        const double md_o = lumped ? 0.5 * m_o : m_o / 3.0, md_i = lumped ? 0.5 * m_i : m_i / 3.0;
        rc.ad_o = md_o + cA * k_o; rc.ad_i = md_i + cA * k_i;
        rc.ao_o = mo_o - cA * k_o; rc.ao_i = mo_i - cA * k_i;
        rc.rd_o = (KIND == PDEB_KIND_HEAT) ? md_o - cR * k_o : -k_o;
        rc.rd_i = (KIND == PDEB_KIND_HEAT) ? md_i - cR * k_i : -k_i;
        rc.ro_o = (KIND == PDEB_KIND_HEAT) ? mo_o + cR * k_o : k_o;
        rc.ro_i = (KIND == PDEB_KIND_HEAT) ? mo_i + cR * k_i : k_i;
        rc.kq_o = (KIND == PDEB_KIND_HEAT) ? dt * k_o : k_o;
        rc.kq_i = (KIND == PDEB_KIND_HEAT) ? dt * k_i : k_i;
        rc.h = h; rc.x_a = x_a; rc.x_b = x_b; rc.N = (status == PDEB_OK) ? N : 0; rc.sampling = cv.material_sampling;
    }
    if (status != PDEB_OK) { S = 0; np = 0; }
    {
        /* class of this lane: rows g0r .. g0r+R-1 with g0r = base + 1 - shift */
        const int g_first = base + 1 - shift, g_last = g_first + R - 1;
        bool special = false;
        if (g_last < 0) cls = 2;                                           /* all padding (a lane that holds the Dirichlet row needs the flux coefficient of element 0: special) */
        else if (g_first >= 2 && g_last <= rc.N - 1) {
            /* elements g_first-1 .. g_last all exist: uniform if they share one material */
            const bool o0 = rc.outer(g_first - 1);
            bool uni = true;
            for (int e = g_first; e <= g_last; ++e) uni = uni && (rc.outer(e) == o0);
            if (uni) cls = o0 ? 0 : 1; else special = true;
        } else special = true;
        const unsigned sp = __ballot_sync(PDEB_FULL, special);
        if (special) cls = 3 + __popc(sp & ((1u << lane) - 1u));
        if (__any_sync(PDEB_FULL, cls >= PDEB_CHUNK_NCLS)) { status = PDEB_ST_BADPARAM; S = 0; np = 0; cls = 2; rc.N = 0; }
    }
    /* ---- initial state (uses the shared scratch before the class tables are built) ---- */
#pragma unroll
    for (int l = 0; l < R; ++l) { x[l] = 0.0; y[l] = 0.0; }
    if (KIND == PDEB_KIND_WAVE) {
#pragma unroll
        for (int l = 0; l < R; ++l) { vv[l] = 0.0; aa[l] = 0.0; }
        if (status == PDEB_OK) {
            /* d = 0 on the free nodes (deviation w0 = -g0), v = 0, a0 = M^-1 (f0 e_N - K w0) */
            const int g_first = base + 1 - shift;
            const double wdev = -g0;
            const double m_o = rho_out * h, m_i = rho_in * h;
            double f0;
            { PdebPwl<double> p0; p0.init(kt, ky, nk, st, T); f0 = p0.eval(0.0, kt, ky, st); }
            double* scr_r = tab;                       /* rhs / solution by position */
            double* scr_c = tab + (G::CAP + 1);        /* Thomas multipliers by position */
            /* every lane: r_g = (R w0)_g (+ f0 at row N) for its R interior rows and its separator */
            for (int l = 0; l <= R; ++l) {
                if (l == R && lane == 31) break;
                const int g = g_first + l;
                double r = 0.0;
                if (g >= 1 && g <= N) {
                    double Ad, Ao, Rd, Ro, Adm, Aom, Rdm, Rom;
                    rc.row(g, Ad, Ao, Rd, Ro);
                    rc.row(g - 1, Adm, Aom, Rdm, Rom);
                    r = wdev * (Rd + ((g - 1 >= 1) ? Rom : 0.0) + ((g + 1 <= N) ? Ro : 0.0));
                    if (g == N) r += f0;
                    const double ml = rc.outer(g - 1) ? m_o : m_i, mr = (g < N) ? (rc.outer(g) ? m_o : m_i) : 0.0;
                    if (lumped) r = r / (0.5 * (ml + mr));
                }
                scr_r[base + 1 + l] = r;
            }
            __syncwarp();
            if (!lumped && lane == 0) {
                /* consistent mass: exact serial Thomas solve with M over rows 1..N (set-up only) */
                double cp_prev = 0.0, x_prev = 0.0, mo_prev = 0.0;
                for (int g = 1; g <= N; ++g) {
                    const double ml = rc.outer(g - 1) ? m_o : m_i, mr = (g < N) ? (rc.outer(g) ? m_o : m_i) : 0.0;
                    const double md = (ml + mr) / 3.0, mo = mr / 6.0;
                    const double den = md - mo_prev * cp_prev;
                    cp_prev = mo / den;
                    x_prev = (scr_r[g + shift] - mo_prev * x_prev) / den;
                    scr_c[g + shift] = cp_prev;
                    scr_r[g + shift] = x_prev;
                    mo_prev = mo;
                }
                for (int g = N - 1; g >= 1; --g) scr_r[g + shift] -= scr_c[g + shift] * scr_r[g + 1 + shift];
            }
            __syncwarp();
#pragma unroll
            for (int l = 0; l < R; ++l) {
                const int g = g_first + l;
                x[l] = (g >= 1 && g <= N) ? wdev : 0.0;
                aa[l] = scr_r[base + 1 + l];
            }
            if (lane < 31) {
                const int g = g_first + R;
                xs = (g >= 1 && g <= N) ? wdev : 0.0;
                as = scr_r[base + 1 + R];
            }
            __syncwarp();
        }
    }
    /* this lane's class table: three arrays of 16-byte pairs indexed by the local row */
    PdebD2* t_ed = reinterpret_cast<PdebD2*>(tab + cls * G::CLS_STRIDE);      /* { e_l : multiplier of edge (l, l+1), 1/pivot_l } */
    PdebD2* t_vw = t_ed + R;                                                   /* { v_l, w_l } spikes */
    PdebD2* t_rr = t_ed + 2 * R;                                               /* { R_ll, R_l,l+1 } */
    constexpr int M = G::M;
    /*
     * In-place TWISTED solve of the interior block with the factors in t_ed, run with run-time loops over shared
     * memory (set-up only).  sel = 0 works on the v slots, 1 on the w slots; on entry they hold the right-hand side.
     */
    auto smem_solve = [&](int sel) {
        double* yv = reinterpret_cast<double*>(t_vw) + sel;          /* element l at yv[2 l] */
        double z = 0.0, zt = 0.0, zb = 0.0;
        for (int l = 0; l < M; ++l) {                                 /* top chain */
            z = yv[2 * l] - ((l > 0) ? t_ed[l - 1].x * zt : 0.0);
            zt = z;
            yv[2 * l] = z * t_ed[l].y;
        }
        for (int l = R - 1; l > M; --l) {                             /* bottom chain */
            z = yv[2 * l] - ((l < R - 1) ? t_ed[l].x * zb : 0.0);
            zb = z;
            yv[2 * l] = z * t_ed[l].y;
        }
        z = yv[2 * M] - ((M > 0) ? t_ed[M - 1].x * zt : 0.0) - ((M < R - 1) ? t_ed[M].x * zb : 0.0);
        yv[2 * M] = z * t_ed[M].y;
        for (int l = M - 1; l >= 0; --l) yv[2 * l] = yv[2 * l] - t_ed[l].x * yv[2 * (l + 1)];
        for (int l = M + 1; l < R; ++l) yv[2 * l] = yv[2 * l] - t_ed[l - 1].x * yv[2 * (l - 1)];
    };
    {
        const int g_first = base + 1 - shift;
        /* pass 1: twisted L D L^T of the interior block (top sweep, bottom sweep, twist row M), R-operator rows */
        bool ok = true;
        double dprev = 0.0, ao_prev = 0.0, a_first = 0.0;
        for (int l = 0; l < R; ++l) {
            double Ad, Ao, Rd, Ro;
            rc.row(g_first + l, Ad, Ao, Rd, Ro);
            PdebD2 rr; rr.x = 0.0; rr.y = rc.kq(g_first + l);            /* flux coefficient of the element right of row l */
            t_rr[l] = rr;
            if (l == R - 1) a_last = Ao;
        }
        for (int l = 0; l < M; ++l) {                                 /* top sweep */
            double Ad, Ao, Rd, Ro;
            rc.row(g_first + l, Ad, Ao, Rd, Ro);
            double piv = Ad;
            if (l > 0) piv -= (ao_prev * dprev) * ao_prev;
            if (!(piv > 0.0)) ok = false;
            dprev = 1.0 / piv;
            PdebD2 ed; ed.x = Ao * dprev; ed.y = dprev;               /* edge (l, l+1) eliminated downwards */
            t_ed[l] = ed;
            ao_prev = Ao;
        }
        const double dtop = dprev, aotop = ao_prev;
        dprev = 0.0;
        double ao_this = 0.0;
        for (int l = R - 1; l > M; --l) {                             /* bottom sweep */
            double Ad, Ao, Rd, Ro;
            rc.row(g_first + l, Ad, Ao, Rd, Ro);
            double piv = Ad;
            if (l < R - 1) piv -= (Ao * dprev) * Ao;                  /* edge (l, l+1) eliminated upwards */
            if (!(piv > 0.0)) ok = false;
            const double e_up = (l < R - 1) ? Ao * dprev : 0.0;
            dprev = 1.0 / piv;
            PdebD2 ed; ed.x = e_up; ed.y = dprev;
            t_ed[l] = ed;
            ao_this = Ao;
        }
        (void)ao_this;
        {
            double Ad, Ao, Rd, Ro;
            rc.row(g_first + M, Ad, Ao, Rd, Ro);                      /* twist row; Ao couples M with M+1 */
            const double e_dn = (M > 0) ? aotop * dtop : 0.0;         /* already stored in t_ed[M-1].x */
            const double e_up = (M < R - 1) ? Ao * dprev : 0.0;
            const double piv = Ad - ((M > 0) ? e_dn * aotop : 0.0) - ((M < R - 1) ? e_up * Ao : 0.0);
            if (!(piv > 0.0)) ok = false;
            const double dm = 1.0 / piv;
            if (!isfinite(dm) || !isfinite(dtop) || !isfinite(dprev)) ok = false;
            PdebD2 ed; ed.x = e_up; ed.y = dm;
            t_ed[M] = ed;
        }
        /* couplings to the previous separator (row g_first - 1) */
        {
            double Adp, Rdp, rfirst;
            rc.row(g_first - 1, Adp, a_first, Rdp, rfirst);
            if (lane == 0) a_first = 0.0;
            (void)rfirst;
            rol0 = rc.kq(g_first - 1);                                   /* element between the previous separator and row 0 */
        }
        __syncwarp();
        /* pass 2: spikes v = A_loc^-1 (a_first e_0), w = A_loc^-1 (a_last e_{R-1}) */
        for (int l = 0; l < R; ++l) {
            PdebD2 vw; vw.x = (l == 0) ? a_first : 0.0; vw.y = (l == R - 1) ? a_last : 0.0;
            t_vw[l] = vw;
        }
        __syncwarp();
        smem_solve(0);
        __syncwarp();
        smem_solve(1);
        if (__any_sync(PDEB_FULL, !ok) && status == PDEB_OK) { status = PDEB_ST_NONFINITE; S = 0; np = 0; }
        __syncwarp();
        /* separator row and Schur complement */
        double As, Ao_s;
        rc.row(base + R + 1 - shift, As, Ao_s, rs_d, rs_l);
        rs_r = rc.kq(base + R + 1 - shift);                             /* element between the separator and the next lane's row 0 */
        if (lane == 31) { As = 1.0; Ao_s = 0.0; rs_r = 0.0; a_last = 0.0; }
        a_first_next = Ao_s;
        const double v_last = t_vw[R - 1].x, w_last = t_vw[R - 1].y;
        const double v0 = t_vw[0].x, w0 = t_vw[0].y;
        const double v0_n = pdeb_shfl_dn1(v0, lane), w0_n = pdeb_shfl_dn1(w0, lane);
        double Sdi = As - a_last * w_last - a_first_next * v0_n;
        double Slo = (lane == 0) ? 0.0 : -a_last * v_last;
        double Sup = -a_first_next * w0_n;
        if (lane == 31) { Sdi = 1.0; Slo = 0.0; Sup = 0.0; }
        if (lane == 30) Sup = 0.0;                                        /* separator 31 does not exist */
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
        if (__any_sync(PDEB_FULL, !ok) && status == PDEB_OK) { status = PDEB_ST_NONFINITE; S = 0; np = 0; }
        /*
         * The Dirichlet node (row 0, an identity row that is decoupled from row 1) receives the flux of element 0 on
         * its right-hand side in the element-flux form; its increment must stay 0, so its pivot inverse is zeroed
         * (its lane is a special class: the table is its own).
         */
        __syncwarp();
        if (g_first <= 0 && g_first + R - 1 >= 0) t_ed[-g_first].y = 0.0;
        if (base + R + 1 - shift == 0) dinvS = 0.0;
        __syncwarp();
    }

    PdebPwl<double> in;
    in.init(kt, ky, (status == PDEB_OK) ? nk : 0, st, T);
    double f_prev = in.eval(0.0, kt, ky, st);

    /* clause data: lane j keeps clause j */
    double c_mh = 0.0, c_c0 = 0.0, c_acc = 0.0;
    int c_klo = 0, c_khi = -1, c_ilo = 0, c_ihi = -1;
    uint32_t c_info = 0;
    if (lane < np) {
        const int j = lane;
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
        c_mh = m * h;
        c_c0 = bb - g0;
        c_acc = (top == PDEB_TOP_G) ? INFINITY : -INFINITY;
        c_klo = (int)a0; c_khi = (int)a1;
        c_ilo = none ? 1 : (int)ilo; c_ihi = none ? 0 : (int)ihi;
        c_info = rel | (top << 2) | (endp << 3);
    }

    /* simulate only up to the last step any clause observes */
    int32_t S32 = (lane < np) ? c_khi : 0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { const int32_t other = __shfl_xor_sync(PDEB_FULL, S32, o); S32 = other > S32 ? other : S32; }
    const double eq_eps = (cv.eq_rule == 1) ? cv.eq_eps : 0.0;

    bool bad_state = false;     /* sticky: state not finite at an observed step (the maxima below drop NaN margins) */
    for (int32_t k = 0;; ++k) {
        /* ---- (1) clauses whose window contains step k ---- */
        unsigned open = __ballot_sync(PDEB_FULL, lane < np && k >= c_klo && k <= c_khi);
        if (open) bad_state = bad_state || !(xs - xs == 0.0) || !(x[0] - x[0] == 0.0) || !(x[R - 1] - x[R - 1] == 0.0);
        while (open) {
            const int j = __ffs(open) - 1;
            open &= open - 1u;
            const double mh = pdeb_shfl(c_mh, j), c0 = pdeb_shfl(c_c0, j);
            const int ilo = __shfl_sync(PDEB_FULL, c_ilo, j), ihi = __shfl_sync(PDEB_FULL, c_ihi, j);
            const uint32_t info = __shfl_sync(PDEB_FULL, c_info, j);
            const uint32_t rel = info & 3u;
            const uint32_t and_mask = (rel == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
            const uint32_t xor_mask = (rel == PDEB_REL_GT) ? 0x80000000u : 0u;
            /* node g sits at local index l = g + shift - base - 1; nodes ilo..ihi */
            const int l_lo = ilo + shift - base - 1, l_hi = ihi + shift - base - 1;
            const double off = fma(mh, (double)(base + 1 - shift), c0);
            double hi4[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
            /* bit l: local row l lies inside the clause interval (R <= 32) */
            const int b_lo = l_lo < 0 ? 0 : l_lo, b_hi = l_hi > R - 1 ? R - 1 : l_hi;
            const uint32_t rowmask = (b_lo <= b_hi) ? ((b_hi >= 31 ? 0xffffffffu : ((1u << (b_hi + 1)) - 1u)) & ~((1u << b_lo) - 1u)) : 0u;
            PdebChunkMargin<0, R>::run(x, mh, off, and_mask, xor_mask, rowmask, hi4);
            {
                const double d = fma(-mh, (double)R, xs - off);
                const double g = pdeb_signop(d, and_mask, xor_mask);
                const bool tk = (lane < 31) & (R >= l_lo) & (R <= l_hi) & (g > hi4[R & 3]);
                hi4[R & 3] = tk ? g : hi4[R & 3];
            }
            double a0 = hi4[0] > hi4[1] ? hi4[0] : hi4[1], a1 = hi4[2] > hi4[3] ? hi4[2] : hi4[3];
            double gmax = a0 > a1 ? a0 : a1;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const double other = __shfl_xor_sync(PDEB_FULL, gmax, o);
                gmax = other > gmax ? other : gmax;
            }
            double r = (rel == PDEB_REL_EQ) ? eq_eps - gmax : -gmax;
            if (info & 8u) {
                /* uncommon: interpolated end points; the two nodes of an element are fetched from their owner lanes */
                if (ilo > ihi) r = INFINITY;
                const double xe[2] = { b.pred_xlo[(int64_t)j * st + s], b.pred_xhi[(int64_t)j * st + s] };
                const double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
                const double rh = 1.0 / h;
                for (int q = 0; q < 2; ++q) {
                    const double xx = xe[q] < 0 ? 0 : (xe[q] > L ? L : xe[q]);
                    int e = (int)(xx * rh);
                    if (e < 0) e = 0;
                    if (e > N - 1) e = N - 1;
                    const double xi = (xx - e * h) * rh;
                    double u2[2];
                    for (int d2 = 0; d2 < 2; ++d2) {
                        const int p = e + d2 + shift;                 /* position of node e + d2 */
                        const int ow = (p - 1) / (R + 1), lo = (p - 1) % (R + 1);
                        double mine = xs;
#pragma unroll
                        for (int l = 0; l < R; ++l) mine = (lo == l) ? x[l] : mine;
                        u2[d2] = pdeb_shfl(mine, ow);
                    }
                    const double u = g0 + (u2[0] + xi * (u2[1] - u2[0]));
                    const double pr = m * xx + bb;
                    double mg;
                    if (rel == PDEB_REL_GT) mg = u - pr;
                    else if (rel == PDEB_REL_LT) mg = pr - u;
                    else { const double lo2 = u - (pr - eq_eps), hi2 = (pr + eq_eps) - u; mg = lo2 < hi2 ? lo2 : hi2; }
                    if (mg < r) r = mg;
                }
            }
            if (lane == j) {
                const bool take = (((info >> 2) & 1u) == PDEB_TOP_G) ? (r < c_acc) : (r > c_acc);
                if (take || r != r) c_acc = r;
            }
        }
        if (k >= S32) break;
        /* ---- (2) advance to step k+1 ---- */
        const double f_next = in.eval((double)(k + 1) * dt, kt, ky, st);
        double fN;
        if (KIND == PDEB_KIND_HEAT) {
            const double th_dt = cv.theta * dt;
            fN = fma(dt - th_dt, f_prev, th_dt * f_next);
        } else {
            const double cw1 = dt, cw2 = 0.5 * (1.0 - 2.0 * cv.newmark_beta) * dt * dt, cw3 = (1.0 - cv.newmark_gamma) * dt;
#pragma unroll
            for (int l = 0; l < R; ++l) {
                x[l] = fma(cw2, aa[l], fma(cw1, vv[l], x[l]));
                vv[l] = fma(cw3, aa[l], vv[l]);
            }
            xs = fma(cw2, as, fma(cw1, vs, xs));
            vs = fma(cw3, as, vs);
            fN = f_next;
        }
        f_prev = f_next;
        {
            const double xs_prev = pdeb_shfl_up1(xs, lane), x_nf = pdeb_shfl_dn1(x[0], lane);
            /*
             * Right-hand side in ELEMENT-FLUX form: Q_e = kq_e (x_{e+1} - x_e), row i gets Q_i - Q_{i-1} (+ load at the
             * free end).  Heat solves for the INCREMENT  A (x' - x) = dt (f - K x), waves for  A a' = f' - K d~.
             * The assembled form R x = R_ii x_i + R_i,i-1 x_{i-1} + R_i,i+1 x_{i+1} costs the same three operations
             * but cancels: on refined meshes |R_ii x| is ~ (k dt / m) times the result, and 10,000 steps carried that
             * rounding to 1.6e-9 of the parity scale; differences of neighbouring values do not cancel.
             */
            const double q_sep = t_rr[R - 1].y * (xs - x[R - 1]);              /* element between row R-1 and the separator */
            const double rsep = rs_r * (x_nf - xs) - q_sep;
            /*
             * Twisted solve with the right-hand side formed row by row inside the two elimination chains (top rows
             * ascending, bottom rows descending, meeting at row M).
             */
            {
                double zt = 0.0, zb = 0.0, e_top = 0.0;
                double q_top = rol0 * (x[0] - xs_prev);       /* flux of the element left of the top chain's row */
                double q_bot = q_sep;                         /* flux of the element right of the bottom chain's row */
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const int it = j, ib = R - 1 - j;
                    if (it < M) {
                        const PdebD2 ed = t_ed[it];
                        const double q = t_rr[it].y * (x[it + 1] - x[it]);
                        const double r = q - q_top;
                        const double z = (it > 0) ? fma(-e_top, zt, r) : r;
                        zt = z;
                        y[it] = z * ed.y;
                        e_top = ed.x;
                        q_top = q;
                    }
                    if (ib > M) {
                        const PdebD2 ed = t_ed[ib];
                        const double q = t_rr[ib - 1].y * (x[ib] - x[ib - 1]);
                        double r = q_bot - q;
                        if (ib == R - 1) r += (lane == 31) ? fN : 0.0;
                        const double z = (ib < R - 1) ? fma(-ed.x, zb, r) : r;
                        zb = z;
                        y[ib] = z * ed.y;
                        q_bot = q;
                    }
                }
                {
                    const PdebD2 ed = t_ed[M];
                    double z = q_bot - q_top;
                    if (M == R - 1) z += (lane == 31) ? fN : 0.0;
                    if (M > 0) z = fma(-e_top, zt, z);
                    if (M < R - 1) z = fma(-ed.x, zb, z);
                    y[M] = z * ed.y;
                }
#pragma unroll
                for (int l = M - 1; l >= 0; --l) y[l] = fma(-t_ed[l].x, y[l + 1], y[l]);
#pragma unroll
                for (int l = M + 1; l < R; ++l) y[l] = fma(-t_ed[l - 1].x, y[l - 1], y[l]);
            }
            /* Schur right-hand side and PCR across the lanes */
            const double y_nf = pdeb_shfl_dn1(y[0], lane);
            double sh = (lane == 31) ? 0.0 : fma(-a_last, y[R - 1], fma(-a_first_next, y_nf, rsep));
#pragma unroll
            for (int lev = 0; lev < 5; ++lev) {
                const int sft = 1 << lev;
                const double sm = __shfl_up_sync(PDEB_FULL, sh, sft), sp = __shfl_down_sync(PDEB_FULL, sh, sft);
                sh = fma(-pcr_a[lev], sm, fma(-pcr_g[lev], sp, sh));    /* coefficients are 0 where no neighbour exists */
            }
            const double xs_new = sh * dinvS;
            const double xs_new_prev = pdeb_shfl_up1(xs_new, lane);
            if (KIND == PDEB_KIND_HEAT) {
#pragma unroll
                for (int l = 0; l < R; ++l) {
                    const PdebD2 vw = t_vw[l];
                    x[l] += fma(-vw.x, xs_new_prev, fma(-vw.y, xs_new, y[l]));       /* the solve returned the increment */
                }
                xs += xs_new;
            } else {
                const double cw4 = cv.newmark_beta * dt * dt, cw5 = cv.newmark_gamma * dt;
#pragma unroll
                for (int l = 0; l < R; ++l) {
                    const PdebD2 vw = t_vw[l];
                    const double an = fma(-vw.x, xs_new_prev, fma(-vw.y, xs_new, y[l]));
                    aa[l] = an;
                    x[l] = fma(cw4, an, x[l]);
                    vv[l] = fma(cw5, an, vv[l]);
                }
                as = xs_new;
                xs = fma(cw4, xs_new, xs);
                vs = fma(cw5, xs_new, vs);
            }
        }
    }

    /* ---- formula tree and outputs (lane 0) ---- */
    double val[PDEB_MAX_TREE];
    double rob = PDEB_FAILED_ROBUSTNESS;
    if (status == PDEB_OK) {
        for (int j = 0; j < nt; ++j) {
            const int op = b.tree_op[(int64_t)j * st + s], lh = b.tree_lhs[(int64_t)j * st + s], rh = b.tree_rhs[(int64_t)j * st + s];
            double xv = NAN;
            if (op == PDEB_TREE_LEAF) {
                const double a = pdeb_shfl(c_acc, (lh >= 0 && lh < np) ? lh : 0);
                if (lh >= 0 && lh < np) xv = a; else status = PDEB_ST_BADPARAM;
            } else if ((op == PDEB_TREE_AND || op == PDEB_TREE_OR) && lh >= 0 && lh < j && rh >= 0 && rh < j) {
                const double a0 = val[lh], a1 = val[rh];
                if (a0 != a0 || a1 != a1) xv = NAN;
                else if (op == PDEB_TREE_AND) xv = a0 < a1 ? a0 : a1;
                else xv = a0 > a1 ? a0 : a1;
            } else status = PDEB_ST_BADPARAM;
            val[j] = xv;
        }
        if (status == PDEB_OK) {
            rob = val[nt - 1];
            if (!isfinite(rob)) status = PDEB_ST_NONFINITE;
            if (__any_sync(PDEB_FULL, (lane < np && !isfinite(c_acc)) || bad_state)) status = PDEB_ST_NONFINITE;
        }
    }
    if (lane == 0) {
        out.robustness[so] = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
        if (out.status) out.status[so] = status;
    }
    if (out.clause && lane < b.pmax)
        out.clause[(int64_t)lane * st + so] = (status == PDEB_OK && lane < np) ? c_acc : PDEB_FAILED_ROBUSTNESS;
}

template <int R>
inline size_t pdeb_chunk_smem_bytes() {
    return sizeof(double) * (size_t)PDEB_CHUNK_WARPS * PdebChunkGeom<R>::WARP_REALS;
}

#endif /* PDEB_CHUNK_CUH */
