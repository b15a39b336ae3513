/*
 * pdeb_core.h — per-trajectory device code of the fused simulate + STL-robustness path.
 *
 * One THREAD owns one trajectory ("register family"): the whole state vector and the
 * LDL^T factors of the constant step matrix live in registers (every loop over nodes is
 * fully unrolled with the mesh size as a template parameter), the per-clause data live in
 * a per-thread column of shared memory, and the trajectory itself is never stored: each
 * new state is reduced into the clause accumulators on the spot.
 *
 * What it computes (reference call sites: heatlinfem.heatlinfem_mix / mechlinfem.mechlinfem /
 * logic.APCont / cspec in utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3 and
 * one_d_wave/examples.jsonl:1-3; driven by eval_robustness_wrapper.py:16):
 *   heat  generalised trapezoidal rule in d-form on the deviation w = d - g0:
 *           (M + th dt K) w' = (M - (1-th) dt K) w + dt ((1-th) f + th f') e_N
 *         (algebraically the v-form the oracle states)
 *   wave  Newmark a-form with (beta, gamma):  predictor, (M + beta dt^2 K) a' = f' e_N - K dp, corrector
 *   STL   clause margin = min over nodes of the interval of +-(u - (m x + b)), G = running min,
 *         F = running max over the window steps, formula tree of min / max.
 *
 * The file is also compiled by g++ (PDEB_HOST_EMU) for the CPU-side logic tests in tests/:
 * there a "warp" is one lane.  That build is test scaffolding, not a product path.
 */
#ifndef PDEB_CORE_H
#define PDEB_CORE_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pdeb200.h"

#if defined(__CUDACC__)
#define PDEB_HD __host__ __device__ __forceinline__
#define PDEB_UNROLL _Pragma("unroll")
#else
#define PDEB_HD inline
#define PDEB_UNROLL
#endif

#if defined(__CUDA_ARCH__)
#define PDEB_ANY(x) __any_sync(0xffffffffu, (x))
#else
#define PDEB_ANY(x) (x)
#endif

#define PDEB_EPS 1e-9

/* device view of a batch: the PdebBatch arrays, on the device */
struct PdebDevBatch {
    int64_t n, stride;
    int32_t kmax, pmax, tmax, pad;
    const uint8_t* kind;
    const int32_t* n_elem;
    const double *L, *dt, *T, *g0, *x_a, *x_b, *rho_in, *rho_out, *E_in, *E_out;
    const int32_t* n_knots;
    const double *knot_t, *knot_y;
    const int32_t* n_pred;
    const double *pred_xlo, *pred_xhi, *pred_m, *pred_b, *pred_tlo, *pred_thi;
    const uint8_t *pred_rel, *pred_top;
    const int32_t* n_tree;
    const uint8_t* tree_op;
    const int8_t *tree_lhs, *tree_rhs;
};

struct PdebDevOut {
    double* robustness;   /* [n] */
    double* clause;       /* [pmax][stride] or NULL */
    int32_t* status;      /* [n] */
};

/*
 * Per-thread clause slots in shared memory, laid out [clause][field][thread] in two areas
 * (reals, then 32-bit words) so that a thread needs only two base pointers:
 *   reals: 0 mh = m*h (profile slope per node index), 1 c0 = b - g0, 2 acc (running temporal min / max)
 *   words: 0 klo, 1 khi (window steps), 2 nmask (bit i: node i inside [x_lo,x_hi]),
 *          3 info (bits 0-1 rel, bit 2 top, bit 3 needs the interpolated end points)
 */
#define PDEB_SLOT_REALS 3
#define PDEB_SLOT_WORDS 4
template <typename real>
struct PdebClauseSlots {
    real* rbase;       /* + tid */
    uint32_t* wbase;   /* + tid */
    int stride;        /* threads per block */
    PDEB_HD real& R(int p, int f) const { return rbase[(p * PDEB_SLOT_REALS + f) * stride]; }
    PDEB_HD uint32_t& W(int p, int f) const { return wbase[(p * PDEB_SLOT_WORDS + f) * stride]; }
};

template <typename real>
PDEB_HD size_t pdeb_clause_smem_bytes(int threads, int pmax) {
    return (size_t)threads * pmax * (PDEB_SLOT_REALS * sizeof(real) + PDEB_SLOT_WORDS * sizeof(uint32_t));
}

template <typename real>
PDEB_HD PdebClauseSlots<real> pdeb_clause_slots(void* base, int threads, int pmax, int tid) {
    PdebClauseSlots<real> s;
    real* r = (real*)base;
    s.rbase = r + tid;
    s.wbase = (uint32_t*)(r + (size_t)PDEB_SLOT_REALS * threads * pmax) + tid;
    s.stride = threads;
    return s;
}

PDEB_HD int64_t pdeb_step_count(double T, double dt, int rule) {
    double q = T / dt;
    double s;
    if (rule == 1) s = floor(q + 0.5);
    else if (rule == 2) s = floor(q + PDEB_EPS);
    else s = ceil(q - PDEB_EPS);
    if (!(s >= 1.0)) s = 1.0;
    if (s > 4.0e9) s = 4.0e9;
    return (int64_t)s;
}

/* host-visible validity test of the scalar parameters (same rules as the oracle's unpack) */
PDEB_HD bool pdeb_params_ok(int kind, int N, double L, double dt, double T, double g0, double rho_in, double rho_out,
                            double E_in, double E_out) {
    if (kind != PDEB_KIND_HEAT && kind != PDEB_KIND_WAVE) return false;
    if (N < 1 || N > PDEB_MAX_ELEM) return false;
    if (!(L > 0) || !(dt > 0) || !(T > 0)) return false;
    if (!(rho_in > 0) || !(rho_out > 0) || !(E_in > 0) || !(E_out > 0)) return false;
    if (!isfinite(L) || !isfinite(dt) || !isfinite(T) || !isfinite(g0)) return false;
    return true;
}

/*
 * Running piecewise-linear input (C11): value held before the first and after the last knot.
 * Only the current segment lives in registers: y0 + slope (t - t0) until t reaches t1.  The knot BEHIND the
 * current segment is fetched one segment ahead, so that crossing a knot never waits for memory: lanes of a warp
 * cross their knots at different steps, which makes the warp take this path in most steps.
 */
template <typename real>
struct PdebPwl {
    double t1, tn;        /* end of the current segment (+inf on the last); time of the knot after it (+inf: none) */
    real t0, y0, slope;   /* current segment */
    real y1, yn;          /* values at t1 and tn */
    int nk, j;            /* j: index of the knot at t1 */

    PDEB_HD void init(const double* kt, const double* ky, int nk_, int64_t stride) {
        nk = nk_; j = 0;
        t0 = 0; slope = 0; y0 = 0; y1 = 0; yn = 0; t1 = INFINITY; tn = INFINITY;
        if (nk >= 1) { y0 = y1 = (real)ky[0]; t1 = kt[0]; }
        if (nk >= 2) { tn = kt[stride]; yn = (real)ky[stride]; }
    }
    /* kt, ky already point at this sample's first knot; stride = plane stride */
    PDEB_HD real eval(double t, const double* kt, const double* ky, int64_t stride) {
        while (t >= t1) {
            t0 = (real)t1; y0 = y1;
            if (tn < INFINITY) {
                slope = (real)(((double)yn - (double)y1) / (tn - t1));
                t1 = tn; y1 = yn;
                ++j;
                if (j + 1 < nk) { tn = kt[(int64_t)(j + 1) * stride]; yn = (real)ky[(int64_t)(j + 1) * stride]; }
                else tn = INFINITY;
            } else { slope = 0; t1 = INFINITY; }
        }
        return fma(slope, (real)(t - (double)t0), y0);
    }
};

struct alignas(16) PdebD2 { double x, y; };
struct alignas(8) PdebF2 { float x, y; };
template <typename real> struct PdebVec2;
template <> struct PdebVec2<double> { typedef PdebD2 type; };
template <> struct PdebVec2<float> { typedef PdebF2 type; };

/* shared memory of one block of the register family: [row table | clause slots] */
template <typename real>
PDEB_HD size_t pdeb_table_smem_bytes(int threads, int n_rows) {
    return (size_t)threads * n_rows * sizeof(typename PdebVec2<real>::type);
}

/*
 * Register-family trajectory.  KIND: PDEB_KIND_*; N: elements = free dofs (nodes 1..N), 4 <= N <= 31.
 * Arrays are indexed by node; index 0 is the Dirichlet node (deviation 0).
 *
 * The constant step matrix A (SPD, tridiagonal) is factored ONCE per sample as a TWISTED
 * L D L^T: rows 1..M-1 are eliminated downwards, rows N..M+1 upwards, and the two sweeps meet
 * at row M.  Each time step then runs two independent dependency chains of half the length
 * (instruction-level parallelism for a kernel that is bound by FP64 latency/issue).
 *   registers      : state, the N-1 edge multipliers c[], two off-diagonal values of R
 *   shared (column): per row { diagonal of R, 1/pivot }   -> one 16-byte load per row and step
 */
template <int KIND, int N, typename real>
struct PdebTrajReg {
    static_assert(N >= 4 && N <= 31, "register family handles 4..31 elements");
    static constexpr int M = (N + 1) / 2;   /* twist row */
    typedef typename PdebVec2<real>::type real2;

    real ro[2];        /* off-diagonal of the right-hand-side operator R per element: [0] outer, [1] inner */
    uint32_t omask;    /* bit e: element e is of the outer material */
    real c[N];         /* c[i], i = 1..N-1: elimination multiplier of the edge (i, i+1) */
    real w[N + 2];     /* heat: temperature deviation d - g0; wave: displacement deviation */
    real v[KIND == PDEB_KIND_WAVE ? N + 2 : 1];
    real a[KIND == PDEB_KIND_WAVE ? N + 2 : 1];
    real2* tab;        /* shared: tab[(i-1)*tstride] = { R_ii, 1/pivot_i } */
    int tstride;

    PDEB_HD bool outer(int e) const { return (omask >> e) & 1u; }
    PDEB_HD real sel_ro(int e) const { return outer(e) ? ro[0] : ro[1]; }

    /* (R x)_i with the diagonal entry rd given; x[0] = 0 */
    PDEB_HD real rrow(const real* x, int i, real rd) const {
        real r = rd * x[i];
        if (i > 1) r = fma(sel_ro(i - 1), x[i - 1], r);
        if (i < N) r = fma(sel_ro(i), x[i + 1], r);
        return r;
    }

    /*
     * z[1..N] <- A^-1 (R x + fN e_N), the two chains interleaved row by row.
     * x and z may not alias (x is read at i-1, i, i+1).  Per row: 3 FMA (R x), 1 FMA (eliminate),
     * 1 MUL (1/pivot), 1 FMA (back substitute) = 6 FP64 operations.
     */
    PDEB_HD void step_solve(const real* x, real fN, real* z) const {
        /* right-hand side + forward elimination; z[i] leaves this loop already divided by its pivot */
        real zt = 0, zb = 0;    /* undivided value of the last row of each chain */
        PDEB_UNROLL
        for (int j = 0; j < N; ++j) {
            /* top row it = 1 + j (while it < M), bottom row ib = N - j (while ib > M) */
            const int it = 1 + j, ib = N - j;
            if (it < M) {
                const real2 t = tab[(it - 1) * tstride];
                real r = rrow(x, it, t.x);
                if (it > 1) r = fma(-c[it - 1], zt, r);
                zt = r;
                z[it] = r * t.y;
            }
            if (ib > M) {
                const real2 t = tab[(ib - 1) * tstride];
                real r = rrow(x, ib, t.x);
                if (ib == N) r += fN;
                if (ib < N) r = fma(-c[ib], zb, r);
                zb = r;
                z[ib] = r * t.y;
            }
        }
        {
            const real2 t = tab[(M - 1) * tstride];
            real r = rrow(x, M, t.x);
            r = fma(-c[M - 1], zt, r);
            r = fma(-c[M], zb, r);
            z[M] = r * t.y;
        }
        /* back substitution outwards from the twist row */
        PDEB_UNROLL
        for (int j = 1; j < N; ++j) {
            const int it = M - j, ib = M + j;
            if (it >= 1) z[it] = fma(-c[it], z[it + 1], z[it]);
            if (ib <= N) z[ib] = fma(-c[ib - 1], z[ib - 1], z[ib]);
        }
    }

    /*
     * Set up coefficients and factors.  Returns false on a non-positive pivot.
     * A = M + cA K; R = M - cR K for heat, R = -K for wave.
     */
    PDEB_HD bool setup(double L, double x_a, double x_b, double rho_in, double rho_out, double E_in, double E_out,
                       int lumped, int sampling, double cA, double cR) {
        const double h = L / N;
        omask = 0;
        PDEB_UNROLL
        for (int e = 0; e < N; ++e) {
            double xs = (sampling == 1) ? e * h : (sampling == 2) ? (e + 1) * h : (e + 0.5) * h;
            if (xs < x_a || xs > x_b) omask |= (1u << e);
        }
        /* per-material element coefficients: suffix o = outer material, i = inner */
        const double m_o = rho_out * h, k_o = E_out / h, m_i = rho_in * h, k_i = E_in / h;
        const double mo_o = lumped ? 0.0 : m_o / 6.0, mo_i = lumped ? 0.0 : m_i / 6.0;
        const double md_o = lumped ? 0.5 * m_o : m_o / 3.0, md_i = lumped ? 0.5 * m_i : m_i / 3.0;
        const double ad_o = md_o + cA * k_o, ad_i = md_i + cA * k_i;       /* A = M + cA K, half diagonal */
        const double ao_o = mo_o - cA * k_o, ao_i = mo_i - cA * k_i;       /* ... off diagonal */
        const double rd_o = (KIND == PDEB_KIND_HEAT) ? md_o - cR * k_o : -k_o;
        const double rd_i = (KIND == PDEB_KIND_HEAT) ? md_i - cR * k_i : -k_i;
        ro[0] = (real)((KIND == PDEB_KIND_HEAT) ? mo_o + cR * k_o : k_o);
        ro[1] = (real)((KIND == PDEB_KIND_HEAT) ? mo_i + cR * k_i : k_i);
#define PDEB_EL(e, vo, vi) (outer(e) ? (vo) : (vi))
        /* twisted L D L^T, streamed: only the previous pivot of each sweep is kept live */
        bool ok = true;
        double dtop = 0.0, dbot = 0.0;     /* 1/pivot of the last eliminated row of each sweep */
        PDEB_UNROLL
        for (int i = 1; i < M; ++i) {
            double piv = PDEB_EL(i - 1, ad_o, ad_i) + PDEB_EL(i, ad_o, ad_i);
            if (i > 1) { const double aol = PDEB_EL(i - 1, ao_o, ao_i), ci = aol * dtop; c[i - 1] = (real)ci; piv -= ci * aol; }
            if (!(piv > 0.0)) ok = false;
            dtop = 1.0 / piv;
            real2 t;
            t.x = (real)(PDEB_EL(i - 1, rd_o, rd_i) + PDEB_EL(i, rd_o, rd_i));
            t.y = (real)dtop;
            tab[(i - 1) * tstride] = t;
        }
        PDEB_UNROLL
        for (int i = N; i > M; --i) {
            double piv = PDEB_EL(i - 1, ad_o, ad_i) + ((i < N) ? PDEB_EL(i, ad_o, ad_i) : 0.0);
            if (i < N) { const double aor = PDEB_EL(i, ao_o, ao_i), ci = aor * dbot; c[i] = (real)ci; piv -= ci * aor; }
            if (!(piv > 0.0)) ok = false;
            dbot = 1.0 / piv;
            real2 t;
            t.x = (real)(PDEB_EL(i - 1, rd_o, rd_i) + ((i < N) ? PDEB_EL(i, rd_o, rd_i) : 0.0));
            t.y = (real)dbot;
            tab[(i - 1) * tstride] = t;
        }
        {
            const double aol = PDEB_EL(M - 1, ao_o, ao_i), aor = PDEB_EL(M, ao_o, ao_i);
            const double ct = aol * dtop, cb = aor * dbot;
            c[M - 1] = (real)ct;
            c[M] = (real)cb;
            const double piv = PDEB_EL(M - 1, ad_o, ad_i) + PDEB_EL(M, ad_o, ad_i) - ct * aol - cb * aor;
            if (!(piv > 0.0)) ok = false;
            const double dm = 1.0 / piv;
            if (!isfinite(dm) || !isfinite(dtop) || !isfinite(dbot)) ok = false;
            real2 t;
            t.x = (real)(PDEB_EL(M - 1, rd_o, rd_i) + PDEB_EL(M, rd_o, rd_i));
            t.y = (real)dm;
            tab[(M - 1) * tstride] = t;
        }
#undef PDEB_EL
        return ok;
    }
};

/* sign surgery on an IEEE value: hi word <- (hi & and_mask) ^ xor_mask  (identity, negate or abs) */
PDEB_HD double pdeb_signop(double x, uint32_t and_mask, uint32_t xor_mask) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)((((uint32_t)__double2hiint(x)) & and_mask) ^ xor_mask), __double2loint(x));
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    u = (u & 0xffffffffull) | ((uint64_t)((((uint32_t)(u >> 32)) & and_mask) ^ xor_mask) << 32);
    memcpy(&x, &u, 8);
    return x;
#endif
}
PDEB_HD float pdeb_signop(float x, uint32_t and_mask, uint32_t xor_mask) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float((__float_as_uint(x) & and_mask) ^ xor_mask);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u & and_mask) ^ xor_mask;
    memcpy(&x, &u, 4);
    return x;
#endif
}

/*
 * Spatial margin of one clause on the current state u = g0 + w.
 * With d_i = u_i - p(x_i) = w_i - mh*i - c0, every relation is ONE masked max-reduction:
 *     ">"  margin = min_i  d_i   = -max_i (-d_i)
 *     "<"  margin = min_i -d_i   = -max_i ( d_i)
 *     "="  margin = -max_i |d_i| (+ eq_eps for the band rule)
 * so the per-lane relation only chooses how the sign bit of d_i is treated (one LOP3).
 * Four independent accumulators keep the compare/select chain short.
 */
template <int N, typename real>
PDEB_HD real pdeb_clause_margin(const real* w, real mh, real c0, uint32_t nmask, uint32_t rel, real eq_eps) {
    const uint32_t and_mask = (rel == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
    const uint32_t xor_mask = (rel == PDEB_REL_GT) ? 0x80000000u : 0u;
    real hi[4];
    for (int q = 0; q < 4; ++q) hi[q] = (real)(-INFINITY);
    PDEB_UNROLL
    for (int i = 0; i <= N; ++i) {
        const real d = ((i == 0) ? (real)0 : fma(-mh, (real)i, w[i])) - c0;
        const real g = pdeb_signop(d, and_mask, xor_mask);
        const bool t = ((nmask & (1u << i)) != 0u) & (g > hi[i & 3]);
        hi[i & 3] = t ? g : hi[i & 3];
    }
    const real a0 = hi[0] > hi[1] ? hi[0] : hi[1], a1 = hi[2] > hi[3] ? hi[2] : hi[3];
    const real gmax = a0 > a1 ? a0 : a1;
    return (rel == PDEB_REL_EQ) ? eq_eps - gmax : -gmax;     /* eq_eps = 0 unless eq_rule == 1 */
}

/* interpolated end points of a clause interval (C8): uncommon, so everything is re-read from the batch */
template <int N, typename real>
PDEB_HD real pdeb_endpoint_margin(const PdebDevBatch& b, int64_t s, int j, const real* w, uint32_t rel, double eq_eps) {
    const int64_t st = b.stride;
    const double L = b.L[s], g0 = b.g0[s];
    const double h = L / N, rh = 1.0 / h;
    const double xe[2] = { b.pred_xlo[(int64_t)j * st + s], b.pred_xhi[(int64_t)j * st + s] };
    const double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
    /* the state is copied to a local array so that it can be indexed by the element number */
    real wl[N + 2];
    PDEB_UNROLL
    for (int i = 0; i <= N; ++i) wl[i] = (i == 0) ? (real)0 : w[i];
    double r = INFINITY;
    for (int q = 0; q < 2; ++q) {
        const double xx = xe[q] < 0 ? 0 : (xe[q] > L ? L : xe[q]);
        int e = (int)(xx * rh);
        if (e < 0) e = 0;
        if (e > N - 1) e = N - 1;
        const double xi = (xx - e * h) * rh;
        const double ue = (double)wl[e], uf = (double)wl[e + 1];
        const double u = g0 + (ue + xi * (uf - ue));
        const double pr = m * xx + bb;
        double mg;
        if (rel == PDEB_REL_GT) mg = u - pr;
        else if (rel == PDEB_REL_LT) mg = pr - u;
        else { const double lo = u - (pr - eq_eps), hi = (pr + eq_eps) - u; mg = lo < hi ? lo : hi; }
        if (mg < r) r = mg;
    }
    return (real)r;
}

/*
 * Whole trajectory of sample s.  `slots` is this thread's clause column in shared memory,
 * `table` its row table.  Writes robustness / clause / status.
 *
 * Time loop, per step k:  (1) clauses whose window is open are reduced on the state w^k (a warp
 * runs as many reduction passes as its busiest lane has open clauses, usually one),
 * (2) w^{k+1} is solved for.
 * Lanes whose trajectory is over keep stepping harmlessly until the warp is done (the batch is
 * ordered by step count, so that is a handful of steps): their observations are switched off.
 */
template <int KIND, int N, typename real>
PDEB_HD void pdeb_run_reg(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, int64_t s,
                          typename PdebVec2<real>::type* table, int table_stride, PdebClauseSlots<real> slots, bool live) {
    /* `live` is false for padding lanes of a partially filled warp: they run the loop with S = 0 */
    const int64_t st = b.stride;
    int status = PDEB_OK;
    double dt = 1;
    int np = 0;
    int32_t S32 = 0;          /* steps actually simulated: up to the last observed one */
    PdebTrajReg<KIND, N, real> tr;
    tr.tab = table;
    tr.tstride = table_stride;
    PdebPwl<real> in;
    in.init(nullptr, nullptr, 0, 0);
    const double* kt = b.knot_t + (live ? s : 0);
    const double* ky = b.knot_y + (live ? s : 0);
    PDEB_UNROLL
    for (int i = 0; i <= N + 1; ++i) tr.w[i] = 0;
    if (KIND == PDEB_KIND_WAVE) {
        PDEB_UNROLL
        for (int i = 0; i <= N + 1; ++i) { tr.v[i] = 0; tr.a[i] = 0; }
    }

    if (live) {
        const double L = b.L[s], T = b.T[s], g0 = b.g0[s];
        dt = b.dt[s];
        const double x_a = b.x_a[s], x_b = b.x_b[s];
        const double rho_in = b.rho_in[s], rho_out = b.rho_out[s], E_in = b.E_in[s], E_out = b.E_out[s];
        np = b.n_pred[s];
        const int nk = b.n_knots[s];
        const int nt = b.n_tree[s];
        int64_t S = 0;
        if (!pdeb_params_ok(KIND, N, L, dt, T, g0, rho_in, rho_out, E_in, E_out)) status = PDEB_ST_BADPARAM;
        if (np < 1 || np > b.pmax || nk < 0 || nk > b.kmax || nt < 1 || nt > b.tmax) status = PDEB_ST_BADPARAM;
        if (status == PDEB_OK) {
            for (int j = 1; j < nk; ++j)
                if (!(kt[(int64_t)j * st] > kt[(int64_t)(j - 1) * st])) status = PDEB_ST_BADPARAM;
            for (int j = 0; j < np; ++j)
                if (b.pred_rel[(int64_t)j * st + s] > 2 || b.pred_top[(int64_t)j * st + s] > 1) status = PDEB_ST_BADPARAM;
        }
        if (status == PDEB_OK) {
            S = pdeb_step_count(T, dt, cv.steps_rule);
            if (S > PDEB_MAX_STEPS) status = PDEB_ST_BADPARAM;
        }
        const double h = L / N;
        if (status == PDEB_OK) {
            const int lumped = (KIND == PDEB_KIND_HEAT) ? cv.mass_heat : cv.mass_wave;
            const double cA = (KIND == PDEB_KIND_HEAT) ? cv.theta * dt : cv.newmark_beta * dt * dt;
            const double cR = (1.0 - cv.theta) * dt;
            const bool ok = tr.setup(L, x_a, x_b, rho_in, rho_out, E_in, E_out, lumped, cv.material_sampling, cA, cR);
            if (!ok) status = PDEB_ST_NONFINITE;
            in.init(kt, ky, nk, st);
            if (KIND == PDEB_KIND_WAVE) {
                /* wave initial state: d = 0 on the free nodes (deviation -g0), v = 0, a0 = M^-1 (f0 e_N - K w0) */
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) tr.w[i] = (real)(-g0);
                const real f0 = in.eval(0.0, kt, ky, st);
                const double m_o = rho_out * h, m_i = rho_in * h;
                if (lumped) {
                    PDEB_UNROLL
                    for (int i = 1; i <= N; ++i) {
                        real r = tr.rrow(tr.w, i, tr.tab[(i - 1) * tr.tstride].x);
                        if (i == N) r += f0;
                        const double md = 0.5 * ((tr.outer(i - 1) ? m_o : m_i) + ((i < N) ? (tr.outer(i) ? m_o : m_i) : 0.0));
                        tr.a[i] = (real)((double)r / md);
                    }
                } else {
                    /* consistent mass: one Thomas solve with M (set-up only) */
                    double cp[N + 1], x[N + 1];
                    double mo_prev = 0.0;
                    PDEB_UNROLL
                    for (int i = 1; i <= N; ++i) {
                        real r = tr.rrow(tr.w, i, tr.tab[(i - 1) * tr.tstride].x);
                        if (i == N) r += f0;
                        const double ml = tr.outer(i - 1) ? m_o : m_i, mr = (i < N) ? (tr.outer(i) ? m_o : m_i) : 0.0;
                        const double md = (ml + mr) / 3.0, mo_i = mr / 6.0;
                        const double den = md - ((i > 1) ? mo_prev * cp[i - 1] : 0.0);
                        cp[i] = mo_i / den;
                        x[i] = ((double)r - ((i > 1) ? mo_prev * x[i - 1] : 0.0)) / den;
                        mo_prev = mo_i;
                    }
                    PDEB_UNROLL
                    for (int i = N - 1; i >= 1; --i) x[i] -= cp[i] * x[i + 1];
                    PDEB_UNROLL
                    for (int i = 1; i <= N; ++i) tr.a[i] = (real)x[i];
                }
            }
        }
        if (status != PDEB_OK) { S = 0; np = 0; }
        int32_t k_end = 0;         /* last step any clause observes: nothing after it can change the result */
        /* clause slots */
        for (int j = 0; j < np; ++j) {
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
            uint32_t nm = 0;
            if (ilo <= ihi) {
                const int i0 = (int)ilo, i1 = (int)ihi;
                nm = (i1 >= 31 ? 0xffffffffu : ((1u << (i1 + 1)) - 1u)) & ~((1u << i0) - 1u);
            }
            const uint32_t top = b.pred_top[(int64_t)j * st + s], rel = b.pred_rel[(int64_t)j * st + s];
            const uint32_t endp = (nm == 0 || cv.region_rule == 1) ? 1u : 0u;
            slots.R(j, 0) = (real)(m * h);
            slots.R(j, 1) = (real)(bb - g0);
            slots.R(j, 2) = (top == PDEB_TOP_G) ? (real)INFINITY : (real)(-INFINITY);
            slots.W(j, 0) = (uint32_t)(int32_t)a0;
            slots.W(j, 1) = (uint32_t)(int32_t)a1;
            slots.W(j, 2) = nm;
            slots.W(j, 3) = rel | (top << 2) | (endp << 3);
            if ((int32_t)a1 > k_end) k_end = (int32_t)a1;
        }
        S32 = k_end;
    }

    real f_prev = in.eval(0.0, kt, ky, st);
    int32_t next_ev = 0;      /* next step at which the set of open windows may change */
    uint32_t open = 0;        /* bit j: step k is inside clause j's window */

    for (int32_t k = 0;; ++k) {
        const bool on = (k <= S32);
        if (on && k == next_ev) {
            open = 0;
            next_ev = 0x7fffffff;
            for (int j = 0; j < np; ++j) {
                const int32_t a0 = (int32_t)slots.W(j, 0), a1 = (int32_t)slots.W(j, 1);
                if (k >= a0 && k <= a1) { open |= (1u << j); if (a1 + 1 < next_ev) next_ev = a1 + 1; }
                else if (k < a0 && a0 < next_ev) next_ev = a0;
            }
        }
        /* ---- (1) reduce the open clauses on the state of step k ---- */
        uint32_t todo = on ? open : 0u;
        while (PDEB_ANY(todo != 0u)) {
            const bool has = todo != 0u;
            int j = 0;
            if (has) { j = __builtin_ffs((int)todo) - 1; todo &= todo - 1u; }
            real mh = 0, c0 = 0, acc = 0;
            uint32_t nm = 0, info = 0;
            if (has) {
                mh = slots.R(j, 0);
                c0 = slots.R(j, 1);
                acc = slots.R(j, 2);
                nm = slots.W(j, 2);
                info = slots.W(j, 3);
            }
            const uint32_t rel = info & 3u;
            const real eq_eps = (cv.eq_rule == 1) ? (real)cv.eq_eps : (real)0;
            real r = pdeb_clause_margin<N, real>(tr.w, mh, c0, nm, rel, eq_eps);
            if (PDEB_ANY(has && (info & 8u))) {
                if (has && (info & 8u)) {
                    const real re = pdeb_endpoint_margin<N, real>(b, s, j, tr.w, rel, (double)eq_eps);
                    if (nm == 0u || re < r) r = re;
                }
            }
            if (has) {
                const bool take = (((info >> 2) & 1u) == PDEB_TOP_G) ? (r < acc) : (r > acc);
                if (take || r != r) slots.R(j, 2) = r;
            }
        }
        if (!PDEB_ANY(k < S32)) break;
        /* ---- (2) advance to step k+1 (lanes that are done keep stepping; nobody looks at them) ---- */
        {
            const real f_next = in.eval((double)(k + 1) * dt, kt, ky, st);
            if (KIND == PDEB_KIND_HEAT) {
                const real th_dt = (real)(cv.theta * dt);
                const real fb = fma((real)dt - th_dt, f_prev, th_dt * f_next);
                real z[N + 2];
                z[0] = 0; z[N + 1] = 0;
                tr.step_solve(tr.w, fb, z);
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) tr.w[i] = z[i];
            } else {
                const real cw1 = (real)dt;
                const real cw2 = (real)(0.5 * (1.0 - 2.0 * cv.newmark_beta) * dt * dt);
                const real cw3 = (real)((1.0 - cv.newmark_gamma) * dt);
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) {
                    tr.w[i] = fma(cw2, tr.a[i], fma(cw1, tr.v[i], tr.w[i]));   /* predictor d~ */
                    tr.v[i] = fma(cw3, tr.a[i], tr.v[i]);                       /* predictor v~ */
                }
                tr.step_solve(tr.w, f_next, tr.a);                              /* a' = A^-1 (f' e_N - K d~) */
                const real cw4 = (real)(cv.newmark_beta * dt * dt);
                const real cw5 = (real)(cv.newmark_gamma * dt);
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) {
                    tr.w[i] = fma(cw4, tr.a[i], tr.w[i]);
                    tr.v[i] = fma(cw5, tr.a[i], tr.v[i]);
                }
            }
            f_prev = f_next;
        }
    }

    if (!live) return;
    /* ---- formula tree and outputs ---- */
    double val[PDEB_MAX_TREE];
    double rob = PDEB_FAILED_ROBUSTNESS;
    if (status == PDEB_OK) {
        const int nt = b.n_tree[s];
        for (int j = 0; j < nt; ++j) {
            const int op = b.tree_op[(int64_t)j * st + s], lh = b.tree_lhs[(int64_t)j * st + s], rh = b.tree_rhs[(int64_t)j * st + s];
            double x = NAN;
            if (op == PDEB_TREE_LEAF) { if (lh >= 0 && lh < np) x = (double)slots.R(lh, 2); else status = PDEB_ST_BADPARAM; }
            else if ((op == PDEB_TREE_AND || op == PDEB_TREE_OR) && lh >= 0 && lh < j && rh >= 0 && rh < j) {
                const double a0 = val[lh], a1 = val[rh];
                if (a0 != a0 || a1 != a1) x = NAN;
                else if (op == PDEB_TREE_AND) x = a0 < a1 ? a0 : a1;
                else x = a0 > a1 ? a0 : a1;
            } else status = PDEB_ST_BADPARAM;
            val[j] = x;
        }
        if (status == PDEB_OK) {
            rob = val[nt - 1];
            if (!isfinite(rob)) status = PDEB_ST_NONFINITE;
            for (int j = 0; j < np; ++j) if (!isfinite((double)slots.R(j, 2))) status = PDEB_ST_NONFINITE;
        }
    }
    out.robustness[s] = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
    if (out.status) out.status[s] = status;
    if (out.clause) {
        for (int j = 0; j < b.pmax; ++j)
            out.clause[(int64_t)j * st + s] = (status == PDEB_OK && j < np) ? (double)slots.R(j, 2) : PDEB_FAILED_ROBUSTNESS;
    }
}

#endif /* PDEB_CORE_H */
