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

/* per-thread clause slots in shared memory; element p of a field is at field[p * stride] */
template <typename real>
struct PdebClauseSlots {
    real* mh;        /* m * h: profile slope per node index */
    real* c0;        /* b - g0 */
    real* acc;       /* running temporal min (G) / max (F) */
    int32_t* klo;    /* first window step */
    int32_t* khi;    /* last window step */
    uint32_t* nmask; /* bit i set: node i is inside [x_lo, x_hi] */
    uint32_t* info;  /* bits 0-1 rel, bit 2 top, bit 3 use interpolated end points */
    int stride;
};

template <typename real>
PDEB_HD size_t pdeb_clause_smem_bytes(int threads, int pmax) {
    return (size_t)threads * pmax * (3 * sizeof(real) + 4 * sizeof(int32_t));
}

template <typename real>
PDEB_HD PdebClauseSlots<real> pdeb_clause_slots(void* base, int threads, int pmax, int tid) {
    PdebClauseSlots<real> s;
    real* r = (real*)base;
    s.mh = r + tid;
    s.c0 = r + (size_t)threads * pmax + tid;
    s.acc = r + (size_t)2 * threads * pmax + tid;
    int32_t* q = (int32_t*)(r + (size_t)3 * threads * pmax);
    s.klo = q + tid;
    s.khi = q + (size_t)threads * pmax + tid;
    s.nmask = (uint32_t*)(q + (size_t)2 * threads * pmax) + tid;
    s.info = (uint32_t*)(q + (size_t)3 * threads * pmax) + tid;
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

/* running piecewise-linear input: keeps the current segment in registers */
template <typename real>
struct PdebPwl {
    const double* kt;
    const double* ky;
    int64_t stride;
    int nk, j;
    double t1;        /* end of the current segment */
    real t0, y0, slope;
    double tfirst, tlast;

    PDEB_HD void load_segment() {
        double a = kt[(int64_t)j * stride], b = kt[(int64_t)(j + 1) * stride];
        double ya = ky[(int64_t)j * stride], yb = ky[(int64_t)(j + 1) * stride];
        t0 = (real)a;
        y0 = (real)ya;
        slope = (real)((yb - ya) / (b - a));
        t1 = b;
    }
    PDEB_HD void init(const double* kt_, const double* ky_, int64_t stride_, int nk_) {
        kt = kt_; ky = ky_; stride = stride_; nk = nk_; j = 0;
        t0 = 0; y0 = 0; slope = 0; t1 = 1e300; tfirst = 0; tlast = 0;
        if (nk == 1) { y0 = (real)ky[0]; }
        if (nk >= 2) { tfirst = kt[0]; tlast = kt[(int64_t)(nk - 1) * stride]; load_segment(); }
    }
    PDEB_HD real eval(double t) {
        if (nk < 2) return y0;
        double tc = t < tfirst ? tfirst : (t > tlast ? tlast : t);
        while (j + 2 < nk && tc >= t1) { ++j; load_segment(); }
        return (real)fma((double)slope, tc - (double)t0, (double)y0);
    }
};

/*
 * Register-family trajectory.  KIND: PDEB_KIND_*; N: elements = free dofs (nodes 1..N), N <= 31.
 * Arrays are indexed by node, index 0 is the Dirichlet node (deviation 0).
 */
template <int KIND, int N, typename real>
struct PdebTrajReg {
    static_assert(N >= 2 && N <= 31, "register family handles 2..31 elements");

    /* two-material coefficient tables, index 0 = outer, 1 = inner */
    real ro[2];        /* off-diagonal of the right-hand-side operator, per element */
    real rd2[3];       /* its diagonal for rows between two elements: oo, mixed, ii */
    real rdN[2];       /* diagonal of row N (one element only) */
    uint32_t omask;    /* bit e: element e is of the outer material */
    real l[N + 1];     /* LDL^T of the step matrix: l[i] multiplies row i-1 into row i (i = 2..N) */
    real dinv[N + 1];
    real w[N + 2];     /* heat: deviation; wave: displacement deviation */
    real v[KIND == PDEB_KIND_WAVE ? N + 2 : 1];
    real a[KIND == PDEB_KIND_WAVE ? N + 2 : 1];

    PDEB_HD bool outer(int e) const { return (omask >> e) & 1u; }
    PDEB_HD real sel_ro(int e) const { return outer(e) ? ro[0] : ro[1]; }
    PDEB_HD real sel_rd(int i) const {   /* row i, 1 <= i <= N */
        if (i == N) return outer(N - 1) ? rdN[0] : rdN[1];
        bool ol = outer(i - 1), orr = outer(i);
        return ol ? (orr ? rd2[0] : rd2[1]) : (orr ? rd2[1] : rd2[2]);
    }

    /* y = R x on the free rows, R tridiagonal from the tables above; x[0] = 0 and x[N+1] unused */
    PDEB_HD real rrow(const real* x, int i) const {
        real r = sel_rd(i) * x[i];
        if (i > 1) r = fma(sel_ro(i - 1), x[i - 1], r);
        if (i < N) r = fma(sel_ro(i), x[i + 1], r);
        return r;
    }

    /* solve (L D L^T) x = z in place; on entry z[i] already holds the forward-eliminated value */
    PDEB_HD void back_substitute(real* z) const {
        z[N] = z[N] * dinv[N];
        PDEB_UNROLL
        for (int i = N - 1; i >= 1; --i) z[i] = fma(-l[i + 1], z[i + 1], z[i] * dinv[i]);
    }

    /*
     * Set up coefficients and factors.  Returns false on a singular / non-finite factorisation.
     * cA multiplies K in the step matrix A = M + cA K; the right-hand-side operator is
     * R = M - cR K for heat (cR = (1-th) dt) and R = -K for wave.
     */
    PDEB_HD bool setup(double L, double x_a, double x_b, double rho_in, double rho_out, double E_in, double E_out,
                       int lumped, int sampling, double cA, double cR, double* Mdiag_out /* [N+1] or NULL */) {
        double h = L / N;
        omask = 0;
        PDEB_UNROLL
        for (int e = 0; e < N; ++e) {
            double xs = (sampling == 1) ? e * h : (sampling == 2) ? (e + 1) * h : (e + 0.5) * h;
            if (xs < x_a || xs > x_b) omask |= (1u << e);
        }
        double md[2], mo[2], ke[2], ad[2], ao[2];
        double rho[2] = { rho_out, rho_in }, E[2] = { E_out, E_in };
        for (int q = 0; q < 2; ++q) {
            double m = rho[q] * h;
            ke[q] = E[q] / h;
            md[q] = lumped ? 0.5 * m : m / 3.0;
            mo[q] = lumped ? 0.0 : m / 6.0;
            ad[q] = md[q] + cA * ke[q];
            ao[q] = mo[q] - cA * ke[q];
            double rdq, roq;
            if (KIND == PDEB_KIND_HEAT) { rdq = md[q] - cR * ke[q]; roq = mo[q] + cR * ke[q]; }
            else { rdq = -ke[q]; roq = ke[q]; }
            ro[q] = (real)roq;
            rdN[q] = (real)rdq;
            if (q == 0) { rd2[0] = (real)(2.0 * rdq); }
            else { rd2[2] = (real)(2.0 * rdq); }
        }
        {
            double rd0 = (KIND == PDEB_KIND_HEAT) ? md[0] - cR * ke[0] : -ke[0];
            double rd1 = (KIND == PDEB_KIND_HEAT) ? md[1] - cR * ke[1] : -ke[1];
            rd2[1] = (real)(rd0 + rd1);
        }
        /* LDL^T of A over rows 1..N (double precision, then rounded to real) */
        double dprev = 0.0, lo_prev = 0.0;
        bool ok = true;
        PDEB_UNROLL
        for (int i = 1; i <= N; ++i) {
            int ql = outer(i - 1) ? 0 : 1;
            double diag = ad[ql];
            if (i < N) diag += ad[outer(i) ? 0 : 1];
            double li = 0.0;
            if (i > 1) { li = lo_prev * dprev; diag -= li * lo_prev; }   /* lo_prev = A[i][i-1], dprev = 1/D[i-1] */
            double di = 1.0 / diag;
            if (!(diag > 0.0) || !isfinite(di)) ok = false;
            l[i] = (real)li;
            dinv[i] = (real)di;
            dprev = di;
            lo_prev = (i < N) ? ao[outer(i) ? 0 : 1] : 0.0;
            if (Mdiag_out) {
                double mdi = md[ql];
                if (i < N) mdi += md[outer(i) ? 0 : 1];
                Mdiag_out[i] = mdi;
            }
        }
        return ok;
    }
};

/* evaluate one clause on the current state u = g0 + w; returns the spatial margin */
template <int N, typename real>
PDEB_HD real pdeb_clause_margin(const real* w, real mh, real c0, uint32_t nmask, uint32_t info, int eq_rule, real eq_eps,
                                bool need_min, bool need_max) {
    const real BIG = (real)1e300 < (real)INFINITY ? (real)1e300 : (real)3.0e38;
    real vmin = BIG, vmax = -BIG;
    if (need_min && need_max) {
        PDEB_UNROLL
        for (int i = 0; i <= N; ++i) {
            real val = (i == 0) ? (real)0 : fma(-mh, (real)i, w[i]);
            bool in = (nmask >> i) & 1u;
            if (in && val < vmin) vmin = val;
            if (in && val > vmax) vmax = val;
        }
    } else if (need_min) {
        PDEB_UNROLL
        for (int i = 0; i <= N; ++i) {
            real val = (i == 0) ? (real)0 : fma(-mh, (real)i, w[i]);
            bool in = (nmask >> i) & 1u;
            if (in && val < vmin) vmin = val;
        }
    } else {
        PDEB_UNROLL
        for (int i = 0; i <= N; ++i) {
            real val = (i == 0) ? (real)0 : fma(-mh, (real)i, w[i]);
            bool in = (nmask >> i) & 1u;
            if (in && val > vmax) vmax = val;
        }
    }
    uint32_t rel = info & 3u;
    real r;
    if (rel == PDEB_REL_GT) r = vmin - c0;
    else if (rel == PDEB_REL_LT) r = c0 - vmax;
    else {
        real lo = vmin - (c0 - eq_eps), hi = (c0 + eq_eps) - vmax;   /* eq_eps = 0 unless eq_rule == 1 */
        r = lo < hi ? lo : hi;
        (void)eq_rule;
    }
    return r;
}

/* state value at node index e (runtime) by a select chain: no dynamic register indexing */
template <int N, typename real>
PDEB_HD real pdeb_pick(const real* w, int e) {
    real r = 0;
    PDEB_UNROLL
    for (int i = 1; i <= N; ++i) r = (i == e) ? w[i] : r;
    return r;
}

/*
 * Whole trajectory of sample s.  `slots` is this thread's clause column in shared memory.
 * Writes robustness / clause / status.
 */
template <int KIND, int N, typename real>
PDEB_HD void pdeb_run_reg(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, int64_t s,
                          PdebClauseSlots<real> slots, bool live) {
    /* `live` is false for padding lanes of a partially filled warp: they run the loop with S = 0 */
    const int64_t st = b.stride;
    int status = PDEB_OK;
    double L = 1, dt = 1, T = 1, g0 = 0;
    int np = 0;
    int64_t S = 0;
    PdebTrajReg<KIND, N, real> tr;
    PdebPwl<real> in;
    in.init(nullptr, nullptr, 0, 0);
    double h = 1;
    real cw1 = 0, cw2 = 0, cw3 = 0, cw4 = 0, cw5 = 0;   /* integrator constants */
    PDEB_UNROLL
    for (int i = 0; i <= N + 1; ++i) tr.w[i] = 0;

    if (live) {
        L = b.L[s]; dt = b.dt[s]; T = b.T[s]; g0 = b.g0[s];
        double x_a = b.x_a[s], x_b = b.x_b[s];
        double rho_in = b.rho_in[s], rho_out = b.rho_out[s], E_in = b.E_in[s], E_out = b.E_out[s];
        np = b.n_pred[s];
        int nk = b.n_knots[s];
        int nt = b.n_tree[s];
        if (!pdeb_params_ok(KIND, N, L, dt, T, g0, rho_in, rho_out, E_in, E_out)) status = PDEB_ST_BADPARAM;
        if (np < 1 || np > b.pmax || nk < 0 || nk > b.kmax || nt < 1 || nt > b.tmax) status = PDEB_ST_BADPARAM;
        if (status == PDEB_OK) {
            for (int j = 1; j < nk; ++j)
                if (!(b.knot_t[(int64_t)j * st + s] > b.knot_t[(int64_t)(j - 1) * st + s])) status = PDEB_ST_BADPARAM;
            for (int j = 0; j < np; ++j)
                if (b.pred_rel[(int64_t)j * st + s] > 2 || b.pred_top[(int64_t)j * st + s] > 1) status = PDEB_ST_BADPARAM;
        }
        if (status == PDEB_OK) {
            S = pdeb_step_count(T, dt, cv.steps_rule);
            if (S > PDEB_MAX_STEPS) status = PDEB_ST_BADPARAM;
        }
        if (status == PDEB_OK) {
            h = L / N;
            int lumped = (KIND == PDEB_KIND_HEAT) ? cv.mass_heat : cv.mass_wave;
            double cA = (KIND == PDEB_KIND_HEAT) ? cv.theta * dt : cv.newmark_beta * dt * dt;
            double cR = (1.0 - cv.theta) * dt;
            double Md[N + 1];
            bool ok = tr.setup(L, x_a, x_b, rho_in, rho_out, E_in, E_out, lumped, cv.material_sampling, cA, cR,
                               KIND == PDEB_KIND_WAVE ? Md : nullptr);
            if (!ok) status = PDEB_ST_NONFINITE;
            in.init(b.knot_t + s, b.knot_y + s, st, nk);
            if (KIND == PDEB_KIND_HEAT) {
                cw1 = (real)((1.0 - cv.theta) * dt);
                cw2 = (real)(cv.theta * dt);
            } else {
                cw1 = (real)dt;
                cw2 = (real)(0.5 * (1.0 - 2.0 * cv.newmark_beta) * dt * dt);
                cw3 = (real)((1.0 - cv.newmark_gamma) * dt);
                cw4 = (real)(cv.newmark_beta * dt * dt);
                cw5 = (real)(cv.newmark_gamma * dt);
                /* wave initial state: d = 0 on the free nodes (deviation -g0), v = 0, a0 = M^-1 (f0 e_N - K w0) */
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) { tr.w[i] = (real)(-g0); tr.v[i] = 0; }
                tr.v[0] = 0; tr.v[N + 1] = 0; tr.a[0] = 0; tr.a[N + 1] = 0;
                real f0 = in.eval(0.0);
                if (lumped) {
                    PDEB_UNROLL
                    for (int i = 1; i <= N; ++i) {
                        real r = tr.rrow(tr.w, i);
                        if (i == N) r += f0;
                        tr.a[i] = (real)((double)r / Md[i]);
                    }
                } else {
                    /* consistent mass: one Thomas solve with M (set-up only) */
                    double cp[N + 1], x[N + 1];
                    double mo_prev = 0.0;
                    PDEB_UNROLL
                    for (int i = 1; i <= N; ++i) {
                        real r = tr.rrow(tr.w, i);
                        if (i == N) r += f0;
                        double mo_i = (i < N) ? (tr.outer(i) ? rho_out : rho_in) * h / 6.0 : 0.0;
                        double den = Md[i] - ((i > 1) ? mo_prev * cp[i - 1] : 0.0);
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
        /* clause slots */
        for (int j = 0; j < np; ++j) {
            double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
            double xlo = b.pred_xlo[(int64_t)j * st + s], xhi = b.pred_xhi[(int64_t)j * st + s];
            double qlo = b.pred_tlo[(int64_t)j * st + s] / dt, qhi = b.pred_thi[(int64_t)j * st + s] / dt;
            double a0, a1;
            if (cv.window_rule == 1) { a0 = trunc(qlo); a1 = trunc(qhi); }
            else { a0 = ceil(qlo - PDEB_EPS); a1 = floor(qhi + PDEB_EPS); }
            if (!(a0 >= 0.0)) a0 = 0.0;
            if (!(a1 <= (double)S)) a1 = (double)S;
            if (a0 > a1) {
                double k = floor(0.5 * (b.pred_tlo[(int64_t)j * st + s] + b.pred_thi[(int64_t)j * st + s]) / dt + 0.5);
                if (!(k >= 0.0)) k = 0.0;
                if (k > (double)S) k = (double)S;
                a0 = a1 = k;
            }
            double ilo = ceil(xlo / h - PDEB_EPS), ihi = floor(xhi / h + PDEB_EPS);
            if (!(ilo >= 0.0)) ilo = 0.0;
            if (!(ihi <= (double)N)) ihi = (double)N;
            uint32_t nm = 0;
            if (ilo <= ihi) {
                int i0 = (int)ilo, i1 = (int)ihi;
                nm = (i1 >= 31 ? 0xffffffffu : ((1u << (i1 + 1)) - 1u)) & ~((1u << i0) - 1u);
            }
            uint32_t top = b.pred_top[(int64_t)j * st + s], rel = b.pred_rel[(int64_t)j * st + s];
            uint32_t endp = (nm == 0 || cv.region_rule == 1) ? 1u : 0u;
            slots.mh[j * slots.stride] = (real)(m * h);
            slots.c0[j * slots.stride] = (real)(bb - g0);
            slots.acc[j * slots.stride] = (top == PDEB_TOP_G) ? (real)INFINITY : (real)(-INFINITY);
            slots.klo[j * slots.stride] = (int32_t)a0;
            slots.khi[j * slots.stride] = (int32_t)a1;
            slots.nmask[j * slots.stride] = nm;
            slots.info[j * slots.stride] = rel | (top << 2) | (endp << 3);
        }
    }

    const real eq_eps = (cv.eq_rule == 1) ? (real)cv.eq_eps : (real)0;
    real f_prev = in.eval(0.0);
    int32_t next_ev = 0;      /* next step at which the set of open windows may change */
    uint32_t open = 0;        /* bit j: step k is inside clause j's window */
    const int32_t S32 = (int32_t)S;

    for (int32_t k = 0;; ++k) {
        const bool on = (k <= S32);
        /* ---- observe step k ---- */
        if (on && k == next_ev) {
            open = 0;
            next_ev = 0x7fffffff;
            for (int j = 0; j < np; ++j) {
                int32_t a0 = slots.klo[j * slots.stride], a1 = slots.khi[j * slots.stride];
                if (k >= a0 && k <= a1) { open |= (1u << j); if (a1 + 1 < next_ev) next_ev = a1 + 1; }
                else if (k < a0 && a0 < next_ev) next_ev = a0;
            }
        }
        uint32_t todo = on ? open : 0u;
        while (PDEB_ANY(todo != 0u)) {
            const bool has = todo != 0u;
            int j = 0;
            if (has) { j = __builtin_ffs((int)todo) - 1; todo &= todo - 1u; }
            real mh = 0, c0 = 0, acc = 0;
            uint32_t nm = 0, info = 0;
            if (has) {
                mh = slots.mh[j * slots.stride];
                c0 = slots.c0[j * slots.stride];
                acc = slots.acc[j * slots.stride];
                nm = slots.nmask[j * slots.stride];
                info = slots.info[j * slots.stride];
            }
            const uint32_t rel = info & 3u;
            const bool want_min = has && rel != PDEB_REL_LT, want_max = has && rel != PDEB_REL_GT;
            const bool need_min = PDEB_ANY(want_min), need_max = PDEB_ANY(want_max);
            real r = pdeb_clause_margin<N, real>(tr.w, mh, c0, nm, info, cv.eq_rule, eq_eps, need_min, need_max);
            if (PDEB_ANY(has && (info & 8u))) {
                /* rare: interpolated interval end points (no node inside, or region_rule 1) */
                if (has && (info & 8u)) {
                    if (nm == 0) r = (real)INFINITY;
                    double xe[2] = { b.pred_xlo[(int64_t)j * st + s], b.pred_xhi[(int64_t)j * st + s] };
                    double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
                    for (int q = 0; q < 2; ++q) {
                        double xx = xe[q] < 0 ? 0 : (xe[q] > L ? L : xe[q]);
                        int e = (int)floor(xx / h);
                        if (e < 0) e = 0;
                        if (e > N - 1) e = N - 1;
                        double xi = (xx - e * h) / h;
                        double ue = (e == 0) ? 0.0 : (double)pdeb_pick<N, real>(tr.w, e);
                        double uf = (double)pdeb_pick<N, real>(tr.w, e + 1);
                        double u = g0 + (ue + xi * (uf - ue));
                        double pr = m * xx + bb;
                        double mg;
                        if (rel == PDEB_REL_GT) mg = u - pr;
                        else if (rel == PDEB_REL_LT) mg = pr - u;
                        else { double lo = u - (pr - (double)eq_eps), hi = (pr + (double)eq_eps) - u; mg = lo < hi ? lo : hi; }
                        if ((real)mg < r) r = (real)mg;
                    }
                }
            }
            if (has) {
                bool isG = ((info >> 2) & 1u) == PDEB_TOP_G;
                bool take = isG ? (r < acc) : (r > acc);
                if (take || r != r) slots.acc[j * slots.stride] = r;
            }
        }
        if (!PDEB_ANY(k < S32)) break;
        /* ---- advance to step k+1 ---- */
        if (k < S32) {
            real f_next = in.eval((double)(k + 1) * dt);
            if (KIND == PDEB_KIND_HEAT) {
                real fb = cw1 * f_prev + cw2 * f_next;
                /* forward elimination fused with the right-hand side: z_i = (R w)_i - l_i z_{i-1} */
                real z[N + 2];
                z[0] = 0;
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) {
                    real r = tr.rrow(tr.w, i);
                    if (i == N) r += fb;
                    z[i] = (i > 1) ? fma(-tr.l[i], z[i - 1], r) : r;
                }
                tr.back_substitute(z);
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) tr.w[i] = z[i];
            } else {
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) {
                    tr.w[i] = fma(cw2, tr.a[i], fma(cw1, tr.v[i], tr.w[i]));   /* predictor d~ */
                    tr.v[i] = fma(cw3, tr.a[i], tr.v[i]);                       /* predictor v~ */
                }
                real z[N + 2];
                z[0] = 0;
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) {
                    real r = tr.rrow(tr.w, i);                                  /* -K d~ */
                    if (i == N) r += f_next;
                    z[i] = (i > 1) ? fma(-tr.l[i], z[i - 1], r) : r;
                }
                tr.back_substitute(z);
                PDEB_UNROLL
                for (int i = 1; i <= N; ++i) {
                    tr.a[i] = z[i];
                    tr.w[i] = fma(cw4, z[i], tr.w[i]);
                    tr.v[i] = fma(cw5, z[i], tr.v[i]);
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
        real chk = 0;
        PDEB_UNROLL
        for (int i = 1; i <= N; ++i) chk = fma(tr.w[i], (real)0, chk);
        if (!(chk == (real)0)) status = PDEB_ST_NONFINITE;
        int nt = b.n_tree[s];
        for (int j = 0; j < nt; ++j) {
            int op = b.tree_op[(int64_t)j * st + s], lh = b.tree_lhs[(int64_t)j * st + s], rh = b.tree_rhs[(int64_t)j * st + s];
            double x = NAN;
            if (op == PDEB_TREE_LEAF) { if (lh >= 0 && lh < np) x = (double)slots.acc[lh * slots.stride]; else status = PDEB_ST_BADPARAM; }
            else if ((op == PDEB_TREE_AND || op == PDEB_TREE_OR) && lh >= 0 && lh < j && rh >= 0 && rh < j) {
                double a0 = val[lh], a1 = val[rh];
                if (a0 != a0 || a1 != a1) x = NAN;
                else if (op == PDEB_TREE_AND) x = a0 < a1 ? a0 : a1;
                else x = a0 > a1 ? a0 : a1;
            } else status = PDEB_ST_BADPARAM;
            val[j] = x;
        }
        if (status == PDEB_OK) {
            rob = val[nt - 1];
            if (!isfinite(rob)) status = PDEB_ST_NONFINITE;
            for (int j = 0; j < np; ++j) if (!isfinite((double)slots.acc[j * slots.stride])) status = PDEB_ST_NONFINITE;
        }
    }
    out.robustness[s] = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
    if (out.status) out.status[s] = status;
    if (out.clause) {
        for (int j = 0; j < b.pmax; ++j)
            out.clause[(int64_t)j * st + s] = (status == PDEB_OK && j < np) ? (double)slots.acc[j * slots.stride] : PDEB_FAILED_ROBUSTNESS;
    }
}

#endif /* PDEB_CORE_H */
