/*
 * pde_oracle.c — FP64 CPU restatement of PDE-Controller's utility-evaluation path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA path, and the
 * timed CPU baseline of bench.py.  Nothing under pde_controller_b200/ may include,
 * link or call it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do.
 *
 * PARITY UNPINNED.  The arithmetic of the reference path lives in the Python-2
 * package `control.femformal` (upstream: femformal + stlmilp, no version pinned
 * anywhere in the reference: requirements.txt, train_test_combined.yml and
 * README.md:71-75 are silent).  That package is NOT in the reference snapshot;
 * only its call sites are:
 *     test/PDEcontrol/evaluation/eval/eval_robustness_wrapper.py:5,10,16
 *     utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3   (heatlinfem.heatlinfem_mix, logic.APCont, cspec)
 *     utils/few_shot_prompts/examples/one_d_wave/examples.jsonl:1-3   (mechlinfem.mechlinfem)
 * and the reference has no test, golden vector or labelled sample for this path.
 * This oracle therefore restates the PUBLISHED algorithms those call sites name
 * (linear FEM on a uniform 1D mesh, generalised trapezoidal rule, Newmark,
 * standard quantitative STL semantics) under the declared conventions C1..C12 of
 * SURVEY.md §8c, and is pinned only by analytical known-answer tests
 * (tests/test_oracle_kat.py) and an independent numpy restatement
 * (oracle/pde_oracle_np.py).
 *
 * Per sample (all FP64):
 *   mesh      x_i = i L/N, i = 0..N                      [call site: xpart = np.linspace(0, L, N + 1)]
 *   material  element e takes (rho,E)_out if its sampling point xs has xs < x_a or xs > x_b,
 *             else (rho,E)_in                            [call site: rho = lambda x: A if x < a or x > b else B]
 *   M, K      P1 elements: K_e = E_e/h [[1,-1],[-1,1]];  M_e = rho_e h/6 [[2,1],[1,2]] (consistent)
 *             or rho_e h/2 I (lumped)                    [heatlinfem_mix / mechlinfem call sites]
 *   BCs       d_0 = g0 for all t (g = [g0, None]); node N free, loaded by the PWL input u(t)
 *   IC        heat d = g0; wave d = 0, v = 0             [figures/case_study.png]
 *   heat      generalised trapezoidal, v-form (Hughes):  v0 = M^-1 (F0 - K d0);
 *             dp = d + (1-th) dt v;  (M + th dt K) v' = F' - K dp;  d' = dp + th dt v'
 *   wave      Newmark a-form: a0 = M^-1 (F0 - K d0);
 *             dp = d + dt v + (1-2b)/2 dt^2 a;  vp = v + (1-g) dt a;
 *             (M + b dt^2 K) a' = F' - K dp;  d' = dp + b dt^2 a';  v' = vp + g dt a'
 *   horizon   only states up to the last step any clause observes can influence the result; a blow-up after that
 *             step is not an error (status NONFINITE means a non-finite OBSERVED value)
 *   STL       clause = G|F over [t_lo,t_hi] of  forall x in [x_lo,x_hi]: u(x) - (m x + b) REL 0
 *             margin(>) = u-p, margin(<) = p-u, margin(=) = -|u-p|;  forall x = min over nodes
 *             in the interval; G = min over window steps, F = max; & = min, | = max.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/pdeb200.h"   /* data layout only (PdebBatch, PdebConventions, enums) */

#define PDEO_EPS 1e-9

void pdeo_default_conventions(PdebConventions* c) {
    memset(c, 0, sizeof(*c));
    c->material_sampling = 0;
    c->mass_heat = 0;
    c->mass_wave = 1;
    c->steps_rule = 0;
    c->window_rule = 0;
    c->region_rule = 0;
    c->eq_rule = 0;
    c->eq_eps = 0.0;
    c->theta = 0.5;
    c->newmark_beta = 0.25;
    c->newmark_gamma = 0.5;
}

/* ---- one sample, unpacked ------------------------------------------------ */
typedef struct {
    int kind, N;
    double L, dt, T, g0, x_a, x_b, rho_in, rho_out, E_in, E_out;
    int nk;
    double kt[64], ky[64];
    int np;
    double xlo[PDEB_MAX_PRED], xhi[PDEB_MAX_PRED], pm[PDEB_MAX_PRED], pb[PDEB_MAX_PRED];
    double tlo[PDEB_MAX_PRED], thi[PDEB_MAX_PRED];
    int rel[PDEB_MAX_PRED], top[PDEB_MAX_PRED];
    int nt;
    int top_op[PDEB_MAX_TREE], lhs[PDEB_MAX_TREE], rhs[PDEB_MAX_TREE];
} Sample;

static int unpack(const PdebBatch* in, int64_t s, Sample* p) {
    int64_t st = in->stride;
    p->kind = in->kind[s];
    p->N = in->n_elem[s];
    p->L = in->L[s]; p->dt = in->dt[s]; p->T = in->T[s]; p->g0 = in->g0[s];
    p->x_a = in->x_a[s]; p->x_b = in->x_b[s];
    p->rho_in = in->rho_in[s]; p->rho_out = in->rho_out[s];
    p->E_in = in->E_in[s]; p->E_out = in->E_out[s];
    p->nk = in->n_knots[s];
    if (p->nk < 0 || p->nk > in->kmax || p->nk > 64) return -1;
    for (int j = 0; j < p->nk; ++j) {
        /* knot_t == NULL: uniform knots, exactly numpy's linspace(0, T, nk) (C11) */
        if (in->knot_t) p->kt[j] = in->knot_t[j * st + s];
        else p->kt[j] = (j == p->nk - 1 && p->nk > 1) ? p->T : (double)j * (p->T / (double)(p->nk - 1 > 0 ? p->nk - 1 : 1));
        p->ky[j] = in->knot_y[j * st + s];
    }
    p->np = in->n_pred[s];
    if (p->np < 1 || p->np > in->pmax || p->np > PDEB_MAX_PRED) return -1;
    for (int j = 0; j < p->np; ++j) {
        p->xlo[j] = in->pred_xlo[j * st + s]; p->xhi[j] = in->pred_xhi[j * st + s];
        p->pm[j] = in->pred_m[j * st + s];    p->pb[j] = in->pred_b[j * st + s];
        p->tlo[j] = in->pred_tlo[j * st + s]; p->thi[j] = in->pred_thi[j * st + s];
        p->rel[j] = in->pred_rel[j * st + s]; p->top[j] = in->pred_top[j * st + s];
        if (p->rel[j] > 2 || p->top[j] > 1) return -1;
    }
    p->nt = in->n_tree[s];
    if (p->nt < 1 || p->nt > in->tmax || p->nt > PDEB_MAX_TREE) return -1;
    for (int j = 0; j < p->nt; ++j) {
        p->top_op[j] = in->tree_op[j * st + s];
        p->lhs[j] = in->tree_lhs[j * st + s];
        p->rhs[j] = in->tree_rhs[j * st + s];
        if (p->top_op[j] == PDEB_TREE_LEAF) {
            if (p->lhs[j] < 0 || p->lhs[j] >= p->np) return -1;
        } else if (p->top_op[j] == PDEB_TREE_AND || p->top_op[j] == PDEB_TREE_OR) {
            if (p->lhs[j] < 0 || p->lhs[j] >= j || p->rhs[j] < 0 || p->rhs[j] >= j) return -1;
        } else return -1;
    }
    if (p->kind != PDEB_KIND_HEAT && p->kind != PDEB_KIND_WAVE) return -1;
    if (p->N < 1 || p->N > PDEB_MAX_ELEM) return -1;
    if (!(p->L > 0) || !(p->dt > 0) || !(p->T > 0)) return -1;
    if (!(p->rho_in > 0) || !(p->rho_out > 0) || !(p->E_in > 0) || !(p->E_out > 0)) return -1;
    if (!isfinite(p->L) || !isfinite(p->dt) || !isfinite(p->T) || !isfinite(p->g0)) return -1;
    for (int j = 1; j < p->nk; ++j) if (!(p->kt[j] > p->kt[j - 1])) return -1;
    return 0;
}

/* C6: number of time steps */
static int64_t step_count(double T, double dt, int rule) {
    double q = T / dt;
    int64_t S;
    if (rule == 1) S = (int64_t)floor(q + 0.5);
    else if (rule == 2) S = (int64_t)floor(q + PDEO_EPS);
    else S = (int64_t)ceil(q - PDEO_EPS);
    if (S < 1) S = 1;
    return S;
}

/* C11: piecewise-linear input, held constant outside the knots */
static double pwl(const Sample* p, double t) {
    if (p->nk == 0) return 0.0;
    if (t <= p->kt[0]) return p->ky[0];
    if (t >= p->kt[p->nk - 1]) return p->ky[p->nk - 1];
    int j = 0;
    while (t >= p->kt[j + 1]) ++j;
    return p->ky[j] + (p->ky[j + 1] - p->ky[j]) * ((t - p->kt[j]) / (p->kt[j + 1] - p->kt[j]));
}

/* tridiagonal LU of (lo, di, up) of size n without pivoting (SPD here) */
static void tri_factor(int n, const double* lo, const double* di, const double* up, double* cp, double* inv) {
    inv[0] = 1.0 / di[0];
    cp[0] = up[0] * inv[0];
    for (int i = 1; i < n; ++i) {
        inv[i] = 1.0 / (di[i] - lo[i] * cp[i - 1]);
        cp[i] = up[i] * inv[i];
    }
}
static void tri_solve(int n, const double* lo, const double* cp, const double* inv, double* x /* rhs in, solution out */) {
    x[0] = x[0] * inv[0];
    for (int i = 1; i < n; ++i) x[i] = (x[i] - lo[i] * x[i - 1]) * inv[i];
    for (int i = n - 2; i >= 0; --i) x[i] = x[i] - cp[i] * x[i + 1];
}

static double margin(int rel, double u, double p, const PdebConventions* c) {
    if (rel == PDEB_REL_GT) return u - p;
    if (rel == PDEB_REL_LT) return p - u;
    if (c->eq_rule == 1) { double a = u - (p - c->eq_eps), b = (p + c->eq_eps) - u; return a < b ? a : b; }
    return -fabs(u - p);
}

static double interp(const double* u, int N, double h, double x) {
    int e = (int)floor(x / h);
    if (e < 0) e = 0;
    if (e > N - 1) e = N - 1;
    double xi = (x - e * h) / h;
    return u[e] + xi * (u[e + 1] - u[e]);
}

/* forall x in [x_lo,x_hi]: min over the discrete sample points of the interval (C8) */
static double spatial_margin(const Sample* p, int j, const double* u, const double* x, double h, const PdebConventions* c) {
    int N = p->N;
    int ilo = (int)ceil(p->xlo[j] / h - PDEO_EPS), ihi = (int)floor(p->xhi[j] / h + PDEO_EPS);
    if (ilo < 0) ilo = 0;
    if (ihi > N) ihi = N;
    double r = INFINITY;
    for (int i = ilo; i <= ihi; ++i) {
        double m = margin(p->rel[j], u[i], p->pm[j] * x[i] + p->pb[j], c);
        if (m < r) r = m;
    }
    if (ilo > ihi || c->region_rule == 1) {
        double xe[2] = { p->xlo[j], p->xhi[j] };
        for (int q = 0; q < 2; ++q) {
            double xx = xe[q] < 0 ? 0 : (xe[q] > p->L ? p->L : xe[q]);
            double m = margin(p->rel[j], interp(u, N, h, xx), p->pm[j] * xx + p->pb[j], c);
            if (m < r) r = m;
        }
    }
    return r;
}

/* K d in assembled element-flux form, rows 0..N */
static void apply_K(int N, const double* ke, const double* d, double* out) {
    for (int i = 0; i <= N; ++i) out[i] = 0.0;
    for (int e = 0; e < N; ++e) {
        double q = ke[e] * (d[e + 1] - d[e]);
        out[e] -= q;
        out[e + 1] += q;
    }
}

/*
 * Simulate one sample.  traj (optional) receives d at every step, (S+1) x (N+1);
 * vel (optional) the wave velocity.  clause[j] receives per-clause robustness.
 */
/* per-thread workspace: grown on demand, reused across samples (the per-sample path allocates nothing) */
typedef struct { double* buf; size_t cap; } Workspace;
static double* ws_get(Workspace* w, size_t doubles) {
    if (w->cap < doubles) {
        free(w->buf);
        w->buf = (double*)malloc(doubles * sizeof(double));
        w->cap = w->buf ? doubles : 0;
    }
    return w->buf;
}

static int eval_one(const Sample* p, const PdebConventions* c, double* rob_out, double* clause,
                    double* traj, double* vel, int64_t traj_cap_rows, int64_t* S_out, Workspace* ws) {
    int N = p->N, n = N; /* n free dofs: nodes 1..N */
    double h = p->L / N;
    int64_t S = step_count(p->T, p->dt, c->steps_rule);
    if (S_out) *S_out = S;
    if (S > PDEB_MAX_STEPS) return PDEB_ST_BADPARAM;
    double dt = p->dt;
    int lumped = (p->kind == PDEB_KIND_HEAT) ? c->mass_heat : c->mass_wave;

    double* buf = ws_get(ws, (size_t)(N + 1) * 18);
    if (!buf) return PDEB_ST_BADPARAM;
    memset(buf, 0, sizeof(double) * (size_t)(N + 1) * 18);
    double *x = buf, *ke = x + (N + 1), *me = ke + (N + 1);
    double *Mlo = me + (N + 1), *Mdi = Mlo + (N + 1), *Mup = Mdi + (N + 1);
    double *Alo = Mup + (N + 1), *Adi = Alo + (N + 1), *Aup = Adi + (N + 1);
    double *cpA = Aup + (N + 1), *invA = cpA + (N + 1);
    double *d = invA + (N + 1), *v = d + (N + 1), *a = v + (N + 1), *w = a + (N + 1), *Kd = w + (N + 1);

    for (int i = 0; i <= N; ++i) x[i] = i * h;
    x[N] = p->L;
    for (int e = 0; e < N; ++e) {
        double xs = (c->material_sampling == 1) ? e * h : (c->material_sampling == 2) ? (e + 1) * h : (e + 0.5) * h;
        int outer = (xs < p->x_a) || (xs > p->x_b);
        ke[e] = (outer ? p->E_out : p->E_in) / h;
        me[e] = (outer ? p->rho_out : p->rho_in) * h;
    }
    /* bands over all nodes 0..N; band entry "lo[i]" couples i with i-1, "up[i]" couples i with i+1 */
    double coefK = (p->kind == PDEB_KIND_HEAT) ? c->theta * dt : c->newmark_beta * dt * dt;
    for (int i = 0; i <= N; ++i) {
        double ml = (i > 0) ? me[i - 1] : 0.0, mr = (i < N) ? me[i] : 0.0;
        double kl = (i > 0) ? ke[i - 1] : 0.0, kr = (i < N) ? ke[i] : 0.0;
        if (lumped) { Mdi[i] = 0.5 * (ml + mr); Mlo[i] = 0.0; Mup[i] = 0.0; }
        else        { Mdi[i] = (ml + mr) / 3.0; Mlo[i] = ml / 6.0; Mup[i] = mr / 6.0; }
        Adi[i] = Mdi[i] + coefK * (kl + kr);
        Alo[i] = Mlo[i] - coefK * kl;
        Aup[i] = Mup[i] - coefK * kr;
    }
    /* factor the free-dof blocks (rows/cols 1..N): pointers offset by one */
    double *cpM = Kd + (N + 1), *invM = cpM + (N + 1);
    tri_factor(n, Mlo + 1, Mdi + 1, Mup + 1, cpM + 1, invM + 1);
    tri_factor(n, Alo + 1, Adi + 1, Aup + 1, cpA + 1, invA + 1);

    /* window step ranges (C7) and temporal accumulators */
    int64_t klo[PDEB_MAX_PRED], khi[PDEB_MAX_PRED];
    double acc[PDEB_MAX_PRED];
    for (int j = 0; j < p->np; ++j) {
        double qlo = p->tlo[j] / dt, qhi = p->thi[j] / dt;
        int64_t a0, a1;
        if (c->window_rule == 1) { a0 = (int64_t)qlo; a1 = (int64_t)qhi; }
        else { a0 = (int64_t)ceil(qlo - PDEO_EPS); a1 = (int64_t)floor(qhi + PDEO_EPS); }
        if (a0 < 0) a0 = 0;
        if (a1 > S) a1 = S;
        if (a0 > a1) {
            int64_t k = (int64_t)floor(0.5 * (p->tlo[j] + p->thi[j]) / dt + 0.5);
            if (k < 0) k = 0;
            if (k > S) k = S;
            a0 = a1 = k;
        }
        klo[j] = a0; khi[j] = a1;
        acc[j] = (p->top[j] == PDEB_TOP_G) ? INFINITY : -INFINITY;
    }
    /* nothing after the last observed step can change the result: the evaluation stops there (the same work the
     * CUDA kernels do); a trajectory dump marches on to the horizon S */
    int64_t k_stop = 0;
    for (int j = 0; j < p->np; ++j) if (khi[j] > k_stop) k_stop = khi[j];
    if (traj || vel) k_stop = S;

    /* initial state */
    double d0 = (p->kind == PDEB_KIND_HEAT) ? p->g0 : 0.0;
    for (int i = 0; i <= N; ++i) { d[i] = d0; v[i] = 0.0; a[i] = 0.0; }
    d[0] = p->g0; /* Dirichlet value, constant in time (its velocity and acceleration are 0) */

    /* w = M^-1 (F(0) - K d) on the free dofs: heat -> v0, wave -> a0 */
    apply_K(N, ke, d, Kd);
    for (int i = 1; i <= N; ++i) w[i] = -Kd[i];
    w[N] += pwl(p, 0.0);
    tri_solve(n, Mlo + 1, cpM + 1, invM + 1, w + 1);
    if (p->kind == PDEB_KIND_HEAT) for (int i = 1; i <= N; ++i) v[i] = w[i];
    else                           for (int i = 1; i <= N; ++i) a[i] = w[i];

    int status = PDEB_OK;
    for (int64_t k = 0;; ++k) {
        /* observe step k */
        if (traj && k < traj_cap_rows) memcpy(traj + k * (N + 1), d, sizeof(double) * (N + 1));
        if (vel && k < traj_cap_rows) memcpy(vel + k * (N + 1), v, sizeof(double) * (N + 1));
        for (int j = 0; j < p->np; ++j) {
            if (k < klo[j] || k > khi[j]) continue;
            /* a state that is not finite at an observed step is an error of the sample (status NONFINITE), whatever
             * the min / max below make of NaN margins */
            for (int i = 1; i <= N; ++i) if (!isfinite(d[i])) status = PDEB_ST_NONFINITE;
            double r = spatial_margin(p, j, d, x, h, c);
            if (p->top[j] == PDEB_TOP_G) { if (r < acc[j] || isnan(r)) acc[j] = r; }
            else                         { if (r > acc[j] || isnan(r)) acc[j] = r; }
        }
        if (k >= k_stop) break;
        double t1 = (double)(k + 1) * dt;
        double f1 = pwl(p, t1);
        if (p->kind == PDEB_KIND_HEAT) {
            double th = c->theta;
            for (int i = 1; i <= N; ++i) d[i] = d[i] + (1.0 - th) * dt * v[i];
            apply_K(N, ke, d, Kd);
            for (int i = 1; i <= N; ++i) w[i] = -Kd[i];
            w[N] += f1;
            tri_solve(n, Alo + 1, cpA + 1, invA + 1, w + 1);
            for (int i = 1; i <= N; ++i) { v[i] = w[i]; d[i] = d[i] + th * dt * v[i]; }
        } else {
            double be = c->newmark_beta, ga = c->newmark_gamma;
            for (int i = 1; i <= N; ++i) {
                d[i] = d[i] + dt * v[i] + 0.5 * (1.0 - 2.0 * be) * dt * dt * a[i];
                v[i] = v[i] + (1.0 - ga) * dt * a[i];
            }
            apply_K(N, ke, d, Kd);
            for (int i = 1; i <= N; ++i) w[i] = -Kd[i];
            w[N] += f1;
            tri_solve(n, Alo + 1, cpA + 1, invA + 1, w + 1);
            for (int i = 1; i <= N; ++i) {
                a[i] = w[i];
                d[i] = d[i] + be * dt * dt * a[i];
                v[i] = v[i] + ga * dt * a[i];
            }
        }
    }

    /* boolean tree */
    double val[PDEB_MAX_TREE];
    for (int j = 0; j < p->nt; ++j) {
        if (p->top_op[j] == PDEB_TREE_LEAF) val[j] = acc[p->lhs[j]];
        else {
            double l = val[p->lhs[j]], r = val[p->rhs[j]];
            if (isnan(l) || isnan(r)) val[j] = NAN;
            else if (p->top_op[j] == PDEB_TREE_AND) val[j] = l < r ? l : r;
            else val[j] = l > r ? l : r;
        }
    }
    double rob = val[p->nt - 1];
    if (!isfinite(rob)) status = PDEB_ST_NONFINITE;
    for (int j = 0; j < p->np; ++j) if (!isfinite(acc[j])) status = PDEB_ST_NONFINITE;
    if (clause) for (int j = 0; j < p->np; ++j) clause[j] = acc[j];
    *rob_out = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
    return status;
}

/* ---- exported ------------------------------------------------------------ */

int pdeo_eval_batch(const PdebBatch* in, const PdebConventions* conv, double* robustness_out,
                    double* clause_out, int32_t* status_out, int nthreads) {
    PdebConventions c;
    if (conv) c = *conv; else pdeo_default_conventions(&c);
    if (!in || !robustness_out || in->n < 0 || in->stride < in->n) return PDEB_E_ARG;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
    int64_t n = in->n;
#pragma omp parallel
    {
    Workspace ws = { NULL, 0 };
#pragma omp for schedule(dynamic, 64)
    for (int64_t s = 0; s < n; ++s) {
        Sample p;
        double cl[PDEB_MAX_PRED];
        double r = PDEB_FAILED_ROBUSTNESS;
        int st;
        for (int j = 0; j < PDEB_MAX_PRED; ++j) cl[j] = PDEB_FAILED_ROBUSTNESS;
        if (unpack(in, s, &p) != 0) st = PDEB_ST_BADPARAM;
        else st = eval_one(&p, &c, &r, cl, NULL, NULL, 0, NULL, &ws);
        robustness_out[s] = r;
        if (status_out) status_out[s] = st;
        if (clause_out) for (int j = 0; j < in->pmax; ++j) clause_out[j * in->stride + s] = (j < PDEB_MAX_PRED) ? cl[j] : PDEB_FAILED_ROBUSTNESS;
    }
    free(ws.buf);
    }
    return PDEB_E_OK;
}

/* number of steps of sample s, or -1 */
int64_t pdeo_steps(const PdebBatch* in, int64_t s, const PdebConventions* conv) {
    PdebConventions c;
    if (conv) c = *conv; else pdeo_default_conventions(&c);
    Sample p;
    if (unpack(in, s, &p) != 0) return -1;
    return step_count(p.T, p.dt, c.steps_rule);
}

/* full trajectory of sample s: traj[(S+1)][(N+1)] (and velocity for waves if vel != NULL) */
int pdeo_trajectory(const PdebBatch* in, int64_t s, const PdebConventions* conv, double* traj, double* vel,
                    int64_t cap_rows, double* robustness_out, double* clause_out) {
    PdebConventions c;
    if (conv) c = *conv; else pdeo_default_conventions(&c);
    Sample p;
    if (unpack(in, s, &p) != 0) return PDEB_ST_BADPARAM;
    double r, cl[PDEB_MAX_PRED];
    Workspace ws = { NULL, 0 };
    int st = eval_one(&p, &c, &r, cl, traj, vel, cap_rows, NULL, &ws);
    free(ws.buf);
    if (robustness_out) *robustness_out = r;
    if (clause_out) for (int j = 0; j < p.np; ++j) clause_out[j] = cl[j];
    return st;
}

int pdeo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
