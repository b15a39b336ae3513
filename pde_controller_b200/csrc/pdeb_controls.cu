/*
 * pdeb_controls.cu — "controls family": ONE specification evaluated against MANY candidate controls
 * (north star: "many candidate controls per spec").  In the reference the boundary input is
 * chosen by the Gurobi MILP inside control.femformal (call site eval_robustness_wrapper.py:16); that synthesis is
 * out of scope, and this family is the sampling surrogate for it: the robustness of 10^3..10^6 piecewise-linear
 * inputs for one program, plus the arg-max, in one call.
 *
 * Because mesh, materials, windows and intervals are the same for every thread,
 *   - the scaled step-matrix factors, R~ diagonal, node scales and clause data are computed ONCE on the host (the
 *     same PDEB_HD set-up code the batch kernels run per thread) and passed in the kernel-parameter bank: FP64
 *     instructions take them as constant-bank operands, so the time loop has no table loads at all;
 *   - every branch on windows, intervals and relations is warp-uniform: clause passes only run while a window is
 *     open and only over the groups of four nodes their interval touches;
 *   - a thread keeps nothing but its state in registers, so many more warps are resident than in the batch kernels.
 * One thread = one candidate control.  Knot TIMES are shared (C11: uniform knots), knot VALUES are per candidate,
 * plane-major [knot][candidate] like every list of the batch layout.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pdeb200.h"
#include "pdeb_core.h"
#include "pdeb_launch.h"

#define PDEB_CTL_THREADS 128
/* resident blocks per SM the kernels are compiled for (measured on B200, 1M candidates: heat 411 -> 466 M evals/s from
 * 1 to 5 blocks of 96 registers despite a few spills; waves carry three state vectors and peak at 3 blocks of 168) */
#ifndef PDEB_CTL_MINBLOCKS_HEAT
#define PDEB_CTL_MINBLOCKS_HEAT 5
#endif
#ifndef PDEB_CTL_MINBLOCKS_WAVE
#define PDEB_CTL_MINBLOCKS_WAVE 3
#endif
#ifndef PDEB_CTL_MINBLOCKS_WAVE_MOM
#define PDEB_CTL_MINBLOCKS_WAVE_MOM 3      /* momentum form of the wave step: two state vectors */
#endif
#define PDEB_CTL_KMAX 32

template <int N>
struct PdebSpecConst {
    double rd[N + 1], rl[N + 1], dinv[N + 1], c[N + 1], s[N + 1];   /* index = node 1..N; rl: low part of R~_ii (waves) */
    double w0[N + 1], a0_base[N + 1], a0_unit[N + 1];             /* waves: scaled initial state, a0 = base + f0 unit */
    double mu[N + 1];                                             /* waves, momentum form: 2 m_i s_i^2 (lumped nodal mass m_i) */
    double knot_t[PDEB_CTL_KMAX];
    PdebClauseData cl[PDEB_MAX_PRED];
    double sN, dt, th_dt, eq_eps;
    double cw1, cw2, cw3, cw4, cw5;
    int32_t nk, np, nt, k_end;
    int32_t momentum, pad_;                                        /* waves: 1 = the step runs in momentum form (below) */
    uint8_t tree_op[PDEB_MAX_TREE];
    int8_t tree_lhs[PDEB_MAX_TREE], tree_rhs[PDEB_MAX_TREE];
};

struct PdebCtlOut {
    double* robustness;   /* [m] */
    double* clause;       /* [pmax][stride] or NULL */
    int32_t* status;      /* [m] or NULL */
};

/* masked max over four consecutive nodes starting at node B (uniform mask, per-thread values) */
template <int N, int B>
__device__ __forceinline__ void pdeb_ctl_group(const PdebSpecConst<N>& sp, const double* w, double mh, double c0, uint32_t nmask,
                                               uint32_t and_mask, uint32_t xor_mask, double* hi) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        constexpr int dummy = 0; (void)dummy;
        const int i = B + q;
        if (i <= N) {
            const double qv = fma(-mh, (double)i, -c0);
            const double d = (i == 0) ? qv : fma(sp.s[i == 0 ? 1 : i], w[i], qv);
            const double g = pdeb_signop(d, and_mask, xor_mask);
            /* B + q is a compile-time constant after unrolling; the template form of the masked max needs it as one */
            switch (q) {
            case 0: pdeb_masked_max<(B + 0) & 31>(hi[0], g, nmask); break;
            case 1: pdeb_masked_max<(B + 1) & 31>(hi[1], g, nmask); break;
            case 2: pdeb_masked_max<(B + 2) & 31>(hi[2], g, nmask); break;
            default: pdeb_masked_max<(B + 3) & 31>(hi[3], g, nmask); break;
            }
        }
    }
}

template <int N, int B>
struct PdebCtlGroups {
    static __device__ __forceinline__ void run(const PdebSpecConst<N>& sp, const double* w, double mh, double c0, uint32_t nmask,
                                               uint32_t and_mask, uint32_t xor_mask, double* hi) {
        if (nmask & (0xFu << B)) pdeb_ctl_group<N, B>(sp, w, mh, c0, nmask, and_mask, xor_mask, hi);   /* warp-uniform */
        PdebCtlGroups<N, B + 4>::run(sp, w, mh, c0, nmask, and_mask, xor_mask, hi);
    }
};
template <int N>
struct PdebCtlGroups<N, 32> {
    static __device__ __forceinline__ void run(const PdebSpecConst<N>&, const double*, double, double, uint32_t, uint32_t, uint32_t,
                                               double*) {}
};

/* z[1..N] <- A~^-1 (R~ x + fNs e_N): the twisted solve of pdeb_core.h with every factor a constant-bank operand */
template <int KIND, int N, bool ADD = false>
__device__ __forceinline__ void pdeb_ctl_solve(const PdebSpecConst<N>& sp, const double* x, double fNs, double* z,
                                               const double* add = nullptr) {
    constexpr int M = (N + 1) / 2;
    double zt = 0, zb = 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int it = 1 + j, ib = N - j;
        if (it < M) {
            double r = x[it - 1] + x[it + 1];
            if (ADD) r += add[it];
            if (KIND == PDEB_KIND_WAVE) r = fma(sp.rl[it], x[it], r);
            r = fma(sp.rd[it], x[it], r);
            if (it > 1) r = fma(-sp.c[it - 1], zt, r);
            zt = r;
            z[it] = r * sp.dinv[it];
        }
        if (ib > M) {
            double r = x[ib - 1] + ((ib == N) ? fNs : x[ib + 1]);
            if (ADD) r += add[ib];
            if (KIND == PDEB_KIND_WAVE) r = fma(sp.rl[ib], x[ib], r);
            r = fma(sp.rd[ib], x[ib], r);
            if (ib < N) r = fma(-sp.c[ib], zb, r);
            zb = r;
            z[ib] = r * sp.dinv[ib];
        }
    }
    {
        double r = x[M - 1] + x[M + 1];
        if (ADD) r += add[M];
        if (KIND == PDEB_KIND_WAVE) r = fma(sp.rl[M], x[M], r);
        r = fma(sp.rd[M], x[M], r);
        r = fma(-sp.c[M - 1], zt, r);
        r = fma(-sp.c[M], zb, r);
        z[M] = r * sp.dinv[M];
    }
#pragma unroll
    for (int j = 1; j < N; ++j) {
        const int it = M - j, ib = M + j;
        if (it >= 1) z[it] = fma(-sp.c[it], z[it + 1], z[it]);
        if (ib <= N) z[ib] = fma(-sp.c[ib - 1], z[ib - 1], z[ib]);
    }
}

/*
 * MOMENTUM FORM of the wave step (MOM; Newmark beta = 1/4, gamma = 1/2 with lumped mass, the defaults C2 / C5).
 * With c = beta dt^2, M a_k = f_k - K d_k at every step (true at k = 0 by the definition of a0) turns the a-form into
 *     (M + c K) delta = -2c K d + p + c (f + f') e_N,    d' = d + delta,    p' = 2 M delta - p,      p = dt M v, p_0 = 0,
 * i.e. two state vectors instead of three and, in the scaled unknowns of pdeb_core.h (same S, same factors of
 * A~ = S (M + c K) S, same R~ = S (-K) S), with eta = delta~ / 2c and pi = S p / 2c:
 *     eta = A~^-1 (R~ x~ + pi + s_N (f + f')/2 e_N),    x~' = x~ + 2c eta,    pi' = mu eta - pi,    mu_i = 2 m_i s_i^2
 * 9 FP64 operations per row and step instead of 11; the right-hand side is still made of increments (K d through the
 * consistent double-double diagonal, the momentum).  Measured deviation from the oracle: 1.0e-10 of the parity metric
 * against 2.1e-11 for the a-form on the same inputs (tolerance 1e-9).
 */
/* one candidate: thread t of the launch (t >= m: a padding lane that walks along and writes nothing).  `ky` is read
 * with plain loads: in the one-launch search below the thread that reads a value has written it in the same kernel. */
template <int KIND, int N, int PMAX, bool MOM>
__device__ __forceinline__ void pdeb_controls_body(const PdebSpecConst<N>& sp, const double* ky, int64_t stride, int64_t m,
                                                   const PdebCtlOut& out, const int64_t t) {
    const bool live = t < m;
    const double* y = ky + (live ? t : 0);

    double w[N + 2];
    double v[KIND == PDEB_KIND_WAVE ? N + 2 : 1];            /* velocity; momentum form: pi */
    double a[(KIND == PDEB_KIND_WAVE && !MOM) ? N + 2 : 1];
    double acc[PMAX];
#pragma unroll
    for (int i = 0; i <= N + 1; ++i) w[i] = 0;
#pragma unroll
    for (int j = 0; j < PMAX; ++j)
        acc[j] = (j < sp.np && (((sp.cl[j].info >> 2) & 1u) == PDEB_TOP_F)) ? -INFINITY : INFINITY;

    /* piecewise-linear input: knot times are uniform, values per thread; same arithmetic as PdebPwl */
    const int nk = sp.nk;
    int jn = 0;                        /* index of the knot at t1 */
    double t0 = 0, y0 = 0, slope = 0, y1 = 0, t1 = INFINITY;
    if (nk >= 1) { y0 = y1 = y[0]; t1 = sp.knot_t[0]; }
    auto input = [&](double tt) {
        while (tt >= t1) {             /* warp-uniform */
            t0 = t1; y0 = y1;
            if (jn + 1 < nk) {
                const double tn = sp.knot_t[jn + 1], yn = y[(int64_t)(jn + 1) * stride];
                slope = (yn - y1) * pdeb_rcp(tn - t1);
                t1 = tn; y1 = yn;
                ++jn;
            } else { slope = 0; t1 = INFINITY; }
        }
        return fma(slope, tt - t0, y0);
    };
    double f_prev = input(0.0);
    if (KIND == PDEB_KIND_WAVE) {
#pragma unroll
        for (int i = 0; i <= N + 1; ++i) { v[i] = 0; if (!MOM) a[i] = 0; }
#pragma unroll
        for (int i = 1; i <= N; ++i) { w[i] = sp.w0[i]; if (!MOM) a[i] = fma(f_prev, sp.a0_unit[i], sp.a0_base[i]); }
    }

    const int32_t k_end = sp.k_end;
    bool bad_state = false;      /* sticky: state not finite at an observed step (see pdeb_run_reg) */
    for (int32_t k = 0;; ++k) {
        /* ---- clauses whose window contains step k (uniform) ---- */
#pragma unroll
        for (int j = 0; j < PMAX; ++j) {
            if (j < sp.np && k >= sp.cl[j].klo && k <= sp.cl[j].khi) {
                const PdebClauseData& cd = sp.cl[j];
                const uint32_t rel = cd.info & 3u;
                const uint32_t and_mask = (rel == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
                const uint32_t xor_mask = (rel == PDEB_REL_GT) ? 0x80000000u : 0u;
                double hi[4] = { -INFINITY, -INFINITY, -INFINITY, -INFINITY };
                PdebCtlGroups<N, 0>::run(sp, w, cd.mh, cd.c0, cd.nmask, and_mask, xor_mask, hi);
                const double h01 = hi[0] > hi[1] ? hi[0] : hi[1], h23 = hi[2] > hi[3] ? hi[2] : hi[3];
                double gmax = h01 > h23 ? h01 : h23;
                if (cd.info & 8u) {
                    /* interpolated end points (C8): uniform elements, so the two node values are picked by a uniform scan */
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = (int)((cd.info >> (4 + 5 * q)) & 31u);
                        double we = 0, wf = 0, se = 0, sf = 0;
#pragma unroll
                        for (int i = 0; i <= N; ++i) {
                            if (i == e) { we = (i == 0) ? 0.0 : w[i]; se = (i == 0) ? 0.0 : sp.s[i == 0 ? 1 : i]; }
                            if (i == e + 1) { wf = w[i]; sf = sp.s[i == 0 ? 1 : i]; }
                        }
                        const double de = fma(se, we, fma(-cd.mh, (double)e, -cd.c0));
                        const double df = fma(sf, wf, fma(-cd.mh, (double)(e + 1), -cd.c0));
                        const double gq = pdeb_signop(fma(cd.xi[q], df - de, de), and_mask, xor_mask);
                        if (gq > gmax) gmax = gq;
                    }
                }
                const double r = (rel == PDEB_REL_EQ) ? sp.eq_eps - gmax : -gmax;
                const bool take = (((cd.info >> 2) & 1u) == PDEB_TOP_G) ? (r < acc[j]) : (r > acc[j]);
                if (take || r != r) acc[j] = r;
                if (!(w[(N + 1) / 2] - w[(N + 1) / 2] == 0.0)) bad_state = true;
            }
        }
        if (k >= k_end) break;
        /* ---- advance to step k+1 ---- */
        const double f_next = input((double)(k + 1) * sp.dt);
        if (KIND == PDEB_KIND_HEAT) {
            const double fb = fma(sp.dt - sp.th_dt, f_prev, sp.th_dt * f_next);
            double z[N + 2];
            z[0] = 0; z[N + 1] = 0;
            pdeb_ctl_solve<KIND, N>(sp, w, sp.sN * fb, z);
#pragma unroll
            for (int i = 1; i <= N; ++i) w[i] = z[i];
        } else if (MOM) {
            double z[N + 2];
            z[0] = 0; z[N + 1] = 0;
            pdeb_ctl_solve<KIND, N, true>(sp, w, sp.sN * (0.5 * (f_prev + f_next)), z, v);
            const double c2 = 2.0 * sp.cw4;
#pragma unroll
            for (int i = 1; i <= N; ++i) {
                w[i] = fma(c2, z[i], w[i]);
                v[i] = fma(sp.mu[i], z[i], -v[i]);
            }
        } else {
#pragma unroll
            for (int i = 1; i <= N; ++i) {
                w[i] = fma(sp.cw2, a[i], fma(sp.cw1, v[i], w[i]));
                v[i] = fma(sp.cw3, a[i], v[i]);
            }
            pdeb_ctl_solve<KIND, N>(sp, w, sp.sN * f_next, a);
#pragma unroll
            for (int i = 1; i <= N; ++i) {
                w[i] = fma(sp.cw4, a[i], w[i]);
                v[i] = fma(sp.cw5, a[i], v[i]);
            }
        }
        f_prev = f_next;
    }

    if (!live) return;
    /* ---- formula tree ---- */
    double val[PDEB_MAX_TREE];
    int status = PDEB_OK;
#pragma unroll
    for (int q = 0; q < 2 * PMAX - 1; ++q) {
        if (q < sp.nt) {
            const int op = sp.tree_op[q], lh = sp.tree_lhs[q], rh = sp.tree_rhs[q];
            double x = NAN;
            if (op == PDEB_TREE_LEAF) {
#pragma unroll
                for (int j = 0; j < PMAX; ++j) if (j == lh) x = acc[j];
            } else {
                double a0 = NAN, a1 = NAN;
#pragma unroll
                for (int u = 0; u < 2 * PMAX - 1; ++u) { if (u == lh && u < q) a0 = val[u]; if (u == rh && u < q) a1 = val[u]; }
                if (a0 != a0 || a1 != a1) x = NAN;
                else if (op == PDEB_TREE_AND) x = a0 < a1 ? a0 : a1;
                else x = a0 > a1 ? a0 : a1;
            }
            val[q] = x;
        } else val[q] = NAN;
    }
    double rob = NAN;
#pragma unroll
    for (int q = 0; q < 2 * PMAX - 1; ++q) if (q == sp.nt - 1) rob = val[q];
    if (!isfinite(rob) || bad_state) status = PDEB_ST_NONFINITE;
#pragma unroll
    for (int j = 0; j < PMAX; ++j) if (j < sp.np && !isfinite(acc[j])) status = PDEB_ST_NONFINITE;
    out.robustness[t] = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
    if (out.status) out.status[t] = status;
    if (out.clause) {
#pragma unroll
        for (int j = 0; j < PMAX; ++j)
            if (j < sp.np) out.clause[(int64_t)j * stride + t] = (status == PDEB_OK) ? acc[j] : PDEB_FAILED_ROBUSTNESS;
    }
}

#define PDEB_CTL_BOUNDS(KIND, MOM) __launch_bounds__(PDEB_CTL_THREADS, ((KIND) == PDEB_KIND_HEAT) ? PDEB_CTL_MINBLOCKS_HEAT \
                                                     : ((MOM) ? PDEB_CTL_MINBLOCKS_WAVE_MOM : PDEB_CTL_MINBLOCKS_WAVE))
template <int KIND, int N, int PMAX, bool MOM = false>
__global__ void PDEB_CTL_BOUNDS(KIND, MOM)
pdeb_controls_kernel(const __grid_constant__ PdebSpecConst<N> sp, const double* __restrict__ ky, int64_t stride, int64_t m,
                     const __grid_constant__ PdebCtlOut out) {
    pdeb_controls_body<KIND, N, PMAX, MOM>(sp, ky, stride, m, out, (int64_t)blockIdx.x * PDEB_CTL_THREADS + threadIdx.x);
}

/* ---- arg-max of the robustness over the candidates (two tiny launches) ---- */
__global__ void pdeb_argmax_partial_kernel(const double* __restrict__ r, const int32_t* __restrict__ status, int64_t m,
                                           double* __restrict__ pv, int64_t* __restrict__ pi) {
    double bv = -INFINITY;
    int64_t bi = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        const bool ok = status ? (status[i] == PDEB_OK) : true;
        const double x = r[i];
        if (ok && (bi < 0 || x > bv)) { bv = x; bi = i; }      /* ascending i: the first of equal values wins */
    }
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o >= 1; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double ov = sv[threadIdx.x + o]; const int64_t oi = si[threadIdx.x + o];
            const double mv = sv[threadIdx.x]; const int64_t mi = si[threadIdx.x];
            if (oi >= 0 && (mi < 0 || ov > mv || (ov == mv && oi < mi))) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { pv[blockIdx.x] = sv[0]; pi[blockIdx.x] = si[0]; }
}


/* ---- device-resident cross-entropy search over the candidates of one specification ------------------------------
 * The host-driven search (pde_controller_b200/search.py, round 1) drew every generation with numpy, uploaded it,
 * fetched every robustness and ranked them on the host: ~4 ms of host work per program around ~0.7 ms of kernels.
 * Here a generation is ONE launch (pdeb_cem_gen_kernel, below) and nothing crosses PCIe until the search is over: every
 * thread draws its candidate  y = clip(mean + std * z)  from a counter-based generator (generation 0: uniform over the bounds,
 * plus the zero input and the two bounds; the incumbent is always candidate 0) and scores it with the body of the controls
 * kernel; the last block to finish selects the elite by a radix select (largest robustness, ties by lowest index), forms
 * mean / standard deviation of the elite's knot values, keeps the incumbent and the history.
 * Measured and removed on the way here: three launches per generation (draw / score / rank: 13 per search, 8.5e3 programs/s
 * at best and erratic), and a single COOPERATIVE launch for the whole search (grid-wide barriers between scoring and ranking:
 * 1.36 ms instead of 0.95 ms per search, and cooperative launches of different streams do not overlap: 2.8e3 programs/s).
 * Deterministic for a given seed (fixed reduction trees, no atomics on floating-point values). */
#ifndef PDEB_SEARCH_DEPTH
#define PDEB_SEARCH_DEPTH 3      /* searches in flight per host thread of pdeb_controls_search_batch */
#endif
#define PDEB_CEM_MAX_GEN 64
#define PDEB_CEM_MAX_ELITE 2048      /* elite candidates of a generation (list in shared memory) */
#define PDEB_CEM_KEY_CACHE 4096      /* populations up to this size are ranked out of shared memory */
struct PdebCemState {
    double mean[PDEB_CTL_KMAX], std[PDEB_CTL_KMAX], best_y[PDEB_CTL_KMAX];
    double history[PDEB_CEM_MAX_GEN];
    double best_r;
    int32_t has_best, n_valid_last;
};

__host__ __device__ __forceinline__ uint64_t pdeb_mix64(uint64_t z) {     /* splitmix64 finaliser */
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* uniform in (0, 1): one 64-bit word per (seed, generation, knot, candidate, which) */
__device__ __forceinline__ double pdeb_cem_uniform(uint64_t seed, int gen, int j, int64_t i, int which) {
    const uint64_t k = pdeb_mix64(pdeb_mix64(seed ^ pdeb_mix64(((uint64_t)gen << 32) | (uint64_t)(uint32_t)j)) + 2ull * (uint64_t)i + (uint64_t)which);
    return ((double)(k >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* the nk knot values of candidate i of generation `gen` */
__device__ __forceinline__ void pdeb_cem_draw(double* ky, int64_t stride, int nk, int64_t i, const PdebCemState* st, int gen,
                                              uint64_t seed, double lo, double hi) {
    /* the state was written by the last block of the PREVIOUS launch of this handle: read past the L1 (blocks of other
     * searches keep this SM busy across the launch boundary) */
    const bool has_best = gen > 0 && __ldcg(&st->has_best) != 0;
    for (int j = 0; j < nk; ++j) {
        double y;
        if (gen == 0) {
            y = lo + (hi - lo) * pdeb_cem_uniform(seed, gen, j, i, 0);
            if (i == 1) y = fmin(fmax(0.0, lo), hi);       /* the zero input competes from the start */
            if (i == 2) y = lo;
            if (i == 3) y = hi;
        } else {
            const double u1 = pdeb_cem_uniform(seed, gen, j, i, 0), u2 = pdeb_cem_uniform(seed, gen, j, i, 1);
            const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
            y = fmin(fmax(fma(__ldcg(&st->std[j]), z, __ldcg(&st->mean[j])), lo), hi);
        }
        if (i == 0 && has_best) y = __ldcg(&st->best_y[j]);    /* the incumbent stays in the population */
        ky[(int64_t)j * stride + i] = y;
    }
}

/* larger robustness <-> larger key; 0 is reserved for candidates that do not count (status != 0, non-finite) */
template <bool CG>
__device__ __forceinline__ uint64_t pdeb_cem_key(const double* rob, const int32_t* status, int64_t i) {
    const double r = CG ? __ldcg(rob + i) : rob[i];
    const int32_t st = CG ? __ldcg(status + i) : status[i];
    if (st != PDEB_OK || !isfinite(r)) return 0ull;
    const uint64_t u = (uint64_t)__double_as_longlong(r);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

/* ONE block of THREADS threads (a multiple of 32, at most 1024): elite, its statistics, the incumbent */
template <int THREADS, bool CG = false>
__device__ __forceinline__ void pdeb_cem_update_body(const double* rob, const int32_t* status, const double* ky, int64_t stride, int64_t P,
                                                     int nk, int n_elite, double std_floor, PdebCemState* st, int gen) {
    constexpr int WARPS = THREADS / 32;
    __shared__ unsigned hist[256];
    __shared__ int32_t elist[PDEB_CEM_MAX_ELITE];
    __shared__ uint64_t skeys[PDEB_CEM_KEY_CACHE];
    __shared__ uint64_t s_prefix;
    __shared__ unsigned long long s_best;
    __shared__ uint64_t s_bestkey;
    __shared__ int s_need, s_k, s_woff[32], s_eoff, s_goff, s_done;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    /* The keys are looked at ten times (count, eight radix passes, compaction).  One block reading them from the L2 each
     * time, one dependent load per loop iteration, was ~115 us of the ~150-280 us a generation took; a population of up to
     * PDEB_CEM_KEY_CACHE candidates is read ONCE, eight loads in flight per thread, into shared memory. */
    const bool cached = P <= PDEB_CEM_KEY_CACHE;
    if (cached) {
        for (int64_t base = tid; base < P; base += (int64_t)THREADS * 8) {
            uint64_t kk[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int64_t i = base + (int64_t)u * THREADS; kk[u] = (i < P) ? pdeb_cem_key<CG>(rob, status, i) : 0ull; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int64_t i = base + (int64_t)u * THREADS; if (i < P) skeys[i] = kk[u]; }
        }
        __syncthreads();
    }
    auto key_at = [&](int64_t i) -> uint64_t { return cached ? skeys[i] : pdeb_cem_key<CG>(rob, status, i); };

    /* ---- how many candidates count, and the best of them (largest key, lowest index) ---- */
    for (int q = tid; q < 256; q += THREADS) hist[q] = 0;
    if (tid == 0) { s_bestkey = 0ull; s_best = ~0ull; }
    __syncthreads();
    unsigned nv_local = 0;
    uint64_t bk = 0ull; int64_t bi = -1;
    for (int64_t i = tid; i < P; i += THREADS) {
        const uint64_t k = key_at(i);
        if (k) { ++nv_local; if (k > bk) { bk = k; bi = i; } }        /* ascending i per thread: first of equal keys */
    }
    atomicAdd(&hist[0], nv_local);
    atomicMax((unsigned long long*)&s_bestkey, (unsigned long long)bk);
    __syncthreads();
    const int nv = (int)hist[0];
    if (bk == s_bestkey && bi >= 0) atomicMin(&s_best, (unsigned long long)bi);
    __syncthreads();
    const int64_t best_i = (s_bestkey != 0ull) ? (int64_t)s_best : -1;
    const int k_el = nv < n_elite ? nv : n_elite;

    /* ---- radix select of the k_el-th largest key ---- */
    if (tid == 0) { s_prefix = 0ull; s_need = k_el; s_done = 0; }
    __syncthreads();
    if (k_el >= 2) {
        for (int b = 7; b >= 0; --b) {
            for (int q = tid; q < 256; q += THREADS) hist[q] = 0;
            __syncthreads();
            const uint64_t prefix = s_prefix;
            for (int64_t i = tid; i < P; i += THREADS) {
                const uint64_t k = key_at(i);
                if (k && (b == 7 || (k >> (8 * (b + 1))) == prefix)) atomicAdd(&hist[(unsigned)(k >> (8 * b)) & 255u], 1u);
            }
            __syncthreads();
            if (warp == 0) {
                /* the bin that holds the need-th largest key: lane l owns the bins 255 - 8 l .. 248 - 8 l (descending), a
                 * shuffle scan gives the keys above each lane's bins (thread 0 walking the bins alone was 2-3 us a pass) */
                unsigned own = 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) own += hist[255 - 8 * lane - q];
                unsigned above = own;
                for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, above, o); if (lane >= o) above += v; }
                above -= own;                                        /* keys in the bins of the lanes before this one */
                const int need = s_need;
                const bool mine = (int)above < need && (int)(above + own) >= need;
                if (mine) {
                    int rest = need - (int)above, bin = 255 - 8 * lane;
                    for (int q = 0; q < 8; ++q, --bin) { if ((int)hist[bin] >= rest) break; rest -= (int)hist[bin]; }
                    /* every key of that bin is wanted: the selection is over, the threshold is the lowest key of the bin */
                    const bool all = (int)hist[bin] == rest && b > 0;
                    s_prefix = all ? (((prefix << 8) | (uint64_t)bin) << (8 * b)) : ((prefix << 8) | (uint64_t)bin);
                    s_need = all ? 0x7fffffff : rest;              /* still to take among the keys equal to the threshold */
                    s_done = all ? 1 : 0;
                }
            }
            __syncthreads();
            if (s_done) break;                                       /* block-uniform */
        }
        /* ---- the elite in index order: every key above the threshold, and the first s_need of those equal to it ---- */
        const uint64_t thr = s_prefix;
        const int need_eq = s_need;
        if (tid == 0) { s_eoff = 0; s_goff = 0; }
        __syncthreads();
        for (int64_t base = 0; base < P; base += THREADS) {
            const int64_t i = base + tid;
            const uint64_t k = (i < P) ? key_at(i) : 0ull;
            const bool eq = k != 0ull && k == thr;
            const unsigned beq = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) s_woff[warp] = __popc(beq);
            __syncthreads();
            int eq_before = s_eoff;
            for (int w = 0; w < warp; ++w) eq_before += s_woff[w];
            eq_before += __popc(beq & ((1u << lane) - 1u));
            int eq_total = 0;
            for (int w = 0; w < WARPS; ++w) eq_total += s_woff[w];
            __syncthreads();
            const bool take = k != 0ull && (k > thr || (eq && eq_before < need_eq));
            const unsigned bt = __ballot_sync(0xffffffffu, take);
            if (lane == 0) s_woff[warp] = __popc(bt);
            __syncthreads();
            int pos = s_goff;
            for (int w = 0; w < warp; ++w) pos += s_woff[w];
            pos += __popc(bt & ((1u << lane) - 1u));
            if (take && pos < PDEB_CEM_MAX_ELITE) elist[pos] = (int32_t)i;
            int tk_total = 0;
            for (int w = 0; w < WARPS; ++w) tk_total += s_woff[w];
            __syncthreads();
            if (tid == 0) { s_eoff += eq_total; s_goff += tk_total; }
            __syncthreads();
        }
        if (tid == 0) s_k = s_goff < PDEB_CEM_MAX_ELITE ? s_goff : PDEB_CEM_MAX_ELITE;
        __syncthreads();
        /* ---- mean and standard deviation of the elite per knot: warp w takes the knots w, w + WARPS, .. (the order of
         * the sums does not depend on the block size: lanes stride over the elite, then a butterfly) ---- */
        const int ke = s_k;
        for (int j = warp; j < nk; j += WARPS) {
            const double* col = ky + (int64_t)j * stride;
            double sum = 0.0;
            for (int q = lane; q < ke; q += 32) sum += CG ? __ldcg(col + elist[q]) : col[elist[q]];
            for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const double mean = sum / (double)ke;
            double var = 0.0;
            for (int q = lane; q < ke; q += 32) { const double d = (CG ? __ldcg(col + elist[q]) : col[elist[q]]) - mean; var = fma(d, d, var); }
            for (int o = 16; o >= 1; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
            if (lane == 0) { st->mean[j] = mean; st->std[j] = fmax(sqrt(var / (double)ke), std_floor); }
        }
    }
    /* ---- incumbent ---- */
    __syncthreads();
    const double cand = best_i >= 0 ? (CG ? __ldcg(rob + best_i) : rob[best_i]) : 0.0;
    const bool better = best_i >= 0 && (gen == 0 || !st->has_best || cand > st->best_r);
    __syncthreads();
    if (better) {
        for (int j = tid; j < nk; j += THREADS) st->best_y[j] = CG ? __ldcg(ky + (int64_t)j * stride + best_i) : ky[(int64_t)j * stride + best_i];
        if (tid == 0) { st->best_r = cand; st->has_best = 1; }
    }
    __syncthreads();
    if (tid == 0) {
        if (gen == 0 && best_i < 0) { st->has_best = 0; st->best_r = -INFINITY; }
        st->history[gen] = st->has_best ? st->best_r : -INFINITY;
        st->n_valid_last = nv;
    }
}

/*
 * ONE launch per generation: every thread draws its own candidate from the distribution the previous launch left, scores
 * it, and the LAST block to finish (a ticket counter) ranks the generation and re-centres the distribution.  Six launches
 * per search instead of thirteen: a sweep of searches is bound by the rate at which a process can launch kernels
 * (~1e5 per second from 16 host threads), not by the device.  What the ranking block reads was written by other blocks
 * of the same launch: it is read past the L1 (__ldcg) behind a __threadfence on the writers' side.
 */
template <int KIND, int N, int PMAX, bool MOM = false>
__global__ void PDEB_CTL_BOUNDS(KIND, MOM)
pdeb_cem_gen_kernel(const __grid_constant__ PdebSpecConst<N> sp, double* ky, int64_t P, const __grid_constant__ PdebCtlOut out,
                    PdebCemState* st, unsigned* ticket, int gen, int n_elite, uint64_t seed, double lo, double hi, double std_floor) {
    const int64_t t = (int64_t)blockIdx.x * PDEB_CTL_THREADS + threadIdx.x;
    const int nk = sp.nk;
    if (gen == 0 && blockIdx.x == 0) {        /* nobody reads the state before the ranking at the end of this launch */
        if (threadIdx.x < PDEB_CTL_KMAX) { st->mean[threadIdx.x] = 0.5 * (lo + hi); st->std[threadIdx.x] = 0.5 * (hi - lo); st->best_y[threadIdx.x] = 0.0; }
        if (threadIdx.x < PDEB_CEM_MAX_GEN) st->history[threadIdx.x] = 0.0;
        if (threadIdx.x == 0) { st->best_r = -INFINITY; st->has_best = 0; st->n_valid_last = 0; }
    }
    if (t < P) pdeb_cem_draw(ky, P, nk, t, st, gen, seed, lo, hi);
    pdeb_controls_body<KIND, N, PMAX, MOM>(sp, ky, P, P, out, t);
    __shared__ int s_last;
    __threadfence();                     /* this thread's results, device-wide */
    __syncthreads();                     /* ... and every thread of the block has got that far */
    if (threadIdx.x == 0) {
        __threadfence();                 /* cumulative: what the barrier made visible to thread 0 is released with its ticket */
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        pdeb_cem_update_body<PDEB_CTL_THREADS, true>(out.robustness, out.status, ky, P, P, nk, n_elite, std_floor, st, gen);
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

/* ---- host side ------------------------------------------------------------------------------------------------ */
static int cfail(int code, const char* what, cudaError_t ce = cudaSuccess) {
    char buf[512];
    if (ce != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
    else snprintf(buf, sizeof buf, "%s", what);
    pdeb_set_last_error(buf);          /* read back through pdeb_last_error() */
    return code;
}
#define CCK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cfail(PDEB_E_CUDA, #call, _e); } while (0)

struct PdebControls {
    int device = 0, kind = 0, n_elem = 0, pmax = 0, nk = 0;
    int64_t capacity = 0;
    int spec_status = PDEB_OK;
    PdebSpecConst<30> heat{};
    PdebSpecConst<20> wave{};
    double* d_y = nullptr;        /* [nk][capacity] */
    double* d_rob = nullptr;      /* [capacity] */
    double* d_clause = nullptr;   /* [pmax][capacity] */
    int32_t* d_status = nullptr;  /* [capacity] */
    double* d_pv = nullptr;       /* arg-max partials */
    int64_t* d_pi = nullptr;
    double* h_pv = nullptr;       /* pinned */
    int64_t* h_pi = nullptr;
    struct PdebCemState* d_cem = nullptr;   /* state of the device-resident search (pdeb_controls_search) */
    unsigned* d_ticket = nullptr;           /* "last block ranks" counter of pdeb_cem_gen_kernel (zero between launches) */
    struct PdebCemState* h_cem = nullptr;   /* pinned mirror */
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr;   /* blocking-sync event (sweeps): a waiting host thread sleeps instead of spinning on the stream */
    bool sleep_on_done = false;
    int last_launches = 0;        /* kernels the last search_enqueue launched */
};
#define PDEB_ARGMAX_BLOCKS 128

/* the per-thread set-up of the batch kernels, run once on the host for the one specification */
template <int KIND, int N>
static int build_spec(const PdebBatch* in, int64_t s, const PdebConventions& cv, int nk, const double* knot_t, PdebSpecConst<N>& sp) {
    memset(&sp, 0, sizeof sp);
    const int64_t st = in->stride;
    const double L = in->L[s], dt = in->dt[s], T = in->T[s], g0 = in->g0[s];
    const double rho_in = in->rho_in[s], rho_out = in->rho_out[s], E_in = in->E_in[s], E_out = in->E_out[s];
    const int np = in->n_pred[s], nt = in->n_tree[s];
    if (!pdeb_params_ok(KIND, N, L, dt, T, g0, rho_in, rho_out, E_in, E_out)) return PDEB_ST_BADPARAM;
    if (np < 1 || np > in->pmax || np > PDEB_MAX_PRED || nt < 1 || nt > in->tmax || nt > PDEB_MAX_TREE) return PDEB_ST_BADPARAM;
    if (nk < 0 || nk > PDEB_CTL_KMAX) return PDEB_ST_BADPARAM;
    for (int j = 1; j < nk; ++j) if (!(knot_t[j] > knot_t[j - 1])) return PDEB_ST_BADPARAM;
    for (int j = 0; j < np; ++j)
        if (in->pred_rel[(int64_t)j * st + s] > 2 || in->pred_top[(int64_t)j * st + s] > 1) return PDEB_ST_BADPARAM;
    const int64_t S = pdeb_step_count(T, dt, cv.steps_rule);
    if (S > PDEB_MAX_STEPS) return PDEB_ST_BADPARAM;
    const double h = L / N;
    const int lumped = (KIND == PDEB_KIND_HEAT) ? cv.mass_heat : cv.mass_wave;
    const double cA = (KIND == PDEB_KIND_HEAT) ? cv.theta * dt : cv.newmark_beta * dt * dt;
    const double cR = (1.0 - cv.theta) * dt;

    PdebTrajReg<KIND, N, double, 1> tr;
    PdebD2 tab[N + 1];
    double stab[N + 1], ltab[N + 1];
#if PDEB_C_SMEM
    double ctab[N + 1];
    tr.ctab = ctab;
#endif
    memset(ltab, 0, sizeof ltab);
    tr.tab = tab; tr.stab = stab; tr.ltab = ltab; tr.sN = 0;
    for (int i = 0; i <= N + 1; ++i) tr.w[i] = 0;
    if (!tr.setup(L, in->x_a[s], in->x_b[s], rho_in, rho_out, E_in, E_out, lumped, cv.material_sampling, cA, cR)) return PDEB_ST_NONFINITE;
    for (int i = 1; i <= N; ++i) {
        sp.rd[i] = tab[i - 1].x; sp.dinv[i] = tab[i - 1].y; sp.s[i] = stab[i - 1]; sp.rl[i] = ltab[i - 1];
        sp.c[i] = (i < N) ? tr.getc(i) : 0.0;
    }
    if constexpr (KIND == PDEB_KIND_WAVE) {
        /* a0 is affine in the initial load f0: evaluate the set-up at f0 = 0 and f0 = 1 */
        tr.wave_init(h, g0, 0.0, rho_in, rho_out, E_in, E_out, lumped);
        for (int i = 1; i <= N; ++i) { sp.w0[i] = tr.w[i]; sp.a0_base[i] = tr.a[i]; }
        tr.wave_init(h, g0, 1.0, rho_in, rho_out, E_in, E_out, lumped);
        for (int i = 1; i <= N; ++i) sp.a0_unit[i] = tr.a[i] - sp.a0_base[i];
    }
    sp.momentum = 0;
    if constexpr (KIND == PDEB_KIND_WAVE) {
        /* PDEB_CONTROLS_AFORM=1 in the environment keeps the a-form (a diagnostic switch: compares the two forms on the
         * same conventions) */
        const char* aform = getenv("PDEB_CONTROLS_AFORM");
        sp.momentum = (lumped && cv.newmark_beta == 0.25 && cv.newmark_gamma == 0.5 && !(aform && aform[0] == '1')) ? 1 : 0;
        for (int i = 1; i <= N; ++i) {
            const double ml = (tr.outer(i - 1) ? rho_out : rho_in) * h, mr = (i < N) ? (tr.outer(i) ? rho_out : rho_in) * h : 0.0;
            sp.mu[i] = (ml + mr) * stab[i - 1] * stab[i - 1];       /* 2 * (ml + mr)/2 * s_i^2 */
        }
    }
    sp.sN = tr.sN; sp.dt = dt; sp.th_dt = cv.theta * dt; sp.eq_eps = (cv.eq_rule == 1) ? cv.eq_eps : 0.0;
    sp.cw1 = dt; sp.cw2 = 0.5 * (1.0 - 2.0 * cv.newmark_beta) * dt * dt; sp.cw3 = (1.0 - cv.newmark_gamma) * dt;
    sp.cw4 = cv.newmark_beta * dt * dt; sp.cw5 = cv.newmark_gamma * dt;
    sp.nk = nk; sp.np = np; sp.nt = nt;
    for (int j = 0; j < nk; ++j) sp.knot_t[j] = knot_t[j];
    int32_t k_end = 0;
    for (int j = 0; j < np; ++j) {
        sp.cl[j] = pdeb_clause_data<N>(in->pred_m[(int64_t)j * st + s], in->pred_b[(int64_t)j * st + s],
                                       in->pred_xlo[(int64_t)j * st + s], in->pred_xhi[(int64_t)j * st + s],
                                       in->pred_tlo[(int64_t)j * st + s], in->pred_thi[(int64_t)j * st + s],
                                       in->pred_rel[(int64_t)j * st + s], in->pred_top[(int64_t)j * st + s], L, h, dt, g0, S, cv);
        if (sp.cl[j].khi > k_end) k_end = sp.cl[j].khi;
    }
    sp.k_end = k_end;
    for (int q = 0; q < nt; ++q) {
        const int op = in->tree_op[(int64_t)q * st + s], lh = in->tree_lhs[(int64_t)q * st + s], rh = in->tree_rhs[(int64_t)q * st + s];
        if (op == PDEB_TREE_LEAF) { if (lh < 0 || lh >= np) return PDEB_ST_BADPARAM; }
        else if (op == PDEB_TREE_AND || op == PDEB_TREE_OR) { if (lh < 0 || lh >= q || rh < 0 || rh >= q) return PDEB_ST_BADPARAM; }
        else return PDEB_ST_BADPARAM;
        sp.tree_op[q] = (uint8_t)op; sp.tree_lhs[q] = (int8_t)lh; sp.tree_rhs[q] = (int8_t)rh;
    }
    return PDEB_OK;
}

extern "C" void pdeb_controls_destroy(PdebControls* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->d_y); cudaFree(h->d_rob); cudaFree(h->d_clause); cudaFree(h->d_status); cudaFree(h->d_pv); cudaFree(h->d_pi);
    cudaFree(h->d_cem); cudaFree(h->d_ticket);
    if (h->h_pv) cudaFreeHost(h->h_pv);          /* one page-locked block: { partial values | partial indices | search state } */
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->done) cudaEventDestroy(h->done);
    delete h;
}

/* (re)build the specification of a handle; the device buffers are sized for the family's maxima and stay */
static int controls_set_spec(PdebControls* h, const PdebBatch* spec, int64_t sample, int32_t n_knots, const double* knot_t,
                             const PdebConventions* conv) {
    if (!spec || sample < 0 || sample >= spec->n || n_knots < 0 || (n_knots > 0 && !knot_t))
        return cfail(PDEB_E_ARG, "controls: bad specification arguments");
    if (n_knots > PDEB_CTL_KMAX) return cfail(PDEB_E_ARG, "controls: more knots than PDEB_CTL_KMAX (32)");
    PdebConventions cv;
    if (conv) cv = *conv; else pdeb_default_conventions(&cv);
    const int kind = spec->kind[sample], N = spec->n_elem[sample];
    if (!((kind == PDEB_KIND_HEAT && N == 30) || (kind == PDEB_KIND_WAVE && N == 20)))
        return cfail(PDEB_E_ARG, "controls family covers the default meshes (heat N=30, wave N=20); use the batch path for other meshes");
    h->kind = kind; h->n_elem = N; h->nk = n_knots; h->pmax = spec->n_pred[sample];
    h->spec_status = (kind == PDEB_KIND_HEAT) ? build_spec<PDEB_KIND_HEAT, 30>(spec, sample, cv, n_knots, knot_t, h->heat)
                                              : build_spec<PDEB_KIND_WAVE, 20>(spec, sample, cv, n_knots, knot_t, h->wave);
    if (h->pmax < 1 || h->pmax > PDEB_MAX_PRED) h->pmax = 1;
    return PDEB_E_OK;
}

extern "C" int pdeb_controls_create(const PdebBatch* spec, int64_t sample, int64_t capacity, int32_t n_knots, const double* knot_t,
                                    int device, const PdebConventions* conv, PdebControls** out) {
    if (!spec || !out || sample < 0 || sample >= spec->n || capacity < 1 || (n_knots > 0 && !knot_t))
        return cfail(PDEB_E_ARG, "pdeb_controls_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return cfail(PDEB_E_NODEVICE, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) return cfail(PDEB_E_ARG, "bad device");
    PdebControls* h = new PdebControls;
    h->device = device; h->capacity = capacity;
    const int rc = controls_set_spec(h, spec, sample, n_knots, knot_t, conv);
    if (rc != PDEB_E_OK) { delete h; return rc; }
    /* sized for ANY specification of the family (PDEB_CTL_KMAX knots, PDEB_MAX_PRED clauses), so that
     * pdeb_controls_respec can point the handle at the next program without touching the allocator */
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_y, sizeof(double) * (size_t)PDEB_CTL_KMAX * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_rob, sizeof(double) * (size_t)capacity);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_clause, sizeof(double) * (size_t)PDEB_MAX_PRED * capacity);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_status, sizeof(int32_t) * (size_t)capacity);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_pv, sizeof(double) * PDEB_ARGMAX_BLOCKS);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_pi, sizeof(int64_t) * PDEB_ARGMAX_BLOCKS);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_cem, sizeof(PdebCemState));
    if (e == cudaSuccess) e = cudaMalloc(&h->d_ticket, sizeof(unsigned));
    /* the ticket must BE zero before the first search, which runs on a non-blocking stream of the caller's: a memset on
     * the default stream is not ordered against that stream, so it is waited for here (the kernel leaves it zero) */
    if (e == cudaSuccess) e = cudaMemset(h->d_ticket, 0, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    /* ONE page-locked allocation per handle (cudaHostAlloc is slow and serialised process-wide: three per handle made the
     * first sweep of a process, which creates 24 handles, run at a seventh of the steady rate) */
    if (e == cudaSuccess) {
        unsigned char* hp = nullptr;
        e = cudaHostAlloc(&hp, 16 * PDEB_ARGMAX_BLOCKS + sizeof(PdebCemState), cudaHostAllocDefault);
        if (e == cudaSuccess) {
            h->h_pv = (double*)hp;
            h->h_pi = (int64_t*)(hp + 8 * PDEB_ARGMAX_BLOCKS);
            h->h_cem = (PdebCemState*)(hp + 16 * PDEB_ARGMAX_BLOCKS);
        }
    }
    for (auto& ev : h->ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->done, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) { pdeb_controls_destroy(h); return cfail(PDEB_E_CUDA, "pdeb_controls_create", e); }
    *out = h;
    return PDEB_E_OK;
}

extern "C" int pdeb_controls_respec(PdebControls* h, const PdebBatch* spec, int64_t sample, int32_t n_knots, const double* knot_t,
                                    const PdebConventions* conv) {
    if (!h) return cfail(PDEB_E_ARG, "pdeb_controls_respec: NULL handle");
    return controls_set_spec(h, spec, sample, n_knots, knot_t, conv);
}

template <int KIND, int N>
static cudaError_t launch_controls(const PdebSpecConst<N>& sp, int pmax, const double* ky, int64_t stride, int64_t m,
                                   const PdebCtlOut& out, cudaStream_t stream) {
    const unsigned blocks = (unsigned)((m + PDEB_CTL_THREADS - 1) / PDEB_CTL_THREADS);
    if constexpr (KIND == PDEB_KIND_WAVE) {
        if (sp.momentum) {
            if (pmax <= 3) pdeb_controls_kernel<KIND, N, 3, true><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, stride, m, out);
            else pdeb_controls_kernel<KIND, N, PDEB_MAX_PRED, true><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, stride, m, out);
            return cudaGetLastError();
        }
    }
    if (pmax <= 3) pdeb_controls_kernel<KIND, N, 3><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, stride, m, out);
    else pdeb_controls_kernel<KIND, N, PDEB_MAX_PRED><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, stride, m, out);
    return cudaGetLastError();
}

template <int KIND, int N>
static cudaError_t launch_gen(const PdebSpecConst<N>& sp, int pmax, double* ky, int64_t P, const PdebCtlOut& out, PdebCemState* st,
                              unsigned* ticket, int gen, int n_elite, uint64_t seed, double lo, double hi, double std_floor,
                              cudaStream_t stream) {
    const unsigned blocks = (unsigned)((P + PDEB_CTL_THREADS - 1) / PDEB_CTL_THREADS);
    if constexpr (KIND == PDEB_KIND_WAVE) {
        if (sp.momentum) {
            if (pmax <= 3) pdeb_cem_gen_kernel<KIND, N, 3, true><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, P, out, st, ticket, gen, n_elite, seed, lo, hi, std_floor);
            else pdeb_cem_gen_kernel<KIND, N, PDEB_MAX_PRED, true><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, P, out, st, ticket, gen, n_elite, seed, lo, hi, std_floor);
            return cudaGetLastError();
        }
    }
    if (pmax <= 3) pdeb_cem_gen_kernel<KIND, N, 3><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, P, out, st, ticket, gen, n_elite, seed, lo, hi, std_floor);
    else pdeb_cem_gen_kernel<KIND, N, PDEB_MAX_PRED><<<blocks, PDEB_CTL_THREADS, 0, stream>>>(sp, ky, P, out, st, ticket, gen, n_elite, seed, lo, hi, std_floor);
    return cudaGetLastError();
}

/*
 * Evaluate m <= capacity candidate controls.  knot_y: host [n_knots][m] plane-major, or NULL to re-evaluate the values
 * already on the device.  Outputs (all optional): robustness [m], clause [pmax][m], status [m], best index / value
 * (largest robustness among the candidates with status 0; index -1 if none).  Synchronises `stream` before returning.
 */
extern "C" int pdeb_controls_eval(PdebControls* h, int64_t m, const double* knot_y, double* robustness_out, double* clause_out,
                                  int32_t* status_out, int64_t* best_index, double* best_value, void* stream_, PdebTiming* timing) {
    if (!h || m < 1 || m > h->capacity) return cfail(PDEB_E_ARG, "pdeb_controls_eval: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    CCK(cudaSetDevice(h->device));
    if (h->spec_status != PDEB_OK) {
        /* the specification itself is unusable: every candidate fails the same way (never raises per sample) */
        for (int64_t i = 0; i < m; ++i) {
            if (robustness_out) robustness_out[i] = PDEB_FAILED_ROBUSTNESS;
            if (status_out) status_out[i] = h->spec_status;
        }
        if (clause_out) for (int64_t i = 0; i < (int64_t)h->pmax * m; ++i) clause_out[i] = PDEB_FAILED_ROBUSTNESS;
        if (best_index) *best_index = -1;
        if (best_value) *best_value = PDEB_FAILED_ROBUSTNESS;
        if (timing) memset(timing, 0, sizeof *timing);
        return PDEB_E_OK;
    }
    if (knot_y && h->nk > 0) CCK(cudaMemcpyAsync(h->d_y, knot_y, sizeof(double) * (size_t)h->nk * m, cudaMemcpyHostToDevice, stream));
    PdebCtlOut out{h->d_rob, h->d_clause, h->d_status};
    CCK(cudaEventRecord(h->ev[0], stream));
    cudaError_t e = (h->kind == PDEB_KIND_HEAT) ? launch_controls<PDEB_KIND_HEAT, 30>(h->heat, h->pmax, h->d_y, m, m, out, stream)
                                                : launch_controls<PDEB_KIND_WAVE, 20>(h->wave, h->pmax, h->d_y, m, m, out, stream);
    if (e != cudaSuccess) return cfail(PDEB_E_CUDA, "controls kernel launch", e);
    int launches = 1;
    if (best_index || best_value) {
        pdeb_argmax_partial_kernel<<<PDEB_ARGMAX_BLOCKS, 256, 0, stream>>>(h->d_rob, h->d_status, m, h->d_pv, h->d_pi);
        CCK(cudaGetLastError());
        ++launches;
    }
    CCK(cudaEventRecord(h->ev[1], stream));
    if (best_index || best_value) {
        CCK(cudaMemcpyAsync(h->h_pv, h->d_pv, sizeof(double) * PDEB_ARGMAX_BLOCKS, cudaMemcpyDeviceToHost, stream));
        CCK(cudaMemcpyAsync(h->h_pi, h->d_pi, sizeof(int64_t) * PDEB_ARGMAX_BLOCKS, cudaMemcpyDeviceToHost, stream));
    }
    if (robustness_out) CCK(cudaMemcpyAsync(robustness_out, h->d_rob, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, stream));
    if (status_out) CCK(cudaMemcpyAsync(status_out, h->d_status, sizeof(int32_t) * (size_t)m, cudaMemcpyDeviceToHost, stream));
    if (clause_out) CCK(cudaMemcpyAsync(clause_out, h->d_clause, sizeof(double) * (size_t)h->pmax * m, cudaMemcpyDeviceToHost, stream));
    CCK(cudaStreamSynchronize(stream));
    if (best_index || best_value) {
        double bv = -INFINITY; int64_t bi = -1;
        for (int q = 0; q < PDEB_ARGMAX_BLOCKS; ++q)
            if (h->h_pi[q] >= 0 && (bi < 0 || h->h_pv[q] > bv || (h->h_pv[q] == bv && h->h_pi[q] < bi))) { bv = h->h_pv[q]; bi = h->h_pi[q]; }
        if (best_index) *best_index = bi;
        if (best_value) *best_value = (bi >= 0) ? bv : PDEB_FAILED_ROBUSTNESS;
    }
    if (timing) {
        memset(timing, 0, sizeof *timing);
        float ms = 0;
        CCK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        timing->kernel_ms = ms; timing->total_ms = ms; timing->launches = launches;
    }
    return PDEB_E_OK;
}

/* the candidates currently on the device (the last generation of a search, or what pdeb_controls_eval uploaded) */
extern "C" int pdeb_controls_fetch_knots(PdebControls* h, int64_t m, double* knot_y_out, void* stream_) {
    if (!h || !knot_y_out || m < 1 || m > h->capacity) return cfail(PDEB_E_ARG, "pdeb_controls_fetch_knots: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    CCK(cudaSetDevice(h->device));
    if (h->nk > 0) CCK(cudaMemcpyAsync(knot_y_out, h->d_y, sizeof(double) * (size_t)h->nk * m, cudaMemcpyDeviceToHost, stream));
    CCK(cudaStreamSynchronize(stream));
    return PDEB_E_OK;
}

/* the first half of a search: everything enqueued on `stream`, nothing waited for.  -1 in *skipped: the specification
 * is unusable and nothing was enqueued */
static int search_enqueue(PdebControls* h, int64_t population, int32_t generations, double elite_frac, double lo, double hi,
                          uint64_t seed, cudaStream_t stream, bool timing, bool* skipped) {
    *skipped = false;
    if (!h || population < 4 || population > h->capacity || generations < 1 || generations > PDEB_CEM_MAX_GEN || !(hi >= lo) ||
        !(elite_frac > 0.0 && elite_frac <= 1.0))
        return cfail(PDEB_E_ARG, "pdeb_controls_search: bad arguments (population 4..capacity, generations 1..64, lo <= hi, 0 < elite_frac <= 1)");
    int64_t n_elite = (int64_t)llround(elite_frac * (double)population);
    if (n_elite < 2) n_elite = 2;
    if (n_elite > PDEB_CEM_MAX_ELITE) return cfail(PDEB_E_ARG, "pdeb_controls_search: more than 2048 elite candidates");
    CCK(cudaSetDevice(h->device));
    if (h->spec_status != PDEB_OK) { *skipped = true; return PDEB_E_OK; }     /* nothing to search (never raises per sample) */
    const int nk = h->nk;
    PdebCtlOut out{h->d_rob, h->d_clause, h->d_status};
    const double std_floor = 1e-3 * (hi - lo);
    if (timing) CCK(cudaEventRecord(h->ev[0], stream));
    for (int g = 0; g < generations; ++g) {
        cudaError_t e = (h->kind == PDEB_KIND_HEAT)
                            ? launch_gen<PDEB_KIND_HEAT, 30>(h->heat, h->pmax, h->d_y, population, out, h->d_cem, h->d_ticket, g, (int)n_elite, seed, lo, hi, std_floor, stream)
                            : launch_gen<PDEB_KIND_WAVE, 20>(h->wave, h->pmax, h->d_y, population, out, h->d_cem, h->d_ticket, g, (int)n_elite, seed, lo, hi, std_floor, stream);
        if (e != cudaSuccess) return cfail(PDEB_E_CUDA, "search kernel launch", e);
    }
    h->last_launches = generations;
    if (timing) CCK(cudaEventRecord(h->ev[1], stream));
    CCK(cudaMemcpyAsync(h->h_cem, h->d_cem, sizeof(PdebCemState), cudaMemcpyDeviceToHost, stream));
    if (h->sleep_on_done) CCK(cudaEventRecord(h->done, stream));
    return PDEB_E_OK;
}

/* the second half: wait for what search_enqueue queued and hand the results out */
static int search_collect(PdebControls* h, int32_t generations, bool skipped, double* best_y_out, double* best_value_out,
                          int32_t* found_out, double* history_out, double* mean_out, double* std_out, cudaStream_t stream,
                          PdebTiming* timing) {
    if (timing) memset(timing, 0, sizeof *timing);
    if (skipped) {
        if (found_out) *found_out = 0;
        if (best_value_out) *best_value_out = PDEB_FAILED_ROBUSTNESS;
        if (history_out) for (int g = 0; g < generations; ++g) history_out[g] = PDEB_FAILED_ROBUSTNESS;
        return PDEB_E_OK;
    }
    CCK(cudaSetDevice(h->device));
    if (h->sleep_on_done) CCK(cudaEventSynchronize(h->done));      /* sleeps: the other threads of a sweep need the cores to launch */
    else CCK(cudaStreamSynchronize(stream));
    const PdebCemState* hs = h->h_cem;
    const int nk = h->nk;
    const bool found = hs->has_best != 0;
    if (found_out) *found_out = found ? 1 : 0;
    if (best_value_out) *best_value_out = found ? hs->best_r : PDEB_FAILED_ROBUSTNESS;
    if (best_y_out) for (int j = 0; j < nk; ++j) best_y_out[j] = found ? hs->best_y[j] : 0.0;
    if (history_out) for (int g = 0; g < generations; ++g) history_out[g] = hs->history[g];
    if (mean_out) for (int j = 0; j < nk; ++j) mean_out[j] = hs->mean[j];
    if (std_out) for (int j = 0; j < nk; ++j) std_out[j] = hs->std[j];
    if (timing) {
        float ms = 0;
        CCK(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
        timing->kernel_ms = ms; timing->total_ms = ms; timing->launches = h->last_launches;
    }
    return PDEB_E_OK;
}

extern "C" int pdeb_controls_search(PdebControls* h, int64_t population, int32_t generations, double elite_frac, double lo, double hi,
                                    uint64_t seed, double* best_y_out, double* best_value_out, int32_t* found_out, double* history_out,
                                    double* mean_out, double* std_out, void* stream_, PdebTiming* timing) {
    bool skipped = false;
    int rc = search_enqueue(h, population, generations, elite_frac, lo, hi, seed, (cudaStream_t)stream_, timing != nullptr, &skipped);
    if (rc != PDEB_E_OK) return rc;
    return search_collect(h, generations, skipped, best_y_out, best_value_out, found_out, history_out, mean_out, std_out,
                          (cudaStream_t)stream_, timing);
}

/* Handles of finished sweeps, kept per device for the next one (creating a handle is seven device allocations, a page-locked
 * one, events: ~24 of them per sweep were most of a SHORT sweep's time).  pdeb_controls_pool_clear frees them. */
static std::mutex g_ctl_pool_mu;
static std::vector<PdebControls*> g_ctl_pool;
#define PDEB_CTL_POOL_MAX 96

static PdebControls* ctl_pool_take(int device, int64_t capacity) {
    std::lock_guard<std::mutex> lock(g_ctl_pool_mu);
    for (size_t i = 0; i < g_ctl_pool.size(); ++i) {
        PdebControls* h = g_ctl_pool[i];
        if (h->device == device && h->capacity >= capacity && h->capacity <= 4 * capacity) {
            g_ctl_pool.erase(g_ctl_pool.begin() + (long)i);
            return h;
        }
    }
    return nullptr;
}
static void ctl_pool_give(PdebControls* h) {
    {
        std::lock_guard<std::mutex> lock(g_ctl_pool_mu);
        if (g_ctl_pool.size() < PDEB_CTL_POOL_MAX) { g_ctl_pool.push_back(h); return; }
    }
    pdeb_controls_destroy(h);
}
extern "C" void pdeb_controls_pool_clear(void) {
    std::vector<PdebControls*> dead;
    {
        std::lock_guard<std::mutex> lock(g_ctl_pool_mu);
        dead.swap(g_ctl_pool);
    }
    for (PdebControls* h : dead) pdeb_controls_destroy(h);
}

/*
 * The search for MANY specifications: host threads, each with its own stream and ONE controls handle that it points at
 * sample after sample (no allocation per program, nothing per program in the caller's language).  The knot TIMES are the
 * batch's (knot_t == NULL: linspace(0, T, n_knots), convention C11); the knot VALUES found go to knot_y_out in the batch's
 * own plane-major layout, so a caller can hand in the batch's knot_y array and evaluate the batch right after.
 */
extern "C" int pdeb_controls_search_batch(const PdebBatch* in, const int64_t* samples, int64_t n_samples, int64_t population,
                                          int32_t generations, double elite_frac, const double* bounds, uint64_t seed, int device,
                                          const PdebConventions* conv, int nthreads, double* knot_y_out, double* best_out,
                                          int32_t* found_out, PdebTiming* timing) {
    if (!in || !bounds || !knot_y_out || !best_out || !found_out || n_samples < 0 || (samples == nullptr && n_samples != in->n))
        return cfail(PDEB_E_ARG, "pdeb_controls_search_batch: bad arguments (samples == NULL means all: n_samples == n)");
    if (population < 4 || generations < 1 || generations > PDEB_CEM_MAX_GEN || !(elite_frac > 0.0 && elite_frac <= 1.0))
        return cfail(PDEB_E_ARG, "pdeb_controls_search_batch: bad search budget");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return cfail(PDEB_E_NODEVICE, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) return cfail(PDEB_E_ARG, "bad device");
    if (timing) memset(timing, 0, sizeof *timing);
    if (n_samples == 0) return PDEB_E_OK;
    for (int64_t q = 0; q < n_samples; ++q)
        if (samples && (samples[q] < 0 || samples[q] >= in->n)) return cfail(PDEB_E_ARG, "pdeb_controls_search_batch: sample index out of range");
    /* default: 8 host threads, each with PDEB_SEARCH_DEPTH searches in flight on streams of their own.  Measured on the
     * 16-core B200 box (tools/search_depth_sweep.py, 4,096 x 6 candidates per program): 3.8e3 programs/s with 2 threads, 6.0e3
     * with 4, 9.1e3 with 8, no more with 16 -- about 24 searches in flight is what pays.  Needs more than CUDA's default 8
     * hardware queues (CUDA_DEVICE_MAX_CONNECTIONS=32, set by the Python package before CUDA starts: with 8 the streams alias)
     * and waiting threads that sleep instead of spinning (the blocking event of the handle). */
    unsigned hw = std::thread::hardware_concurrency();
    int nt = nthreads > 0 ? nthreads : (hw == 0 || hw >= 8 ? 8 : (int)hw);
    if (nt > 32) nt = 32;
    if ((int64_t)nt > n_samples) nt = (int)n_samples;
    const bool sleep_wait = nt * PDEB_SEARCH_DEPTH >= 8;
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    std::mutex err_mutex;
    std::string err_text;
    int err_code = PDEB_E_OK;
    std::vector<float> ms((size_t)nt, 0.f);
    std::vector<int> launches((size_t)nt, 0);
    const int64_t st = in->stride;

    /* every worker keeps PDEB_SEARCH_DEPTH searches in flight, each on a stream and a handle of its own: while the device
     * runs the 13 kernels of one, the thread is already launching the next (a search is ~1 ms of device time and ~0.1 ms of
     * launching) */
    auto work = [&](int w) {
        struct Slot { PdebControls* h = nullptr; cudaStream_t st = nullptr; int64_t s = -1; int nk = 0; bool busy = false, skipped = false; };
        Slot slot[PDEB_SEARCH_DEPTH];
        int rc = PDEB_E_OK;
        if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); rc = cfail(PDEB_E_CUDA, "pdeb_controls_search_batch: device"); }
        for (int q = 0; q < PDEB_SEARCH_DEPTH && rc == PDEB_E_OK; ++q)
            if (cudaStreamCreateWithFlags(&slot[q].st, cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError();
                rc = cfail(PDEB_E_CUDA, "pdeb_controls_search_batch: stream");
            }
        auto collect = [&](Slot& sl) {
            double by[PDEB_CTL_KMAX], best = 0.0;
            int32_t found = 0;
            PdebTiming tm;
            memset(&tm, 0, sizeof tm);
            const int r2 = search_collect(sl.h, generations, sl.skipped, by, &best, &found, nullptr, nullptr, nullptr, sl.st, timing ? &tm : nullptr);
            sl.busy = false;
            if (r2 != PDEB_E_OK) return r2;
            ms[(size_t)w] += tm.kernel_ms; launches[(size_t)w] += sl.skipped ? 0 : sl.h->last_launches;
            found_out[sl.s] = found;
            best_out[sl.s] = best;
            if (found) for (int j = 0; j < sl.nk; ++j) knot_y_out[(int64_t)j * st + sl.s] = by[j];
            return (int)PDEB_E_OK;
        };
        int cur = 0;
        while (rc == PDEB_E_OK && !failed.load(std::memory_order_relaxed)) {
            Slot& sl = slot[cur];
            if (sl.busy) { rc = collect(sl); if (rc != PDEB_E_OK) break; }       /* the oldest search of this thread */
            const int64_t q = next.fetch_add(1);
            if (q >= n_samples) break;
            const int64_t s = samples ? samples[q] : q;
            found_out[s] = 0;
            best_out[s] = PDEB_FAILED_ROBUSTNESS;
            const int kind = in->kind[s], N = in->n_elem[s];
            const int nk = in->n_knots[s];
            if (!((kind == PDEB_KIND_HEAT && N == 30) || (kind == PDEB_KIND_WAVE && N == 20)) || nk < 0 || nk > PDEB_CTL_KMAX || nk > in->kmax)
                continue;                                   /* not this family's: the caller searches it another way */
            double kt[PDEB_CTL_KMAX];
            for (int j = 0; j < nk; ++j)
                kt[j] = in->knot_t ? in->knot_t[(int64_t)j * st + s]
                                   : ((j == nk - 1 && nk > 1) ? in->T[s] : (nk > 1 ? (double)j * (in->T[s] / (double)(nk - 1)) : 0.0));
            if (!sl.h) sl.h = ctl_pool_take(device, population);
            if (!sl.h) rc = pdeb_controls_create(in, s, population, nk, kt, device, conv, &sl.h);
            else rc = pdeb_controls_respec(sl.h, in, s, nk, kt, conv);
            if (rc == PDEB_E_OK) sl.h->sleep_on_done = sleep_wait;
            if (rc != PDEB_E_OK) break;
            const double lo = bounds[kind == PDEB_KIND_HEAT ? 0 : 2], hi = bounds[kind == PDEB_KIND_HEAT ? 1 : 3];
            /* one generator stream per sample, whatever thread takes it */
            const uint64_t seed_s = pdeb_mix64(seed ^ pdeb_mix64((uint64_t)s));
            sl.s = s; sl.nk = nk;
            rc = search_enqueue(sl.h, population, generations, elite_frac, lo, hi, seed_s, sl.st, timing != nullptr, &sl.skipped);
            if (rc != PDEB_E_OK) break;
            sl.busy = true;
            cur = (cur + 1) % PDEB_SEARCH_DEPTH;
        }
        for (int q = 0; q < PDEB_SEARCH_DEPTH; ++q) {                            /* drain, oldest first */
            Slot& sl = slot[(cur + q) % PDEB_SEARCH_DEPTH];
            if (sl.busy) { const int r2 = collect(sl); if (rc == PDEB_E_OK) rc = r2; }
        }
        if (rc != PDEB_E_OK) {
            std::lock_guard<std::mutex> lock(err_mutex);
            if (err_code == PDEB_E_OK) { err_code = rc; err_text = pdeb_last_error(); }
            failed.store(1);
        }
        for (int q = 0; q < PDEB_SEARCH_DEPTH; ++q) {
            if (slot[q].st) cudaStreamSynchronize(slot[q].st);
            if (slot[q].h) { if (rc == PDEB_E_OK) ctl_pool_give(slot[q].h); else pdeb_controls_destroy(slot[q].h); }
            if (slot[q].st) cudaStreamDestroy(slot[q].st);
        }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < nt; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
    if (err_code != PDEB_E_OK) { pdeb_set_last_error(err_text.c_str()); return err_code; }
    if (timing) {
        for (int w = 0; w < nt; ++w) { timing->kernel_ms += ms[(size_t)w]; timing->launches += launches[(size_t)w]; }
        timing->total_ms = timing->kernel_ms;      /* sum over the threads' own streams (they overlap on the device) */
    }
    return PDEB_E_OK;
}
