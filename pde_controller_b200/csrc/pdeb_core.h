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
#define PDEB_WARP_OR(x) __reduce_or_sync(0xffffffffu, (x))
#define PDEB_WARP_MAX(x) __reduce_max_sync(0xffffffffu, (x))
#else
#define PDEB_ANY(x) (x)
#define PDEB_WARP_OR(x) (x)
#define PDEB_WARP_MAX(x) (x)
#endif

#define PDEB_EPS 1e-9

/* independent accumulators of the masked max over the nodes: four (2 and 8 measured within 2 %); the in-element end-point
 * rule relies on consecutive nodes falling into different accumulators (pdeb_inelement_gmax) */
#define PDEB_MARGIN_ACC 4
/* 1: the elimination multipliers of the register family live in a shared-memory column instead of registers */
#ifndef PDEB_C_SMEM
#define PDEB_C_SMEM 0
#endif
/* A/B switch of tools/variant_bench.py (default = what measured faster on B200):
 *   PDEB_EVENT_LIST 1: window events from a presorted per-lane list (one word per event), 0: every event re-scans the
 *                   windows of all clauses */
/*
 * Software pipelining across the basic blocks of a step (round 2; ptxas schedules inside a block only, and the step
 * is cut into blocks by the window events, the knot crossing and the pass): per kind (H heat, W wave)
 *   PDEB_PRE_TAB_x    rows per chain whose { R~_ii, 1/pivot } are loaded BEFORE the events / crossing blocks, so the solve
 *                     starts on operands that have arrived instead of on its own first loads;
 *   PDEB_PRE_SLOTS_x  the clause slots of the next pass are loaded behind the solve, not at the top of the pass;
 *   PDEB_PRE_TIME_x   the time of the next step is formed behind the solve of the current one.
 */
#ifndef PDEB_PRE_TAB_H
#define PDEB_PRE_TAB_H 3
#endif
#ifndef PDEB_PRE_TAB_W
#define PDEB_PRE_TAB_W 2
#endif
#ifndef PDEB_PRE_SLOTS_H
#define PDEB_PRE_SLOTS_H 1
#endif
#ifndef PDEB_PRE_SLOTS_W
#define PDEB_PRE_SLOTS_W 1
#endif
#ifndef PDEB_PRE_TIME_H
#define PDEB_PRE_TIME_H 1
#endif
#ifndef PDEB_PRE_TIME_W
#define PDEB_PRE_TIME_W 1
#endif
#define PDEB_PRE_TAB_MAX 8
/*   PDEB_ROTATE            the loop is rotated: the pass over step k+1 closes the iteration that advanced to it (the
 *                          unrotated form tests "last step?" between pass and advance: one more branch per step);
 *   PDEB_ENDPOINTS_MERGED  one warp-wide test skips both interpolated-end-point forms of the pass, and the in-element
 *                          form is a select instead of a divergent branch.
 * Measured on B200, 65,536 heat samples (tools/variant_bench.py): 0.3645 ms without any of these, 0.3535 with the three
 * prefetches, 0.3463 with all five; waves 1.108 -> 1.086 ms. */
#ifndef PDEB_PWL_SELECT
#define PDEB_PWL_SELECT 1
#endif
#ifndef PDEB_ROTATE
#define PDEB_ROTATE 1
#endif
#ifndef PDEB_ENDPOINTS_MERGED
#define PDEB_ENDPOINTS_MERGED 1
#endif
#ifndef PDEB_EVENT_LIST
#define PDEB_EVENT_LIST 1
#endif


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
 * Per-thread clause slots in shared memory: 7 cells of 8 bytes per clause, laid out [clause][cell][thread], so that
 * ONE per-thread base serves every field and the cells of the clause a lane is working on are base + immediate:
 *   0 mh = m*h (profile slope per node index)     1 c0 = b - g0
 *   2 A  running maximum over the window steps of (G: gmax, F: -gmax), gmax = max over the nodes of the
 *        relation-transformed margin; turned into the clause value (G: -A, F: A, + eq_eps for the band rule) once,
 *        after the time loop
 *   3, 4 xi of the two interval end points inside their elements (interpolated end points only)
 *   5 { nmask (bit i: node i inside [x_lo,x_hi]), info }
 *   6 two words of the lane's WINDOW-EVENT LIST: the steps at which the set of open windows changes, ascending, each
 *     as  step << 9 | last << 8 | open mask  (bit j: clause j's window holds that step; PDEB_MAX_STEPS < 2^23 and
 *     PDEB_MAX_PRED = 8 make it one word).  A clause contributes two boundaries (k_lo, k_hi + 1), so the 2 pmax words
 *     of the pmax cells always suffice.  The time loop reads ONE word per event instead of re-scanning every window.
 *   info: bits 0-1 rel, bit 2 top, bit 3 needs the interpolated end points (general form),
 *         bits 4-8 / 9-13 element of end point 0 / 1, bit 14 interval inside ONE element (nmask then holds that
 *         element's two nodes)
 * Reals are stored as `real` at the start of their cell (the FP32 build leaves the upper halves unused).
 */
#define PDEB_SLOT_CELLS 7
struct alignas(8) PdebU2 { uint32_t x, y; };
template <typename real, int STRIDE>
struct PdebClauseSlots {
    unsigned char* base;   /* + tid * 8 */
    PDEB_HD unsigned char* cell(int p, int f) const { return base + (size_t)((p * PDEB_SLOT_CELLS + f) * STRIDE) * 8; }
    PDEB_HD real& R(int p, int f) const { return *reinterpret_cast<real*>(cell(p, f)); }
    PDEB_HD PdebU2& NI(int p) const { return *reinterpret_cast<PdebU2*>(cell(p, 5)); }     /* { nmask, info } */
    /* cell 6 of clause p holds event words 2p and 2p + 1 of the lane's window-event list (below) */
    PDEB_HD uint32_t& EV(int e) const { return *reinterpret_cast<uint32_t*>(cell(e >> 1, 6) + (e & 1) * 4); }
    PDEB_HD PdebU2& WIN(int p) const { return *reinterpret_cast<PdebU2*>(cell(p, 6)); }    /* { klo, khi } (PDEB_EVENT_LIST 0) */
};

template <typename real>
PDEB_HD size_t pdeb_clause_smem_bytes(int threads, int pmax) {
    return (size_t)threads * pmax * PDEB_SLOT_CELLS * 8;
}

template <typename real, int STRIDE>
PDEB_HD PdebClauseSlots<real, STRIDE> pdeb_clause_slots(void* base, int pmax, int tid) {
    PdebClauseSlots<real, STRIDE> s;
    (void)pmax;
    s.base = (unsigned char*)base + (size_t)tid * 8;
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
/* 1/x to within an ulp: hardware seed + two Newton steps (5 instructions instead of the ~25 of an IEEE division with its
 * slow-path call).  Used where a segment slope is formed at a knot crossing, which happens in most steps of a warp
 * because its lanes cross their knots at different times, and for the ~60 (heat) / ~120 (waves) reciprocals of the
 * per-sample set-up (scales, pivots, the double-double quotients of the wave diagonal). */
PDEB_HD double pdeb_rcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}

/* A/B switch of tools/variant_bench.py: -DPDEB_SETUP_IEEE keeps IEEE divisions in the set-up */
PDEB_HD double pdeb_rcp_setup(double x) {
#ifdef PDEB_SETUP_IEEE
    return 1.0 / x;
#else
    return pdeb_rcp(x);
#endif
}

template <typename real>
struct PdebPwl {
    double t1, tn;        /* end of the current segment (+inf on the last); time of the knot after it (+inf: none) */
    real t0, y0, slope;   /* current segment */
    real y1, yn;          /* values at t1 and tn */
    int nk, j;            /* j: index of the knot at t1 */
    double ustep, uT;     /* uniform knots (knot_t == NULL in the batch): linspace(0, T, nk), exactly as numpy forms it */

    /* time of knot q: from the batch, or q * (T / (nk-1)) with the last one set to T (one knot: t = 0) */
    PDEB_HD double knot_time(const double* kt, int q, int64_t stride) const {
        if (kt) return kt[(int64_t)q * stride];
        return (q == nk - 1 && nk > 1) ? uT : (double)q * ustep;
    }
    PDEB_HD void init(const double* kt, const double* ky, int nk_, int64_t stride, double T = 0.0) {
        nk = nk_; j = 0;
        uT = T; ustep = (nk > 1) ? T / (double)(nk - 1) : 0.0;
        t0 = 0; slope = 0; y0 = 0; y1 = 0; yn = 0; t1 = INFINITY; tn = INFINITY;
        if (nk >= 1) { y0 = y1 = (real)ky[0]; t1 = knot_time(kt, 0, stride); }
        if (nk >= 2) { tn = knot_time(kt, 1, stride); yn = (real)ky[stride]; }
    }
    /* kt, ky already point at this sample's first knot (kt may be NULL); stride = plane stride */
    PDEB_HD real eval(double t, const double* kt, const double* ky, int64_t stride) {
        while (t >= t1) {
            t0 = (real)t1; y0 = y1;
            if (tn < INFINITY) {
                slope = (real)(((double)yn - (double)y1) * pdeb_rcp(tn - t1));
                t1 = tn; y1 = yn;
                ++j;
                if (j + 1 < nk) { tn = knot_time(kt, j + 1, stride); yn = (real)ky[(int64_t)(j + 1) * stride]; }
                else tn = INFINITY;
            } else { slope = 0; t1 = INFINITY; }
        }
        return fma(slope, (real)(t - (double)t0), y0);
    }
};

/*
 * The same input for the register family, with the NEXT knot staged in shared memory.
 * In that family the lanes of a warp cross their knots at different steps, so a warp runs the crossing code in about
 * half of its steps, and a knot fetched into a register there stalls the whole warp for a global-memory latency: the
 * value has to be merged into the register the loop carries, which the compiler does with a move that waits for the
 * load right where it was issued (measured: 7.6 % of the heat kernel in that one stall).  Here the crossing reads the
 * knot from a per-thread shared-memory slot (two doubles: value, time) and refills the slot with an ASYNCHRONOUS copy
 * (cp.async, global -> shared, no register destination) that has the ~20 steps until the next crossing to land.
 * Arithmetic and knot semantics are those of PdebPwl.
 */
template <typename real, int STRIDE>
struct PdebPwlStaged {
    double t1;            /* end of the current segment (+inf on the last) */
    real t0, y0, slope;   /* current segment */
    real y1;              /* value at t1 */
    int nk, j;            /* j: index of the knot at t1 */
    double ustep, uT;     /* uniform knots, as in PdebPwl */
    /* shared, + tid: slot[0] value of knot j+1, slot[STRIDE] its time (explicit knot times only) */
#if defined(__CUDA_ARCH__)
    uint32_t slot;        /* shared-window address */
#else
    double* slot;
#endif

    PDEB_HD double knot_time(const double* kt, int q, int64_t stride) const {
        if (kt) return kt[(int64_t)q * stride];
        return (q == nk - 1 && nk > 1) ? uT : (double)q * ustep;
    }
    /* start the copy of knot q into the slot */
    PDEB_HD void stage(const double* kt, const double* ky, int q, int64_t stride) {
#if defined(__CUDA_ARCH__)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(slot), "l"(ky + (int64_t)q * stride) : "memory");
        if (kt)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(slot + (uint32_t)(STRIDE * sizeof(double))),
                         "l"(kt + (int64_t)q * stride) : "memory");
#else
        slot[0] = ky[(int64_t)q * stride];
        if (kt) slot[STRIDE] = kt[(int64_t)q * stride];
#endif
    }
    /* staged value (which = 0) or time (which = 1), once the copy has landed */
    PDEB_HD double staged(int which, bool wait = true) const {
#if defined(__CUDA_ARCH__)
        double v;
        if (wait) asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(slot + (uint32_t)(which * STRIDE * sizeof(double))) : "memory");
        return v;
#else
        return slot[which * STRIDE];
#endif
    }
    PDEB_HD void init(const double* kt, const double* ky, int nk_, int64_t stride, double T, double* slot_) {
        nk = nk_; j = 0;
#if defined(__CUDA_ARCH__)
        slot = (uint32_t)__cvta_generic_to_shared(slot_);
#else
        slot = slot_;
#endif
        uT = T; ustep = (nk > 1) ? T / (double)(nk - 1) : 0.0;
        t0 = 0; slope = 0; y0 = 0; y1 = 0; t1 = INFINITY;
        if (nk >= 1) { y0 = y1 = (real)ky[0]; t1 = knot_time(kt, 0, stride); }
        if (nk >= 2) stage(kt, ky, 1, stride);
    }
    PDEB_HD real eval(double t, const double* kt, const double* ky, int64_t stride) {
        while (t >= t1) {
            t0 = (real)t1; y0 = y1;
            double tn = INFINITY;
            real yn = 0;
            if (j + 1 < nk) {
#if PDEB_PWL_SELECT
                /* both forms of the knot time, then a select: no branch on the (batch-wide) kind of knots */
                const double tu = knot_time(nullptr, j + 1, 0);
                yn = (real)staged(0);
                const double ts = staged(1, false);
                tn = kt ? ts : tu;
#else
                yn = (real)staged(0);
                tn = kt ? staged(1) : knot_time(nullptr, j + 1, 0);
#endif
            }
            if (tn < INFINITY) {
                slope = (real)(((double)yn - (double)y1) * pdeb_rcp(tn - t1));
                t1 = tn; y1 = yn;
                ++j;
                if (j + 1 < nk) stage(kt, ky, j + 1, stride);
            } else { slope = 0; t1 = INFINITY; }
        }
        return fma(slope, (real)(t - (double)t0), y0);
    }
};
#define PDEB_PWL_SLOT_BYTES (2 * sizeof(double))   /* per thread */

struct alignas(16) PdebD2 { double x, y; };
struct alignas(8) PdebF2 { float x, y; };
template <typename real> struct PdebVec2;
template <> struct PdebVec2<double> { typedef PdebD2 type; };
template <> struct PdebVec2<float> { typedef PdebF2 type; };

/* shared memory of one block of the register family: [row table | clause slots] */
template <typename real>
PDEB_HD size_t pdeb_table_smem_bytes(int threads, int n_rows, int kind) {
    /* { R~_ii, 1/pivot_i } rows, the column of scales s_i, for waves a column for the low part of R~_ii, then
     * (optionally) the multipliers */
    const int cols = 1 + (kind == PDEB_KIND_WAVE ? 1 : 0) + (PDEB_C_SMEM ? 1 : 0);
    return (size_t)threads * n_rows * (sizeof(typename PdebVec2<real>::type) + cols * sizeof(real));
}

/*
 * Register-family trajectory.  KIND: PDEB_KIND_*; N: elements = free dofs (nodes 1..N), 4 <= N <= 31.
 * Arrays are indexed by node; index 0 is the Dirichlet node (deviation 0).
 *
 * SCALED UNKNOWNS.  The step  A x' = R x + f e_N  (A = M + cA K SPD tridiagonal, R = M - cR K or -K, both assembled
 * from two materials) is solved for x~ = S^-1 x with the diagonal scaling S chosen such that
 *     s_i s_{i+1} R_{i,i+1} = 1        (R's off-diagonals are positive: m/6 + cR k, or k)
 * i.e. the scaled operator R~ = S R S has UNIT off-diagonals and  (R~ x~)_i = R~_ii x~_i + (x~_{i-1} + x~_{i+1}):
 * one add and one FMA per row, no per-element material constant in the time loop at all (the unscaled form needs
 * the off-diagonal of every element: 60 registers of hoisted selects, or three instructions per element and step).
 * A~ = S A S stays SPD tridiagonal with off-diagonals ao_e / ro_e.  It is factored ONCE per sample as a TWISTED
 * L D L^T: rows 1..M-1 are eliminated downwards, rows N..M+1 upwards, and the two sweeps meet at row M, so each
 * time step runs two independent dependency chains of half the length.  Per row and step: 1 ADD + 1 FMA (R~ x~),
 * 1 FMA (eliminate), 1 MUL (1/pivot), 1 FMA (back substitute) = 5 FP64 operations.
 * The physical value s_i x~_i is only formed where a clause looks at it (one FMA with the profile folded in).
 *   registers      : scaled state, the N-1 edge multipliers c[]
 *   shared (column): per row { R~_ii, 1/pivot_i } (one 16-byte load per row and step) and s_i (clause passes only)
 */
template <int KIND, int N, typename real, int STRIDE>
struct PdebTrajReg {
    static_assert(N >= 4 && N <= 31, "register family handles 4..31 elements");
    static constexpr int M = (N + 1) / 2;   /* twist row */
    typedef typename PdebVec2<real>::type real2;

    uint32_t omask;    /* bit e: element e is of the outer material (set-up only) */
#if PDEB_C_SMEM
    real* ctab;        /* shared column: ctab[(i-1)*STRIDE] = elimination multiplier of the edge (i, i+1), i = 1..N-1 */
    PDEB_HD real getc(int i) const { return ctab[(i - 1) * STRIDE]; }
    PDEB_HD void setc(int i, real v) { ctab[(i - 1) * STRIDE] = v; }
#else
    real c[N];         /* c[i], i = 1..N-1: elimination multiplier of the edge (i, i+1) of A~ */
    PDEB_HD real getc(int i) const { return c[i]; }
    PDEB_HD void setc(int i, real v) { c[i] = v; }
#endif
    real w[N + 2];     /* SCALED deviation: heat temperature d - g0 = s_i w[i]; wave displacement likewise */
    real v[KIND == PDEB_KIND_WAVE ? N + 2 : 1];
    real a[KIND == PDEB_KIND_WAVE ? N + 2 : 1];
    real sN;           /* s_N: the load enters row N as s_N f */
    real2* tab;        /* shared: tab[(i-1)*tstride] = { R~_ii, 1/pivot_i } */
    real* stab;        /* shared: stab[(i-1)*tstride] = s_i */
    real* ltab;        /* shared, waves only: low-order part of R~_ii (see setup) */
    static constexpr bool DD = (KIND == PDEB_KIND_WAVE);   /* R~_ii carried as an unevaluated sum hi + lo */
    static constexpr int tstride = STRIDE;   /* threads per block: rows are base + immediate offset */

    PDEB_HD bool outer(int e) const { return (omask >> e) & 1u; }
    PDEB_HD real scale(int i) const { return stab[(i - 1) * tstride]; }

    /*
     * z[1..N] <- A~^-1 (R~ x + fNs e_N), fNs = s_N f, the two chains interleaved row by row.
     * x and z may not alias (x is read at i-1, i, i+1); x[0] must be 0.
     */
    template <int PRE = 0>
    PDEB_HD void step_solve(const real* x, real fNs, real* z, const real2* pre_t = nullptr, const real2* pre_b = nullptr) const {
        /* right-hand side + forward elimination; z[i] leaves this loop already divided by its pivot */
        real zt = 0, zb = 0;    /* undivided value of the last row of each chain */
        PDEB_UNROLL
        for (int j = 0; j < N; ++j) {
            /* top row it = 1 + j (while it < M), bottom row ib = N - j (while ib > M) */
            const int it = 1 + j, ib = N - j;
            if (it < M) {
                const real2 t = (j < PRE) ? pre_t[j < PRE ? j : 0] : tab[(it - 1) * tstride];
                real r = x[it - 1] + x[it + 1];
                if (DD) r = fma(ltab[(it - 1) * tstride], x[it], r);
                r = fma(t.x, x[it], r);
                if (it > 1) r = fma(-getc(it - 1), zt, r);
                zt = r;
                z[it] = r * t.y;
            }
            if (ib > M) {
                const real2 t = (j < PRE) ? pre_b[j < PRE ? j : 0] : tab[(ib - 1) * tstride];
                real r = x[ib - 1] + ((ib == N) ? fNs : x[ib + 1]);
                if (DD) r = fma(ltab[(ib - 1) * tstride], x[ib], r);
                r = fma(t.x, x[ib], r);
                if (ib < N) r = fma(-getc(ib), zb, r);
                zb = r;
                z[ib] = r * t.y;
            }
        }
        {
            const real2 t = tab[(M - 1) * tstride];
            real r = x[M - 1] + x[M + 1];
            if (DD) r = fma(ltab[(M - 1) * tstride], x[M], r);
            r = fma(t.x, x[M], r);
            r = fma(-getc(M - 1), zt, r);
            r = fma(-getc(M), zb, r);
            z[M] = r * t.y;
        }
        /* back substitution outwards from the twist row */
        PDEB_UNROLL
        for (int j = 1; j < N; ++j) {
            const int it = M - j, ib = M + j;
            if (it >= 1) z[it] = fma(-getc(it), z[it + 1], z[it]);
            if (ib <= N) z[ib] = fma(-getc(ib - 1), z[ib - 1], z[ib]);
        }
    }

    /* per-material element coefficients of A = M + cA K and of R (M - cR K for heat, -K for waves) */
    struct Coef {
        double ad_o, ad_i, ao_o, ao_i;   /* A: half diagonal, off diagonal; o = outer material, i = inner */
        double rd_o, rd_i, ro_o, ro_i;   /* R likewise */
    };
    static PDEB_HD Coef coef(double h, double rho_in, double rho_out, double E_in, double E_out, int lumped, double cA, double cR) {
        Coef q;
        const double m_o = rho_out * h, k_o = E_out / h, m_i = rho_in * h, k_i = E_in / h;
        const double mo_o = lumped ? 0.0 : m_o / 6.0, mo_i = lumped ? 0.0 : m_i / 6.0;
        const double md_o = lumped ? 0.5 * m_o : m_o / 3.0, md_i = lumped ? 0.5 * m_i : m_i / 3.0;
        q.ad_o = md_o + cA * k_o; q.ad_i = md_i + cA * k_i;
        q.ao_o = mo_o - cA * k_o; q.ao_i = mo_i - cA * k_i;
        q.rd_o = (KIND == PDEB_KIND_HEAT) ? md_o - cR * k_o : -k_o;
        q.rd_i = (KIND == PDEB_KIND_HEAT) ? md_i - cR * k_i : -k_i;
        q.ro_o = (KIND == PDEB_KIND_HEAT) ? mo_o + cR * k_o : k_o;
        q.ro_i = (KIND == PDEB_KIND_HEAT) ? mo_i + cR * k_i : k_i;
        return q;
    }

    /*
     * Set up scaling, coefficients and factors.  Returns false on a non-positive pivot.
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
        const Coef q = coef(h, rho_in, rho_out, E_in, E_out, lumped, cA, cR);
#define PDEB_EL(e, vo, vi) (outer(e) ? (vo) : (vi))
        bool ok = true;
        /* scaling: s_1 = ro_1^-1/2 (uniform rods get a uniform scale), s_i = 1 / (ro_{i-1} s_{i-1}); element e joins nodes e, e+1 */
        {
            double sp = 1.0 / sqrt(PDEB_EL(1, q.ro_o, q.ro_i));
            stab[0] = (real)sp;
            PDEB_UNROLL
            for (int i = 2; i <= N; ++i) {
                sp = pdeb_rcp_setup(PDEB_EL(i - 1, q.ro_o, q.ro_i) * sp);
                stab[(i - 1) * tstride] = (real)sp;
                if (i == N) sN = (real)sp;
            }
            if (!isfinite(sp) || !(sp > 0.0)) ok = false;
        }
        /* scaled off-diagonal of A per material: s_i s_{i+1} ao_e = ao_e / ro_e */
        const double as_o = q.ao_o / q.ro_o, as_i = q.ao_i / q.ro_i;
        /* twisted L D L^T of A~, streamed: only the previous pivot of each sweep is kept live */
        double dtop = 0.0, dbot = 0.0;     /* 1/pivot of the last eliminated row of each sweep */
        PDEB_UNROLL
        for (int i = 1; i < M; ++i) {
            const double si = (double)stab[(i - 1) * tstride], s2 = si * si;
            double piv = s2 * (PDEB_EL(i - 1, q.ad_o, q.ad_i) + PDEB_EL(i, q.ad_o, q.ad_i));
            if (i > 1) { const double aol = PDEB_EL(i - 1, as_o, as_i), ci = aol * dtop; setc(i - 1, (real)ci); piv -= ci * aol; }
            if (!(piv > 0.0)) ok = false;
            dtop = pdeb_rcp_setup(piv);
            real2 t;
            t.x = (real)(s2 * (PDEB_EL(i - 1, q.rd_o, q.rd_i) + PDEB_EL(i, q.rd_o, q.rd_i)));
            t.y = (real)dtop;
            tab[(i - 1) * tstride] = t;
        }
        PDEB_UNROLL
        for (int i = N; i > M; --i) {
            const double si = (double)stab[(i - 1) * tstride], s2 = si * si;
            double piv = s2 * (PDEB_EL(i - 1, q.ad_o, q.ad_i) + ((i < N) ? PDEB_EL(i, q.ad_o, q.ad_i) : 0.0));
            if (i < N) { const double aor = PDEB_EL(i, as_o, as_i), ci = aor * dbot; setc(i, (real)ci); piv -= ci * aor; }
            if (!(piv > 0.0)) ok = false;
            dbot = pdeb_rcp_setup(piv);
            real2 t;
            t.x = (real)(s2 * (PDEB_EL(i - 1, q.rd_o, q.rd_i) + ((i < N) ? PDEB_EL(i, q.rd_o, q.rd_i) : 0.0)));
            t.y = (real)dbot;
            tab[(i - 1) * tstride] = t;
        }
        {
            const double si = (double)stab[(M - 1) * tstride], s2 = si * si;
            const double aol = PDEB_EL(M - 1, as_o, as_i), aor = PDEB_EL(M, as_o, as_i);
            const double ct = aol * dtop, cb = aor * dbot;
            setc(M - 1, (real)ct);
            setc(M, (real)cb);
            const double piv = s2 * (PDEB_EL(M - 1, q.ad_o, q.ad_i) + PDEB_EL(M, q.ad_o, q.ad_i)) - ct * aol - cb * aor;
            if (!(piv > 0.0)) ok = false;
            const double dm = pdeb_rcp_setup(piv);
            if (!isfinite(dm) || !isfinite(dtop) || !isfinite(dbot)) ok = false;
            real2 t;
            t.x = (real)(s2 * (PDEB_EL(M - 1, q.rd_o, q.rd_i) + PDEB_EL(M, q.rd_o, q.rd_i)));
            t.y = (real)dm;
            tab[(M - 1) * tstride] = t;
        }
        if (DD) {
            /*
             * Waves: R = -K must annihilate the rigid translation, or rounding R~_ii acts as a spring to ground of
             * relative stiffness ~1e-16 whose phase error the undamped Newmark scheme accumulates over the run
             * (measured: ~1e-11 of the field amplitude after 800 steps).  With the off-diagonals taken as exactly 1
             * the consistent diagonal is  R~_ii = -(s_i/s_{i-1} + s_i/s_{i+1})  (node 1: s_1^2 k_0 for the element at
             * the Dirichlet end); it is carried as hi + lo to twice the working precision.
             */
            PDEB_UNROLL
            for (int i = 1; i <= N; ++i) {
                const double si = (double)stab[(i - 1) * tstride];
                double ah, al, bh = 0.0, bl = 0.0;
                if (i > 1) {
                    const double sm = (double)stab[(i - 2) * tstride], rm = pdeb_rcp_setup(sm);
                    ah = si * rm;
                    al = fma(-ah, sm, si) * rm;
                } else {
                    const double k0 = -PDEB_EL(0, q.rd_o, q.rd_i);
                    const double ph = si * si, pl = fma(si, si, -ph);
                    ah = ph * k0;
                    al = fma(ph, k0, -ah) + pl * k0;
                }
                if (i < N) {
                    const double sp = (double)stab[i * tstride], rp = pdeb_rcp_setup(sp);
                    bh = si * rp;
                    bl = fma(-bh, sp, si) * rp;
                }
                const double sh = ah + bh, bb = sh - ah;
                const double err = (ah - (sh - bb)) + (bh - bb);
                real2 t = tab[(i - 1) * tstride];
                t.x = (real)(-sh);
                tab[(i - 1) * tstride] = t;
                ltab[(i - 1) * tstride] = (real)(-(err + al + bl) - ((double)(real)sh - sh));
            }
        }
#undef PDEB_EL
        return ok;
    }

    /*
     * Wave initial state in scaled unknowns: d = 0 on the free nodes (deviation -g0), v = 0,
     * a0 = M^-1 (f0 e_N - K w0), computed in physical variables and divided by s_i (set-up only).
     */
    PDEB_HD void wave_init(double h, double g0, double f0, double rho_in, double rho_out, double E_in, double E_out, int lumped) {
        const Coef q = coef(h, rho_in, rho_out, E_in, E_out, lumped, 0.0, 0.0);      /* R = -K */
        const double m_o = rho_out * h, m_i = rho_in * h;
        const double wd = -g0;
        double cp[N + 1], x[N + 1];
        double mo_prev = 0.0;
        PDEB_UNROLL
        for (int i = 1; i <= N; ++i) {
            const bool ol = outer(i - 1), orr = (i < N) ? outer(i) : false;
            /* (R w0)_i with w0 = wd on nodes 1..N and 0 on the Dirichlet node */
            double r = wd * ((ol ? q.rd_o : q.rd_i) + ((i < N) ? (orr ? q.rd_o : q.rd_i) : 0.0)
                             + ((i > 1) ? (ol ? q.ro_o : q.ro_i) : 0.0) + ((i < N) ? (orr ? q.ro_o : q.ro_i) : 0.0));
            if (i == N) r += f0;
            const double ml = ol ? m_o : m_i, mr = (i < N) ? (orr ? m_o : m_i) : 0.0;
            if (lumped) {
                x[i] = r / (0.5 * (ml + mr));
            } else {
                /* consistent mass: one Thomas solve with M */
                const double md = (ml + mr) / 3.0, mo_i = mr / 6.0;
                const double den = md - ((i > 1) ? mo_prev * cp[i - 1] : 0.0);
                cp[i] = mo_i / den;
                x[i] = (r - ((i > 1) ? mo_prev * x[i - 1] : 0.0)) / den;
                mo_prev = mo_i;
            }
        }
        if (!lumped) {
            PDEB_UNROLL
            for (int i = N - 1; i >= 1; --i) x[i] -= cp[i] * x[i + 1];
        }
        PDEB_UNROLL
        for (int i = 1; i <= N; ++i) {
            const double si = (double)scale(i);
            w[i] = (real)(wd / si);
            a[i] = (real)(x[i] / si);
        }
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
/* hi <- g where bit I of nmask is set and g > hi.  The FP64 device form is spelt in PTX: one LOP3 with a predicate
 * result, one DSETP that takes it as its AND input, one 64-bit select (the compiler's own rendering of the C form
 * below spent eleven instructions per node on shifts, predicate logic and register re-pairing). */
template <int I>
PDEB_HD void pdeb_masked_max(double& hi, double g, uint32_t nmask) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .pred pm, pg;\n\t.reg .b32 t;\n\t"
        "and.b32 t, %2, %3;\n\t"
        "setp.ne.u32 pm, t, 0;\n\t"
        "setp.gt.and.f64 pg, %1, %0, pm;\n\t"
        "selp.f64 %0, %1, %0, pg;\n\t}" : "+d"(hi) : "d"(g), "r"(nmask), "n"(1u << I));
#else
    if (((nmask >> I) & 1u) && g > hi) hi = g;
#endif
}
template <int I>
PDEB_HD void pdeb_masked_max(float& hi, float g, uint32_t nmask) {
    if (((nmask >> I) & 1u) && g > hi) hi = g;
}
template <int I, int END, typename real, int STRIDE>
struct PdebMarginLoop {
    /* w is the SCALED state, stab the column of scales: d_I = s_I w_I - (mh I + c0) */
    static PDEB_HD void run(const real* w, const real* stab, real mh, real c0, uint32_t nmask, uint32_t and_mask,
                            uint32_t xor_mask, real* hi) {
        const real q = fma(-mh, (real)I, -c0);
        const real d = (I == 0) ? q : fma(stab[(I == 0 ? 0 : I - 1) * STRIDE], w[I], q);
        pdeb_masked_max<I>(hi[I % PDEB_MARGIN_ACC], pdeb_signop(d, and_mask, xor_mask), nmask);
        PdebMarginLoop<I + 1, END, real, STRIDE>::run(w, stab, mh, c0, nmask, and_mask, xor_mask, hi);
    }
};
template <int END, typename real, int STRIDE>
struct PdebMarginLoop<END, END, real, STRIDE> {
    static PDEB_HD void run(const real*, const real*, real, real, uint32_t, uint32_t, uint32_t, real*) {}
};

/* masked max over the nodes, left in PDEB_MARGIN_ACC partial accumulators (node I goes to hi[I % PDEB_MARGIN_ACC]; -inf where
 * no node of that residue class is in the mask) */
template <int N, typename real, int STRIDE>
PDEB_HD void pdeb_clause_partials(const real* w, const real* stab, real mh, real c0, uint32_t nmask, uint32_t and_mask,
                                  uint32_t xor_mask, real* hi) {
    for (int q = 0; q < PDEB_MARGIN_ACC; ++q) hi[q] = (real)(-INFINITY);
    PdebMarginLoop<0, N + 1, real, STRIDE>::run(w, stab, mh, c0, nmask, and_mask, xor_mask, hi);
}
template <typename real>
PDEB_HD real pdeb_partials_max(const real* hi) {
    static_assert(PDEB_MARGIN_ACC == 4, "four partial accumulators");
    const real a = hi[0] > hi[2] ? hi[0] : hi[2], b = hi[1] > hi[3] ? hi[1] : hi[3];
    return a > b ? a : b;
}

/*
 * Interval strictly inside ONE element e (it holds no node, C8): the margin lives on the two interpolated end points
 *     d(xi) = d_e + xi (d_{e+1} - d_e).
 * The masked loop itself delivers the two RAW node margins: with the mask {e, e+1} and the identity sign treatment,
 * consecutive nodes fall into different partial accumulators, hi[e % ACC] = d_e and hi[(e+1) % ACC] = d_{e+1}.
 * So the common case of the end-point rule costs a handful of selects after the loop every lane runs anyway,
 * instead of a copy of the state to local memory (pdeb_endpoint_gmax below, kept for region_rule 1).
 */
template <typename real>
PDEB_HD real pdeb_pick4(const real* hi, uint32_t q) {
    static_assert(PDEB_MARGIN_ACC == 4, "pdeb_pick4 selects among four partial accumulators");
    const real lo2 = (q & 1u) ? hi[1] : hi[0], hi2 = (q & 1u) ? hi[3] : hi[2];
    return (q & 2u) ? hi2 : lo2;
}
template <typename real>
PDEB_HD real pdeb_inelement_gmax(const real* hi, uint32_t e, real xi0, real xi1, uint32_t and_mask, uint32_t xor_mask) {
    const real de = pdeb_pick4<real>(hi, e), df = pdeb_pick4<real>(hi, e + 1u);
    const real dd = df - de;
    const real g0 = pdeb_signop(fma(xi0, dd, de), and_mask, xor_mask);
    const real g1 = pdeb_signop(fma(xi1, dd, de), and_mask, xor_mask);
    real g = (real)(-INFINITY);
    if (g0 > g) g = g0;
    if (g1 > g) g = g1;
    return g;
}

/*
 * Interpolated end points of a clause interval (C8: intervals that hold no node, or region_rule 1).
 * The profile is affine, so the margin at a point of element e is the linear interpolation of the NODE margins:
 * d(x) = d_e + xi (d_{e+1} - d_e).  The state is copied to a local array so that it can be indexed by the element
 * number; element numbers and xi were fixed at set-up (clause slots), nothing is re-read from the batch.
 */
template <int N, typename real, int STRIDE>
PDEB_HD real pdeb_endpoint_gmax(const real* w, const real* stab, real mh, real c0, uint32_t info, real xi0, real xi1,
                                uint32_t and_mask, uint32_t xor_mask) {
    real wl[N + 2];
    PDEB_UNROLL
    for (int i = 0; i <= N; ++i) wl[i] = (i == 0) ? (real)0 : w[i];
    real g = (real)(-INFINITY);
    for (int q = 0; q < 2; ++q) {
        const int e = (int)((info >> (4 + 5 * q)) & 31u);            /* 0 .. N-1 */
        const real xi = q ? xi1 : xi0;
        const real se = (e == 0) ? (real)0 : stab[(e == 0 ? 0 : e - 1) * STRIDE], sf = stab[e * STRIDE];
        const real de = fma(se, wl[e], fma(-mh, (real)e, -c0));
        const real df = fma(sf, wl[e + 1], fma(-mh, (real)(e + 1), -c0));
        const real gq = pdeb_signop(fma(xi, df - de, de), and_mask, xor_mask);
        if (gq > g) g = gq;
    }
    return g;
}

/*
 * Discrete form of one clause  TOP_[t_lo,t_hi] ( forall x in [x_lo,x_hi]: u(x) - (m x + b) REL 0 )  on a mesh of N
 * elements of size h with S steps of size dt (conventions C7, C8): window steps, node mask, profile per node
 * index, and for intervals that need interpolated end points their elements and local coordinates.
 */
struct PdebClauseData {
    double mh, c0;        /* profile per node index: p(x_i) - g0 = mh i + c0 */
    double xi[2];         /* local coordinate of each interval end inside its element (info bit 3) */
    int32_t klo, khi;     /* window, in steps */
    uint32_t nmask;       /* bit i: node i inside the interval */
    uint32_t info;        /* bits 0-1 rel, bit 2 top, bit 3 end points needed, bits 4-8 / 9-13 their elements */
};
template <int N>
PDEB_HD PdebClauseData pdeb_clause_data(double m, double bb, double xlo, double xhi, double tlo, double thi, uint32_t rel,
                                        uint32_t top, double L, double h, double dt, double g0, int64_t S,
                                        const PdebConventions& cv) {
    PdebClauseData cd;
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
    const uint32_t endp = (nm == 0 || cv.region_rule == 1) ? 1u : 0u;
    uint32_t ebits = 0;
    cd.xi[0] = 0.0; cd.xi[1] = 0.0;
    if (endp) {
        const double xe[2] = { xlo, xhi };
        for (int q = 0; q < 2; ++q) {
            const double xx = xe[q] < 0 ? 0 : (xe[q] > L ? L : xe[q]);
            int e = (int)floor(xx / h);
            if (e < 0) e = 0;
            if (e > N - 1) e = N - 1;
            cd.xi[q] = (xx - e * h) / h;
            ebits |= (uint32_t)e << (4 + 5 * q);
        }
    }
    cd.mh = m * h;
    cd.c0 = bb - g0;
    cd.klo = (int32_t)a0;
    cd.khi = (int32_t)a1;
    cd.nmask = nm;
    cd.info = rel | (top << 2) | (endp << 3) | ebits;
    return cd;
}

/*
 * Whole trajectory of the sample at position s of the input image; results go to index so of the outputs (the image
 * the kernels read is in plan order, the outputs in the caller's).  `slots` is this thread's clause column in shared
 * memory, `table` its row table.  Writes robustness / clause / status.
 *
 * Time loop, per step k:  (1) clauses whose window is open are reduced on the state w^k (a warp
 * runs as many reduction passes as its busiest lane has open clauses, usually one),
 * (2) w^{k+1} is solved for.
 * Lanes whose trajectory is over keep stepping harmlessly until the warp is done (the batch is
 * ordered by step count, so that is a handful of steps): their observations are switched off.
 */
template <int KIND, int N, typename real, int STRIDE>
PDEB_HD void pdeb_run_reg(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, int64_t s, int64_t so,
                          typename PdebVec2<real>::type* table, real* stab, PdebClauseSlots<real, STRIDE> slots, bool live,
                          double* pwl_slot, real* ctab = nullptr) {
    /* `live` is false for padding lanes of a partially filled warp: they run the loop with S = 0 */
    const int64_t st = b.stride;
    int status = PDEB_OK;
    double dt = 1;
    int np = 0;
    int32_t S32 = 0;          /* steps actually simulated: up to the last observed one */
    PdebTrajReg<KIND, N, real, STRIDE> tr;
    tr.tab = table;
    tr.stab = stab;
    tr.ltab = stab + (size_t)N * STRIDE;
    tr.sN = 0;
#if PDEB_C_SMEM
    tr.ctab = ctab;
#endif
    PdebPwlStaged<real, STRIDE> in;
    in.init(nullptr, nullptr, 0, 0, 0.0, pwl_slot);
    uint32_t evw = 0x100u;     /* the NEXT window event of this lane (clause-slot cell 6); none yet */
    int32_t next_ev = 0x7fffffff;
    int evi = 0;
    const double* kt = b.knot_t ? b.knot_t + (live ? s : 0) : nullptr;     /* NULL: uniform knots */
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
            for (int j = 1; j < nk && kt; ++j)
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
            in.init(kt, ky, nk, st, T, pwl_slot);
            if (KIND == PDEB_KIND_WAVE && ok) {
                const double f0 = (double)in.eval(0.0, kt, ky, st);
                tr.wave_init(h, g0, f0, rho_in, rho_out, E_in, E_out, lumped);
            }
        }
        if (status != PDEB_OK) { S = 0; np = 0; }
        int32_t k_end = 0;         /* last step any clause observes: nothing after it can change the result */
        int32_t wlo[PDEB_MAX_PRED], whi[PDEB_MAX_PRED];    /* window steps of the clauses (set-up only) */
        /* clause slots */
        for (int j = 0; j < np; ++j) {
            const double m = b.pred_m[(int64_t)j * st + s], bb = b.pred_b[(int64_t)j * st + s];
            const double xlo = b.pred_xlo[(int64_t)j * st + s], xhi = b.pred_xhi[(int64_t)j * st + s];
            const double tlo = b.pred_tlo[(int64_t)j * st + s], thi = b.pred_thi[(int64_t)j * st + s];
            const PdebClauseData cd = pdeb_clause_data<N>(m, bb, xlo, xhi, tlo, thi, b.pred_rel[(int64_t)j * st + s],
                                                          b.pred_top[(int64_t)j * st + s], L, h, dt, g0, S, cv);
            slots.R(j, 0) = (real)cd.mh;
            slots.R(j, 1) = (real)cd.c0;
            slots.R(j, 2) = (real)(-INFINITY);
            slots.R(j, 3) = (real)cd.xi[0];
            slots.R(j, 4) = (real)cd.xi[1];
            wlo[j] = cd.klo; whi[j] = cd.khi;
            uint32_t nm = cd.nmask, info = cd.info;
#ifndef PDEB_NO_INEL   /* A/B switch of tools/variant_bench.py */
            if ((info & 8u) && nm == 0u && ((info >> 4) & 31u) == ((info >> 9) & 31u))
#else
            if (false)
#endif
            {
                /* interval inside one element: the masked loop extracts the two node margins (pdeb_inelement_gmax) */
                const uint32_t e = (info >> 4) & 31u;
                nm = (1u << e) | (1u << (e + 1u));
                info = (info & ~8u) | (1u << 14);
            }
            PdebU2 ni;
            ni.x = nm; ni.y = info;
            slots.NI(j) = ni;
            const int32_t a1 = cd.khi;
            if (a1 > k_end) k_end = a1;
        }
        S32 = k_end;
        /* window-event list: the distinct boundaries k_lo, k_hi + 1 of all clauses in ascending order, each with the set
         * of windows open from that step on (selection by repeated minimum: at most 16 boundaries, set-up only) */
        int ne = 0;
        int32_t prev = -1;
        for (;;) {
            int32_t nxt = 0x7fffffff;
            for (int j = 0; j < np; ++j) {
                if (wlo[j] > prev && wlo[j] < nxt) nxt = wlo[j];
                if (whi[j] + 1 > prev && whi[j] + 1 < nxt) nxt = whi[j] + 1;
            }
            if (nxt == 0x7fffffff) break;
            uint32_t m = 0;
            for (int j = 0; j < np; ++j) if (nxt >= wlo[j] && nxt <= whi[j]) m |= 1u << j;
            slots.EV(ne++) = ((uint32_t)nxt << 9) | m;
            prev = nxt;
        }
        if (ne > 0) {
            slots.EV(ne - 1) |= 0x100u;
            evw = slots.EV(0);
            next_ev = (int32_t)(evw >> 9);
        }
#if !PDEB_EVENT_LIST
        for (int j = 0; j < np; ++j) { PdebU2 win; win.x = (uint32_t)wlo[j]; win.y = (uint32_t)whi[j]; slots.WIN(j) = win; }
        next_ev = 0;
#endif
    }

    real f_prev = in.eval(0.0, kt, ky, st);
    uint32_t open = 0;        /* bit j: step k is inside clause j's window */
    int cur = 0;              /* lowest open clause (0 when none is open) */
    /* what this lane needs from the pass of the current step: bit 0 a clause is open, bit 1 more than one is,
     * bit 2 / 3 the clause `cur` takes its end points from one element / in the general interpolated form.
     * It changes at window events only; each step ONE warp-wide OR of it decides every branch of the pass. */
    uint32_t flags = 0;
    /* sticky: the state was not finite at an OBSERVED step (status NONFINITE).  The running maxima below drop NaN
     * margins, so this is tested on the state itself: the step matrix couples every row (both elimination chains and
     * the back substitution run through non-zero multipliers), so one non-finite value (an inf / NaN control knot
     * reaching the load) makes EVERY row non-finite within that step, and it never heals: the twist row stands for
     * all of them */
    bool bad_state = false;
    const real eq_shift = (cv.eq_rule == 1) ? (real)cv.eq_eps : (real)0;

    /* window events of step k: which clauses are open from here on, and what the pass needs to know about them */
#if !PDEB_EVENT_LIST
    auto window_events = [&](int32_t k) {
        if (k == next_ev) {
            open = 0;
            next_ev = 0x7fffffff;
            for (int j = 0; j < np; ++j) {
                const PdebU2 win = slots.WIN(j);
                const int32_t a0 = (int32_t)win.x, a1 = (int32_t)win.y;
                if (k >= a0 && k <= a1) { open |= (1u << j); if (a1 + 1 < next_ev) next_ev = a1 + 1; }
                else if (k < a0 && a0 < next_ev) next_ev = a0;
            }
            cur = open ? __builtin_ffs((int)open) - 1 : 0;
            flags = 0;
            if (open) {
                const uint32_t info = slots.NI(cur).y;
                flags = 1u | ((open & (open - 1u)) ? 2u : 0u) | (((info >> 14) & 1u) << 2) | (info & 8u);
            }
        }
    };
    (void)evw; (void)evi;
#else
    auto window_events = [&](int32_t k) {
        if (k == next_ev) {
            open = evw & 0xffu;
            cur = open ? __builtin_ffs((int)open) - 1 : 0;
            flags = 0;
            if (open) {
                const uint32_t info = slots.NI(cur).y;
                flags = 1u | ((open & (open - 1u)) ? 2u : 0u) | (((info >> 14) & 1u) << 2) | (info & 8u);
            }
            if (evw & 0x100u) next_ev = 0x7fffffff;                     /* that was the lane's last event */
            else { evw = slots.EV(++evi); next_ev = (int32_t)(evw >> 9); }
        }
    };
#endif
    const int32_t Smax = PDEB_WARP_MAX(S32);     /* the warp steps until its longest lane is done */
    window_events(0);
    uint32_t wf = PDEB_WARP_OR(flags);
    constexpr int PRE_TAB = (KIND == PDEB_KIND_HEAT) ? PDEB_PRE_TAB_H : PDEB_PRE_TAB_W;
    constexpr bool PRE_SLOTS = (KIND == PDEB_KIND_HEAT) ? PDEB_PRE_SLOTS_H : PDEB_PRE_SLOTS_W;
    constexpr bool PRE_TIME = (KIND == PDEB_KIND_HEAT) ? PDEB_PRE_TIME_H : PDEB_PRE_TIME_W;
    static_assert(PRE_TAB >= 0 && PRE_TAB <= PDEB_PRE_TAB_MAX && PRE_TAB < (N + 1) / 2 - 1, "prefetched rows lie inside the two chains");
    real p_mh = slots.R(cur, 0), p_c0 = slots.R(cur, 1), p_A = slots.R(cur, 2);
    PdebU2 p_ni = slots.NI(cur);
    double t_nx = dt;

    /* ---- (1) reduce the open clauses on the state of the current step ---- */
    auto clause_pass = [&]() {
            /* every lane runs the pass on its clause `cur`; a lane without an open clause does so with an empty mask */
            const bool has = flags & 1u;
            const real mh = PRE_SLOTS ? p_mh : slots.R(cur, 0), c0 = PRE_SLOTS ? p_c0 : slots.R(cur, 1), A = PRE_SLOTS ? p_A : slots.R(cur, 2);
            const PdebU2 ni = PRE_SLOTS ? p_ni : slots.NI(cur);
            const uint32_t nm = has ? ni.x : 0u, info = ni.y;
            const uint32_t rel = info & 3u;
            const uint32_t and_mask = (rel == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
            const uint32_t xor_mask = (rel == PDEB_REL_GT) ? 0x80000000u : 0u;
            const bool inel = flags & 4u;            /* raw node margins of one element wanted: identity sign treatment */
            real hi[PDEB_MARGIN_ACC];
            pdeb_clause_partials<N, real, STRIDE>(tr.w, tr.stab, mh, c0, nm, inel ? 0xffffffffu : and_mask, inel ? 0u : xor_mask, hi);
            real gmax = pdeb_partials_max<real>(hi);
#if PDEB_ENDPOINTS_MERGED
            /* interpolated end points: ONE warp-wide test skips both forms (two thirds of the passes see neither) */
            if (wf & 12u) {
                if (wf & 4u) {
                    /* every lane runs the few selects; only the lanes whose interval lies inside one element keep the result */
                    const real gi = pdeb_inelement_gmax<real>(hi, (info >> 4) & 31u, slots.R(cur, 3), slots.R(cur, 4), and_mask, xor_mask);
                    gmax = inel ? gi : gmax;
                }
                if (wf & 8u) {
                    if (flags & 8u) {
                        const real ge = pdeb_endpoint_gmax<N, real, STRIDE>(tr.w, tr.stab, mh, c0, info, slots.R(cur, 3), slots.R(cur, 4),
                                                                            and_mask, xor_mask);
                        if (ge > gmax) gmax = ge;
                    }
                }
            }
#else
            if (wf & 4u) {
                if (inel) gmax = pdeb_inelement_gmax<real>(hi, (info >> 4) & 31u, slots.R(cur, 3), slots.R(cur, 4), and_mask, xor_mask);
            }
            if (wf & 8u) {
                if (flags & 8u) {
                    const real ge = pdeb_endpoint_gmax<N, real, STRIDE>(tr.w, tr.stab, mh, c0, info, slots.R(cur, 3), slots.R(cur, 4),
                                                                        and_mask, xor_mask);
                    if (ge > gmax) gmax = ge;
                }
            }
#endif
            /* G: min over steps of -gmax = -(max gmax); F: max over steps of -gmax: one running max either way */
            const real v = (((info >> 2) & 1u) == PDEB_TOP_G) ? gmax : -gmax;
            if (has && v > A) slots.R(cur, 2) = v;
            {
                const real wm = tr.w[PdebTrajReg<KIND, N, real, STRIDE>::M];
                if (has && !(wm - wm == (real)0)) bad_state = true;
            }
            if (wf & 2u) {
                /* lanes with several open windows: their remaining clauses, one pass per clause */
                uint32_t todo = (flags & 2u) ? (open & (open - 1u)) : 0u;
                while (PDEB_ANY(todo != 0u)) {
                    const bool has2 = todo != 0u;
                    int j = 0;
                    if (has2) { j = __builtin_ffs((int)todo) - 1; todo &= todo - 1u; }
                    const real mh2 = slots.R(j, 0), c02 = slots.R(j, 1);
                    const PdebU2 ni2 = slots.NI(j);
                    const uint32_t nm2 = has2 ? ni2.x : 0u, info2 = ni2.y;
                    const uint32_t rel2 = info2 & 3u;
                    const uint32_t am2 = (rel2 == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
                    const uint32_t xm2 = (rel2 == PDEB_REL_GT) ? 0x80000000u : 0u;
                    const bool inel2 = has2 && ((info2 >> 14) & 1u);
                    real hj[PDEB_MARGIN_ACC];
                    pdeb_clause_partials<N, real, STRIDE>(tr.w, tr.stab, mh2, c02, nm2, inel2 ? 0xffffffffu : am2, inel2 ? 0u : xm2, hj);
                    real g2 = pdeb_partials_max<real>(hj);
                    if (PDEB_ANY(inel2)) {
                        if (inel2) g2 = pdeb_inelement_gmax<real>(hj, (info2 >> 4) & 31u, slots.R(j, 3), slots.R(j, 4), am2, xm2);
                    }
                    if (PDEB_ANY(has2 && (info2 & 8u))) {
                        if (has2 && (info2 & 8u)) {
                            const real ge = pdeb_endpoint_gmax<N, real, STRIDE>(tr.w, tr.stab, mh2, c02, info2, slots.R(j, 3),
                                                                                slots.R(j, 4), am2, xm2);
                            if (ge > g2) g2 = ge;
                        }
                    }
                    const real v2 = (((info2 >> 2) & 1u) == PDEB_TOP_G) ? g2 : -g2;
                    if (has2 && v2 > slots.R(j, 2)) slots.R(j, 2) = v2;
                }
            }
            };
#if PDEB_ROTATE
    /* rotated loop: the pass of step k+1 closes the iteration that advanced to it (one branch per step less) */
    if (wf & 1u) clause_pass();
    for (int32_t k = 0; k < Smax; ++k) {
#else
    for (int32_t k = 0;; ++k) {
        if (wf & 1u) clause_pass();
        if (k >= Smax) break;
#endif
        typename PdebTrajReg<KIND, N, real, STRIDE>::real2 pre_t[PDEB_PRE_TAB_MAX], pre_b[PDEB_PRE_TAB_MAX];
        PDEB_UNROLL
        for (int q = 0; q < PRE_TAB; ++q) { pre_t[q] = tr.tab[q * STRIDE]; pre_b[q] = tr.tab[(N - 1 - q) * STRIDE]; }
        /* the windows of step k+1 are settled before the solve */
        window_events(k + 1);
        /* ---- (2) advance to step k+1 (lanes that are done keep stepping; nobody looks at them) ---- */
        {
            const real f_next = in.eval(PRE_TIME ? t_nx : (double)(k + 1) * dt, kt, ky, st);
            /* the warp-wide OR opens the block of the solve, whose length hides its latency (between the events and the
             * crossing it ended a short block, and the copy of its result out of the uniform register stalled there) */
            wf = PDEB_WARP_OR(flags);
            if (KIND == PDEB_KIND_HEAT) {
                const real th_dt = (real)(cv.theta * dt);
                const real fb = tr.sN * fma((real)dt - th_dt, f_prev, th_dt * f_next);
                real z[N + 2];
                z[0] = 0; z[N + 1] = 0;
                tr.template step_solve<PRE_TAB>(tr.w, fb, z, pre_t, pre_b);
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
                tr.template step_solve<PRE_TAB>(tr.w, tr.sN * f_next, tr.a, pre_t, pre_b);   /* a' = A^-1 (f' e_N - K d~), scaled */
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
        if (PRE_SLOTS) { p_mh = slots.R(cur, 0); p_c0 = slots.R(cur, 1); p_A = slots.R(cur, 2); p_ni = slots.NI(cur); }
        if (PRE_TIME) t_nx = (double)(k + 2) * dt;
#if PDEB_ROTATE
        if (wf & 1u) clause_pass();
#endif
    }

    if (!live) return;
    /* ---- clause values from the running maxima: G -A, F A; "=" under the band rule adds eq_eps ---- */
    for (int j = 0; j < np; ++j) {
        const uint32_t info = slots.NI(j).y;
        const real A = slots.R(j, 2);
        const real r = (((info >> 2) & 1u) == PDEB_TOP_G) ? -A : A;
        slots.R(j, 2) = ((info & 3u) == PDEB_REL_EQ) ? r + eq_shift : r;
    }
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
            if (!isfinite(rob) || bad_state) status = PDEB_ST_NONFINITE;
            for (int j = 0; j < np; ++j) if (!isfinite((double)slots.R(j, 2))) status = PDEB_ST_NONFINITE;
        }
    }
    /* s is the sample's position in the (plan-ordered) input image, so its index in the caller's order */
    out.robustness[so] = (status == PDEB_OK) ? rob : PDEB_FAILED_ROBUSTNESS;
    if (out.status) out.status[so] = status;
    if (out.clause) {
        for (int j = 0; j < b.pmax; ++j)
            out.clause[(int64_t)j * st + so] = (status == PDEB_OK && j < np) ? (double)slots.R(j, 2) : PDEB_FAILED_ROBUSTNESS;
    }
}

#endif /* PDEB_CORE_H */
