/*
 * pdeb200.h — C ABI of the B200-native utility-evaluation path of PDE-Controller.
 *
 * One call evaluates a BATCH of independent samples.  One sample = one 1D heat
 * or 1D wave finite-element forward simulation (femformal style: two-material
 * rod, Dirichlet at x=0, controlled flux/force at x=L) plus the quantitative
 * robustness of one STL specification over the simulated field.
 *
 * What each entry point replaces in the reference (paths relative to the
 * reference tree):
 *
 *   pdeb_eval_batch          the per-sample chain
 *                              eval_robustness_helper  test/PDEcontrol/evaluation/eval/eval_script.py:43-79
 *                              -> wrapper main         test/PDEcontrol/evaluation/eval/eval_robustness_wrapper.py:13-21
 *                              -> control.femformal.eval_robustness.eval_robustness   (call site wrapper:16)
 *                            fanned out one process per sample by
 *                              evaluate_future         test/PDEcontrol/evaluation/infer/simulate_gt.py:62-89
 *   PdebBatch                the femformal program dialect that chain receives as text
 *                              heatlinfem.heatlinfem_mix(xpart, rho, E, g, f_nodal, dt)
 *                              mechlinfem.mechlinfem(xpart, rho, E, g, f_nodal, dt)
 *                              logic.APCont([x_lo,x_hi], rel, p, dp), cregions, cspec
 *                            (call sites: utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3,
 *                             one_d_wave/examples.jsonl:1-3), already parsed to numbers.
 *   robustness_out/status    the {"robustness": r, "runtime": t} line of wrapper:18-21 and the
 *                            -100 failure label of simulate_gt.py:55-59.
 *
 * All pointers are caller-owned.  Nothing is retained after a call returns
 * (except inside an explicit PdebResident handle).  No torch types.
 *
 * Layout: structure of arrays.  Per-sample scalars are arrays of length n.
 * Per-sample lists (control knots, predicates, formula-tree nodes) are
 * PLANE-MAJOR: element j of sample s lives at  base[j*stride + s]  so that
 * consecutive samples are consecutive in memory (coalesced on the device).
 */
#ifndef PDEB200_H
#define PDEB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDEB_ABI_VERSION 1

/* kind */
#define PDEB_KIND_HEAT 0   /* rho*c u_t = (E u_x)_x   first order,  heatlinfem_mix */
#define PDEB_KIND_WAVE 1   /* rho u_tt  = (E u_x)_x   second order, mechlinfem     */

/* pred_rel */
#define PDEB_REL_GT 0      /* u(x) - (m x + b) > 0 */
#define PDEB_REL_LT 1      /* u(x) - (m x + b) < 0 */
#define PDEB_REL_EQ 2      /* u(x) - (m x + b) = 0 */

/* pred_top (temporal operator of the clause) */
#define PDEB_TOP_G 0       /* always   : min over the window */
#define PDEB_TOP_F 1       /* eventually: max over the window */

/* tree_op; nodes are stored children-before-parent, root = node n_tree-1 */
#define PDEB_TREE_LEAF 0   /* lhs = predicate (clause) index */
#define PDEB_TREE_AND  1   /* min(lhs, rhs), operands are earlier tree nodes */
#define PDEB_TREE_OR   2   /* max(lhs, rhs) */

/* status_out */
#define PDEB_OK            0
#define PDEB_ST_PARSE      1   /* set by the host front-end only (program text not understood) */
#define PDEB_ST_NONFINITE  2   /* simulation or robustness produced inf/nan */
#define PDEB_ST_BADPARAM   3   /* mesh / material / step parameters out of the supported range */

#define PDEB_FAILED_ROBUSTNESS (-100.0)  /* simulate_gt.py:59 */

/* limits of this build */
#define PDEB_MAX_PRED   8
#define PDEB_MAX_TREE   (2 * PDEB_MAX_PRED - 1)
#define PDEB_MAX_ELEM   1054          /* heat; 32 lanes x 32 rows + 31 separators - the Dirichlet node (waves: 542) */
#define PDEB_MAX_STEPS  (1 << 22)

/* return codes */
#define PDEB_E_OK        0
#define PDEB_E_ARG      -1
#define PDEB_E_CUDA     -2
#define PDEB_E_NOMEM    -3
#define PDEB_E_NODEVICE -4

/*
 * Convention flags.  The reference's femformal source
 * is not public in the snapshot, so every choice that moves the result is an
 * explicit parameter.  pdeb_default_conventions() fills the "v0" defaults.
 */
typedef struct PdebConventions {
    int32_t material_sampling; /* C1: 0 element midpoint (v0), 1 left node, 2 right node */
    int32_t mass_heat;         /* C2: 0 consistent (v0), 1 lumped */
    int32_t mass_wave;         /* C2: 0 consistent, 1 lumped (v0) */
    int32_t steps_rule;        /* C6: 0 ceil(T/dt - 1e-9) (v0), 1 round, 2 floor(T/dt + 1e-9) */
    int32_t window_rule;       /* C7: 0 [ceil, floor] inclusive (v0), 1 int() truncation of both ends */
    int32_t region_rule;       /* C8: 0 nodes inside, end-point interpolation only if none (v0);
                                      1 nodes inside plus both interpolated end points */
    int32_t eq_rule;           /* C9: 0 -|u-p| (v0), 1 band: min(u-(p-eps), (p+eps)-u) */
    int32_t reserved0;
    double  eq_eps;            /* C9 band half width, used when eq_rule == 1 */
    double  theta;             /* C4: heat generalised-trapezoidal parameter, v0 = 0.5 */
    double  newmark_beta;      /* C5: v0 = 0.25 */
    double  newmark_gamma;     /* C5: v0 = 0.5  */
} PdebConventions;

typedef struct PdebBatch {
    int64_t n;        /* number of samples */
    int64_t stride;   /* plane stride of the 2-D arrays, >= n */
    int32_t kmax;     /* planes in knot_t / knot_y */
    int32_t pmax;     /* planes in pred_*  (<= PDEB_MAX_PRED) */
    int32_t tmax;     /* planes in tree_*  (<= PDEB_MAX_TREE) */
    int32_t reserved0;

    /* optional: if all arrays below live in one contiguous host block, give it
       here and the library moves it with a single copy; else NULL / 0 */
    const void* slab_base;
    int64_t     slab_bytes;

    const uint8_t* kind;     /* [n] PDEB_KIND_* */
    const int32_t* n_elem;   /* [n] N: elements; nodes x_i = i L/N, i = 0..N  (xpart = linspace(0,L,N+1)) */
    const double*  L;        /* [n] rod length */
    const double*  dt;       /* [n] time step */
    const double*  T;        /* [n] horizon */
    const double*  g0;       /* [n] Dirichlet value at x = 0   (g = [g0, None]) */
    const double*  x_a;      /* [n] inner segment is x_a <= x <= x_b, outer is x < x_a or x > x_b */
    const double*  x_b;
    const double*  rho_in;   /* [n] rho(x) on the inner segment (heat: rho*c product) */
    const double*  rho_out;
    const double*  E_in;
    const double*  E_out;

    const int32_t* n_knots;  /* [n] 0 = zero input */
    const double*  knot_t;   /* [kmax][stride] ascending knot times of the PWL input at x = L;
                                NULL = uniform knots linspace(0, T, n_knots) (convention C11), formed on the device:
                                120 of the ~500 bytes per sample of the default heat layout need not cross PCIe */
    const double*  knot_y;   /* [kmax][stride] knot values (heat: flux, wave: force), added to F[N] */

    const int32_t* n_pred;   /* [n] clauses, 1..pmax */
    const double*  pred_xlo; /* [pmax][stride] APCont interval */
    const double*  pred_xhi;
    const double*  pred_m;   /* profile p(x) = m x + b */
    const double*  pred_b;
    const double*  pred_tlo; /* temporal window of the clause */
    const double*  pred_thi;
    const uint8_t* pred_rel; /* PDEB_REL_* */
    const uint8_t* pred_top; /* PDEB_TOP_* */

    const int32_t* n_tree;   /* [n] nodes, 1..tmax */
    const uint8_t* tree_op;  /* [tmax][stride] */
    const int8_t*  tree_lhs;
    const int8_t*  tree_rhs;
} PdebBatch;

/* timing detail of one call, all in milliseconds, filled when non-NULL */
typedef struct PdebTiming {
    float h2d_ms;
    float kernel_ms;     /* CUDA events around the kernel launches only */
    float d2h_ms;
    float total_ms;      /* events around h2d + kernels + d2h */
    int32_t launches;    /* kernels launched by this call */
    int32_t reserved0;
} PdebTiming;

int  pdeb_version(void);
int  pdeb_device_count(void);
const char* pdeb_last_error(void);       /* thread-local, valid until the next call on this thread */
/* "dddd:bb:dd.f" of a device (cudaDeviceGetPCIBusId), so that a host process can place itself and its page-locked
   buffers on the NUMA node the device hangs off (native.bind_host_near; the reference's process pool,
   simulate_gt.py:62-89, leaves placement to the OS).  len >= 13. */
int  pdeb_device_pci_bus_id(int device, char* buf, int len);
void pdeb_default_conventions(PdebConventions* out);

/*
 * Evaluate a host batch on one device: H2D, kernels, D2H.
 *   precision       64 or 32 (arithmetic of the time loop; inputs/outputs stay f64)
 *   conv            NULL = v0 defaults
 *   robustness_out  [n]          formula robustness, PDEB_FAILED_ROBUSTNESS where status != 0
 *   clause_out      [pmax][stride] per-clause robustness, or NULL
 *   status_out      [n] or NULL
 *   stream          cudaStream_t or NULL (default stream); the call synchronises it before returning
 */
int pdeb_eval_batch(const PdebBatch* in, int device, int precision, const PdebConventions* conv,
                    double* robustness_out, double* clause_out, int32_t* status_out,
                    void* stream, PdebTiming* timing);

/*
 * Same over several devices of one box: contiguous shards of equal COST (time steps x mesh rows, x 1.5 for waves:
 * SURVEY.md section 8e; pdeb_shard_bounds gives the cut points), one host thread and one stream per device, every shard
 * through its device's page-locked staging mirror (one H2D copy), results written straight to the samples' own indices
 * of the caller's arrays: no inter-device traffic, nothing to gather.  devices == NULL means 0..ndev-1.
 * Replaces the process pool of evaluate_future (simulate_gt.py:62-89 of the reference).
 */
int pdeb_shard_bounds(const PdebBatch* in, int ndev, const PdebConventions* conv, int64_t* bounds /* [ndev + 1] */);
/* one-shot calls keep small resident images per device and shape (no allocation on the next call of that shape);
   this frees them */
void pdeb_cache_clear(void);
int pdeb_eval_batch_multi(const PdebBatch* in, int ndev, const int* devices, int precision,
                          const PdebConventions* conv, double* robustness_out, double* clause_out,
                          int32_t* status_out, PdebTiming* timing_per_device);

/*
 * Resident form: upload once, evaluate many times (used to time the kernels with
 * inputs already in HBM, and by the many-controls search that rewrites only the knots).
 */
typedef struct PdebResident PdebResident;

int  pdeb_resident_create(const PdebBatch* in, int device, void* stream, PdebResident** out);
int  pdeb_resident_upload(PdebResident* r, const PdebBatch* in, void* stream);  /* same shape: refill the image */

/*
 * Packed TRANSFER form of a batch.  The end-to-end rate of this path is host->device bandwidth divided by the bytes per
 * sample (the kernels wait for PCIe, not the other way round), and a third of the plane-major layout is slots a sample
 * does not use (knot planes beyond n_knots, clause planes beyond n_pred, tree planes beyond n_tree).  Here every
 * per-sample LIST is ragged: sample after sample, only the entries in use; the small integers are narrow.  The default
 * heat set is 280 instead of 382 bytes per sample.  The library copies the block with ONE transfer, forms the list
 * offsets on the device (prefix sums of the counts) and scatters the lists into the plane-major resident image, so the
 * kernels, the plan and every other entry point are untouched.  What it carries is what eval_robustness_helper's two
 * strings carry in the reference (eval/eval_script.py:43-79), as numbers.
 */
typedef struct PdebPackedBatch {
    int64_t n;
    int32_t kmax, pmax, tmax, reserved0;             /* list bounds: the shape of the resident image this refills */
    int64_t total_knots, total_pred, total_tree;      /* lengths of the flat arrays: sums of the counts */
    const void* slab_base;                            /* REQUIRED: every array below inside one contiguous host block */
    int64_t     slab_bytes;                           /* (page-locked for an asynchronous copy: pdeb_host_alloc) */
    const uint8_t*  kind;                             /* [n] */
    const uint16_t* n_elem;                           /* [n] */
    /* real-valued arrays: stored as enc[] says (below); the value every consumer sees is the f64 it encodes, bit for bit */
    const void *L, *dt, *T, *g0, *x_a, *x_b, *rho_in, *rho_out, *E_in, *E_out;      /* [n] each */
    const uint16_t* n_knots;                          /* [n] */
    const uint8_t*  n_pred;                           /* [n] */
    const uint8_t*  n_tree;                           /* [n] */
    const void* knot_t;                               /* [total_knots], or NULL = uniform knots (as at creation) */
    const void* knot_y;                               /* [total_knots] */
    const void *pred_xlo, *pred_xhi, *pred_m, *pred_b, *pred_tlo, *pred_thi;        /* [total_pred] each */
    const uint8_t *pred_rel, *pred_top;               /* [total_pred] */
    const uint8_t* tree_op;                           /* [total_tree] */
    const int8_t *tree_lhs, *tree_rhs;                /* [total_tree] */
    /* LOSSLESS storage of each real-valued array, index PDEB_PK_*: PDEB_ENC_F64 as is; PDEB_ENC_F32 as float, allowed
       only if every value is exactly representable (lengths, interval ends, Dirichlet values are integers in the
       reference's data); PDEB_ENC_ONE a single f64 standing for every entry (a set sharing one time step) */
    uint8_t enc[24];
} PdebPackedBatch;
#define PDEB_ENC_F64 0
#define PDEB_ENC_F32 1
#define PDEB_ENC_ONE 2
enum { PDEB_PK_L = 0, PDEB_PK_DT, PDEB_PK_T, PDEB_PK_G0, PDEB_PK_XA, PDEB_PK_XB, PDEB_PK_RHO_IN, PDEB_PK_RHO_OUT, PDEB_PK_E_IN,
       PDEB_PK_E_OUT, PDEB_PK_KNOT_T, PDEB_PK_KNOT_Y, PDEB_PK_XLO, PDEB_PK_XHI, PDEB_PK_M, PDEB_PK_B, PDEB_PK_TLO, PDEB_PK_THI,
       PDEB_PK_COUNT };

/*
 * Host-side producer of the packed form (no device involved): packs the plane-major batch `in` into `slab`
 * (`capacity` bytes, 8-byte aligned; page-locked if it is to be uploaded asynchronously) and fills `*out` with pointers
 * into it, choosing each real-valued array's storage tag by bit comparison.  pdeb_packed_capacity is the worst case of a
 * shape, so one buffer serves every batch of that shape.  Returns PDEB_E_ARG if a count does not fit the narrow
 * integers or the slab is too small.
 */
int64_t pdeb_packed_capacity(int64_t n, int32_t kmax, int32_t pmax, int32_t tmax, int32_t with_knot_t);
int  pdeb_pack_batch(const PdebBatch* in, void* slab, int64_t capacity, PdebPackedBatch* out);

/* refill the image of `r` (same n / kmax / pmax / tmax) from the packed form; one H2D copy + 3 small launches + the plan */
int  pdeb_resident_upload_packed(PdebResident* r, const PdebPackedBatch* in, void* stream);
/* upload + evaluate + fetch of a batch of the resident shape in one pipelined pass (H2D / kernels / D2H of
   successive chunks overlap on two internal streams); synchronises before returning */
int  pdeb_resident_run(PdebResident* r, const PdebBatch* in, int precision, const PdebConventions* conv,
                       double* robustness_out, double* clause_out, int32_t* status_out, void* stream, PdebTiming* timing);
int  pdeb_resident_eval(PdebResident* r, int precision, const PdebConventions* conv,
                        void* stream, PdebTiming* timing);   /* asynchronous on stream unless timing != NULL */
int  pdeb_resident_set_knots(PdebResident* r, const double* knot_t, const double* knot_y,
                             const int32_t* n_knots, void* stream);      /* host planes, same kmax/stride */
int  pdeb_resident_fetch(PdebResident* r, double* robustness_out, double* clause_out,
                         int32_t* status_out, void* stream);
/* the same copies without the final synchronisation (pair with pdeb_stream_sync) */
int  pdeb_resident_fetch_async(PdebResident* r, double* robustness_out, double* clause_out,
                               int32_t* status_out, void* stream);
/* library-owned non-blocking streams, for callers that do not bring their own (native.BlockPipeline) */
void* pdeb_stream_create(int device);
void  pdeb_stream_destroy(void* stream);
int   pdeb_stream_sync(void* stream);
/* the order the kernels walk the batch in (sample ids grouped by kernel variant, longest trajectories first), as formed
   under the conventions of the last evaluation: diagnostics and tests of the plan.  order_out [n]; synchronises. */
int  pdeb_resident_plan_order(PdebResident* r, int32_t* order_out, void* stream);
/* device pointers of the results (f64[n], f64[pmax][stride], i32[n]) for zero-copy consumers */
int  pdeb_resident_device_ptrs(PdebResident* r, void** robustness_dev, void** clause_dev, void** status_dev);
void pdeb_resident_destroy(PdebResident* r);

/*
 * Many candidate controls for ONE specification (north star "many candidate controls per spec").
 * In the reference the boundary input is synthesised by Gurobi inside control.femformal
 * (eval_robustness_wrapper.py:16); that MILP is out of scope, and this is its sampling surrogate: the robustness
 * of every one of m piecewise-linear inputs for sample `sample` of `spec`, and the arg-max.
 *   knot_t   [n_knots] knot times shared by all candidates (C11), ascending; the spec's own knots are ignored
 *   knot_y   host [n_knots][m], plane-major: value j of candidate i at knot_y[j*m + i]; NULL = values already there
 *   outputs  robustness [m], clause [n_pred][m], status [m], best index / value: all optional (NULL)
 * Default meshes only (heat N=30, wave N=20), FP64; other meshes go through pdeb_eval_batch with a replicated spec.
 */
typedef struct PdebControls PdebControls;
int  pdeb_controls_create(const PdebBatch* spec, int64_t sample, int64_t capacity, int32_t n_knots, const double* knot_t,
                          int device, const PdebConventions* conv, PdebControls** out);
int  pdeb_controls_eval(PdebControls* h, int64_t m, const double* knot_y, double* robustness_out, double* clause_out,
                        int32_t* status_out, int64_t* best_index, double* best_value, void* stream, PdebTiming* timing);
void pdeb_controls_destroy(PdebControls* h);
/* point an existing handle at another specification of the family (heat N=30 / wave N=20, up to 32 knots): no allocation,
   the candidates on the device stay.  What a sweep over many programs calls between searches. */
int  pdeb_controls_respec(PdebControls* h, const PdebBatch* spec, int64_t sample, int32_t n_knots, const double* knot_t,
                          const PdebConventions* conv);
/*
 * The whole cross-entropy search for the handle's specification on the device (the sampling surrogate of the reference's
 * control synthesis, eval_robustness_wrapper.py:16): `generations` times { draw `population` candidates
 * y = clip(mean + std z, lo, hi) from a counter-based generator (generation 0: uniform over [lo, hi], plus the zero input
 * and both bounds; the incumbent is always candidate 0) -> their robustness -> the elite_frac best re-centre mean / std }.
 * ONE launch per generation (every thread draws and scores its candidate, the last block to finish ranks), nothing crosses
 * PCIe in between, deterministic for a given seed.
 *   best_y_out [n_knots], best_value_out, found_out (0: no candidate had status 0; then best_value = -100),
 *   history_out [generations] best robustness after each generation (never decreases), mean_out / std_out [n_knots]
 *   the final sampling distribution: all optional.  Synchronises `stream` before returning.
 */
int  pdeb_controls_search(PdebControls* h, int64_t population, int32_t generations, double elite_frac, double lo, double hi,
                          uint64_t seed, double* best_y_out, double* best_value_out, int32_t* found_out, double* history_out,
                          double* mean_out, double* std_out, void* stream, PdebTiming* timing);
/*
 * pdeb_controls_search for MANY specifications: samples[0..n_samples) of `in` (NULL = every sample), `nthreads` host threads
 * (<= 0: 8; at most 32), each with three searches in flight with one stream and one controls handle each.  Knot TIMES are the batch's own (n_knots, knot_t;
 * knot_t == NULL: linspace(0, T, n_knots)), the knot VALUES found are written to knot_y_out in the batch's plane-major
 * layout [kmax][stride] (it may BE in->knot_y), best_out [n] / found_out [n] per sample.  bounds = { heat lo, heat hi, wave lo,
 * wave hi }.  Sample s draws from the generator stream mix(seed, s): the result does not depend on the thread count.
 * A sample outside the controls family (or with more than 32 knots) is left with found = 0 for the caller to handle.
 * Replaces one femformal + Gurobi process per sample (simulate_gt.py:62-89 -> eval_robustness_wrapper.py:16).
 */
int  pdeb_controls_search_batch(const PdebBatch* in, const int64_t* samples, int64_t n_samples, int64_t population,
                                int32_t generations, double elite_frac, const double* bounds, uint64_t seed, int device,
                                const PdebConventions* conv, int nthreads, double* knot_y_out, double* best_out,
                                int32_t* found_out, PdebTiming* timing);
/* pdeb_controls_search_batch keeps the handles of a finished sweep for the next one (per device); this frees them */
void pdeb_controls_pool_clear(void);
/* the candidates currently on the device, host [n_knots][m] plane-major (diagnostics, tests) */
int  pdeb_controls_fetch_knots(PdebControls* h, int64_t m, double* knot_y_out, void* stream);

/*
 * Native front-end for the CANONICAL program template (host only).  The reference's call surface is program TEXT
 * (eval_robustness_helper(program, nl), eval/eval_script.py:43-79), and every program of its records is the same dozen
 * lines with different numbers (utils/few_shot_prompts/examples/one_d_heat/examples.jsonl:1-3, ...).  Program i of `texts`
 * (`lengths[i]` bytes, need not be NUL-terminated) that IS that template is parsed, never executed, straight into sample
 * first + i of `out` (the PdebBatch arrays must be writable host memory; n_knots / knot_* are not touched: the control
 * input is the caller's) and accepted[i] = 1; any other text leaves accepted[i] = 0 and is the business of the general
 * front-end (pde_controller_b200/frontend.py), which this matcher never contradicts.  nthreads <= 0: all host threads.
 */
int  pdeb_parse_programs(const char* const* texts, const int64_t* lengths, int64_t n, const PdebBatch* out, int64_t first,
                         uint8_t* accepted, int nthreads);

/*
 * Boolean grids of the reference's IoU metric between two sSTL formulas (eval_IOU, eval/eval_script.py:233-281;
 * check_conditions :437-487; G_array / F_array :159-219), batched: for every clause and every linear test profile
 * u(t) = m t + b, the bit per time step the reference's bit array holds.  The caller (pde_controller_b200/iou.py) parses
 * the formulas, forms the time / position grids with the reference's own numpy calls and combines the clauses' arrays;
 * this entry evaluates the ~1.4 M cells per clause, which the reference does with one Python eval() each.
 *   clauses   one record per clause; `item` selects the grid set (one per formula pair)
 *   tgrid     [n_items][2][tmax]  times: direction 0 = forwards (G clauses), 1 = backwards (F clauses), in iteration order
 *   xgrid     [n_items][2][xmax]  positions, likewise
 *   ngrid     [n_items][4]        lengths { nt forwards, nt backwards, nx forwards, nx backwards };  tmax <= 64
 *   m_vals, b_vals                the profile grid; profile index = im * nb + ib  (itertools.product(m, b))
 *   bits_out  [n_clauses][nm * nb][tmax]  0/1 per profile and time, times ASCENDING for both clause types
 * Bit-exact with the reference: IEEE double operations in Python's order, no contraction.
 */
typedef struct PdebIouClause {
    double t0, t1;          /* time interval of the clause */
    double p0, p1;          /* position interval (the reference truncates its ends to integers) */
    double coef, c, thr;    /* expr(x) = coef * x + c (PDEB_IOU_MUL) or x / coef + c (PDEB_IOU_DIV); no profile: coef = c = 0 */
    int32_t item;
    uint8_t kind, neg, op, backwards;   /* neg: x enters negated ("-x / 8"); op: PDEB_IOU_*; backwards: 1 for F clauses */
} PdebIouClause;
#define PDEB_IOU_MUL 0
#define PDEB_IOU_DIV 1
enum { PDEB_IOU_LE = 0, PDEB_IOU_LT, PDEB_IOU_GE, PDEB_IOU_GT, PDEB_IOU_EQ, PDEB_IOU_NE };
int  pdeb_iou_grids(const PdebIouClause* clauses, int64_t n_clauses, const double* tgrid, const double* xgrid,
                    const int32_t* ngrid, int64_t n_items, int32_t tmax, int32_t xmax, const double* m_vals, int32_t nm,
                    const double* b_vals, int32_t nb, uint8_t* bits_out, int device, void* stream, PdebTiming* timing);

/* pinned host memory for batches that are uploaded every step */
void* pdeb_host_alloc(int64_t bytes);
void  pdeb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PDEB200_H */
