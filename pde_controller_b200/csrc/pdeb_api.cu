/*
 * pdeb_api.cu — the C ABI of include/pdeb200.h: device memory, copies, launches, sharding.
 *
 * Replaces the process-per-sample fan-out of the reference
 * (evaluate_future, test/PDEcontrol/evaluation/infer/simulate_gt.py:62-89, around
 *  eval_robustness_helper, eval/eval_script.py:43-79) by one batched call.
 * There is no CPU path in this library: without a CUDA device every entry point fails.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pdeb200.h"
#include "pdeb_core.h"
#include "pdeb_launch.h"

static thread_local std::string g_err;

static int fail(int code, const char* what, cudaError_t ce = cudaSuccess) {
    char buf[512];
    if (ce != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
    else snprintf(buf, sizeof buf, "%s", what);
    g_err = buf;
    return code;
}

#define CK(call)                                                         \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) return fail(PDEB_E_CUDA, #call, _e);      \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void pdeb_set_last_error(const char* msg) { g_err = msg ? msg : ""; }

extern "C" int pdeb_version(void) { return PDEB_ABI_VERSION; }

extern "C" int pdeb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* pdeb_last_error(void) { return g_err.c_str(); }

extern "C" int pdeb_device_pci_bus_id(int device, char* buf, int len) {
    if (!buf || len < 13) return fail(PDEB_E_ARG, "pdeb_device_pci_bus_id: buffer of at least 13 bytes wanted");
    if (device < 0 || device >= pdeb_device_count()) return fail(PDEB_E_NODEVICE, "no such CUDA device");
    const cudaError_t e = cudaDeviceGetPCIBusId(buf, len, device);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(PDEB_E_CUDA, cudaGetErrorString(e)); }
    return 0;
}

extern "C" void pdeb_default_conventions(PdebConventions* c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
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

extern "C" void* pdeb_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void pdeb_host_free(void* p) { if (p) cudaFreeHost(p); }

/* ---- resident batch ------------------------------------------------------ */
struct PdebResident;
static int fetch_range_fwd(PdebResident* r, double* robustness_out, double* clause_out, int32_t* status_out, cudaStream_t stream);

struct PdebResident {
    int device = 0;
    int64_t n = 0, stride = 0, host_stride = 0;
    int kmax = 0, pmax = 0, tmax = 0;
    unsigned char* slab = nullptr;   /* inputs as uploaded: the caller's order (and, with a host slab, the caller's layout) */
    size_t slab_bytes = 0;
    unsigned char* img = nullptr;    /* the same inputs in PLAN ORDER (dimg): consecutive lanes read consecutive addresses.
                                        Formed by pdeb_permute when an upload is evaluated a SECOND time (the permutation costs
                                        more than the gathers of one evaluation: a batch that passes through once is read where
                                        it lies, a resident batch that is evaluated again and again is permuted once) */
    PdebDevBatch dimg{};
    bool img_valid = false;          /* dimg holds the current upload in the current order */
    int evals_since_upload = 0;
    unsigned char* outs = nullptr;   /* robustness | clause | status */
    PdebDevBatch dev{};
    PdebDevOut out{};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool host_slab = false;          /* device slab mirrors the caller's host slab byte for byte */
    unsigned char* plan = nullptr;   /* keys | order | PDEB_MAX_CHUNKS x (hist | vstart_count) */
    PdebPlanBuffers pb{};            /* of the whole batch (chunk 0's histogram) */
    int32_t* hist[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t vsc_host[2 * PDEB_NUM_VARIANTS] = {0};   /* first index / count of each variant inside `order` (whole batch) */
    int plan_launches = 0;
    bool plan_stale = false;         /* the last run planned chunk by chunk: whole-batch order must be rebuilt */
    int plan_steps_rule = 0, plan_window_rule = 0;   /* conventions the current order was formed under (they decide the
                                                        last observed step and where windows open: the plan's keys) */
    unsigned char* pstage = nullptr; /* packed uploads: the transferred block | list offsets | block sums (lazy) */
    size_t pstage_cap = 0, pstage_off = 0;
    size_t soff[28] = {0};           /* byte offset of every input array inside `slab` */
    unsigned char* hstage = nullptr; /* page-locked host mirror of `slab` (compact layout): batches that are not one pinned slab
                                        of this shape (shards, pageable memory, strided views) are gathered here, so that an
                                        upload is ONE asynchronous copy whatever the caller's layout (lazy) */
    cudaStream_t pipe[2] = {nullptr, nullptr};        /* internal streams of the pipelined run */
    cudaEvent_t pev[3] = {nullptr, nullptr, nullptr};
};
#define PDEB_MAX_CHUNKS 4
#define PDEB_PIPELINE_MIN 16384     /* samples per chunk below which splitting stops paying */

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int check_batch(const PdebBatch* in) {
    if (!in) return fail(PDEB_E_ARG, "batch is NULL");
    if (in->n < 0 || in->stride < in->n) return fail(PDEB_E_ARG, "bad n / stride");
    if (in->kmax < 1 || in->pmax < 1 || in->pmax > PDEB_MAX_PRED || in->tmax < 1 || in->tmax > PDEB_MAX_TREE)
        return fail(PDEB_E_ARG, "kmax / pmax / tmax out of range");
    if (in->n == 0) return 0;
    /* knot_t may be NULL: uniform knots linspace(0, T, n_knots) (C11) */
    const void* req[] = {in->kind, in->n_elem, in->L, in->dt, in->T, in->g0, in->x_a, in->x_b, in->rho_in, in->rho_out,
                         in->E_in, in->E_out, in->n_knots, in->knot_y, in->n_pred, in->pred_xlo, in->pred_xhi,
                         in->pred_m, in->pred_b, in->pred_tlo, in->pred_thi, in->pred_rel, in->pred_top, in->n_tree,
                         in->tree_op, in->tree_lhs, in->tree_rhs};
    for (const void* p : req) if (!p) return fail(PDEB_E_ARG, "a batch array is NULL");
    return 0;
}

struct Field { const void* host; size_t elem; int planes; size_t offset; };

extern "C" void pdeb_resident_destroy(PdebResident* r) {
    if (!r) return;
    DeviceGuard g(r->device);
    for (auto& e : r->ev) if (e) cudaEventDestroy(e);
    if (r->slab) cudaFree(r->slab);
    if (r->img) cudaFree(r->img);
    if (r->outs) cudaFree(r->outs);
    if (r->plan) cudaFree(r->plan);
    if (r->pstage) cudaFree(r->pstage);
    if (r->hstage) cudaFreeHost(r->hstage);
    for (auto& st : r->pipe) if (st) cudaStreamDestroy(st);
    for (auto& e : r->pev) if (e) cudaEventDestroy(e);
    delete r;
}

/* host -> device copy of samples [lo, hi) of every input array (or of the knot arrays only) */
static int upload(PdebResident* r, const PdebBatch* in, int64_t lo, int64_t hi, cudaStream_t stream, bool knots_only) {
    const int64_t n = hi - lo;
    if (n <= 0) return 0;
    PdebDevBatch& d = r->dev;
    const bool whole = lo == 0 && hi == r->n;
    const bool fast = whole && !knots_only && r->host_slab && in->slab_base && in->slab_bytes > 0 && in->stride == r->n &&
                      (size_t)in->slab_bytes == r->slab_bytes && ((in->knot_t != nullptr) == (r->dev.knot_t != nullptr)) &&
                      (const unsigned char*)in->L - (const unsigned char*)in->slab_base ==
                          (const unsigned char*)r->dev.L - r->slab;
    if (fast) {
        CK(cudaMemcpyAsync(r->slab, in->slab_base, r->slab_bytes, cudaMemcpyHostToDevice, stream));
        return 0;
    }
    auto copy = [&](const void* dst, const void* src, size_t elem, int planes) -> cudaError_t {
        unsigned char* dp = (unsigned char*)dst + elem * (size_t)lo;
        const unsigned char* sp = (const unsigned char*)src + elem * (size_t)lo;
        if (planes == 1) return cudaMemcpyAsync(dp, sp, elem * (size_t)n, cudaMemcpyHostToDevice, stream);
        if (whole && in->stride == r->stride && r->stride == r->n)
            return cudaMemcpyAsync(dp, sp, elem * (size_t)n * planes, cudaMemcpyHostToDevice, stream);
        return cudaMemcpy2DAsync(dp, elem * r->stride, sp, elem * in->stride, elem * n, planes, cudaMemcpyHostToDevice, stream);
    };
    CK(copy(d.n_knots, in->n_knots, 4, 1));
    if ((in->knot_t != nullptr) != (d.knot_t != nullptr)) return fail(PDEB_E_ARG, "knot_t must be given (or NULL) as at creation");
    if (d.knot_t) CK(copy(d.knot_t, in->knot_t, 8, r->kmax));
    CK(copy(d.knot_y, in->knot_y, 8, r->kmax));
    if (knots_only) return 0;
    CK(copy(d.kind, in->kind, 1, 1));
    CK(copy(d.n_elem, in->n_elem, 4, 1));
    CK(copy(d.L, in->L, 8, 1));
    CK(copy(d.dt, in->dt, 8, 1));
    CK(copy(d.T, in->T, 8, 1));
    CK(copy(d.g0, in->g0, 8, 1));
    CK(copy(d.x_a, in->x_a, 8, 1));
    CK(copy(d.x_b, in->x_b, 8, 1));
    CK(copy(d.rho_in, in->rho_in, 8, 1));
    CK(copy(d.rho_out, in->rho_out, 8, 1));
    CK(copy(d.E_in, in->E_in, 8, 1));
    CK(copy(d.E_out, in->E_out, 8, 1));
    CK(copy(d.n_pred, in->n_pred, 4, 1));
    CK(copy(d.pred_xlo, in->pred_xlo, 8, r->pmax));
    CK(copy(d.pred_xhi, in->pred_xhi, 8, r->pmax));
    CK(copy(d.pred_m, in->pred_m, 8, r->pmax));
    CK(copy(d.pred_b, in->pred_b, 8, r->pmax));
    CK(copy(d.pred_tlo, in->pred_tlo, 8, r->pmax));
    CK(copy(d.pred_thi, in->pred_thi, 8, r->pmax));
    CK(copy(d.pred_rel, in->pred_rel, 1, r->pmax));
    CK(copy(d.pred_top, in->pred_top, 1, r->pmax));
    CK(copy(d.n_tree, in->n_tree, 4, 1));
    CK(copy(d.tree_op, in->tree_op, 1, r->tmax));
    CK(copy(d.tree_lhs, in->tree_lhs, 1, r->tmax));
    CK(copy(d.tree_rhs, in->tree_rhs, 1, r->tmax));
    return 0;
}

/* is this host pointer page-locked (cudaHostAlloc / cudaHostRegister)?  Pageable memory cannot be copied asynchronously. */
static bool host_is_pinned(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

/*
 * Upload of samples [lo, hi) of ANY host batch (a shard, pageable memory, a strided view) into a resident image of the
 * compact layout with n = hi - lo: the arrays are gathered into the image's page-locked host mirror (plain memcpy, one per
 * array and plane) and cross PCIe as ONE asynchronous copy, instead of 28 strided copies from pageable memory.
 */
static int upload_staged(PdebResident* r, const PdebBatch* in, int64_t lo, int64_t hi, cudaStream_t stream) {
    const int64_t n = hi - lo;
    if (n != r->n || r->host_slab) return fail(PDEB_E_ARG, "staged upload needs a compact image of the shard's size");
    if (n <= 0) return 0;
    if ((in->knot_t != nullptr) != (r->dev.knot_t != nullptr)) return fail(PDEB_E_ARG, "knot_t must be given (or NULL) as at creation");
    /* arrays that already live in page-locked memory need no mirror: every plane segment goes straight to its place in the image */
    const bool direct = host_is_pinned(in->L) && host_is_pinned(in->knot_y) && host_is_pinned(in->tree_rhs) && host_is_pinned(in->kind);
    if (!direct && !r->hstage) {
        cudaError_t ce = cudaHostAlloc((void**)&r->hstage, r->slab_bytes, cudaHostAllocDefault);
        if (ce != cudaSuccess) { r->hstage = nullptr; return fail(PDEB_E_NOMEM, "page-locked staging buffer", ce); }
    }
    const void* hostp[] = {in->kind, in->n_elem, in->L, in->dt, in->T, in->g0, in->x_a, in->x_b, in->rho_in, in->rho_out,
                           in->E_in, in->E_out, in->n_knots, in->knot_t, in->knot_y, in->n_pred, in->pred_xlo, in->pred_xhi,
                           in->pred_m, in->pred_b, in->pred_tlo, in->pred_thi, in->pred_rel, in->pred_top, in->n_tree,
                           in->tree_op, in->tree_lhs, in->tree_rhs};
    const size_t elem[] = {1, 4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 4, 8, 8, 4, 8, 8, 8, 8, 8, 8, 1, 1, 4, 1, 1, 1};
    const int planes[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, in->knot_t ? r->kmax : 0, r->kmax, 1, r->pmax, r->pmax, r->pmax,
                          r->pmax, r->pmax, r->pmax, r->pmax, r->pmax, 1, r->tmax, r->tmax, r->tmax};
    struct Seg { size_t off; const unsigned char* src; size_t bytes; };
    std::vector<Seg> segs;
    segs.reserve(96);
    for (int f = 0; f < 28; ++f) {
        if (!hostp[f]) continue;
        for (int p = 0; p < planes[f]; ++p)
            segs.push_back({r->soff[f] + elem[f] * (size_t)n * p,
                            (const unsigned char*)hostp[f] + elem[f] * ((size_t)p * (size_t)in->stride + (size_t)lo), elem[f] * (size_t)n});
    }
    if (direct) {
        for (const Seg& g : segs) CK(cudaMemcpyAsync(r->slab + g.off, g.src, g.bytes, cudaMemcpyHostToDevice, stream));
        return 0;
    }
    /* the gather into the mirror is a 25 MB memcpy per 65,536 samples: above a few MB it is dealt out to four threads
     * (1,048,576 pageable samples through the chunked call: 6.7e7 -> 8.2e7 evals/s) */
    const int nt = (r->slab_bytes >= ((size_t)4 << 20)) ? 4 : 1;
    auto copy_part = [&](int w) {
        for (size_t q = (size_t)w; q < segs.size(); q += (size_t)nt) memcpy(r->hstage + segs[q].off, segs[q].src, segs[q].bytes);
    };
    if (nt == 1) copy_part(0);
    else {
        std::thread th[3];
        for (int w = 1; w < nt; ++w) th[w - 1] = std::thread(copy_part, w);
        copy_part(0);
        for (int w = 1; w < nt; ++w) th[w - 1].join();
    }
    CK(cudaMemcpyAsync(r->slab, r->hstage, r->slab_bytes, cudaMemcpyHostToDevice, stream));
    return 0;
}

/* a contiguous range of the resident batch seen as a batch of its own (pointer offsets, same plane stride) */
struct RangeView {
    PdebDevBatch dev;      /* as uploaded */
    PdebDevBatch dimg;     /* plan order */
    PdebDevOut out;
    PdebPlanBuffers pb;
    int32_t vsc[2 * PDEB_NUM_VARIANTS];
};

static RangeView view_of(const PdebResident* r, int64_t lo, int64_t hi, int chunk) {
    RangeView v;
    v.dev = r->dev;
    v.dev.n = hi - lo;
    v.dimg = r->dimg;
    v.dimg.n = hi - lo;
#define OFF(f) v.dev.f = r->dev.f + lo; v.dimg.f = r->dimg.f + lo
    OFF(kind); OFF(n_elem); OFF(L); OFF(dt); OFF(T); OFF(g0); OFF(x_a); OFF(x_b); OFF(rho_in); OFF(rho_out); OFF(E_in);
    OFF(E_out); OFF(n_knots); if (r->dev.knot_t) { OFF(knot_t); } OFF(knot_y); OFF(n_pred); OFF(pred_xlo); OFF(pred_xhi); OFF(pred_m); OFF(pred_b);
    OFF(pred_tlo); OFF(pred_thi); OFF(pred_rel); OFF(pred_top); OFF(n_tree); OFF(tree_op); OFF(tree_lhs); OFF(tree_rhs);
#undef OFF
    v.out.robustness = r->out.robustness + lo;
    v.out.clause = r->out.clause + lo;
    v.out.status = r->out.status + lo;
    v.pb.keys = r->pb.keys + lo;
    v.pb.order = r->pb.order + lo;
    v.pb.hist = r->hist[chunk];
    v.pb.vstart_count = r->hist[chunk] + PDEB_PLAN_BINS;
    memset(v.vsc, 0, sizeof v.vsc);
    return v;
}

/*
 * Which kernel variants samples [lo, hi) need and how many of each: a pass over two small HOST arrays.
 * The device-side plan orders the samples; because the variant of a sample depends on (kind, mesh) only,
 * the host knows every variant's place inside that order without reading anything back.
 */
static void host_variant_counts(const PdebBatch* in, int64_t lo, int64_t hi, int32_t* vsc, int flags) {
    int32_t cnt[PDEB_NUM_VARIANTS] = {0};
    int last_kind = -1, last_ne = -1, last_v = -1;      /* the variant is looked up only where (kind, mesh) changes */
    int64_t len = 0;
    for (int64_t s = lo; s < hi; ++s) {
        const int k = in->kind[s], ne = in->n_elem[s];
        if (k != last_kind || ne != last_ne) {
            if (last_v >= 0) cnt[last_v] += (int32_t)len;
            len = 0;
            last_kind = k; last_ne = ne;
            last_v = pdeb_variant_of(k, ne, flags);
        }
        ++len;
    }
    if (last_v >= 0) cnt[last_v] += (int32_t)len;
    int32_t run = 0;
    for (int v = 0; v < PDEB_NUM_VARIANTS; ++v) { vsc[2 * v] = run; vsc[2 * v + 1] = cnt[v]; run += cnt[v]; }
}

/* The FP32 time loop exists for the register family only (default meshes: heat N = 30, wave N = 20).  A batch that holds
 * other meshes is refused rather than silently evaluated in FP64 under an FP32 label. */
static int check_precision(int precision, const int32_t* vsc) {
    if (precision != 32) return 0;
    for (int v = 2; v < PDEB_NUM_VARIANTS; ++v)
        if (vsc[2 * v + 1] > 0)
            return fail(PDEB_E_ARG, "precision 32 is implemented for the default meshes only (heat N=30, wave N=20): this batch holds other meshes");
    return 0;
}

/* order a range by kernel variant and step count (device side, 3 launches, asynchronous) */
static int plan_view(PdebResident* r, const RangeView& v, const int32_t* vsc, cudaStream_t stream, bool keys_done = false,
                     const PdebConventions* conv = nullptr) {
    if (v.dev.n == 0) return 0;
    PdebConventions cv;
    if (conv) cv = *conv; else pdeb_default_conventions(&cv);
    r->plan_steps_rule = cv.steps_rule;
    r->plan_window_rule = cv.window_rule;
    int nl = 0;
    uint32_t live = 0;                       /* variants this range holds (host counts) */
    for (int q = 0; q < PDEB_NUM_VARIANTS; ++q) if (vsc[2 * q + 1] > 0) live |= 1u << q;
    int rc = pdeb_plan(v.dev, cv, v.out, v.pb, live, stream, &nl, keys_done);
    if (rc) return fail(PDEB_E_CUDA, "plan launch", (cudaError_t)rc);
    r->plan_launches = nl;
    r->img_valid = false;                    /* a new order (and usually a new upload): the plan-ordered image is stale */
    r->evals_since_upload = 0;
    return 0;
}

/* the plan-ordered image of the whole batch, formed on demand (samples no variant covers are not part of it) */
static int permute_image(PdebResident* r, bool knots_only, cudaStream_t stream) {
    int64_t ordered = 0;
    for (int q = 0; q < PDEB_NUM_VARIANTS; ++q) ordered += r->vsc_host[2 * q + 1];
    const int rc = pdeb_permute(r->dev, r->dimg, r->pb.order, ordered, knots_only, stream);
    if (rc) return fail(PDEB_E_CUDA, "permute launch", (cudaError_t)rc);
    return 0;
}

/* allocate the device image of `in` (no copy yet) */
static int resident_alloc(const PdebBatch* in, int device, PdebResident** outp, bool force_compact = false) {
    if (!outp) return fail(PDEB_E_ARG, "out is NULL");
    *outp = nullptr;
    int rc = check_batch(in);
    if (rc) return rc;
    if (device < 0 || device >= pdeb_device_count()) return fail(PDEB_E_NODEVICE, "no such CUDA device");
    DeviceGuard g(device);
    if (!g.ok) return fail(PDEB_E_NODEVICE, "cudaSetDevice failed");
    PdebResident* r = new PdebResident();
    r->device = device;
    r->n = in->n;
    r->host_stride = in->stride;
    r->kmax = in->kmax; r->pmax = in->pmax; r->tmax = in->tmax;
    const int64_t n = in->n;

    /* device layout: the host slab verbatim when there is one, else a compact slab with stride n */
    const bool slab = !force_compact && in->slab_base && in->slab_bytes > 0 && in->stride == n;
    r->stride = n;
    PdebDevBatch& d = r->dev;
    d.n = n; d.stride = r->stride; d.kmax = in->kmax; d.pmax = in->pmax; d.tmax = in->tmax;
    { const char* e = getenv("PDEB_CTA"); d.pad = (e && e[0] == '1') ? PDEB_FLAG_CTA : ((e && e[0] == '2') ? PDEB_FLAG_CTA2 : 0); }   /* opt-in kernel family (pdeb_launch.h) */
    const void* hostp[] = {in->kind, in->n_elem, in->L, in->dt, in->T, in->g0, in->x_a, in->x_b, in->rho_in, in->rho_out,
                           in->E_in, in->E_out, in->n_knots, in->knot_t, in->knot_y, in->n_pred, in->pred_xlo, in->pred_xhi,
                           in->pred_m, in->pred_b, in->pred_tlo, in->pred_thi, in->pred_rel, in->pred_top, in->n_tree,
                           in->tree_op, in->tree_lhs, in->tree_rhs};
    const size_t elem[] = {1, 4, 8, 8, 8, 8, 8, 8, 8, 8, 8, 8, 4, 8, 8, 4, 8, 8, 8, 8, 8, 8, 1, 1, 4, 1, 1, 1};
    const int planes[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, in->knot_t ? in->kmax : 0, in->kmax, 1, in->pmax, in->pmax, in->pmax,
                          in->pmax, in->pmax, in->pmax, in->pmax, in->pmax, 1, in->tmax, in->tmax, in->tmax};
    const int NF = 28;
    size_t off[NF];
    size_t total = 0;
    if (slab) {
        for (int f = 0; f < NF; ++f) {
            const unsigned char* p = (const unsigned char*)hostp[f];
            const unsigned char* base = (const unsigned char*)in->slab_base;
            if (!p) { off[f] = 0; continue; }               /* knot_t == NULL */
            if (n > 0 && (p < base || p + elem[f] * (size_t)n * planes[f] > base + in->slab_bytes)) {
                delete r;
                return fail(PDEB_E_ARG, "a batch array lies outside slab_base..slab_base+slab_bytes");
            }
            off[f] = n > 0 ? (size_t)(p - base) : 0;
        }
        total = (size_t)in->slab_bytes;
    } else {
        for (int f = 0; f < NF; ++f) {
            off[f] = total;
            total = align_up(total + elem[f] * (size_t)(n > 0 ? n : 1) * planes[f], 256);
        }
    }
    r->slab_bytes = total;
    for (int f = 0; f < NF; ++f) r->soff[f] = off[f];
    /* the plan-ordered image: always the compact layout */
    size_t ioff[NF];
    size_t itotal = 0;
    for (int f = 0; f < NF; ++f) {
        ioff[f] = itotal;
        itotal = align_up(itotal + elem[f] * (size_t)(n > 0 ? n : 1) * planes[f], 256);
    }
    const size_t out_rob = 0;
    const size_t out_clause = align_up(out_rob + 8 * (size_t)(n > 0 ? n : 1), 256);
    const size_t out_status = align_up(out_clause + 8 * (size_t)(n > 0 ? n : 1) * in->pmax, 256);
    const size_t out_total = align_up(out_status + 4 * (size_t)(n > 0 ? n : 1), 256);
    cudaError_t ce = cudaMalloc((void**)&r->slab, total > 0 ? total : 256);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&r->img, itotal > 0 ? itotal : 256);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&r->outs, out_total);
    const size_t n1 = (size_t)(n > 0 ? n : 1);
    const size_t plan_keys = 0, plan_order = align_up(4 * n1, 256), plan_hist = plan_order + align_up(4 * n1, 256);
    const size_t hist_bytes = align_up(sizeof(int32_t) * (PDEB_PLAN_BINS + 2 * PDEB_VARIANT_SLOTS), 256);
    if (ce == cudaSuccess) ce = cudaMalloc((void**)&r->plan, plan_hist + PDEB_MAX_CHUNKS * hist_bytes);
    for (int i = 0; i < 2 && ce == cudaSuccess; ++i) ce = cudaStreamCreateWithFlags(&r->pipe[i], cudaStreamNonBlocking);
    for (int i = 0; i < 3 && ce == cudaSuccess; ++i) ce = cudaEventCreateWithFlags(&r->pev[i], cudaEventDisableTiming);
    if (ce == cudaSuccess) {
        r->pb.keys = (int32_t*)(r->plan + plan_keys);
        r->pb.order = (int32_t*)(r->plan + plan_order);
        for (int c = 0; c < PDEB_MAX_CHUNKS; ++c) r->hist[c] = (int32_t*)(r->plan + plan_hist + c * hist_bytes);
        r->pb.hist = r->hist[0];
        r->pb.vstart_count = r->hist[0] + PDEB_PLAN_BINS;
    }
    for (int i = 0; i < 4 && ce == cudaSuccess; ++i) ce = cudaEventCreate(&r->ev[i]);
    if (ce != cudaSuccess) {
        pdeb_resident_destroy(r);
        return fail(ce == cudaErrorMemoryAllocation ? PDEB_E_NOMEM : PDEB_E_CUDA, "device allocation", ce);
    }
    unsigned char* s = r->slab;
    d.kind = (const uint8_t*)(s + off[0]);        d.n_elem = (const int32_t*)(s + off[1]);
    d.L = (const double*)(s + off[2]);            d.dt = (const double*)(s + off[3]);
    d.T = (const double*)(s + off[4]);            d.g0 = (const double*)(s + off[5]);
    d.x_a = (const double*)(s + off[6]);          d.x_b = (const double*)(s + off[7]);
    d.rho_in = (const double*)(s + off[8]);       d.rho_out = (const double*)(s + off[9]);
    d.E_in = (const double*)(s + off[10]);        d.E_out = (const double*)(s + off[11]);
    d.n_knots = (const int32_t*)(s + off[12]);    d.knot_t = in->knot_t ? (const double*)(s + off[13]) : nullptr;
    d.knot_y = (const double*)(s + off[14]);      d.n_pred = (const int32_t*)(s + off[15]);
    d.pred_xlo = (const double*)(s + off[16]);    d.pred_xhi = (const double*)(s + off[17]);
    d.pred_m = (const double*)(s + off[18]);      d.pred_b = (const double*)(s + off[19]);
    d.pred_tlo = (const double*)(s + off[20]);    d.pred_thi = (const double*)(s + off[21]);
    d.pred_rel = (const uint8_t*)(s + off[22]);   d.pred_top = (const uint8_t*)(s + off[23]);
    d.n_tree = (const int32_t*)(s + off[24]);     d.tree_op = (const uint8_t*)(s + off[25]);
    d.tree_lhs = (const int8_t*)(s + off[26]);    d.tree_rhs = (const int8_t*)(s + off[27]);
    {
        PdebDevBatch& g = r->dimg;
        unsigned char* q = r->img;
        g = d;
        g.kind = (const uint8_t*)(q + ioff[0]);        g.n_elem = (const int32_t*)(q + ioff[1]);
        g.L = (const double*)(q + ioff[2]);            g.dt = (const double*)(q + ioff[3]);
        g.T = (const double*)(q + ioff[4]);            g.g0 = (const double*)(q + ioff[5]);
        g.x_a = (const double*)(q + ioff[6]);          g.x_b = (const double*)(q + ioff[7]);
        g.rho_in = (const double*)(q + ioff[8]);       g.rho_out = (const double*)(q + ioff[9]);
        g.E_in = (const double*)(q + ioff[10]);        g.E_out = (const double*)(q + ioff[11]);
        g.n_knots = (const int32_t*)(q + ioff[12]);    g.knot_t = in->knot_t ? (const double*)(q + ioff[13]) : nullptr;
        g.knot_y = (const double*)(q + ioff[14]);      g.n_pred = (const int32_t*)(q + ioff[15]);
        g.pred_xlo = (const double*)(q + ioff[16]);    g.pred_xhi = (const double*)(q + ioff[17]);
        g.pred_m = (const double*)(q + ioff[18]);      g.pred_b = (const double*)(q + ioff[19]);
        g.pred_tlo = (const double*)(q + ioff[20]);    g.pred_thi = (const double*)(q + ioff[21]);
        g.pred_rel = (const uint8_t*)(q + ioff[22]);   g.pred_top = (const uint8_t*)(q + ioff[23]);
        g.n_tree = (const int32_t*)(q + ioff[24]);     g.tree_op = (const uint8_t*)(q + ioff[25]);
        g.tree_lhs = (const int8_t*)(q + ioff[26]);    g.tree_rhs = (const int8_t*)(q + ioff[27]);
    }
    r->out.robustness = (double*)(r->outs + out_rob);
    r->out.clause = (double*)(r->outs + out_clause);
    r->out.status = (int32_t*)(r->outs + out_status);
    r->host_slab = slab;
    *outp = r;
    return 0;
}

extern "C" int pdeb_resident_create(const PdebBatch* in, int device, void* stream_, PdebResident** outp) {
    PdebResident* r = nullptr;
    int rc = resident_alloc(in, device, &r);
    if (rc) return rc;
    DeviceGuard g(device);
    rc = upload(r, in, 0, r->n, (cudaStream_t)stream_, false);
    if (rc == 0) { host_variant_counts(in, 0, r->n, r->vsc_host, r->dev.pad); rc = plan_view(r, view_of(r, 0, r->n, 0), r->vsc_host, (cudaStream_t)stream_); }
    if (rc) { pdeb_resident_destroy(r); return rc; }
    *outp = r;
    return 0;
}

extern "C" int pdeb_resident_upload(PdebResident* r, const PdebBatch* in, void* stream_) {
    if (!r) return fail(PDEB_E_ARG, "resident is NULL");
    int rc = check_batch(in);
    if (rc) return rc;
    if (in->n != r->n || in->kmax != r->kmax || in->pmax != r->pmax || in->tmax != r->tmax || in->stride != r->host_stride)
        return fail(PDEB_E_ARG, "batch shape differs from the resident image");
    const bool slab = in->slab_base && in->slab_bytes > 0 && in->stride == in->n;
    if (slab != r->host_slab || (slab && (size_t)in->slab_bytes != r->slab_bytes))
        return fail(PDEB_E_ARG, "batch slab layout differs from the resident image");
    DeviceGuard g(r->device);
    rc = upload(r, in, 0, r->n, (cudaStream_t)stream_, false);
    if (rc == 0) { host_variant_counts(in, 0, r->n, r->vsc_host, r->dev.pad); rc = plan_view(r, view_of(r, 0, r->n, 0), r->vsc_host, (cudaStream_t)stream_); }
    return rc;
}

/* ---- host-side producer of the packed transfer form ------------------------------------------------------------- */
static size_t pk_align(size_t x) { return (x + 63) / 64 * 64; }

extern "C" int64_t pdeb_packed_capacity(int64_t n, int32_t kmax, int32_t pmax, int32_t tmax, int32_t with_knot_t) {
    const size_t n1 = (size_t)(n > 0 ? n : 1);
    size_t t = 0;
    t = pk_align(t + n1);                                   /* kind */
    t = pk_align(t + 2 * n1);                               /* n_elem */
    for (int q = 0; q < 10; ++q) t = pk_align(t + 8 * n1);  /* real scalars */
    t = pk_align(t + 2 * n1);                               /* n_knots */
    t = pk_align(t + n1);                                   /* n_pred */
    t = pk_align(t + n1);                                   /* n_tree */
    for (int q = 0; q < (with_knot_t ? 2 : 1); ++q) t = pk_align(t + 8 * n1 * (size_t)kmax);
    for (int q = 0; q < 6; ++q) t = pk_align(t + 8 * n1 * (size_t)pmax);
    for (int q = 0; q < 2; ++q) t = pk_align(t + n1 * (size_t)pmax);
    for (int q = 0; q < 3; ++q) t = pk_align(t + n1 * (size_t)tmax);
    return (int64_t)t;
}

/* storage tag of `cnt` doubles produced by get(i): one value / float-exact / as is (bit comparison) */
template <typename Get>
static int pk_choose(size_t cnt, Get get) {
    if (cnt == 0) return PDEB_ENC_F64;
    bool same = cnt > 1, f32 = true;
    uint64_t first;
    { const double v0 = get(0); memcpy(&first, &v0, 8); }
    for (size_t i = 0; i < cnt && (same || f32); ++i) {
        const double v = get(i);
        uint64_t b, rb;
        memcpy(&b, &v, 8);
        if (b != first) same = false;
        const double r = (double)(float)v;
        memcpy(&rb, &r, 8);
        if (rb != b) f32 = false;
    }
    return same ? PDEB_ENC_ONE : (f32 ? PDEB_ENC_F32 : PDEB_ENC_F64);
}

extern "C" int pdeb_pack_batch(const PdebBatch* in, void* slab_, int64_t capacity, PdebPackedBatch* out) {
    if (!out || !slab_) return fail(PDEB_E_ARG, "NULL argument");
    int rc = check_batch(in);
    if (rc) return rc;
    if ((uintptr_t)slab_ & 7u) return fail(PDEB_E_ARG, "the slab must be 8-byte aligned");
    const int64_t n = in->n, st = in->stride;
    const int kmax = in->kmax, pmax = in->pmax, tmax = in->tmax;
    if (capacity < pdeb_packed_capacity(n, kmax, pmax, tmax, in->knot_t ? 1 : 0)) return fail(PDEB_E_ARG, "slab smaller than pdeb_packed_capacity");
    unsigned char* slab = (unsigned char*)slab_;
    memset(out, 0, sizeof *out);
    out->n = n; out->kmax = kmax; out->pmax = pmax; out->tmax = tmax;
    out->slab_base = slab;
    size_t off = 0;
    auto take = [&](size_t bytes) { unsigned char* p = slab + off; off = pk_align(off + (bytes ? bytes : 1)); return p; };
    /* counts: range check, clamp to the planes for the list lengths, offsets */
    std::vector<int32_t> ok((size_t)n + 1), op((size_t)n + 1), ot((size_t)n + 1);
    for (int64_t s = 0; s < n; ++s) {
        const int32_t a = in->n_knots[s], b = in->n_pred[s], c = in->n_tree[s], e = in->n_elem[s];
        if (a < 0 || a > 65535 || b < 0 || b > 255 || c < 0 || c > 255 || e < 0 || e > 65535)
            return fail(PDEB_E_ARG, "a count does not fit the packed form's narrow integers");
        ok[s + 1] = ok[s] + (a < kmax ? a : kmax);
        op[s + 1] = op[s] + (b < pmax ? b : pmax);
        ot[s + 1] = ot[s] + (c < tmax ? c : tmax);
    }
    out->total_knots = ok[n]; out->total_pred = op[n]; out->total_tree = ot[n];
    { uint8_t* d = take((size_t)n); memcpy(d, in->kind, (size_t)n); out->kind = d; }
    { uint16_t* d = (uint16_t*)take(2 * (size_t)n); for (int64_t s = 0; s < n; ++s) d[s] = (uint16_t)in->n_elem[s]; out->n_elem = d; }
    /* a real-valued array of `cnt` entries produced by get(i), stored under its tag */
    auto put_real = [&](int slot, size_t cnt, auto get) -> const void* {
        const int e = pk_choose(cnt, get);
        out->enc[slot] = (uint8_t)e;
        if (e == PDEB_ENC_ONE) { double* d = (double*)take(8); d[0] = get(0); return d; }
        if (e == PDEB_ENC_F32) { float* d = (float*)take(4 * cnt); for (size_t i = 0; i < cnt; ++i) d[i] = (float)get(i); return d; }
        double* d = (double*)take(8 * cnt);
        for (size_t i = 0; i < cnt; ++i) d[i] = get(i);
        return d;
    };
    const double* scal[10] = {in->L, in->dt, in->T, in->g0, in->x_a, in->x_b, in->rho_in, in->rho_out, in->E_in, in->E_out};
    const void** scal_out[10] = {&out->L, &out->dt, &out->T, &out->g0, &out->x_a, &out->x_b, &out->rho_in, &out->rho_out, &out->E_in, &out->E_out};
    for (int q = 0; q < 10; ++q) { const double* a = scal[q]; *scal_out[q] = put_real(q, (size_t)n, [a](size_t i) { return a[i]; }); }
    { uint16_t* d = (uint16_t*)take(2 * (size_t)n); for (int64_t s = 0; s < n; ++s) d[s] = (uint16_t)in->n_knots[s]; out->n_knots = d; }
    { uint8_t* d = take((size_t)n); for (int64_t s = 0; s < n; ++s) d[s] = (uint8_t)in->n_pred[s]; out->n_pred = d; }
    { uint8_t* d = take((size_t)n); for (int64_t s = 0; s < n; ++s) d[s] = (uint8_t)in->n_tree[s]; out->n_tree = d; }
    /* lists: entry i of the flat array is (sample, plane) in sample-major order; an index table turns i into its source */
    auto list_index = [&](const std::vector<int32_t>& offs) {
        std::vector<int64_t> src((size_t)offs[n]);
        for (int64_t s = 0; s < n; ++s)
            for (int32_t j = 0, c = offs[s + 1] - offs[s]; j < c; ++j) src[(size_t)offs[s] + j] = (int64_t)j * st + s;
        return src;
    };
    {
        const std::vector<int64_t> src = list_index(ok);
        if (in->knot_t) { const double* a = in->knot_t; out->knot_t = put_real(PDEB_PK_KNOT_T, src.size(), [a, &src](size_t i) { return a[src[i]]; }); }
        const double* a = in->knot_y;
        out->knot_y = put_real(PDEB_PK_KNOT_Y, src.size(), [a, &src](size_t i) { return a[src[i]]; });
    }
    {
        const std::vector<int64_t> src = list_index(op);
        const double* pr[6] = {in->pred_xlo, in->pred_xhi, in->pred_m, in->pred_b, in->pred_tlo, in->pred_thi};
        const void** po[6] = {&out->pred_xlo, &out->pred_xhi, &out->pred_m, &out->pred_b, &out->pred_tlo, &out->pred_thi};
        for (int q = 0; q < 6; ++q) { const double* a = pr[q]; *po[q] = put_real(PDEB_PK_XLO + q, src.size(), [a, &src](size_t i) { return a[src[i]]; }); }
        uint8_t* d = take(src.size()); for (size_t i = 0; i < src.size(); ++i) d[i] = in->pred_rel[src[i]]; out->pred_rel = d;
        d = take(src.size()); for (size_t i = 0; i < src.size(); ++i) d[i] = in->pred_top[src[i]]; out->pred_top = d;
    }
    {
        const std::vector<int64_t> src = list_index(ot);
        uint8_t* d = take(src.size()); for (size_t i = 0; i < src.size(); ++i) d[i] = in->tree_op[src[i]]; out->tree_op = d;
        int8_t* e = (int8_t*)take(src.size()); for (size_t i = 0; i < src.size(); ++i) e[i] = in->tree_lhs[src[i]]; out->tree_lhs = e;
        e = (int8_t*)take(src.size()); for (size_t i = 0; i < src.size(); ++i) e[i] = in->tree_rhs[src[i]]; out->tree_rhs = e;
    }
    out->slab_bytes = (int64_t)off;
    return 0;
}

/* ---- packed transfer form (PdebPackedBatch): one copy, offsets by prefix sums on the device, scatter into the image ---- */
#define PDEB_SCAN_BLOCK 1024

struct PdebPackedDev {
    const uint8_t* kind; const uint16_t* n_elem;
    const void* real[PDEB_PK_COUNT];     /* the real-valued arrays, index PDEB_PK_*, stored as enc[] says */
    const uint16_t* n_knots; const uint8_t *n_pred, *n_tree;
    const uint8_t *pred_rel, *pred_top, *tree_op;
    const int8_t *tree_lhs, *tree_rhs;
    uint8_t enc[24];
};
/* entry i of real-valued array a (the encoding is a kernel parameter: uniform, no divergence) */
__device__ __forceinline__ double pk_ld(const PdebPackedDev& p, int a, int64_t i) {
    const int e = p.enc[a];
    if (e == PDEB_ENC_F64) return ((const double*)p.real[a])[i];
    if (e == PDEB_ENC_F32) return (double)((const float*)p.real[a])[i];
    return ((const double*)p.real[a])[0];
}

/* exclusive prefix sums of the three count arrays inside each block of PDEB_SCAN_BLOCK samples + the block totals */
__global__ void __launch_bounds__(PDEB_SCAN_BLOCK)
pdeb_pack_scan_kernel(int64_t n, const uint16_t* __restrict__ nk, const uint8_t* __restrict__ np, const uint8_t* __restrict__ nt,
                      int kmax, int pmax, int tmax, int32_t* __restrict__ off /* [3][n] */, int32_t* __restrict__ bsum /* [3][blocks] */) {
    __shared__ int32_t wsum[3][32];
    const int64_t i = (int64_t)blockIdx.x * PDEB_SCAN_BLOCK + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t v[3] = {0, 0, 0};
    /* a list holds at most as many entries as the image has planes (larger counts are flagged by the kernels) */
    if (i < n) { v[0] = min((int)nk[i], kmax); v[1] = min((int)np[i], pmax); v[2] = min((int)nt[i], tmax); }
    int32_t inc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int32_t x = v[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        inc[q] = x;
        if (lane == 31) wsum[q][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            int32_t x = wsum[q][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            wsum[q][lane] = x;                                  /* inclusive over the warps */
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int32_t before = (warp > 0 ? wsum[q][warp - 1] : 0) + inc[q] - v[q];
        if (i < n) off[(int64_t)q * n + i] = before;
        if (threadIdx.x == PDEB_SCAN_BLOCK - 1) bsum[(int64_t)q * gridDim.x + blockIdx.x] = wsum[q][31];
    }
}

/* one block: block totals -> exclusive offsets of the blocks (in place), any number of blocks */
__global__ void __launch_bounds__(PDEB_SCAN_BLOCK) pdeb_pack_blocksum_kernel(int32_t* __restrict__ bsum, int blocks) {
    __shared__ int32_t wsum[3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t carry[3] = {0, 0, 0};
    for (int base = 0; base < blocks; base += PDEB_SCAN_BLOCK) {
        const int i = base + threadIdx.x;
        int32_t v[3], inc[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            v[q] = (i < blocks) ? bsum[(int64_t)q * blocks + i] : 0;
            int32_t x = v[q];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            inc[q] = x;
            if (lane == 31) wsum[q][warp] = x;
        }
        __syncthreads();
        if (warp < 3) {
            int32_t x = wsum[warp][lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            wsum[warp][lane] = x;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (i < blocks) bsum[(int64_t)q * blocks + i] = carry[q] + (warp > 0 ? wsum[q][warp - 1] : 0) + inc[q] - v[q];
            carry[q] += wsum[q][31];
        }
        __syncthreads();
    }
}

/* scatter one sample per thread into the plane-major image (unused list slots keep whatever they held: nobody reads them) */
__global__ void pdeb_unpack_kernel(const __grid_constant__ PdebPackedDev p, const __grid_constant__ PdebDevBatch d,
                                   const int32_t* __restrict__ off, const int32_t* __restrict__ bsum, int blocks,
                                   const __grid_constant__ PdebConventions cv, const __grid_constant__ PdebDevOut out,
                                   int32_t* __restrict__ keys, int32_t* __restrict__ hist) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = d.n, st = d.stride;
    if (s >= n) return;
    const int blk = (int)(s / PDEB_SCAN_BLOCK);
    ((uint8_t*)d.kind)[s] = p.kind[s];
    ((int32_t*)d.n_elem)[s] = p.n_elem[s];
    ((double*)d.L)[s] = pk_ld(p, PDEB_PK_L, s);             ((double*)d.dt)[s] = pk_ld(p, PDEB_PK_DT, s);
    ((double*)d.T)[s] = pk_ld(p, PDEB_PK_T, s);             ((double*)d.g0)[s] = pk_ld(p, PDEB_PK_G0, s);
    ((double*)d.x_a)[s] = pk_ld(p, PDEB_PK_XA, s);          ((double*)d.x_b)[s] = pk_ld(p, PDEB_PK_XB, s);
    ((double*)d.rho_in)[s] = pk_ld(p, PDEB_PK_RHO_IN, s);   ((double*)d.rho_out)[s] = pk_ld(p, PDEB_PK_RHO_OUT, s);
    ((double*)d.E_in)[s] = pk_ld(p, PDEB_PK_E_IN, s);       ((double*)d.E_out)[s] = pk_ld(p, PDEB_PK_E_OUT, s);
    int nk = p.n_knots[s], np = p.n_pred[s], nt = p.n_tree[s];
    ((int32_t*)d.n_knots)[s] = nk; ((int32_t*)d.n_pred)[s] = np; ((int32_t*)d.n_tree)[s] = nt;
    /* counts beyond the image's planes are flagged by the kernels (BADPARAM); never write outside the image */
    if (nk > d.kmax) nk = d.kmax;
    if (np > d.pmax) np = d.pmax;
    if (nt > d.tmax) nt = d.tmax;
    const int64_t ok = (int64_t)off[s] + bsum[blk];
    const int64_t op = (int64_t)off[n + s] + bsum[blocks + blk];
    const int64_t ot = (int64_t)off[2 * n + s] + bsum[2 * blocks + blk];
    for (int j = 0; j < nk; ++j) {
        ((double*)d.knot_y)[(int64_t)j * st + s] = pk_ld(p, PDEB_PK_KNOT_Y, ok + j);
        if (d.knot_t) ((double*)d.knot_t)[(int64_t)j * st + s] = pk_ld(p, PDEB_PK_KNOT_T, ok + j);
    }
    for (int j = 0; j < np; ++j) {
        const int64_t q = (int64_t)j * st + s;
        ((double*)d.pred_xlo)[q] = pk_ld(p, PDEB_PK_XLO, op + j);  ((double*)d.pred_xhi)[q] = pk_ld(p, PDEB_PK_XHI, op + j);
        ((double*)d.pred_m)[q] = pk_ld(p, PDEB_PK_M, op + j);      ((double*)d.pred_b)[q] = pk_ld(p, PDEB_PK_B, op + j);
        ((double*)d.pred_tlo)[q] = pk_ld(p, PDEB_PK_TLO, op + j);  ((double*)d.pred_thi)[q] = pk_ld(p, PDEB_PK_THI, op + j);
        ((uint8_t*)d.pred_rel)[q] = p.pred_rel[op + j]; ((uint8_t*)d.pred_top)[q] = p.pred_top[op + j];
    }
    for (int j = 0; j < nt; ++j) {
        const int64_t q = (int64_t)j * st + s;
        ((uint8_t*)d.tree_op)[q] = p.tree_op[ot + j];
        ((int8_t*)d.tree_lhs)[q] = p.tree_lhs[ot + j];
        ((int8_t*)d.tree_rhs)[q] = p.tree_rhs[ot + j];
    }
    /* the plan's sort key and histogram of this sample, while its data are at hand (saves the plan's keys pass) */
    pdeb_plan_count(d, cv, out, s, keys, hist);
}

extern "C" int pdeb_resident_upload_packed(PdebResident* r, const PdebPackedBatch* in, void* stream_) {
    if (!r || !in) return fail(PDEB_E_ARG, "NULL argument");
    if (in->n != r->n || in->kmax != r->kmax || in->pmax != r->pmax || in->tmax != r->tmax)
        return fail(PDEB_E_ARG, "packed batch shape differs from the resident image");
    if ((in->knot_t != nullptr) != (r->dev.knot_t != nullptr)) return fail(PDEB_E_ARG, "knot_t must be given (or NULL) as at creation");
    if (r->n == 0) return 0;
    if (!in->slab_base || in->slab_bytes <= 0) return fail(PDEB_E_ARG, "the packed form needs slab_base / slab_bytes");
    if (in->total_knots < 0 || in->total_pred < 0 || in->total_tree < 0 || in->total_knots > (int64_t)r->n * r->kmax ||
        in->total_pred > (int64_t)r->n * r->pmax || in->total_tree > (int64_t)r->n * r->tmax || in->total_knots > 0x7fffffff)
        return fail(PDEB_E_ARG, "packed list totals out of range");
    const unsigned char* base = (const unsigned char*)in->slab_base;
    const int64_t n = r->n;
    struct Arr { const void* p; size_t bytes; };
    const size_t tk = (size_t)in->total_knots, tp = (size_t)in->total_pred, tt = (size_t)in->total_tree, nn = (size_t)n;
    const void* realp[PDEB_PK_COUNT] = {in->L, in->dt, in->T, in->g0, in->x_a, in->x_b, in->rho_in, in->rho_out, in->E_in,
                                        in->E_out, in->knot_t, in->knot_y, in->pred_xlo, in->pred_xhi, in->pred_m, in->pred_b,
                                        in->pred_tlo, in->pred_thi};
    const size_t realn[PDEB_PK_COUNT] = {nn, nn, nn, nn, nn, nn, nn, nn, nn, nn, tk, tk, tp, tp, tp, tp, tp, tp};
    std::vector<Arr> arrs = {{in->kind, nn}, {in->n_elem, 2 * nn}, {in->n_knots, 2 * nn}, {in->n_pred, nn}, {in->n_tree, nn},
                             {in->pred_rel, tp}, {in->pred_top, tp}, {in->tree_op, tt}, {in->tree_lhs, tt}, {in->tree_rhs, tt}};
    for (int a = 0; a < PDEB_PK_COUNT; ++a) {
        if (in->enc[a] > PDEB_ENC_ONE) return fail(PDEB_E_ARG, "unknown encoding of a packed array");
        if (a == PDEB_PK_KNOT_T && !in->knot_t) continue;
        const size_t bytes = in->enc[a] == PDEB_ENC_F64 ? 8 * realn[a] : in->enc[a] == PDEB_ENC_F32 ? 4 * realn[a] : (realn[a] ? 8 : 0);
        arrs.push_back({realp[a], bytes});
        if (realp[a] && ((uintptr_t)realp[a] & 7u)) return fail(PDEB_E_ARG, "a packed array is not 8-byte aligned");
    }
    for (const Arr& a : arrs) {
        if (!a.p) { if (a.bytes) return fail(PDEB_E_ARG, "a packed array is NULL"); continue; }
        const unsigned char* q = (const unsigned char*)a.p;
        if (q < base || q + a.bytes > base + in->slab_bytes) return fail(PDEB_E_ARG, "a packed array lies outside the slab");
    }
    /* one pass over the small host arrays: the variants of the samples (as for the plain form) and the list totals, which
     * must be the sums of the counts (the device forms the offsets from the counts: a wrong total would read past a list) */
    int32_t cnt[PDEB_NUM_VARIANTS] = {0};
    int64_t sk = 0, sp = 0, stt = 0;
    {
        /* this runs on the submitting host thread once per upload, i.e. inside the end-to-end rate: plain loops the
         * compiler vectorises, and the variant of a sample is looked up only where (kind, mesh) changes */
        const int kmax = r->kmax, pmax = r->pmax, tmax = r->tmax;
        const uint16_t* nk = in->n_knots;
        const uint8_t *npd = in->n_pred, *ntr = in->n_tree;
        for (int64_t s = 0; s < n; ++s) sk += nk[s] < kmax ? nk[s] : kmax;
        for (int64_t s = 0; s < n; ++s) sp += npd[s] < pmax ? npd[s] : pmax;
        for (int64_t s = 0; s < n; ++s) stt += ntr[s] < tmax ? ntr[s] : tmax;
        int last_kind = -1, last_ne = -1, last_v = -1;
        int64_t run = 0;
        for (int64_t s = 0; s < n; ++s) {
            const int k = in->kind[s], ne = in->n_elem[s];
            if (k != last_kind || ne != last_ne) {
                if (last_v >= 0) cnt[last_v] += (int32_t)run;
                run = 0;
                last_kind = k; last_ne = ne;
                last_v = pdeb_variant_of(k, ne, r->dev.pad);
            }
            ++run;
        }
        if (last_v >= 0) cnt[last_v] += (int32_t)run;
    }
    if (sk != in->total_knots || sp != in->total_pred || stt != in->total_tree)
        return fail(PDEB_E_ARG, "packed list totals are not the sums of the (clamped) counts");
    DeviceGuard g(r->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    /* staging: worst-case block (every list full, 64 bytes of alignment slack per array) | offsets [3][n] | block sums */
    const int blocks = (int)((n + PDEB_SCAN_BLOCK - 1) / PDEB_SCAN_BLOCK);
    if (!r->pstage) {
        const size_t worst = nn * (1 + 2 + 80 + 2 + 1 + 1) + 8 * nn * r->kmax * (r->dev.knot_t ? 2 : 1) + 50 * nn * r->pmax +
                             3 * nn * r->tmax + 64 * 32;
        r->pstage_off = align_up(worst, 256);
        r->pstage_cap = r->pstage_off + align_up(12 * nn, 256) + align_up(12 * (size_t)blocks, 256);
        cudaError_t ce = cudaMalloc((void**)&r->pstage, r->pstage_cap);
        if (ce != cudaSuccess) { r->pstage = nullptr; return fail(PDEB_E_NOMEM, "packed staging buffer", ce); }
    }
    if ((size_t)in->slab_bytes > r->pstage_off) return fail(PDEB_E_ARG, "packed slab larger than the worst case of this shape");
    CK(cudaMemcpyAsync(r->pstage, base, (size_t)in->slab_bytes, cudaMemcpyHostToDevice, stream));
    PdebPackedDev p;
#define DEVP(f) p.f = in->f ? (decltype(p.f))(r->pstage + ((const unsigned char*)in->f - base)) : nullptr
    DEVP(kind); DEVP(n_elem); DEVP(n_knots); DEVP(n_pred); DEVP(n_tree); DEVP(pred_rel); DEVP(pred_top);
    DEVP(tree_op); DEVP(tree_lhs); DEVP(tree_rhs);
#undef DEVP
    for (int a = 0; a < PDEB_PK_COUNT; ++a)
        p.real[a] = realp[a] ? (const void*)(r->pstage + ((const unsigned char*)realp[a] - base)) : nullptr;
    memcpy(p.enc, in->enc, sizeof p.enc);
    int32_t* off = (int32_t*)(r->pstage + r->pstage_off);
    int32_t* bsum = (int32_t*)(r->pstage + r->pstage_off + align_up(12 * nn, 256));
    PdebConventions pcv;
    pdeb_default_conventions(&pcv);                      /* the plan orders by the default step / window rules */
    CK(cudaMemsetAsync(r->pb.hist, 0, sizeof(int32_t) * PDEB_PLAN_BINS, stream));
    pdeb_pack_scan_kernel<<<blocks, PDEB_SCAN_BLOCK, 0, stream>>>(n, p.n_knots, p.n_pred, p.n_tree, r->kmax, r->pmax, r->tmax, off, bsum);
    pdeb_pack_blocksum_kernel<<<1, PDEB_SCAN_BLOCK, 0, stream>>>(bsum, blocks);
    pdeb_unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, r->dev, off, bsum, blocks, pcv, r->out, r->pb.keys, r->pb.hist);
    CK(cudaGetLastError());
    int32_t run = 0;
    for (int v = 0; v < PDEB_NUM_VARIANTS; ++v) { r->vsc_host[2 * v] = run; r->vsc_host[2 * v + 1] = cnt[v]; run += cnt[v]; }
    r->plan_stale = false;
    return plan_view(r, view_of(r, 0, r->n, 0), r->vsc_host, stream, true);
}

extern "C" int pdeb_resident_set_knots(PdebResident* r, const double* knot_t, const double* knot_y,
                                       const int32_t* n_knots, void* stream_) {
    if (!r || !knot_y || !n_knots) return fail(PDEB_E_ARG, "NULL argument");
    DeviceGuard g(r->device);
    PdebBatch in{};
    in.n = r->n; in.stride = r->host_stride; in.kmax = r->kmax; in.pmax = r->pmax; in.tmax = r->tmax;
    in.knot_t = knot_t; in.knot_y = knot_y; in.n_knots = n_knots;
    int rc = upload(r, &in, 0, r->n, (cudaStream_t)stream_, true);
    if (rc) return rc;
    /* the knots move into the plan-ordered image under the current order (the order depends on windows and steps, not
     * on the control) */
    if (r->plan_stale) {
        rc = plan_view(r, view_of(r, 0, r->n, 0), r->vsc_host, (cudaStream_t)stream_);
        if (rc) return rc;
        r->plan_stale = false;
        return 0;
    }
    if (r->img_valid) return permute_image(r, true, (cudaStream_t)stream_);
    return 0;
}

extern "C" int pdeb_resident_eval(PdebResident* r, int precision, const PdebConventions* conv, void* stream_,
                                  PdebTiming* timing) {
    if (!r) return fail(PDEB_E_ARG, "resident is NULL");
    if (precision != 64 && precision != 32) return fail(PDEB_E_ARG, "precision must be 64 or 32");
    DeviceGuard g(r->device);
    if (!g.ok) return fail(PDEB_E_NODEVICE, "cudaSetDevice failed");
    cudaStream_t stream = (cudaStream_t)stream_;
    PdebConventions cv;
    if (conv) cv = *conv; else pdeb_default_conventions(&cv);
    int launches = 0;
    { int prc = check_precision(precision, r->vsc_host); if (prc) return prc; }
    /* the order was formed at upload under the default step / window rules: a caller that evaluates under others gets
     * the batch re-ordered under ITS rules (results do not depend on the order, warp homogeneity does) */
    if (r->plan_stale || cv.steps_rule != r->plan_steps_rule || cv.window_rule != r->plan_window_rule) {
        int prc = plan_view(r, view_of(r, 0, r->n, 0), r->vsc_host, stream, false, &cv);
        if (prc) return prc;
        r->plan_stale = false;
    }
    if (!r->img_valid && r->evals_since_upload >= 1) {
        int prc = permute_image(r, false, stream);      /* evaluated again: from now on the kernels read coalesced */
        if (prc) return prc;
        r->img_valid = true;
    }
    ++r->evals_since_upload;
    if (timing) CK(cudaEventRecord(r->ev[0], stream));
    int rc = pdeb_launch_all(r->img_valid ? r->dimg : r->dev, cv, r->out, r->pb, r->vsc_host, precision, r->img_valid ? 0 : 1, stream, &launches);
    if (rc) return fail(PDEB_E_CUDA, "kernel launch", (cudaError_t)rc);
    if (timing) {
        CK(cudaEventRecord(r->ev[1], stream));
        CK(cudaEventSynchronize(r->ev[1]));
        memset(timing, 0, sizeof *timing);
        CK(cudaEventElapsedTime(&timing->kernel_ms, r->ev[0], r->ev[1]));
        timing->total_ms = timing->kernel_ms;
        timing->launches = launches;
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int pdeb_resident_fetch(PdebResident* r, double* robustness_out, double* clause_out, int32_t* status_out,
                                   void* stream_) {
    if (!r || !robustness_out) return fail(PDEB_E_ARG, "NULL argument");
    DeviceGuard g(r->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t n = r->n;
    if (n > 0) {
        CK(cudaMemcpyAsync(robustness_out, r->out.robustness, 8 * (size_t)n, cudaMemcpyDeviceToHost, stream));
        if (status_out) CK(cudaMemcpyAsync(status_out, r->out.status, 4 * (size_t)n, cudaMemcpyDeviceToHost, stream));
        if (clause_out)
            CK(cudaMemcpy2DAsync(clause_out, 8 * (size_t)r->host_stride, r->out.clause, 8 * (size_t)r->stride, 8 * (size_t)n,
                                 r->pmax, cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int pdeb_resident_fetch_async(PdebResident* r, double* robustness_out, double* clause_out, int32_t* status_out,
                                         void* stream_) {
    if (!r || !robustness_out) return fail(PDEB_E_ARG, "NULL argument");
    DeviceGuard g(r->device);
    return fetch_range_fwd(r, robustness_out, clause_out, status_out, (cudaStream_t)stream_);
}

extern "C" void* pdeb_stream_create(int device) {
    DeviceGuard g(device);
    cudaStream_t st = nullptr;
    if (!g.ok || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (void*)st;
}
extern "C" void pdeb_stream_destroy(void* stream) { if (stream) cudaStreamDestroy((cudaStream_t)stream); }
extern "C" int pdeb_stream_sync(void* stream) {
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    CK(cudaGetLastError());
    return 0;
}

extern "C" int pdeb_resident_plan_order(PdebResident* r, int32_t* order_out, void* stream_) {
    if (!r || !order_out) return fail(PDEB_E_ARG, "NULL argument");
    DeviceGuard g(r->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (r->n > 0) CK(cudaMemcpyAsync(order_out, r->pb.order, 4 * (size_t)r->n, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return 0;
}

extern "C" int pdeb_resident_device_ptrs(PdebResident* r, void** robustness_dev, void** clause_dev, void** status_dev) {
    if (!r) return fail(PDEB_E_ARG, "resident is NULL");
    if (robustness_dev) *robustness_dev = r->out.robustness;
    if (clause_dev) *clause_dev = r->out.clause;
    if (status_dev) *status_dev = r->out.status;
    return 0;
}

/* ---- upload + evaluate + fetch, pipelined ---------------------------------- */

static int fetch_range(PdebResident* r, int64_t lo, int64_t hi, double* robustness_out, double* clause_out,
                       int32_t* status_out, cudaStream_t stream) {
    const int64_t n = hi - lo;
    if (n <= 0) return 0;
    CK(cudaMemcpyAsync(robustness_out + lo, r->out.robustness + lo, 8 * (size_t)n, cudaMemcpyDeviceToHost, stream));
    if (status_out) CK(cudaMemcpyAsync(status_out + lo, r->out.status + lo, 4 * (size_t)n, cudaMemcpyDeviceToHost, stream));
    if (clause_out)
        CK(cudaMemcpy2DAsync(clause_out + lo, 8 * (size_t)r->host_stride, r->out.clause + lo, 8 * (size_t)r->stride,
                             8 * (size_t)n, r->pmax, cudaMemcpyDeviceToHost, stream));
    return 0;
}

static int fetch_range_fwd(PdebResident* r, double* robustness_out, double* clause_out, int32_t* status_out, cudaStream_t stream) {
    return fetch_range(r, 0, r->n, robustness_out, clause_out, status_out, stream);
}

/*
 * One pass over a host batch through an existing resident image: the batch is cut into up to
 * PDEB_MAX_CHUNKS contiguous chunks that alternate between two internal streams, so the H2D copy of one
 * chunk overlaps the kernels of the previous one and the D2H of the one before.  Synchronises before returning.
 */
extern "C" int pdeb_resident_run(PdebResident* r, const PdebBatch* in, int precision, const PdebConventions* conv,
                                 double* robustness_out, double* clause_out, int32_t* status_out, void* stream_,
                                 PdebTiming* timing) {
    if (!r || !robustness_out) return fail(PDEB_E_ARG, "NULL argument");
    if (precision != 64 && precision != 32) return fail(PDEB_E_ARG, "precision must be 64 or 32");
    int rc = check_batch(in);
    if (rc) return rc;
    if (in->n != r->n || in->kmax != r->kmax || in->pmax != r->pmax || in->tmax != r->tmax || in->stride != r->host_stride)
        return fail(PDEB_E_ARG, "batch shape differs from the resident image");
    DeviceGuard g(r->device);
    if (!g.ok) return fail(PDEB_E_NODEVICE, "cudaSetDevice failed");
    cudaStream_t user = (cudaStream_t)stream_;
    PdebConventions cv;
    if (conv) cv = *conv; else pdeb_default_conventions(&cv);
    const int64_t n = r->n;
    /* Chunking inside one call pays only when the per-field copies are not the bottleneck: a batch that lives in
       one host slab moves fastest as ONE copy (measured on B200: 28 strided copies per chunk cost more CPU time than
       the overlap wins), so the default is a single chunk; PDEB_CHUNKS=2..4 turns the in-call pipeline on.
       Overlap across batches is what native.BlockPipeline is for. */
    static const int env_chunks = [] { const char* e = getenv("PDEB_CHUNKS"); return e ? atoi(e) : 1; }();
    int nch = env_chunks;
    if (nch > (int)(n / PDEB_PIPELINE_MIN)) nch = (int)(n / PDEB_PIPELINE_MIN);
    if (nch < 1) nch = 1;
    if (nch > PDEB_MAX_CHUNKS) nch = PDEB_MAX_CHUNKS;
    int launches = 0;
    {
        int32_t vall[2 * PDEB_NUM_VARIANTS];
        host_variant_counts(in, 0, n, vall, r->dev.pad);
        rc = check_precision(precision, vall);
        if (rc) return rc;
    }
    if (timing) CK(cudaEventRecord(r->ev[0], user));
    CK(cudaEventRecord(r->pev[0], user));
    for (int i = 0; i < 2; ++i) CK(cudaStreamWaitEvent(r->pipe[i], r->pev[0], 0));
    for (int c = 0; c < nch; ++c) {
        const int64_t lo = n * c / nch, hi = n * (c + 1) / nch;
        cudaStream_t st = r->pipe[c & 1];
        rc = upload(r, in, lo, hi, st, false);
        if (rc) return rc;
        RangeView v = view_of(r, lo, hi, c);
        host_variant_counts(in, lo, hi, v.vsc, r->dev.pad);
        rc = plan_view(r, v, v.vsc, st, false, &cv);
        if (rc) return rc;
        int nl = 0;
        const int lrc = pdeb_launch_all(v.dev, cv, v.out, v.pb, v.vsc, precision, 1, st, &nl);
        if (lrc) return fail(PDEB_E_CUDA, "kernel launch", (cudaError_t)lrc);
        launches += nl + r->plan_launches;
        rc = fetch_range(r, lo, hi, robustness_out, clause_out, status_out, st);
        if (rc) return rc;
    }
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventRecord(r->pev[1 + i], r->pipe[i]));
        CK(cudaStreamWaitEvent(user, r->pev[1 + i], 0));
    }
    if (timing) CK(cudaEventRecord(r->ev[3], user));
    CK(cudaStreamSynchronize(user));
    CK(cudaGetLastError());
    /* the whole-batch order and counts are stale now (each chunk was planned on its own): rebuild lazily */
    host_variant_counts(in, 0, n, r->vsc_host, r->dev.pad);
    r->plan_stale = nch > 1;
    if (timing) {
        memset(timing, 0, sizeof *timing);
        CK(cudaEventElapsedTime(&timing->total_ms, r->ev[0], r->ev[3]));
        timing->launches = launches;
    }
    return 0;
}

/* ---- one-shot ------------------------------------------------------------ */

/*
 * Resident images of recent one-shot calls, kept per device and keyed by shape, so that a caller that evaluates batch
 * after batch of one shape (the per-sample helpers of the reference's call surface: eval_robustness_helper evaluates
 * ONE program per call) does not pay three cudaMalloc, two streams and seven events every time.  Small images only
 * (PDEB_CACHE_MAX_BYTES of device memory each, PDEB_CACHE_ENTRIES per device, least recently used out); an image in
 * use by another thread is never shared: that caller gets a private one.  pdeb_cache_clear() frees them.
 */
#define PDEB_CACHE_ENTRIES 8
#define PDEB_CACHE_MAX_BYTES ((size_t)256 << 20)
#define PDEB_API_MAX_DEVICES 64
struct CacheKey {
    int64_t n;
    int kmax, pmax, tmax, has_kt, slab;
    size_t slab_bytes, sig0, sig1;
    bool operator==(const CacheKey& o) const {
        return n == o.n && kmax == o.kmax && pmax == o.pmax && tmax == o.tmax && has_kt == o.has_kt && slab == o.slab &&
               slab_bytes == o.slab_bytes && sig0 == o.sig0 && sig1 == o.sig1;
    }
};
struct CacheEntry { CacheKey key; PdebResident* r; bool busy; uint64_t stamp; };
static std::mutex g_cache_mu;
static std::vector<CacheEntry> g_cache[PDEB_API_MAX_DEVICES];
static uint64_t g_cache_stamp = 0;

static int cache_acquire(const PdebBatch* view, const CacheKey& key, int device, bool compact, PdebResident** outp, bool* cached) {
    *cached = false;
    if (device >= 0 && device < PDEB_API_MAX_DEVICES) {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        for (CacheEntry& e : g_cache[device])
            if (!e.busy && e.key == key) { e.busy = true; e.stamp = ++g_cache_stamp; *outp = e.r; *cached = true; return 0; }
    }
    PdebResident* r = nullptr;
    int rc = resident_alloc(view, device, &r, compact);
    if (rc) return rc;
    *outp = r;
    const size_t bytes = 2 * r->slab_bytes + (size_t)r->n * (20 + 8 * (size_t)r->pmax);
    if (device >= 0 && device < PDEB_API_MAX_DEVICES && bytes <= PDEB_CACHE_MAX_BYTES) {
        PdebResident* evict = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_cache_mu);
            std::vector<CacheEntry>& v = g_cache[device];
            if (v.size() >= PDEB_CACHE_ENTRIES) {
                int lru = -1;
                for (int i = 0; i < (int)v.size(); ++i) if (!v[i].busy && (lru < 0 || v[i].stamp < v[lru].stamp)) lru = i;
                if (lru >= 0) { evict = v[lru].r; v.erase(v.begin() + lru); }
            }
            if (v.size() < PDEB_CACHE_ENTRIES) { v.push_back({key, r, true, ++g_cache_stamp}); *cached = true; }
        }
        if (evict) pdeb_resident_destroy(evict);
    }
    return 0;
}

static void cache_release(PdebResident* r, bool cached) {
    if (!cached) { pdeb_resident_destroy(r); return; }
    std::lock_guard<std::mutex> lock(g_cache_mu);
    for (CacheEntry& e : g_cache[r->device]) if (e.r == r) { e.busy = false; return; }
}

extern "C" void pdeb_cache_clear(void) {
    std::vector<PdebResident*> dead;
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        for (auto& v : g_cache) {
            for (size_t i = 0; i < v.size();) {
                if (!v[i].busy) { dead.push_back(v[i].r); v.erase(v.begin() + i); } else ++i;
            }
        }
    }
    for (PdebResident* r : dead) pdeb_resident_destroy(r);
}

/*
 * Samples [lo, hi) of a host batch on one device: H2D (one copy), plan, kernels, D2H into the caller's arrays at the
 * samples' own indices; synchronises.  A batch that is one page-locked slab of exactly this range goes across as it
 * lies; anything else (a shard, pageable memory, a strided view) through the image's page-locked mirror (upload_staged).
 */
static int run_range(const PdebBatch* in, int64_t lo, int64_t hi, int device, int precision, const PdebConventions* conv,
                     double* robustness_out, double* clause_out, int32_t* status_out, cudaStream_t user, PdebTiming* timing) {
    const int64_t n = hi - lo;
    if (timing) memset(timing, 0, sizeof *timing);
    if (n <= 0) return 0;
    const bool whole = lo == 0 && hi == in->n;
    const bool slab = whole && in->slab_base && in->slab_bytes > 0 && in->stride == in->n && host_is_pinned(in->slab_base);
    PdebBatch view = *in;                       /* shape of the image: n samples */
    view.n = n;
    if (!slab) { view.slab_base = nullptr; view.slab_bytes = 0; }
    CacheKey key{n, in->kmax, in->pmax, in->tmax, in->knot_t ? 1 : 0, slab ? 1 : 0, 0, 0, 0};
    if (slab) {
        key.slab_bytes = (size_t)in->slab_bytes;
        key.sig0 = (size_t)((const unsigned char*)in->L - (const unsigned char*)in->slab_base);
        key.sig1 = (size_t)((const unsigned char*)in->tree_rhs - (const unsigned char*)in->slab_base);
    }
    PdebResident* r = nullptr;
    bool cached = false;
    int rc = cache_acquire(&view, key, device, !slab, &r, &cached);
    if (rc) return rc;
    r->host_stride = in->stride;
    if (slab) {
        rc = pdeb_resident_run(r, in, precision, conv, robustness_out, clause_out, status_out, user, timing);
        cache_release(r, cached);
        return rc;
    }
    DeviceGuard g(device);
    PdebConventions cv;
    if (conv) cv = *conv; else pdeb_default_conventions(&cv);
    auto done = [&](int code) { cache_release(r, cached); return code; };
    if (timing) { cudaError_t ce = cudaEventRecord(r->ev[0], user); if (ce != cudaSuccess) return done(fail(PDEB_E_CUDA, "event record", ce)); }
    rc = upload_staged(r, in, lo, hi, user);
    if (rc) return done(rc);
    {
        /* variants of the shard: the counting pass reads the caller's arrays at their own indices */
        host_variant_counts(in, lo, hi, r->vsc_host, r->dev.pad);
        rc = check_precision(precision, r->vsc_host);
        if (rc) return done(rc);
        rc = plan_view(r, view_of(r, 0, n, 0), r->vsc_host, user, false, &cv);
        if (rc) return done(rc);
        r->plan_stale = false;
    }
    int launches = 0;
    const int lrc = pdeb_launch_all(r->dev, cv, r->out, r->pb, r->vsc_host, precision, 1, user, &launches);
    if (lrc) return done(fail(PDEB_E_CUDA, "kernel launch", (cudaError_t)lrc));
    {
        cudaError_t ce = cudaMemcpyAsync(robustness_out + lo, r->out.robustness, 8 * (size_t)n, cudaMemcpyDeviceToHost, user);
        if (ce == cudaSuccess && status_out) ce = cudaMemcpyAsync(status_out + lo, r->out.status, 4 * (size_t)n, cudaMemcpyDeviceToHost, user);
        if (ce == cudaSuccess && clause_out)
            ce = cudaMemcpy2DAsync(clause_out + lo, 8 * (size_t)in->stride, r->out.clause, 8 * (size_t)r->stride, 8 * (size_t)n,
                                   r->pmax, cudaMemcpyDeviceToHost, user);
        if (ce == cudaSuccess && timing) ce = cudaEventRecord(r->ev[3], user);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(user);
        if (ce == cudaSuccess) ce = cudaGetLastError();
        if (ce != cudaSuccess) return done(fail(PDEB_E_CUDA, "download", ce));
    }
    if (timing) {
        cudaEventElapsedTime(&timing->total_ms, r->ev[0], r->ev[3]);
        timing->launches = launches + r->plan_launches;
    }
    return done(0);
}

/*
 * A LARGE one-shot range: chunks of PDEB_ONESHOT_CHUNK samples dealt out to a few host threads, each with a stream and a
 * cached image of the chunk's shape.  One thread copies its chunk into its page-locked mirror while the chunk of another
 * crosses PCIe and a third computes, so the call runs at the rate of its slowest stage instead of the sum of all three
 * (measured, 1,048,576 heat samples from pageable memory: 4.1e6 evals/s as ONE chunk -- a 400 MB image and mirror allocated and
 * freed per call, outside the cache's size limit -- against 8.2e7 in chunks on 6 threads; 262,144 samples: 1.8e7 -> 4.3e7; from a
 * page-locked slab 6.0e7 -> 6.6e7; the 65,536-sample call is unchanged: tools/oneshot_rate.py).
 * Samples are independent, so the results are those of the single-range call.
 */
#define PDEB_ONESHOT_CHUNK 65536
static int run_chunked(const PdebBatch* in, int64_t lo, int64_t hi, int device, int precision, const PdebConventions* conv,
                       double* robustness_out, double* clause_out, int32_t* status_out, cudaStream_t user, PdebTiming* timing) {
    if (timing) memset(timing, 0, sizeof *timing);
    {
        /* what the caller queued on its stream comes first; the chunks then run on streams of their own */
        DeviceGuard g(device);
        cudaError_t ce = cudaStreamSynchronize(user);
        if (ce != cudaSuccess) return fail(PDEB_E_CUDA, "stream synchronise", ce);
    }
    const int64_t n = hi - lo, nchunks = (n + PDEB_ONESHOT_CHUNK - 1) / PDEB_ONESHOT_CHUNK;
    static const int want = [] { const char* e = getenv("PDEB_ONESHOT_THREADS"); const int v = e ? atoi(e) : 0; return v > 0 ? (v > 8 ? 8 : v) : 6; }();
    const int nt = (int)(nchunks < want ? nchunks : want);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    std::mutex err_mutex;
    std::string err_text;
    int err_code = 0;
    std::vector<int> launches((size_t)nt, 0);
    const auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int w) {
        cudaStream_t st = nullptr;
        int rc = 0;
        if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(PDEB_E_CUDA, "pdeb_eval_batch: stream for a chunk");
        }
        while (rc == 0 && !failed.load(std::memory_order_relaxed)) {
            const int64_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            const int64_t a = lo + c * PDEB_ONESHOT_CHUNK, b = (a + PDEB_ONESHOT_CHUNK < hi) ? a + PDEB_ONESHOT_CHUNK : hi;
            PdebTiming tm;
            rc = run_range(in, a, b, device, precision, conv, robustness_out, clause_out, status_out, st, timing ? &tm : nullptr);
            if (rc == 0 && timing) launches[(size_t)w] += tm.launches;
        }
        if (rc != 0) {
            std::lock_guard<std::mutex> lock(err_mutex);
            if (err_code == 0) { err_code = rc; err_text = g_err; }
            failed.store(1);
        }
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    };
    std::vector<std::thread> pool;
    for (int w = 1; w < nt; ++w) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
    if (err_code != 0) { g_err = err_text; return err_code; }
    if (timing) {
        timing->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();   /* wall clock of the call */
        for (int w = 0; w < nt; ++w) timing->launches += launches[(size_t)w];
    }
    return 0;
}

extern "C" int pdeb_eval_batch(const PdebBatch* in, int device, int precision, const PdebConventions* conv,
                               double* robustness_out, double* clause_out, int32_t* status_out, void* stream_,
                               PdebTiming* timing) {
    if (!robustness_out) return fail(PDEB_E_ARG, "robustness_out is NULL");
    if (precision != 64 && precision != 32) return fail(PDEB_E_ARG, "precision must be 64 or 32");
    int rc = check_batch(in);
    if (rc) return rc;
    if (device < 0 || device >= pdeb_device_count()) return fail(PDEB_E_NODEVICE, "no such CUDA device");
    if (in->n >= 2 * PDEB_ONESHOT_CHUNK)
        return run_chunked(in, 0, in->n, device, precision, conv, robustness_out, clause_out, status_out, (cudaStream_t)stream_, timing);
    return run_range(in, 0, in->n, device, precision, conv, robustness_out, clause_out, status_out, (cudaStream_t)stream_, timing);
}

/* ---- several devices of one box: contiguous shards balanced by cost, no inter-device traffic ---- */

/*
 * Cut points of `ndev` contiguous shards carrying equal COST ("for mixed batches balance by estimated
 * cost S n rather than by count"): cost of a sample = time steps x mesh rows (x 1.5 for waves: three state vectors).
 * bounds[r] .. bounds[r+1] is shard r.  pdeb_shard_bounds exports the rule (the Python mirror is sharding.cost_balanced_bounds).
 */
extern "C" int pdeb_shard_bounds(const PdebBatch* in, int ndev, const PdebConventions* conv, int64_t* bounds) {
    int rc = check_batch(in);
    if (rc) return rc;
    if (ndev < 1 || !bounds) return fail(PDEB_E_ARG, "bad ndev / bounds");
    PdebConventions cv;
    if (conv) cv = *conv; else pdeb_default_conventions(&cv);
    const int64_t n = in->n;
    std::vector<double> acc((size_t)n + 1, 0.0);
    for (int64_t s = 0; s < n; ++s) {
        const double T = in->T[s], dt = in->dt[s];
        double c = 1.0;
        if (T > 0 && dt > 0 && std::isfinite(T) && std::isfinite(dt)) {
            const double steps = (double)pdeb_step_count(T, dt, cv.steps_rule);
            const int ne = in->n_elem[s] > 0 ? in->n_elem[s] : 1;
            c = steps * (double)ne * (in->kind[s] == PDEB_KIND_WAVE ? 1.5 : 1.0);
        }
        acc[(size_t)s + 1] = acc[(size_t)s] + c;
    }
    const double total = acc[(size_t)n];
    bounds[0] = 0;
    int64_t cut = 0;
    for (int r = 1; r < ndev; ++r) {
        const double want = total * (double)r / (double)ndev;
        while (cut < n && acc[(size_t)cut + 1] < want) ++cut;   /* the sample whose inclusion reaches the target opens the next shard */
        bounds[r] = cut;
    }
    bounds[ndev] = n;
    return 0;
}

extern "C" int pdeb_eval_batch_multi(const PdebBatch* in, int ndev, const int* devices, int precision,
                                     const PdebConventions* conv, double* robustness_out, double* clause_out,
                                     int32_t* status_out, PdebTiming* timing_per_device) {
    int rc = check_batch(in);
    if (rc) return rc;
    {
        /* devices == NULL: shard r on device r; an explicit list may name a device more than once (its shards then run
         * concurrently on that device, each on an image of its own) */
        const int have = pdeb_device_count();
        if (ndev < 1 || ndev > 1024) return fail(PDEB_E_ARG, "bad ndev");
        if (!devices && ndev > have) return fail(PDEB_E_NODEVICE, "ndev exceeds the visible devices");
        for (int r = 0; devices && r < ndev; ++r)
            if (devices[r] < 0 || devices[r] >= have) return fail(PDEB_E_NODEVICE, "no such CUDA device in the device list");
    }
    if (!robustness_out) return fail(PDEB_E_ARG, "robustness_out is NULL");
    if (precision != 64 && precision != 32) return fail(PDEB_E_ARG, "precision must be 64 or 32");
    std::vector<int64_t> bounds((size_t)ndev + 1);
    rc = pdeb_shard_bounds(in, ndev, conv, bounds.data());
    if (rc) return rc;
    std::vector<int> rcs(ndev, 0);
    std::vector<std::string> errs(ndev);
    std::vector<std::thread> th;
    for (int r = 0; r < ndev; ++r) {
        th.emplace_back([&, r]() {
            const int64_t lo = bounds[(size_t)r], hi = bounds[(size_t)r + 1];
            if (timing_per_device) memset(&timing_per_device[r], 0, sizeof(PdebTiming));
            if (hi <= lo) return;
            const int dev = devices ? devices[r] : r;
            cudaStream_t stream = nullptr;
            if (cudaSetDevice(dev) != cudaSuccess || cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) {
                cudaGetLastError();
                rcs[r] = PDEB_E_NODEVICE; errs[r] = "cannot use the device";
                return;
            }
            /* the shard goes through its device's page-locked mirror: one H2D copy, results straight into the caller's arrays */
            rcs[r] = (hi - lo >= 2 * PDEB_ONESHOT_CHUNK)
                         ? run_chunked(in, lo, hi, dev, precision, conv, robustness_out, clause_out, status_out, stream,
                                       timing_per_device ? &timing_per_device[r] : nullptr)
                         : run_range(in, lo, hi, dev, precision, conv, robustness_out, clause_out, status_out, stream,
                                     timing_per_device ? &timing_per_device[r] : nullptr);
            if (rcs[r]) errs[r] = g_err;
            cudaStreamDestroy(stream);
        });
    }
    for (auto& t : th) t.join();
    for (int r = 0; r < ndev; ++r)
        if (rcs[r]) { g_err = "device shard " + std::to_string(r) + ": " + errs[r]; return rcs[r]; }
    return 0;
}
