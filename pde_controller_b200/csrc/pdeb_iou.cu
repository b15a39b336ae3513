/*
 * pdeb_iou.cu — the boolean grids of the reference's IoU metric, batched on the GPU (optional part).
 *
 * The reference compares two sSTL formulas by sampling them (test/PDEcontrol/evaluation/eval/eval_script.py:233-281,
 * eval_IOU): for every clause, every linear test profile u(t) = m t + b of a 50 x 50 grid, every time of a 51-step grid
 * over the common time domain and every position of a 51-step grid over the common position domain, it evaluates
 *     time in [t0, t1] and pos in [p0, p1]  ->  (u - expr(pos)) OP threshold            (check_conditions, :437-487)
 * with one Python eval() per cell inside the clause's window (up to ~7 M per clause: minutes), folds the positions into "holds for all positions inside"
 * / "fails somewhere" / "vacuous", runs the G or F bit-array state machine over the times of a profile (:159-219) and
 * finally combines the clauses' bit arrays and counts intersection over union.
 *
 * Here one THREAD owns one (clause, profile) pair: it walks the times in the order the reference does (forwards for G,
 * backwards for F), folds the positions, and writes the profile's final bits.  Every arithmetic operation is the IEEE
 * double operation Python performs on the same operands in the same order (__dmul_rn / __dadd_rn / __dsub_rn / __ddiv_rn:
 * no FMA contraction), and the grids themselves are formed on the host by the same numpy calls, so the bits are EXACTLY
 * the reference's (tests/golden/iou_v0.json holds outputs of the reference's own eval_IOU).
 * Integer / boolean work: no tensor cores, nothing to fuse; the kernel is launch-and-forget next to the simulation path.
 */
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "../../include/pdeb200.h"
#include "pdeb_launch.h"

#define PDEB_IOU_THREADS 128

__device__ __forceinline__ bool iou_compare(int op, double a, double b) {
    switch (op) {
    case PDEB_IOU_LE: return a <= b;
    case PDEB_IOU_LT: return a < b;
    case PDEB_IOU_GE: return a >= b;
    case PDEB_IOU_GT: return a > b;
    case PDEB_IOU_EQ: return a == b;
    default: return a != b;
    }
}

__global__ void __launch_bounds__(PDEB_IOU_THREADS)
pdeb_iou_kernel(const PdebIouClause* __restrict__ clauses, int64_t n_clauses, const double* __restrict__ tgrid,
                const double* __restrict__ xgrid, const int32_t* __restrict__ ngrid, int tmax, int xmax,
                const double* __restrict__ m_vals, int nm, const double* __restrict__ b_vals, int nb,
                uint8_t* __restrict__ bits) {
    const int n_prof = nm * nb;
    const int64_t gid = (int64_t)blockIdx.x * PDEB_IOU_THREADS + threadIdx.x;
    if (gid >= n_clauses * n_prof) return;
    const int64_t ci = gid / n_prof;
    const int prof = (int)(gid % n_prof);
    const PdebIouClause c = clauses[ci];
    const int dir = c.backwards ? 1 : 0;
    const double* tg = tgrid + ((int64_t)c.item * 2 + dir) * tmax;
    const double* xg = xgrid + ((int64_t)c.item * 2 + dir) * xmax;
    const int nt = ngrid[c.item * 4 + dir], nx = ngrid[c.item * 4 + 2 + dir];
    const double m = m_vals[prof / nb], b = b_vals[prof % nb];      /* itertools.product(m, b): m outer, b inner */
    /* per time (in the reference's iteration order): does the clause hold for all positions inside / fail / is it vacuous */
    uint64_t vac = 0, tru = 0, fal = 0;
    for (int ti = 0; ti < nt; ++ti) {
        const double time = tg[ti];
        const double u = __dadd_rn(__dmul_rn(m, time), b);
        int flag = -1;                                              /* None */
        if (time >= c.t0 && time <= c.t1) {
            for (int xi = 0; xi < nx; ++xi) {
                const double pos = xg[xi];
                if (pos >= c.p0 && pos <= c.p1) {
                    const double xs = c.neg ? -pos : pos;
                    const double a = (c.kind == PDEB_IOU_DIV) ? __ddiv_rn(xs, c.coef) : __dmul_rn(c.coef, xs);
                    const double expr = __dadd_rn(a, c.c);
                    const bool ok = iou_compare(c.op, __dsub_rn(u, expr), c.thr);
                    if (ok) { if (flag < 0) flag = 1; }
                    else flag = 0;
                }
            }
        }
        const uint64_t bit = 1ull << ti;
        if (flag < 0) vac |= bit; else if (flag) tru |= bit; else fal |= bit;
    }
    uint8_t* out = bits + ((int64_t)ci * n_prof + prof) * tmax;
    if (!c.backwards) {
        /* G_array (eval_script.py:159-187), times ascending: a failure zeroes the profile's segment up to itself and turns
         * every later non-vacuous entry into 0; vacuous entries behind the last failure stay 1 */
        const int last_fail = fal ? 63 - __clzll((long long)fal) : -1;
        for (int ti = 0; ti < nt; ++ti) {
            uint8_t v;
            if (ti <= last_fail) v = 0;
            else if ((vac >> ti) & 1ull) v = 1;
            else v = (last_fail >= 0) ? 0 : 1;
            out[ti] = v;
        }
    } else {
        /* F_array (:190-210), times DESCENDING, inserted at the front: an entry is 1 if it is vacuous, holds, or a later
         * time (already visited) held; the array is in ascending time order: position nt - 1 - ti */
        bool seen = false;
        for (int ti = 0; ti < nt; ++ti) {
            uint8_t v;
            if ((vac >> ti) & 1ull) v = 1;
            else if ((tru >> ti) & 1ull) { seen = true; v = 1; }
            else v = seen ? 1 : 0;
            out[nt - 1 - ti] = v;
        }
    }
    for (int ti = nt; ti < tmax; ++ti) out[ti] = 0;
}

static int ifail(int code, const char* what, cudaError_t ce = cudaSuccess) {
    char buf[512];
    if (ce != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
    else snprintf(buf, sizeof buf, "%s", what);
    pdeb_set_last_error(buf);
    return code;
}

extern "C" int pdeb_iou_grids(const PdebIouClause* clauses, int64_t n_clauses, const double* tgrid, const double* xgrid,
                              const int32_t* ngrid, int64_t n_items, int32_t tmax, int32_t xmax, const double* m_vals, int32_t nm,
                              const double* b_vals, int32_t nb, uint8_t* bits_out, int device, void* stream_, PdebTiming* timing) {
    if (!clauses || !tgrid || !xgrid || !ngrid || !m_vals || !b_vals || !bits_out) return ifail(PDEB_E_ARG, "pdeb_iou_grids: NULL argument");
    if (n_clauses < 0 || n_items < 1 || tmax < 1 || tmax > 64 || xmax < 1 || nm < 1 || nb < 1)
        return ifail(PDEB_E_ARG, "pdeb_iou_grids: bad sizes (time grids hold at most 64 points)");
    for (int64_t i = 0; i < n_clauses; ++i)
        if (clauses[i].item < 0 || clauses[i].item >= n_items || clauses[i].op > PDEB_IOU_NE || clauses[i].kind > PDEB_IOU_DIV)
            return ifail(PDEB_E_ARG, "pdeb_iou_grids: bad clause record");
    for (int64_t i = 0; i < n_items; ++i)
        for (int q = 0; q < 4; ++q)
            if (ngrid[i * 4 + q] < 0 || ngrid[i * 4 + q] > (q < 2 ? tmax : xmax)) return ifail(PDEB_E_ARG, "pdeb_iou_grids: grid length out of range");
    if (timing) memset(timing, 0, sizeof *timing);
    if (n_clauses == 0) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return ifail(PDEB_E_NODEVICE, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) return ifail(PDEB_E_ARG, "bad device");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n_prof = nm * nb;
    const size_t b_cl = sizeof(PdebIouClause) * (size_t)n_clauses, b_t = 8 * (size_t)n_items * 2 * tmax, b_x = 8 * (size_t)n_items * 2 * xmax;
    const size_t b_n = 4 * (size_t)n_items * 4, b_m = 8 * (size_t)nm, b_b = 8 * (size_t)nb, b_out = (size_t)n_clauses * n_prof * tmax;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    unsigned char* d = nullptr;
    cudaError_t ce = cudaMalloc((void**)&d, up(b_cl) + up(b_t) + up(b_x) + up(b_n) + up(b_m) + up(b_b) + up(b_out));
    if (ce != cudaSuccess) { if (prev >= 0) cudaSetDevice(prev); return ifail(PDEB_E_NOMEM, "pdeb_iou_grids: device allocation", ce); }
    unsigned char* p = d;
    PdebIouClause* d_cl = (PdebIouClause*)p; p += up(b_cl);
    double* d_t = (double*)p; p += up(b_t);
    double* d_x = (double*)p; p += up(b_x);
    int32_t* d_n = (int32_t*)p; p += up(b_n);
    double* d_m = (double*)p; p += up(b_m);
    double* d_b = (double*)p; p += up(b_b);
    uint8_t* d_out = (uint8_t*)p;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
    ce = cudaMemcpyAsync(d_cl, clauses, b_cl, cudaMemcpyHostToDevice, stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_t, tgrid, b_t, cudaMemcpyHostToDevice, stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_x, xgrid, b_x, cudaMemcpyHostToDevice, stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_n, ngrid, b_n, cudaMemcpyHostToDevice, stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_m, m_vals, b_m, cudaMemcpyHostToDevice, stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_b, b_vals, b_b, cudaMemcpyHostToDevice, stream);
    if (ce == cudaSuccess) {
        if (timing) cudaEventRecord(e0, stream);
        const int64_t threads = n_clauses * n_prof;
        pdeb_iou_kernel<<<(unsigned)((threads + PDEB_IOU_THREADS - 1) / PDEB_IOU_THREADS), PDEB_IOU_THREADS, 0, stream>>>(
            d_cl, n_clauses, d_t, d_x, d_n, tmax, xmax, d_m, nm, d_b, nb, d_out);
        ce = cudaGetLastError();
        if (timing) cudaEventRecord(e1, stream);
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(bits_out, d_out, b_out, cudaMemcpyDeviceToHost, stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(stream);
    if (ce == cudaSuccess && timing) { cudaEventElapsedTime(&timing->kernel_ms, e0, e1); timing->total_ms = timing->kernel_ms; timing->launches = 1; }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d);
    if (prev >= 0) cudaSetDevice(prev);
    if (ce != cudaSuccess) return ifail(PDEB_E_CUDA, "pdeb_iou_grids", ce);
    return 0;
}
