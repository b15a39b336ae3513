/* pdeb_launch.h — interface between the C-ABI layer (pdeb_api.cu) and the kernels (pdeb_kernels.cu). */
#ifndef PDEB_LAUNCH_H
#define PDEB_LAUNCH_H

#include <cuda_runtime.h>

#include "pdeb_core.h"

#ifndef PDEB_REG_THREADS
#define PDEB_REG_THREADS 128
#endif
#ifndef PDEB_REG_MINBLOCKS
#define PDEB_REG_MINBLOCKS 2   /* resident blocks per SM the register family is compiled for */
#endif

/* sets the thread-local message pdeb_last_error() returns (pdeb_api.cu) */
void pdeb_set_last_error(const char* msg);

/*
 * Kernel variants.  0: heat N=30 and 1: wave N=20 are the register family (thread per trajectory);
 * 2..7: chunk family heat with R = 1,2,4,8,16,32 rows per lane; 8..12: chunk family wave, R = 1..16.
 * A mesh of N elements needs 32 R + 30 >= N.
 */
#define PDEB_NUM_VARIANTS 15
#define PDEB_VARIANT_SLOTS 16
/* 13: CTA family (one block of 128 threads per trajectory, 8 rows per lane, two-level substructuring; pdeb_cta.cuh), heat
 * meshes beyond 542 elements (BASELINE.json configs[3]).  Measured on B200 it is SLOWER than the chunk family at R = 32 on
 * that configuration (62.8 ms chunk, 68.4 ms 2 x 16, 91.9 ms 4 x 8; profiles/r02_cta_refined_summary.txt: the per-step fixed
 * cost -- cyclic reduction, exchange, barrier, clause pass -- is paid by every warp of the block), so it is an OPT-IN:
 * PDEB_FLAG_CTA / PDEB_FLAG_CTA2 in the image's flags (environment PDEB_CTA=1 / 2 when the resident image is created).  Kept
 * because it is the mapping the north star names and the parity tests cover it. */
#define PDEB_VARIANT_CTA_HEAT 13   /* 4 warps x 8 rows per lane */
#define PDEB_VARIANT_CTA2_HEAT 14  /* 2 warps x 16 rows per lane */
#define PDEB_FLAG_CTA 1            /* PdebDevBatch.pad: refined heat meshes go to the CTA family, 4 x 8 shape (PDEB_CTA=1) */
#define PDEB_FLAG_CTA2 2           /* ... 2 x 16 shape (PDEB_CTA=2) */
PDEB_HD int pdeb_variant_of(int kind, int n_elem, int flags = 0) {
    if (n_elem < 1) return -1;
    if (kind == PDEB_KIND_HEAT) {
        if (n_elem == 30) return 0;
        if ((flags & PDEB_FLAG_CTA) && n_elem > 32 * 16 + 30 && n_elem <= PDEB_MAX_ELEM) return PDEB_VARIANT_CTA_HEAT;
        if ((flags & PDEB_FLAG_CTA2) && n_elem > 32 * 16 + 30 && n_elem <= PDEB_MAX_ELEM) return PDEB_VARIANT_CTA2_HEAT;
        for (int q = 0; q < 6; ++q) if (n_elem <= 32 * (1 << q) + 30) return 2 + q;
        return -1;
    }
    if (kind == PDEB_KIND_WAVE) {
        if (n_elem == 20) return 1;
        for (int q = 0; q < 5; ++q) if (n_elem <= 32 * (1 << q) + 30) return 8 + q;
        return -1;
    }
    return -1;
}

/*
 * Plan bins of one variant: PDEB_PLAN_SBUCKETS buckets of PDEB_PLAN_SWIDTH steps (longest first), each split into
 * PDEB_PLAN_FBINS classes of "where in the run the first window opens" (pdeb_plan_key).
 */
#define PDEB_PLAN_SBINS 4096
#ifndef PDEB_PLAN_FBINS
#define PDEB_PLAN_FBINS 16
#endif
#ifndef PDEB_PLAN_SWIDTH
#define PDEB_PLAN_SWIDTH 8
#endif
#define PDEB_PLAN_SBUCKETS (PDEB_PLAN_SBINS / PDEB_PLAN_FBINS)
#define PDEB_PLAN_BINS (PDEB_VARIANT_SLOTS * PDEB_PLAN_SBINS)   /* multiple of 1024 */

#if defined(__CUDACC__)
/* sort key of sample s (pdeb_plan): variant, step bucket (longest first), where the first window opens; -1: no variant */
__device__ __forceinline__ int pdeb_plan_key(const PdebDevBatch& b, const PdebConventions& cv, int64_t s) {
    const int kind = b.kind[s], N = b.n_elem[s];
    const int v = pdeb_variant_of(kind, N, b.pad);
    if (v < 0) return -1;
    const double T = b.T[s], dt = b.dt[s];
    int64_t S = 1;
    int sub = 0;
    if (T > 0 && dt > 0 && isfinite(T) && isfinite(dt)) {
        S = pdeb_step_count(T, dt, cv.steps_rule);
        /* the kernels stop at the last observed step: order by that */
        const int np = b.n_pred[s];
        double last = 0.0, first = 4.0e9;
        for (int j = 0; j < np && j < b.pmax; ++j) {
            const double q = b.pred_thi[(int64_t)j * b.stride + s] / dt;
            const double a1 = (cv.window_rule == 1) ? trunc(q) : floor(q + PDEB_EPS);
            if (a1 > last) last = a1;
            const double p = b.pred_tlo[(int64_t)j * b.stride + s] / dt;
            const double a0 = (cv.window_rule == 1) ? trunc(p) : ceil(p - PDEB_EPS);
            if (a0 < first) first = a0;
        }
        if (last < (double)S) S = (int64_t)last;
        if (S < 1) S = 1;
        /* second key: where in the run the first window opens.  A warp runs a clause pass in every step in which ANY
         * of its lanes has an open window, so lanes whose windows open together share their idle prefix. */
        if (first > 0.0) {
            const double f = first * PDEB_PLAN_FBINS / ((double)S + 1.0);
            sub = (f >= (double)(PDEB_PLAN_FBINS - 1)) ? PDEB_PLAN_FBINS - 1 : (int)f;
        }
    }
    /* first key: steps in buckets, descending inside a variant (longest first).  Half of the buckets are
     * PDEB_PLAN_SWIDTH steps wide, the next quarters 8x and 512x that (then one bucket for everything longer) */
    constexpr int64_t H = PDEB_PLAN_SBUCKETS / 2, Q = PDEB_PLAN_SBUCKETS / 4, W = PDEB_PLAN_SWIDTH;
    int64_t sb = S / W;
    if (sb >= H) {
        sb = H + (S - H * W) / (8 * W);
        if (sb >= H + Q) sb = H + Q + (S - H * W - Q * 8 * W) / (512 * W);
    }
    if (sb > PDEB_PLAN_SBUCKETS - 1) sb = PDEB_PLAN_SBUCKETS - 1;
    return v * PDEB_PLAN_SBINS + (PDEB_PLAN_SBUCKETS - 1 - (int)sb) * PDEB_PLAN_FBINS + sub;
}

/* key + histogram of sample s; a sample no variant covers is flagged here (status BADPARAM, robustness -100).
 * Shared by the plan's keys kernel and by the unpack kernel of the packed transfer form, which has just written the
 * sample it counts (a thread reads its own writes). */
__device__ __forceinline__ void pdeb_plan_count(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, int64_t s,
                                                int32_t* __restrict__ keys, int32_t* __restrict__ hist) {
    const int key = pdeb_plan_key(b, cv, s);
    keys[s] = key;
    if (key >= 0) { atomicAdd(&hist[key], 1); return; }
    out.robustness[s] = PDEB_FAILED_ROBUSTNESS;
    if (out.status) out.status[s] = PDEB_ST_BADPARAM;
    if (out.clause)
        for (int j = 0; j < b.pmax; ++j) out.clause[(int64_t)j * b.stride + s] = PDEB_FAILED_ROBUSTNESS;
}
#endif

/* gather: 0 = `b` is the plan-ordered image (thread t reads position first + t), 1 = `b` is the image as uploaded
 * (thread t reads its sample's own index order[first + t]) */
cudaError_t pdeb_launch_chunk_heat(int R, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                   const int32_t* order, int64_t first, int64_t count, int gather, cudaStream_t stream);
cudaError_t pdeb_launch_cta_heat(int shape, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                 const int32_t* order, int64_t first, int64_t count, int gather, cudaStream_t stream);
cudaError_t pdeb_launch_chunk_wave(int R, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                   const int32_t* order, int64_t first, int64_t count, int gather, cudaStream_t stream);

/* device scratch of the plan (owned by the resident batch) */
struct PdebPlanBuffers {
    int32_t* keys;          /* [n] */
    int32_t* order;         /* [n] sample ids, grouped by variant, longest trajectories first */
    int32_t* hist;          /* [PDEB_PLAN_BINS] histogram, then scatter cursors */
    int32_t* vstart_count;  /* [2 * PDEB_NUM_VARIANTS] first index / count of each variant inside order */
};

/* order the batch (3 launches); flags samples no variant covers.  live_variants: bit v set if the batch may hold samples
 * of variant v (known on the host from kind / n_elem); only those variants' bins are scanned.  keys_done: the keys and
 * the histogram have already been formed on `stream` (pdeb_plan_count from the unpack kernel): scan + scatter only */
int pdeb_plan(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
              uint32_t live_variants, cudaStream_t stream, int* launches, bool keys_done = false);

/* dst[p] <- src[order[p]], p < count: the resident image in plan order (see pdeb_permute_kernel) */
int pdeb_permute(const PdebDevBatch& src, const PdebDevBatch& dst, const int32_t* order, int64_t count, bool knots_only,
                 cudaStream_t stream);

/* enqueue the compute kernels of every non-empty variant over the image `b` (gather: see above); returns 0 or a cudaError_t value */
int pdeb_launch_all(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
                    const int32_t* vstart_count_host, int precision, int gather, cudaStream_t stream, int* launches);

#endif
