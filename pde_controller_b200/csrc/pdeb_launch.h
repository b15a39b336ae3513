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
#define PDEB_NUM_VARIANTS 13
#define PDEB_VARIANT_SLOTS 16
PDEB_HD int pdeb_variant_of(int kind, int n_elem) {
    if (n_elem < 1) return -1;
    if (kind == PDEB_KIND_HEAT) {
        if (n_elem == 30) return 0;
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

cudaError_t pdeb_launch_chunk_heat(int R, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                   const int32_t* order, int64_t first, int64_t count, cudaStream_t stream);
cudaError_t pdeb_launch_chunk_wave(int R, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                   const int32_t* order, int64_t first, int64_t count, cudaStream_t stream);

/* device scratch of the plan (owned by the resident batch) */
struct PdebPlanBuffers {
    int32_t* keys;          /* [n] */
    int32_t* order;         /* [n] sample ids, grouped by variant, longest trajectories first */
    int32_t* hist;          /* [PDEB_PLAN_BINS] histogram, then scatter cursors */
    int32_t* vstart_count;  /* [2 * PDEB_NUM_VARIANTS] first index / count of each variant inside order */
};

/* order the batch (3 launches); flags samples no variant covers.  live_variants: bit v set if the batch may hold samples
 * of variant v (known on the host from kind / n_elem); only those variants' bins are scanned */
int pdeb_plan(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
              uint32_t live_variants, cudaStream_t stream, int* launches);

/* enqueue the compute kernels of every non-empty variant; returns 0 or a cudaError_t value */
int pdeb_launch_all(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
                    const int32_t* vstart_count_host, int precision, cudaStream_t stream, int* launches);

#endif
