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

/* which kernel variant owns a (kind, mesh) pair; -1 = none in this build */
PDEB_HD int pdeb_variant_of(int kind, int n_elem) {
    if (kind == PDEB_KIND_HEAT && n_elem == 30) return 0;
    if (kind == PDEB_KIND_WAVE && n_elem == 20) return 1;
    return -1;
}

#define PDEB_NUM_VARIANTS 2
#define PDEB_PLAN_SBINS 4096                                   /* step-count bins per variant */
#define PDEB_PLAN_BINS (PDEB_NUM_VARIANTS * PDEB_PLAN_SBINS)   /* multiple of 1024 */

/* device scratch of the plan (owned by the resident batch) */
struct PdebPlanBuffers {
    int32_t* keys;          /* [n] */
    int32_t* order;         /* [n] sample ids, grouped by variant, longest trajectories first */
    int32_t* hist;          /* [PDEB_PLAN_BINS] histogram, then scatter cursors */
    int32_t* vstart_count;  /* [2 * PDEB_NUM_VARIANTS] first index / count of each variant inside order */
};

/* order the batch (3 launches); flags samples no variant covers */
int pdeb_plan(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
              cudaStream_t stream, int* launches);

/* enqueue the compute kernels of every non-empty variant; returns 0 or a cudaError_t value */
int pdeb_launch_all(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
                    const int32_t* vstart_count_host, int precision, cudaStream_t stream, int* launches);

#endif
