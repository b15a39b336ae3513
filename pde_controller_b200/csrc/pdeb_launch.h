/* pdeb_launch.h — interface between the C-ABI layer (pdeb_api.cu) and the kernels (pdeb_kernels.cu). */
#ifndef PDEB_LAUNCH_H
#define PDEB_LAUNCH_H

#include <cuda_runtime.h>

#include "pdeb_core.h"

#define PDEB_REG_THREADS 128

/* which kernel variant owns a (kind, mesh) pair; -1 = none in this build */
PDEB_HD int pdeb_variant_of(int kind, int n_elem) {
    if (kind == PDEB_KIND_HEAT && n_elem == 30) return 0;
    if (kind == PDEB_KIND_WAVE && n_elem == 20) return 1;
    return -1;
}

/* enqueue every kernel needed for batch b on `stream`; returns 0 or a cudaError_t value */
int pdeb_launch_all(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, int precision,
                    cudaStream_t stream, int* launches);

#endif
