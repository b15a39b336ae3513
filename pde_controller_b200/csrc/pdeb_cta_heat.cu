/* pdeb_cta_heat.cu — instantiation of the CTA family (block per trajectory, refined meshes) for the heat equation. */
#include "pdeb_cta.cuh"
#include "pdeb_launch.h"

cudaError_t pdeb_launch_cta_heat(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                 const int32_t* order, int64_t first, int64_t count, int gather, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    pdeb_cta_kernel<PDEB_KIND_HEAT><<<(unsigned)count, PDEB_CTA_THREADS, 0, stream>>>(b, cv, out, order, first, count, gather);
    return cudaGetLastError();
}
