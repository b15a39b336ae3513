/* pdeb_cta_heat.cu — instantiations of the CTA family (block per trajectory, refined meshes) for the heat equation. */
#include "pdeb_cta.cuh"
#include "pdeb_launch.h"

/* shape 0: 4 warps x 8 rows per lane (every table in registers); shape 1: 2 warps x 16 rows (spikes and flux coefficients in
 * shared-memory columns) */
cudaError_t pdeb_launch_cta_heat(int shape, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                 const int32_t* order, int64_t first, int64_t count, int gather, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    if (shape == 0) pdeb_cta_kernel<PDEB_KIND_HEAT, 4, 8, PDEB_CTA_MINBLOCKS><<<(unsigned)count, 128, 0, stream>>>(b, cv, out, order, first, count, gather);
    else pdeb_cta_kernel<PDEB_KIND_HEAT, 2, 16, PDEB_CTA2_MINBLOCKS><<<(unsigned)count, 64, 0, stream>>>(b, cv, out, order, first, count, gather);
    return cudaGetLastError();
}
