/* pdeb_chunk_heat.cu — instantiations of the chunk family (warp per trajectory) for the heat equation. */
#include "pdeb_chunk.cuh"
#include "pdeb_launch.h"

template <int R>
static cudaError_t launch(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const int32_t* order,
                          int64_t first, int64_t count, int gather, cudaStream_t stream) {
    const size_t smem = pdeb_chunk_smem_bytes<R>();
    auto kern = pdeb_chunk_kernel<PDEB_KIND_HEAT, R>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int64_t blocks = (count + PDEB_CHUNK_WARPS - 1) / PDEB_CHUNK_WARPS;
    kern<<<(unsigned)blocks, 32 * PDEB_CHUNK_WARPS, smem, stream>>>(b, cv, out, order, first, count, gather);
    return cudaGetLastError();
}

cudaError_t pdeb_launch_chunk_heat(int R, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                   const int32_t* order, int64_t first, int64_t count, int gather, cudaStream_t stream) {
    switch (R) {
    case 1: return launch<1>(b, cv, out, order, first, count, gather, stream);
    case 2: return launch<2>(b, cv, out, order, first, count, gather, stream);
    case 4: return launch<4>(b, cv, out, order, first, count, gather, stream);
    case 8: return launch<8>(b, cv, out, order, first, count, gather, stream);
    case 16: return launch<16>(b, cv, out, order, first, count, gather, stream);
    case 32: return launch<32>(b, cv, out, order, first, count, gather, stream);
    default: return cudaErrorInvalidValue;
    }
}
