/*
 * pdeb_kernels.cu — sm_100a kernels of the fused simulate + STL-robustness path and their launchers.
 *
 * Register family (pdeb_reg_kernel): one thread = one trajectory, state and LDL^T factors in
 * registers, clause slots in shared memory (see pdeb_core.h).  Replaces, for a whole batch at
 * once, the per-sample femformal run behind eval_robustness_wrapper.py:16 of the reference.
 */
#include <cuda_runtime.h>

#include "pdeb_core.h"
#include "pdeb_launch.h"

template <int KIND, int N, typename real>
__global__ void __launch_bounds__(PDEB_REG_THREADS)
pdeb_reg_kernel(const __grid_constant__ PdebDevBatch b, const __grid_constant__ PdebConventions cv,
                const __grid_constant__ PdebDevOut out, const int32_t* __restrict__ order, int64_t first, int64_t count) {
    extern __shared__ __align__(16) unsigned char pdeb_smem[];
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool live = t < count;
    int64_t s = 0;
    if (live) {
        s = order ? (int64_t)order[first + t] : first + t;
        /* a sample belongs to exactly one kernel variant; the others skip it */
        if (b.kind[s] != KIND || b.n_elem[s] != N) live = false;
    }
    if (!__any_sync(0xffffffffu, live)) return;
    PdebClauseSlots<real> slots = pdeb_clause_slots<real>(pdeb_smem, blockDim.x, b.pmax, threadIdx.x);
    pdeb_run_reg<KIND, N, real>(b, cv, out, s, slots, live);
}

/* samples no compute variant of this build covers: flag them, never drop them silently */
__global__ void pdeb_unsupported_kernel(const __grid_constant__ PdebDevBatch b, const __grid_constant__ PdebDevOut out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n) return;
    const int kind = b.kind[s], N = b.n_elem[s];
    if (pdeb_variant_of(kind, N) >= 0) return;
    out.robustness[s] = PDEB_FAILED_ROBUSTNESS;
    if (out.status) out.status[s] = PDEB_ST_BADPARAM;
    if (out.clause)
        for (int j = 0; j < b.pmax; ++j) out.clause[(int64_t)j * b.stride + s] = PDEB_FAILED_ROBUSTNESS;
}

template <int KIND, int N, typename real>
static cudaError_t launch_reg(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                              const int32_t* order, int64_t first, int64_t count, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    const size_t smem = pdeb_clause_smem_bytes<real>(PDEB_REG_THREADS, b.pmax);
    auto kern = pdeb_reg_kernel<KIND, N, real>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int64_t blocks = (count + PDEB_REG_THREADS - 1) / PDEB_REG_THREADS;
    kern<<<(unsigned)blocks, PDEB_REG_THREADS, smem, stream>>>(b, cv, out, order, first, count);
    return cudaGetLastError();
}

int pdeb_launch_all(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, int precision,
                    cudaStream_t stream, int* launches) {
    cudaError_t e = cudaSuccess;
    int nl = 0;
    if (b.n <= 0) { if (launches) *launches = 0; return 0; }
    {
        const int threads = 256;
        pdeb_unsupported_kernel<<<(unsigned)((b.n + threads - 1) / threads), threads, 0, stream>>>(b, out);
        e = cudaGetLastError();
        ++nl;
    }
    if (e == cudaSuccess) {
        if (precision == 32) e = launch_reg<PDEB_KIND_HEAT, 30, float>(b, cv, out, nullptr, 0, b.n, stream);
        else e = launch_reg<PDEB_KIND_HEAT, 30, double>(b, cv, out, nullptr, 0, b.n, stream);
        ++nl;
    }
    if (e == cudaSuccess) {
        if (precision == 32) e = launch_reg<PDEB_KIND_WAVE, 20, float>(b, cv, out, nullptr, 0, b.n, stream);
        else e = launch_reg<PDEB_KIND_WAVE, 20, double>(b, cv, out, nullptr, 0, b.n, stream);
        ++nl;
    }
    if (launches) *launches = nl;
    return e == cudaSuccess ? 0 : (int)e;
}
