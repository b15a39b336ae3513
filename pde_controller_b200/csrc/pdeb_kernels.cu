/*
 * pdeb_kernels.cu — sm_100a kernels of the fused simulate + STL-robustness path and their launchers.
 *
 * Register family (pdeb_reg_kernel): one thread = one trajectory, state and LDL^T factors in
 * registers, clause slots in shared memory (see pdeb_core.h).  Replaces, for a whole batch at
 * once, the per-sample femformal run behind eval_robustness_wrapper.py:16 of the reference.
 */
#include <cuda_runtime.h>

#include <mutex>

#include "pdeb_core.h"
#include "pdeb_launch.h"

#ifdef PDEB_REG_MAXNREG
#define PDEB_REG_BOUNDS __maxnreg__(PDEB_REG_MAXNREG)
#else
/* FP32 mode: half the registers and 544 instead of 904 bytes of shared memory per thread, so a third (fourth) block fits */
#ifndef PDEB_REG_MINBLOCKS_F32
#define PDEB_REG_MINBLOCKS_F32 3
#endif
#define PDEB_REG_BOUNDS __launch_bounds__(PDEB_REG_THREADS, sizeof(real) == 4 ? PDEB_REG_MINBLOCKS_F32 : PDEB_REG_MINBLOCKS)
#endif

template <int KIND, int N, typename real>
__global__ void PDEB_REG_BOUNDS
pdeb_reg_kernel(const __grid_constant__ PdebDevBatch b, const __grid_constant__ PdebConventions cv,
                const __grid_constant__ PdebDevOut out, const int32_t* __restrict__ order, int64_t first, int64_t count,
                int gather) {
    extern __shared__ __align__(16) unsigned char pdeb_smem[];
    const int64_t t = (int64_t)blockIdx.x * PDEB_REG_THREADS + threadIdx.x;
    const bool live = t < count;                      /* padding lanes of the last warp idle through the loop */
    /* gather = 0: `b` is the PLAN-ORDERED image: thread t of this launch reads position first + t of every array
     * (consecutive lanes, consecutive addresses); gather = 1: `b` is the image as uploaded and the thread reads its
     * sample's own index (a batch that is evaluated once is not worth permuting).  Results always go to the sample's own
     * index order[first + t]. */
    const int64_t so = live ? (int64_t)order[first + t] : 0;
    const int64_t s = live ? (gather ? so : first + t) : 0;
    typedef typename PdebVec2<real>::type real2;
    real2* table = reinterpret_cast<real2*>(pdeb_smem) + threadIdx.x;
    PdebClauseSlots<real, PDEB_REG_THREADS> slots = pdeb_clause_slots<real, PDEB_REG_THREADS>(
        pdeb_smem + pdeb_table_smem_bytes<real>(PDEB_REG_THREADS, N, KIND), b.pmax, threadIdx.x);
    real* stab = reinterpret_cast<real*>(pdeb_smem + sizeof(real2) * (size_t)PDEB_REG_THREADS * N) + threadIdx.x;
    real* ctab = nullptr;
#if PDEB_C_SMEM
    ctab = stab + (size_t)PDEB_REG_THREADS * N * (KIND == PDEB_KIND_WAVE ? 2 : 1);
#endif
    /* staged knot of the piecewise-linear input: two doubles per thread behind the clause slots (8-byte aligned:
     * every area before it is a multiple of 8 bytes per thread) */
    double* pwl_slot = reinterpret_cast<double*>(pdeb_smem + pdeb_table_smem_bytes<real>(PDEB_REG_THREADS, N, KIND) +
                                                 pdeb_clause_smem_bytes<real>(PDEB_REG_THREADS, b.pmax)) + threadIdx.x;
    pdeb_run_reg<KIND, N, real, PDEB_REG_THREADS>(b, cv, out, s, so, table, stab, slots, live, pwl_slot, ctab);
}

/*
 * Plan: order the samples so that every warp gets trajectories of the same kernel variant, (nearly) the same
 * number of time steps, longest first, and windows that open at about the same point of the run.  Three tiny kernels: keys + histogram,
 * exclusive scan of the PDEB_PLAN_BINS bins, scatter.  Samples no variant of this build covers
 * are flagged here (status BADPARAM, robustness -100), never dropped silently.
 */
__global__ void pdeb_plan_keys_kernel(const __grid_constant__ PdebDevBatch b, const __grid_constant__ PdebConventions cv,
                                      const __grid_constant__ PdebDevOut out, int32_t* __restrict__ keys,
                                      int32_t* __restrict__ hist) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= b.n) return;
    pdeb_plan_count(b, cv, out, s, keys, hist);
}

/*
 * One block: hist -> exclusive offsets (in place), per-variant start / count.  Only the bins of the variants the batch
 * holds are touched (`live`, bit v; the host knows the variants from two small arrays without reading anything back):
 * a default batch has one or two live variants of the 16 slots, and scanning all 65,536 bins from one block cost 73 us
 * per plan, a fifth of the heat kernel it precedes in the end-to-end path.  Per live variant: 4 bins per thread as one
 * 16-byte access, warp shuffles, two barriers.
 */
__global__ void __launch_bounds__(1024) pdeb_plan_scan_kernel(int32_t* __restrict__ hist, int32_t* __restrict__ vstart_count,
                                                              uint32_t live) {
    static_assert(PDEB_PLAN_SBINS == 4 * 1024, "four bins per thread");
    __shared__ int32_t wsum[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int32_t carry = 0;                                   /* samples in the variants before this one (same in every thread) */
    for (int v = 0; v < PDEB_NUM_VARIANTS; ++v) {
        const int32_t start = carry;
        if ((live >> v) & 1u) {
            int4* bins = reinterpret_cast<int4*>(hist + (size_t)v * PDEB_PLAN_SBINS) + t;
            const int4 c = *bins;
            const int32_t sum = c.x + c.y + c.z + c.w;
            int32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
            if (lane == 31) wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int32_t x = wsum[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                wsum[lane] = x;
            }
            __syncthreads();
            const int32_t before = carry + (warp > 0 ? wsum[warp - 1] : 0) + inc - sum;
            int4 e;
            e.x = before; e.y = before + c.x; e.z = e.y + c.y; e.w = e.z + c.z;
            *bins = e;
            carry += wsum[31];
            __syncthreads();                             /* wsum is rewritten by the next live variant */
        }
        if (t == 0) { vstart_count[2 * v] = start; vstart_count[2 * v + 1] = carry - start; }
    }
}

__global__ void pdeb_plan_scatter_kernel(int64_t n, const int32_t* __restrict__ keys, int32_t* __restrict__ cursor,
                                         int32_t* __restrict__ order) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int key = keys[s];
    if (key < 0) return;
    order[atomicAdd(&cursor[key], 1)] = (int32_t)s;
}

/*
 * Permutation of the resident image into PLAN ORDER: position p of `dst` receives sample order[p] of `src` (the image as
 * uploaded: the caller's order).  One thread per position; the reads are gathers from an image that has just been
 * written (L2), the writes are coalesced, and from then on every kernel launch reads its per-sample parameters as
 * consecutive lanes -> consecutive addresses.  Only the list entries a sample uses are moved.  knots_only: the refill
 * of pdeb_resident_set_knots.
 */
__global__ void pdeb_permute_kernel(const __grid_constant__ PdebDevBatch src, const __grid_constant__ PdebDevBatch dst,
                                    const int32_t* __restrict__ order, int64_t count, int knots_only) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= count) return;
    const int64_t s = order[p], ss = src.stride, ds = dst.stride;
    int nk = src.n_knots[s];
    ((int32_t*)dst.n_knots)[p] = nk;
    if (nk > src.kmax) nk = src.kmax;                  /* counts beyond the planes are flagged by the kernels */
    for (int j = 0; j < nk; ++j) {
        ((double*)dst.knot_y)[(int64_t)j * ds + p] = src.knot_y[(int64_t)j * ss + s];
        if (src.knot_t) ((double*)dst.knot_t)[(int64_t)j * ds + p] = src.knot_t[(int64_t)j * ss + s];
    }
    if (knots_only) return;
    ((uint8_t*)dst.kind)[p] = src.kind[s];
    ((int32_t*)dst.n_elem)[p] = src.n_elem[s];
    ((double*)dst.L)[p] = src.L[s];             ((double*)dst.dt)[p] = src.dt[s];
    ((double*)dst.T)[p] = src.T[s];             ((double*)dst.g0)[p] = src.g0[s];
    ((double*)dst.x_a)[p] = src.x_a[s];         ((double*)dst.x_b)[p] = src.x_b[s];
    ((double*)dst.rho_in)[p] = src.rho_in[s];   ((double*)dst.rho_out)[p] = src.rho_out[s];
    ((double*)dst.E_in)[p] = src.E_in[s];       ((double*)dst.E_out)[p] = src.E_out[s];
    int np = src.n_pred[s], nt = src.n_tree[s];
    ((int32_t*)dst.n_pred)[p] = np;
    ((int32_t*)dst.n_tree)[p] = nt;
    if (np > src.pmax) np = src.pmax;
    if (nt > src.tmax) nt = src.tmax;
    for (int j = 0; j < np; ++j) {
        const int64_t a = (int64_t)j * ss + s, d = (int64_t)j * ds + p;
        ((double*)dst.pred_xlo)[d] = src.pred_xlo[a];   ((double*)dst.pred_xhi)[d] = src.pred_xhi[a];
        ((double*)dst.pred_m)[d] = src.pred_m[a];       ((double*)dst.pred_b)[d] = src.pred_b[a];
        ((double*)dst.pred_tlo)[d] = src.pred_tlo[a];   ((double*)dst.pred_thi)[d] = src.pred_thi[a];
        ((uint8_t*)dst.pred_rel)[d] = src.pred_rel[a];  ((uint8_t*)dst.pred_top)[d] = src.pred_top[a];
    }
    for (int j = 0; j < nt; ++j) {
        const int64_t a = (int64_t)j * ss + s, d = (int64_t)j * ds + p;
        ((uint8_t*)dst.tree_op)[d] = src.tree_op[a];
        ((int8_t*)dst.tree_lhs)[d] = src.tree_lhs[a];
        ((int8_t*)dst.tree_rhs)[d] = src.tree_rhs[a];
    }
}

int pdeb_permute(const PdebDevBatch& src, const PdebDevBatch& dst, const int32_t* order, int64_t count, bool knots_only,
                 cudaStream_t stream) {
    if (count <= 0) return 0;
    pdeb_permute_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(src, dst, order, count, knots_only ? 1 : 0);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

int pdeb_plan(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
              uint32_t live_variants, cudaStream_t stream, int* launches, bool keys_done) {
    if (b.n <= 0) { if (launches) *launches = 0; return 0; }
    cudaError_t e = cudaSuccess;
    const int threads = 256;
    const unsigned blocks = (unsigned)((b.n + threads - 1) / threads);
    if (!keys_done) {
        e = cudaMemsetAsync(pb.hist, 0, sizeof(int32_t) * PDEB_PLAN_BINS, stream);
        if (e != cudaSuccess) return (int)e;
        pdeb_plan_keys_kernel<<<blocks, threads, 0, stream>>>(b, cv, out, pb.keys, pb.hist);
    }
    pdeb_plan_scan_kernel<<<1, 1024, 0, stream>>>(pb.hist, pb.vstart_count, live_variants);
    pdeb_plan_scatter_kernel<<<blocks, threads, 0, stream>>>(b.n, pb.keys, pb.hist, pb.order);
    if (launches) *launches = keys_done ? 2 : 3;
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <int KIND, int N, typename real>
static cudaError_t launch_reg(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                              const int32_t* order, int64_t first, int64_t count, int gather, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    const size_t smem = pdeb_table_smem_bytes<real>(PDEB_REG_THREADS, N, KIND) + pdeb_clause_smem_bytes<real>(PDEB_REG_THREADS, b.pmax) +
                        PDEB_PWL_SLOT_BYTES * PDEB_REG_THREADS;
    auto kern = pdeb_reg_kernel<KIND, N, real>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int64_t blocks = (count + PDEB_REG_THREADS - 1) / PDEB_REG_THREADS;
    kern<<<(unsigned)blocks, PDEB_REG_THREADS, smem, stream>>>(b, cv, out, order, first, count, gather);
    return cudaGetLastError();
}


/*
 * Side streams for batches that hold several kernel variants (a mixed heat + wave set): the variants are independent,
 * so the second one's blocks fill the SMs the first one's tail leaves idle (the longest trajectories of a launch
 * finish alone).  One small pool per device, created on first use; fork / join with events on the caller's stream.
 */
#define PDEB_SIDE_STREAMS 3
#define PDEB_MAX_DEVICES 64
struct PdebSidePool {
    cudaStream_t st[PDEB_SIDE_STREAMS] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[PDEB_SIDE_STREAMS] = {nullptr, nullptr, nullptr};
    bool ready = false;
};
static PdebSidePool g_side[PDEB_MAX_DEVICES];
static std::mutex g_side_mutex;

/* call with g_side_mutex held */
static PdebSidePool* side_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PDEB_MAX_DEVICES) return nullptr;
    PdebSidePool& p = g_side[dev];
    if (!p.ready) {
        cudaError_t e = cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming);
        for (int i = 0; i < PDEB_SIDE_STREAMS && e == cudaSuccess; ++i) {
            e = cudaStreamCreateWithFlags(&p.st[i], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p.join[i], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
        p.ready = true;
    }
    return &p;
}

static cudaError_t launch_variant(int v, const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out,
                                  const PdebPlanBuffers& pb, int64_t first, int64_t count, int precision, int gather, cudaStream_t stream) {
    switch (v) {
    case 0:
        return (precision == 32) ? launch_reg<PDEB_KIND_HEAT, 30, float>(b, cv, out, pb.order, first, count, gather, stream)
                                 : launch_reg<PDEB_KIND_HEAT, 30, double>(b, cv, out, pb.order, first, count, gather, stream);
    case 1:
        return (precision == 32) ? launch_reg<PDEB_KIND_WAVE, 20, float>(b, cv, out, pb.order, first, count, gather, stream)
                                 : launch_reg<PDEB_KIND_WAVE, 20, double>(b, cv, out, pb.order, first, count, gather, stream);
    default:
        if (v >= 2 && v <= 7) return pdeb_launch_chunk_heat(1 << (v - 2), b, cv, out, pb.order, first, count, gather, stream);
        if (v >= 8 && v <= 12) return pdeb_launch_chunk_wave(1 << (v - 8), b, cv, out, pb.order, first, count, gather, stream);
        if (v == PDEB_VARIANT_CTA_HEAT) return pdeb_launch_cta_heat(0, b, cv, out, pb.order, first, count, gather, stream);
        if (v == PDEB_VARIANT_CTA2_HEAT) return pdeb_launch_cta_heat(1, b, cv, out, pb.order, first, count, gather, stream);
        return cudaSuccess;
    }
}

int pdeb_launch_all(const PdebDevBatch& b, const PdebConventions& cv, const PdebDevOut& out, const PdebPlanBuffers& pb,
                    const int32_t* vstart_count_host, int precision, int gather, cudaStream_t stream, int* launches) {
    int live[PDEB_NUM_VARIANTS], nlive = 0;
    for (int v = 0; v < PDEB_NUM_VARIANTS; ++v) if (vstart_count_host[2 * v + 1] > 0) live[nlive++] = v;
    /* a stream that is being captured into a graph cannot be forked this way: keep those on one stream */
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (nlive > 1 && cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusActive; }
    /* the pool's events are shared by every caller on this device: the whole fork / launch / join sequence is enqueued
     * under the lock (an event wait takes the record that is current when the wait is enqueued) */
    std::unique_lock<std::mutex> lock(g_side_mutex, std::defer_lock);
    const bool fork = nlive > 1 && cap == cudaStreamCaptureStatusNone;
    if (fork) lock.lock();
    PdebSidePool* pool = fork ? side_pool() : nullptr;
    cudaError_t e = cudaSuccess;
    if (pool) e = cudaEventRecord(pool->fork, stream);
    int used = 0;
    /* the variant with the most work per sample last in the list goes first on the caller's stream: longest first */
    for (int q = 0; q < nlive && e == cudaSuccess; ++q) {
        const int v = live[q];
        const int64_t first = vstart_count_host[2 * v], count = vstart_count_host[2 * v + 1];
        cudaStream_t st = stream;
        if (pool && q > 0 && used < PDEB_SIDE_STREAMS) {
            st = pool->st[used];
            e = cudaStreamWaitEvent(st, pool->fork, 0);
            if (e != cudaSuccess) break;
        }
        e = launch_variant(v, b, cv, out, pb, first, count, precision, gather, st);
        if (st != stream && e == cudaSuccess) {
            e = cudaEventRecord(pool->join[used], st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, pool->join[used], 0);
            ++used;
        }
    }
    if (launches) *launches = nlive;
    return e == cudaSuccess ? 0 : (int)e;
}
