/*
 * solve_probe.cu — development harness: times variants of the per-step tridiagonal solve of the register family in
 * isolation (no clauses, no input), at the occupancy of the real kernel (2 blocks x 128 threads per SM).
 * Prints SMSP-cycles per warp-step for each variant; the FP64-pipe floor is 2 cycles per FP64 warp instruction.
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../../pde_controller_b200/csrc -o solve_probe solve_probe.cu
 */
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <cuda_runtime.h>
#include "pdeb_core.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#define THREADS 128
constexpr int N = 30;
constexpr int M = (N + 1) / 2;

/* ---- variant 1: unit off-diagonal right-hand side (scaled unknowns), c in registers, table {rd, dinv} ---- */
struct V1 {
    double c[N];
    PdebD2* tab;
    __device__ __forceinline__ void step(const double* x, double fN, double* z) const {
        double zt = 0, zb = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int it = 1 + j, ib = N - j;
            if (it < M) {
                const PdebD2 t = tab[(it - 1) * THREADS];
                double r = fma(t.x, x[it], x[it - 1] + x[it + 1]);
                if (it > 1) r = fma(-c[it - 1], zt, r);
                zt = r;
                z[it] = r * t.y;
            }
            if (ib > M) {
                const PdebD2 t = tab[(ib - 1) * THREADS];
                double r = fma(t.x, x[ib], x[ib - 1] + ((ib == N) ? fN : x[ib + 1]));
                if (ib < N) r = fma(-c[ib], zb, r);
                zb = r;
                z[ib] = r * t.y;
            }
        }
        {
            const PdebD2 t = tab[(M - 1) * THREADS];
            double r = fma(t.x, x[M], x[M - 1] + x[M + 1]);
            r = fma(-c[M - 1], zt, r);
            r = fma(-c[M], zb, r);
            z[M] = r * t.y;
        }
#pragma unroll
        for (int j = 1; j < N; ++j) {
            const int it = M - j, ib = M + j;
            if (it >= 1) z[it] = fma(-c[it], z[it + 1], z[it]);
            if (ib <= N) z[ib] = fma(-c[ib - 1], z[ib - 1], z[ib]);
        }
    }
};

/* ---- variant 2: as V1 but the right-hand sides of ALL rows are formed first (one independent batch), then the chains ---- */
struct V2 {
    double c[N];
    PdebD2* tab;
    __device__ __forceinline__ void step(const double* x, double fN, double* z) const {
        double dinv[N + 1];
#pragma unroll
        for (int i = 1; i <= N; ++i) {
            const PdebD2 t = tab[(i - 1) * THREADS];
            z[i] = fma(t.x, x[i], x[i - 1] + ((i == N) ? fN : x[i + 1]));
            dinv[i] = t.y;
        }
        double zt = 0, zb = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int it = 1 + j, ib = N - j;
            if (it < M) { double r = z[it]; if (it > 1) r = fma(-c[it - 1], zt, r); zt = r; z[it] = r * dinv[it]; }
            if (ib > M) { double r = z[ib]; if (ib < N) r = fma(-c[ib], zb, r); zb = r; z[ib] = r * dinv[ib]; }
        }
        { double r = fma(-c[M - 1], zt, z[M]); r = fma(-c[M], zb, r); z[M] = r * dinv[M]; }
#pragma unroll
        for (int j = 1; j < N; ++j) {
            const int it = M - j, ib = M + j;
            if (it >= 1) z[it] = fma(-c[it], z[it + 1], z[it]);
            if (ib <= N) z[ib] = fma(-c[ib - 1], z[ib - 1], z[ib]);
        }
    }
};

template <int VAR>
__global__ void __launch_bounds__(THREADS, 2) probe(double* out, const double* par, int steps) {
    extern __shared__ __align__(16) unsigned char smem[];
    PdebD2* table = reinterpret_cast<PdebD2*>(smem) + threadIdx.x;
    const int t = blockIdx.x * THREADS + threadIdx.x;
    const double L = par[0] + (t & 63), x_a = par[1] + (t & 15), x_b = par[2] + (t & 31);
    double w[N + 2], z[N + 2];
#pragma unroll
    for (int i = 0; i <= N + 1; ++i) { w[i] = 0; z[i] = 0; }
    double fN = par[7];
    double* stab0 = reinterpret_cast<double*>(smem + sizeof(PdebD2) * THREADS * N) + threadIdx.x;
    if (VAR == 0) {
        PdebTrajReg<PDEB_KIND_HEAT, N, double, THREADS> tr;
        tr.tab = table; tr.stab = stab0; tr.ltab = stab0;
        tr.setup(L, x_a, x_b, par[3], par[4], par[5], par[6], 0, 0, 0.025, 0.025);
        for (int k = 0; k < steps; k += 2) {
            tr.step_solve(w, fN, z);
            tr.step_solve(z, fN, w);
            fN += 1.0;
        }
    } else {
        /* the scaled variants only need plausible SPD factors: reuse the real set-up, then treat R as unit off-diagonal */
        PdebTrajReg<PDEB_KIND_HEAT, N, double, THREADS> tr;
        tr.tab = table; tr.stab = stab0; tr.ltab = stab0;
        tr.setup(L, x_a, x_b, par[3], par[4], par[5], par[6], 0, 0, 0.025, 0.025);
        typename std::conditional<VAR == 1, V1, V2>::type v;
        v.tab = table;
#pragma unroll
        for (int i = 1; i < N; ++i) v.c[i] = tr.c[i];
        v.c[0] = 0;
#pragma unroll
        for (int i = 1; i <= N; ++i) { PdebD2 q = table[(i - 1) * THREADS]; q.x = -1.9 * (1.0 + 1e-3 * i); table[(i - 1) * THREADS] = q; }
        fN *= 1e-6;
        for (int k = 0; k < steps; k += 2) {
            v.step(w, fN, z);
            v.step(z, fN, w);
            fN += 1e-6;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 1; i <= N; ++i) s += w[i];
    out[t] = s;
}

/* solve + one clause pass per step through the real device functions (scaled state, PTX masked max) */
template <int PASSES>
__global__ void __launch_bounds__(THREADS, 2) probe_clause(double* out, const double* par, int steps) {
    extern __shared__ __align__(16) unsigned char smem[];
    PdebD2* table = reinterpret_cast<PdebD2*>(smem) + threadIdx.x;
    double* stab = reinterpret_cast<double*>(smem + sizeof(PdebD2) * THREADS * N) + threadIdx.x;
    const int t = blockIdx.x * THREADS + threadIdx.x;
    const double L = par[0] + (t & 63), x_a = par[1] + (t & 15), x_b = par[2] + (t & 31);
    PdebTrajReg<PDEB_KIND_HEAT, N, double, THREADS> tr;
    tr.tab = table; tr.stab = stab; tr.ltab = stab;
#pragma unroll
    for (int i = 0; i <= N + 1; ++i) tr.w[i] = 0;
    tr.setup(L, x_a, x_b, par[3], par[4], par[5], par[6], 0, 0, 0.025, 0.025);
    double fN = par[7] * tr.sN;
    const double mh = par[1] * 1e-3, c0 = par[2] * 1e-2;
    const uint32_t nm = 0x00ffff00u >> (t & 7), rel = t % 3;
    const uint32_t and_mask = (rel == PDEB_REL_EQ) ? 0x7fffffffu : 0xffffffffu;
    const uint32_t xor_mask = (rel == PDEB_REL_GT) ? 0x80000000u : 0u;
    double acc = INFINITY;
    double z[N + 2];
    z[0] = 0; z[N + 1] = 0;
    for (int k = 0; k < steps; ++k) {
#pragma unroll
        for (int q = 0; q < PASSES; ++q) {
            const double g = pdeb_clause_gmax<N, double, THREADS>(tr.w, tr.stab, mh + q, c0, nm, and_mask, xor_mask);
            const double r = -g;
            if (r < acc) acc = r;
        }
        tr.step_solve(tr.w, fN, z);
#pragma unroll
        for (int i = 1; i <= N; ++i) tr.w[i] = z[i];
        fN += 1e-3;
    }
    out[t] = acc;
}

template <int PASSES>
static void run_clause(const char* name, int sms, double* dout, double* dpar, int clock_khz) {
    const size_t smem = (sizeof(PdebD2) + sizeof(double)) * THREADS * N;
    CK(cudaFuncSetAttribute(probe_clause<PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int steps = 2000, blocks = sms * 2;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_clause<PASSES><<<blocks, THREADS, smem>>>(dout, dpar, steps);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe_clause<PASSES><<<blocks, THREADS, smem>>>(dout, dpar, steps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cycles = ms * 1e-3 * clock_khz * 1e3;
    printf("%-28s %8.3f ms   %7.1f SMSP-cycles per warp-step\n", name, ms, cycles / steps / 2.0);
}

template <int VAR>
static void run(const char* name, int sms, double* dout, double* dpar, int clock_khz) {
    const size_t smem = (sizeof(PdebD2) + sizeof(double)) * THREADS * N;
    CK(cudaFuncSetAttribute(probe<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int steps = 2000, blocks = sms * 2;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe<VAR><<<blocks, THREADS, smem>>>(dout, dpar, steps);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe<VAR><<<blocks, THREADS, smem>>>(dout, dpar, steps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cycles = ms * 1e-3 * clock_khz * 1e3;
    /* 2 warps per SMSP run concurrently: SMSP-cycles per warp-step = elapsed / steps / 2 */
    printf("%-28s %8.3f ms   %7.1f SMSP-cycles per warp-step\n", name, ms, cycles / steps / 2.0);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    double h[8] = { 100.0, 20.0, 40.0, 1800.0, 2100.0, 9e5, 1.2e6, 3e5 };
    double *dpar, *dout;
    CK(cudaMalloc(&dpar, sizeof h)); CK(cudaMemcpy(dpar, h, sizeof h, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dout, sizeof(double) * p.multiProcessorCount * 2 * THREADS));
    run<0>("v0 library step_solve", p.multiProcessorCount, dout, dpar, p.clockRate);
    run<1>("v1 unit off-diagonal", p.multiProcessorCount, dout, dpar, p.clockRate);
    run<2>("v2 unit, rhs batch first", p.multiProcessorCount, dout, dpar, p.clockRate);
    run_clause<0>("real solve, 0 clause passes", p.multiProcessorCount, dout, dpar, p.clockRate);
    run_clause<1>("real solve + 1 clause pass", p.multiProcessorCount, dout, dpar, p.clockRate);
    run_clause<2>("real solve + 2 clause passes", p.multiProcessorCount, dout, dpar, p.clockRate);
    return 0;
}
