/*
 * fp64_probe.cu — measures, on the GPU it runs on, the numbers the fused kernels are bounded by:
 *   - dependent-issue latency of DFMA, DSETP+select (the masked max of the clause reduction), LDS.128
 *   - FP64 FMA throughput per SM as a function of resident warps per scheduler and independent chains per thread
 *   - peak DFMA rate of the whole device (the denominator of the "fp64" roofline bench.py reports)
 * Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu ; prints one JSON object.
 */
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int CHAINS>
__global__ void dfma_kernel(double* out, double a, double b, int iters, long long* cyc) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void dsetp_sel_kernel(double* out, const double* in, int iters, long long* cyc) {
    double hi = -1e300;
    double g[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) g[u] = in[threadIdx.x * 8 + u];
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %0;\n\tselp.f64 %0, %1, %0, p;\n\t}" : "+d"(hi) : "d"(g[u]));
        }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = hi;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

__global__ void lds_kernel(double* out, int iters, long long* cyc) {
    __shared__ double2 tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = make_double2((double)((i * 7 + 1) & 255), 0.0);
    __syncthreads();
    int idx = threadIdx.x & 255;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) idx = (int)tab[idx].x;
    }
    const long long t1 = clock64();
    out[threadIdx.x] = idx;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
static double run_dfma(int blocks, int threads, int iters, double* dout, long long* dcyc, double* cycles_per_iter) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    dfma_kernel<CHAINS><<<blocks, threads>>>(dout, 0.999999, 1e-7, iters, dcyc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    dfma_kernel<CHAINS><<<blocks, threads>>>(dout, 0.999999, 1e-7, iters, dcyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long cyc;
    CK(cudaMemcpy(&cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost));
    if (cycles_per_iter) *cycles_per_iter = (double)cyc / ((double)iters * 8);
    return (double)blocks * threads * iters * 8.0 * CHAINS / (ms * 1e-3);   /* DFMA lane-ops per second */
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* dout; long long* dcyc; double* din;
    CK(cudaMalloc(&dout, sizeof(double) * 1024 * 4096));
    CK(cudaMalloc(&din, sizeof(double) * 8 * 1024));
    CK(cudaMemset(din, 0, sizeof(double) * 8 * 1024));
    CK(cudaMalloc(&dcyc, sizeof(long long)));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d", p.name, sms, p.clockRate / 1000);

    double cpi;
    run_dfma<1>(1, 32, 20000, dout, dcyc, &cpi);
    printf(", \"dfma_dependent_latency_cycles\": %.2f", cpi);
    long long cyc;
    dsetp_sel_kernel<<<1, 32>>>(dout, din, 20000, dcyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost));
    printf(", \"dsetp_selp_chain_cycles\": %.2f", (double)cyc / (20000.0 * 8));
    lds_kernel<<<1, 32>>>(dout, 20000, dcyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&cyc, dcyc, sizeof cyc, cudaMemcpyDeviceToHost));
    printf(", \"lds128_dependent_latency_cycles\": %.2f", (double)cyc / (20000.0 * 8));

    /* per-SM throughput: w warps per scheduler (4 schedulers), c chains per thread; one block per SM */
    printf(", \"dfma_per_clk_per_sm\": {");
    bool first = true;
    for (int w = 1; w <= 4; ++w) {
        const int threads = 32 * 4 * w;
        double r1 = run_dfma<1>(sms, threads, 4000, dout, dcyc, nullptr);
        double r2 = run_dfma<2>(sms, threads, 4000, dout, dcyc, nullptr);
        double r4 = run_dfma<4>(sms, threads, 4000, dout, dcyc, nullptr);
        double r8 = run_dfma<8>(sms, threads, 4000, dout, dcyc, nullptr);
        const double den = (double)sms * p.clockRate * 1e3;
        printf("%s\"w%d\": {\"c1\": %.1f, \"c2\": %.1f, \"c4\": %.1f, \"c8\": %.1f}", first ? "" : ", ", w, r1 / den, r2 / den, r4 / den, r8 / den);
        first = false;
    }
    printf("}");
    /* device peak: 16 warps per SM x 8 chains */
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        const double r = run_dfma<8>(sms * 2, 512, 8000, dout, dcyc, nullptr);
        if (r > best) best = r;
    }
    printf(", \"dfma_peak_per_s\": %.4e, \"fp64_peak_tflops\": %.2f}\n", best, 2.0 * best / 1e12);
    return 0;
}
