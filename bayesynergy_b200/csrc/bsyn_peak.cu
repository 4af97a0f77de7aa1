// bsyn_peak.cu -- register-resident DFMA microbenchmark: the measured FP64 roofline denominator
// (MEASURED_PEAKS.json has no FP64 figure, so it is measured on the box by bench.py).
#include "../../include/bsyn.h"
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// fma_per_sec: achieved double-precision FMA instructions (per lane) per second; sm_mhz unused here.
extern "C" int bsyn_measure_fp64_peak(int32_t device, double *fma_per_sec)
{
    if (!fma_per_sec) return BSYN_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BSYN_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BSYN_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *d = nullptr;
    if (cudaMalloc(&d, sizeof(double) * blocks * threads) != cudaSuccess) return BSYN_ERR_ALLOC;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return BSYN_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = (double)blocks * threads * iters * 64.0;
        if (rep > 0 && ms > 0) best = fmas / (ms * 1e-3) > best ? fmas / (ms * 1e-3) : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *fma_per_sec = best;
    return BSYN_OK;
}
