// dmma_check.cu -- is the FP64 tensor-core MMA (mma.sync.m8n8k4.f64) worth using for the 10 x 10 Kronecker products?
// Measures on the device: latency of a dependent DMMA chain, throughput with 1..8 warps per scheduler issuing
// independent DMMAs, the same for DFMA, and checks one product against a scalar reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_check tools/dmma_check.cu && ./dmma_check
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b, double c0, double c1)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}

template <int ILP> __global__ void k_dmma(double *out, int iters, long long *cyc)
{
    double c0[ILP], c1[ILP];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = i; c1[i] = -i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a, b, c0[i], c1[i]);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> __global__ void k_dfma(double *out, int iters, long long *cyc)
{
    double c[ILP];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// correctness: D(8x8) = A(8x4) B(4x8) with the fragment layout the kernels will use
__global__ void k_check(const double *A, const double *B, double *D)
{
    const int l = threadIdx.x;
    double d0 = 0, d1 = 0;
    dmma(d0, d1, A[(l >> 2) * 4 + (l & 3)], B[(l & 3) * 8 + (l >> 2)], 0.0, 0.0);
    D[(l >> 2) * 8 + 2 * (l & 3)] = d0;
    D[(l >> 2) * 8 + 2 * (l & 3) + 1] = d1;
}

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    // latency: 1 warp, dependent chain
    k_dmma<1><<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA dependent chain: %.1f cycles per DMMA\n", (double)h / iters);
    k_dfma<1><<<1, 32>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent chain: %.1f cycles per DFMA\n", (double)h / iters);
    // throughput per SM: W warps per CTA, one CTA, ILP 4
    for (int w : {1, 2, 4, 8, 12, 16}) {
        k_dmma<4><<<1, 32 * w>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double dm = (double)h / (iters * 4.0 * w);
        k_dfma<4><<<1, 32 * w>>>(out, iters, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double df = (double)h / (iters * 4.0 * w);
        printf("%2d warps/SM: %.2f cycles per DMMA per SM (= %.1f FMA/cycle/SM), %.2f cycles per DFMA per SM (= %.1f FMA/cycle/SM)\n", w, dm,
               256.0 / dm, df, 32.0 / df);
    }
    // whole chip
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k_dmma<4><<<148 * 4, 256>>>(out, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("chip DMMA: %.2f TFLOP/s\n", 2.0 * 256.0 * 4 * iters * 148 * 4 * 8 / (ms * 1e-3) / 1e12);
    cudaEventRecord(e0); k_dfma<8><<<148 * 4, 256>>>(out, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("chip DFMA: %.2f TFLOP/s\n", 2.0 * 32.0 * 8 * iters * 148 * 4 * 8 / (ms * 1e-3) / 1e12);
    double hA[32], hB[32], hD[64], *dA, *dB, *dD;
    for (int i = 0; i < 32; ++i) { hA[i] = 0.1 * i + 1; hB[i] = 1.0 / (i + 1); }
    cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dD, 512);
    cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice);
    k_check<<<1, 32>>>(dA, dB, dD); cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) { double s = 0; for (int k = 0; k < 4; ++k) s += hA[r * 4 + k] * hB[k * 8 + c]; err = fmax(err, fabs(s - hD[r * 8 + c])); }
    printf("layout check: max abs err %.3e (%s)\n", err, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
