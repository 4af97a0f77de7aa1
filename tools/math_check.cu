// Accuracy of the branch-free FP64 helpers in bsyn_math.cuh against CUDA's IEEE / libm versions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/math_check tools/math_check.cu && /tmp/math_check
#include <cstdio>
#include <cmath>
#include "../bayesynergy_b200/csrc/bsyn_math.cuh"

__device__ double u01(unsigned long long &s)
{
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return ((s >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__global__ void check(double *out)
{
    unsigned long long s = 1234567ULL + 7919ULL * (blockIdx.x * blockDim.x + threadIdx.x);
    double e_rcp = 0, e_rsq = 0, e_exp = 0, e_log = 0, e_div = 0;
    for (int i = 0; i < 20000; ++i) {
        const double m = 1.0 + u01(s), sc = exp(40.0 * (u01(s) - 0.5));
        const double x = m * sc;
        e_rcp = fmax(e_rcp, fabs(bsyn::fast_rcp(x) * x - 1.0));
        const double r = bsyn::fast_rsqrt(x);
        e_rsq = fmax(e_rsq, fabs(r * r * x - 1.0) * 0.5);
        const double a = 60.0 * (u01(s) - 0.5);
        const double ex = exp(a);
        e_exp = fmax(e_exp, fabs(bsyn::fast_exp(a) - ex) / ex);
        const double lg = log(x);
        if (fabs(lg) > 1e-3) e_log = fmax(e_log, fabs(bsyn::fast_log(x) - lg) / fabs(lg));
        const double y = 1.0 + u01(s);
        e_div = fmax(e_div, fabs(bsyn::fast_div(y, x) - y / x) / (y / x));
    }
    atomicMax((unsigned long long *)&out[0], __double_as_longlong(e_rcp));
    atomicMax((unsigned long long *)&out[1], __double_as_longlong(e_rsq));
    atomicMax((unsigned long long *)&out[2], __double_as_longlong(e_exp));
    atomicMax((unsigned long long *)&out[3], __double_as_longlong(e_log));
    atomicMax((unsigned long long *)&out[4], __double_as_longlong(e_div));
}
int main()
{
    double *d, h[5];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0, sizeof(h));
    check<<<64, 128>>>(d);
    if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
    printf("max relative error over 1.6e8 samples: rcp %.3g  rsqrt %.3g  exp %.3g  log %.3g  div %.3g  (2^-53 = 1.11e-16)\n",
           h[0], h[1], h[2], h[3], h[4]);
    return 0;
}
