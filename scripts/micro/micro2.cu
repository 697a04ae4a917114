// micro2.cu -- B200 micro-benchmarks behind the element-kernel occupancy analysis (not product code):
//   per-warp DFMA issue rate and dependent-issue latency as a function of independent chains and warps per
//   scheduler; LDS.64 latency; DFMA with constant-bank operand.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false micro2.cu -o micro2
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cD[32];

template <int NCH, bool CONSTOP>
__global__ void k_chain(double *out, int iters, long long *cyc, double s) {
    double a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b0 = s, b1 = s * 1.1;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < NCH; ++i) a[i] = CONSTOP ? fma(a[i], cD[(u * NCH + i) & 31], b1) : fma(a[i], b0, b1);
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
    double r = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void k_lds_lat(int *out, int iters, long long *cyc) {
    __shared__ int sm[1024];
    for (int x = threadIdx.x; x < 1024; x += blockDim.x) sm[x] = (x * 37 + 11) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) p = sm[p];
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = p;
}

int main() {
    double *out; long long *cyc; int *iout;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 32 * 8); cudaMalloc(&iout, 148 * 1024 * 4);
    double h[32];
    for (int i = 0; i < 32; ++i) h[i] = 1.0 + i * 1e-9;
    cudaMemcpyToSymbol(cD, h, sizeof h);
    const int iters = 4000;
    long long hc[32];
#define RUN(NCH, C, NT)                                                                                      \
    {                                                                                                        \
        k_chain<NCH, C><<<148, NT>>>(out, iters, cyc, 1.0000001);                                            \
        cudaDeviceSynchronize();                                                                             \
        cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);                                              \
        printf("dfma chains %2d constop %d warps/SM %2d: %.2f cyc per DFMA per warp (=> %.2f cyc latency if 1 chain)\n", NCH, (int)C, NT / 32, \
               (double)hc[0] / (iters * 8.0 * NCH), (double)hc[0] / (iters * 8.0));                          \
    }
    RUN(1, false, 32) RUN(2, false, 32) RUN(4, false, 32) RUN(8, false, 32) RUN(12, false, 32) RUN(16, false, 32) RUN(24, false, 32)
    RUN(8, true, 32) RUN(16, true, 32)
    RUN(4, false, 128) RUN(8, false, 128) RUN(16, false, 128)
    RUN(4, false, 256) RUN(8, false, 256) RUN(16, false, 256) RUN(16, true, 256)
    RUN(4, false, 512) RUN(8, false, 512)
    k_lds_lat<<<1, 32>>>(iout, 10000, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(hc, cyc, 8, cudaMemcpyDeviceToHost);
    printf("LDS.32 dependent latency: %.1f cyc\n", (double)hc[0] / 10000);
    return 0;
}
