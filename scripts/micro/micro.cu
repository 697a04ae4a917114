// micro.cu -- B200 micro-benchmarks that decide the element-kernel layout (not product code):
//   shared-memory wavefronts for broadcast LDS.64/LDS.128 patterns, DFMA issue rate with register /
//   constant operands, RED.F64 throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 micro.cu -o micro
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ double cD[32];

template <int MODE>
__global__ void k_lds(double *out, int iters, long long *cyc) {
    extern __shared__ __align__(16) double sm[];
    const int t = threadIdx.x, lane = t & 31;
    for (int x = t; x < 4096; x += blockDim.x) sm[x] = x * 0.5;
    __syncthreads();
    int idx;
    // MODE 0: LDS.64 all lanes distinct consecutive; 1: LDS.64, 7 distinct rows stride 5 (lane/5);
    // 2: LDS.64 7 distinct rows stride 6; 3: LDS.128 distinct consecutive; 4: LDS.128 7 rows stride 6 (48 B)
    // 5: LDS.64 25 distinct consecutive + 7 more (i.e. lane) stride 1 == mode 0; 6: LDS.128 all same address
    // 7: LDS.64 5 distinct, stride 6 (lane%5)  8: LDS.128 lane%5 stride 6
    if (MODE == 0 || MODE == 3) idx = lane;
    else if (MODE == 1) idx = (lane / 5) * 5;
    else if (MODE == 2 || MODE == 4) idx = (lane / 5) * 6;
    else if (MODE == 6) idx = 0;
    else idx = (lane % 5) * 6;
    long long acc = 0, acc2 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (MODE == 3 || MODE == 4 || MODE == 6 || MODE == 8) {
                const longlong2 v = *reinterpret_cast<const longlong2 *>(&sm[(idx * (MODE == 3 ? 2 : 1) + u * 64) & 4094]);
                acc ^= v.x; acc2 ^= v.y;
            } else {
                acc ^= reinterpret_cast<const long long *>(sm)[(idx + u * 64) & 4095];
            }
        }
        idx += (int)(acc & 0x4000000000000000LL ? 1 : 0) * 0;  if (acc == 0x1234567) idx++;
    }
    long long t1 = clock64();
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + t] = (double)(acc ^ acc2);
}

template <int MODE>
__global__ void k_dfma(double *out, int iters, long long *cyc, double s) {
    const int t = threadIdx.x;
    double a[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = t * 1e-3 + i;
    double b0 = s, b1 = s * 1.1, b2 = s * 0.9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                if (MODE == 0) a[i] = fma(a[i], b0, b1);
                else a[i] = fma(a[i], cD[(u * 12 + i) & 31], b2);
            }
        }
    }
    long long t1 = clock64();
    if (t == 0) cyc[blockIdx.x] = t1 - t0;
    double r = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + t] = r;
}

// RED.F64: element-like pattern: lane -> node id = base + (lane%5) + 1000*(lane/5), 5 arrays apart
__global__ void k_red(double *du, long long npoin, int iters, int mode) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        long long w = (g / 32 + (long long)it * gridDim.x * (blockDim.x / 32));
        long long ip;
        if (mode == 0) ip = (w * 32 + lane) % npoin;                        // fully coalesced
        else ip = (w * 5 + (lane % 5) + 293LL * (lane / 5)) % npoin;       // 5-node runs, 293 apart
        atomicAdd(&du[ip], 1.0);
    }
}
__global__ void k_st(double *du, long long npoin, int iters, int mode) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        long long w = (g / 32 + (long long)it * gridDim.x * (blockDim.x / 32));
        long long ip;
        if (mode == 0) ip = (w * 32 + lane) % npoin;
        else ip = (w * 5 + (lane % 5) + 293LL * (lane / 5)) % npoin;
        du[ip] = 1.0;
    }
}

template <class F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 8 * 1024 * 8);
    cudaMalloc(&cyc, 148 * 8 * 8);
    double h[32];
    for (int i = 0; i < 32; ++i) h[i] = 1.0 + i * 1e-9;
    cudaMemcpyToSymbol(cD, h, sizeof h);
    const int iters = 2000;
    long long hc[148 * 8];
#define RUN_LDS(M, NT)                                                                        \
    {                                                                                         \
        float ms = timeit([&] { k_lds<M><<<148, NT, 4096 * 8>>>(out, iters, cyc); });         \
        cudaMemcpy(hc, cyc, 148 * 8, cudaMemcpyDeviceToHost);                                 \
        printf("lds mode %d nt %d: %.3f ms, %.2f cyc per warp-instr per SM (1 CTA/SM)\n", M, NT, ms, \
               (double)hc[0] / (iters * 16.0 * (NT / 32)));                                   \
    }
    RUN_LDS(0, 256) RUN_LDS(1, 256) RUN_LDS(2, 256) RUN_LDS(3, 256) RUN_LDS(4, 256) RUN_LDS(6, 256) RUN_LDS(7, 256) RUN_LDS(8, 256)
    RUN_LDS(0, 512) RUN_LDS(4, 512)
#define RUN_DFMA(M, NT)                                                                       \
    {                                                                                         \
        float ms = timeit([&] { k_dfma<M><<<148, NT>>>(out, iters, cyc, 1.0000001); });      \
        cudaMemcpy(hc, cyc, 148 * 8, cudaMemcpyDeviceToHost);                                 \
        double fl = 148.0 * NT * iters * 96.0 * 2;                                            \
        printf("dfma mode %d nt %d: %.3f ms, %.2f TFLOP/s, %.3f cyc per warp-DFMA per SM\n", M, NT, ms, fl / ms / 1e9, \
               (double)hc[0] / (iters * 96.0 * (NT / 32)));                                   \
    }
    RUN_DFMA(0, 128) RUN_DFMA(0, 256) RUN_DFMA(0, 512) RUN_DFMA(1, 128) RUN_DFMA(1, 256) RUN_DFMA(1, 512)
    const long long npoin = 25153757;
    double *du;
    cudaMalloc(&du, npoin * 8);
    cudaMemset(du, 0, npoin * 8);
    for (int mode = 0; mode < 2; ++mode) {
        const int it2 = 64, grid = 148 * 16, nt = 256;
        float ms = timeit([&] { k_red<<<grid, nt>>>(du, npoin, it2, mode); });
        double n = (double)grid * nt * it2;
        printf("red.f64 mode %d: %.3f ms, %.1f G red/s (%.1f GB/s of 8-byte updates)\n", mode, ms, n / ms / 1e6, n * 8 / ms / 1e6);
        ms = timeit([&] { k_st<<<grid, nt>>>(du, npoin, it2, mode); });
        printf("st.f64  mode %d: %.3f ms, %.1f G st/s\n", mode, ms, n / ms / 1e6);
    }
    return 0;
}
