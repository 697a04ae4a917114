// stage_bench.cu -- which shape of the fused 2N stage update (tmp = A tmp + dt du; u += B tmp; du = 0; aux = EOS(u)) reaches
// the HBM rate?  31 concurrent streams (one sweep over all five equations) against per-equation sweeps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I../../include stage_bench.cu -o stage_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "jxpow.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int NEQ = 5;
struct Args { double *u, *tmp, *du, *aux; long long n; double A, B, dt, c0, gam; };

// (A) everything in one sweep, one node per thread
__global__ void __launch_bounds__(256) k_all(Args a) {
    for (long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x; ip < a.n; ip += (long long)gridDim.x * blockDim.x) {
        double d[NEQ], tm[NEQ], q[NEQ];
#pragma unroll
        for (int e = 0; e < NEQ; ++e) d[e] = __ldcs(a.du + e * a.n + ip);
#pragma unroll
        for (int e = 0; e < NEQ; ++e) tm[e] = __ldcs(a.tmp + e * a.n + ip);
#pragma unroll
        for (int e = 0; e < NEQ; ++e) q[e] = a.u[e * a.n + ip];
#pragma unroll
        for (int e = 0; e < NEQ; ++e) {
            tm[e] = a.A * tm[e] + a.dt * d[e];
            q[e] = q[e] + a.B * tm[e];
            __stcs(a.tmp + e * a.n + ip, tm[e]);
            a.u[e * a.n + ip] = q[e];
            a.du[e * a.n + ip] = 0.0;
        }
        a.aux[ip] = a.c0 * jx_pow(q[0] * (q[4] / q[0]), a.gam);
    }
}

// (B) flat sweep over a range of the [eq][node] arrays (6 streams)
__global__ void __launch_bounds__(256) k_flat(Args a, long long lo, long long hi) {
    for (long long t = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < hi; t += (long long)gridDim.x * blockDim.x) {
        const double tm = a.A * __ldcs(a.tmp + t) + a.dt * __ldcs(a.du + t);
        __stcs(a.tmp + t, tm);
        a.u[t] = a.u[t] + a.B * tm;
        a.du[t] = 0.0;
    }
}
// flat sweep, two values per thread (16-byte accesses; lo, hi even)
__global__ void __launch_bounds__(256) k_flat2(Args a, long long lo, long long hi) {
    for (long long t = lo / 2 + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < hi / 2; t += (long long)gridDim.x * blockDim.x) {
        const double2 tp = __ldcs(reinterpret_cast<const double2 *>(a.tmp) + t), d = __ldcs(reinterpret_cast<const double2 *>(a.du) + t);
        double2 u = reinterpret_cast<double2 *>(a.u)[t], tm;
        tm.x = a.A * tp.x + a.dt * d.x; tm.y = a.A * tp.y + a.dt * d.y;
        u.x = u.x + a.B * tm.x; u.y = u.y + a.B * tm.y;
        __stcs(reinterpret_cast<double2 *>(a.tmp) + t, tm);
        reinterpret_cast<double2 *>(a.u)[t] = u;
        reinterpret_cast<double2 *>(a.du)[t] = make_double2(0.0, 0.0);
    }
}
// equations 0 and 4 + aux (13 streams)
__global__ void __launch_bounds__(256) k_two_aux(Args a) {
    for (long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x; ip < a.n; ip += (long long)gridDim.x * blockDim.x) {
        double q[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const long long o = (k == 0 ? 0 : 4) * a.n + ip;
            const double tm = a.A * __ldcs(a.tmp + o) + a.dt * __ldcs(a.du + o);
            __stcs(a.tmp + o, tm);
            q[k] = a.u[o] + a.B * tm;
            a.u[o] = q[k];
            a.du[o] = 0.0;
        }
        a.aux[ip] = a.c0 * jx_pow(q[0] * (q[1] / q[0]), a.gam);
    }
}
// aux only (reads u0, u4)
__global__ void __launch_bounds__(256) k_aux(Args a) {
    for (long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x; ip < a.n; ip += (long long)gridDim.x * blockDim.x) {
        const double r = a.u[ip], rt = a.u[4 * a.n + ip];
        a.aux[ip] = a.c0 * jx_pow(r * (rt / r), a.gam);
    }
}
// old pair: k_lsrk_update (flat, no zero-fill) + k_node_aux (aux + zero-fill)
__global__ void __launch_bounds__(256) k_old_update(Args a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n * NEQ) return;
    const double tm = a.A * a.tmp[t] + a.dt * a.du[t];
    a.tmp[t] = tm;
    a.u[t] = a.u[t] + a.B * tm;
}
__global__ void __launch_bounds__(256) k_old_aux(Args a) {
    for (long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x; ip < a.n; ip += (long long)gridDim.x * blockDim.x) {
        const double r = a.u[ip], rt = a.u[4 * a.n + ip];
        a.aux[ip] = a.c0 * jx_pow(r * (rt / r), a.gam);
#pragma unroll
        for (int e = 0; e < NEQ; ++e) a.du[e * a.n + ip] = 0.0;
    }
}
__global__ void k_init(Args a) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < a.n * NEQ; t += (long long)gridDim.x * blockDim.x) {
        a.u[t] = 1.0 + 1e-3 * (double)(t % 97);
        a.tmp[t] = 0.0;
        a.du[t] = 0.0;
    }
}

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 25153757LL;
    Args a;
    a.n = n; a.A = -0.4; a.B = 0.3; a.dt = 0.0; a.c0 = 3.0; a.gam = 1.4;
    CK(cudaMalloc(&a.u, n * NEQ * 8)); CK(cudaMalloc(&a.tmp, n * NEQ * 8)); CK(cudaMalloc(&a.du, n * NEQ * 8)); CK(cudaMalloc(&a.aux, n * 8));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    k_init<<<sms * 8, 256>>>(a);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, double bytes_per_node, auto fn) {
        for (int i = 0; i < 3; ++i) fn();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int K = 20;
        for (int i = 0; i < K; ++i) fn();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= K;
        printf("%-44s %7.3f ms  %7.1f GB/s (%.0f B/node)\n", name, ms, bytes_per_node * n / ms / 1e6, bytes_per_node);
    };
    const unsigned gflat = (unsigned)((n * NEQ + 255) / 256);
    for (int per = 4; per <= 16; per *= 2) {
        char nm[96];
        snprintf(nm, sizeof nm, "A all-in-one, grid %d/SM", per);
        timeit(nm, 248, [&]() { k_all<<<sms * per, 256>>>(a); });
    }
    timeit("A all-in-one, one node per thread", 248, [&]() { k_all<<<(unsigned)((n + 255) / 256), 256>>>(a); });
    timeit("B flat(all eq)+zero, then aux", 264, [&]() { k_flat<<<sms * 16, 256>>>(a, 0, n * NEQ); k_aux<<<sms * 8, 256>>>(a); });
    timeit("B' flat(all eq)+zero only", 240, [&]() { k_flat<<<sms * 16, 256>>>(a, 0, n * NEQ); });
    timeit("B'' flat2(all eq)+zero only (16 B)", 240, [&]() { k_flat2<<<sms * 16, 256>>>(a, 0, n * NEQ / 2 * 2); });
    timeit("B''' flat(all eq), one value per thread", 240, [&]() { k_flat<<<gflat, 256>>>(a, 0, n * NEQ); });
    timeit("C flat(eq1-3) + two_aux(eq0,4)", 248, [&]() { k_flat<<<sms * 16, 256>>>(a, n, 4 * n); k_two_aux<<<sms * 8, 256>>>(a); });
    timeit("C' two_aux(eq0,4) alone", 2 * 48 + 8, [&]() { k_two_aux<<<sms * 8, 256>>>(a); });
    timeit("D old: lsrk_update + node_aux(zero)", 288, [&]() { k_old_update<<<gflat, 256>>>(a); k_old_aux<<<sms * 8, 256>>>(a); });
    timeit("D' old lsrk_update alone", 200, [&]() { k_old_update<<<gflat, 256>>>(a); });
    timeit("D'' old node_aux alone", 64, [&]() { k_old_aux<<<sms * 8, 256>>>(a); });
    timeit("E aux alone", 24, [&]() { k_aux<<<sms * 8, 256>>>(a); });
    return 0;
}
