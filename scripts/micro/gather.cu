// gather.cu -- B200 micro-benchmark (not product code): how should the element kernel bring the state of the
// 250 nodes of an element pair into shared memory?  Candidates, all double-buffered one group ahead and consumed by a
// node-parallel read (the flux phase's access pattern):
//   0  cp.async 8 B (LDGSTS) per node and component from the SoA state (round-1 kernel)
//   1  cp.async.bulk 48 B per node from a 48-byte-row node image (TMA bulk copy, mbarrier completion)
//   2  cp.async.bulk 64 B per node from a 64-byte-row image
//   3  cp.async.bulk.tensor.2d tile::gather4 from a 64-byte-row image (one op per 4 nodes)
//   4  LDG.128 x 3 per node from the 48-byte-row image straight into registers (no staging; latency exposed)
//   5  LDG.64 x 6 per node from the SoA state straight into registers
// Reported: nodes/s gathered by the whole GPU with nothing else running (the ceiling of the method) -- the element
// kernel needs ~32 G nodes/s -- and, under ncu, l1tex__data_pipe_lsu_wavefronts per group.
// Node ids follow the reference's category numbering (vertices, edge, face, volume interiors; jexpresso_b200/sem/mesh.py).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo gather.cu -o gather
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include <vector>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } \
    } while (0)

constexpr int NP = 125, EPB = 2, NN = EPB * NP, NT = 128, NCOMP = 6;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(b)), "r"(ph)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src),
                 "r"(bytes), "r"(s32(b))
                 : "memory");
}
__device__ __forceinline__ void gather4(void *dst, const CUtensorMap *tm, int c0, int r0, int r1, int r2, int r3, uint64_t *b) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            s32(dst)),
        "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(s32(b))
        : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s32(dst)), "l"(src) : "memory");
}

struct Args {
    const double *soa;     // [NCOMP][npoin]
    const double *aos48;   // [npoin][6]
    const double *aos64;   // [npoin][8]
    const int32_t *ids;    // [ngroups][256] (250 used)
    double *out;
    int64_t npoin;
    int ngroups;
};

template <int V>
__global__ void __launch_bounds__(NT) k_gather(const __grid_constant__ Args a, const __grid_constant__ CUtensorMap tm) {
    constexpr int ROWB = (V == 1 || V == 4) ? 48 : 64;
    constexpr int TILE_B = V == 0 ? NCOMP * 256 * 8 : (V == 3 ? 64 * 256 : ROWB * 256);
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *tile[2] = {sm, sm + TILE_B};
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 2 * TILE_B);
    const int t = threadIdx.x;
    if (t == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    double acc = 0.0;
    uint32_t ph[2] = {0, 0};
    auto issue = [&](int g, int b) {
        const int32_t *id = a.ids + (size_t)g * 256;
        if constexpr (V == 0) {
            for (int n = t; n < NN; n += NT) {
                const int ip = id[n];
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) cp_async8(tile[b] + ((size_t)c * 256 + n) * 8, a.soa + (size_t)c * a.npoin + ip);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else if constexpr (V == 1 || V == 2) {
            if (t == 0) mbar_expect(&bar[b], NN * ROWB);
            __syncwarp();
            // the expect_tx must precede the completions: thread 0 is in warp 0; other warps are ordered by the block barrier
            // of the consume step (every issue happens after a __syncthreads that follows thread 0's previous wait)
            const unsigned char *src = reinterpret_cast<const unsigned char *>(V == 1 ? a.aos48 : a.aos64);
            for (int n = t; n < NN; n += NT) bulk_g2s(tile[b] + (size_t)n * ROWB, src + (size_t)id[n] * ROWB, ROWB, &bar[b]);
        } else if constexpr (V == 3) {
            if (t == 0) mbar_expect(&bar[b], 63 * 4 * 64);
            __syncwarp();
            for (int o = t; o < 63; o += NT) {
                const int4 r = *reinterpret_cast<const int4 *>(id + 4 * o);
                gather4(tile[b] + (size_t)o * 256, &tm, 0, r.x, r.y, r.z, r.w, &bar[b]);
            }
        }
    };
    auto consume = [&](int g, int b) {
        if constexpr (V == 0) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            for (int n = t; n < NN; n += NT)
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) acc += reinterpret_cast<double *>(tile[b])[c * 256 + n];
        } else if constexpr (V >= 1 && V <= 3) {
            mbar_wait(&bar[b], ph[b]);
            ph[b] ^= 1;
            for (int n = t; n < NN; n += NT) {
                const double2 *r = reinterpret_cast<const double2 *>(tile[b] + (size_t)n * ROWB);
                const double2 x = r[0], y = r[1], z = r[2];
                acc += x.x + x.y + y.x + y.y + z.x + z.y;
            }
        } else if constexpr (V == 4) {
            const int32_t *id = a.ids + (size_t)g * 256;
            for (int n = t; n < NN; n += NT) {
                const double2 *r = reinterpret_cast<const double2 *>(a.aos48 + (size_t)id[n] * 6);
                const double2 x = __ldg(r), y = __ldg(r + 1), z = __ldg(r + 2);
                acc += x.x + x.y + y.x + y.y + z.x + z.y;
            }
        } else {
            const int32_t *id = a.ids + (size_t)g * 256;
            for (int n = t; n < NN; n += NT) {
                const int ip = id[n];
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) acc += __ldg(a.soa + (size_t)c * a.npoin + ip);
            }
        }
    };
    int g = blockIdx.x, b = 0;
    if (g < a.ngroups) issue(g, 0);
    for (; g < a.ngroups; g += gridDim.x, b ^= 1) {
        const int gn = g + gridDim.x;
        if (gn < a.ngroups) issue(gn, b ^ 1);
        else if (V == 0) asm volatile("cp.async.commit_group;" ::: "memory");
        if constexpr (V == 0) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncthreads();
            for (int n = t; n < NN; n += NT)
#pragma unroll
                for (int c = 0; c < NCOMP; ++c) acc += reinterpret_cast<double *>(tile[b])[c * 256 + n];
        } else consume(g, b);
        __syncthreads();   // tile b free again
    }
    a.out[(size_t)blockIdx.x * NT + t] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int nel = argc > 1 ? atoi(argv[1]) : 48;
    const int boxrows = argc > 2 ? atoi(argv[2]) : 1;
    const int p = 4, nv = nel + 1, q = p - 1;
    const int64_t nE = (int64_t)nel * nv * nv, nF = (int64_t)nel * nel * nv;
    const int64_t off_ex = (int64_t)nv * nv * nv, off_ey = off_ex + nE * q, off_ez = off_ey + nE * q, off_fxy = off_ez + nE * q,
                  off_fxz = off_fxy + nF * q * q, off_fyz = off_fxz + nF * q * q, off_vol = off_fyz + nF * q * q;
    const int64_t npoin = off_vol + (int64_t)nel * nel * nel * q * q * q;
    auto gid = [&](int I, int J, int K) -> int64_t {
        const int ex = I / p, rx = I % p, ey = J / p, ry = J % p, ez = K / p, rz = K % p;
        const bool a = rx == 0, b = ry == 0, c = rz == 0;
        if (a && b && c) return ex + (int64_t)nv * (ey + (int64_t)nv * ez);
        if (!a && b && c) return off_ex + (ex + (int64_t)nel * (ey + (int64_t)nv * ez)) * q + (rx - 1);
        if (a && !b && c) return off_ey + (ex + (int64_t)nv * (ey + (int64_t)nel * ez)) * q + (ry - 1);
        if (a && b && !c) return off_ez + (ex + (int64_t)nv * (ey + (int64_t)nv * ez)) * q + (rz - 1);
        if (!a && !b && c) return off_fxy + (ex + (int64_t)nel * (ey + (int64_t)nel * ez)) * q * q + (rx - 1) + q * (ry - 1);
        if (!a && b && !c) return off_fxz + (ex + (int64_t)nel * (ey + (int64_t)nv * ez)) * q * q + (rx - 1) + q * (rz - 1);
        if (a && !b && !c) return off_fyz + (ex + (int64_t)nv * (ey + (int64_t)nel * ez)) * q * q + (ry - 1) + q * (rz - 1);
        return off_vol + (ex + (int64_t)nel * (ey + (int64_t)nel * ez)) * q * q * q + (rx - 1) + q * ((ry - 1) + q * (rz - 1));
    };
    const int64_t nelem = (int64_t)nel * nel * nel;
    const int ngroups = (int)((nelem + 1) / 2);
    std::vector<int32_t> ids((size_t)ngroups * 256, 0);
    for (int64_t e = 0; e < nelem; ++e) {
        const int ex = (int)(e % nel), ey = (int)((e / nel) % nel), ez = (int)(e / ((int64_t)nel * nel));
        for (int k = 0; k < 5; ++k)
            for (int j = 0; j < 5; ++j)
                for (int i = 0; i < 5; ++i)   // local i along -x, j along +z, k along +y (mesh.py)
                    ids[(size_t)(e / 2) * 256 + (e % 2) * NP + i + 5 * j + 25 * k] = (int32_t)gid(p * ex + (p - i), p * ey + k, p * ez + j);
    }
    printf("nel %d, npoin %lld, groups %d\n", nel, (long long)npoin, ngroups);
    std::vector<double> soa((size_t)npoin * NCOMP), a48((size_t)npoin * 6), a64((size_t)npoin * 8, 0.0);
    for (int64_t ip = 0; ip < npoin; ++ip)
        for (int c = 0; c < NCOMP; ++c) {
            const double v = (double)((ip * 7 + c * 13) % 1000) * 1e-3;
            soa[(size_t)c * npoin + ip] = v; a48[(size_t)ip * 6 + c] = v; a64[(size_t)ip * 8 + c] = v;
        }
    double want = 0.0;
    for (int g = 0; g < ngroups; ++g)
        for (int n = 0; n < NN; ++n)
            for (int c = 0; c < NCOMP; ++c) want += soa[(size_t)c * npoin + ids[(size_t)g * 256 + n]];
    Args a;
    double *d_soa, *d_48, *d_64, *d_out;
    int32_t *d_ids;
    const int grid = 148 * 4;
    CK(cudaMalloc(&d_soa, soa.size() * 8)); CK(cudaMalloc(&d_48, a48.size() * 8)); CK(cudaMalloc(&d_64, a64.size() * 8));
    CK(cudaMalloc(&d_ids, ids.size() * 4)); CK(cudaMalloc(&d_out, (size_t)grid * NT * 8));
    CK(cudaMemcpy(d_soa, soa.data(), soa.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_48, a48.data(), a48.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_64, a64.data(), a64.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice));
    a.soa = d_soa; a.aos48 = d_48; a.aos64 = d_64; a.ids = d_ids; a.out = d_out; a.npoin = npoin; a.ngroups = ngroups;

    CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        cuuint64_t dims[2] = {8, (cuuint64_t)npoin}, strides[1] = {64};
        cuuint32_t box[2] = {8, (cuuint32_t)boxrows}, es[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d_64, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("cuTensorMapEncodeTiled(box rows %d) -> %d\n", boxrows, (int)r);
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<double> out((size_t)grid * NT);
    auto run = [&](int v, auto kern, size_t smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            kern<<<grid, NT, smem>>>(a, tm);
            CK(cudaEventRecord(e1));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); exit(2); }
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        CK(cudaMemcpy(out.data(), d_out, out.size() * 8, cudaMemcpyDeviceToHost));
        double got = 0.0;
        for (double x : out) got += x;
        printf("variant %d: %.3f ms, %.1f G nodes/s, checksum %s (%.6e vs %.6e)\n", v, best, (double)ngroups * NN / best * 1e-6,
               fabs(got - want) <= 1e-9 * fabs(want) ? "ok" : "MISMATCH", got, want);
    };
    const int only = argc > 3 ? atoi(argv[3]) : -1;
    if (only < 0 || only == 0) run(0, k_gather<0>, 2 * NCOMP * 256 * 8 + 64);
    if (only < 0 || only == 1) run(1, k_gather<1>, 2 * 48 * 256 + 64);
    if (only < 0 || only == 2) run(2, k_gather<2>, 2 * 64 * 256 + 64);
    if (only < 0 || only == 4) run(4, k_gather<4>, 2 * 48 * 256 + 64);
    if (only < 0 || only == 5) run(5, k_gather<5>, 2 * 64 * 256 + 64);
    if (only < 0 || only == 3) run(3, k_gather<3>, 2 * 64 * 256 + 64);
    return 0;
}
