// tools/gather_lab.cu -- development tool: how fast can the 19 populations of the cells of a z-face (one 8-byte value per
// 4 KB row and slot) be gathered and scattered?  (a) plain loads, (b) cp.async 8 B, (c) cp.async.bulk 16 B + mbarrier.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_lab.cu -o tools/bin/gather_lab && tools/bin/gather_lab
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int Q = 19;

__device__ __forceinline__ double work(const double* f) {
    double s = 0.;
#pragma unroll
    for (int i = 0; i < Q; ++i) s += f[i] * (1.0 + 1e-3 * i);
    return s;
}

template <int MODE, bool STORE>
__global__ void __launch_bounds__(128) gather(double* F, long long S, int P, int nz, int nCells, int ny, int zsel) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    const bool act = t < nCells;
    const int X = 1 + t / ny, y = t % ny;
    const long long c = (long long)X * P + (long long)y * nz + zsel;
    double f[Q];
    if (MODE == 0) {
        if (act) {
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = F[i * S + c];
        }
    } else if (MODE == 1) {
        double* sh = reinterpret_cast<double*>(smem) + threadIdx.x;     // [Q][128]
        if (act) {
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(sh + i * 128);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(F + i * S + c));
            }
        }
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;");
        if (act) {
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = sh[i * 128];
        }
    } else if (MODE == 2) {
        double2* sh = reinterpret_cast<double2*>(smem) + threadIdx.x;   // [Q][128] pairs
        const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;");
        }
        __syncthreads();
        const int nAct = min(128, nCells - (int)(blockIdx.x * blockDim.x));
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(nAct * Q * 16));
        if (act) {
            const long long ca = c & ~1ll;
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(sh + i * 128);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];"
                             ::"r"(dst), "l"(F + i * S + ca), "r"(mb) : "memory");
            }
        }
        unsigned ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb), "r"(0) : "memory");
        }
        if (act) {
#pragma unroll
            for (int i = 0; i < Q; ++i) { const double2 v = sh[i * 128]; f[i] = (c & 1) ? v.y : v.x; }
        }
    }
    if (MODE == 3) {
        // 256-bit accesses (sm_100): the whole 32-byte sector of each population, so that the store is a full-sector write
        if (act) {
            const long long ca = c & ~3ll;
            double s = 0.;
            double v[Q][4];
#pragma unroll
            for (int i = 0; i < Q; ++i)
                asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[i][0]), "=d"(v[i][1]), "=d"(v[i][2]), "=d"(v[i][3]) : "l"(F + i * S + ca));
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = v[i][c & 3];
            s = work(f);
            if (STORE) {
#pragma unroll
                for (int i = 0; i < Q; ++i) {
                    v[i][c & 3] = f[i] + 1e-9 * s;
                    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(F + i * S + ca), "d"(v[i][0]), "d"(v[i][1]), "d"(v[i][2]), "d"(v[i][3]) : "memory");
                }
            } else if (s == 12345.678) F[c] = s;
        }
        return;
    }
    if (act) {
        const double s = work(f);
        if (STORE) {
#pragma unroll
            for (int i = 0; i < Q; ++i) F[i * S + c] = f[i] + 1e-9 * s;
        } else if (s == 12345.678) F[c] = s;
    }
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 512;
    const int P = n * n, nz = n, ny = n;
    const long long S = (long long)(n + 2) * P;
    double* F;
    CK(cudaMalloc(&F, Q * S * sizeof(double)));
    CK(cudaMemset(F, 0, Q * S * sizeof(double)));
    const int nCells = n * n;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double* flush; const size_t flushN = 512ull << 20; CK(cudaMalloc(&flush, flushN));
    auto run = [&](const char* name, auto kern, size_t smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float best = 1e9f, sum = 0.f;
        for (int it = 0; it < 6; ++it) {
            CK(cudaMemset(flush, it, flushN));
            CK(cudaEventRecord(e0));
            kern<<<(nCells + 127) / 128, 128, smem>>>(F, S, P, nz, nCells, ny, it & 1 ? nz - 1 : 0);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%-28s one z-face of %d cells: best %.3f ms  mean %.3f ms\n", name, nCells, best, sum / 5);
    };
    run("ldg          gather only", gather<0, false>, 0);
    run("ldg          gather+scatter", gather<0, true>, 0);
    run("cp.async 8B  gather only", gather<1, false>, Q * 128 * 8);
    run("cp.async 8B  gather+scatter", gather<1, true>, Q * 128 * 8);
    run("256-bit      gather only", gather<3, false>, 0);
    run("256-bit      gather+scatter", gather<3, true>, 0);
    run("bulk 16B     gather only", gather<2, false>, Q * 128 * 16);
    run("bulk 16B     gather+scatter", gather<2, true>, Q * 128 * 16);
    return 0;
}
