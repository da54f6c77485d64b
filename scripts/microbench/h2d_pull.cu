// What the PCIe link gives to (a) the copy engine, (b) SM 16-byte loads (the shipped gather kernel's access pattern),
// (c) TMA bulk copies host -> shared -> device.  Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../dctdomain_b200/csrc/dctd_tma.cuh"
using namespace dctd::tma;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) lsu_pull(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long n16, long long piece16) {
    const long long npieces = (n16 + piece16 - 1) / piece16;
    for (long long e = blockIdx.x; e < npieces; e += gridDim.x) {
        const uint4 *s = src + e * piece16;
        uint4 *d = dst + e * piece16;
        const long long m = min(piece16, n16 - e * piece16);
        constexpr int U = 8;
        for (long long i0 = threadIdx.x; i0 < m; i0 += (long long)blockDim.x * U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + (long long)u * blockDim.x;
                if (i < m)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(s + i));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long i = i0 + (long long)u * blockDim.x;
                if (i < m) d[i] = v[u];
            }
        }
    }
}

// one thread per CTA drives a ring of S stages of CH bytes: bulk load host -> shared (mbarrier), bulk store shared -> device
template <int S>
__global__ void __launch_bounds__(32) tma_pull(const unsigned char *__restrict__ src, unsigned char *__restrict__ dst, long long nbytes, int CH) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long full[S];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long nch = (nbytes + CH - 1) / CH;
    // chunk c of this CTA = blockIdx.x + c * gridDim.x
    long long issued = 0, done = 0;
    const long long mine = (nch - blockIdx.x + gridDim.x - 1) / gridDim.x;
    unsigned int ph[S];
    for (int s = 0; s < S; ++s) ph[s] = 0;
    auto issue = [&](long long c) {
        const int s = (int)(c % S);
        const long long g = (blockIdx.x + c * gridDim.x) * (long long)CH;
        const unsigned int b = (unsigned int)min((long long)CH, nbytes - g);
        mbar_expect_tx(&full[s], b);
        bulk_g2s(smem + (size_t)s * CH, src + g, b, &full[s]);
    };
    for (; issued < mine && issued < S; ++issued) issue(issued);
    for (; done < mine; ++done) {
        const int s = (int)(done % S);
        mbar_wait(&full[s], ph[s]);
        ph[s] ^= 1u;
        const long long g = (blockIdx.x + done * gridDim.x) * (long long)CH;
        const unsigned int b = (unsigned int)min((long long)CH, nbytes - g);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + g), "r"(smem_u32(smem + (size_t)s * CH)), "r"(b) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (issued < mine) {
            // the stage that chunk `issued` reuses was stored S chunks ago: wait until its store has READ shared memory
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(0) : "memory");
            issue(issued);
            ++issued;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    const long long nbytes = 1431LL << 20;
    unsigned char *h, *d;
    CK(cudaHostAlloc(&h, nbytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d, nbytes));
    for (long long i = 0; i < nbytes; i += 4096) h[i] = (unsigned char)(i >> 12);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](const char *name, auto fn) {
        float best = 1e9f;
        for (int r = 0; r < 4; ++r) {
            CK(cudaEventRecord(e0));
            fn();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("%-44s %7.2f ms  %5.1f GB/s\n", name, best, nbytes / best / 1e6);
    };
    timeit("copy engine, one cudaMemcpyAsync", [&] { CK(cudaMemcpyAsync(d, h, nbytes, cudaMemcpyHostToDevice)); });
    for (int g : {4, 8, 16, 32, 64, 148, 148 * 2, 148 * 8}) {
        char nm[96]; snprintf(nm, sizeof nm, "SM 16-byte loads, %d CTAs x 256, 256 KB pieces", g);
        timeit(nm, [&] { lsu_pull<<<g, 256>>>((const uint4 *)h, (uint4 *)d, nbytes / 16, (256 << 10) / 16); });
    }
    for (int CH : {8 << 10, 16 << 10, 32 << 10}) {
        for (int g : {148, 148 * 2, 148 * 4}) {
            const int smem = 4 * CH;
            if (smem * (g / 148) > 200 << 10) continue;
            CK(cudaFuncSetAttribute(tma_pull<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            char nm[96]; snprintf(nm, sizeof nm, "TMA bulk, %d CTAs, 4 stages x %d KB", g, CH >> 10);
            timeit(nm, [&] { tma_pull<4><<<g, 32, smem>>>(h, d, nbytes, CH); });
        }
    }
    // check the last variant's result
    unsigned char *back = (unsigned char *)malloc(nbytes);
    CK(cudaMemset(d, 0, nbytes));
    CK(cudaFuncSetAttribute(tma_pull<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (16 << 10)));
    tma_pull<4><<<296, 32, 4 * (16 << 10)>>>(h, d, nbytes, 16 << 10);
    CK(cudaMemcpy(back, d, nbytes, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (long long i = 0; i < nbytes; ++i) bad += back[i] != h[i];
    printf("TMA pull check: %lld bytes differ\n", bad);
    return 0;
}
