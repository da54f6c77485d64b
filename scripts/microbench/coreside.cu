// While a PCIe-bound pull kernel (128 threads x 148 CTAs, no shared memory) runs on one stream: how long does
// (a) a small copy-engine H2D transfer, (b) a one-CTA-per-SM kernel with 223 KB of shared memory and 608 threads x 96
// registers wait on another (high-priority) stream?  Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1); } } while (0)

template <int U, int MINB>
__global__ void __launch_bounds__(128, MINB) pull(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long n16) {
    for (long long i0 = (long long)blockIdx.x * blockDim.x * U + threadIdx.x; i0 < n16; i0 += (long long)gridDim.x * blockDim.x * U) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + (long long)u * blockDim.x;
            if (i < n16) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + i));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + (long long)u * blockDim.x;
            if (i < n16) dst[i] = v[u];
        }
    }
}

// stands in for fp_ws_kernel: one CTA per SM, 608 threads, ~96 registers (forced by the array), big dynamic shared memory
__global__ void __launch_bounds__(608, 1) big(float *out, int iters) {
    extern __shared__ float sm[];
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = acc[i] * 1.0001f + sm[(threadIdx.x + i * 37 + it) & 1023];
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = sm[5];
}

int main() {
    const long long nbytes = 1431LL << 20;
    unsigned char *h, *d, *hs, *ds;
    float *o;
    CK(cudaHostAlloc(&h, nbytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&hs, 1 << 20, cudaHostAllocDefault));
    CK(cudaMalloc(&d, nbytes));
    CK(cudaMalloc(&ds, 1 << 20));
    CK(cudaMalloc(&o, 4096));
    int lo, hi;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t sx, sy;
    CK(cudaStreamCreateWithPriority(&sx, cudaStreamNonBlocking, lo));
    CK(cudaStreamCreateWithPriority(&sy, cudaStreamNonBlocking, hi));
    cudaEvent_t e[6];
    for (auto &x : e) CK(cudaEventCreate(&x));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, big));
    printf("big: %d registers, pull: ", fa.numRegs);
    CK(cudaFuncGetAttributes(&fa, pull<4, 16>));
    printf("%d registers\n", fa.numRegs);
    for (int smem_kb : {223, 100}) {
        for (int carve = 0; carve < 2; ++carve) {
            for (int split = 0; split < 2; ++split) {
                CK(cudaFuncSetAttribute(big, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb << 10));
                CK(cudaFuncSetAttribute(pull<4, 16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        carve ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault));
                // warm up
                big<<<148, 608, smem_kb << 10, sy>>>(o, 10);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e[0], sx));
                const int nl = split ? 16 : 1;         // the pull as one launch or as 16 back-to-back launches
                for (int l = 0; l < nl; ++l)
                    pull<4, 16><<<148, 128, 0, sx>>>((const uint4 *)(h + nbytes / nl / 16 * 16 * l), (uint4 *)(d + nbytes / nl / 16 * 16 * l), nbytes / nl / 16);
                CK(cudaEventRecord(e[1], sx));
                CK(cudaEventRecord(e[2], sy));
                CK(cudaMemcpyAsync(ds, hs, 1 << 20, cudaMemcpyHostToDevice, sy));
                CK(cudaEventRecord(e[3], sy));
                big<<<148, 608, smem_kb << 10, sy>>>(o, 2000);
                CK(cudaEventRecord(e[4], sy));
                CK(cudaDeviceSynchronize());
                CK(cudaGetLastError());
                float t_pull, t_dma, t_big, t_end;
                CK(cudaEventElapsedTime(&t_pull, e[0], e[1]));
                CK(cudaEventElapsedTime(&t_dma, e[2], e[3]));
                CK(cudaEventElapsedTime(&t_big, e[3], e[4]));
                CK(cudaEventElapsedTime(&t_end, e[0], e[4]));
                printf("big smem %3d KB, pull carveout %s, pull in %2d launch(es): pull %6.2f ms | 1 MB DMA beside it %6.3f ms | big kernel after it %6.3f ms (done %6.2f ms after the pull began)\n",
                       smem_kb, carve ? "max-shared" : "default   ", nl, t_pull, t_dma, t_big, t_end);
            }
        }
    }
    // the big kernel alone
    CK(cudaEventRecord(e[0], sy));
    big<<<148, 608, 223 << 10, sy>>>(o, 2000);
    CK(cudaEventRecord(e[1], sy));
    CK(cudaDeviceSynchronize());
    float t;
    CK(cudaEventElapsedTime(&t, e[0], e[1]));
    printf("big kernel alone: %.3f ms\n", t);
    return 0;
}
