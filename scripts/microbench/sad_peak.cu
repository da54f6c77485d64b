// Microbenchmark: peak issue rate of VABSDIFF4.U8.ACC (the instruction the L1 scan kernel is made of).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sad_peak sad_peak.cu ; run on the B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned sad4(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

template <int ACC>
__global__ void k(unsigned *out, int iters, unsigned seed) {
    unsigned acc[ACC], a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ACC; ++i) acc[i] = sad4(a + i, b + r, acc[i]);
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < ACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
void run(int warps_per_sm, int sms, float mhz) {
    unsigned *out;
    const int threads = 256, blocks = sms * warps_per_sm * 32 / threads;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    const int iters = 4096;
    k<ACC><<<blocks, threads>>>(out, 16, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<ACC><<<blocks, threads>>>(out, iters, 7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8.0 * ACC;
    printf("{\"acc\": %d, \"warps_per_sm\": %d, \"sad4_per_s\": %.4g, \"lanes_per_clk_per_sm_at_%.0fMHz\": %.2f}\n",
           ACC, warps_per_sm, ops / (ms * 1e-3), mhz, ops / (ms * 1e-3) / sms / (mhz * 1e6));
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float mhz = khz / 1000.f;
    for (int w : {8, 16, 32, 64}) { run<8>(w, p.multiProcessorCount, mhz); run<16>(w, p.multiProcessorCount, mhz); }
    return 0;
}
