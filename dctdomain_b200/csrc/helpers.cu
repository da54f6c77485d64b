// libdctd: GPU versions of the small public helpers of reference src/fingerprint.py that are part of its
// API surface but not of the batched hot path: Fingerprint.scale (:110-123) and Fingerprint.idct_quant
// (:126-142) on arbitrary float64 matrices.  Plain float64 kernels, not tuned: the hot path is
// dctd_fp_execute (fingerprint.cu), which fuses both idct_quant calls of quantize().
#include <math.h>

#include "dctd_internal.cuh"

namespace {

// out[i] = (x[i] - min) / (max - min) over the whole vector; one CTA
__global__ void scale_kernel(const double *__restrict__ x, long long n, double *__restrict__ out) {
    __shared__ double s_mn[32], s_mx[32];
    double mn = INFINITY, mx = -INFINITY;
    bool nan = false;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = x[i];
        nan = nan || !(v == v);
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    if (nan) { mn = NAN; mx = NAN; }      // numpy min/max propagate NaN
    for (int o = 16; o > 0; o >>= 1) {
        const double a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = (a != a || mn != mn) ? NAN : fmin(mn, a);
        mx = (b != b || mx != mx) ? NAN : fmax(mx, b);
    }
    if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = s_mn[0]; mx = s_mx[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        mn = (s_mn[w] != s_mn[w] || mn != mn) ? NAN : fmin(mn, s_mn[w]);
        mx = (s_mx[w] != s_mx[w] || mx != mx) ? NAN : fmax(mx, s_mx[w]);
    }
    for (long long i = threadIdx.x; i < n; i += blockDim.x) out[i] = (x[i] - mn) / (mx - mn);
}

// One CTA per column c of x [R, C]: DCT-II (ortho) along the rows, first `num` coefficients, length-`num`
// inverse, min-max over the num values.  out [num, C].
__global__ void idct_quant_kernel(const double *__restrict__ x, int R, int C, int num, double *__restrict__ out) {
    extern __shared__ double sh[];          // u[num] | y[num]
    double *u = sh, *y = sh + num;
    const int c = blockIdx.x;
    for (int k = threadIdx.x; k < num; k += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < R; ++r) {
            const long long q = ((long long)(2 * r + 1) * k) % (4LL * R);
            s = fma(x[(long long)r * C + c], cospi((double)q / (2.0 * R)), s);
        }
        u[k] = s * (k == 0 ? sqrt(1.0 / R) : sqrt(2.0 / R));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < num; j += blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < num; ++k) {
            const long long q = ((long long)(2 * j + 1) * k) % (4LL * num);
            s = fma(u[k] * (k == 0 ? sqrt(1.0 / num) : sqrt(2.0 / num)), cospi((double)q / (2.0 * num)), s);
        }
        y[j] = s;
    }
    __syncthreads();
    double mn = INFINITY, mx = -INFINITY;
    bool nan = false;
    for (int j = 0; j < num; ++j) {
        nan = nan || !(y[j] == y[j]);
        mn = fmin(mn, y[j]);
        mx = fmax(mx, y[j]);
    }
    if (nan) { mn = NAN; mx = NAN; }
    for (int j = threadIdx.x; j < num; j += blockDim.x) out[(long long)j * C + c] = (y[j] - mn) / (mx - mn);
}

}  // namespace

extern "C" {

int dctd_scale_f64(const double *d_x, int64_t n, double *d_out, void *stream) {
    if (n < 0 || (n > 0 && (!d_x || !d_out))) return DCTD_ERR_ARG;
    if (n == 0) return DCTD_OK;
    scale_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_x, n, d_out);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

int dctd_idct_quant_f64(const double *d_x, int32_t rows, int32_t cols, int32_t num, double *d_out, void *stream) {
    if (rows < 1 || cols < 1 || num < 1 || !d_x || !d_out) return DCTD_ERR_ARG;
    if (num > rows) num = rows;             // scipy: f[:, :num] has at most `rows` coefficients
    if (num > 4096) return DCTD_ERR_UNSUPPORTED;
    const int threads = num <= 32 ? 32 : (num <= 64 ? 64 : 128);
    idct_quant_kernel<<<cols, threads, 2 * (size_t)num * sizeof(double), (cudaStream_t)stream>>>(d_x, rows, cols, num, d_out);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

}  // extern "C"
