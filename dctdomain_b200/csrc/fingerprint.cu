// libdctd hot path 1: batched DCT fingerprints ("quant2D") for sm_100a.
//
// What the reference computes per (domain, layer) - src/fingerprint.py:174-201 - restated:
//   X [L, D] rows of the domain (get_doms order)                       fingerprint.py:145-171
//   Y = idct_n(dct(X along L)[:n])   -> [n, D], per column min-max     fingerprint.py:126-142, :192
//   Z = idct_m(dct(Y along D)[:m])   -> [n, m], per row min-max        fingerprint.py:193
//   out[j*m + c] = trunc(Z[j, c] * 127) as int8                        fingerprint.py:194-195
// Both min-max steps are invariant to adding a constant and to a positive scale, so the DC
// coefficient and the orthonormal scale factors drop out:
//   u_k[d] = sum_l (x[l,d] - x[0,d]) * cos(pi (2l+1) k / 2L),  k = 1..n-1      ("pass 1")
//   Y[j,d] ~ sum_k cos(pi (2j+1) k / 2n) u_k[d]
//   F[j,k] = sum_d Y'[j,d] * cos(pi (2d+1) k / 2D),             k = 1..m-1      ("pass 2")
//   Z[j,c] ~ sum_k cos(pi (2c+1) k / 2m) F[j,k]
// Pass 1 is the HBM-bound part (every embedding element is read exactly once, n-1 FMAs each);
// everything after it works on O(n*D) values per (domain, layer) and stays in shared memory.
//
// Kernel organisation (B200: 148 SMs, persistent CTAs pulling work items from an atomic queue):
//   item = (domain, layer, row range <= 512 rows); a CTA streams the item's rows with 16-byte
//   no-allocate loads, U rows in flight per thread, float32 FMAs flushed into float64
//   accumulators every U rows; domains longer than 512 rows are split over several items whose
//   partial sums meet in the workspace, the last-arriving CTA (atomic ticket) finishing the
//   domain.  The maxlen windows of src/embedding.py:153-192 are consumed in place: rows covered
//   by two windows are loaded from both and averaged (a + b) * 0.5f exactly as embedding.py:186.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "dctd_internal.cuh"

namespace {

constexpr int kRowsPerItem = 512;  // max rows of one work item (size of the per-item basis table)
constexpr int kMaxN = 8;           // n in [2, 8]
constexpr int kMaxM = 128;         // m in [2, 128]
constexpr int kMaxThreads = 640;

struct Piece {        // a run of consecutive domain rows read from one (or two averaged) sources
    int32_t src_a, row_a;
    int32_t src_b, row_b;   // src_b < 0: single source
    int32_t nrows, l0;      // l0 = index of the first row within the concatenated domain
};
struct DomInfo {
    int32_t piece_off, n_pieces;
    int32_t L;              // rows of the concatenated domain
    int32_t nsplit;
    int32_t slab0;          // first partial-sum slab (layer-major: slab0 + layer*nsplit + split)
    int32_t counter0;       // arrival counter index base (counter0 + layer), -1 if nsplit == 1
    int32_t pad0, pad1;
};
struct Item {
    int32_t dom, layer;
    int32_t r0, r1;         // domain-local row range
    int32_t split, piece_first;
    int32_t pad0, pad1;
};

struct Layout {             // thread / shared-memory layout derived from (D, n, m)
    int vec;                // 4: float4 loads, 1: scalar loads (D % 4 != 0 or misaligned rows)
    int G;                  // column groups = ceil(D / vec)
    int CT;                 // column tiles per thread
    int RL;                 // row lanes (threads sharing a column group, interleaved rows)
    int T;                  // threads per CTA
    size_t smem;            // dynamic shared memory bytes
    size_t off_t4d, off_y, off_cb, off_scr, off_tm, off_mj;
};

struct Params {
    const float *const *src;
    const Piece *pieces;
    const DomInfo *doms;
    const Item *items;
    int *counters;
    double *partials;
    int8_t *out;
    int64_t ld, out_stride;
    int32_t n_src, n_items, n_layers;
    int32_t D, n, m;
    Layout lay;
};

}  // namespace

struct dctd_fp_plan {
    int32_t n_layers, D, n, m, n_src, n_dom;
    std::vector<Piece> pieces;
    std::vector<DomInfo> doms;
    std::vector<Item> items;
    int32_t n_counters;       // 1 (work queue) + split arrival counters
    int64_t n_slabs;          // partial-sum slabs of (n-1)*D doubles
    int64_t algo_bytes;
    // device blob layout (bytes from the workspace base)
    size_t off_pieces, off_doms, off_items, off_src, off_counters, off_partials, total;
    void *blob;               // host copy of [pieces | doms | items], pinned when possible
    bool blob_pinned;
    size_t blob_bytes;
};

namespace {

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
template <int VEC>
struct Vec;
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ static Vec load(const float *p) {
        Vec r;
        // streaming read: data is used exactly once, keep it out of L1
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3])
                     : "l"(p));
        return r;
    }
};
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ static Vec load(const float *p) {
        Vec r;
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p));
        return r;
    }
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> zero_vec() {
    Vec<VEC> r;
#pragma unroll
    for (int i = 0; i < VEC; ++i) r.v[i] = 0.f;
    return r;
}

// one row of the domain at this thread's columns: single source or the overlap average
template <int VEC, bool DUAL>
__device__ __forceinline__ Vec<VEC> load_row(const float *pa, const float *pb) {
    Vec<VEC> a = Vec<VEC>::load(pa);
    if (DUAL) {
        Vec<VEC> b = Vec<VEC>::load(pb);
#pragma unroll
        for (int i = 0; i < VEC; ++i) a.v[i] = (a.v[i] + b.v[i]) * 0.5f;   // embedding.py:186
    }
    return a;
}

// Streams rows [0, nr) of one piece (this thread: rows rl, rl+RL, ...), accumulating
// acc[k][v] += (x - pivot) * cb[row][k].
template <int K, int VEC, int U, bool DUAL>
__device__ __forceinline__ void stream_piece(const float *pa, const float *pb, int64_t ld, int nr,
                                             int rl, int RL, const float *cb,
                                             const float (&piv)[VEC], double (&acc)[K][VEC]) {
    int i = rl;
    const int64_t step = (int64_t)RL * ld;
    pa += (int64_t)rl * ld;
    if (DUAL) pb += (int64_t)rl * ld;
    for (; i < nr; i += U * RL) {
        Vec<VEC> x[U];
        const bool full = (i + (U - 1) * RL) < nr;
        if (full) {
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = load_row<VEC, DUAL>(pa + u * step, pb + u * step);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i + u * RL < nr) x[u] = load_row<VEC, DUAL>(pa + u * step, pb + u * step);
                else x[u] = zero_vec<VEC>();
            }
        }
        float a32[K][VEC];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v) a32[k][v] = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int row = i + u * RL;
            const bool ok = full || row < nr;
            float c[K];
#pragma unroll
            for (int k = 0; k < K; ++k) c[k] = ok ? cb[row * K + k] : 0.f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float t = x[u].v[v] - piv[v];
#pragma unroll
                for (int k = 0; k < K; ++k) a32[k][v] = fmaf(t, c[k], a32[k][v]);
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[k][v] += (double)a32[k][v];
        pa += U * step;
        if (DUAL) pb += U * step;
    }
}

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
// MAXT/MINB: launch bounds.  CTAs of <= 320 threads are compiled for 3 CTAs per SM (<= 64
// registers) so that ~3 x 320 threads x U x 16 B of loads are in flight per SM.
template <int K, int VEC, int U, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) fp_kernel(const Params p) {
    constexpr int N = K + 1;
    extern __shared__ __align__(16) unsigned char smem[];
    float *T4D = reinterpret_cast<float *>(smem + p.lay.off_t4d);    // cos(pi i / 2D), i < 4D
    float *Y = reinterpret_cast<float *>(smem + p.lay.off_y);        // [N][D] pass-1 result - 0.5
    float *cb = reinterpret_cast<float *>(smem + p.lay.off_cb);      // [rows][K] basis of the item
    double *scr = reinterpret_cast<double *>(smem + p.lay.off_scr);  // u sums, later F / Z
    double *Tm = reinterpret_cast<double *>(smem + p.lay.off_tm);    // cos(pi i / 2m), i < 4m
    double *Mj = reinterpret_cast<double *>(smem + p.lay.off_mj);    // [N][K] cos(pi (2j+1) k / 2n)
    __shared__ int s_item, s_flag, s_last;

    const int tid = threadIdx.x, T = blockDim.x;
    const int D = p.D, m = p.m;
    const int G = p.lay.G, RL = p.lay.RL, CT = p.lay.CT;

    // ---- per-CTA tables (the CTA is persistent: built once) ----
    for (int i = tid; i < 4 * D; i += T) T4D[i] = (float)cospi((double)i / (2.0 * D));
    for (int i = tid; i < 4 * m; i += T) Tm[i] = cospi((double)i / (2.0 * m));
    for (int i = tid; i < N * K; i += T) {
        const int j = i / K, k = i % K + 1;
        Mj[i] = cospi((double)((2 * j + 1) * k) / (2.0 * N));
    }

    const int rl = (RL > 1) ? tid / G : 0;
    const int g0 = (RL > 1) ? tid % G : tid;
    const bool lane_ok = (RL > 1) ? (rl < RL) : true;

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            s_item = atomicAdd(&p.counters[0], 1);
            s_flag = 0;
        }
        __syncthreads();
        const int it = s_item;
        if (it >= p.n_items) break;
        const Item item = p.items[it];
        const DomInfo dom = p.doms[item.dom];
        const int L = dom.L, r0 = item.r0, r1 = item.r1;

        // ---- basis for this item's rows: cos(pi (2l+1) k / 2L), argument reduced exactly ----
        for (int i = tid; i < (r1 - r0) * K; i += T) {
            const int l = r0 + i / K, k = i % K + 1;
            const long long q = ((long long)(2 * l + 1) * k) % (4LL * L);
            cb[i] = (float)cospi((double)q / (2.0 * L));
        }
        __syncthreads();

        const float *const *src = p.src + (int64_t)item.layer * p.n_src;
        const Piece first = p.pieces[dom.piece_off];
        for (int ct = 0; ct < CT; ++ct) {
            const int g = g0 + ct * T;
            const bool active = lane_ok && g < G;
            const int col = g * VEC;
            double acc[K][VEC];
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[k][v] = 0.0;
            if (active) {
                // pivot = row 0 of the domain (same for every split of the domain)
                float piv[VEC];
                {
                    const float *pa = src[first.src_a] + (int64_t)first.row_a * p.ld + col;
                    Vec<VEC> a = Vec<VEC>::load(pa);
                    if (first.src_b >= 0) {
                        Vec<VEC> b = Vec<VEC>::load(src[first.src_b] + (int64_t)first.row_b * p.ld + col);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) a.v[v] = (a.v[v] + b.v[v]) * 0.5f;
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) piv[v] = a.v[v];
                }
                for (int pi = item.piece_first; pi < dom.n_pieces; ++pi) {
                    const Piece pc = p.pieces[dom.piece_off + pi];
                    if (pc.l0 >= r1) break;
                    const int a = max(pc.l0, r0), b = min(pc.l0 + pc.nrows, r1);
                    if (a >= b) continue;
                    const float *pa = src[pc.src_a] + (int64_t)(pc.row_a + (a - pc.l0)) * p.ld + col;
                    const float *cbp = cb + (a - r0) * K;
                    if (pc.src_b < 0) {
                        stream_piece<K, VEC, U, false>(pa, pa, p.ld, b - a, rl, RL, cbp, piv, acc);
                    } else {
                        const float *pb = src[pc.src_b] + (int64_t)(pc.row_b + (a - pc.l0)) * p.ld + col;
                        stream_piece<K, VEC, U, true>(pa, pb, p.ld, b - a, rl, RL, cbp, piv, acc);
                    }
                }
            }
            // ---- u[k][d] into shared memory, row lanes added in a fixed order ----
            if (RL == 1) {
                if (active) {
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            if (col + v < D) scr[k * D + col + v] = acc[k][v];
                }
            } else {
                for (int r = 0; r < RL; ++r) {
                    if (active && rl == r) {
#pragma unroll
                        for (int k = 0; k < K; ++k)
#pragma unroll
                            for (int v = 0; v < VEC; ++v)
                                if (col + v < D) {
                                    if (r == 0) scr[k * D + col + v] = acc[k][v];
                                    else scr[k * D + col + v] += acc[k][v];
                                }
                    }
                    __syncthreads();
                }
            }
        }
        __syncthreads();

        // ---- domains split over several items: partial sums meet in the workspace ----
        if (dom.nsplit > 1) {
            double *slab = p.partials + ((int64_t)dom.slab0 + (int64_t)item.layer * dom.nsplit) * (K * D);
            double *mine = slab + (int64_t)item.split * (K * D);
            for (int i = tid; i < K * D; i += T) mine[i] = scr[i];
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const int ticket = atomicAdd(&p.counters[dom.counter0 + item.layer], 1);
                s_last = (ticket == dom.nsplit - 1);
            }
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
            for (int i = tid; i < K * D; i += T) {
                double s = 0.0;
                for (int sp = 0; sp < dom.nsplit; ++sp) s += __ldcg(slab + (int64_t)sp * (K * D) + i);
                scr[i] = s;
            }
            __syncthreads();
        }

        // ---- length-n inverse + per-column min-max (fingerprint.py:138-140 on [D, n]) ----
        for (int d = tid; d < D; d += T) {
            double u[K], y[N];
#pragma unroll
            for (int k = 0; k < K; ++k) u[k] = scr[k * D + d];
            double mn = INFINITY, mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) s = fma(Mj[j * K + k], u[k], s);
                y[j] = s;
                mn = fmin(mn, s);
                mx = fmax(mx, s);
            }
            bool bad = !(mx > mn);
#pragma unroll
            for (int j = 0; j < N; ++j) bad = bad || isnan(y[j]);
            if (bad) s_flag = 1;   // constant / non-finite column: the reference yields NaN -> all 0
#pragma unroll
            for (int j = 0; j < N; ++j) Y[j * D + d] = (float)((y[j] - mn) / (mx - mn) - 0.5);
        }
        __syncthreads();

        // ---- pass 2a: F[j][k] = sum_d Y[j][d] cos(pi (2d+1) k / 2D), k = 1..m-1 ----
        const int nk = m - 1;
        const int DS = max(1, T / nk);
        double *Fp = scr;                       // [DS][N][nk]
        double *Z = scr + (size_t)DS * N * nk;  // [N][m]
        for (int w = tid; w < nk * DS; w += T) {
            const int k = 1 + w % nk, ds = w / nk;
            const int d0 = (int)((int64_t)D * ds / DS), d1 = (int)((int64_t)D * (ds + 1) / DS);
            int idx = (int)(((long long)(2 * d0 + 1) * k) % (4LL * D));
            const int stepk = 2 * k;
            double f64[N];
            float f32[N];
#pragma unroll
            for (int j = 0; j < N; ++j) { f64[j] = 0.0; f32[j] = 0.f; }
            int run = 0;
            for (int d = d0; d < d1; ++d) {
                const float c = T4D[idx];
                idx += stepk;
                if (idx >= 4 * D) idx -= 4 * D;
#pragma unroll
                for (int j = 0; j < N; ++j) f32[j] = fmaf(Y[j * D + d], c, f32[j]);
                if (++run == 32) {
#pragma unroll
                    for (int j = 0; j < N; ++j) { f64[j] += (double)f32[j]; f32[j] = 0.f; }
                    run = 0;
                }
            }
#pragma unroll
            for (int j = 0; j < N; ++j) Fp[((size_t)ds * N + j) * nk + (k - 1)] = f64[j] + (double)f32[j];
        }
        __syncthreads();

        // ---- pass 2b: Z[j][c] = sum_k cos(pi (2c+1) k / 2m) F[j][k] ----
        for (int w = tid; w < N * m; w += T) {
            const int j = w / m, c = w % m;
            int idx = 0;
            const int stepc = 2 * c + 1;
            double z = 0.0;
            for (int k = 1; k <= nk; ++k) {
                idx += stepc;
                if (idx >= 4 * m) idx -= 4 * m;
                double f = 0.0;
                for (int ds = 0; ds < DS; ++ds) f += Fp[((size_t)ds * N + j) * nk + (k - 1)];
                z = fma(Tm[idx], f, z);
            }
            Z[w] = z;
        }
        __syncthreads();

        // ---- per-row min-max, *127, truncating int8 cast (fingerprint.py:193-195) ----
        const bool layer_bad = s_flag != 0;
        int8_t *out = p.out + (int64_t)item.dom * p.out_stride + (int64_t)item.layer * (N * m);
        for (int w = tid; w < N * m; w += T) {
            const int j = w / m;
            double mn = INFINITY, mx = -INFINITY;
            bool bad = layer_bad;
            for (int c = 0; c < m; ++c) {
                const double z = Z[j * m + c];
                bad = bad || isnan(z);
                mn = fmin(mn, z);
                mx = fmax(mx, z);
            }
            bad = bad || !(mx > mn);
            int q = 0;
            if (!bad) q = (int)(((Z[w] - mn) / (mx - mn)) * 127.0);
            out[w] = (int8_t)q;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host: layout, planner, launch
// ------------------------------------------------------------------------------------------
Layout make_layout(int D, int n, int m, bool vec4) {
    Layout l{};
    const int K = n - 1;
    l.vec = vec4 ? 4 : 1;
    l.G = (D + l.vec - 1) / l.vec;
    if (l.G <= kMaxThreads) {
        l.CT = 1;
        l.RL = std::max(1, 320 / l.G);
        l.T = (l.G * l.RL + 31) / 32 * 32;
    } else {
        l.CT = (l.G + kMaxThreads - 1) / kMaxThreads;
        l.RL = 1;
        l.T = ((l.G + l.CT - 1) / l.CT + 31) / 32 * 32;
    }
    l.T = std::max(l.T, 128);
    const int nk = m - 1;
    const int DS = std::max(1, l.T / nk);
    size_t off = 0;
    l.off_t4d = off; off += dctd::align_up((size_t)4 * D * sizeof(float), 16);
    l.off_y = off;   off += dctd::align_up((size_t)n * D * sizeof(float), 16);
    l.off_cb = off;  off += dctd::align_up((size_t)kRowsPerItem * K * sizeof(float), 16);
    const size_t scr = std::max((size_t)K * D, (size_t)DS * n * nk + (size_t)n * m) * sizeof(double);
    l.off_scr = off; off += dctd::align_up(scr, 16);
    l.off_tm = off;  off += dctd::align_up((size_t)4 * m * sizeof(double), 16);
    l.off_mj = off;  off += dctd::align_up((size_t)n * K * sizeof(double), 16);
    l.smem = off;
    return l;
}

int g_variant = 0;   // tuning hook, see dctd_fp_set_variant
typedef void (*KernelFn)(const Params);
template <int VEC, int U, int MAXT, int MINB>
KernelFn pick_k(int K) {
    switch (K) {
        case 1: return fp_kernel<1, VEC, U, MAXT, MINB>;
        case 2: return fp_kernel<2, VEC, U, MAXT, MINB>;
        case 3: return fp_kernel<3, VEC, U, MAXT, MINB>;
        case 4: return fp_kernel<4, VEC, U, MAXT, MINB>;
        case 5: return fp_kernel<5, VEC, U, MAXT, MINB>;
        case 6: return fp_kernel<6, VEC, U, MAXT, MINB>;
        case 7: return fp_kernel<7, VEC, U, MAXT, MINB>;
    }
    return nullptr;
}

}  // namespace

extern "C" {

int dctd_fp_plan_create(const dctd_fp_geometry *geo, dctd_fp_plan **out_plan) {
    if (!geo || !out_plan) return DCTD_ERR_ARG;
    *out_plan = nullptr;
    if (geo->n_layers < 1 || geo->D < 1 || geo->n < 2 || geo->n > kMaxN || geo->m < 2 ||
        geo->m > kMaxM || geo->m > geo->D || geo->n_src < 0 || geo->n_prot < 0 || geo->n_dom < 0)
        return DCTD_ERR_ARG;
    if (geo->n_dom > 0 && (!geo->dom_prot || !geo->dom_seg_off || !geo->seg_beg || !geo->seg_end ||
                           !geo->src_rows || !geo->prot_src0 || !geo->prot_nsrc))
        return DCTD_ERR_ARG;
    const int stride = geo->maxlen - geo->overlap;
    dctd_fp_plan *pl = new (std::nothrow) dctd_fp_plan();
    if (!pl) return DCTD_ERR_NOMEM;
    pl->n_layers = geo->n_layers; pl->D = geo->D; pl->n = geo->n; pl->m = geo->m;
    pl->n_src = geo->n_src; pl->n_dom = geo->n_dom;
    pl->blob = nullptr; pl->blob_pinned = false; pl->blob_bytes = 0;
    pl->algo_bytes = 0;
    int rc = DCTD_OK;
    try {
        // protein lengths
        std::vector<int64_t> plen((size_t)geo->n_prot);
        for (int p = 0; p < geo->n_prot && rc == DCTD_OK; ++p) {
            const int s0 = geo->prot_src0[p], ns = geo->prot_nsrc[p];
            if (ns < 1 || s0 < 0 || (int64_t)s0 + ns > geo->n_src) { rc = DCTD_ERR_ARG; break; }
            if (ns == 1) { plen[p] = geo->src_rows[s0]; continue; }
            if (geo->overlap < 0 || stride <= 0) { rc = DCTD_ERR_ARG; break; }
            if (stride < geo->overlap) { rc = DCTD_ERR_UNSUPPORTED; break; }  // rows in > 2 windows
            for (int c = 0; c < ns - 1; ++c)
                if (geo->src_rows[s0 + c] != geo->maxlen) { rc = DCTD_ERR_ARG; break; }
            const int last = geo->src_rows[s0 + ns - 1];
            if (last <= geo->overlap || last > geo->maxlen) rc = DCTD_ERR_ARG;
            plen[p] = (int64_t)(ns - 1) * stride + last;
        }
        int32_t n_counters = 1;
        int64_t n_slabs = 0;
        pl->doms.resize((size_t)geo->n_dom);
        for (int i = 0; i < geo->n_dom && rc == DCTD_OK; ++i) {
            const int p = geo->dom_prot[i];
            if (p < 0 || p >= geo->n_prot) { rc = DCTD_ERR_ARG; break; }
            const int s0 = geo->prot_src0[p], ns = geo->prot_nsrc[p];
            DomInfo di{};
            di.piece_off = (int32_t)pl->pieces.size();
            int64_t l0 = 0;
            for (int j = geo->dom_seg_off[i]; j < geo->dom_seg_off[i + 1]; ++j) {
                int64_t b = geo->seg_beg[j], e = geo->seg_end[j];
                if (b < 0 || e > plen[p] || e < b) { rc = DCTD_ERR_ARG; break; }
                while (b < e) {
                    Piece pc{};
                    int64_t run_end;
                    if (ns == 1) {
                        pc.src_a = s0; pc.row_a = (int32_t)b; pc.src_b = -1; pc.row_b = 0;
                        run_end = e;
                    } else {
                        int64_t c = std::min<int64_t>(b / stride, ns - 1);
                        const int64_t off = b - c * stride;
                        if (c >= 1 && off < geo->overlap) {   // covered by windows c-1 and c
                            pc.src_a = s0 + (int32_t)c - 1; pc.row_a = (int32_t)(off + stride);
                            pc.src_b = s0 + (int32_t)c;     pc.row_b = (int32_t)off;
                            run_end = std::min<int64_t>(e, c * stride + geo->overlap);
                        } else {
                            pc.src_a = s0 + (int32_t)c; pc.row_a = (int32_t)off; pc.src_b = -1; pc.row_b = 0;
                            run_end = (c < ns - 1) ? std::min<int64_t>(e, (c + 1) * stride) : e;
                        }
                    }
                    pc.nrows = (int32_t)(run_end - b);
                    pc.l0 = (int32_t)l0;
                    pl->pieces.push_back(pc);
                    l0 += pc.nrows;
                    b = run_end;
                }
            }
            if (rc != DCTD_OK) break;
            if (l0 < geo->n || l0 > (1 << 26)) { rc = DCTD_ERR_ARG; break; }  // reference: reshape fails for L < n
            di.n_pieces = (int32_t)pl->pieces.size() - di.piece_off;
            di.L = (int32_t)l0;
            di.nsplit = (int32_t)((l0 + kRowsPerItem - 1) / kRowsPerItem);
            di.slab0 = -1; di.counter0 = -1;
            if (di.nsplit > 1) {
                di.slab0 = (int32_t)n_slabs;
                n_slabs += (int64_t)di.nsplit * geo->n_layers;
                di.counter0 = n_counters;
                n_counters += geo->n_layers;
            }
            pl->doms[i] = di;
            pl->algo_bytes += (int64_t)geo->n_layers * l0 * geo->D * 4;
            const int rps = (int)((l0 + di.nsplit - 1) / di.nsplit);
            for (int layer = 0; layer < geo->n_layers; ++layer) {
                int pf = 0;
                for (int s = 0; s < di.nsplit; ++s) {
                    Item itm{};
                    itm.dom = i; itm.layer = layer; itm.split = s;
                    itm.r0 = s * rps; itm.r1 = (int32_t)std::min<int64_t>(l0, (int64_t)(s + 1) * rps);
                    while (pf < di.n_pieces) {
                        const Piece &pc = pl->pieces[di.piece_off + pf];
                        if (pc.l0 + pc.nrows > itm.r0) break;
                        ++pf;
                    }
                    itm.piece_first = pf;
                    pl->items.push_back(itm);
                }
            }
        }
        if (rc == DCTD_OK) {
            // longest items first: the atomic work queue then behaves like LPT scheduling
            std::stable_sort(pl->items.begin(), pl->items.end(), [](const Item &a, const Item &b) {
                return (a.r1 - a.r0) > (b.r1 - b.r0);
            });
            pl->n_counters = n_counters;
            pl->n_slabs = n_slabs;
            size_t off = 0;
            pl->off_pieces = off; off += dctd::align_up(pl->pieces.size() * sizeof(Piece), 256);
            pl->off_doms = off;   off += dctd::align_up(pl->doms.size() * sizeof(DomInfo), 256);
            pl->off_items = off;  off += dctd::align_up(pl->items.size() * sizeof(Item), 256);
            pl->blob_bytes = off;
            pl->off_src = off;      off += dctd::align_up((size_t)geo->n_layers * geo->n_src * sizeof(void *), 256);
            pl->off_counters = off; off += dctd::align_up((size_t)n_counters * sizeof(int), 256);
            pl->off_partials = off; off += dctd::align_up((size_t)n_slabs * (geo->n - 1) * geo->D * sizeof(double), 256);
            pl->total = off;
            if (pl->blob_bytes) {
                void *h = nullptr;
                if (cudaHostAlloc(&h, pl->blob_bytes, cudaHostAllocDefault) == cudaSuccess) {
                    pl->blob_pinned = true;
                } else {
                    (void)cudaGetLastError();   // no device (CPU-only planning): plain host memory
                    h = malloc(pl->blob_bytes);
                    if (!h) rc = DCTD_ERR_NOMEM;
                }
                if (h) {
                    memset(h, 0, pl->blob_bytes);
                    if (!pl->pieces.empty()) memcpy((char *)h + pl->off_pieces, pl->pieces.data(), pl->pieces.size() * sizeof(Piece));
                    if (!pl->doms.empty()) memcpy((char *)h + pl->off_doms, pl->doms.data(), pl->doms.size() * sizeof(DomInfo));
                    if (!pl->items.empty()) memcpy((char *)h + pl->off_items, pl->items.data(), pl->items.size() * sizeof(Item));
                    pl->blob = h;
                }
            }
        }
    } catch (const std::bad_alloc &) {
        rc = DCTD_ERR_NOMEM;
    }
    if (rc != DCTD_OK) {
        dctd_fp_plan_destroy(pl);
        return rc;
    }
    *out_plan = pl;
    return DCTD_OK;
}

void dctd_fp_plan_destroy(dctd_fp_plan *plan) {
    if (!plan) return;
    if (plan->blob) {
        if (plan->blob_pinned) cudaFreeHost(plan->blob);
        else free(plan->blob);
    }
    delete plan;
}

size_t dctd_fp_workspace_bytes(const dctd_fp_plan *plan) { return plan ? plan->total : 0; }
int64_t dctd_fp_algorithmic_bytes(const dctd_fp_plan *plan) { return plan ? plan->algo_bytes : 0; }
int32_t dctd_fp_num_items(const dctd_fp_plan *plan) { return plan ? (int32_t)plan->items.size() : 0; }

/* tuning hook (not in dctd.h): selects the pass-1 unroll / occupancy variant for n == 3 */
int dctd_fp_set_variant(int v) { g_variant = v; return DCTD_OK; }

/* test hook: copies the plan's pieces (6 int32 each) / items (8 int32 each) to host buffers */
int dctd_fp_plan_dump(const dctd_fp_plan *plan, int32_t *pieces, int64_t max_pieces, int32_t *items,
                      int64_t max_items, int64_t *n_pieces, int64_t *n_items) {
    if (!plan) return DCTD_ERR_ARG;
    if (n_pieces) *n_pieces = (int64_t)plan->pieces.size();
    if (n_items) *n_items = (int64_t)plan->items.size();
    if (pieces) memcpy(pieces, plan->pieces.data(), std::min<size_t>(max_pieces, plan->pieces.size()) * sizeof(Piece));
    if (items) memcpy(items, plan->items.data(), std::min<size_t>(max_items, plan->items.size()) * sizeof(Item));
    return DCTD_OK;
}

int dctd_fp_execute(const dctd_fp_plan *plan, const void *const *h_src_ptrs, int64_t ld, int8_t *d_out,
                    int64_t out_stride, void *d_workspace, size_t workspace_bytes, uint32_t flags,
                    void *stream_) {
    if (!plan || !d_workspace) return DCTD_ERR_ARG;
    if (plan->n_dom == 0) return DCTD_OK;
    if (!h_src_ptrs || !d_out || ld < plan->D || out_stride < (int64_t)plan->n_layers * plan->n * plan->m)
        return DCTD_ERR_ARG;
    if (workspace_bytes < plan->total) return DCTD_ERR_WORKSPACE;
    if (((uintptr_t)d_workspace & 255) != 0) return DCTD_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    char *ws = (char *)d_workspace;

    // float4 path needs 16-byte aligned rows in every source
    bool vec4 = (plan->D % 4 == 0) && (ld % 4 == 0);
    const size_t nptr = (size_t)plan->n_layers * plan->n_src;
    for (size_t i = 0; i < nptr && vec4; ++i)
        if (((uintptr_t)h_src_ptrs[i] & 15) != 0) vec4 = false;

    Params prm{};
    prm.lay = make_layout(plan->D, plan->n, plan->m, vec4);
    int dev = 0, max_smem = 0, n_sm = 0;
    DCTD_CUDA_TRY(cudaGetDevice(&dev));
    DCTD_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    DCTD_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    if (prm.lay.smem > (size_t)max_smem) return DCTD_ERR_UNSUPPORTED;

    if (!(flags & DCTD_FP_TABLES_RESIDENT))
        DCTD_CUDA_TRY(cudaMemcpyAsync(ws, plan->blob, plan->blob_bytes, cudaMemcpyHostToDevice, stream));
    DCTD_CUDA_TRY(cudaMemcpyAsync(ws + plan->off_src, h_src_ptrs, nptr * sizeof(void *), cudaMemcpyHostToDevice, stream));
    DCTD_CUDA_TRY(cudaMemsetAsync(ws + plan->off_counters, 0, (size_t)plan->n_counters * sizeof(int), stream));

    prm.src = (const float *const *)(ws + plan->off_src);
    prm.pieces = (const Piece *)(ws + plan->off_pieces);
    prm.doms = (const DomInfo *)(ws + plan->off_doms);
    prm.items = (const Item *)(ws + plan->off_items);
    prm.counters = (int *)(ws + plan->off_counters);
    prm.partials = (double *)(ws + plan->off_partials);
    prm.out = d_out;
    prm.ld = ld;
    prm.out_stride = out_stride;
    prm.n_src = plan->n_src;
    prm.n_items = (int32_t)plan->items.size();
    prm.n_layers = plan->n_layers;
    prm.D = plan->D; prm.n = plan->n; prm.m = plan->m;

    const int K = plan->n - 1;
    KernelFn fn;
    if (!vec4) fn = pick_k<1, 8, kMaxThreads, 1>(K);
    else if (prm.lay.T <= 320) {
        // tuning variants of the common case n = 3 (see DESIGN.md "pass-1 occupancy")
        if (K == 2 && g_variant == 1) fn = fp_kernel<2, 4, 4, 320, 3>;
        else if (K == 2 && g_variant == 2) fn = fp_kernel<2, 4, 8, 320, 3>;
        else fn = pick_k<4, 8, 320, 2>(K);
    } else fn = pick_k<4, 8, kMaxThreads, 1>(K);
    if (!fn) return DCTD_ERR_UNSUPPORTED;
    DCTD_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prm.lay.smem));
    int per_sm = 0;
    DCTD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, prm.lay.T, prm.lay.smem));
    if (per_sm < 1) return DCTD_ERR_UNSUPPORTED;
    const int grid = std::max(1, std::min(prm.n_items, n_sm * per_sm));
    fn<<<grid, prm.lay.T, prm.lay.smem, stream>>>(prm);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

}  // extern "C"
