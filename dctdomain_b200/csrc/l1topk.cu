// libdctd hot path 2: exhaustive int8 L1 top-k for sm_100a.
//
// Stands in for faiss IndexFlat + METRIC_L1 `index.search(x, k)` as called by the reference
// (src/query_db.py:75-76,87; bench/cathdb/run_dct.py:51-60) on the flat index built at
// src/database.py:240-243, and for the pairwise scorer of src/dct-sim.py:12-50.
// Semantics: per query the k smallest database positions by (L1 distance, position), ascending,
// padded with (FLT_MAX, -1).
//
// Design
//   * The database lives on the device as int8 (the reference's fingerprints ARE int8; faiss
//     widens them to float32, 4x the bytes) in a packed layout: groups of 32 vectors interleaved
//     in 16-byte chunks, bytes XOR 0x80 so that an unsigned packed-byte SAD gives |a-b| for any
//     int8 pair.  A group tile is contiguous, so one TMA bulk copy (cp.async.bulk, mbarrier
//     completion) stages it in shared memory, and lane v of a warp reads vector v of the group
//     with conflict-free 16-byte shared loads.
//   * Distances: one VABSDIFF4.U8.ACC (__vsadu4) per 4 bytes per (query, vector) pair, register
//     tile of TQ queries x TD vectors per thread, queries broadcast from shared memory.
//   * Selection: 64-bit keys (dist << 40 | position) make (dist, position) order a plain integer
//     compare.  Each warp owns its queries' candidate buffers in shared memory: a candidate is
//     appended only if its key beats the query's current k-th key (warp ballot + prefix), and a
//     full buffer is re-sorted (bitonic, warp-synchronous) to tighten the threshold.
//   * The database is split over grid.x; per-split sorted lists meet in a merge kernel that also
//     converts to faiss' (float32 distance, int64 id) output.  The same merge kernel folds the
//     per-rank results of a sharded database after the NCCL all-gather.
#include <float.h>

#include <algorithm>

#include "dctd_internal.cuh"

namespace {

constexpr int kWarps = 8;            // warps per CTA of the scan kernel
constexpr int kTDmax = 2;            // database groups (of 32 vectors) per shared-memory stage (1 for wide d)
constexpr int kStagesMax = 4;        // TMA stages (2 for wide d)
constexpr int kIdBits = 40;          // key = dist << 40 | position
constexpr unsigned long long kKeyMax = ~0ULL;
constexpr unsigned long long kIdMask = (1ULL << kIdBits) - 1ULL;

__host__ __device__ inline int chunks_of(int d) { return (d + 15) / 16; }

// ------------------------------------------------------------------------------------------
// pack / unpack
// ------------------------------------------------------------------------------------------
__global__ void pack_kernel(const int8_t *__restrict__ rows, long long n, int d, long long n_offset,
                            uint4 *__restrict__ packed) {
    const int C = chunks_of(d);
    const long long v0 = n_offset / 32 * 32;                 // first vector of the first touched group
    const long long v1 = (n_offset + n + 31) / 32 * 32;      // end of the last touched group
    const long long total = (v1 - v0) * C;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long vec = v0 + t / C;
        const int c = (int)(t % C);
        if (vec < n_offset) continue;                        // already packed by an earlier add
        unsigned int w[4] = {0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u};
        if (vec < n_offset + n) {
            const int8_t *src = rows + (vec - n_offset) * d + c * 16;
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const int col = c * 16 + b;
                const unsigned int byte = (col < d) ? ((unsigned int)(unsigned char)src[b] ^ 0x80u) : 0x80u;
                w[b >> 2] = (w[b >> 2] & ~(0xffu << ((b & 3) * 8))) | (byte << ((b & 3) * 8));
            }
        }
        const long long g = vec / 32;
        const int lane = (int)(vec % 32);
        packed[(g * C + c) * 32 + lane] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void unpack_kernel(const uint4 *__restrict__ packed, long long n, int d, int8_t *__restrict__ rows) {
    const int C = chunks_of(d);
    const long long total = n * C;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long vec = t / C;
        const int c = (int)(t % C);
        const uint4 q = packed[((vec / 32) * C + c) * 32 + (vec % 32)];
        const unsigned int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const int col = c * 16 + b;
            if (col < d) rows[vec * d + col] = (int8_t)(((w[b >> 2] >> ((b & 3) * 8)) & 0xffu) ^ 0x80u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// warp-synchronous bitonic sort of `cap` (power of two) keys in shared memory, ascending
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_sort(unsigned long long *buf, int cap, int lane) {
    for (int k2 = 2; k2 <= cap; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < (cap >> 1); i += 32) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const bool up = (lo & k2) == 0;
                const unsigned long long a = buf[lo], b = buf[hi];
                if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
            }
            __syncwarp();
        }
    }
}

// sorts the `cnt` live keys of a candidate buffer, keeps the k best, returns the new count
__device__ __forceinline__ int warp_refine(unsigned long long *buf, int cnt, int cap, int k, int lane,
                                           unsigned long long *thr) {
    for (int i = cnt + lane; i < cap; i += 32) buf[i] = kKeyMax;
    __syncwarp();
    warp_sort(buf, cap, lane);
    const int keep = min(cnt, k);
    if (lane == 0) *thr = (keep == k) ? buf[k - 1] : kKeyMax;
    __syncwarp();
    return keep;
}

// ------------------------------------------------------------------------------------------
// mbarrier / TMA bulk copy (cp.async.bulk: SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void *p) {
    return (unsigned int)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
    unsigned int done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// producer-side wait: sleeps between probes so that the spinning lane does not take issue slots from
// the compute warps of its scheduler
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long *bar, unsigned int parity) {
    unsigned int done = 0;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(400);
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// acc + sum_i |a.byte[i] - b.byte[i]| (unsigned bytes): SASS VABSDIFF4.U8.ACC
__device__ __forceinline__ unsigned int sad4(unsigned int a, unsigned int b, unsigned int acc) {
    unsigned int r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
}

// ------------------------------------------------------------------------------------------
// scan kernel
// ------------------------------------------------------------------------------------------
struct ScanParams {
    const int8_t *q;            // [nq, d] row-major
    const uint4 *packed;        // packed database
    unsigned long long *parts;  // [S, nq, k] sorted keys per database split
    long long nq, n;
    int d, k, cap;
    long long n_groups, groups_per_split;
};

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised scan: warp kWarps is the TMA producer (one elected lane), warps 0..kWarps-1 compute.
// Stages are handed over with full/empty mbarriers, so compute warps never meet at a CTA barrier and
// drift apart: while one warp filters candidates, the others keep the integer pipe busy.
template <int TQ, int kTD, int STAGES>
__global__ void __launch_bounds__((kWarps + 1) * 32, 1) l1_scan_kernel(const ScanParams p) {
    constexpr int QT = kWarps * TQ;
    extern __shared__ __align__(128) unsigned char smem[];
    const int C = chunks_of(p.d);
    const int dpad = C * 16;
    const int tile_bytes = kTD * 32 * dpad;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);           // [STAGES]
    unsigned long long *empty = full + STAGES;                                          // [STAGES]
    unsigned char *qs = smem + 128;                                                     // [QT][dpad]
    unsigned char *st = qs + (size_t)QT * dpad;                                         // [STAGES][tile]
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(st + (size_t)STAGES * tile_bytes);
    unsigned long long *thr = buf + (size_t)QT * p.cap;                                 // [QT]
    int *cnt = reinterpret_cast<int *>(thr + QT);                                       // [QT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long q0 = (long long)blockIdx.y * QT;
    const long long g_begin = (long long)blockIdx.x * p.groups_per_split;
    const long long g_end = min(p.n_groups, g_begin + p.groups_per_split);
    const int n_tiles = (int)((g_end - g_begin + kTD - 1) / kTD);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // queries -> shared memory, biased by 0x80, zero-distance padding beyond d / nq
    for (int i = tid; i < QT * dpad; i += blockDim.x) {
        const int qi = i / dpad, col = i % dpad;
        unsigned char b = 0x80;
        if (q0 + qi < p.nq && col < p.d) b = (unsigned char)p.q[(q0 + qi) * p.d + col] ^ 0x80;
        qs[i] = b;
    }
    for (int i = tid; i < QT; i += blockDim.x) { thr[i] = kKeyMax; cnt[i] = 0; }
    __syncthreads();

    if (warp == kWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % STAGES;
                if (t >= STAGES) mbar_wait_backoff(&empty[s], (unsigned int)(((t / STAGES) - 1) & 1));
                const long long g = g_begin + (long long)t * kTD;
                const int ng = (int)min((long long)kTD, g_end - g);
                const unsigned int bytes = (unsigned int)ng * 32u * (unsigned int)dpad;
                mbar_expect_tx(&full[s], bytes);
                bulk_g2s(st + (size_t)s * tile_bytes, p.packed + g * C * 32, bytes, &full[s]);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const uint4 *qs4 = reinterpret_cast<const uint4 *>(qs) + (size_t)warp * TQ * C;
    unsigned long long *wbuf = buf + (size_t)warp * TQ * p.cap;
    unsigned long long *wthr = thr + warp * TQ;
    int *wcnt = cnt + warp * TQ;
    unsigned int tdist[TQ];      // distance part of each query's current k-th key (fast reject)
#pragma unroll
    for (int a = 0; a < TQ; ++a) tdist[a] = 0xffffffffu;

    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (unsigned int)((t / STAGES) & 1));
        const uint4 *st4 = reinterpret_cast<const uint4 *>(st + (size_t)s * tile_bytes);
        unsigned int acc[TQ][kTD];
#pragma unroll
        for (int a = 0; a < TQ; ++a)
#pragma unroll
            for (int b = 0; b < kTD; ++b) acc[a][b] = 0u;
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            uint4 dv[kTD], qv[TQ];
#pragma unroll
            for (int b = 0; b < kTD; ++b) dv[b] = st4[(b * C + c) * 32 + lane];
#pragma unroll
            for (int a = 0; a < TQ; ++a) qv[a] = qs4[a * C + c];
            // word-major order: TQ*kTD independent accumulators between two updates of the same one
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < kTD; ++b) acc[a][b] = sad4(qv[a].x, dv[b].x, acc[a][b]);
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < kTD; ++b) acc[a][b] = sad4(qv[a].y, dv[b].y, acc[a][b]);
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < kTD; ++b) acc[a][b] = sad4(qv[a].z, dv[b].z, acc[a][b]);
#pragma unroll
            for (int a = 0; a < TQ; ++a)
#pragma unroll
                for (int b = 0; b < kTD; ++b) acc[a][b] = sad4(qv[a].w, dv[b].w, acc[a][b]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);     // this warp no longer reads the stage

        // ---- candidate filter (warp-private buffers: no CTA-level synchronisation) ----
        bool hit = false;
#pragma unroll
        for (int a = 0; a < TQ; ++a)
#pragma unroll
            for (int b = 0; b < kTD; ++b) hit = hit || (acc[a][b] <= tdist[a]);
        if (!__any_sync(0xffffffffu, hit)) continue;          // common case once thresholds are tight
        const long long gbase = g_begin + (long long)t * kTD;
#pragma unroll
        for (int a = 0; a < TQ; ++a) {
#pragma unroll
            for (int b = 0; b < kTD; ++b) {
                const long long id = (gbase + b) * 32 + lane;
                const bool valid = (gbase + b) < g_end && id < p.n;
                const unsigned long long key = ((unsigned long long)acc[a][b] << kIdBits) | (unsigned long long)id;
                const bool pass = valid && key < wthr[a];
                const unsigned int vote = __ballot_sync(0xffffffffu, pass);
                if (vote) {
                    int cn = wcnt[a];
                    if (cn + 32 > p.cap) cn = warp_refine(wbuf + (size_t)a * p.cap, cn, p.cap, p.k, lane, &wthr[a]);
                    // keys that stopped passing after a refine are harmless: they sort behind the k-th key
                    if (pass) wbuf[(size_t)a * p.cap + cn + __popc(vote & ((1u << lane) - 1u))] = key;
                    __syncwarp();
                    if (lane == 0) wcnt[a] = cn + __popc(vote);
                    __syncwarp();
                }
            }
            const unsigned long long th = wthr[a];
            tdist[a] = (th == kKeyMax) ? 0xffffffffu : (unsigned int)(th >> kIdBits);
        }
    }

    // ---- final: sort each query's candidates and publish the k best keys of this split ----
    for (int a = 0; a < TQ; ++a) {
        const long long qi = q0 + warp * TQ + a;
        unsigned long long *b = wbuf + (size_t)a * p.cap;
        const int keep = warp_refine(b, wcnt[a], p.cap, p.k, lane, &wthr[a]);
        if (qi < p.nq) {
            unsigned long long *out = p.parts + ((long long)blockIdx.x * p.nq + qi) * p.k;
            for (int i = lane; i < p.k; i += 32) out[i] = (i < keep) ? b[i] : kKeyMax;
        }
    }
}

// ------------------------------------------------------------------------------------------
// merge kernel: `parts` sorted lists per query -> k best, converted to (float32, int64)
// ------------------------------------------------------------------------------------------
struct MergeParams {
    const unsigned long long *key_parts;   // [parts, nq, k] or null
    const float *dist_parts;               // [parts, nq, k] (PAIRS input)
    const long long *id_parts;
    int parts;
    long long nq;
    int k, cap;                            // cap = power of two >= 2k
    long long id_base;
    float *dist;
    long long *ids;
};

template <bool PAIRS>
__global__ void __launch_bounds__(128) l1_merge_kernel(const MergeParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(smem) + (size_t)warp * p.cap;
    const long long qi = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (qi >= p.nq) return;
    auto load = [&](int part, int i) -> unsigned long long {
        const long long o = ((long long)part * p.nq + qi) * p.k + i;
        if (!PAIRS) return p.key_parts[o];
        const long long id = p.id_parts[o];
        if (id < 0) return kKeyMax;
        return ((unsigned long long)(unsigned int)p.dist_parts[o] << kIdBits) | (unsigned long long)id;
    };
    for (int i = lane; i < p.k; i += 32) buf[i] = load(0, i);
    for (int s = 1; s < p.parts; ++s) {
        for (int i = lane; i < p.k; i += 32) buf[p.k + i] = load(s, i);
        for (int i = 2 * p.k + lane; i < p.cap; i += 32) buf[i] = kKeyMax;
        __syncwarp();
        warp_sort(buf, p.cap, lane);
    }
    __syncwarp();
    for (int i = lane; i < p.k; i += 32) {
        const unsigned long long key = buf[i];
        const long long o = qi * p.k + i;
        if (key == kKeyMax) {
            p.dist[o] = FLT_MAX;
            p.ids[o] = -1;
        } else {
            p.dist[o] = (float)(unsigned int)(key >> kIdBits);
            p.ids[o] = (long long)(key & kIdMask) + p.id_base;
        }
    }
}

// ------------------------------------------------------------------------------------------
// pairwise scorer (dct-sim.py:12-50): one warp per protein pair
// ------------------------------------------------------------------------------------------
__global__ void pair_scores_kernel(const int8_t *__restrict__ fps, int d, const long long *__restrict__ off,
                                   const int *__restrict__ pa, const int *__restrict__ pb, long long n_pairs,
                                   int *__restrict__ min_dist, int *__restrict__ last_dist) {
    const int lane = threadIdx.x & 31;
    const long long pr = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pr >= n_pairs) return;
    const long long a0 = off[pa[pr]], a1 = off[pa[pr] + 1], b0 = off[pb[pr]], b1 = off[pb[pr] + 1];
    int best = 0x7fffffff, last = 0;
    for (long long i = a0; i < a1; ++i)
        for (long long j = b0; j < b1; ++j) {
            int s = 0;
            for (int c = lane; c < d; c += 32) s += abs((int)fps[i * d + c] - (int)fps[j * d + c]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            best = min(best, s);
            last = s;
        }
    if (lane == 0) {
        min_dist[pr] = best;
        last_dist[pr] = last;
    }
}

// ------------------------------------------------------------------------------------------
// host-side configuration shared by workspace_bytes and topk
// ------------------------------------------------------------------------------------------
struct ScanConfig {
    int tq, td, stages, cap;
    long long n_qtiles, n_groups, splits, groups_per_split;
    size_t smem;
};

bool make_config(long long nq, long long n, int d, int k, ScanConfig *cfg) {
    if (k < 1 || k > 992 || d < 1 || d > 2048) return false;
    int cap = 128, tq = 8;
    if (k > 96) { cap = 256; tq = 4; }
    if (k > 224) { cap = 512; tq = 2; }
    if (k > 480) { cap = 1024; tq = 1; }
    while (tq > 1 && (long long)kWarps * (tq / 2) >= nq) tq /= 2;   // few queries: smaller tiles
    const int dpad = chunks_of(d) * 16;
    int td = kTDmax, stages = kStagesMax;
    auto smem_of = [&](int tq_, int td_, int st_) {
        const size_t qt = (size_t)kWarps * tq_;
        return (size_t)128 + qt * dpad + (size_t)st_ * td_ * 32 * dpad + qt * cap * 8 + qt * 12 + 64;
    };
    const size_t lim = 227 * 1024;
    if (smem_of(1, td, stages) > lim) { td = 1; stages = 2; }      // wide vectors
    while (tq > 1 && smem_of(tq, td, stages) > lim) tq /= 2;
    if (smem_of(tq, td, stages) > lim) return false;
    const int kTD = td;
    cfg->tq = tq;
    cfg->td = td;
    cfg->stages = stages;
    cfg->cap = cap;
    cfg->smem = smem_of(tq, td, stages);
    cfg->n_qtiles = (nq + (long long)kWarps * tq - 1) / ((long long)kWarps * tq);
    cfg->n_groups = (n + 31) / 32;
    const long long tiles = std::max<long long>(1, (cfg->n_groups + kTD - 1) / kTD);
    // One CTA per SM is resident (shared memory), so pick the number of database splits that fills
    // whole waves of 148 CTAs: the smallest count (>= 1 wave, >= 8 tiles per split so the per-split
    // warm-up of the candidate buffers stays small) whose last wave is >= 96 % full, else the fullest.
    const long long sms = 148;
    const long long s_max = std::max<long long>(1, std::min<long long>(1024, tiles / 8));
    const long long s_min = std::min(s_max, std::max<long long>(1, (sms + cfg->n_qtiles - 1) / cfg->n_qtiles));
    long long best_s = s_min;
    double best_eff = -1.0;
    for (long long sp = s_min; sp <= std::min(s_max, s_min + 4 * sms); ++sp) {
        const long long tps = (tiles + sp - 1) / sp;
        const long long real = (tiles + tps - 1) / tps;            // splits actually launched
        const double waves = (double)(real * cfg->n_qtiles) / (double)sms;
        const double eff = waves / (double)((long long)(waves + 0.999999));
        if (eff > best_eff + 1e-9) { best_eff = eff; best_s = sp; }
        if (eff >= 0.96) { best_s = sp; break; }
    }
    const long long tiles_per_split = (tiles + best_s - 1) / best_s;
    cfg->groups_per_split = tiles_per_split * kTD;
    cfg->splits = (cfg->n_groups + cfg->groups_per_split - 1) / std::max<long long>(1, cfg->groups_per_split);
    if (cfg->splits < 1) cfg->splits = 1;
    return true;
}

int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace

extern "C" {

size_t dctd_l1_packed_bytes(int64_t n, int32_t d) {
    if (n < 0 || d < 1) return 0;
    return (size_t)((n + 31) / 32) * 32 * (size_t)chunks_of(d) * 16;
}

int dctd_l1_pack(const int8_t *d_rows, int64_t n, int32_t d, int64_t n_offset, void *d_packed, void *stream) {
    if (n < 0 || d < 1 || n_offset < 0 || (n > 0 && (!d_rows || !d_packed))) return DCTD_ERR_ARG;
    if (n == 0) return DCTD_OK;
    const long long total = ((n_offset + n + 31) / 32 * 32 - n_offset / 32 * 32) * chunks_of(d);
    const int block = 256;
    const int grid = (int)std::min<long long>((total + block - 1) / block, 148 * 16);
    pack_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_rows, n, d, n_offset, (uint4 *)d_packed);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

int dctd_l1_unpack(const void *d_packed, int64_t n, int32_t d, int8_t *d_rows, void *stream) {
    if (n < 0 || d < 1 || (n > 0 && (!d_rows || !d_packed))) return DCTD_ERR_ARG;
    if (n == 0) return DCTD_OK;
    const long long total = n * chunks_of(d);
    const int block = 256;
    const int grid = (int)std::min<long long>((total + block - 1) / block, 148 * 16);
    unpack_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint4 *)d_packed, n, d, d_rows);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

size_t dctd_l1_topk_workspace_bytes(int64_t nq, int64_t n, int32_t d, int32_t k) {
    ScanConfig cfg;
    if (nq <= 0 || n < 0 || !make_config(nq, n, d, k, &cfg)) return 0;
    return dctd::align_up((size_t)cfg.splits * (size_t)nq * (size_t)k * 8, 256) + 256;
}

int dctd_l1_topk(const int8_t *d_q, int64_t nq, const void *d_packed, int64_t n, int32_t d, int32_t k,
                 int64_t id_base, float *d_dist, int64_t *d_ids, void *d_workspace, size_t workspace_bytes,
                 void *stream_) {
    if (nq < 0 || n < 0 || k < 1 || d < 1) return DCTD_ERR_ARG;
    if (nq == 0) return DCTD_OK;
    if (!d_q || !d_dist || !d_ids || (n > 0 && !d_packed)) return DCTD_ERR_ARG;
    if (n >= (1LL << kIdBits)) return DCTD_ERR_UNSUPPORTED;
    ScanConfig cfg;
    if (!make_config(nq, n, d, k, &cfg)) return DCTD_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t need = dctd_l1_topk_workspace_bytes(nq, n, d, k);
    if (!d_workspace || workspace_bytes < need) return DCTD_ERR_WORKSPACE;
    if (((uintptr_t)d_workspace & 255) != 0 || ((uintptr_t)d_packed & 15) != 0) return DCTD_ERR_ARG;

    MergeParams mp{};
    mp.nq = nq; mp.k = k; mp.id_base = id_base; mp.dist = d_dist; mp.ids = (long long *)d_ids;
    mp.cap = next_pow2(2 * k);
    if (n == 0) {
        // empty database: all slots are padding.  Reuse the merge kernel on one all-MAX part.
        DCTD_CUDA_TRY(cudaMemsetAsync(d_workspace, 0xff, (size_t)nq * k * 8, stream));
        mp.key_parts = (const unsigned long long *)d_workspace;
        mp.parts = 1;
    } else {
        ScanParams sp{};
        sp.q = d_q; sp.packed = (const uint4 *)d_packed; sp.parts = (unsigned long long *)d_workspace;
        sp.nq = nq; sp.n = n; sp.d = d; sp.k = k; sp.cap = cfg.cap;
        sp.n_groups = cfg.n_groups; sp.groups_per_split = cfg.groups_per_split;
        if (cfg.n_qtiles > 65535) return DCTD_ERR_UNSUPPORTED;   // caller batches queries
        dim3 grid((unsigned)cfg.splits, (unsigned)cfg.n_qtiles);
        void (*fn)(const ScanParams) = nullptr;
        if (cfg.td == kTDmax) {
            switch (cfg.tq) {
                case 1: fn = l1_scan_kernel<1, kTDmax, kStagesMax>; break;
                case 2: fn = l1_scan_kernel<2, kTDmax, kStagesMax>; break;
                case 4: fn = l1_scan_kernel<4, kTDmax, kStagesMax>; break;
                default: fn = l1_scan_kernel<8, kTDmax, kStagesMax>; break;
            }
        } else {
            switch (cfg.tq) {
                case 1: fn = l1_scan_kernel<1, 1, 2>; break;
                case 2: fn = l1_scan_kernel<2, 1, 2>; break;
                case 4: fn = l1_scan_kernel<4, 1, 2>; break;
                default: fn = l1_scan_kernel<8, 1, 2>; break;
            }
        }
        DCTD_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
        fn<<<grid, (kWarps + 1) * 32, cfg.smem, stream>>>(sp);
        DCTD_LAUNCH_CHECK();
        mp.key_parts = sp.parts;
        mp.parts = (int)cfg.splits;
    }
    const int warps = 4;
    const size_t msmem = (size_t)warps * mp.cap * 8;
    DCTD_CUDA_TRY(cudaFuncSetAttribute(l1_merge_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    l1_merge_kernel<false><<<(unsigned)((nq + warps - 1) / warps), warps * 32, msmem, stream>>>(mp);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

int dctd_l1_topk_merge(const float *d_dist_parts, const int64_t *d_ids_parts, int32_t parts, int64_t nq,
                       int32_t k, float *d_dist, int64_t *d_ids, void *stream) {
    if (parts < 1 || nq < 0 || k < 1 || k > 1024) return DCTD_ERR_ARG;
    if (nq == 0) return DCTD_OK;
    if (!d_dist_parts || !d_ids_parts || !d_dist || !d_ids) return DCTD_ERR_ARG;
    MergeParams mp{};
    mp.dist_parts = d_dist_parts; mp.id_parts = (const long long *)d_ids_parts;
    mp.parts = parts; mp.nq = nq; mp.k = k; mp.cap = next_pow2(2 * k); mp.id_base = 0;
    mp.dist = d_dist; mp.ids = (long long *)d_ids;
    const int warps = 4;
    const size_t msmem = (size_t)warps * mp.cap * 8;
    DCTD_CUDA_TRY(cudaFuncSetAttribute(l1_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    l1_merge_kernel<true><<<(unsigned)((nq + warps - 1) / warps), warps * 32, msmem, (cudaStream_t)stream>>>(mp);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

int dctd_l1_pair_scores(const int8_t *d_fps, int32_t d, const int64_t *d_off, const int32_t *d_pair_a,
                        const int32_t *d_pair_b, int64_t n_pairs, int32_t *d_min_dist, int32_t *d_last_dist,
                        void *stream) {
    if (n_pairs < 0 || d < 1) return DCTD_ERR_ARG;
    if (n_pairs == 0) return DCTD_OK;
    if (!d_fps || !d_off || !d_pair_a || !d_pair_b || !d_min_dist || !d_last_dist) return DCTD_ERR_ARG;
    const int warps = 4;
    pair_scores_kernel<<<(unsigned)((n_pairs + warps - 1) / warps), warps * 32, 0, (cudaStream_t)stream>>>(
        d_fps, d, (const long long *)d_off, d_pair_a, d_pair_b, n_pairs, d_min_dist, d_last_dist);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

}  // extern "C"
