// Fused streaming search for a handful of queries (nq <= 16): the HBM-bound regime of hot path 2.
// Included by l1topk.cu inside its anonymous namespace (uses ScanParams, SelectParams, sad4, select_query, ...).
//
// The reference calls index.search with the 1-13 fingerprints of ONE protein (src/query_db.py:87): the database is
// streamed once per call and every byte meets only a few queries, so the call is bound by HBM, not by the SAD pipe.
// One cooperative launch of one persistent CTA per SM does the whole call:
//
//   phase A  threshold sample: every gstride-th group goes through the ring; each lane keeps the smallest distance
//            it has seen per query -> `mins` (one value per warp and CTA: each is the distance of a distinct real
//            vector, so the k-th smallest of them bounds the true k-th best distance from above);
//   grid barrier; CTA q sorts the minima of query q and publishes its bound; grid barrier;
//   phase B  every group goes through the ring; vectors within the bound are appended to the query's candidate list
//            (warp-aggregated atomics; ~k * gstride expected);
//   grid barrier; CTA q selects the k best of query q's candidates and writes the result.
//
// Warp roles: warp NW is the producer - one lane issues TMA bulk copies (cp.async.bulk; a 32-vector group is one
// contiguous slab, staged as NH column parts) into a ring of S part-sized slots with full / empty mbarriers; it never
// takes part in the grid barriers, so the ring is already full of phase-B groups when the consumers come back from
// them.  Warps 0..NW-1 consume: warp w takes turns w, w + NW, ... of TD groups each, lane = vector, TQ x TD
// accumulators, queries broadcast from shared memory.  Bytes in flight are set by the ring, not by registers.
#pragma once

struct StreamParams {
    ScanParams sp;              // q, packed, nq, n, d, k, gstride (sample stride), thr_dist (external bound or null),
                                // cand, cnt, cmax
    SelectParams se;
    unsigned int *mins;         // [16][kMinSlots]
    unsigned int *thr_out;      // [16] bounds computed in-kernel
    unsigned int *bar;          // grid barrier counter (zeroed before the launch); bar[1]: next phase-B group to hand out
    int stages;                 // ring slots
    int m;                      // minima per query = gridDim.x * NW
    int kscr_n;                 // entries of the k-th scratch (>= m)
};

constexpr int kStreamMaxSlots = 64;
constexpr int kStreamHeader = 1664;    // full / empty mbarriers and the staged group of every slot, issued-slot counter
constexpr int kStreamGrab = 8;         // consecutive groups a CTA takes per grab of the phase-B work counter

// barrier over all CTAs of the (cooperative, hence co-resident) grid, for the nthr consumer threads of every CTA that the
// named barrier 1 joins; `target` = number of this barrier (1, 2, ..) x gridDim.x, the counter only ever grows
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int target, int tid, int nthr) {
    __threadfence();                                  // this thread's global writes, before the CTA-level hand-over
    named_bar_sync(1, nthr);
    if (tid == 0) {
        atomicAdd(bar, 1u);
        unsigned int v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v >= target) break;
            __nanosleep(64);
        }
        __threadfence();
    }
    named_bar_sync(1, nthr);
}

constexpr int kDistBits = 20;           // distances are < 2^20 (d <= 2048, |a - b| <= 255)

// k-th smallest (k >= 1) of the m values in v[] (shared memory; values >= 2^kDistBits = no value), by ONE warp, no
// barriers: the answer is built bit by bit from the top - the largest x with count(v < x) < k - starting at the top bit
// of the largest value present.  One pass over v per bit, eight independent loads in flight per lane (~1.5 us for
// m ~ 1000).  Fewer than k values: 2^kDistBits - 1 (no bound).
template <typename T, int SHIFT>
__device__ __forceinline__ unsigned int warp_kth_smallest_t(const T *v, int m, int k, int lane) {
    unsigned int vmax = 0u;
    int valid = 0;
#pragma unroll 8
    for (int i = lane; i < m; i += 32) {
        const unsigned int x = (unsigned int)(v[i] >> SHIFT);
        if (x < (1u << kDistBits)) {
            vmax = max(vmax, x);
            ++valid;
        }
    }
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    valid = __reduce_add_sync(0xffffffffu, valid);
    if (valid < k) return (1u << kDistBits) - 1u;
    unsigned int ans = 0u;
    for (int b = 31 - __clz(vmax | 1u); b >= 0; --b) {
        const unsigned int cand = ans | (1u << b);
        int c = 0;
#pragma unroll 8
        for (int i = lane; i < m; i += 32) c += ((unsigned int)(v[i] >> SHIFT) < cand) ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c < k) ans = cand;
    }
    return ans;
}

__device__ __forceinline__ unsigned int warp_kth_smallest(const unsigned int *v, int m, int k, int lane) {
    return warp_kth_smallest_t<unsigned int, 0>(v, m, k, lane);
}

// The same for the distance part of 64-bit keys (dist << 40 | position).
__device__ __forceinline__ unsigned int warp_kth_smallest_dist(const unsigned long long *keys, int m, int k, int lane) {
    return warp_kth_smallest_t<unsigned long long, kIdBits>(keys, m, k, lane);
}

constexpr int kFastSortCap = 1024;

// Query qi's candidate list -> its k best, by the consumer threads of one CTA.  Fast path (warp 0 alone, no barriers
// after the load): the k-th smallest DISTANCE d_k of the list, then only the keys with distance <= d_k (k plus the
// ties at d_k: a few dozen) are compacted and sorted.  Lists with more than kFastSortCap such keys (massive ties) take
// the general selection (select_smallest, all threads).
// keys: shared memory for cmax + kFastSortCap keys.
template <typename Sync>
__device__ __forceinline__ void select_query_fast(const SelectParams &p, long long qi, unsigned long long *keys, int tid,
                                                  int nthr, Sync sync) {
    const int m = __ldcg(&p.cnt[qi]);
    if (m > p.cmax) {
        if (tid == 0) p.qflags[qi] = 1;
        return;
    }
    if (tid == 0) p.qflags[qi] = 0;
#pragma unroll 8
    for (int i = tid; i < m; i += nthr) keys[i] = __ldcg(&p.cand[qi * p.cmax + i]);     // independent loads in flight
    sync();
    const int lane = tid & 31;
    unsigned long long *out = keys + p.cmax;
    __shared__ int s_c;
    int c = m;                                   // keys that can still be among the k best
    if (tid < 32) {
        if (m > p.k) {
            const unsigned int dk = warp_kth_smallest_dist(keys, m, p.k, lane);
            c = 0;
            for (int i0 = 0; i0 < m; i0 += 32) {
                const int i = i0 + lane;
                const unsigned long long key = (i < m) ? keys[i] : kKeyMax;
                const bool take = i < m && (unsigned int)(key >> kIdBits) <= dk;
                const unsigned int vote = __ballot_sync(0xffffffffu, take);
                const int slot = c + __popc(vote & ((1u << lane) - 1u));
                if (take && slot < kFastSortCap) out[slot] = key;
                c += __popc(vote);
            }
        } else {
            for (int i = lane; i < m; i += 32) out[i] = keys[i];
        }
        if (c <= kFastSortCap) {
            int Pn = 32;
            while (Pn < c) Pn <<= 1;
            for (int i = c + lane; i < Pn; i += 32) out[i] = kKeyMax;
            __syncwarp();
            warp_sort(out, Pn, lane);
            for (int i = lane; i < p.k; i += 32) {
                const unsigned long long key = (i < c) ? out[i] : kKeyMax;
                const long long o = qi * p.k + i;
                if (p.key_out) {
                    p.key_out[o] = (key == kKeyMax) ? kKeyMax : key + (unsigned long long)p.id_base;
                } else if (key == kKeyMax) {
                    p.dist[o] = FLT_MAX;
                    p.ids[o] = -1;
                } else {
                    p.dist[o] = (float)(unsigned int)(key >> kIdBits);
                    p.ids[o] = (long long)(key & kIdMask) + p.id_base;
                }
            }
        }
        if (lane == 0) s_c = c;
    }
    sync();
    if (s_c > kFastSortCap) select_query(p, qi, keys, tid, nthr, sync);     // rare: reloads the list, all threads
}

// acc + sum_i |a.byte[i] - b.byte[i]|, kept in program order (volatile): the word-major order below puts TQ x TD
// independent accumulators between two updates of the same one, so that a single warp keeps the ALU pipe busy on its own
__device__ __forceinline__ unsigned int sad4v(unsigned int a, unsigned int b, unsigned int acc) {
    unsigned int r;
    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
}

// distances of TQ queries against lane's vector of TD staged groups.  Shared-memory wavefronts are what this kernel
// runs out of first with TD = 1 and 8 or more queries (a 16-byte broadcast load costs 2, a 512-byte row of a group 4):
// with TD = 2 every query load serves two vectors.
// VAR (tuning): 0 = program order kept (volatile), query-major; 1 = order left to the compiler; 2 = volatile, group-major
template <int TQ, int TD, int CT, int VAR>
__device__ __forceinline__ void sad_groups(const uint4 *const (&st)[TD], const uint4 *qs4, int C_, int lane,
                                           unsigned int (&acc)[TQ][TD]) {
    const int C = CT ? CT : C_;
#pragma unroll
    for (int a = 0; a < TQ; ++a)
#pragma unroll
        for (int b = 0; b < TD; ++b) acc[a][b] = 0u;
    // unrolled only as far as the instruction cache likes it: a fully unrolled 30-chunk body of 64 SADs per chunk
    // stalled on instruction fetch (ncu: no_instruction 1.2 per issue)
#pragma unroll(CT ? (TQ * TD <= 8 ? CT : (TQ * TD <= 16 ? 6 : 3)) : 2)
    for (int c = 0; c < C; ++c) {
        uint4 qv[TQ], dv[TD];
#pragma unroll
        for (int b = 0; b < TD; ++b) dv[b] = st[b][c * 32 + lane];
#pragma unroll
        for (int a = 0; a < TQ; ++a) qv[a] = qs4[a * C + c];
#define DCTD_WORD(W)                                                                                       \
        if (VAR == 2) {                                                                                    \
            _Pragma("unroll") for (int b = 0; b < TD; ++b)                                                 \
                _Pragma("unroll") for (int a = 0; a < TQ; ++a) acc[a][b] = sad4v(qv[a].W, dv[b].W, acc[a][b]); \
        } else {                                                                                           \
            _Pragma("unroll") for (int a = 0; a < TQ; ++a)                                                 \
                _Pragma("unroll") for (int b = 0; b < TD; ++b)                                             \
                    acc[a][b] = VAR == 1 ? sad4(qv[a].W, dv[b].W, acc[a][b]) : sad4v(qv[a].W, dv[b].W, acc[a][b]); \
        }
        DCTD_WORD(x)
        DCTD_WORD(y)
        DCTD_WORD(z)
        DCTD_WORD(w)
#undef DCTD_WORD
    }
}

// TQ queries per pass, NW consumer warps, TD groups per warp and turn; CT = compile-time chunk count (0: runtime)
template <int TQ, int NW, int TD, int CT, int VAR = 0>
__global__ void __launch_bounds__((NW + 1) * 32, 1) l1_stream_fused_kernel(const StreamParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const ScanParams &p = P.sp;
    const int C = chunks_of(p.d);
    const int dpad = C * 16;
    const unsigned int pbytes = 32u * (unsigned int)dpad;                                // bytes per group = per slot
    const unsigned int S = (unsigned int)P.stages;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);            // [kStreamMaxSlots]
    unsigned long long *empty = full + kStreamMaxSlots;                                  // [kStreamMaxSlots]
    long long *slot_g = reinterpret_cast<long long *>(empty + kStreamMaxSlots);          // group staged in each slot
    volatile unsigned int *issued = reinterpret_cast<volatile unsigned int *>(slot_g + kStreamMaxSlots);   // slots issued
    unsigned char *qs = smem + kStreamHeader;                                            // [TQ][dpad]
    unsigned int *kscr = reinterpret_cast<unsigned int *>(qs + (size_t)TQ * dpad);       // [kscr_n] minima of one query
    unsigned char *ring = reinterpret_cast<unsigned char *>(kscr + P.kscr_n);            // [S][pbytes]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NC = NW * 32;                                                          // consumer threads
    constexpr unsigned int TS = TD;                                                      // slots per turn
#ifdef DCTD_TUNING
    // phase stamps of CTA 0 (globaltimer, ns) behind the barrier counter and the bounds: dctd_l1_stream_stamps
    unsigned long long *stamps = reinterpret_cast<unsigned long long *>(P.bar + 32);
    int n_stamp = 0;
#define DCTD_STAMP()                                                                        \
    do {                                                                                    \
        if (blockIdx.x == 0 && tid == 0) {                                                  \
            unsigned long long t_;                                                          \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
            stamps[n_stamp++] = t_;                                                         \
        }                                                                                   \
    } while (0)
#else
#define DCTD_STAMP() do {} while (0)
#endif
    const long long G = gridDim.x, cta = blockIdx.x;
    const bool own_bound = p.thr_dist == nullptr;
    const long long n_groups = (p.n + 31) / 32;
    const long long sg = own_bound ? (n_groups + p.gstride - 1) / p.gstride : 0;         // sampled groups
    const long long mineA = cta < sg ? (sg - cta + G - 1) / G : 0;                       // this CTA's sampled groups
    const unsigned int turnsA = (unsigned int)((mineA + TD - 1) / TD);
    // A slot serves different warps in successive rounds unless S is a multiple of the slots all warps take per round.
    // A wait on the parity of a round that has not been armed yet would pass at once (it names the PREVIOUS phase), so
    // in that case consumers first wait until the producer has issued their slots.
    const bool wait_issued = (S % (NW * TS)) != 0;

    if (tid == 0) {
        for (unsigned int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        *issued = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_queries<1, TQ>(p, qs, 0, dpad);
    __syncthreads();

    if (warp == NW) {
        // ---- producer: one lane, one running slot counter, no divisions in the loop.  A turn = TD groups.
        //      Phase A: this CTA's static share of the
        //      sampled groups.  Phase B: groups handed out dynamically, kStreamChunk turns per grab of a global counter
        //      (SMs do not stream at the same rate: a static split leaves the fast ones idle for ~10 % of the call). ----
        if (lane == 0) {
            unsigned int t = 0, s = 0, par = 0;                      // slots issued, slot = t % S, parity of round t / S
            auto issue_turn = [&](const long long (&g)[TD]) {        // g[b] < 0: no data (-1 end marker, -2 absent)
#pragma unroll
                for (int b = 0; b < TD; ++b) {
                    if (t >= S) mbar_wait_hint(&empty[s], par ^ 1u, 2000u);
                    slot_g[s] = g[b];
                    if (g[b] >= 0) {
                        mbar_expect_tx(&full[s], pbytes);
                        bulk_g2s(ring + (size_t)s * pbytes, p.packed + g[b] * C * 32, pbytes, &full[s]);
                    } else {
                        mbar_arrive(&full[s]);
                    }
                    ++t;
                    if (++s == S) { s = 0; par ^= 1u; }
                }
                if (wait_issued) {
                    __threadfence_block();
                    *issued = t;
                }
            };
            long long g[TD];
            for (long long i = 0; i < mineA; i += TD) {
#pragma unroll
                for (int b = 0; b < TD; ++b) g[b] = (i + b < mineA) ? (cta + (i + b) * G) * p.gstride : -2;
                issue_turn(g);
            }
            unsigned long long *next = reinterpret_cast<unsigned long long *>(P.bar + 2);
            constexpr unsigned long long kGrab = (unsigned long long)kStreamGrab;
            long long cur = (long long)atomicAdd(next, kGrab);
            while (cur < n_groups) {
                const long long nxt = (long long)atomicAdd(next, kGrab);     // one grab ahead: its latency is hidden
                for (long long g0 = cur; g0 < min(n_groups, cur + (long long)kGrab); g0 += TD) {
#pragma unroll
                    for (int b = 0; b < TD; ++b) g[b] = (g0 + b < n_groups) ? g0 + b : -2;
                    issue_turn(g);
                }
                cur = nxt;
            }
#pragma unroll
            for (int b = 0; b < TD; ++b) g[b] = -1;
            for (int w = 0; w < NW; ++w) issue_turn(g);              // every consumer warp meets exactly one end marker
        }
        return;
    }

    // ---- consumers: warp w takes turns w, w + NW, ... (running index over both phases) ----
    const uint4 *qs4 = reinterpret_cast<const uint4 *>(qs);
    unsigned int tdist[TQ];
    unsigned int barrier_no = 0;
    unsigned int cs = (unsigned int)warp * TS % S, cpar = ((unsigned int)warp * TS / S) & 1u;   // first slot of the next turn
    unsigned int cturn = (unsigned int)warp;
    // distances of this warp's next turn; g[b] = group of accumulator column b (< 0: none); false at the end marker
    auto consume = [&](unsigned int (&acc)[TQ][TD], long long (&g)[TD]) -> bool {
        if (wait_issued) {
            while (*issued < (cturn + 1u) * TS) __nanosleep(20);
        }
        const uint4 *st[TD];
        unsigned int s2 = cs, p2 = cpar;
#pragma unroll
        for (int b = 0; b < TD; ++b) {
            mbar_wait(&full[s2], p2);
            st[b] = reinterpret_cast<const uint4 *>(ring + (size_t)s2 * pbytes);
            g[b] = slot_g[s2];
            if (++s2 == S) { s2 = 0; p2 ^= 1u; }
        }
        const bool live = g[0] != -1;
        if (live) {
            sad_groups<TQ, TD, CT, VAR>(st, qs4, C, lane, acc);
            __syncwarp();
            if (lane == 0) {
                unsigned int s3 = cs;
#pragma unroll
                for (int b = 0; b < TD; ++b) {
                    mbar_arrive(&empty[s3]);
                    if (++s3 == S) s3 = 0;
                }
            }
        }
        // on to this warp's next turn: skip the slots of the other warps' turns
        cs = s2 + (NW - 1) * TS;
        cpar = p2;
        while (cs >= S) { cs -= S; cpar ^= 1u; }
        cturn += NW;
        return live;
    };

    DCTD_STAMP();   // 0: start
    if (own_bound) {
        // ---- phase A: lane minima over the sampled groups ----
        unsigned int best[TQ];
#pragma unroll
        for (int a = 0; a < TQ; ++a) best[a] = 0xffffffffu;
        for (unsigned int i = warp; i < turnsA; i += NW) {
            unsigned int acc[TQ][TD];
            long long g[TD];
            consume(acc, g);
#pragma unroll
            for (int b = 0; b < TD; ++b) {
                if (g[b] >= 0 && g[b] * 32 + lane < p.n) {        // padding lanes of the last group are not vectors
#pragma unroll
                    for (int a = 0; a < TQ; ++a) best[a] = min(best[a], acc[a][b]);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < TQ; ++a) {
            unsigned int v = best[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0 && a < p.nq) P.mins[(long long)a * kMinSlots + cta * NW + warp] = v;
        }
        DCTD_STAMP();   // 1: sample done
        grid_barrier(P.bar, (++barrier_no) * (unsigned int)G, tid, NC);
        DCTD_STAMP();   // 2: barrier 1
        // ---- CTA q: k-th smallest of query q's minima -> bound ----
        if (cta < p.nq) {
#pragma unroll 8
            for (int i = tid; i < P.m; i += NC) kscr[i] = __ldcg(&P.mins[cta * kMinSlots + i]);
            named_bar_sync(1, NC);
            if (warp == 0) {
                const unsigned int kth = warp_kth_smallest(kscr, P.m, p.k, lane);
                if (lane == 0) P.thr_out[cta] = (p.k <= P.m) ? kth : 0xffffffffu;
            }
        }
        DCTD_STAMP();   // 3: k-th
        grid_barrier(P.bar, (++barrier_no) * (unsigned int)G, tid, NC);
        DCTD_STAMP();   // 4: barrier 2
    }
    {
        const unsigned int *thr = own_bound ? P.thr_out : p.thr_dist;
#pragma unroll
        for (int a = 0; a < TQ; ++a) tdist[a] = (a < p.nq) ? __ldcg(&thr[a]) : 0u;
    }
    // ---- phase B: every group; append what lies within the bound ----
    for (;;) {
        unsigned int acc[TQ][TD];
        long long g[TD];
        if (!consume(acc, g)) break;
        unsigned int mask = 0u;
#pragma unroll
        for (int a = 0; a < TQ; ++a)
#pragma unroll
            for (int b = 0; b < TD; ++b) mask |= (acc[a][b] <= tdist[a]) ? (1u << (a * TD + b)) : 0u;
        const unsigned int any = __reduce_or_sync(0xffffffffu, mask);
        if (!any) continue;
#pragma unroll
        for (int a = 0; a < TQ; ++a) {
#pragma unroll
            for (int b = 0; b < TD; ++b) {
                if (any & (1u << (a * TD + b))) {          // warp-uniform
                    const long long id = g[b] * 32 + lane;
                    const bool pass = g[b] >= 0 && id < p.n && a < p.nq && acc[a][b] <= tdist[a];
                    append_candidates(p, a, pass, ((unsigned long long)acc[a][b] << kIdBits) | (unsigned long long)id, lane);
                }
            }
        }
    }
    DCTD_STAMP();   // 5: stream done (this CTA)
    grid_barrier(P.bar, (++barrier_no) * (unsigned int)G, tid, NC);
    DCTD_STAMP();   // 6: barrier 3
    // ---- CTA q: exact selection of query q (the ring is idle: it becomes the sort buffer) ----
    if (cta < p.nq)
        select_query_fast(P.se, cta, reinterpret_cast<unsigned long long *>(ring), tid, NC, [] { named_bar_sync(1, NC); });
    DCTD_STAMP();   // 7: select
}
