// Warp-specialised fingerprint kernel (included by fingerprint.cu inside its anonymous namespace).
//
// Same arithmetic as fp_kernel (see the header of fingerprint.cu), different machine mapping: one
// persistent CTA per SM whose warps have fixed roles and only meet at mbarriers, so the HBM stream
// never pauses for the per-domain epilogue:
//
//   warp 0            producer   pulls items from the atomic work queue (fat 128-byte item records,
//                                fetched one item ahead), computes the per-row cosine basis and
//                                moves the rows HBM -> shared memory with TMA bulk copies
//                                (cp.async.bulk, SASS UBLKCP) into a ring of ~20 KB stages;
//   warps 1..NCW      consumers  one thread per float4 column group; pivot subtraction and the
//                                K (2K with a riding global fingerprint) projections as packed
//                                FFMA2 from shared memory, float32 partial sums flushed into
//                                float64 every 8 rows; at the end of an item the sums are handed
//                                to the finishers through a double-buffered shared array;
//   last 8 warps      finishers  everything after pass 1 (split-domain / rider bookkeeping, the
//                                length-n inverse + column min-max, pass 2a against a cosine
//                                table in shared memory, pass 2b, row min-max, int8 output).
//
// Bytes in flight are set by the ring (6 stages x 20 KB per SM at D = 1280), not by registers, and
// the epilogue of item i overlaps the stream of items i+1, i+2.
#pragma once

constexpr int kWsFinWarps = 8;
constexpr int kWsFinThreads = kWsFinWarps * 32;
constexpr int kWsDescSlots = 16;     // item records in flight between producer and finishers
#ifndef WS_MAX_STAGES
#define WS_MAX_STAGES 10
#endif
constexpr int kWsMaxStages = WS_MAX_STAGES;
constexpr int kWsMaxUBufs = 4;        // hand-over buffers between consumers and finishers (WsCfg::NB of them are used)
constexpr int kWsBasisRows = 256;    // basis ring (rows); >= 32 + stages x rows per stage + 32, power of two

// stage flags
constexpr int kStLast = 1;           // last stage of its item: hand the sums to the finishers
constexpr int kStDual = 2;           // rows come from two windows and are averaged (embedding.py:186)
constexpr int kStPivotOnly = 4;      // the stage holds only the item's pivot row
constexpr int kStPivotInline = 8;    // row 0 of the stage is also the item's pivot row
constexpr int kStExit = 16;          // no more items
constexpr int kStRider = 32;         // the item carries its protein's global fingerprint

// words of a fat item record (WsItem, 32 x int32 = 128 bytes, one word per producer lane)
enum {
    kWDom = 0, kWLayer, kWR0, kWR1, kWL, kWFlags, kWSplit, kWNsplit, kWSlabBase, kWCounter,
    kWRiderDom, kWRiderSlab, kWRiderNsplit, kWRiderSlabBase, kWRiderCounter, kWLg,
    kWPivSrcA, kWPivRowA, kWPivSrcB, kWPivRowB, kWNRuns, kWPieceAbs,
    kWRunSrcA, kWRunRowA, kWRunSrcB, kWRunRowB, kWRunRows, kWRunL0, kWRunG0, kWSpare0, kWSpare1, kWSpare2
};
constexpr int kWfRider = 1, kWfPivotInline = 2, kWfSplit = 4;

template <int K, int DC, bool RIDER>
struct WsCfg {
    static constexpr int KS = RIDER ? 2 * K : K;
    static constexpr int NCT = DC / 4;                  // consumer threads: one per float4 column group
    static constexpr int NCW = (NCT + 31) / 32;         // the last consumer warp may be partly idle (D = 320 / 480)
#ifndef WS_STAGE_ROW_BYTES
#define WS_STAGE_ROW_BYTES 5120
#endif
    static constexpr int R0 = WS_STAGE_ROW_BYTES / DC / 2 * 2;
    static constexpr int R = R0 < 2 ? 2 : (R0 > 8 ? 8 : R0);   // rows per stage (even)
    static constexpr int T = 32 * (1 + NCW + kWsFinWarps);
    static constexpr int STAGE_BYTES = R * DC * 4;
#ifndef WS_FLUSH_ROWS
#define WS_FLUSH_ROWS 8
#endif
#ifndef WS_FLUSH_ROWS_RIDER
#define WS_FLUSH_ROWS_RIDER WS_FLUSH_ROWS
#endif
#ifndef WS_UBUFS
#define WS_UBUFS 2
#endif
#ifndef WS_UBUFS_RIDER
#define WS_UBUFS_RIDER WS_UBUFS
#endif
    static constexpr int NB = RIDER ? WS_UBUFS_RIDER : WS_UBUFS;      // hand-over buffers (each K * D doubles)
    static_assert(NB >= 2 && NB <= kWsMaxUBufs, "hand-over buffers");
    static constexpr int FLUSH_ROWS = RIDER ? WS_FLUSH_ROWS_RIDER : WS_FLUSH_ROWS;   // float32 chain length of pass 1
    static constexpr int FLUSH_EVERY = (FLUSH_ROWS / R) < 1 ? 1 : (FLUSH_ROWS / R);
    static_assert(DC % 16 == 0, "float4 steps over the D/4 folded columns of pass 2a");
    static_assert(R % 2 == 0, "dual stages hold R/2 rows of each window");
};

// per-role cycle counters (DCTD_FP_TIMING builds; scripts/fp_phases.py): one thread per role accumulates locally and
// adds to Params::timing at exit.  Slots: 0 producer waiting for a free stage, 1 producer total, 2 consumer waiting
// for a full stage, 3 consumer waiting for a free hand-over buffer, 4 consumer total, 5 finisher idle, 6 finisher
// total, 7 finisher stage 1 (+ split / rider bookkeeping), 8 pass 2a, 9 row min-max + output (warp 0), 10 items, 11 reduce, 12 pass 2b.
#ifdef DCTD_FP_TIMING
#define WS_T0() long long _wt = clock64()
#define WS_ACC(var) do { const long long _n = clock64(); (var) += _n - _wt; _wt = _n; } while (0)
#define WS_PUT(slot, var) atomicAdd((unsigned long long *)&p.timing[slot], (unsigned long long)(var))
#else
#define WS_T0() do {} while (0)
#define WS_ACC(var) do {} while (0)
#define WS_PUT(slot, var) do {} while (0)
#endif

// wait flavour per site (0 producer / free stage, 1 consumer / full stage, 2 consumer / hand-over buffer,
// 3 finisher / idle), selectable at build time for A/B runs: -DWS_WAIT_MODE=0 spin, 1 suspend hint, 2 sleep
#ifndef WS_WAIT_MODE
#define WS_WAIT_MODE 1
#endif
#if WS_WAIT_MODE == 0
#define WS_WAIT(bar, par, site) mbar_wait(bar, par)
#elif WS_WAIT_MODE == 1
#define WS_WAIT(bar, par, site) mbar_wait_hint(bar, par, 20000u)
#else
#define WS_WAIT(bar, par, site) do { if ((site) == 1 || (site) == 2) mbar_wait(bar, par); else mbar_wait_sleep(bar, par, (site) == 3 ? 256u : 96u); } while (0)
#endif

// read-only table lookup at a shared-memory byte address (the table never changes after the kernel prologue, so the
// compiler may schedule these freely)
__device__ __forceinline__ float lds_f32(unsigned int addr) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_pk4(unsigned int addr, pk2 &lo, pk2 &hi) {
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr));
}

template <int K, int DC, bool RIDER>
__global__ void __launch_bounds__(WsCfg<K, DC, RIDER>::T, 1) fp_ws_kernel(const Params p) {
    using Cfg = WsCfg<K, DC, RIDER>;
    using namespace dctd::tma;
    constexpr int N = K + 1, KS = Cfg::KS, R = Cfg::R, NCW = Cfg::NCW, NFT = kWsFinThreads;
    constexpr int D = DC, half = DC / 2, quarter = DC / 4;
    extern __shared__ __align__(16) unsigned char smem[];
    const WsLayout wl = p.wl;
    const int NST = wl.nst;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + wl.off_bars);   // [NST]
    unsigned long long *empty = full + kWsMaxStages;                                        // [NST]
    unsigned long long *u_full = empty + kWsMaxStages;                                      // [NB]
    unsigned long long *u_empty = u_full + kWsMaxUBufs;                                     // [NB]
    int4 *metas = reinterpret_cast<int4 *>(smem + wl.off_meta);            // [NST] {nrows, flags, desc slot | basis slot << 8, rider slab}
    float *basis = reinterpret_cast<float *>(smem + wl.off_basis);         // [kWsBasisRows][KS] ring
    int *desc = reinterpret_cast<int *>(smem + wl.off_desc);               // [kWsDescSlots][32]
    unsigned char *ring = smem + wl.off_ring;                              // [NST][STAGE_BYTES]
    double *ubuf = reinterpret_cast<double *>(smem + wl.off_u);            // [NB][K][D]
    float *Ye = reinterpret_cast<float *>(smem + wl.off_ye);               // [2][N][D/4]  EE | EO (see stage1)
    float *Yo = reinterpret_cast<float *>(smem + wl.off_yo);               // [2][N][D/4]  OD | OR
    double *Fs = reinterpret_cast<double *>(smem + wl.off_f);              // pass-2a partial sums | Fr [N][nk] | Z [N][m]
    float *TT = reinterpret_cast<float *>(smem + wl.off_tt);               // [even i | odd i] halves of cos(pi (i mod 4D) / 2D), i < 4D + 8m
    double *Tm = reinterpret_cast<double *>(smem + wl.off_tm);             // cos(pi i / 2m), i < 4m, skewed (see below)
    double *Mj = reinterpret_cast<double *>(smem + wl.off_mj);             // [N][K] cos(pi (2j+1) k / 2n)
    __shared__ int u_slot[kWsMaxUBufs];
    __shared__ int s_flag, s_last, s_rlast;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = p.m, nk = m - 1;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        for (int b = 0; b < Cfg::NB; ++b) { mbar_init(&u_full[b], NCW); mbar_init(&u_empty[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // cos(pi i / 2D) split by the parity of i: odd k only ever look up odd i = (2d+1) k, even k even i, and inside
    // one parity class lanes with consecutive k are an odd number of entries apart (conflict-free)
    for (int i = tid; i < 4 * D + 8 * m; i += blockDim.x) TT[(i & 1) * (2 * D + 4 * m) + (i >> 1)] = __ldg(p.table + i);
    // cos(pi i / 2m) for pass 2b, entry i stored at i + (i >> 4): the lanes of a warp look up (2c + 1) k for 16 consecutive
    // c, i.e. at a stride of 2k entries, and for k = 8, 16, ... all of them would meet in one bank of a plain table
    for (int i = tid; i < 4 * m; i += blockDim.x) Tm[i + (i >> 4)] = cospi((double)i / (2.0 * m));
    if (tid == 0) s_flag = 0;
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int j = i / K, k = i % K + 1;
        Mj[i] = cospi((double)((2 * j + 1) * k) / (2.0 * N));
    }
    __syncthreads();

    if (warp == 0) {
        // =============================== producer ===============================
        auto fetch = [&]() -> int {
            int v = 0;
            if (lane == 0) v = atomicAdd(&p.counters[0], 1);
            return __shfl_sync(0xffffffffu, v, 0);
        };
        int s = 0;
        unsigned int eph = 1;            // parity of the "previous" phase: a fresh barrier passes at once
        int seq = 0;
        int rowseq = 0;                  // data rows emitted so far: position in the basis ring
        long long t_wait = 0, t_busy = 0;
        (void)t_wait; (void)t_busy;
        WS_T0();
        // emits one stage: nr rows from pa (and pb); bslot = ring position of the basis of its first row
        auto emit = [&](const float *pa, const float *pb, int nr, int flags, int slot, int bslot, int aux) {
            WS_ACC(t_busy);
            WS_WAIT(&empty[s], eph, 0);
            WS_ACC(t_wait);
            if (lane == 0) {
                metas[s] = make_int4(nr, flags, slot | (bslot << 8), aux);
                const unsigned int rb = (unsigned int)nr * D * 4u;
                unsigned char *dst = ring + (size_t)s * Cfg::STAGE_BYTES;
                mbar_expect_tx(&full[s], (flags & kStExit) ? 0u : (pb ? 2u * rb : rb));
                if (!(flags & kStExit)) {
                    bulk_g2s(dst, pa, rb, &full[s]);
                    if (pb) bulk_g2s(dst + (R / 2) * D * 4, pb, rb, &full[s]);
                }
            }
            if (++s == NST) { s = 0; eph ^= 1u; }
        };
        int it_cur = fetch();
        int it_nxt = fetch();
        int rec = (it_cur < p.n_items) ? __ldg(p.wsitems + (size_t)it_cur * 32 + lane) : 0;
        for (;;) {
            if (it_cur >= p.n_items) {
                emit(nullptr, nullptr, 0, kStExit, 0, 0, 0);
                break;
            }
            const int my = rec;
            // one item ahead: the next record is in flight while this item streams
            const int it_after = fetch();
            int rec_n = 0;
            if (it_nxt < p.n_items) rec_n = __ldg(p.wsitems + (size_t)it_nxt * 32 + lane);
#define WS_W(i) __shfl_sync(0xffffffffu, my, (i))
            const int slot = seq & (kWsDescSlots - 1);
            desc[slot * 32 + lane] = my;
            const int iflags = WS_W(kWFlags), layer = WS_W(kWLayer);
            const int L = WS_W(kWL), Lg = WS_W(kWLg);
            const int r0 = WS_W(kWR0), r1 = WS_W(kWR1);
            const bool rider = RIDER && (iflags & kWfRider);
            const int rider_slab = WS_W(kWRiderSlab);
            const float *const *src = p.src + (int64_t)layer * p.n_src;
            const int base_flags = rider ? kStRider : 0;
            __syncwarp();       // the record in desc[] is complete before lane 0 publishes the item's first stage
            if (!(iflags & kWfPivotInline)) {
                const int sa = WS_W(kWPivSrcA), ra = WS_W(kWPivRowA), sb = WS_W(kWPivSrcB), rb = WS_W(kWPivRowB);
                const float *pa = src[sa] + (int64_t)ra * D;
                const float *pb = (sb >= 0) ? src[sb] + (int64_t)rb * D : nullptr;
                emit(pa, pb, 1, base_flags | kStPivotOnly | (pb ? kStDual : 0), slot, 0, rider_slab);
            }
            const int n_runs = WS_W(kWNRuns), piece_abs = WS_W(kWPieceAbs);
            bool first_stage = true;
            for (int run = 0; run < n_runs; ++run) {
                int sa, ra, sb, rb, nrows, l0, g0;
                if (run == 0) {
                    sa = WS_W(kWRunSrcA); ra = WS_W(kWRunRowA); sb = WS_W(kWRunSrcB); rb = WS_W(kWRunRowB);
                    nrows = WS_W(kWRunRows); l0 = WS_W(kWRunL0); g0 = WS_W(kWRunG0);
                } else {
                    const Piece pc = p.pieces[piece_abs + run];       // later pieces of the item, clipped to [r0, r1)
                    const int a = max(pc.l0, r0), b = min(pc.l0 + pc.nrows, r1);
                    sa = pc.src_a; ra = pc.row_a + (a - pc.l0); sb = pc.src_b; rb = pc.row_b + (a - pc.l0);
                    nrows = b - a; l0 = a; g0 = pc.g0 + (a - pc.l0);
                }
                const float *pa = src[sa] + (int64_t)ra * D;
                const float *pb = (sb >= 0) ? src[sb] + (int64_t)rb * D : nullptr;
                const int cap = pb ? R / 2 : R;
                int have = 0;                   // rows of this run whose basis is in the ring
                for (int done = 0; done < nrows; done += cap) {
                    const int nr = min(cap, nrows - done);
                    if (done + nr > have) {
                        // basis of the next 32 rows, one row per lane: cos(pi (2l+1) k / 2L), k = 1..K, by the
                        // Chebyshev recurrence from one cospi (float64); the ring slots written here belonged to
                        // rows >= kWsBasisRows - 32 back, whose stages were handed back long ago
                        const int row = have + lane;
                        if (row < nrows) {
                            float *bp = basis + (size_t)((rowseq + row) & (kWsBasisRows - 1)) * KS;
#pragma unroll
                            for (int which = 0; which < (RIDER ? 2 : 1); ++which) {
                                if (which == 1 && !rider) {
#pragma unroll
                                    for (int k = 0; k < K; ++k) bp[K + k] = 0.f;
                                    break;
                                }
                                const int idx = (which == 0 ? l0 : g0) + row;
                                const int len = which == 0 ? L : Lg;
                                const double c1 = cospi((double)(2 * idx + 1) / (2.0 * len));
                                double ckm2 = 1.0, ckm1 = c1;
                                bp[which * K] = (float)c1;
#pragma unroll
                                for (int k = 2; k <= K; ++k) {
                                    const double ck = 2.0 * c1 * ckm1 - ckm2;
                                    bp[which * K + k - 1] = (float)ck;
                                    ckm2 = ckm1;
                                    ckm1 = ck;
                                }
                            }
                        }
                        have += 32;
                        __syncwarp();
                    }
                    const bool last = (run == n_runs - 1) && (done + nr == nrows);
                    int fl = base_flags | (pb ? kStDual : 0) | (last ? kStLast : 0);
                    if (first_stage && (iflags & kWfPivotInline)) fl |= kStPivotInline;
                    first_stage = false;
                    emit(pa + (int64_t)done * D, pb ? pb + (int64_t)done * D : nullptr, nr, fl, slot,
                         (rowseq + done) & (kWsBasisRows - 1), rider_slab);
                }
                rowseq += nrows;
            }
#undef WS_W
            ++seq;
            it_cur = it_nxt;
            it_nxt = it_after;
            rec = rec_n;
        }
        WS_ACC(t_busy);
        if (lane == 0) { WS_PUT(0, t_wait); WS_PUT(1, t_wait + t_busy); WS_PUT(10, seq); }
    } else if (warp <= NCW) {
        // =============================== consumers ===============================
        const int ctid = tid - 32;
        // D not a multiple of 128: the lanes of the last consumer warp beyond the last column group run along with
        // column group 0 (same loads, no stores), so that every warp-level step stays uniform
        const bool c_active = ctid < Cfg::NCT;
        const int col = (c_active ? ctid : 0) * 4;
        const unsigned int ring_u32 = smem_u32(ring) + (unsigned int)col * 4u;
        int s = 0, ub = 0, nflush = 0;
        unsigned int fph = 0, uph = 1;
        double acc[KS][4];
        pk2 a0[KS], a1[KS];
        const pk2 zero = pk(0.f, 0.f), hf = pk(0.5f, 0.5f), neg = pk(-1.f, -1.f);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            a0[k] = zero; a1[k] = zero;
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[k][v] = 0.0;
        }
        pk2 npiv0 = zero, npiv1 = zero;
#ifdef DCTD_DEBUG_U
        pk2 chk0 = zero, chk1 = zero, chkx0 = zero, chkx1 = zero;
#endif
        long long t_full = 0, t_u = 0, t_busy = 0;
        (void)t_full; (void)t_u; (void)t_busy;
        WS_T0();
        for (;;) {
            WS_ACC(t_busy);
            WS_WAIT(&full[s], fph, 1);
            WS_ACC(t_full);
            const int4 mt = metas[s];
            const int nrows = mt.x, flags = mt.y;
            if (flags & kStExit) {
                WS_WAIT(&u_empty[ub], uph, 2);
                if (ctid == 0) u_slot[ub] = -1;
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_full[ub]);
                break;
            }
            const unsigned int st = ring_u32 + (unsigned int)s * Cfg::STAGE_BYTES;
            const int bslot = mt.z >> 8;
            // All R row slots of the stage are loaded before the stage's record is inspected: the loads only depend on
            // the stage index, so their latency overlaps that of the record (slots beyond nrows hold stale rows, which
            // nothing below uses).  The next stage index is computed here as well, off the critical path.
            pk2 x0[R], x1[R];
#pragma unroll
            for (int r = 0; r < R; ++r) lds_pk4(st + r * D * 4, x0[r], x1[r]);
            const int s_cur = s;
            if (++s == NST) { s = 0; fph ^= 1u; }
            if (nrows == R && !(flags & (kStDual | kStPivotOnly))) {
                // the common stage: R rows of one source.  Straight-line code: every load is issued before the
                // first use, and the stage goes back to the producer as soon as the loads have been PERFORMED
                // (the basis ring is not part of the stage: its slots are recycled 256 rows later, see the producer)
                // Without a rider all basis values are fetched up front as well (measured: +2 % at D = 1280, +6 % at
                // D = 640); with a rider (twice the projections) that would spill, so they are fetched row by row.
                constexpr int CR = RIDER ? 1 : R;
                pk2 c[CR][KS];
                if constexpr (!RIDER) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        load_basis<KS>(basis + (size_t)((bslot + r) & (kWsBasisRows - 1)) * KS, c[r]);
                }
                if (flags & kStPivotInline) {
                    npiv0 = mul2(x0[0], neg);
                    npiv1 = mul2(x1[0], neg);
                }
                // The pivot subtraction comes BEFORE the stage is handed back: it reads every register the stage's loads
                // write, so the loads have been performed when the release is issued.  (Releasing right after ISSUING
                // the loads let the refill of the slot overtake them under shared-memory pressure: 1 launch in 5 of the
                // protein-shaped batch had a few hundred bytes of one (domain, layer) off - found by scripts/fp_check.py.)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    x0[r] = add2(x0[r], npiv0);
                    x1[r] = add2(x1[r], npiv1);
                }
                // without a rider the stage goes back here; with one (twice the FMAs, basis fetched row by row) handing it
                // back after the FMA loop measured faster (5620 vs 5430 GB/s on the protein-shaped batch)
                if constexpr (!RIDER) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s_cur]);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if constexpr (RIDER) load_basis<KS>(basis + (size_t)((bslot + r) & (kWsBasisRows - 1)) * KS, c[0]);
                    const pk2 t0 = x0[r], t1 = x1[r];
#ifdef DCTD_DEBUG_U
                    chk0 = add2(chk0, t0); chk1 = add2(chk1, t1);
                    chkx0 = add2(chkx0, c[RIDER ? 0 : r][0]); chkx1 = add2(chkx1, c[RIDER ? 0 : r][1]);
#endif
#pragma unroll
                    for (int k = 0; k < KS; ++k) {
                        a0[k] = fma2(t0, c[RIDER ? 0 : r][k], a0[k]);
                        a1[k] = fma2(t1, c[RIDER ? 0 : r][k], a1[k]);
                    }
                }
                if constexpr (RIDER) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s_cur]);
                }
            } else {
                // partial stages, rows averaged from two windows, pivot-only stages
                if (flags & kStDual) {
#pragma unroll
                    for (int r = 0; r < R / 2; ++r) {
                        x0[r] = mul2(add2(x0[r], x0[R / 2 + r]), hf);      // embedding.py:186: (prev + cur) / 2 in float32
                        x1[r] = mul2(add2(x1[r], x1[R / 2 + r]), hf);
                    }
                }
                if (flags & (kStPivotOnly | kStPivotInline)) {
                    npiv0 = mul2(x0[0], neg);
                    npiv1 = mul2(x1[0], neg);
                }
                if (!(flags & kStPivotOnly)) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (r < nrows) {
                            const pk2 t0 = add2(x0[r], npiv0), t1 = add2(x1[r], npiv1);
                            pk2 c[KS];
                            load_basis<KS>(basis + (size_t)((bslot + r) & (kWsBasisRows - 1)) * KS, c);
#ifdef DCTD_DEBUG_U
                            chk0 = add2(chk0, t0); chk1 = add2(chk1, t1);
                            chkx0 = add2(chkx0, c[0]); chkx1 = add2(chkx1, c[1]);
#endif
#pragma unroll
                            for (int k = 0; k < KS; ++k) {
                                a0[k] = fma2(t0, c[k], a0[k]);
                                a1[k] = fma2(t1, c[k], a1[k]);
                            }
                        }
                    }
                }
                // the stage's rows and basis are consumed: give it back to the producer
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s_cur]);
            }
            if (flags & kStPivotOnly) continue;
            if (++nflush == Cfg::FLUSH_EVERY || (flags & kStLast)) {
                nflush = 0;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    float f0, f1, f2, f3;
                    unpk(a0[k], f0, f1);
                    unpk(a1[k], f2, f3);
                    acc[k][0] += (double)f0; acc[k][1] += (double)f1;
                    acc[k][2] += (double)f2; acc[k][3] += (double)f3;
                    a0[k] = zero; a1[k] = zero;
                }
            }
            if (flags & kStLast) {
                // hand the item's sums to the finishers
                WS_ACC(t_busy);
                WS_WAIT(&u_empty[ub], uph, 2);
                WS_ACC(t_u);
                double *u = ubuf + (size_t)ub * K * D;
                if (c_active) {
                    // a thread owns 32 bytes per row; lanes 4..7 of every eight write their second half first so that the
                    // eight 16-byte stores of a quarter warp fall into eight different bank groups
#ifdef WS_OLD_STORE
                    const int hx = 0;
#else
                    const int hx = (lane >> 2) & 1;
#endif
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const double2 lo = make_double2(acc[k][0], acc[k][1]), hi = make_double2(acc[k][2], acc[k][3]);
                        *reinterpret_cast<double2 *>(u + k * D + col + 2 * hx) = hx ? hi : lo;
                        *reinterpret_cast<double2 *>(u + k * D + col + 2 - 2 * hx) = hx ? lo : hi;
                    }
                }
                if constexpr (RIDER) {
                    if ((flags & kStRider) && c_active) {     // the rider's partial sums go straight to its workspace slab
                        double *slab = p.partials + (int64_t)mt.w * (K * D);
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            *reinterpret_cast<double2 *>(slab + k * D + col) = make_double2(acc[K + k][0], acc[K + k][1]);
                            *reinterpret_cast<double2 *>(slab + k * D + col + 2) = make_double2(acc[K + k][2], acc[K + k][3]);
                        }
                        // no fence here: these stores reach the finishers through the u_full mbarrier (release / acquire
                        // at CTA scope); the finisher thread that takes the rider ticket fences at GPU scope before its
                        // atomicAdd, and that release is cumulative over everything ordered before it
                    }
                }
#ifdef DCTD_DEBUG_U
                if (c_active) {
                    const int *dd = desc + (mt.z & 255) * 32;
                    double *dbg = p.debug_u + ((size_t)(dd[kWDom] * p.n_layers + dd[kWLayer]) * 4 + min(dd[kWSplit], 3)) * (K * D);
#pragma unroll
                    for (int k = 0; k < K; ++k)
#pragma unroll
                        for (int v = 0; v < 4; ++v) dbg[k * D + col + v] = acc[k][v];
                    // checksums: plain float32 sums of (x - pivot) per column (slot 2) and of the two basis values (slot 3)
                    float c0, c1, c2, c3;
                    unpk(chk0, c0, c1); unpk(chk1, c2, c3);
                    double *d2 = dbg + (size_t)(2 - min(dd[kWSplit], 3)) * (K * D);
                    if (dd[kWSplit] == 0) { d2[col] = c0; d2[col + 1] = c1; d2[col + 2] = c2; d2[col + 3] = c3; }
                    unpk(chkx0, c0, c1); unpk(chkx1, c2, c3);
                    if (dd[kWSplit] == 0) { d2[D + col] = c0; d2[D + col + 1] = c1; d2[D + col + 2] = c2; d2[D + col + 3] = c3; }
                }
                chk0 = zero; chk1 = zero; chkx0 = zero; chkx1 = zero;
#endif
                if (ctid == 0) u_slot[ub] = mt.z & 255;
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_full[ub]);
                if (++ub == Cfg::NB) { ub = 0; uph ^= 1u; }
#pragma unroll
                for (int k = 0; k < KS; ++k)
#pragma unroll
                    for (int v = 0; v < 4; ++v) acc[k][v] = 0.0;
            }
        }
        WS_ACC(t_busy);
        if (ctid == 0) { WS_PUT(2, t_full); WS_PUT(3, t_u); WS_PUT(4, t_full + t_u + t_busy); }
    } else {
        // =============================== finishers ===============================
        const int ftid = tid - 32 * (1 + NCW);
        const int fwarp = ftid >> 5;
        const int DSo = wl.DSo, DSe = wl.DSe, H = wl.H, HP = wl.H / 2;
        const int TH = 2 * D + 4 * m;                      // entries of one parity half of the cosine table
        constexpr int QQ = quarter / 4;                    // float4 steps over d < D/4
        const unsigned int tt_u32 = smem_u32(TT);
        double *Fp = Fs;                                   // odd k [DSo][N][H] | even k [DSe][N][H]
        double *Fr = Fs + (size_t)(DSo + DSe) * N * H;     // [N][nk]
        double *Z = Fr + (size_t)N * nk;                   // [N][m]
        auto fin_bar = [&]() { named_bar_sync(1, NFT); };

        // length-n inverse + per-column min-max of one column (fingerprint.py:138-140): y' - 0.5
        auto column = [&](const double (&u)[K], double (&yo)[N]) {
            double y[N];
            double mn = INFINITY, mx = -INFINITY;
            bool bad = false;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k < K; ++k) sacc = fma(Mj[j * K + k], u[k], sacc);
                y[j] = sacc;
                bad = bad || !(sacc == sacc);
                mn = fmin(mn, sacc);
                mx = fmax(mx, sacc);
            }
            bad = bad || !(mx > mn);
            if (bad) s_flag = 1;       // constant / non-finite column: the reference yields NaN -> all 0
            const double inv = 1.0 / (mx - mn);
#pragma unroll
            for (int j = 0; j < N; ++j) yo[j] = (y[j] - mn) * inv - 0.5;
        };
        // Folded inputs of pass 2a from the u sums.  With x = pi (2d+1) k / 2D:
        //   cos at column D-1-d   = (-1)^k cos x                      -> e[d] = y'[d] + y'[D-1-d] (even k), o[d] = y'[d] - y'[D-1-d] (odd k), d < D/2
        //   cos at column D/2-1-d = cos(k pi/2 - x) = +cos x, sin x, -cos x, -sin x for k = 0, 1, 2, 3 (mod 4)
        //                                                              -> EE = e[d] + e[D/2-1-d] (k = 0 mod 4), EO = e[d] - e[D/2-1-d] (k = 2 mod 4),
        //                                                                 OD = o[d] with cos x and OR = o[D/2-1-d] with +-sin x (odd k), d < D/4
        // Ye = [EE | EO], Yo = [OD | OR], each [N][D/4]: even k need half the products of the one-fold form.
        auto stage1 = [&](auto getu) {
            for (int d = ftid; d < quarter; d += NFT) {
                const int c[4] = {d, half - 1 - d, half + d, D - 1 - d};
                double y[4][N];            // folded in float64, rounded to float32 once
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double u[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) u[k] = getu(k, c[i]);
                    column(u, y[i]);
                }
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double e0 = y[0][j] + y[3][j], o0 = y[0][j] - y[3][j];     // columns d, D-1-d
                    const double e1 = y[1][j] + y[2][j], o1 = y[1][j] - y[2][j];     // columns D/2-1-d, D/2+d
                    Ye[j * quarter + d] = (float)(e0 + e1);
                    Ye[(N + j) * quarter + d] = (float)(e0 - e1);
                    Yo[j * quarter + d] = (float)o0;
                    Yo[(N + j) * quarter + d] = (float)o1;
                }
            }
        };
        // passes 2a / 2b, row min-max and the int8 output of one (domain, layer) whose Ye / Yo are ready
        long long t_idle = 0, t_s1 = 0, t_2a = 0, t_rest = 0, t_red = 0, t_2b = 0;
        (void)t_idle; (void)t_s1; (void)t_2a; (void)t_rest; (void)t_red; (void)t_2b;
        WS_T0();
        auto finish_rest = [&](int dom_index, int layer) {
            WS_ACC(t_s1);
            // ---- pass 2a: F[j][k] = sum_d y'[j][d] cos(pi (2d+1) k / 2D) on the folded inputs, d < D/4.  A thread owns
            //      two coefficients of one class mod 4, k and k + H (H = 0 mod 4), so one broadcast load of the inputs
            //      feeds both; lanes hold consecutive k of one parity; cosines come from the parity-split shared table at
            //      byte addresses (entry ((2d+1) k - parity) / 2, four columns per step at +k entries each; the table is
            //      extended by 4m entries so only the first address wraps).  Odd k also need sin x = cos(x -+ pi/2): the
            //      same table half, D/2 entries back (k = 1 mod 4) or ahead (k = 3 mod 4, which supplies the minus sign).
            //      An odd pair costs twice an even one per column, so the odd pairs split the columns twice as finely. ----
            for (int w = ftid; w < HP * (DSo + DSe); w += NFT) {
                const bool odd = w < HP * DSo;
                const int wi = odd ? w : w - HP * DSo;
                const int DSx = odd ? DSo : DSe;
                const int ds = wi / HP, ka = 2 * (wi % HP) + (odd ? 1 : 2);
                if (ka > nk) continue;
                const bool hasb = ka + H <= nk;
                const int kb = hasb ? ka + H : ka;
                const int q0 = QQ * ds / DSx, q1 = QQ * (ds + 1) / DSx;
                const unsigned int tbase = tt_u32 + (odd ? 4u * (unsigned int)TH : 0u), tend = tbase + 8u * D;
                unsigned int oa = tbase + 4u * (unsigned int)((((long long)(8 * q0 + 1) * ka - (odd ? 1 : 0)) / 2) % (2LL * D));
                unsigned int ob = tbase + 4u * (unsigned int)((((long long)(8 * q0 + 1) * kb - (odd ? 1 : 0)) / 2) % (2LL * D));
                const unsigned int sa = 4u * ka, sb = 4u * kb;          // k entries in bytes
                double fa[N], fb[N];
#pragma unroll
                for (int j = 0; j < N; ++j) { fa[j] = 0.0; fb[j] = 0.0; }
                const pk2 zero = pk(0.f, 0.f);
                pk2 aa[N][2], ab[N][2];
#pragma unroll
                for (int j = 0; j < N; ++j) { aa[j][0] = zero; aa[j][1] = zero; ab[j][0] = zero; ab[j][1] = zero; }
                auto flush = [&]() {
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        float l0, h0, l1, h1;
                        unpk(aa[j][0], l0, h0);
                        unpk(aa[j][1], l1, h1);
                        fa[j] += (double)((l0 + h0) + (l1 + h1));
                        unpk(ab[j][0], l0, h0);
                        unpk(ab[j][1], l1, h1);
                        fb[j] += (double)((l0 + h0) + (l1 + h1));
                        aa[j][0] = zero; aa[j][1] = zero; ab[j][0] = zero; ab[j][1] = zero;
                    }
                };
                auto advance = [&](unsigned int &o, unsigned int step) {
                    o += 4 * step;
                    if (o >= tend) o -= 8u * D;
                };
                if (odd) {
                    // sin x at -D/2 entries for k = 1 (mod 4), -sin x at +D/2 entries for k = 3 (mod 4)
                    const unsigned int shift = ((ka & 3) == 1) ? 6u * D : 2u * D;     // -2D or +2D bytes, modulo 8D
                    unsigned int osa = oa + shift, osb = ob + shift;
                    if (osa >= tend) osa -= 8u * D;
                    if (osb >= tend) osb -= 8u * D;
                    const float *yd = Yo + 4 * q0, *yr = Yo + N * quarter + 4 * q0;
                    int q = q0;
                    while (q < q1) {
                        const int qe = min(q1, q + 4);              // 16 columns x (cos, sin) in float32, then float64
                        for (; q < qe; ++q) {
                            const pk2 ca01 = pk(lds_f32(oa), lds_f32(oa + sa)), ca23 = pk(lds_f32(oa + 2 * sa), lds_f32(oa + 3 * sa));
                            const pk2 na01 = pk(lds_f32(osa), lds_f32(osa + sa)), na23 = pk(lds_f32(osa + 2 * sa), lds_f32(osa + 3 * sa));
                            const pk2 cb01 = pk(lds_f32(ob), lds_f32(ob + sb)), cb23 = pk(lds_f32(ob + 2 * sb), lds_f32(ob + 3 * sb));
                            const pk2 nb01 = pk(lds_f32(osb), lds_f32(osb + sb)), nb23 = pk(lds_f32(osb + 2 * sb), lds_f32(osb + 3 * sb));
                            advance(oa, sa); advance(osa, sa); advance(ob, sb); advance(osb, sb);
#pragma unroll
                            for (int j = 0; j < N; ++j) {
                                const float4 v = *reinterpret_cast<const float4 *>(yd + j * quarter);
                                const float4 r = *reinterpret_cast<const float4 *>(yr + j * quarter);
                                const pk2 v01 = pk(v.x, v.y), v23 = pk(v.z, v.w), r01 = pk(r.x, r.y), r23 = pk(r.z, r.w);
                                aa[j][0] = fma2(v01, ca01, aa[j][0]);
                                aa[j][1] = fma2(v23, ca23, aa[j][1]);
                                aa[j][0] = fma2(r01, na01, aa[j][0]);
                                aa[j][1] = fma2(r23, na23, aa[j][1]);
                                ab[j][0] = fma2(v01, cb01, ab[j][0]);
                                ab[j][1] = fma2(v23, cb23, ab[j][1]);
                                ab[j][0] = fma2(r01, nb01, ab[j][0]);
                                ab[j][1] = fma2(r23, nb23, ab[j][1]);
                            }
                            yd += 4;
                            yr += 4;
                        }
                        flush();
                    }
                } else {
                    const float *ye = Ye + (((ka & 3) == 0) ? 0 : N * quarter) + 4 * q0;      // EE or EO
                    int q = q0;
                    while (q < q1) {
                        const int qe = min(q1, q + 8);              // 32 columns in float32 (4 chains of 8), then float64
                        for (; q < qe; ++q) {
                            const pk2 ca01 = pk(lds_f32(oa), lds_f32(oa + sa)), ca23 = pk(lds_f32(oa + 2 * sa), lds_f32(oa + 3 * sa));
                            const pk2 cb01 = pk(lds_f32(ob), lds_f32(ob + sb)), cb23 = pk(lds_f32(ob + 2 * sb), lds_f32(ob + 3 * sb));
                            advance(oa, sa); advance(ob, sb);
#pragma unroll
                            for (int j = 0; j < N; ++j) {
                                const float4 v = *reinterpret_cast<const float4 *>(ye + j * quarter);
                                const pk2 v01 = pk(v.x, v.y), v23 = pk(v.z, v.w);
                                aa[j][0] = fma2(v01, ca01, aa[j][0]);
                                aa[j][1] = fma2(v23, ca23, aa[j][1]);
                                ab[j][0] = fma2(v01, cb01, ab[j][0]);
                                ab[j][1] = fma2(v23, cb23, ab[j][1]);
                            }
                            ye += 4;
                        }
                        flush();
                    }
                }
                double *fp = (odd ? Fp : Fp + (size_t)DSo * N * H) + (size_t)ds * N * H;       // [ds][j][k slot]
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    fp[j * H + (ka - 1) / 2] = fa[j];
                    if (hasb) fp[j * H + (kb - 1) / 2] = fb[j];
                }
            }
            fin_bar();
            WS_ACC(t_2a);
            const bool layer_bad = s_flag != 0;       // read by everyone before thread 0 can start the next item
            for (int i = ftid; i < N * nk; i += NFT) {
                const int j = i / nk, k = 1 + i % nk;
                const double *fp = ((k & 1) ? Fp : Fp + (size_t)DSo * N * H) + j * H + (k - 1) / 2;
                const int nds = (k & 1) ? DSo : DSe;
                double f = 0.0;
                for (int ds = 0; ds < nds; ++ds) f += fp[(size_t)ds * N * H];
                Fr[i] = f;
            }
            fin_bar();
            // everyone has read s_flag (layer_bad): reset it for the next stage 1, which no thread reaches before the
            // barrier after pass 2b
#ifndef WS_OLD_TICKET
            if (ftid == 0) s_flag = 0;
#endif
            WS_ACC(t_red);
            // ---- pass 2b: Z[j][c] = sum_k cos(pi (2c+1) k / 2m) F[j][k].  cos(pi (2(m-1-c)+1) k / 2m) =
            //      (-1)^k cos(pi (2c+1) k / 2m), so columns c and m-1-c share their products: with E / O the sums over
            //      the even / odd k, Z[c] = E + O and Z[m-1-c] = E - O.  Two adjacent lanes per column pair, one per
            //      parity, two float64 chains each ----
            {
                const int MP = (m + 1) / 2;
                const int w = ftid;                      // one pass: N * MP * 2 <= NFT is checked on the host
                const bool act = w < N * MP * 2;
                const int h = w & 1, t = act ? (w >> 1) : 0;
                const int j = t / MP, c = t % MP;
                double sum = 0.0;
                if (act) {
                    const int stepc = 2 * c + 1, step2 = 2 * stepc, m4 = 4 * m;
                    const double *fr = Fr + j * nk;
                    int k = h ? 1 : 2;
                    int idx = stepc * k;                 // < 4m
                    double z0 = 0.0, z1 = 0.0;
                    for (; k + 2 <= nk; k += 4) {
                        z0 = fma(Tm[idx + (idx >> 4)], fr[k - 1], z0);
                        idx += step2;
                        if (idx >= m4) idx -= m4;
                        z1 = fma(Tm[idx + (idx >> 4)], fr[k + 1], z1);
                        idx += step2;
                        if (idx >= m4) idx -= m4;
                    }
                    if (k <= nk) z0 = fma(Tm[idx + (idx >> 4)], fr[k - 1], z0);
                    sum = z0 + z1;
                }
                const double other = __shfl_xor_sync(0xffffffffu, sum, 1);
                if (act && h == 0) {
                    Z[j * m + c] = sum + other;                              // E + O
                    if (m - 1 - c != c) Z[j * m + (m - 1 - c)] = sum - other;   // E - O
                }
            }
            fin_bar();
            WS_ACC(t_2b);
            // ---- per-row min-max, *127 and the truncating int8 cast (fingerprint.py:193-195): one warp per row, no
            //      further CTA-level barrier; the other warps go ahead to the next item ----
            if (fwarp < N) {
                const int j = fwarp;
                double mn = INFINITY, mx = -INFINITY;
                int bad = 0;
                for (int c = lane; c < m; c += 32) {
                    const double z = Z[j * m + c];
                    bad |= !(z == z);
                    mn = fmin(mn, z);
                    mx = fmax(mx, z);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
                }
                const bool zero_row = layer_bad || bad || !(mx > mn);
                int8_t *out = p.out + (int64_t)dom_index * p.out_stride + (int64_t)layer * (N * m) + j * m;
                for (int c = lane; c < m; c += 32) {
                    int qv = 0;
                    if (!zero_row) qv = (int)(((Z[j * m + c] - mn) / (mx - mn)) * 127.0);
                    out[c] = (int8_t)qv;
                }
            }
            WS_ACC(t_rest);
        };

        // dst[i] = sum over the nsl partial-sum slabs, added in slot order (so the result does not depend on which
        // CTA arrives last); the loads of one slab row are independent, EPT of them in flight per thread
        auto sum_slabs = [&](double *dst, const double *slab, int nsl) {
            constexpr int EPT = (K * D + NFT - 1) / NFT;
            double sacc[EPT];
#pragma unroll
            for (int e = 0; e < EPT; ++e) sacc[e] = 0.0;
            // two slab rows in flight (the adds stay in slot order)
            int sp = 0;
#ifndef WS_DBG_SLAB
            for (; sp + 2 <= nsl; sp += 2) {
                const double *row0 = slab + (int64_t)sp * (K * D), *row1 = row0 + K * D;
                double v0[EPT], v1[EPT];
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int i = ftid + e * NFT;
                    v0[e] = (i < K * D) ? __ldcg(row0 + i) : 0.0;
                }
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int i = ftid + e * NFT;
                    v1[e] = (i < K * D) ? __ldcg(row1 + i) : 0.0;
                }
#pragma unroll
                for (int e = 0; e < EPT; ++e) sacc[e] = (sacc[e] + v0[e]) + v1[e];
            }
#endif
            for (; sp < nsl; ++sp) {
                const double *row = slab + (int64_t)sp * (K * D);
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int i = ftid + e * NFT;
                    sacc[e] += (i < K * D) ? __ldcg(row + i) : 0.0;
                }
            }
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int i = ftid + e * NFT;
                if (i < K * D) dst[i] = sacc[e];
            }
        };

        int fb = 0;
        unsigned int fph = 0;
        for (;;) {
            WS_ACC(t_s1);
            WS_WAIT(&u_full[fb], fph, 3);
            WS_ACC(t_idle);
            const int slot = u_slot[fb];
            if (slot < 0) break;
#ifdef WS_NOFIN
            fin_bar();
            if (ftid == 0) mbar_arrive(&u_empty[fb]);
            if (++fb == Cfg::NB) { fb = 0; fph ^= 1u; }
            continue;
#endif
            const int *ds_ = desc + slot * 32;
            const int dom = ds_[kWDom], layer = ds_[kWLayer], iflags = ds_[kWFlags];
            double *u = ubuf + (size_t)fb * K * D;
            // The rider ticket is taken by one thread of the last finisher warp (the warp with the least stage-1 work) and
            // only looked at after stage 1: the fence, the atomic's round trip to L2 and stage 1 overlap, and no barrier
            // is needed before stage 1 (s_flag is reset inside finish_rest, s_rlast is written after stage 1).
            // The consumers wrote this item's rider sums to the slab before the hand-over through the CTA-scope mbarrier;
            // the release at GPU scope needs the fence in the thread that takes the ticket.
            int ticket = -1;
#ifdef WS_OLD_TICKET
            const bool ticket_thread = false;
            if (ftid == 0) {
                s_flag = 0;
                s_rlast = 0;
                if (RIDER && (iflags & kWfRider)) {
                    __threadfence();
                    const int t0 = atomicAdd(&p.counters[ds_[kWRiderCounter]], 1);
                    s_rlast = (t0 == ds_[kWRiderNsplit] - 1);
                }
            }
            if (!(iflags & kWfSplit)) fin_bar();
#else
            const bool ticket_thread = RIDER && ftid == NFT - 32 && (iflags & kWfRider);
            if (ticket_thread) {
                __threadfence();
                ticket = atomicAdd(&p.counters[ds_[kWRiderCounter]], 1);
            }
#endif
            bool own_ready = true;
            if (iflags & kWfSplit) {
                // a domain assembled from several items: publish this item's sums, the last to arrive adds them up
                const int nsplit = ds_[kWNsplit];
                double *slab = p.partials + (int64_t)ds_[kWSlabBase] * (K * D);
                double *mine = slab + (int64_t)ds_[kWSplit] * (K * D);
                for (int i = ftid; i < K * D; i += NFT) mine[i] = u[i];
                __threadfence();
                fin_bar();
                if (ftid == 0) {
                    const int t2 = atomicAdd(&p.counters[ds_[kWCounter]], 1);
                    s_last = (t2 == nsplit - 1);
                }
                fin_bar();
                own_ready = s_last != 0;
                if (own_ready) {
                    __threadfence();
                    sum_slabs(u, slab, nsplit);     // the hand-over buffer is ours until it is released below
                    fin_bar();
                }
            }
#ifdef WS_DBG_BAR
            fin_bar();
#endif
            auto from_u = [&](int k, int d) { return u[k * D + d]; };
            if (own_ready) stage1(from_u);
#ifndef WS_OLD_TICKET
            if (RIDER && ftid == NFT - 32) s_rlast = ticket_thread && (ticket == ds_[kWRiderNsplit] - 1);
#endif
            fin_bar();
            const bool rider_last = RIDER && (s_rlast != 0);
            if constexpr (RIDER) {
                if (rider_last) {
                    // last contributor of the protein's global fingerprint: finish this item's own domain, then add
                    // the rider slabs in slot order into the (still owned) hand-over buffer and finish the global one
                    if (own_ready) finish_rest(dom, layer);
#ifdef WS_OLD_TICKET
                    if (ftid == 0) s_flag = 0;
#endif
                    __threadfence();
                    sum_slabs(u, p.partials + (int64_t)ds_[kWRiderSlabBase] * (K * D), ds_[kWRiderNsplit]);
                    fin_bar();
                    stage1(from_u);
                    fin_bar();
                }
            }
            if (ftid == 0) mbar_arrive(&u_empty[fb]);      // every finisher thread is done with ubuf[fb]
            if (++fb == Cfg::NB) { fb = 0; fph ^= 1u; }
            if (rider_last) finish_rest(ds_[kWRiderDom], layer);
            else if (own_ready) finish_rest(dom, layer);
            else fin_bar();                                 // s_last / s_rlast are rewritten by the next iteration
        }
        WS_ACC(t_s1);
        if (ftid == 0) { WS_PUT(5, t_idle); WS_PUT(6, t_idle + t_s1 + t_2a + t_rest + t_red + t_2b); WS_PUT(7, t_s1); WS_PUT(8, t_2a); WS_PUT(9, t_rest);
                          WS_PUT(11, t_red); WS_PUT(12, t_2b); }
    }
}
