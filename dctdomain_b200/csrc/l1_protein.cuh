// Protein-level scores of reference src/dct-sim.py:28-50 for ALL pairs of two protein sets (db_search / all_sim,
// dct-sim.py:126-176), on the SAD scan machinery of l1topk.cu (included there, inside its anonymous namespace).
//
//   phase 1  l1_protein_kernel   the same TMA-staged, warp-specialised tile loop as the threshold scan (queries = the
//                                fingerprints of the query proteins, database = the packed fingerprints of the other set):
//                                every (query fingerprint, database fingerprint) distance is computed exactly once at the
//                                SAD issue rate.  A warp owns TQ consecutive query fingerprints; a run of them that belongs
//                                to one protein is a SEGMENT, reduced to its minimum in registers (lane = database vector,
//                                so no shuffles) and stored as one row of `mseg` [segments, columns]; the distances of a
//                                protein's LAST fingerprint go to `mlast` [proteins, columns].  Plain coalesced stores, no
//                                atomics; ~2.4 rows per 8 query fingerprints for typical proteins.
//   phase 2  l1_protein_reduce_kernel   min over a protein pair's block (segments of a x columns of b) and the
//                                last-vs-last entry -> int32 [n_qprot, n_dbprot] each.
//
// The intermediate matrices cost 2 x 0.1 ms of HBM time per ms of SAD time at 4-5 fingerprints per protein; the host never
// sees a per-pair index array.
#pragma once

struct ProtParams {
    ScanParams sp;              // q / nq: this chunk's query fingerprints; packed / n / n_groups / groups_per_split: database
    const int4 *slices;         // per warp slice of TQ query slots: {first segment row, first protein row, segend | last << 8, 0}
    unsigned int *mseg;         // [segments of the chunk, ld]
    unsigned int *mlast;        // [proteins of the chunk, ld]
    long long ld;               // columns = database groups * 32
};

template <int NW, int TQ, int kTD, int STAGES, int CT = 0>
__global__ void __launch_bounds__((NW + 1) * 32, 1) l1_protein_kernel(const ProtParams pp) {
    static_assert(TQ <= 8, "segment masks are 8 bits wide");
    const ScanParams &p = pp.sp;
    constexpr int QT = NW * TQ;
    extern __shared__ __align__(128) unsigned char smem[];
    const int C = chunks_of(p.d);
    const int dpad = C * 16;
    const int tile_bytes = kTD * 32 * dpad;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem);
    unsigned long long *empty = full + STAGES;
    unsigned char *qs = smem + 128;
    unsigned char *st = qs + (size_t)QT * dpad;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long q0 = (long long)blockIdx.y * QT;
    const long long g_begin = (long long)blockIdx.x * p.groups_per_split;
    const long long g_end = min(p.n_groups, g_begin + p.groups_per_split);
    const int n_tiles = (int)((g_end - g_begin + kTD - 1) / kTD);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_queries<NW, TQ>(p, qs, q0, dpad);
    __syncthreads();

    if (warp == NW) {
        if (lane == 0) producer_loop<kTD, STAGES>(p, full, empty, st, tile_bytes, C, dpad, g_begin, g_end, n_tiles);
        return;
    }
    const uint4 *qs4 = reinterpret_cast<const uint4 *>(qs) + (size_t)warp * TQ * C;
    const int4 sl = pp.slices[q0 / TQ + warp];
    const unsigned int segend = (unsigned int)sl.z & 0xffu, lastm = ((unsigned int)sl.z >> 8) & 0xffu;
    for (int t = 0; t < n_tiles; ++t) {
        const int s = t % STAGES;
        mbar_wait(&full[s], (unsigned int)((t / STAGES) & 1));
        unsigned int acc[TQ][kTD];
        sad_tile<TQ, kTD, CT>(reinterpret_cast<const uint4 *>(st + (size_t)s * tile_bytes), qs4, C, lane, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        const long long vbase = g_begin + (long long)t * kTD;
#pragma unroll
        for (int b = 0; b < kTD; ++b) {
            if (vbase + b >= g_end) break;                       // warp-uniform
            const long long col = (vbase + b) * 32 + lane;
            unsigned int run = 0xffffffffu;
            unsigned int *ps = pp.mseg + (long long)sl.x * pp.ld + col;
            unsigned int *pl = pp.mlast + (long long)sl.y * pp.ld + col;
#pragma unroll
            for (int a = 0; a < TQ; ++a) {
                run = min(run, acc[a][b]);
                if (segend & (1u << a)) {                       // warp-uniform
                    *ps = run;
                    ps += pp.ld;
                    run = 0xffffffffu;
                }
                if (lastm & (1u << a)) {
                    *pl = acc[a][b];
                    pl += pp.ld;
                }
            }
        }
    }
}

// out_min[a, b] = min over (segments of query protein a) x (columns of database protein b) of mseg,
// out_last[a, b] = mlast[last_row[a], last column of b]; INT32_MAX where either protein has no fingerprints
__global__ void __launch_bounds__(256) l1_protein_reduce_kernel(const unsigned int *__restrict__ mseg,
                                                                const unsigned int *__restrict__ mlast, long long ld,
                                                                const int *__restrict__ prot_seg_off,
                                                                const int *__restrict__ last_row,
                                                                const long long *__restrict__ doff, long long n_dbprot,
                                                                int *__restrict__ out_min, int *__restrict__ out_last,
                                                                long long out_ld) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long a = blockIdx.y;
    if (b >= n_dbprot) return;
    const int s0 = prot_seg_off[a], s1 = prot_seg_off[a + 1];
    const long long c0 = doff[b], c1 = doff[b + 1];
    unsigned int m = 0x7fffffffu, last = 0x7fffffffu;
    for (int s = s0; s < s1; ++s) {
        const unsigned int *row = mseg + (long long)s * ld;
        for (long long c = c0; c < c1; ++c) m = min(m, row[c]);
    }
    if (s1 > s0 && c1 > c0) last = mlast[(long long)last_row[a] * ld + c1 - 1];
    out_min[a * out_ld + b] = (int)m;
    out_last[a * out_ld + b] = (int)last;
}
