// libdctd: version / error plumbing of the C ABI (include/dctd.h).
#include "dctd_internal.cuh"

namespace dctd {
static thread_local cudaError_t g_last_cuda = cudaSuccess;
static thread_local int64_t g_launches = 0;
void set_cuda_error(cudaError_t e) { g_last_cuda = e; }
void count_launch(int n) { g_launches += n; }
}  // namespace dctd

extern "C" {

int dctd_version(void) { return DCTD_VERSION; }

const char *dctd_strerror(int code) {
    switch (code) {
        case DCTD_OK: return "ok";
        case DCTD_ERR_ARG: return "invalid argument";
        case DCTD_ERR_CUDA: return "CUDA error (see dctd_last_cuda_error_string)";
        case DCTD_ERR_WORKSPACE: return "workspace too small";
        case DCTD_ERR_UNSUPPORTED: return "configuration not supported by this kernel set";
        case DCTD_ERR_NOMEM: return "host allocation failed";
        default: return "unknown dctd error code";
    }
}

int dctd_last_cuda_error(void) { return (int)dctd::g_last_cuda; }
const char *dctd_last_cuda_error_string(void) { return cudaGetErrorString(dctd::g_last_cuda); }

/* Host -> device staging for the Python surface: n independent copies (one cudaMemcpyAsync each, in order, on
 * `stream`) of h_src[i] (nbytes[i] bytes, pinned or pageable) to d_base + d_off[i].  Saves the per-tensor
 * interpreter overhead when a batch arrives as hundreds of separate host arrays. */
int dctd_h2d_rows(const void *const *h_src, const int64_t *nbytes, int64_t n, void *d_base, const int64_t *d_off,
                  void *stream) {
    if (n < 0 || (n > 0 && (!h_src || !nbytes || !d_base || !d_off))) return DCTD_ERR_ARG;
    for (int64_t i = 0; i < n; ++i) {
        if (nbytes[i] == 0) continue;
        DCTD_CUDA_TRY(cudaMemcpyAsync((char *)d_base + d_off[i], h_src[i], (size_t)nbytes[i], cudaMemcpyHostToDevice,
                                      (cudaStream_t)stream));
    }
    return DCTD_OK;
}

/* The same staging as a gather KERNEL for host arrays the device can read (pinned memory): the SMs pull the pieces
 * over PCIe with 16-byte loads, many pieces in flight, so the link stays busy across array boundaries (a DMA copy per
 * array leaves a gap of a few microseconds between arrays).  d_table: n descriptors in device-accessible memory
 * (pinned host memory or device memory), every src / dst / nbytes a multiple of 16. */
namespace {
__global__ void __launch_bounds__(128, 16) gather_kernel(const dctd_copy_desc *__restrict__ table, long long n) {
    for (long long e = blockIdx.x; e < n; e += gridDim.x) {
        const dctd_copy_desc dsc = table[e];
        const uint4 *src = reinterpret_cast<const uint4 *>(dsc.src);
        uint4 *dst = reinterpret_cast<uint4 *>(dsc.dst);
        const unsigned int m = (unsigned int)(dsc.nbytes / 16);        // pieces are far below 64 GB
        constexpr int U = 4;       // 64 bytes in flight per thread: with one CTA per SM still 1.2 MB, the link needs ~1 MB
        for (unsigned int i0 = threadIdx.x; i0 < m; i0 += blockDim.x * U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int i = i0 + u * blockDim.x;
                if (i < m)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                                 : "l"(src + i));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int i = i0 + u * blockDim.x;
                if (i < m) dst[i] = v[u];
            }
        }
    }
}
}  // namespace

int dctd_h2d_gather(const dctd_copy_desc *d_table, int64_t n, void *stream) {
    if (n < 0 || (n > 0 && !d_table)) return DCTD_ERR_ARG;
    if (n == 0) return DCTD_OK;
    // The link is saturated by ~1 MB in flight (measured: 32 CTAs x 256 threads x 128 bytes reach the 51.4 GB/s that
    // 1184 CTAs reach; scripts/microbench/h2d_pull.cu), so the gather stays small enough to run NEXT TO the fingerprint
    // kernel's one CTA per SM (608 threads x 96 registers, 223 KB of shared memory): one CTA of 128 threads per SM, no
    // shared memory, and <= 32 registers - the register file is split over the four SM sub-partitions, the fingerprint
    // kernel's 19 warps leave 1024 registers in three of them, i.e. room for one warp of 32 registers each (with 48
    // registers the big kernel never got onto an SM before the whole gather launch had ended:
    // scripts/microbench/coreside.cu).  In quantize_stream the copies of batch i+1 then run beside the fingerprint
    // kernel of batch i instead of keeping it off the SMs.
    int dev = 0, n_sm = 0;
    DCTD_CUDA_TRY(cudaGetDevice(&dev));
    DCTD_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int grid = (int)(n < (int64_t)n_sm ? n : (int64_t)n_sm);
    // An SM changes its L1 / shared-memory split only when it is idle.  The fingerprint kernel needs the maximum shared
    // carveout; asking for the same split here (the loads bypass L1 anyway) lets its CTAs start on SMs that run gather
    // CTAs (same microbenchmark: with the default carveout the big kernel waits for the end of the gather launch, 30 ms;
    // with this attribute it is done 1.7 ms after the gather began).
    DCTD_CUDA_TRY(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       (int)cudaSharedmemCarveoutMaxShared));
    gather_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_table, (long long)n);
    DCTD_LAUNCH_CHECK();
    return DCTD_OK;
}

/* RecCut domain strings -> segment arrays with get_doms' rules (reference src/fingerprint.py:160-171), for a whole
 * batch in one call.  See dctd.h. */
int dctd_parse_domains(const char *text, int64_t text_len, int32_t n_str, const int32_t *str_prot,
                       const int32_t *prot_len, int32_t n_prot, int32_t *dom_str, int32_t *dom_seg_off,
                       int32_t *seg_beg, int32_t *seg_end, int64_t max_segs, int32_t *n_dom_out,
                       int32_t *n_irregular_out) {
    if (n_str < 0 || text_len < 0 || max_segs < 0 || !n_dom_out || !n_irregular_out) return DCTD_ERR_ARG;
    if (n_str > 0 && (!text || !str_prot || !prot_len || !dom_str || !dom_seg_off || !seg_beg || !seg_end))
        return DCTD_ERR_ARG;
    int32_t n_dom = 0, n_irr = 0;
    int64_t n_seg = 0, pos = 0;
    if (dom_seg_off) dom_seg_off[0] = 0;
    for (int32_t i = 0; i < n_str; ++i) {
        const int32_t p = str_prot[i];
        if (p < 0 || p >= n_prot) return DCTD_ERR_ARG;
        const int64_t rows_p = prot_len[p];
        const int64_t seg0 = n_seg;
        int64_t rows = 0;
        bool regular = true, at_end = false;
        // one string: "beg-end" segments separated by ',', terminated by '\n' (or the end of the text)
        while (!at_end) {
            int64_t v[2] = {0, 0};
            for (int part = 0; part < 2 && regular; ++part) {
                int digits = 0;
                while (pos < text_len && text[pos] >= '0' && text[pos] <= '9') {
                    if (v[part] < (1LL << 40)) v[part] = v[part] * 10 + (text[pos] - '0');
                    ++pos;
                    ++digits;
                }
                if (digits == 0) regular = false;
                if (part == 0) {
                    if (pos < text_len && text[pos] == '-') ++pos;
                    else regular = false;
                }
            }
            if (regular) {
                // reference: a segment whose begin lies beyond the protein (or begin 0) takes get_doms' special paths
                if (v[0] < 1 || v[0] > rows_p) {
                    regular = false;
                } else if (n_seg >= max_segs) {
                    return DCTD_ERR_WORKSPACE;
                } else {
                    const int64_t lo = v[0] - 1, hi = v[1] < rows_p ? v[1] : rows_p;
                    seg_beg[n_seg] = (int32_t)lo;
                    seg_end[n_seg] = (int32_t)(hi > lo ? hi : lo);
                    rows += seg_end[n_seg] - seg_beg[n_seg];
                    ++n_seg;
                }
            }
            if (regular && pos < text_len && text[pos] == ',') {
                ++pos;
                continue;
            }
            if (!(pos >= text_len || text[pos] == '\n')) regular = false;      // anything else: leave it to the caller
            while (pos < text_len && text[pos] != '\n') ++pos;               // skip to the end of this string
            if (pos < text_len) ++pos;
            at_end = true;
        }
        if (!regular) {
            ++n_irr;
            n_seg = seg0;
            continue;
        }
        if (rows == 0) {          // reference: an empty embedding slice -> the domain is skipped
            n_seg = seg0;
            continue;
        }
        dom_str[n_dom] = i;
        dom_seg_off[n_dom + 1] = (int32_t)n_seg;
        ++n_dom;
    }
    *n_dom_out = n_dom;
    *n_irregular_out = n_irr;
    return DCTD_OK;
}

int64_t dctd_launch_count(int reset) {
    int64_t v = dctd::g_launches;
    if (reset) dctd::g_launches = 0;
    return v;
}

}  // extern "C"
