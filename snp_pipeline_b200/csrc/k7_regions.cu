// k7_regions.cu -- K7: the arithmetic of filter_regions -- which SNPs lie in an abnormal region.
//
// Replaces find_dense_regions (filter_regions.py:17-71), utils.merge_regions (utils.py:1168-1282) and utils.in_region
// (utils.py:1285-1318) as filter_regions.py:386-428 and :296-303 / :375-383 use them: a window of `window` positions
// that starts at a SNP and holds more than `max_snps` SNPs of one sample's contig makes the region
// [pos[i], pos[i + max_snps]] dense; a SNP is removed when it lies in a dense region or an edge region of its contig --
// of any sample (mode "all") or of its own sample (mode "each").  Merging regions does not change which integer positions
// they cover, so membership is tested against the union directly.
//   k7_dense_kernel    per SNP and (max_snps, window) pair: the dense-region test on the sample's sorted positions
//                      -> (key of the start, end) pairs; pairs that are no region get the key ~0 and sort to the back
//   k2_sort_pairs      the regions by (group, contig, start)  (K2's stable radix sort, k2_merge.cu)
//   k7_prefmax_kernel  running maximum of (group, contig, end) along the sorted regions: the furthest end so far
//   k7_query_kernel    per SNP: the last region that starts at or before it (binary search) -- inside iff the running
//                      maximum there belongs to the SNP's (group, contig) and reaches it
// Keys: group << 48 | contig rank << 32 | position (group 0 in mode "all", the sample index in mode "each").
// A few thousand SNPs per sample: latency-bound, a handful of microseconds.
#include "internal.h"

namespace snpgpu {

constexpr unsigned long long K7_NONE = ~0ull;

size_t k7_workspace_bytes(size_t n, int n_params, size_t n_extra) {
    const size_t m = n * (size_t)(n_params > 0 ? n_params : 1) + n_extra + 1;
    return 2 * ((m * 8 + 255) & ~(size_t)255) + 2 * ((m * 4 + 255) & ~(size_t)255) + k2_workspace_bytes(m) + 512;
}

__global__ void k7_dense_kernel(const unsigned long long *snp, const uint32_t *seg_last, size_t n, const int32_t *max_snps,
                                const int32_t *window, int n_params, const unsigned long long *extra_keys,
                                const uint32_t *extra_end, size_t n_extra, unsigned long long *rkey, uint32_t *rend) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t m = n * (size_t)n_params;
    if (t < m) {
        const size_t i = t / (size_t)n_params;
        const int q = (int)(t % (size_t)n_params);
        unsigned long long key = K7_NONE;
        uint32_t end = 0;
        const long long j = (long long)i + max_snps[q];       // (idx + max_allowed_snps) < snp_count, filter_regions.py:65
        if (max_snps[q] >= 0 && j <= (long long)seg_last[i]) {
            const long long start = (long long)(snp[i] & 0xffffffffull), stop = (long long)(snp[j] & 0xffffffffull);
            if (start + (long long)window[q] - 1 >= stop) { key = snp[i]; end = (uint32_t)stop; }      // :67
        }
        rkey[t] = key;
        rend[t] = end;
    } else if (t < m + n_extra) {                             // the contigs' edge regions (filter_regions.py:411-417)
        rkey[t] = extra_keys[t - m];
        rend[t] = extra_end[t - m];
    }
}

// running maximum of (key's group and contig | end) over the sorted regions, one block
__global__ void __launch_bounds__(1024) k7_prefmax_kernel(const unsigned long long *rkey, const uint32_t *rend, size_t m,
                                                          unsigned long long *pm) {
    __shared__ unsigned long long part[1024];
    const size_t per = (m + 1023) / 1024;
    const size_t lo = (size_t)threadIdx.x * per, hi = lo + per < m ? lo + per : m;
    unsigned long long best = 0;
    for (size_t i = lo; i < hi; i++) {
        const unsigned long long v = rkey[i] == K7_NONE ? 0ull : ((rkey[i] & ~0xffffffffull) | rend[i]);
        best = v > best ? v : best;
    }
    part[threadIdx.x] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; t++) { const unsigned long long v = part[t]; part[t] = run; run = v > run ? v : run; }
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (size_t i = lo; i < hi; i++) {
        const unsigned long long v = rkey[i] == K7_NONE ? 0ull : ((rkey[i] & ~0xffffffffull) | rend[i]);
        run = v > run ? v : run;
        pm[i] = run;
    }
}

__global__ void k7_query_kernel(const unsigned long long *snp, size_t n, const unsigned long long *rkey,
                                const unsigned long long *pm, size_t m, uint8_t *removed) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = snp[i];
    size_t lo = 0, hi = m;                                    // first region whose start key is > k
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (rkey[mid] <= k) lo = mid + 1; else hi = mid;
    }
    bool in = false;
    if (lo > 0) {
        const unsigned long long best = pm[lo - 1];           // the furthest end among the regions that start at or before k
        in = (best >> 32) == (k >> 32) && (best & 0xffffffffull) >= (k & 0xffffffffull);      // utils.py:1314-1316
    }
    removed[i] = in ? 1 : 0;
}

int k7_launch(cudaStream_t stream, const uint64_t *snp_keys, const uint32_t *seg_last, size_t n, const int32_t *max_snps,
              const int32_t *window, int n_params, const uint64_t *extra_keys, const uint32_t *extra_end, size_t n_extra,
              uint8_t *removed_out, void *tmp, size_t tmp_bytes, int *launches) {
    if (n == 0) return 0;
    const size_t m = n * (size_t)n_params + n_extra;
    if (m == 0) return cudaMemsetAsync(removed_out, 0, n, stream) == cudaSuccess ? 0 : SNPGPU_E_CUDA;
    if (tmp_bytes < k7_workspace_bytes(n, n_params, n_extra)) return SNPGPU_E_NOMEM;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    uint8_t *t = reinterpret_cast<uint8_t *>(tmp);
    unsigned long long *rkey = reinterpret_cast<unsigned long long *>(t);
    unsigned long long *pm = reinterpret_cast<unsigned long long *>(t + up(m * 8));
    uint32_t *rend = reinterpret_cast<uint32_t *>(t + 2 * up(m * 8));
    uint32_t *rend_sorted = reinterpret_cast<uint32_t *>(t + 2 * up(m * 8) + up(m * 4));
    void *sort_tmp = t + 2 * up(m * 8) + 2 * up(m * 4);
    k7_dense_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>((const unsigned long long *)snp_keys, seg_last, n, max_snps, window,
                                                                   n_params, (const unsigned long long *)extra_keys, extra_end, n_extra,
                                                                   rkey, rend);
    const unsigned long long *sorted = nullptr;
    void *spare = nullptr;
    uint32_t *hist = nullptr;
    if (int rc = k2_sort_pairs(stream, (const uint64_t *)rkey, rend, m, rend_sorted, sort_tmp, tmp_bytes - (size_t)((uint8_t *)sort_tmp - t),
                               &sorted, &spare, &hist, launches))
        return rc;
    k7_prefmax_kernel<<<1, 1024, 0, stream>>>(sorted, rend_sorted, m, pm);
    k7_query_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const unsigned long long *)snp_keys, n, sorted, pm, m, removed_out);
    *launches += 3;
    return cudaGetLastError() == cudaSuccess ? 0 : SNPGPU_E_CUDA;
}

}  // namespace snpgpu
