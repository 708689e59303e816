// k2_merge.cu -- K2: union of the samples' variant sites.
//
// Replaces the dict-of-lists loop of merge_sites.py:94-116 and the sorted() of utils.py:1068.  Input: every
// sample's (chrom_rank << 32 | pos) keys, samples concatenated in sorted-sample-directory order, plus the owning
// sample of each key.  A stable LSD radix sort by key keeps the sample order inside each key (that is the order
// merge_sites appends names in); a run-length pass yields the unique keys and their sample counts.
// HBM traffic: 12 B/key per radix pass (8 passes worst case) + 12 B/key for the run-length pass -- a few MB for
// the 1000-sample configuration; latency-bound, not bandwidth-bound.
// The sort and the run-length encode are CUB device primitives (header-only, compiled into this library for
// sm_100a); everything else in the library is hand-written.
#include "internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

namespace snpgpu {

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void k2_cub_bytes(size_t n, size_t *sort_b, size_t *rle_b) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n);
    cub::DeviceRunLengthEncode::Encode(nullptr, b, (const uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr,
                                       (unsigned long long *)nullptr, (int64_t)n);
    *sort_b = a; *rle_b = b;
}

size_t k2_workspace_bytes(size_t n) {
    size_t a, b;
    k2_cub_bytes(n, &a, &b);
    return align256(n * sizeof(uint64_t)) + align256(a > b ? a : b) + 256;
}

int k2_launch(cudaStream_t stream, const uint64_t *keys, const uint32_t *sample_of, size_t n, uint64_t *uniq_out,
              uint32_t *count_out, uint32_t *samples_out, unsigned long long *n_uniq_dev, void *tmp, size_t tmp_bytes,
              int *launches) {
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(n_uniq_dev, 0, sizeof(unsigned long long), stream);
        return e == cudaSuccess ? 0 : SNPGPU_E_CUDA;
    }
    size_t a, b;
    k2_cub_bytes(n, &a, &b);
    uint64_t *sorted = reinterpret_cast<uint64_t *>(tmp);
    void *cub_tmp = reinterpret_cast<uint8_t *>(tmp) + align256(n * sizeof(uint64_t));
    size_t cub_bytes = tmp_bytes - align256(n * sizeof(uint64_t));
    if (cub_bytes < (a > b ? a : b)) return SNPGPU_E_NOMEM;
    size_t sb = cub_bytes;
    if (cub::DeviceRadixSort::SortPairs(cub_tmp, sb, keys, sorted, sample_of, samples_out, (int64_t)n, 0, 64, stream) !=
        cudaSuccess)
        return SNPGPU_E_CUDA;
    size_t rb = cub_bytes;
    if (cub::DeviceRunLengthEncode::Encode(cub_tmp, rb, sorted, uniq_out, count_out, n_uniq_dev, (int64_t)n, stream) !=
        cudaSuccess)
        return SNPGPU_E_CUDA;
    *launches += 0;    // (the sort and the run-length encode are CUB's kernels: not counted among this library's own launches)
    return 0;
}

}  // namespace snpgpu
