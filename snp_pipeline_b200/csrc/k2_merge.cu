// k2_merge.cu -- K2: union of the samples' variant sites.
//
// Replaces the dict-of-lists loop of merge_sites.py:94-116 and the sorted() of utils.py:1068.  Input: every
// sample's (chrom_rank << 32 | pos) keys, samples concatenated in sorted-sample-directory order, plus the owning
// sample of each key.  A STABLE least-significant-digit radix sort by key keeps the sample order inside each key (that
// is the order merge_sites appends names in); a unique pass yields the distinct keys and their sample counts.
//
// Hand-written (round 1 used CUB here):
//   k2_or_kernel        OR of all keys -> which of the eight key bytes carry information at all (a single 5 Mbp contig:
//                       three passes instead of eight)
//   per 8-bit digit:
//   k2_hist_kernel      a block counts the digits of its 2048 keys in shared memory -> hist[digit][block]
//   k2_scan_kernel      exclusive prefix over hist in (digit, block) order: where each block's keys of each digit go
//   k2_scatter_kernel   the block walks its keys in order, 256 at a time; rank within the warp by __match_any_sync,
//                       warps ordered through shared counters: a stable scatter of (key, sample)
//   then:
//   k2_heads_kernel     run heads (key differs from its left neighbour) counted per block
//   k2_scan_kernel      prefix over the blocks' counts
//   k2_unique_kernel    unique keys and the start of each run;  k2_counts_kernel  run lengths
// HBM traffic: 24 B/key per pass (12 in, 12 out; the histogram pass re-reads 8) + 20 B/key for the unique pass -- a few MB
// for configs[1], 0.5 GB for the 10 M keys of the 1000-sample configuration: bandwidth-bound, < 1 % of a step.
#include "internal.h"

namespace snpgpu {

constexpr int K2_THREADS = 256;
constexpr int K2_KPT = 8;                              // keys per thread
constexpr int K2_TILE = K2_THREADS * K2_KPT;           // keys per block
constexpr int K2_SCAN_THREADS = 1024;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t k2_blocks(size_t n) { return (n + K2_TILE - 1) / K2_TILE; }

// workspace: two key buffers, one value buffer, the histogram / block counts, the OR word
size_t k2_workspace_bytes(size_t n) {
    return 2 * align256(n * sizeof(uint64_t)) + align256(n * sizeof(uint32_t)) + align256((256 * k2_blocks(n) + 1) * sizeof(uint32_t)) + 512;
}

__global__ void __launch_bounds__(K2_THREADS) k2_or_kernel(const unsigned long long *keys, size_t n, unsigned long long *key_or) {
    unsigned long long v = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v |= keys[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicOr(key_or, v);
}

__global__ void __launch_bounds__(K2_THREADS) k2_hist_kernel(const unsigned long long *keys, size_t n, int shift, uint32_t *hist,
                                                             size_t n_blocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * K2_TILE;
#pragma unroll
    for (int r = 0; r < K2_KPT; r++) {
        const size_t i = base + (size_t)r * K2_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive prefix sum over a[0 .. m), in place, one block; *total (nullable) receives the sum
__global__ void __launch_bounds__(K2_SCAN_THREADS) k2_scan_kernel(uint32_t *a, size_t m, unsigned long long *total) {
    __shared__ unsigned long long part[K2_SCAN_THREADS];
    const size_t per = (m + K2_SCAN_THREADS - 1) / K2_SCAN_THREADS;
    const size_t lo = (size_t)threadIdx.x * per, hi = lo + per < m ? lo + per : m;
    unsigned long long s = 0;
    for (size_t i = lo; i < hi; i++) s += a[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {                                   // (1024 partial sums: serial is cheaper than it looks)
        unsigned long long run = 0;
        for (int t = 0; t < K2_SCAN_THREADS; t++) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        if (total) *total = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (size_t i = lo; i < hi; i++) { const uint32_t v = a[i]; a[i] = (uint32_t)run; run += v; }
}

__global__ void __launch_bounds__(K2_THREADS) k2_scatter_kernel(const unsigned long long *keys_in, const uint32_t *vals_in, size_t n,
                                                                int shift, const uint32_t *offs, size_t n_blocks,
                                                                unsigned long long *keys_out, uint32_t *vals_out) {
    __shared__ uint32_t base_of[256];                         // where the block's next key of each digit goes
    __shared__ uint32_t wcount[K2_THREADS / 32][256];         // this round's keys per warp and digit
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base_of[threadIdx.x] = offs[(size_t)threadIdx.x * n_blocks + blockIdx.x];
    const size_t tile0 = (size_t)blockIdx.x * K2_TILE;
    for (int r = 0; r < K2_KPT; r++) {
        for (int w = 0; w < K2_THREADS / 32; w++) wcount[w][threadIdx.x] = 0;
        __syncthreads();
        const size_t i = tile0 + (size_t)r * K2_THREADS + threadIdx.x;
        const bool have = i < n;
        unsigned long long k = 0;
        uint32_t v = 0, d = 0, rank = 0;
        if (have) { k = keys_in[i]; v = vals_in[i]; d = (unsigned)(k >> shift) & 0xffu; }
        const uint32_t act = __ballot_sync(0xffffffffu, have);
        if (have) {
            const uint32_t same = __match_any_sync(act, d);
            rank = (uint32_t)__popc(same & ((1u << lane) - 1u));
            if (rank == 0) wcount[warp][d] = (uint32_t)__popc(same);      // (the lowest lane of the digit's group)
        }
        __syncthreads();
        if (have) {
            uint32_t before = 0;
            for (int w = 0; w < warp; w++) before += wcount[w][d];
            const size_t o = (size_t)base_of[d] + before + rank;
            keys_out[o] = k;
            vals_out[o] = v;
        }
        __syncthreads();
        uint32_t sum = 0;
        for (int w = 0; w < K2_THREADS / 32; w++) sum += wcount[w][threadIdx.x];
        base_of[threadIdx.x] += sum;
        __syncthreads();
    }
}

// run heads per block
__global__ void __launch_bounds__(K2_THREADS) k2_heads_kernel(const unsigned long long *keys, size_t n, uint32_t *block_heads) {
    uint32_t c = 0;
    const size_t base = (size_t)blockIdx.x * K2_TILE;
#pragma unroll
    for (int r = 0; r < K2_KPT; r++) {
        const size_t i = base + (size_t)threadIdx.x * K2_KPT + r;         // (a thread owns 8 consecutive keys)
        if (i < n && (i == 0 || keys[i] != keys[i - 1])) c++;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ uint32_t ws[K2_THREADS / 32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < K2_THREADS / 32; w++) s += ws[w];
        block_heads[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(K2_THREADS) k2_unique_kernel(const unsigned long long *keys, size_t n, const uint32_t *block_first,
                                                               unsigned long long *uniq_out, uint32_t *start) {
    __shared__ uint32_t ws[K2_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * K2_TILE + (size_t)threadIdx.x * K2_KPT;
    uint32_t heads = 0, mine = 0;
#pragma unroll
    for (int r = 0; r < K2_KPT; r++) {
        const size_t i = base + r;
        if (i < n && (i == 0 || keys[i] != keys[i - 1])) { heads |= 1u << r; mine++; }
    }
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t before = block_first[blockIdx.x] + incl - mine;
    for (int w = 0; w < warp; w++) before += ws[w];
#pragma unroll
    for (int r = 0; r < K2_KPT; r++) {
        if ((heads >> r) & 1u) {
            uniq_out[before] = keys[base + r];
            start[before] = (uint32_t)(base + r);
            before++;
        }
    }
}

__global__ void k2_counts_kernel(const uint32_t *start, const unsigned long long *n_uniq, size_t n, uint32_t *count_out) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long nu = *n_uniq;
    if (k < nu) count_out[k] = (k + 1 < nu ? start[k + 1] : (uint32_t)n) - start[k];
}

// Stable sort of (key, value) pairs by key.  tmp: k2_workspace_bytes(n).  *sorted_keys points into tmp (or at `keys` when no
// key byte carries information); the values land in vals_out; *spare: a key-sized buffer of tmp the result does not use.
int k2_sort_pairs(cudaStream_t stream, const uint64_t *keys, const uint32_t *vals, size_t n, uint32_t *vals_out, void *tmp,
                  size_t tmp_bytes, const unsigned long long **sorted_keys, void **spare, uint32_t **hist_out, int *launches) {
    if (n >= ((size_t)1 << 32)) return SNPGPU_E_ARG;
    if (tmp_bytes < k2_workspace_bytes(n)) return SNPGPU_E_NOMEM;
    const size_t nb = k2_blocks(n);
    uint8_t *t = reinterpret_cast<uint8_t *>(tmp);
    unsigned long long *kbuf[2] = {reinterpret_cast<unsigned long long *>(t),
                                   reinterpret_cast<unsigned long long *>(t + align256(n * sizeof(uint64_t)))};
    uint32_t *vtmp = reinterpret_cast<uint32_t *>(t + 2 * align256(n * sizeof(uint64_t)));
    uint32_t *hist = reinterpret_cast<uint32_t *>(t + 2 * align256(n * sizeof(uint64_t)) + align256(n * sizeof(uint32_t)));
    unsigned long long *key_or = reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(hist) +
                                                                         align256((256 * nb + 1) * sizeof(uint32_t)));
    // ---- which key bytes carry information?  (the one synchronisation of K2; merge_sites hands n_uniq back anyway)
    if (cudaMemsetAsync(key_or, 0, sizeof(unsigned long long), stream) != cudaSuccess) return SNPGPU_E_CUDA;
    k2_or_kernel<<<(unsigned)(nb < 1184 ? nb : 1184), K2_THREADS, 0, stream>>>((const unsigned long long *)keys, n, key_or);
    unsigned long long h_or = 0;
    if (cudaMemcpyAsync(&h_or, key_or, sizeof(h_or), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return SNPGPU_E_CUDA;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return SNPGPU_E_CUDA;
    int shifts[8], n_pass = 0;
    for (int b = 0; b < 8; b++)
        if ((h_or >> (8 * b)) & 0xffull) shifts[n_pass++] = 8 * b;
    *launches += 1;
    // ---- the passes: (keys, values) ping-pong so that the last pass writes the values into vals_out
    const unsigned long long *kin = (const unsigned long long *)keys;
    const uint32_t *vin = vals;
    for (int p = 0; p < n_pass; p++) {
        unsigned long long *kout = kbuf[p & 1];
        uint32_t *vout = ((n_pass - 1 - p) & 1) ? vtmp : vals_out;
        k2_hist_kernel<<<(unsigned)nb, K2_THREADS, 0, stream>>>(kin, n, shifts[p], hist, nb);
        k2_scan_kernel<<<1, K2_SCAN_THREADS, 0, stream>>>(hist, 256 * nb, nullptr);
        k2_scatter_kernel<<<(unsigned)nb, K2_THREADS, 0, stream>>>(kin, vin, n, shifts[p], hist, nb, kout, vout);
        kin = kout; vin = vout;
        *launches += 3;
    }
    if (n_pass == 0 && cudaMemcpyAsync(vals_out, vals, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
        return SNPGPU_E_CUDA;
    *sorted_keys = kin;
    *spare = kin == kbuf[0] ? kbuf[1] : kbuf[0];
    *hist_out = hist;
    return cudaGetLastError() == cudaSuccess ? 0 : SNPGPU_E_CUDA;
}

int k2_launch(cudaStream_t stream, const uint64_t *keys, const uint32_t *sample_of, size_t n, uint64_t *uniq_out,
              uint32_t *count_out, uint32_t *samples_out, unsigned long long *n_uniq_dev, void *tmp, size_t tmp_bytes,
              int *launches) {
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(n_uniq_dev, 0, sizeof(unsigned long long), stream);
        return e == cudaSuccess ? 0 : SNPGPU_E_CUDA;
    }
    const unsigned long long *kin = nullptr;
    void *spare = nullptr;
    uint32_t *hist = nullptr;
    if (int rc = k2_sort_pairs(stream, keys, sample_of, n, samples_out, tmp, tmp_bytes, &kin, &spare, &hist, launches)) return rc;
    const size_t nb = k2_blocks(n);
    // ---- unique keys, run starts, run lengths (the key buffer the last pass did not write holds the starts)
    uint32_t *start = reinterpret_cast<uint32_t *>(spare);
    k2_heads_kernel<<<(unsigned)nb, K2_THREADS, 0, stream>>>(kin, n, hist);
    k2_scan_kernel<<<1, K2_SCAN_THREADS, 0, stream>>>(hist, nb, n_uniq_dev);
    k2_unique_kernel<<<(unsigned)nb, K2_THREADS, 0, stream>>>(kin, n, hist, (unsigned long long *)uniq_out, start);
    k2_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(start, n_uniq_dev, n, count_out);
    *launches += 4;
    return cudaGetLastError() == cudaSuccess ? 0 : SNPGPU_E_CUDA;
}

}  // namespace snpgpu
