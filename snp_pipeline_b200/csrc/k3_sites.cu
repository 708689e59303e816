// k3_sites.cu -- the site table built on the device from K2's output (no trip through the host).
//
// Between merge_sites and call_consensus the reference passes the site list through snplist.txt
// (utils.write_list_of_snps utils.py:1056-1070 -> utils.read_snp_position_list utils.py:1073-1088) and every
// call_consensus process builds its set of positions from that file (call_consensus.py:133, :147-151).  When both
// steps run in one process on one GPU the list never has to leave HBM: the sorted unique keys K2 wrote are turned
// into the table K1 probes (sites.cuh) by a few small kernels: set the bits, prefix sum of the words' popcounts (block sums,
// their offsets, per-block scan: 156 k words for 5 Mbp), pack the per-word records, resolve every snplist entry to its unique-site index.
#include "internal.h"

namespace snpgpu {

// one bit per site; every key is a snplist entry and the list is already in snplist order
__global__ void k3_set_bits_kernel(const unsigned long long *keys, size_t n, int n_contigs, const int64_t *bit_base,
                                   const int64_t *max_pos, uint32_t *bits, uint8_t *flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    flags[i] = SITE_SNP;                                     // (unique-site indices are ranks in the bitmap: at most n of them)
    if (i == n) return;
    const unsigned long long k = keys[i];
    const int c = (int)(k >> 32);
    const int64_t p = (int64_t)(k & 0xffffffffull);
    if (c < n_contigs && p <= max_pos[c]) {                  // (a key outside the caller's contig lengths gets no bit: no line can hit it)
        const int64_t b = bit_base[c] + p;
        atomicOr(&bits[b >> 5], 1u << (b & 31));
    }
}

// snplist entry i -> its unique-site index = the rank of its bit (after the scan); a key that got no bit points at the
// spare cell n, which nothing ever writes: its matrix column reads '-' like any position without a pileup line
__global__ void k3_unique_kernel(const unsigned long long *keys, size_t n, int n_contigs, const int64_t *bit_base,
                                 const int64_t *max_pos, const uint32_t *bits, const uint32_t *rank, int32_t *snp_unique) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const int c = (int)(k >> 32);
    const int64_t p = (int64_t)(k & 0xffffffffull);
    int32_t u = (int32_t)n;
    if (c < n_contigs && p <= max_pos[c]) {
        const int64_t b = bit_base[c] + p;
        u = (int32_t)(rank[b >> 5] + (uint32_t)__popc(bits[b >> 5] & ((1u << (b & 31)) - 1u)));
    }
    snp_unique[i] = u;
}

// rank[w] = number of set bits in bits[0 .. w): exclusive prefix over the words' popcounts.  Three small kernels: every
// block sums the popcounts of its K3_CHUNK words; one block turns the sums into exclusive offsets; every block scans its
// own words from its offset (a single block over the 156 k words of a 5 Mbp contig took 0.17 ms per site table).
constexpr int K3_SCAN_THREADS = 256;
constexpr int K3_PER_THREAD = 8;
constexpr int K3_CHUNK = K3_SCAN_THREADS * K3_PER_THREAD;
// exclusive prefix of v over the block's threads (in thread order); *total = the block's sum
__device__ __forceinline__ uint32_t k3_block_exclusive(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
    const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t before = 0, all = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
        const uint32_t s = warp_sums[w];
        if (w < warp) before += s;
        all += s;
    }
    __syncthreads();
    *total = all;
    return before + incl - v;
}
__global__ void __launch_bounds__(K3_SCAN_THREADS) k3_rank_sum_kernel(const uint32_t *bits, size_t n_words, uint32_t *part) {
    __shared__ uint32_t ws[K3_SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * K3_CHUNK + (size_t)threadIdx.x * K3_PER_THREAD;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < K3_PER_THREAD; j++) if (base + j < n_words) s += (uint32_t)__popc(bits[base + j]);
    uint32_t total;
    k3_block_exclusive(s, ws, &total);
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k3_rank_offsets_kernel(uint32_t *part, size_t n_blocks) {    // one block, in place
    __shared__ uint32_t ws[32];
    uint32_t carry = 0;
    for (size_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const size_t i = b0 + threadIdx.x;
        const uint32_t v = i < n_blocks ? part[i] : 0u;
        uint32_t total;
        const uint32_t ex = k3_block_exclusive(v, ws, &total);
        if (i < n_blocks) part[i] = carry + ex;
        carry += total;
    }
}
__global__ void __launch_bounds__(K3_SCAN_THREADS) k3_rank_kernel(const uint32_t *bits, size_t n_words, const uint32_t *part,
                                                                  uint32_t *rank) {
    __shared__ uint32_t ws[K3_SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * K3_CHUNK + (size_t)threadIdx.x * K3_PER_THREAD;
    uint32_t c[K3_PER_THREAD], s = 0;
#pragma unroll
    for (int j = 0; j < K3_PER_THREAD; j++) { c[j] = base + j < n_words ? (uint32_t)__popc(bits[base + j]) : 0u; s += c[j]; }
    uint32_t total;
    uint32_t run = part[blockIdx.x] + k3_block_exclusive(s, ws, &total);
#pragma unroll
    for (int j = 0; j < K3_PER_THREAD; j++) { if (base + j < n_words) rank[base + j] = run; run += c[j]; }
}

// bits + rank -> the packed words of the first-tier parser (every site is a snplist entry, none is excluded)
__global__ void k3_pack_words_kernel(const uint32_t *bits, const uint32_t *rank, size_t n_words, SiteWord *words) {
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n_words) words[w] = SiteWord{bits[w], bits[w], 0u, rank[w]};
}

size_t k3_scan_bytes(size_t n_words) { return ((n_words + K3_CHUNK - 1) / K3_CHUNK + 1) * sizeof(uint32_t); }   // the blocks' popcount sums

// bits / rank: n_words words each, bits zeroed here; returns the number of kernels launched, or < 0
int k3_launch(cudaStream_t stream, const unsigned long long *keys, size_t n, int n_contigs, const int64_t *bit_base,
              const int64_t *max_pos, uint32_t *bits, uint32_t *rank, size_t n_words, uint8_t *flags,
              int32_t *snp_unique, SiteWord *words, void *tmp, size_t tmp_bytes) {
    if (cudaMemsetAsync(bits, 0, n_words * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    int launches = 0;
    if (n) {
        k3_set_bits_kernel<<<(unsigned)((n + 256) / 256), 256, 0, stream>>>(keys, n, n_contigs, bit_base, max_pos, bits, flags);
        launches++;
    }
    const size_t n_blocks = (n_words + K3_CHUNK - 1) / K3_CHUNK;
    if (tmp_bytes < k3_scan_bytes(n_words)) return -1;
    uint32_t *part = reinterpret_cast<uint32_t *>(tmp);
    if (n_blocks) {
        k3_rank_sum_kernel<<<(unsigned)n_blocks, K3_SCAN_THREADS, 0, stream>>>(bits, n_words, part);
        k3_rank_offsets_kernel<<<1, 1024, 0, stream>>>(part, n_blocks);
        k3_rank_kernel<<<(unsigned)n_blocks, K3_SCAN_THREADS, 0, stream>>>(bits, n_words, part, rank);
    }
    k3_pack_words_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, stream>>>(bits, rank, n_words, words);
    if (n) {
        k3_unique_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(keys, n, n_contigs, bit_base, max_pos, bits, rank, snp_unique);
        launches++;
    }
    return launches + (n_blocks ? 4 : 1);
}

// ---- reference bases at the snplist positions (utils.write_reference_snp_file, utils.py:1091-1110): out[k] =
//      upper(seq[pos[k] - 1]) with Python's indexing -- position 0 and negative positions count from the end -- and the
//      first position outside the sequence reported the way the reference's IndexError would surface --------------
__global__ void k3_reference_bases_kernel(const uint8_t *seq, long long seq_len, const long long *pos, size_t n,
                                          uint8_t *out, unsigned long long *first_bad) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    long long i = pos[k] - 1;
    if (i < 0) i += seq_len;
    if (i < 0 || i >= seq_len) { atomicMin(first_bad, (unsigned long long)k); out[k] = '?'; return; }
    const unsigned c = seq[i];
    out[k] = (uint8_t)((c - 'a' < 26u) ? c - 32u : c);
}

int k3_launch_reference_bases(cudaStream_t stream, const uint8_t *seq, size_t seq_len, const long long *pos, size_t n,
                              uint8_t *out, unsigned long long *first_bad) {
    if (!n) return 0;
    k3_reference_bases_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(seq, (long long)seq_len, pos, n, out, first_bad);
    return 1;
}

}  // namespace snpgpu
