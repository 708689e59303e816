// k4_distance.cu -- K4: all-pairs SNP distance over the samples x sites matrix.
//
// Replaces utils.calculate_sequence_distance (utils.py:1135-1165) under itertools.combinations
// (distance.py:93-96): for every pair of rows, the number of columns where BOTH bytes are one of A C G T
// (case-insensitive) and differ.  '-', 'N', IUPAC codes never count.
//
// Two kernels:
//   k4_pack     bytes -> three bit planes per 32 sites (valid, code bit 0, code bit 1), stored WORD-MAJOR
//               ([plane][word][row]) so that a tile of rows for one word is one contiguous, coalesced run;
//   k4_pairs    64 x 64 pair tiles per CTA, 4 x 4 pairs per thread; the planes of a 32-word slab are staged in
//               shared memory; per pair and word: 3 LOP3 + POPC + IADD  (xor, xor-or, and-and, popc, add).
// Bound: integer ALU / POPC issue, not HBM -- the packed matrix (3 bits per cell) is read once per tile row /
// column from L2.  No tensor cores (integer bit work; BASELINE.json north_star).
#include "internal.h"

namespace snpgpu {

constexpr int K4_TILE = 64;        // pairs per CTA edge
constexpr int K4_KW   = 32;        // words (of 32 sites) per shared-memory slab
constexpr int K4_THREADS = 256;

static size_t k4_pad_rows(size_t n) { return (n + K4_TILE - 1) / K4_TILE * K4_TILE; }
static size_t k4_words(size_t s) { return (s + 31) / 32; }

size_t k4_workspace_bytes(size_t n_rows, size_t n_sites) {
    size_t w = k4_words(n_sites);
    w = (w + K4_KW - 1) / K4_KW * K4_KW;
    return 3 * w * k4_pad_rows(n_rows) * sizeof(uint32_t) + 256;
}

// one warp packs 32 words of one row: lane l reads the 32 bytes of word (w0 + l)
__global__ void k4_pack_kernel(const uint8_t *__restrict__ matrix, size_t n_rows, size_t n_sites, size_t row_stride,
                               size_t n_rows_pad, size_t n_words_pad, uint32_t *__restrict__ planes) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // threads along rows: coalesced stores
    const size_t word = blockIdx.y;
    if (row >= n_rows_pad) return;
    uint32_t v = 0, lo = 0, hi = 0;
    if (row < n_rows) {
        const uint8_t *p = matrix + row * row_stride + word * 32;
        const size_t left = word * 32 < n_sites ? n_sites - word * 32 : 0;
        const int n = left < 32 ? (int)left : 32;
        for (int i = 0; i < n; i++) {
            unsigned c = p[i] & 0xdfu;                                     // upper() for letters (utils.py:1153-1154)
            unsigned is_acgt = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
            // (c >> 1) & 3:  A -> 0, C -> 1, T -> 2, G -> 3
            unsigned code = (c >> 1) & 3u;
            // p[i] & 0xdf maps non-letters onto other bytes, e.g. '!' (0x21) -> 0x01; none of them lands on ACGT
            // unless the original was a letter or one of 0x61..0x7a / 0x41..0x5a: check the original is a letter
            unsigned o = p[i];
            unsigned letter = ((o | 0x20u) - 'a') < 26u;
            is_acgt &= letter;
            v |= is_acgt << i;
            lo |= (is_acgt & code & 1u) << i;
            hi |= (is_acgt & (code >> 1)) << i;
        }
    }
    const size_t plane = n_words_pad * n_rows_pad;
    planes[0 * plane + word * n_rows_pad + row] = v;
    planes[1 * plane + word * n_rows_pad + row] = lo;
    planes[2 * plane + word * n_rows_pad + row] = hi;
}

struct K4Smem {
    uint32_t a[3][K4_KW][K4_TILE];     // [plane][word][row of the i tile]
    uint32_t b[3][K4_KW][K4_TILE];
};

__global__ void __launch_bounds__(K4_THREADS) k4_pairs_kernel(const uint32_t *__restrict__ planes, size_t n_rows,
                                                             size_t n_rows_pad, size_t n_words_pad, size_t row_begin,
                                                             size_t row_end, int triangle, int32_t *__restrict__ dist) {
    extern __shared__ __align__(16) uint8_t k4_smem_raw[];
    K4Smem &sm = *reinterpret_cast<K4Smem *>(k4_smem_raw);
    const size_t i0 = row_begin / K4_TILE * K4_TILE + (size_t)blockIdx.y * K4_TILE;
    const size_t j0 = (size_t)blockIdx.x * K4_TILE;
    if (triangle && j0 < i0) return;                       // the mirror tile writes both halves
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const size_t plane = n_words_pad * n_rows_pad;
    uint32_t acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) acc[x][y] = 0;

    for (size_t w0 = 0; w0 < n_words_pad; w0 += K4_KW) {
        // stage: 3 planes x 32 words x 64 rows for each side; rows are contiguous in the word-major layout
        for (int idx = tid; idx < 3 * K4_KW * (K4_TILE / 4); idx += K4_THREADS) {
            const int r4 = idx % (K4_TILE / 4);
            const int k = (idx / (K4_TILE / 4)) % K4_KW;
            const int p = idx / (K4_TILE / 4 * K4_KW);
            const uint32_t *src = planes + p * plane + (w0 + k) * n_rows_pad;
            *reinterpret_cast<uint4 *>(&sm.a[p][k][r4 * 4]) = *reinterpret_cast<const uint4 *>(src + i0 + r4 * 4);
            *reinterpret_cast<uint4 *>(&sm.b[p][k][r4 * 4]) = *reinterpret_cast<const uint4 *>(src + j0 + r4 * 4);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < K4_KW; k++) {
            const uint4 av = *reinterpret_cast<const uint4 *>(&sm.a[0][k][ty * 4]);
            const uint4 al = *reinterpret_cast<const uint4 *>(&sm.a[1][k][ty * 4]);
            const uint4 ah = *reinterpret_cast<const uint4 *>(&sm.a[2][k][ty * 4]);
            const uint4 bv = *reinterpret_cast<const uint4 *>(&sm.b[0][k][tx * 4]);
            const uint4 bl = *reinterpret_cast<const uint4 *>(&sm.b[1][k][tx * 4]);
            const uint4 bh = *reinterpret_cast<const uint4 *>(&sm.b[2][k][tx * 4]);
            const uint32_t avv[4] = {av.x, av.y, av.z, av.w}, alv[4] = {al.x, al.y, al.z, al.w},
                           ahv[4] = {ah.x, ah.y, ah.z, ah.w};
            const uint32_t bvv[4] = {bv.x, bv.y, bv.z, bv.w}, blv[4] = {bl.x, bl.y, bl.z, bl.w},
                           bhv[4] = {bh.x, bh.y, bh.z, bh.w};
#pragma unroll
            for (int x = 0; x < 4; x++)
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    const uint32_t d = ((alv[x] ^ blv[y]) | (ahv[x] ^ bhv[y])) & avv[x] & bvv[y];
                    acc[x][y] += (uint32_t)__popc(d);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < 4; x++) {
        const size_t i = i0 + (size_t)ty * 4 + x;
#pragma unroll
        for (int y = 0; y < 4; y++) {
            const size_t j = j0 + (size_t)tx * 4 + y;
            if (i >= n_rows || j >= n_rows) continue;
            const int32_t d = i == j ? 0 : (int32_t)acc[x][y];
            if (i >= row_begin && i < row_end) dist[(i - row_begin) * n_rows + j] = d;
            if (triangle && j0 != i0 && j >= row_begin && j < row_end) dist[(j - row_begin) * n_rows + i] = d;
        }
    }
}

int k4_launch(cudaStream_t stream, const uint8_t *matrix, size_t n_rows, size_t n_sites, size_t row_stride,
              size_t row_begin, size_t row_end, int32_t *dist_out, void *tmp, int *launches) {
    if (n_rows == 0 || row_end <= row_begin) return 0;
    const size_t n_rows_pad = k4_pad_rows(n_rows);
    size_t n_words = k4_words(n_sites);
    const size_t n_words_pad = (n_words + K4_KW - 1) / K4_KW * K4_KW;
    uint32_t *planes = reinterpret_cast<uint32_t *>(tmp);
    if (n_words_pad == 0) {
        if (cudaMemsetAsync(dist_out, 0, (row_end - row_begin) * n_rows * sizeof(int32_t), stream) != cudaSuccess)
            return SNPGPU_E_CUDA;
        return 0;
    }
    if (n_words_pad > 65535) {
        // grid.y limit of the pack kernel: 65535 words = 2.09 M sites per call is far above any SNP matrix the
        // pipeline produces; refuse rather than silently truncate
        return SNPGPU_E_ARG;
    }
    dim3 pg((unsigned)((n_rows_pad + 127) / 128), (unsigned)n_words_pad);
    k4_pack_kernel<<<pg, 128, 0, stream>>>(matrix, n_rows, n_sites, row_stride, n_rows_pad, n_words_pad, planes);
    const int triangle = (row_begin == 0 && row_end == n_rows) ? 1 : 0;
    const size_t i_first = row_begin / K4_TILE, i_last = (row_end + K4_TILE - 1) / K4_TILE;
    dim3 grid((unsigned)(n_rows_pad / K4_TILE), (unsigned)(i_last - i_first));
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k4_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem));
        attr_set = true;
    }
    k4_pairs_kernel<<<grid, K4_THREADS, sizeof(K4Smem), stream>>>(planes, n_rows, n_rows_pad, n_words_pad, row_begin,
                                                                 row_end, triangle, dist_out);
    *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : SNPGPU_E_CUDA;
}

}  // namespace snpgpu
