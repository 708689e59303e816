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
    return 3 * w * k4_pad_rows(n_rows) * sizeof(uint32_t) + 256;     // (K4_TILE rows and K4_KW words are multiples of the pack tile's 32)
}

// plane bits of one byte: bit 0 valid (A C G T, either case), bit 8 code bit 0, bit 16 code bit 1.
// (c & 0xdf) == 'A' only for 'A' and 'a' (bit 5 is the only one cleared), likewise C G T: no other byte lands on them.
// (c >> 1) & 3:  A -> 0, C -> 1, T -> 2, G -> 3
__host__ __device__ inline uint32_t k4_byte_entry(unsigned o) {
    const unsigned c = o & 0xdfu;                                          // upper() for letters (utils.py:1153-1154)
    const unsigned ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T');
    const unsigned code = (c >> 1) & 3u;
    return ok | ((ok & code & 1u) << 8) | ((ok & (code >> 1)) << 16);
}

// generic pack (any alignment): one thread per (row, word), byte loads
__global__ void k4_pack_bytes_kernel(const uint8_t *__restrict__ matrix, size_t n_rows, size_t n_sites, size_t row_stride,
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
            const uint32_t e = k4_byte_entry(p[i]);
            v |= (e & 1u) << i;
            lo |= ((e >> 8) & 1u) << i;
            hi |= ((e >> 16) & 1u) << i;
        }
    }
    const size_t plane = n_words_pad * n_rows_pad;
    planes[0 * plane + word * n_rows_pad + row] = v;
    planes[1 * plane + word * n_rows_pad + row] = lo;
    planes[2 * plane + word * n_rows_pad + row] = hi;
}

// Fast pack (matrix and row_stride 16-byte aligned): a CTA turns 32 rows x 32 words (1024 sites).  A warp takes one row
// at a time: lane l loads the 32 bytes of word l as two 16-byte loads (one coalesced 1 KiB run per warp), looks every
// byte up in a 256-entry table in shared memory (8 bytes share one accumulator: acc = 2 acc + entry keeps the three
// planes' bits in three byte lanes), and parks the three words in a padded shared tile; then every warp writes one
// word's 32 rows as one 128-byte run into the word-major planes.
constexpr int K4_PACK_THREADS = 256;
__global__ void __launch_bounds__(K4_PACK_THREADS) k4_pack_kernel(const uint8_t *__restrict__ matrix, size_t n_rows, size_t n_sites,
                                                                  size_t row_stride, size_t n_rows_pad, size_t n_words_pad,
                                                                  uint32_t *__restrict__ planes) {
    __shared__ uint32_t tab[256];
    __shared__ uint32_t tile[3][32][33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    tab[tid] = k4_byte_entry((unsigned)tid);
    __syncthreads();
    const size_t row0 = (size_t)blockIdx.x * 32, word0 = (size_t)blockIdx.y * 32;
    for (int r = warp; r < 32; r += K4_PACK_THREADS / 32) {
        const size_t row = row0 + r, site0 = (word0 + lane) * 32;
        uint32_t pl[3] = {0u, 0u, 0u};
        if (row < n_rows && site0 < n_sites) {
            const uint8_t *p = matrix + row * row_stride + site0;
            uint32_t w[8];
            if (site0 + 32 <= n_sites) {
                const uint4 x = *reinterpret_cast<const uint4 *>(p), y = *reinterpret_cast<const uint4 *>(p + 16);
                w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
            } else {                                                       // the row's last, partial word
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    uint32_t v = 0;
                    for (int b = 0; b < 4; b++)
                        if (site0 + 4 * k + b < n_sites) v |= (uint32_t)p[4 * k + b] << (8 * b);
                    w[k] = v;                                              // (bytes behind the row: 0, not a base)
                }
            }
#pragma unroll
            for (int g = 0; g < 4; g++) {                                  // 8 bytes per accumulator, highest byte first
                uint32_t acc = 0;
#pragma unroll
                for (int b = 7; b >= 0; b--) acc = acc * 2u + tab[(w[2 * g + (b >> 2)] >> (8 * (b & 3))) & 0xffu];
                pl[0] |= (acc & 0xffu) << (8 * g);
                pl[1] |= ((acc >> 8) & 0xffu) << (8 * g);
                pl[2] |= ((acc >> 16) & 0xffu) << (8 * g);
            }
        }
        tile[0][lane][r] = pl[0]; tile[1][lane][r] = pl[1]; tile[2][lane][r] = pl[2];
    }
    __syncthreads();
    const size_t plane = n_words_pad * n_rows_pad;
    for (int k = warp; k < 32; k += K4_PACK_THREADS / 32) {
#pragma unroll
        for (int p = 0; p < 3; p++) planes[p * plane + (word0 + k) * n_rows_pad + row0 + lane] = tile[p][k][lane];
    }
}

constexpr int K4_MAX_LIST = 480;   // tile rows one launch of the list form takes
struct K4TileList {                // which 64-row tile rows a launch computes: a stripe (n == 0) or a list
    int n;
    uint16_t t[K4_MAX_LIST];
};

struct K4Smem {
    uint32_t a[3][K4_KW][K4_TILE];     // [plane][word][row of the i tile]
    uint32_t b[3][K4_KW][K4_TILE];
};

__global__ void __launch_bounds__(K4_THREADS) k4_pairs_kernel(const uint32_t *__restrict__ planes, size_t n_rows,
                                                             size_t n_rows_pad, size_t n_words_pad, size_t row_begin,
                                                             size_t row_end, int triangle, int upper_only, int slabs_per_cta,
                                                             const __grid_constant__ K4TileList list,
                                                             int32_t *__restrict__ dist) {
    extern __shared__ __align__(16) uint8_t k4_smem_raw[];
    K4Smem &sm = *reinterpret_cast<K4Smem *>(k4_smem_raw);
    // list form: tile row list.t[y], its 64 output rows at y * 64; stripe form: tile rows from row_begin's on, output row
    // i - row_begin
    const size_t i0 = list.n ? (size_t)list.t[blockIdx.y] * K4_TILE : row_begin / K4_TILE * K4_TILE + (size_t)blockIdx.y * K4_TILE;
    const size_t out0 = list.n ? i0 - (size_t)blockIdx.y * K4_TILE : row_begin;      // output row of matrix row i: i - out0
    if (list.n) { row_begin = i0; row_end = i0 + K4_TILE; }
    const size_t j0 = (size_t)blockIdx.x * K4_TILE;
    if ((triangle || upper_only) && j0 < i0) return;       // (triangle: the mirror tile writes both halves)
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const size_t plane = n_words_pad * n_rows_pad;
    uint32_t acc[4][4];
#pragma unroll
    for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) acc[x][y] = 0;

    const size_t w_begin = (size_t)blockIdx.z * slabs_per_cta * K4_KW;
    const size_t w_stop = w_begin + (size_t)slabs_per_cta * K4_KW < n_words_pad ? w_begin + (size_t)slabs_per_cta * K4_KW : n_words_pad;
    const bool split = gridDim.z > 1;                      // the words are dealt to several CTAs: they add into zeroed cells
    for (size_t w0 = w_begin; w0 < w_stop; w0 += K4_KW) {
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
            if (split) {
                if (i >= row_begin && i < row_end) atomicAdd(&dist[(i - out0) * n_rows + j], d);
                if (triangle && j0 != i0 && j >= row_begin && j < row_end) atomicAdd(&dist[(j - out0) * n_rows + i], d);
            } else {
                if (i >= row_begin && i < row_end) dist[(i - out0) * n_rows + j] = d;
                if (triangle && j0 != i0 && j >= row_begin && j < row_end) dist[(j - out0) * n_rows + i] = d;
            }
        }
    }
}

// bytes -> planes, once per matrix
static int k4_pack(cudaStream_t stream, const uint8_t *matrix, size_t n_rows, size_t n_sites, size_t row_stride,
                   size_t n_rows_pad, size_t n_words_pad, uint32_t *planes) {
    if (n_words_pad > 65535 * 32) return SNPGPU_E_ARG;     // (grid.y of the pack kernel: 67 M sites)
    if ((((uintptr_t)matrix | (uintptr_t)row_stride) & 15u) == 0) {
        dim3 pg((unsigned)(n_rows_pad / 32), (unsigned)(n_words_pad / 32));
        k4_pack_kernel<<<pg, K4_PACK_THREADS, 0, stream>>>(matrix, n_rows, n_sites, row_stride, n_rows_pad, n_words_pad, planes);
    } else {
        if (n_words_pad > 65535) return SNPGPU_E_ARG;
        dim3 pg((unsigned)((n_rows_pad + 127) / 128), (unsigned)n_words_pad);
        k4_pack_bytes_kernel<<<pg, 128, 0, stream>>>(matrix, n_rows, n_sites, row_stride, n_rows_pad, n_words_pad, planes);
    }
    return 0;
}

// Stripe form (tiles == nullptr): rows [row_begin, row_end) x all columns (the whole matrix: the upper triangle + mirror).
// List form: the 64-row tile rows tiles[0 .. n_tiles), upper part only -- of each only the cells in and to the right of
// its diagonal tile are written (the others are left as they are), 64 output rows per entry; the multi-GPU driver deals
// whole tile rows to the ranks and mirrors.
int k4_launch(cudaStream_t stream, const uint8_t *matrix, size_t n_rows, size_t n_sites, size_t row_stride,
              size_t row_begin, size_t row_end, const uint32_t *tiles, size_t n_tiles, int32_t *dist_out, void *tmp,
              int n_sms, int *launches) {
    const bool list_form = tiles != nullptr;
    if (n_rows == 0 || (!list_form && row_end <= row_begin) || (list_form && n_tiles == 0)) return 0;
    const size_t n_rows_pad = k4_pad_rows(n_rows);
    size_t n_words = k4_words(n_sites);
    const size_t n_words_pad = (n_words + K4_KW - 1) / K4_KW * K4_KW;
    uint32_t *planes = reinterpret_cast<uint32_t *>(tmp);
    const size_t out_rows = list_form ? n_tiles * K4_TILE : row_end - row_begin;
    if (n_words_pad == 0) {
        if (cudaMemsetAsync(dist_out, 0, out_rows * n_rows * sizeof(int32_t), stream) != cudaSuccess) return SNPGPU_E_CUDA;
        return 0;
    }
    if (int rc = k4_pack(stream, matrix, n_rows, n_sites, row_stride, n_rows_pad, n_words_pad, planes)) return rc;
    *launches += 1;
    const int triangle = (!list_form && row_begin == 0 && row_end == n_rows) ? 1 : 0;
    const size_t n_slabs = n_words_pad / K4_KW;
    const size_t gx = n_rows_pad / K4_TILE;
    cudaFuncSetAttribute(k4_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem));   // (per device; cheap)
    for (size_t done = 0; done < (list_form ? n_tiles : 1); done += K4_MAX_LIST) {
        K4TileList list;
        list.n = 0;
        size_t gy;
        if (list_form) {
            gy = n_tiles - done < (size_t)K4_MAX_LIST ? n_tiles - done : (size_t)K4_MAX_LIST;
            list.n = (int)gy;
            for (size_t k = 0; k < gy; k++) {
                if ((size_t)tiles[done + k] * K4_TILE >= n_rows || tiles[done + k] > 65535u) return SNPGPU_E_ARG;
                list.t[k] = (uint16_t)tiles[done + k];
            }
        } else {
            gy = (row_end + K4_TILE - 1) / K4_TILE - row_begin / K4_TILE;
        }
        // few tiles (a small batch, one tile row): the words are dealt to several CTAs so that every SM has work
        size_t n_ctas = gx * gy;
        if (triangle || list_form) n_ctas = n_ctas / 2 + 1;
        size_t ksplit = n_ctas >= (size_t)(3 * n_sms) ? 1 : ((size_t)(3 * n_sms) + n_ctas - 1) / n_ctas;
        if (ksplit > n_slabs) ksplit = n_slabs;
        if (ksplit > 65535) ksplit = 65535;
        const size_t slabs_per_cta = (n_slabs + ksplit - 1) / ksplit;
        ksplit = (n_slabs + slabs_per_cta - 1) / slabs_per_cta;
        int32_t *out = dist_out + done * K4_TILE * n_rows;
        if (ksplit > 1) {
            if (cudaMemsetAsync(out, 0, (list_form ? gy * K4_TILE : out_rows) * n_rows * sizeof(int32_t), stream) != cudaSuccess) return SNPGPU_E_CUDA;
            *launches += 1;
        }
        dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)ksplit);
        k4_pairs_kernel<<<grid, K4_THREADS, sizeof(K4Smem), stream>>>(planes, n_rows, n_rows_pad, n_words_pad, row_begin, row_end, triangle,
                                                                     list_form ? 1 : 0, (int)slabs_per_cta, list, out);
        *launches += 1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : SNPGPU_E_CUDA;
}

}  // namespace snpgpu
