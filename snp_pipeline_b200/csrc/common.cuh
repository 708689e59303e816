// common.cuh -- shared device-side structures and PTX helpers of libsnpgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/snpgpu.h"
#include "hd.cuh"
#include "sites.cuh"

namespace snpgpu {

// Device-side status of one pileup call.  first_error orders errors by file offset so the one reported is
// the one the reference would have raised first.
struct PileupStatusDev {
    unsigned long long first_error_inv;   // ~((byte offset of the line << 8) | code), 0 when clean (atomicMax): the
                                          // whole status starts as zeros, like the other per-call scratch
    unsigned long long n_lines;
    unsigned long long n_parsed;
    unsigned long long n_general;
    unsigned long long arena_used;    // bytes of scratch claimed by the exact-splice path
    unsigned int       arena_overflow;
    unsigned int       next_tile;     // the pileup kernel's tile tickets (warps take tiles in increasing order)
    unsigned long long over_used;     // entries claimed in the per-line results' overflow list
};

// Tile geometry of the pileup kernel: every warp is its own pipeline.
#ifndef K1_CFG_CHUNKS          // (tuning builds override these four on the nvcc command line)
#define K1_CFG_CHUNKS 21
#define K1_CFG_CTAS 4
#define K1_CFG_LHCAP 10
#define K1_CFG_QCAP 64
#endif
constexpr int K1_LANE_CHUNKS = K1_CFG_CHUNKS;             // 16-byte chunks one lane scans for newlines ...
constexpr int K1_LANE_BYTES  = 16 * K1_LANE_CHUNKS;   // ... an odd number: stride = 4 mod 8 words, quarter warps hit disjoint banks
constexpr int K1_TILE     = 32 * K1_LANE_BYTES;       // bytes of text whose line starts one tile owns
#ifndef K1_CFG_LOOK
#define K1_CFG_LOOK 896
#define K1_CFG_WCAP 192
#endif
constexpr int K1_LOOK     = K1_CFG_LOOK;              // extra bytes staged so that the last owned line is complete
constexpr int K1_PAD      = 32;                // '\n' sentinels after the staged bytes (word over-reads land here)
constexpr int K1_WARPS    = 4;                 // independent warps per CTA
constexpr int K1_THREADS  = 32 * K1_WARPS;
constexpr int K1_CTAS_PER_SM = K1_CFG_CTAS;              // 16 warps x 13.6 KiB of shared memory per SM, 128 registers per thread
constexpr int K1_WCAP     = K1_CFG_WCAP;               // line starts a warp lists per pass (more -> another pass)
constexpr int K1_LHCAP    = K1_CFG_LHCAP;                // line starts one lane lists per tile (more -> byte-wise path)
constexpr int K1_ORDER_TILES = 512;            // tiles per group of the ordering pass (k1_tile_prefix_kernel)
constexpr int K1_STAGE_CAP = 320;              // per-line results a tile keeps in its row of the staging array (a tile of the
                                               // scan's own path owns at most 32 (K1_LHCAP - 1) + 1 lines; more -> overflow list)
constexpr int K1_NAMEW    = 16;                // words of the expected contig's name a warp keeps in shared memory
constexpr int K1_DRAIN_AT = 24;                // queued lines that trigger a drain between two tiles (32: also inside a tile)
constexpr int K1_QCAP     = K1_CFG_QCAP;                // per-warp queue slots (drained whenever 32 are filled)

struct PileupArgs {
    const uint8_t      *text;          // 16-byte aligned
    unsigned long long  nbytes;
    SiteTable           sites;
    CallParams          p;
    int                 mode;          // SNPGPU_MODE_SITES / SNPGPU_MODE_ALL
    int                 n_tiles;
    unsigned long long *site_cells;    // n_unique, zero-initialised: ((line offset + 1) << 8) | cell  (atomicMax:
                                       // the last line in file order wins, like the dict of call_consensus.py:169)
    uint16_t           *line_out;      // null, or one uint16 per line in file order: cell | fail << 8
    unsigned long long  line_out_cap;
    uint32_t           *tile_lines;    // [n_tiles]: lines each tile owns (written by the tile's warp, when line_out)
    unsigned long long *group_lines;   // zero-initialised: lines per group of K1_ORDER_TILES tiles (atomic adds)
    uint16_t           *stage;         // [n_tiles][K1_STAGE_CAP]: per-line results, tile by tile (when line_out)
    unsigned long long *over;          // overflow list of (tile << 32 | index << 16 | result) for tiles above K1_STAGE_CAP ...
    unsigned long long  over_cap;      // ... its capacity (PileupStatusDev::over_used counts the claims)
    unsigned long long *tile_first;    // [n_tiles]: file-order index of a tile's first line (k1_tile_prefix_kernel)
    unsigned long long *rec_off;       // null, or: file offset of every parsed line is appended here (any order) ...
    unsigned long long *rec_count;     // ... through this counter (the consensus-VCF pass, k5_vcf.cu)
    unsigned long long  rec_cap;
    PileupStatusDev    *st;
    uint8_t            *arena;
    unsigned long long  arena_cap;
};

// ---- PTX: mbarrier + 1-D bulk async copy (TMA engine, UBLKCP in SASS) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// waits with a short sleep between polls, so that a waiting warp leaves the issue slots to the others
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(200);
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared, bytes a multiple of 16, both addresses 16-byte aligned; completion lands on bar
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// atomicAdd(p, 1) whose result may stay in flight: written as PTX so that the compiler does not turn it into its
// warp-aggregated form, which shuffles the result out right away and so waits for it
__device__ __forceinline__ uint32_t atom_inc_u32(unsigned int *p) {
    uint32_t old;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p) : "memory");
    return old;
}

// pull [src, src + bytes) into L2 ahead of the bulk copy that will want it (bytes a multiple of 16)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

}  // namespace snpgpu
