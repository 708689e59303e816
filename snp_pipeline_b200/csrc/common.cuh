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
    unsigned long long n_called;      // snplist positions that received a call (the finish kernel counts them) ...
    unsigned int       finish_done;   // ... its blocks that are through: the last one writes the caller's stats
};

// ---- geometry of the pileup kernel (k1_pileup.cu): every warp is its own pipeline --------------
#ifndef K1_CFG_WARPS           // (tuning builds override these on the nvcc command line)
#define K1_CFG_WARPS 20
#define K1_CFG_LANE_BYTES 336
#define K1_CFG_LOOK 208
#endif
#define K1_CFG_CTAS 1
constexpr int K1_WARPS       = K1_CFG_WARPS;         // independent warps per CTA; one CTA per SM (it shares the filter tables)
constexpr int K1_THREADS     = 32 * K1_WARPS;
constexpr int K1_CTAS_PER_SM = K1_CFG_CTAS;
constexpr int K1_LANE_BYTES  = K1_CFG_LANE_BYTES;    // bytes of a tile one lane owns: lines whose preceding '\n' lies there are its
constexpr int K1_TILE        = 32 * K1_LANE_BYTES;   // bytes of text whose line starts one tile (= one warp-step) owns
constexpr int K1_LOOK        = K1_CFG_LOOK;          // extra bytes staged so that the last owned line is complete
constexpr int K1_WIN         = K1_TILE + K1_LOOK;
constexpr int K1_PAD         = 32;                   // '\n' sentinels after the staged bytes (word over-reads land here)
#ifndef K1_CFG_LCAP
#define K1_CFG_LCAP 8
#endif
constexpr int K1_LCAP        = K1_CFG_LCAP;          // per-line results a lane keeps in the staging array per tile (more -> overflow list)
constexpr int K1_ORDER_TILES = 512;                  // tiles per group of the ordering pass (k1_tile_prefix_kernel)
constexpr int K1_NAMEW       = 16;                   // words of the expected contig's name a warp keeps in shared memory
#ifndef K1_CFG_BATCH
#define K1_CFG_BATCH 64
#endif
constexpr int K1_BATCH       = K1_CFG_BATCH;         // samples one launch of the pileup kernel takes (the descriptors are kernel parameters: 10 KB of the 32 KB)
static_assert(K1_LANE_BYTES % 16 == 0 && K1_LANE_BYTES <= 1008 && K1_WIN % 16 == 0, "tile geometry");

// What the second / third parser tiers (line_fast.cuh, line_general.cuh) see of one sample: built per line by the
// follow-up kernel from the batch descriptor below.
struct PileupArgs {
    const uint8_t      *text;          // 16-byte aligned
    unsigned long long  nbytes;
    SiteTable           sites;
    CallParams          p;
    int                 mode;          // SNPGPU_MODE_SITES / SNPGPU_MODE_ALL
    unsigned long long *site_cells;    // n_unique, zero-initialised: ((line offset + 1) << 8) | cell  (atomicMax:
                                       // the last line in file order wins, like the dict of call_consensus.py:169)
    uint16_t           *stage;         // null, or [n_tiles][K1_LCAP][32]: per-line results, slot k of lane L of a tile
    unsigned long long *over;          // overflow list of (tile * 32 + lane) << 32 | k << 16 | result for lanes above K1_LCAP lines ...
    unsigned long long  over_cap;      // ... its capacity (PileupStatusDev::over_used counts the claims)
    unsigned long long *rec_off;       // null, or: file offset of every parsed line is appended here (any order) ...
    unsigned long long *rec_count;     // ... through this counter (the consensus-VCF pass, k5_vcf.cu)
    unsigned long long  rec_cap;
    PileupStatusDev    *st;
    uint8_t            *arena;
    unsigned long long  arena_cap;
    PileupStatusDev    *arena_st;      // whose arena_used / arena_overflow count the claims (the arena is shared by a batch)
};

// One sample of a batch, as the kernels see it
struct K1Samp {
    const uint8_t      *text;          // 16-byte aligned
    unsigned long long  nbytes;
    int                 tile0;         // first ticket of the batch that belongs to this sample
    int                 n_tiles;
    unsigned long long *site_cells;    // n_unique + 1, zero-initialised
    PileupStatusDev    *st;            // zero-initialised
    uint16_t           *stage;         // null (no per-line results wanted), or [n_tiles][K1_LCAP][32]
    uint8_t            *lane_lines;    // [n_tiles][32]: lines each lane owns
    uint32_t           *tile_lines;    // [n_tiles]
    unsigned long long *group_lines;   // zero-initialised: lines per group of K1_ORDER_TILES tiles
    unsigned long long *tile_first;    // [n_tiles]: file-order index of a tile's first line (k1_tile_prefix_kernel)
    unsigned long long *over;          // overflow list of this sample, over_cap entries
    uint16_t           *line_out;      // one uint16 per line in file order: cell | fail << 8
    unsigned long long  line_out_cap;
    unsigned long long *rec_off;       // see PileupArgs
    unsigned long long *rec_count;
    unsigned long long  rec_cap;
    uint8_t            *row_out;       // n_snp bytes, snplist order
    unsigned long long  n_unique;      // unique sites of the table
    snpgpu_pileup_stats *stats_out;    // nullable
};

struct K1Batch {
    K1Samp              s[K1_BATCH];   // tile0 ascending
    int                 n_samples;
    int                 total_tiles;
    int                 mode;
    int                 has_qual;      // min_base_qual > 0: line_fast.cuh pairs every base with its quality
    int                 all_rest;      // every line that is parsed goes to the follow-up kernel (has_qual, or the VCF pass's line list is wanted)
    uint32_t            one;           // 1: the multiplier of line_quick3.cuh's IMAD adds (a value the compiler cannot fold)
    unsigned int       *next_tile;     // zero-initialised ticket counter
    unsigned long long *queue;         // lines for the follow-up kernel: 2 words per entry (k1_entry, sample | flags << 32)
    unsigned long long  queue_cap;     // ... entries
    unsigned long long *queue_count;   // zero-initialised; may run past queue_cap (reported as SNPGPU_E_NOMEM)
    unsigned long long  over_cap;
    uint8_t            *arena;         // splice scratch of line_general.cuh, shared by the batch (claimed through s[0].st)
    unsigned long long  arena_cap;
    const int32_t      *snp_unique;    // n_snp: unique-site index of snplist entry k
    unsigned long long  n_snp;
    SiteTable           sites;
    CallParams          p;
};
static_assert(sizeof(K1Batch) <= 32000, "the batch descriptor travels as a kernel parameter (32 764 bytes at most since CUDA 12.1)");

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
#ifndef K1_CFG_SLEEP
#define K1_CFG_SLEEP 200
#endif
    while (!mbar_try_wait(bar, parity)) { if (K1_CFG_SLEEP) __nanosleep(K1_CFG_SLEEP); }
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
    // (inc, not add: ptxas turns an add at a uniform address into a warp-aggregated one whose result is shuffled out -- and
    //  so waited for -- on the spot; the ticket is wanted a whole tile later)
    asm volatile("atom.global.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(old) : "l"(p) : "memory");
    return old;
}

// pull [src, src + bytes) into L2 ahead of the bulk copy that will want it (bytes a multiple of 16)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

}  // namespace snpgpu
