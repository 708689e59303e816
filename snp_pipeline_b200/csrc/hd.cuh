// hd.cuh -- host/device portability shims of libsnpgpu.
//
// The per-line parsers (line_general.cuh, line_fast.cuh) are written once and compiled twice: by nvcc for
// sm_100a (the product) and by g++ for tests/cpu_sim (a debugging harness that checks the very same
// functions against the oracle where no GPU exists; it is never loaded by the product package).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define SNP_HD __host__ __device__ __forceinline__
#define SNP_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define SNP_HD inline
#define SNP_HD_NOINLINE inline
#endif

namespace snpgpu {

// status codes (mirror include/snpgpu.h; kept numeric here so the header stays self-contained)
enum : int { ST_OK = 0, ST_VALUE = 1, ST_INDEX = 2, ST_UNPACK = 3, ST_DOMAIN = 4, ST_LONECR = 5 };
// internal outcomes that are not errors: the general parser wants scratch space / the fast parser declines
// the line / the line is not at a wanted site (SITES mode)
enum : int { ST_NEED_ARENA = 64, ST_FALLBACK = 65, ST_SKIP = 66 };

enum : uint8_t { FAIL_RAWDPTH = 1, FAIL_VARFREQ = 2, FAIL_DEPTH = 4, FAIL_STRDPTH = 8, FAIL_STRBIAS = 16,
                 FAIL_REGION = 32 };

struct CallParams {
    int32_t min_base_qual;
    int32_t min_cons_depth;
    int32_t min_cons_strand_depth;
    int32_t pad;
    double  min_cons_freq;
    double  min_cons_strand_bias;
};

SNP_HD unsigned up8(unsigned c) { return (c - 'a' < 26u) ? c - 32u : c; }
SNP_HD unsigned low8(unsigned c) { return (c - 'A' < 26u) ? c + 32u : c; }
// Python str.isspace() restricted to ASCII (what str.split() / rstrip() treat as separators)
SNP_HD bool py_space(unsigned c) { return c == 0x20u || (c - 9u) < 5u || (c - 0x1cu) < 4u; }
SNP_HD bool is_digit(unsigned c) { return (c - '0') < 10u; }

// 32-bit funnel shift right by (sh & 31) in {0, 8, 16, 24} bits: the little-endian word that starts sh/8 bytes into lo
SNP_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31u;
    return sh ? (lo >> sh) | (hi << (32u - sh)) : lo;
#endif
}
// the 4 bytes at buf[off .. off+3] of a 4-byte aligned buffer, as a little-endian word (reads up to buf[off+7])
SNP_HD uint32_t load_u32(const uint8_t *buf, uint32_t off) {
    const uint32_t *p = reinterpret_cast<const uint32_t *>(buf + (off & ~3u));
    return funnel_r(p[0], p[1], (off & 3u) * 8u);
}
SNP_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// index of the lowest set bit (x != 0)
SNP_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// pileup.py:556-584 on one allele's counts.  good = Record.good_depth, cons/fwd/rev = counts of the
// consensus base.  The two threshold products are IEEE double multiplies, as in the Python.
SNP_HD uint8_t filter_mask(uint32_t good, uint32_t cons, uint32_t fwd, uint32_t rev, const CallParams &p) {
    uint8_t m = 0;
    if ((double)cons < (double)good * p.min_cons_freq) m |= FAIL_VARFREQ;
    if ((int64_t)cons < (int64_t)p.min_cons_depth) m |= FAIL_DEPTH;
    if ((int64_t)fwd < (int64_t)p.min_cons_strand_depth || (int64_t)rev < (int64_t)p.min_cons_strand_depth)
        m |= FAIL_STRDPTH;
    double msb = (double)cons * p.min_cons_strand_bias;
    if ((double)fwd < msb || (double)rev < msb) m |= FAIL_STRBIAS;
    return m;
}

}  // namespace snpgpu
