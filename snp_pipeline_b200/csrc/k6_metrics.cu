// k6_metrics.cu -- K6: the pileup by-product of collect_metrics: the sum of the raw-depth column.
//
// Replaces the loop of collect_metrics.py:322-329 -- for line in f: tokens = line.split(); depth_sum += int(tokens[3]),
// ValueError / IndexError ignored -- whose result, divided by the reference length, is the sample's "avePileupDepth"
// (collect_metrics.py:333-338).  Exact for any byte input under the package's byte domain (a byte >= 0x80 is reported,
// like in K1): lines end at '\n', '\r' or "\r\n" (universal newlines; the empty line between a '\r' and its '\n' has no
// fourth token and adds nothing), tokens are split on Python's ASCII whitespace, int() takes an optional sign, digits and
// single underscores between digits.
// Shape: one thread per 128 bytes of text.  A thread owns the lines that START in its bytes; for each it walks the text
// forward to the fourth token (about 45 bytes of a pileup line, an L1 hit after the first touch), and it checks its own
// bytes for the domain.  HBM traffic: the text once.  Not on the hot path (the reference reads the file a second time
// for it, too).
#include "internal.h"
#include "line_general.cuh"

namespace snpgpu {

constexpr int K6_CHUNK = 128;

__device__ __forceinline__ bool k6_term(unsigned c) { return c == '\n' || c == '\r'; }

__global__ void __launch_bounds__(256) k6_depth_sum_kernel(const uint8_t *__restrict__ text, unsigned long long nbytes,
                                                           unsigned long long *out /* [0] sum, [1] lines counted, [2] ~first error */) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long lo = t * K6_CHUNK;
    long long sum = 0;
    unsigned long long counted = 0;
    if (lo < nbytes) {
        const unsigned long long hi = lo + K6_CHUNK < nbytes ? lo + K6_CHUNK : nbytes;
        unsigned high = 0;
        for (unsigned long long p = lo; p < hi; p++) {
            const unsigned c = text[p];
            high |= c;
            const bool start = p == 0 ? true : k6_term(text[p - 1]);
            if (!start || k6_term(c)) continue;
            // the line that starts at p: its fourth whitespace-separated token
            unsigned long long i = p;
            int tok = 0;
            for (;;) {
                while (i < nbytes && py_space(text[i]) && !k6_term(text[i])) i++;
                if (i >= nbytes || k6_term(text[i])) break;
                const unsigned long long b = i;
                while (i < nbytes && !py_space(text[i])) i++;
                if (++tok == 4) {
                    int64_t v = 0;
                    const int st = py_int(text + b, (int64_t)(i - b), &v);
                    if (st == ST_OK) { sum += v; counted++; }
                    else if (st == ST_DOMAIN) atomicMax(&out[2], ~((p << 8) | (unsigned long long)ST_DOMAIN));   // beyond int64
                    break;
                }
            }
        }
        if (high & 0x80u) {                                   // which line? the one that holds the first such byte of the chunk
            unsigned long long p = lo;
            while (p < hi && text[p] < 0x80u) p++;
            unsigned long long s = p;
            while (s > 0 && !k6_term(text[s - 1])) s--;
            atomicMax(&out[2], ~((s << 8) | (unsigned long long)ST_DOMAIN));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        counted += __shfl_xor_sync(0xffffffffu, counted, d);
    }
    if ((threadIdx.x & 31) == 0 && counted) {
        atomicAdd(&out[0], (unsigned long long)sum);
        atomicAdd(&out[1], counted);
    }
}

int k6_launch_depth_sum(cudaStream_t stream, const uint8_t *text, size_t nbytes, unsigned long long *out3) {
    if (cudaMemsetAsync(out3, 0, 3 * sizeof(unsigned long long), stream) != cudaSuccess) return -1;
    if (!nbytes) return 0;
    const size_t threads = (nbytes + K6_CHUNK - 1) / K6_CHUNK;
    k6_depth_sum_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(text, nbytes, out3);
    return 1;
}

}  // namespace snpgpu
