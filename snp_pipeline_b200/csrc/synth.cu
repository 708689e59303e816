// synth.cu -- synthetic pileup generator (bench / tests only; SURVEY.md section 8d's input spec).
//
// One sample's `samtools mpileup`-style text for positions 1..genome_len of one contig, written straight into
// HBM.  Every random draw is a counter-based hash of (seed, sample, position, slot), so any sample can be
// regenerated on any rank and the host can recompute which pool sites a sample carries.
//   pass 1  length of every line        (one thread per position)
//   scan    exclusive prefix of the lengths (CUB DeviceScan)
//   pass 2  the bytes                   (one thread per position, same draws as pass 1)
// Mix (rates from the bundled lambda-virus pileups): depth ~ Binomial(64, 3/8) scaled to mean_depth and clipped
// to [0, 60]; '.'/',' by strand; 0.7 % substitutions; 0.8 % N/n; 0.15 % '*'; '^'+MAPQ and '$' at 1/130 per
// read; one indel token on 0.1 % of the lines; zero-depth lines at 0.08 %; qualities Phred 13..39.
#include "internal.h"
#include "synth_line.cuh"
#include <cub/device/device_scan.cuh>

namespace snpgpu {

__global__ void synth_len_kernel(const SynthArgs a, unsigned long long *len) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.genome_len) return;
    len[i] = synth_line<false>(a, i + 1u, nullptr);
}

__global__ void synth_write_kernel(const SynthArgs a, const unsigned long long *off, uint8_t *text, size_t cap,
                                   unsigned long long *nbytes) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.genome_len) return;
    unsigned long long o = off[i];
    uint8_t line[448];
    uint32_t n = synth_line<true>(a, i + 1u, line);
    if (i == a.genome_len - 1u) *nbytes = o + n;
    if (o + n > cap) return;
    for (uint32_t k = 0; k < n; k++) text[o + k] = line[k];
}

size_t synth_workspace_bytes(uint32_t genome_len) {
    size_t scan = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                  (int)genome_len);
    return 2 * (((size_t)genome_len * 8 + 255) / 256 * 256) + scan + 512;
}

int synth_launch(cudaStream_t stream, const snpgpu_synth_spec &spec, const char *contig_name, uint8_t *text_dev,
                 size_t cap, unsigned long long *nbytes_dev, void *tmp, size_t tmp_bytes, int *launches) {
    if (spec.genome_len == 0) return SNPGPU_E_ARG;
    SynthArgs a = synth_make_args(spec, contig_name);
    uint8_t *t = reinterpret_cast<uint8_t *>(tmp);
    unsigned long long *len = reinterpret_cast<unsigned long long *>(t);
    size_t o1 = ((size_t)spec.genome_len * 8 + 255) / 256 * 256;
    unsigned long long *off = reinterpret_cast<unsigned long long *>(t + o1);
    size_t o2 = 2 * o1;
    if (tmp_bytes < o2 + 256) return SNPGPU_E_NOMEM;
    size_t scan_bytes = tmp_bytes - o2;
    unsigned grid = (spec.genome_len + 127u) / 128u;
    synth_len_kernel<<<grid, 128, 0, stream>>>(a, len);
    cudaError_t se = cub::DeviceScan::ExclusiveSum(t + o2, scan_bytes, (const unsigned long long *)len, off,
                                                   (int)spec.genome_len, stream);
    if (se != cudaSuccess) return SNPGPU_E_CUDA;
    synth_write_kernel<<<grid, 128, 0, stream>>>(a, off, text_dev, cap, nbytes_dev);
    *launches += 4;
    return cudaGetLastError() == cudaSuccess ? 0 : SNPGPU_E_CUDA;
}

void synth_host_sites(const snpgpu_synth_spec &spec, uint32_t *pos_out, size_t cap, size_t *n_out) {
    SynthArgs a = synth_make_args(spec, nullptr);
    size_t n = 0;
    for (uint32_t pos = 1; pos <= spec.genome_len; pos++) {
        if (synth_carries(a, pos)) {
            if (n < cap) pos_out[n] = pos;
            n++;
        }
    }
    *n_out = n;
}

}  // namespace snpgpu
