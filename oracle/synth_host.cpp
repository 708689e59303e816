// oracle/synth_host.cpp -- TEST INFRASTRUCTURE (not product): the synthetic pileup of bench.py written on the HOST.
//
// bench.py --impl reference times the CPU restatement of the reference's path; its inputs must not come out of the GPU
// library, so this file compiles the generator's own line function (snp_pipeline_b200/csrc/synth_line.cuh, host/device)
// for the host and writes the very same bytes snpgpu_synth_pileup_dev writes into HBM (tests/test_gpu_parity.py compares
// the two).  Built by oracle.build_synth() into oracle/_build/libsynthhost.so.
#include "../snp_pipeline_b200/csrc/synth_line.cuh"
#include <thread>
#include <vector>

using namespace snpgpu;

extern "C" {

// the text of one sample; returns its length in bytes (nothing past `cap` is written: call again with a larger buffer)
unsigned long long synth_host_pileup(const snpgpu_synth_spec *spec, const char *contig_name, uint8_t *out,
                                     unsigned long long cap, int threads) {
    const SynthArgs a = synth_make_args(*spec, contig_name);
    const uint32_t G = a.genome_len;
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > G) threads = (int)(G ? G : 1u);
    std::vector<unsigned long long> part((size_t)threads + 1, 0ull);
    auto lo = [&](int t) { return (uint32_t)((unsigned long long)G * (unsigned)t / (unsigned)threads); };
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] {
                unsigned long long n = 0;
                for (uint32_t i = lo(t); i < lo(t + 1); i++) n += synth_line<false>(a, i + 1u, nullptr);
                part[(size_t)t + 1] = n;
            });
        for (auto &th : pool) th.join();
    }
    for (int t = 0; t < threads; t++) part[(size_t)t + 1] += part[(size_t)t];
    const unsigned long long total = part[(size_t)threads];
    if (total > cap || !out) return total;
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] {
                unsigned long long o = part[(size_t)t];
                for (uint32_t i = lo(t); i < lo(t + 1); i++) o += synth_line<true>(a, i + 1u, out + o);
            });
        for (auto &th : pool) th.join();
    }
    return total;
}

// the pool sites the sample carries (1-based positions, ascending); returns their number
unsigned long long synth_host_sites(const snpgpu_synth_spec *spec, uint32_t *pos_out, unsigned long long cap) {
    const SynthArgs a = synth_make_args(*spec, nullptr);
    unsigned long long n = 0;
    for (uint32_t pos = 1; pos <= a.genome_len; pos++)
        if (synth_carries(a, pos)) {
            if (n < cap && pos_out) pos_out[n] = pos;
            n++;
        }
    return n;
}

}  // extern "C"
