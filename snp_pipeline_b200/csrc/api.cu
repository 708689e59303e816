// api.cu -- the C ABI of libsnpgpu.so (include/snpgpu.h): contexts, the device-resident site table, and the
// host-side sequencing of the kernels.  No kernels of its own.
#include "internal.h"
#include "sites_host.h"

#include <algorithm>
#include <new>
#include <stdio.h>
#include <string>
#include <string.h>
#include <utility>
#include <vector>

using namespace snpgpu;

namespace {

struct DevBuf {
    void  *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = (n + (n >> 3) + 4095) & ~(size_t)4095;       // a little slack so growth is rare
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct snpgpu_ctx {
    int          device = 0;
    int          n_sms = 148;
    int          k1_blocks = 1;              // resident CTAs per SM of the pileup kernel
    cudaStream_t own_stream = nullptr, stream = nullptr;
    uint64_t     launches = 0;
    std::string  err;
    DevBuf k1_zero, tile_first, tile_lines, lane_lines, stage, over, arena, queue, stats;   // k1_zero: counters | per sample status, site cells: one memset per batch
    DevBuf text, row, lines;                  // staging of the host-buffer entry points
    size_t text_nbytes = 0;                   // bytes of the last text snpgpu_pileup_consensus() staged ...
    bool   text_valid = false;                // ... still there (snpgpu_pileup_vcf_records works on it)
    size_t vcf_text_bytes = 0;                // formatted data lines kept in vcf_text for a second call with enough room
    size_t vcf_text_recs = 0;
    uint64_t vcf_text_key = 0;                // what it was formatted with (the arguments' hash)
    bool   vcf_text_valid = false;
    bool   want_rec = false;                  // snpgpu_pileup_want_vcf_records: list the parsed lines during the call itself
    bool   rec_valid = false;                 // the last snpgpu_pileup_consensus() left its line list in rec_off ...
    size_t rec_listed = 0;                    // ... this many entries
    int    rec_mode = 0;
    DevBuf rec_off, rec_sorted, rec_out, alt_out, k5_tmp, k5_state, vcf_text;
    size_t rec_cap = 0;                       // entries of rec_off
    DevBuf k2_tmp, k2_keys, k2_samp, k2_uniq, k2_cnt, k2_out, k2_n;
    DevBuf k4_tmp, k4_mat, k4_dist;
    DevBuf synth_tmp, synth_n;
    DevBuf k3_tmp;
    // pipelined host-buffer calls (snpgpu_pileup_consensus_begin / _end): two child contexts, each with its own stream,
    // staging buffers and kernel scratch, so that one call's copy in overlaps the other's kernels and copies out
    struct Pending {
        bool active = false;
        const void *text = nullptr; size_t nbytes = 0; const snpgpu_sites *sites = nullptr; snpgpu_params params;
        int mode = 0; uint8_t *row_out = nullptr; uint16_t *line_out = nullptr; size_t line_out_cap = 0;
        snpgpu_pileup_stats *stats = nullptr;
    };
    snpgpu_ctx *lane[2] = {nullptr, nullptr};
    Pending pending[2];
    snpgpu_pileup_stats *lane_stats[2] = {nullptr, nullptr};   // pinned
    cudaEvent_t lane_event = nullptr;
    std::vector<std::pair<void *, size_t>> sites_pool;   // blobs of destroyed device-built site tables, reused in stream order
    size_t arena_want = 1 << 20;
    size_t over_want = 1 << 16;              // entries of a sample's overflow list of per-line results (lanes with many tiny lines)
    size_t queue_want = 0;                   // entries of the follow-up kernel's queue, once a batch has asked for more
    bool   timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed[2];     // pending event pairs per kernel id
    std::vector<cudaEvent_t> spare_events;
};

namespace {
struct TimedLaunch {       // brackets one launch with events when timing is on
    snpgpu_ctx *ctx; int kernel; cudaEvent_t a = nullptr, b = nullptr;
    cudaEvent_t get() {
        if (!ctx->spare_events.empty()) { cudaEvent_t e = ctx->spare_events.back(); ctx->spare_events.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    TimedLaunch(snpgpu_ctx *c, int k) : ctx(c), kernel(k) {
        if (ctx->timing) { a = get(); b = get(); cudaEventRecord(a, ctx->stream); }
    }
    ~TimedLaunch() {
        if (a) { cudaEventRecord(b, ctx->stream); ctx->timed[kernel].push_back({a, b}); }
    }
};
}  // namespace

struct snpgpu_sites {
    snpgpu_ctx *ctx = nullptr;
    size_t n_snp = 0, n_unique = 0;
    int n_contigs = 0;
    void *blob = nullptr;                     // one device allocation holding every array below
    size_t blob_bytes = 0;
    bool pooled = false;                      // the blob goes back to the context's pool when the table is destroyed
    SiteTable table;                          // device pointers
    int32_t *snp_unique = nullptr;            // device, n_snp: unique-site index of snplist entry k
};

static int fail(snpgpu_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess) {
    if (ctx) {
        ctx->err = what;
        if (e != cudaSuccess) { ctx->err += ": "; ctx->err += cudaGetErrorString(e); }
    }
    return code;
}

#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? SNPGPU_E_NOMEM : SNPGPU_E_CUDA, #call, e_); \
    } while (0)

extern "C" {

int snpgpu_abi_version(void) { return SNPGPU_ABI_VERSION; }

int snpgpu_create(int device, snpgpu_ctx **out) {
    if (!out) return SNPGPU_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) return SNPGPU_E_CUDA;   // no CPU fallback: fail loudly
    if (device < 0 || device >= n) return SNPGPU_E_ARG;
    snpgpu_ctx *ctx = new (std::nothrow) snpgpu_ctx();
    if (!ctx) return SNPGPU_E_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return SNPGPU_E_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return SNPGPU_E_CUDA; }
    ctx->n_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SNPGPU_E_CUDA; }
    ctx->stream = ctx->own_stream;
    ctx->k1_blocks = std::max(1, k1_blocks_per_sm());
    if (cudaGetLastError() != cudaSuccess) { cudaStreamDestroy(ctx->own_stream); delete ctx; return SNPGPU_E_CUDA; }
    *out = ctx;
    return SNPGPU_OK;
}

void snpgpu_destroy(snpgpu_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int k = 0; k < 2; k++) {
        if (ctx->lane[k]) snpgpu_destroy(ctx->lane[k]);
        if (ctx->lane_stats[k]) cudaFreeHost(ctx->lane_stats[k]);
    }
    if (ctx->lane_event) cudaEventDestroy(ctx->lane_event);
    DevBuf *all[] = {&ctx->k1_zero, &ctx->tile_first, &ctx->tile_lines, &ctx->lane_lines, &ctx->stage, &ctx->over, &ctx->arena, &ctx->queue,
                     &ctx->stats, &ctx->text, &ctx->row, &ctx->lines, &ctx->rec_off, &ctx->rec_sorted, &ctx->rec_out, &ctx->alt_out, &ctx->vcf_text,
                     &ctx->k5_tmp, &ctx->k5_state, &ctx->k2_tmp, &ctx->k2_keys, &ctx->k2_samp,
                     &ctx->k2_uniq, &ctx->k2_cnt, &ctx->k2_out, &ctx->k2_n, &ctx->k4_tmp, &ctx->k4_mat, &ctx->k4_dist,
                     &ctx->synth_tmp, &ctx->synth_n, &ctx->k3_tmp};
    for (DevBuf *b : all) b->release();
    for (auto &pr : ctx->sites_pool) cudaFree(pr.first);
    ctx->sites_pool.clear();
    for (int k = 0; k < 2; k++)
        for (auto &pr : ctx->timed[k]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    for (cudaEvent_t e : ctx->spare_events) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *snpgpu_last_error(const snpgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

int snpgpu_set_stream(snpgpu_ctx *ctx, void *stream) {
    if (!ctx) return SNPGPU_E_ARG;
    cudaStream_t next = stream ? (cudaStream_t)stream : ctx->own_stream;
    if (next != ctx->stream) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }   // pooled blobs are reused in stream order
    ctx->stream = next;
    return SNPGPU_OK;
}

int snpgpu_sync(snpgpu_ctx *ctx) {
    if (!ctx) return SNPGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNPGPU_OK;
}

int snpgpu_host_alloc(snpgpu_ctx *ctx, size_t nbytes, void **out) {
    if (!ctx || !out) return SNPGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaHostAlloc(out, nbytes ? nbytes : 1, cudaHostAllocDefault));
    return SNPGPU_OK;
}

int snpgpu_host_free(snpgpu_ctx *ctx, void *p) {
    if (!ctx) return SNPGPU_E_ARG;
    if (p) CK(cudaFreeHost(p));
    return SNPGPU_OK;
}

uint64_t snpgpu_launch_count(const snpgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int snpgpu_enable_timing(snpgpu_ctx *ctx, int on) {
    if (!ctx) return SNPGPU_E_ARG;
    ctx->timing = on != 0;
    return SNPGPU_OK;
}

int snpgpu_kernel_time(snpgpu_ctx *ctx, int kernel, double *ms_out, uint64_t *launches_out) {
    if (!ctx || kernel < 0 || kernel > 1 || !ms_out || !launches_out) return fail(ctx, SNPGPU_E_ARG, "kernel_time: bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    double total = 0;
    for (auto &pr : ctx->timed[kernel]) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
        total += ms;
        ctx->spare_events.push_back(pr.first);
        ctx->spare_events.push_back(pr.second);
    }
    *ms_out = total;
    *launches_out = ctx->timed[kernel].size();
    ctx->timed[kernel].clear();
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ site table
int snpgpu_sites_create(snpgpu_ctx *ctx, const char *contig_names, const int32_t *name_off, int32_t n_contigs,
                        const int32_t *snp_contig, const int64_t *snp_pos, size_t n_snp, const int32_t *exc_contig,
                        const int64_t *exc_pos, size_t n_exc, snpgpu_sites **out) {
    if (!ctx || !out || n_contigs < 0) return fail(ctx, SNPGPU_E_ARG, "sites_create: bad argument");
    if ((n_contigs && (!contig_names || !name_off)) || (n_snp && (!snp_contig || !snp_pos)) ||
        (n_exc && (!exc_contig || !exc_pos)))
        return fail(ctx, SNPGPU_E_ARG, "sites_create: null array");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    HostSites h;
    const char *why = "";
    int hrc = build_host_sites(contig_names, name_off, n_contigs, snp_contig, snp_pos, n_snp, exc_contig, exc_pos, n_exc,
                               &h, &why);
    if (hrc) {
        static thread_local std::string msg;
        msg = std::string("sites_create: ") + why;
        return fail(ctx, hrc == 1 ? SNPGPU_E_ARG : hrc, msg.c_str());
    }
    const size_t n_unique = h.n_unique;
    // one blob, 256-byte aligned sections
    struct Sec { const void *src; size_t bytes; size_t off; };
    std::vector<Sec> secs;
    size_t total = 0;
    auto add = [&](const void *src, size_t bytes) { secs.push_back({src, bytes, total}); total += (bytes + 255) & ~(size_t)255; return secs.size() - 1; };
    size_t s_names4 = add(h.names4.data(), h.names4.size() * 4);
    size_t s_off4 = add(h.off4.data(), h.off4.size() * 4);
    size_t s_len1 = add(h.len1.data(), h.len1.size() * 4);
    size_t s_names = add(h.names.data(), h.names.size());
    size_t s_noff = add(h.name_off.data(), h.name_off.size() * 4);
    size_t s_base = add(h.bit_base.data(), h.bit_base.size() * 8);
    size_t s_max = add(h.max_pos.data(), h.max_pos.size() * 8);
    size_t s_bits = add(h.bits.data(), h.bits.size() * 4);
    size_t s_rank = add(h.rank.data(), h.rank.size() * 4);
    size_t s_flags = add(h.flags.data(), h.flags.size());
    size_t s_words = add(h.words.data(), h.words.size() * sizeof(SiteWord));
    size_t s_su = add(h.snp_unique.data(), h.snp_unique.size() * 4);
    size_t s_q3 = add(h.q3rows.data(), h.q3rows.size() * 4);
    snpgpu_sites *s = new (std::nothrow) snpgpu_sites();
    if (!s) return fail(ctx, SNPGPU_E_NOMEM, "sites_create: host allocation");
    // the blob of a destroyed table when one is large enough (cudaFree stalls the whole device -- up to hundreds of
    // milliseconds next to page-locked copies -- so a loop that builds a table per batch must not free one per batch)
    cudaError_t e = cudaSuccess;
    for (size_t k = 0; k < ctx->sites_pool.size(); k++) {
        if (ctx->sites_pool[k].second >= total) {
            s->blob = ctx->sites_pool[k].first; s->blob_bytes = ctx->sites_pool[k].second;
            ctx->sites_pool.erase(ctx->sites_pool.begin() + (long)k);
            break;
        }
    }
    if (!s->blob) {
        const size_t want = (total + (total >> 3) + 4095) & ~(size_t)4095;
        e = cudaMalloc(&s->blob, want);
        if (e != cudaSuccess) { delete s; return fail(ctx, SNPGPU_E_NOMEM, "sites_create: cudaMalloc", e); }
        s->blob_bytes = want;
    }
    s->pooled = true;
    for (const Sec &x : secs) {
        if (!x.bytes || !x.src) continue;
        e = cudaMemcpyAsync((uint8_t *)s->blob + x.off, x.src, x.bytes, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFree(s->blob); delete s; return fail(ctx, SNPGPU_E_CUDA, "sites_create: copy", e); }
    }
    e = cudaStreamSynchronize(ctx->stream);                      // the host vectors die with this scope
    if (e != cudaSuccess) { cudaFree(s->blob); delete s; return fail(ctx, SNPGPU_E_CUDA, "sites_create: sync", e); }
    auto at = [&](size_t i) { return (uint8_t *)s->blob + secs[i].off; };
    s->ctx = ctx; s->n_snp = n_snp; s->n_unique = n_unique; s->n_contigs = n_contigs;
    s->table.n_contigs = n_contigs; s->table.n_unique = (int32_t)n_unique;
    s->table.names4 = (const uint32_t *)at(s_names4); s->table.off4 = (const int32_t *)at(s_off4);
    s->table.len1 = (const int32_t *)at(s_len1); s->table.names = (const uint8_t *)at(s_names);
    s->table.name_off = (const int32_t *)at(s_noff); s->table.bit_base = (const int64_t *)at(s_base);
    s->table.max_pos = (const int64_t *)at(s_max); s->table.bits = (const uint32_t *)at(s_bits);
    s->table.rank = (const uint32_t *)at(s_rank); s->table.flags = (const uint8_t *)at(s_flags);
    s->table.words = (const SiteWord *)at(s_words);
    s->table.q3rows = (const uint32_t *)at(s_q3);
    s->snp_unique = (int32_t *)at(s_su);
    *out = s;
    return SNPGPU_OK;
}

void snpgpu_sites_destroy(snpgpu_sites *sites) {
    if (!sites) return;
    if (sites->pooled && sites->ctx && sites->blob && sites->ctx->sites_pool.size() < 4 && !sites->ctx->pending[0].active &&
        !sites->ctx->pending[1].active) {
        // whatever still reads the table was enqueued on the context's stream, and so is whatever reuses the blob
        sites->ctx->sites_pool.emplace_back(sites->blob, sites->blob_bytes);
        delete sites;
        return;
    }
    if (sites->ctx) { cudaSetDevice(sites->ctx->device); cudaStreamSynchronize(sites->ctx->stream); }
    if (sites->blob) cudaFree(sites->blob);
    delete sites;
}

// The site table straight from K2's output (k3_sites.cu): keys_dev = n_keys sorted unique (chrom_rank << 32 | pos)
// keys in device memory, as snpgpu_merge_sites_dev writes them; contig_len[c] bounds the positions of contig c.
// Every key is a snplist entry, in this order (the order of snplist.txt, utils.py:1068); no exclude list.
// Nothing is copied back and nothing is synchronised: the table is ready in stream order.
int snpgpu_sites_create_from_keys_dev(snpgpu_ctx *ctx, const char *contig_names, const int32_t *name_off,
                                      int32_t n_contigs, const int64_t *contig_len, const uint64_t *keys_dev,
                                      size_t n_keys, snpgpu_sites **out) {
    if (!ctx || !out || n_contigs <= 0 || !contig_names || !name_off || !contig_len || (n_keys && !keys_dev))
        return fail(ctx, SNPGPU_E_ARG, "sites_create_from_keys_dev: bad argument");
    if (n_keys >= ((size_t)1 << 31)) return fail(ctx, SNPGPU_E_ARG, "sites_create_from_keys_dev: too many keys");
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    HostSites h;
    const char *why = "";
    int hrc = build_host_sites(contig_names, name_off, n_contigs, nullptr, nullptr, 0, nullptr, nullptr, 0, &h, &why);
    if (hrc) {
        static thread_local std::string msg;
        msg = std::string("sites_create_from_keys_dev: ") + why;
        return fail(ctx, hrc == 1 ? SNPGPU_E_ARG : hrc, msg.c_str());
    }
    int64_t total_bits = 0;
    for (int c = 0; c < n_contigs; c++) {
        if (contig_len[c] < 0 || contig_len[c] >= ((int64_t)1 << 31))
            return fail(ctx, SNPGPU_E_DOMAIN, "sites_create_from_keys_dev: contig length outside [0, 2^31)");
        h.max_pos[c] = contig_len[c];
        h.bit_base[c] = total_bits;
        total_bits += (contig_len[c] + 1 + 31) / 32 * 32;
    }
    if (total_bits > ((int64_t)1 << 33)) return fail(ctx, 18, "sites_create_from_keys_dev: site bitmap above 1 GiB");
    const size_t n_words = (size_t)(total_bits / 32) + 1;
    // small host-built sections, packed into one staging buffer -> one copy
    std::vector<uint8_t> stage;
    auto put = [&](const void *src, size_t bytes) { size_t off = stage.size(); stage.resize(off + ((bytes + 255) & ~(size_t)255));
                                                     if (bytes) memcpy(stage.data() + off, src, bytes); return off; };
    const size_t o_names4 = put(h.names4.data(), h.names4.size() * 4), o_off4 = put(h.off4.data(), h.off4.size() * 4);
    const size_t o_len1 = put(h.len1.data(), h.len1.size() * 4), o_names = put(h.names.data(), h.names.size());
    const size_t o_noff = put(h.name_off.data(), h.name_off.size() * 4), o_base = put(h.bit_base.data(), h.bit_base.size() * 8);
    const size_t o_max = put(h.max_pos.data(), h.max_pos.size() * 8);
    const size_t o_q3 = put(h.q3rows.data(), h.q3rows.size() * 4);
    const size_t small = stage.size();
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_bits = small, o_rank = o_bits + up(n_words * 4), o_flags = o_rank + up(n_words * 4);
    const size_t o_su = o_flags + up(n_keys + 1), o_words = o_su + up((n_keys + 1) * 4);
    const size_t total = o_words + up(n_words * sizeof(SiteWord));
    snpgpu_sites *s = new (std::nothrow) snpgpu_sites();
    if (!s) return fail(ctx, SNPGPU_E_NOMEM, "sites_create_from_keys_dev: host allocation");
    for (size_t k = 0; k < ctx->sites_pool.size(); k++) {
        if (ctx->sites_pool[k].second >= total) {
            s->blob = ctx->sites_pool[k].first; s->blob_bytes = ctx->sites_pool[k].second;
            ctx->sites_pool.erase(ctx->sites_pool.begin() + (long)k);
            break;
        }
    }
    if (!s->blob) {
        const size_t want = total + total / 4;                  // room for the next, slightly larger, list
        cudaError_t e = cudaMalloc(&s->blob, want);
        if (e != cudaSuccess) { delete s; return fail(ctx, SNPGPU_E_NOMEM, "sites_create_from_keys_dev: cudaMalloc", e); }
        s->blob_bytes = want;
    }
    s->pooled = true;
    auto at = [&](size_t off) { return (uint8_t *)s->blob + off; };
    cudaError_t e = cudaMemcpyAsync(s->blob, stage.data(), small, cudaMemcpyHostToDevice, ctx->stream);   // pageable: staged before return
    if (e == cudaSuccess && ctx->k3_tmp.ensure(k3_scan_bytes(n_words) + 256) != cudaSuccess) e = cudaErrorMemoryAllocation;
    int launched = -1;
    if (e == cudaSuccess)
        launched = k3_launch(ctx->stream, (const unsigned long long *)keys_dev, n_keys, n_contigs, (const int64_t *)at(o_base),
                             (const int64_t *)at(o_max), (uint32_t *)at(o_bits), (uint32_t *)at(o_rank), n_words,
                             at(o_flags), (int32_t *)at(o_su), (SiteWord *)at(o_words), ctx->k3_tmp.p, ctx->k3_tmp.cap);
    if (e != cudaSuccess || launched < 0) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(s->blob); delete s;
        return fail(ctx, SNPGPU_E_CUDA, "sites_create_from_keys_dev: launch", e != cudaSuccess ? e : cudaGetLastError());
    }
    ctx->launches += (uint64_t)launched;
    s->ctx = ctx; s->n_snp = n_keys; s->n_unique = n_keys; s->n_contigs = n_contigs;
    s->table.n_contigs = n_contigs; s->table.n_unique = (int32_t)n_keys;
    s->table.names4 = (const uint32_t *)at(o_names4); s->table.off4 = (const int32_t *)at(o_off4);
    s->table.len1 = (const int32_t *)at(o_len1); s->table.names = (const uint8_t *)at(o_names);
    s->table.name_off = (const int32_t *)at(o_noff); s->table.bit_base = (const int64_t *)at(o_base);
    s->table.max_pos = (const int64_t *)at(o_max); s->table.bits = (const uint32_t *)at(o_bits);
    s->table.rank = (const uint32_t *)at(o_rank); s->table.flags = at(o_flags);
    s->table.words = (const SiteWord *)at(o_words);
    s->table.q3rows = (const uint32_t *)at(o_q3);
    s->snp_unique = (int32_t *)at(o_su);
    *out = s;
    return SNPGPU_OK;
}

size_t snpgpu_sites_n_snp(const snpgpu_sites *sites) { return sites ? sites->n_snp : 0; }

// ------------------------------------------------------------------------------------------ reference bases (snp_reference)
int snpgpu_reference_bases(snpgpu_ctx *ctx, const uint8_t *seq, size_t seq_len, const int64_t *pos, size_t n,
                           uint8_t *out, size_t *bad_index) {
    if (!ctx || (seq_len && !seq) || (n && (!pos || !out))) return fail(ctx, SNPGPU_E_ARG, "reference_bases: null argument");
    if (bad_index) *bad_index = (size_t)-1;
    if (!n) return SNPGPU_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t o_pos = (seq_len + 255) & ~(size_t)255, o_out = o_pos + ((n * 8 + 255) & ~(size_t)255);
    const size_t o_bad = o_out + ((n + 255) & ~(size_t)255);
    CK(ctx->k3_tmp.ensure(o_bad + 256));
    uint8_t *b = (uint8_t *)ctx->k3_tmp.p;
    unsigned long long none = ~0ull, bad = ~0ull;
    if (seq_len) CK(cudaMemcpyAsync(b, seq, seq_len, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_pos, pos, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_bad, &none, 8, cudaMemcpyHostToDevice, st));
    ctx->launches += (uint64_t)k3_launch_reference_bases(st, b, seq_len, (const long long *)(b + o_pos), n, b + o_out,
                                                         (unsigned long long *)(b + o_bad));
    CK(cudaMemcpyAsync(out, b + o_out, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&bad, b + o_bad, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (bad != ~0ull) {
        if (bad_index) *bad_index = (size_t)bad;
        return fail(ctx, SNPGPU_E_INDEX, "reference_bases: position outside the sequence (the reference raises IndexError)");
    }
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ K1
// One launch sequence over up to K1_BATCH samples: memset of the scratch that starts as zeros, pileup kernel + follow-up
// kernel, the two ordering kernels (per-line results wanted), finish kernel.  rec_*: the consensus-VCF pass's line list
// (single sample only).
static int k1_run_batch(snpgpu_ctx *ctx, const snpgpu_pileup_sample *smp, size_t n, const snpgpu_sites *sites,
                        const snpgpu_params *params, int mode, unsigned long long *rec_off, unsigned long long *rec_count,
                        size_t rec_cap) {
    cudaStream_t st = ctx->stream;
    const bool has_qual = params->min_base_qual > 0;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // ---- geometry
    int tile0[K1_BATCH + 1], max_tiles = 0;
    size_t total_bytes = 0, stage_tiles = 0;
    bool want_lines = false;
    tile0[0] = 0;
    for (size_t i = 0; i < n; i++) {
        const size_t nt = (smp[i].nbytes + K1_TILE - 1) / K1_TILE;
        if (nt + (size_t)tile0[i] >= ((size_t)1 << 31)) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: batch too large");
        tile0[i + 1] = tile0[i] + (int)nt;
        max_tiles = std::max(max_tiles, (int)nt);
        total_bytes += smp[i].nbytes;
        if (mode == SNPGPU_MODE_ALL && smp[i].line_out_dev) { want_lines = true; stage_tiles += nt + 1; }
    }
    // ---- scratch that starts as zeros: ticket + queue counters | per sample: status, lines per tile group, site cells
    const size_t z_head = 256;
    const size_t z_groups = want_lines ? up(((size_t)max_tiles / K1_ORDER_TILES + 2) * sizeof(unsigned long long)) : 0;
    const size_t z_cells = up((sites->n_unique + 1) * sizeof(unsigned long long));
    const size_t z_per = 256 + z_groups + z_cells;
    static_assert(sizeof(PileupStatusDev) <= 256, "status block");
    const size_t z_total = z_head + n * z_per;
    CK(ctx->k1_zero.ensure(z_total));
    CK(ctx->arena.ensure(ctx->arena_want));
    // follow-up queue: a share of the lines (all of them with a minimum base quality) + the blocks of 64 slots every warp
    // of the grid may hold half filled (k1_pileup.cu claims slots K1_QBLOCK at a time)
    const size_t q_blocks = (size_t)ctx->n_sms * ctx->k1_blocks * K1_WARPS * 64 * 2;
    const size_t q_cap = std::max<size_t>(ctx->queue_want, (has_qual || rec_off ? total_bytes / 8 : total_bytes / 256) + 65536 + q_blocks);
    CK(ctx->queue.ensure(q_cap * 16));
    if (want_lines) {
        CK(ctx->tile_first.ensure(stage_tiles * sizeof(unsigned long long)));
        CK(ctx->tile_lines.ensure(stage_tiles * sizeof(uint32_t)));
        CK(ctx->lane_lines.ensure(stage_tiles * 32));
        CK(ctx->stage.ensure(stage_tiles * K1_LCAP * 32 * sizeof(uint16_t)));
        CK(ctx->over.ensure(n * ctx->over_want * sizeof(unsigned long long)));
    }
    CK(cudaMemsetAsync(ctx->k1_zero.p, 0, z_total, st));
    uint8_t *zb = (uint8_t *)ctx->k1_zero.p;
    K1Batch g;
    memset(&g, 0, sizeof(g));
    size_t tiles_before = 0;
    for (size_t i = 0; i < n; i++) {
        K1Samp &S = g.s[i];
        uint8_t *zs = zb + z_head + i * z_per;
        const bool lines = mode == SNPGPU_MODE_ALL && smp[i].line_out_dev;
        S.text = (const uint8_t *)smp[i].text_dev;
        S.nbytes = smp[i].nbytes;
        S.tile0 = tile0[i];
        S.n_tiles = tile0[i + 1] - tile0[i];
        S.st = (PileupStatusDev *)zs;
        S.group_lines = lines ? (unsigned long long *)(zs + 256) : nullptr;
        S.site_cells = (unsigned long long *)(zs + 256 + z_groups);
        if (lines) {
            S.tile_first = (unsigned long long *)ctx->tile_first.p + tiles_before;
            S.tile_lines = (uint32_t *)ctx->tile_lines.p + tiles_before;
            S.lane_lines = (uint8_t *)ctx->lane_lines.p + tiles_before * 32;
            S.stage = (uint16_t *)ctx->stage.p + tiles_before * K1_LCAP * 32;
            S.over = (unsigned long long *)ctx->over.p + i * ctx->over_want;
            S.line_out = smp[i].line_out_dev;
            S.line_out_cap = smp[i].line_out_cap;
            tiles_before += (size_t)S.n_tiles + 1;
        }
        S.rec_off = rec_off; S.rec_count = rec_count; S.rec_cap = rec_cap;
        S.row_out = smp[i].row_out_dev;
        S.n_unique = sites->n_unique;
        S.stats_out = smp[i].stats_dev;
    }
    g.n_samples = (int)n;
    g.total_tiles = tile0[n];
    g.mode = mode;
    g.has_qual = has_qual ? 1 : 0;
    g.all_rest = (has_qual || rec_off) ? 1 : 0;
    g.one = 1u;
    g.next_tile = (unsigned int *)zb;
    g.queue_count = (unsigned long long *)(zb + 8);
    g.queue = (unsigned long long *)ctx->queue.p;
    g.queue_cap = q_cap;
    g.over_cap = want_lines ? ctx->over_want : 0;
    g.arena = (uint8_t *)ctx->arena.p;
    g.arena_cap = ctx->arena.cap;
    g.snp_unique = sites->snp_unique;
    g.n_snp = sites->n_snp;
    g.sites = sites->table;
    memcpy(&g.p, params, sizeof(CallParams));
    {
        TimedLaunch t(ctx, SNPGPU_KERNEL_PILEUP);
        ctx->launches += (uint64_t)k1_launch(st, g, ctx->n_sms * ctx->k1_blocks, ctx->n_sms);
    }
    ctx->launches += (uint64_t)k1_launch_finish(st, g, max_tiles, want_lines);
    CK(cudaGetLastError());
    return SNPGPU_OK;
}

static int k1_check(snpgpu_ctx *ctx, const snpgpu_pileup_sample *smp, size_t n, const snpgpu_sites *sites,
                    const snpgpu_params *params, int mode) {
    if (!ctx || !sites || !params || (n && !smp)) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: null argument");
    if (mode != SNPGPU_MODE_SITES && mode != SNPGPU_MODE_ALL) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: bad mode");
    for (size_t i = 0; i < n; i++) {
        if (smp[i].nbytes && !smp[i].text_dev) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: null text");
        if (((uintptr_t)smp[i].text_dev & 15u) != 0) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: text must be 16-byte aligned");
        if (sites->n_snp && !smp[i].row_out_dev) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: row_out is null");
        if (smp[i].nbytes >= ((size_t)1 << 38)) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: text of 256 GiB or more");
    }
    return SNPGPU_OK;
}

int snpgpu_pileup_consensus_batch_dev(snpgpu_ctx *ctx, const snpgpu_pileup_sample *samples, size_t n_samples,
                                      const snpgpu_sites *sites, const snpgpu_params *params, int mode) {
    int rc = k1_check(ctx, samples, n_samples, sites, params, mode);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    // launches of equal size (a launch's fixed cost -- the pileup kernel's last tiles, the follow-up kernel's latency -- is
    // shared by the samples it covers: 100 samples go as 50 + 50, not 64 + 36)
    const size_t n_launch = (n_samples + K1_BATCH - 1) / K1_BATCH;
    for (size_t b = 0, i = 0; b < n_launch; b++) {
        const size_t take = (n_samples - i + (n_launch - b) - 1) / (n_launch - b);
        rc = k1_run_batch(ctx, samples + i, take, sites, params, mode, nullptr, nullptr, 0);
        if (rc) return rc;
        i += take;
    }
    return SNPGPU_OK;
}

static int k1_run(snpgpu_ctx *ctx, const void *text_dev, size_t nbytes, const snpgpu_sites *sites,
                  const snpgpu_params *params, int mode, uint8_t *row_out_dev, uint16_t *line_out_dev,
                  size_t line_out_cap, snpgpu_pileup_stats *stats_dev, unsigned long long *rec_off,
                  unsigned long long *rec_count, size_t rec_cap) {
    snpgpu_pileup_sample s;
    s.text_dev = text_dev; s.nbytes = nbytes; s.row_out_dev = row_out_dev;
    s.line_out_dev = mode == SNPGPU_MODE_ALL ? line_out_dev : nullptr; s.line_out_cap = line_out_cap; s.stats_dev = stats_dev;
    int rc = k1_check(ctx, &s, 1, sites, params, mode);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    return k1_run_batch(ctx, &s, 1, sites, params, mode, rec_off, rec_count, rec_cap);
}

int snpgpu_pileup_consensus_dev(snpgpu_ctx *ctx, const void *text_dev, size_t nbytes, const snpgpu_sites *sites,
                                const snpgpu_params *params, int mode, uint8_t *row_out_dev, uint16_t *line_out_dev,
                                size_t line_out_cap, snpgpu_pileup_stats *stats_dev) {
    return k1_run(ctx, text_dev, nbytes, sites, params, mode, row_out_dev, line_out_dev, line_out_cap, stats_dev, nullptr,
                  nullptr, 0);
}

// what a batch that came back with SNPGPU_E_NOMEM asked for (k1_finish_kernel)
static void k1_grow(snpgpu_ctx *ctx, const snpgpu_pileup_stats &hs) {
    if (hs.reserved < 0) { ctx->queue_want = (size_t)hs.error_offset + 65536; return; }
    if (hs.error_offset) ctx->arena_want = (size_t)hs.error_offset + (1 << 20);
    if (hs.reserved > 0) ctx->over_want = ((size_t)hs.reserved << 10) + (1 << 16);
}

int snpgpu_pileup_consensus(snpgpu_ctx *ctx, const void *text, size_t nbytes, const snpgpu_sites *sites,
                            const snpgpu_params *params, int mode, uint8_t *row_out, uint16_t *line_out,
                            size_t line_out_cap, snpgpu_pileup_stats *stats) {
    if (!ctx || !sites || !params || (nbytes && !text)) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: null argument");
    if (sites->n_snp && !row_out) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus: row_out is null");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool want_lines = mode == SNPGPU_MODE_ALL && line_out != nullptr && line_out_cap > 0;
    bool normalized = false, skip_copy = false;
    ctx->text_valid = false;
    ctx->vcf_text_valid = false;
    for (int attempt = 0; attempt < 4; attempt++) {
        CK(ctx->text.ensure(nbytes + 64));
        CK(ctx->row.ensure(sites->n_snp + 16));
        CK(ctx->stats.ensure(sizeof(snpgpu_pileup_stats)));
        if (want_lines) CK(ctx->lines.ensure(line_out_cap * sizeof(uint16_t)));
        if (nbytes && !skip_copy) CK(cudaMemcpyAsync(ctx->text.p, text, nbytes, cudaMemcpyHostToDevice, st));
        // the consensus-VCF pass wants the list of parsed lines: in sites mode those are the lines at sites, which the
        // follow-up kernel handles anyway -- listing them costs this run nothing, and K5 need not run K1 again
        const bool list_now = ctx->want_rec && mode == SNPGPU_MODE_SITES;
        ctx->rec_valid = false;
        if (list_now) {
            ctx->rec_cap = std::max(ctx->rec_cap, 4 * sites->n_unique + 1024);
            CK(ctx->rec_off.ensure(ctx->rec_cap * sizeof(unsigned long long)));
            CK(ctx->k5_state.ensure(256 + sizeof(PileupStatusDev)));
            CK(cudaMemsetAsync(ctx->k5_state.p, 0, 256 + sizeof(PileupStatusDev), st));
        }
        int rc = k1_run(ctx, ctx->text.p, nbytes, sites, params, mode, (uint8_t *)ctx->row.p,
                        want_lines ? (uint16_t *)ctx->lines.p : nullptr, line_out_cap, (snpgpu_pileup_stats *)ctx->stats.p,
                        list_now ? (unsigned long long *)ctx->rec_off.p : nullptr,
                        list_now ? (unsigned long long *)ctx->k5_state.p : nullptr, list_now ? ctx->rec_cap : 0);
        if (rc) return rc;
        snpgpu_pileup_stats hs;
        unsigned long long listed = 0;
        if (list_now) CK(cudaMemcpyAsync(&listed, ctx->k5_state.p, sizeof(listed), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&hs, ctx->stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
        if (sites->n_snp) CK(cudaMemcpyAsync(row_out, ctx->row.p, sites->n_snp, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (list_now && listed > ctx->rec_cap && attempt < 3) {   // the line list was too small: grow, redo
            ctx->rec_cap = (size_t)listed + 1024;
            skip_copy = true;
            continue;
        }
        if (list_now) { ctx->rec_valid = true; ctx->rec_listed = (size_t)listed; ctx->rec_mode = mode; }
        if (hs.error_code == SNPGPU_E_LONECR && !normalized) {   // classic-Mac line ends: universal newlines, redo
            ctx->launches += (uint64_t)k1_launch_normalize(st, (uint8_t *)ctx->text.p, nbytes);
            normalized = true;
            attempt--;
            skip_copy = true;
            continue;
        }
        if (hs.error_code == SNPGPU_E_NOMEM && attempt < 3) {     // the follow-up queue, the splice scratch or the overflow list
            k1_grow(ctx, hs);                                     // was too small: grow, redo
            skip_copy = true;
            continue;
        }
        if (hs.error_code == SNPGPU_E_NOMEM) return fail(ctx, SNPGPU_E_NOMEM, "pileup_consensus: device scratch kept growing");
        if (want_lines && hs.error_code == 0) {
            size_t n = (size_t)std::min<uint64_t>(hs.n_lines, line_out_cap);
            if (n) CK(cudaMemcpy(line_out, ctx->lines.p, n * sizeof(uint16_t), cudaMemcpyDeviceToHost));
        }
        if (stats) *stats = hs;
        ctx->text_nbytes = nbytes;
        ctx->text_valid = hs.error_code == 0;
        ctx->vcf_text_valid = false;
        if (hs.error_code) {
            char msg[160];
            snprintf(msg, sizeof msg, "pileup_consensus: the reference raises (code %d) on the line at byte offset %llu",
                     hs.error_code, (unsigned long long)hs.error_offset);
            return fail(ctx, hs.error_code, msg);
        }
        return SNPGPU_OK;
    }
    return fail(ctx, SNPGPU_E_NOMEM, "pileup_consensus: splice scratch");
}

// ---- the same call, pipelined: _begin enqueues copy in + kernels + copies out on one of two lanes and returns,
//      _end waits for that call and reports like snpgpu_pileup_consensus.  A caller that streams samples keeps one call
//      ahead: the next sample's text crosses PCIe while this sample's kernels run and its results go back.
int snpgpu_pileup_consensus_begin(snpgpu_ctx *ctx, const void *text, size_t nbytes, const snpgpu_sites *sites,
                                  const snpgpu_params *params, int mode, uint8_t *row_out, uint16_t *line_out,
                                  size_t line_out_cap, snpgpu_pileup_stats *stats, int *slot_out) {
    if (!ctx || !sites || !params || !slot_out || (nbytes && !text))
        return fail(ctx, SNPGPU_E_ARG, "pileup_consensus_begin: null argument");
    if (sites->n_snp && !row_out) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus_begin: row_out is null");
    int k = !ctx->pending[0].active ? 0 : (!ctx->pending[1].active ? 1 : -1);
    if (k < 0) return fail(ctx, SNPGPU_E_ARG, "pileup_consensus_begin: two calls are already in flight");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->lane[k]) {
        int rc = snpgpu_create(ctx->device, &ctx->lane[k]);
        if (rc) return fail(ctx, rc, "pileup_consensus_begin: lane context");
        CK(cudaHostAlloc((void **)&ctx->lane_stats[k], sizeof(snpgpu_pileup_stats), cudaHostAllocDefault));
    }
    if (!ctx->lane_event) CK(cudaEventCreateWithFlags(&ctx->lane_event, cudaEventDisableTiming));
    snpgpu_ctx *L = ctx->lane[k];
    cudaStream_t st = L->stream;
    CK(cudaEventRecord(ctx->lane_event, ctx->stream));         // (a site table built in stream order on the parent is ready)
    CK(cudaStreamWaitEvent(st, ctx->lane_event, 0));
    const bool want_lines = mode == SNPGPU_MODE_ALL && line_out != nullptr && line_out_cap > 0;
    L->text_valid = false;
    L->vcf_text_valid = false;
    CK(L->text.ensure(nbytes + 64));
    CK(L->row.ensure(sites->n_snp + 16));
    CK(L->stats.ensure(sizeof(snpgpu_pileup_stats)));
    if (want_lines) CK(L->lines.ensure(line_out_cap * sizeof(uint16_t)));
    if (nbytes) CK(cudaMemcpyAsync(L->text.p, text, nbytes, cudaMemcpyHostToDevice, st));
    int rc = snpgpu_pileup_consensus_dev(L, L->text.p, nbytes, sites, params, mode, (uint8_t *)L->row.p,
                                         want_lines ? (uint16_t *)L->lines.p : nullptr, line_out_cap,
                                         (snpgpu_pileup_stats *)L->stats.p);
    if (rc) { ctx->err = L->err; return rc; }
    CK(cudaMemcpyAsync(ctx->lane_stats[k], L->stats.p, sizeof(snpgpu_pileup_stats), cudaMemcpyDeviceToHost, st));
    if (sites->n_snp) CK(cudaMemcpyAsync(row_out, L->row.p, sites->n_snp, cudaMemcpyDeviceToHost, st));
    snpgpu_ctx::Pending &p = ctx->pending[k];
    p.active = true; p.text = text; p.nbytes = nbytes; p.sites = sites; p.params = *params; p.mode = mode;
    p.row_out = row_out; p.line_out = line_out; p.line_out_cap = line_out_cap; p.stats = stats;
    *slot_out = k;
    return SNPGPU_OK;
}

int snpgpu_pileup_consensus_end(snpgpu_ctx *ctx, int slot) {
    if (!ctx || slot < 0 || slot > 1 || !ctx->pending[slot].active)
        return fail(ctx, SNPGPU_E_ARG, "pileup_consensus_end: no such call in flight");
    snpgpu_ctx::Pending p = ctx->pending[slot];
    ctx->pending[slot].active = false;
    snpgpu_ctx *L = ctx->lane[slot];
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(L->stream));
    const snpgpu_pileup_stats hs = *ctx->lane_stats[slot];
    ctx->launches += L->launches; L->launches = 0;
    if (hs.error_code == SNPGPU_E_LONECR || hs.error_code == SNPGPU_E_NOMEM) {
        // lone-CR line ends or a splice scratch that has to grow: the plain call knows how to redo the sample
        int rc = snpgpu_pileup_consensus(L, p.text, p.nbytes, p.sites, &p.params, p.mode, p.row_out, p.line_out,
                                         p.line_out_cap, p.stats);
        ctx->launches += L->launches; L->launches = 0;
        if (rc) ctx->err = L->err;
        return rc;
    }
    const bool want_lines = p.mode == SNPGPU_MODE_ALL && p.line_out != nullptr && p.line_out_cap > 0;
    if (want_lines && hs.error_code == 0) {
        size_t n = (size_t)std::min<uint64_t>(hs.n_lines, p.line_out_cap);
        if (n) {
            CK(cudaMemcpyAsync(p.line_out, L->lines.p, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, L->stream));
            CK(cudaStreamSynchronize(L->stream));
        }
    }
    if (p.stats) *p.stats = hs;
    if (hs.error_code) {
        char msg[160];
        snprintf(msg, sizeof msg, "pileup_consensus: the reference raises (code %d) on the line at byte offset %llu",
                 hs.error_code, (unsigned long long)hs.error_offset);
        return fail(ctx, hs.error_code, msg);
    }
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ K5
// the records and their ALT entries of the staged text, in file order, in ctx->rec_out / ctx->alt_out (device)
static int k5_records_dev(snpgpu_ctx *ctx, const snpgpu_sites *sites, const snpgpu_params *params, int mode, size_t alt_hint,
                          size_t *n_rec, size_t *n_alt) {
    if (!ctx->text_valid) return fail(ctx, SNPGPU_E_ARG, "pileup_vcf_records: no staged text (call snpgpu_pileup_consensus first)");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t nbytes = ctx->text_nbytes;
    // state block: [0] lines listed by K1, [1] ALT entries claimed, then the tally kernel's status
    CK(ctx->k5_state.ensure(256 + sizeof(PileupStatusDev)));
    unsigned long long *counts = (unsigned long long *)ctx->k5_state.p;
    PileupStatusDev *k5st = (PileupStatusDev *)((uint8_t *)ctx->k5_state.p + 256);
    CK(ctx->row.ensure(sites->n_snp + 16));
    // ---- which lines did K1 parse?  The list of the call itself when it was asked for (snpgpu_pileup_want_vcf_records,
    //      sites mode); else K1 runs again with the list switched on (the list may have to grow once)
    unsigned long long listed = 0;
    if (ctx->rec_valid && ctx->rec_mode == mode) {
        listed = ctx->rec_listed;
    } else {
        size_t guess = mode == SNPGPU_MODE_ALL ? nbytes / 8 + 16 : 4 * sites->n_unique + 1024;
        for (int attempt = 0; attempt < 2; attempt++) {
            CK(ctx->rec_off.ensure(guess * sizeof(unsigned long long)));
            CK(cudaMemsetAsync(ctx->k5_state.p, 0, 256 + sizeof(PileupStatusDev), st));
            int rc = k1_run(ctx, ctx->text.p, nbytes, sites, params, mode, (uint8_t *)ctx->row.p, nullptr, 0, nullptr,
                            (unsigned long long *)ctx->rec_off.p, counts, guess);
            if (rc) return rc;
            CK(cudaMemcpyAsync(&listed, counts, sizeof(listed), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (listed <= guess) break;
            guess = (size_t)listed;
            if (attempt == 1) return fail(ctx, SNPGPU_E_CUDA, "pileup_vcf_records: line list kept growing");
        }
    }
    const size_t n = (size_t)listed;
    *n_rec = n;
    *n_alt = 0;
    if (n == 0) return SNPGPU_OK;
    // ---- file order: K2's radix sort over the offsets (the values it carries along are not looked at)
    const size_t sort_b = k2_workspace_bytes(n) + 2 * ((n * sizeof(uint32_t) + 255) & ~(size_t)255);
    CK(ctx->k5_tmp.ensure(sort_b + 256));
    CK(ctx->rec_sorted.ensure(n * sizeof(unsigned long long)));
    {
        uint32_t *vin = (uint32_t *)ctx->k5_tmp.p, *vout = vin + ((n + 63) & ~(size_t)63);
        void *tmp = (uint8_t *)ctx->k5_tmp.p + 2 * ((n * sizeof(uint32_t) + 255) & ~(size_t)255);
        const unsigned long long *sorted = nullptr;
        void *spare = nullptr;
        uint32_t *hist = nullptr;
        int launches = 0;
        CK(cudaMemsetAsync(vin, 0, n * sizeof(uint32_t), st));
        if (k2_sort_pairs(st, (const uint64_t *)ctx->rec_off.p, vin, n, vout, tmp, k2_workspace_bytes(n), &sorted, &spare, &hist, &launches))
            return fail(ctx, SNPGPU_E_CUDA, "pileup_vcf_records: sort failed");
        ctx->launches += (uint64_t)launches;
        CK(cudaMemcpyAsync(ctx->rec_sorted.p, sorted, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    }
    // ---- tallies; ALT entries are claimed with a counter, so the device buffer may have to grow once as well
    CK(ctx->rec_out.ensure(n * sizeof(snpgpu_vcf_record)));
    size_t alt_room = std::max<size_t>(alt_hint, 2 * n + 64);
    for (int attempt = 0; attempt < 3; attempt++) {
        CK(ctx->alt_out.ensure(alt_room * sizeof(snpgpu_vcf_alt)));
        CK(ctx->arena.ensure(ctx->arena_want));
        CK(cudaMemsetAsync((uint8_t *)ctx->k5_state.p + 8, 0, 248 + sizeof(PileupStatusDev), st));
        ctx->launches += (uint64_t)k5_launch_tally(st, (const uint8_t *)ctx->text.p, nbytes, sites->table,
                                                   *reinterpret_cast<const CallParams *>(params),
                                                   (const unsigned long long *)ctx->rec_sorted.p, n,
                                                   (snpgpu_vcf_record *)ctx->rec_out.p, (snpgpu_vcf_alt *)ctx->alt_out.p,
                                                   alt_room, counts + 1, k5st, (uint8_t *)ctx->arena.p, ctx->arena.cap);
        CK(cudaGetLastError());
        unsigned long long claimed = 0;
        PileupStatusDev hst;
        CK(cudaMemcpyAsync(&claimed, counts + 1, sizeof(claimed), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&hst, k5st, sizeof(hst), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (hst.arena_overflow) { ctx->arena_want = (size_t)hst.arena_used + (1 << 20); continue; }
        if (hst.first_error_inv) {
            const unsigned long long e = ~hst.first_error_inv;
            char msg[160];
            snprintf(msg, sizeof msg, "pileup_vcf_records: the reference raises (code %d) on the line at byte offset %llu",
                     (int)(e & 0xff), (unsigned long long)(e >> 8));
            return fail(ctx, (int)(e & 0xff), msg);
        }
        if (claimed > alt_room) { alt_room = (size_t)claimed; continue; }
        *n_alt = (size_t)claimed;
        return SNPGPU_OK;
    }
    return fail(ctx, SNPGPU_E_NOMEM, "pileup_vcf_records: scratch kept growing");
}

int snpgpu_pileup_vcf_records(snpgpu_ctx *ctx, const snpgpu_sites *sites, const snpgpu_params *params, int mode,
                              snpgpu_vcf_record *rec_out, size_t rec_cap, size_t *n_rec, snpgpu_vcf_alt *alt_out,
                              size_t alt_cap, size_t *n_alt) {
    if (!ctx || !sites || !params || !n_rec || !n_alt) return fail(ctx, SNPGPU_E_ARG, "pileup_vcf_records: null argument");
    if ((rec_cap && !rec_out) || (alt_cap && !alt_out)) return fail(ctx, SNPGPU_E_ARG, "pileup_vcf_records: null output");
    if (int rc = k5_records_dev(ctx, sites, params, mode, alt_cap, n_rec, n_alt)) return rc;
    if (*n_rec == 0) return SNPGPU_OK;
    if (*n_rec > rec_cap || *n_alt > alt_cap) return fail(ctx, SNPGPU_E_NOMEM, "pileup_vcf_records: output capacity too small");
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(rec_out, ctx->rec_out.p, *n_rec * sizeof(snpgpu_vcf_record), cudaMemcpyDeviceToHost, st));
    if (*n_alt) CK(cudaMemcpyAsync(alt_out, ctx->alt_out.p, *n_alt * sizeof(snpgpu_vcf_alt), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SNPGPU_OK;
}

// The data lines as text (vcf_writer.py:295-379 + PyVCF3's Writer.write_record), formatted on the device.
int snpgpu_pileup_vcf_text(snpgpu_ctx *ctx, const snpgpu_sites *sites, const snpgpu_params *params, int mode,
                           const char *filter_text, int failed_snp_gt, int preserve_ref_case, char *text_out, size_t text_cap,
                           size_t *n_text, size_t *n_rec) {
    if (!ctx || !sites || !params || !filter_text || !n_text || !n_rec || (text_cap && !text_out))
        return fail(ctx, SNPGPU_E_ARG, "pileup_vcf_text: null argument");
    size_t n_alt = 0;
    *n_text = 0;
    uint64_t key = 1469598103934665603ull;                    // FNV-1a over everything the text depends on
    auto mix = [&key](const void *p, size_t nb) { for (size_t i = 0; i < nb; i++) { key ^= ((const uint8_t *)p)[i]; key *= 1099511628211ull; } };
    mix(&sites, sizeof sites); mix(params, sizeof *params); mix(&mode, sizeof mode); mix(&failed_snp_gt, sizeof failed_snp_gt);
    mix(&preserve_ref_case, sizeof preserve_ref_case); mix(filter_text, SNPGPU_VCF_FILTER_MASKS * SNPGPU_VCF_FILTER_TEXT);
    if (ctx->vcf_text_valid && ctx->vcf_text_key != key) ctx->vcf_text_valid = false;
    if (ctx->vcf_text_valid) {                                // the call before this one found text_cap too small: the text is still there
        *n_text = ctx->vcf_text_bytes;
        *n_rec = ctx->vcf_text_recs;
        if (ctx->vcf_text_bytes > text_cap) return fail(ctx, SNPGPU_E_NOMEM, "pileup_vcf_text: output capacity too small");
        CK(cudaSetDevice(ctx->device));
        CK(cudaMemcpyAsync(text_out, ctx->vcf_text.p, ctx->vcf_text_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->vcf_text_valid = false;
        return SNPGPU_OK;
    }
    if (int rc = k5_records_dev(ctx, sites, params, mode, 0, n_rec, &n_alt)) return rc;
    const size_t n = *n_rec;
    if (n == 0) return SNPGPU_OK;
    cudaStream_t st = ctx->stream;
    const size_t nb = k5_text_blocks(n);
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_len = up(SNPGPU_VCF_FILTER_MASKS * SNPGPU_VCF_FILTER_TEXT), o_bsum = o_len + up(n * sizeof(uint32_t));
    const size_t o_total = o_bsum + up(nb * sizeof(unsigned long long));
    CK(ctx->k5_tmp.ensure(o_total + 256));
    uint8_t *b = (uint8_t *)ctx->k5_tmp.p;
    CK(cudaMemcpyAsync(b, filter_text, SNPGPU_VCF_FILTER_MASKS * SNPGPU_VCF_FILTER_TEXT, cudaMemcpyHostToDevice, st));
    K5TextArgs a;
    a.text = (const uint8_t *)ctx->text.p; a.rec = (const snpgpu_vcf_record *)ctx->rec_out.p; a.alt = (const snpgpu_vcf_alt *)ctx->alt_out.p;
    a.n_rec = n; a.filter_text = (const char *)b; a.failed_snp_gt = (char)failed_snp_gt; a.preserve_ref_case = preserve_ref_case != 0;
    a.len = (uint32_t *)(b + o_len); a.block_sum = (unsigned long long *)(b + o_bsum); a.total = (unsigned long long *)(b + o_total);
    a.out = nullptr;
    ctx->launches += (uint64_t)k5_launch_text_sizes(st, a);
    unsigned long long total = 0;
    CK(cudaMemcpyAsync(&total, a.total, sizeof(total), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_text = (size_t)total;
    CK(ctx->vcf_text.ensure((size_t)total + 16));
    a.out = (char *)ctx->vcf_text.p;
    ctx->launches += (uint64_t)k5_launch_text_write(st, a);
    CK(cudaGetLastError());
    if (total > text_cap) {                                   // keep it: the caller comes back with *n_text bytes of room
        CK(cudaStreamSynchronize(st));
        ctx->vcf_text_bytes = (size_t)total; ctx->vcf_text_recs = n; ctx->vcf_text_key = key; ctx->vcf_text_valid = true;
        return fail(ctx, SNPGPU_E_NOMEM, "pileup_vcf_text: output capacity too small");
    }
    CK(cudaMemcpyAsync(text_out, ctx->vcf_text.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SNPGPU_OK;
}


int snpgpu_pileup_want_vcf_records(snpgpu_ctx *ctx, int on) {
    if (!ctx) return SNPGPU_E_ARG;
    ctx->want_rec = on != 0;
    if (!on) ctx->rec_valid = false;
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ K7
int snpgpu_filter_regions(snpgpu_ctx *ctx, const uint64_t *snp_keys, const uint32_t *seg_last, size_t n, const int32_t *max_snps,
                          const int32_t *window, int32_t n_params, const uint64_t *edge_keys, const uint32_t *edge_end,
                          size_t n_edges, uint8_t *removed_out) {
    if (!ctx || n_params < 0 || (n && (!snp_keys || !seg_last || !removed_out)) || (n_params && (!max_snps || !window)) ||
        (n_edges && (!edge_keys || !edge_end)))
        return fail(ctx, SNPGPU_E_ARG, "filter_regions: null argument");
    if (n == 0) return SNPGPU_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_last = up(n * 8), o_max = o_last + up(n * 4), o_win = o_max + up((size_t)n_params * 4 + 4);
    const size_t o_ek = o_win + up((size_t)n_params * 4 + 4), o_ee = o_ek + up(n_edges * 8 + 8), o_out = o_ee + up(n_edges * 4 + 4);
    const size_t o_tmp = o_out + up(n);
    const size_t tb = k7_workspace_bytes(n, n_params, n_edges);
    CK(ctx->k2_tmp.ensure(o_tmp + tb));
    uint8_t *b = (uint8_t *)ctx->k2_tmp.p;
    CK(cudaMemcpyAsync(b, snp_keys, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b + o_last, seg_last, n * 4, cudaMemcpyHostToDevice, st));
    if (n_params) {
        CK(cudaMemcpyAsync(b + o_max, max_snps, (size_t)n_params * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b + o_win, window, (size_t)n_params * 4, cudaMemcpyHostToDevice, st));
    }
    if (n_edges) {
        CK(cudaMemcpyAsync(b + o_ek, edge_keys, n_edges * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(b + o_ee, edge_end, n_edges * 4, cudaMemcpyHostToDevice, st));
    }
    int launches = 0;
    int rc = k7_launch(st, (const uint64_t *)b, (const uint32_t *)(b + o_last), n, (const int32_t *)(b + o_max),
                       (const int32_t *)(b + o_win), n_params, (const uint64_t *)(b + o_ek), (const uint32_t *)(b + o_ee), n_edges,
                       b + o_out, b + o_tmp, tb, &launches);
    if (rc) return fail(ctx, rc, "filter_regions: launch failed");
    ctx->launches += (uint64_t)launches;
    CK(cudaMemcpyAsync(removed_out, b + o_out, n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ K6
int snpgpu_pileup_depth_sum_dev(snpgpu_ctx *ctx, const void *text_dev, size_t nbytes, int64_t *sum_out, uint64_t *lines_out,
                                uint64_t *error_offset) {
    if (!ctx || (nbytes && !text_dev) || !sum_out) return fail(ctx, SNPGPU_E_ARG, "pileup_depth_sum: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->k5_state.ensure(256 + sizeof(PileupStatusDev)));
    unsigned long long *out3 = (unsigned long long *)ctx->k5_state.p;
    const int launched = k6_launch_depth_sum(ctx->stream, (const uint8_t *)text_dev, nbytes, out3);
    if (launched < 0) return fail(ctx, SNPGPU_E_CUDA, "pileup_depth_sum: launch failed");
    ctx->launches += (uint64_t)launched;
    unsigned long long h[3];
    CK(cudaMemcpyAsync(h, out3, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    *sum_out = (int64_t)h[0];
    if (lines_out) *lines_out = h[1];
    if (error_offset) *error_offset = ~0ull;
    if (h[2]) {
        const unsigned long long e = ~h[2];
        if (error_offset) *error_offset = e >> 8;
        return fail(ctx, (int)(e & 0xff), "pileup_depth_sum: a line outside the byte / int64 domain");
    }
    return SNPGPU_OK;
}

int snpgpu_pileup_depth_sum(snpgpu_ctx *ctx, const void *text, size_t nbytes, int64_t *sum_out, uint64_t *lines_out,
                            uint64_t *error_offset) {
    if (!ctx || (nbytes && !text) || !sum_out) return fail(ctx, SNPGPU_E_ARG, "pileup_depth_sum: null argument");
    CK(cudaSetDevice(ctx->device));
    ctx->text_valid = false;
    ctx->vcf_text_valid = false;
    CK(ctx->text.ensure(nbytes + 64));
    if (nbytes) CK(cudaMemcpyAsync(ctx->text.p, text, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    return snpgpu_pileup_depth_sum_dev(ctx, ctx->text.p, nbytes, sum_out, lines_out, error_offset);
}

int snpgpu_normalize_newlines_dev(snpgpu_ctx *ctx, void *text_dev, size_t nbytes) {
    if (!ctx || (nbytes && !text_dev)) return fail(ctx, SNPGPU_E_ARG, "normalize_newlines: null argument");
    CK(cudaSetDevice(ctx->device));
    ctx->launches += (uint64_t)k1_launch_normalize(ctx->stream, (uint8_t *)text_dev, nbytes);
    CK(cudaGetLastError());
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ K2
int snpgpu_merge_sites_dev(snpgpu_ctx *ctx, const uint64_t *keys_dev, const uint32_t *sample_of_dev, size_t n,
                           uint64_t *uniq_out_dev, uint32_t *count_out_dev, uint32_t *samples_out_dev,
                           size_t *n_uniq_out) {
    if (!ctx || !n_uniq_out || (n && (!keys_dev || !sample_of_dev || !uniq_out_dev || !count_out_dev || !samples_out_dev)))
        return fail(ctx, SNPGPU_E_ARG, "merge_sites: null argument");
    if (n >= ((size_t)1 << 31)) return fail(ctx, SNPGPU_E_ARG, "merge_sites: more than 2^31 keys");
    CK(cudaSetDevice(ctx->device));
    size_t tb = k2_workspace_bytes(n);
    CK(ctx->k2_tmp.ensure(tb));
    CK(ctx->k2_n.ensure(sizeof(unsigned long long)));
    int launches = 0;
    int rc = k2_launch(ctx->stream, keys_dev, sample_of_dev, n, uniq_out_dev, count_out_dev, samples_out_dev,
                       (unsigned long long *)ctx->k2_n.p, ctx->k2_tmp.p, ctx->k2_tmp.cap, &launches);
    if (rc) return fail(ctx, rc, "merge_sites: sort / run-length encode failed");
    ctx->launches += (uint64_t)launches;
    unsigned long long nu = 0;
    CK(cudaMemcpyAsync(&nu, ctx->k2_n.p, sizeof(nu), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n_uniq_out = (size_t)nu;
    return SNPGPU_OK;
}

int snpgpu_merge_sites(snpgpu_ctx *ctx, const uint64_t *keys, const uint32_t *sample_of, size_t n, uint64_t *uniq_out,
                       uint32_t *count_out, uint32_t *samples_out, size_t *n_uniq_out) {
    if (!ctx || !n_uniq_out || (n && (!keys || !sample_of || !uniq_out || !count_out || !samples_out)))
        return fail(ctx, SNPGPU_E_ARG, "merge_sites: null argument");
    CK(cudaSetDevice(ctx->device));
    if (n == 0) { *n_uniq_out = 0; return SNPGPU_OK; }
    CK(ctx->k2_keys.ensure(n * 8)); CK(ctx->k2_samp.ensure(n * 4)); CK(ctx->k2_uniq.ensure(n * 8));
    CK(ctx->k2_cnt.ensure(n * 4)); CK(ctx->k2_out.ensure(n * 4));
    CK(cudaMemcpyAsync(ctx->k2_keys.p, keys, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->k2_samp.p, sample_of, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    int rc = snpgpu_merge_sites_dev(ctx, (const uint64_t *)ctx->k2_keys.p, (const uint32_t *)ctx->k2_samp.p, n,
                                    (uint64_t *)ctx->k2_uniq.p, (uint32_t *)ctx->k2_cnt.p, (uint32_t *)ctx->k2_out.p,
                                    n_uniq_out);
    if (rc) return rc;
    CK(cudaMemcpyAsync(uniq_out, ctx->k2_uniq.p, *n_uniq_out * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(count_out, ctx->k2_cnt.p, *n_uniq_out * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(samples_out, ctx->k2_out.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ K4
static int k4_run(snpgpu_ctx *ctx, const uint8_t *matrix_dev, size_t n_rows, size_t n_sites, size_t row_stride,
                  size_t row_begin, size_t row_end, const uint32_t *tiles, size_t n_tiles, int32_t *dist_out_dev) {
    if (!ctx || (n_rows && n_sites && !matrix_dev) || (n_rows && !dist_out_dev))
        return fail(ctx, SNPGPU_E_ARG, "pairwise_distance: null argument");
    if (row_stride < n_sites || row_begin > row_end || row_end > n_rows)
        return fail(ctx, SNPGPU_E_ARG, "pairwise_distance: bad stride or row range");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->k4_tmp.ensure(k4_workspace_bytes(n_rows, n_sites)));
    int launches = 0;
    int rc;
    {
        TimedLaunch t(ctx, SNPGPU_KERNEL_DISTANCE);
        rc = k4_launch(ctx->stream, matrix_dev, n_rows, n_sites, row_stride, row_begin, row_end, tiles, n_tiles, dist_out_dev,
                       ctx->k4_tmp.p, ctx->n_sms, &launches);
    }
    if (rc) return fail(ctx, rc, "pairwise_distance: launch failed");
    ctx->launches += (uint64_t)launches;
    return SNPGPU_OK;
}

int snpgpu_pairwise_distance_dev(snpgpu_ctx *ctx, const uint8_t *matrix_dev, size_t n_rows, size_t n_sites,
                                 size_t row_stride, size_t row_begin, size_t row_end, int32_t *dist_out_dev) {
    return k4_run(ctx, matrix_dev, n_rows, n_sites, row_stride, row_begin, row_end, nullptr, 0, dist_out_dev);
}

int snpgpu_pairwise_distance_tiles_dev(snpgpu_ctx *ctx, const uint8_t *matrix_dev, size_t n_rows, size_t n_sites,
                                       size_t row_stride, const uint32_t *tile_rows, size_t n_tile_rows,
                                       int32_t *dist_out_dev) {
    if (n_tile_rows && !tile_rows) return fail(ctx, SNPGPU_E_ARG, "pairwise_distance_tiles: null tile list");
    return k4_run(ctx, matrix_dev, n_rows, n_sites, row_stride, 0, 0, tile_rows ? tile_rows : (const uint32_t *)"", n_tile_rows, dist_out_dev);
}

int snpgpu_pairwise_distance(snpgpu_ctx *ctx, const uint8_t *matrix, size_t n_rows, size_t n_sites, size_t row_stride,
                             int32_t *dist_out) {
    if (!ctx || (n_rows && n_sites && !matrix) || (n_rows && !dist_out))
        return fail(ctx, SNPGPU_E_ARG, "pairwise_distance: null argument");
    if (row_stride < n_sites) return fail(ctx, SNPGPU_E_ARG, "pairwise_distance: row_stride < n_sites");
    if (n_rows == 0) return SNPGPU_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t mb = n_rows * row_stride, db = n_rows * n_rows * sizeof(int32_t);
    CK(ctx->k4_mat.ensure(mb + 64));
    CK(ctx->k4_dist.ensure(db));
    if (mb) CK(cudaMemcpyAsync(ctx->k4_mat.p, matrix, mb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = snpgpu_pairwise_distance_dev(ctx, (const uint8_t *)ctx->k4_mat.p, n_rows, n_sites, row_stride, 0, n_rows,
                                          (int32_t *)ctx->k4_dist.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(dist_out, ctx->k4_dist.p, db, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SNPGPU_OK;
}

// ------------------------------------------------------------------------------------------ synthetic input
int snpgpu_synth_pileup_dev(snpgpu_ctx *ctx, const snpgpu_synth_spec *spec, const char *contig_name, void *text_dev,
                            size_t cap, size_t *nbytes_out) {
    if (!ctx || !spec || !contig_name || !text_dev || !nbytes_out) return fail(ctx, SNPGPU_E_ARG, "synth: null argument");
    CK(cudaSetDevice(ctx->device));
    size_t tb = synth_workspace_bytes(spec->genome_len);
    CK(ctx->synth_tmp.ensure(tb));
    CK(ctx->synth_n.ensure(sizeof(unsigned long long)));
    int launches = 0;
    int rc = synth_launch(ctx->stream, *spec, contig_name, (uint8_t *)text_dev, cap, (unsigned long long *)ctx->synth_n.p,
                          ctx->synth_tmp.p, ctx->synth_tmp.cap, &launches);
    if (rc) return fail(ctx, rc, "synth: launch failed");
    ctx->launches += (uint64_t)launches;
    unsigned long long n = 0;
    CK(cudaMemcpyAsync(&n, ctx->synth_n.p, sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *nbytes_out = (size_t)n;
    if (n > cap) return fail(ctx, SNPGPU_E_NOMEM, "synth: text buffer too small");
    return SNPGPU_OK;
}

int snpgpu_synth_sample_sites(snpgpu_ctx *ctx, const snpgpu_synth_spec *spec, uint32_t *pos_out, size_t cap,
                              size_t *n_out) {
    if (!spec || !n_out || (cap && !pos_out)) return fail(ctx, SNPGPU_E_ARG, "synth_sample_sites: null argument");
    synth_host_sites(*spec, pos_out, cap, n_out);
    return SNPGPU_OK;
}

}  // extern "C"
