// C ABI of the device path (include/lesseq_b200.h, lsq_* entry points): context, table upload,
// read-set upload, and the filter -> index -> count -> finalize -> solve pipeline.
//
// No CPU fallback: every compute entry point needs a CUDA device and fails with LSQ_ERR_CUDA
// otherwise.  Host code here only lays tables out and orchestrates launches.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/lesseq_b200.h"
#include "lsq_kernels.cuh"
#include "lsq_tables_host.hpp"

static_assert(lsq::ROUTE_MAX_RANKS == LSQ_MAX_RANKS, "rank limit of the router");

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define LSQ_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(LSQ_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

struct DevBuf {  // grow-only device allocation
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        else p = nullptr;
        return e;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

template <class T>
cudaError_t upload(DevBuf &b, const std::vector<T> &v, cudaStream_t s) {
    cudaError_t e = b.ensure(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

struct ReadSet {
    bool declared = false, injected = false, borrowed = false;
    bool streamed = false;  // blk[] is still being copied when lsq_count launches the filter (LSQ_READS_STREAM)
    bool copy_pending = false;  // a streamed copy into this set's buffers was enqueued and no lsq_count has waited for it
    uint32_t type = 0;
    uint64_t read_length = 0;
    uint64_t n_reads = 0, n_blocks = 0;
    uint32_t n_file_strands = 0;
    // inputs (owned copies or borrowed device pointers)
    DevBuf d_blk_off, d_blk, d_read_cs, d_blk_cs, d_line_num, d_blk_cnt, d_read_cs8, d_blk_start, d_blk_len;
    DevBuf d_grp_base, d_sup_base;  // block-offset index of a compact read set (lsq_group_sum_kernel)
    const uint32_t *blk_off = nullptr, *read_cs = nullptr, *blk_cs = nullptr, *line_num = nullptr;
    const uint8_t *blk_cnt = nullptr;
    const int2 *blk = nullptr;
    const uint8_t *read_cs8 = nullptr;  // compact wire format
    const int32_t *blk_start = nullptr;
    const uint16_t *blk_len = nullptr;
    const uint8_t *blk_len8 = nullptr;
    DevBuf d_chrom_remap, d_file_meta, d_strand_tbl;
    std::vector<std::string> chrom_names, strand_names;  // the file dictionaries (kept for lsq_shard_reads)
    // injected classes (output order CSR)
    DevBuf d_inj_off, d_inj_mask, d_inj_cnt;
    std::vector<uint64_t> h_inj_off;
};

}  // namespace

struct lsq_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int n_sms = 0;
    // tables
    bool have_tables = false;
    lsq::Tables T{};
    uint32_t E = 0, F = 0;
    lsq::HostTables H;  // host mirror (internal order): headers, permutation, form segment lengths
    std::map<std::string, uint32_t> chrom_id;
    std::vector<std::string> strand_names;
    DevBuf d_hdr, d_seg, d_fmask, d_pfx, d_tie, d_bp_chrom_off, d_stab_off, d_stab, d_cov_off, d_bp_bkt_off, d_bp_bkt, d_bp_sub, d_pinfo, d_chrom_meta,
        d_cov_start, d_cov_end, d_form_len, d_sig_cut, d_sig_piece;
    // read sets
    ReadSet sets[LSQ_MAX_METHODS];
    uint32_t M = 0;  // highest declared index + 1
    // per-run state
    DevBuf d_slot_cnt, d_slot_mask, d_bases, d_G, d_counters;
    // router (multi-GPU; this context holds the tables of the whole selection)
    uint32_t route_ranks = 0;
    std::vector<uint32_t> shard_begin, form_begin;  // [n_ranks+1] event / form boundaries in output order
    DevBuf d_owner, d_cursor, d_gather;
    DevBuf d_out[LSQ_MAX_RANKS][6], d_in[6];  // outboxes per destination, inbox: cnt, cs, line, start|blk, len, blk_cs
    DevBuf d_ext, d_scratch, d_gen_list;  // d_scratch: signature counters | hash | general-list length (cleared per read set)
    uint32_t scan_flags = 0;
    DevBuf d_n_valid, d_bases_out, d_iso, d_theta, d_rpkm, d_se, d_ll, d_llpr, d_iters;
    DevBuf d_mid;  // work order of lsq_solve_mid_kernel
    DevBuf d_flag, d_exact, d_exact_rec, d_cand;  // order-exact EM replay: flag list + per-event flags; spans, segments; records
    uint32_t exact_events = 0;  // events replayed by the last lsq_solve (statistics)
    float exact_ms = 0.f;       // host wall time of that replay (collect + sort + replay, its round trips included)
    bool counted = false, finalized = false, solved = false;
    // streamed read sets: copies run on their own stream and publish their progress to the filter
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_small = nullptr, ev_all[LSQ_MAX_METHODS] = {};
    DevBuf d_progress;            // [LSQ_MAX_METHODS] blocks of blk[] landed so far
    uint32_t *h_progress = nullptr;  // pinned: the values the copy stream writes after each chunk
    // timing: events per read set (clear | scan | expand | general | end), then finalize and solve; read lazily
    cudaEvent_t ev_set[LSQ_MAX_METHODS][5] = {};
    cudaEvent_t ev[4] = {};  // finalize begin/end, solve begin/end
    bool tm_count_pending = false, tm_solve_pending = false;
    bool set_counted[LSQ_MAX_METHODS] = {};
    lsq_timings tm{};
};

namespace {

constexpr uint32_t STREAM_MAX_CHUNKS = 64;           // progress values per streamed read set
constexpr size_t STREAM_MIN_CHUNK_BLOCKS = 2u << 20;  // 16 MB of blk[] per chunk at least

}  // namespace

extern "C" {

const char *lsq_last_error(void) { return g_err.c_str(); }
const char *lsq_version(void) { return "lesseq_b200 0.1 (sm_100a)"; }

uint64_t lsq_ars_length(uint32_t read_type, uint64_t R, const uint64_t *len, uint32_t n) {
    return lsq::ars_length(read_type, R, len, n);
}

int lsq_ctx_create(int device, lsq_ctx **out) {
    if (!out) return fail(LSQ_ERR_ARG, "lsq_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(LSQ_ERR_CUDA, std::string("no CUDA device available (lesseq_b200 has no CPU fallback): ") +
                                      cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(LSQ_ERR_ARG, "lsq_ctx_create: bad device index");
    LSQ_CUDA(cudaSetDevice(device));
    lsq_ctx *c = new lsq_ctx();
    c->device = device;
    cudaDeviceProp prop;
    LSQ_CUDA(cudaGetDeviceProperties(&prop, device));
    c->n_sms = prop.multiProcessorCount;
    LSQ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    for (auto &ev : c->ev) LSQ_CUDA(cudaEventCreate(&ev));
    for (auto &row : c->ev_set)
        for (auto &ev : row) LSQ_CUDA(cudaEventCreate(&ev));
    if (const char *f = getenv("LSQ_SCAN_FLAGS")) c->scan_flags = (uint32_t)strtoul(f, nullptr, 0);  // debugging (SCAN_NO_*)
    LSQ_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    LSQ_CUDA(cudaEventCreateWithFlags(&c->ev_small, cudaEventDisableTiming));
    for (auto &ev : c->ev_all) LSQ_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    LSQ_CUDA(c->d_progress.ensure(LSQ_MAX_METHODS * sizeof(uint32_t)));
    LSQ_CUDA(cudaMallocHost(&c->h_progress, LSQ_MAX_METHODS * STREAM_MAX_CHUNKS * sizeof(uint32_t)));
    *out = c;
    return LSQ_OK;
}

int lsq_ctx_destroy(lsq_ctx *c) {
    if (!c) return LSQ_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->ev_small) cudaEventDestroy(c->ev_small);
    for (auto &ev : c->ev_all)
        if (ev) cudaEventDestroy(ev);
    if (c->h_progress) cudaFreeHost(c->h_progress);
    for (auto &ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto &row : c->ev_set)
        for (auto &ev : row)
            if (ev) cudaEventDestroy(ev);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return LSQ_OK;
}

int lsq_ctx_set_stream(lsq_ctx *c, void *s) {
    if (!c) return fail(LSQ_ERR_ARG, "null context");
    LSQ_CUDA(cudaSetDevice(c->device));
    LSQ_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (s) {
        c->stream = (cudaStream_t)s;
        c->own_stream = false;
    } else {
        LSQ_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return LSQ_OK;
}

int lsq_ctx_sync(lsq_ctx *c) {
    if (!c) return fail(LSQ_ERR_ARG, "null context");
    LSQ_CUDA(cudaSetDevice(c->device));
    LSQ_CUDA(cudaStreamSynchronize(c->stream));
    return LSQ_OK;
}

int lsq_upload_events(lsq_ctx *c, const lsq_event_tables *t) {
    if (!c || !t) return fail(LSQ_ERR_ARG, "lsq_upload_events: null argument");
    LSQ_CUDA(cudaSetDevice(c->device));
    lsq::HostTables &H = c->H;
    std::string err;
    int rc = lsq::build_host_tables(t, H, err);
    if (rc) return fail(rc, err);
    cudaStream_t s = c->stream;
    LSQ_CUDA(upload(c->d_hdr, H.hdr, s));
    LSQ_CUDA(upload(c->d_seg, H.seg, s));
    LSQ_CUDA(upload(c->d_fmask, H.fmask, s));
    LSQ_CUDA(upload(c->d_pfx, H.pfx, s));
    LSQ_CUDA(upload(c->d_tie, H.tie, s));
    LSQ_CUDA(upload(c->d_bp_chrom_off, H.bp_chrom_off, s));
    LSQ_CUDA(upload(c->d_pinfo, H.pinfo, s));
    LSQ_CUDA(upload(c->d_chrom_meta, H.chrom_meta, s));
    {
        std::vector<uint32_t> padded(H.stab_off);  // slack so 16-byte aligned bulk copies stay in bounds
        padded.resize(padded.size() + 8, padded.back());
        LSQ_CUDA(upload(c->d_stab_off, padded, s));
        LSQ_CUDA(cudaStreamSynchronize(s));
    }
    LSQ_CUDA(upload(c->d_stab, H.stab, s));
    LSQ_CUDA(upload(c->d_bp_bkt_off, H.bp_bkt_off, s));
    LSQ_CUDA(upload(c->d_bp_bkt, H.bp_bkt, s));
    LSQ_CUDA(upload(c->d_bp_sub, H.bp_sub, s));
    LSQ_CUDA(upload(c->d_cov_off, H.cov_off, s));
    LSQ_CUDA(upload(c->d_cov_start, H.cov_start, s));
    LSQ_CUDA(upload(c->d_cov_end, H.cov_end, s));
    LSQ_CUDA(upload(c->d_form_len, H.form_len, s));
    LSQ_CUDA(upload(c->d_sig_cut, H.sig_cut, s));
    LSQ_CUDA(upload(c->d_sig_piece, H.sig_piece, s));
    LSQ_CUDA(cudaStreamSynchronize(s));

    lsq::Tables &T = c->T;
    T = H.host_view();  // sizes; pointers are replaced by the device copies below
    T.hdr = c->d_hdr.as<lsq::EvHdr>();
    T.seg = c->d_seg.as<int2>();
    T.fmask = c->d_fmask.as<uint64_t>();
    T.pfx = c->d_pfx.as<uint32_t>();
    T.tie_pool = c->d_tie.as<char>();
    T.bp_chrom_off = c->d_bp_chrom_off.as<uint32_t>();
    T.pinfo = c->d_pinfo.as<lsq::PieceInfo>();
    T.chrom_meta = c->d_chrom_meta.as<uint4>();
    T.stab_off = c->d_stab_off.as<uint32_t>();
    T.stab = c->d_stab.as<lsq::StabEntry>();
    T.bp_bkt_off = c->d_bp_bkt_off.as<uint32_t>();
    T.bp_bkt = c->d_bp_bkt.as<uint2>();
    T.bp_sub = c->d_bp_sub.as<uint2>();
    T.cov_off = c->d_cov_off.as<uint32_t>();
    T.cov_start = c->d_cov_start.as<int32_t>();
    T.cov_end = c->d_cov_end.as<int32_t>();
    T.sig_cut = c->d_sig_cut.as<int32_t>();
    T.sig_piece = c->d_sig_piece.as<uint32_t>();
    c->E = H.E;
    c->F = H.F;
    c->chrom_id.clear();
    for (uint32_t i = 0; i < t->n_chroms; ++i) c->chrom_id[t->chrom_names[i]] = i;
    c->strand_names.clear();
    for (uint32_t i = 0; i < t->n_strands; ++i) c->strand_names.push_back(t->strand_names[i]);
    c->have_tables = true;
    c->counted = c->finalized = c->solved = false;
    for (auto &rs : c->sets) rs.declared = false;
    c->M = 0;
    return LSQ_OK;
}

static int upload_G(lsq_ctx *c, uint32_t m, uint32_t read_type, uint64_t read_length) {
    std::vector<double> G;
    lsq::form_weights(c->H, read_type, read_length, G);
    LSQ_CUDA(c->d_G.ensure((size_t)LSQ_MAX_METHODS * std::max<uint32_t>(c->F, 1) * sizeof(double)));
    if (c->F)
        LSQ_CUDA(cudaMemcpyAsync(c->d_G.as<double>() + (size_t)m * c->F, G.data(), c->F * sizeof(double),
                                 cudaMemcpyHostToDevice, c->stream));  // pageable source: staged before the call returns
    return LSQ_OK;
}

int lsq_set_reads(lsq_ctx *c, uint32_t m, uint32_t read_type, uint64_t read_length, const lsq_packed_reads *r) {
    if (!c || !r) return fail(LSQ_ERR_ARG, "lsq_set_reads: null argument");
    if (!c->have_tables) return fail(LSQ_ERR_STATE, "lsq_set_reads before lsq_upload_events");
    if (m >= LSQ_MAX_METHODS) return fail(LSQ_ERR_RANGE, "read set index >= LSQ_MAX_METHODS");
    if (read_type != LSQ_READ_MEDIUM && read_type != LSQ_READ_SHORT)
        return fail(LSQ_ERR_FORMAT, "unknown read type");
    if (r->n_blocks + 2 * r->n_reads > 0xFFFFFFF0ull)
        return fail(LSQ_ERR_RANGE, "read set exceeds 2^32 (blocks + 2 * reads)");
    if (r->n_strand_names > 256 || r->n_chrom_names > 65536)
        return fail(LSQ_ERR_RANGE, "read dictionaries too large");
    LSQ_CUDA(cudaSetDevice(c->device));
    ReadSet &rs = c->sets[m];
    cudaStream_t s = c->stream;
    if (rs.copy_pending) {  // the set is replaced before it was counted: its streamed copy must not land on the new data
        LSQ_CUDA(cudaStreamWaitEvent(s, c->ev_all[m], 0));
        rs.copy_pending = false;
    }
    rs.declared = true;
    rs.injected = false;
    rs.type = read_type;
    rs.read_length = read_length;
    rs.n_reads = r->n_reads;
    rs.n_blocks = r->n_blocks;
    rs.n_file_strands = r->n_strand_names;
    rs.borrowed = r->on_device != 0;
    rs.streamed = !rs.borrowed && (r->flags & LSQ_READS_STREAM) != 0 && r->n_blocks > 0;
    if (!r->blk_off && !r->blk_cnt) return fail(LSQ_ERR_ARG, "lsq_set_reads: blk_off or blk_cnt is required");
    const bool compact_blk = r->blk_start != nullptr && (r->blk_len != nullptr || r->blk_len8 != nullptr);
    const bool len8 = compact_blk && r->blk_len8 != nullptr;
    if (!compact_blk && !r->blk && r->n_blocks) return fail(LSQ_ERR_ARG, "lsq_set_reads: blk or blk_start + blk_len/blk_len8 is required");
    if (!r->read_cs && !r->read_cs8 && r->n_reads) return fail(LSQ_ERR_ARG, "lsq_set_reads: read_cs or read_cs8 is required");
    if (r->read_cs8 && (r->n_chrom_names > 64 || r->n_strand_names > 2))
        return fail(LSQ_ERR_RANGE, "read_cs8 holds 64 chromosome and 2 strand ids at most");
    {
        const size_t n = r->n_reads, nb = r->n_blocks;
        const cudaMemcpyKind kind = rs.borrowed ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        rs.blk = nullptr;
        rs.blk_start = nullptr;
        rs.blk_len = nullptr;
        rs.blk_len8 = nullptr;
        rs.read_cs = nullptr;
        rs.read_cs8 = nullptr;
        rs.blk_cs = nullptr;
        rs.line_num = nullptr;
        rs.blk_off = nullptr;
        rs.blk_cnt = nullptr;
        if (rs.borrowed) {
            if (compact_blk) {
                rs.blk_start = r->blk_start;
                if (len8) rs.blk_len8 = r->blk_len8;
                else rs.blk_len = r->blk_len;
            } else {
                rs.blk = reinterpret_cast<const int2 *>(r->blk);
            }
            // the one-byte per-read arrays are read four reads at a time: own, padded copies (2 bytes per read)
            if (r->read_cs8) {
                LSQ_CUDA(rs.d_read_cs8.ensure(n + 4));
                if (n) LSQ_CUDA(cudaMemcpyAsync(rs.d_read_cs8.p, r->read_cs8, n, kind, s));
                rs.read_cs8 = rs.d_read_cs8.as<uint8_t>();
            } else {
                rs.read_cs = r->read_cs;
            }
            rs.blk_cs = r->blk_cs;
            rs.line_num = r->line_num;
            rs.blk_off = r->blk_off;
            if (r->blk_cnt) {
                LSQ_CUDA(rs.d_blk_cnt.ensure(n + 4));
                if (n) LSQ_CUDA(cudaMemcpyAsync(rs.d_blk_cnt.p, r->blk_cnt, n, kind, s));
                rs.blk_cnt = rs.d_blk_cnt.as<uint8_t>();
            }
        } else {
            // device buffers first (an allocation synchronises the device), then the copies: the per-read
            // arrays, then the blocks.  A streamed set (LSQ_READS_STREAM) puts all copies on the copy stream,
            // the blocks in chunks, each followed by a 4-byte write of the number of blocks landed so far; the
            // compute stream only waits for the per-read arrays and the filter kernel polls the progress word.
            if (compact_blk) {
                LSQ_CUDA(rs.d_blk_start.ensure(std::max<size_t>(nb, 1) * 4));
                LSQ_CUDA(rs.d_blk_len.ensure(std::max<size_t>(nb, 1) * (len8 ? 1 : 2)));
                rs.blk_start = rs.d_blk_start.as<int32_t>();
                if (len8) rs.blk_len8 = rs.d_blk_len.as<uint8_t>();
                else rs.blk_len = rs.d_blk_len.as<uint16_t>();
            } else {
                LSQ_CUDA(rs.d_blk.ensure(std::max<size_t>(nb, 1) * 8));
                rs.blk = rs.d_blk.as<int2>();
            }
            bool any_mixed = false;
            if (r->blk_cs) {
                if (r->read_cs8) for (size_t i = 0; i < n && !any_mixed; ++i) any_mixed = (r->read_cs8[i] & 0x80u) != 0;
                else for (size_t i = 0; i < n && !any_mixed; ++i) any_mixed = (r->read_cs[i] & LSQ_READ_MIXED) != 0;
            }
            if (r->read_cs8) LSQ_CUDA(rs.d_read_cs8.ensure(n + 4));
            else LSQ_CUDA(rs.d_read_cs.ensure(std::max<size_t>(n, 1) * 4));
            if (r->line_num) LSQ_CUDA(rs.d_line_num.ensure(std::max<size_t>(n, 1) * 4));
            if (any_mixed) LSQ_CUDA(rs.d_blk_cs.ensure(std::max<size_t>(nb, 1) * 4));
            const bool use_cnt = r->blk_cnt != nullptr;  // a quarter of the bytes of blk_off
            if (use_cnt) LSQ_CUDA(rs.d_blk_cnt.ensure(n + 4));
            else LSQ_CUDA(rs.d_blk_off.ensure((n + 1) * 4));
            if (rs.streamed) {
                s = c->copy_stream;
                LSQ_CUDA(cudaMemsetAsync(c->d_progress.as<uint32_t>() + m, 0, sizeof(uint32_t), s));
            }
            if (r->read_cs8) {
                if (n) LSQ_CUDA(cudaMemcpyAsync(rs.d_read_cs8.p, r->read_cs8, n, kind, s));
                rs.read_cs8 = rs.d_read_cs8.as<uint8_t>();
            } else {
                if (n) LSQ_CUDA(cudaMemcpyAsync(rs.d_read_cs.p, r->read_cs, n * 4, kind, s));
                rs.read_cs = rs.d_read_cs.as<uint32_t>();
            }
            if (r->line_num) {
                if (n) LSQ_CUDA(cudaMemcpyAsync(rs.d_line_num.p, r->line_num, n * 4, kind, s));
                rs.line_num = rs.d_line_num.as<uint32_t>();
            }
            if (any_mixed) {
                LSQ_CUDA(cudaMemcpyAsync(rs.d_blk_cs.p, r->blk_cs, nb * 4, kind, s));
                rs.blk_cs = rs.d_blk_cs.as<uint32_t>();
            }
            if (!use_cnt) {
                LSQ_CUDA(cudaMemcpyAsync(rs.d_blk_off.p, r->blk_off, (n + 1) * 4, kind, s));
                rs.blk_off = rs.d_blk_off.as<uint32_t>();
            } else {
                if (n) LSQ_CUDA(cudaMemcpyAsync(rs.d_blk_cnt.p, r->blk_cnt, n, kind, s));
                rs.blk_cnt = rs.d_blk_cnt.as<uint8_t>();
            }
            if (rs.streamed) LSQ_CUDA(cudaEventRecord(c->ev_small, s));
            size_t chunk = std::max<size_t>(nb, 1);
            if (rs.streamed) {
                chunk = std::max<size_t>(STREAM_MIN_CHUNK_BLOCKS, (nb + STREAM_MAX_CHUNKS - 1) / STREAM_MAX_CHUNKS);
                chunk = (chunk + 127) & ~(size_t)127;  // whole 128-byte lines of every array: no line straddles two chunks
            }
            uint32_t *hp = c->h_progress + (size_t)m * STREAM_MAX_CHUNKS;
            uint32_t k = 0;
            for (size_t b0 = 0; b0 < nb; b0 += chunk, ++k) {
                const size_t b1 = std::min(nb, b0 + chunk);
                if (compact_blk) {
                    LSQ_CUDA(cudaMemcpyAsync(rs.d_blk_start.as<int32_t>() + b0, r->blk_start + b0, (b1 - b0) * 4, kind, s));
                    if (len8) LSQ_CUDA(cudaMemcpyAsync(rs.d_blk_len.as<uint8_t>() + b0, r->blk_len8 + b0, b1 - b0, kind, s));
                    else LSQ_CUDA(cudaMemcpyAsync(rs.d_blk_len.as<uint16_t>() + b0, r->blk_len + b0, (b1 - b0) * 2, kind, s));
                } else {
                    LSQ_CUDA(cudaMemcpyAsync(rs.d_blk.as<int2>() + b0, r->blk + 2 * b0, (b1 - b0) * 8, kind, s));
                }
                if (rs.streamed) {
                    hp[k] = (uint32_t)b1;
                    LSQ_CUDA(cudaMemcpyAsync(c->d_progress.as<uint32_t>() + m, hp + k, sizeof(uint32_t), kind, s));
                }
            }
            if (rs.streamed) {
                LSQ_CUDA(cudaEventRecord(c->ev_all[m], s));
                rs.copy_pending = true;
                s = c->stream;  // everything below belongs to the compute stream, after the per-read arrays
                LSQ_CUDA(cudaStreamWaitEvent(s, c->ev_small, 0));
            }
        }
        if (rs.blk_cnt) {  // block-offset index: first block of every group of SCAN_GROUP reads
            const uint32_t n_groups = (uint32_t)((n + lsq::SCAN_GROUP - 1) / lsq::SCAN_GROUP);
            const uint32_t n_super = (n_groups + lsq::SCAN_SUPER - 1) / lsq::SCAN_SUPER;
            LSQ_CUDA(rs.d_grp_base.ensure(std::max<size_t>(n_groups, 1) * 4));
            LSQ_CUDA(rs.d_sup_base.ensure(std::max<size_t>(n_super, 1) * 4));
            if (n_groups) {
                lsq::lsq_group_sum_kernel<<<n_super, lsq::SCAN_SUPER, 0, s>>>(rs.blk_cnt, n, n_groups, rs.d_grp_base.as<uint32_t>(),
                                                                           rs.d_sup_base.as<uint32_t>());
                LSQ_CUDA(cudaGetLastError());
                lsq::lsq_super_scan_kernel<<<1, 1024, 0, s>>>(rs.d_sup_base.as<uint32_t>(), n_super);
                LSQ_CUDA(cudaGetLastError());
            }
        } else if (!rs.blk_off) {
            return fail(LSQ_ERR_ARG, "lsq_set_reads: blk_off or blk_cnt is required");
        }
    }
    // dictionaries: file chromosome ids -> annotation ids; strand order table
    {
        std::vector<std::string> cn(r->n_chrom_names), sn(r->n_strand_names);  // copied first: r may alias rs' own vectors
        for (uint32_t i = 0; i < r->n_chrom_names; ++i) cn[i] = r->chrom_names[i];
        for (uint32_t i = 0; i < r->n_strand_names; ++i) sn[i] = r->strand_names[i];
        rs.chrom_names.swap(cn);
        rs.strand_names.swap(sn);
    }
    std::vector<uint32_t> remap(std::max<uint32_t>(r->n_chrom_names, 1), 0xFFFFFFFFu);
    for (uint32_t i = 0; i < r->n_chrom_names; ++i) {
        auto it = c->chrom_id.find(r->chrom_names[i]);
        if (it != c->chrom_id.end()) remap[i] = it->second;
    }
    std::vector<uint8_t> tbl;
    lsq::strand_table(r->strand_names, r->n_strand_names, c->strand_names, tbl);
    std::vector<uint4> fmeta(remap.size(), make_uint4(0u, 0u, 0u, 0u));
    for (size_t i = 0; i < remap.size(); ++i)
        if (remap[i] != 0xFFFFFFFFu) fmeta[i] = c->H.chrom_meta[remap[i]];
    LSQ_CUDA(upload(rs.d_chrom_remap, remap, s));
    LSQ_CUDA(upload(rs.d_file_meta, fmeta, s));
    LSQ_CUDA(upload(rs.d_strand_tbl, tbl, s));
    // host buffers of a plain (not streamed) set are free again when this returns
    if (!rs.borrowed && !rs.streamed) LSQ_CUDA(cudaStreamSynchronize(s));
    c->M = std::max(c->M, m + 1);
    c->counted = c->finalized = c->solved = false;
    return upload_G(c, m, read_type, read_length);
}

int lsq_set_classes(lsq_ctx *c, uint32_t m, uint32_t read_type, uint64_t read_length, const uint64_t *cls_off,
                    const uint64_t *cls_mask, const uint64_t *cls_cnt, const uint64_t *bases) {
    if (!c || !cls_off || !bases) return fail(LSQ_ERR_ARG, "lsq_set_classes: null argument");
    if (!c->have_tables) return fail(LSQ_ERR_STATE, "lsq_set_classes before lsq_upload_events");
    if (m >= LSQ_MAX_METHODS) return fail(LSQ_ERR_RANGE, "read set index >= LSQ_MAX_METHODS");
    if (read_type != LSQ_READ_MEDIUM && read_type != LSQ_READ_SHORT)
        return fail(LSQ_ERR_FORMAT, "unknown read type");
    LSQ_CUDA(cudaSetDevice(c->device));
    ReadSet &rs = c->sets[m];
    cudaStream_t s = c->stream;
    const uint32_t E = c->E;
    const uint64_t nc = cls_off[E];
    if (nc && (!cls_mask || !cls_cnt)) return fail(LSQ_ERR_ARG, "lsq_set_classes: null class arrays");
    rs.declared = true;
    rs.injected = true;
    rs.type = read_type;
    rs.read_length = read_length;
    rs.h_inj_off.assign(cls_off, cls_off + E + 1);
    LSQ_CUDA(rs.d_inj_off.ensure((E + 1) * 8));
    LSQ_CUDA(rs.d_inj_mask.ensure(std::max<uint64_t>(nc, 1) * 8));
    LSQ_CUDA(rs.d_inj_cnt.ensure(std::max<uint64_t>(nc, 1) * 8));
    LSQ_CUDA(cudaMemcpyAsync(rs.d_inj_off.p, cls_off, (E + 1) * 8, cudaMemcpyHostToDevice, s));
    if (nc) {
        LSQ_CUDA(cudaMemcpyAsync(rs.d_inj_mask.p, cls_mask, nc * 8, cudaMemcpyHostToDevice, s));
        LSQ_CUDA(cudaMemcpyAsync(rs.d_inj_cnt.p, cls_cnt, nc * 8, cudaMemcpyHostToDevice, s));
    }
    // bases arrive in output order; the device keeps them in internal order
    std::vector<unsigned long long> b(std::max<uint32_t>(E, 1));
    for (uint32_t ie = 0; ie < E; ++ie) b[ie] = bases[c->H.order[ie]];
    LSQ_CUDA(c->d_bases.ensure((size_t)LSQ_MAX_METHODS * std::max<uint32_t>(E, 1) * 8));
    LSQ_CUDA(cudaMemcpyAsync(c->d_bases.as<unsigned long long>() + (size_t)m * E, b.data(), (size_t)E * 8,
                             cudaMemcpyHostToDevice, s));
    LSQ_CUDA(cudaStreamSynchronize(s));
    c->M = std::max(c->M, m + 1);
    c->finalized = c->solved = false;
    return upload_G(c, m, read_type, read_length);
}

static int ensure_run_buffers(lsq_ctx *c) {
    const size_t E = std::max<uint32_t>(c->E, 1), F = std::max<uint32_t>(c->F, 1);
    // class slot tables only for read sets that are counted on the device (injected class tables
    // bring their own storage); indexed by read set, so size them up to the last counted one
    uint32_t m_slots = 0;
    for (uint32_t m = 0; m < c->M; ++m)
        if (c->sets[m].declared && !c->sets[m].injected) m_slots = m + 1;
    const size_t slots = std::max<uint64_t>(c->T.total_slots, 1);
    LSQ_CUDA(c->d_slot_cnt.ensure(std::max<size_t>((size_t)m_slots * slots, 1) * 8));
    LSQ_CUDA(c->d_slot_mask.ensure(std::max<size_t>((size_t)m_slots * slots, 1) * 8));
    LSQ_CUDA(c->d_bases.ensure((size_t)LSQ_MAX_METHODS * E * 8));
    LSQ_CUDA(c->d_counters.ensure((size_t)(LSQ_MAX_METHODS * lsq::CTR_N + 32) * 8));
    LSQ_CUDA(c->d_n_valid.ensure((size_t)LSQ_MAX_METHODS * E * 8));
    LSQ_CUDA(c->d_bases_out.ensure((size_t)LSQ_MAX_METHODS * E * 8));
    LSQ_CUDA(c->d_iso.ensure(F * 8));
    LSQ_CUDA(c->d_theta.ensure(F * 8));
    LSQ_CUDA(c->d_rpkm.ensure(F * 8));
    LSQ_CUDA(c->d_se.ensure(F * 8));
    LSQ_CUDA(c->d_ll.ensure(E * 8));
    LSQ_CUDA(c->d_llpr.ensure(E * 8));
    LSQ_CUDA(c->d_iters.ensure(E * 4));
    LSQ_CUDA(c->d_flag.ensure((2 * E + 8) * 4));
    return LSQ_OK;
}

static void fill_solve_params(lsq_ctx *c, lsq::SolveParams &P) {
    memset(&P, 0, sizeof P);
    P.T = c->T;
    P.M = c->M;
    P.slot_cnt = c->d_slot_cnt.as<unsigned long long>();
    P.slot_mask = c->d_slot_mask.as<unsigned long long>();
    P.bases = c->d_bases.as<unsigned long long>();
    P.G = c->d_G.as<double>();
    P.form_len = c->d_form_len.as<uint64_t>();
    for (uint32_t m = 0; m < c->M; ++m) {
        if (c->sets[m].injected) {
            P.inj_off[m] = c->sets[m].d_inj_off.as<uint64_t>();
            P.inj_mask[m] = c->sets[m].d_inj_mask.as<uint64_t>();
            P.inj_cnt[m] = c->sets[m].d_inj_cnt.as<uint64_t>();
        }
    }
    P.n_valid = c->d_n_valid.as<unsigned long long>();
    P.bases_out = c->d_bases_out.as<unsigned long long>();
    P.iso_count = c->d_iso.as<unsigned long long>();
    P.theta = c->d_theta.as<double>();
    P.rpkm = c->d_rpkm.as<double>();
    P.se = c->d_se.as<double>();
    P.loglik = c->d_ll.as<double>();
    P.ll_per_read = c->d_llpr.as<double>();
    P.iters = c->d_iters.as<uint32_t>();
    P.max_iters = 10000000u;
    P.flag_n = c->d_flag.as<uint32_t>();
    P.flag_list = P.flag_n + 4;
    P.ev_flag = P.flag_list + std::max<uint32_t>(c->E, 1);
    P.pairs_valid = c->d_counters.as<unsigned long long>() + LSQ_MAX_METHODS * lsq::CTR_N + 2;
}

static int run_finalize(lsq_ctx *c) {
    if (c->finalized) return LSQ_OK;
    lsq::SolveParams P;
    fill_solve_params(c, P);
    if (c->E) {
        cudaMemsetAsync(P.pairs_valid, 0, 8, c->stream);
        lsq::lsq_finalize_kernel<<<(c->E + 127) / 128, 128, 0, c->stream>>>(P);
        LSQ_CUDA(cudaGetLastError());
        c->tm.launches++;
    }
    c->finalized = true;
    return LSQ_OK;
}

static inline size_t pad16(size_t bytes) { return (bytes + 15) & ~(size_t)15; }

int lsq_count(lsq_ctx *c) {
    if (!c) return fail(LSQ_ERR_ARG, "null context");
    if (!c->have_tables) return fail(LSQ_ERR_STATE, "lsq_count before lsq_upload_events");
    if (c->M == 0) return fail(LSQ_ERR_STATE, "lsq_count without read sets");
    for (uint32_t m = 0; m < c->M; ++m)
        if (!c->sets[m].declared) return fail(LSQ_ERR_STATE, "read set indices must be contiguous from 0");
    LSQ_CUDA(cudaSetDevice(c->device));
    int rc = ensure_run_buffers(c);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    const lsq::Tables &T = c->T;
    lsq_timings &tm = c->tm;
    memset(&tm, 0, sizeof tm);
    const uint32_t E = c->E;
    for (uint32_t m = 0; m < c->M; ++m) {
        ReadSet &rs = c->sets[m];
        c->set_counted[m] = false;
        if (rs.injected) continue;
        const size_t n = rs.n_reads, nb = rs.n_blocks;
        LSQ_CUDA(c->d_ext.ensure((nb + 2 * n + 2) * 8));
        LSQ_CUDA(c->d_gen_list.ensure(std::max<size_t>(n, 1) * sizeof(uint2)));
        // signature scratch: counters of the single-block signatures, hash of the two-block ones
        const size_t n_sig = std::max<uint32_t>(T.n_sig, 1);
        size_t want = std::min<size_t>(4 * (size_t)T.n_pieces, 2 * n);
        size_t hcap = 1024;
        while (hcap < want && hcap < ((size_t)1 << 28)) hcap <<= 1;
        const size_t o_sig_dev = pad16(n_sig * 4), o_hash = o_sig_dev + pad16(n_sig * 8), o_hash_dev = o_hash + hcap * 16,
                     o_gen_n = o_hash_dev + hcap * 8, scratch_bytes = o_gen_n + 16;
        LSQ_CUDA(c->d_scratch.ensure(scratch_bytes));
        unsigned char *sc = c->d_scratch.as<unsigned char>();

        lsq::ScanParams P;
        memset(&P, 0, sizeof P);
        P.T = T;
        P.n_reads = n;
        P.blk_off = rs.blk_cnt ? nullptr : rs.blk_off;
        P.blk_cnt = rs.blk_cnt;
        P.grp_base = rs.d_grp_base.as<uint32_t>();
        P.sup_base = rs.d_sup_base.as<uint32_t>();
        P.blk = rs.blk;
        P.read_cs = rs.read_cs;
        P.read_cs8 = rs.read_cs8;
        P.blk_start = rs.blk_start;
        P.blk_len = rs.blk_len;
        P.blk_len8 = rs.blk_len8;
        P.blk_cs = rs.blk_cs;
        P.chrom_remap = rs.d_chrom_remap.as<uint32_t>();
        P.file_meta = rs.d_file_meta.as<uint4>();
        P.progress = rs.streamed ? c->d_progress.as<uint32_t>() + m : nullptr;
        P.line_num = rs.line_num;
        P.strand_tbl = rs.d_strand_tbl.as<uint8_t>();
        P.ext = c->d_ext.as<int2>();
        P.sig_cnt = reinterpret_cast<uint32_t *>(sc);
        P.sig_dev = reinterpret_cast<unsigned long long *>(sc + o_sig_dev);
        P.hash = reinterpret_cast<ulonglong2 *>(sc + o_hash);
        P.hash_dev = reinterpret_cast<unsigned long long *>(sc + o_hash_dev);
        P.hash_mask = (uint32_t)(hcap - 1);
        P.len0 = (uint32_t)std::min<uint64_t>(rs.read_length, 0x7FFFFFFFull);
        P.gen_list = c->d_gen_list.as<uint2>();
        P.gen_n = reinterpret_cast<uint32_t *>(sc + o_gen_n);
        P.gen_cap = (uint32_t)std::min<size_t>(n, 0xFFFFFFFFull);
        P.flags = c->scan_flags;
        P.counters = c->d_counters.as<unsigned long long>() + (size_t)m * lsq::CTR_N;
        P.slot_cnt = c->d_slot_cnt.as<unsigned long long>() + (size_t)m * T.total_slots;
        P.slot_mask = c->d_slot_mask.as<unsigned long long>() + (size_t)m * T.total_slots;
        P.bases = c->d_bases.as<unsigned long long>() + (size_t)m * E;

        LSQ_CUDA(cudaEventRecord(c->ev_set[m][0], s));
        {   // one launch clears the set's class tables, its counters and the signature scratch
            lsq::ClearList L;
            memset(&L, 0, sizeof L);
            auto add = [&](void *p, size_t bytes) {
                L.p[L.n] = reinterpret_cast<unsigned long long *>(p);
                L.n8[L.n] = bytes / 8;
                ++L.n;
            };
            add(P.slot_cnt, (size_t)T.total_slots * 8);
            add(P.slot_mask, (size_t)T.total_slots * 8);
            add(P.bases, (size_t)E * 8);
            add(P.counters, lsq::CTR_N * 8);
            add(sc, scratch_bytes);
            lsq::lsq_clear_kernel<<<c->n_sms * 4, 256, 0, s>>>(L);
            LSQ_CUDA(cudaGetLastError());
            tm.launches++;
        }
        LSQ_CUDA(cudaEventRecord(c->ev_set[m][1], s));
        if (n) {
            const uint64_t n_groups = (n + lsq::SCAN_GROUP - 1) / lsq::SCAN_GROUP;
            const unsigned grid = (unsigned)std::min<uint64_t>((n_groups + lsq::SCAN_WARPS - 1) / lsq::SCAN_WARPS, (uint64_t)c->n_sms * lsq::SCAN_MINB);
            const bool compact = P.blk_cnt && P.read_cs8 && P.blk_start && P.blk_len8;
            if (compact) lsq::lsq_scan_kernel<true><<<grid, lsq::SCAN_BLOCK, 0, s>>>(P);
            else lsq::lsq_scan_kernel<false><<<grid, lsq::SCAN_BLOCK, 0, s>>>(P);
            LSQ_CUDA(cudaGetLastError());
            tm.launches++;
        }
        if (rs.streamed) {  // the caller's buffers are free again once the copy has finished
            LSQ_CUDA(cudaStreamWaitEvent(s, c->ev_all[m], 0));
        }
        LSQ_CUDA(cudaEventRecord(c->ev_set[m][2], s));
        if (n) {
            const uint64_t n_exp = (uint64_t)T.n_sig + hcap;
            lsq::lsq_expand_kernel<<<(unsigned)((n_exp + 127) / 128), 128, 0, s>>>(P);
            LSQ_CUDA(cudaGetLastError());
            tm.launches++;
        }
        LSQ_CUDA(cudaEventRecord(c->ev_set[m][3], s));
        if (n) {
            lsq::lsq_general_kernel<<<c->n_sms * 8, 128, 0, s>>>(P);
            LSQ_CUDA(cudaGetLastError());
            tm.launches++;
        }
        LSQ_CUDA(cudaEventRecord(c->ev_set[m][4], s));
        c->set_counted[m] = true;
        tm.reads_in += n;
        tm.blocks_in += nb;
    }
    c->counted = true;
    c->finalized = false;
    c->solved = false;
    LSQ_CUDA(cudaEventRecord(c->ev[0], s));
    rc = run_finalize(c);
    if (rc) return rc;
    LSQ_CUDA(cudaEventRecord(c->ev[1], s));
    c->tm_count_pending = true;
    c->tm_solve_pending = false;
    // streamed sets: the contract is that the host buffers are free when lsq_count returns
    for (uint32_t m = 0; m < c->M; ++m) {
        ReadSet &rs = c->sets[m];
        if (!rs.injected && rs.streamed && rs.copy_pending) {
            LSQ_CUDA(cudaEventSynchronize(c->ev_all[m]));
            rs.copy_pending = false;
        }
    }
    return LSQ_OK;
}

int lsq_get_count(lsq_ctx *c, lsq_count_out *o) {
    if (!c || !o) return fail(LSQ_ERR_ARG, "lsq_get_count: null argument");
    bool any_inj = false, all_ready = true;
    for (uint32_t m = 0; m < c->M; ++m) {
        any_inj |= c->sets[m].injected;
        all_ready &= c->sets[m].injected || c->counted;
    }
    (void)any_inj;
    if (!all_ready || c->M == 0) return fail(LSQ_ERR_STATE, "lsq_get_count before lsq_count");
    LSQ_CUDA(cudaSetDevice(c->device));
    int rc = ensure_run_buffers(c);
    if (rc) return rc;
    rc = run_finalize(c);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    const size_t E = c->E, F = c->F, M = c->M;
    if (o->n_valid && E) LSQ_CUDA(cudaMemcpyAsync(o->n_valid, c->d_n_valid.p, M * E * 8, cudaMemcpyDeviceToHost, s));
    if (o->bases && E) LSQ_CUDA(cudaMemcpyAsync(o->bases, c->d_bases_out.p, M * E * 8, cudaMemcpyDeviceToHost, s));
    if (o->iso_count && F) LSQ_CUDA(cudaMemcpyAsync(o->iso_count, c->d_iso.p, F * 8, cudaMemcpyDeviceToHost, s));
    LSQ_CUDA(cudaStreamSynchronize(s));
    return LSQ_OK;
}

int lsq_get_classes(lsq_ctx *c, uint32_t m, uint64_t *cls_off, uint64_t *cls_mask, uint64_t *cls_cnt) {
    if (!c || !cls_off) return fail(LSQ_ERR_ARG, "lsq_get_classes: null argument");
    if (m >= c->M || !c->sets[m].declared) return fail(LSQ_ERR_ARG, "lsq_get_classes: unknown read set");
    LSQ_CUDA(cudaSetDevice(c->device));
    const uint32_t E = c->E;
    ReadSet &rs = c->sets[m];
    if (rs.injected) {
        std::copy(rs.h_inj_off.begin(), rs.h_inj_off.end(), cls_off);
        if (cls_mask && cls_cnt && cls_off[E]) {
            LSQ_CUDA(cudaMemcpy(cls_mask, rs.d_inj_mask.p, cls_off[E] * 8, cudaMemcpyDeviceToHost));
            LSQ_CUDA(cudaMemcpy(cls_cnt, rs.d_inj_cnt.p, cls_off[E] * 8, cudaMemcpyDeviceToHost));
        }
        return LSQ_OK;
    }
    if (!c->counted) return fail(LSQ_ERR_STATE, "lsq_get_classes before lsq_count");
    const size_t slots = c->T.total_slots;
    std::vector<unsigned long long> cnt(std::max<size_t>(slots, 1)), msk(std::max<size_t>(slots, 1));
    LSQ_CUDA(cudaStreamSynchronize(c->stream));
    if (slots) {
        LSQ_CUDA(cudaMemcpy(cnt.data(), c->d_slot_cnt.as<unsigned long long>() + (size_t)m * slots, slots * 8,
                            cudaMemcpyDeviceToHost));
        LSQ_CUDA(cudaMemcpy(msk.data(), c->d_slot_mask.as<unsigned long long>() + (size_t)m * slots, slots * 8,
                            cudaMemcpyDeviceToHost));
    }
    // output order: event e is internal event inv[e]
    std::vector<uint32_t> inv(E);
    for (uint32_t ie = 0; ie < E; ++ie) inv[c->H.order[ie]] = ie;
    uint64_t tot = 0;
    for (uint32_t e = 0; e < E; ++e) {
        cls_off[e] = tot;
        const lsq::EvHdr &h = c->H.hdr[inv[e]];
        std::vector<std::pair<uint64_t, uint64_t> > cl;
        for (uint32_t i = 0; i < h.slot_cap; ++i) {
            uint64_t n = cnt[h.slot_off + i];
            if (!n) continue;
            uint64_t mk = (h.flags & lsq::EV_HASH) ? msk[h.slot_off + i] : (uint64_t)(i + 1);
            cl.push_back(std::make_pair(mk, n));
        }
        std::sort(cl.begin(), cl.end());
        if (cls_mask && cls_cnt)
            for (size_t i = 0; i < cl.size(); ++i) {
                cls_mask[tot + i] = cl[i].first;
                cls_cnt[tot + i] = cl[i].second;
            }
        tot += cl.size();
    }
    cls_off[E] = tot;
    return LSQ_OK;
}

// Order-exact EM replay of the flagged events (lsq_kernels.cuh: lsq_collect_kernel, lsq_exact_kernel).
static int exact_replay(lsq_ctx *c, const lsq::SolveParams &P, uint32_t n_flag) {
    cudaStream_t s = c->stream;
    // LSQ_EXACT_TIMING=1: wall-clock milestones of the replay on stderr (each one synchronises the stream)
    const bool timing = getenv("LSQ_EXACT_TIMING") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(s);
        fprintf(stderr, "[exact] %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    const uint32_t E = c->E, M = c->M;
    const lsq::HostTables &H = c->H;
    std::vector<uint32_t> flagged(n_flag);
    std::vector<unsigned long long> nv((size_t)M * E);
    LSQ_CUDA(cudaMemcpyAsync(flagged.data(), P.flag_list, (size_t)n_flag * 4, cudaMemcpyDeviceToHost, s));
    LSQ_CUDA(cudaMemcpyAsync(nv.data(), c->d_n_valid.p, nv.size() * 8, cudaMemcpyDeviceToHost, s));
    LSQ_CUDA(cudaStreamSynchronize(s));
    // record segments: an event holds exactly its valid reads (of the read sets counted on the device)
    std::vector<uint64_t> seg_off(n_flag + 1, 0);
    for (uint32_t i = 0; i < n_flag; ++i) {
        uint64_t tot = 0;
        for (uint32_t m = 0; m < M; ++m)
            if (!c->sets[m].injected) tot += nv[(size_t)m * E + H.hdr[flagged[i]].out_ev];
        seg_off[i + 1] = seg_off[i] + tot;
    }
    // merged spans of the flagged events per chromosome: the collect kernel's pre-check
    const uint32_t nch = H.n_chroms;
    std::vector<std::vector<std::pair<int32_t, int32_t> > > spans(std::max<uint32_t>(nch, 1));
    for (uint32_t i = 0; i < n_flag; ++i) {
        const lsq::EvHdr &h = H.hdr[flagged[i]];
        spans[H.ev_chrom[flagged[i]]].push_back(std::make_pair(h.gs, h.ge));
    }
    std::vector<int32_t> iv_start, iv_end;
    std::vector<uint32_t> iv_off(nch + 1, 0);
    for (uint32_t ch = 0; ch < nch; ++ch) {
        std::sort(spans[ch].begin(), spans[ch].end());
        for (auto &sp : spans[ch]) {
            if (iv_start.size() > iv_off[ch] && sp.first <= iv_end.back()) iv_end.back() = std::max(iv_end.back(), sp.second);
            else {
                iv_start.push_back(sp.first);
                iv_end.push_back(sp.second);
            }
        }
        iv_off[ch + 1] = (uint32_t)iv_start.size();
    }
    // coarse test of the collect pass: one bit per bin of 2^COLLECT_BIN_SHIFT positions that a flagged span overlaps on
    // any chromosome
    uint32_t n_bins = 0;
    for (size_t i = 0; i < iv_end.size(); ++i)
        if (iv_end[i] >= 0) n_bins = std::max(n_bins, ((uint32_t)iv_end[i] >> lsq::COLLECT_BIN_SHIFT) + 1u);
    std::vector<uint32_t> bin_bits((n_bins + 31) / 32 + 1, 0);
    for (size_t i = 0; i < iv_end.size(); ++i) {
        if (iv_end[i] < 0) continue;
        const uint32_t b0 = (uint32_t)std::max(iv_start[i], 0) >> lsq::COLLECT_BIN_SHIFT, b1 = (uint32_t)iv_end[i] >> lsq::COLLECT_BIN_SHIFT;
        for (uint32_t b = b0; b <= b1; ++b) bin_bits[b >> 5] |= 1u << (b & 31u);
    }
    // device layout: seg_off | seg_cur | iv_off | iv_start | iv_end | strand ranks (one row per read set) | candidate counters | bin_bits | pre_off
    const size_t niv = iv_start.size();
    const size_t o_cur = (size_t)(n_flag + 1) * 8, o_ivoff = o_cur + (((size_t)n_flag * 4 + 7) & ~(size_t)7), o_ivs = o_ivoff + (size_t)(nch + 2) * 4,
                 o_ive = o_ivs + niv * 4 + 4, o_sr = o_ive + niv * 4 + 4, o_candn = o_sr + (size_t)LSQ_MAX_METHODS * 256,
                 o_bits = o_candn + (size_t)LSQ_MAX_METHODS * 4, o_poff = (o_bits + bin_bits.size() * 4 + 7) & ~(size_t)7,
                 bytes = o_poff + (size_t)(n_flag + 1) * 8;
    LSQ_CUDA(c->d_exact.ensure(bytes));
    const uint64_t n_rec = std::max<uint64_t>(seg_off[n_flag], 1);  // records, then one class byte per record, then the chunk tables
    std::vector<uint64_t> pre_off(n_flag + 1, 0);                   // rows of EXACT_MAX_CLS counters (ordered_sum)
    uint64_t n_max = 0;
    for (uint32_t i = 0; i < n_flag; ++i) {
        const uint64_t n_i = seg_off[i + 1] - seg_off[i];
        pre_off[i + 1] = pre_off[i] + (n_i >= lsq::ORD_MIN_N ? n_i / lsq::ORD_CHUNK + 2 : 0);
        n_max = std::max(n_max, n_i);
    }
    const uint64_t o_cls = n_rec * sizeof(lsq::ExactRec), o_pre = (o_cls + n_rec + 32ull * n_flag + 64 + 15) & ~(uint64_t)15;
    const uint64_t o_ctab = (o_pre + (pre_off[n_flag] + 1) * lsq::EXACT_MAX_CLS * 4 + 15) & ~(uint64_t)15;
    LSQ_CUDA(c->d_exact_rec.ensure(o_ctab + (uint64_t)n_flag * lsq::EXACT_CTAB * 8));
    unsigned char *d = c->d_exact.as<unsigned char>();
    std::vector<unsigned char> hbuf(bytes, 0);
    memcpy(hbuf.data(), seg_off.data(), (size_t)(n_flag + 1) * 8);
    memcpy(hbuf.data() + o_ivoff, iv_off.data(), (size_t)(nch + 1) * 4);
    if (niv) {
        memcpy(hbuf.data() + o_ivs, iv_start.data(), niv * 4);
        memcpy(hbuf.data() + o_ive, iv_end.data(), niv * 4);
    }
    memcpy(hbuf.data() + o_bits, bin_bits.data(), bin_bits.size() * 4);
    memcpy(hbuf.data() + o_poff, pre_off.data(), (size_t)(n_flag + 1) * 8);
    for (uint32_t m = 0; m < M; ++m) {  // rank of every file strand string (bytewise order, count.cpp:64-85)
        const std::vector<std::string> &sn = c->sets[m].strand_names;
        for (size_t i = 0; i < sn.size() && i < 256; ++i) {
            unsigned rank = 0;
            for (size_t j = 0; j < sn.size(); ++j) rank += sn[j] < sn[i] ? 1u : 0u;
            hbuf[o_sr + (size_t)m * 256 + i] = (unsigned char)rank;
        }
    }
    LSQ_CUDA(cudaMemcpyAsync(d, hbuf.data(), bytes, cudaMemcpyHostToDevice, s));
    if (timing) fprintf(stderr, "[exact] events %u records %llu largest %llu\n", n_flag, (unsigned long long)seg_off[n_flag], (unsigned long long)n_max);
    mark("tables on the device");
    for (uint32_t m = 0; m < M; ++m) {
        ReadSet &rs = c->sets[m];
        if (rs.injected || rs.n_reads == 0) continue;
        lsq::ScanParams SP;
        memset(&SP, 0, sizeof SP);
        SP.T = c->T;
        SP.n_reads = rs.n_reads;
        SP.blk_off = rs.blk_cnt ? nullptr : rs.blk_off;
        SP.blk_cnt = rs.blk_cnt;
        SP.grp_base = rs.d_grp_base.as<uint32_t>();
        SP.sup_base = rs.d_sup_base.as<uint32_t>();
        SP.blk = rs.blk;
        SP.read_cs = rs.read_cs;
        SP.read_cs8 = rs.read_cs8;
        SP.blk_start = rs.blk_start;
        SP.blk_len = rs.blk_len;
        SP.blk_len8 = rs.blk_len8;
        SP.blk_cs = rs.blk_cs;
        SP.chrom_remap = rs.d_chrom_remap.as<uint32_t>();
        SP.file_meta = rs.d_file_meta.as<uint4>();
        SP.line_num = rs.line_num;
        SP.strand_tbl = rs.d_strand_tbl.as<uint8_t>();
        LSQ_CUDA(c->d_ext.ensure((rs.n_blocks + 2 * rs.n_reads + 2) * 8));
        SP.ext = c->d_ext.as<int2>();
        lsq::CollectParams C;
        C.ev_flag = P.ev_flag;
        C.iv_chrom_off = reinterpret_cast<const uint32_t *>(d + o_ivoff);
        C.iv_start = reinterpret_cast<const int32_t *>(d + o_ivs);
        C.iv_end = reinterpret_cast<const int32_t *>(d + o_ive);
        C.bin_bits = reinterpret_cast<const uint32_t *>(d + o_bits);
        C.n_bins = n_bins;
        C.seg_off = reinterpret_cast<const uint64_t *>(d);
        C.seg_cur = reinterpret_cast<uint32_t *>(d + o_cur);
        C.rec = c->d_exact_rec.as<lsq::ExactRec>();
        C.strand_rank = d + o_sr + (size_t)m * 256;
        C.m = m;
        C.cand_n = reinterpret_cast<uint32_t *>(d + o_candn) + m;
        // pass 1 lists the blocks that pass the coarse test; the list is sized for 1 block in 16 and grown on demand
        uint32_t n_cand = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            const uint64_t cap = attempt == 0 ? std::max<uint64_t>(rs.n_blocks / 16, 1u << 20) : (uint64_t)n_cand;
            LSQ_CUDA(c->d_cand.ensure(cap * 4));
            C.cand = c->d_cand.as<uint32_t>();
            C.cand_cap = (uint32_t)std::min<uint64_t>(cap, 0xFFFFFFFFull);
            if (attempt) LSQ_CUDA(cudaMemsetAsync(C.cand_n, 0, 4, s));
            const unsigned grid = (unsigned)std::min<uint64_t>((rs.n_blocks / 4 + 255) / 256 + 1, (uint64_t)c->n_sms * 16);
            lsq::lsq_collect_list_kernel<<<grid, 256, 0, s>>>(SP, C, rs.n_blocks);
            LSQ_CUDA(cudaGetLastError());
            c->tm.launches++;
            LSQ_CUDA(cudaMemcpyAsync(&n_cand, C.cand_n, 4, cudaMemcpyDeviceToHost, s));
            LSQ_CUDA(cudaStreamSynchronize(s));
            if (n_cand <= C.cand_cap) break;
        }
        if (n_cand) {
            const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)n_cand + 127) / 128, (uint64_t)c->n_sms * 16);
            lsq::lsq_collect_kernel<<<grid, 128, 0, s>>>(SP, C);
            LSQ_CUDA(cudaGetLastError());
            c->tm.launches++;
        }
    }
    mark("records collected");
    lsq::ExactParams X;
    X.n_flag = n_flag;
    X.flag_list = P.flag_list;
    X.seg_off = reinterpret_cast<const uint64_t *>(d);
    X.seg_cur = reinterpret_cast<const uint32_t *>(d + o_cur);
    X.rec = c->d_exact_rec.as<lsq::ExactRec>();
    X.cls = c->d_exact_rec.as<uint8_t>() + o_cls;
    X.pre = reinterpret_cast<uint32_t *>(c->d_exact_rec.as<uint8_t>() + o_pre);
    X.pre_off = reinterpret_cast<const uint64_t *>(d + o_poff);
    X.ctab = reinterpret_cast<unsigned long long *>(c->d_exact_rec.as<uint8_t>() + o_ctab);
    X.total = seg_off[n_flag];
    X.scalar_sums = getenv("LSQ_EXACT_SCALAR_SUMS") ? 1u : 0u;
    X.debug = timing ? 1u : 0u;
    X.presorted = n_max > 8192 ? 1u : 0u;  // one block sorts a short event by itself; long ones take the whole device
    if (X.presorted) {
        const unsigned grid = (unsigned)((X.total + 255) / 256);
        for (uint64_t k = 2; (k >> 1) < n_max; k <<= 1) {
            X.sort_k = (uint32_t)k;
            X.sort_j = 0;
            lsq::lsq_exact_sort_kernel<<<grid, 256, 0, s>>>(X);
            for (uint32_t j = (uint32_t)(k >> 2); j > 0; j >>= 1) {
                X.sort_j = j;
                lsq::lsq_exact_sort_kernel<<<grid, 256, 0, s>>>(X);
            }
            c->tm.launches += 1;  // counted per merge level, not per step
        }
        LSQ_CUDA(cudaMemsetAsync(X.ctab, 0, (size_t)n_flag * lsq::EXACT_CTAB * 8, s));
        lsq::lsq_exact_classify_kernel<<<grid, 256, 0, s>>>(X);
        LSQ_CUDA(cudaGetLastError());
        c->tm.launches++;
    }
    mark("sorted and classified");
    lsq::lsq_exact_kernel<<<n_flag, lsq::EXACT_BLOCK, 0, s>>>(P, X);
    LSQ_CUDA(cudaGetLastError());
    c->tm.launches++;
    LSQ_CUDA(cudaStreamSynchronize(s));  // hbuf and friends go out of scope
    mark("replayed");
    c->exact_events = n_flag;
    return LSQ_OK;
}

int lsq_solve(lsq_ctx *c, const double *total_read_bases, uint32_t n_methods) {
    if (!c || !total_read_bases) return fail(LSQ_ERR_ARG, "lsq_solve: null argument");
    if (!c->have_tables || c->M == 0) return fail(LSQ_ERR_STATE, "lsq_solve without tables / read sets");
    if (n_methods != c->M) return fail(LSQ_ERR_ARG, "lsq_solve: n_methods differs from the declared read sets");
    for (uint32_t m = 0; m < c->M; ++m) {
        if (!c->sets[m].declared) return fail(LSQ_ERR_STATE, "read set indices must be contiguous from 0");
        if (!c->sets[m].injected && !c->counted) return fail(LSQ_ERR_STATE, "lsq_solve before lsq_count");
    }
    LSQ_CUDA(cudaSetDevice(c->device));
    int rc = ensure_run_buffers(c);
    if (rc) return rc;
    cudaStream_t s = c->stream;
    rc = run_finalize(c);
    if (rc) return rc;
    lsq::SolveParams P;
    fill_solve_params(c, P);
    double mb = 0.0;  // solve.cpp:810-812
    for (uint32_t m = 0; m < c->M; ++m) mb += total_read_bases[m] / 1.0E6;
    P.total_read_mbases = mb;
    bool small = c->M <= 2;
    for (uint32_t m = 0; m < c->M; ++m) small = small && !c->sets[m].injected;
    P.small_enabled = small ? 1 : 0;
    P.work_counter = reinterpret_cast<unsigned int *>(c->d_counters.as<unsigned long long>() + LSQ_MAX_METHODS * lsq::CTR_N);
    LSQ_CUDA(cudaEventRecord(c->ev[2], s));
    LSQ_CUDA(cudaMemsetAsync(c->d_flag.p, 0, (size_t)(2 * std::max<uint32_t>(c->E, 1) + 8) * 4, s));
    if (c->E && small) {
        LSQ_CUDA(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), s));
        unsigned grid = std::min<unsigned>((c->E + 127) / 128, (unsigned)c->n_sms * 8u);
        if (c->M == 1) lsq::lsq_solve_small_kernel<1><<<grid, 128, 0, s>>>(P);
        else lsq::lsq_solve_small_kernel<2><<<grid, 128, 0, s>>>(P);
        LSQ_CUDA(cudaGetLastError());
        c->tm.launches++;
    }
    if (c->E && (!small || c->H.max_forms > 2)) {
        // events with up to 16 forms: one thread per event; the warp-per-event kernel then adds their standard errors and
        // solves the larger events
        P.mid_enabled = getenv("LSQ_NO_MID_SOLVE") ? 0 : 1;
        if (P.mid_enabled) {
            P.work_counter = P.work_counter + 1;  // its own fetch counter
            LSQ_CUDA(cudaMemsetAsync(P.work_counter, 0, sizeof(unsigned int), s));
            // work order: counting sort of the events by (classes, forms)
            LSQ_CUDA(c->d_mid.ensure(((size_t)lsq::MID_KEY_BINS + 4 + std::max<uint32_t>(c->E, 1)) * 4));
            uint32_t *hist = c->d_mid.as<uint32_t>(), *n_mid = hist + lsq::MID_KEY_BINS, *order = n_mid + 4;
            LSQ_CUDA(cudaMemsetAsync(hist, 0, (lsq::MID_KEY_BINS + 4) * 4, s));
            lsq::lsq_mid_key_kernel<<<(c->E + 255) / 256, 256, 0, s>>>(P, hist);
            lsq::lsq_mid_scan_kernel<<<1, 1024, 0, s>>>(hist, n_mid);
            lsq::lsq_mid_scatter_kernel<<<(c->E + 255) / 256, 256, 0, s>>>(P, hist, order);
            LSQ_CUDA(cudaGetLastError());
            c->tm.launches += 3;
            P.mid_order = order;
            P.mid_n = n_mid;
            unsigned grid = std::min<unsigned>((c->E + lsq::MID_BLOCK - 1) / lsq::MID_BLOCK, (unsigned)c->n_sms * 4u);
            lsq::lsq_solve_mid_kernel<<<grid, lsq::MID_BLOCK, 0, s>>>(P);
            LSQ_CUDA(cudaGetLastError());
            c->tm.launches++;
        }
        constexpr int WARPS = 8;
        unsigned grid = std::min<unsigned>((c->E + WARPS - 1) / WARPS, (unsigned)c->n_sms * 8u);
        lsq::lsq_solve_kernel<WARPS><<<grid, WARPS * 32, 0, s>>>(P);
        LSQ_CUDA(cudaGetLastError());
        c->tm.launches++;
        grid = std::min<unsigned>((c->E + lsq::SE_WARPS - 1) / lsq::SE_WARPS, (unsigned)c->n_sms * 8u);
        lsq::lsq_se_kernel<<<grid, lsq::SE_WARPS * 32, 0, s>>>(P);
        LSQ_CUDA(cudaGetLastError());
        c->tm.launches++;
    }
    // events whose stop decision is within summation-order noise of its threshold are replayed read by read (rare);
    // this is the one host round trip of the lsq_count + lsq_solve pair
    c->exact_events = 0;
    c->exact_ms = 0.f;
    if (c->E && !getenv("LSQ_NO_EXACT_EM")) {
        uint32_t n_flag = 0;
        LSQ_CUDA(cudaMemcpyAsync(&n_flag, P.flag_n, 4, cudaMemcpyDeviceToHost, s));
        LSQ_CUDA(cudaStreamSynchronize(s));
        if (n_flag) {
            const auto t0 = std::chrono::steady_clock::now();
            rc = exact_replay(c, P, n_flag);
            if (rc) return rc;
            c->exact_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    }
    LSQ_CUDA(cudaEventRecord(c->ev[3], s));
    c->tm_solve_pending = true;
    c->solved = true;
    return LSQ_OK;
}

int lsq_get_solve(lsq_ctx *c, lsq_solve_out *o) {
    if (!c || !o) return fail(LSQ_ERR_ARG, "lsq_get_solve: null argument");
    if (!c->solved) return fail(LSQ_ERR_STATE, "lsq_get_solve before lsq_solve");
    LSQ_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t E = c->E, F = c->F;
    if (o->theta && F) LSQ_CUDA(cudaMemcpyAsync(o->theta, c->d_theta.p, F * 8, cudaMemcpyDeviceToHost, s));
    if (o->rpkm && F) LSQ_CUDA(cudaMemcpyAsync(o->rpkm, c->d_rpkm.p, F * 8, cudaMemcpyDeviceToHost, s));
    if (o->loglik && E) LSQ_CUDA(cudaMemcpyAsync(o->loglik, c->d_ll.p, E * 8, cudaMemcpyDeviceToHost, s));
    if (o->ll_per_read && E) LSQ_CUDA(cudaMemcpyAsync(o->ll_per_read, c->d_llpr.p, E * 8, cudaMemcpyDeviceToHost, s));
    if (o->iters && E) LSQ_CUDA(cudaMemcpyAsync(o->iters, c->d_iters.p, E * 4, cudaMemcpyDeviceToHost, s));
    if (o->se && F) LSQ_CUDA(cudaMemcpyAsync(o->se, c->d_se.p, F * 8, cudaMemcpyDeviceToHost, s));
    LSQ_CUDA(cudaStreamSynchronize(s));
    return LSQ_OK;
}

int lsq_get_device_results(lsq_ctx *c, lsq_device_results *o) {
    if (!c || !o) return fail(LSQ_ERR_ARG, "lsq_get_device_results: null argument");
    if (!c->finalized) return fail(LSQ_ERR_STATE, "lsq_get_device_results before lsq_count");
    o->n_valid = c->d_n_valid.as<uint64_t>();
    o->bases = c->d_bases_out.as<uint64_t>();
    o->iso_count = c->d_iso.as<uint64_t>();
    o->theta = c->d_theta.as<double>();
    o->rpkm = c->d_rpkm.as<double>();
    o->loglik = c->d_ll.as<double>();
    o->ll_per_read = c->d_llpr.as<double>();
    o->iters = c->d_iters.as<uint32_t>();
    return LSQ_OK;
}

int lsq_measure_fp64_peak(lsq_ctx *c, double *tflops) {
    if (!c || !tflops) return fail(LSQ_ERR_ARG, "lsq_measure_fp64_peak: null argument");
    LSQ_CUDA(cudaSetDevice(c->device));
    const int blocks = c->n_sms * 8, iters = 100000;
    DevBuf out;
    LSQ_CUDA(out.ensure((size_t)blocks * 256 * 8));
    cudaStream_t s = c->stream;
    lsq::lsq_dfma_kernel<<<blocks, 256, 0, s>>>(out.as<double>(), 1000);  // warm-up
    LSQ_CUDA(cudaEventRecord(c->ev[2], s));
    lsq::lsq_dfma_kernel<<<blocks, 256, 0, s>>>(out.as<double>(), iters);
    LSQ_CUDA(cudaEventRecord(c->ev[3], s));
    LSQ_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    LSQ_CUDA(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
    *tflops = 2.0 * blocks * 256.0 * (double)iters * 8.0 / (ms * 1e-3) / 1e12;
    return LSQ_OK;
}

int lsq_get_timings(lsq_ctx *c, lsq_timings *o) {
    if (!c || !o) return fail(LSQ_ERR_ARG, "lsq_get_timings: null argument");
    LSQ_CUDA(cudaSetDevice(c->device));
    if (c->tm_count_pending || c->tm_solve_pending) LSQ_CUDA(cudaStreamSynchronize(c->stream));
    lsq_timings &tm = c->tm;
    if (c->tm_count_pending) {
        std::vector<unsigned long long> hc((size_t)LSQ_MAX_METHODS * lsq::CTR_N + 4, 0);
        LSQ_CUDA(cudaMemcpy(hc.data(), c->d_counters.p, hc.size() * 8, cudaMemcpyDeviceToHost));
        tm.clear_ms = tm.scan_ms = tm.expand_ms = tm.general_ms = 0.f;
        tm.reads_single = tm.reads_pair = tm.reads_triple = tm.reads_general = tm.hash_full = 0;
        for (uint32_t m = 0; m < c->M; ++m) {
            if (!c->set_counted[m]) continue;
            float f;
            LSQ_CUDA(cudaEventElapsedTime(&f, c->ev_set[m][0], c->ev_set[m][1]));
            tm.clear_ms += f;
            LSQ_CUDA(cudaEventElapsedTime(&f, c->ev_set[m][1], c->ev_set[m][2]));
            tm.scan_ms += f;
            LSQ_CUDA(cudaEventElapsedTime(&f, c->ev_set[m][2], c->ev_set[m][3]));
            tm.expand_ms += f;
            LSQ_CUDA(cudaEventElapsedTime(&f, c->ev_set[m][3], c->ev_set[m][4]));
            tm.general_ms += f;
            const unsigned long long *k = hc.data() + (size_t)m * lsq::CTR_N;
            tm.reads_single += k[lsq::CTR_SINGLE_SIG];
            tm.reads_pair += k[lsq::CTR_PAIR_SIG];
            tm.reads_triple += k[lsq::CTR_TRIPLE_SIG];
            tm.reads_general += k[lsq::CTR_GENERAL];
            tm.hash_full += k[lsq::CTR_HASH_FULL];
        }
        tm.pairs_valid = hc[(size_t)LSQ_MAX_METHODS * lsq::CTR_N + 2];
        LSQ_CUDA(cudaEventElapsedTime(&tm.finalize_ms, c->ev[0], c->ev[1]));
        c->tm_count_pending = false;
    }
    if (c->tm_solve_pending) {
        tm.exact_events = c->exact_events;
        tm.exact_ms = c->exact_ms;
        LSQ_CUDA(cudaEventElapsedTime(&tm.solve_ms, c->ev[2], c->ev[3]));
        c->tm_solve_pending = false;
    }
    *o = tm;
    return LSQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU: NCCL is loaded on first use, so that single-GPU callers do not need it installed.
// ---------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load() {
    if (g_nccl.h) return LSQ_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(LSQ_ERR_CUDA, std::string("NCCL not available: ") + dlerror());
#define LSQ_SYM(field, name)                                                              \
    *reinterpret_cast<void **>(&g_nccl.field) = dlsym(h, name);                          \
    if (!g_nccl.field) return fail(LSQ_ERR_CUDA, std::string("NCCL symbol missing: ") + name)
    LSQ_SYM(GetUniqueId, "ncclGetUniqueId");
    LSQ_SYM(CommInitRank, "ncclCommInitRank");
    LSQ_SYM(CommDestroy, "ncclCommDestroy");
    LSQ_SYM(GroupStart, "ncclGroupStart");
    LSQ_SYM(GroupEnd, "ncclGroupEnd");
    LSQ_SYM(Send, "ncclSend");
    LSQ_SYM(Recv, "ncclRecv");
    LSQ_SYM(AllGather, "ncclAllGather");
    LSQ_SYM(GetErrorString, "ncclGetErrorString");
#undef LSQ_SYM
    g_nccl.h = h;
    return LSQ_OK;
}

#define LSQ_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r_ = (expr);                                                                        \
        if (r_ != ncclSuccess) return fail(LSQ_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

}  // namespace

struct lsq_comm {
    ncclComm_t comm = nullptr;
    int n_ranks = 0, rank = 0, device = 0;
    bool owned = false;
};

extern "C" {

int lsq_comm_unique_id(void *id) {
    if (!id) return fail(LSQ_ERR_ARG, "lsq_comm_unique_id: null argument");
    static_assert(sizeof(ncclUniqueId) == LSQ_COMM_ID_BYTES, "ncclUniqueId size");
    int rc = nccl_load();
    if (rc) return rc;
    LSQ_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId *>(id)));
    return LSQ_OK;
}

int lsq_comm_create(const void *id, int n_ranks, int rank, int device, lsq_comm **out) {
    if (!id || !out) return fail(LSQ_ERR_ARG, "lsq_comm_create: null argument");
    if (n_ranks < 1 || n_ranks > (int)LSQ_MAX_RANKS || rank < 0 || rank >= n_ranks)
        return fail(LSQ_ERR_RANGE, "lsq_comm_create: 1..16 ranks");
    int rc = nccl_load();
    if (rc) return rc;
    LSQ_CUDA(cudaSetDevice(device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm;
    LSQ_NCCL(g_nccl.CommInitRank(&comm, n_ranks, uid, rank));
    lsq_comm *c = new lsq_comm();
    c->comm = comm;
    c->n_ranks = n_ranks;
    c->rank = rank;
    c->device = device;
    c->owned = true;
    *out = c;
    return LSQ_OK;
}

int lsq_comm_attach(void *nccl_comm, int n_ranks, int rank, int device, lsq_comm **out) {
    if (!nccl_comm || !out) return fail(LSQ_ERR_ARG, "lsq_comm_attach: null argument");
    if (n_ranks < 1 || n_ranks > (int)LSQ_MAX_RANKS || rank < 0 || rank >= n_ranks)
        return fail(LSQ_ERR_RANGE, "lsq_comm_attach: 1..16 ranks");
    int rc = nccl_load();
    if (rc) return rc;
    lsq_comm *c = new lsq_comm();
    c->comm = reinterpret_cast<ncclComm_t>(nccl_comm);
    c->n_ranks = n_ranks;
    c->rank = rank;
    c->device = device;
    c->owned = false;
    *out = c;
    return LSQ_OK;
}

int lsq_comm_destroy(lsq_comm *c) {
    if (!c) return LSQ_OK;
    if (c->owned && c->comm && g_nccl.h) g_nccl.CommDestroy(c->comm);
    delete c;
    return LSQ_OK;
}

int lsq_router_build(lsq_ctx *c, uint32_t n_ranks, const uint32_t *shard_begin) {
    if (!c || !shard_begin) return fail(LSQ_ERR_ARG, "lsq_router_build: null argument");
    if (!c->have_tables) return fail(LSQ_ERR_STATE, "lsq_router_build before lsq_upload_events");
    if (n_ranks < 1 || n_ranks > LSQ_MAX_RANKS) return fail(LSQ_ERR_RANGE, "lsq_router_build: 1..16 ranks");
    if (shard_begin[0] != 0 || shard_begin[n_ranks] != c->E) return fail(LSQ_ERR_ARG, "lsq_router_build: shard_begin must span all events");
    for (uint32_t d = 0; d < n_ranks; ++d)
        if (shard_begin[d] > shard_begin[d + 1]) return fail(LSQ_ERR_ARG, "lsq_router_build: shard_begin must ascend");
    LSQ_CUDA(cudaSetDevice(c->device));
    const lsq::HostTables &H = c->H;
    c->route_ranks = n_ranks;
    c->shard_begin.assign(shard_begin, shard_begin + n_ranks + 1);
    // forms per shard: the events' form ranges in output order
    std::vector<uint32_t> form_off(c->E + 1, 0);
    for (uint32_t x = 0; x < c->E; ++x) form_off[H.hdr[x].out_ev + 1] = H.hdr[x].K;
    for (uint32_t e = 0; e < c->E; ++e) form_off[e + 1] += form_off[e];
    c->form_begin.resize(n_ranks + 1);
    for (uint32_t d = 0; d <= n_ranks; ++d) c->form_begin[d] = form_off[shard_begin[d]];
    // owner mask of every piece
    const size_t np = H.bp_pos.size();
    std::vector<unsigned long long> owner(np + 1, 0ull);
    for (size_t p = 0; p < np; ++p)
        for (uint32_t q = H.stab_off[p]; q < H.stab_off[p + 1]; ++q) {
            const uint32_t e = H.hdr[H.stab[q].x].out_ev;
            const uint32_t d = (uint32_t)(std::upper_bound(shard_begin, shard_begin + n_ranks + 1, e) - shard_begin) - 1u;
            owner[p] |= 1ull << std::min(d, n_ranks - 1);
        }
    LSQ_CUDA(upload(c->d_owner, owner, c->stream));
    LSQ_CUDA(c->d_cursor.ensure((size_t)(LSQ_MAX_RANKS + LSQ_MAX_RANKS * LSQ_MAX_RANKS + 2 + LSQ_MAX_RANKS) * 8));
    LSQ_CUDA(cudaStreamSynchronize(c->stream));
    return LSQ_OK;
}

int lsq_shard_reads(lsq_ctx *full, uint32_t m_src, lsq_ctx *shard, uint32_t m_dst, lsq_comm *comm, uint64_t first_line,
                    lsq_shard_stats *stats) {
    if (!full || !shard || !comm) return fail(LSQ_ERR_ARG, "lsq_shard_reads: null argument");
    if (full->route_ranks != (uint32_t)comm->n_ranks) return fail(LSQ_ERR_STATE, "lsq_shard_reads: lsq_router_build with the communicator's rank count first");
    if (m_src >= LSQ_MAX_METHODS || !full->sets[m_src].declared || full->sets[m_src].injected)
        return fail(LSQ_ERR_STATE, "lsq_shard_reads: read set m_src is not declared on the full context");
    if (full->device != shard->device) return fail(LSQ_ERR_ARG, "lsq_shard_reads: both contexts must be on the communicator's device");
    LSQ_CUDA(cudaSetDevice(full->device));
    ReadSet &rs = full->sets[m_src];
    cudaStream_t s = full->stream;
    const uint32_t P = (uint32_t)comm->n_ranks, me = (uint32_t)comm->rank;
    const size_t n = rs.n_reads, nb = rs.n_blocks;
    if (first_line + n > 0xFFFFFFFFull) return fail(LSQ_ERR_RANGE, "line numbers exceed 32 bits");
    // The exchange format must be the same on every rank (each posts the sends its peers expect): the narrowest one
    // all slices fit.  One tiny all-gather of the local capabilities.
    bool compact_blk = rs.blk_start != nullptr, len8 = rs.blk_len8 != nullptr, cs8 = rs.read_cs8 != nullptr;
    bool mixed = rs.blk_cs != nullptr;
    {
        unsigned long long *d_fmt = full->d_cursor.as<unsigned long long>() + LSQ_MAX_RANKS + LSQ_MAX_RANKS * LSQ_MAX_RANKS;
        const unsigned long long mine = (compact_blk ? 1ull : 0ull) | (len8 ? 2ull : 0ull) | (cs8 ? 4ull : 0ull) | (mixed ? 8ull : 0ull);
        LSQ_CUDA(cudaMemcpyAsync(d_fmt, &mine, 8, cudaMemcpyHostToDevice, s));
        LSQ_NCCL(g_nccl.AllGather(d_fmt, d_fmt + 1, 1, ncclUint64, comm->comm, s));
        unsigned long long all[LSQ_MAX_RANKS];
        LSQ_CUDA(cudaMemcpyAsync(all, d_fmt + 1, (size_t)P * 8, cudaMemcpyDeviceToHost, s));
        LSQ_CUDA(cudaStreamSynchronize(s));
        for (uint32_t d = 0; d < P; ++d) {
            compact_blk = compact_blk && (all[d] & 1ull);
            len8 = len8 && (all[d] & 2ull);
            cs8 = cs8 && (all[d] & 4ull);
            mixed = mixed || (all[d] & 8ull);
        }
        len8 = len8 && compact_blk;  // else 16-bit lengths when every slice is compact, wide blocks otherwise
    }
    const size_t blk_bytes = compact_blk ? 4 : 8, len_bytes = compact_blk ? (len8 ? 1 : 2) : 0, cs_bytes = cs8 ? 1 : 4;
    // outboxes: worst case every read goes to one destination
    lsq::RouteParams R;
    memset(&R, 0, sizeof R);
    R.n_ranks = P;
    R.line_base = (uint32_t)(first_line - 1);
    R.owner = full->d_owner.as<unsigned long long>();
    R.cursor = full->d_cursor.as<unsigned long long>();
    for (uint32_t d = 0; d < P; ++d) {
        DevBuf *o = full->d_out[d];
        LSQ_CUDA(o[0].ensure(n + 16));
        LSQ_CUDA(o[1].ensure(n * cs_bytes + 16));
        LSQ_CUDA(o[2].ensure(n * 4 + 16));
        LSQ_CUDA(o[3].ensure(nb * blk_bytes + 16));
        if (len_bytes) LSQ_CUDA(o[4].ensure(nb * len_bytes + 16));
        if (mixed) LSQ_CUDA(o[5].ensure(nb * 4 + 16));
        lsq::RouteBox &B = R.box[d];
        B.blk_cnt = o[0].as<uint8_t>();
        if (cs8) B.read_cs8 = o[1].as<uint8_t>();
        else B.read_cs = o[1].as<uint32_t>();
        B.line_num = o[2].as<uint32_t>();
        if (compact_blk) {
            B.blk_start = o[3].as<int32_t>();
            if (len8) B.blk_len8 = o[4].as<uint8_t>();
            else B.blk_len = o[4].as<uint16_t>();
        } else {
            B.blk = o[3].as<int2>();
        }
        if (mixed) B.blk_cs = o[5].as<uint32_t>();
    }
    lsq::ScanParams SP;
    memset(&SP, 0, sizeof SP);
    SP.T = full->T;
    SP.n_reads = n;
    SP.blk_off = rs.blk_cnt ? nullptr : rs.blk_off;
    SP.blk_cnt = rs.blk_cnt;
    SP.grp_base = rs.d_grp_base.as<uint32_t>();
    SP.sup_base = rs.d_sup_base.as<uint32_t>();
    SP.blk = rs.blk;
    SP.read_cs = rs.read_cs;
    SP.read_cs8 = rs.read_cs8;
    SP.blk_start = rs.blk_start;
    SP.blk_len = rs.blk_len;
    SP.blk_len8 = rs.blk_len8;
    SP.blk_cs = rs.blk_cs;
    SP.chrom_remap = rs.d_chrom_remap.as<uint32_t>();
    SP.file_meta = rs.d_file_meta.as<uint4>();
    SP.line_num = rs.line_num;
    cudaEvent_t *ev = full->ev_set[m_src];
    if (rs.copy_pending) {  // a streamed slice: the router needs all of it
        LSQ_CUDA(cudaStreamWaitEvent(s, full->ev_all[m_src], 0));
    }
    LSQ_CUDA(cudaEventRecord(ev[0], s));
    LSQ_CUDA(cudaMemsetAsync(R.cursor, 0, LSQ_MAX_RANKS * 8, s));
    if (n) {
        const uint64_t n_groups = (n + lsq::SCAN_GROUP - 1) / lsq::SCAN_GROUP;
        const unsigned grid = (unsigned)std::min<uint64_t>((n_groups + lsq::SCAN_WARPS - 1) / lsq::SCAN_WARPS, (uint64_t)full->n_sms * 4);
        lsq::lsq_route_kernel<<<grid, lsq::SCAN_BLOCK, 0, s>>>(SP, R);
        LSQ_CUDA(cudaGetLastError());
    }
    LSQ_CUDA(cudaEventRecord(ev[1], s));
    // sizes: every rank learns how much every rank sends to every rank
    unsigned long long *d_mat = R.cursor + LSQ_MAX_RANKS;
    LSQ_NCCL(g_nccl.AllGather(R.cursor, d_mat, LSQ_MAX_RANKS, ncclUint64, comm->comm, s));
    std::vector<unsigned long long> mat((size_t)P * LSQ_MAX_RANKS);
    LSQ_CUDA(cudaMemcpyAsync(mat.data(), d_mat, mat.size() * 8, cudaMemcpyDeviceToHost, s));
    LSQ_CUDA(cudaStreamSynchronize(s));
    if (rs.copy_pending) rs.copy_pending = false;
    auto sent_reads = [&](uint32_t src, uint32_t dst) { return (size_t)(mat[(size_t)src * LSQ_MAX_RANKS + dst] >> 32); };
    auto sent_blocks = [&](uint32_t src, uint32_t dst) { return (size_t)(mat[(size_t)src * LSQ_MAX_RANKS + dst] & 0xFFFFFFFFull); };
    size_t in_reads = 0, in_blocks = 0, out_reads = 0;
    std::vector<size_t> roff(P + 1, 0), boff(P + 1, 0);
    for (uint32_t src = 0; src < P; ++src) {
        roff[src] = in_reads;
        boff[src] = in_blocks;
        in_reads += sent_reads(src, me);
        in_blocks += sent_blocks(src, me);
    }
    for (uint32_t d = 0; d < P; ++d) out_reads += sent_reads(me, d);
    if (in_blocks + 2 * in_reads > 0xFFFFFFF0ull) return fail(LSQ_ERR_RANGE, "received read set exceeds 2^32 (blocks + 2 * reads)");
    DevBuf *in = full->d_in;
    LSQ_CUDA(in[0].ensure(in_reads + 16));
    LSQ_CUDA(in[1].ensure(in_reads * cs_bytes + 16));
    LSQ_CUDA(in[2].ensure(in_reads * 4 + 16));
    LSQ_CUDA(in[3].ensure(in_blocks * blk_bytes + 16));
    if (len_bytes) LSQ_CUDA(in[4].ensure(in_blocks * len_bytes + 16));
    if (mixed) LSQ_CUDA(in[5].ensure(in_blocks * 4 + 16));
    LSQ_CUDA(cudaEventRecord(ev[2], s));
    LSQ_NCCL(g_nccl.GroupStart());
    for (uint32_t peer = 0; peer < P; ++peer) {
        const size_t sr = sent_reads(me, peer), sb = sent_blocks(me, peer), rr = sent_reads(peer, me), rb = sent_blocks(peer, me);
        DevBuf *o = full->d_out[peer];
        const size_t unit[6] = {1, cs_bytes, 4, blk_bytes, len_bytes, (size_t)(mixed ? 4 : 0)};
        for (int a = 0; a < 6; ++a) {
            if (!unit[a]) continue;
            const bool per_read = a < 3;
            const size_t sbytes = (per_read ? sr : sb) * unit[a], rbytes = (per_read ? rr : rb) * unit[a];
            const size_t roffb = (per_read ? roff[peer] : boff[peer]) * unit[a];
            if (peer == me) {  // this rank's own bucket never leaves the device
                if (sbytes) LSQ_CUDA(cudaMemcpyAsync(in[a].as<unsigned char>() + roffb, o[a].p, sbytes, cudaMemcpyDeviceToDevice, s));
                continue;
            }
            if (sbytes) LSQ_NCCL(g_nccl.Send(o[a].p, sbytes, ncclUint8, (int)peer, comm->comm, s));
            if (rbytes) LSQ_NCCL(g_nccl.Recv(in[a].as<unsigned char>() + roffb, rbytes, ncclUint8, (int)peer, comm->comm, s));
        }
    }
    LSQ_NCCL(g_nccl.GroupEnd());
    LSQ_CUDA(cudaEventRecord(ev[3], s));
    LSQ_CUDA(cudaStreamSynchronize(s));
    // what arrived is read set m_dst of the shard context (borrowed device arrays)
    lsq_packed_reads v;
    memset(&v, 0, sizeof v);
    v.n_reads = in_reads;
    v.n_blocks = in_blocks;
    v.blk_cnt = in[0].as<uint8_t>();
    if (cs8) v.read_cs8 = in[1].as<uint8_t>();
    else v.read_cs = in[1].as<uint32_t>();
    v.line_num = in[2].as<uint32_t>();
    if (compact_blk) {
        v.blk_start = in[3].as<int32_t>();
        if (len8) v.blk_len8 = in[4].as<uint8_t>();
        else v.blk_len = in[4].as<uint16_t>();
    } else {
        v.blk = in[3].as<int32_t>();
    }
    if (mixed) v.blk_cs = in[5].as<uint32_t>();
    std::vector<const char *> cn(full->sets[m_src].chrom_names.size()), sn(full->sets[m_src].strand_names.size());
    for (size_t i = 0; i < cn.size(); ++i) cn[i] = full->sets[m_src].chrom_names[i].c_str();
    for (size_t i = 0; i < sn.size(); ++i) sn[i] = full->sets[m_src].strand_names[i].c_str();
    v.n_chrom_names = (uint32_t)cn.size();
    v.n_strand_names = (uint32_t)sn.size();
    v.chrom_names = cn.data();
    v.strand_names = sn.data();
    v.on_device = 1;
    int rc = lsq_set_reads(shard, m_dst, rs.type, rs.read_length, &v);
    if (rc) return rc;
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->reads_local = n;
        stats->reads_sent = out_reads;
        stats->reads_received = in_reads;
        stats->blocks_received = in_blocks;
        LSQ_CUDA(cudaEventElapsedTime(&stats->route_ms, ev[0], ev[1]));
        LSQ_CUDA(cudaEventElapsedTime(&stats->exchange_ms, ev[2], ev[3]));
    }
    return LSQ_OK;
}

int lsq_gather(lsq_ctx *full, lsq_ctx *shard, lsq_comm *comm, int root, lsq_count_out *count, lsq_solve_out *solve) {
    if (!full || !shard || !comm) return fail(LSQ_ERR_ARG, "lsq_gather: null argument");
    if (full->route_ranks != (uint32_t)comm->n_ranks) return fail(LSQ_ERR_STATE, "lsq_gather: lsq_router_build first");
    if (!shard->finalized) return fail(LSQ_ERR_STATE, "lsq_gather before lsq_count");
    LSQ_CUDA(cudaSetDevice(shard->device));
    const uint32_t P = (uint32_t)comm->n_ranks, me = (uint32_t)comm->rank, M = shard->M;
    const std::vector<uint32_t> &eb = full->shard_begin, &fb = full->form_begin;
    const size_t E = full->E, F = full->F;
    if (shard->E != eb[me + 1] - eb[me] || shard->F != fb[me + 1] - fb[me]) return fail(LSQ_ERR_ARG, "lsq_gather: the shard context does not hold this rank's events");
    cudaStream_t s = shard->stream;
    const bool solved = shard->solved;
    // root buffer: n_valid [M][E] | bases [M][E] | iso [F] | theta [F] | rpkm [F] | se [F] | ll [E] | llpr [E] | iters [E] (u32)
    const size_t o_nv = 0, o_bs = o_nv + M * E * 8, o_iso = o_bs + M * E * 8, o_th = o_iso + F * 8, o_rp = o_th + F * 8, o_se = o_rp + F * 8,
                 o_ll = o_se + F * 8, o_lp = o_ll + E * 8, o_it = o_lp + E * 8, total = o_it + E * 4 + 16;
    unsigned char *g = nullptr;
    if ((int)me == root) {
        LSQ_CUDA(full->d_gather.ensure(total));
        g = full->d_gather.as<unsigned char>();
    }
    struct Arr {
        const void *src;
        size_t off, unit;
        bool per_event, per_m;
    };
    const Arr arrs[9] = {{shard->d_n_valid.p, o_nv, 8, true, true}, {shard->d_bases_out.p, o_bs, 8, true, true},
                         {shard->d_iso.p, o_iso, 8, false, false},  {shard->d_theta.p, o_th, 8, false, false},
                         {shard->d_rpkm.p, o_rp, 8, false, false},   {shard->d_se.p, o_se, 8, false, false},
                         {shard->d_ll.p, o_ll, 8, true, false},      {shard->d_llpr.p, o_lp, 8, true, false},
                         {shard->d_iters.p, o_it, 4, true, false}};
    LSQ_NCCL(g_nccl.GroupStart());
    for (int a = 0; a < 9; ++a) {
        if (a >= 3 && !solved) continue;
        const Arr &A = arrs[a];
        const size_t mine = A.per_event ? shard->E : shard->F;
        const uint32_t reps = A.per_m ? M : 1;
        for (uint32_t m = 0; m < reps; ++m) {
            if (mine) LSQ_NCCL(g_nccl.Send(static_cast<const unsigned char *>(A.src) + (size_t)m * mine * A.unit, mine * A.unit, ncclUint8, root, comm->comm, s));
            if ((int)me == root)
                for (uint32_t d = 0; d < P; ++d) {
                    const size_t cnt = A.per_event ? eb[d + 1] - eb[d] : fb[d + 1] - fb[d];
                    const size_t at = A.per_event ? eb[d] : fb[d];
                    const size_t row = A.per_event ? E : F;
                    if (cnt) LSQ_NCCL(g_nccl.Recv(g + A.off + ((size_t)m * row + at) * A.unit, cnt * A.unit, ncclUint8, (int)d, comm->comm, s));
                }
        }
    }
    LSQ_NCCL(g_nccl.GroupEnd());
    if ((int)me == root) {
        auto d2h = [&](void *dst, size_t off, size_t bytes) {
            return (dst && bytes) ? cudaMemcpyAsync(dst, g + off, bytes, cudaMemcpyDeviceToHost, s) : cudaSuccess;
        };
        if (count) {
            LSQ_CUDA(d2h(count->n_valid, o_nv, M * E * 8));
            LSQ_CUDA(d2h(count->bases, o_bs, M * E * 8));
            LSQ_CUDA(d2h(count->iso_count, o_iso, F * 8));
        }
        if (solve && solved) {
            LSQ_CUDA(d2h(solve->theta, o_th, F * 8));
            LSQ_CUDA(d2h(solve->rpkm, o_rp, F * 8));
            LSQ_CUDA(d2h(solve->se, o_se, F * 8));
            LSQ_CUDA(d2h(solve->loglik, o_ll, E * 8));
            LSQ_CUDA(d2h(solve->ll_per_read, o_lp, E * 8));
            LSQ_CUDA(d2h(solve->iters, o_it, E * 4));
        }
    }
    LSQ_CUDA(cudaStreamSynchronize(s));
    return LSQ_OK;
}

}  // extern "C"

namespace lsqh {
void set_error(const std::string &msg) { g_err = msg; }
}  // namespace lsqh
