// Host front end: on-disk packed cache of a parsed MRF_SINGLE read set.
//
// The reference parses the same text file once in `count` and again in `solve`
// (count/count.cpp:279-339, solve/solve.cpp:429-486); after the kernels, text parsing is what is
// left of the wall clock.  A pack file holds exactly the arrays lsqh_mrf_parse produces (the input
// of the covered-region filter, which depends on the gene range and therefore stays on the
// device), stamped with size and mtime of the text it was made from.  Opt-in: the drop-in programs
// only look for / write `<reads_path>.lsqpack` when LSQ_READ_CACHE=1, because the reference
// writes no files.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "../../../include/lesseq_b200.h"
#include "mrf.hpp"

namespace lsqh {

namespace {

struct PackHeader {
    char magic[8];
    uint32_t version, flags;  // flags bit 0: blk_cs stored (some read is LSQ_READ_MIXED)
    uint64_t n_reads, n_blocks;
    uint64_t src_size;
    int64_t src_mtime_ns;
    uint32_t n_chroms, n_strands;
    uint64_t check;       // FNV-1a over the per-piece hashes of the arrays, in file order
    uint64_t arrays_off;  // file offset of the first array (8-byte aligned)
};
const char MAGIC[8] = {'L', 'S', 'Q', 'P', 'A', 'C', 'K', '2'};
const size_t PIECE = 1u << 20;  // arrays are hashed (and loaded) in pieces of 1 MiB by all host threads

uint64_t fnv(uint64_t h, const void *p, size_t n) {
    // 8 bytes per step: a cache check, not a cryptographic hash
    const unsigned char *b = static_cast<const unsigned char *>(p);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t v;
        memcpy(&v, b + i, 8);
        h = (h ^ v) * 0x100000001B3ull;
    }
    for (; i < n; ++i) h = (h ^ b[i]) * 0x100000001B3ull;
    return h;
}
const uint64_t FNV0 = 0xCBF29CE484222325ull;

bool source_stamp(const char *path, uint64_t &size, int64_t &mtime_ns) {
    struct stat st;
    if (!path || stat(path, &st) != 0) return false;
    size = (uint64_t)st.st_size;
    mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    return true;
}

unsigned host_threads() {
    unsigned n = std::thread::hardware_concurrency();
    if (const char *e = getenv("LSQ_THREADS")) n = (unsigned)atoi(e);
    return std::max(1u, std::min(n, 64u));
}

// One array of the pack: where it lives in memory and how many bytes it has.
struct Span {
    const unsigned char *src;  // save: the array; load: the mapped file
    unsigned char *dst;        // load: the array to fill (null when saving)
    size_t bytes;
};

// Hash of every 1 MiB piece of every span (copying it to dst first when loading), computed by all host threads;
// the pack's checksum chains the piece hashes in file order.
uint64_t hash_spans(const std::vector<Span> &spans) {
    std::vector<std::pair<size_t, size_t> > pieces;  // (span, offset)
    for (size_t i = 0; i < spans.size(); ++i)
        for (size_t off = 0; off < spans[i].bytes; off += PIECE) pieces.push_back(std::make_pair(i, off));
    std::vector<uint64_t> h(pieces.size());
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (size_t k; (k = next.fetch_add(1)) < pieces.size();) {
            const Span &sp = spans[pieces[k].first];
            const size_t off = pieces[k].second, n = std::min(PIECE, sp.bytes - off);
            if (sp.dst) memcpy(sp.dst + off, sp.src + off, n);
            h[k] = fnv(FNV0, sp.src + off, n);
        }
    };
    const unsigned nt = (unsigned)std::min<size_t>(host_threads(), std::max<size_t>(pieces.size(), 1));
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &x : th) x.join();
    return fnv(FNV0, h.data(), h.size() * sizeof(uint64_t));
}

// fn(lo, hi) over [0, n) split evenly among the host threads
template <class F>
void parallel_ranges(size_t n, F fn) {
    const unsigned nt = (unsigned)std::min<size_t>(host_threads(), std::max<size_t>(n >> 16, 1));
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back([&fn, n, nt, t]() { fn(n * t / nt, n * (t + 1) / nt); });
    fn(0, n / nt);
    for (auto &x : th) x.join();
}

bool put_names(FILE *f, const std::vector<std::string> &names) {
    for (const std::string &s : names) {
        const uint32_t len = (uint32_t)s.size();
        if (fwrite(&len, 4, 1, f) != 1 || (len && fwrite(s.data(), 1, len, f) != len)) return false;
    }
    return true;
}
// names from a memory range; returns the position after them or null
const unsigned char *get_names(const unsigned char *p, const unsigned char *end, uint32_t n, Dictionary &d) {
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t len;
        if ((size_t)(end - p) < 4) return nullptr;
        memcpy(&len, p, 4);
        p += 4;
        if (len > (1u << 20) || (size_t)(end - p) < len) return nullptr;
        if (d.intern(std::string(reinterpret_cast<const char *>(p), len)) != i) return nullptr;  // duplicate: not ours
        p += len;
    }
    return p;
}

template <class V>
Span span_of(const V &v, size_t n) {
    return Span{reinterpret_cast<const unsigned char *>(v.data()), nullptr, n * sizeof(typename V::value_type)};
}

}  // namespace

int save_reads(const lsqh_reads &R, const char *pack_path, const char *source_path) {
    const size_t n = R.read_cs.size(), nb = R.blk_cs.size();
    if (R.blk_off.size() < n + 1 || R.line_num.size() < n || R.blk.size() < 2 * nb) {
        set_error("lsqh_reads_save: inconsistent read set");
        return LSQ_ERR_ARG;
    }
    PackHeader H;
    memset(&H, 0, sizeof H);
    memcpy(H.magic, MAGIC, 8);
    H.version = 2;
    H.flags = R.any_mixed ? 1u : 0u;
    H.n_reads = n;
    H.n_blocks = nb;
    H.n_chroms = (uint32_t)R.chroms.names.size();
    H.n_strands = (uint32_t)R.strands.names.size();
    if (source_path && !source_stamp(source_path, H.src_size, H.src_mtime_ns)) {
        set_error(std::string("cannot stat ") + source_path);
        return LSQ_ERR_IO;
    }
    std::vector<Span> spans = {span_of(R.blk_off, n + 1), span_of(R.read_cs, n), span_of(R.line_num, n), span_of(R.blk, 2 * nb)};
    if (R.any_mixed) spans.push_back(span_of(R.blk_cs, nb));
    H.check = hash_spans(spans);
    const std::string tmp = std::string(pack_path) + ".tmp." + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) {
        set_error("cannot create " + tmp);
        return LSQ_ERR_IO;
    }
    bool ok = fwrite(&H, sizeof H, 1, f) == 1 && put_names(f, R.chroms.names) && put_names(f, R.strands.names);
    if (ok) {  // arrays start on an 8-byte boundary
        long pos = ftell(f);
        static const char pad[8] = {0};
        const size_t npad = (size_t)((8 - pos % 8) % 8);
        ok = pos >= 0 && (npad == 0 || fwrite(pad, 1, npad, f) == npad);
        H.arrays_off = (uint64_t)pos + npad;
    }
    for (size_t i = 0; ok && i < spans.size(); ++i) ok = spans[i].bytes == 0 || fwrite(spans[i].src, 1, spans[i].bytes, f) == spans[i].bytes;
    if (ok) ok = fseek(f, 0, SEEK_SET) == 0 && fwrite(&H, sizeof H, 1, f) == 1;  // arrays_off is known now
    ok = (fclose(f) == 0) && ok;
    if (ok) ok = rename(tmp.c_str(), pack_path) == 0;
    if (!ok) {
        remove(tmp.c_str());
        set_error(std::string("cannot write ") + pack_path);
        return LSQ_ERR_IO;
    }
    return LSQ_OK;
}

int load_reads(const char *pack_path, const char *source_path, lsqh_reads &R) {
    int fd = open(pack_path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || (size_t)st.st_size < sizeof(PackHeader)) {
        if (fd >= 0) close(fd);
        set_error(std::string("cannot open ") + pack_path);
        return LSQ_ERR_IO;
    }
    const size_t fsize = (size_t)st.st_size;
    void *map = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) {
        set_error(std::string("cannot map ") + pack_path);
        return LSQ_ERR_IO;
    }
    madvise(map, fsize, MADV_WILLNEED);
    const unsigned char *base = static_cast<const unsigned char *>(map), *end = base + fsize;
    auto bad = [&](const char *why) {
        munmap(map, fsize);
        set_error(std::string(pack_path) + ": " + why);
        return LSQ_ERR_IO;
    };
    PackHeader H;
    memcpy(&H, base, sizeof H);
    if (memcmp(H.magic, MAGIC, 8) != 0 || H.version != 2) return bad("not a read pack");
    if (source_path) {
        uint64_t size;
        int64_t mtime;
        if (!source_stamp(source_path, size, mtime) || size != H.src_size || mtime != H.src_mtime_ns)
            return bad("stale (the text file changed)");
    }
    if (H.n_blocks + 2 * H.n_reads > 0xFFFFFFF0ull || H.n_chroms > 0x10000 || H.n_strands > 0x100) return bad("corrupt header");
    R = lsqh_reads();
    const unsigned char *p = get_names(base + sizeof H, end, H.n_chroms, R.chroms);
    if (p) p = get_names(p, end, H.n_strands, R.strands);
    if (!p) return bad("corrupt dictionaries");
    const size_t n = (size_t)H.n_reads, nb = (size_t)H.n_blocks;
    R.any_mixed = (H.flags & 1u) != 0;
    const size_t need = ((n + 1) + n + n + 2 * nb + (R.any_mixed ? nb : 0)) * 4;
    if (H.arrays_off % 8 != 0 || H.arrays_off < (uint64_t)(p - base) || H.arrays_off > fsize || fsize - H.arrays_off != need)
        return bad("truncated");
    R.blk_off.resize(n + 1);  // lsqh::Array: not zero-filled
    R.read_cs.resize(n);
    R.line_num.resize(n);
    R.blk.resize(2 * nb);
    if (R.any_mixed) R.blk_cs.resize(nb);
    std::vector<Span> spans;
    const unsigned char *q = base + H.arrays_off;
    auto take = [&](void *dst, size_t bytes) {
        spans.push_back(Span{q, static_cast<unsigned char *>(dst), bytes});
        q += bytes;
    };
    take(R.blk_off.data(), (n + 1) * 4);
    take(R.read_cs.data(), n * 4);
    take(R.line_num.data(), n * 4);
    take(R.blk.data(), 2 * nb * 4);
    if (R.any_mixed) take(R.blk_cs.data(), nb * 4);
    const uint64_t h = hash_spans(spans);  // copies and hashes, one piece per thread at a time
    munmap(map, fsize);
    map = nullptr;
    auto bad2 = [&](const char *why) {
        set_error(std::string(pack_path) + ": " + why);
        return LSQ_ERR_IO;
    };
    if (h != H.check) return bad2("checksum mismatch");
    if (R.blk_off[0] != 0 || R.blk_off[n] != nb) return bad2("corrupt offsets");
    std::atomic<bool> monotone(true);
    parallel_ranges(n, [&](size_t lo, size_t hi) {
        bool ok = true;
        for (size_t i = lo; i < hi; ++i) ok &= R.blk_off[i] <= R.blk_off[i + 1];
        if (!ok) monotone = false;
    });
    if (!monotone) return bad2("corrupt offsets");
    if (!R.any_mixed) {  // blk_cs always has one entry per block: the read's ids
        R.blk_cs.resize(nb);
        parallel_ranges(n, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
                for (uint32_t b = R.blk_off[i]; b < R.blk_off[i + 1]; ++b) R.blk_cs[b] = R.read_cs[i];
        });
    }
    R.chrom_ptrs = R.chroms.c_ptrs();
    R.strand_ptrs = R.strands.c_ptrs();
    return LSQ_OK;
}

}  // namespace lsqh
