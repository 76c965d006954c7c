// Host front end: MRF_SINGLE text -> packed reads (lsq_packed_reads).
//
// Grammar as consumed by count/count.cpp:279-339: the first line is skipped; every further
// '\n'-terminated line advances the read number (comment lines "#..." and lines equal to
// "AlignmentBlocks" too, :286-290); a read is a comma separated list of blocks
// chr:strand:start:end[:...]; start/end are 1-based inclusive and must be plain integers
// (boost::lexical_cast<long>, :312-318) else the run ends with exit status 1.  The block is
// stored as [start-1, end).  The covered-region filter of :319 is NOT applied here -- it
// depends on the selected gene range and runs on the device (filter kernel).
#include "mrf.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "annotation.hpp"

namespace lsqh {

namespace {

// lexical_cast<long>: optional sign, digits only, nothing else.
inline bool strict_long(const char *b, const char *e, long &out) {
    if (b >= e) return false;
    bool neg = false;
    if (*b == '+' || *b == '-') {
        neg = (*b == '-');
        ++b;
        if (b == e) return false;
    }
    unsigned long v = 0;
    for (; b < e; ++b) {
        unsigned d = (unsigned)(*b - '0');
        if (d > 9) return false;
        if (v > (unsigned long)LONG_MAX / 10) return false;
        v = v * 10 + d;
        if (v > (unsigned long)LONG_MAX + (neg ? 1 : 0)) return false;
    }
    out = neg ? -(long)v : (long)v;
    return true;
}

inline const char *find_ch(const char *b, const char *e, char c) {
    if (b >= e) return nullptr;
    return (const char *)memchr(b, c, e - b);
}
// A numeric field that starts at b and ends at the next ':' or at the end of the line: returns where it ends
// (':' or le) with the value in `out`, or nullptr if it is not a lexical_cast<long> number.  Fields of up to 18
// digits cannot overflow and are converted on the fly; longer ones take strict_long.
inline const char *field_long(const char *b, const char *le, long &out) {
    const char *q = b;
    bool neg = false;
    if (q < le && (*q == '+' || *q == '-')) neg = (*q++ == '-');
    const char *d0 = q;
    unsigned long v = 0;
    while (q < le) {
        const unsigned d = (unsigned)(*q - '0');
        if (d > 9) break;
        v = v * 10 + d;
        ++q;
    }
    if (q < le && *q != ':') return nullptr;  // something that is neither a digit nor the end of the field
    if (q == d0) return nullptr;              // empty field, or a sign only
    if (q - d0 > 18) return strict_long(b, q, out) ? q : nullptr;
    out = neg ? -(long)v : (long)v;
    return q;
}
inline bool same_bytes(const char *a, const char *b, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (a[i] != b[i]) return false;
    return true;
}
// the same for the few bytes of one field: a byte loop beats a library call
inline const char *scan_ch(const char *b, const char *e, char c) {
    while (b < e && *b != c) ++b;
    return b < e ? b : nullptr;
}

}  // namespace

namespace {

// Dictionary ids of the chromosome / strand strings of a chunk.  Reads arrive in no particular order, so the
// name changes on almost every line; a small direct-mapped table of the names seen (pointers into the mapped
// file) answers without building a std::string and hashing it.
struct NameCache {
    enum { N = 256 };
    const char *ptr[N];
    uint32_t len[N], id[N];
    NameCache() { memset(ptr, 0, sizeof ptr); }
    uint32_t lookup(const char *b, size_t n, Dictionary &d) {
        uint32_t h = (uint32_t)n * 0x9E3779B1u;
        for (size_t i = n > 4 ? n - 4 : 0; i < n; ++i) h = (h ^ (unsigned char)b[i]) * 0x01000193u;  // the last bytes differ
        const uint32_t k = (h >> 8) & (N - 1);
        if (ptr[k] && len[k] == n && same_bytes(ptr[k], b, n)) return id[k];
        const uint32_t v = d.intern(std::string(b, n));
        ptr[k] = b;
        len[k] = (uint32_t)n;
        id[k] = v;
        return v;
    }
};

struct ChunkOut {  // per-thread result, merged in file order
    Dictionary chroms, strands;
    std::vector<uint32_t> nblk, read_cs, blk_cs, line_num;
    std::vector<int32_t> blk;
    bool any_mixed = false;
    uint64_t n_lines = 0;  // '\n'-terminated lines of the chunk
    int err = LSQ_OK;      // LSQ_ERR_PARSE: not a number; LSQ_ERR_RANGE: coordinate outside int32
    uint64_t err_line = 0; // chunk-local number of the offending line
};

// Parses the '\n'-terminated lines of [p, end).  Line numbers are local to the chunk (first line = 1); the merge
// adds the number of lines before the chunk, so the file is not scanned twice.
void parse_chunk(const char *p, const char *end, ChunkOut &R) {
    uint64_t line_num = 0;
    NameCache chrom_cache, strand_cache;
    {   // room for the whole chunk at typical line lengths; growth beyond that is the vector's
        const size_t bytes = (size_t)(end - p);
        const size_t er = bytes / 24 + 16, eb = bytes / 20 + 16;
        R.nblk.reserve(er);
        R.read_cs.reserve(er);
        R.line_num.reserve(er);
        R.blk_cs.reserve(eb);
        R.blk.reserve(2 * eb);
    }
    auto failed = [&](int code) {
        R.err = code;
        R.err_line = line_num;
        R.n_lines = line_num;
    };
    while (p < end) {
        const char *le = find_ch(p, end, '\n');
        if (!le) break;  // unterminated last line is never processed (:285)
        ++line_num;
        const char *ls = p;
        p = le + 1;
        if (ls < le && *ls == '#') continue;
        if ((size_t)(le - ls) == 15 && memcmp(ls, "AlignmentBlocks", 15) == 0) continue;
        if (line_num > UINT32_MAX) {  // reported as "too many lines" (err_line 0)
            failed(LSQ_ERR_RANGE);
            R.err_line = 0;
            return;
        }
        uint32_t first_cs = 0;
        bool mixed = false;
        uint32_t nb = 0;
        const char *b0 = ls;
        while (true) {
            // chr ':' strand ':' start ':' end [':' ...] [',' next block] -- one pass over the bytes of the block
            const char *c1 = scan_ch(b0, le, ':');
            const char *c2 = c1 ? scan_ch(c1 + 1, le, ':') : nullptr;
            long s = 0, e = 0;
            const char *c3 = c2 ? field_long(c2 + 1, le, s) : nullptr;  // the ':' that ends the start field
            const char *c4 = nullptr;                                   // the ':' after the end field, or le
            if (c3 && c3 < le) c4 = field_long(c3 + 1, le, e);
            if (!c4) return failed(LSQ_ERR_PARSE);
            if (s - 1 < INT32_MIN || s - 1 > INT32_MAX || e < INT32_MIN || e > INT32_MAX) return failed(LSQ_ERR_RANGE);
            const uint32_t cid = chrom_cache.lookup(b0, (size_t)(c1 - b0), R.chroms);
            const uint32_t sid = strand_cache.lookup(c1 + 1, (size_t)(c2 - c1 - 1), R.strands);
            uint32_t cs = cid | (sid << 16);  // chunk-local ids, remapped when chunks are merged
            if (nb == 0) first_cs = cs;
            else if (cs != first_cs) mixed = true;
            R.blk.push_back((int32_t)(s - 1));
            R.blk.push_back((int32_t)e);
            R.blk_cs.push_back(cs);
            ++nb;
            if (c4 == le) break;  // no further field: the read ends with this block
            const char *comma = scan_ch(c4, le, ',');  // :326
            if (!comma) break;
            b0 = comma + 1;
        }
        R.any_mixed |= mixed;
        R.read_cs.push_back(first_cs | (mixed ? LSQ_READ_MIXED : 0u));
        R.line_num.push_back((uint32_t)line_num);
        R.nblk.push_back(nb);
    }
    R.n_lines = line_num;
}

struct MappedFile {  // read-only view of the whole file: mapped when it is a regular file, read otherwise (pipes)
    const char *data = nullptr;
    size_t size = 0;
    bool ok = false;
    void *map = nullptr;
    std::string owned;
    explicit MappedFile(const char *path) {
        int fd = open(path, O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (fstat(fd, &st) != 0) {
            close(fd);
            return;
        }
        ok = true;
        if (S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                map = m;
                size = (size_t)st.st_size;
                data = static_cast<const char *>(m);
                madvise(m, size, MADV_SEQUENTIAL);
                madvise(m, size, MADV_WILLNEED);
            }
        }
        if (!map) {  // a pipe, a character device, or a file system without mmap
            char tmp[1 << 16];
            ssize_t n;
            while ((n = read(fd, tmp, sizeof tmp)) > 0) owned.append(tmp, (size_t)n);
            ok = n == 0;
            data = owned.data();
            size = owned.size();
        }
        close(fd);
    }
    ~MappedFile() {
        if (map) munmap(map, size);
    }
    MappedFile(const MappedFile &) = delete;
    MappedFile &operator=(const MappedFile &) = delete;
};

template <class F>
void run_threads(unsigned n, F fn) {
    std::vector<std::thread> th;
    for (unsigned t = 1; t < n; ++t) th.emplace_back([&fn, t]() { fn(t); });
    fn(0);
    for (auto &x : th) x.join();
}

}  // namespace

int parse_mrf(const char *path, lsqh_reads &R) {
    MappedFile mf(path);
    if (!mf.ok) {
        set_error(std::string("cannot read reads file ") + path);
        return LSQ_ERR_IO;
    }
    if (mf.size == 0) {  // an empty file passes the reference's open assertion and holds no reads (count.cpp:278-283)
        R = lsqh_reads();
        R.blk_off.assign(1, 0u);
        return LSQ_OK;
    }
    const bool timing = getenv("LSQ_PARSE_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    const char *p = mf.data, *end = p + mf.size;
    const char *nl = find_ch(p, end, '\n');
    p = nl ? nl + 1 : end;  // first line skipped unconditionally (:283)

    // split the body at line boundaries; every chunk is parsed by its own thread
    unsigned nthreads = std::thread::hardware_concurrency();
    if (const char *e = getenv("LSQ_THREADS")) nthreads = (unsigned)atoi(e);
    nthreads = std::max(1u, std::min(nthreads, 64u));
    const size_t body = (size_t)(end - p);
    if (body < (1u << 20)) nthreads = 1;
    std::vector<const char *> cut(nthreads + 1, end);
    cut[0] = p;
    for (unsigned t = 1; t < nthreads; ++t) {
        const char *q = p + body / nthreads * t;
        if (q < cut[t - 1]) q = cut[t - 1];
        const char *n2 = find_ch(q, end, '\n');
        cut[t] = n2 ? n2 + 1 : end;
    }
    const double t1 = now();
    std::vector<ChunkOut> out(nthreads);
    run_threads(nthreads, [&](unsigned t) { parse_chunk(cut[t], cut[t + 1], out[t]); });
    const double t2 = now();
    std::vector<uint64_t> first_line(nthreads + 1, 0);  // lines before each chunk
    for (unsigned t = 0; t < nthreads; ++t) {  // the first failing line in file order decides, as in a serial scan
        const ChunkOut &o = out[t];
        if (o.err) {
            const std::string ln = std::to_string(first_line[t] + o.err_line);
            set_error(o.err == LSQ_ERR_PARSE
                          ? "Lexical_cast error when converting arguments to numeric values (MRF line " + ln + ")"
                          : o.err_line ? "MRF coordinate outside int32 at line " + ln : "more than 2^32-1 lines in one MRF file");
            return o.err;
        }
        first_line[t + 1] = first_line[t] + o.n_lines;
        if (first_line[t + 1] > UINT32_MAX && !o.read_cs.empty()) {
            set_error("more than 2^32-1 lines in one MRF file");
            return LSQ_ERR_RANGE;
        }
    }
    // merge in file order: global dictionary ids follow first appearance; the arrays are copied by the threads
    std::vector<size_t> read_base(nthreads + 1, 0), blk_base(nthreads + 1, 0);
    std::vector<std::vector<uint32_t> > cmap(nthreads), smap(nthreads);
    bool any_mixed = false;
    for (unsigned t = 0; t < nthreads; ++t) {
        const ChunkOut &o = out[t];
        read_base[t + 1] = read_base[t] + o.read_cs.size();
        blk_base[t + 1] = blk_base[t] + o.blk_cs.size();
        cmap[t].resize(o.chroms.names.size());
        smap[t].resize(o.strands.names.size());
        for (size_t i = 0; i < cmap[t].size(); ++i) cmap[t][i] = R.chroms.intern(o.chroms.names[i]);
        for (size_t i = 0; i < smap[t].size(); ++i) smap[t][i] = R.strands.intern(o.strands.names[i]);
        any_mixed |= o.any_mixed;
    }
    if (R.chroms.names.size() > 0x10000 || R.strands.names.size() > 0x100) {
        set_error("too many distinct chromosome (65536) or strand (256) strings in MRF file");
        return LSQ_ERR_RANGE;
    }
    const size_t n_reads = read_base[nthreads], n_blocks = blk_base[nthreads];
    if (n_blocks > UINT32_MAX) {
        set_error("more than 2^32-1 alignment blocks in one MRF file");
        return LSQ_ERR_RANGE;
    }
    R.blk_off.resize(n_reads + 1);
    R.read_cs.resize(n_reads);
    R.line_num.resize(n_reads);
    R.blk_cs.resize(n_blocks);
    R.blk.resize(2 * n_blocks);
    R.blk_off[0] = 0;
    run_threads(nthreads, [&](unsigned t) {
        ChunkOut &o = out[t];
        const std::vector<uint32_t> &cm = cmap[t], &sm = smap[t];
        // two reads of different chunks may have used the same local id for different names, so
        // "mixed" was derived within a read, which stays inside one chunk
        auto remap = [&](uint32_t cs) { return cm[cs & 0xFFFFu] | (sm[(cs >> 16) & 0xFFu] << 16) | (cs & LSQ_READ_MIXED); };
        const size_t r0 = read_base[t], b0 = blk_base[t];
        uint32_t off = (uint32_t)b0;
        for (size_t i = 0; i < o.read_cs.size(); ++i) {
            R.read_cs[r0 + i] = remap(o.read_cs[i]);
            R.line_num[r0 + i] = (uint32_t)(first_line[t] + o.line_num[i]);
            off += o.nblk[i];
            R.blk_off[r0 + i + 1] = off;
        }
        for (size_t i = 0; i < o.blk_cs.size(); ++i) R.blk_cs[b0 + i] = remap(o.blk_cs[i]);
        if (!o.blk.empty()) memcpy(R.blk.data() + 2 * b0, o.blk.data(), o.blk.size() * sizeof(int32_t));
        o = ChunkOut();
    });
    R.any_mixed = any_mixed;
    R.chrom_ptrs = R.chroms.c_ptrs();
    R.strand_ptrs = R.strands.c_ptrs();
    if (timing)
        fprintf(stderr, "parse_mrf: %u threads, map %.3f s, parse %.3f s, merge %.3f s\n", nthreads, t1 - t0, t2 - t1,
                now() - t2);
    return LSQ_OK;
}

// Compact wire format: every array in its narrowest lossless form, or left empty when a value
// does not fit (the view then carries only the wide array).
void compact_reads(lsqh_reads &R) {
    const size_t n = R.read_cs.size(), nb = R.blk_cs.size();
    R.blk_cnt.clear();
    R.read_cs8.clear();
    R.blk_start.clear();
    R.blk_len.clear();
    R.blk_len8.clear();
    bool ok = true;
    R.blk_cnt.resize(n);
    for (size_t i = 0; i < n && ok; ++i) {
        const uint32_t c = R.blk_off[i + 1] - R.blk_off[i];
        ok = c < 256u;
        R.blk_cnt[i] = (uint8_t)c;
    }
    if (!ok) R.blk_cnt.clear();
    if (R.chroms.names.size() <= 64 && R.strands.names.size() <= 2) {
        R.read_cs8.resize(n);
        for (size_t i = 0; i < n; ++i) {
            const uint32_t cs = R.read_cs[i];
            R.read_cs8[i] = (uint8_t)((cs & 0x3Fu) | (((cs >> 16) & 1u) << 6) | ((cs & LSQ_READ_MIXED) ? 0x80u : 0u));
        }
    }
    ok = true;
    R.blk_start.resize(nb);
    R.blk_len.resize(nb);
    for (size_t b = 0; b < nb && ok; ++b) {
        const int64_t len = (int64_t)R.blk[2 * b + 1] - (int64_t)R.blk[2 * b];
        ok = len >= 0 && len < 65536;
        R.blk_start[b] = R.blk[2 * b];
        R.blk_len[b] = (uint16_t)len;
    }
    if (!ok) {
        R.blk_start.clear();
        R.blk_len.clear();
        return;
    }
    uint16_t longest = 0;
    for (uint16_t l : R.blk_len) longest = std::max(longest, l);
    if (longest < 256 && nb) {  // short reads: one byte per block length
        R.blk_len8.resize(nb);
        for (size_t b = 0; b < nb; ++b) R.blk_len8[b] = (uint8_t)R.blk_len[b];
    }
}

}  // namespace lsqh
