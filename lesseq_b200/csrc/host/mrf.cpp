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

}  // namespace

namespace {

struct ChunkOut {  // per-thread result, merged in file order
    Dictionary chroms, strands;
    std::vector<uint32_t> nblk, read_cs, blk_cs, line_num;
    std::vector<int32_t> blk;
    int err = LSQ_OK;
    uint64_t err_line = 0;
    std::string err_msg;
};

// Parses the '\n'-terminated lines of [p, end); `line_num` is the number of the line before p.
void parse_chunk(const char *p, const char *end, uint64_t line_num, ChunkOut &R) {
    std::string last_chr, last_strand;
    uint32_t last_chr_id = 0, last_strand_id = 0;
    bool have_chr = false, have_strand = false;
    auto failed = [&](int code, const std::string &msg) {
        R.err = code;
        R.err_line = line_num;
        R.err_msg = msg;
    };
    while (p < end) {
        const char *le = find_ch(p, end, '\n');
        if (!le) break;  // unterminated last line is never processed (:285)
        ++line_num;
        const char *ls = p;
        p = le + 1;
        if (ls < le && *ls == '#') continue;
        if ((size_t)(le - ls) == 15 && memcmp(ls, "AlignmentBlocks", 15) == 0) continue;
        if (line_num > UINT32_MAX) return failed(LSQ_ERR_RANGE, "more than 2^32-1 lines in one MRF file");
        uint32_t first_cs = 0;
        bool mixed = false;
        uint32_t nb = 0;
        const char *b0 = ls;
        while (true) {
            const char *c1 = find_ch(b0, le, ':');
            const char *c2 = c1 ? find_ch(c1 + 1, le, ':') : nullptr;
            const char *c3 = c2 ? find_ch(c2 + 1, le, ':') : nullptr;
            if (!c3)
                return failed(LSQ_ERR_PARSE, "Lexical_cast error when converting arguments to numeric values (MRF line " +
                                                 std::to_string(line_num) + ")");
            const char *c4 = find_ch(c3 + 1, le, ':');
            long s, e;
            if (!strict_long(c2 + 1, c3, s) || !strict_long(c3 + 1, c4 ? c4 : le, e))
                return failed(LSQ_ERR_PARSE, "Lexical_cast error when converting arguments to numeric values (MRF line " +
                                                 std::to_string(line_num) + ")");
            if (s - 1 < INT32_MIN || s - 1 > INT32_MAX || e < INT32_MIN || e > INT32_MAX)
                return failed(LSQ_ERR_RANGE, "MRF coordinate outside int32 at line " + std::to_string(line_num));
            uint32_t cid, sid;
            if (have_chr && last_chr.size() == (size_t)(c1 - b0) && memcmp(last_chr.data(), b0, c1 - b0) == 0) {
                cid = last_chr_id;
            } else {
                last_chr.assign(b0, c1 - b0);
                cid = last_chr_id = R.chroms.intern(last_chr);
                have_chr = true;
            }
            if (have_strand && last_strand.size() == (size_t)(c2 - c1 - 1) &&
                memcmp(last_strand.data(), c1 + 1, c2 - c1 - 1) == 0) {
                sid = last_strand_id;
            } else {
                last_strand.assign(c1 + 1, c2 - c1 - 1);
                sid = last_strand_id = R.strands.intern(last_strand);
                have_strand = true;
            }
            uint32_t cs = cid | (sid << 16);  // chunk-local ids, remapped when chunks are merged
            if (nb == 0) first_cs = cs;
            else if (cs != first_cs) mixed = true;
            R.blk.push_back((int32_t)(s - 1));
            R.blk.push_back((int32_t)e);
            R.blk_cs.push_back(cs);
            ++nb;
            if (!c4) break;
            const char *comma = find_ch(c4, le, ',');  // :326
            if (!comma) break;
            b0 = comma + 1;
        }
        R.read_cs.push_back(first_cs | (mixed ? LSQ_READ_MIXED : 0u));
        R.line_num.push_back((uint32_t)line_num);
        R.nblk.push_back(nb);
    }
}

}  // namespace

int parse_mrf(const char *path, lsqh_reads &R) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_error(std::string("cannot read reads file ") + path);
        return LSQ_ERR_IO;
    }
    std::string buf;
    {
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (sz > 0) {
            buf.resize((size_t)sz);
            size_t got = fread(&buf[0], 1, (size_t)sz, f);
            buf.resize(got);
        }
        fclose(f);
    }
    if (buf.empty()) {
        set_error(std::string("empty reads file ") + path);
        return LSQ_ERR_IO;
    }
    const char *p = buf.data(), *end = p + buf.size();
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
    std::vector<uint64_t> first_line(nthreads + 1, 0);  // lines before each chunk
    {
        std::vector<uint64_t> lines(nthreads, 0);
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() {
                uint64_t n = 0;
                for (const char *q = cut[t]; q < cut[t + 1];) {
                    const char *x = (const char *)memchr(q, '\n', cut[t + 1] - q);
                    if (!x) break;
                    ++n;
                    q = x + 1;
                }
                lines[t] = n;
            });
        for (auto &x : th) x.join();
        for (unsigned t = 0; t < nthreads; ++t) first_line[t + 1] = first_line[t] + lines[t];
    }
    std::vector<ChunkOut> out(nthreads);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; ++t)
            th.emplace_back([&, t]() { parse_chunk(cut[t], cut[t + 1], first_line[t], out[t]); });
        for (auto &x : th) x.join();
    }
    for (unsigned t = 0; t < nthreads; ++t)  // the first failing line in file order decides, as in a serial scan
        if (out[t].err) {
            set_error(out[t].err_msg);
            return out[t].err;
        }
    size_t n_reads = 0, n_blocks = 0;
    for (auto &o : out) {
        n_reads += o.read_cs.size();
        n_blocks += o.blk_cs.size();
    }
    if (n_blocks > UINT32_MAX) {
        set_error("more than 2^32-1 alignment blocks in one MRF file");
        return LSQ_ERR_RANGE;
    }
    R.blk_off.reserve(n_reads + 1);
    R.read_cs.reserve(n_reads);
    R.line_num.reserve(n_reads);
    R.blk_cs.reserve(n_blocks);
    R.blk.reserve(2 * n_blocks);
    R.blk_off.push_back(0);
    bool any_mixed = false;
    for (auto &o : out) {  // merge in file order: global dictionary ids follow first appearance
        std::vector<uint32_t> cmap(o.chroms.names.size()), smap(o.strands.names.size());
        for (size_t i = 0; i < cmap.size(); ++i) cmap[i] = R.chroms.intern(o.chroms.names[i]);
        for (size_t i = 0; i < smap.size(); ++i) smap[i] = R.strands.intern(o.strands.names[i]);
        if (R.chroms.names.size() > 0x10000 || R.strands.names.size() > 0x100) {
            set_error("too many distinct chromosome (65536) or strand (256) strings in MRF file");
            return LSQ_ERR_RANGE;
        }
        auto remap = [&](uint32_t cs) {
            return cmap[cs & 0xFFFFu] | (smap[(cs >> 16) & 0xFFu] << 16) | (cs & LSQ_READ_MIXED);
        };
        // two reads of different chunks may have used the same local id for different names, so
        // "mixed" has to be re-derived only within a read, which stays inside one chunk
        for (size_t i = 0; i < o.read_cs.size(); ++i) {
            R.read_cs.push_back(remap(o.read_cs[i]));
            any_mixed |= (o.read_cs[i] & LSQ_READ_MIXED) != 0;
            R.line_num.push_back(o.line_num[i]);
            R.blk_off.push_back(R.blk_off.back() + o.nblk[i]);
        }
        for (uint32_t cs : o.blk_cs) R.blk_cs.push_back(remap(cs));
        R.blk.insert(R.blk.end(), o.blk.begin(), o.blk.end());
        o = ChunkOut();
    }
    R.any_mixed = any_mixed;
    R.chrom_ptrs = R.chroms.c_ptrs();
    R.strand_ptrs = R.strands.c_ptrs();
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
