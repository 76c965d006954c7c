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
#include <cstring>

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

int parse_mrf(const char *path, lsqh_reads &R) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_error(std::string("cannot read reads file ") + path);
        return LSQ_ERR_IO;
    }
    std::string buf;
    {
        char tmp[1 << 16];
        size_t n;
        while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
        fclose(f);
    }
    if (buf.empty()) {
        set_error(std::string("empty reads file ") + path);
        return LSQ_ERR_IO;
    }
    const char *p = buf.data(), *end = p + buf.size();
    const char *nl = find_ch(p, end, '\n');
    p = nl ? nl + 1 : end;  // first line skipped unconditionally (:283)
    uint64_t line_num = 0;
    R.blk_off.push_back(0);
    bool any_mixed = false;
    std::string key;
    // small caches: consecutive blocks nearly always repeat chromosome and strand
    std::string last_chr, last_strand;
    uint32_t last_chr_id = 0, last_strand_id = 0;
    bool have_chr = false, have_strand = false;
    while (p < end) {
        const char *le = find_ch(p, end, '\n');
        if (!le) break;  // unterminated last line is never processed (:285)
        ++line_num;
        const char *ls = p;
        p = le + 1;
        if (ls < le && *ls == '#') continue;
        if ((size_t)(le - ls) == 15 && memcmp(ls, "AlignmentBlocks", 15) == 0) continue;
        if (line_num > UINT32_MAX) {
            set_error("more than 2^32-1 lines in one MRF file");
            return LSQ_ERR_RANGE;
        }
        uint32_t first_cs = 0;
        bool mixed = false;
        size_t nb = 0;
        const char *b0 = ls;
        while (true) {
            const char *c1 = find_ch(b0, le, ':');
            const char *c2 = c1 ? find_ch(c1 + 1, le, ':') : nullptr;
            const char *c3 = c2 ? find_ch(c2 + 1, le, ':') : nullptr;
            if (!c3) {
                set_error("Lexical_cast error when converting arguments to numeric values (MRF line " +
                          std::to_string(line_num) + ")");
                return LSQ_ERR_PARSE;
            }
            const char *c4 = find_ch(c3 + 1, le, ':');
            long s, e;
            if (!strict_long(c2 + 1, c3, s) || !strict_long(c3 + 1, c4 ? c4 : le, e)) {
                set_error("Lexical_cast error when converting arguments to numeric values (MRF line " +
                          std::to_string(line_num) + ")");
                return LSQ_ERR_PARSE;
            }
            if (s - 1 < INT32_MIN || s - 1 > INT32_MAX || e < INT32_MIN || e > INT32_MAX) {
                set_error("MRF coordinate outside int32 at line " + std::to_string(line_num));
                return LSQ_ERR_RANGE;
            }
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
            if (cid > 0xFFFF || sid > 0xFF) {
                set_error("too many distinct chromosome (65536) or strand (256) strings in MRF file");
                return LSQ_ERR_RANGE;
            }
            uint32_t cs = cid | (sid << 16);
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
        if (R.blk_cs.size() > UINT32_MAX) {
            set_error("more than 2^32-1 alignment blocks in one MRF file");
            return LSQ_ERR_RANGE;
        }
        R.read_cs.push_back(first_cs | (mixed ? LSQ_READ_MIXED : 0u));
        R.line_num.push_back((uint32_t)line_num);
        R.blk_off.push_back((uint32_t)R.blk_cs.size());
        any_mixed |= mixed;
    }
    R.any_mixed = any_mixed;
    R.chrom_ptrs = R.chroms.c_ptrs();
    R.strand_ptrs = R.strands.c_ptrs();
    return LSQ_OK;
}

}  // namespace lsqh
