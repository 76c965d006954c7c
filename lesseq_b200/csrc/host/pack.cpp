// Host front end: on-disk packed cache of a parsed MRF_SINGLE read set.
//
// The reference parses the same text file once in `count` and again in `solve`
// (count/count.cpp:279-339, solve/solve.cpp:429-486); after the kernels, text parsing is what is
// left of the wall clock.  A pack file holds exactly the arrays lsqh_mrf_parse produces (the input
// of the covered-region filter, which depends on the gene range and therefore stays on the
// device), stamped with size and mtime of the text it was made from.  Opt-in: the drop-in programs
// only look for / write `<reads_path>.lsqpack` when LSQ_READ_CACHE=1, because the reference
// writes no files.
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstring>

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
    uint64_t check;  // FNV-1a over the arrays, in file order
};
const char MAGIC[8] = {'L', 'S', 'Q', 'P', 'A', 'C', 'K', '1'};

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

bool source_stamp(const char *path, uint64_t &size, int64_t &mtime_ns) {
    struct stat st;
    if (!path || stat(path, &st) != 0) return false;
    size = (uint64_t)st.st_size;
    mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
    return true;
}

template <class V>
bool put(FILE *f, const V &v, size_t n, uint64_t &h) {
    typedef typename V::value_type T;
    if (v.size() < n) return false;
    h = fnv(h, v.data(), n * sizeof(T));
    return n == 0 || fwrite(v.data(), sizeof(T), n, f) == n;
}
template <class V>
bool get(FILE *f, V &v, size_t n, uint64_t &h) {
    typedef typename V::value_type T;
    v.resize(n);  // lsqh::Array: not zero-filled
    if (n && fread(v.data(), sizeof(T), n, f) != n) return false;
    h = fnv(h, v.data(), n * sizeof(T));
    return true;
}
bool put_names(FILE *f, const std::vector<std::string> &names) {
    for (const std::string &s : names) {
        const uint32_t len = (uint32_t)s.size();
        if (fwrite(&len, 4, 1, f) != 1 || (len && fwrite(s.data(), 1, len, f) != len)) return false;
    }
    return true;
}
bool get_names(FILE *f, uint32_t n, Dictionary &d) {
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t len;
        if (fread(&len, 4, 1, f) != 1 || len > (1u << 20)) return false;
        std::string s(len, '\0');
        if (len && fread(&s[0], 1, len, f) != len) return false;
        if (d.intern(s) != i) return false;  // duplicate name: not a file we wrote
    }
    return true;
}

}  // namespace

int save_reads(const lsqh_reads &R, const char *pack_path, const char *source_path) {
    const size_t n = R.read_cs.size(), nb = R.blk_cs.size();
    PackHeader H;
    memset(&H, 0, sizeof H);
    memcpy(H.magic, MAGIC, 8);
    H.version = 1;
    H.flags = R.any_mixed ? 1u : 0u;
    H.n_reads = n;
    H.n_blocks = nb;
    H.n_chroms = (uint32_t)R.chroms.names.size();
    H.n_strands = (uint32_t)R.strands.names.size();
    if (source_path && !source_stamp(source_path, H.src_size, H.src_mtime_ns)) {
        set_error(std::string("cannot stat ") + source_path);
        return LSQ_ERR_IO;
    }
    const std::string tmp = std::string(pack_path) + ".tmp." + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) {
        set_error("cannot create " + tmp);
        return LSQ_ERR_IO;
    }
    uint64_t h = 0xCBF29CE484222325ull;
    bool ok = fwrite(&H, sizeof H, 1, f) == 1 && put_names(f, R.chroms.names) && put_names(f, R.strands.names) &&
              put(f, R.blk_off, n + 1, h) && put(f, R.read_cs, n, h) && put(f, R.line_num, n, h) && put(f, R.blk, 2 * nb, h) &&
              (!R.any_mixed || put(f, R.blk_cs, nb, h));
    if (ok) {  // the checksum is known last: rewrite the header
        H.check = h;
        ok = fseek(f, 0, SEEK_SET) == 0 && fwrite(&H, sizeof H, 1, f) == 1;
    }
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
    FILE *f = fopen(pack_path, "rb");
    if (!f) {
        set_error(std::string("cannot open ") + pack_path);
        return LSQ_ERR_IO;
    }
    auto bad = [&](const char *why) {
        fclose(f);
        set_error(std::string(pack_path) + ": " + why);
        return LSQ_ERR_IO;
    };
    PackHeader H;
    if (fread(&H, sizeof H, 1, f) != 1 || memcmp(H.magic, MAGIC, 8) != 0 || H.version != 1) return bad("not a read pack");
    if (source_path) {
        uint64_t size;
        int64_t mtime;
        if (!source_stamp(source_path, size, mtime) || size != H.src_size || mtime != H.src_mtime_ns)
            return bad("stale (the text file changed)");
    }
    if (H.n_blocks + 2 * H.n_reads > 0xFFFFFFF0ull || H.n_chroms > 0x10000 || H.n_strands > 0x100) return bad("corrupt header");
    R = lsqh_reads();
    if (!get_names(f, H.n_chroms, R.chroms) || !get_names(f, H.n_strands, R.strands)) return bad("corrupt dictionaries");
    const size_t n = (size_t)H.n_reads, nb = (size_t)H.n_blocks;
    uint64_t h = 0xCBF29CE484222325ull;
    R.any_mixed = (H.flags & 1u) != 0;
    if (!get(f, R.blk_off, n + 1, h) || !get(f, R.read_cs, n, h) || !get(f, R.line_num, n, h) || !get(f, R.blk, 2 * nb, h) ||
        (R.any_mixed && !get(f, R.blk_cs, nb, h)))
        return bad("truncated");
    if (h != H.check || fgetc(f) != EOF) return bad("checksum mismatch");
    if (R.blk_off[0] != 0 || R.blk_off[n] != nb) return bad("corrupt offsets");
    for (size_t i = 0; i < n; ++i)
        if (R.blk_off[i] > R.blk_off[i + 1]) return bad("corrupt offsets");
    if (!R.any_mixed) {  // blk_cs always has one entry per block: the read's ids
        R.blk_cs.resize(nb);
        for (size_t i = 0; i < n; ++i)
            for (uint32_t b = R.blk_off[i]; b < R.blk_off[i + 1]; ++b) R.blk_cs[b] = R.read_cs[i];
    }
    fclose(f);
    R.chrom_ptrs = R.chroms.c_ptrs();
    R.strand_ptrs = R.strands.c_ptrs();
    return LSQ_OK;
}

}  // namespace lsqh
