// Host front end: packed MRF_SINGLE reads (owner of the arrays an lsq_packed_reads views).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "annotation.hpp"

namespace lsqh {
// std::vector whose resize() leaves trivially constructible elements uninitialised: the merge of the parser's
// per-thread chunks and the pack loader overwrite every element anyway, and zero-filling gigabytes first is not free
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U>
    struct rebind {
        typedef NoInitAlloc<U> other;
    };
    template <class U, class... A>
    void construct(U *p, A &&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
        else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
template <class T>
using Array = std::vector<T, NoInitAlloc<T> >;
}  // namespace lsqh

struct lsqh_reads {
    lsqh::Dictionary chroms, strands;
    lsqh::Array<uint32_t> blk_off, read_cs, blk_cs, line_num;
    lsqh::Array<int32_t> blk;
    bool any_mixed = false;
    // compact wire format (lsqh_reads_compact); empty when not derived or when a value does not fit
    std::vector<uint8_t> blk_cnt, read_cs8, blk_len8;
    std::vector<int32_t> blk_start;
    std::vector<uint16_t> blk_len;
    std::vector<const char *> chrom_ptrs, strand_ptrs;
};

namespace lsqh {
int parse_mrf(const char *path, lsqh_reads &out);
void compact_reads(lsqh_reads &r);
// on-disk packed cache (pack.cpp); source_path stamps / validates the text file the pack was made from
int save_reads(const lsqh_reads &r, const char *pack_path, const char *source_path);
int load_reads(const char *pack_path, const char *source_path, lsqh_reads &out);
}
