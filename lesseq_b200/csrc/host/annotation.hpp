// Host front end: annotation model (LH_GENE_TXT forms + UCSC_GENE2ISOFORM map -> event tables).
// Mirrors what count/count.cpp:136-258 builds, as flat arrays ready for lsq_upload_events.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/lesseq_b200.h"

namespace lsqh {

struct FormRow {  // one LH_GENE_TXT line (jsc/bioinfo/gene_anno.hpp:35-51)
    std::string name, chrom, strand;
    uint64_t exon_count = 0;
    std::vector<long> starts, ends;
};

// Sorted disjoint interval set with the reference's add/contains semantics
// (jsc/util/interval_list.hpp:396-422, 462-503).  Abutting intervals fuse only when the later
// one starts where an earlier one ends.
struct IntervalSet {
    std::vector<long> starts, ends;
    void add(long s, long e);
    bool contains(long s, long e) const;
};

struct Dictionary {
    std::vector<std::string> names;
    std::unordered_map<std::string, uint32_t> index;
    uint32_t intern(const std::string &s);
    std::vector<const char *> c_ptrs() const;
};

}  // namespace lsqh

struct lsqh_annotation {
    lsqh::Dictionary chroms, strands;
    uint64_t n_selected = 0;  // genes in the CLI range [begin,end)
    // shard events, output order
    std::vector<std::string> ev_name;
    std::vector<uint32_t> ev_chrom, ev_strand, ev_seg_off, ev_form_off, ev_name_off;
    std::vector<int32_t> ev_gs, ev_ge;
    std::string ev_names_pool;
    std::vector<int32_t> seg_start, seg_end;
    std::vector<std::string> form_name;
    std::vector<uint64_t> form_len;
    std::vector<uint32_t> form_seg_off, form_seg_idx;
    std::vector<uint32_t> cov_off;
    std::vector<int32_t> cov_start, cov_end;
    std::vector<const char *> chrom_ptrs, strand_ptrs;
};

namespace lsqh {
// Returns LSQ_OK or a negative LSQ_ERR_* (message via set_error()).
int load_annotation(const char *forms_path, const char *map_path, uint64_t begin, uint64_t end,
                    uint64_t shard_begin, uint64_t shard_end, lsqh_annotation &out);
// the same with the reference's format tokens (solve.cpp:158-296, 310-329); LSQ_ERR_FORMAT for an unknown token
int load_annotation_fmt(const char *forms_format, const char *forms_path, const char *map_format, const char *map_path,
                        uint64_t begin, uint64_t end, uint64_t shard_begin, uint64_t shard_end, lsqh_annotation &A);
void set_error(const std::string &msg);  // shared thread-local error slot (abi.cu)
}  // namespace lsqh
