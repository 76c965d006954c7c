// Device-side table layouts shared by the kernels and the host orchestration (lsq_abi.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifdef __CUDACC__
#define LSQ_TYPES_HD __host__ __device__ __forceinline__
#else
#define LSQ_TYPES_HD inline
#endif

namespace lsq {

constexpr int DENSE_MAX_FORMS = 8;  // events with K <= 8 use a direct-indexed class table
constexpr int LSQ_MAX_FORMS_PER_EVENT_DEV = 64;

// Event header, 48 bytes, events sorted by (chromosome, gs, ge).  hdr[E] is a sentinel whose
// seg_off / form_off close the last event's ranges.
struct __align__(16) EvHdr {
    int32_t gs, ge;        // candidate reads have gs <= start <= ge (count.cpp:431-432)
    uint32_t seg_off;      // into seg[]
    uint32_t form_off;     // into fmask[] / G[] (internal form order)
    uint16_t S, K;         // segments, forms
    uint8_t strand;        // annotation strand id of the first form
    uint8_t flags;         // EV_* below
    uint16_t tie_len;      // length of the name suffix for tie mode 2
    uint32_t slot_off;     // first class slot of this event (per read set: + m * total_slots)
    uint32_t slot_cap;     // dense: 2^K-1 slots indexed by mask-1; hash: power of two
    uint32_t tie_off;      // into tie_pool
    uint32_t pfx_off;      // into pfx[] when EV_WIDE (K rows of S+1 prefix counts)
    uint32_t out_ev;       // index of the event in the caller's (output) order
    uint32_t out_form;     // index of the event's first form in the caller's order
};
static_assert(sizeof(EvHdr) == 48, "EvHdr layout");

constexpr uint8_t EV_TIE_MASK = 0x3;   // 0: "read-N" never < name, 1: always, 2: compare digits with suffix
constexpr uint8_t EV_WIDE = 0x4;       // S > 64: prefix-count compatibility test
constexpr uint8_t EV_HASH = 0x8;       // class table is an open-addressed hash (K > DENSE_MAX_FORMS)

// Stab entry, 8 bytes: one (piece, event) pair of the stab arrangement.  Pieces are the elementary
// intervals cut by every segment boundary and span end of the events of a chromosome (and by the
// covered-region boundaries), so for a read whose first block starts in the piece, the segment j0 of
// event x that contains the start is known in advance.  A piece lists only the events for which it
// lies inside a segment: a first block that starts outside every segment of an event can never be
// matched by the walk (common/read.h:226-262).
struct __align__(8) StabEntry {
    uint32_t x;      // internal event id
    uint32_t meta;   // bits 0-15 j0 | SE_AT_GS
};
static_assert(sizeof(StabEntry) == 8, "StabEntry layout");
constexpr uint32_t SE_AT_GS = 1u << 25;  // the piece starts at the event start (tie-break possible)

// Per-piece record read by the scan kernel: ONE 16-byte load answers "is this block start inside a covered
// region, and up to where", "is a read starting here a candidate of any event" and -- through the cut list
// -- "which outcome class does a block that starts here and ends at e fall into" (see lsq_scan_kernel).
// Pieces are cut by the covered-region boundaries too, so a piece is covered entirely or not.
//
// Cut list of a piece: for every stab entry (event x, segment j0 holding the piece) the ends of the
// segments of the chain of abutting segments that starts at j0 (common/read.h:226-262: a block can only
// run from one segment into the next when they abut).  All entries' ends, distinct and ascending, are the
// piece's cuts q_0 < q_1 < ...: the result of the walk for a block [s, e) with s inside the piece depends on
// e only through the number of cuts below e (and on whether e hits a cut exactly when a further block
// follows).  Signature slot sig_off + k counts the single-block reads with q_{k-1} < e <= q_k.
struct __align__(16) PieceInfo {
    int32_t pos;        // first position of the piece
    int32_t cov_end;    // end of the covered region containing the piece, INT32_MIN if uncovered
    uint32_t sigw;      // bits 0-25 first signature slot | bits 26-29 number of cuts (PIECE_NO_SIG: none usable)
                        // | bit 30 PIECE_CAND: the piece has stab entries | bit 31 PIECE_TIE: an entry's event
                        // starts here, so the lower_bound tie-break (count.cpp:429-432) can apply
    uint32_t cutw;      // cuts 0 and 1 as 16-bit offsets from pos (further cuts: Tables::sig_cut)
};
constexpr uint32_t PIECE_SIG_BITS = 26, PIECE_SIG_MASK = (1u << PIECE_SIG_BITS) - 1u;
constexpr uint32_t PIECE_NO_SIG = 15u;   // cut count that disables the signature paths (too many / too far)
constexpr uint32_t PIECE_MAX_CUTS = 14u;
constexpr uint32_t PIECE_CAND = 1u << 30, PIECE_TIE = 1u << 31;
static_assert(sizeof(PieceInfo) == 16, "PieceInfo layout");
LSQ_TYPES_HD uint32_t piece_ncuts(uint32_t sigw) { return (sigw >> PIECE_SIG_BITS) & 15u; }

// Signature of a two-block read (lsq_scan_kernel): piece of block 0 | cut hit by its end | piece of
// block 1 (which starts exactly on the piece start) | cut category of its end.  Never 0: piece 1 > piece 0.
LSQ_TYPES_HD uint64_t pair_key(uint32_t p0, uint32_t k0, uint32_t p1, uint32_t k1) {
    return (uint64_t)p0 | ((uint64_t)k0 << 26) | ((uint64_t)p1 << 30) | ((uint64_t)k1 << 56);
}
constexpr uint32_t HASH_MAX_PROBES = 32;
// Signature of a three-block read: the hash slot that interns its (piece 0, cut 0, piece 1, cut 1) prefix -- a pair
// key, present with count 0 when no two-block read shares it -- then piece 2 and the cut category of its end.
// Bit 63 tells the two kinds of key apart (a pair key uses bits 0-59).
constexpr uint64_t TRIPLE_KEY_BIT = 1ull << 63;
LSQ_TYPES_HD uint64_t triple_key(uint32_t prefix_slot, uint32_t p2, uint32_t k2) {
    return TRIPLE_KEY_BIT | (uint64_t)prefix_slot | ((uint64_t)p2 << 30) | ((uint64_t)k2 << 56);
}

struct Tables {
    uint32_t E, F, n_pieces, n_chroms, n_ann_strands, n_sig;
    uint64_t total_slots;
    const EvHdr *hdr;
    const int2 *seg;
    const uint64_t *fmask;
    const uint32_t *pfx;
    const char *tie_pool;
    const uint32_t *bp_chrom_off;  // [n_chroms+1] piece range per chromosome
    const PieceInfo *pinfo;        // [n_pieces] sorted by position within a chromosome
    const uint32_t *stab_off;      // [n_pieces+1]
    const StabEntry *stab;         // candidate events of each piece
    const int32_t *sig_cut;        // [n_sig] cut position of every signature slot
    const uint32_t *sig_piece;     // [n_sig] piece of every signature slot
    const uint32_t *cov_off;       // [n_chroms+1]
    const int32_t *cov_start, *cov_end;
    // bucket index: position >> shift -> rank of the bucket start in the sorted piece starts plus the
    // in-bucket starts, so the piece of a position comes from ONE 8-byte load (bucket_word)
    uint32_t bp_shift;
    const uint4 *chrom_meta;       // [n_chroms] {first piece, end piece, first bucket, buckets - 1}
    const uint32_t *bp_bkt_off;    // [n_chroms+1] first bucket of each chromosome
    const uint2 *bp_bkt;           // [buckets + n_chroms] x: rank of the first piece starting in or after the bucket,
                                   // y: the piece starts inside the bucket (bucket_word)
    const uint2 *bp_sub;           // sub-buckets of the crowded buckets (same entry format, bucket_sub_shift)
};

// Bucket word: bits 29-31 = number of piece starts inside the bucket, then bucket_capacity(shift)
// offsets of `shift` bits each (start - bucket start), ascending.  A bucket with more starts than fit
// (BKT_MANY) keeps in bits 0-28 the index of its 2^(shift - sub_shift) sub-buckets in bp_sub[], which
// have the same entry format over 2^sub_shift positions each; only a crowded SUB-bucket needs a search.
constexpr uint32_t BKT_MANY = 7u;
LSQ_TYPES_HD uint32_t bucket_capacity(uint32_t shift) {
    return shift == 0u ? 4u : (29u / shift < 4u ? 29u / shift : 4u);
}
LSQ_TYPES_HD uint32_t bucket_sub_shift(uint32_t shift) { return shift > 3u ? shift - 3u : 0u; }
LSQ_TYPES_HD bool bucket_many(uint32_t w) { return (w >> 29) == BKT_MANY; }
// Number of piece starts at or below in-bucket offset xo (the word is not BKT_MANY).  Branch-free:
// every lane of a warp takes the same four compares.
LSQ_TYPES_HD uint32_t bucket_rank(uint32_t shift, uint32_t w, uint32_t xo) {
    const uint32_t cnt = w >> 29, mask = (1u << shift) - 1u;
    const uint32_t s2 = 2u * shift < 31u ? 2u * shift : 31u, s3 = 3u * shift < 31u ? 3u * shift : 31u;
    // offsets ascend, so the tests are nested: bitwise ands keep the four compares free of branches
    uint32_t n = (uint32_t)(cnt > 0u) & (uint32_t)(xo >= (w & mask));
    n += (uint32_t)(cnt > 1u) & (uint32_t)(xo >= ((w >> shift) & mask));
    n += (uint32_t)(cnt > 2u) & (uint32_t)(xo >= ((w >> s2) & mask));
    n += (uint32_t)(cnt > 3u) & (uint32_t)(xo >= ((w >> s3) & mask));
    return n;
}

// Reads are processed in groups of SCAN_GROUP consecutive reads (one warp, four reads per lane).
constexpr uint32_t SCAN_GROUP = 128;
constexpr uint32_t SCAN_SUPER = 256;  // groups per super-group of the block-offset index (grp_base / sup_base)

// Counters of the scan (statistics; per read set)
constexpr int CTR_SINGLE_SIG = 0, CTR_PAIR_SIG = 1, CTR_GENERAL = 2, CTR_NONE = 3, CTR_HASH_FULL = 4, CTR_TRIPLE_SIG = 5, CTR_N = 8;

// Everything the count path reads and writes for one read set.  Inputs come in two forms:
//   wide     blk_off u32[n+1] (CSR), blk int2[B], read_cs u32[n]
//   compact  blk_cnt u8[n] (+ the block-offset index grp_base/sup_base built by lsq_group_base_kernels),
//            read_cs8 u8[n], blk_start i32[B] + blk_len u16[B] or blk_len8 u8[B]
// and any mix of them (load_group / load_block / load_read_cs pick what is present).
struct ScanParams {
    Tables T;
    uint64_t n_reads;
    const uint32_t *blk_off;
    const uint8_t *blk_cnt;       // with grp_base/sup_base: instead of blk_off (padded to a multiple of 4 reads)
    const uint32_t *grp_base;     // [groups] blocks before the group within its super-group
    const uint32_t *sup_base;     // [super-groups] blocks before the super-group
    const int2 *blk;
    const uint32_t *read_cs;
    const uint32_t *blk_cs;       // may be null
    const uint32_t *chrom_remap;  // file chromosome id -> annotation id or 0xFFFFFFFF
    const uint4 *file_meta;       // file chromosome id -> chrom_meta of its annotation id (zeros: none)
    const uint8_t *read_cs8;      // padded to a multiple of 4 reads
    const int32_t *blk_start;
    const uint16_t *blk_len;
    const uint8_t *blk_len8;      // instead of blk_len when every block is shorter than 256
    const uint32_t *progress;     // null, or: blocks of blk[] that have landed so far (streamed host->device copy)
    const uint32_t *line_num;     // null: read i is "read-<i+1>"
    const uint8_t *strand_tbl;    // [file strands x annotation strands] bit0: read < event, bit1: equal
    int2 *ext;                    // [n_blocks + 2 n_reads] scratch: interval lists of reads with > 2 kept blocks
    // signature counters (per launch scratch, cleared before the scan)
    uint32_t *sig_cnt;                // [n_sig] reads per (piece, cut category)
    unsigned long long *sig_dev;      // [n_sig] sum of (read length - len0) over those reads
    ulonglong2 *hash;                 // [hash_mask + 1] {pair_key, reads}; key 0 = empty
    unsigned long long *hash_dev;     // [hash_mask + 1] sum of (read length - len0)
    uint32_t hash_mask;
    uint32_t len0;                    // the read set's nominal read length (the CLI's expected_read_length)
    uint2 *gen_list;                  // [gen_cap] {read, first block} of the reads left to lsq_general_kernel
    uint32_t *gen_n;
    uint32_t gen_cap;
    uint32_t flags;                   // SCAN_* below
    unsigned long long *counters;     // [CTR_N]
    // results
    unsigned long long *slot_cnt;   // this read set's class counters
    unsigned long long *slot_mask;  // hash keys (EV_HASH events)
    unsigned long long *bases;      // [E] internal order
};
constexpr uint32_t SCAN_NO_SINGLE_SIG = 1u, SCAN_NO_PAIR_SIG = 2u, SCAN_NO_TRIPLE_SIG = 4u;  // debugging: send those reads to lsq_general_kernel

// Read bucketing per gene-range shard (count.cpp:202-215: every process of the reference reads the whole file and
// keeps what its range covers).  The router runs on the tables of the WHOLE selection; owner[p] has bit d set when
// rank d owns an event with a stab entry at piece p.  A read is sent to every rank owning a candidate event of ANY
// of its blocks' pieces -- a superset of the ranks whose events it can count for (its first KEPT block decides
// that), so each rank's tables, built with the covered regions of the whole selection, give the same tallies as
// the single-GPU run.  Outboxes keep the read set's own layout (compact or wide).
constexpr uint32_t ROUTE_MAX_RANKS = 16;  // = LSQ_MAX_RANKS of the public header
struct RouteBox {
    uint8_t *blk_cnt;      // [cap_reads]
    uint8_t *read_cs8;     // compact ids ...
    uint32_t *read_cs;     // ... or wide ids
    uint32_t *line_num;    // [cap_reads] the read's line number in the whole file (its name, count.cpp:293-295)
    int32_t *blk_start;    // compact blocks: start + length (u8 or u16) ...
    uint8_t *blk_len8;
    uint16_t *blk_len;
    int2 *blk;             // ... or wide blocks
    uint32_t *blk_cs;      // per-block ids (only when the read set has chimeric reads)
};
struct RouteParams {
    uint32_t n_ranks;
    uint32_t line_base;               // line number of local read 0 minus 1 when the set carries no line numbers
    const unsigned long long *owner;  // [n_pieces]
    unsigned long long *cursor;       // [n_ranks] appended per destination: reads << 32 | blocks
    RouteBox box[ROUTE_MAX_RANKS];
};

struct SolveParams {
    Tables T;
    uint32_t M;
    const unsigned long long *slot_cnt;   // [M * total_slots]
    const unsigned long long *slot_mask;  // [M * total_slots]
    const unsigned long long *bases;      // [M * E] internal order
    const double *G;                      // [M * F] internal form order, 1/ARS or 0
    const uint64_t *form_len;             // [F] internal form order
    // injected class tables (output-order CSR), per read set; null when counted on the device
    const uint64_t *inj_off[16];
    const uint64_t *inj_mask[16];
    const uint64_t *inj_cnt[16];
    double total_read_mbases;
    // outputs, caller (output) order
    unsigned long long *n_valid;    // [M * E]
    unsigned long long *bases_out;  // [M * E]
    unsigned long long *iso_count;  // [F]
    double *theta, *rpkm, *se;      // [F]
    double *loglik, *ll_per_read;   // [E]
    uint32_t *iters;                // [E]
    uint32_t max_iters;
    int with_se;
    int small_enabled;             // events with K <= 2 are solved by lsq_solve_small_kernel
    int mid_enabled;               // events with K <= 16 are solved by lsq_solve_mid_kernel (the warp kernel adds their se)
    const uint32_t *mid_order;     // those events sorted by (classes, forms)
    const uint32_t *mid_n;
    unsigned int *work_counter;    // dynamic event fetch of lsq_solve_small_kernel
    unsigned long long *pairs_valid;  // finalize: total valid (read, event) pairs (statistics)
    // events whose stop decision is fragile (em_stop_is_fragile): replayed read by read by lsq_exact_kernel
    uint32_t *flag_n;     // null: no flagging
    uint32_t *flag_list;  // [E] internal event ids
    uint32_t *ev_flag;    // [E] index + 1 in flag_list, 0: not flagged
};

}  // namespace lsq
