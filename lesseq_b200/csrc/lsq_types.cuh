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

// Stab entry, 32 bytes: one (piece, event) pair of the stab arrangement.  Pieces are the
// elementary intervals cut by every segment boundary and span end of the events of a chromosome,
// so for a read whose first block starts in the piece, the segment j0 of event x that contains
// the start is known in advance, and so is the chain of abutting segments a single block can run
// through (common/read.h:226-262).  FAST entries carry the chain ends c[i] and the form
// compatibility mask cmask[i] of a read covering chain segments 0..i, which decides a
// single-block read with a few compares; everything else takes the general walk.
struct __align__(16) StabEntry {
    uint32_t x;         // internal event id
    uint32_t slot_off;  // class slots of x (copy of EvHdr::slot_off)
    int32_t ge;         // event end, first test of the lower_bound tie-break
    uint32_t meta;      // bits 0-15 j0 | bits 16-17 chain length (1..3) | SE_* flags
    int32_t c[3];       // chain segment ends, ascending
    uint8_t cmask[3];   // form mask (K <= 8) when the block ends within chain segment i
    uint8_t K;          // forms of x
};
static_assert(sizeof(StabEntry) == 32, "StabEntry layout");
constexpr uint32_t SE_FAST = 1u << 24;   // table-driven single-block evaluation is valid
constexpr uint32_t SE_AT_GS = 1u << 25;  // the piece starts at the event start (tie-break possible)

// Per-piece record read by the load filter: one 16-byte load answers "is this block start inside
// a covered region, and up to where" as well as "which stab list does a read starting here get".
// Pieces are cut by the covered-region boundaries too, so a piece is covered entirely or not.
struct __align__(16) PieceInfo {
    int32_t pos;        // first position of the piece
    int32_t cov_end;    // end of the covered region containing the piece, INT32_MIN if uncovered
    uint32_t stab_off;  // first stab entry
    uint32_t stab_n;    // number of stab entries (0: reads starting here are no candidates);
                        // bit 31 (PIECE_ALL_FAST): every entry is table-driven (SE_FAST)
};
constexpr uint32_t PIECE_ALL_FAST = 0x80000000u;
static_assert(sizeof(PieceInfo) == 16, "PieceInfo layout");

// Read record (8 bytes).  The load filter drops every read that is a candidate of no event and appends the others to the
// bin (read class, stab bucket) they belong to -- see PartState.  Read classes:
//   FAST    one kept block, shorter than 64 Ki, not starting on the first position of its piece,
//           every stab entry of the piece table-driven:
//           x = start of the block, y = length (bits 0-15) | file strand id (bits 16-23)
//   PAIR    two kept blocks that fit the record (pack_pair)
//   GENERAL everything else: x = offset of the read's entry in ext[]:
//           ext[x] = {original read index, (strand << 24) | n}, ext[x+1 .. x+n] = kept blocks
typedef uint2 ReadRec;

// Partition of the kept reads by (class, stab bucket).  A stab bucket is a run of consecutive
// pieces whose stab entries fall into one window of 2^bkt_shift entries, so the tables a bucket
// needs (its stab entries, their offsets and four counters per entry) fit in shared memory.
// Bins grow in chunks of PART_CHUNK records taken from one pool: state[bin] packs the bin's
// current chunk and its fill, so one 64-bit atomicAdd reserves room for a whole tile's records;
// the reservation that takes the last slot of a chunk installs the bin's spare chunk (and then
// allocates a new spare, off the critical path).  Each installed chunk is one work item of the
// count kernel of its class (wl lists chunk ids per class).
constexpr uint32_t PART_CHUNK_LOG2 = 11;
constexpr uint32_t PART_CHUNK = 1u << PART_CHUNK_LOG2;
constexpr uint32_t PART_MAX_BINS = 12288;  // shared-memory budget of the filter kernel (10 bytes per bin)
constexpr uint32_t BIN_NONE = 0xFFFFu;     // read dropped

struct PartState {
    uint32_t n_bins;              // 3 * n_bkt: bin = class * n_bkt + (stab_off >> bkt_shift)
    uint32_t pool_cap;            // chunks in the pool
    unsigned long long *state;    // [n_bins] (current chunk << 32) | fill
    uint32_t *next_chunk;         // [n_bins] spare chunk, installed when the current one is full
    uint32_t *pool_next;          // chunks allocated so far
    uint32_t *owner;              // [pool_cap] bin of chunk j
    uint32_t *wl;                 // [3 * pool_cap] chunk ids per class
    uint32_t *wl_n;               // [3] list lengths
    uint4 *pool;                  // [pool_cap * PART_CHUNK] PartRec
};

// Partitioned record, 16 bytes (half a sector; one 128-bit access):
//   x  first stab entry of the read's piece (bits 0-25) | number of entries (bits 26-31);
//      FAST and PAIR reads whose piece does not fit are demoted to GENERAL
//   y  PAIR: first position of the piece; GENERAL: the piece (its stab list is looked up)
//   z, w  the ReadRec
constexpr uint32_t PREC_SO_BITS = 26;
constexpr uint32_t PREC_MAX_SO = (1u << PREC_SO_BITS) - 1u, PREC_MAX_N = 63u;

struct Tables {
    uint32_t E, F, n_pieces, n_chroms, n_ann_strands;
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
    const uint32_t *cov_off;       // [n_chroms+1]
    const int32_t *cov_start, *cov_end;
    // bucket indexes: position >> shift -> rank of the bucket start in the sorted array, so a
    // lookup is one table load plus a short forward scan instead of a full binary search
    uint32_t bp_shift;
    const uint4 *chrom_meta;       // [n_chroms] {first piece, end piece, first bucket, buckets - 1}
    const uint32_t *bp_bkt_off;    // [n_chroms+1] first bucket of each chromosome
    const uint2 *bp_bkt;           // [buckets + n_chroms] x: rank of the first piece starting in or after the bucket,
                                   // y: the piece starts inside the bucket (bucket_word), so the piece of a
                                   // position comes from ONE 8-byte load
    const uint2 *bp_sub;           // sub-buckets of the crowded buckets (same entry format, bucket_sub_shift)
    // stab buckets (PartState): bucket of piece p = stab_off[p] >> bkt_shift
    uint32_t bkt_shift, n_bkt;
    const uint32_t *bkt_piece;     // [n_bkt+1] first piece whose stab_off is >= b << bkt_shift
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

struct FilterParams {
    Tables T;
    uint64_t n_reads;
    const uint32_t *blk_off;
    const int2 *blk;
    const uint32_t *read_cs;
    const uint32_t *blk_cs;       // may be null
    const uint32_t *chrom_remap;  // file chromosome id -> annotation id or 0xFFFFFFFF
    const uint4 *file_meta;       // file chromosome id -> chrom_meta of its annotation id (zeros: none)
    // compact wire format (include/lesseq_b200.h): alternatives to read_cs and blk, null when absent
    const uint8_t *read_cs8;
    const int32_t *blk_start;
    const uint16_t *blk_len;
    const uint8_t *blk_len8;      // instead of blk_len when every block is shorter than 256
    const uint32_t *progress;     // null, or: blocks of blk[] that have landed so far (streamed host->device copy)
    int2 *ext;                    // [n_blocks + 2 n_reads] worst case (header + blocks per general read)
    PartState part;
    unsigned long long *counters; // per read set: [0] FAST reads, [1] ext slots read, [2] PAIR reads, [3] GENERAL reads
};

struct CountParams {
    Tables T;
    PartState part;
    const int2 *ext;
    const uint32_t *line_num;
    const uint8_t *strand_tbl;   // [file strands x annotation strands] bit0: read < event, bit1: equal
    unsigned long long *slot_cnt;   // this read set's class counters
    unsigned long long *slot_mask;  // hash keys (EV_HASH events)
    unsigned long long *bases;      // [E] internal order
    unsigned long long *stats;      // ext slots read by the GENERAL instantiation (byte accounting)
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
    unsigned int *work_counter;    // dynamic event fetch of lsq_solve_small_kernel
    unsigned long long *pairs_valid;  // finalize: total valid (read, event) pairs (statistics)
};

}  // namespace lsq
