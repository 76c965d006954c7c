// Per-read / per-pair device functions of the count path (no warp intrinsics here).
//
// Everything in this header is written against LSQ_HD so that tests/emul/ can compile the very
// same logic with g++ and replay it sequentially on the CPU against the oracle while no GPU is
// attached to the development container.  That harness is test infrastructure: the product
// only ever runs these functions inside the CUDA kernels of lsq_kernels.cuh.
#pragma once
#include "lsq_types.cuh"

#ifdef __CUDACC__
#define LSQ_HD __device__ __forceinline__
#define LSQ_COLD __device__ __noinline__  /* rare paths: kept out of line so the hot loop stays I-cache resident */
#define LSQ_LD(p) __ldg(p)
#define LSQ_LD_STREAM(p) __ldcs(p) /* read-once input streams: evict-first, keep L2 for the tables */
#define LSQ_NO_UNROLL _Pragma("unroll 1")
#else
#define LSQ_HD inline
#define LSQ_COLD inline
#define LSQ_LD(p) (*(p))
#define LSQ_LD_STREAM(p) (*(p))
#define LSQ_NO_UNROLL
#endif

namespace lsq {

LSQ_HD int lsq_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
LSQ_HD double lsq_div(double a, double b) {  // IEEE division, never contracted or approximated
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
LSQ_HD int32_t lsq_max(int32_t a, int32_t b) { return a > b ? a : b; }
LSQ_HD int32_t lsq_min(int32_t a, int32_t b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
// number of elements < x in sorted a[lo,hi)
LSQ_HD uint32_t lower_bound_i32(const int32_t *__restrict__ a, uint32_t lo, uint32_t hi,
                                                    int32_t x) {
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (LSQ_LD(a + mid) < x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// number of elements <= x in sorted a[lo,hi)
LSQ_HD uint32_t upper_bound_i32(const int32_t *__restrict__ a, uint32_t lo, uint32_t hi,
                                                    int32_t x) {
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (LSQ_LD(a + mid) <= x) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Piece containing position x on chromosome c (index into T.pinfo), or 0xFFFFFFFF when x lies
// before the first piece.  One 8-byte bucket entry names the piece (bucket_rank), a crowded bucket
// takes one more entry (its sub-bucket); a crowded sub-bucket, a negative position or one past the
// last bucket takes a binary search over the piece starts.
LSQ_HD uint32_t find_piece(const Tables &T, uint32_t c, int32_t x, PieceInfo &out) {
    const uint4 cm = LSQ_LD(T.chrom_meta + c);
    const uint32_t lo = cm.x, hi = cm.y;
    uint32_t i = lo, j = hi;
    uint32_t n = 0xFFFFFFFFu;  // pieces of [i, j) starting at or below x, when a bucket word tells
    if (x >= 0) {
        const uint32_t b0 = cm.z, nb = cm.w;
        const uint32_t b = (uint32_t)x >> T.bp_shift;
        if (b >= nb) {
            i = j = hi;
            n = 0;
        } else {
            const uint2 e = LSQ_LD(T.bp_bkt + b0 + b);
            const uint32_t xo = (uint32_t)x - (b << T.bp_shift);
            i = e.x;
            if (!bucket_many(e.y)) {
                n = bucket_rank(T.bp_shift, e.y, xo);
            } else {
                const uint32_t ss = bucket_sub_shift(T.bp_shift);
                const uint2 e2 = LSQ_LD(T.bp_sub + (e.y & 0x1FFFFFFFu) + (xo >> ss));
                i = e2.x;
                if (!bucket_many(e2.y)) n = bucket_rank(ss, e2.y, xo & ((1u << ss) - 1u));
                else j = i + (1u << ss) < hi ? i + (1u << ss) : hi;  // at most 2^ss starts in the sub-bucket
            }
        }
    }
    if (n == 0xFFFFFFFFu) {  // number of pieces with pos <= x among [i, j)
        uint32_t a = i, b = j;
        while (a < b) {
            const uint32_t mid = (a + b) >> 1;
            if (LSQ_LD(&T.pinfo[mid].pos) <= x) a = mid + 1;
            else b = mid;
        }
        n = a - i;
    }
    if (i + n == lo) return 0xFFFFFFFFu;
    const uint32_t p = i + n - 1u;
    const uint4 raw = LSQ_LD(reinterpret_cast<const uint4 *>(T.pinfo + p));
    out.pos = (int32_t)raw.x;
    out.cov_end = (int32_t)raw.y;
    out.stab_off = raw.z;
    out.stab_n = raw.w;
    return p;
}

// In-place "erase [b,e) then optionally insert v at b" on one component (.x or .y) of an int2 list.
LSQ_HD void erase_insert(int32_t *comp /* stride 2 */, uint32_t n, uint32_t b, uint32_t e,
                                    bool ins, int32_t v) {
    int shift = (ins ? 1 : 0) - (int)(e - b);
    if (shift < 0) {
        for (uint32_t i = e; i < n; ++i) comp[2 * (i + shift)] = comp[2 * i];
    } else if (shift > 0) {
        for (uint32_t i = n; i > e; --i) comp[2 * (i - 1 + shift)] = comp[2 * (i - 1)];
    }
    if (ins) comp[2 * b] = v;
}

// interval_list::add_interval on a list of n intervals stored in L[0..n) (capacity >= n+1).
// Returns the new length.  General (order dependent) semantics; the caller handles the
// common append/extend case without calling this.
LSQ_HD uint32_t list_add(int2 *L, uint32_t n, int32_t s, int32_t e) {
    uint32_t ss = 0, se = 0, es = 0, ee = 0;
    for (uint32_t i = 0; i < n; ++i) {
        int2 v = L[i];
        ss += v.x < s;
        se += v.y < s;
        es += v.x < e;
        ee += v.y < e;
    }
    bool s_in = (ss - se) == 1u, e_in = (es - ee) == 1u;
    int32_t *base = reinterpret_cast<int32_t *>(L);
    erase_insert(base, n, ss, es, !s_in, s);
    erase_insert(base + 1, n, se, ee, !e_in, e);
    return ss + (s_in ? 0u : 1u) + (n - es);
}


// Read classes (see ReadRec): the table-driven evaluator handles FAST reads, everything else
// (several kept blocks, a start on the first position of a piece -- the only place the
// lower_bound tie-break can apply --, or an over-long block) goes through the general walk.
constexpr uint32_t FAST_MAX_LEN = 0xFFFFu;
LSQ_HD bool is_fast_read(uint32_t n, uint32_t at_piece_start, int2 first, uint32_t piece_stab_n) {
    return n == 1u && !at_piece_start && (uint32_t)(first.y - first.x) <= FAST_MAX_LEN &&
           (piece_stab_n & PIECE_ALL_FAST) != 0;
}

// Slot of read r's ext entry (GENERAL class): even, so the entry is 16-byte aligned; entries of
// consecutive reads cannot overlap because an entry has at most 1 + (raw blocks of r) slots.
LSQ_HD uint32_t ext_slot(uint32_t blk_off_r, uint32_t r) { return (blk_off_r + 2u * r) & ~1u; }

// Sort key of a read that is a candidate of no event: one past the largest real key.
LSQ_HD uint32_t KEY_DROPPED(const Tables &T) { return T.n_pieces * 3u; }

// PAIR class: exactly two kept blocks, not on the first position of the piece, and everything fits
// the 8-byte record: x = (start - piece start) [16 bits] | length of block 0 [11 bits] << 16,
// y = (gap between the blocks - 1) [22 bits] | length of block 1 [... 8 low bits << 22 | strand << 30]
// -- see pack_pair / unpack_pair.  Such reads are evaluated by the general walk straight from
// registers, without an ext entry.
constexpr uint32_t PAIR_MAX_OFF = 0xFFFFu, PAIR_MAX_LEN = 0x7FFu, PAIR_MAX_GAP = 0x3FFFFFu;
LSQ_HD bool is_pair_read(uint32_t n, uint32_t at_piece_start, int2 b0, int2 b1, int32_t piece_pos, uint32_t strand) {
    return n == 2u && !at_piece_start && (uint32_t)(b0.x - piece_pos) <= PAIR_MAX_OFF &&
           (uint32_t)(b0.y - b0.x) <= PAIR_MAX_LEN && (uint32_t)(b1.y - b1.x) <= PAIR_MAX_LEN && b1.x > b0.y &&
           (uint32_t)(b1.x - b0.y - 1) <= PAIR_MAX_GAP && strand < 4u;
}
LSQ_HD ReadRec pack_pair(int2 b0, int2 b1, int32_t piece_pos, uint32_t strand) {
    ReadRec r;
    const uint32_t len1 = (uint32_t)(b1.y - b1.x);
    r.x = (uint32_t)(b0.x - piece_pos) | ((uint32_t)(b0.y - b0.x) << 16) | ((len1 & 0x1Fu) << 27);
    r.y = (uint32_t)(b1.x - b0.y - 1) | ((len1 >> 5) << 22) | (strand << 28);
    return r;
}
LSQ_HD void unpack_pair(ReadRec r, int32_t piece_pos, int2 &b0, int2 &b1, uint32_t &strand) {
    const uint32_t len0 = (r.x >> 16) & 0x7FFu, len1 = ((r.x >> 27) & 0x1Fu) | (((r.y >> 22) & 0x3Fu) << 5);
    b0.x = piece_pos + (int32_t)(r.x & 0xFFFFu);
    b0.y = b0.x + (int32_t)len0;
    b1.x = b0.y + 1 + (int32_t)(r.y & 0x3FFFFFu);
    b1.y = b1.x + (int32_t)len1;
    strand = r.y >> 28;
}

// Result of the per-read load filter.
constexpr uint32_t CLS_FAST = 0, CLS_PAIR = 1, CLS_GENERAL = 2, CLS_DROPPED = 3;
struct FilteredRead {
    uint32_t n;       // kept intervals (0: read dropped)
    uint32_t key;     // piece (FAST), n_pieces + piece (PAIR), 2 n_pieces + piece (GENERAL), KEY_DROPPED
    uint32_t cls;     // CLS_*: key / n_pieces
    uint32_t piece;   // key - cls * n_pieces (undefined when dropped)
    uint32_t stab_off;  // first stab entry of the piece: selects the stab bucket
    uint32_t stab_n;    // stab entries of the piece
    int2 second;      // second kept interval (n >= 2)
    int32_t piece_pos;  // first position of the read's piece
    int2 first;       // first kept interval
    uint32_t strand;  // file strand id of the last registered block
    uint32_t at_piece_start;  // 1 when the first kept start is the first position of its piece
};

// count.cpp:297-327 for one read: every block is tested against the covered regions of its
// chromosome; kept blocks are merged into the read's interval list (work[0..n), used only when
// the read has more than one raw block); the read's chromosome/strand are those of the last
// registered block; the stab piece is looked up for the first kept start.
// The filter is written against a read context that says where its inputs come from, so that the
// kernel can hand it blocks, lookup answers and piece records it has already fetched (several
// reads at a time) while the CPU replay and the fallback path fetch on demand:
//   ctx.blk(j)          raw block j of the read
//   ctx.lookup(j, c, b) for block j = [b.x, b.y) on chromosome c: the piece holding b.x in bits
//                       0-30 (PIECE_NONE: before the first piece) and, in bit 31, whether the
//                       covered region of that piece reaches b.y
//   ctx.pinfo(p)        the 16-byte PieceInfo of piece p
constexpr uint32_t PIECE_NONE = 0x7FFFFFFFu;

// Raw block i and the ids of read r, from the wide or the compact representation of the read set.
LSQ_HD int2 load_block(const FilterParams &P, uint32_t i) {
    if (P.blk_start != nullptr) {
        const int32_t s = LSQ_LD(P.blk_start + i);
        const int32_t len = P.blk_len8 != nullptr ? (int32_t)LSQ_LD(P.blk_len8 + i) : (int32_t)LSQ_LD(P.blk_len + i);
        return make_int2(s, s + len);
    }
    return LSQ_LD(P.blk + i);
}
LSQ_HD int2 load_block_stream(const FilterParams &P, uint32_t i) {
    if (P.blk_start != nullptr) {
        const int32_t s = LSQ_LD_STREAM(P.blk_start + i);
        const int32_t len = P.blk_len8 != nullptr ? (int32_t)LSQ_LD_STREAM(P.blk_len8 + i) : (int32_t)LSQ_LD_STREAM(P.blk_len + i);
        return make_int2(s, s + len);
    }
    return LSQ_LD_STREAM(P.blk + i);
}
LSQ_HD uint32_t load_read_cs(const FilterParams &P, uint64_t r) {
    if (P.read_cs8 != nullptr) {
        const uint32_t v = LSQ_LD_STREAM(P.read_cs8 + r);
        return (v & 0x3Fu) | ((v & 0x40u) << 10) | ((v & 0x80u) << 24);
    }
    return LSQ_LD_STREAM(P.read_cs + r);
}

LSQ_HD uint32_t lookup_direct(const Tables &T, uint32_t c, int2 b) {
    PieceInfo pi;
    const uint32_t pidx = find_piece(T, c, b.x, pi);
    if (pidx == 0xFFFFFFFFu) return PIECE_NONE;
    return pidx | (pi.cov_end >= b.y ? 0x80000000u : 0u);
}

struct DirectCtx {  // everything on demand, lookups through find_piece
    const FilterParams &P;
    uint32_t base;
    LSQ_HD int2 blk(uint32_t j) const { return load_block_stream(P, base + j); }
    LSQ_HD uint32_t lookup(uint32_t, uint32_t c, int2 b) const { return lookup_direct(P.T, c, b); }
    LSQ_HD uint4 pinfo(uint32_t p) const { return LSQ_LD(reinterpret_cast<const uint4 *>(P.T.pinfo + p)); }
};

template <class Ctx>
LSQ_HD FilteredRead filter_read_t(const FilterParams &P, uint64_t r, uint32_t base, uint32_t nraw, uint32_t cs0,
                                  const Ctx &ctx) {
    const Tables &T = P.T;
    FilteredRead out;
    out.n = 0;
    out.key = KEY_DROPPED(T);
    out.first = make_int2(0, 0);
    out.strand = 0;
    out.at_piece_start = 0;
    out.second = make_int2(0, 0);
    out.piece_pos = 0;
    out.cls = CLS_DROPPED;
    out.piece = 0;
    out.stab_off = 0;
    out.stab_n = 0;
    const bool mixed = (cs0 & 0x80000000u) != 0;
    uint32_t chrom = 0xFFFFFFFFu;  // annotation id of the last registered block
    bool registered = false;
    int32_t last_end = 0;
    uint32_t n = 0;
    int2 first = make_int2(0, 0), second = make_int2(0, 0);
    // The read's interval list lives in registers while it has at most two intervals; a third
    // interval or an out-of-order block spills it to the read's ext entry (blocks start one slot
    // after the header), where the general list update works in place.
    int2 *work = P.ext + ext_slot(base, (uint32_t)r) + 1;
    bool spilled = false;
    uint32_t first_chrom = 0xFFFFFFFFu;
    uint32_t first_piece = 0xFFFFFFFFu;  // piece of `first` when it was found while filtering
    bool first_valid = false;
    for (uint32_t j = 0; j < nraw; ++j) {
        const int2 b = ctx.blk(j);
        const uint32_t cs = mixed ? LSQ_LD(P.blk_cs + base + j) : cs0;
        const uint32_t c = LSQ_LD(P.chrom_remap + (cs & 0xFFFFu));
        bool keep;
        uint32_t pidx = 0xFFFFFFFFu;
        if (!(b.x < b.y)) {
            keep = true;  // a nil block is "contained" (interval_list.hpp:398-401)
        } else {
            // contains_interval (interval_list.hpp:396-422): the covered region holding the block
            // start must reach the block end
            keep = false;
            if (c != 0xFFFFFFFFu) {
                const uint32_t res = ctx.lookup(j, c, b);
                pidx = res & PIECE_NONE;
                keep = pidx != PIECE_NONE && (res >> 31) != 0u;
            }
        }
        if (!keep) continue;
        registered = true;
        chrom = c;
        out.strand = (cs >> 16) & 0xFFu;
        if (!(b.x < b.y)) continue;  // add_interval ignores nil intervals
        if (n == 0) {
            first = b;
            last_end = b.y;
            n = 1;
            first_piece = pidx;
            first_valid = true;
            first_chrom = c;
        } else if (b.x >= last_end) {
            if (b.x == last_end) {  // starts where the last interval ends: the two fuse
                last_end = b.y;
                if (n == 1) first.y = b.y;
                if (n == 2) second.y = b.y;
                if (spilled) work[n - 1].y = b.y;
            } else {
                if (n == 1) {
                    second = b;
                    if (spilled) work[1] = b;  // an earlier merge already moved the list to the ext entry
                } else {
                    if (!spilled) {
                        work[0] = first;
                        work[1] = second;
                        spilled = true;
                    }
                    work[n] = b;
                }
                last_end = b.y;
                ++n;
            }
        } else {  // out of order or overlapping: general list update
            if (!spilled) {
                work[0] = first;
                if (n >= 2) work[1] = second;
                spilled = true;
            }
            n = list_add(work, n, b.x, b.y);
            first_valid = first_valid && work[0].x == first.x;
            first = work[0];
            if (n >= 2) second = work[1];
            last_end = work[n - 1].y;
        }
    }
    out.n = n;
    out.first = first;
    if (registered && n > 0 && chrom != 0xFFFFFFFFu) {
        // the read's chromosome is that of the LAST registered block; the piece found while
        // filtering is reused when it was looked up on that chromosome for this very start
        uint32_t p = first_piece;
        PieceInfo pi;
        if (!first_valid || first_chrom != chrom) {
            p = find_piece(T, chrom, first.x, pi);
        } else {
            const uint4 raw = ctx.pinfo(p);
            pi.pos = (int32_t)raw.x;
            pi.cov_end = (int32_t)raw.y;
            pi.stab_off = raw.z;
            pi.stab_n = raw.w;
        }
        if (p != 0xFFFFFFFFu && (pi.stab_n & ~PIECE_ALL_FAST) > 0) {
            out.at_piece_start = pi.pos == first.x ? 1u : 0u;
            out.piece_pos = pi.pos;
            if (n >= 2) out.second = second;
            if (is_fast_read(n, out.at_piece_start, first, pi.stab_n)) out.cls = CLS_FAST;
            else if (is_pair_read(n, out.at_piece_start, first, out.second, pi.pos, out.strand)) out.cls = CLS_PAIR;
            else out.cls = CLS_GENERAL;
            out.key = out.cls * T.n_pieces + p;
            out.piece = p;
            out.stab_off = pi.stab_off;
            out.stab_n = pi.stab_n & ~PIECE_ALL_FAST;
        }
    }
    return out;
}

// The ext entry of a GENERAL read: header {read index, (strand << 24) | n} at its slot, then the
// kept blocks.  Blocks 2.. were spilled by the filter; the first two come from its registers.
LSQ_HD void write_ext_entry(int2 *ext, uint32_t off, uint32_t r, const FilteredRead &fr) {
    ext[off] = make_int2((int32_t)r, (int32_t)((fr.strand << 24) | (fr.n & 0xFFFFFFu)));
    ext[off + 1] = fr.first;
    if (fr.n >= 2u) ext[off + 2] = fr.second;
}

LSQ_HD FilteredRead filter_read(const FilterParams &P, uint64_t r) {
    const uint32_t base = LSQ_LD_STREAM(P.blk_off + r);
    const uint32_t nraw = LSQ_LD_STREAM(P.blk_off + r + 1) - base;
    const uint32_t cs0 = load_read_cs(P, r);
    const DirectCtx ctx{P, base};
    return filter_read_t(P, r, base, nraw, cs0, ctx);
}

// ---------------------------------------------------------------------------------------------
// The piece lookup of one block, split into stages (chromosome -> bucket entry -> piece
// record) so that the kernel can issue the loads of one stage for several blocks
// before it consumes any of them: the chain of dependent L2 accesses, not arithmetic, bounds
// the filter.  A stage hands the block over to find_piece (SP_DEFER) when it meets a case it
// does not cover; both give identical answers (tests/emul replays block_lookup).
// ---------------------------------------------------------------------------------------------
constexpr int SP_DROP = 0, SP_GO = 1, SP_DEFER = 2, SP_SUB = 3;
struct BlockProbe {
    int2 b;          // the block
    uint32_t c;      // file chromosome id
    uint32_t lo;     // first piece of the chromosome
    uint32_t slot;   // bucket slot of b.x (SP_SUB: sub-bucket slot)
    uint32_t xo;     // offset of b.x inside its bucket (SP_SUB: inside its sub-bucket)
};

LSQ_HD int block_stage_chrom(const FilterParams &P, int2 b, uint32_t cs, BlockProbe &q) {
    q.b = b;
    q.c = cs & 0xFFFFu;
    q.lo = 0;
    q.slot = 0;
    q.xo = 0;
    if (!(b.x < b.y)) return SP_DROP;         // nil blocks are never looked up
    const uint4 cm = LSQ_LD(P.file_meta + q.c);  // chrom_meta of the file chromosome's annotation id, or zeros
    if (cm.x == cm.y) return SP_DROP;         // chromosome without annotation (no pieces)
    if (b.x < 0) return SP_DEFER;
    const uint32_t bk = (uint32_t)b.x >> P.T.bp_shift;
    if (bk >= cm.w) return SP_DEFER;
    q.lo = cm.x;
    q.slot = cm.z + bk;
    q.xo = (uint32_t)b.x - (bk << P.T.bp_shift);
    return SP_GO;
}

// The piece of the block's start from its bucket entry e: SP_GO with p (PIECE_NONE: before the first
// piece of the chromosome), or SP_SUB for a crowded bucket: q then names the sub-bucket entry.
LSQ_HD int block_stage_piece(const Tables &T, BlockProbe &q, uint2 e, uint32_t &p) {
    p = PIECE_NONE;
    if (bucket_many(e.y)) {
        const uint32_t ss = bucket_sub_shift(T.bp_shift);
        q.slot = (e.y & 0x1FFFFFFFu) + (q.xo >> ss);
        q.xo &= (1u << ss) - 1u;
        return SP_SUB;
    }
    const uint32_t n = bucket_rank(T.bp_shift, e.y, q.xo);
    if (e.x + n != q.lo) p = e.x + n - 1u;
    return SP_GO;
}
// ... from its sub-bucket entry: SP_GO with p, or SP_DEFER (find_piece searches a crowded sub-bucket).
LSQ_HD int block_stage_sub(const Tables &T, const BlockProbe &q, uint2 e, uint32_t &p) {
    p = PIECE_NONE;
    if (bucket_many(e.y)) return SP_DEFER;
    const uint32_t n = bucket_rank(bucket_sub_shift(T.bp_shift), e.y, q.xo);
    if (e.x + n != q.lo) p = e.x + n - 1u;
    return SP_GO;
}

// SP_DEFER: the full search.
LSQ_HD uint32_t block_stage_defer(const FilterParams &P, const BlockProbe &q) {
    return lookup_direct(P.T, LSQ_LD(P.chrom_remap + q.c), q.b);
}

// The lookup answer from the piece and its covered-region end.
LSQ_HD uint32_t block_stage_answer(const BlockProbe &q, uint32_t p, int32_t cov_end) {
    return p == PIECE_NONE ? PIECE_NONE : (p | (cov_end >= q.b.y ? 0x80000000u : 0u));
}

// Sequential composition of the stages (CPU replay; the kernel runs the same stages batched).
LSQ_HD uint32_t block_lookup(const FilterParams &P, int2 b, uint32_t cs) {
    BlockProbe q;
    int st = block_stage_chrom(P, b, cs, q);
    if (st == SP_DROP) return PIECE_NONE;
    uint32_t res = PIECE_NONE;
    if (st == SP_GO) {
        uint32_t p;
        st = block_stage_piece(P.T, q, LSQ_LD(P.T.bp_bkt + q.slot), p);
        if (st == SP_SUB) st = block_stage_sub(P.T, q, LSQ_LD(P.T.bp_sub + q.slot), p);
        if (st == SP_GO && p != PIECE_NONE) res = block_stage_answer(q, p, LSQ_LD(&P.T.pinfo[p].cov_end));
    }
    if (st == SP_DEFER) res = block_stage_defer(P, q);
    return res;
}

// The filter for the common read -- exactly one raw block b, chromosome/strand ids from read_cs --
// given the lookup answer `res` of that block (nil blocks and blocks on chromosomes without
// annotation answer PIECE_NONE) and, for a kept block, the PieceInfo `raw` of the answered piece.
// Same result as filter_read_t on such a read, without its interval-list machinery.
LSQ_HD bool single_kept(uint32_t res) { return (res & PIECE_NONE) != PIECE_NONE && (res >> 31) != 0u; }
LSQ_HD FilteredRead single_decide(const Tables &T, int2 b, uint32_t cs0, uint32_t res, uint4 raw) {
    FilteredRead out;
    out.n = 0;
    out.key = KEY_DROPPED(T);
    out.first = b;
    out.strand = (cs0 >> 16) & 0xFFu;
    out.at_piece_start = 0;
    out.second = make_int2(0, 0);
    out.piece_pos = 0;
    out.cls = CLS_DROPPED;
    out.piece = 0;
    out.stab_off = 0;
    out.stab_n = 0;
    if (!single_kept(res)) return out;
    out.n = 1;
    const uint32_t stab_n = raw.w;
    if ((stab_n & ~PIECE_ALL_FAST) == 0u) return out;  // candidate of no event
    const uint32_t p = res & PIECE_NONE;
    out.at_piece_start = (int32_t)raw.x == b.x ? 1u : 0u;
    out.piece_pos = (int32_t)raw.x;
    out.cls = is_fast_read(1u, out.at_piece_start, b, stab_n) ? CLS_FAST : CLS_GENERAL;
    out.key = out.cls * T.n_pieces + p;
    out.piece = p;
    out.stab_off = raw.z;
    out.stab_n = stab_n & ~PIECE_ALL_FAST;
    return out;
}

// The load filter as the kernel runs it: every block answered by block_lookup, single-block reads
// decided by single_decide, all others by filter_read_t.
struct StagedCtx {
    const FilterParams &P;
    uint32_t base, cs0;
    LSQ_HD int2 blk(uint32_t j) const { return load_block_stream(P, base + j); }
    LSQ_HD uint32_t lookup(uint32_t j, uint32_t, int2 b) const {
        const uint32_t cs = (cs0 & 0x80000000u) ? LSQ_LD(P.blk_cs + base + j) : cs0;
        return block_lookup(P, b, cs);
    }
    LSQ_HD uint4 pinfo(uint32_t p) const { return LSQ_LD(reinterpret_cast<const uint4 *>(P.T.pinfo + p)); }
};
LSQ_HD FilteredRead filter_dispatch(const FilterParams &P, uint64_t r) {
    const uint32_t base = LSQ_LD_STREAM(P.blk_off + r);
    const uint32_t nraw = LSQ_LD_STREAM(P.blk_off + r + 1) - base;
    const uint32_t cs0 = load_read_cs(P, r);
    if (nraw == 1u && !(cs0 & 0x80000000u)) {
        const int2 b = load_block_stream(P, base);
        const uint32_t res = block_lookup(P, b, cs0);
        const uint4 raw = single_kept(res) ? LSQ_LD(reinterpret_cast<const uint4 *>(P.T.pinfo + (res & PIECE_NONE)))
                                           : make_uint4(0u, 0u, 0u, 0u);
        return single_decide(P.T, b, cs0, res, raw);
    }
    const StagedCtx ctx{P, base, cs0};
    return filter_read_t(P, r, base, nraw, cs0, ctx);
}

// ---------------------------------------------------------------------------------------------
// Per (read, event) pair
// ---------------------------------------------------------------------------------------------
struct ReadView {
    int2 b0, b1, b2;  // kept blocks 0..2 held in registers (reads with more blocks read the rest from blk)
    uint32_t n;       // kept blocks
    uint32_t ridx;    // original read index (names the read: "read-<line_num[ridx]>")
    const int2 *blk;  // kept blocks 0..n-1 (GENERAL class: inside the ext entry)
    LSQ_HD int2 block(uint32_t j) const { return j == 0 ? b0 : (j == 1 ? b1 : (j == 2 ? b2 : LSQ_LD(blk + j))); }
    LSQ_HD int32_t last_end() const { return block(n - 1).y; }
    LSQ_HD uint32_t read_index() const { return ridx; }
};

// "read-<N>" < name for tie mode 2: the name is "read-" + suffix, so compare the decimal digits
// of N with the suffix bytewise (std::string operator<).
LSQ_COLD bool digits_less(uint32_t N, const char *__restrict__ sfx, uint32_t len) {
    char d[10];
    int nd = 0;
    do {
        d[nd++] = (char)('0' + N % 10u);
        N /= 10u;
    } while (N);
    for (int i = 0; i < nd; ++i) {
        if ((uint32_t)i >= len) return false;  // suffix is a proper prefix of the digits
        unsigned char a = (unsigned char)d[nd - 1 - i], b = (unsigned char)sfx[i];
        if (a != b) return a < b;
    }
    return (uint32_t)nd < len;  // digits are a proper prefix of the suffix
}

// Read_single::build (common/read.h:204-274) for S <= 64: segment bit mask + matched length.
// `seg` may point to shared or global memory.
LSQ_HD void walk_mask(const ReadView &rv, const int2 *seg, uint32_t S, uint64_t &mask,
                                          uint32_t &matched) {
    uint32_t it = 0, b = 0;
    int32_t cont = 0;  // the reference's cur_exon_start
    bool found = false;
    uint64_t m = 0;
    uint32_t ml = 0;
    LSQ_NO_UNROLL
    while (b < rv.n) {
        int2 blk = rv.block(b);
        int32_t cs = blk.x;
        const int32_t ce = blk.y;
        bool consumed = false;
        LSQ_NO_UNROLL
        while (it < S) {
            int2 sg = seg[it];
            if (!(sg.x < ce)) break;
            cont = lsq_max(cont, sg.x);
            if (cs >= cont && cs < sg.y) {
                if (!found) found = true;
                else if (cs > cont) break;
                m |= 1ull << it;
                cont = lsq_min(sg.y, ce);
                ml += (uint32_t)(cont - cs);
                if (ce < sg.y) {  // block ends inside this segment: the cursor stays on it
                    consumed = true;
                    break;
                } else if (ce == sg.y) {
                    consumed = true;
                    cs = ce;
                } else {
                    cs = sg.y;  // only an abutting next segment can continue the block
                }
            } else if (cont > sg.x && cont < sg.y) {
                break;
            }
            ++it;
        }
        if (!consumed) break;  // block not fully matched: the read is truncated here
        ++b;
    }
    mask = m;
    matched = ml;
}

// Wide events (S > 64): the same walk, checking one form on the fly with its prefix counts
// pf[s] = number of form segments with index < s.  The read's index list must be a contiguous
// run of the form's list (common/read.h:44-79).
// (Cold functions take scalars and return packed values: a reference to a caller's local would
// force that local into thread-private memory in the hot kernel.)  Returns (ok << 32) | matched.
LSQ_COLD uint64_t walk_wide(int2 b0, int2 b1, int2 b2, uint32_t n, const int2 *blk_ptr, const int2 *seg, uint32_t S,
                            const uint32_t *__restrict__ pf) {
    ReadView rv;
    rv.b0 = b0;
    rv.b1 = b1;
    rv.b2 = b2;
    rv.n = n;
    rv.ridx = 0;
    rv.blk = blk_ptr;
    uint32_t it = 0, b = 0;
    int32_t cont = 0;
    bool found = false, ok = true;
    uint32_t ml = 0;
    uint32_t last_added = 0xFFFFFFFFu, last_rank = 0;
    while (b < rv.n) {
        int2 blk = rv.block(b);
        int32_t cs = blk.x;
        const int32_t ce = blk.y;
        bool consumed = false;
        while (it < S) {
            int2 sg = seg[it];
            if (!(sg.x < ce)) break;
            cont = lsq_max(cont, sg.x);
            if (cs >= cont && cs < sg.y) {
                if (!found) found = true;
                else if (cs > cont) break;
                if (it != last_added) {
                    uint32_t rk = LSQ_LD(pf + it);
                    bool member = LSQ_LD(pf + it + 1) != rk;
                    if (!member || (last_added != 0xFFFFFFFFu && rk != last_rank + 1)) ok = false;
                    last_added = it;
                    last_rank = rk;
                }
                cont = lsq_min(sg.y, ce);
                ml += (uint32_t)(cont - cs);
                if (ce < sg.y) {
                    consumed = true;
                    break;
                } else if (ce == sg.y) {
                    consumed = true;
                    cs = ce;
                } else {
                    cs = sg.y;
                }
            } else if (cont > sg.x && cont < sg.y) {
                break;
            }
            ++it;
        }
        if (!consumed) break;
        ++b;
    }
    return ((uint64_t)((ok && last_added != 0xFFFFFFFFu) ? 1u : 0u) << 32) | ml;
}

// Table-driven evaluation of a FAST-class read [s,e) against a FAST stab entry: the block starts
// inside chain segment 0 (not on the event start, so no tie-break applies); it is matched up to
// min(e, end of the chain) and covers chain segments 0..i where c[i] is the first chain end >= e.
// Same result as walk_mask + the compatibility test + the validity rule of eval_pair, without
// touching the event tables.  Returns the form mask of a valid read (0 otherwise).
LSQ_HD uint32_t eval_fast(const StabEntry &en, int32_t s, int32_t e, uint32_t &matched) {
    matched = 0;
    const int32_t cend = en.c[2];  // c[] is padded with the last chain end
    uint32_t cm;
    if (e <= en.c[0]) cm = en.cmask[0];
    else if (e <= en.c[1]) cm = en.cmask[1];
    else cm = en.cmask[2];
    if (cm == 0) return 0;
    const uint32_t total = (uint32_t)(e - s);
    uint32_t ml = total;
    if (e > cend) {  // the block overhangs the chain: truncated match, IEEE rule of count.cpp:441-442
        ml = (uint32_t)(cend - s);
        if (!(lsq_div((double)ml, (double)total) > 0.98)) return 0;
    }
    matched = ml;
    return cm;
}

// Two-block (PAIR class) read against a FAST stab entry of a plain event (S <= 64): block 0 starts
// in chain segment 0 and is resolved with the entry's chain ends exactly as in eval_fast; block 1
// can only attach when block 0 ended exactly on a segment end, and then only at the exact start
// of a later segment (read.h:226-262: `cur_start > cur_exon_start -> break`), from where it runs
// through abutting segments.  Same result as walk_mask + the tests of eval_pair for such reads,
// without the generic block/segment loops.  No tie-break: PAIR reads never start on a piece start.
LSQ_HD uint64_t eval_pair2(const StabEntry &en, const EvHdr &h, const int2 *seg, const uint64_t *fm, int2 b0,
                           int2 b1, uint32_t total_len, uint32_t &matched) {
    matched = 0;
    const uint32_t j0 = en.meta & 0xFFFFu, nchain = (en.meta >> 16) & 3u;
    const int32_t cend = en.c[2];
    uint32_t i0 = b0.y <= en.c[0] ? 0u : (b0.y <= en.c[1] ? 1u : 2u);
    if (i0 >= nchain) i0 = nchain - 1u;
    uint64_t rmask = ((2ull << (j0 + i0)) - 1ull) & ~((1ull << j0) - 1ull);
    uint32_t ml;
    if (b0.y > cend) {
        ml = (uint32_t)(cend - b0.x);  // block 0 overhangs its chain: the read is truncated there
    } else {
        ml = (uint32_t)(b0.y - b0.x);
        const int32_t ci = i0 == 0u ? en.c[0] : (i0 == 1u ? en.c[1] : en.c[2]);
        if (b0.y == ci) {  // consumed on a segment end: block 1 may attach at a later segment start
            uint32_t it = j0 + i0 + 1u;
            const uint32_t S = h.S;
            LSQ_NO_UNROLL
            while (it < S) {
                int2 sg = seg[it];
                if (!(sg.x < b1.y)) break;
                if (b1.x < sg.y) {
                    if (b1.x == sg.x) {
                        int32_t cs = b1.x;
                        while (true) {
                            rmask |= 1ull << it;
                            const int32_t m = lsq_min(sg.y, b1.y);
                            ml += (uint32_t)(m - cs);
                            if (b1.y <= sg.y) break;  // block 1 consumed
                            cs = sg.y;
                            ++it;
                            if (!(it < S)) break;
                            const int2 nx = seg[it];
                            if (nx.x != sg.y) break;  // no abutting segment: truncated
                            sg = nx;
                        }
                    }
                    break;  // starts inside or before this segment without matching its start
                }
                ++it;
            }
        }
    }
    const uint32_t hi = 63u - (uint32_t)lsq_clz64(rmask);
    const uint64_t low = rmask & (0ull - rmask);
    const uint64_t span = ((2ull << hi) - 1ull) & ~(low - 1ull);
    uint64_t cm;
    if (h.K <= 2u) {
        cm = (uint64_t)((fm[0] & span) == rmask);
        if (h.K == 2u) cm |= (uint64_t)((fm[1] & span) == rmask) << 1;
    } else {
        cm = 0;
        LSQ_NO_UNROLL
        for (uint32_t k = 0; k < h.K; ++k) cm |= (uint64_t)((fm[k] & span) == rmask) << k;
    }
    if (cm == 0) return 0;
    if (ml != total_len && !(lsq_div((double)ml, (double)total_len) > 0.98)) return 0;
    matched = ml;
    return cm;
}

LSQ_HD uint32_t hash_mask(uint64_t m) {
    m ^= m >> 33;
    m *= 0xff51afd7ed558ccdULL;
    m ^= m >> 33;
    return (uint32_t)m;
}

// The lower_bound skip of count.cpp:429-432: a read starting exactly at the event start is
// passed over when its key (start, end, strand, name) sorts before the probe
// (gs, ge, event strand, event name) under the comparator of count.cpp:64-85.
LSQ_COLD bool tie_resolve(uint32_t strand_rel, uint32_t tie_mode, uint32_t line_num, const char *sfx,
                          uint32_t sfx_len) {
    if (strand_rel & 1u) return true;      // read strand string < event strand string
    if (!(strand_rel & 2u)) return false;  // greater
    if (tie_mode == 0u) return false;
    if (tie_mode == 1u) return true;
    return digits_less(line_num, sfx, sfx_len);
}

LSQ_HD bool tie_skipped(const CountParams &P, const EvHdr &h, const ReadView &rv, uint32_t strand) {
    if (rv.b0.x != h.gs) return false;
    int32_t r_end = rv.last_end();
    if (r_end < h.ge) return true;
    if (r_end > h.ge) return false;
    // a read set without comment/header lines carries no line numbers: read i is "read-<i+1>"
    const uint32_t ridx = rv.read_index();
    const uint32_t ln = P.line_num ? LSQ_LD(P.line_num + ridx) : ridx + 1u;
    return tie_resolve(LSQ_LD(P.strand_tbl + strand * P.T.n_ann_strands + h.strand), h.flags & EV_TIE_MASK, ln,
                       P.T.tie_pool + h.tie_off, h.tie_len);
}

// One (read, event) pair -> form compatibility mask of a VALID read (0 otherwise) and matched bases.
// Validity (count.cpp:438-448): compatible with some form and matched/total > 0.98 as IEEE doubles.
LSQ_HD uint64_t eval_pair(const CountParams &P, const EvHdr &h, const int2 *seg,
                                              const uint64_t *fm, const ReadView &rv, uint32_t strand,
                                              uint32_t total_len, uint32_t &matched) {
    matched = 0;
    if (tie_skipped(P, h, rv, strand)) return 0;
    uint64_t cm = 0;
    uint32_t ml = 0;
    if (!(h.flags & EV_WIDE)) {
        uint64_t rmask;
        walk_mask(rv, seg, h.S, rmask, ml);
        if (rmask == 0) return 0;
        // contiguous-run test (read.h:44-79) as a bit identity: within the span of the read's
        // segments the form must contain exactly the read's segments.
        uint32_t hi = 63u - (uint32_t)lsq_clz64(rmask);
        uint64_t low = rmask & (0ull - rmask);
        uint64_t span = ((2ull << hi) - 1ull) & ~(low - 1ull);
        LSQ_NO_UNROLL
        for (uint32_t k = 0; k < h.K; ++k) cm |= (uint64_t)((fm[k] & span) == rmask) << k;
    } else {
        const uint32_t *pf = P.T.pfx + h.pfx_off;
        LSQ_NO_UNROLL
        for (uint32_t k = 0; k < h.K; ++k) {
            const uint64_t w = walk_wide(rv.b0, rv.b1, rv.b2, rv.n, rv.blk, seg, h.S, pf + (size_t)k * (h.S + 1u));
            ml = (uint32_t)w;
            cm |= (w >> 32) << k;
        }
    }
    if (cm == 0) return 0;
    // count.cpp:441-442 as an IEEE double division; a fully matched read has ratio exactly 1
    if (ml != total_len && !(lsq_div((double)ml, (double)total_len) > 0.98)) return 0;
    matched = ml;
    return cm;
}


}  // namespace lsq
