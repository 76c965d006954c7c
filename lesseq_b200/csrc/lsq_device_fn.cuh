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
LSQ_HD double lsq_mul(double a, double b) {  // IEEE multiply / add that nvcc cannot contract into an FMA
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
LSQ_HD double lsq_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
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
    out.sigw = raw.z;
    out.cutw = raw.w;
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


// Slot of read r's scratch entry in ext[]: even, so the entry is 16-byte aligned; entries of
// consecutive reads cannot overlap because an entry has at most 1 + (raw blocks of r) slots.
LSQ_HD uint32_t ext_slot(uint32_t blk_off_r, uint32_t r) { return (blk_off_r + 2u * r) & ~1u; }

// Result of the per-read load filter.
struct FilteredRead {
    uint32_t n;       // kept intervals (0: read dropped)
    uint32_t piece;   // piece holding the start of the first kept interval; PIECE_NONE: candidate of no event
    uint32_t piece2;  // piece holding the start of the second kept interval when known, else PIECE_NONE
    uint32_t strand;  // file strand id of the last registered block
    uint32_t mixed;   // the read's blocks carry their own chromosome/strand ids (chimeric read)
    int2 first;       // first kept interval
    int2 second;      // second kept interval (n >= 2)
    uint4 raw;        // PieceInfo of `piece` (pos, cov_end, sigw, cutw)
};

// count.cpp:297-327 for one read: every block is tested against the covered regions of its
// chromosome; kept blocks are merged into the read's interval list (work[0..n), used only when
// the read has more than one raw block); the read's chromosome/strand are those of the last
// registered block; the stab piece is looked up for the first kept start.
// The filter is written against a read context that says where its inputs come from, so that the
// kernel can hand it blocks, lookup answers and piece records it has already fetched (several
// reads at a time) while the CPU replay and the general kernel fetch on demand:
//   ctx.blk(j)          raw block j of the read
//   ctx.lookup(j, c, b) for block j = [b.x, b.y) on chromosome c: the piece holding b.x in bits
//                       0-30 (PIECE_NONE: before the first piece) and, in bit 31, whether the
//                       covered region of that piece reaches b.y
//   ctx.pinfo(p)        the 16-byte PieceInfo of piece p
constexpr uint32_t PIECE_NONE = 0x7FFFFFFFu;

// Raw block i and the ids of read r, from the wide or the compact representation of the read set.
LSQ_HD int2 load_block(const ScanParams &P, uint32_t i) {
    if (P.blk_start != nullptr) {
        const int32_t s = LSQ_LD(P.blk_start + i);
        const int32_t len = P.blk_len8 != nullptr ? (int32_t)LSQ_LD(P.blk_len8 + i) : (int32_t)LSQ_LD(P.blk_len + i);
        return make_int2(s, s + len);
    }
    return LSQ_LD(P.blk + i);
}
LSQ_HD int2 load_block_stream(const ScanParams &P, uint32_t i) {
    if (P.blk_start != nullptr) {
        const int32_t s = LSQ_LD_STREAM(P.blk_start + i);
        const int32_t len = P.blk_len8 != nullptr ? (int32_t)LSQ_LD_STREAM(P.blk_len8 + i) : (int32_t)LSQ_LD_STREAM(P.blk_len + i);
        return make_int2(s, s + len);
    }
    return LSQ_LD_STREAM(P.blk + i);
}
LSQ_HD uint32_t expand_cs8(uint32_t v) { return (v & 0x3Fu) | ((v & 0x40u) << 10) | ((v & 0x80u) << 24); }
LSQ_HD uint32_t load_read_cs(const ScanParams &P, uint64_t r) {
    if (P.read_cs8 != nullptr) return expand_cs8(LSQ_LD_STREAM(P.read_cs8 + r));
    return LSQ_LD_STREAM(P.read_cs + r);
}
// Raw blocks of read r (the general kernel's access path; the scan kernel reads whole groups).
LSQ_HD uint32_t load_read_nraw(const ScanParams &P, uint64_t r) {
    if (P.blk_cnt != nullptr) return LSQ_LD(P.blk_cnt + r);
    return LSQ_LD(P.blk_off + r + 1) - LSQ_LD(P.blk_off + r);
}

LSQ_HD uint32_t lookup_direct(const Tables &T, uint32_t c, int2 b) {
    PieceInfo pi;
    const uint32_t pidx = find_piece(T, c, b.x, pi);
    if (pidx == 0xFFFFFFFFu) return PIECE_NONE;
    return pidx | (pi.cov_end >= b.y ? 0x80000000u : 0u);
}

struct DirectCtx {  // everything on demand, lookups through find_piece
    const ScanParams &P;
    uint32_t base;
    LSQ_HD int2 blk(uint32_t j) const { return load_block(P, base + j); }
    LSQ_HD uint32_t lookup(uint32_t, uint32_t c, int2 b) const { return lookup_direct(P.T, c, b); }
    LSQ_HD uint4 pinfo(uint32_t p) const { return LSQ_LD(reinterpret_cast<const uint4 *>(P.T.pinfo + p)); }
};

template <class Ctx>
LSQ_HD FilteredRead filter_read_t(const ScanParams &P, uint64_t r, uint32_t base, uint32_t nraw, uint32_t cs0,
                                  const Ctx &ctx) {
    const Tables &T = P.T;
    FilteredRead out;
    out.n = 0;
    out.piece = PIECE_NONE;
    out.piece2 = PIECE_NONE;
    out.first = make_int2(0, 0);
    out.strand = 0;
    out.second = make_int2(0, 0);
    out.raw = make_uint4(0u, 0u, 0u, 0u);
    const bool mixed = (cs0 & 0x80000000u) != 0;
    out.mixed = mixed ? 1u : 0u;
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
    uint32_t second_piece = PIECE_NONE;  // piece of `second` likewise
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
                    second_piece = pidx;
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
            second_piece = PIECE_NONE;
            last_end = work[n - 1].y;
        }
    }
    out.n = n;
    out.first = first;
    if (n >= 2) {
        out.second = second;
        out.piece2 = second_piece;
    }
    if (registered && n > 0 && chrom != 0xFFFFFFFFu) {
        // the read's chromosome is that of the LAST registered block; the piece found while
        // filtering is reused when it was looked up on that chromosome for this very start
        uint32_t p = first_piece;
        if (!first_valid || first_chrom != chrom) {
            PieceInfo pi;
            p = find_piece(T, chrom, first.x, pi);
            out.raw = make_uint4((uint32_t)pi.pos, (uint32_t)pi.cov_end, pi.sigw, pi.cutw);
            out.piece2 = PIECE_NONE;
        } else {
            out.raw = ctx.pinfo(p);
        }
        if (p != 0xFFFFFFFFu && (out.raw.z & PIECE_CAND)) out.piece = p;
    }
    return out;
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

LSQ_HD int block_stage_chrom(const ScanParams &P, int2 b, uint32_t cs, BlockProbe &q) {
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
LSQ_HD uint32_t block_stage_defer(const ScanParams &P, const BlockProbe &q) {
    return lookup_direct(P.T, LSQ_LD(P.chrom_remap + q.c), q.b);
}

// The lookup answer from the piece and its covered-region end.
LSQ_HD uint32_t block_stage_answer(const BlockProbe &q, uint32_t p, int32_t cov_end) {
    return p == PIECE_NONE ? PIECE_NONE : (p | (cov_end >= q.b.y ? 0x80000000u : 0u));
}

// Sequential composition of the stages (CPU replay; the kernel runs the same stages batched).
LSQ_HD uint32_t block_lookup(const ScanParams &P, int2 b, uint32_t cs) {
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

// ---------------------------------------------------------------------------------------------
// Signature routing.  The walk of common/read.h:204-274 for a read whose first kept block starts in
// piece p depends on the read only through (a) where its block ends fall among the cuts of their pieces
// (PieceInfo), (b) whether a later block starts exactly on a piece start, and -- through the validity
// rule matched/total > 0.98 (count.cpp:441-442) -- (c) the lengths, but only when some event matches the
// read PARTLY.  So reads are counted per SIGNATURE and every distinct signature is evaluated once, on a
// representative read, by the general evaluator (sig_rep_single / sig_rep_pair + eval_pair in
// lsq_expand_kernel).  A read takes that route only when every partial match is certainly invalid
// (the missing part is at least total/50 long, so matched/total <= 0.98 in exact arithmetic and the IEEE
// quotient cannot exceed the double nearest 0.98); then a (read, event) pair is valid iff it matches fully,
// which the signature decides, and its matched length is the read's total length.  Everything else goes
// to lsq_general_kernel, which walks the read itself.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t RT_NONE = 0, RT_SINGLE = 1, RT_PAIR = 2, RT_GENERAL = 3, RT_TRIPLE = 4;

// Number of cuts of the piece `raw` below block end e (the block starts inside the piece); exact: e hits cut k;
// q_prev: the largest cut below e (k > 0).  The first two cuts come with the piece record.
LSQ_HD uint32_t cut_category(const Tables &T, uint4 raw, int32_t e, bool &exact, int32_t &q_prev) {
    const uint32_t n = piece_ncuts(raw.z);
    const int32_t pos = (int32_t)raw.x;
    const int32_t q0 = pos + (int32_t)(raw.w & 0xFFFFu), q1 = pos + (int32_t)(raw.w >> 16);
    q_prev = pos;
    exact = false;
    if (n == 0u) return 0u;
    if (e <= q0) {
        exact = e == q0;
        return 0u;
    }
    q_prev = q0;
    if (n == 1u) return 1u;
    if (e <= q1) {
        exact = e == q1;
        return 1u;
    }
    q_prev = q1;
    uint32_t k = 2;
    const uint32_t so = raw.z & PIECE_SIG_MASK;
    LSQ_NO_UNROLL
    while (k < n) {  // one piece in twenty has more than two cuts
        const int32_t q = LSQ_LD(T.sig_cut + so + k);
        if (e <= q) {
            exact = e == q;
            break;
        }
        q_prev = q;
        ++k;
    }
    return k;
}

// missing * 50 >= total: a match that misses `missing` of `total` bases fails matched/total > 0.98
LSQ_HD bool surely_invalid(uint32_t missing, uint32_t total) { return (uint64_t)missing * 50ull >= (uint64_t)total; }

// One kept block b of a read whose piece record is `raw` (a candidate piece): RT_SINGLE with the signature
// slot, RT_NONE (valid for no event), or RT_GENERAL.
LSQ_HD uint32_t route_single(const Tables &T, int2 b, uint4 raw, uint32_t &sig) {
    sig = 0;
    const uint32_t n = piece_ncuts(raw.z);
    if (n == PIECE_NO_SIG) return RT_GENERAL;
    if ((raw.z & PIECE_TIE) && (int32_t)raw.x == b.x) return RT_GENERAL;  // the lower_bound tie-break may apply
    bool exact;
    int32_t q_prev;
    const uint32_t k = cut_category(T, raw, b.y, exact, q_prev);
    // events whose chain ends below b.y match the read up to a cut <= q_prev only
    if (k > 0u && !surely_invalid((uint32_t)(b.y - q_prev), (uint32_t)(b.y - b.x))) return RT_GENERAL;
    if (k == n) return RT_NONE;  // the block runs past every chain
    sig = (raw.z & PIECE_SIG_MASK) + k;
    return RT_SINGLE;
}

// Two kept blocks b0 < b1 (disjoint, not abutting) on one chromosome; raw0 / raw1 the piece records of
// their starts (raw0 a candidate piece).  RT_PAIR with the signature key, RT_NONE or RT_GENERAL.
LSQ_HD uint32_t route_pair(const Tables &T, int2 b0, int2 b1, uint32_t p0, uint4 raw0, uint32_t p1, uint4 raw1,
                           uint64_t &key) {
    key = 0;
    const uint32_t n0 = piece_ncuts(raw0.z);
    if (n0 == PIECE_NO_SIG) return RT_GENERAL;
    if ((raw0.z & PIECE_TIE) && (int32_t)raw0.x == b0.x) return RT_GENERAL;
    const uint32_t len0 = (uint32_t)(b0.y - b0.x), len1 = (uint32_t)(b1.y - b1.x), total = len0 + len1;
    // an event that does not take block 1 whole-heartedly matches at most block 0
    if (!surely_invalid(len1, total)) return RT_GENERAL;
    bool exact0, exact1;
    int32_t qp;
    const uint32_t k0 = cut_category(T, raw0, b0.y, exact0, qp);
    // block 0 must end exactly on a segment end of the event for block 1 to attach (read.h:226-262); an end
    // that is no cut is a segment end of no event; so is one past every chain
    if (k0 == n0 || !exact0) return RT_NONE;
    // ... and block 1 must start exactly on a later segment start, which is a piece start
    if ((int32_t)raw1.x != b1.x || !(raw1.z & PIECE_CAND)) return RT_NONE;
    const uint32_t n1 = piece_ncuts(raw1.z);
    if (n1 == PIECE_NO_SIG) return RT_GENERAL;
    const uint32_t k1 = cut_category(T, raw1, b1.y, exact1, qp);
    if (k1 > 0u && !surely_invalid((uint32_t)(b1.y - qp), total)) return RT_GENERAL;
    if (k1 == n1) return RT_NONE;
    key = pair_key(p0, k0, p1, k1);
    return RT_PAIR;
}

// Three kept blocks b0 < b1 < b2 (disjoint, not abutting) on one chromosome.  RT_TRIPLE with the prefix key
// (pair_key of blocks 0 and 1) and the cut category k2 of block 2, RT_NONE or RT_GENERAL.  Block 1 must be a whole
// chain link for any event to take all three: it starts on a piece start and ends exactly on a cut; whatever
// breaks that leaves at most blocks 0 and 1 matched, which is certainly invalid when block 2 is at least total/50.
LSQ_HD uint32_t route_triple(const Tables &T, int2 b0, int2 b1, int2 b2, uint32_t p0, uint4 raw0, uint32_t p1, uint4 raw1,
                             uint4 raw2, uint64_t &prefix, uint32_t &k2) {
    prefix = 0;
    k2 = 0;
    const uint32_t n0 = piece_ncuts(raw0.z);
    if (n0 == PIECE_NO_SIG) return RT_GENERAL;
    if ((raw0.z & PIECE_TIE) && (int32_t)raw0.x == b0.x) return RT_GENERAL;
    const uint32_t len2 = (uint32_t)(b2.y - b2.x);
    const uint32_t total = (uint32_t)(b0.y - b0.x) + (uint32_t)(b1.y - b1.x) + len2;
    if (!surely_invalid(len2, total)) return RT_GENERAL;
    bool exact;
    int32_t qp;
    const uint32_t k0 = cut_category(T, raw0, b0.y, exact, qp);
    if (k0 == n0 || !exact) return RT_NONE;
    if ((int32_t)raw1.x != b1.x || !(raw1.z & PIECE_CAND)) return RT_NONE;
    const uint32_t n1 = piece_ncuts(raw1.z);
    if (n1 == PIECE_NO_SIG) return RT_GENERAL;
    const uint32_t k1 = cut_category(T, raw1, b1.y, exact, qp);
    if (k1 == n1 || !exact) return RT_NONE;
    if ((int32_t)raw2.x != b2.x || !(raw2.z & PIECE_CAND)) return RT_NONE;
    const uint32_t n2 = piece_ncuts(raw2.z);
    if (n2 == PIECE_NO_SIG) return RT_GENERAL;
    k2 = cut_category(T, raw2, b2.y, exact, qp);
    if (k2 > 0u && !surely_invalid((uint32_t)(b2.y - qp), total)) return RT_GENERAL;
    if (k2 == n2) return RT_NONE;
    prefix = pair_key(p0, k0, p1, k1);
    return RT_TRIPLE;
}

// Route of a filtered read.  raw1 is only read for two-block reads whose second piece is known.
template <class PinfoFn>
LSQ_HD uint32_t route_read(const ScanParams &P, const FilteredRead &fr, PinfoFn pinfo, uint32_t &sig, uint64_t &key,
                           uint32_t &total) {
    sig = 0;
    key = 0;
    total = 0;
    if (fr.n == 0u || fr.piece == PIECE_NONE) return RT_NONE;
    if (fr.mixed) return RT_GENERAL;
    if (fr.n == 1u) {
        total = (uint32_t)(fr.first.y - fr.first.x);
        if (P.flags & SCAN_NO_SINGLE_SIG) return RT_GENERAL;
        return route_single(P.T, fr.first, fr.raw, sig);
    }
    if (fr.n == 2u && fr.piece2 != PIECE_NONE && !(P.flags & SCAN_NO_PAIR_SIG)) {
        total = (uint32_t)(fr.first.y - fr.first.x) + (uint32_t)(fr.second.y - fr.second.x);
        return route_pair(P.T, fr.first, fr.second, fr.piece, fr.raw, fr.piece2, pinfo(fr.piece2), key);
    }
    return RT_GENERAL;
}

// A read with exactly two raw blocks on one chromosome, as the scan kernel decides it without the interval-list
// machinery: p / raw = piece and PieceInfo of each block's start (PIECE_NONE: none), kept = the block lies inside
// one covered region (count.cpp:319).  Same outcome as filter_read_t + route_read: two kept blocks in order form a
// pair, abutting ones fuse (interval_list.hpp:480-500), one kept block is a single-block read; overlapping or
// out-of-order blocks go to the general path.  Nil blocks are never kept and never registered as intervals.
LSQ_HD uint32_t route_double(const ScanParams &P, int2 b0, int2 b1, uint32_t p0, uint4 raw0, bool kept0, uint32_t p1,
                             uint4 raw1, bool kept1, uint32_t &sig, uint64_t &key, uint32_t &total) {
    sig = 0;
    key = 0;
    total = 0;
    if (kept0 && kept1) {
        if (b1.x > b0.y) {
            if (!(raw0.z & PIECE_CAND)) return RT_NONE;
            total = (uint32_t)(b0.y - b0.x) + (uint32_t)(b1.y - b1.x);
            if (P.flags & SCAN_NO_PAIR_SIG) return RT_GENERAL;
            return route_pair(P.T, b0, b1, p0, raw0, p1, raw1, key);
        }
        if (b1.x != b0.y) return RT_GENERAL;  // overlapping or out of order: general list update
        b0.y = b1.y;                          // the two fuse into one interval
    } else if (kept1) {
        b0 = b1;
        raw0 = raw1;
    } else if (!kept0) {
        return RT_NONE;
    }
    if (!(raw0.z & PIECE_CAND)) return RT_NONE;
    total = (uint32_t)(b0.y - b0.x);
    if (P.flags & SCAN_NO_SINGLE_SIG) return RT_GENERAL;
    return route_single(P.T, b0, raw0, sig);
}

// kept: the lookup answer names a piece whose covered region reaches the block end
LSQ_HD bool single_kept(uint32_t res) { return (res & PIECE_NONE) != PIECE_NONE && (res >> 31) != 0u; }

// The scan kernel's decision for read r, replayed sequentially (tests/emul): single-block and two-block reads are
// routed inline, everything else is left to the general kernel.
LSQ_HD uint32_t scan_route(const ScanParams &P, uint64_t r, uint32_t base, uint32_t nraw, uint32_t &sig, uint64_t &key,
                           uint32_t &total) {
    sig = 0;
    key = 0;
    total = 0;
    if (nraw == 0u) return RT_NONE;
    const uint32_t cs0 = load_read_cs(P, r);
    if (nraw > 3u || (cs0 & 0x80000000u)) return RT_GENERAL;
    int2 b[3];
    uint32_t p[3];
    uint4 raw[3];
    bool kept[3];
    for (uint32_t j = 0; j < nraw; ++j) {
        b[j] = load_block_stream(P, base + j);
        const uint32_t res = block_lookup(P, b[j], cs0);
        kept[j] = single_kept(res);
        p[j] = res & PIECE_NONE;
        raw[j] = p[j] != PIECE_NONE ? LSQ_LD(reinterpret_cast<const uint4 *>(P.T.pinfo + p[j])) : make_uint4(0u, 0u, 0u, 0u);
    }
    if (nraw == 1u) {
        if (!kept[0] || !(raw[0].z & PIECE_CAND)) return RT_NONE;
        total = (uint32_t)(b[0].y - b[0].x);
        if (P.flags & SCAN_NO_SINGLE_SIG) return RT_GENERAL;
        return route_single(P.T, b[0], raw[0], sig);
    }
    if (nraw == 2u) return route_double(P, b[0], b[1], p[0], raw[0], kept[0], p[1], raw[1], kept[1], sig, key, total);
    // three raw blocks: the signature route needs all three kept, in order, apart; anything else is general
    if (!(kept[0] && kept[1] && kept[2] && b[1].x > b[0].y && b[2].x > b[1].y) || (P.flags & SCAN_NO_TRIPLE_SIG)) return RT_GENERAL;
    if (!(raw[0].z & PIECE_CAND)) return RT_NONE;
    total = (uint32_t)(b[0].y - b[0].x) + (uint32_t)(b[1].y - b[1].x) + (uint32_t)(b[2].y - b[2].x);
    uint32_t k2;
    const uint32_t rt = route_triple(P.T, b[0], b[1], b[2], p[0], raw[0], p[1], raw[1], raw[2], key, k2);
    sig = (p[2] << 4) | k2;  // the caller keys the triple by (prefix = key, piece 2, k2)
    return rt;
}

// ---------------------------------------------------------------------------------------------
// Per (read, event) pair
// ---------------------------------------------------------------------------------------------
struct ReadView {
    int2 b0, b1, b2;  // kept blocks 0..2 held in registers (reads with more blocks read the rest from blk)
    uint32_t n;       // kept blocks
    uint32_t ridx;    // original read index (names the read: "read-<line_num[ridx]>")
    const int2 *blk;  // kept blocks 0..n-1 (reads with more than two kept blocks: the read's ext entry, which the
                      // same thread wrote in this kernel -- so plain loads, not the read-only path)
    LSQ_HD int2 block(uint32_t j) const { return j == 0 ? b0 : (j == 1 ? b1 : (j == 2 ? b2 : blk[j])); }
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

LSQ_HD bool tie_skipped(const ScanParams &P, const EvHdr &h, const ReadView &rv, uint32_t strand) {
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
LSQ_HD uint64_t eval_pair(const ScanParams &P, const EvHdr &h, const int2 *seg,
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

// ---------------------------------------------------------------------------------------------
// Signature expansion: representative reads.  Any read of the signature gives the same full-match
// outcome for every candidate event; the representative starts on the first position of the piece that
// cannot trigger the tie-break and ends exactly on the cut that closes its category.
// ---------------------------------------------------------------------------------------------
LSQ_HD bool sig_rep_single(const Tables &T, uint32_t sig, ReadView &rv, uint32_t &piece) {
    piece = LSQ_LD(T.sig_piece + sig);
    const uint4 raw = LSQ_LD(reinterpret_cast<const uint4 *>(T.pinfo + piece));
    const int32_t s = (int32_t)raw.x + ((raw.z & PIECE_TIE) ? 1 : 0), q = LSQ_LD(T.sig_cut + sig);
    rv.b0 = make_int2(s, q);
    rv.b1 = rv.b2 = make_int2(0, 0);
    rv.n = 1;
    rv.ridx = 0;
    rv.blk = nullptr;
    return s < q;
}
LSQ_HD bool sig_rep_pair(const Tables &T, uint64_t key, ReadView &rv, uint32_t &piece) {
    const uint32_t p0 = (uint32_t)key & PIECE_SIG_MASK, k0 = (uint32_t)(key >> 26) & 15u;
    const uint32_t p1 = (uint32_t)(key >> 30) & PIECE_SIG_MASK, k1 = (uint32_t)(key >> 56) & 15u;
    piece = p0;
    const uint4 raw0 = LSQ_LD(reinterpret_cast<const uint4 *>(T.pinfo + p0));
    const uint4 raw1 = LSQ_LD(reinterpret_cast<const uint4 *>(T.pinfo + p1));
    const int32_t s0 = (int32_t)raw0.x + ((raw0.z & PIECE_TIE) ? 1 : 0);
    rv.b0 = make_int2(s0, LSQ_LD(T.sig_cut + (raw0.z & PIECE_SIG_MASK) + k0));
    rv.b1 = make_int2((int32_t)raw1.x, LSQ_LD(T.sig_cut + (raw1.z & PIECE_SIG_MASK) + k1));
    rv.b2 = make_int2(0, 0);
    rv.n = 2;
    rv.ridx = 0;
    rv.blk = nullptr;
    return rv.b0.x < rv.b0.y && rv.b0.y < rv.b1.x && rv.b1.x < rv.b1.y;
}

// Representative of a three-block signature: `prefix` is the pair key its prefix slot holds.
LSQ_HD bool sig_rep_triple(const Tables &T, uint64_t prefix, uint64_t key, ReadView &rv, uint32_t &piece) {
    if (!sig_rep_pair(T, prefix, rv, piece)) return false;
    const uint32_t p2 = (uint32_t)(key >> 30) & PIECE_SIG_MASK, k2 = (uint32_t)(key >> 56) & 15u;
    const uint4 raw2 = LSQ_LD(reinterpret_cast<const uint4 *>(T.pinfo + p2));
    rv.b2 = make_int2((int32_t)raw2.x, LSQ_LD(T.sig_cut + (raw2.z & PIECE_SIG_MASK) + k2));
    rv.n = 3;
    return rv.b1.y < rv.b2.x && rv.b2.x < rv.b2.y;
}

// Every candidate event of `piece` that the representative read matches FULLY: sink(event id, header, form mask).
template <class Sink>
LSQ_HD void expand_signature(const ScanParams &P, const ReadView &rv, uint32_t piece, Sink &&sink) {
    const Tables &T = P.T;
    const uint32_t total = (uint32_t)(rv.b0.y - rv.b0.x) + (rv.n > 1u ? (uint32_t)(rv.b1.y - rv.b1.x) : 0u) +
                           (rv.n > 2u ? (uint32_t)(rv.b2.y - rv.b2.x) : 0u);
    const uint32_t e0 = LSQ_LD(T.stab_off + piece), e1 = LSQ_LD(T.stab_off + piece + 1);
    LSQ_NO_UNROLL
    for (uint32_t q = e0; q < e1; ++q) {
        const uint32_t x = LSQ_LD(&T.stab[q].x);
        const EvHdr h = T.hdr[x];
        uint32_t ml;
        const uint64_t cm = eval_pair(P, h, T.seg + h.seg_off, T.fmask + h.form_off, rv, 0u, total, ml);
        if (cm != 0ull && ml == total) sink(x, h, cm);
    }
}

// ---------------------------------------------------------------------------------------------
// Order-exact EM (common/read.h:592-660).  The solve kernels sum over compatibility CLASSES; the reference sums
// over READS in index order (count.cpp:64-85: chromosome, start, end, strand string, read name), so their
// log-likelihoods differ in the last bits.  That never matters -- except when the stop rule
// |1 - ll_old / ll| > 1e-6 is decided within that noise: near the threshold, or when ll itself is rounding noise
// (a boundary solution with all G_k = 1).  The solve kernels flag such events (em_stop_is_fragile); their valid
// reads are then collected, put in the reference's order and the EM is replayed read by read with the
// reference's own operation order, which makes theta bit-identical to the reference.
// ---------------------------------------------------------------------------------------------
struct ExactRec {
    uint64_t k1;    // read set << 32 | first kept start (biased); after sorting bits 40-51 hold the read's class index
    uint64_t k2;    // last kept end (biased) << 8 | rank of the read's strand string among the file's strand strings
    uint64_t k3;    // name_sort_key of "read-<line_num>"
    uint64_t mask;  // forms the read is compatible with
};

// "read-<a>" < "read-<b>" as byte strings: compare the decimal digit strings.  Left-aligned to ten digits the
// numbers compare like the strings; a proper prefix ("1" vs "10") is the smaller one.
LSQ_HD uint64_t name_sort_key(uint32_t n) {
    uint32_t nd = 1;
    uint64_t p = 10;
    while (nd < 10u && (uint64_t)n >= p) {
        p *= 10;
        ++nd;
    }
    uint64_t v = n;
    for (uint32_t i = nd; i < 10u; ++i) v *= 10;
    return (v << 4) | nd;
}
LSQ_HD ExactRec make_exact_rec(uint32_t m, int32_t start, int32_t end, uint32_t strand_rank, uint32_t line_num, uint64_t mask) {
    ExactRec e;
    e.k1 = ((uint64_t)m << 32) | (uint64_t)((uint32_t)start ^ 0x80000000u);
    e.k2 = ((uint64_t)((uint32_t)end ^ 0x80000000u) << 8) | (uint64_t)(strand_rank & 0xFFu);
    e.k3 = name_sort_key(line_num);
    e.mask = mask;
    return e;
}
LSQ_HD uint32_t exact_rec_set(const ExactRec &r) { return (uint32_t)(r.k1 >> 32) & 0xFFu; }
LSQ_HD bool exact_rec_less(const ExactRec &a, const ExactRec &b) {
    if (a.k1 != b.k1) return a.k1 < b.k1;
    if (a.k2 != b.k2) return a.k2 < b.k2;
    return a.k3 < b.k3;
}
// In-place heapsort (no recursion, no extra memory): read names are unique, so the order is total.
LSQ_HD void exact_sort(ExactRec *a, uint32_t n) {
    if (n < 2u) return;
    auto sift = [&](uint32_t root, uint32_t end) {
        const ExactRec v = a[root];
        while (true) {
            uint32_t child = 2u * root + 1u;
            if (child >= end) break;
            if (child + 1u < end && exact_rec_less(a[child], a[child + 1u])) ++child;
            if (!exact_rec_less(v, a[child])) break;
            a[root] = a[child];
            root = child;
        }
        a[root] = v;
    };
    for (uint32_t i = n / 2u; i-- > 0u;) sift(i, n);
    for (uint32_t end = n - 1u; end > 0u; --end) {
        const ExactRec t = a[0];
        a[0] = a[end];
        a[end] = t;
        sift(0u, end);
    }
}

// Is the EM's stop decision at distance `near` from the threshold (min over the iterations of
// | |1 - ll_old/ll| - 1e-6 |), with final log-likelihood ll over n reads, safe against summation-order noise?
LSQ_HD bool em_stop_is_fragile(double near, double ll, double n_reads) {
    const double margin = 1E-12 > 4.5E-16 * n_reads ? 1E-12 : 4.5E-16 * n_reads;
    return near < margin || (ll > -1E-8 && ll < 1E-8);
}

// ---------------------------------------------------------------------------------------------
// A stop that em_stop_is_fragile cannot vouch for may still come out the same in EVERY summation order.  The common
// case: forms not longer than a read (G = 1, or G = 0 for a form no read can start in) drive theta to a vertex in
// one step, every read's probability becomes exactly 1, the log-likelihood exactly 0, and the stop test divides by
// zero -- "fragile" by value, yet nothing in that run ever rounded.  em_result_is_order_free re-runs the
// class-collapsed EM of such an event (they have few iterations) and certifies that
//   * every E-step sum is exact: all its per-read terms are multiples of one power of two, 2^low, and their total
//     stays below 2^(low + 52), so every partial sum in any order is representable -- the reference's sum and ours
//     are the same number, and so is the next theta (one correctly rounded division);
//   * every stop decision is the same in any order: a log-likelihood whose terms are all exactly 0 is exactly 0
//     however it is added up; one that is not must be clearly non-zero, and the test value must keep the margin of
//     em_stop_is_fragile from the threshold unless the division by an exact zero already decides it.
// Certified events need no read-by-read replay.  visit(f) calls f(read set, form mask, reads) for every class with
// reads; g(m, k) is G of form k for read set m.  Events with more than CERT_MAX_FORMS forms or CERT_MAX_ITERS
// iterations are not examined.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t CERT_MAX_FORMS = 8, CERT_MAX_ITERS = 8;

// exponent of the lowest set bit of x != 0 (x = odd * 2^result)
LSQ_HD int low_bit_exponent(double x) {
    int e;
    uint64_t mant = (uint64_t)ldexp(frexp(x < 0.0 ? -x : x, &e), 53);
    int tz = 0;
    while (!(mant & 1ull)) {
        mant >>= 1;
        ++tz;
    }
    return e - 53 + tz;
}

template <class Visit, class GFn>
LSQ_HD bool em_result_is_order_free(uint32_t K, double n_total, uint32_t iters_done, uint32_t max_iters, Visit visit, GFn g) {
    if (K < 2u || K > CERT_MAX_FORMS || iters_done == 0u || iters_done > CERT_MAX_ITERS) return false;
    double th[CERT_MAX_FORMS], z[CERT_MAX_FORMS];
    int low[CERT_MAX_FORMS];
    for (uint32_t k = 0; k < K; ++k) th[k] = lsq_div(1.0, (double)K);
    auto prob = [&](uint32_t m, uint64_t mk) {
        double sc = 0.0;
        for (uint32_t k = 0; k < K; ++k)
            if ((mk >> k) & 1ull) sc = lsq_add(sc, lsq_mul(th[k], g(m, k)));
        return sc;
    };
    // log-likelihood at th; zero: every term is exactly 0
    auto loglik = [&](bool &zero) {
        double ll = 0.0;
        bool all_zero = true;
        visit([&](uint32_t m, uint64_t mk, double n_c) {
            const double L = log(prob(m, mk));
            if (L != 0.0) all_zero = false;
            ll = lsq_add(ll, lsq_mul(n_c, L));
        });
        zero = all_zero;
        return ll;
    };
    const double margin = 1E-12 > 4.5E-16 * n_total ? 1E-12 : 4.5E-16 * n_total;
    bool old_zero;
    double ll_old = loglik(old_zero);
    if (!old_zero && !(ll_old < -1E-8 && ll_old > -1E300)) return false;
    for (uint32_t it = 1;; ++it) {
        for (uint32_t k = 0; k < K; ++k) {
            z[k] = 0.0;
            low[k] = 4096;
        }
        visit([&](uint32_t m, uint64_t mk, double n_c) {
            const double sc = prob(m, mk);
            if (!(sc > 0.0)) return;
            for (uint32_t k = 0; k < K; ++k) {
                if (!((mk >> k) & 1ull)) continue;
                const double lp = lsq_mul(th[k], g(m, k));
                if (!(lp > 0.0)) continue;
                const double t = lsq_div(lp, sc);
                const int lb = low_bit_exponent(t);
                if (lb < low[k]) low[k] = lb;
                z[k] = lsq_add(z[k], lsq_mul(n_c, t));
            }
        });
        for (uint32_t k = 0; k < K; ++k) {
            if (z[k] != 0.0 && !(z[k] < ldexp(1.0, low[k] + 52))) return false;  // some partial sum may round
            th[k] = lsq_div(z[k], n_total);
        }
        bool zero;
        const double ll = loglik(zero);
        bool stop;
        if (zero) {
            stop = old_zero;  // 0 / 0: the test is false; x / 0 with x clearly non-zero: infinite, the test is true
        } else {
            if (!(ll < -1E-8 && ll > -1E300)) return false;
            if (old_zero) {
                stop = false;  // |1 - 0 / ll| = 1
            } else {
                const double d = 1.0 - lsq_div(ll_old, ll), tk = d < 0.0 ? -d : d, off = tk - 1E-6;
                if ((off < 0.0 ? -off : off) < margin) return false;
                stop = !(tk > 1E-6);
            }
        }
        if (stop || it >= max_iters) return it == iters_done;
        if (it >= iters_done) return false;  // the kernel stopped earlier than this run would: leave it to the replay
        ll_old = ll;
        old_zero = zero;
    }
}

// The reference evaluates, per read, s = sum_k theta_k M(i,k) (every form's term added in form order, zeros included),
// log s for the log-likelihood (read.h:620-636) and theta_k M(i,k) / s for the E-step (read.h:592-618), and adds them up
// read by read.  Those VALUES depend only on the read's class (read set, form mask): they are tabulated per class,
// and only the additions run over the reads, in the reference's order -- same values, same order, same bits.
constexpr uint32_t EXACT_MAX_CLS = 64, EXACT_MAX_TERMS = 1024;  // classes, classes * forms of the tabulated path

LSQ_HD double exact_class_prob(uint64_t mask, uint32_t K, const double *theta, const double *g) {
    double s = 0.0;
    for (uint32_t k = 0; k < K; ++k) s = lsq_add(s, lsq_mul(theta[k], ((mask >> k) & 1ull) ? g[k] : 0.0));
    return s;
}

// ---------------------------------------------------------------------------------------------
// The running sum  s <- fl(s + term[cls[i]]),  i = 0 .. n-1,  bit for bit, without n dependent additions.
// All terms have one sign and take C <= 64 distinct values.  While s stays inside one binade [2^e, 2^(e+1)) it is an
// integer multiple S of u = ulp(s), and fl(s + t) = (S + inc_t) u with the INTEGER inc_t = t/u rounded to nearest
// (a term that is an exact half-way case rounds by the parity of S: such binades, rare and short, are added up one
// by one).  So inside a binade the sum is integer arithmetic, order does not matter, and the only question is where
// the binade ends: the first read at which S + (sum of inc) reaches 2^53.  A table of per-class read counts ahead of
// every chunk of ORD_CHUNK reads answers that with a binary search over the chunks plus a walk through one chunk;
// the addition that crosses into the next binade is done in floating point, and the next binade starts.  A sum over
// n reads costs O(binades * (C log(n / ORD_CHUNK) + ORD_CHUNK)) instead of n.
// pre[j * C + c]: reads of class c in chunks 0 .. j-1 (j = 0 .. n_chunks).  term[c * stride]: the term of class c; the
// result is the sum of the MAGNITUDES (the caller gives it the terms' common sign).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t ORD_CHUNK = 16;  // the walk through a chunk is the serial part: short chunks, a deeper search
constexpr uint32_t ORD_MIN_N = 4096;  // shorter sequences are summed one by one

LSQ_HD uint64_t ord_sat_mul(uint64_t inc, uint64_t cnt, uint64_t cap) {  // min(inc * cnt, cap)
#ifdef __CUDA_ARCH__
    if (__umul64hi(inc, cnt) != 0ull) return cap;
    const uint64_t lo = inc * cnt;
#else
    const unsigned __int128 w = (unsigned __int128)inc * cnt;
    if (w >> 64) return cap;
    const uint64_t lo = (uint64_t)w;
#endif
    return lo > cap ? cap : lo;
}

LSQ_HD uint32_t ord_min(uint32_t a, uint32_t b) { return a < b ? a : b; }

constexpr uint32_t ORD_MAX_GUESS = 96;  // binades of one sum whose end is remembered from the previous sweep

// guess (optional, ORD_MAX_GUESS entries, zero-initialised): where each binade ended in the previous sweep of the same
// running sum -- theta moves little between EM iterations, so the search usually ends after one look around that place.
// On the device cls must be 16-byte aligned (a chunk of classes is one vector load).
LSQ_HD double ordered_sum(const uint8_t *cls, uint32_t n, const double *term, uint32_t stride, uint32_t C, const uint32_t *pre,
                          uint32_t *guess = nullptr) {
    auto mag = [&](uint32_t c) { return fabs(term[c * stride]); };
    double s = 0.0;
    uint32_t i = 0, n_searches = 0;
    const uint32_t n_chunks = (n + ORD_CHUNK - 1u) / ORD_CHUNK;
    uint64_t inc[EXACT_MAX_CLS];
    while (i < n) {
        // While s is 0 every zero term leaves it 0 and the first other term is taken exactly (0 + t): the same search with
        // "increment 1 for a class whose term is not zero" and room for none finds that read (a form whose theta has
        // gone to 0 has nothing but zero terms: no read is visited at all).
        const bool zero_mode = s == 0.0;
        int e = 0;  // s = m 2^e, 0.5 <= m < 1: the binade is [2^(e-1), 2^e), u = 2^(e-53)
        if (!zero_mode) (void)frexp(s, &e);
        uint64_t S = zero_mode ? 0ull : (uint64_t)ldexp(s, 53 - e);
        const uint64_t rem = zero_mode ? 1ull : (1ull << 53) - S;  // the sum leaves the binade when the increments reach rem
        bool tie = false;
        for (uint32_t c = 0; c < C; ++c) {
            if (zero_mode) {
                inc[c] = mag(c) != 0.0 ? 1ull : 0ull;
                continue;
            }
            const double sc = ldexp(mag(c), 53 - e);  // t / u, exact (or underflows towards 0: far below 1/2)
            if (!(sc < 9007199254740992.0)) {
                inc[c] = rem;  // t alone leaves the binade
            } else {
                const double q = floor(sc), r = sc - q;
                inc[c] = (uint64_t)q + (r > 0.5 ? 1ull : 0ull);
                if (r == 0.5) tie = true;
                if (inc[c] > rem) inc[c] = rem;
            }
        }
        const double top = ldexp(1.0, e);
        if (!zero_mode && (tie || e < -900)) {  // one by one until the binade is left (also near the subnormal range, where u is not 2^(e-53))
            while (i < n && s < top) {
                s = lsq_add(s, mag(cls[i]));
                ++i;
            }
            continue;
        }
        uint64_t acc = 0;  // increments taken inside this binade, < rem
        bool crossed = false;
        // read by read up to `end` (at most the end of i's chunk); stops before the read that would cross
        auto walk = [&](uint32_t end) {
#ifdef __CUDA_ARCH__
            // the chunk's 16 classes in one load, their increments fetched together, then the running test in registers
            const uint32_t c0 = i & ~(ORD_CHUNK - 1u);
            const uint4 v = *reinterpret_cast<const uint4 *>(cls + c0);  // the buffer is padded to a whole chunk
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            uint64_t d[ORD_CHUNK];
#pragma unroll
            for (uint32_t j = 0; j < ORD_CHUNK; ++j) d[j] = inc[(w[j >> 2] >> (8u * (j & 3u))) & (EXACT_MAX_CLS - 1u)];
#pragma unroll
            for (uint32_t j = 0; j < ORD_CHUNK; ++j) {
                if (c0 + j < i || c0 + j >= end || crossed) continue;
                if (d[j] >= rem - acc) {
                    crossed = true;
                } else {
                    acc += d[j];
                    ++i;
                }
            }
#else
            while (i < end) {
                const uint64_t d = inc[cls[i]];
                if (d >= rem - acc) {
                    crossed = true;
                    return;
                }
                acc += d;
                ++i;
            }
#endif
        };
        walk(ord_min(n, (i / ORD_CHUNK + 1u) * ORD_CHUNK));  // to the end of the current chunk
        if (!crossed && i < n) {
            // i is a chunk boundary j0: the last boundary j whose reads [j0, j) all stay inside the binade
            const uint32_t j0 = i / ORD_CHUNK;
            const uint64_t room = rem - acc;
            auto span = [&](uint32_t j) {  // min(sum of increments of chunks [j0, j), room)
                uint64_t g = 0;
                for (uint32_t c = 0; c < C && g < room; ++c) {
                    const uint32_t cnt = LSQ_LD(pre + (size_t)j * C + c) - LSQ_LD(pre + (size_t)j0 * C + c);
                    if (cnt) g += ord_sat_mul(inc[c], cnt, room);
                }
                return g < room ? g : room;
            };
            // invariant: span(lo) < room; hi_known: span(hi) >= room.  The answer is the largest boundary below room.
            uint32_t lo = j0, hi = n_chunks;
            bool hi_known = false;
            const uint32_t gs = (guess != nullptr && n_searches < ORD_MAX_GUESS) ? guess[n_searches] : 0u;
#ifdef __CUDA_ARCH__
            // on the device a whole warp runs this function on the same arguments (everything above is uniform): the
            // lanes probe 32 boundaries at once -- first a window around the previous sweep's answer, then 32 evenly
            // spaced ones per round -- so the search takes one or two round trips to L2 instead of log2(n / 16)
            const uint32_t lane = threadIdx.x & 31u;
            auto clampb = [&](uint32_t b) { return b <= lo ? lo + 1u : (b > hi ? hi : b); };
            bool window = gs > lo && gs <= hi;
            while (lo < hi && !(hi_known && hi - lo == 1u)) {
                const uint32_t step = (hi - lo + 31u) / 32u;
                const uint32_t first = gs > lo + 8u ? gs - 8u : lo + 1u;
                const uint32_t mine = window ? clampb(first + lane) : ord_min(hi, lo + (lane + 1u) * step);
                const bool below = span(mine) < room;  // true for a prefix of the lanes: the boundaries do not decrease
                const uint32_t nb = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, below));
                const uint32_t b_last_below = __shfl_sync(0xFFFFFFFFu, mine, nb ? (int)nb - 1 : 0);
                const uint32_t b_first_above = __shfl_sync(0xFFFFFFFFu, mine, nb < 32u ? (int)nb : 31);
                if (nb) lo = b_last_below;
                if (nb < 32u) {
                    hi = b_first_above;
                    hi_known = true;
                }  // nb == 32: lo is now lane 31's boundary -- hi itself in an evenly spaced round, which ends the search
                window = false;
            }
#else
            if (gs > lo && gs <= hi) {  // the previous sweep's answer and its successor first
                if (span(gs) < room) {
                    lo = gs;
                    if (gs < hi && !(span(gs + 1u) < room)) {
                        hi = gs + 1u;
                        hi_known = true;
                    }
                } else {
                    hi = gs;
                    hi_known = true;
                }
            }
            if (!hi_known) {
                if (span(hi) < room) lo = hi;
                else hi_known = true;
            }
            while (hi - lo > 1u) {
                const uint32_t mid = lo + (hi - lo) / 2u;
                if (span(mid) < room) lo = mid;
                else hi = mid;
            }
#endif
            if (guess != nullptr && n_searches < ORD_MAX_GUESS) guess[n_searches] = lo;
            ++n_searches;
            acc += span(lo);
            i = ord_min(n, lo * ORD_CHUNK);
            if (i < n) walk(ord_min(n, (lo + 1u) * ORD_CHUNK));  // the crossing read is in this chunk
        }
        if (!zero_mode) s = ldexp((double)(S + acc), e - 53);  // exact: S + acc < 2^53
        if (crossed) {
            s = lsq_add(s, mag(cls[i]));
            ++i;
        }
    }
    return s;
}

// compute_iso_probs_EM (read.h:638-660) over the sorted records, every term evaluated per read: for events with more
// classes than the tables below hold.  theta (K doubles) receives the result, old_theta is scratch.  Returns the
// iterations executed; ll_out = the final log-likelihood.
LSQ_HD uint32_t exact_em_per_read(const ExactRec *rec, uint32_t n, double n_total, uint32_t K, const double *G, size_t g_stride,
                                  double *theta, double *old_theta, double &ll_out, uint32_t max_iters) {
    auto loglik = [&](const double *th) {
        double ll = 0.0;
        for (uint32_t i = 0; i < n; ++i)
            ll = lsq_add(ll, log(exact_class_prob(rec[i].mask, K, th, G + (size_t)exact_rec_set(rec[i]) * g_stride)));
        return ll;
    };
    for (uint32_t k = 0; k < K; ++k) theta[k] = lsq_div(1.0, (double)K);
    uint32_t iters = 0;
    double ll = loglik(theta);
    while (true) {
        for (uint32_t k = 0; k < K; ++k) old_theta[k] = theta[k];
        const double old_ll = ll;
        for (uint32_t k = 0; k < K; ++k) {
            double sum_zeta = 0.0;
            for (uint32_t i = 0; i < n; ++i) {
                const double *g = G + (size_t)exact_rec_set(rec[i]) * g_stride;
                const double sc = exact_class_prob(rec[i].mask, K, old_theta, g);
                if (sc > 0.0) {
                    const double lp = lsq_mul(old_theta[k], ((rec[i].mask >> k) & 1ull) ? g[k] : 0.0);
                    if (lp > 0.0) sum_zeta = lsq_add(sum_zeta, lsq_div(lp, sc));
                }
            }
            theta[k] = lsq_div(sum_zeta, n_total);
        }
        ll = loglik(theta);
        ++iters;
        const double d = 1.0 - lsq_div(old_ll, ll);
        if (!((d < 0.0 ? -d : d) > 1E-6) || iters >= max_iters) break;
    }
    ll_out = ll;
    return iters;
}

// The same with the terms tabulated per class (one thread; lsq_exact_kernel runs this computation with one thread per
// running sum, the CPU replay of tests/emul runs it as written).
LSQ_HD uint32_t exact_em(ExactRec *rec, uint32_t n, double n_total, uint32_t K, const double *G, size_t g_stride,
                         double *theta, double *old_theta, double &ll_out, uint32_t max_iters) {
    // classes of the event: distinct (read set, mask) pairs; a read's class index goes into bits 40-51 of k1
    uint64_t cmask[EXACT_MAX_CLS];
    uint32_t cset[EXACT_MAX_CLS];
    uint32_t C = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t m = exact_rec_set(rec[i]);
        uint32_t c = 0;
        while (c < C && !(cmask[c] == rec[i].mask && cset[c] == m)) ++c;
        if (c == C) {
            if (C == EXACT_MAX_CLS || (C + 1u) * (K + 1u) > EXACT_MAX_TERMS + EXACT_MAX_CLS)
                return exact_em_per_read(rec, n, n_total, K, G, g_stride, theta, old_theta, ll_out, max_iters);
            cmask[C] = rec[i].mask;
            cset[C] = m;
            ++C;
        }
        rec[i].k1 = (rec[i].k1 & 0xFFFF00FFFFFFFFFFull) | ((uint64_t)c << 40);
    }
    double term[EXACT_MAX_TERMS + EXACT_MAX_CLS];  // [class][form] E-step terms
    double lterm[EXACT_MAX_CLS];                   // log-likelihood term per class
    auto loglik = [&](const double *th) {
        double ll = 0.0;
        for (uint32_t c = 0; c < C; ++c) lterm[c] = log(exact_class_prob(cmask[c], K, th, G + (size_t)cset[c] * g_stride));
        for (uint32_t i = 0; i < n; ++i) ll = lsq_add(ll, lterm[(rec[i].k1 >> 40) & 0xFFu]);
        return ll;
    };
    for (uint32_t k = 0; k < K; ++k) theta[k] = lsq_div(1.0, (double)K);
    uint32_t iters = 0;
    double ll = loglik(theta);  // the reference re-evaluates old_ll every iteration: the same theta in the same order gives the same bits
    while (true) {
        for (uint32_t k = 0; k < K; ++k) old_theta[k] = theta[k];
        const double old_ll = ll;
        // compute_iso_probs_EM_step, read.h:592-618: a skipped term (s <= 0 or lp <= 0) adds nothing, like + 0.0
        for (uint32_t c = 0; c < C; ++c) {
            const double *g = G + (size_t)cset[c] * g_stride;
            const double sc = exact_class_prob(cmask[c], K, old_theta, g);
            for (uint32_t k = 0; k < K; ++k) {
                const double lp = lsq_mul(old_theta[k], ((cmask[c] >> k) & 1ull) ? g[k] : 0.0);
                term[c * K + k] = (sc > 0.0 && lp > 0.0) ? lsq_div(lp, sc) : 0.0;
            }
        }
        for (uint32_t k = 0; k < K; ++k) theta[k] = 0.0;
        for (uint32_t i = 0; i < n; ++i) {  // K independent running sums, each in read order
            const double *t = term + ((rec[i].k1 >> 40) & 0xFFu) * K;
            for (uint32_t k = 0; k < K; ++k) theta[k] = lsq_add(theta[k], t[k]);
        }
        for (uint32_t k = 0; k < K; ++k) theta[k] = lsq_div(theta[k], n_total);
        ll = loglik(theta);
        ++iters;
        const double d = 1.0 - lsq_div(old_ll, ll);
        if (!((d < 0.0 ? -d : d) > 1E-6) || iters >= max_iters) break;
    }
    ll_out = ll;
    return iters;
}

}  // namespace lsq
