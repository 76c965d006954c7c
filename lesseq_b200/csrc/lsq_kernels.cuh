// Hand-written sm_100a kernels of the count + solve hot path.
//
//   lsq_filter_kernel  per-block covered-region filter + per-read interval list + stab piece key,
//                      partition of the kept reads by (read class, stab bucket)
//                      (count/count.cpp:319-324; jsc/util/interval_list.hpp:396-422, 462-503;
//                       read index of count.cpp:346-364)
//   lsq_count_kernel   candidate scan, Read_single::build walk, compatibility, validity, class
//                      histogram (count.cpp:421-482; common/read.h:44-79, 204-274)
//   lsq_finalize_kernel per-event n_valid and per-form tallies from the class histogram
//                      (count.cpp:467-482)
//   lsq_solve_kernel   EM on the class histogram + RPKM + log-likelihood, FP64 without FMA
//                      contraction (common/read.h:592-660; solve/solve.cpp:796-826)
#pragma once
#include "lsq_device_fn.cuh"
#include "lsq_tma.cuh"

namespace lsq {

// ---------------------------------------------------------------------------------------------
// Filter + partition kernel: one CTA per tile of FILT_TILE consecutive reads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

#ifndef LSQ_FILT_BLOCK
#define LSQ_FILT_BLOCK 512
#endif
constexpr int FILT_BLOCK = LSQ_FILT_BLOCK;  // 1024 threads per SM either way (64 registers per thread)
constexpr int FILT_ITEMS = 4;                         // reads per thread (all kept in flight together)
#ifndef LSQ_FILT_LBATCH
#define LSQ_FILT_LBATCH 2
#endif
constexpr int FILT_LBATCH = LSQ_FILT_LBATCH;                        // block lookups of multi-block reads kept in flight per thread
#ifndef LSQ_FILT_RBATCH
#define LSQ_FILT_RBATCH 3
#endif
constexpr int FILT_RBATCH = LSQ_FILT_RBATCH;                        // bin reservations kept in flight per thread
constexpr int FILT_TILE = FILT_BLOCK * FILT_ITEMS;    // 2048 reads
constexpr int FILT_WREADS = 32 * FILT_ITEMS;          // 128 reads per warp
constexpr int FILT_WBLK = 320;                        // raw blocks of a warp's reads whose lookups are staged
constexpr int FILT_MAX_BLK = FILT_WBLK * (FILT_BLOCK / 32);
static_assert((uint32_t)FILT_TILE <= PART_CHUNK, "a tile's reservation may span at most two chunks");

// dynamic shared memory of the filter kernel
__host__ __device__ inline size_t filter_smem_bytes(uint32_t n_bins) {
    return (size_t)FILT_TILE * (16 + 4 + 2) + (size_t)FILT_MAX_BLK * (4 + 2) + (size_t)n_bins * (4 + 4 + 2) + 64;  // <= 200 KB at PART_MAX_BINS
}

// Read context of the filter kernel (see filter_read_t): blocks and piece records are fetched on
// demand, lookup answers come from the tile's shared staging.
struct TileCtx {
    const FilterParams &P;
    const uint32_t *s_blk;
    uint32_t base, tbase;
    int2 b0, b1;  // the first two raw blocks, fetched by the caller
    bool staged;
    __device__ __forceinline__ int2 blk(uint32_t j) const { return j == 0u ? b0 : (j == 1u ? b1 : load_block(P, base + j)); }
    __device__ __forceinline__ uint32_t lookup(uint32_t j, uint32_t c, int2 b) const {
        return staged ? s_blk[base - tbase + j] : lookup_direct(P.T, c, b);
    }
    __device__ __forceinline__ uint4 pinfo(uint32_t p) const {
        return __ldg(reinterpret_cast<const uint4 *>(P.T.pinfo + p));
    }
};

__global__ void lsq_part_init_kernel(PartState S, uint32_t n_bkt) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.n_bins) return;
    // every bin starts with chunk b, which is also the first work item of its class
    S.state[b] = (unsigned long long)b << 32;
    S.next_chunk[b] = S.n_bins + b;
    S.owner[b] = b;
    const uint32_t cls = b / n_bkt;
    S.wl[(size_t)cls * S.pool_cap + (b - cls * n_bkt)] = b;
    if (b < 3u) S.wl_n[b] = n_bkt;
    if (b == 0u) *S.pool_next = 2u * S.n_bins;
}

// Phases A-C run per WARP on its own 128 reads of the tile, with warp-level synchronisation only,
// so the sixteen warps of a block drift apart and cover each other's memory latencies:
// Phase A: every lane loads offsets and chromosome/strand ids of its four reads.  Reads with one
//          raw block (4 in 5) are "singles"; the raw blocks of all other reads are appended to a
//          warp-local list (their ids expanded per block in shared memory).
// Phase S: singles, one lane per read, four reads in flight: block -> bucket entry (ONE 8-byte
//          gather names the piece, lsq_types.cuh:bucket_rank) -> PieceInfo (ONE 16-byte gather:
//          covered test, stab list, piece start) -> single_decide.  Two dependent gathers per
//          read, nothing staged.
// Phase B: the listed raw blocks are spread densely over the lanes regardless of the reads they
//          belong to; same two gathers per block, answers go to shared memory.
// Phase C: the multi-block reads, spread densely over the lanes, take filter_read_t (interval
//          list, read class) with the staged answers.  Records of kept reads are staged in
//          shared memory.
// Then, block-wide:
// Phase D: one thread per non-empty bin reserves room for the tile's records of that bin with a
//          single 64-bit atomicAdd on (chunk, fill); the reservation that takes the last slot of
//          a chunk installs the bin's spare chunk.
// Phase E: records go from shared memory to their slots (writes of one bin are contiguous; L2
//          merges them into full sectors before they reach HBM).
__global__ void __launch_bounds__(FILT_BLOCK, 1024 / FILT_BLOCK) lsq_filter_kernel(FilterParams P) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const Tables &T = P.T;
    const PartState &S = P.part;
    const uint32_t n_bins = S.n_bins;
    uint4 *s_rec = reinterpret_cast<uint4 *>(s_raw);  // PartRec of the tile's reads
    uint32_t *s_br = reinterpret_cast<uint32_t *>(s_rec + FILT_TILE);  // bin | rank << 16
    uint32_t *s_blk_all = s_br + FILT_TILE;          // per warp and raw block: chromosome/strand ids, then the lookup answer
    uint32_t *s_A = s_blk_all + FILT_MAX_BLK;        // per bin: records of the tile, then first slot
    uint32_t *s_B = s_A + n_bins;                    // per bin: slot base after a chunk switch
    uint16_t *s_multi_all = reinterpret_cast<uint16_t *>(s_B + n_bins);  // per warp: reads for filter_read_t (tile-local index)
    uint16_t *s_split = s_multi_all + FILT_TILE;     // per bin: records before the chunk switch
    uint16_t *s_list_all = s_split + n_bins;         // per warp: raw blocks of the multi-block reads (warp-local index)
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t i = tid; i < (n_bins + 3u) / 4u; i += FILT_BLOCK)  // may run into s_B, which is written before use
        reinterpret_cast<uint4 *>(s_A)[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint64_t tile0 = (uint64_t)blockIdx.x * FILT_TILE;
    const uint64_t tile1 = min(tile0 + (uint64_t)FILT_TILE, P.n_reads);
    if (P.progress != nullptr && tid == 0u) {
        // streamed read set: the copy engine is still filling blk[]; wait until the tile's last block
        // has landed (the per-read arrays were complete before the launch)
        const uint32_t need = __ldg(P.blk_off + tile1);
        uint32_t spins = 0;
        while (*reinterpret_cast<const volatile uint32_t *>(P.progress) < need) {
            __nanosleep(400);
            if (++spins > 100000000u) __trap();  // ~40 s without progress: the copy died; fail instead of hanging
        }
        __threadfence();
    }
    __syncthreads();
    const uint32_t n_bkt = T.n_bkt;
    // the warp's reads: [w0, w1), lane l owns w0 + l + 32 k
    const uint64_t w0 = min(tile0 + (uint64_t)warp * FILT_WREADS, tile1), w1 = min(w0 + (uint64_t)FILT_WREADS, tile1);
    uint32_t *s_blk = s_blk_all + warp * FILT_WBLK;
    uint16_t *s_multi = s_multi_all + warp * FILT_WREADS;
    uint16_t *s_list = s_list_all + warp * FILT_WBLK;

    // ---- phase A --------------------------------------------------------------------------
    uint32_t rbase[FILT_ITEMS], rn[FILT_ITEMS], rcs[FILT_ITEMS];  // absolute block offset, raw blocks, ids
#pragma unroll
    for (int k = 0; k < FILT_ITEMS; ++k) {
        const uint64_t r = w0 + lane + (uint32_t)k * 32u;
        rbase[k] = 0u;
        rn[k] = 0u;
        rcs[k] = 0u;
        if (r < w1) {
            rbase[k] = __ldcs(P.blk_off + r);
            rn[k] = __ldcs(P.blk_off + r + 1) - rbase[k];
            rcs[k] = load_read_cs(P, r);
        }
    }
    const uint32_t wbase = __shfl_sync(0xFFFFFFFFu, rbase[0], 0);
    const uint32_t w_nblk = w1 > w0 ? __ldg(P.blk_off + w1) - wbase : 0u;
    const bool staged = w_nblk <= (uint32_t)FILT_WBLK;  // warp-uniform
    bool single[FILT_ITEMS];
    uint32_t n_mb = 0;  // listed raw blocks, warp-uniform
    {
        uint32_t mb = 0;
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k) {
            single[k] = rn[k] == 1u && !(rcs[k] & 0x80000000u);
            mb += single[k] ? 0u : rn[k];
        }
        if (staged) {
            uint32_t incl = mb;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            n_mb = __shfl_sync(0xFFFFFFFFu, incl, 31);
            uint32_t pos = incl - mb;
#pragma unroll
            for (int k = 0; k < FILT_ITEMS; ++k) {
                if (single[k] || rn[k] == 0u) continue;
                const uint32_t rel = rbase[k] - wbase;
                if (rcs[k] & 0x80000000u) {  // per-block ids (chimeric read)
                    for (uint32_t j = 0; j < rn[k]; ++j) s_blk[rel + j] = __ldg(P.blk_cs + rbase[k] + j);
                } else {
                    s_blk[rel] = rcs[k];
                    if (rn[k] >= 2u) s_blk[rel + 1] = rcs[k];
                    for (uint32_t j = 2; j < rn[k]; ++j) s_blk[rel + j] = rcs[k];
                }
                s_list[pos] = (uint16_t)rel;
                if (rn[k] >= 2u) s_list[pos + 1] = (uint16_t)(rel + 1u);
                for (uint32_t j = 2; j < rn[k]; ++j) s_list[pos + j] = (uint16_t)(rel + j);
                pos += rn[k];
            }
        }
    }
    __syncwarp();
    // ---- phase B --------------------------------------------------------------------------
#pragma unroll 1
    for (uint32_t k0 = lane; k0 < n_mb; k0 += 32u * FILT_LBATCH) {
        int st[FILT_LBATCH];
        BlockProbe q[FILT_LBATCH];
        uint32_t rel[FILT_LBATCH];
#pragma unroll
        for (int u = 0; u < FILT_LBATCH; ++u) {
            const uint32_t k = k0 + (uint32_t)u * 32u;
            st[u] = k < n_mb ? SP_GO : -1;
            rel[u] = st[u] == SP_GO ? (uint32_t)s_list[k] : 0u;
        }
        int2 b[FILT_LBATCH];
#pragma unroll
        for (int u = 0; u < FILT_LBATCH; ++u) b[u] = st[u] == SP_GO ? load_block(P, wbase + rel[u]) : make_int2(0, 0);
#pragma unroll
        for (int u = 0; u < FILT_LBATCH; ++u)
            if (st[u] == SP_GO) st[u] = block_stage_chrom(P, b[u], s_blk[rel[u]], q[u]);
        uint2 be[FILT_LBATCH];
#pragma unroll
        for (int u = 0; u < FILT_LBATCH; ++u) be[u] = st[u] == SP_GO ? __ldg(T.bp_bkt + q[u].slot) : make_uint2(0u, 0u);
        uint32_t pp[FILT_LBATCH];
        int32_t ce[FILT_LBATCH];
#pragma unroll
        for (int u = 0; u < FILT_LBATCH; ++u) {
            pp[u] = PIECE_NONE;
            ce[u] = 0;
            if (st[u] == SP_GO) st[u] = block_stage_piece(T, q[u], be[u], pp[u]);
            if (st[u] == SP_SUB) st[u] = block_stage_sub(T, q[u], __ldg(T.bp_sub + q[u].slot), pp[u]);
            if (st[u] == SP_GO && pp[u] != PIECE_NONE) ce[u] = __ldg(&T.pinfo[pp[u]].cov_end);
        }
#pragma unroll
        for (int u = 0; u < FILT_LBATCH; ++u) {
            if (st[u] < 0) continue;
            uint32_t res = PIECE_NONE;
            if (st[u] == SP_GO) res = block_stage_answer(q[u], pp[u], ce[u]);
            if (st[u] == SP_DEFER) res = block_stage_defer(P, q[u]);
            s_blk[rel[u]] = res;
        }
    }
    __syncwarp();
    // ---- phase C --------------------------------------------------------------------------
    // record of a kept read -> shared staging (and the ext entry of a GENERAL read)
    auto emit = [&](uint32_t li, uint64_t r, uint32_t base, const FilteredRead &fr) {
        uint32_t br = BIN_NONE;
        if (fr.cls != CLS_DROPPED) {
            uint32_t cls = fr.cls;
            const uint32_t so = fr.stab_off, nent = fr.stab_n;
            if (so > PREC_MAX_SO || nent > PREC_MAX_N) cls = CLS_GENERAL;  // the piece does not fit the record
            uint4 rec;
            rec.x = so | (nent << PREC_SO_BITS);
            if (cls == CLS_PAIR) {
                const ReadRec pr = pack_pair(fr.first, fr.second, fr.piece_pos, fr.strand);  // two blocks inline
                rec.y = (uint32_t)fr.piece_pos;
                rec.z = pr.x;
                rec.w = pr.y;
            } else if (cls == CLS_GENERAL) {
                // the ext entry of read r lives at the even slot (blk_off[r] + 2r) & ~1: 1 header +
                // n <= 1 + raw blocks, so no allocation is needed, entries of different reads never
                // overlap, and header + first block form one aligned 128-bit access
                const uint32_t off = ext_slot(base, (uint32_t)r);
                *reinterpret_cast<int4 *>(P.ext + off) = make_int4(
                    (int32_t)(uint32_t)r, (int32_t)((fr.strand << 24) | (fr.n & 0xFFFFFFu)), fr.first.x, fr.first.y);
                if (fr.n >= 2u) P.ext[off + 2] = fr.second;  // further blocks were spilled by the filter
                rec.x = 0u;
                rec.y = fr.piece;
                rec.z = off;
                rec.w = fr.n;
            } else {
                rec.y = 0u;
                rec.z = (uint32_t)fr.first.x;
                rec.w = (uint32_t)(fr.first.y - fr.first.x) | (fr.strand << 16);
            }
            const uint32_t bin = cls * n_bkt + (so >> T.bkt_shift);
            const uint32_t rank = atomicAdd(&s_A[bin], 1u);
            s_rec[li] = rec;
            br = bin | (rank << 16);
        }
        s_br[li] = br;
    };
    // ---- phase S --------------------------------------------------------------------------
    uint32_t n_listed = 0;  // warp-uniform
    {
        int st[FILT_ITEMS];
        int2 b0[FILT_ITEMS];
        BlockProbe q[FILT_ITEMS];
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k) {
            st[k] = single[k] ? SP_GO : -1;
            b0[k] = single[k] ? load_block_stream(P, rbase[k]) : make_int2(0, 0);
        }
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k)
            if (st[k] == SP_GO) st[k] = block_stage_chrom(P, b0[k], rcs[k], q[k]);
        uint2 be[FILT_ITEMS];
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k) be[k] = st[k] == SP_GO ? __ldg(T.bp_bkt + q[k].slot) : make_uint2(0u, 0u);
        uint32_t pp[FILT_ITEMS];
        uint4 raw[FILT_ITEMS];
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k) {
            pp[k] = PIECE_NONE;
            if (st[k] == SP_GO) st[k] = block_stage_piece(T, q[k], be[k], pp[k]);
        }
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k)  // crowded buckets (1 in 8 blocks): one more entry
            if (st[k] == SP_SUB) be[k] = __ldg(T.bp_sub + q[k].slot);
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k) {
            if (st[k] == SP_SUB) st[k] = block_stage_sub(T, q[k], be[k], pp[k]);
            if (st[k] == SP_DEFER) pp[k] = block_stage_defer(P, q[k]) & PIECE_NONE;  // searched
            raw[k] = pp[k] != PIECE_NONE ? __ldg(reinterpret_cast<const uint4 *>(T.pinfo + pp[k])) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < FILT_ITEMS; ++k) {
            const uint32_t li = warp * FILT_WREADS + lane + (uint32_t)k * 32u;  // tile-local index
            const uint64_t r = tile0 + li;
            const bool listed = !single[k] && r < w1 && rn[k] != 0u;
            const unsigned lm = __ballot_sync(0xFFFFFFFFu, listed);
            if (single[k]) {
                const uint32_t res = block_stage_answer(q[k], pp[k], (int32_t)raw[k].y);
                emit(li, r, rbase[k], single_decide(T, b0[k], rcs[k], res, raw[k]));
            } else if (listed) {
                s_multi[n_listed + __popc(lm & ((1u << lane) - 1u))] = (uint16_t)li;
                s_rec[li] = make_uint4(rbase[k], rn[k], rcs[k], 0u);  // the read's inputs travel with the list entry
            } else {
                s_br[li] = BIN_NONE;  // no blocks, or past the end of the read set
            }
            n_listed += __popc(lm);
        }
    }
    // ---- phase C --------------------------------------------------------------------------
    __syncwarp();
#pragma unroll 1
    for (uint32_t k = lane; k < n_listed; k += 32u) {
        const uint32_t li = s_multi[k];
        const uint64_t r = tile0 + li;
        const uint4 in = s_rec[li];
        const uint32_t base = in.x, nraw = in.y;
        const int2 b0 = load_block(P, base);
        const int2 b1 = nraw >= 2u ? load_block(P, base + 1) : make_int2(0, 0);
        const TileCtx ctx{P, s_blk, base, wbase, b0, b1, staged};
        emit(li, r, base, filter_read_t(P, r, base, nraw, in.z, ctx));
    }
    __syncthreads();
    // ---- phase D --------------------------------------------------------------------------
#pragma unroll 1
    for (uint32_t b0 = tid; b0 < n_bins; b0 += FILT_BLOCK * FILT_RBATCH) {
        uint32_t c[FILT_RBATCH];
        unsigned long long old[FILT_RBATCH];
#pragma unroll
        for (int u = 0; u < FILT_RBATCH; ++u) {
            const uint32_t bin = b0 + (uint32_t)u * FILT_BLOCK;
            c[u] = bin < n_bins ? s_A[bin] : 0u;
            old[u] = 0ull;
            if (c[u]) old[u] = atomicAdd(S.state + bin, (unsigned long long)c[u]);
        }
#pragma unroll
        for (int u = 0; u < FILT_RBATCH; ++u) {
            if (!c[u]) continue;
            const uint32_t bin = b0 + (uint32_t)u * FILT_BLOCK;
            const uint32_t cls = (bin >= n_bkt ? 1u : 0u) + (bin >= 2u * n_bkt ? 1u : 0u);
            unsigned long long *st = S.state + bin;
            unsigned long long o = old[u];
            while ((uint32_t)o >= PART_CHUNK) {
                // the chunk was full: wait until the reservation that filled it has installed the
                // next chunk, then try again
                while ((uint32_t)(*reinterpret_cast<volatile unsigned long long *>(st)) >= PART_CHUNK) {
                }
                o = atomicAdd(st, (unsigned long long)c[u]);
            }
            const uint32_t f = (uint32_t)o, chunk = (uint32_t)(o >> 32);
            uint32_t slotB = 0, first = 0xFFFFu;
            if (f + c[u] >= PART_CHUNK) {
                // this reservation takes the chunk's last slot: install the bin's spare chunk.  Others
                // wait for the exchange, so nothing else happens before it; a spare equal to the
                // current chunk is the previous switch's, whose successor is not published yet.
                first = PART_CHUNK - f;
                volatile uint32_t *nx = S.next_chunk + bin;
                uint32_t j = *nx;
                while (j == chunk) j = *nx;
                atomicExch(st, ((unsigned long long)j << 32) | (unsigned long long)(c[u] - first));
                slotB = j * PART_CHUNK - first;  // slot of rank r >= first is slotB + r
                S.owner[j] = bin;
                const uint32_t w = atomicAdd(S.wl_n + cls, 1u);
                S.wl[(size_t)cls * S.pool_cap + w] = j;
                *nx = atomicAdd(S.pool_next, 1u);
                __threadfence();
            }
            s_A[bin] = chunk * PART_CHUNK + f;
            s_B[bin] = slotB;
            s_split[bin] = (uint16_t)first;
        }
    }
    __syncthreads();
    // ---- phase E --------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < FILT_ITEMS; ++k) {
        const uint32_t li = tid + (uint32_t)k * FILT_BLOCK;
        const uint32_t br = s_br[li];
        const uint32_t bin = br & 0xFFFFu, rank = br >> 16;
        if (bin == BIN_NONE) continue;
        const uint32_t slot = (rank < (uint32_t)s_split[bin] ? s_A[bin] : s_B[bin]) + rank;
        S.pool[slot] = s_rec[li];
    }
}

// Reads kept per class after the filter: every installed chunk of a bin is full except the current
// one.  counters[0] = FAST, [2] = PAIR, [3] = GENERAL reads (statistics and byte accounting).
__global__ void lsq_part_stats_kernel(PartState S, uint32_t n_bkt, unsigned long long *counters) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per bin; counters start at zero
    uint32_t v[3] = {0u, 0u, 0u};
    if (b < S.n_bins) {
        const uint32_t cls = (b >= n_bkt ? 1u : 0u) + (b >= 2u * n_bkt ? 1u : 0u);
        v[cls] = min((uint32_t)S.state[b], PART_CHUNK);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned long long t = __reduce_add_sync(0xFFFFFFFFu, v[k]);
        if (b == 0u) t += (unsigned long long)(S.wl_n[k] - n_bkt) * PART_CHUNK;
        if (lane_id() == 0u && t) atomicAdd(counters + (k == 0 ? 0 : k + 1), t);
    }
}

// ---------------------------------------------------------------------------------------------
// Count kernel
// ---------------------------------------------------------------------------------------------
// Per-lane update of a class table with more than 3 classes (dense K <= 8, or hashed).  Cold:
// scalar arguments only, so the hot kernel keeps its event header in registers.
__device__ __noinline__ void bump_class(unsigned long long *slot_cnt, unsigned long long *slot_mask,
                                        uint32_t slot_cap, bool hashed, uint64_t cm) {
    if (!hashed) {
        atomicAdd(slot_cnt + (uint32_t)(cm - 1ull), 1ull);
        return;
    }
    const uint32_t capm = slot_cap - 1u;
    uint32_t i = hash_mask(cm) & capm;
    while (true) {
        unsigned long long *kp = slot_mask + i;
        unsigned long long cur = *((volatile unsigned long long *)kp);
        if (cur == 0ull) cur = atomicCAS(kp, 0ull, (unsigned long long)cm);
        if (cur == 0ull || cur == cm) {
            atomicAdd(slot_cnt + i, 1ull);
            return;
        }
        i = (i + 1u) & capm;
    }
}

// Count kernel.  A work item is one chunk of one bin: up to PART_CHUNK reads of one class whose
// first kept start lies in one stab bucket, in no particular order.  The bucket's stab entries
// are contiguous in memory and are staged in shared memory by a bulk async copy (TMA, completion
// on an mbarrier) while the threads fetch their reads; every (read, stab entry) pair is
// evaluated by one lane and tallied with shared-memory atomics into four block-local counters
// per staged entry (classes {0}, {1}, {0,1} of events with at most two forms, and matched
// bases), which are flushed to the global 64-bit counters once per item.  Events with more
// forms update their class tables directly.  Buckets whose entries do not fit (pathological
// nesting) are read through L1/L2 and tallied with global atomics.
//
// FAST reads meet 1.7 entries on average but up to five, so a lane-per-read loop would leave most
// lanes idle: instead every warp first expands its reads into a queue of (entry, start, length)
// pairs in shared memory (offsets from a warp scan of the list lengths) and then evaluates the
// pairs densely, one per lane.
constexpr int COUNT_BLOCK = 256;
constexpr int COUNT_WARPS = COUNT_BLOCK / 32;
constexpr int COUNT_ITEMS = (int)PART_CHUNK / COUNT_BLOCK;  // 8 reads per thread and item
constexpr int COUNT_ROUND = 4;          // FAST: reads per thread expanded together
constexpr int COUNT_QCAP = 448;         // FAST: pair queue per warp (8 bytes per pair; dynamic shared memory)
constexpr size_t COUNT_FAST_DYN_SMEM = (size_t)COUNT_QCAP * 8 * (COUNT_BLOCK / 32);
constexpr int COUNT_MAX_ENT = 576;      // staged stab entries (18 KB): a bucket window + the longest stab list
constexpr int PAIR_QCAP = 256;          // PAIR: pair queue per warp (one read per lane and round)
constexpr uint32_t CTR_MAX_BASES = 1u << 20;  // larger matched lengths bypass the 32-bit block counters
#ifndef LSQ_CTR_PACKED
#define LSQ_CTR_PACKED 1
#endif
// packed class counter: reads in bits 20-31 (an item has at most PART_CHUNK = 2^11 reads), their matched bases in
// bits 0-19 (2^11 reads of fewer than 2^9 matched bases each stay below 2^20)
constexpr uint32_t CTR_PACK_SHIFT = 20, CTR_PACK_ONE = 1u << CTR_PACK_SHIFT, CTR_PACK_MAX_ML = 512u;
static_assert(PART_CHUNK <= (1u << 11), "packed counters hold 2^11 reads per item");

struct ItemTables {
    uint32_t chunk, n;   // pool chunk and its fill
    uint32_t e0, e1;     // stab entry range of the bucket
    bool staged;
};

template <int CLS>
__device__ __forceinline__ ItemTables item_tables(const CountParams &P, uint32_t it) {
    const PartState &S = P.part;
    ItemTables c;
    c.chunk = __ldg(S.wl + (size_t)CLS * S.pool_cap + it);
    const uint32_t bin = __ldg(S.owner + c.chunk);
    const unsigned long long st = __ldg(S.state + bin);
    c.n = (uint32_t)(st >> 32) == c.chunk ? min((uint32_t)st, PART_CHUNK) : PART_CHUNK;
    const uint32_t bkt = bin - (uint32_t)CLS * P.T.n_bkt;
    const uint32_t p_lo = __ldg(P.T.bkt_piece + bkt), p_hi = __ldg(P.T.bkt_piece + bkt + 1);
    c.e0 = __ldg(P.T.stab_off + p_lo);
    c.e1 = __ldg(P.T.stab_off + p_hi);
    c.staged = (c.e1 - c.e0) <= (uint32_t)COUNT_MAX_ENT && c.e1 > c.e0;
    return c;
}

// CLS 0: FAST reads -- purely table driven.  CLS 1: PAIR reads -- two blocks held in the record.
// CLS 2: GENERAL reads -- the walk of read.h:204-274 over the ext entries.  Separate
// instantiations keep the lean kernel's register count (and so its occupancy) independent of
// the walk's.
template <int CLS>
__global__ void __launch_bounds__(COUNT_BLOCK, CLS == 0 ? 4 : 3) lsq_count_kernel(CountParams P) {
    __shared__ __align__(128) StabEntry s_ent[COUNT_MAX_ENT];
    __shared__ __align__(16) uint32_t s_ctr[COUNT_MAX_ENT][4];  // block-local counters of the staged entries
    extern __shared__ __align__(16) uint2 s_q[];  // FAST: pair queues, COUNT_QCAP entries per warp
    __shared__ __align__(16) uint4 s_pr[CLS == 1 ? COUNT_WARPS : 1][CLS == 1 ? 32 : 1];        // PAIR: parked reads
    __shared__ uint32_t s_pq[CLS == 1 ? COUNT_WARPS : 1][CLS == 1 ? PAIR_QCAP : 1];            // PAIR: pair queues
    __shared__ __align__(8) uint64_t s_full;
    const Tables &T = P.T;
    const PartState &S = P.part;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    for (uint32_t i = tid; i < (uint32_t)COUNT_MAX_ENT * 4u; i += COUNT_BLOCK) (&s_ctr[0][0])[i] = 0u;
    if (tid == 0) mbar_init(&s_full, 1);
    __syncthreads();
    const uint32_t n_items = __ldg(S.wl_n + CLS);
    uint32_t my_slots = 0;  // GENERAL: ext slots read (byte accounting)
    uint32_t phase = 0;
#pragma unroll 1
    for (uint32_t it = blockIdx.x; it < n_items; it += gridDim.x) {
        const ItemTables ct = item_tables<CLS>(P, it);
        if (ct.n == 0u) continue;  // block-uniform
        if (ct.staged && tid == 0) {
            const uint32_t ebytes = (ct.e1 - ct.e0) * (uint32_t)sizeof(StabEntry);
            mbar_expect_tx(&s_full, ebytes);
            bulk_g2s(s_ent, T.stab + ct.e0, ebytes, &s_full);
        }
        // FAST: all of the thread's reads of this item are fetched while the entries are in flight;
        // the other classes keep one read in flight ahead of the one being walked
        constexpr int NPRE = CLS == 0 ? COUNT_ITEMS : 1;
        uint4 rec[NPRE];
        const uint4 *pool = S.pool + (size_t)ct.chunk * PART_CHUNK;
#pragma unroll
        for (int k = 0; k < NPRE; ++k) {
            const uint32_t i = tid + (uint32_t)k * COUNT_BLOCK;
            rec[k] = make_uint4(0u, 0u, 0u, 0u);  // no entries: past the end of the item
            if (i < ct.n) rec[k] = __ldcs(pool + i);
        }
        if (ct.staged) {
            mbar_wait(&s_full, phase);
            phase ^= 1u;
        }
        const StabEntry *ent = ct.staged ? s_ent - ct.e0 : T.stab;
        uint32_t(*ctr)[4] = ct.staged ? s_ctr - ct.e0 : nullptr;
        // tally of one valid (read, entry) pair with at most two forms
        auto tally = [&](const StabEntry &en, uint32_t e, uint32_t cm, uint32_t ml) {
#if LSQ_CTR_PACKED
            if (ctr && ml < CTR_PACK_MAX_ML) {
                atomicAdd(&ctr[e][cm - 1u], CTR_PACK_ONE | ml);  // class count and matched bases in one shared atomic
            } else if (ctr && ml < CTR_MAX_BASES) {
                atomicAdd(&ctr[e][cm - 1u], CTR_PACK_ONE);
                atomicAdd(&ctr[e][3], ml);
            } else {
#else
            if (ctr && ml < CTR_MAX_BASES) {
                atomicAdd(&ctr[e][cm - 1u], 1u);
                atomicAdd(&ctr[e][3], ml);
            } else {
#endif
                atomicAdd(P.slot_cnt + en.slot_off + (cm - 1u), 1ull);
                atomicAdd(P.bases + en.x, (unsigned long long)ml);
            }
        };
        auto tally_bases = [&](const StabEntry &en, uint32_t e, uint32_t ml) {
            if (ctr && ml < CTR_MAX_BASES) atomicAdd(&ctr[e][3], ml);
            else atomicAdd(P.bases + en.x, (unsigned long long)ml);
        };
        auto pair_fast = [&](uint32_t e, int32_t s0, int32_t e0) {  // one FAST read against stab entry e
            const StabEntry en = ent[e];  // SE_FAST by construction of the class
            uint32_t ml;
            const uint32_t cm = eval_fast(en, s0, e0, ml);
            if (cm == 0u) return;
            if (en.K > 2) {  // dense class table with more than 3 classes
                atomicAdd(P.slot_cnt + en.slot_off + (cm - 1u), 1ull);
                tally_bases(en, e, ml);
            } else {
                tally(en, e, cm, ml);
            }
        };
        if (CLS == 0) {
            uint2 *wq = s_q + warp * COUNT_QCAP;
#pragma unroll
            for (int h = 0; h < COUNT_ITEMS / COUNT_ROUND; ++h) {
                // pairs of this round: warp scan of the threads' list lengths
                uint32_t mine = 0;
#pragma unroll
                for (int u = 0; u < COUNT_ROUND; ++u) mine += rec[h * COUNT_ROUND + u].x >> PREC_SO_BITS;
                uint32_t incl = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                    if (lane >= (uint32_t)d) incl += t;
                }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                if (total == 0u) continue;  // warp-uniform
                if (ct.staged && total <= (uint32_t)COUNT_QCAP) {
                    // queue entry: x = entry relative to the bucket | length << 16, y = start
                    uint32_t pos = incl - mine;
#pragma unroll
                    for (int u = 0; u < COUNT_ROUND; ++u) {
                        const uint4 rc = rec[h * COUNT_ROUND + u];
                        const uint32_t rel = (rc.x & PREC_MAX_SO) - ct.e0, nent = rc.x >> PREC_SO_BITS;
                        const uint32_t hi = rc.w << 16;
#pragma unroll 1
                        for (uint32_t q = 0; q < nent; ++q) wq[pos++] = make_uint2((rel + q) | hi, rc.z);
                    }
                    __syncwarp();
#pragma unroll 1
                    for (uint32_t p = lane; p < total; p += 32u) {
                        const uint2 d = wq[p];
                        const int32_t s0 = (int32_t)d.y;
                        pair_fast(ct.e0 + (d.x & 0xFFFFu), s0, s0 + (int32_t)(d.x >> 16));
                    }
                    __syncwarp();  // the queue is rewritten by the next round
                } else {  // lane-per-read loop (unstaged bucket or an overfull queue)
#pragma unroll
                    for (int u = 0; u < COUNT_ROUND; ++u) {
                        const uint4 rc = rec[h * COUNT_ROUND + u];
                        const uint32_t so = rc.x & PREC_MAX_SO, nent = rc.x >> PREC_SO_BITS;
                        const int32_t s0 = (int32_t)rc.z, e0 = s0 + (int32_t)(rc.w & 0xFFFFu);
#pragma unroll 1
                        for (uint32_t q = 0; q < nent; ++q) pair_fast(so + q, s0, e0);
                    }
                }
            }
        } else {
            // one (read, stab entry) pair
            auto eval_entry = [&](const ReadView &rv, uint32_t strand, uint32_t total_len, uint32_t e) {
                const StabEntry en = ent[e];
                const EvHdr h = T.hdr[en.x];
                uint32_t ml;
                const uint64_t c64 =
                    (CLS == 1 && (en.meta & SE_FAST))
                        ? eval_pair2(en, h, T.seg + h.seg_off, T.fmask + h.form_off, rv.b0, rv.b1, total_len, ml)
                        : eval_pair(P, h, T.seg + h.seg_off, T.fmask + h.form_off, rv, strand, total_len, ml);
                if (c64 == 0ull) return;
                if (en.K > 2) {
                    bump_class(P.slot_cnt + h.slot_off, P.slot_mask + h.slot_off, h.slot_cap, (h.flags & EV_HASH) != 0,
                               c64);
                    tally_bases(en, e, ml);
                } else {
                    tally(en, e, (uint32_t)c64, ml);
                }
            };
            // PAIR: the two blocks of a record, relative to the first position of the piece
            auto pair_view = [&](const uint4 rc, ReadView &rv, uint32_t &strand, uint32_t &total_len) {
                rv.b0 = rv.b1 = rv.b2 = make_int2(0, 0);
                rv.ridx = 0;
                rv.blk = nullptr;
                unpack_pair(make_uint2(rc.z, rc.w), (int32_t)rc.y, rv.b0, rv.b1, strand);
                rv.n = 2;
                total_len = (uint32_t)(rv.b0.y - rv.b0.x) + (uint32_t)(rv.b1.y - rv.b1.x);
            };
            auto walk_entry = [&](const uint4 rc, uint32_t e) {  // PAIR read against stab entry e
                ReadView rv;
                uint32_t strand, total_len;
                pair_view(rc, rv, strand, total_len);
                eval_entry(rv, strand, total_len, e);
            };
            // one PAIR / GENERAL read against the stab entries of its piece, lane per read
            auto walk = [&](const uint4 rc) {
                uint32_t so = rc.x & PREC_MAX_SO, nent = rc.x >> PREC_SO_BITS;
                ReadView rv;
                uint32_t strand = 0, total_len = 0;
                if (CLS == 1) {
                    pair_view(rc, rv, strand, total_len);
                } else {
                    rv.b0 = rv.b1 = rv.b2 = make_int2(0, 0);
                    so = __ldg(T.stab_off + rc.y);  // the record names the piece
                    nent = __ldg(T.stab_off + rc.y + 1) - so;
                    // one pass over the ext entry: header and up to three blocks go to registers
                    const int2 *entry = P.ext + rc.z;  // 16-byte aligned
                    const int4 q0 = __ldg(reinterpret_cast<const int4 *>(entry));
                    const int2 hd = make_int2(q0.x, q0.y);
                    rv.b0 = make_int2(q0.z, q0.w);
                    const uint32_t nn = rc.w;  // kept blocks, also stored in the record
                    if (nn > 2) {
                        const int4 q1 = __ldg(reinterpret_cast<const int4 *>(entry + 2));
                        rv.b1 = make_int2(q1.x, q1.y);
                        rv.b2 = make_int2(q1.z, q1.w);
                    } else if (nn > 1) {
                        rv.b1 = __ldg(entry + 2);
                    }
                    rv.ridx = (uint32_t)hd.x;
                    rv.n = nn;
                    my_slots += nn + 1u;
                    strand = (uint32_t)hd.y >> 24;
                    rv.blk = entry + 1;
                    total_len = (uint32_t)(rv.b0.y - rv.b0.x) + (uint32_t)(rv.b1.y - rv.b1.x) +
                                (uint32_t)(rv.b2.y - rv.b2.x);
                    for (uint32_t b = 3; b < nn; ++b) {
                        const int2 blk = __ldg(rv.blk + b);
                        total_len += (uint32_t)(blk.y - blk.x);
                    }
                }
#pragma unroll 1
                for (uint32_t q = 0; q < nent; ++q) eval_entry(rv, strand, total_len, so + q);
            };
            if (CLS == 1 && ct.staged) {
                // PAIR reads meet about two entries each and their evaluation is long, so the same
                // expansion pays: per round every lane parks one read in the warp's shared slots and
                // queues its (entry, slot) pairs; the pairs are then evaluated densely, one per lane.
                uint4 *wr = s_pr[warp];
                uint32_t *wq = s_pq[warp];
                uint4 rec_c = rec[0];
#pragma unroll 1
                for (uint32_t i0 = tid & ~31u; i0 < ct.n; i0 += COUNT_BLOCK) {  // warp-uniform
                    const uint32_t i = i0 + lane;
                    const uint4 rc = i < ct.n ? rec_c : make_uint4(0u, 0u, 0u, 0u);
                    if (i + COUNT_BLOCK < ct.n) rec_c = __ldcs(pool + i + COUNT_BLOCK);
                    const uint32_t mine = rc.x >> PREC_SO_BITS;
                    uint32_t incl = mine;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= (uint32_t)d) incl += t;
                    }
                    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                    if (total == 0u) continue;
                    if (total <= (uint32_t)PAIR_QCAP) {
                        wr[lane] = rc;
                        const uint32_t rel = (rc.x & PREC_MAX_SO) - ct.e0;
                        uint32_t pos = incl - mine;
#pragma unroll 1
                        for (uint32_t q = 0; q < mine; ++q) wq[pos++] = (rel + q) | (lane << 16);
                        __syncwarp();
#pragma unroll 1
                        for (uint32_t p = lane; p < total; p += 32u) {
                            const uint32_t d = wq[p];
                            walk_entry(wr[d >> 16], ct.e0 + (d & 0xFFFFu));
                        }
                        __syncwarp();  // slots and queue are rewritten by the next round
                    } else {
                        walk(rc);
                    }
                }
            } else {
                uint4 rec_c = rec[0];
#pragma unroll 1
                for (uint32_t i = tid; i < ct.n; i += COUNT_BLOCK) {
                    const uint4 rc = rec_c;
                    if (i + COUNT_BLOCK < ct.n) rec_c = __ldcs(pool + i + COUNT_BLOCK);
                    walk(rc);
                }
            }
        }
        __syncthreads();
        if (ct.staged) {  // flush the block-local counters of this item
            const uint32_t nctr = (ct.e1 - ct.e0) * 4u;
            uint32_t *flat = &s_ctr[0][0];
            for (uint32_t i = tid; i < nctr; i += COUNT_BLOCK) {
                const uint32_t v = flat[i];
                if (!v) continue;
                flat[i] = 0u;
                const StabEntry &en = s_ent[i >> 2];
#if LSQ_CTR_PACKED
                if ((i & 3u) == 3u) {
                    atomicAdd(P.bases + en.x, (unsigned long long)v);
                } else {
                    atomicAdd(P.slot_cnt + en.slot_off + (i & 3u), (unsigned long long)(v >> CTR_PACK_SHIFT));
                    if (v & (CTR_PACK_ONE - 1u)) atomicAdd(P.bases + en.x, (unsigned long long)(v & (CTR_PACK_ONE - 1u)));
                }
#else
                if ((i & 3u) == 3u) atomicAdd(P.bases + en.x, (unsigned long long)v);
                else atomicAdd(P.slot_cnt + en.slot_off + (i & 3u), (unsigned long long)v);
#endif
            }
        }
        __syncthreads();  // the staged entries may be overwritten by the next item
    }
    if (CLS == 2) {
        const uint32_t sl = __reduce_add_sync(0xFFFFFFFFu, my_slots);
        if (lane == 0 && sl) atomicAdd(P.stats, (unsigned long long)sl);
    }
}

// ---------------------------------------------------------------------------------------------
// Class iteration shared by finalize and solve: visits (mask, count) of event h in read set m.
// ---------------------------------------------------------------------------------------------
struct ClassRange {
    const unsigned long long *cnt;
    const unsigned long long *mask;  // null: dense table, mask = index + 1
    uint32_t n;
};

__device__ __forceinline__ ClassRange class_range(const SolveParams &P, const EvHdr &h, uint32_t m) {
    ClassRange r;
    if (P.inj_off[m]) {
        uint64_t a = P.inj_off[m][h.out_ev], b = P.inj_off[m][h.out_ev + 1];
        r.cnt = (const unsigned long long *)P.inj_cnt[m] + a;
        r.mask = (const unsigned long long *)P.inj_mask[m] + a;
        r.n = (uint32_t)(b - a);
    } else {
        size_t o = (size_t)m * P.T.total_slots + h.slot_off;
        r.cnt = P.slot_cnt + o;
        r.mask = (h.flags & EV_HASH) ? P.slot_mask + o : nullptr;
        r.n = h.slot_cap;
    }
    return r;
}

// One thread per event: n_valid[m], bases[m] and per-form tallies in the caller's order.
__global__ void __launch_bounds__(128) lsq_finalize_kernel(SolveParams P) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long pairs = 0;
    if (x < P.T.E) {
        const EvHdr h = P.T.hdr[x];
        unsigned long long iso[LSQ_MAX_FORMS_PER_EVENT_DEV];
        for (uint32_t k = 0; k < h.K; ++k) iso[k] = 0;
        for (uint32_t m = 0; m < P.M; ++m) {
            ClassRange cr = class_range(P, h, m);
            unsigned long long nv = 0;
            for (uint32_t i = 0; i < cr.n; ++i) {
                unsigned long long c = cr.cnt[i];
                if (!c) continue;
                unsigned long long mk = cr.mask ? cr.mask[i] : (unsigned long long)(i + 1u);
                nv += c;
                for (uint32_t k = 0; k < h.K; ++k)
                    if ((mk >> k) & 1ull) iso[k] += c;
            }
            P.n_valid[(size_t)m * P.T.E + h.out_ev] = nv;
            P.bases_out[(size_t)m * P.T.E + h.out_ev] = P.bases[(size_t)m * P.T.E + x];
            pairs += nv;
        }
        for (uint32_t k = 0; k < h.K; ++k) P.iso_count[h.out_form + k] = iso[k];
    }
    // statistics: total valid (read, event) pairs, one atomic per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, d);
    if ((threadIdx.x & 31u) == 0 && pairs && P.pairs_valid) atomicAdd(P.pairs_valid, pairs);
}

// ---------------------------------------------------------------------------------------------
// Solve kernel: one warp per event, lanes over compatibility classes.
// All floating point goes through __dmul_rn/__dadd_rn/__ddiv_rn so that nvcc cannot contract
// multiply-add pairs into FMAs: per class the arithmetic is bit-identical to the reference's
// per-read arithmetic (read.h:604-613, 628-632); only the order of the sum over reads differs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xFFFFFFFFu, v, d));
    return v;
}

// s = sum_k theta[k] * M(i,k) over the forms of the mask in ascending k (zero terms add nothing)
__device__ __forceinline__ double class_prob(uint64_t mk, const double *theta, const double *G) {
    double s = 0.0;
    while (mk) {
        int k = __ffsll((long long)mk) - 1;
        mk &= mk - 1ull;
        s = __dadd_rn(s, __dmul_rn(theta[k], G[k]));
    }
    return s;
}

constexpr int SE_MAX_FORMS = 16;  // standard errors are computed for events with at most 16 forms

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 3) lsq_solve_kernel(SolveParams P) {
    __shared__ double s_theta[WARPS][2][LSQ_MAX_FORMS_PER_EVENT_DEV];
    __shared__ double s_fim[WARPS][SE_MAX_FORMS - 1][2 * (SE_MAX_FORMS - 1)];  // [I | identity]
    const uint32_t lane = lane_id();
    const uint32_t w = threadIdx.x >> 5;
    const uint32_t E = P.T.E;
    for (uint32_t x = blockIdx.x * WARPS + w; x < E; x += gridDim.x * WARPS) {
        const EvHdr h = P.T.hdr[x];
        const uint32_t K = h.K;
        if (P.small_enabled && K <= 2) continue;  // solved by lsq_solve_small_kernel
        double *th_old = s_theta[w][0], *th_new = s_theta[w][1];
        // N = total number of valid reads (read.h:599-602), sum of supports (solve.cpp:826)
        double n_total = 0.0, sum_supports = 0.0;
        bool any = false;
        for (uint32_t m = 0; m < P.M; ++m) {
            unsigned long long nv = P.n_valid[(size_t)m * E + h.out_ev];
            if (nv) {
                any = true;
                n_total = __dadd_rn(n_total, (double)nv);
            }
            sum_supports = __dadd_rn(sum_supports, (double)nv);
        }
        __syncwarp();
        for (uint32_t k = lane; k < K; k += 32) th_new[k] = __ddiv_rn(1.0, (double)K);
        __syncwarp();
        uint32_t iters = 0;
        double ll = 0.0;
        // Each lane keeps the (mask, count) classes it owns in registers together with the class
        // probability s_c at the current theta, so an EM iteration evaluates s_c once per class
        // instead of once per (class, form).  Lanes own classes in the same order as the strided
        // loops of the general path below, so both paths add in the same order (same bits).
        constexpr int CPL = 3;  // classes per lane held in registers: up to 96 classes per event
        uint64_t cmk[CPL];
        double ccnt[CPL], cs[CPL];
        const double *cG[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            cmk[j] = 0;
            ccnt[j] = 0.0;
            cs[j] = 1.0;
            cG[j] = P.G;
        }
        uint32_t nown = 0;
        for (uint32_t m = 0; m < P.M; ++m) {
            ClassRange cr = class_range(P, h, m);
            const double *G = P.G + (size_t)m * P.T.F + h.form_off;
            for (uint32_t i = lane; i < cr.n; i += 32) {
                unsigned long long c = cr.cnt[i];
                if (!c) continue;
                uint64_t mk = cr.mask ? cr.mask[i] : (uint64_t)(i + 1u);
#pragma unroll
                for (int j = 0; j < CPL; ++j)
                    if (nown == (uint32_t)j) {
                        cmk[j] = mk;
                        ccnt[j] = (double)c;
                        cG[j] = G;
                    }
                ++nown;
            }
        }
        const bool cached = !__any_sync(0xFFFFFFFFu, nown > (uint32_t)CPL);
        auto refresh = [&](const double *th) {  // s_c at theta and the log-likelihood (read.h:620-636)
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < CPL; ++j)
                if ((uint32_t)j < nown) {
                    cs[j] = class_prob(cmk[j], th, cG[j]);
                    part = __dadd_rn(part, __dmul_rn(ccnt[j], log(cs[j])));
                }
            return warp_sum(part);
        };
        auto loglik = [&](const double *th) {  // read.h:620-636
            double part = 0.0;
            for (uint32_t m = 0; m < P.M; ++m) {
                ClassRange cr = class_range(P, h, m);
                const double *G = P.G + (size_t)m * P.T.F + h.form_off;
                for (uint32_t i = lane; i < cr.n; i += 32) {
                    unsigned long long c = cr.cnt[i];
                    if (!c) continue;
                    uint64_t mk = cr.mask ? cr.mask[i] : (uint64_t)(i + 1u);
                    part = __dadd_rn(part, __dmul_rn((double)c, log(class_prob(mk, th, G))));
                }
            }
            return warp_sum(part);
        };
        if (!any) {
            // no valid read in any read set: uniform theta, log-likelihood 0 (solve.cpp:798-800)
        } else if (K == 1) {
            if (lane == 0) th_new[0] = 1.0;  // solve.cpp:801-802
            __syncwarp();
            ll = cached ? refresh(th_new) : loglik(th_new);
        } else if (cached) {
            double ll_old = refresh(th_new);
            while (true) {  // read.h:647-659
                __syncwarp();
                for (uint32_t k = lane; k < K; k += 32) th_old[k] = th_new[k];
                __syncwarp();
                for (uint32_t k = 0; k < K; ++k) {  // read.h:597-617; cs[] holds s_c at th_old
                    double part = 0.0;
                    const double tk = th_old[k];
#pragma unroll
                    for (int j = 0; j < CPL; ++j)
                        if ((uint32_t)j < nown && ((cmk[j] >> k) & 1ull) && cs[j] > 0.0) {
                            const double lp = __dmul_rn(tk, cG[j][k]);
                            if (lp > 0.0) part = __dadd_rn(part, __dmul_rn(ccnt[j], __ddiv_rn(lp, cs[j])));
                        }
                    const double z = warp_sum(part);
                    if (lane == 0) th_new[k] = __ddiv_rn(z, n_total);
                }
                __syncwarp();
                ll = refresh(th_new);
                ++iters;
                const double ratio = __ddiv_rn(ll_old, ll);
                if (!(fabs(__dsub_rn(1.0, ratio)) > 1E-6) || iters >= P.max_iters) break;
                ll_old = ll;
            }
        } else {
            double ll_old = loglik(th_new);
            while (true) {  // read.h:647-659
                __syncwarp();
                for (uint32_t k = lane; k < K; k += 32) th_old[k] = th_new[k];
                __syncwarp();
                for (uint32_t k = 0; k < K; ++k) {  // read.h:597-617
                    double part = 0.0;
                    for (uint32_t m = 0; m < P.M; ++m) {
                        ClassRange cr = class_range(P, h, m);
                        const double *G = P.G + (size_t)m * P.T.F + h.form_off;
                        for (uint32_t i = lane; i < cr.n; i += 32) {
                            unsigned long long c = cr.cnt[i];
                            if (!c) continue;
                            uint64_t mk = cr.mask ? cr.mask[i] : (uint64_t)(i + 1u);
                            if (!((mk >> k) & 1ull)) continue;
                            double s = class_prob(mk, th_old, G);
                            if (s > 0.0) {
                                double lp = __dmul_rn(th_old[k], G[k]);
                                if (lp > 0.0) part = __dadd_rn(part, __dmul_rn((double)c, __ddiv_rn(lp, s)));
                            }
                        }
                    }
                    double z = warp_sum(part);
                    if (lane == 0) th_new[k] = __ddiv_rn(z, n_total);
                }
                __syncwarp();
                ll = loglik(th_new);
                ++iters;
                double ratio = __ddiv_rn(ll_old, ll);
                if (!(fabs(__dsub_rn(1.0, ratio)) > 1E-6) || iters >= P.max_iters) break;
                ll_old = ll;  // the next iteration's old_ll is this very value (same theta, same order)
            }
        }
        __syncwarp();
        // Standard errors (extension, no oracle in the reference): observed Fisher information of
        // the free parameters theta_0..theta_{K-2} (theta_{K-1} eliminated), per-read term as in
        // common/fim.h:322-367: I_pq = sum_c n_c (v_cp - v_cK)(v_cq - v_cK) / s_c^2, inverted by
        // Gauss-Jordan with partial pivoting (the reference's linalg.h uses a GSL LU for the same).
        {
            const uint32_t D = K - 1u;
            if (K == 1u) {
                if (lane == 0) P.se[h.out_form] = 0.0;
            } else if (!any || K > (uint32_t)SE_MAX_FORMS) {
                for (uint32_t k = lane; k < K; k += 32) P.se[h.out_form + k] = nan("");
            } else {
                double(*A)[2 * (SE_MAX_FORMS - 1)] = s_fim[w];
                for (uint32_t p = 0; p < D; ++p)
                    for (uint32_t q = p; q < D; ++q) {
                        double part = 0.0;
                        if (cached) {  // cs[] holds s_c at the final theta
#pragma unroll
                            for (int j = 0; j < CPL; ++j)
                                if ((uint32_t)j < nown && cs[j] > 0.0) {
                                    const double *G = cG[j];
                                    const double vK = ((cmk[j] >> D) & 1ull) ? G[D] : 0.0;
                                    const double vp = (((cmk[j] >> p) & 1ull) ? G[p] : 0.0) - vK;
                                    const double vq = (((cmk[j] >> q) & 1ull) ? G[q] : 0.0) - vK;
                                    part += ccnt[j] * (vp / cs[j]) * (vq / cs[j]);
                                }
                        } else
                        for (uint32_t m = 0; m < P.M; ++m) {
                            ClassRange cr = class_range(P, h, m);
                            const double *G = P.G + (size_t)m * P.T.F + h.form_off;
                            for (uint32_t i = lane; i < cr.n; i += 32) {
                                unsigned long long c = cr.cnt[i];
                                if (!c) continue;
                                uint64_t mk = cr.mask ? cr.mask[i] : (uint64_t)(i + 1u);
                                double sc = class_prob(mk, th_new, G);
                                if (!(sc > 0.0)) continue;
                                double vK = ((mk >> D) & 1ull) ? G[D] : 0.0;
                                double vp = (((mk >> p) & 1ull) ? G[p] : 0.0) - vK;
                                double vq = (((mk >> q) & 1ull) ? G[q] : 0.0) - vK;
                                part += (double)c * (vp / sc) * (vq / sc);
                            }
                        }
                        double v = warp_sum(part);
                        if (lane == 0) {
                            A[p][q] = v;
                            A[q][p] = v;
                        }
                    }
                __syncwarp();
                for (uint32_t i = lane; i < D * D; i += 32) A[i / D][D + i % D] = (i / D == i % D) ? 1.0 : 0.0;
                __syncwarp();
                bool singular = false;
                for (uint32_t r = 0; r < D; ++r) {  // lane j owns column j of the augmented matrix
                    uint32_t piv = r;
                    double best = fabs(A[r][r]);
                    for (uint32_t i = r + 1; i < D; ++i)
                        if (fabs(A[i][r]) > best) {
                            best = fabs(A[i][r]);
                            piv = i;
                        }
                    if (!(best > 0.0)) {
                        singular = true;
                        break;
                    }
                    __syncwarp();
                    if (lane < 2 * D) {
                        double a = A[r][lane], b = A[piv][lane];
                        A[r][lane] = b;
                        A[piv][lane] = a;
                    }
                    __syncwarp();
                    const double pv = A[r][r];
                    __syncwarp();
                    if (lane < 2 * D) A[r][lane] /= pv;
                    __syncwarp();
                    for (uint32_t i = 0; i < D; ++i) {
                        if (i == r) continue;
                        const double f = A[i][r];
                        __syncwarp();
                        if (lane < 2 * D) A[i][lane] -= f * A[r][lane];
                        __syncwarp();
                    }
                }
                if (singular) {
                    for (uint32_t k = lane; k < K; k += 32) P.se[h.out_form + k] = nan("");
                } else {
                    double tot = 0.0;
                    for (uint32_t i = 0; i < D * D; ++i) tot += A[i / D][D + i % D];
                    for (uint32_t k = lane; k < D; k += 32) P.se[h.out_form + k] = sqrt(A[k][D + k]);
                    if (lane == 0) P.se[h.out_form + D] = sqrt(tot);
                }
            }
        }
        __syncwarp();
        // RPKM (solve.cpp:809-821) and outputs
        for (uint32_t k = lane; k < K; k += 32) {
            double th = th_new[k];
            double r = 0.0;
            for (uint32_t m = 0; m < P.M; ++m)
                r = __dadd_rn(r, __dmul_rn((double)P.bases[(size_t)m * E + x], th));
            r = __ddiv_rn(r, __ddiv_rn((double)P.form_len[h.form_off + k], 1.0E3));
            r = __ddiv_rn(r, P.total_read_mbases);
            P.theta[h.out_form + k] = th;
            P.rpkm[h.out_form + k] = r;
        }
        if (lane == 0) {
            P.loglik[h.out_ev] = ll;
            P.ll_per_read[h.out_ev] = (sum_supports > 1E-5) ? __ddiv_rn(ll, sum_supports) : 0.0;
            P.iters[h.out_ev] = iters;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Solve kernel for the common small events (K <= 2 forms, M <= 2 counted read sets): one THREAD
// per event, the whole EM state in registers, and lanes fetch a new event as soon as theirs has
// converged, so iteration-count skew between events does not idle a warp.  Arithmetic per class
// is the same as in lsq_solve_kernel (and the reference's per-read arithmetic); only the order
// of the sums over classes differs.
// ---------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(128) lsq_solve_small_kernel(SolveParams P) {
    const uint32_t E = P.T.E;
    while (true) {
        const uint32_t x = atomicAdd(P.work_counter, 1u);
        if (x >= E) return;
        const EvHdr h = P.T.hdr[x];
        const uint32_t K = h.K;
        if (K > 2) continue;
        double cnt[M][3], G[M][2];
        double n_total = 0.0, sum_supports = 0.0;
        bool any = false;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const unsigned long long *sc = P.slot_cnt + (size_t)m * P.T.total_slots + h.slot_off;
            cnt[m][0] = (double)sc[0];
            cnt[m][1] = K == 2 ? (double)sc[1] : 0.0;
            cnt[m][2] = K == 2 ? (double)sc[2] : 0.0;
            const double *g = P.G + (size_t)m * P.T.F + h.form_off;
            G[m][0] = g[0];
            G[m][1] = K == 2 ? g[1] : 0.0;
            const unsigned long long nv = P.n_valid[(size_t)m * E + h.out_ev];
            if (nv) {
                any = true;
                n_total = __dadd_rn(n_total, (double)nv);
            }
            sum_supports = __dadd_rn(sum_supports, (double)nv);
        }
        double th0 = __ddiv_rn(1.0, (double)K), th1 = th0;
        // log-likelihood at (t0, t1): sum over classes of count * log(sum_k theta_k G_k)
        auto loglik = [&](double t0, double t1) {
            double ll = 0.0;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const double p0 = __dmul_rn(t0, G[m][0]), p1 = __dmul_rn(t1, G[m][1]);
                if (cnt[m][0] != 0.0) ll = __dadd_rn(ll, __dmul_rn(cnt[m][0], log(p0)));
                if (cnt[m][1] != 0.0) ll = __dadd_rn(ll, __dmul_rn(cnt[m][1], log(p1)));
                if (cnt[m][2] != 0.0) ll = __dadd_rn(ll, __dmul_rn(cnt[m][2], log(__dadd_rn(p0, p1))));
            }
            return ll;
        };
        uint32_t iters = 0;
        double ll = 0.0;
        if (!any) {
        } else if (K == 1) {
            th0 = 1.0;
            ll = loglik(1.0, 0.0);
        } else {
            double ll_old = loglik(th0, th1);
            while (true) {
                double z0 = 0.0, z1 = 0.0;
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const double p0 = __dmul_rn(th0, G[m][0]), p1 = __dmul_rn(th1, G[m][1]);
                    const double s3 = __dadd_rn(p0, p1);
                    // form 0: classes {0} and {0,1}; form 1: classes {1} and {0,1}
                    if (p0 > 0.0) {
                        if (cnt[m][0] != 0.0) z0 = __dadd_rn(z0, __dmul_rn(cnt[m][0], __ddiv_rn(p0, p0)));
                        if (cnt[m][2] != 0.0 && s3 > 0.0) z0 = __dadd_rn(z0, __dmul_rn(cnt[m][2], __ddiv_rn(p0, s3)));
                    }
                    if (p1 > 0.0) {
                        if (cnt[m][1] != 0.0) z1 = __dadd_rn(z1, __dmul_rn(cnt[m][1], __ddiv_rn(p1, p1)));
                        if (cnt[m][2] != 0.0 && s3 > 0.0) z1 = __dadd_rn(z1, __dmul_rn(cnt[m][2], __ddiv_rn(p1, s3)));
                    }
                }
                th0 = __ddiv_rn(z0, n_total);
                th1 = __ddiv_rn(z1, n_total);
                ll = loglik(th0, th1);
                ++iters;
                const double ratio = __ddiv_rn(ll_old, ll);
                if (!(fabs(__dsub_rn(1.0, ratio)) > 1E-6) || iters >= P.max_iters) break;
                ll_old = ll;
            }
        }
        // standard error from the observed Fisher information of the free parameter theta_0
        // (theta_1 = 1 - theta_0): I = sum_c n_c (v_c0 - v_c1)^2 / s_c^2 with v_ck = G_k [k in c]
        // (per-read term of common/fim.h:322-367); se_0 = se_1 = I^{-1/2}.  Extension: the reference
        // never evaluates fim.h, so this has no oracle.
        double se = 0.0;
        if (K == 2) {
            double info = 0.0;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const double p0 = th0 * G[m][0], p1 = th1 * G[m][1], s3 = p0 + p1;
                if (cnt[m][0] != 0.0 && p0 > 0.0) info += cnt[m][0] * (G[m][0] / p0) * (G[m][0] / p0);
                if (cnt[m][1] != 0.0 && p1 > 0.0) info += cnt[m][1] * (G[m][1] / p1) * (G[m][1] / p1);
                if (cnt[m][2] != 0.0 && s3 > 0.0) {
                    const double d = (G[m][0] - G[m][1]) / s3;
                    info += cnt[m][2] * d * d;
                }
            }
            se = any && info > 0.0 ? sqrt(1.0 / info) : nan("");
        }
        for (uint32_t k = 0; k < K; ++k) {
            const double th = k == 0 ? th0 : th1;
            P.se[h.out_form + k] = se;
            double r = 0.0;
#pragma unroll
            for (int m = 0; m < M; ++m) r = __dadd_rn(r, __dmul_rn((double)P.bases[(size_t)m * E + x], th));
            r = __ddiv_rn(r, __ddiv_rn((double)P.form_len[h.form_off + k], 1.0E3));
            r = __ddiv_rn(r, P.total_read_mbases);
            P.theta[h.out_form + k] = th;
            P.rpkm[h.out_form + k] = r;
        }
        P.loglik[h.out_ev] = ll;
        P.ll_per_read[h.out_ev] = (sum_supports > 1E-5) ? __ddiv_rn(ll, sum_supports) : 0.0;
        P.iters[h.out_ev] = iters;
    }
}

}  // namespace lsq
