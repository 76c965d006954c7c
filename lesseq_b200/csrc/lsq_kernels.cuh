// Hand-written sm_100a kernels of the count + solve hot path.
//
//   lsq_filter_kernel  per-block covered-region filter + per-read interval list + stab piece key
//                      (count/count.cpp:319-324; jsc/util/interval_list.hpp:396-422, 462-503;
//                       index key of count.cpp:346-364)
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
// Filter kernel: one thread per read.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__global__ void __launch_bounds__(256) lsq_filter_kernel(FilterParams P) {
    const Tables &T = P.T;
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.n_reads) return;
    const FilteredRead fr = filter_read(P, r);
    const bool kept = fr.key != KEY_DROPPED(T);
    ReadRec rec;
    if (kept && fr.key >= T.n_pieces && fr.key < 2u * T.n_pieces) {
        rec = pack_pair(fr.first, fr.second, fr.piece_pos, fr.strand);  // PAIR class: two blocks inline
    } else if (kept && fr.key >= 2u * T.n_pieces) {
        // GENERAL class: the ext entry of read r lives at the even slot (blk_off[r] + 2r) & ~1:
        // 1 header + n <= 1 + raw blocks, so no allocation is needed, entries of different reads
        // never overlap, and header + first block form one aligned 128-bit access.  Blocks 1..n-1
        // were left in place by filter_read.
        const uint32_t off = ext_slot(__ldg(P.blk_off + r), (uint32_t)r);
        *reinterpret_cast<int4 *>(P.ext + off) =
            make_int4((int32_t)(uint32_t)r, (int32_t)((fr.strand << 24) | (fr.n & 0xFFFFFFu)), fr.first.x, fr.first.y);
        rec.x = off;
        rec.y = fr.n;
    } else {
        rec.x = (uint32_t)fr.first.x;
        rec.y = (uint32_t)(fr.first.y - fr.first.x) | (fr.strand << 16);
    }
    __stcs(P.rkey + r, fr.key);  // streamed to the sort: do not displace the tables in L2
    __stcs(P.rec + r, rec);
}

// After the sort: class sizes from the sorted keys (three binary searches by one thread) --
// counters[0] = kept reads (key < 3 n_pieces), counters[2] = PAIR reads, counters[3] = GENERAL reads.
__global__ void lsq_class_bounds_kernel(const uint32_t *__restrict__ g_key, unsigned long long n, uint32_t n_pieces,
                                        unsigned long long *counters) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    auto lower = [&](uint32_t v) {
        unsigned long long lo = 0, hi = n;
        while (lo < hi) {
            unsigned long long mid = (lo + hi) >> 1;
            if (g_key[mid] < v) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    };
    const unsigned long long kept = lower(3u * n_pieces), fast = lower(n_pieces), pair_end = lower(2u * n_pieces);
    counters[0] = kept;
    counters[2] = pair_end - fast;
    counters[3] = kept - pair_end;
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

// Count kernel.  Reads arrive sorted by (stab piece, multi-block flag): a warp holds a few RUNS of
// reads with the same key.  Every lane evaluates its own read against the j-th stab entry of its
// own piece (lanes of a run share the entry, so the loads broadcast); per-class counts of a run
// are formed with __ballot_sync/__popc restricted to the run's lane mask, the matched bases with a
// segmented shuffle reduction, and the run's head lane issues the global reductions.
//
// A CTA works on chunks of COUNT_CHUNK consecutive reads.  The stab entries and offsets of a
// chunk's piece range are contiguous in memory and are staged in shared memory by bulk async
// copies (TMA) one chunk ahead: full[stage] completes when the bytes have landed, empty[stage]
// when all warps have released the stage.  Chunks whose tables do not fit are read through L1/L2.
constexpr int COUNT_BLOCK = 256;
constexpr int COUNT_WARPS = COUNT_BLOCK / 32;
constexpr int COUNT_ITEMS = 8;
constexpr int COUNT_CHUNK = COUNT_BLOCK * COUNT_ITEMS;
constexpr int COUNT_MAX_ENT = 256;     // staged stab entries per stage (8 KB)
constexpr int COUNT_MAX_OFF = 512;     // staged offsets per stage (2 KB)

struct ChunkTables {
    uint32_t e0, e1;   // stab entry range
    uint32_t pa, noff; // 16-byte aligned offset slice
    bool staged;
};

__device__ __forceinline__ ChunkTables chunk_tables(const CountParams &P, unsigned long long base,
                                                    unsigned long long end, uint32_t key_bias) {
    ChunkTables c;
    const uint32_t pf = __ldg(P.g_key + base) - key_bias, pl = __ldg(P.g_key + end - 1) - key_bias;
    c.e0 = __ldg(P.T.stab_off + pf);
    c.e1 = __ldg(P.T.stab_off + pl + 1);
    c.pa = pf & ~3u;
    c.noff = ((pl + 2 - c.pa) + 3u) & ~3u;
    c.staged = (c.e1 - c.e0) <= (uint32_t)COUNT_MAX_ENT && c.noff <= (uint32_t)COUNT_MAX_OFF && c.e1 > c.e0;
    return c;
}

// One q-step of the run aggregation: per-class counts of each run via ballots restricted to the
// run's lane mask, matched bases via a segmented shuffle reduction, reductions by the head lane.
// Staged chunks accumulate in block-local shared-memory counters (4 per staged stab entry: three
// classes + bases) that are flushed once per chunk; other chunks reduce straight to global memory.
__device__ __forceinline__ void aggregate_runs(const CountParams &P, uint32_t lane, bool head_has, unsigned run_mask,
                                               uint32_t run_end, bool single_run, uint32_t cm, uint32_t ml,
                                               uint32_t slot_off, uint32_t x, uint32_t *s_ctr /* or null */) {
    // cm is 1, 2 or 3 for events with at most two forms (bit k = compatible with form k);
    // larger tables were updated by the caller and arrive as the 0x80000000 marker
    const unsigned f0 = __ballot_sync(0xFFFFFFFFu, (cm & 1u) != 0u);
    const unsigned f1 = __ballot_sync(0xFFFFFFFFu, (cm & 2u) != 0u);
    uint32_t v;
    if (single_run) {  // warp-uniform: the whole warp is one run
        v = __reduce_add_sync(0xFFFFFFFFu, ml);
    } else {
        v = ml;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_down_sync(0xFFFFFFFFu, v, d);
            if (lane + (uint32_t)d < run_end) v += t;
        }
    }
    if (head_has) {
        const uint32_t c1 = __popc(f0 & ~f1 & run_mask), c2 = __popc(f1 & ~f0 & run_mask),
                       c3 = __popc(f0 & f1 & run_mask);
        if (s_ctr) {
            if (c1) atomicAdd(s_ctr + 0, c1);
            if (c2) atomicAdd(s_ctr + 1, c2);
            if (c3) atomicAdd(s_ctr + 2, c3);
            if (v) atomicAdd(s_ctr + 3, v);
        } else {
            if (c1) atomicAdd(P.slot_cnt + slot_off, (unsigned long long)c1);
            if (c2) atomicAdd(P.slot_cnt + slot_off + 1u, (unsigned long long)c2);
            if (c3) atomicAdd(P.slot_cnt + slot_off + 2u, (unsigned long long)c3);
            if (v) atomicAdd(P.bases + x, (unsigned long long)v);
        }
    }
}

// CLS 0: FAST-class reads, positions [0, n_fast) of the sorted arrays -- purely table driven.
// CLS 1: PAIR-class reads [n_fast, n_fast + n_pair) -- the general walk over two blocks held in the
// sorted record.  CLS 2: GENERAL reads [.., n_kept) -- the general walk over the ext entries.
// Separate instantiations keep the lean kernel's register count (and so its occupancy)
// independent of the walk's.
template <int CLS>
__global__ void __launch_bounds__(COUNT_BLOCK, CLS ? 3 : 5) lsq_count_kernel(CountParams P) {
    constexpr bool GENERAL = CLS != 0;
    __shared__ __align__(128) StabEntry s_ent[2][COUNT_MAX_ENT];
    __shared__ __align__(16) uint32_t s_off[2][COUNT_MAX_OFF];
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    __shared__ uint32_t s_ctr[2][COUNT_MAX_ENT][4];  // block-local counters of the staged entries
    __shared__ uint32_t s_done[2];                   // warps that finished the chunk of a stage
    const Tables &T = P.T;
    for (uint32_t i = threadIdx.x; i < 2u * COUNT_MAX_ENT * 4u; i += COUNT_BLOCK) (&s_ctr[0][0][0])[i] = 0u;
    if (threadIdx.x < 2) s_done[threadIdx.x] = 0u;
    const unsigned long long n_fast = P.counters[0] - P.counters[3] - P.counters[2];
    const unsigned long long n_pair_end = n_fast + P.counters[2];
    const unsigned long long r_begin = CLS == 0 ? 0ull : (CLS == 1 ? n_fast : n_pair_end);
    const unsigned long long r_end = CLS == 0 ? n_fast : (CLS == 1 ? n_pair_end : P.counters[0]);
    const uint32_t key_bias = (uint32_t)CLS * T.n_pieces;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const bool producer = threadIdx.x == 0;
    if (producer) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(&s_empty[0], COUNT_WARPS);
        mbar_init(&s_empty[1], COUNT_WARPS);
    }
    __syncthreads();
    const unsigned long long n_chunks = (r_end - r_begin + COUNT_CHUNK - 1) / COUNT_CHUNK;
    auto issue = [&](unsigned long long chunk, uint32_t stage) {  // producer only
        const unsigned long long base = r_begin + chunk * COUNT_CHUNK;
        const unsigned long long end = min(base + (unsigned long long)COUNT_CHUNK, r_end);
        const ChunkTables c = chunk_tables(P, base, end, key_bias);
        if (c.staged) {
            const uint32_t ebytes = (c.e1 - c.e0) * (uint32_t)sizeof(StabEntry), obytes = c.noff * 4u;
            mbar_expect_tx(&s_full[stage], ebytes + obytes);
            bulk_g2s(s_ent[stage], T.stab + c.e0, ebytes, &s_full[stage]);
            bulk_g2s(s_off[stage], T.stab_off + c.pa, obytes, &s_full[stage]);
        } else {
            mbar_arrive(&s_full[stage]);
        }
    };
    if (producer && blockIdx.x < n_chunks) issue(blockIdx.x, 0);
    uint32_t j = 0;
    uint32_t my_slots = 0;  // GENERAL: ext slots read (byte accounting)
    for (unsigned long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, ++j) {
        const uint32_t stage = j & 1u;
        if (producer && chunk + gridDim.x < n_chunks) {
            const uint32_t ns = (j + 1u) & 1u;
            if (j + 1u >= 2u) mbar_wait(&s_empty[ns], (((j + 1u) >> 1) - 1u) & 1u);
            issue(chunk + gridDim.x, ns);
        }
        const unsigned long long base = r_begin + chunk * COUNT_CHUNK;
        const unsigned long long end = min(base + (unsigned long long)COUNT_CHUNK, r_end);
        const ChunkTables ct = chunk_tables(P, base, end, key_bias);
        if (lane == 0) mbar_wait(&s_full[stage], (j >> 1) & 1u);  // one poller per warp
        __syncwarp();
        const StabEntry *ent = ct.staged ? s_ent[stage] - ct.e0 : T.stab;
        const uint32_t *off = ct.staged ? s_off[stage] - ct.pa : T.stab_off;
        uint32_t(*ctr)[4] = ct.staged ? s_ctr[stage] - ct.e0 : nullptr;
        // software pipeline: the next iteration's key and record are loaded before this one is used
        unsigned long long i = base + (unsigned long long)(warp * COUNT_ITEMS * 32u) + lane;
        uint32_t key_n = 0xFFFFFFFFu;
        ReadRec rec_n = make_uint2(0u, 0u);
        if (i < end) {
            key_n = __ldcs(P.g_key + i);
            rec_n = __ldcs(P.g_rec + i);
        }
        for (int it = 0; it < COUNT_ITEMS; ++it, i += 32u) {
            if (i - lane >= end) break;  // warp-uniform
            const bool act = i < end;
            const uint32_t key = key_n;
            const ReadRec rec = rec_n;
            key_n = 0xFFFFFFFFu;
            if (it + 1 < COUNT_ITEMS && i + 32u < end) {
                key_n = __ldcs(P.g_key + i + 32u);
                rec_n = __ldcs(P.g_rec + i + 32u);
            }
            uint32_t so = 0, nent = 0;
            if (act) {
                const uint32_t piece = key - key_bias;
                so = off[piece];
                nent = off[piece + 1] - so;
            }
            // runs of equal keys: head lane, lane mask and end of the run
            const uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
            const bool head = lane == 0 || key != prev;
            const unsigned heads = __ballot_sync(0xFFFFFFFFu, head);
            const bool single_run = heads == 1u;
            const unsigned above = lane == 31 ? 0u : (heads >> (lane + 1));
            const uint32_t run_end = above ? lane + (uint32_t)__ffs(above) : 32u;  // first lane of the next run
            const unsigned run_mask = (run_end == 32u ? 0xFFFFFFFFu : ((1u << run_end) - 1u)) & ~((1u << lane) - 1u);
            const uint32_t max_ent = __reduce_max_sync(0xFFFFFFFFu, nent);
            if (!GENERAL) {
                const int32_t s0 = (int32_t)rec.x, e0 = s0 + (int32_t)(rec.y & 0xFFFFu);
                for (uint32_t q = 0; q < max_ent; ++q) {
                    const bool has = q < nent;
                    uint32_t cm = 0, ml = 0, slot_off = 0, x = 0;
                    if (has) {
                        const StabEntry en = ent[so + q];  // SE_FAST by construction of the class
                        slot_off = en.slot_off;
                        x = en.x;
                        cm = eval_fast(en, s0, e0, ml);
                        if (cm != 0 && en.K > 2) {  // dense class table with more than 3 classes
                            atomicAdd(P.slot_cnt + slot_off + (cm - 1u), 1ull);
                            cm = 0x80000000u;
                        }
                    }
                    aggregate_runs(P, lane, head && has, run_mask, run_end, single_run, cm, ml, slot_off, x,
                                   ctr ? ctr[so + q] : nullptr);
                }
            } else {
                ReadView rv;
                rv.b0 = rv.b1 = rv.b2 = make_int2(0, 0);
                rv.n = 1;
                rv.ridx = 0;
                rv.blk = nullptr;
                uint32_t strand = 0, total_len = 0;
                if (CLS == 1) {
                    if (act) {  // two blocks, relative to the first position of the piece
                        const int32_t ppos = __ldg(&T.pinfo[key - key_bias].pos);
                        unpack_pair(rec, ppos, rv.b0, rv.b1, strand);
                        rv.n = 2;
                        total_len = (uint32_t)(rv.b0.y - rv.b0.x) + (uint32_t)(rv.b1.y - rv.b1.x);
                    }
                } else if (act) {
                    // one pass over the ext entry: header and up to three blocks go to registers
                    const int2 *entry = P.ext + rec.x;  // 16-byte aligned
                    const int4 q0 = __ldg(reinterpret_cast<const int4 *>(entry));
                    const int2 hd = make_int2(q0.x, q0.y);
                    rv.b0 = make_int2(q0.z, q0.w);
                    const uint32_t nn = rec.y;  // kept blocks, also stored in the sorted record
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
                for (uint32_t q = 0; q < max_ent; ++q) {
                    const bool has = q < nent;
                    uint32_t cm = 0, ml = 0, slot_off = 0, x = 0;
                    if (has) {
                        const StabEntry en = ent[so + q];
                        slot_off = en.slot_off;
                        x = en.x;
                        const EvHdr h = T.hdr[x];
                        const uint64_t c64 =
                            (CLS == 1 && (en.meta & SE_FAST))
                                ? eval_pair2(en, h, T.seg + h.seg_off, T.fmask + h.form_off, rv.b0, rv.b1, total_len, ml)
                                : eval_pair(P, h, T.seg + h.seg_off, T.fmask + h.form_off, rv, strand, total_len, ml);
                        cm = (uint32_t)c64;
                        if (c64 != 0 && en.K > 2) {
                            bump_class(P.slot_cnt + h.slot_off, P.slot_mask + h.slot_off, h.slot_cap,
                                       (h.flags & EV_HASH) != 0, c64);
                            cm = 0x80000000u;
                        }
                    }
                    aggregate_runs(P, lane, head && has, run_mask, run_end, single_run, cm, ml, slot_off, x,
                                   ctr ? ctr[so + q] : nullptr);
                }
            }
        }
        __syncwarp();
        // the last warp of the block to finish this chunk flushes the stage's shared counters to
        // global memory; the stage is released (empty barrier) only after the flush, so the
        // counters and the staged entries they refer to cannot be overwritten before
        uint32_t last = 0;
        if (lane == 0) {
            __threadfence_block();
            last = atomicAdd(&s_done[stage], 1u) == (uint32_t)(COUNT_WARPS - 1);
        }
        last = __shfl_sync(0xFFFFFFFFu, last, 0);
        if (last) {
            __threadfence_block();
            if (ct.staged) {
                const uint32_t nctr = (ct.e1 - ct.e0) * 4u;
                uint32_t *flat = &s_ctr[stage][0][0];
                for (uint32_t i = lane; i < nctr; i += 32u) {
                    const uint32_t v = flat[i];
                    if (!v) continue;
                    flat[i] = 0u;
                    const StabEntry &en = s_ent[stage][i >> 2];
                    if ((i & 3u) == 3u) atomicAdd(P.bases + en.x, (unsigned long long)v);
                    else atomicAdd(P.slot_cnt + en.slot_off + (i & 3u), (unsigned long long)v);
                }
            }
            __syncwarp();
            if (lane == 0) s_done[stage] = 0u;
            __syncwarp();
        }
        if (lane == 0) mbar_arrive(&s_empty[stage]);  // this warp no longer reads the stage
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
