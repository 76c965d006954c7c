// Hand-written sm_100a kernels of the count + solve hot path.
//
//   lsq_group_sum_kernel / lsq_super_scan_kernel
//                      block-offset index of a compact read set (blocks per read -> first block of
//                      every group of 128 reads); replaces the CSR offsets of the wide form
//   lsq_scan_kernel    the whole per-read work in ONE pass over the packed reads: per-block covered-
//                      region filter + per-read interval list (count/count.cpp:319-324;
//                      jsc/util/interval_list.hpp:396-422, 462-503), candidate lookup (the read index
//                      of count.cpp:346-364 and the scan of :421-464 become one table lookup), and the
//                      tally of the read's SIGNATURE (lsq_device_fn.cuh: signature routing)
//   lsq_expand_kernel  every distinct signature, once: Read_single::build walk, compatibility,
//                      validity on a representative read (common/read.h:44-79, 204-274;
//                      count.cpp:438-482) -> class histogram and matched bases of the events
//   lsq_general_kernel the few reads whose outcome is not a function of their signature (tie-break
//                      candidates, partial matches near the 0.98 rule, three and more kept blocks):
//                      the same walk on the read itself
//   lsq_finalize_kernel per-event n_valid and per-form tallies from the class histogram
//                      (count.cpp:467-482)
//   lsq_solve_kernel   EM on the class histogram + RPKM + log-likelihood, FP64 without FMA
//                      contraction (common/read.h:592-660; solve/solve.cpp:796-826)
#pragma once
#include "lsq_device_fn.cuh"

namespace lsq {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ---------------------------------------------------------------------------------------------
// One launch clears every per-run scratch region (class counters, signature counters, hash, ...).
// ---------------------------------------------------------------------------------------------
struct ClearList {
    unsigned long long *p[8];
    unsigned long long n8[8];  // 8-byte units
    int n;
};
__global__ void __launch_bounds__(256) lsq_clear_kernel(ClearList L) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < L.n; ++k)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.n8[k]; i += stride) L.p[k][i] = 0ull;
}

// ---------------------------------------------------------------------------------------------
// Block-offset index of a compact read set: grp_base[g] = blocks of the reads before group g inside its
// super-group of SCAN_SUPER groups, sup_base[s] = blocks before super-group s.  One warp sums one group
// (one 32-bit load of four block counts per lane), one block scans one super-group; a single block then
// scans the super-group totals.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_SUPER) lsq_group_sum_kernel(const uint8_t *__restrict__ blk_cnt, uint64_t n_reads,
                                                                  uint32_t n_groups, uint32_t *__restrict__ grp_base,
                                                                  uint32_t *__restrict__ sup_total) {
    __shared__ uint32_t s_sum[SCAN_SUPER];
    __shared__ uint32_t s_wtot[SCAN_SUPER / 32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // phase 1: warp w of the block sums groups w, w + 8, ... of this super-group
    for (uint32_t gi = warp; gi < SCAN_SUPER; gi += SCAN_SUPER / 32) {
        const uint32_t g = blockIdx.x * SCAN_SUPER + gi;
        uint32_t v = 0;
        if (g < n_groups) {
            const uint64_t r0 = (uint64_t)g * SCAN_GROUP + 4u * lane;
            if (r0 < n_reads) {
                uint32_t c4 = __ldg(reinterpret_cast<const uint32_t *>(blk_cnt + r0));  // padded to 4 reads
                if (r0 + 4 > n_reads) c4 &= 0xFFFFFFFFu >> (8u * (uint32_t)(r0 + 4 - n_reads));
                v = __vsadu4(c4, 0u);  // sum of the four bytes
            }
        }
        v = __reduce_add_sync(0xFFFFFFFFu, v);
        if (lane == 0) s_sum[gi] = v;
    }
    __syncthreads();
    // phase 2: exclusive scan of the 256 group sums
    const uint32_t mine = s_sum[tid];
    uint32_t incl = mine;
#pragma unroll
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31u) s_wtot[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += s_wtot[w];
    const uint32_t g = blockIdx.x * SCAN_SUPER + tid;
    if (g < n_groups) grp_base[g] = before + incl - mine;
    if (tid == SCAN_SUPER - 1) sup_total[blockIdx.x] = before + incl;
}
// exclusive scan of sup_total[0..n) in place (single block)
__global__ void __launch_bounds__(1024) lsq_super_scan_kernel(uint32_t *__restrict__ sup, uint32_t n) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024u) {
        const uint32_t i = base + tid;
        const uint32_t mine = i < n ? sup[i] : 0u;
        uint32_t incl = mine;
#pragma unroll
        for (uint32_t d = 1; d < 32u; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31u) s_w[warp] = incl;
        __syncthreads();
        uint32_t before = s_carry;
        for (uint32_t w = 0; w < warp; ++w) before += s_w[w];
        if (i < n) sup[i] = before + incl - mine;
        __syncthreads();
        if (tid == 1023u) s_carry = before + incl;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Scan kernel.  A warp owns a group of 128 consecutive reads at a time (grid-stride over the groups);
// warps share nothing, so there is no block-wide barrier anywhere: the warps of an SM drift apart and
// cover each other's L2 latencies.  The loops are rolled on purpose: the whole kernel is ~1 k SASS
// instructions and stays resident in the instruction cache.
// Phase A: lane l owns reads 4l..4l+3 of the group: ONE 32-bit load brings their four block counts, one
//          more their chromosome/strand ids; a warp scan of the counts gives every read's first block.
// Phase S: reads with one raw block (4 in 5), SCAN_FLIGHT per lane and round: block -> bucket entry (ONE 8-byte
//          gather names the piece, lsq_types.cuh:bucket_rank) -> PieceInfo (ONE 16-byte gather: covered
//          test, candidate test, cut list) -> route_single -> ONE 32-bit reduction on the signature
//          counter.  Two dependent gathers and one reduction per read: the floor of the kernel is the SM's
//          rate of scattered L1TEX operations (one lane-operation per cycle, tools/microbench_gather.cu),
//          not HBM.
// Phase D: reads with two raw blocks are first listed densely in the warp's shared-memory slice, then
//          taken one per lane and round: the same two gathers for both blocks (in flight together) ->
//          route_double (pair signature -> hash; fused or half-kept reads -> single signature).
// Reads with more blocks, chimeric reads and every read the routing functions decline are listed for
// lsq_general_kernel.  Rare lookup cases (crowded bucket, position outside the bucket range) take
// lookup_cold, out of line.
// A streamed read set (lsq_set_reads with LSQ_READS_STREAM) is consumed while the copy engine is still
// delivering it: lane 0 polls the landed-blocks word before the group's blocks are touched.
// ---------------------------------------------------------------------------------------------
#ifndef LSQ_SCAN_BLOCK
#define LSQ_SCAN_BLOCK 256
#endif
#ifndef LSQ_SCAN_MINB
#define LSQ_SCAN_MINB 5  // resident CTAs per SM (the kernel needs 48 registers; measured best of 3..6)
#endif
constexpr int SCAN_BLOCK = LSQ_SCAN_BLOCK;
constexpr int SCAN_MINB = LSQ_SCAN_MINB;
constexpr int SCAN_WARPS = SCAN_BLOCK / 32;
constexpr int SCAN_ITEMS = 4;     // reads per lane
#ifndef LSQ_SCAN_TRIPLES
// 1: three-block reads get a signature route of their own (phase T) instead of the general path.  Correct and tested,
// but off: at C4 it takes 1.0 ms off lsq_general_kernel and puts 1.2 ms on the scan (one more dependent chain of
// lookups per group for ~4 reads of 128), at C2 it costs 0.13 ms for nothing (profiles/r02_elimination_experiments.txt).
#define LSQ_SCAN_TRIPLES 0
#endif
#ifndef LSQ_SCAN_FLIGHT
#define LSQ_SCAN_FLIGHT 2
#endif
constexpr int SCAN_FLIGHT = LSQ_SCAN_FLIGHT;  // single-block reads whose lookup chains a lane keeps in flight together (1, 2 or 4)
static_assert(SCAN_ITEMS % SCAN_FLIGHT == 0, "rounds of SCAN_FLIGHT reads");
constexpr uint32_t SCAN_GEN_FLUSH = 64;                          // the warp's general list is flushed when it holds this many reads
constexpr uint32_t SCAN_GEN_CAP = SCAN_GEN_FLUSH + SCAN_GROUP;   // ... so a whole group always fits
static_assert(SCAN_ITEMS * 32 == (int)SCAN_GROUP, "a lane owns four consecutive reads of the group");

// The piece of position x on file chromosome c by the full search (find_piece): positions below zero or past the
// last bucket, crowded sub-buckets.  Out of line: the hot loop stays small.
__device__ __noinline__ uint32_t lookup_cold(const ScanParams &P, uint32_t c, int32_t x) {
    PieceInfo pi;
    const uint32_t p = find_piece(P.T, __ldg(P.chrom_remap + c), x, pi);
    return p == 0xFFFFFFFFu ? PIECE_NONE : p;
}

// The hot lookup, in two stages so that a lane can have the bucket entries of two blocks in flight together.
// Stage 1: what to load.  how = 0: no piece can hold the block (nil block, chromosome without annotation);
// 1: bucket entry at `slot`; 2: lookup_cold.
struct Probe {
    uint32_t slot, lo, how;
};
__device__ __forceinline__ Probe probe_of(const ScanParams &P, int2 b, uint4 cm) {
    Probe q;
    const uint32_t bk = (uint32_t)b.x >> P.T.bp_shift;
    q.lo = cm.x;
    q.slot = cm.z + bk;
    q.how = (!(b.x < b.y) || cm.x == cm.y) ? 0u : ((b.x < 0 || bk >= cm.w) ? 2u : 1u);
    return q;
}
// Stage 2: the piece from the bucket entry e (a crowded bucket names its sub-bucket: one more 8-byte load).
__device__ __forceinline__ uint32_t piece_of(const ScanParams &P, int2 b, uint32_t c, Probe q, uint2 e) {
    if (q.how != 1u) return q.how == 0u ? PIECE_NONE : lookup_cold(P, c, b.x);
    uint32_t sh = P.T.bp_shift, xo = (uint32_t)b.x & ((1u << sh) - 1u);
    if (bucket_many(e.y)) {
        sh = bucket_sub_shift(sh);
        e = __ldg(P.T.bp_sub + (e.y & 0x1FFFFFFFu) + (xo >> sh));
        xo &= (1u << sh) - 1u;
        if (bucket_many(e.y)) return lookup_cold(P, c, b.x);
    }
    const uint32_t n = e.x + bucket_rank(sh, e.y, xo);
    return n != q.lo ? n - 1u : PIECE_NONE;
}

// Block i of the read set: the compact layout (start i32 + length u8) without any format test.
template <bool COMPACT>
__device__ __forceinline__ int2 scan_block(const ScanParams &P, uint32_t i) {
    // cached loads (L1): a lane's four reads, and its neighbours', share 128-byte lines, so three rounds of four hit L1
    // instead of paying an HBM round trip each (a streaming, evict-first load would)
    if (COMPACT) {
        const int32_t s = __ldg(P.blk_start + i);
        return make_int2(s, s + (int32_t)__ldg(P.blk_len8 + i));
    }
    return load_block(P, i);
}

// Signature hash: open addressing, linear probing; a slot's key and count share one 16-byte line segment, so a resident
// key costs one 8-byte load (and one 64-bit reduction to count it).  Returns the key's slot, or 0xFFFFFFFF when
// HASH_MAX_PROBES slots were taken by other keys.
__device__ __forceinline__ uint32_t sig_hash_slot(const ScanParams &P, uint64_t key) {
    uint32_t i = hash_mask(key) & P.hash_mask;
#pragma unroll 1
    for (uint32_t probe = 0; probe < HASH_MAX_PROBES; ++probe) {
        unsigned long long *kp = &P.hash[i].x;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(kp);
        if (cur == 0ull) cur = atomicCAS(kp, 0ull, (unsigned long long)key);
        if (cur == 0ull || cur == key) return i;
        i = (i + 1u) & P.hash_mask;
    }
    return 0xFFFFFFFFu;
}
__device__ __forceinline__ void sig_hash_count(const ScanParams &P, uint32_t slot, uint32_t total) {
    atomicAdd(&P.hash[slot].y, 1ull);
    if (total != P.len0) atomicAdd(P.hash_dev + slot, (unsigned long long)(long long)((int)total - (int)P.len0));
}

// A read with three raw blocks (rare outside long-read sets): out of line, so that its registers do not weigh on the
// one- and two-block paths.
template <bool COMPACT>
__device__ __noinline__ uint32_t scan_triple(const ScanParams &P, uint32_t base, uint32_t c, uint32_t &sig, uint64_t &key,
                                             uint32_t &total) {
    const PieceInfo *const pinfo = P.T.pinfo;
    sig = 0;
    key = 0;
    total = 0;
    int2 b[3];
    Probe q[3];
    uint2 e[3];
    uint32_t p[3];
    uint4 raw[3];
    const uint4 cm = __ldg(P.file_meta + c);
#pragma unroll
    for (int j = 0; j < 3; ++j) b[j] = scan_block<COMPACT>(P, base + (uint32_t)j);
#pragma unroll
    for (int j = 0; j < 3; ++j) q[j] = probe_of(P, b[j], cm);
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = q[j].how == 1u ? __ldg(P.T.bp_bkt + q[j].slot) : make_uint2(0u, 0u);
    bool all_kept = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        p[j] = piece_of(P, b[j], c, q[j], e[j]);
        raw[j] = p[j] != PIECE_NONE ? __ldg(reinterpret_cast<const uint4 *>(pinfo + p[j])) : make_uint4(0u, 0u, 0u, 0u);
        all_kept = all_kept && p[j] != PIECE_NONE && (int32_t)raw[j].y >= b[j].y;
    }
    // the signature route needs all three blocks kept, in order, apart; anything else takes the general path
    uint32_t rt = RT_GENERAL;
    if (all_kept && b[1].x > b[0].y && b[2].x > b[1].y) {
        if (!(raw[0].z & PIECE_CAND)) {
            rt = RT_NONE;
        } else {
            total = (uint32_t)(b[0].y - b[0].x) + (uint32_t)(b[1].y - b[1].x) + (uint32_t)(b[2].y - b[2].x);
            uint32_t k2;
            rt = route_triple(P.T, b[0], b[1], b[2], p[0], raw[0], p[1], raw[1], raw[2], key, k2);
            sig = (p[2] << 4) | k2;
        }
    }
    return rt;
}

// COMPACT: the read set is in the compact layout throughout (blk_cnt + index, read_cs8, blk_start + blk_len8) --
// what the host front end and the benchmark produce; the other instantiation takes any mix of layouts.
template <bool COMPACT>
__global__ void __launch_bounds__(SCAN_BLOCK, SCAN_MINB) lsq_scan_kernel(const __grid_constant__ ScanParams P) {
    __shared__ uint2 s_dbl[SCAN_WARPS][SCAN_GROUP];  // per warp: its two-block reads {first block, chromosome id | group-local index << 24}
    __shared__ uint2 s_tri[SCAN_WARPS][LSQ_SCAN_TRIPLES ? SCAN_GROUP : 1];    // per warp: its three-block reads, same format
    __shared__ uint2 s_gen[SCAN_WARPS][SCAN_GEN_CAP];  // per warp: reads for the general kernel {read, first block}
    __shared__ uint32_t s_ngen[SCAN_WARPS];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t n_groups = (P.n_reads + SCAN_GROUP - 1) / SCAN_GROUP;
    uint32_t c_single = 0, c_pair = 0, c_triple = 0, c_general = 0, c_full = 0;  // statistics, flushed once per warp
    const PieceInfo *const pinfo = P.T.pinfo;
    const bool no_single = (P.flags & SCAN_NO_SINGLE_SIG) != 0u, no_triple = !LSQ_SCAN_TRIPLES || (P.flags & SCAN_NO_TRIPLE_SIG) != 0u;

    // general reads are parked in the warp's shared-memory slice (a shared atomic returns in ~30 cycles); the warp
    // reserves slots in the global list only when SCAN_GEN_FLUSH reads have gathered -- a returning global atomic on
    // the one list counter is a full L2 round trip for the whole warp
    auto to_general = [&](uint64_t r, uint32_t base) {
#ifndef LSQ_EXP_NO_GENERAL  /* timing experiments only */
        s_gen[warp][atomicAdd(&s_ngen[warp], 1u)] = make_uint2((uint32_t)r, base);
#endif
    };
    auto flush_general = [&]() {  // called by the whole warp
        const uint32_t n_gen = s_ngen[warp];
        uint32_t slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(P.gen_n, n_gen);
        slot0 = __shfl_sync(0xFFFFFFFFu, slot0, 0);
        for (uint32_t i = lane; i < n_gen; i += 32u)
            if (slot0 + i < P.gen_cap) P.gen_list[slot0 + i] = s_gen[warp][i];
        c_general += lane == 0 ? n_gen : 0u;
        __syncwarp();
        if (lane == 0) s_ngen[warp] = 0u;
    };
    if (lane == 0) s_ngen[warp] = 0u;
    __syncwarp();
    auto tally = [&](uint32_t rt, uint32_t sig, uint64_t key, uint32_t total, uint64_t r, uint32_t base) {
        if (rt == RT_SINGLE) {
#ifndef LSQ_EXP_NO_SIG_RED  /* timing experiments only */
            atomicAdd(P.sig_cnt + sig, 1u);
#endif
            if (total != P.len0) atomicAdd(P.sig_dev + sig, (unsigned long long)(long long)((int)total - (int)P.len0));
            ++c_single;
        } else if (rt == RT_PAIR) {
            const uint32_t slot = sig_hash_slot(P, key);
            if (slot != 0xFFFFFFFFu) {
                sig_hash_count(P, slot, total);
                ++c_pair;
            } else {
                ++c_full;
                rt = RT_GENERAL;
            }
        } else if (rt == RT_TRIPLE) {  // key: the prefix (pair key of blocks 0 and 1); sig: piece 2 << 4 | cut category
            const uint32_t ps = sig_hash_slot(P, key);  // interned, not counted
            const uint32_t slot = ps != 0xFFFFFFFFu ? sig_hash_slot(P, triple_key(ps, sig >> 4, sig & 15u)) : 0xFFFFFFFFu;
            if (slot != 0xFFFFFFFFu) {
                sig_hash_count(P, slot, total);
                ++c_triple;
            } else {
                ++c_full;
                rt = RT_GENERAL;
            }
        }
        if (rt == RT_GENERAL) to_general(r, base);
    };

#pragma unroll 1
    for (uint64_t g = (uint64_t)blockIdx.x * SCAN_WARPS + warp; g < n_groups; g += (uint64_t)gridDim.x * SCAN_WARPS) {
        const uint64_t g0 = g * SCAN_GROUP, g1 = min(g0 + (uint64_t)SCAN_GROUP, P.n_reads);
        const uint64_t r0 = g0 + 4u * lane;
        // ---- phase A --------------------------------------------------------------------------
        uint32_t c4 = 0, v4 = 0;  // block counts and one-byte ids of my four reads (compact form)
        uint32_t o;               // first block of my first read
        uint32_t g_end;           // one past the group's last block (warp-uniform)
        uint32_t dmask = 0;       // bit 8k+7: my read k has two raw blocks and is not chimeric
        uint32_t tmask = 0;       // ... three raw blocks
        const bool compact_cnt = COMPACT || P.blk_cnt != nullptr, compact_cs = COMPACT || P.read_cs8 != nullptr;
        if (compact_cs && r0 < g1) v4 = __ldcs(reinterpret_cast<const uint32_t *>(P.read_cs8 + r0));
        if (compact_cnt) {
            if (r0 < g1) {
                c4 = __ldcs(reinterpret_cast<const uint32_t *>(P.blk_cnt + r0));
                if (r0 + 4 > g1) c4 &= 0xFFFFFFFFu >> (8u * (uint32_t)(r0 + 4 - g1));
            }
            const uint32_t mine = __vsadu4(c4, 0u);
            uint32_t incl = mine;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t gbase = __ldg(P.sup_base + (g / SCAN_SUPER)) + __ldg(P.grp_base + g);
            g_end = gbase + __shfl_sync(0xFFFFFFFFu, incl, 31);
            o = gbase + incl - mine;
            if (compact_cs) {  // bytes of c4 equal to 2 (exact zero-byte test on c4 ^ 0x02020202), not chimeric
                const uint32_t t = c4 ^ 0x02020202u, t3 = c4 ^ 0x03030303u;
                dmask = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu) & ~v4 & 0x80808080u;
                if (!no_triple) tmask = ~(((t3 & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t3 | 0x7F7F7F7Fu) & ~v4 & 0x80808080u;
            }
        } else {
            o = r0 < g1 ? __ldcs(P.blk_off + r0) : 0u;
            g_end = __ldg(P.blk_off + g1);
        }
        if (P.progress != nullptr) {
            // streamed read set: the copy engine is still filling the block arrays; wait until the group's
            // last block has landed (the per-read arrays were complete before the launch)
            if (lane == 0) {
                uint32_t spins = 0;
                while (*reinterpret_cast<const volatile uint32_t *>(P.progress) < g_end) {
                    __nanosleep(400);
                    if (++spins > 100000000u) __trap();  // ~40 s without progress: the copy died; fail instead of hanging
                }
                __threadfence();
            }
            __syncwarp();
        }
        // where my two-block reads go in the warp's list: known before the loop when both per-read arrays are compact
        uint32_t nd = __popc(dmask), n_dbl = 0, dpos = 0, n_tri = 0, tpos = 0;
        if (compact_cnt && compact_cs) {  // both counts in one scan: doubles in the low half, triples in the high half
            const uint32_t mine = nd | ((uint32_t)__popc(tmask) << 16);
            uint32_t incl = mine;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
            n_dbl = tot & 0xFFFFu;
            n_tri = tot >> 16;
            dpos = (incl - mine) & 0xFFFFu;
            tpos = (incl - mine) >> 16;
        }
        // ---- phase S --------------------------------------------------------------------------
        // SCAN_FLIGHT reads per round: their block loads, then their bucket gathers, then their piece-record gathers
        // are issued together, so a lane has SCAN_FLIGHT dependent chains in flight instead of one
#pragma unroll 1
        for (uint32_t k0 = 0; k0 < (uint32_t)SCAN_ITEMS; k0 += SCAN_FLIGHT) {
            uint32_t base[SCAN_FLIGHT], c[SCAN_FLIGHT];
            bool act[SCAN_FLIGHT];
#pragma unroll
            for (int u = 0; u < SCAN_FLIGHT; ++u) {
                const uint32_t k = k0 + (uint32_t)u;
                const uint64_t r = r0 + k;
                uint32_t rn;
                bool plain;
                if (compact_cnt) rn = (c4 >> (8u * k)) & 0xFFu;
                else rn = (r < g1 ? __ldcs(P.blk_off + r + 1) : o) - o;
                if (compact_cs) {
                    const uint32_t v = v4 >> (8u * k);
                    c[u] = v & 0x3Fu;
                    plain = !(v & 0x80u);
                } else {
                    const uint32_t cs = r < g1 ? __ldcs(P.read_cs + r) : 0u;
                    c[u] = cs & 0xFFFFu;
                    plain = !(cs & 0x80000000u);
                }
                base[u] = o;
                o += rn;
                act[u] = rn == 1u && plain;
                if (!act[u]) {
                    if (rn == 2u && plain) {
                        const uint2 d = make_uint2(base[u], c[u] | ((4u * lane + k) << 24));
                        if (compact_cnt && compact_cs) {
                            s_dbl[warp][dpos++] = d;
                        } else {  // mixed layouts: sparse list, compacted below
                            s_dbl[warp][4u * lane + nd] = d;
                            ++nd;
                        }
                    } else if (rn == 3u && plain && compact_cnt && compact_cs && !no_triple) {
                        s_tri[warp][tpos++] = make_uint2(base[u], c[u] | ((4u * lane + k) << 24));
                    } else if (rn != 0u) {
                        to_general(r, base[u]);  // more blocks, chimeric reads
                    }
                }
            }
            int2 b[SCAN_FLIGHT];
            Probe q[SCAN_FLIGHT];
            uint2 e[SCAN_FLIGHT];
            uint32_t p[SCAN_FLIGHT];
            uint4 raw[SCAN_FLIGHT];
#pragma unroll
            for (int u = 0; u < SCAN_FLIGHT; ++u) b[u] = act[u] ? scan_block<COMPACT>(P, base[u]) : make_int2(0, 0);
#pragma unroll
            for (int u = 0; u < SCAN_FLIGHT; ++u) q[u] = probe_of(P, b[u], __ldg(P.file_meta + c[u]));
#pragma unroll
            for (int u = 0; u < SCAN_FLIGHT; ++u) e[u] = q[u].how == 1u ? __ldg(P.T.bp_bkt + q[u].slot) : make_uint2(0u, 0u);
#pragma unroll
            for (int u = 0; u < SCAN_FLIGHT; ++u) {
                p[u] = piece_of(P, b[u], c[u], q[u], e[u]);
#ifdef LSQ_EXP_NO_PINFO  /* timing experiments only */
                if (p[u] != PIECE_NONE) ++c_full;
                p[u] = PIECE_NONE;
#endif
                raw[u] = p[u] != PIECE_NONE ? __ldg(reinterpret_cast<const uint4 *>(pinfo + p[u])) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < SCAN_FLIGHT; ++u) {
                // kept: inside a covered region; candidate of some event
                if (p[u] == PIECE_NONE || (int32_t)raw[u].y < b[u].y || !(raw[u].z & PIECE_CAND)) continue;
                uint32_t sig = 0;
                const uint32_t rt = no_single ? RT_GENERAL : route_single(P.T, b[u], raw[u], sig);
                tally(rt, sig, 0ull, (uint32_t)(b[u].y - b[u].x), r0 + k0 + (uint32_t)u, base[u]);
            }
        }
        // ---- phase D --------------------------------------------------------------------------
        if (!(compact_cnt && compact_cs)) {  // mixed layouts: compact the sparse list (every lane wrote at 4 * lane)
            uint32_t incl = nd;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            n_dbl = __shfl_sync(0xFFFFFFFFu, incl, 31);
            uint2 mine[SCAN_ITEMS];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) mine[j] = s_dbl[warp][4u * lane + j];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j)
                if ((uint32_t)j < nd) s_dbl[warp][incl - nd + j] = mine[j];
        }
        __syncwarp();
#pragma unroll 1
        for (uint32_t i = lane; i < n_dbl; i += 32u) {
            const uint2 d = s_dbl[warp][i];
            const uint32_t base = d.x, c = d.y & 0xFFFFu;
            const uint64_t r = g0 + (d.y >> 24);
            const int2 b0 = scan_block<COMPACT>(P, base), b1 = scan_block<COMPACT>(P, base + 1u);
            const uint4 cm = __ldg(P.file_meta + c);
            const Probe q0 = probe_of(P, b0, cm), q1 = probe_of(P, b1, cm);
            const uint2 e0 = q0.how == 1u ? __ldg(P.T.bp_bkt + q0.slot) : make_uint2(0u, 0u);
            const uint2 e1 = q1.how == 1u ? __ldg(P.T.bp_bkt + q1.slot) : make_uint2(0u, 0u);
            const uint32_t p0 = piece_of(P, b0, c, q0, e0), p1 = piece_of(P, b1, c, q1, e1);
            const uint4 raw0 = p0 != PIECE_NONE ? __ldg(reinterpret_cast<const uint4 *>(pinfo + p0)) : make_uint4(0u, 0u, 0u, 0u);
            const uint4 raw1 = p1 != PIECE_NONE ? __ldg(reinterpret_cast<const uint4 *>(pinfo + p1)) : make_uint4(0u, 0u, 0u, 0u);
            const bool kept0 = p0 != PIECE_NONE && (int32_t)raw0.y >= b0.y, kept1 = p1 != PIECE_NONE && (int32_t)raw1.y >= b1.y;
            uint32_t sig, total;
            uint64_t key;
            const uint32_t rt = route_double(P, b0, b1, p0, raw0, kept0, p1, raw1, kept1, sig, key, total);
            tally(rt, sig, key, total, r, base);
        }
        // ---- phase T: three-block reads, one per lane and round ---------------------------------------------
#pragma unroll 1
        for (uint32_t i = lane; LSQ_SCAN_TRIPLES && i < n_tri; i += 32u) {
            const uint2 d = s_tri[warp][i];
            const uint32_t base = d.x, c = d.y & 0xFFFFu;
            const uint64_t r = g0 + (d.y >> 24);
            uint32_t sig, total;
            uint64_t key;
            const uint32_t rt = scan_triple<COMPACT>(P, base, c, sig, key, total);
            tally(rt, sig, key, total, r, base);
        }
        // ---- flush the warp's general reads when enough have gathered --------------------------------
        __syncwarp();
        if (s_ngen[warp] >= SCAN_GEN_FLUSH) flush_general();  // warp-uniform
        __syncwarp();  // the lists are rewritten by the next group
    }
    __syncwarp();
    if (s_ngen[warp] != 0u) flush_general();
    c_single = __reduce_add_sync(0xFFFFFFFFu, c_single);
    c_pair = __reduce_add_sync(0xFFFFFFFFu, c_pair);
    c_triple = __reduce_add_sync(0xFFFFFFFFu, c_triple);
    c_general = __reduce_add_sync(0xFFFFFFFFu, c_general);
    c_full = __reduce_add_sync(0xFFFFFFFFu, c_full);
    if (lane == 0) {
        if (c_single) atomicAdd(P.counters + CTR_SINGLE_SIG, (unsigned long long)c_single);
        if (c_pair) atomicAdd(P.counters + CTR_PAIR_SIG, (unsigned long long)c_pair);
        if (c_triple) atomicAdd(P.counters + CTR_TRIPLE_SIG, (unsigned long long)c_triple);
        if (c_general) atomicAdd(P.counters + CTR_GENERAL, (unsigned long long)c_general);
        if (c_full) atomicAdd(P.counters + CTR_HASH_FULL, (unsigned long long)c_full);
    }
}

// ---------------------------------------------------------------------------------------------
// Route kernel (multi-GPU): one pass over the local slice of the read set; every read is appended to the
// outbox of each rank that owns a candidate event at the piece of any of its blocks.  Same group / lane
// layout and the same two-gather lookup as the scan kernel, plus one 8-byte gather of the piece's owner
// mask.  Appends are reserved once per group of 128 reads: a warp scan per destination gives every lane
// its offsets, lane d reserves destination d's slots (all destinations in ONE round trip to L2), a second
// scan places the records.  Reads and blocks are reserved together (reads << 32 | blocks): the receiver
// pairs them by prefix sums, so both orders must agree.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long warp_incl_scan64(unsigned long long v, uint32_t lane) {
#pragma unroll
    for (uint32_t s = 1; s < 32u; s <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, v, s);
        if (lane >= s) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(SCAN_BLOCK, 4) lsq_route_kernel(const __grid_constant__ ScanParams P,
                                                                  const __grid_constant__ RouteParams R) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint64_t n_groups = (P.n_reads + SCAN_GROUP - 1) / SCAN_GROUP;
    const PieceInfo *const pinfo = P.T.pinfo;
#pragma unroll 1
    for (uint64_t g = (uint64_t)blockIdx.x * SCAN_WARPS + warp; g < n_groups; g += (uint64_t)gridDim.x * SCAN_WARPS) {
        const uint64_t g0 = g * SCAN_GROUP, g1 = min(g0 + (uint64_t)SCAN_GROUP, P.n_reads);
        const uint64_t r0 = g0 + 4u * lane;
        uint32_t c4 = 0, v4 = 0, o;
        const bool compact_cnt = P.blk_cnt != nullptr, compact_cs = P.read_cs8 != nullptr;
        if (compact_cnt) {
            if (r0 < g1) {
                c4 = __ldcs(reinterpret_cast<const uint32_t *>(P.blk_cnt + r0));
                if (r0 + 4 > g1) c4 &= 0xFFFFFFFFu >> (8u * (uint32_t)(r0 + 4 - g1));
            }
            const uint32_t mine = __vsadu4(c4, 0u);
            uint32_t incl = mine;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            o = __ldg(P.sup_base + (g / SCAN_SUPER)) + __ldg(P.grp_base + g) + incl - mine;
        } else {
            o = r0 < g1 ? __ldcs(P.blk_off + r0) : 0u;
        }
        if (compact_cs && r0 < g1) v4 = __ldcs(reinterpret_cast<const uint32_t *>(P.read_cs8 + r0));
        // ---- owners of my four reads ---------------------------------------------------------------
        uint32_t rn[SCAN_ITEMS], base[SCAN_ITEMS], cs[SCAN_ITEMS], mask[SCAN_ITEMS];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const uint64_t r = r0 + k;
            if (compact_cnt) rn[k] = (c4 >> (8 * k)) & 0xFFu;
            else rn[k] = (r < g1 ? __ldcs(P.blk_off + r + 1) : o) - o;
            if (compact_cs) cs[k] = expand_cs8((v4 >> (8 * k)) & 0xFFu);
            else cs[k] = r < g1 ? __ldcs(P.read_cs + r) : 0u;
            base[k] = o;
            o += rn[k];
        }
        // owner mask of block `at` of read k, stage by stage so that the four reads' lookups are in flight together
        auto owners_of = [&](const int2 b, const uint32_t c, const Probe q, const uint2 e) -> uint32_t {
            const uint32_t p = piece_of(P, b, c, q, e);
            // a piece without candidate events (no PIECE_CAND) has no stab entry, hence no owner: no need to look at its record
            return p == PIECE_NONE ? 0u : (uint32_t)__ldg(R.owner + p);  // at most ROUTE_MAX_RANKS = 16 ranks
        };
        {
            int2 b[SCAN_ITEMS];
            uint32_t c[SCAN_ITEMS];
            Probe q[SCAN_ITEMS];
            uint2 e[SCAN_ITEMS];
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                b[k] = rn[k] ? load_block(P, base[k]) : make_int2(0, 0);
                c[k] = ((rn[k] && (cs[k] & 0x80000000u)) ? __ldg(P.blk_cs + base[k]) : cs[k]) & 0xFFFFu;
            }
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) q[k] = probe_of(P, b[k], __ldg(P.file_meta + c[k]));
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) e[k] = (rn[k] && q[k].how == 1u) ? __ldg(P.T.bp_bkt + q[k].slot) : make_uint2(0u, 0u);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) mask[k] = rn[k] ? owners_of(b[k], c[k], q[k], e[k]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)  // the later blocks of spliced reads, one at a time
#pragma unroll 1
            for (uint32_t j = 1; j < rn[k]; ++j) {
                const int2 b = load_block(P, base[k] + j);
                const uint32_t c = ((cs[k] & 0x80000000u) ? __ldg(P.blk_cs + base[k] + j) : cs[k]) & 0xFFFFu;
                const Probe q = probe_of(P, b, __ldg(P.file_meta + c));
                const uint2 e = q.how == 1u ? __ldg(P.T.bp_bkt + q.slot) : make_uint2(0u, 0u);
                mask[k] |= owners_of(b, c, q, e);
            }
        uint32_t present = mask[0] | mask[1] | mask[2] | mask[3];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) present |= __shfl_xor_sync(0xFFFFFFFFu, present, d);
        if (present == 0u) continue;  // warp-uniform
        // what I send to destination d, as one half of a 64-bit word: reads << 20 | blocks (a group holds at most 128 reads
        // and 32 640 blocks), so ONE warp scan serves two destinations and both passes use its result
        auto my_half = [&](uint32_t d) {
            uint32_t v = 0u;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k)
                if ((mask[k] >> d) & 1u) v += (1u << 20) | rn[k];
            return v;
        };
        auto widen = [](uint32_t half) { return ((unsigned long long)(half >> 20) << 32) | (unsigned long long)(half & 0xFFFFFu); };
        // ---- pass 1: the group's totals per destination; lane d reserves destination d ----------------
        unsigned long long excl[ROUTE_MAX_RANKS / 2];  // my exclusive prefixes, two destinations per word
        unsigned long long my_total = 0ull;
#pragma unroll
        for (int w = 0; w < (int)ROUTE_MAX_RANKS / 2; ++w) {
            excl[w] = 0ull;
            if ((uint32_t)(2 * w) >= R.n_ranks || !((present >> (2 * w)) & 3u)) continue;  // warp-uniform
            const unsigned long long v = (unsigned long long)my_half(2 * w) | ((unsigned long long)my_half(2 * w + 1) << 32);
            const unsigned long long incl = warp_incl_scan64(v, lane);
            excl[w] = incl - v;
            const unsigned long long tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if ((lane >> 1) == (uint32_t)w) my_total = widen((lane & 1u) ? (uint32_t)(tot >> 32) : (uint32_t)tot);
        }
        unsigned long long my_rsv = 0ull;
        if (my_total) my_rsv = atomicAdd(R.cursor + lane, my_total);
        // ---- pass 2: place the records -----------------------------------------------------------------
#pragma unroll 1
        for (uint32_t pm = present; pm; pm &= pm - 1u) {
            const uint32_t d = (uint32_t)__ffs((int)pm) - 1u;
            unsigned long long ew = 0ull;
#pragma unroll
            for (int w = 0; w < (int)ROUTE_MAX_RANKS / 2; ++w)
                if ((d >> 1) == (uint32_t)w) ew = excl[w];
            const unsigned long long at = __shfl_sync(0xFFFFFFFFu, my_rsv, (int)d) + widen((d & 1u) ? (uint32_t)(ew >> 32) : (uint32_t)ew);
            const uint32_t v = my_half(d);
            if (v == 0u) continue;
            const RouteBox &B = R.box[d];
            size_t ri = (size_t)(at >> 32), bi = (size_t)(at & 0xFFFFFFFFull);
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                if (!((mask[k] >> d) & 1u)) continue;
                B.blk_cnt[ri] = (uint8_t)rn[k];
                if (B.read_cs8) B.read_cs8[ri] = (uint8_t)((cs[k] & 0x3Fu) | ((cs[k] >> 10) & 0x40u) | ((cs[k] >> 24) & 0x80u));
                else B.read_cs[ri] = cs[k];
                B.line_num[ri] = P.line_num ? __ldg(P.line_num + r0 + k) : R.line_base + (uint32_t)(r0 + k) + 1u;
                ++ri;
#pragma unroll 1
                for (uint32_t j = 0; j < rn[k]; ++j, ++bi) {
                    const int2 b = load_block(P, base[k] + j);
                    if (B.blk_start) {
                        B.blk_start[bi] = b.x;
                        if (B.blk_len8) B.blk_len8[bi] = (uint8_t)(b.y - b.x);
                        else B.blk_len[bi] = (uint16_t)(b.y - b.x);
                    } else {
                        B.blk[bi] = b;
                    }
                    if (B.blk_cs) B.blk_cs[bi] = (cs[k] & 0x80000000u) ? __ldg(P.blk_cs + base[k] + j) : cs[k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Class histogram updates
// ---------------------------------------------------------------------------------------------
// `n` reads of class `cm` for the event with header h (dense table: slot mask-1; K > 8: open-addressed hash)
__device__ __forceinline__ void bump_class(const ScanParams &P, const EvHdr &h, uint64_t cm, unsigned long long n) {
    unsigned long long *slot_cnt = P.slot_cnt + h.slot_off;
    if (!(h.flags & EV_HASH)) {
        atomicAdd(slot_cnt + (uint32_t)(cm - 1ull), n);
        return;
    }
    unsigned long long *slot_mask = P.slot_mask + h.slot_off;
    const uint32_t capm = h.slot_cap - 1u;
    uint32_t i = hash_mask(cm) & capm;
    while (true) {
        unsigned long long *kp = slot_mask + i;
        unsigned long long cur = *((volatile unsigned long long *)kp);
        if (cur == 0ull) cur = atomicCAS(kp, 0ull, (unsigned long long)cm);
        if (cur == 0ull || cur == cm) {
            atomicAdd(slot_cnt + i, n);
            return;
        }
        i = (i + 1u) & capm;
    }
}

// ---------------------------------------------------------------------------------------------
// Expansion: one thread per signature slot (single-block signatures, then the hash of the two-block
// ones).  A slot that was hit takes the walk of read.h:204-274 once, on its representative read, for
// every candidate event of its piece; events matched fully receive the slot's reads and bases.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) lsq_expand_kernel(const __grid_constant__ ScanParams P) {
    const Tables &T = P.T;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_hash = (uint64_t)P.hash_mask + 1u;
    ReadView rv;
    uint32_t piece;
    unsigned long long cnt, dev;
    if (i < T.n_sig) {
        cnt = P.sig_cnt[i];
        if (cnt == 0ull) return;
        dev = P.sig_dev[i];
        if (!sig_rep_single(T, (uint32_t)i, rv, piece)) return;
    } else if (i - T.n_sig < n_hash) {
        const ulonglong2 e = P.hash[i - T.n_sig];
        if (e.x == 0ull || e.y == 0ull) return;  // empty, or a prefix that only three-block reads use
        cnt = e.y;
        dev = P.hash_dev[i - T.n_sig];
        if (e.x & TRIPLE_KEY_BIT) {
            if (!sig_rep_triple(T, P.hash[(uint32_t)e.x & 0x3FFFFFFFu].x, e.x, rv, piece)) return;
        } else if (!sig_rep_pair(T, e.x, rv, piece)) {
            return;
        }
    } else {
        return;
    }
    const unsigned long long bases = cnt * (unsigned long long)P.len0 + dev;  // sum of the reads' lengths (dev wraps)
    expand_signature(P, rv, piece, [&](uint32_t x, const EvHdr &h, uint64_t cm) {
        bump_class(P, h, cm, cnt);
        atomicAdd(P.bases + x, bases);
    });
}

// ---------------------------------------------------------------------------------------------
// General kernel: one thread per listed read.  The read is filtered again (its interval list is built in
// its ext entry) and walked against every candidate event of its piece: tie-break (count.cpp:429-432),
// Read_single::build, compatibility, IEEE validity rule.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) lsq_general_kernel(const __grid_constant__ ScanParams P) {
    const Tables &T = P.T;
    const uint32_t n = min(*P.gen_n, P.gen_cap);
#pragma unroll 1
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 it = P.gen_list[i];
        const uint32_t r = it.x, base = it.y;
        const uint32_t nraw = load_read_nraw(P, r);
        const uint32_t cs0 = load_read_cs(P, r);
        const DirectCtx ctx{P, base};
        const FilteredRead fr = filter_read_t(P, r, base, nraw, cs0, ctx);
        if (fr.n == 0u || fr.piece == PIECE_NONE) continue;
        ReadView rv;
        rv.blk = P.ext + ext_slot(base, r) + 1;  // the interval list of a read with more than two kept blocks
        rv.b0 = fr.first;
        rv.b1 = fr.n > 1u ? fr.second : make_int2(0, 0);
        rv.b2 = fr.n > 2u ? rv.blk[2] : make_int2(0, 0);
        rv.n = fr.n;
        rv.ridx = r;
        uint32_t total = 0;
        for (uint32_t b = 0; b < fr.n; ++b) {
            const int2 blk = rv.block(b);
            total += (uint32_t)(blk.y - blk.x);
        }
        const uint32_t e0 = __ldg(T.stab_off + fr.piece), e1 = __ldg(T.stab_off + fr.piece + 1);
#pragma unroll 1
        for (uint32_t q = e0; q < e1; ++q) {
            const uint32_t x = __ldg(&T.stab[q].x);
            const EvHdr h = T.hdr[x];
            uint32_t ml;
            const uint64_t cm = eval_pair(P, h, T.seg + h.seg_off, T.fmask + h.form_off, rv, fr.strand, total, ml);
            if (cm == 0ull) continue;
            bump_class(P, h, cm, 1ull);
            atomicAdd(P.bases + x, (unsigned long long)ml);
        }
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
constexpr int MID_MAX_FORMS = 16; // events with at most 16 forms: one thread per event (lsq_solve_mid_kernel)
constexpr int MID_BLOCK = 128;

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 3) lsq_solve_kernel(SolveParams P) {
    __shared__ double s_theta[WARPS][2][LSQ_MAX_FORMS_PER_EVENT_DEV];
    const uint32_t lane = lane_id();
    const uint32_t w = threadIdx.x >> 5;
    const uint32_t E = P.T.E;
    for (uint32_t x = blockIdx.x * WARPS + w; x < E; x += gridDim.x * WARPS) {
        const EvHdr h = P.T.hdr[x];
        const uint32_t K = h.K;
        if (P.small_enabled && K <= 2) continue;  // solved by lsq_solve_small_kernel
        // events with at most MID_MAX_FORMS forms are solved by lsq_solve_mid_kernel
        if (P.mid_enabled && K <= (uint32_t)MID_MAX_FORMS) continue;
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
        double ll = 0.0, near = 1E300;  // near: distance of the stop decision from its threshold (em_stop_is_fragile)
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
                const double tk = fabs(__dsub_rn(1.0, ratio));
                near = fmin(near, fabs(tk - 1E-6));
                if (!(tk > 1E-6) || iters >= P.max_iters) break;
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
                const double tk = fabs(__dsub_rn(1.0, ratio));
                near = fmin(near, fabs(tk - 1E-6));
                if (!(tk > 1E-6) || iters >= P.max_iters) break;
                ll_old = ll;  // the next iteration's old_ll is this very value (same theta, same order)
            }
        }
        __syncwarp();
        // standard errors: lsq_se_kernel (events with at most SE_MAX_FORMS forms)
        if (K > (uint32_t)SE_MAX_FORMS)
            for (uint32_t k = lane; k < K; k += 32) P.se[h.out_form + k] = nan("");
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
            if (iters != 0u && P.flag_n != nullptr && em_stop_is_fragile(near, ll, n_total) &&
                !em_result_is_order_free(K, n_total, iters, P.max_iters, [&](auto f) {
                    for (uint32_t m = 0; m < P.M; ++m) {
                        const ClassRange cr = class_range(P, h, m);
                        for (uint32_t i = 0; i < cr.n; ++i) {
                            const unsigned long long c = cr.cnt[i];
                            if (c) f(m, cr.mask ? (uint64_t)cr.mask[i] : (uint64_t)(i + 1u), (double)c);
                        }
                    }
                }, [&](uint32_t m, uint32_t k) { return P.G[(size_t)m * P.T.F + h.form_off + k]; })) {
                const uint32_t fi = atomicAdd(P.flag_n, 1u);  // replayed read by read in the reference's order
                P.flag_list[fi] = x;
                P.ev_flag[x] = fi + 1u;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Standard errors (the north star's Fisher-information step; common/fim.h and linalg.h are dead code in the
// reference, so there is no reference output to pin this to: the tests hold it against a numpy restatement).
// Observed information of the free parameters theta_0..theta_{K-2} (theta_{K-1} = 1 - the rest) at the estimate,
// per-read term as common/fim.h:322-367:  I_pq = sum_c n_c (v_cp - v_cK)(v_cq - v_cK) / s_c^2,  v_ck = G_k M(c,k);
// se_k = sqrt((I^-1)_kk), se_{K-1} = sqrt(1' I^-1 1) (linalg.h:28-71 inverts with a GSL LU).
// One warp per event, the augmented matrix [I | identity] in REGISTERS: lane j holds column j (2 (K-1) <= 30 columns),
// so Gauss-Jordan elimination needs no shared memory and no barrier -- the pivot row's factors travel by shuffle.  Runs after the solve kernels, for every event with 3..SE_MAX_FORMS forms.
// ---------------------------------------------------------------------------------------------
constexpr int SE_DIM = SE_MAX_FORMS - 1;
constexpr int SE_WARPS = 8;

__global__ void __launch_bounds__(SE_WARPS * 32) lsq_se_kernel(const __grid_constant__ SolveParams P) {
    const uint32_t lane = lane_id(), E = P.T.E;
    for (uint32_t x = blockIdx.x * SE_WARPS + (threadIdx.x >> 5); x < E; x += gridDim.x * SE_WARPS) {
        const EvHdr h = P.T.hdr[x];
        const uint32_t K = h.K;
        if ((P.small_enabled && K <= 2u) || K > (uint32_t)SE_MAX_FORMS) continue;
        if (K == 1u) {
            if (lane == 0u) P.se[h.out_form] = 0.0;
            continue;
        }
        bool any = false;
        for (uint32_t m = 0; m < P.M; ++m) any = any || P.n_valid[(size_t)m * E + h.out_ev] != 0ull;
        const uint32_t D = K - 1u;
        const double th = lane < K ? P.theta[h.out_form + lane] : 0.0;
        double col[SE_DIM];
#pragma unroll
        for (int i = 0; i < SE_DIM; ++i) col[i] = 0.0;
        for (uint32_t m = 0; any && m < P.M; ++m) {
            const ClassRange cr = class_range(P, h, m);
            const double g = lane < K ? __ldg(P.G + (size_t)m * P.T.F + h.form_off + lane) : 0.0;
            for (uint32_t i = 0; i < cr.n; ++i) {
                const unsigned long long c = __ldg(cr.cnt + i);
                if (!c) continue;
                const uint64_t mk = cr.mask ? __ldg(cr.mask + i) : (uint64_t)(i + 1u);
                const double v = ((mk >> lane) & 1ull) ? g : 0.0;  // lanes >= K: g = 0
                const double sc = warp_sum(th * v);
                if (!(sc > 0.0)) continue;
                const double w = (v - __shfl_sync(0xFFFFFFFFu, v, D)) / sc;
                const double a = (double)c * w;
#pragma unroll
                for (int p = 0; p < SE_DIM; ++p)
                    if ((uint32_t)p < D) col[p] = fma(a, __shfl_sync(0xFFFFFFFFu, w, p), col[p]);  // lane q < D: I[p][q]
            }
        }
        if (lane >= D) {
#pragma unroll
            for (int i = 0; i < SE_DIM; ++i) col[i] = (lane - D == (uint32_t)i) ? 1.0 : 0.0;
        }
        bool singular = !any;
#pragma unroll
        for (int r = 0; r < SE_DIM; ++r) {
            if ((uint32_t)r < D && !singular) {
                // the information matrix is symmetric positive semi-definite: the diagonal entry is the pivot, no row
                // exchange is needed for stability; a pivot that is not positive means a singular matrix
                const double pv = __shfl_sync(0xFFFFFFFFu, col[r], r);
                if (!(pv > 0.0)) {
                    singular = true;
                } else {
                    col[r] /= pv;
#pragma unroll
                    for (int i = 0; i < SE_DIM; ++i)
                        if (i != r && (uint32_t)i < D) col[i] = fma(-__shfl_sync(0xFFFFFFFFu, col[i], r), col[r], col[i]);
                }
            }
        }
        if (singular) {
            if (lane < K) P.se[h.out_form + lane] = nan("");
            continue;
        }
        // lane D + k holds column k of the inverse
        double diag = 0.0, sum = 0.0;
#pragma unroll
        for (int i = 0; i < SE_DIM; ++i)
            if ((uint32_t)i < D) {
                if (lane - D == (uint32_t)i) diag = col[i];
                sum += col[i];
            }
        const double tot = warp_sum((lane >= D && lane < 2u * D) ? sum : 0.0);
        if (lane >= D && lane < 2u * D) P.se[h.out_form + lane - D] = sqrt(diag);
        if (lane == 0u) P.se[h.out_form + D] = sqrt(tot);
    }
}

// ---------------------------------------------------------------------------------------------
// Work order of lsq_solve_mid_kernel: events sorted by (classes, forms), so that the 32 events a warp works on at
// any time run class loops of the same length (lanes of one warp share an instruction stream).  Counting sort:
// lsq_mid_key_kernel histograms the keys, one block scans the 2^12 bins, lsq_mid_scatter_kernel places the events.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t MID_KEY_BINS = 4096;  // key = min(classes, 255) << 4 | (forms - 1)
__device__ __forceinline__ uint32_t mid_key(const SolveParams &P, const EvHdr &h) {
    uint32_t c = 0;
    for (uint32_t m = 0; m < P.M; ++m) {
        if (P.inj_off[m]) c += (uint32_t)(P.inj_off[m][h.out_ev + 1] - P.inj_off[m][h.out_ev]);
        else c += h.slot_cap;  // dense or hashed table: its capacity bounds the loop length
    }
    return (min(c, 255u) << 4) | ((h.K - 1u) & 15u);
}
__global__ void __launch_bounds__(256) lsq_mid_key_kernel(const __grid_constant__ SolveParams P, uint32_t *hist) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.T.E) return;
    const EvHdr h = P.T.hdr[x];
    if (h.K > (uint32_t)MID_MAX_FORMS) return;
    atomicAdd(hist + mid_key(P, h), 1u);
}
__global__ void __launch_bounds__(1024) lsq_mid_scan_kernel(uint32_t *hist, uint32_t *n_mid) {  // exclusive scan of MID_KEY_BINS bins
    __shared__ uint32_t s_w[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t v[4], sum = 0;
    for (int j = 0; j < 4; ++j) {
        v[j] = hist[tid * 4 + j];
        sum += v[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (uint32_t d = 1; d < 32u; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31u) s_w[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += s_w[w];
    uint32_t run = before + incl - sum;
    for (int j = 0; j < 4; ++j) {
        hist[tid * 4 + j] = run;
        run += v[j];
    }
    if (tid == 1023u) *n_mid = run;
}
__global__ void __launch_bounds__(256) lsq_mid_scatter_kernel(const __grid_constant__ SolveParams P, uint32_t *cursor, uint32_t *order) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.T.E) return;
    const EvHdr h = P.T.hdr[x];
    if (h.K > (uint32_t)MID_MAX_FORMS) return;
    order[atomicAdd(cursor + mid_key(P, h), 1u)] = x;
}

// ---------------------------------------------------------------------------------------------
// Solve kernel for events with at most MID_MAX_FORMS forms: one THREAD per event, dynamic event fetch (a lane takes
// the next event as soon as its own has converged, so the spread of iteration counts -- mean 42, tail in the
// thousands at BASELINE config 5 -- idles nobody).  One pass over the event's classes per iteration: the class
// probability s_c = sum_k theta_k G_k [k in c] is formed once and gives both the log-likelihood term n_c log s_c and,
// through t_c = n_c / s_c, every form's numerator z_k += theta_k G_k t_c -- so the pass that evaluates ll(theta_i)
// for the stop rule also yields theta_{i+1} (read.h:592-660).  theta, theta*G and the numerators live in shared
// memory, [form][thread], so the dynamic form index costs no local memory and lanes never conflict.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MID_BLOCK) lsq_solve_mid_kernel(const __grid_constant__ SolveParams P) {
    __shared__ double s_th[MID_MAX_FORMS][MID_BLOCK];
    const uint32_t E = P.T.E, t = threadIdx.x;
    // Every lane runs its own event, and every trip of the loop below is ONE likelihood pass for all lanes of the warp:
    // a lane whose event has converged writes it out and takes the next event at the top of the next trip, so the
    // lanes stay converged on the pass (the expensive part) although their events need different numbers of
    // iterations.  phase: 0 = wants an event, 1 = first pass (ll at the uniform start), 2 = iterating (read.h:647-659).
    uint32_t phase = 0, x = 0, iters = 0;
    EvHdr h = {};
    double n_total = 0.0, sum_supports = 0.0, ll_old = 0.0, near = 1E300;
    bool alive = true;
    auto finish = [&](double ll, bool any) {
        const uint32_t K = h.K;
        if (iters != 0u && P.flag_n != nullptr && em_stop_is_fragile(near, ll, n_total) &&
            !em_result_is_order_free(K, n_total, iters, P.max_iters, [&](auto f) {
                for (uint32_t m = 0; m < P.M; ++m) {
                    const ClassRange cr = class_range(P, h, m);
                    for (uint32_t i = 0; i < cr.n; ++i) {
                        const unsigned long long c = cr.cnt[i];
                        if (c) f(m, cr.mask ? (uint64_t)cr.mask[i] : (uint64_t)(i + 1u), (double)c);
                    }
                }
            }, [&](uint32_t m, uint32_t k) { return P.G[(size_t)m * P.T.F + h.form_off + k]; })) {
            const uint32_t fi = atomicAdd(P.flag_n, 1u);  // replayed read by read in the reference's order
            P.flag_list[fi] = x;
            P.ev_flag[x] = fi + 1u;
        }
        for (uint32_t k = 0; k < K; ++k) {  // RPKM (solve.cpp:809-821) and outputs
            const double th = s_th[k][t];
            double r = 0.0;
            for (uint32_t m = 0; m < P.M; ++m) r = __dadd_rn(r, __dmul_rn((double)P.bases[(size_t)m * E + x], th));
            r = __ddiv_rn(r, __ddiv_rn((double)P.form_len[h.form_off + k], 1.0E3));
            r = __ddiv_rn(r, P.total_read_mbases);
            P.theta[h.out_form + k] = th;
            P.rpkm[h.out_form + k] = r;
            if (K == 1u) P.se[h.out_form] = 0.0;
            else if (!any) P.se[h.out_form + k] = nan("");
        }
        P.loglik[h.out_ev] = ll;
        P.ll_per_read[h.out_ev] = (sum_supports > 1E-5) ? __ddiv_rn(ll, sum_supports) : 0.0;
        P.iters[h.out_ev] = iters;
    };
    while (true) {
        if (phase == 0u && alive) {
            while (true) {
                const uint32_t wi = atomicAdd(P.work_counter, 1u);
                if (wi >= *P.mid_n) {
                    alive = false;
                    break;
                }
                x = P.mid_order[wi];  // events in (classes, forms) order: see lsq_mid_key_kernel
                h = P.T.hdr[x];
                const uint32_t K = h.K;
                if (P.small_enabled && K <= 2u) continue;
                n_total = 0.0;
                sum_supports = 0.0;
                bool any = false;
                for (uint32_t m = 0; m < P.M; ++m) {
                    const unsigned long long nv = P.n_valid[(size_t)m * E + h.out_ev];
                    if (nv) {
                        any = true;
                        n_total = __dadd_rn(n_total, (double)nv);
                    }
                    sum_supports = __dadd_rn(sum_supports, (double)nv);
                }
                for (uint32_t k = 0; k < K; ++k) s_th[k][t] = __ddiv_rn(1.0, (double)K);
                iters = 0;
                near = 1E300;
                if (!any) {  // no valid read in any read set: uniform theta, log-likelihood 0 (solve.cpp:798-800)
                    finish(0.0, false);
                    continue;
                }
                if (K == 1u) s_th[0][t] = 1.0;  // solve.cpp:801-802
                phase = 1u;
                break;
            }
        }
        if (__all_sync(0xFFFFFFFFu, !alive)) return;
        if (!alive) continue;
        // one pass at the theta in s_th: ll(theta) and the numerators z of the next theta.  w_k = theta_k G_k and z_k
        // live in registers: the loops over the forms are unrolled to MID_MAX_FORMS with the class's mask bit as the
        // predicate (adding nothing for a form outside the class, exactly like skipping it) and cut short, uniformly
        // for the warp, at the largest K among its lanes' events
        const uint32_t K = h.K, Kmax = __reduce_max_sync(__activemask(), K);
        double z[MID_MAX_FORMS], w[MID_MAX_FORMS];
#pragma unroll
        for (int k = 0; k < MID_MAX_FORMS; ++k) z[k] = 0.0;
        double ll = 0.0;
        for (uint32_t m = 0; m < P.M; ++m) {
            const ClassRange cr = class_range(P, h, m);
            const double *G = P.G + (size_t)m * P.T.F + h.form_off;
#pragma unroll
            for (int k = 0; k < MID_MAX_FORMS; ++k) {
                if ((uint32_t)k >= Kmax) break;
                w[k] = (uint32_t)k < K ? __dmul_rn(s_th[k][t], __ldg(G + k)) : 0.0;
            }
#pragma unroll 1
            for (uint32_t i = 0; i < cr.n; ++i) {
                const unsigned long long c = __ldg(cr.cnt + i);
                if (!c) continue;
                const uint32_t mk = cr.mask ? (uint32_t)__ldg(cr.mask + i) : i + 1u;  // K <= 16: the low bits
                double sm = 0.0;
#pragma unroll
                for (int k = 0; k < MID_MAX_FORMS; ++k) {
                    if ((uint32_t)k >= Kmax) break;
                    if (mk & (1u << k)) sm = __dadd_rn(sm, w[k]);
                }
                const double n_c = (double)c;
                ll = __dadd_rn(ll, __dmul_rn(n_c, log(sm)));
                if (sm > 0.0) {
                    const double tc = __ddiv_rn(n_c, sm);
#pragma unroll
                    for (int k = 0; k < MID_MAX_FORMS; ++k) {
                        if ((uint32_t)k >= Kmax) break;
                        if ((mk & (1u << k)) && w[k] > 0.0) z[k] = __dadd_rn(z[k], __dmul_rn(w[k], tc));
                    }
                }
            }
        }
        bool stop;
        if (phase == 1u) {
            stop = K == 1u;
            phase = 2u;
        } else {
            ++iters;
            const double tk = fabs(__dsub_rn(1.0, __ddiv_rn(ll_old, ll)));
            near = fmin(near, fabs(tk - 1E-6));
            stop = !(tk > 1E-6) || iters >= P.max_iters;
        }
        if (stop) {
            finish(ll, true);
            phase = 0u;
        } else {
            ll_old = ll;
#pragma unroll
            for (int k = 0; k < MID_MAX_FORMS; ++k) {
                if ((uint32_t)k >= Kmax) break;
                if ((uint32_t)k < K) s_th[k][t] = __ddiv_rn(z[k], n_total);
            }
        }
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
        double ll = 0.0, near = 1E300;
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
                const double tk = fabs(__dsub_rn(1.0, ratio));
                near = fmin(near, fabs(tk - 1E-6));
                if (!(tk > 1E-6) || iters >= P.max_iters) break;
                ll_old = ll;
            }
        }
        if (iters != 0u && P.flag_n != nullptr && em_stop_is_fragile(near, ll, n_total) &&
            !em_result_is_order_free(K, n_total, iters, P.max_iters, [&](auto f) {
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (cnt[m][j] != 0.0) f((uint32_t)m, (uint64_t)(j + 1), cnt[m][j]);
            }, [&](uint32_t m, uint32_t k) { return G[m][k]; })) {
            const uint32_t fi = atomicAdd(P.flag_n, 1u);  // replayed read by read in the reference's order
            P.flag_list[fi] = x;
            P.ev_flag[x] = fi + 1u;
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

// ---------------------------------------------------------------------------------------------
// Order-exact replay of the EM for the events the solve kernels flagged (lsq_device_fn.cuh: order-exact EM).
// lsq_collect_kernel walks the read set once more: a read with a block inside the span of a flagged event takes
// the general path (filter, walk) and leaves one record per flagged event it is valid for; lsq_exact_kernel puts
// an event's records in the reference's index order and replays the EM read by read.  Rare: an event is flagged
// when its stop decision lies within summation-order noise of the threshold.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t COLLECT_BIN_SHIFT = 12;
struct CollectParams {
    const uint32_t *ev_flag;      // [E] index + 1 of the event in the flag list, 0: not flagged
    const int32_t *iv_start;      // merged spans of the flagged events, sorted by (chromosome, start) ...
    const int32_t *iv_end;        // ... inclusive ends
    const uint32_t *iv_chrom_off; // [n_chroms + 1] interval range per annotation chromosome
    const uint32_t *bin_bits;     // coarse test: bit b set = a flagged span overlaps positions [b, b + 1) << COLLECT_BIN_SHIFT
    uint32_t n_bins;              //  on some chromosome
    uint32_t *cand;               // blocks that pass the coarse test
    uint32_t *cand_n;
    uint32_t cand_cap;
    const uint64_t *seg_off;      // [n_flag + 1] record range per flagged event
    uint32_t *seg_cur;            // [n_flag] records written so far
    ExactRec *rec;
    const uint8_t *strand_rank;   // file strand id -> rank of its string
    uint32_t m;
};

// One read of the collect pass: exact pre-check of its blocks against the flagged events' spans, then the general
// walk; a record for every flagged event the read is valid for.  first_hit: the block the candidate list named (the
// read is taken only from its FIRST block that passes the coarse test, so that it is collected once).
__device__ __forceinline__ bool collect_coarse_hit(const CollectParams &C, int32_t x) {
    if (x < 0) return false;
    const uint32_t bin = (uint32_t)x >> COLLECT_BIN_SHIFT;
    return bin < C.n_bins && ((__ldg(C.bin_bits + (bin >> 5)) >> (bin & 31u)) & 1u);
}

__device__ void collect_read(const ScanParams &P, const CollectParams &C, uint64_t r, uint32_t base, uint32_t nraw, uint32_t first_hit) {
    const Tables &T = P.T;
    const uint32_t cs0 = load_read_cs(P, r);
    for (uint32_t j = 0; j < first_hit; ++j)
        if (collect_coarse_hit(C, load_block(P, base + j).x)) return;  // an earlier candidate entry owns this read
    bool hit = false;
    for (uint32_t j = first_hit; j < nraw && !hit; ++j) {
        const int2 b = load_block(P, base + j);
        const uint32_t cs = (cs0 & 0x80000000u) ? __ldg(P.blk_cs + base + j) : cs0;
        const uint32_t c = __ldg(P.chrom_remap + (cs & 0xFFFFu));
        if (c == 0xFFFFFFFFu) continue;
        uint32_t lo = __ldg(C.iv_chrom_off + c), hi = __ldg(C.iv_chrom_off + c + 1);
        while (lo < hi) {  // last interval with start <= b.x
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(C.iv_start + mid) <= b.x) lo = mid + 1;
            else hi = mid;
        }
        hit = lo > __ldg(C.iv_chrom_off + c) && b.x <= __ldg(C.iv_end + lo - 1);
    }
    if (!hit) return;
    const DirectCtx ctx{P, base};
    const FilteredRead fr = filter_read_t(P, r, base, nraw, cs0, ctx);
    if (fr.n == 0u || fr.piece == PIECE_NONE) return;
    ReadView rv;
    rv.blk = P.ext + ext_slot(base, (uint32_t)r) + 1;
    rv.b0 = fr.first;
    rv.b1 = fr.n > 1u ? fr.second : make_int2(0, 0);
    rv.b2 = fr.n > 2u ? rv.blk[2] : make_int2(0, 0);
    rv.n = fr.n;
    rv.ridx = (uint32_t)r;
    uint32_t total = 0;
    for (uint32_t b = 0; b < fr.n; ++b) {
        const int2 blk = rv.block(b);
        total += (uint32_t)(blk.y - blk.x);
    }
    const uint32_t e0 = __ldg(T.stab_off + fr.piece), e1 = __ldg(T.stab_off + fr.piece + 1);
    for (uint32_t q = e0; q < e1; ++q) {
        const uint32_t x = __ldg(&T.stab[q].x);
        const uint32_t fi = __ldg(C.ev_flag + x);
        if (fi == 0u) continue;
        const EvHdr h = T.hdr[x];
        uint32_t ml;
        const uint64_t cm = eval_pair(P, h, T.seg + h.seg_off, T.fmask + h.form_off, rv, fr.strand, total, ml);
        if (cm == 0ull) continue;
        const uint64_t at = C.seg_off[fi - 1u] + atomicAdd(C.seg_cur + (fi - 1u), 1u);
        if (at >= C.seg_off[fi]) continue;  // cannot happen: the segment holds the event's n_valid records
        C.rec[at] = make_exact_rec(C.m, fr.first.x, rv.last_end(), __ldg(C.strand_rank + fr.strand),
                                   P.line_num ? __ldg(P.line_num + r) : (uint32_t)r + 1u, cm);
    }
}

// Pass 1 of the collect: the block starts of the whole read set streamed once, one coarse test each (one bit per
// 2^COLLECT_BIN_SHIFT positions that a flagged event's span overlaps on ANY chromosome -- the table stays in L1);
// the blocks that pass (a fraction of a percent) are listed.  HBM-bound: 4 bytes per block.
__global__ void __launch_bounds__(256) lsq_collect_list_kernel(const __grid_constant__ ScanParams P, const __grid_constant__ CollectParams C,
                                                               uint64_t n_blocks) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4u;
    for (uint64_t b = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4u; b < n_blocks; b += stride) {
        int32_t x[4];
        if (P.blk_start != nullptr && b + 4u <= n_blocks && (reinterpret_cast<uintptr_t>(P.blk_start) & 15u) == 0u) {
            const int4 v = __ldcs(reinterpret_cast<const int4 *>(P.blk_start + b));
            x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                x[j] = b + j < n_blocks ? (P.blk_start != nullptr ? __ldcs(P.blk_start + b + j) : __ldcs(&P.blk[b + j].x)) : -1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (collect_coarse_hit(C, x[j])) {
                const uint32_t at = atomicAdd(C.cand_n, 1u);
                if (at < C.cand_cap) C.cand[at] = (uint32_t)(b + j);
            }
    }
}

// Pass 2: one thread per listed block: find its read in the block-offset index, collect the read.
__global__ void __launch_bounds__(128) lsq_collect_kernel(const __grid_constant__ ScanParams P, const __grid_constant__ CollectParams C) {
    const uint32_t n_cand = min(*C.cand_n, C.cand_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cand; i += gridDim.x * blockDim.x) {
        const uint32_t bi = C.cand[i];
        uint64_t r;
        uint32_t base;
        if (P.blk_cnt != nullptr) {
            const uint32_t n_groups = (uint32_t)((P.n_reads + SCAN_GROUP - 1) / SCAN_GROUP), n_sup = (n_groups + SCAN_SUPER - 1) / SCAN_SUPER;
            uint32_t lo = 0, hi = n_sup;  // last super-group that starts at or before block bi
            while (hi - lo > 1u) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(P.sup_base + mid) <= bi) lo = mid;
                else hi = mid;
            }
            const uint32_t sb = __ldg(P.sup_base + lo);
            uint32_t g = lo * SCAN_SUPER, ge = min(n_groups, (lo + 1u) * SCAN_SUPER);
            while (ge - g > 1u) {  // last group of it that does
                const uint32_t mid = (g + ge) >> 1;
                if (sb + __ldg(P.grp_base + mid) <= bi) g = mid;
                else ge = mid;
            }
            base = sb + __ldg(P.grp_base + g);
            r = (uint64_t)g * SCAN_GROUP;
            uint32_t c = __ldg(P.blk_cnt + r);
            while (base + c <= bi) {  // the read of the group that holds the block
                base += c;
                ++r;
                c = __ldg(P.blk_cnt + r);
            }
        } else {
            uint64_t lo = 0, hi = P.n_reads;  // last read with blk_off <= bi (reads without blocks share offsets: take the one that holds it)
            while (hi - lo > 1u) {
                const uint64_t mid = (lo + hi) >> 1;
                if (__ldg(P.blk_off + mid) <= bi) lo = mid;
                else hi = mid;
            }
            r = lo;
            base = __ldg(P.blk_off + r);
        }
        const uint32_t nraw = P.blk_cnt != nullptr ? (uint32_t)__ldg(P.blk_cnt + r) : __ldg(P.blk_off + r + 1) - base;
        collect_read(P, C, r, base, nraw, bi - base);
    }
}

struct ExactParams {
    uint32_t n_flag;
    const uint32_t *flag_list;
    const uint64_t *seg_off;
    const uint32_t *seg_cur;
    ExactRec *rec;
    uint8_t *cls;             // class index of every sorted record (same offsets as rec)
    uint32_t *pre;            // per event: class counts ahead of every ORD_CHUNK reads (ordered_sum), rows of EXACT_MAX_CLS
    const uint64_t *pre_off;  // [n_flag] first row of the event in pre
    uint32_t presorted;       // the records were sorted and classified by grid-wide launches (large events)
    uint32_t scalar_sums;     // LSQ_EXACT_SCALAR_SUMS: every lane runs ordered_sum by itself (the form the CPU tests check)
    uint32_t debug;           // LSQ_EXACT_TIMING: one line per event from the kernel
    unsigned long long *ctab; // per event EXACT_CTAB u64: class keys (form mask | read set << 56) in order of discovery, then an overflow flag
    uint32_t sort_k, sort_j;  // lsq_exact_sort_kernel: the step of the bitonic network (j = 0: first step of merge k)
    uint64_t total;           // records of all events
};

constexpr int EXACT_BLOCK = 256;

// Where the class bytes of flagged event fi start in ExactParams::cls: 16-byte aligned (ordered_sum loads a chunk of 16
// classes at once) with room for a whole last chunk.
__device__ __forceinline__ uint64_t exact_cls_offset(uint64_t seg_off_fi, uint32_t fi) {
    return ((seg_off_fi + 15ull) & ~15ull) + 32ull * fi;
}

__device__ __forceinline__ void exact_cmpxchg(ExactRec *rec, uint32_t i, uint32_t l) {
    const ExactRec a = rec[i], b = rec[l];
    if (exact_rec_less(b, a)) {
        rec[i] = b;
        rec[l] = a;
    }
}

// One step of the bitonic network (same all-ascending form as in lsq_exact_kernel) over the records of ALL flagged
// events at once, one thread per record: for events too large for one block to sort in reasonable time.  The host
// launches the steps in order; an event shorter than the step's span is left untouched by it.
__global__ void __launch_bounds__(256) lsq_exact_sort_kernel(const __grid_constant__ ExactParams X) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= X.total) return;
    uint32_t lo = 0, hi = X.n_flag;  // the event of record g: last f with seg_off[f] <= g
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(X.seg_off + mid) <= g) lo = mid;
        else hi = mid;
    }
    const uint64_t o = __ldg(X.seg_off + lo);
    const uint32_t n = min((uint32_t)(__ldg(X.seg_off + lo + 1) - o), __ldg(X.seg_cur + lo)), i = (uint32_t)(g - o);
    if (i >= n) return;
    const uint32_t l = X.sort_j ? i ^ X.sort_j : i ^ (X.sort_k - 1u);
    if (l > i && l < n) exact_cmpxchg(X.rec + o, i, l);
}

// Classes of the sorted records of all flagged events (large events): one thread per record finds or appends its
// (form mask, read set) key in the event's table -- append-only, the first empty slot is claimed with a compare-and-swap.
constexpr uint32_t EXACT_CTAB = EXACT_MAX_CLS + 1;
__global__ void __launch_bounds__(256) lsq_exact_classify_kernel(const __grid_constant__ ExactParams X) {
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= X.total) return;
    uint32_t lo = 0, hi = X.n_flag;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(X.seg_off + mid) <= g) lo = mid;
        else hi = mid;
    }
    const uint64_t o = __ldg(X.seg_off + lo);
    const uint32_t n = min((uint32_t)(__ldg(X.seg_off + lo + 1) - o), __ldg(X.seg_cur + lo));
    if ((uint32_t)(g - o) >= n) return;
    const ExactRec r = X.rec[g];
    unsigned long long *tab = X.ctab + (size_t)lo * EXACT_CTAB;
    if (r.mask >> 56) {  // no room for the read set in the key: the event takes the per-read path
        tab[EXACT_MAX_CLS] = 1ull;
        return;
    }
    const unsigned long long key = r.mask | ((unsigned long long)exact_rec_set(r) << 56);
    for (uint32_t q = 0; q < EXACT_MAX_CLS; ++q) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tab + q);
        if (cur == 0ull) cur = atomicCAS(tab + q, 0ull, key);
        if (cur == 0ull || cur == key) {
            X.cls[exact_cls_offset(o, lo) + (g - o)] = (uint8_t)q;
            return;
        }
    }
    tab[EXACT_MAX_CLS] = 1ull;  // more classes than the table holds
}

// ordered_sum (lsq_device_fn.cuh) with the work of a binade spread over the lanes of a warp -- same algorithm, same
// result: the increments of the classes are computed by the lanes in parallel (into s_inc, 64 entries of this warp), a
// chunk's 16 reads are taken with one warp scan instead of 16 dependent steps, the boundary search probes 32 places at
// once.  All lanes call it with the same arguments and get the same value.  cls must be 16-byte aligned.
__device__ double ordered_sum_warp(const uint8_t *cls, uint32_t n, const double *term, uint32_t stride, uint32_t C,
                                   const uint32_t *pre, uint32_t *guess, unsigned long long *s_inc) {
    const uint32_t lane = threadIdx.x & 31u;
    auto mag = [&](uint32_t c) { return fabs(term[c * stride]); };
    double s = 0.0;
    uint32_t i = 0, n_searches = 0;
    const uint32_t n_chunks = (n + ORD_CHUNK - 1u) / ORD_CHUNK;
    while (i < n) {
        const bool zero_mode = s == 0.0;  // skip to the first read whose term is not zero (see ordered_sum)
        int e = 0;
        if (!zero_mode) (void)frexp(s, &e);
        const uint64_t S = zero_mode ? 0ull : (uint64_t)ldexp(s, 53 - e);
        const uint64_t rem = zero_mode ? 1ull : (1ull << 53) - S;
        bool tie = false;
        __syncwarp();
        for (uint32_t c = lane; c < C; c += 32u) {
            const double sc = zero_mode ? 0.0 : ldexp(mag(c), 53 - e);
            uint64_t v;
            if (zero_mode) {
                v = mag(c) != 0.0 ? 1ull : 0ull;
            } else if (!(sc < 9007199254740992.0)) {
                v = rem;
            } else {
                const double q = floor(sc), r = sc - q;
                v = (uint64_t)q + (r > 0.5 ? 1ull : 0ull);
                if (r == 0.5) tie = true;
                if (v > rem) v = rem;
            }
            s_inc[c] = v;
        }
        tie = __any_sync(0xFFFFFFFFu, tie);
        __syncwarp();
        const double top = ldexp(1.0, e);
        if (!zero_mode && (tie || e < -900)) {
            while (i < n && s < top) {
                s = lsq_add(s, mag(cls[i]));
                ++i;
            }
            continue;
        }
        uint64_t acc = 0;
        bool crossed = false;
        // the reads [i, end) of i's chunk: lane j < 16 takes position c0 + j; stops before the read that would cross
        auto walk = [&](uint32_t end) {
            const uint32_t c0 = i & ~(ORD_CHUNK - 1u), pos = c0 + (lane & (ORD_CHUNK - 1u));
            const bool valid = lane < ORD_CHUNK && pos >= i && pos < end;
            const unsigned long long d = valid ? s_inc[cls[pos]] : 0ull;
            unsigned long long incl = d;
#pragma unroll
            for (uint32_t sft = 1; sft < ORD_CHUNK; sft <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, sft);
                if (lane >= sft) incl += t;
            }
            const uint32_t over = __ballot_sync(0xFFFFFFFFu, valid && incl >= rem - acc);
            if (over) {
                const int lc = __ffs((int)over) - 1;
                acc += __shfl_sync(0xFFFFFFFFu, incl - d, lc);
                i = c0 + (uint32_t)lc;
                crossed = true;
            } else {
                acc += __shfl_sync(0xFFFFFFFFu, incl, (int)ORD_CHUNK - 1);
                i = end;
            }
        };
        walk(min(n, (i / ORD_CHUNK + 1u) * ORD_CHUNK));
        if (!crossed && i < n) {
            const uint32_t j0 = i / ORD_CHUNK;
            const uint64_t room = rem - acc;
            auto span = [&](uint32_t j) {
                uint64_t g = 0;
                for (uint32_t c = 0; c < C && g < room; ++c) {
                    const uint32_t cnt = __ldg(pre + (size_t)j * C + c) - __ldg(pre + (size_t)j0 * C + c);
                    if (cnt) g += ord_sat_mul(s_inc[c], cnt, room);
                }
                return g < room ? g : room;
            };
            uint32_t lo = j0, hi = n_chunks;
            bool hi_known = false;
            const uint32_t gs = (guess != nullptr && n_searches < ORD_MAX_GUESS) ? guess[n_searches] : 0u;
            auto clampb = [&](uint32_t b) { return b <= lo ? lo + 1u : (b > hi ? hi : b); };
            bool window = gs > lo && gs <= hi;
            while (lo < hi && !(hi_known && hi - lo == 1u)) {
                const uint32_t step = (hi - lo + 31u) / 32u;
                const uint32_t first = gs > lo + 8u ? gs - 8u : lo + 1u;
                const uint32_t mine = window ? clampb(first + lane) : min(hi, lo + (lane + 1u) * step);
                const bool below = span(mine) < room;
                const uint32_t nb = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, below));
                const uint32_t b_last_below = __shfl_sync(0xFFFFFFFFu, mine, nb ? (int)nb - 1 : 0);
                const uint32_t b_first_above = __shfl_sync(0xFFFFFFFFu, mine, nb < 32u ? (int)nb : 31);
                if (nb) lo = b_last_below;
                if (nb < 32u) {
                    hi = b_first_above;
                    hi_known = true;
                }
                window = false;
            }
            __syncwarp();
            if (guess != nullptr && n_searches < ORD_MAX_GUESS && lane == 0u) guess[n_searches] = lo;
            ++n_searches;
            // span(lo), the lanes sharing the classes
            {
                unsigned long long g = 0;
                for (uint32_t c = lane; c < C; c += 32u) {
                    const uint32_t cnt = __ldg(pre + (size_t)lo * C + c) - __ldg(pre + (size_t)j0 * C + c);
                    g += (unsigned long long)cnt * s_inc[c];  // below room in total: no overflow
                }
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) g += __shfl_xor_sync(0xFFFFFFFFu, g, sft);
                acc += g;
            }
            i = min(n, lo * ORD_CHUNK);
            if (i < n) walk(min(n, (lo + 1u) * ORD_CHUNK));
        }
        if (!zero_mode) s = ldexp((double)(S + acc), e - 53);
        if (crossed) {
            s = lsq_add(s, mag(cls[i]));
            ++i;
        }
    }
    return s;
}

// One block per flagged event (they are few): sort the event's records into the reference's read order, replay the
// EM with the reference's own operation order, overwrite the event's results.  What is sequential in the reference
// stays sequential here -- every running sum adds its terms read by read -- but the K + 1 running sums of one sweep
// (the E-step numerators of the K forms and the log-likelihood, all at the same theta) run side by side, one thread
// each, over per-class tables of the terms (exact_em in lsq_device_fn.cuh is the same computation on one thread; it
// stays the fallback for events with more classes than the tables hold).
__global__ void __launch_bounds__(EXACT_BLOCK) lsq_exact_kernel(const __grid_constant__ SolveParams P, const __grid_constant__ ExactParams X) {
    __shared__ double s_tbl[EXACT_MAX_TERMS + 2 * EXACT_MAX_CLS];  // [class][K + 1]: E-step term per form, then the log term
    __shared__ double s_theta[LSQ_MAX_FORMS_PER_EVENT_DEV], s_acc[LSQ_MAX_FORMS_PER_EVENT_DEV + 1];
    __shared__ uint64_t s_cmask[EXACT_MAX_CLS];
    __shared__ uint32_t s_cset[EXACT_MAX_CLS];
    __shared__ uint32_t s_C, s_tab;
    __shared__ uint32_t s_guess[EXACT_BLOCK / 32][ORD_MAX_GUESS];  // ordered_sum: binade ends of the previous sweep, per running sum
    __shared__ unsigned long long s_inc[EXACT_BLOCK / 32][EXACT_MAX_CLS];  // ordered_sum_warp: increments of the classes
    const uint32_t fi = blockIdx.x, t = threadIdx.x, lane = t & 31u;
    for (uint32_t i = t; i < (EXACT_BLOCK / 32) * ORD_MAX_GUESS; i += EXACT_BLOCK) (&s_guess[0][0])[i] = 0u;
    if (fi >= X.n_flag) return;
    const uint32_t x = X.flag_list[fi];
    const EvHdr h = P.T.hdr[x];
    const uint32_t K = h.K, E = P.T.E;
    ExactRec *rec = X.rec + X.seg_off[fi];
    uint8_t *cls = X.cls + exact_cls_offset(X.seg_off[fi], fi);
    const uint32_t n = min((uint32_t)(X.seg_off[fi + 1] - X.seg_off[fi]), X.seg_cur[fi]);
    double n_total = 0.0, sum_supports = 0.0;
    for (uint32_t m = 0; m < P.M; ++m) {
        const unsigned long long nv = P.n_valid[(size_t)m * E + h.out_ev];
        if (nv) n_total = __dadd_rn(n_total, (double)nv);
        sum_supports = __dadd_rn(sum_supports, (double)nv);
    }
    if ((double)n != n_total) return;  // injected class tables have no reads to replay: keep the class-collapsed result
    // ---- sort: bitonic network in its all-ascending form (first step of a merge pairs i with i ^ (k - 1)), so the
    // missing tail of a non-power-of-two n acts as +infinity and is never touched; names are unique, the order total
    for (uint32_t k = 2; !X.presorted && (k >> 1) < n; k <<= 1) {
        for (uint32_t i = t; i < n; i += EXACT_BLOCK) {
            const uint32_t l = i ^ (k - 1u);
            if (l > i && l < n) exact_cmpxchg(rec, i, l);
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0u; j >>= 1) {
            for (uint32_t i = t; i < n; i += EXACT_BLOCK) {
                const uint32_t l = i ^ j;
                if (l > i && l < n) exact_cmpxchg(rec, i, l);
            }
            __syncthreads();
        }
    }
    // ---- classes of the event: distinct (read set, form mask) pairs, found by one warp, 32 records at a time
    if (X.presorted) {  // lsq_exact_classify_kernel has filled cls and the event's class table
        if (t == 0u) {
            const unsigned long long *tab = X.ctab + (size_t)fi * EXACT_CTAB;
            uint32_t C = 0;
            while (C < EXACT_MAX_CLS && tab[C] != 0ull) {
                s_cmask[C] = tab[C] & 0x00FFFFFFFFFFFFFFull;
                s_cset[C] = (uint32_t)(tab[C] >> 56);
                ++C;
            }
            s_C = C;
            s_tab = (tab[EXACT_MAX_CLS] == 0ull && (C + 1u) * (K + 1u) <= EXACT_MAX_TERMS + EXACT_MAX_CLS) ? 1u : 0u;
        }
    } else if (t < 32u) {
        uint32_t C = 0;
        bool tab = true;
        for (uint32_t base = 0; base < n && tab; base += 32u) {
            const uint32_t i = base + lane;
            const bool have = i < n;
            uint64_t mk = 0;
            uint32_t m = 0, c = 0xFFFFFFFFu, from = 0;
            if (have) {
                mk = rec[i].mask;
                m = exact_rec_set(rec[i]);
            }
            while (true) {
                if (have && c == 0xFFFFFFFFu)
                    for (uint32_t q = from; q < C; ++q)
                        if (s_cmask[q] == mk && s_cset[q] == m) {
                            c = q;
                            break;
                        }
                const uint32_t unfound = __ballot_sync(0xFFFFFFFFu, have && c == 0xFFFFFFFFu);
                if (unfound == 0u) break;
                if (C == EXACT_MAX_CLS || (C + 1u) * (K + 1u) > EXACT_MAX_TERMS + EXACT_MAX_CLS) {
                    tab = false;
                    break;
                }
                if (lane == (uint32_t)__ffs((int)unfound) - 1u) {
                    s_cmask[C] = mk;
                    s_cset[C] = m;
                }
                __syncwarp();
                from = C++;
            }
            if (tab && have) cls[i] = (uint8_t)c;
        }
        if (lane == 0u) {
            s_C = C;
            s_tab = tab ? 1u : 0u;
        }
    }
    __syncthreads();
    // long events: class counts ahead of every chunk of reads, for ordered_sum
    const bool fast_sums = s_tab && n >= ORD_MIN_N;
    uint32_t *pre = X.pre + X.pre_off[fi] * EXACT_MAX_CLS;
    if (fast_sums) {
        const uint32_t C = s_C, n_chunks = (n + ORD_CHUNK - 1u) / ORD_CHUNK;
        for (uint32_t j = t; j < n_chunks; j += EXACT_BLOCK) {
            uint32_t cnt[EXACT_MAX_CLS];
            for (uint32_t c = 0; c < C; ++c) cnt[c] = 0;
            const uint32_t i1 = min(n, (j + 1u) * ORD_CHUNK);
            for (uint32_t i = j * ORD_CHUNK; i < i1; ++i) cnt[cls[i]]++;
            for (uint32_t c = 0; c < C; ++c) pre[(size_t)(j + 1u) * C + c] = cnt[c];
        }
        __syncthreads();
        if (t < C) {  // running totals down the chunks, one class per thread
            uint32_t run = 0;
            pre[t] = 0;
            for (uint32_t j = 1; j <= n_chunks; ++j) {
                run += pre[(size_t)j * C + t];
                pre[(size_t)j * C + t] = run;
            }
        }
        __syncthreads();
    }
    uint32_t iters = 0;
    double ll = 0.0;
    if (!s_tab) {
        if (t == 0u) {
            double theta[LSQ_MAX_FORMS_PER_EVENT_DEV], old_theta[LSQ_MAX_FORMS_PER_EVENT_DEV];
            iters = exact_em_per_read(rec, n, n_total, K, P.G + h.form_off, P.T.F, theta, old_theta, ll, P.max_iters);
            for (uint32_t k = 0; k < K; ++k) s_theta[k] = theta[k];
        }
    } else {
        const uint32_t C = s_C, W = K + 1u;
        if (t < K) s_theta[t] = lsq_div(1.0, (double)K);
        double ll_old = 0.0;
        bool first = true;
        while (true) {
            __syncthreads();
            // tables at the current theta: per class the probability s (every form's term added in form order, zeros
            // included -- read.h:620-636), log s, and theta_k M(c,k) / s (read.h:592-618: a skipped term adds nothing)
            for (uint32_t c = t; c < C; c += EXACT_BLOCK)
                s_tbl[c * W + K] = exact_class_prob(s_cmask[c], K, s_theta, P.G + h.form_off + (size_t)s_cset[c] * P.T.F);
            __syncthreads();
            for (uint32_t q = t; q < C * K; q += EXACT_BLOCK) {
                const uint32_t c = q / K, k = q - c * K;
                const double sc = s_tbl[c * W + K];
                const double g = ((s_cmask[c] >> k) & 1ull) ? P.G[h.form_off + (size_t)s_cset[c] * P.T.F + k] : 0.0;
                const double lp = lsq_mul(s_theta[k], g);
                s_tbl[c * W + k] = (sc > 0.0 && lp > 0.0) ? lsq_div(lp, sc) : 0.0;
            }
            __syncthreads();
            for (uint32_t c = t; c < C; c += EXACT_BLOCK) s_tbl[c * W + K] = log(s_tbl[c * W + K]);
            __syncthreads();
            // the sweep: one warp per running sum (form a's E-step terms for a < K, the log-likelihood for a = K), each in
            // read order; the lanes of the warp compute the same thing except inside ordered_sum's search
            for (uint32_t a = t >> 5; a < W; a += EXACT_BLOCK / 32) {
                double acc = 0.0;
                const double *tb = s_tbl + a;
                // the fast path needs terms of one sign, all finite (the log terms are <= 0 unless a class probability
                // rounds above 1)
                bool pos = false, neg = false, fin = true;
                for (uint32_t c = 0; c < C; ++c) {
                    const double v = tb[c * W];
                    pos = pos || v > 0.0;
                    neg = neg || v < 0.0;
                    fin = fin && (v - v == 0.0);
                }
                if (X.debug && lane == 0u && !(fast_sums && fin && !(pos && neg)) && fast_sums)
                    printf("[exact] event %u sum %u of %u: read-by-read sweep (finite %d, positive %d, negative %d)\n", fi, a, W, (int)fin, (int)pos, (int)neg);
                if (fast_sums && fin && !(pos && neg)) {
                    uint32_t *const gs = a < (uint32_t)(EXACT_BLOCK / 32) ? s_guess[a] : nullptr;
                    acc = X.scalar_sums ? ordered_sum(cls, n, tb, W, C, pre, gs) : ordered_sum_warp(cls, n, tb, W, C, pre, gs, s_inc[t >> 5]);
                    if (neg && acc != 0.0) acc = -acc;
                } else {
                    uint32_t i = 0;
                    if ((reinterpret_cast<uintptr_t>(cls) & 15u) == 0u)
                        for (; i + 16u <= n; i += 16u) {
                            const uint4 v = *reinterpret_cast<const uint4 *>(cls + i);
                            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int j = 0; j < 16; ++j) acc = lsq_add(acc, tb[((w[j >> 2] >> (8 * (j & 3))) & 0xFFu) * W]);
                        }
                    for (; i < n; ++i) acc = lsq_add(acc, tb[(uint32_t)cls[i] * W]);
                }
                if (lane == 0u) s_acc[a] = acc;
            }
            __syncthreads();
            ll = s_acc[K];
            bool stop = false;
            if (first) {
                first = false;
            } else {
                ++iters;
                const double d = 1.0 - lsq_div(ll_old, ll);
                stop = !((d < 0.0 ? -d : d) > 1E-6) || iters >= P.max_iters;
            }
            if (stop) break;
            ll_old = ll;
            __syncthreads();
            if (t < K) s_theta[t] = lsq_div(s_acc[t], n_total);
        }
    }
    __syncthreads();
    if (t < K) {  // RPKM (solve.cpp:809-821) and outputs, as the solve kernels
        const double th = s_theta[t];
        double r = 0.0;
        for (uint32_t m = 0; m < P.M; ++m) r = __dadd_rn(r, __dmul_rn((double)P.bases[(size_t)m * E + x], th));
        r = __ddiv_rn(r, __ddiv_rn((double)P.form_len[h.form_off + t], 1.0E3));
        r = __ddiv_rn(r, P.total_read_mbases);
        P.theta[h.out_form + t] = th;
        P.rpkm[h.out_form + t] = r;
    }
    if (t == 0u && X.debug) printf("[exact] event %u: %u reads, %u forms, %u classes, tabulated %u, %u iterations\n", fi, n, K, s_C, s_tab, iters);
    if (t == 0u) {
        P.loglik[h.out_ev] = ll;
        P.ll_per_read[h.out_ev] = (sum_supports > 1E-5) ? __ddiv_rn(ll, sum_supports) : 0.0;
        P.iters[h.out_ev] = iters;
    }
}

// FP64 FMA throughput of the device (8 independent chains per thread): the denominator of the solve roofline.
__global__ void __launch_bounds__(256) lsq_dfma_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace lsq
