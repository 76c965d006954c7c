// CPU replay harness -- TEST INFRASTRUCTURE ONLY (never built into or loaded by lesseq_b200).
//
// The development container has no GPU, so kernel logic bugs would otherwise surface only on
// gpurun calls.  This program compiles the product's own per-read / per-pair device functions
// (lesseq_b200/csrc/lsq_device_fn.cuh, via the LSQ_HD switch) and the product's host table
// layout (lsq_tables_host.hpp) with g++ and replays the scan (filter, routing, signature tally) ->
// expand -> general pipeline sequentially; the EM is restated over the class histogram with the same
// per-class arithmetic as lsq_solve_kernel.  tests/test_emul_vs_oracle.py compares its output with the
// oracle.  It takes the same positional arguments as the reference programs.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <chrono>
#include <random>
#include <string>
#include <vector>

#include "../../lesseq_b200/csrc/host/annotation.hpp"
#include "../../lesseq_b200/csrc/host/mrf.hpp"
#include "../../lesseq_b200/csrc/lsq_device_fn.cuh"
#include "../../lesseq_b200/csrc/lsq_tables_host.hpp"

static std::string g_err;
namespace lsqh {
void set_error(const std::string &m) { g_err = m; }
}  // namespace lsqh
extern "C" const char *lsq_last_error(void) { return g_err.c_str(); }

using namespace lsq;

static double class_prob_h(uint64_t mk, const double *theta, const double *G) {
    double s = 0.0;
    for (int k = 0; k < 64; ++k)
        if ((mk >> k) & 1ull) s = s + theta[k] * G[k];
    return s;
}

// Self-test of the piece lookup (two-level bucket index of lsq_tables_host.hpp / find_piece and the
// staged block_lookup of lsq_device_fn.cuh) against a linear scan, on clustered piece starts that
// force large bucket shifts, crowded buckets and crowded sub-buckets.
static int bucket_selftest(unsigned seed) {
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&](uint32_t n) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        return (uint32_t)((st >> 33) % n);
    };
    const uint32_t nch = 1 + rnd(3);
    std::vector<int32_t> pos;
    std::vector<uint32_t> off(nch + 1, 0);
    const uint32_t gap_scale = 1u << rnd(24);
    for (uint32_t c = 0; c < nch; ++c) {
        off[c] = (uint32_t)pos.size();
        int32_t x = (int32_t)rnd(1000) - (rnd(4) == 0 ? 300 : 0);
        const uint32_t clusters = 1 + rnd(60);
        for (uint32_t k = 0; k < clusters; ++k) {
            const uint32_t m = 1 + rnd(40), step = 1 + rnd(rnd(2) ? 3 : 40);
            for (uint32_t i = 0; i < m; ++i) {
                pos.push_back(x);
                x += 1 + (int32_t)rnd(step);
            }
            x += (int32_t)rnd(gap_scale + 1);
        }
    }
    off[nch] = (uint32_t)pos.size();
    uint32_t shift = 0;
    std::vector<uint32_t> bkt_off;
    std::vector<uint2> bkt, sub;
    build_bucket_index(pos, off, shift, bkt_off, bkt, sub);
    std::vector<PieceInfo> pinfo(pos.size() + 1);
    for (size_t i = 0; i < pos.size(); ++i) {
        pinfo[i].pos = pos[i];
        pinfo[i].cov_end = (i % 3) ? pos[i] + 50 : INT32_MIN;
        pinfo[i].sigw = PIECE_CAND;
        pinfo[i].cutw = 0;
    }
    std::vector<uint4> meta(nch);
    std::vector<uint32_t> remap(nch);
    for (uint32_t c = 0; c < nch; ++c) {
        meta[c] = make_uint4(off[c], off[c + 1], bkt_off[c], bkt_off[c + 1] - bkt_off[c] - 1u);
        remap[c] = c;
    }
    ScanParams P{};
    P.T.bp_shift = shift;
    P.T.bp_bkt = bkt.data();
    P.T.bp_sub = sub.data();
    P.T.pinfo = pinfo.data();
    P.T.chrom_meta = meta.data();
    P.chrom_remap = remap.data();
    P.file_meta = meta.data();
    size_t checked = 0;
    for (uint32_t c = 0; c < nch; ++c) {
        const uint32_t lo = off[c], hi = off[c + 1];
        if (lo == hi) continue;
        const int32_t x0 = pos[lo] - 20, x1 = pos[hi - 1] + 2000;
        const int32_t stride = (x1 - x0) > 4000000 ? 1 + (x1 - x0) / 4000000 : 1;
        auto check = [&](int32_t x) {
            uint32_t want = 0xFFFFFFFFu;  // last piece with pos <= x
            {
                uint32_t a = lo, b = hi;
                while (a < b) {
                    const uint32_t mid = (a + b) >> 1;
                    if (pos[mid] <= x) a = mid + 1;
                    else b = mid;
                }
                if (a > lo) want = a - 1;
            }
            PieceInfo pi;
            const uint32_t got = find_piece(P.T, c, x, pi);
            const int2 blk = make_int2(x, x + 30);
            const uint32_t r1 = block_lookup(P, blk, c), r2 = lookup_direct(P.T, c, blk);
            const uint32_t wr = want == 0xFFFFFFFFu ? PIECE_NONE : (want | (pinfo[want].cov_end >= x + 30 ? 0x80000000u : 0u));
            if (got != want || r1 != wr || r2 != wr) {
                fprintf(stderr, "bucket selftest: seed %u chrom %u x %d shift %u: find_piece %u, block_lookup %x, direct %x, want %u / %x\n",
                        seed, c, x, shift, got, r1, r2, want, wr);
                return false;
            }
            ++checked;
            return true;
        };
        for (int32_t x = x0; x <= x1; x += stride)
            if (!check(x)) return 1;
        for (uint32_t i = lo; i < hi; ++i)  // every piece start and its neighbours
            for (int32_t d = -1; d <= 1; ++d)
                if (!check(pos[i] + d)) return 1;
    }
    printf("ok seed %u shift %u pieces %zu buckets %zu sub %zu checked %zu\n", seed, shift, pos.size(), bkt.size(), sub.size(), checked);
    return 0;
}

static lsq_event_tables tables_of(const lsqh_annotation &ann) {
    lsq_event_tables t{};
    t.n_events = (uint32_t)ann.ev_chrom.size();
    t.n_forms = (uint32_t)ann.form_name.size();
    t.n_segments = (uint32_t)ann.seg_start.size();
    t.n_chroms = (uint32_t)ann.chroms.names.size();
    t.n_strands = (uint32_t)ann.strands.names.size();
    t.ev_chrom = ann.ev_chrom.data();
    t.ev_strand = ann.ev_strand.data();
    t.ev_gs = ann.ev_gs.data();
    t.ev_ge = ann.ev_ge.data();
    t.ev_seg_off = ann.ev_seg_off.data();
    t.ev_form_off = ann.ev_form_off.data();
    t.ev_name_off = ann.ev_name_off.data();
    t.ev_names = ann.ev_names_pool.data();
    t.seg_start = ann.seg_start.data();
    t.seg_end = ann.seg_end.data();
    t.form_seg_off = ann.form_seg_off.data();
    t.form_seg_idx = ann.form_seg_idx.data();
    t.cov_off = ann.cov_off.data();
    t.cov_start = ann.cov_start.data();
    t.cov_end = ann.cov_end.data();
    t.chrom_names = ann.chrom_ptrs.data();
    t.strand_names = ann.strand_ptrs.data();
    return t;
}

// ordered_sum (the binade-exact running sum of lsq_exact_kernel) against the plain loop, bit for bit, on random class
// sequences: terms spread over many binades, short ones (so that exact half-way cases occur), zeros, long runs.
static int ordered_sum_selftest(unsigned seed) {
    std::mt19937_64 rng(seed);
    const uint32_t C = 1 + rng() % 64, n = 1 + rng() % (seed % 7 == 0 ? 3000000u : 200000u), stride = 1 + rng() % 3;
    std::vector<double> term((size_t)C * stride, 0.0);
    const int style = (int)(rng() % 4);
    for (uint32_t c = 0; c < C; ++c) {
        double t;
        if (style == 0) t = ldexp((double)(1 + rng() % 7), -(int)(rng() % 12));                      // short: ties galore
        else if (style == 1) t = ldexp((double)(rng() >> 11) / 9007199254740992.0, -(int)(rng() % 40));  // full mantissas
        else if (style == 2) t = (rng() % 5 == 0) ? 0.0 : -log((double)(1 + rng() % 1000) / 1001.0);   // like log terms
        else t = (double)(1 + rng() % 1000) / (double)(1 + rng() % 1000);
        term[(size_t)c * stride] = (seed & 1) ? -t : t;
    }
    if (seed % 11 == 0)  // a form whose theta went to 0: nothing but zero terms
        for (uint32_t c = 0; c < C; ++c) term[(size_t)c * stride] = 0.0;
    if (seed % 11 == 1)  // ... or all but one class
        for (uint32_t c = 1; c < C; ++c) term[(size_t)c * stride] = (seed & 1) ? -0.0 : 0.0;
    std::vector<uint8_t> cls(n);
    uint32_t run = 0, cur = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (run == 0) {
            cur = (uint32_t)(rng() % C);
            run = 1 + (uint32_t)(rng() % (style == 3 ? 5000 : 8));
        }
        cls[i] = (uint8_t)cur;
        --run;
    }
    const uint32_t n_chunks = (n + lsq::ORD_CHUNK - 1) / lsq::ORD_CHUNK;
    std::vector<uint32_t> pre((size_t)(n_chunks + 1) * C, 0);
    for (uint32_t j = 0; j < n_chunks; ++j) {
        for (uint32_t c = 0; c < C; ++c) pre[(size_t)(j + 1) * C + c] = pre[(size_t)j * C + c];
        for (uint32_t i = j * lsq::ORD_CHUNK; i < std::min(n, (j + 1) * lsq::ORD_CHUNK); ++i) pre[(size_t)(j + 1) * C + cls[i]]++;
    }
    double seq = 0.0;
    for (uint32_t i = 0; i < n; ++i) seq = seq + fabs(term[(size_t)cls[i] * stride]);
    std::vector<uint32_t> guess(lsq::ORD_MAX_GUESS, 0);
    const double fast = lsq::ordered_sum(cls.data(), n, term.data(), stride, C, pre.data(), (seed & 2) ? guess.data() : nullptr);
    if (memcmp(&seq, &fast, 8) != 0) {
        fprintf(stderr, "ordered_sum selftest: seed %u C %u n %u style %d: %.17g != %.17g\n", seed, C, n, style, fast, seq);
        return 1;
    }
    // the next sweep of an EM: slightly different terms, the remembered binade ends of this one (right, nearly right or stale)
    for (int sweep = 0; sweep < 3; ++sweep) {
        for (uint32_t c = 0; c < C; ++c) term[(size_t)c * stride] *= 1.0 + (double)((int)(rng() % 2001) - 1000) * (sweep == 2 ? 1e-3 : 1e-9);
        double seq2 = 0.0;
        for (uint32_t i = 0; i < n; ++i) seq2 = seq2 + fabs(term[(size_t)cls[i] * stride]);
        const double fast2 = lsq::ordered_sum(cls.data(), n, term.data(), stride, C, pre.data(), guess.data());
        if (memcmp(&seq2, &fast2, 8) != 0) {
            fprintf(stderr, "ordered_sum selftest (remembered ends): seed %u sweep %d: %.17g != %.17g\n", seed, sweep, fast2, seq2);
            return 1;
        }
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    if (std::string(argv[1]) == "bucket-selftest") {
        const unsigned n = argc > 2 ? (unsigned)atoi(argv[2]) : 50;
        for (unsigned seed = 1; seed <= n; ++seed)
            if (bucket_selftest(seed)) return 1;
        return 0;
    }
    if (std::string(argv[1]) == "ordered-sum-selftest") {
        const unsigned n = argc > 2 ? (unsigned)atoi(argv[2]) : 200;
        for (unsigned seed = 1; seed <= n; ++seed)
            if (ordered_sum_selftest(seed)) return 1;
        return 0;
    }
    const bool solve = std::string(argv[1]) == "solve";
    argv += 1;
    argc -= 1;
    const int group = solve ? 5 : 4;
    if (argc < 10 + group) return 1;
    int argi = 4;
    std::string form_path = argv[argi + 1], map_path = argv[argi + 3];
    unsigned long begin = strtoul(argv[8], nullptr, 10), end = strtoul(argv[9], nullptr, 10);
    argi = 10;
    struct RA {
        std::string type, path;
        unsigned long len;
        double trb;
    };
    std::vector<RA> sets;
    while (argi < argc) {
        RA r;
        argi++;
        r.type = argv[argi++];
        r.len = strtoul(argv[argi++], nullptr, 10);
        r.path = argv[argi++];
        r.trb = solve ? atof(argv[argi++]) : 0;
        sets.push_back(r);
    }
    // LSQ_EMUL_SHARDS=P: the multi-GPU data flow, sequentially.  The reads are routed with the tables of the whole
    // selection (owner mask of every block's piece, as lsq_route_kernel), then every gene-rank shard counts its own
    // bucket with its own tables; the concatenated tables must equal the unsharded run.
    uint32_t n_shards = 1;
    if (const char *e = getenv("LSQ_EMUL_SHARDS")) n_shards = std::max(1, atoi(e));
    std::vector<lsqh_reads> parsed(sets.size());
    for (size_t m = 0; m < sets.size(); ++m)
        if (lsqh::parse_mrf(sets[m].path.c_str(), parsed[m])) {
            fprintf(stderr, "emul: %s\n", g_err.c_str());
            return 1;
        }
    std::vector<std::vector<uint64_t> > owner_mask(sets.size());  // per read: ranks it is sent to
    uint64_t n_sel = 0;
    if (n_shards > 1) {
        lsqh_annotation full;
        if (lsqh::load_annotation_fmt(argv[4], form_path.c_str(), argv[6], map_path.c_str(), begin, end, 0, UINT64_MAX, full)) {
            fprintf(stderr, "emul: %s\n", g_err.c_str());
            return 1;
        }
        n_sel = full.ev_chrom.size();
        lsq_event_tables ft = tables_of(full);
        HostTables FH;
        std::string ferr;
        if (build_host_tables(&ft, FH, ferr)) return 1;
        const Tables FT = FH.host_view();
        std::vector<uint64_t> owner(FH.bp_pos.size() + 1, 0);
        for (size_t p = 0; p < FH.bp_pos.size(); ++p)
            for (uint32_t q = FH.stab_off[p]; q < FH.stab_off[p + 1]; ++q) {
                const uint64_t e = FH.hdr[FH.stab[q].x].out_ev;
                uint32_t d = 0;
                while (d + 1 < n_shards && e >= (n_sel * (d + 1)) / n_shards) ++d;
                owner[p] |= 1ull << d;
            }
        std::map<std::string, uint32_t> cid;
        for (uint32_t i = 0; i < ft.n_chroms; ++i) cid[ft.chrom_names[i]] = i;
        for (size_t m = 0; m < sets.size(); ++m) {
            const lsqh_reads &R = parsed[m];
            const size_t n = R.read_cs.size();
            owner_mask[m].assign(n, 0);
            std::vector<uint32_t> remap(std::max<size_t>(R.chroms.names.size(), 1), 0xFFFFFFFFu);
            for (size_t i = 0; i < R.chroms.names.size(); ++i) {
                auto it = cid.find(R.chroms.names[i]);
                if (it != cid.end()) remap[i] = it->second;
            }
            std::vector<uint4> fmeta(remap.size(), make_uint4(0u, 0u, 0u, 0u));
            for (size_t i = 0; i < remap.size(); ++i)
                if (remap[i] != 0xFFFFFFFFu) fmeta[i] = FH.chrom_meta[remap[i]];
            ScanParams RP{};
            RP.T = FT;
            RP.chrom_remap = remap.data();
            RP.file_meta = fmeta.data();
            for (size_t r = 0; r < n; ++r)
                for (uint32_t j = R.blk_off[r]; j < R.blk_off[r + 1]; ++j) {
                    const uint32_t cs = (R.read_cs[r] & 0x80000000u) ? R.blk_cs[j] : R.read_cs[r];
                    const int2 b = make_int2(R.blk[2 * j], R.blk[2 * j + 1]);
                    const uint32_t p = block_lookup(RP, b, cs) & PIECE_NONE;
                    if (p != PIECE_NONE && (FT.pinfo[p].sigw & PIECE_CAND)) owner_mask[m][r] |= owner[p];
                }
        }
    }
    for (uint32_t shard = 0; shard < n_shards; ++shard) {
    const uint64_t sb = n_shards > 1 ? (n_sel * shard) / n_shards : 0, se = n_shards > 1 ? (n_sel * (shard + 1)) / n_shards : UINT64_MAX;
    lsqh_annotation ann;
    int rc = lsqh::load_annotation_fmt(argv[4], form_path.c_str(), argv[6], map_path.c_str(), begin, end, sb, se, ann);
    if (rc) {
        fprintf(stderr, "emul: %s\n", g_err.c_str());
        return 1;
    }
    lsq_event_tables t = tables_of(ann);
    HostTables H;
    std::string err;
    const auto tb0 = std::chrono::steady_clock::now();
    rc = build_host_tables(&t, H, err);
    if (getenv("LSQ_CLI_TIMING"))
        fprintf(stderr, "[timing] emul_host_tables_ms %.1f\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tb0).count());
    if (rc) {
        fprintf(stderr, "emul: %s\n", err.c_str());
        return 1;
    }
    const Tables T = H.host_view();
    const uint32_t E = H.E, M = (uint32_t)sets.size();
    if (E == 0) continue;
    std::map<std::string, uint32_t> chrom_id;
    for (uint32_t i = 0; i < t.n_chroms; ++i) chrom_id[t.chrom_names[i]] = i;
    std::vector<std::string> ann_strands(ann.strands.names);

    // class histogram per (m, internal event): mask -> count
    std::vector<std::vector<std::map<uint64_t, uint64_t> > > hist(M, std::vector<std::map<uint64_t, uint64_t> >(E));
    std::vector<std::vector<uint64_t> > bases(M, std::vector<uint64_t>(E, 0));
    std::vector<std::vector<double> > Gm(M);
    struct SetCtx {  // what the exact replay needs to walk the reads of a set again
        std::vector<uint32_t> remap;
        std::vector<uint4> fmeta;
        std::vector<uint8_t> stbl;
        std::vector<int2> ext2;
        std::vector<uint32_t> strand_rank;
        ScanParams FP;
    };
    std::vector<SetCtx> sctx(M);
    for (uint32_t m = 0; m < M; ++m) {
        uint32_t type;
        if (sets[m].type == "MEDIUM_READ") type = LSQ_READ_MEDIUM;
        else if (sets[m].type == "SHORT_READ") type = LSQ_READ_SHORT;
        else return 1;
        form_weights(H, type, sets[m].len, Gm[m]);
        const lsqh_reads &R = parsed[m];
        const size_t n = R.read_cs.size(), nb = R.blk_cs.size();
        std::vector<uint32_t> &remap = sctx[m].remap;
        remap.assign(std::max<size_t>(R.chroms.names.size(), 1), 0xFFFFFFFFu);
        for (size_t i = 0; i < R.chroms.names.size(); ++i) {
            auto it = chrom_id.find(R.chroms.names[i]);
            if (it != chrom_id.end()) remap[i] = it->second;
        }
        std::vector<uint8_t> &stbl = sctx[m].stbl;
        strand_table(R.strand_ptrs.data(), (uint32_t)R.strand_ptrs.size(), ann_strands, stbl);
        std::vector<int2> &ext2 = sctx[m].ext2;
        ext2.assign(nb + 2 * n + 2, make_int2(0, 0));
        sctx[m].strand_rank.assign(R.strands.names.size(), 0);
        for (size_t i = 0; i < R.strands.names.size(); ++i)
            for (size_t j = 0; j < R.strands.names.size(); ++j)
                if (R.strands.names[j] < R.strands.names[i]) sctx[m].strand_rank[i]++;
        ScanParams &FP = sctx[m].FP;
        FP = ScanParams{};
        FP.T = T;
        FP.n_reads = n;
        FP.blk_off = R.blk_off.data();
        FP.blk = reinterpret_cast<const int2 *>(R.blk.data());
        FP.read_cs = R.read_cs.data();
        FP.blk_cs = R.any_mixed ? R.blk_cs.data() : nullptr;
        FP.chrom_remap = remap.data();
        std::vector<uint4> &fmeta = sctx[m].fmeta;
        fmeta.assign(remap.size(), make_uint4(0u, 0u, 0u, 0u));
        for (size_t i = 0; i < remap.size(); ++i)
            if (remap[i] != 0xFFFFFFFFu) fmeta[i] = H.chrom_meta[remap[i]];
        FP.file_meta = fmeta.data();
        FP.ext = ext2.data();
        FP.line_num = R.line_num.data();
        FP.strand_tbl = stbl.data();
        FP.len0 = (uint32_t)sets[m].len;
        if (const char *f = getenv("LSQ_SCAN_FLAGS")) FP.flags = (uint32_t)strtoul(f, nullptr, 0);
        // lsq_scan_kernel, one "thread" at a time: filter, route, tally the signature
        std::vector<uint64_t> sig_cnt(T.n_sig, 0);
        std::vector<int64_t> sig_dev(T.n_sig, 0);
        std::map<uint64_t, std::pair<uint64_t, int64_t> > pair_sig;
        std::map<std::pair<uint64_t, uint32_t>, std::pair<uint64_t, int64_t> > triple_sig;  // (prefix pair key, piece 2 << 4 | cut)
        std::vector<uint2> gen_list;
        uint64_t n_single = 0, n_pair = 0, n_triple = 0;
        unsigned long long gen_by_blocks[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t r = 0; r < n; ++r) {
            if (n_shards > 1 && !((owner_mask[m][r] >> shard) & 1ull)) continue;  // not in this rank's bucket
            const uint32_t base = R.blk_off[r], nraw = R.blk_off[r + 1] - base;
            uint32_t sig, total;
            uint64_t key;
            const uint32_t rt = scan_route(FP, r, base, nraw, sig, key, total);  // inline single / two-block paths, as the kernel
            if (rt == RT_SINGLE) {
                if (sig >= T.n_sig) {
                    fprintf(stderr, "emul: signature slot out of range\n");
                    return 3;
                }
                sig_cnt[sig]++;
                sig_dev[sig] += (int64_t)total - (int64_t)FP.len0;
                ++n_single;
            } else if (rt == RT_PAIR) {
                auto &e = pair_sig[key];
                e.first++;
                e.second += (int64_t)total - (int64_t)FP.len0;
                ++n_pair;
            } else if (rt == RT_TRIPLE) {
                auto &e = triple_sig[std::make_pair(key, sig)];
                e.first++;
                e.second += (int64_t)total - (int64_t)FP.len0;
                ++n_triple;
            } else if (rt == RT_GENERAL) {
                gen_list.push_back(make_uint2((uint32_t)r, base));
                gen_by_blocks[nraw < 7u ? nraw : 7u]++;
            }
        }
        if (getenv("LSQ_EMUL_STATS"))
            fprintf(stderr, "emul: general reads by raw blocks 1:%llu 2:%llu 3:%llu 4:%llu 5:%llu 6:%llu 7+:%llu\n", gen_by_blocks[1],
                    gen_by_blocks[2], gen_by_blocks[3], gen_by_blocks[4], gen_by_blocks[5], gen_by_blocks[6], gen_by_blocks[7]);
        if (getenv("LSQ_EMUL_STATS"))
            fprintf(stderr, "emul: reads %zu single-sig %llu pair-sig %llu (distinct %zu) triple-sig %llu (distinct %zu) general %zu\n", n,
                    (unsigned long long)n_single, (unsigned long long)n_pair, pair_sig.size(), (unsigned long long)n_triple,
                    triple_sig.size(), gen_list.size());
        // lsq_expand_kernel
        auto expand = [&](const ReadView &rv, uint32_t piece, uint64_t cnt, int64_t dev) {
            const uint64_t b = (uint64_t)((int64_t)(cnt * (uint64_t)FP.len0) + dev);
            expand_signature(FP, rv, piece, [&](uint32_t x, const EvHdr &, uint64_t cm) {
                hist[m][x][cm] += cnt;
                bases[m][x] += b;
            });
        };
        for (uint32_t i = 0; i < T.n_sig; ++i) {
            if (!sig_cnt[i]) continue;
            ReadView rv;
            uint32_t piece;
            if (!sig_rep_single(T, i, rv, piece)) {
                fprintf(stderr, "emul: counted signature without a representative read\n");
                return 3;
            }
            expand(rv, piece, sig_cnt[i], sig_dev[i]);
        }
        for (auto &kv : pair_sig) {
            ReadView rv;
            uint32_t piece;
            if (!sig_rep_pair(T, kv.first, rv, piece)) {
                fprintf(stderr, "emul: counted pair signature without a representative read\n");
                return 3;
            }
            expand(rv, piece, kv.second.first, kv.second.second);
        }
        for (auto &kv : triple_sig) {
            ReadView rv;
            uint32_t piece;
            if (!sig_rep_triple(T, kv.first.first, triple_key(0u, kv.first.second >> 4, kv.first.second & 15u), rv, piece)) {
                fprintf(stderr, "emul: counted triple signature without a representative read\n");
                return 3;
            }
            expand(rv, piece, kv.second.first, kv.second.second);
        }
        // lsq_general_kernel, one read at a time
        for (const uint2 &it : gen_list) {
            const uint32_t r = it.x, base = it.y;
            const uint32_t nraw = load_read_nraw(FP, r), cs0 = load_read_cs(FP, r);
            const DirectCtx ctx{FP, base};
            const FilteredRead fr = filter_read_t(FP, r, base, nraw, cs0, ctx);
            if (fr.n == 0u || fr.piece == PIECE_NONE) continue;
            ReadView rv;
            rv.blk = FP.ext + ext_slot(base, r) + 1;
            rv.b0 = fr.first;
            rv.b1 = fr.n > 1u ? fr.second : make_int2(0, 0);
            rv.b2 = fr.n > 2u ? rv.blk[2] : make_int2(0, 0);
            rv.n = fr.n;
            rv.ridx = r;
            uint32_t total = 0;
            for (uint32_t b = 0; b < fr.n; ++b) total += (uint32_t)(rv.block(b).y - rv.block(b).x);
            for (uint32_t q = T.stab_off[fr.piece]; q < T.stab_off[fr.piece + 1]; ++q) {
                const uint32_t x = T.stab[q].x;
                const EvHdr &h = T.hdr[x];
                uint32_t ml = 0;
                const uint64_t cm = eval_pair(FP, h, T.seg + h.seg_off, T.fmask + h.form_off, rv, fr.strand, total, ml);
                if (cm) {
                    hist[m][x][cm]++;
                    bases[m][x] += ml;
                }
            }
        }
    }
    // outputs in the caller's order
    std::vector<uint32_t> inv(E);
    for (uint32_t ie = 0; ie < E; ++ie) inv[H.order[ie]] = ie;
    int prec = 6;
    if (const char *p = getenv("LSQ_PRECISION")) prec = atoi(p);
    for (uint32_t e = 0; e < E; ++e) {
        const uint32_t x = inv[e];
        const EvHdr &h = H.hdr[x];
        const uint32_t K = h.K;
        std::vector<uint64_t> nv(M, 0), iso(K, 0);
        for (uint32_t m = 0; m < M; ++m)
            for (auto &kv : hist[m][x]) {
                nv[m] += kv.second;
                for (uint32_t k = 0; k < K; ++k)
                    if ((kv.first >> k) & 1ull) iso[k] += kv.second;
            }
        std::vector<double> theta(K, 1.0 / (double)K), rpkm(K, 0.0);
        double ll = 0.0, sum_supports = 0.0, n_total = 0.0;
        bool any = false;
        for (uint32_t m = 0; m < M; ++m) {
            if (nv[m]) {
                any = true;
                n_total += (double)nv[m];
            }
            sum_supports += (double)nv[m];
        }
        if (solve) {
            auto loglik = [&](const std::vector<double> &th) {
                double s = 0.0;
                for (uint32_t m = 0; m < M; ++m)
                    for (auto &kv : hist[m][x])
                        s += (double)kv.second * log(class_prob_h(kv.first, th.data(), Gm[m].data() + h.form_off));
                return s;
            };
            if (!any) {
            } else if (K == 1) {
                theta[0] = 1.0;
                ll = loglik(theta);
            } else {
                double ll_old = loglik(theta);
                uint32_t iters = 0;
                double near = 1e300;
                while (true) {
                    std::vector<double> old_t = theta;
                    for (uint32_t k = 0; k < K; ++k) {
                        double z = 0.0;
                        for (uint32_t m = 0; m < M; ++m) {
                            const double *G = Gm[m].data() + h.form_off;
                            for (auto &kv : hist[m][x]) {
                                if (!((kv.first >> k) & 1ull)) continue;
                                double s = class_prob_h(kv.first, old_t.data(), G);
                                if (s > 0.0) {
                                    double lp = old_t[k] * G[k];
                                    if (lp > 0.0) z += (double)kv.second * (lp / s);
                                }
                            }
                        }
                        theta[k] = z / n_total;
                    }
                    ll = loglik(theta);
                    ++iters;
                    const double tk = fabs(1.0 - ll_old / ll);
                    if (fabs(tk - 1E-6) < near) near = fabs(tk - 1E-6);
                    if (!(tk > 1E-6)) break;
                    ll_old = ll;
                }
                const bool fragile = em_stop_is_fragile(near, ll, n_total) &&
                                     (getenv("LSQ_EMUL_NO_CERT") != nullptr ||  // test hook: replay what the certificate would clear
                                      !em_result_is_order_free(K, n_total, iters, 0xFFFFFFFFu,
                                                              [&](auto f) {
                                                                  for (uint32_t m = 0; m < M; ++m)
                                                                      for (auto &kv : hist[m][x]) f(m, (uint64_t)kv.first, (double)kv.second);
                                                              },
                                                              [&](uint32_t m, uint32_t k) { return Gm[m][h.form_off + k]; }));
                if (fragile && getenv("LSQ_EMUL_STATS"))
                    fprintf(stderr, "emul: fragile event K %u reads %.0f near %.3g ll %.17g iters %u\n", K, n_total, near, ll, iters);
                if (fragile && !getenv("LSQ_EMUL_NO_EXACT")) {
                    // lsq_collect_kernel + lsq_exact_kernel: the event's valid reads in the reference's index
                    // order, then the EM read by read
                    std::vector<ExactRec> recs;
                    for (uint32_t m = 0; m < M; ++m) {
                        const lsqh_reads &R = parsed[m];
                        const ScanParams &FP = sctx[m].FP;
                        const size_t n = R.read_cs.size();
                        for (size_t r = 0; r < n; ++r) {
                            if (n_shards > 1 && !((owner_mask[m][r] >> shard) & 1ull)) continue;
                            const uint32_t base = R.blk_off[r], nraw = R.blk_off[r + 1] - base;
                            if (nraw == 0) continue;
                            const DirectCtx dctx{FP, base};
                            const FilteredRead fr = filter_read_t(FP, r, base, nraw, load_read_cs(FP, r), dctx);
                            if (fr.n == 0u || fr.piece == PIECE_NONE) continue;
                            ReadView rv;
                            rv.blk = FP.ext + ext_slot(base, (uint32_t)r) + 1;
                            rv.b0 = fr.first;
                            rv.b1 = fr.n > 1u ? fr.second : make_int2(0, 0);
                            rv.b2 = fr.n > 2u ? rv.blk[2] : make_int2(0, 0);
                            rv.n = fr.n;
                            rv.ridx = (uint32_t)r;
                            uint32_t total = 0;
                            for (uint32_t b = 0; b < fr.n; ++b) total += (uint32_t)(rv.block(b).y - rv.block(b).x);
                            for (uint32_t q = T.stab_off[fr.piece]; q < T.stab_off[fr.piece + 1]; ++q) {
                                if (T.stab[q].x != x) continue;
                                uint32_t ml = 0;
                                const uint64_t cm = eval_pair(FP, h, T.seg + h.seg_off, T.fmask + h.form_off, rv, fr.strand, total, ml);
                                if (!cm) continue;
                                recs.push_back(make_exact_rec(m, fr.first.x, rv.last_end(), sctx[m].strand_rank[fr.strand], R.line_num[r], cm));
                            }
                        }
                    }
                    exact_sort(recs.data(), (uint32_t)recs.size());
                    std::vector<double> Gflat((size_t)M * K), old_t(K);
                    for (uint32_t m = 0; m < M; ++m)
                        for (uint32_t k = 0; k < K; ++k) Gflat[(size_t)m * K + k] = Gm[m][h.form_off + k];
                    exact_em(recs.data(), (uint32_t)recs.size(), n_total, K, Gflat.data(), K, theta.data(), old_t.data(), ll, 10000000u);
                }
            }
            double mb = 0.0;
            for (uint32_t m = 0; m < M; ++m) mb += sets[m].trb / 1.0E6;
            for (uint32_t k = 0; k < K; ++k) {
                double r = 0.0;
                for (uint32_t m = 0; m < M; ++m) r += (double)bases[m][x] * theta[k];
                r /= ((double)H.form_len[h.form_off + k] / 1.0E3);
                r /= mb;
                rpkm[k] = r;
            }
        }
        for (uint32_t k = 0; k < K; ++k) {
            printf("%s\t", ann.ev_name[e].c_str());
            for (uint32_t m = 0; m < M; ++m) printf("%.*g\t", prec, (double)nv[m]);
            printf("%s\t", ann.form_name[h.out_form + k].c_str());
            if (!solve) printf("%.*g\n", prec, (double)iso[k]);
            else
                printf("%.*g\t%.*g\t%.*g\n", prec, theta[k], prec, rpkm[k], prec,
                       sum_supports > 1E-5 ? ll / sum_supports : 0.0);
        }
    }
    }  // shards
    return 0;
}
