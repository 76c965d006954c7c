// Host-side layout of the device tables (pure C++, shared by lsq_abi.cu and the CPU replay
// harness under tests/emul/).  Turns the caller's event model (lsq_event_tables, output order)
// into the sorted internal arrays the kernels read: event headers, segments, form masks /
// prefix counts, tie-break data, class-slot layout and the stab arrangement.
#pragma once
#include <atomic>
#include <thread>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lesseq_b200.h"
#include "lsq_types.cuh"

namespace lsq {

struct HostTables {
    uint32_t E = 0, F = 0, n_chroms = 0, n_strands = 0;
    uint64_t slot_total = 0;
    uint32_t max_forms = 0;
    std::vector<uint32_t> order;  // internal -> output event index
    std::vector<uint32_t> ev_chrom;  // chromosome id per event, internal order
    std::vector<EvHdr> hdr;       // internal order + sentinel
    std::vector<int2> seg;
    std::vector<uint64_t> fmask, form_len, form_seg_len;
    std::vector<uint32_t> pfx, form_seg_off;
    std::vector<char> tie;
    std::vector<uint32_t> bp_chrom_off, stab_off, cov_off;
    std::vector<StabEntry> stab;
    std::vector<int32_t> bp_pos, cov_start, cov_end;
    std::vector<PieceInfo> pinfo;
    uint32_t bp_shift = 0;
    std::vector<uint32_t> bp_bkt_off;
    std::vector<uint2> bp_bkt, bp_sub;
    std::vector<uint4> chrom_meta;
    std::vector<int32_t> sig_cut;     // cut position of every signature slot (PieceInfo)
    std::vector<uint32_t> sig_piece;  // piece of every signature slot

    // Table view over the HOST arrays (CPU replay harness only).
    Tables host_view() const {
        Tables T{};
        T.E = E;
        T.F = F;
        T.n_pieces = (uint32_t)bp_pos.size();
        T.n_chroms = n_chroms;
        T.n_ann_strands = n_strands;
        T.n_sig = (uint32_t)sig_cut.size();
        T.sig_cut = sig_cut.data();
        T.sig_piece = sig_piece.data();
        T.total_slots = slot_total;
        T.hdr = hdr.data();
        T.seg = seg.data();
        T.fmask = fmask.data();
        T.pfx = pfx.data();
        T.tie_pool = tie.data();
        T.bp_chrom_off = bp_chrom_off.data();
        T.pinfo = pinfo.data();
        T.stab_off = stab_off.data();
        T.stab = stab.data();
        T.bp_shift = bp_shift;
        T.bp_bkt_off = bp_bkt_off.data();
        T.chrom_meta = chrom_meta.data();
        T.bp_bkt = bp_bkt.data();
        T.bp_sub = bp_sub.data();
        T.cov_off = cov_off.data();
        T.cov_start = cov_start.data();
        T.cov_end = cov_end.data();
        return T;
    }
};

// Bucket index over per-chromosome sorted piece starts (CSR `off`): bucket b of a chromosome covers
// positions [b << shift, (b+1) << shift).  Its entry holds the rank (global index) of the first
// piece starting at or after b << shift and, packed into one word (lsq_types.cuh: bucket word),
// the in-bucket offsets of the pieces that start inside it -- so the piece of a position is found
// with ONE 8-byte load.  A bucket with more starts than fit the word points to eight sub-buckets
// of the same format (one more load); only a crowded sub-bucket falls back to a search.
// One extra bucket per chromosome closes the last range.  The shift is chosen so that there are
// about eight buckets per piece (at most 2^25): pieces cluster inside genes.
inline uint32_t bucket_word(const std::vector<int32_t> &arr, uint32_t i, uint32_t hi, int64_t start, uint32_t shift,
                            uint32_t &cnt) {
    const int64_t end = start + ((int64_t)1 << shift);
    uint32_t j = i;
    while (j < hi && (int64_t)arr[j] < end) ++j;
    cnt = j - i;
    if (cnt > bucket_capacity(shift)) return BKT_MANY << 29;
    uint32_t w = cnt << 29;
    for (uint32_t t = 0; t < cnt; ++t) w |= (uint32_t)((int64_t)arr[i + t] - start) << (t * shift);
    return w;
}
inline void build_bucket_index(const std::vector<int32_t> &arr, const std::vector<uint32_t> &off, uint32_t &shift,
                               std::vector<uint32_t> &bkt_off, std::vector<uint2> &bkt, std::vector<uint2> &sub) {
    const size_t nch = off.size() - 1;
    uint64_t span = 0;
    for (size_t c = 0; c < nch; ++c)
        if (off[c + 1] > off[c]) span += (uint64_t)std::max<int32_t>(arr[off[c + 1] - 1], 0) + 1;
    uint64_t per_piece = 8;
    if (const char *ev = getenv("LSQ_BKT_PER_PIECE")) per_piece = std::max<uint64_t>(1, strtoull(ev, nullptr, 10));  // tuning only
    uint64_t want = std::min<uint64_t>(std::max<uint64_t>(per_piece * arr.size(), 1), 1ull << 25);
    shift = 0;
    while ((span >> shift) > want && shift < 30) ++shift;
    const uint32_t ss = bucket_sub_shift(shift), n_sub = 1u << (shift - ss);
    bkt_off.assign(nch + 1, 0);
    bkt.clear();
    sub.clear();
    for (size_t c = 0; c < nch; ++c) {
        bkt_off[c] = (uint32_t)bkt.size();
        uint32_t lo = off[c], hi = off[c + 1];
        uint32_t nb = hi > lo ? (uint32_t)(((uint64_t)std::max<int32_t>(arr[hi - 1], 0) >> shift) + 1) : 0;
        uint32_t i = lo;
        for (uint32_t b = 0; b <= nb; ++b) {  // nb + 1 entries: the last one is `hi`
            const int64_t start = (int64_t)b << shift;
            while (i < hi && (int64_t)arr[i] < start) ++i;
            uint2 e = make_uint2(b == nb ? hi : i, 0u);
            if (b < nb) {
                uint32_t cnt;
                e.y = bucket_word(arr, i, hi, start, shift, cnt);
                if (bucket_many(e.y)) {
                    e.y |= (uint32_t)sub.size();  // < 2^29: there are at most 2^25 buckets of 8 sub-buckets
                    uint32_t i2 = i;
                    for (uint32_t t = 0; t < n_sub; ++t) {
                        const int64_t s2 = start + ((int64_t)t << ss);
                        while (i2 < hi && (int64_t)arr[i2] < s2) ++i2;
                        uint32_t c2;
                        const uint32_t w2 = bucket_word(arr, i2, hi, s2, ss, c2);
                        sub.push_back(make_uint2(i2, w2));
                    }
                }
            }
            bkt.push_back(e);
        }
    }
    bkt_off[nch] = (uint32_t)bkt.size();
    if (sub.empty()) sub.push_back(make_uint2(0u, 0u));
}

inline int tfail(std::string &err, int code, const std::string &msg) {
    err = msg;
    return code;
}

// "read-<N>" < name ?  0: never, 1: always, 2: decided by comparing the digits of N with the
// suffix of the name after "read-" (count.cpp:64-85 with read names of count.cpp:293-295).
inline uint8_t tie_mode(const std::string &name) {
    static const char P[] = "read-";
    size_t n = std::min<size_t>(5, name.size());
    for (size_t i = 0; i < n; ++i) {
        unsigned char a = (unsigned char)P[i], b = (unsigned char)name[i];
        if (a != b) return a < b ? 1 : 0;
    }
    if (name.size() <= 5) return 0;  // the name is a prefix of every read name
    return 2;
}

// ARS length of one form: closed form of accessible_read_starts.h:48-89 (MEDIUM) and :221-274
// (SHORT, min_partial 0): segments before the first one from which a read of length R no
// longer fits contribute l (MEDIUM) or l+1 (SHORT) starts; that segment contributes
// max(l + 1 - overshoot, 0).
inline uint64_t ars_length(uint32_t read_type, uint64_t R, const uint64_t *len, uint32_t n) {
    uint64_t T = 0;
    for (uint32_t i = 0; i < n; ++i) T += len[i];
    uint64_t total = 0, c = 0;
    for (uint32_t i = 0; i < n; ++i) {
        c += len[i];
        if (c + R > T) {
            long long tail = (long long)len[i] + 1 - (long long)(c + R - T);
            return total + (tail > 0 ? (uint64_t)tail : 0);
        }
        total += len[i] + (read_type == LSQ_READ_SHORT ? 1 : 0);
    }
    return total;
}

// G_k = 1 / ARS_k, or 0 when no read of this length fits (read.h:331-340); internal form order.
inline void form_weights(const HostTables &H, uint32_t read_type, uint64_t read_length, std::vector<double> &G) {
    G.assign(H.F, 0.0);
    for (uint32_t f = 0; f < H.F; ++f) {
        uint32_t a = H.form_seg_off[f], b = H.form_seg_off[f + 1];
        uint64_t ars = ars_length(read_type, read_length, H.form_seg_len.data() + a, b - a);
        double nd = (double)ars;
        G[f] = (nd <= 0) ? 0.0 : 1.0 / nd;
    }
}

// strand order table: bit0 = read strand string < event strand string, bit1 = equal
inline void strand_table(const char *const *file_strands, uint32_t nf, const std::vector<std::string> &ann,
                         std::vector<uint8_t> &tbl) {
    tbl.assign(std::max<size_t>((size_t)nf * ann.size(), 1), 0);
    for (uint32_t i = 0; i < nf; ++i)
        for (size_t j = 0; j < ann.size(); ++j) {
            std::string a = file_strands[i];
            tbl[(size_t)i * ann.size() + j] = (uint8_t)((a < ann[j] ? 1 : 0) | (a == ann[j] ? 2 : 0));
        }
}

inline int build_host_tables(const lsq_event_tables *t, HostTables &H, std::string &err) {
    const uint32_t E = t->n_events;
    if (t->n_chroms > 0xFFFF) return tfail(err, LSQ_ERR_RANGE, "more than 65535 chromosomes");
    if (t->n_strands > 0xFF) return tfail(err, LSQ_ERR_RANGE, "more than 255 strand strings");
    // internal order: (chromosome, gs, ge, output index)
    std::vector<uint32_t> &order = H.order;
    order.assign(E, 0);
    for (uint32_t i = 0; i < E; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (t->ev_chrom[a] != t->ev_chrom[b]) return t->ev_chrom[a] < t->ev_chrom[b];
        if (t->ev_gs[a] != t->ev_gs[b]) return t->ev_gs[a] < t->ev_gs[b];
        if (t->ev_ge[a] != t->ev_ge[b]) return t->ev_ge[a] < t->ev_ge[b];
        return a < b;
    });
    std::vector<lsq::EvHdr> &hdr = H.hdr;
    hdr.assign(E + 1, lsq::EvHdr());
    H.ev_chrom.clear();
    H.max_forms = 0;
    std::vector<int2> &seg = H.seg;
    std::vector<uint64_t> &fmask = H.fmask, &form_len = H.form_len, &form_seg_len = H.form_seg_len;
    std::vector<uint32_t> &pfx = H.pfx, &form_seg_off = H.form_seg_off;
    seg.clear(); fmask.clear(); form_len.clear(); form_seg_len.clear(); pfx.clear();
    form_seg_off.assign(1, 0);
    std::string tie_pool;
    uint64_t slot_total = 0;
    std::vector<uint32_t> member;  // scratch of the loop below
    for (uint32_t ie = 0; ie < E; ++ie) {
        const uint32_t e = order[ie];
        lsq::EvHdr &h = hdr[ie];
        h = lsq::EvHdr();
        const uint32_t s0 = t->ev_seg_off[e], S = t->ev_seg_off[e + 1] - s0;
        const uint32_t f0 = t->ev_form_off[e], K = t->ev_form_off[e + 1] - f0;
        if (K == 0 || K > LSQ_MAX_FORMS_PER_EVENT)
            return tfail(err, LSQ_ERR_RANGE, "event with 0 or more than 64 forms");
        if (S > 0xFFFF) return tfail(err, LSQ_ERR_RANGE, "event with more than 65535 segments");
        if (t->ev_chrom[e] >= t->n_chroms || t->ev_strand[e] >= t->n_strands)
            return tfail(err, LSQ_ERR_ARG, "event chromosome/strand id out of range");
        if (t->ev_ge[e] == INT32_MAX) return tfail(err, LSQ_ERR_RANGE, "event end at INT32_MAX");
        h.gs = t->ev_gs[e];
        h.ge = t->ev_ge[e];
        H.ev_chrom.push_back(t->ev_chrom[e]);
        h.seg_off = (uint32_t)seg.size();
        h.form_off = (uint32_t)fmask.size();
        h.S = (uint16_t)S;
        h.K = (uint16_t)K;
        H.max_forms = std::max(H.max_forms, K);
        h.strand = (uint8_t)t->ev_strand[e];
        h.out_ev = e;
        h.out_form = f0;
        for (uint32_t s = 0; s < S; ++s) {
            if (s && t->seg_start[s0 + s] < t->seg_end[s0 + s - 1])
                return tfail(err, LSQ_ERR_ARG, "segments of an event must be disjoint and ascending");
            seg.push_back(make_int2(t->seg_start[s0 + s], t->seg_end[s0 + s]));
        }
        std::string name(t->ev_names + t->ev_name_off[e], t->ev_names + t->ev_name_off[e + 1]);
        uint8_t mode = tie_mode(name);
        h.flags = mode;
        if (mode == 2) {
            h.tie_off = (uint32_t)tie_pool.size();
            h.tie_len = (uint16_t)std::min<size_t>(name.size() - 5, 0xFFFF);
            tie_pool.append(name, 5, h.tie_len);
        }
        const bool wide = S > 64;
        if (wide) {
            h.flags |= lsq::EV_WIDE;
            h.pfx_off = (uint32_t)pfx.size();
        }
        uint64_t runs_bound = 0;
        for (uint32_t k = 0; k < K; ++k) {
            const uint32_t a = t->form_seg_off[f0 + k], b = t->form_seg_off[f0 + k + 1];
            uint64_t mask = 0, len = 0;
            member.assign(S + 1, 0);
            uint32_t prev = 0;
            for (uint32_t q = a; q < b; ++q) {
                uint32_t si = t->form_seg_idx[q];
                if (si >= S || (q > a && si <= prev))
                    return tfail(err, LSQ_ERR_ARG, "form segment indices must be ascending and < S");
                prev = si;
                if (si < 64) mask |= 1ull << si;
                member[si] = 1;
                uint64_t l = (uint64_t)((int64_t)t->seg_end[s0 + si] - (int64_t)t->seg_start[s0 + si]);
                len += l;
                form_seg_len.push_back(l);
            }
            form_seg_off.push_back((uint32_t)form_seg_len.size());
            fmask.push_back(wide ? 0ull : mask);
            form_len.push_back(len);
            uint64_t nk = b - a;
            runs_bound += nk * (nk + 1) / 2;
            if (wide) {
                uint32_t acc = 0;
                for (uint32_t s = 0; s <= S; ++s) {
                    pfx.push_back(acc);
                    if (s < S) acc += member[s];
                }
            }
        }
        h.slot_off = (uint32_t)slot_total;
        if (K <= (uint32_t)lsq::DENSE_MAX_FORMS) {
            h.slot_cap = (1u << K) - 1u;
        } else {
            // distinct classes <= distinct valid segment runs <= sum_k n_k (n_k + 1) / 2
            uint64_t want = 2 * runs_bound + 2;
            uint64_t cap = 16;
            while (cap < want) cap <<= 1;
            if (cap > (1ull << 30)) return tfail(err, LSQ_ERR_RANGE, "class table of one event exceeds 2^30 slots");
            h.slot_cap = (uint32_t)cap;
            h.flags |= lsq::EV_HASH;
        }
        slot_total += h.slot_cap;
        if (slot_total > 0xFFFFFFFFull) return tfail(err, LSQ_ERR_RANGE, "class tables exceed 2^32 slots");
    }
    hdr[E].seg_off = (uint32_t)seg.size();
    hdr[E].form_off = (uint32_t)fmask.size();

    // stab arrangement: per chromosome, every segment boundary and span end of its events cuts
    // the axis into pieces.  A read is keyed by the piece holding the start of its first kept
    // block; the piece lists the events for which that start lies inside a segment (a start
    // outside every segment of an event can never be matched by the walk of read.h:226-262, so
    // such pairs are not candidates at all).
    std::vector<uint32_t> &bp_chrom_off = H.bp_chrom_off, &stab_off = H.stab_off;
    std::vector<StabEntry> &stab = H.stab;
    std::vector<int32_t> &bp_pos = H.bp_pos;
    bp_chrom_off.assign(t->n_chroms + 1, 0);
    stab_off.assign(1, 0);
    stab.clear();
    bp_pos.clear();
    {
        // the chromosomes are independent: each builds its pieces and stab lists into its own buffers (several threads),
        // then the buffers are joined in chromosome order
        std::vector<uint32_t> ch_first(t->n_chroms + 1, E);
        {
            uint32_t ie = 0;
            for (uint32_t ch = 0; ch < t->n_chroms; ++ch) {
                ch_first[ch] = ie;
                while (ie < E && t->ev_chrom[order[ie]] == ch) ++ie;
            }
            ch_first[t->n_chroms] = ie;
            if (ie != E) return tfail(err, LSQ_ERR_ARG, "event chromosome ids must be < n_chroms");
        }
        struct ChromPart {
            std::vector<int32_t> bps;
            std::vector<StabEntry> stab;
            std::vector<uint32_t> end;  // entries of pieces 0..p, per piece
        };
        std::vector<ChromPart> part(t->n_chroms);
        auto build_chrom = [&](uint32_t ch) {
            ChromPart &P = part[ch];
            const uint32_t first = ch_first[ch], last = ch_first[ch + 1];
            std::vector<int32_t> &bps = P.bps;
            for (uint32_t q = t->cov_off[ch]; q < t->cov_off[ch + 1]; ++q) {
                bps.push_back(t->cov_start[q]);
                bps.push_back(t->cov_end[q]);
            }
            for (uint32_t q = first; q < last; ++q) {
                bps.push_back(hdr[q].gs);
                bps.push_back(hdr[q].ge + 1);
                for (uint32_t j = 0; j < hdr[q].S; ++j) {
                    bps.push_back(seg[hdr[q].seg_off + j].x);
                    bps.push_back(seg[hdr[q].seg_off + j].y);
                }
            }
            std::sort(bps.begin(), bps.end());
            bps.erase(std::unique(bps.begin(), bps.end()), bps.end());
            // the entries of a piece in the order (event, segment) they are met: count, offsets, fill
            std::vector<uint32_t> cnt(bps.size() + 1, 0);
            std::vector<std::pair<uint32_t, uint32_t> > range;  // piece range of every non-empty segment, in that order
            for (uint32_t q = first; q < last; ++q) {
                const EvHdr &h = hdr[q];
                const int2 *sg = seg.data() + h.seg_off;
                for (uint32_t j = 0; j < h.S; ++j) {
                    if (!(sg[j].x < sg[j].y)) continue;
                    const uint32_t a = (uint32_t)(std::lower_bound(bps.begin(), bps.end(), sg[j].x) - bps.begin());
                    const uint32_t b = (uint32_t)(std::lower_bound(bps.begin() + a, bps.end(), sg[j].y) - bps.begin());
                    range.push_back(std::make_pair(a, b));
                    for (uint32_t p = a; p < b; ++p) cnt[p + 1]++;
                }
            }
            for (size_t p = 0; p < bps.size(); ++p) cnt[p + 1] += cnt[p];
            P.stab.resize(cnt[bps.size()]);
            std::vector<uint32_t> cur(cnt.begin(), cnt.end() - 1);
            size_t ri = 0;
            for (uint32_t q = first; q < last; ++q) {
                const EvHdr &h = hdr[q];
                const int2 *sg = seg.data() + h.seg_off;
                for (uint32_t j = 0; j < h.S; ++j) {
                    if (!(sg[j].x < sg[j].y)) continue;
                    StabEntry en;
                    en.x = q;
                    en.meta = j & 0xFFFFu;
                    const std::pair<uint32_t, uint32_t> r = range[ri++];
                    for (uint32_t p = r.first; p < r.second; ++p) {
                        StabEntry e2 = en;
                        if (bps[p] == h.gs) e2.meta |= SE_AT_GS;
                        P.stab[cur[p]++] = e2;
                    }
                }
            }
            P.end.assign(cnt.begin() + 1, cnt.end());
        };
        {
            std::atomic<uint32_t> next(0);
            auto work = [&]() {
                for (uint32_t ch = next++; ch < t->n_chroms; ch = next++) build_chrom(ch);
            };
            unsigned T = std::thread::hardware_concurrency();
            if (const char *e = getenv("LSQ_THREADS")) T = (unsigned)std::max(1, atoi(e));
            T = std::max(1u, std::min(T, std::min(64u, std::max(t->n_chroms, 1u))));
            std::vector<std::thread> th;
            for (unsigned i = 1; i < T; ++i) th.emplace_back(work);
            work();
            for (auto &x : th) x.join();
        }
        for (uint32_t ch = 0; ch < t->n_chroms; ++ch) {
            const ChromPart &P = part[ch];
            bp_chrom_off[ch] = (uint32_t)bp_pos.size();
            if (stab.size() + P.stab.size() > 0x3FFFFFFFull)
                return tfail(err, LSQ_ERR_RANGE, "stab lists exceed 2^30 entries (too many nested events)");
            const size_t stab0 = stab.size();
            bp_pos.insert(bp_pos.end(), P.bps.begin(), P.bps.end());
            stab.insert(stab.end(), P.stab.begin(), P.stab.end());
            for (size_t p = 0; p < P.end.size(); ++p) stab_off.push_back((uint32_t)(stab0 + P.end[p]));
        }
        bp_chrom_off[t->n_chroms] = (uint32_t)bp_pos.size();
    }
    H.cov_off.assign(t->cov_off, t->cov_off + t->n_chroms + 1);
    H.cov_start.assign(t->cov_start, t->cov_start + H.cov_off.back());
    H.cov_end.assign(t->cov_end, t->cov_end + H.cov_off.back());
    H.tie.assign(tie_pool.begin(), tie_pool.end());
    build_bucket_index(H.bp_pos, H.bp_chrom_off, H.bp_shift, H.bp_bkt_off, H.bp_bkt, H.bp_sub);
    H.chrom_meta.resize(std::max<uint32_t>(t->n_chroms, 1));
    for (uint32_t ch = 0; ch < t->n_chroms; ++ch)
        H.chrom_meta[ch] = make_uint4(H.bp_chrom_off[ch], H.bp_chrom_off[ch + 1], H.bp_bkt_off[ch],
                                      H.bp_bkt_off[ch + 1] - H.bp_bkt_off[ch] - 1u);
    H.pinfo.assign(H.bp_pos.size() + 1, PieceInfo());  // one sentinel: clamped candidate loads stay in bounds
    H.sig_cut.clear();
    H.sig_piece.clear();
    std::vector<int32_t> cuts;  // scratch of the loop below
    for (uint32_t ch = 0; ch < t->n_chroms; ++ch) {
        const int32_t *cs = H.cov_start.data() + H.cov_off[ch], *ce = H.cov_end.data() + H.cov_off[ch];
        const size_t nc = H.cov_off[ch + 1] - H.cov_off[ch];
        for (uint32_t p = H.bp_chrom_off[ch]; p < H.bp_chrom_off[ch + 1]; ++p) {
            PieceInfo &pi = H.pinfo[p];
            pi.pos = H.bp_pos[p];
            size_t i = std::upper_bound(cs, cs + nc, pi.pos) - cs;
            pi.cov_end = (i > 0 && ce[i - 1] > pi.pos) ? ce[i - 1] : INT32_MIN;
            pi.sigw = 0;
            pi.cutw = 0;
            if (H.stab_off[p + 1] == H.stab_off[p]) continue;
            // cut list: the segment ends of every entry's chain of abutting segments (lsq_types.cuh: PieceInfo)
            cuts.clear();
            bool tie = false;
            for (uint32_t q = H.stab_off[p]; q < H.stab_off[p + 1]; ++q) {
                const StabEntry &en = H.stab[q];
                const EvHdr &h = hdr[en.x];
                const int2 *sg = seg.data() + h.seg_off;
                tie = tie || (en.meta & SE_AT_GS);
                for (uint32_t j = en.meta & 0xFFFFu;; ++j) {
                    cuts.push_back(sg[j].y);
                    if (j + 1 >= h.S || sg[j + 1].x != sg[j].y) break;
                }
            }
            std::sort(cuts.begin(), cuts.end());
            cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
            uint32_t n = (uint32_t)cuts.size();
            const size_t so = H.sig_cut.size();
            bool usable = n <= PIECE_MAX_CUTS && so + n <= (size_t)PIECE_SIG_MASK;
            for (uint32_t k = 0; k < n && k < 2u && usable; ++k)
                usable = (int64_t)cuts[k] - (int64_t)pi.pos < 0xFFFF && cuts[k] > pi.pos;
            pi.sigw = PIECE_CAND | (tie ? PIECE_TIE : 0u);
            if (!usable) {
                pi.sigw |= PIECE_NO_SIG << PIECE_SIG_BITS;
                continue;
            }
            pi.sigw |= (uint32_t)so | (n << PIECE_SIG_BITS);
            for (uint32_t k = 0; k < n; ++k) {
                H.sig_cut.push_back(cuts[k]);
                H.sig_piece.push_back(p);
                if (k < 2u) pi.cutw |= (uint32_t)(cuts[k] - pi.pos) << (16u * k);
            }
        }
    }
    if (H.sig_cut.empty()) {  // never null on the device
        H.sig_cut.push_back(0);
        H.sig_piece.push_back(0);
    }
    H.slot_total = slot_total;
    H.E = E;
    H.F = (uint32_t)fmask.size();
    H.n_chroms = t->n_chroms;
    H.n_strands = t->n_strands;
    return LSQ_OK;

}

}  // namespace lsq
