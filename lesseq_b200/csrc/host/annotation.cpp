// Host front end: annotation files -> event model.  LH_GENE_TXT + UCSC_GENE2ISOFORM are the formats of `count`
// and of the north-star path; `solve` also reads UCSC_GENE_TXT, UCSC_GFF, WORMBASE_GFF2, GENELETS_GFF3 forms and
// WORMBASE_GENE2ISOFORMS maps (solve/solve.cpp:158-296, 310-329), which only differ in how the form rows are read.
//
// What the reference does at count/count.cpp:136-258 (and identically solve/solve.cpp:152-412
// for these two formats), restated over flat vectors:
//   * forms file: one form per '\n'-terminated line (an unterminated last line is ignored, as
//     the reference's `getline && !eof` loops do); whitespace separated
//     name chrom strand txStart txEnd exonCount exonStarts exonEnds.
//   * map file: `event_id form_name` per line; events are the distinct event ids in bytewise
//     lexicographic order and [begin,end) selects by that rank (count.cpp:204-215).
//   * per event: atomic segments (splicing_graph.h:88-169), per-form segment lists
//     (:318-361), span = [min exon start, max exon end] (count.cpp:246,374-375), chromosome and
//     strand of the first form (:239-242); per chromosome the merged exon union of ALL
//     selected events (count.cpp:244) -- the read load filter.
#include "annotation.hpp"

#include <chrono>
#include <atomic>
#include <thread>
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>

namespace lsqh {

uint32_t Dictionary::intern(const std::string &s) {
    auto it = index.find(s);
    if (it != index.end()) return it->second;
    uint32_t id = (uint32_t)names.size();
    names.push_back(s);
    index.emplace(s, id);
    return id;
}

std::vector<const char *> Dictionary::c_ptrs() const {
    std::vector<const char *> p;
    for (auto &n : names) p.push_back(n.c_str());
    return p;
}

// interval_list.hpp:462-503.  Every boundary inside [s,e) disappears; s becomes a start unless
// it sits in (start,end] of an existing interval, e becomes an end under the same test.
void IntervalSet::add(long s, long e) {
    if (!(s < e)) return;
    auto rank = [](const std::vector<long> &v, long x) {
        return (long)(std::lower_bound(v.begin(), v.end(), x) - v.begin());
    };
    long starts_lt_s = rank(starts, s), ends_lt_s = rank(ends, s);
    long starts_lt_e = rank(starts, e), ends_lt_e = rank(ends, e);
    bool s_inside = (starts_lt_s - ends_lt_s) == 1;
    bool e_inside = (starts_lt_e - ends_lt_e) == 1;
    starts.erase(starts.begin() + starts_lt_s, starts.begin() + starts_lt_e);
    ends.erase(ends.begin() + ends_lt_s, ends.begin() + ends_lt_e);
    if (!s_inside) starts.insert(starts.begin() + starts_lt_s, s);
    if (!e_inside) ends.insert(ends.begin() + ends_lt_s, e);
}

bool IntervalSet::contains(long s, long e) const {  // interval_list.hpp:396-422
    if (!(s < e)) return true;
    long i = (long)(std::upper_bound(starts.begin(), starts.end(), s) - starts.begin()) - 1;
    return i >= 0 && ends[i] >= e;
}

namespace {

typedef std::pair<long, long> Seg;

// Cut the segment that strictly contains p.
void split_at(std::vector<Seg> &v, long p) {
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].first < p && p < v[i].second) {
            Seg right(p, v[i].second);
            v[i].second = p;
            v.insert(v.begin() + i + 1, right);
            return;
        }
        if (v[i].first >= p) return;
    }
}

// ExonSet::insert (splicing_graph.h:88-169) as an arrangement update: existing segments are
// cut at the exon's end points and the parts of the exon not yet covered become new segments.
// A zero-length exon only cuts the segment it falls strictly inside (if one exists already).
void insert_exon(std::vector<Seg> &v, long s, long e) {
    split_at(v, s);
    if (e != s) split_at(v, e);
    if (!(s < e)) return;
    std::vector<Seg> gaps;
    long cur = s;
    for (size_t i = 0; i < v.size() && v[i].first < e; ++i) {
        if (v[i].second <= s) continue;
        if (v[i].first > cur) gaps.push_back(Seg(cur, v[i].first));
        cur = std::max(cur, v[i].second);
    }
    if (cur < e) gaps.push_back(Seg(cur, e));
    for (auto &g : gaps) v.insert(std::upper_bound(v.begin(), v.end(), g), g);
}

struct FileLines {  // '\n'-terminated lines only
    std::string buf;
    bool ok = false;
    explicit FileLines(const char *path) {
        FILE *f = fopen(path, "rb");
        if (!f) return;
        char tmp[1 << 16];
        size_t n;
        while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n);
        fclose(f);
        ok = true;
    }
    template <class F>
    void each(F fn) const {
        each_range(0, buf.size(), fn);
    }
    // the lines that START in [from, to) (from is a line start)
    template <class F>
    void each_range(size_t from, size_t to, F fn) const {
        size_t i = from;
        while (i < to) {
            size_t j = buf.find('\n', i);
            if (j == std::string::npos) break;  // unterminated last line: dropped
            fn(buf.data() + i, j - i);
            i = j + 1;
        }
    }
    // n + 1 line starts that cut the buffer into n pieces of about equal size
    std::vector<size_t> cut(size_t n) const {
        std::vector<size_t> c(n + 1, buf.size());
        c[0] = 0;
        for (size_t k = 1; k < n; ++k) {
            size_t j = buf.find('\n', std::max(c[k - 1], buf.size() * k / n));
            c[k] = j == std::string::npos ? buf.size() : j + 1;
        }
        return c;
    }
};

unsigned host_threads() {
    unsigned t = std::thread::hardware_concurrency();
    if (const char *e = getenv("LSQ_THREADS")) t = (unsigned)std::max(1, atoi(e));
    return std::max(1u, std::min(t, 64u));
}

inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }

// istream >> string tokenisation
struct Tokens {
    const char *p, *end;
    Tokens(const char *s, size_t n) : p(s), end(s + n) {}
    bool next(std::string &out) {
        while (p < end && is_ws(*p)) ++p;
        if (p == end) return false;
        const char *b = p;
        while (p < end && !is_ws(*p)) ++p;
        out.assign(b, p - b);
        return true;
    }
};

bool parse_long_token(const std::string &t, long &v) {
    if (t.empty()) return false;
    char *endp = nullptr;
    v = strtol(t.c_str(), &endp, 10);
    return endp != t.c_str() && *endp == '\0';
}

void split_exon_list(const std::string &s, std::vector<long> &out) {  // count.cpp:154-168
    size_t i = 0;
    while (i < s.size()) {
        size_t j = s.find(',', i);
        if (j == std::string::npos) j = s.size();
        if (j > i) out.push_back(atol(s.substr(i, j - i).c_str()));
        i = j + 1;
    }
}

bool fits_i32(long v) { return v >= INT32_MIN && v <= INT32_MAX; }

}  // namespace

namespace {

// LH_GENE_TXT (count.cpp:141-170) and UCSC_GENE_TXT (jsc/bioinfo/gene_anno.hpp:57-99, two more columns:
// cdsStart cdsEnd): one form per line, name chrom strand txStart txEnd [cdsStart cdsEnd] exonCount starts ends.
// A column that does not parse as a number ends the row as `istream >>` does: the fields after it stay empty.
void read_gene_txt(const FileLines &ff, bool with_cds, std::vector<FormRow> &rows) {
    // the lines are independent: every thread parses a stretch of the file, the rows are joined in file order
    const unsigned T = ff.buf.size() < (1u << 20) ? 1u : host_threads();
    const std::vector<size_t> cuts = ff.cut(T);
    std::vector<std::vector<FormRow> > part(T);
    auto work = [&](unsigned t) {
        std::vector<FormRow> &rows = part[t];
        ff.each_range(cuts[t], cuts[t + 1], [&](const char *s, size_t n) {
        Tokens tk(s, n);
        FormRow r;
        std::string t, cnt, es, ee;
        bool ok = tk.next(r.name) && tk.next(r.chrom) && tk.next(r.strand);
        long v;
        for (int k = 0; ok && k < (with_cds ? 4 : 2); ++k) ok = tk.next(t) && parse_long_token(t, v);
        if (ok && tk.next(cnt) && parse_long_token(cnt, v) && v >= 0) {
            r.exon_count = (uint64_t)v;
            if (tk.next(es)) split_exon_list(es, r.starts);
            if (tk.next(ee)) split_exon_list(ee, r.ends);
        }
        rows.push_back(std::move(r));
        });
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    size_t total = 0;
    for (auto &p : part) total += p.size();
    rows.reserve(rows.size() + total);
    for (auto &p : part)
        for (auto &r : p) rows.push_back(std::move(r));
}

// The GFF-like formats give one exon per line; the exons of a form are merged with interval_list semantics
// and the forms come out in bytewise order of their names (set<string> inames, solve.cpp:161-296).
struct GffForm {
    std::string chrom, strand;
    IntervalSet il;
};
void trim_quotes(std::string &s) {  // trim_if(iname, is_any_of("\""))
    size_t b = 0, e = s.size();
    while (b < e && s[b] == '"') ++b;
    while (e > b && s[e - 1] == '"') --e;
    s = s.substr(b, e - b);
}
int read_gff(const FileLines &ff, const std::string &fmt, std::vector<FormRow> &rows, std::string &err) {
    const bool ucsc = fmt == "UCSC_GFF", wb = fmt == "WORMBASE_GFF2", genelets = fmt == "GENELETS_GFF3";
    std::map<std::string, GffForm> forms;
    size_t line_no = 0;
    int rc = LSQ_OK;
    ff.each([&](const char *s, size_t n) {
        ++line_no;
        if (rc || ((ucsc || genelets) && line_no <= 2)) return;  // two header lines (solve.cpp:164-165, 236-237)
        Tokens tk(s, n);
        std::string chr, tmp, type, t0, t1, strand, tail;
        tk.next(chr);
        tk.next(tmp);
        tk.next(type);
        if (genelets && type != "exon") return;
        long start = 0, end = 0;
        if (!(tk.next(t0) && parse_long_token(t0, start) && tk.next(t1) && parse_long_token(t1, end))) {
            rc = LSQ_ERR_PARSE;  // the reference reads uninitialised longs here
            err = fmt + " line " + std::to_string(line_no) + ": start/end are not numbers";
            return;
        }
        tk.next(tmp);
        tk.next(strand);
        tk.next(tmp);
        if (wb) tk.next(tmp);
        tk.next(tail);  // form name (UCSC_GFF, WORMBASE_GFF2) or attribute list (GENELETS_GFF3)
        auto add = [&](const std::string &iname, const std::string &chrom) {
            GffForm &f = forms[iname];
            f.chrom = chrom;
            f.strand = strand;
            f.il.add(start - 1, end);
        };
        if (!genelets) {
            trim_quotes(tail);
            add(tail, ucsc ? chr : "chr" + chr);
            return;
        }
        size_t i = 0;
        while (i <= tail.size()) {  // infos split at ';', names after "Parent=" split at ','
            size_t j = tail.find(';', i);
            if (j == std::string::npos) j = tail.size();
            if (j - i > 7 && tail.compare(i, 7, "Parent=") == 0) {
                size_t a = i + 7;
                while (a <= j) {
                    size_t b = tail.find(',', a);
                    if (b == std::string::npos || b > j) b = j;
                    if (b > a) add(tail.substr(a, b - a), "chr" + chr);
                    a = b + 1;
                }
            }
            i = j + 1;
        }
    });
    if (rc) return rc;
    for (auto &kv : forms) {
        FormRow r;
        r.name = kv.first;
        r.chrom = kv.second.chrom;
        r.strand = kv.second.strand;
        r.exon_count = kv.second.il.starts.size();
        r.starts = kv.second.il.starts;
        r.ends = kv.second.il.ends;
        rows.push_back(std::move(r));
    }
    return LSQ_OK;
}

}  // namespace

int load_annotation(const char *forms_path, const char *map_path, uint64_t begin, uint64_t end,
                    uint64_t shard_begin, uint64_t shard_end, lsqh_annotation &A) {
    return load_annotation_fmt("LH_GENE_TXT", forms_path, "UCSC_GENE2ISOFORM", map_path, begin, end, shard_begin, shard_end, A);
}

int load_annotation_fmt(const char *forms_format, const char *forms_path, const char *map_format, const char *map_path,
                        uint64_t begin, uint64_t end, uint64_t shard_begin, uint64_t shard_end, lsqh_annotation &A) {
    const std::string ffmt = forms_format ? forms_format : "", mfmt = map_format ? map_format : "";
    const bool timing = getenv("LSQ_CLI_TIMING") != nullptr;  // milestones of the loader on stderr
    const auto t0 = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (timing) fprintf(stderr, "[timing] annotation_%s %.1f\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    };
    // An EMPTY file passes the reference's `assert(ifs.is_open() && !ifs.eof())` (eof is only set by a read): it
    // simply yields no rows -- no forms, or no genes and therefore no output (verified against bin/count, bin/solve).
    FileLines ff(forms_path);
    if (!ff.ok) {
        set_error(std::string("cannot read forms file ") + forms_path);
        return LSQ_ERR_IO;
    }
    std::vector<FormRow> rows;
    if (ffmt == "LH_GENE_TXT" || ffmt == "UCSC_GENE_TXT") {
        read_gene_txt(ff, ffmt == "UCSC_GENE_TXT", rows);
    } else if (ffmt == "UCSC_GFF" || ffmt == "WORMBASE_GFF2" || ffmt == "GENELETS_GFF3") {
        std::string err;
        int rc = read_gff(ff, ffmt, rows, err);
        if (rc) {
            set_error(err);
            return rc;
        }
    } else {
        set_error("Unknown file format error: " + ffmt);
        return LSQ_ERR_FORMAT;
    }
    mark("forms_parsed");
    std::unordered_map<std::string, size_t> name2row;  // later duplicates win (count.cpp:176-179)
    name2row.reserve(rows.size() * 2);
    for (size_t i = 0; i < rows.size(); ++i) name2row[rows[i].name] = i;
    mark("forms_indexed");

    FileLines mf(map_path);
    if (!mf.ok) {
        set_error(std::string("cannot read map file ") + map_path);
        return LSQ_ERR_IO;
    }
    if (mfmt != "UCSC_GENE2ISOFORM" && mfmt != "WORMBASE_GENE2ISOFORMS") {
        set_error("Unknown file format error: " + mfmt);
        return LSQ_ERR_FORMAT;
    }
    const bool wb_map = mfmt == "WORMBASE_GENE2ISOFORMS";  // `gene form;form;...` (solve.cpp:318-329)
    std::map<std::string, std::vector<long> > ev2rows;  // std::map: bytewise lexicographic order
    std::pair<const std::string, std::vector<long> > *last = nullptr;  // the rows of one event usually follow each other
    std::string g, f;
    mf.each([&](const char *s, size_t n) {
        Tokens tk(s, n);
        g.clear();
        f.clear();
        tk.next(g);
        tk.next(f);
        if (last == nullptr || last->first != g) last = &*ev2rows.emplace(g, std::vector<long>()).first;
        std::vector<long> &dst = last->second;
        auto push = [&](const std::string &name) {
            auto it = name2row.find(name);
            dst.push_back(it == name2row.end() ? -1 : (long)it->second);
        };
        if (!wb_map) {
            push(f);
            return;
        }
        size_t i = 0;
        while (i < f.size()) {
            size_t j = f.find(';', i);
            if (j == std::string::npos) j = f.size();
            if (j > i) push(f.substr(i, j - i));
            i = j + 1;
        }
    });

    mark("map_parsed");
    std::vector<const std::pair<const std::string, std::vector<long> > *> selected;
    {
        uint64_t rank = 0;
        for (auto &kv : ev2rows) {
            if (rank >= begin && rank < end) selected.push_back(&kv);
            ++rank;
        }
    }
    A.n_selected = selected.size();
    shard_end = std::min<uint64_t>(shard_end, selected.size());
    shard_begin = std::min(shard_begin, shard_end);

    // covered regions by chromosome id: add_interval depends on the ORDER of the adds (it merges an adjacent interval only
    // when the later one is appended), but the chromosomes are independent -- the adds are recorded in event order and
    // replayed per chromosome, the chromosomes side by side
    std::vector<std::vector<std::pair<long, long> > > cov_adds;
    A.ev_seg_off.push_back(0);
    A.ev_form_off.push_back(0);
    A.ev_name_off.push_back(0);
    A.form_seg_off.push_back(0);
    for (uint64_t g = 0; g < selected.size(); ++g) {
        const std::string &gname = selected[g]->first;
        const std::vector<long> &fr = selected[g]->second;
        const bool in_shard = g >= shard_begin && g < shard_end;
        std::vector<Seg> segs;
        long gs = LONG_MAX, ge = LONG_MIN;
        for (long ri : fr) {
            if (ri < 0) {
                set_error("map file names a form that is not in the forms file (event " + gname + ")");
                return LSQ_ERR_ARG;
            }
            const FormRow &r = rows[ri];
            if (r.exon_count > r.starts.size() || r.exon_count > r.ends.size()) {
                set_error("form " + r.name + ": exonCount exceeds the exon lists");
                return LSQ_ERR_ARG;
            }
            uint32_t cid = A.chroms.intern(r.chrom);
            if (cov_adds.size() <= cid) cov_adds.resize(cid + 1);
            for (uint64_t i = 0; i < r.exon_count; ++i) {
                long s = r.starts[i], e = r.ends[i];
                if (s > e) {
                    set_error("form " + r.name + ": exon start > end (reference: assertion failure)");
                    return LSQ_ERR_ARG;
                }
                if (!fits_i32(s) || !fits_i32(e)) {
                    set_error("form " + r.name + ": coordinate outside int32");
                    return LSQ_ERR_RANGE;
                }
                cov_adds[cid].push_back(std::make_pair(s, e));
                if (in_shard) insert_exon(segs, s, e);
                if (s < e) {
                    gs = std::min(gs, s);
                    ge = std::max(ge, e);
                }
            }
        }
        if (!in_shard) continue;
        if (fr.empty() || gs == LONG_MAX) {
            set_error("event " + gname + " has no exon (reference: undefined behaviour)");
            return LSQ_ERR_ARG;
        }
        if (fr.size() > LSQ_MAX_FORMS_PER_EVENT) {
            set_error("event " + gname + " has more than 64 forms");
            return LSQ_ERR_RANGE;
        }
        const FormRow &first = rows[fr[0]];
        A.ev_name.push_back(gname);
        A.ev_names_pool += gname;
        A.ev_name_off.push_back((uint32_t)A.ev_names_pool.size());
        A.ev_chrom.push_back(A.chroms.intern(first.chrom));
        A.ev_strand.push_back(A.strands.intern(first.strand));
        A.ev_gs.push_back((int32_t)gs);
        A.ev_ge.push_back((int32_t)ge);
        for (auto &sg : segs) {
            A.seg_start.push_back((int32_t)sg.first);
            A.seg_end.push_back((int32_t)sg.second);
        }
        A.ev_seg_off.push_back((uint32_t)A.seg_start.size());
        for (long ri : fr) {  // build_isoform_array, splicing_graph.h:322-360
            const FormRow &r = rows[ri];
            uint64_t cursor = 0, len = 0;
            for (size_t si = 0; si < segs.size(); ++si) {
                for (uint64_t i = cursor; i < r.exon_count; ++i) {
                    if (r.starts[i] <= segs[si].first && segs[si].second <= r.ends[i]) {
                        A.form_seg_idx.push_back((uint32_t)si);
                        len += (uint64_t)(segs[si].second - segs[si].first);
                        cursor = i;
                        break;
                    }
                }
            }
            A.form_seg_off.push_back((uint32_t)A.form_seg_idx.size());
            A.form_name.push_back(r.name);
            A.form_len.push_back(len);
        }
        A.ev_form_off.push_back((uint32_t)A.form_name.size());
    }
    mark("events_built");
    std::vector<IntervalSet> covered(std::max(cov_adds.size(), A.chroms.names.size()));
    {
        std::atomic<size_t> next(0);
        auto work = [&]() {
            for (size_t c = next++; c < cov_adds.size(); c = next++)
                for (auto &iv : cov_adds[c]) covered[c].add(iv.first, iv.second);
        };
        const unsigned T = std::min<size_t>(host_threads(), std::max<size_t>(cov_adds.size(), 1));
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back(work);
        work();
        for (auto &x : th) x.join();
    }
    mark("covered_regions");
    A.cov_off.push_back(0);
    for (auto &c : covered) {
        for (size_t i = 0; i < c.starts.size(); ++i) {
            A.cov_start.push_back((int32_t)c.starts[i]);
            A.cov_end.push_back((int32_t)c.ends[i]);
        }
        A.cov_off.push_back((uint32_t)A.cov_start.size());
    }
    A.chrom_ptrs = A.chroms.c_ptrs();
    A.strand_ptrs = A.strands.c_ptrs();
    return LSQ_OK;
}

}  // namespace lsqh
