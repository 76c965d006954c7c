// lesseq_oracle.cpp -- CPU restatement of LESSeq's count + solve hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under lesseq_b200/ may include, link or execute this
// file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs.  The product path is the CUDA library behind
// include/lesseq_b200.h and has no CPU fallback.
//
// Parity status: PINNED.  This restatement is checked (tests/test_oracle_golden.py) against
// outputs of the reference's own prebuilt binaries (/root/reference/bin/{count,solve}, run
// through oracle/_ref, see oracle/build_ref.sh) at 17 significant digits on the seed cases of
// SURVEY.md Appendix A and on seeded random annotation+read sets (tests/golden/).
//
// Every function cites the reference code it restates (paths relative to /root/reference).
// The arithmetic is plain sequential IEEE double (build with -O2 -ffp-contract=off), which is
// what the gcc-4.4 -O2 SSE2 reference binary executes.
//
// Usage (same positional grammar as the reference programs):
//   lesseq_oracle count log_level proj out_prefix LH_GENE_TXT forms UCSC_GENE2ISOFORM map b e
//                 (MRF_SINGLE read_type read_len reads_path)+
//   lesseq_oracle solve ... (MRF_SINGLE read_type read_len reads_path total_read_bases)+
// Environment: LSQ_ORACLE_PRECISION=17 prints round-trippable doubles (the reference binary
//   prints 6 digits; its 17-digit output is obtained with oracle/_ref/libprec17.so).
//   LSQ_ORACLE_DUMP=<path> additionally writes per-event diagnostics (EM iteration count,
//   mask-class histogram, bases) used by the GPU parity tests.

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace oracle {

// ------------------------------------------------------------------------------------------
// jsc::util::interval_list<long>: sorted disjoint half-open intervals.
// ------------------------------------------------------------------------------------------
struct IntervalList {
    std::vector<long> starts, ends;

    // interval_list.hpp:462-503.  The four lower_bound ranks decide whether each end point
    // of the new interval falls "on" an existing interval (start < p <= end).  Note the
    // asymmetry this produces: [a,b) followed by [b,c) fuses into [a,c), while [b,c) followed
    // by [a,b) leaves two abutting intervals.
    void add(long s, long e) {
        if (!(s < e)) return;
        size_t ss = std::lower_bound(starts.begin(), starts.end(), s) - starts.begin();
        size_t se = std::lower_bound(ends.begin(), ends.end(), s) - ends.begin();
        size_t es = std::lower_bound(starts.begin(), starts.end(), e) - starts.begin();
        size_t ee = std::lower_bound(ends.begin(), ends.end(), e) - ends.begin();
        bool s_on = ((long)ss - (long)se == 1);
        bool e_on = ((long)es - (long)ee == 1);
        starts.erase(starts.begin() + ss, starts.begin() + es);
        ends.erase(ends.begin() + se, ends.begin() + ee);
        if (s_on && e_on) {
        } else if (!s_on && !e_on) {
            starts.insert(starts.begin() + ss, s);
            ends.insert(ends.begin() + se, e);
        } else if (s_on && !e_on) {
            ends.insert(ends.begin() + se, e);
        } else {
            starts.insert(starts.begin() + ss, s);
        }
    }

    // interval_list.hpp:396-422.  A nil query interval is "contained".
    bool contains(long s, long e) const {
        if (!(s < e)) return true;
        long i = std::lower_bound(starts.begin(), starts.end(), s) - starts.begin();
        long n = (long)starts.size();
        if (i >= 0 && i < n && !(s < starts[i]) && !(ends[i] < e)) return true;
        if (i - 1 >= 0 && i - 1 < n && !(s < starts[i - 1]) && !(ends[i - 1] < e)) return true;
        return false;
    }

    // interval_list.hpp:348-357 (returns double in the reference).
    double total_length() const {
        double t = 0;
        for (size_t i = 0; i < starts.size(); ++i) t += ends[i] - starts[i];
        return t;
    }
    size_t size() const { return starts.size(); }
};

// ------------------------------------------------------------------------------------------
// ExonSet::insert (splicing_graph.h:88-169): refine exons into atomic segments.  The set is
// keyed by segment start (ExonComp, :65-70), so an insertion whose start already exists is
// silently dropped (std::set::insert semantics) -- kept here via map::insert.
// ------------------------------------------------------------------------------------------
struct SegmentSet {
    std::map<long, long> seg;  // start -> end

    void insert(long ns, long ne) {
        std::vector<std::pair<long, long> > extra;
        std::map<long, long>::iterator lo = seg.lower_bound(ns);
        if (!seg.empty() && lo != seg.begin()) --lo;  // :94-97
        for (std::map<long, long>::iterator it = lo; it != seg.end() && it->first < ne; ++it) {
            long os = it->first;
            long &oe = it->second;
            if (os == ns) {  // :101-116
                if (oe > ne) {
                    extra.push_back(std::make_pair(ne, oe));
                    oe = ne;
                    ns = ne;
                } else {
                    ns = oe;
                }
            } else if (os < ns) {  // :117-140
                if (oe > ns) {
                    if (oe > ne) {
                        extra.push_back(std::make_pair(ns, ne));
                        extra.push_back(std::make_pair(ne, oe));
                        oe = ns;
                        ns = ne;
                    } else {
                        extra.push_back(std::make_pair(ns, oe));
                        long old_end = oe;
                        oe = ns;
                        ns = old_end;
                    }
                }
            } else {  // os > ns, :141-156
                if (oe > ne) {
                    extra.push_back(std::make_pair(ne, oe));
                    oe = ne;
                    ne = os;
                } else {
                    extra.push_back(std::make_pair(ns, os));
                    ns = oe;
                }
            }
        }
        if (ns < ne) seg.insert(std::make_pair(ns, ne));  // :159-161
        for (size_t i = 0; i < extra.size(); ++i)          // :162-166
            if (extra[i].first < extra[i].second) seg.insert(extra[i]);
    }
};

// One LH_GENE_TXT row (jsc/bioinfo/gene_anno.hpp:35-51; loader count/count.cpp:141-171).
struct Form {
    std::string name, chrom, strand;
    long txStart, txEnd;
    unsigned long exonCount;
    std::vector<long> exonStarts, exonEnds;
    Form() : txStart(0), txEnd(0), exonCount(0) {}
};

static void split_atol(const std::string &s, std::vector<long> &out) {
    // boost::tokenizer<char_separator<char>>(",") drops empty tokens; atol per token
    // (count/count.cpp:154-168).
    size_t i = 0;
    while (i < s.size()) {
        size_t j = s.find(',', i);
        if (j == std::string::npos) j = s.size();
        if (j > i) out.push_back(atol(s.substr(i, j - i).c_str()));
        i = j + 1;
    }
}

// Per-event model: Isoforms::build / build_isoform_array (splicing_graph.h:235-252, 318-361).
struct Event {
    std::string gname;
    std::vector<const Form *> forms;
    std::vector<std::pair<long, long> > segs;             // atomic segments, sorted
    std::vector<unsigned long> seg_len;                    // exon_lengths (:340-346)
    std::vector<std::vector<unsigned long> > form_idx;     // known_iso_exon_indices
    std::vector<std::vector<unsigned long> > form_cum;     // known_iso_exon_total_lengths
    std::string chrom, strand;                             // from the first form (:239-242)
    long gene_start, gene_end;                             // gname2il first start / last end

    void build(const SegmentSet &ss) {
        for (std::map<long, long>::const_iterator it = ss.seg.begin(); it != ss.seg.end(); ++it)
            segs.push_back(*it);
        size_t S = segs.size(), K = forms.size();
        if (K > 0) {
            chrom = forms[0]->chrom;
            strand = forms[0]->strand;
        }
        std::vector<std::vector<int> > member(K, std::vector<int>(S, 0));
        for (size_t k = 0; k < K; ++k) {  // :322-338, monotone exon cursor
            const Form &f = *forms[k];
            unsigned long cur = 0;
            for (size_t s = 0; s < S; ++s) {
                for (unsigned long i = cur; i < f.exonCount; ++i) {
                    if (f.exonStarts[i] <= segs[s].first && segs[s].second <= f.exonEnds[i]) {
                        member[k][s] = 1;
                        cur = i;
                        break;
                    }
                }
            }
        }
        seg_len.assign(S, 0);
        for (size_t s = 0; s < S; ++s) seg_len[s] = segs[s].second - segs[s].first;
        form_idx.resize(K);
        form_cum.resize(K);
        for (size_t k = 0; k < K; ++k) {  // :348-360
            unsigned long tot = 0;
            for (size_t s = 0; s < S; ++s)
                if (member[k][s]) {
                    form_idx[k].push_back(s);
                    tot += seg_len[s];
                    form_cum[k].push_back(tot);
                }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Accessible read starts: number of admissible read start positions per form.
// MEDIUM_READ: AccessibleReadStarts::construct_ARS (accessible_read_starts.h:48-89)
// SHORT_READ : AccessibleShortReadStarts::construct_ARS (:221-274), min_partial_exon_size = 0
// Returned value = get_iso_ARS_total_length (:184-189).
// ------------------------------------------------------------------------------------------
static unsigned long ars_total_length(const Event &ev, size_t k, bool short_read,
                                      unsigned long read_length) {
    const std::vector<unsigned long> &idx = ev.form_idx[k];
    const unsigned long min_partial = 0;
    unsigned long total = 0, iso_len = 0, iso_total = 0;
    for (size_t i = 0; i < idx.size(); ++i) iso_total += ev.seg_len[idx[i]];
    for (size_t i = 0; i < idx.size(); ++i) {
        unsigned long L = ev.seg_len[idx[i]];
        IntervalList ars;  // the reference uses interval_list<unsigned long>; values are >= 0
        iso_len += L;
        if (iso_len + read_length > iso_total) {
            long last2 = std::max((long)L + 1 - (long)(iso_len + read_length - iso_total), (long)0);
            ars.add(0, last2);
            total += (unsigned long)ars.total_length();
            break;  // later segments contribute empty lists
        }
        if (short_read) {
            long last = std::max((long)L - (long)read_length + 1, (long)0);
            ars.add(0, last);
            long pos = std::max((long)L - (long)read_length + (long)min_partial, (long)0);
            if (L > min_partial) ars.add(pos, (long)(L - min_partial + 1));
        } else {
            ars.add(0, (long)L);
        }
        total += (unsigned long)ars.total_length();
    }
    return total;
}

// ------------------------------------------------------------------------------------------
// Read_single::build (common/read.h:204-274): two-pointer walk of the read's blocks over the
// event's segments.  Produces the sorted set of segment indices and the matched length.
// ------------------------------------------------------------------------------------------
struct ReadModel {
    std::vector<unsigned long> seg_indices;
    unsigned long matched;
};

static void build_read(const IntervalList &il, const std::vector<std::pair<long, long> > &segs,
                       ReadModel &rm) {
    std::set<unsigned long> index_set;
    unsigned long matched = 0;
    bool found_start = false;
    size_t b = 0, it = 0;
    long cur_start = 0, cur_end = 0, cont = 0;  // cont == cur_exon_start in the reference
    while (b < il.size()) {
        cur_start = il.starts[b];
        cur_end = il.ends[b];
        while (it < segs.size() && segs[it].first < cur_end) {
            long s = segs[it].first, e = segs[it].second;
            cont = std::max(cont, s);
            if (cur_start >= cont && cur_start < e) {
                if (!found_start) {
                    found_start = true;
                } else if (cur_start > cont) {
                    break;
                }
                index_set.insert(it);
                cont = std::min(e, cur_end);
                matched += cont - cur_start;
                if (cur_end < e) {
                    cur_start = cur_end;
                    ++b;
                    break;
                } else if (cur_end == e) {
                    cur_start = cur_end;
                    ++b;
                } else {
                    cur_start = e;
                }
            } else if (cont > s && cont < e) {
                break;
            }
            ++it;
        }
        if (cur_start == cur_end) continue;
        break;
    }
    rm.seg_indices.assign(index_set.begin(), index_set.end());
    rm.matched = matched;
}

// is_connected_exons_compatible_with_isoform (common/read.h:44-79): the read's index list
// must appear as a contiguous run of the form's index list.
static bool compatible(const std::vector<unsigned long> &r, const std::vector<unsigned long> &f) {
    size_t i = 0, j = 0;
    bool ok = false, found = false;
    while (i < r.size()) {
        if (j == f.size()) {
            ok = false;
            break;
        }
        if (r[i] != f[j]) {
            if (!found) {
                ++j;
            } else {
                ok = false;
                break;
            }
        } else {
            found = true;
            ok = true;
            ++i;
            ++j;
        }
    }
    return ok;
}

// ------------------------------------------------------------------------------------------
// EM (common/read.h:592-660).  rows[m] = dense N_m x K matrices in valid-read order.
// ------------------------------------------------------------------------------------------
typedef std::vector<std::vector<double> > Matrix;  // row-major rows

static double log_likelihood(size_t K, const std::vector<Matrix> &mats,
                             const std::vector<double> &theta) {  // :620-636
    double ll = 0;
    for (size_t m = 0; m < mats.size(); ++m)
        for (size_t i = 0; i < mats[m].size(); ++i) {
            double s = 0;
            for (size_t k = 0; k < K; ++k) s += theta[k] * mats[m][i][k];
            ll += log(s);
        }
    return ll;
}

static void em_step(size_t K, const std::vector<Matrix> &mats, const std::vector<double> &old_t,
                    std::vector<double> &new_t) {  // :592-618
    for (size_t k = 0; k < K; ++k) {
        double sum_zeta = 0, n_total = 0;
        for (size_t m = 0; m < mats.size(); ++m) {
            size_t n = mats[m].size();
            n_total += (double)n;
            for (size_t i = 0; i < n; ++i) {
                double s = 0;
                for (size_t k2 = 0; k2 < K; ++k2) s += old_t[k2] * mats[m][i][k2];
                if (s > 0) {
                    double lp = old_t[k] * mats[m][i][k];
                    if (lp > 0) sum_zeta += lp / s;
                }
            }
        }
        new_t[k] = sum_zeta / n_total;
    }
}

static unsigned long em(size_t K, const std::vector<Matrix> &mats, std::vector<double> &theta) {
    theta.resize(K, 1.0 / (double)K);  // :642-643
    double ll, old_ll;
    std::vector<double> old_t;
    unsigned long iters = 0;
    do {  // :647-659, no iteration cap in the reference
        old_t = theta;
        old_ll = log_likelihood(K, mats, old_t);
        em_step(K, mats, old_t, theta);
        ll = log_likelihood(K, mats, theta);
        ++iters;
    } while (std::fabs(1.0 - old_ll / ll) > 1E-6);
    return iters;
}

// ------------------------------------------------------------------------------------------
// Driver: count/count.cpp:88-501 and solve/solve.cpp:102-856 (LH_GENE_TXT, UCSC_GENE2ISOFORM,
// MRF_SINGLE branches only).
// ------------------------------------------------------------------------------------------
struct LoadedRead {
    unsigned long line_num;
    std::string name;  // "read-<line_num>"
    std::string chrom, strand;
    IntervalList il;
    long start, end;
};

struct ReadSet {
    std::string format, type, path;
    unsigned long read_length;
    double total_read_bases;
    std::vector<LoadedRead> reads;     // retained reads
    std::vector<size_t> index;         // permutation = the reference's std::set order
};

struct ReadKeyLess {  // RinfopComp, count/count.cpp:64-85
    bool operator()(const LoadedRead &a, const LoadedRead &b) const {
        if (a.chrom == b.chrom) {
            if (a.start == b.start) {
                if (a.end == b.end) {
                    if (a.strand == b.strand) return a.name < b.name;
                    return a.strand < b.strand;
                }
                return a.end < b.end;
            }
            return a.start < b.start;
        }
        return a.chrom < b.chrom;
    }
};

static int fail(const std::string &msg) {
    std::cerr << "[LOG ERROR] " << msg << std::endl;
    return 1;
}

// boost::lexical_cast<long>/<unsigned long>: the whole token must be a number.
static bool lex_long(const std::string &s, long &out) {
    if (s.empty()) return false;
    size_t i = 0;
    if (s[0] == '+' || s[0] == '-') i = 1;
    if (i == s.size()) return false;
    for (size_t j = i; j < s.size(); ++j)
        if (s[j] < '0' || s[j] > '9') return false;
    errno = 0;
    out = strtol(s.c_str(), NULL, 10);
    return errno == 0;
}

// MRF_SINGLE loader with the per-block covered-region filter (count/count.cpp:279-339).
static int load_mrf(ReadSet &rs, std::map<std::string, IntervalList> &covered) {
    std::ifstream ifs(rs.path.c_str());
    if (!ifs.is_open() || ifs.eof()) {
        std::cerr << "oracle: cannot open " << rs.path << std::endl;
        abort();  // assert(ifs.is_open() && !ifs.eof()) -> SIGABRT (:278)
    }
    std::string line;
    std::getline(ifs, line);  // first line skipped unconditionally (:283)
    unsigned long line_num = 0;
    const std::string::size_type npos = std::string::npos;
    while (std::getline(ifs, line) && !ifs.eof()) {  // last unterminated line is dropped
        line_num++;
        if ("#" == line.substr(0, 1) || "AlignmentBlocks" == line) continue;
        LoadedRead rd;
        bool registered = false;
        std::string::size_type last_comma = 0, colon, old_colon;
        while (last_comma != npos) {
            std::string::size_type b0 = (last_comma == 0 ? 0 : last_comma + 1);
            colon = line.find(':', last_comma);
            std::string chr = line.substr(b0, colon - b0);
            old_colon = colon;
            colon = line.find(':', colon + 1);
            std::string strand = line.substr(old_colon + 1, colon - old_colon - 1);
            old_colon = colon;
            colon = line.find(':', colon + 1);
            long start, end;
            if (!lex_long(line.substr(old_colon + 1, colon - old_colon - 1), start))
                return fail("Lexical_cast error when converting arguments to numeric values");
            old_colon = colon;
            colon = line.find(':', colon + 1);
            if (!lex_long(line.substr(old_colon + 1, colon - old_colon - 1), end))
                return fail("Lexical_cast error when converting arguments to numeric values");
            if (covered[chr].contains(start - 1, end)) {  // :319-324
                registered = true;
                rd.chrom = chr;
                rd.strand = strand;
                rd.il.add(start - 1, end);
            }
            last_comma = line.find(',', colon);
        }
        if (registered) {
            if (rd.il.size() == 0) {
                // Only nil blocks were "contained": the reference then reads starts[0] of an
                // empty vector (undefined behaviour, :359).  Such reads are dropped here and in
                // the product; documented in DESIGN.md.
                continue;
            }
            rd.line_num = line_num;
            std::ostringstream ss;
            ss << "read-" << line_num;
            rd.name = ss.str();
            rd.start = rd.il.starts[0];
            rd.end = rd.il.ends[rd.il.size() - 1];
            rs.reads.push_back(rd);
        }
    }
    rs.index.resize(rs.reads.size());
    for (size_t i = 0; i < rs.index.size(); ++i) rs.index[i] = i;
    ReadKeyLess less;
    const std::vector<LoadedRead> &R = rs.reads;
    std::sort(rs.index.begin(), rs.index.end(),
              [&](size_t a, size_t b) { return less(R[a], R[b]); });  // :346-364
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Standard errors (dump only; no reference program evaluates fim.h, so there is no reference output to pin this
// to: it restates the reference's FORMULAS so that the product's standard errors have a checker that is not its own).
// ---------------------------------------------------------------------------------------------

// FIM::ofim, common/fim.h:322-367, summed over the event's valid reads: per read the (K-1)x(K-1) matrix
// (v_p - v_K)(v_q - v_K) / s^2 with v_k = G_k [read compatible with form k] and s = sum_k theta_k v_k over the
// terms with v_k > 0 and theta_k > 0; evaluated, like the reference, in the log domain with a sign matrix.
static bool observed_information(size_t K, const std::vector<Matrix> &mats, const std::vector<double> &theta, Matrix &I) {
    I.assign(K - 1, std::vector<double>(K - 1, 0.0));
    for (size_t m = 0; m < mats.size(); ++m)
        for (size_t i = 0; i < mats[m].size(); ++i) {
            const std::vector<double> &v = mats[m][i];
            double sum = 0;
            for (size_t k = 0; k < K; ++k)
                if (v[k] > 0 && theta[k] > 0) sum += theta[k] * v[k];
            if (!(sum > 0)) return false;  // the reference asserts sum > 0
            const double log_prod = log(sum) + log(sum);
            for (size_t p = 0; p + 1 < K; ++p)
                for (size_t q = 0; q + 1 < K; ++q) {
                    double t = (v[p] - v[K - 1]) * (v[q] - v[K - 1]);
                    if (t != 0) {
                        const int sign = t > 0 ? 1 : -1;
                        if (t < 0) t = -t;
                        I[p][q] += sign * exp(log(t) - log_prod);
                    }
                }
        }
    return true;
}

// linalg::invert_matrix, common/linalg.h:28-71 = gsl_linalg_LU_decomp + gsl_linalg_LU_invert (GSL 1.12): LU with
// partial pivoting (largest magnitude in the column), then the inverse column by column from the unit vectors.
static bool lu_invert(const Matrix &A, Matrix &inv) {
    const size_t n = A.size();
    Matrix lu = A;
    std::vector<size_t> perm(n);
    for (size_t i = 0; i < n; ++i) perm[i] = i;
    for (size_t j = 0; j + 1 < n || (j < n && n == 1); ++j) {
        size_t piv = j;
        double best = fabs(lu[j][j]);
        for (size_t i = j + 1; i < n; ++i)
            if (fabs(lu[i][j]) > best) {
                best = fabs(lu[i][j]);
                piv = i;
            }
        if (piv != j) {
            std::swap(lu[j], lu[piv]);
            std::swap(perm[j], perm[piv]);
        }
        if (lu[j][j] != 0.0)
            for (size_t i = j + 1; i < n; ++i) {
                const double f = lu[i][j] / lu[j][j];
                lu[i][j] = f;
                for (size_t c = j + 1; c < n; ++c) lu[i][c] -= f * lu[j][c];
            }
        if (n == 1) break;
    }
    for (size_t i = 0; i < n; ++i)
        if (lu[i][i] == 0.0) return false;  // singular
    inv.assign(n, std::vector<double>(n, 0.0));
    for (size_t c = 0; c < n; ++c) {  // solve L U x = P e_c
        std::vector<double> x(n, 0.0);
        for (size_t i = 0; i < n; ++i) x[i] = perm[i] == c ? 1.0 : 0.0;
        for (size_t i = 0; i < n; ++i)
            for (size_t k = 0; k < i; ++k) x[i] -= lu[i][k] * x[k];
        for (size_t ii = n; ii-- > 0;) {
            for (size_t k = ii + 1; k < n; ++k) x[ii] -= lu[ii][k] * x[k];
            x[ii] /= lu[ii][ii];
        }
        for (size_t i = 0; i < n; ++i) inv[i][c] = x[i];
    }
    return true;
}

// The variance decomposition of FIM::estimate_mle_variance_by_inv (fim.h:75-93: the sum of all entries of the inverse
// plus its trace = var(theta_{K-1}) + sum_p var(theta_p)), form by form: se_p = sqrt(inv(p,p)), se_{K-1} = sqrt(sum of
// all entries).  Returns false where the reference's own code would fail (no reads, s = 0, singular information).
static bool standard_errors(size_t K, const std::vector<Matrix> &mats, const std::vector<double> &theta, std::vector<double> &se) {
    se.assign(K, 0.0);
    if (K == 1) return true;
    if (mats.empty()) return false;
    Matrix I, inv;
    if (!observed_information(K, mats, theta, I) || !lu_invert(I, inv)) return false;
    double total = 0;
    for (size_t p = 0; p + 1 < K; ++p) {
        for (size_t q = 0; q + 1 < K; ++q) total += inv[p][q];
        se[p] = sqrt(inv[p][p]);
    }
    se[K - 1] = sqrt(total);
    return true;
}

static int run(bool solve_mode, int argc, char **argv) {
    int group = solve_mode ? 5 : 4;
    if (argc < 10 + group) return fail("Usage error");  // argc < 14 / 15 incl. program name
    int argi = 1;
    long log_level;
    if (!lex_long(argv[argi++], log_level)) abort();  // uncaught bad_lexical_cast in reference
    argi += 2;                                        // proj_name, out_prefix: unused
    std::string form_format = argv[argi++], form_path = argv[argi++];
    std::string map_format = argv[argi++], map_path = argv[argi++];
    long lb, le;
    if (!lex_long(argv[argi++], lb) || lb < 0)
        return fail("Lexical_cast error when converting arguments to numeric values");
    if (!lex_long(argv[argi++], le) || le < 0)
        return fail("Lexical_cast error when converting arguments to numeric values");
    unsigned long gene_begin = lb, gene_end = le;
    std::vector<ReadSet> sets;
    while (argi < argc) {
        if (argc - argi < group) return fail("Usage error");
        ReadSet rs;
        rs.format = argv[argi++];
        rs.type = argv[argi++];
        long rl;
        if (!lex_long(argv[argi++], rl) || rl < 0)
            return fail("Lexical_cast error when converting arguments to numeric values");
        rs.read_length = rl;
        rs.path = argv[argi++];
        rs.total_read_bases = 0;
        if (solve_mode) {
            char *endp = NULL;
            const char *a = argv[argi++];
            rs.total_read_bases = strtod(a, &endp);
            if (endp == a || *endp != '\0')
                return fail("Lexical_cast error when converting arguments to numeric values");
        }
        sets.push_back(rs);
    }

    // forms (count.cpp:136-180)
    std::vector<Form> forms;
    {
        std::ifstream ifs(form_path.c_str());
        if (!ifs.is_open() || ifs.eof()) abort();
        if (form_format != "LH_GENE_TXT") return fail("Unknown file format error: " + form_format);
        std::string line;
        while (std::getline(ifs, line) && !ifs.eof()) {
            std::istringstream iss(line);
            Form f;
            std::string s1, s2;
            iss >> f.name >> f.chrom >> f.strand >> f.txStart >> f.txEnd >> f.exonCount >> s1 >> s2;
            split_atol(s1, f.exonStarts);
            split_atol(s2, f.exonEnds);
            forms.push_back(f);
        }
    }
    std::map<std::string, const Form *> name2form;  // later duplicates win (:176-179)
    for (size_t i = 0; i < forms.size(); ++i) name2form[forms[i].name] = &forms[i];

    // event -> forms map (count.cpp:182-200)
    std::map<std::string, std::vector<const Form *> > g2forms;
    std::set<std::string> gname_set;
    {
        std::ifstream ifs(map_path.c_str());
        if (!ifs.is_open() || ifs.eof()) abort();
        if (map_format != "UCSC_GENE2ISOFORM")
            return fail("Unknown file format error: " + map_format);
        std::string line;
        while (std::getline(ifs, line) && !ifs.eof()) {
            std::istringstream iss(line);
            std::string g, f;
            iss >> g >> f;
            gname_set.insert(g);
            g2forms[g].push_back(name2form.count(f) ? name2form[f] : NULL);
        }
    }

    // selection by lexicographic rank (count.cpp:204-215)
    std::vector<std::string> gnames;
    {
        unsigned long rank = 0;
        for (std::set<std::string>::const_iterator it = gname_set.begin(); it != gname_set.end();
             ++it, ++rank)
            if (rank >= gene_begin && rank < gene_end) gnames.push_back(*it);
    }

    // per-event structures + covered regions (count.cpp:231-258)
    std::map<std::string, IntervalList> covered;
    std::vector<Event> events(gnames.size());
    for (size_t g = 0; g < gnames.size(); ++g) {
        Event &ev = events[g];
        ev.gname = gnames[g];
        ev.forms = g2forms[gnames[g]];
        SegmentSet ss;
        IntervalList span;
        for (size_t k = 0; k < ev.forms.size(); ++k) {
            const Form *f = ev.forms[k];
            if (!f) {
                std::cerr << "oracle: map row names an unknown form (reference dereferences NULL)"
                          << std::endl;
                abort();
            }
            for (unsigned long i = 0; i < f->exonCount; ++i) {
                ss.insert(f->exonStarts[i], f->exonEnds[i]);
                covered[f->chrom].add(f->exonStarts[i], f->exonEnds[i]);
                span.add(f->exonStarts[i], f->exonEnds[i]);
            }
        }
        ev.build(ss);
        ev.gene_start = span.size() ? span.starts[0] : 0;
        ev.gene_end = span.size() ? span.ends[span.size() - 1] : 0;
    }

    // reads (count.cpp:262-364)
    for (size_t m = 0; m < sets.size(); ++m) {
        if (sets[m].format != "MRF_SINGLE")
            return fail("Unknown file format error: " + sets[m].format);
        int rc = load_mrf(sets[m], covered);
        if (rc) return rc;
    }

    int prec = 6;
    if (const char *p = getenv("LSQ_ORACLE_PRECISION")) prec = atoi(p);
    std::cout.precision(prec);
    FILE *dump = NULL;
    if (const char *p = getenv("LSQ_ORACLE_DUMP")) dump = fopen(p, "w");

    const size_t M = sets.size();
    ReadKeyLess less;
    for (size_t g = 0; g < events.size(); ++g) {  // count.cpp:369-497 / solve.cpp:668-852
        const Event &ev = events[g];
        const size_t K = ev.forms.size();
        std::vector<double> supports, support_bases;
        std::vector<double> iso_count(K, 0.0);
        std::vector<Matrix> mats;
        std::vector<std::map<unsigned long long, unsigned long> > class_hist(M);
        for (size_t m = 0; m < M; ++m) {
            ReadSet &rs = sets[m];
            bool short_read;
            if (rs.type == "MEDIUM_READ") short_read = false;
            else if (rs.type == "SHORT_READ") short_read = true;
            else return fail("Unknown read type error: " + rs.type);
            std::vector<unsigned long> ars(K);
            for (size_t k = 0; k < K; ++k) ars[k] = ars_total_length(ev, k, short_read, rs.read_length);

            LoadedRead probe;  // count.cpp:423-430
            probe.name = ev.gname;
            probe.chrom = ev.chrom;
            probe.strand = ev.strand;
            probe.start = ev.gene_start;
            probe.end = ev.gene_end;
            const std::vector<LoadedRead> &R = rs.reads;
            std::vector<size_t>::const_iterator it = std::lower_bound(
                rs.index.begin(), rs.index.end(), probe,
                [&](size_t a, const LoadedRead &p) { return less(R[a], p); });
            std::vector<size_t> valid;
            double bases = 0;
            ReadModel rm;
            for (; it != rs.index.end() && R[*it].chrom == probe.chrom && R[*it].start <= probe.end;
                 ++it) {
                const LoadedRead &rd = R[*it];
                if (rd.start < probe.start) continue;
                build_read(rd.il, ev.segs, rm);
                for (size_t k = 0; k < K; ++k) {
                    if (compatible(rm.seg_indices, ev.form_idx[k])) {
                        if ((double)rm.matched / rd.il.total_length() > 0.98) {  // :441-446
                            valid.push_back(*it);
                            bases += rm.matched;  // solve.cpp:744
                            break;
                        }
                    }
                }
            }
            Matrix mat(valid.size(), std::vector<double>(K, 0.0));
            for (size_t i = 0; i < valid.size(); ++i) {  // count.cpp:470-481 / solve.cpp:773-787
                const LoadedRead &rd = R[valid[i]];
                build_read(rd.il, ev.segs, rm);
                unsigned long long mask = 0;
                for (size_t k = 0; k < K; ++k) {
                    if (compatible(rm.seg_indices, ev.form_idx[k])) {
                        iso_count[k] += 1;
                        if (k < 64) mask |= 1ull << k;
                        double nd = (double)ars[k];  // prob_generated_by_iso, read.h:331-340
                        mat[i][k] = (nd <= 0) ? 0.0 : (double)1.0 / nd;
                    }
                }
                class_hist[m][mask]++;
            }
            if (!valid.empty()) mats.push_back(mat);
            supports.push_back((double)valid.size());
            support_bases.push_back(bases);
        }

        if (!solve_mode) {  // count.cpp:486-492
            for (size_t k = 0; k < K; ++k) {
                std::cout << ev.gname << "\t";
                for (size_t m = 0; m < M; ++m) std::cout << supports[m] << "\t";
                std::cout << ev.forms[k]->name << "\t" << iso_count[k] << std::endl;
            }
        } else {  // solve.cpp:796-847
            std::vector<double> theta;
            unsigned long iters = 0;
            if (mats.empty()) theta.resize(K, 1.0 / (double)K);
            else if (K == 1) theta.push_back(1);
            else iters = em(K, mats, theta);
            std::vector<double> rpkm(K, 0.0);
            double total_mb = 0.0;
            for (size_t m = 0; m < M; ++m) {
                total_mb += sets[m].total_read_bases / 1.0E6;
                for (size_t k = 0; k < K; ++k) rpkm[k] += (double)(support_bases[m]) * theta[k];
            }
            for (size_t k = 0; k < K; ++k) {
                rpkm[k] /= ((double)(ev.form_cum[k].back()) / 1.0E3);
                rpkm[k] /= total_mb;
            }
            double logll = log_likelihood(K, mats, theta);
            double sum_supports = 0.0;
            for (size_t m = 0; m < M; ++m) sum_supports += supports[m];
            for (size_t k = 0; k < K; ++k) {
                std::cout << ev.gname << "\t";
                for (size_t m = 0; m < M; ++m) std::cout << supports[m] << "\t";
                std::cout << ev.forms[k]->name << "\t" << theta[k] << "\t" << rpkm[k];
                if (sum_supports > 1E-5) std::cout << "\t" << (logll / sum_supports) << std::endl;
                else std::cout << "\t" << 0 << std::endl;
            }
            if (dump) fprintf(dump, "E\t%s\titers\t%lu\n", ev.gname.c_str(), iters);
            if (dump)  // G records: G_k = 1 / ARS_k per read set, as the S records below use it
                for (size_t m = 0; m < M; ++m) {
                    const bool sr = sets[m].type == "SHORT_READ";
                    for (size_t k = 0; k < K; ++k) {
                        const unsigned long a = ars_total_length(ev, k, sr, sets[m].read_length);
                        fprintf(dump, "G\t%s\t%zu\t%zu\t%.17g\n", ev.gname.c_str(), m, k, a ? 1.0 / (double)a : 0.0);
                    }
                }
            if (dump) {  // S records: standard errors from the observed information at the estimate (see standard_errors)
                std::vector<double> se;
                const bool ok = standard_errors(K, mats, theta, se);
                for (size_t k = 0; k < K; ++k) fprintf(dump, "S\t%s\t%zu\t%.17g\n", ev.gname.c_str(), k, ok ? se[k] : NAN);
            }
        }
        if (dump) {
            for (size_t m = 0; m < M; ++m) {
                fprintf(dump, "B\t%s\t%zu\t%.0f\t%.0f\n", ev.gname.c_str(), m, supports[m],
                        support_bases[m]);
                for (std::map<unsigned long long, unsigned long>::const_iterator c =
                         class_hist[m].begin();
                     c != class_hist[m].end(); ++c)
                    fprintf(dump, "C\t%s\t%zu\t%llu\t%lu\n", ev.gname.c_str(), m, c->first,
                            c->second);
            }
        }
    }
    if (dump) fclose(dump);
    return 0;
}

}  // namespace oracle

int main(int argc, char **argv) {
    if (argc < 2) {
        std::cerr << "usage: lesseq_oracle {count|solve} <reference positional arguments>\n";
        return 2;
    }
    std::string cmd = argv[1];
    if (cmd != "count" && cmd != "solve") {
        std::cerr << "unknown subcommand " << cmd << "\n";
        return 2;
    }
    // shift so that argv[1] is log_level, as in the reference programs
    return oracle::run(cmd == "solve", argc - 1, argv + 1);
}
