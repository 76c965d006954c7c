// Host front end entry points (lsqh_* of include/lesseq_b200.h) and the drop-in bodies of the
// reference's two programs: same positional arguments, same stdout tables, same exit codes
// (count/count.cpp:88-501, solve/solve.cpp:102-856; only the LH_GENE_TXT / UCSC_GENE2ISOFORM /
// MRF_SINGLE branches).  All counting and estimation is done by the lsq_* device path.
#include <cerrno>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdarg>
#include <cstdio>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "annotation.hpp"
#include "mrf.hpp"

extern "C" {

int lsqh_annotation_load_fmt(const char *forms_format, const char *forms_path, const char *map_format,
                             const char *map_path, uint64_t begin, uint64_t end, uint64_t shard_begin,
                             uint64_t shard_end, lsqh_annotation **out) {
    if (!forms_format || !forms_path || !map_format || !map_path || !out) {
        lsqh::set_error("lsqh_annotation_load: null argument");
        return LSQ_ERR_ARG;
    }
    lsqh_annotation *a = new lsqh_annotation();
    int rc = lsqh::load_annotation_fmt(forms_format, forms_path, map_format, map_path, begin, end, shard_begin, shard_end, *a);
    if (rc) {
        delete a;
        return rc;
    }
    *out = a;
    return LSQ_OK;
}

int lsqh_annotation_load(const char *forms_path, const char *map_path, uint64_t begin, uint64_t end,
                         uint64_t shard_begin, uint64_t shard_end, lsqh_annotation **out) {
    return lsqh_annotation_load_fmt("LH_GENE_TXT", forms_path, "UCSC_GENE2ISOFORM", map_path, begin, end, shard_begin,
                                    shard_end, out);
}

void lsqh_annotation_free(lsqh_annotation *a) { delete a; }

int lsqh_annotation_tables(const lsqh_annotation *a, lsq_event_tables *t) {
    if (!a || !t) {
        lsqh::set_error("lsqh_annotation_tables: null argument");
        return LSQ_ERR_ARG;
    }
    memset(t, 0, sizeof *t);
    t->n_events = (uint32_t)a->ev_chrom.size();
    t->n_forms = (uint32_t)a->form_name.size();
    t->n_segments = (uint32_t)a->seg_start.size();
    t->n_chroms = (uint32_t)a->chroms.names.size();
    t->n_strands = (uint32_t)a->strands.names.size();
    t->ev_chrom = a->ev_chrom.data();
    t->ev_strand = a->ev_strand.data();
    t->ev_gs = a->ev_gs.data();
    t->ev_ge = a->ev_ge.data();
    t->ev_seg_off = a->ev_seg_off.data();
    t->ev_form_off = a->ev_form_off.data();
    t->ev_name_off = a->ev_name_off.data();
    t->ev_names = a->ev_names_pool.data();
    t->seg_start = a->seg_start.data();
    t->seg_end = a->seg_end.data();
    t->form_seg_off = a->form_seg_off.data();
    t->form_seg_idx = a->form_seg_idx.data();
    t->cov_off = a->cov_off.data();
    t->cov_start = a->cov_start.data();
    t->cov_end = a->cov_end.data();
    t->chrom_names = a->chrom_ptrs.data();
    t->strand_names = a->strand_ptrs.data();
    return LSQ_OK;
}

const char *lsqh_event_name(const lsqh_annotation *a, uint32_t e) {
    return (a && e < a->ev_name.size()) ? a->ev_name[e].c_str() : "";
}
const char *lsqh_form_name(const lsqh_annotation *a, uint32_t f) {
    return (a && f < a->form_name.size()) ? a->form_name[f].c_str() : "";
}
uint64_t lsqh_form_length(const lsqh_annotation *a, uint32_t f) {
    return (a && f < a->form_len.size()) ? a->form_len[f] : 0;
}

int lsqh_mrf_parse(const char *path, lsqh_reads **out) {
    if (!path || !out) {
        lsqh::set_error("lsqh_mrf_parse: null argument");
        return LSQ_ERR_ARG;
    }
    lsqh_reads *r = new lsqh_reads();
    int rc = lsqh::parse_mrf(path, *r);
    if (rc) {
        delete r;
        return rc;
    }
    *out = r;
    return LSQ_OK;
}

void lsqh_reads_free(lsqh_reads *r) { delete r; }

int lsqh_reads_view(const lsqh_reads *r, lsq_packed_reads *v) {
    if (!r || !v) {
        lsqh::set_error("lsqh_reads_view: null argument");
        return LSQ_ERR_ARG;
    }
    memset(v, 0, sizeof *v);
    v->n_reads = r->read_cs.size();
    v->n_blocks = r->blk_cs.size();
    v->blk_off = r->blk_off.data();
    v->blk = r->blk.data();
    v->read_cs = r->read_cs.data();
    v->blk_cs = r->any_mixed ? r->blk_cs.data() : nullptr;
    v->line_num = r->line_num.data();
    v->n_chrom_names = (uint32_t)r->chroms.names.size();
    v->n_strand_names = (uint32_t)r->strands.names.size();
    v->chrom_names = r->chrom_ptrs.data();
    v->strand_names = r->strand_ptrs.data();
    v->on_device = 0;
    v->blk_cnt = r->blk_cnt.size() == r->read_cs.size() && !r->blk_cnt.empty() ? r->blk_cnt.data() : nullptr;
    v->read_cs8 = r->read_cs8.size() == r->read_cs.size() && !r->read_cs8.empty() ? r->read_cs8.data() : nullptr;
    if (r->blk_len.size() == r->blk_cs.size() && !r->blk_len.empty()) {
        v->blk_start = r->blk_start.data();
        v->blk_len = r->blk_len.data();
        if (r->blk_len8.size() == r->blk_len.size()) v->blk_len8 = r->blk_len8.data();
    }
    return LSQ_OK;
}

int lsqh_reads_save(const lsqh_reads *r, const char *pack_path, const char *source_path) {
    if (!r || !pack_path) {
        lsqh::set_error("lsqh_reads_save: null argument");
        return LSQ_ERR_ARG;
    }
    return lsqh::save_reads(*r, pack_path, source_path);
}

int lsqh_reads_load(const char *pack_path, const char *source_path, lsqh_reads **out) {
    if (!pack_path || !out) {
        lsqh::set_error("lsqh_reads_load: null argument");
        return LSQ_ERR_ARG;
    }
    lsqh_reads *r = new lsqh_reads();
    int rc = lsqh::load_reads(pack_path, source_path, *r);
    if (rc) {
        delete r;
        return rc;
    }
    *out = r;
    return LSQ_OK;
}

int lsqh_reads_compact(lsqh_reads *r) {
    if (!r) {
        lsqh::set_error("lsqh_reads_compact: null argument");
        return LSQ_ERR_ARG;
    }
    lsqh::compact_reads(*r);
    return LSQ_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Command line programs
// ---------------------------------------------------------------------------------------------
namespace {

long g_log_level = 2;  // LogLevel::info (jsc/util/log.hpp:22-36)
const char *LEVELS[] = {"ERROR", "WARNING", "INFO", "DEBUG"};

void logf(int level, const char *fmt, ...) {  // log.hpp:58-79 line format, stderr
    if (level > g_log_level) return;
    time_t raw;
    time(&raw);
    struct tm *ti = localtime(&raw);
    char msg[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    fprintf(stderr, "[LOG %d-%02d-%02d %02d:%02d:%02d %s] %s\n", ti->tm_year + 1900, ti->tm_mon + 1, ti->tm_mday,
            ti->tm_hour, ti->tm_min, ti->tm_sec, LEVELS[level], msg);
    fflush(stderr);
}

bool lex_ulong(const char *s, unsigned long &out) {  // boost::lexical_cast<unsigned long>
    if (!s || !*s) return false;
    const char *p = s;
    if (*p == '+') ++p;
    if (!*p) return false;
    for (const char *q = p; *q; ++q)
        if (*q < '0' || *q > '9') return false;
    errno = 0;
    out = strtoul(p, nullptr, 10);
    return errno == 0;
}

bool lex_long(const char *s, long &out) {
    if (!s || !*s) return false;
    const char *p = s;
    if (*p == '+' || *p == '-') ++p;
    if (!*p) return false;
    for (const char *q = p; *q; ++q)
        if (*q < '0' || *q > '9') return false;
    errno = 0;
    out = strtol(s, nullptr, 10);
    return errno == 0;
}

bool lex_double(const char *s, double &out) {  // boost::lexical_cast<double>
    if (!s || !*s || *s == ' ' || *s == '\t') return false;
    char *endp = nullptr;
    out = strtod(s, &endp);
    return endp != s && *endp == '\0';
}

const char *USAGE_COUNT =
    "Usage:\ncount\n\tlog_level(0,1,2,...) proj_name out_prefix\n\tisoform_format isoforms_path g2i_format g2i_path "
    "gene_begin_idx gene_end_idx\n  (read_format read_type expected_read_length reads_path )+";
const char *USAGE_SOLVE =
    "Usage:\nsolve\n\tlog_level(0,1,2,...) proj_name out_prefix\n\tisoform_format isoforms_path g2i_format g2i_path "
    "gene_begin_idx gene_end_idx\n  (read_format read_type expected_read_length reads_path total_read_bases)+";

struct ReadArg {
    std::string format, type, path;
    unsigned long length = 0;
    double total_bases = 0;
};

bool file_openable(const std::string &p) {  // assert(ifs.is_open() && !ifs.eof()), count.cpp:139
    struct stat st;
    if (stat(p.c_str(), &st) == 0 && !S_ISREG(st.st_mode) && !S_ISDIR(st.st_mode))
        return access(p.c_str(), R_OK) == 0;  // a pipe (`<(zcat reads.mrf.gz)`) must be opened exactly once
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

void print_g(std::string &out, double v, int prec) {
    char b[64];
    snprintf(b, sizeof b, "%.*g", prec, v);  // std::ostream << double with precision(prec)
    out += b;
}

// LSQ_CLI_TIMING=1: wall-clock milestones of the program on stderr (milliseconds since start)
struct CliClock {
    double t0;
    bool on;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    CliClock() : t0(now()), on(getenv("LSQ_CLI_TIMING") != nullptr) {}
    void mark(const char *what) const {
        if (on) fprintf(stderr, "[timing] %s %.1f\n", what, now() - t0);
    }
};

int run_program(bool solve, int argc, char **argv) {
    const CliClock clk;
    const int group = solve ? 5 : 4;
    if (argc < 10 + group) {
        logf(0, "%s", solve ? USAGE_SOLVE : USAGE_COUNT);
        return 1;
    }
    int argi = 1;
    if (!lex_long(argv[argi++], g_log_level)) {
        fprintf(stderr, "terminate called after throwing an instance of 'bad_lexical_cast'\n");
        return 134;  // the reference aborts on an uncaught bad_lexical_cast (count.cpp:99)
    }
    argi += 2;  // proj_name, out_prefix: parsed but unused by count/solve
    std::string form_format = argv[argi++], form_path = argv[argi++];
    std::string map_format = argv[argi++], map_path = argv[argi++];
    unsigned long begin = 0, end = 0;
    std::vector<ReadArg> sets;
    bool cast_ok = lex_ulong(argv[argi++], begin) && lex_ulong(argv[argi++], end);
    while (cast_ok && argi < argc) {
        if (argc - argi < group) {
            logf(0, "%s", solve ? USAGE_SOLVE : USAGE_COUNT);
            return 1;
        }
        ReadArg ra;
        ra.format = argv[argi++];
        ra.type = argv[argi++];
        cast_ok = lex_ulong(argv[argi++], ra.length);
        ra.path = argv[argi++];
        if (solve && cast_ok) cast_ok = lex_double(argv[argi++], ra.total_bases);
        sets.push_back(ra);
    }
    if (!cast_ok) {
        logf(0, "Lexical_cast error when converting arguments to numeric values");
        return 1;
    }
    if (sets.size() > LSQ_MAX_METHODS) {
        logf(0, "more than %u read sets are not supported", LSQ_MAX_METHODS);
        return 1;
    }
    int prec = 6;
    if (const char *p = getenv("LSQ_PRECISION")) prec = atoi(p);

    // The read text does not depend on the annotation: it is parsed by a background thread from the start (the parser
    // itself uses every core), while this thread loads the annotation and a third one sets up the device.  Errors are
    // reported where the reference would meet them, so the outcomes are recorded and looked at in its order.
    const uint32_t M = (uint32_t)sets.size();
    struct Loaded {
        int rc = 0;           // LSQ_* of the parser / loader
        int stage = 0;        // 0 ok, 1 file not openable, 2 unknown format, 3 parse / load failure
        bool from_pack = false;
        std::string err, warn;
    };
    std::vector<lsqh_reads *> reads(M, nullptr);
    std::vector<Loaded> loaded(M);
    // Creating the CUDA context takes a few hundred milliseconds and depends on nothing: it starts first.  A failure
    // only matters if there turn out to be events to count.
    int device = 0;
    if (const char *d = getenv("LSQ_DEVICE")) device = atoi(d);
    lsq_ctx *ctx = nullptr;
    int dev_rc = 0;
    std::string dev_err;
    std::thread ctx_thread([&]() {
        dev_rc = lsq_ctx_create(device, &ctx);
        if (dev_rc) dev_err = lsq_last_error();
        clk.mark("device_context");
    });
    struct CtxJoiner {  // every return path waits for the context and gives it back
        std::thread &t;
        lsq_ctx *&c;
        ~CtxJoiner() {
            if (t.joinable()) t.join();
            if (c) lsq_ctx_destroy(c);
            c = nullptr;
        }
    } ctx_joiner{ctx_thread, ctx};
    std::thread read_thread([&]() {
        for (uint32_t m = 0; m < M; ++m) {
            Loaded &L = loaded[m];
            if (!file_openable(sets[m].path)) {
                L.stage = 1;
                break;  // the reference stops here
            }
            if (sets[m].format != "MRF_SINGLE") {
                L.stage = 2;
                break;
            }
            // opt-in packed cache next to the text file: `solve` reuses what `count` parsed
            const char *cache = getenv("LSQ_READ_CACHE");
            const bool use_cache = cache && cache[0] == '1';
            const std::string pack = sets[m].path + ".lsqpack";
            if (use_cache && lsqh_reads_load(pack.c_str(), sets[m].path.c_str(), &reads[m]) == LSQ_OK) {
                L.from_pack = true;
                continue;
            }
            L.rc = lsqh_mrf_parse(sets[m].path.c_str(), &reads[m]);
            if (L.rc == LSQ_OK && use_cache && lsqh_reads_save(reads[m], pack.c_str(), sets[m].path.c_str()) != LSQ_OK)
                L.warn = lsq_last_error();  // a cache that cannot be written is not an error
            if (L.rc) {
                L.stage = 3;
                L.err = lsq_last_error();
                break;
            }
            lsqh_reads_compact(reads[m]);  // the narrow layout where the values fit (9 instead of 22 bytes per read cross PCIe)
        }
        clk.mark("reads_parsed");
    });
    struct Joiner {  // every return path waits for the background loader
        std::thread &t;
        std::vector<lsqh_reads *> &r;
        bool keep = false;
        ~Joiner() {
            if (t.joinable()) t.join();
            if (!keep)
                for (auto &x : r) {
                    lsqh_reads_free(x);
                    x = nullptr;
                }
        }
    } read_joiner{read_thread, reads};

    logf(2, "Loading isoforms...");
    if (!file_openable(form_path)) {
        fprintf(stderr, "Assertion `ifs.is_open() && !ifs.eof()' failed (%s).\n", form_path.c_str());
        return 134;
    }
    // `count` reads LH_GENE_TXT only (count.cpp:141-183); `solve` also the formats of solve.cpp:158-296
    const bool form_ok = form_format == "LH_GENE_TXT" ||
                         (solve && (form_format == "UCSC_GENE_TXT" || form_format == "UCSC_GFF" ||
                                    form_format == "WORMBASE_GFF2" || form_format == "GENELETS_GFF3"));
    if (!form_ok) {
        logf(0, "Unknown file format error: %s", form_format.c_str());
        return 1;
    }
    logf(2, "Loading gene->isoform mappings ...");
    if (!file_openable(map_path)) {
        fprintf(stderr, "Assertion `ifs.is_open() && !ifs.eof()' failed (%s).\n", map_path.c_str());
        return 134;
    }
    if (map_format != "UCSC_GENE2ISOFORM" && !(solve && map_format == "WORMBASE_GENE2ISOFORMS")) {  // solve.cpp:318
        logf(0, "Unknown file format error: %s", map_format.c_str());
        return 1;
    }
    lsqh_annotation *ann = nullptr;
    int rc = lsqh_annotation_load_fmt(form_format.c_str(), form_path.c_str(), map_format.c_str(), map_path.c_str(), begin, end,
                                      0, UINT64_MAX, &ann);
    if (rc == LSQ_ERR_IO) {
        fprintf(stderr, "Assertion `ifs.is_open() && !ifs.eof()' failed: %s\n", lsq_last_error());
        return 134;
    }
    if (rc == LSQ_ERR_ARG) {  // a selected gene names a missing form, has no exon, ...: the reference dies on an assertion
        fprintf(stderr, "Assertion `px != 0' failed: %s\n", lsq_last_error());
        return 134;
    }
    if (rc) {
        logf(0, "%s", lsq_last_error());
        return 1;
    }
    lsq_event_tables tables;
    lsqh_annotation_tables(ann, &tables);
    clk.mark("annotation_loaded");
    logf(2, "Selected %llu gene(s)", (unsigned long long)ann->n_selected);

    // The event tables are uploaded by a helper thread (as soon as the context is there) while the read text is parsed.
    std::thread dev_thread;
    const bool have_events = tables.n_events != 0;
    if (have_events)
        dev_thread = std::thread([&]() {
            if (ctx_thread.joinable()) ctx_thread.join();
            if (!dev_rc) {
                dev_rc = lsq_upload_events(ctx, &tables);
                if (dev_rc) dev_err = lsq_last_error();
            }
            clk.mark("tables_uploaded");
        });
    auto join_dev = [&]() {
        if (dev_thread.joinable()) dev_thread.join();
    };

    logf(2, "Loading the reads from %u sampling method(s)", M);
    auto cleanup = [&]() {
        join_dev();
        lsqh_annotation_free(ann);  // the context goes with ctx_joiner
    };
    if (read_thread.joinable()) read_thread.join();
    for (uint32_t m = 0; m < M; ++m) {
        const Loaded &L = loaded[m];
        logf(2, "Loading the reads from sampling method #%u", m);
        if (L.stage == 1) {
            fprintf(stderr, "Assertion `ifs.is_open() && !ifs.eof()' failed (%s).\n", sets[m].path.c_str());
            cleanup();
            return 134;
        }
        if (L.stage == 2) {
            logf(0, "Unknown file format error: %s", sets[m].format.c_str());
            cleanup();
            return 1;
        }
        if (L.from_pack) logf(3, "Read pack %s.lsqpack", sets[m].path.c_str());
        if (!L.warn.empty()) logf(1, "%s", L.warn.c_str());
        if (L.stage == 3 && L.rc == LSQ_ERR_IO) {
            fprintf(stderr, "Assertion `ifs.is_open() && !ifs.eof()' failed: %s\n", L.err.c_str());
            cleanup();
            return 134;
        }
        if (L.stage == 3) {
            logf(0, "%s", L.err.c_str());
            cleanup();
            return 1;
        }
    }
    if (!have_events) {  // nothing selected: the per-gene loop never runs
        logf(2, "Processed 0 genes... Done");
        cleanup();
        return 0;
    }
    std::vector<uint32_t> types(M);
    for (uint32_t m = 0; m < M; ++m) {
        if (sets[m].type == "MEDIUM_READ") types[m] = LSQ_READ_MEDIUM;
        else if (sets[m].type == "SHORT_READ") types[m] = LSQ_READ_SHORT;
        else {  // detected inside the gene loop before anything is printed (count.cpp:417-419)
            logf(0, "Unknown read type error: %s", sets[m].type.c_str());
            cleanup();
            return 1;
        }
    }

    auto gpu_fail = [&]() {
        logf(0, "lesseq_b200 device path failed: %s", dev_err.empty() ? lsq_last_error() : dev_err.c_str());
        cleanup();
        return 70;  // no CPU fallback
    };
    join_dev();
    if (dev_rc) return gpu_fail();
    for (uint32_t m = 0; m < M; ++m) {
        // line numbers only when the file had comment lines (they ascend, so the last one tells)
        lsq_packed_reads view;
        lsqh_reads_view(reads[m], &view);
        if (view.n_reads && view.line_num && view.line_num[view.n_reads - 1] == view.n_reads) view.line_num = nullptr;
        if (lsq_set_reads(ctx, m, types[m], sets[m].length, &view)) return gpu_fail();
    }
    clk.mark("reads_on_device");
    logf(2, "Processing reads info for genes");
    if (lsq_count(ctx)) return gpu_fail();
    const uint32_t E = tables.n_events, F = tables.n_forms;
    std::vector<uint64_t> n_valid((size_t)M * E), bases((size_t)M * E), iso(F);
    lsq_count_out co{n_valid.data(), bases.data(), iso.data()};
    if (lsq_get_count(ctx, &co)) return gpu_fail();
    std::vector<double> theta(F), rpkm(F), ll(E), llpr(E);
    std::vector<uint32_t> iters(E);
    if (solve) {
        std::vector<double> trb(M);
        for (uint32_t m = 0; m < M; ++m) trb[m] = sets[m].total_bases;
        if (lsq_solve(ctx, trb.data(), M)) return gpu_fail();
        lsq_solve_out so{theta.data(), rpkm.data(), ll.data(), llpr.data(), iters.data(), nullptr};
        if (lsq_get_solve(ctx, &so)) return gpu_fail();
        // the reference warns once per read when a compatible form has no accessible start
        // (read.h:334-337); one summary line is logged here instead
    }
    clk.mark("results_on_host");
    std::string out;
    out.reserve((size_t)F * 64);
    for (uint32_t e = 0; e < E; ++e) {
        for (uint32_t f = tables.ev_form_off[e]; f < tables.ev_form_off[e + 1]; ++f) {
            out += ann->ev_name[e];
            out += '\t';
            for (uint32_t m = 0; m < M; ++m) {
                print_g(out, (double)n_valid[(size_t)m * E + e], prec);
                out += '\t';
            }
            out += ann->form_name[f];
            out += '\t';
            if (!solve) {
                print_g(out, (double)iso[f], prec);
            } else {
                print_g(out, theta[f], prec);
                out += '\t';
                print_g(out, rpkm[f], prec);
                out += '\t';
                print_g(out, llpr[e], prec);
            }
            out += '\n';
        }
        if (out.size() > (1u << 20)) {
            fwrite(out.data(), 1, out.size(), stdout);
            out.clear();
        }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    fflush(stdout);
    logf(2, "Processed %u genes... Done", E);
    clk.mark("table_printed");
    cleanup();
    clk.mark("done");
    return 0;
}

}  // namespace

extern "C" int lsqh_write_mrf(const lsq_packed_reads *r, const char *path) {
    if (!r || !path || r->on_device || !r->blk_off || !r->blk || !r->read_cs) {
        lsqh::set_error("lsqh_write_mrf: host reads in the wide layout (blk_off, blk, read_cs) are required");
        return LSQ_ERR_ARG;
    }
    FILE *f = fopen(path, "w");
    if (!f) {
        lsqh::set_error(std::string("cannot write ") + path);
        return LSQ_ERR_IO;
    }
    std::vector<char> buf(1u << 22);
    setvbuf(f, nullptr, _IOFBF, 1u << 22);
    fputs("AlignmentBlocks\n", f);
    size_t used = 0;
    auto put_num = [&](long v) {
        char tmp[24];
        int n = 0;
        if (v < 0) {
            buf[used++] = '-';
            v = -v;
        }
        do {
            tmp[n++] = (char)('0' + v % 10);
            v /= 10;
        } while (v);
        while (n) buf[used++] = tmp[--n];
    };
    auto put_str = [&](const char *p) {
        while (*p) buf[used++] = *p++;
    };
    for (uint64_t i = 0; i < r->n_reads; ++i) {
        long q = 1;
        for (uint32_t j = r->blk_off[i]; j < r->blk_off[i + 1]; ++j) {
            const uint32_t cs = (r->read_cs[i] & LSQ_READ_MIXED) && r->blk_cs ? r->blk_cs[j] : r->read_cs[i];
            if (j != r->blk_off[i]) buf[used++] = ',';
            const long s0 = (long)r->blk[2 * j] + 1, e0 = r->blk[2 * j + 1];  // 1-based inclusive in the text (count.cpp:299-318)
            put_str(r->chrom_names[cs & 0xFFFFu]);
            buf[used++] = ':';
            put_str(r->strand_names[(cs >> 16) & 0xFFu]);
            buf[used++] = ':';
            put_num(s0);
            buf[used++] = ':';
            put_num(e0);
            buf[used++] = ':';
            put_num(q);
            buf[used++] = ':';
            put_num(q + e0 - s0);
            q += e0 - s0 + 1;
        }
        buf[used++] = '\n';
        if (used > buf.size() - 4096) {
            fwrite(buf.data(), 1, used, f);
            used = 0;
        }
    }
    fwrite(buf.data(), 1, used, f);
    const bool ok = fclose(f) == 0;
    if (!ok) {
        lsqh::set_error(std::string("write error on ") + path);
        return LSQ_ERR_IO;
    }
    return LSQ_OK;
}

int lsqh_count_main(int argc, char **argv) { return run_program(false, argc, argv); }
extern "C" int lsqh_solve_main(int argc, char **argv) { return run_program(true, argc, argv); }
