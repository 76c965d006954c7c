/* lesseq_b200.h -- C ABI of the B200-native count + solve hot path of LESSeq.
 *
 * The reference (gersteinlab/LESSeq) has no plugin / operator / FFI interface: its hot path is
 * inlined in two main() functions (count/count.cpp:88-501, solve/solve.cpp:102-856) and the
 * only stable contract is argv + input file formats + stdout.  This header therefore defines
 * the seams a maintainer would cut into those two programs; every entry point names the
 * reference code it replaces.  The library behind it is CUDA for sm_100a only; there is no
 * CPU fallback -- every call fails with LSQ_ERR_CUDA when no device is usable.
 *
 * Two groups of entry points live in the same shared library (liblesseq_b200.so):
 *
 *   lsq_*   the device path: event tables + packed reads in, per-event / per-form results out.
 *           Plain pointers and sizes only.
 *   lsqh_*  the host front end that produces those inputs from the reference's file formats
 *           (LH_GENE_TXT, UCSC_GENE2ISOFORM, MRF_SINGLE) and formats the reference's stdout
 *           tables; lsqh_count_main / lsqh_solve_main are drop-in bodies for the two programs.
 *
 * Conventions: every function returns 0 on success or a negative LSQ_ERR_* code; the message
 * of the last failure on the calling thread is available from lsq_last_error().  No C++
 * exception crosses the boundary.  A context is bound to one device and is not thread-safe.
 * Coordinates are 0-based half-open int32 (the reference uses long; coordinates outside
 * int32 are rejected by the host front end with LSQ_ERR_RANGE).
 */
#ifndef LESSEQ_B200_H_
#define LESSEQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LSQ_OK 0
#define LSQ_ERR_ARG (-1)     /* bad argument / inconsistent tables */
#define LSQ_ERR_CUDA (-2)    /* CUDA runtime failure (including "no device") */
#define LSQ_ERR_STATE (-3)   /* call order violated (e.g. solve before count) */
#define LSQ_ERR_RANGE (-4)   /* a documented capacity limit was exceeded */
#define LSQ_ERR_IO (-5)      /* host front end: file could not be read */
#define LSQ_ERR_PARSE (-6)   /* host front end: malformed number (reference: lexical_cast error) */
#define LSQ_ERR_FORMAT (-7)  /* host front end: unknown file format / read type token */

#define LSQ_MAX_FORMS_PER_EVENT 64u /* compatibility classes are 64-bit form masks */
#define LSQ_MAX_METHODS 16u

/* read_type tokens of the reference (common/read.h:81-93) that count/solve accept */
#define LSQ_READ_MEDIUM 1u /* "MEDIUM_READ": AccessibleReadStarts (accessible_read_starts.h:48-89) */
#define LSQ_READ_SHORT 2u  /* "SHORT_READ": AccessibleShortReadStarts (:221-274), min_partial 0 */

#define LSQ_READ_MIXED 0x80000000u /* read_cs flag: blocks differ in chromosome/strand */

/* lsq_packed_reads.flags */
#define LSQ_READS_STREAM 1u /* host buffers only: lsq_set_reads returns while the host->device copy is still in
                               flight and the load filter of the next lsq_count consumes the blocks as they land
                               (PCIe transfer and filter overlap).  The caller keeps the buffers valid and
                               unmodified until that lsq_count has returned. */

typedef struct lsq_ctx lsq_ctx;

/* ---------------------------------------------------------------------------------------------
 * Event tables: the per-run immutable model the reference builds at count/count.cpp:231-258
 * (ExonSet::insert, Isoforms::build, covered_regions, gname2il).  Events are given in OUTPUT
 * order (the reference's lexicographic gene-rank order); results come back in the same order.
 * An "event" is the reference's "gene", a "form" its "isoform", a "segment" its atomic "exon".
 * ------------------------------------------------------------------------------------------- */
typedef struct lsq_event_tables {
    uint32_t n_events;   /* E */
    uint32_t n_forms;    /* sum of K_e */
    uint32_t n_segments; /* sum of S_e */
    uint32_t n_chroms;   /* size of the annotation chromosome dictionary */
    uint32_t n_strands;  /* size of the annotation strand dictionary */
    uint32_t reserved0;

    /* per event [E] */
    const uint32_t *ev_chrom;  /* chromosome id of the event's FIRST form (splicing_graph.h:239-242) */
    const uint32_t *ev_strand; /* strand id of the first form */
    const int32_t *ev_gs;      /* smallest exon start of the event (count.cpp:374) */
    const int32_t *ev_ge;      /* largest exon end (count.cpp:375); candidate reads have gs <= start <= ge */
    const uint32_t *ev_seg_off;  /* [E+1] CSR into seg_* */
    const uint32_t *ev_form_off; /* [E+1] CSR into form_* */
    const uint32_t *ev_name_off; /* [E+1] CSR into ev_names: the event name, needed by the
                                    lower_bound tie-break on ("read-N" < gname), count.cpp:64-85,429 */
    const char *ev_names;

    /* per segment [n_segments]: disjoint, ascending within an event (splicing_graph.h:88-169) */
    const int32_t *seg_start;
    const int32_t *seg_end;

    /* per form [n_forms]: ascending segment indices (known_iso_exon_indices, :348-360) */
    const uint32_t *form_seg_off; /* [n_forms+1] CSR into form_seg_idx */
    const uint32_t *form_seg_idx; /* segment index relative to the event's first segment */

    /* merged exon union of the WHOLE selected range per chromosome (covered_regions,
     * count.cpp:231,244): the per-block load filter (count.cpp:319).  When events are a shard
     * of the range, this must still be the union over the full range. */
    const uint32_t *cov_off; /* [n_chroms+1] */
    const int32_t *cov_start;
    const int32_t *cov_end;

    /* dictionaries: ids above index these arrays of NUL-terminated strings */
    const char *const *chrom_names;
    const char *const *strand_names;
} lsq_event_tables;

/* ---------------------------------------------------------------------------------------------
 * Packed MRF_SINGLE reads: what the parser of count/count.cpp:279-339 sees BEFORE the covered-
 * region filter.  Block (start,end) is the file's 1-based inclusive pair already converted to
 * [start-1, end).  line_num is the reference's line counter (first data line = 1, comment and
 * header lines counted; the read's name is "read-<line_num>", count.cpp:286-295).
 * Pointers are host memory unless on_device != 0 (then they are device pointers that stay
 * valid until the read set is replaced; nothing is copied).
 * ------------------------------------------------------------------------------------------- */
typedef struct lsq_packed_reads {
    uint64_t n_reads;
    uint64_t n_blocks;       /* < 2^32 */
    const uint32_t *blk_off; /* [n_reads+1] CSR into blk / blk_cs; may be NULL when blk_cnt is given */
    const int32_t *blk;      /* [2*n_blocks] interleaved (start0, end) */
    const uint32_t *read_cs; /* [n_reads] bits 0-15 chromosome id, 16-23 strand id (file dictionaries),
                                bit 31 LSQ_READ_MIXED: per-block values are in blk_cs */
    const uint32_t *blk_cs;  /* [n_blocks] same encoding per block; may be NULL if no read is MIXED */
    const uint32_t *line_num; /* [n_reads]; NULL = no comment/header lines in the file: line_num[i] = i + 1 */
    uint32_t n_chrom_names;
    uint32_t n_strand_names;
    const char *const *chrom_names;  /* always host memory */
    const char *const *strand_names; /* always host memory */
    uint32_t on_device;
    uint32_t flags;          /* LSQ_READS_* */
    const uint8_t *blk_cnt;  /* [n_reads] blocks per read (< 256); alternative to blk_off that costs a
                                quarter of the host->device bytes (offsets are rebuilt by a device scan) */
    /* Compact wire format (optional; lsqh_reads_compact produces it): the transfer over PCIe, not the
     * kernels, bounds the end-to-end rate, so every array has a narrow alternative. */
    const uint8_t *read_cs8;   /* [n_reads] alternative to read_cs: bits 0-5 chromosome id (< 64), bit 6 strand
                                  id (< 2), bit 7 MIXED (per-block values in blk_cs) */
    const int32_t *blk_start;  /* [n_blocks] with blk_len: alternative to blk -- start0 of every block ... */
    const uint16_t *blk_len;   /* [n_blocks] ... and end - start0 (< 65536) */
    const uint8_t *blk_len8;   /* [n_blocks] instead of blk_len when every block is shorter than 256 (short reads) */
} lsq_packed_reads;

/* Results of the counted path (count.cpp:421-492, solve.cpp:719-793), caller-owned host buffers. */
typedef struct lsq_count_out {
    uint64_t *n_valid;   /* [M*E] supports[m] per event (count.cpp:467,482) */
    uint64_t *bases;     /* [M*E] matched bases of valid reads (solve.cpp:744) */
    uint64_t *iso_count; /* [n_forms] per-form tallies summed over read sets (count.cpp:477) */
} lsq_count_out;

/* Results of the estimator (solve.cpp:796-847; common/read.h:592-660). */
typedef struct lsq_solve_out {
    double *theta;       /* [n_forms] */
    double *rpkm;        /* [n_forms] */
    double *loglik;      /* [E] reads_log_likelihood at the final theta */
    double *ll_per_read; /* [E] the printed column: loglik / sum n_valid, or 0 */
    uint32_t *iters;     /* [E] EM iterations executed (0: closed-form branches) */
    double *se;          /* [n_forms] or NULL: standard errors from the observed Fisher information at
                            theta (per-read term of common/fim.h:322-367, inverse as linalg.h:28-71).
                            Extension: no reference program evaluates fim.h, so there is no oracle.
                            NaN when undefined (no reads, singular information, more than 16 forms). */
} lsq_solve_out;

/* Device timings (CUDA events on the context's stream) and read statistics of the last lsq_count /
 * lsq_solve.  lsq_count and lsq_solve only enqueue work; lsq_get_timings waits for it.
 * scan_ms: the scan kernel -- load filter + candidate lookup + signature tally, the one pass over the reads;
 * expand_ms: the expand kernel -- the walk of every distinct signature; general_ms: the general kernel. */
typedef struct lsq_timings {
    float clear_ms, scan_ms, expand_ms, general_ms, finalize_ms, solve_ms;
    uint32_t launches;     /* kernels launched by the last lsq_count + lsq_solve */
    uint32_t exact_events; /* events whose EM was replayed read by read in the reference's order: their stop test fell
                              inside summation-order noise and the run could not be certified order-free */
    uint64_t reads_in, blocks_in;   /* summed over read sets */
    uint64_t reads_single;          /* reads tallied by a single-block signature */
    uint64_t reads_pair;            /* reads tallied by a two-block signature */
    uint64_t reads_general;         /* reads walked one by one */
    uint64_t hash_full;             /* two-block reads that found no hash slot (walked one by one) */
    uint64_t pairs_valid;           /* valid (read, event) pairs */
    uint64_t reads_triple;          /* reads tallied by a three-block signature (a build option, off by default) */
    float exact_ms;                 /* part of solve_ms: the read-by-read replay of the exact_events (host wall clock) */
} lsq_timings;

const char *lsq_last_error(void);
const char *lsq_version(void);

int lsq_ctx_create(int device, lsq_ctx **out);
int lsq_ctx_destroy(lsq_ctx *ctx);
/* Use an existing CUDA stream (cudaStream_t) for all work of this context; NULL = own stream. */
int lsq_ctx_set_stream(lsq_ctx *ctx, void *cuda_stream);
int lsq_ctx_sync(lsq_ctx *ctx);

/* Replaces count.cpp:231-258 (+ the per-event ARS objects of :393-419 once read sets are known). */
int lsq_upload_events(lsq_ctx *ctx, const lsq_event_tables *tables);

/* Declares read set m (count.cpp:262-344 minus text parsing).  read_type: LSQ_READ_*;
 * read_length: the CLI's expected_read_length.  Host buffers are copied to the device. */
int lsq_set_reads(lsq_ctx *ctx, uint32_t m, uint32_t read_type, uint64_t read_length,
                  const lsq_packed_reads *reads);

/* Load filter + candidate lookup + compatibility + histogram for every declared read set
 * (count.cpp:319-324, 346-364, 421-482).  Asynchronous: the work is enqueued on the context's stream and
 * results stay on the device (lsq_get_count / lsq_get_solve / lsq_get_timings / lsq_ctx_sync wait for it). */
int lsq_count(lsq_ctx *ctx);
int lsq_get_count(lsq_ctx *ctx, lsq_count_out *out);

/* Compatibility-class histogram of read set m: for event e the classes are
 * [cls_off[e], cls_off[e+1]) with form mask (bit k = compatible with form k) and multiplicity.
 * Call with cls_mask == NULL to obtain only cls_off (sizes). */
int lsq_get_classes(lsq_ctx *ctx, uint32_t m, uint64_t *cls_off /*[E+1]*/, uint64_t *cls_mask,
                    uint64_t *cls_cnt);
/* Inject class tables directly (solve-only workloads): replaces the histogram of read set m.
 * read_type/read_length define the ARS normaliser as in lsq_set_reads. */
int lsq_set_classes(lsq_ctx *ctx, uint32_t m, uint32_t read_type, uint64_t read_length,
                    const uint64_t *cls_off /*[E+1]*/, const uint64_t *cls_mask,
                    const uint64_t *cls_cnt, const uint64_t *bases /*[E]*/);

/* EM + RPKM + log-likelihood (solve.cpp:796-826) + standard errors.  total_read_bases[m] is the CLI value.
 * Theta, RPKM and log-likelihood agree with the reference within 1e-9 relative and the EM runs the reference's number
 * of iterations for every event: where the stop test is decided inside summation-order noise the event is either
 * certified order-free or replayed read by read in the reference's own order (one host round trip; lsq_timings). */
int lsq_solve(lsq_ctx *ctx, const double *total_read_bases, uint32_t n_methods);
int lsq_get_solve(lsq_ctx *ctx, lsq_solve_out *out);

/* Device pointers of the result arrays (valid until the next lsq_upload_events), for callers that
 * move results between GPUs without a host round trip (NCCL gather of gene-range shards). */
typedef struct lsq_device_results {
    const uint64_t *n_valid, *bases, *iso_count; /* [M*E], [M*E], [n_forms] */
    const double *theta, *rpkm;                  /* [n_forms] */
    const double *loglik, *ll_per_read;          /* [E] */
    const uint32_t *iters;                       /* [E] */
} lsq_device_results;
int lsq_get_device_results(lsq_ctx *ctx, lsq_device_results *out);

int lsq_get_timings(lsq_ctx *ctx, lsq_timings *out);
/* FP64 FMA throughput of the context's device in TFLOP/s (a DFMA micro-benchmark, ~10 ms): the measured peak the
 * solve kernels' arithmetic rate is quoted against. */
int lsq_measure_fp64_peak(lsq_ctx *ctx, double *tflops);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: gene-range shards with reads bucketed per shard.  The reference scales out by running one
 * process per gene_begin_idx/gene_end_idx range, each reading the WHOLE read file and keeping what its range
 * covers (count.cpp:113-114, 202-215, 319, 429-432).  Here rank d of n_ranks owns the events
 * [shard_begin[d], shard_begin[d+1]) of the selection (output order = gene-rank order), every rank holds an
 * arbitrary slice of the read set, and lsq_shard_reads sends each read to the ranks whose events it can be a
 * candidate of -- over NCCL (NVLink / NVSwitch).  One process per GPU or one thread per GPU.
 *   full  : context with the event tables of the WHOLE selection (it routes; it never counts)
 *   shard : context with the tables of this rank's events, covered regions of the whole selection
 *           (lsqh_annotation_load with shard_begin/shard_end)
 * ------------------------------------------------------------------------------------------- */
typedef struct lsq_comm lsq_comm;
#define LSQ_COMM_ID_BYTES 128
#define LSQ_MAX_RANKS 16
int lsq_comm_unique_id(void *id /* LSQ_COMM_ID_BYTES, filled on one rank and handed to all */);
int lsq_comm_create(const void *id, int n_ranks, int rank, int device, lsq_comm **out); /* collective (ncclCommInitRank) */
int lsq_comm_attach(void *nccl_comm /* an existing ncclComm_t */, int n_ranks, int rank, int device, lsq_comm **out);
int lsq_comm_destroy(lsq_comm *comm);

/* Which rank owns which event: shard_begin[n_ranks + 1], ascending, shard_begin[0] = 0, last = n_events. */
int lsq_router_build(lsq_ctx *full, uint32_t n_ranks, const uint32_t *shard_begin);

typedef struct lsq_shard_stats {
    uint64_t reads_local, reads_sent, reads_received; /* sent counts a read once per destination */
    uint64_t blocks_received;
    float route_ms, exchange_ms;                      /* CUDA events: bucketing kernel, NCCL exchange */
} lsq_shard_stats;
/* Buckets read set m_src of `full` (this rank's slice, declared with lsq_set_reads; first_line = line number of
 * its first read in the whole file, 1 for the first slice) per destination rank, exchanges the buckets and
 * declares what arrives as read set m_dst of `shard` (device resident, with the reads' global line numbers).
 * Collective: every rank of the communicator must call it.  Limits: what a rank receives must satisfy
 * blocks + 2 * reads < 2^32 (LSQ_ERR_RANGE otherwise), and 128 consecutive reads of a slice may hold at most 2^20
 * blocks in total (always true for the compact layout, whose reads have fewer than 256 blocks). */
int lsq_shard_reads(lsq_ctx *full, uint32_t m_src, lsq_ctx *shard, uint32_t m_dst, lsq_comm *comm, uint64_t first_line,
                    lsq_shard_stats *stats /* may be NULL */);
/* Gathers the per-event / per-form results of all shards on `root` in gene-rank order (one NCCL send/recv group;
 * no collective runs inside the count/solve loop).  count/solve: host buffers for the WHOLE selection on root
 * (NULL members and NULL structs are skipped; ignored on the other ranks).  Collective. */
int lsq_gather(lsq_ctx *full, lsq_ctx *shard, lsq_comm *comm, int root, lsq_count_out *count, lsq_solve_out *solve);

/* ARS length of one form (accessible_read_starts.h:48-89, 184-189, 221-274), host helper. */
uint64_t lsq_ars_length(uint32_t read_type, uint64_t read_length, const uint64_t *seg_len, uint32_t n);

/* ---------------------------------------------------------------------------------------------
 * Host front end
 * ------------------------------------------------------------------------------------------- */
typedef struct lsqh_annotation lsqh_annotation;
typedef struct lsqh_reads lsqh_reads;

/* Loads LH_GENE_TXT forms + UCSC_GENE2ISOFORM map, selects gene ranks [begin,end) in bytewise
 * lexicographic order (count.cpp:141-215) and builds the event model for ranks
 * [shard_begin, shard_end) of that selection while covered regions span the whole selection
 * (shard_begin/shard_end are indices INTO the selection; pass 0, UINT64_MAX for everything). */
int lsqh_annotation_load(const char *forms_path, const char *map_path, uint64_t begin, uint64_t end,
                         uint64_t shard_begin, uint64_t shard_end, lsqh_annotation **out);
/* The same with the reference's format tokens.  forms_format: LH_GENE_TXT (the only one `count` accepts,
 * count.cpp:141), or what `solve` also reads (solve.cpp:158-296): UCSC_GENE_TXT, UCSC_GFF, WORMBASE_GFF2,
 * GENELETS_GFF3; map_format: UCSC_GENE2ISOFORM or WORMBASE_GENE2ISOFORMS (solve.cpp:310-329).  LSQ_ERR_FORMAT for
 * any other token. */
int lsqh_annotation_load_fmt(const char *forms_format, const char *forms_path, const char *map_format,
                             const char *map_path, uint64_t begin, uint64_t end, uint64_t shard_begin,
                             uint64_t shard_end, lsqh_annotation **out);
void lsqh_annotation_free(lsqh_annotation *a);
int lsqh_annotation_tables(const lsqh_annotation *a, lsq_event_tables *out); /* views into a */
const char *lsqh_event_name(const lsqh_annotation *a, uint32_t e);
const char *lsqh_form_name(const lsqh_annotation *a, uint32_t f);
uint64_t lsqh_form_length(const lsqh_annotation *a, uint32_t f);

/* Writes host reads in the wide layout as MRF_SINGLE text (header line + one line per read, the format of
 * count.cpp:279-327): the inverse of lsqh_mrf_parse, for test and benchmark inputs. */
int lsqh_write_mrf(const lsq_packed_reads *reads, const char *path);
/* Parses an MRF_SINGLE file into packed reads (count.cpp:279-339 without the filter). */
int lsqh_mrf_parse(const char *path, lsqh_reads **out);
void lsqh_reads_free(lsqh_reads *r);
int lsqh_reads_view(const lsqh_reads *r, lsq_packed_reads *out);
/* Derives the compact wire format (blk_cnt, read_cs8, blk_start + blk_len) of a parsed read set where the
 * values fit; later views carry both representations and lsq_set_reads copies the narrow one. */
int lsqh_reads_compact(lsqh_reads *r);
/* On-disk packed cache of a parsed read set: the reference parses the same text in `count` and again in `solve`
 * (count.cpp:279-339, solve.cpp:429-486).  source_path (may be NULL) stamps the pack with size and mtime of the text
 * it was made from; lsqh_reads_load then fails with LSQ_ERR_IO when the text has changed (or the pack is missing,
 * truncated or fails its checksum) and the caller parses the text instead.  The drop-in programs use
 * `<reads_path>.lsqpack` only when the environment has LSQ_READ_CACHE=1 (the reference writes no files). */
int lsqh_reads_save(const lsqh_reads *r, const char *pack_path, const char *source_path);
int lsqh_reads_load(const char *pack_path, const char *source_path, lsqh_reads **out);

/* Drop-in bodies of the two programs: same positional arguments (argv[0] = program name),
 * same stdout tables, log on stderr, same exit codes (0, 1; 134 is reported as abort()). */
int lsqh_count_main(int argc, char **argv);
int lsqh_solve_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif /* LESSEQ_B200_H_ */
