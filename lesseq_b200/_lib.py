"""ctypes binding of include/lesseq_b200.h.

The shared library is built in-tree by ``python -m lesseq_b200.build`` (nvcc, sm_100a).  There is
no CPU fallback: if the library is missing this module raises, and every compute entry point
returns LSQ_ERR_CUDA (raised as ``LsqError``) when no CUDA device is usable.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LESSEQ_B200_LIB") or os.path.join(PKG, "lib", "liblesseq_b200.so")  # env: tuning builds only

LSQ_READ_MEDIUM = 1
LSQ_READ_SHORT = 2
LSQ_READ_MIXED = 0x80000000
LSQ_READS_STREAM = 1
LSQ_COMM_ID_BYTES = 128
LSQ_MAX_METHODS = 16
READ_TYPES = {"MEDIUM_READ": LSQ_READ_MEDIUM, "SHORT_READ": LSQ_READ_SHORT}

c_u32p = C.POINTER(C.c_uint32)
c_i32p = C.POINTER(C.c_int32)
c_u64p = C.POINTER(C.c_uint64)
c_f64p = C.POINTER(C.c_double)
c_strp = C.POINTER(C.c_char_p)


class EventTables(C.Structure):
    _fields_ = [
        ("n_events", C.c_uint32), ("n_forms", C.c_uint32), ("n_segments", C.c_uint32),
        ("n_chroms", C.c_uint32), ("n_strands", C.c_uint32), ("reserved0", C.c_uint32),
        ("ev_chrom", c_u32p), ("ev_strand", c_u32p), ("ev_gs", c_i32p), ("ev_ge", c_i32p),
        ("ev_seg_off", c_u32p), ("ev_form_off", c_u32p), ("ev_name_off", c_u32p), ("ev_names", C.c_void_p),
        ("seg_start", c_i32p), ("seg_end", c_i32p),
        ("form_seg_off", c_u32p), ("form_seg_idx", c_u32p),
        ("cov_off", c_u32p), ("cov_start", c_i32p), ("cov_end", c_i32p),
        ("chrom_names", c_strp), ("strand_names", c_strp),
    ]


class PackedReads(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint64), ("n_blocks", C.c_uint64),
        ("blk_off", C.c_void_p), ("blk", C.c_void_p), ("read_cs", C.c_void_p), ("blk_cs", C.c_void_p),
        ("line_num", C.c_void_p),
        ("n_chrom_names", C.c_uint32), ("n_strand_names", C.c_uint32),
        ("chrom_names", c_strp), ("strand_names", c_strp),
        ("on_device", C.c_uint32), ("flags", C.c_uint32),
        ("blk_cnt", C.c_void_p), ("read_cs8", C.c_void_p), ("blk_start", C.c_void_p), ("blk_len", C.c_void_p), ("blk_len8", C.c_void_p),
    ]


class CountOut(C.Structure):
    _fields_ = [("n_valid", C.c_void_p), ("bases", C.c_void_p), ("iso_count", C.c_void_p)]


class SolveOut(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("rpkm", C.c_void_p), ("loglik", C.c_void_p),
                ("ll_per_read", C.c_void_p), ("iters", C.c_void_p), ("se", C.c_void_p)]


class Timings(C.Structure):
    _fields_ = [
        ("clear_ms", C.c_float), ("scan_ms", C.c_float), ("expand_ms", C.c_float), ("general_ms", C.c_float),
        ("finalize_ms", C.c_float), ("solve_ms", C.c_float),
        ("launches", C.c_uint32), ("exact_events", C.c_uint32),
        ("reads_in", C.c_uint64), ("blocks_in", C.c_uint64), ("reads_single", C.c_uint64), ("reads_pair", C.c_uint64),
        ("reads_general", C.c_uint64), ("hash_full", C.c_uint64), ("pairs_valid", C.c_uint64), ("reads_triple", C.c_uint64), ("exact_ms", C.c_float),
    ]


class ShardStats(C.Structure):
    _fields_ = [("reads_local", C.c_uint64), ("reads_sent", C.c_uint64), ("reads_received", C.c_uint64),
                ("blocks_received", C.c_uint64), ("route_ms", C.c_float), ("exchange_ms", C.c_float)]


class DeviceResults(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("n_valid", "bases", "iso_count", "theta", "rpkm", "loglik", "ll_per_read",
                                           "iters")]


class LsqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lesseq_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/lesseq_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "lsq_last_error": (C.c_char_p, []),
    "lsq_version": (C.c_char_p, []),
    "lsq_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "lsq_ctx_destroy": (C.c_int, [C.c_void_p]),
    "lsq_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lsq_ctx_sync": (C.c_int, [C.c_void_p]),
    "lsq_upload_events": (C.c_int, [C.c_void_p, C.POINTER(EventTables)]),
    "lsq_set_reads": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(PackedReads)]),
    "lsq_count": (C.c_int, [C.c_void_p]),
    "lsq_get_count": (C.c_int, [C.c_void_p, C.POINTER(CountOut)]),
    "lsq_get_classes": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lsq_set_classes": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "lsq_solve": (C.c_int, [C.c_void_p, c_f64p, C.c_uint32]),
    "lsq_get_solve": (C.c_int, [C.c_void_p, C.POINTER(SolveOut)]),
    "lsq_get_timings": (C.c_int, [C.c_void_p, C.POINTER(Timings)]),
    "lsq_get_device_results": (C.c_int, [C.c_void_p, C.POINTER(DeviceResults)]),
    "lsq_measure_fp64_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "lsq_comm_unique_id": (C.c_int, [C.c_void_p]),
    "lsq_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "lsq_comm_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "lsq_comm_destroy": (C.c_int, [C.c_void_p]),
    "lsq_router_build": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
    "lsq_shard_reads": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                  C.POINTER(ShardStats)]),
    "lsq_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(CountOut), C.POINTER(SolveOut)]),
    "lsq_ars_length": (C.c_uint64, [C.c_uint32, C.c_uint64, c_u64p, C.c_uint32]),
    "lsqh_annotation_load": (C.c_int, [C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                       C.POINTER(C.c_void_p)]),
    "lsqh_annotation_load_fmt": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint64, C.c_uint64,
                                           C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p)]),
    "lsqh_annotation_free": (None, [C.c_void_p]),
    "lsqh_annotation_tables": (C.c_int, [C.c_void_p, C.POINTER(EventTables)]),
    "lsqh_event_name": (C.c_char_p, [C.c_void_p, C.c_uint32]),
    "lsqh_form_name": (C.c_char_p, [C.c_void_p, C.c_uint32]),
    "lsqh_form_length": (C.c_uint64, [C.c_void_p, C.c_uint32]),
    "lsqh_write_mrf": (C.c_int, [C.POINTER(PackedReads), C.c_char_p]),
    "lsqh_mrf_parse": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "lsqh_reads_free": (None, [C.c_void_p]),
    "lsqh_reads_compact": (C.c_int, [C.c_void_p]),
    "lsqh_reads_save": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "lsqh_reads_load": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "lsqh_reads_view": (C.c_int, [C.c_void_p, C.POINTER(PackedReads)]),
    "lsqh_count_main": (C.c_int, [C.c_int, c_strp]),
    "lsqh_solve_main": (C.c_int, [C.c_int, c_strp]),
}

_lib = None


def load():
    """Load liblesseq_b200.so (once).  Raises ImportError with the build hint if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("lesseq_b200: %s not found -- run `python -m lesseq_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise LsqError(rc, load().lsq_last_error().decode("utf-8", "replace"))
