"""Python mirror of the drop-in boundary (include/lesseq_b200.h).

The reference has no library API -- only two programs (count/count.cpp, solve/solve.cpp).  The
classes below wrap the C ABI one-to-one; ``count_main`` / ``solve_main`` take the reference's
positional argument lists.  numpy holds host buffers, torch (if given CUDA tensors) device ones.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import _lib
from ._lib import LsqError, READ_TYPES, check  # noqa: F401

PKG = os.path.dirname(os.path.abspath(__file__))
UINT64_MAX = 2 ** 64 - 1


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Annotation:
    """Event model built from LH_GENE_TXT + UCSC_GENE2ISOFORM files (count/count.cpp:136-258) or the
    other annotation formats of solve/solve.cpp:158-329 (forms_format / map_format tokens)."""

    def __init__(self, forms_path, map_path, begin=0, end=UINT64_MAX, shard=(0, UINT64_MAX),
                 forms_format="LH_GENE_TXT", map_format="UCSC_GENE2ISOFORM"):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.lsqh_annotation_load_fmt(forms_format.encode(), str(forms_path).encode(), map_format.encode(),
                                                 str(map_path).encode(), begin, end, shard[0], shard[1],
                                                 C.byref(self._h)))
        self.tables = _lib.EventTables()
        check(self._lib.lsqh_annotation_tables(self._h, C.byref(self.tables)))
        self.n_events = self.tables.n_events
        self.n_forms = self.tables.n_forms

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lsqh_annotation_free(self._h)
            self._h = None

    def event_names(self):
        return [self._lib.lsqh_event_name(self._h, e).decode() for e in range(self.n_events)]

    def form_names(self):
        return [self._lib.lsqh_form_name(self._h, f).decode() for f in range(self.n_forms)]

    def form_lengths(self):
        return np.array([self._lib.lsqh_form_length(self._h, f) for f in range(self.n_forms)], dtype=np.uint64)

    def form_offsets(self):
        return np.ctypeslib.as_array(self.tables.ev_form_off, shape=(self.n_events + 1,)).copy()

    def array(self, name, n):
        return np.ctypeslib.as_array(getattr(self.tables, name), shape=(n,)).copy()


class ArrayAnnotation:
    """Event model given directly as arrays (no text front end); same `tables` attribute as
    Annotation.  Used for solve-only workloads (BASELINE config 5) and tests."""

    def __init__(self, ev_chrom, ev_strand, ev_gs, ev_ge, ev_seg_off, ev_form_off, seg_start, seg_end,
                 form_seg_off, form_seg_idx, cov_off, cov_start, cov_end, chrom_names, strand_names, ev_names=None):
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)  # noqa: E731
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
        E = len(ev_gs)
        if ev_names is None:
            ev_names = ["e%d" % i for i in range(E)]
        pool = "".join(ev_names).encode()
        name_off = np.zeros(E + 1, dtype=np.uint32)
        name_off[1:] = np.cumsum([len(n.encode()) for n in ev_names])
        self._a = dict(ev_chrom=u32(ev_chrom), ev_strand=u32(ev_strand), ev_gs=i32(ev_gs), ev_ge=i32(ev_ge),
                       ev_seg_off=u32(ev_seg_off), ev_form_off=u32(ev_form_off), ev_name_off=name_off,
                       seg_start=i32(seg_start), seg_end=i32(seg_end), form_seg_off=u32(form_seg_off),
                       form_seg_idx=u32(form_seg_idx), cov_off=u32(cov_off), cov_start=i32(cov_start),
                       cov_end=i32(cov_end))
        self._pool = C.create_string_buffer(pool, max(len(pool), 1))
        self._cn = (C.c_char_p * max(len(chrom_names), 1))(*[s.encode() for s in chrom_names])
        self._sn = (C.c_char_p * max(len(strand_names), 1))(*[s.encode() for s in strand_names])
        t = _lib.EventTables()
        t.n_events, t.n_forms, t.n_segments = E, len(form_seg_off) - 1, len(seg_start)
        t.n_chroms, t.n_strands = len(chrom_names), len(strand_names)
        for k, a in self._a.items():
            setattr(t, k, a.ctypes.data_as(dict(_lib.EventTables._fields_)[k]))
        t.ev_names = C.cast(self._pool, C.c_void_p)
        t.chrom_names = C.cast(self._cn, _lib.c_strp)
        t.strand_names = C.cast(self._sn, _lib.c_strp)
        self.tables = t
        self.n_events, self.n_forms = E, t.n_forms
        self.names = ev_names

    def form_offsets(self):
        return self._a["ev_form_off"].copy()


class Reads:
    """Packed MRF_SINGLE reads (lsq_packed_reads).  Build with ``from_mrf``, ``from_numpy`` or
    ``from_torch`` (CUDA tensors are passed by device pointer without a copy)."""

    def __init__(self):
        self.view = _lib.PackedReads()
        self._keep = []
        self._h = None
        self._lib = _lib.load()

    @classmethod
    def from_mrf(cls, path, compact=False):
        """Parse an MRF_SINGLE file; compact=True also derives the compact wire format
        (lsqh_reads_compact), which lsq_set_reads then prefers."""
        self = cls()
        self._h = C.c_void_p()
        check(self._lib.lsqh_mrf_parse(str(path).encode(), C.byref(self._h)))
        if compact:
            check(self._lib.lsqh_reads_compact(self._h))
        check(self._lib.lsqh_reads_view(self._h, C.byref(self.view)))
        return self

    @classmethod
    def from_pack(cls, pack_path, source_path=None, compact=False):
        """Load a packed read cache (lsqh_reads_load); raises LsqError when it is missing, stale
        with respect to source_path, or corrupt."""
        self = cls()
        self._h = C.c_void_p()
        check(self._lib.lsqh_reads_load(str(pack_path).encode(), str(source_path).encode() if source_path else None,
                                        C.byref(self._h)))
        if compact:
            check(self._lib.lsqh_reads_compact(self._h))
        check(self._lib.lsqh_reads_view(self._h, C.byref(self.view)))
        return self

    def save_pack(self, pack_path, source_path=None):
        """Write the packed read cache of a parsed file (lsqh_reads_save)."""
        assert self._h, "only reads parsed by from_mrf / loaded by from_pack own a host read set"
        check(self._lib.lsqh_reads_save(self._h, str(pack_path).encode(),
                                        str(source_path).encode() if source_path else None))

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lsqh_reads_free(self._h)
            self._h = None

    def _names(self, chrom_names, strand_names):
        cn = (C.c_char_p * len(chrom_names))(*[s.encode() for s in chrom_names])
        sn = (C.c_char_p * len(strand_names))(*[s.encode() for s in strand_names])
        self._keep += [cn, sn]
        self.view.n_chrom_names = len(chrom_names)
        self.view.n_strand_names = len(strand_names)
        self.view.chrom_names = C.cast(cn, _lib.c_strp)
        self.view.strand_names = C.cast(sn, _lib.c_strp)

    @classmethod
    def from_numpy(cls, blk_off, blk, read_cs, line_num, chrom_names, strand_names, blk_cs=None):
        self = cls()
        blk_off = np.ascontiguousarray(blk_off, dtype=np.uint32)
        blk = np.ascontiguousarray(blk, dtype=np.int32).reshape(-1)
        read_cs = np.ascontiguousarray(read_cs, dtype=np.uint32)
        line_num = np.ascontiguousarray(line_num, dtype=np.uint32)
        self._keep += [blk_off, blk, read_cs, line_num]
        v = self.view
        v.n_reads = len(read_cs)
        v.n_blocks = len(blk) // 2
        v.blk_off, v.blk, v.read_cs, v.line_num = _ptr(blk_off), _ptr(blk), _ptr(read_cs), _ptr(line_num)
        if blk_cs is not None:
            blk_cs = np.ascontiguousarray(blk_cs, dtype=np.uint32)
            self._keep.append(blk_cs)
            v.blk_cs = _ptr(blk_cs)
        v.on_device = 0
        self._names(chrom_names, strand_names)
        return self

    @classmethod
    def from_torch(cls, blk_off, blk, read_cs, line_num, chrom_names, strand_names, blk_cs=None, blk_cnt=None,
                   stream=False, read_cs8=None, blk_start=None, blk_len=None, blk_len8=None):
        """Tensors must be contiguous int32 (bit patterns of the uint32 fields; blk_cnt uint8); CUDA
        tensors are borrowed (on_device=1), CPU tensors (pinned or not) are host buffers.  Compact
        wire format: pass blk_off=None with blk_cnt (blocks per read) and/or line_num=None (implicit
        1..n); read_cs8 (uint8) instead of read_cs; blk_start (int32) + blk_len (int16 bit patterns of
        uint16) or blk_len8 (uint8) instead of blk.  stream=True (host tensors): LSQ_READS_STREAM, the copy overlaps the
        next count()."""
        self = cls()
        ts = [t for t in (blk_off, blk, read_cs, line_num, blk_cs, blk_start) if t is not None]
        dev = (blk if blk is not None else blk_start).is_cuda
        for t in ts:
            assert t.is_contiguous() and t.element_size() == 4 and t.is_cuda == dev
        if blk_cnt is not None:
            assert blk_cnt.is_contiguous() and blk_cnt.element_size() == 1 and blk_cnt.is_cuda == dev
            ts.append(blk_cnt)
            self.view.blk_cnt = blk_cnt.data_ptr()
        assert blk_off is not None or blk_cnt is not None
        if read_cs8 is not None:
            assert read_cs8.is_contiguous() and read_cs8.element_size() == 1 and read_cs8.is_cuda == dev
            ts.append(read_cs8)
            self.view.read_cs8 = read_cs8.data_ptr()
        if blk_len is not None:
            assert blk_start is not None and blk_len.is_contiguous() and blk_len.element_size() == 2
            assert blk_len.is_cuda == dev and blk_len.numel() == blk_start.numel()
            ts.append(blk_len)
            self.view.blk_start, self.view.blk_len = blk_start.data_ptr(), blk_len.data_ptr()
        if blk_len8 is not None:  # one byte per block length (every block shorter than 256)
            assert blk_start is not None and blk_len8.is_contiguous() and blk_len8.element_size() == 1
            assert blk_len8.is_cuda == dev and blk_len8.numel() == blk_start.numel()
            ts.append(blk_len8)
            self.view.blk_start, self.view.blk_len8 = blk_start.data_ptr(), blk_len8.data_ptr()
        self._keep += ts
        v = self.view
        v.n_reads = (read_cs if read_cs is not None else read_cs8).numel()
        v.n_blocks = blk.numel() // 2 if blk is not None else blk_start.numel()
        v.blk = blk.data_ptr() if blk is not None else None
        v.read_cs = read_cs.data_ptr() if read_cs is not None else None
        v.blk_off = blk_off.data_ptr() if blk_off is not None else None
        v.line_num = line_num.data_ptr() if line_num is not None else None
        if blk_cs is not None:
            v.blk_cs = blk_cs.data_ptr()
        v.on_device = 1 if dev else 0
        v.flags = _lib.LSQ_READS_STREAM if (stream and not dev) else 0
        self._names(chrom_names, strand_names)
        return self

    def write_mrf(self, path):
        """Write host reads in the wide layout as MRF_SINGLE text (lsqh_write_mrf)."""
        check(self._lib.lsqh_write_mrf(C.byref(self.view), str(path).encode()))

    @property
    def n_reads(self):
        return int(self.view.n_reads)

    @property
    def n_blocks(self):
        return int(self.view.n_blocks)

    def host_bytes(self):
        """Bytes lsq_set_reads copies host->device for this read set."""
        n, nb = self.n_reads, self.n_blocks
        v = self.view
        return ((n if v.blk_cnt else (n + 1) * 4) + (nb * 5 if v.blk_len8 else nb * 6 if v.blk_len else nb * 8) + (n if v.read_cs8 else n * 4) +
                (n * 4 if v.line_num else 0))


class Context:
    """One device context (lsq_ctx).  Call order: upload_events -> set_reads* -> count -> solve."""

    def __init__(self, device=0, stream=None):
        self._lib = _lib.load()
        self._c = C.c_void_p()
        check(self._lib.lsq_ctx_create(device, C.byref(self._c)))
        self.E = self.F = self.M = 0
        if stream is not None:
            check(self._lib.lsq_ctx_set_stream(self._c, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "_c", None):
            self._lib.lsq_ctx_destroy(self._c)
            self._c = None

    __del__ = close

    def upload_events(self, ann):
        t = ann.tables if hasattr(ann, "tables") else ann
        check(self._lib.lsq_upload_events(self._c, C.byref(t)))
        self.E, self.F, self.M = t.n_events, t.n_forms, 0
        self._ann = ann

    def set_reads(self, m, read_type, read_length, reads):
        rt = READ_TYPES[read_type] if isinstance(read_type, str) else read_type
        check(self._lib.lsq_set_reads(self._c, m, rt, read_length, C.byref(reads.view)))
        self.M = max(self.M, m + 1)
        self._reads = getattr(self, "_reads", {})
        self._reads[m] = reads

    def set_classes(self, m, read_type, read_length, cls_off, cls_mask, cls_cnt, bases):
        rt = READ_TYPES[read_type] if isinstance(read_type, str) else read_type
        a = [np.ascontiguousarray(x, dtype=np.uint64) for x in (cls_off, cls_mask, cls_cnt, bases)]
        check(self._lib.lsq_set_classes(self._c, m, rt, read_length, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3])))
        self.M = max(self.M, m + 1)

    def count(self):
        check(self._lib.lsq_count(self._c))

    def get_count(self):
        nv = np.zeros((self.M, self.E), dtype=np.uint64)
        bs = np.zeros((self.M, self.E), dtype=np.uint64)
        iso = np.zeros(self.F, dtype=np.uint64)
        out = _lib.CountOut(_ptr(nv), _ptr(bs), _ptr(iso))
        check(self._lib.lsq_get_count(self._c, C.byref(out)))
        return {"n_valid": nv, "bases": bs, "iso_count": iso}

    def get_classes(self, m):
        off = np.zeros(self.E + 1, dtype=np.uint64)
        check(self._lib.lsq_get_classes(self._c, m, _ptr(off), None, None))
        n = int(off[-1])
        mask = np.zeros(max(n, 1), dtype=np.uint64)
        cnt = np.zeros(max(n, 1), dtype=np.uint64)
        check(self._lib.lsq_get_classes(self._c, m, _ptr(off), _ptr(mask), _ptr(cnt)))
        return off, mask[:n], cnt[:n]

    def solve(self, total_read_bases):
        trb = np.ascontiguousarray(total_read_bases, dtype=np.float64)
        check(self._lib.lsq_solve(self._c, trb.ctypes.data_as(_lib.c_f64p), len(trb)))

    def get_solve(self):
        th = np.zeros(self.F)
        rp = np.zeros(self.F)
        ll = np.zeros(self.E)
        lp = np.zeros(self.E)
        it = np.zeros(self.E, dtype=np.uint32)
        se = np.zeros(self.F)
        out = _lib.SolveOut(_ptr(th), _ptr(rp), _ptr(ll), _ptr(lp), _ptr(it), _ptr(se))
        check(self._lib.lsq_get_solve(self._c, C.byref(out)))
        return {"theta": th, "rpkm": rp, "loglik": ll, "ll_per_read": lp, "iters": it, "se": se}

    def device_results(self):
        """Zero-copy torch views (CUDA) of the result arrays; uint64 arrays are viewed as int64."""
        import torch
        r = _lib.DeviceResults()
        check(self._lib.lsq_get_device_results(self._c, C.byref(r)))
        E, F, M = self.E, self.F, self.M

        class _Cai:  # __cuda_array_interface__ carrier
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False),
                                                 "version": 2}

        def view(ptr, n, typestr):
            return torch.as_tensor(_Cai(ptr, n, typestr), device="cuda") if n else torch.empty(0, device="cuda")

        return {"n_valid": view(r.n_valid, M * E, "<i8"), "bases": view(r.bases, M * E, "<i8"),
                "iso_count": view(r.iso_count, F, "<i8"), "theta": view(r.theta, F, "<f8"),
                "rpkm": view(r.rpkm, F, "<f8"), "loglik": view(r.loglik, E, "<f8"),
                "ll_per_read": view(r.ll_per_read, E, "<f8"), "iters": view(r.iters, E, "<i4")}

    def sync(self):
        check(self._lib.lsq_ctx_sync(self._c))

    def fp64_peak_tflops(self):
        """Measured FP64 FMA throughput of the device (lsq_measure_fp64_peak)."""
        v = C.c_double()
        check(self._lib.lsq_measure_fp64_peak(self._c, C.byref(v)))
        return float(v.value)

    def timings(self):
        t = _lib.Timings()
        check(self._lib.lsq_get_timings(self._c, C.byref(t)))
        return {k: getattr(t, k) for k, _ in _lib.Timings._fields_}


class Comm:
    """NCCL communicator of the gene-range shards (lsq_comm): one rank per GPU.  `uid` comes from
    Comm.unique_id() on one rank and is handed to all (e.g. torch.distributed.broadcast_object_list)."""

    def __init__(self, uid, n_ranks, rank, device):
        self._lib = _lib.load()
        self._c = C.c_void_p()
        self.n_ranks, self.rank = n_ranks, rank
        buf = C.create_string_buffer(bytes(uid), 128)
        check(self._lib.lsq_comm_create(buf, n_ranks, rank, device, C.byref(self._c)))

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        check(_lib.load().lsq_comm_unique_id(buf))
        return bytes(buf.raw)

    def close(self):
        if getattr(self, "_c", None):
            self._lib.lsq_comm_destroy(self._c)
            self._c = None

    __del__ = close


def router_build(full_ctx, shard_begin):
    """Declare which rank owns which events of the whole selection held by `full_ctx`
    (shard_begin: n_ranks + 1 ascending event indices, gene-rank order)."""
    sb = np.ascontiguousarray(shard_begin, dtype=np.uint32)
    check(full_ctx._lib.lsq_router_build(full_ctx._c, len(sb) - 1, _ptr(sb)))
    full_ctx._shard_begin = sb


def shard_reads(full_ctx, m_src, shard_ctx, m_dst, comm, first_line=1):
    """Bucket read set m_src of full_ctx (this rank's slice) per owning rank, exchange over NCCL and declare what
    arrives as read set m_dst of shard_ctx.  Collective.  Returns the lsq_shard_stats as a dict."""
    st = _lib.ShardStats()
    check(full_ctx._lib.lsq_shard_reads(full_ctx._c, m_src, shard_ctx._c, m_dst, comm._c, first_line, C.byref(st)))
    shard_ctx.M = max(shard_ctx.M, m_dst + 1)
    return {k: getattr(st, k) for k, _ in _lib.ShardStats._fields_}


def gather(full_ctx, shard_ctx, comm, root=0, solved=True):
    """Gather all shards' result rows on `root` in gene-rank order (lsq_gather).  Collective.  Returns
    (count, solve) dicts for the whole selection on root, (None, None) elsewhere."""
    M, E, F = shard_ctx.M, full_ctx.E, full_ctx.F
    if comm.rank != root:
        check(full_ctx._lib.lsq_gather(full_ctx._c, shard_ctx._c, comm._c, root, None, None))
        return None, None
    nv = np.zeros((M, E), dtype=np.uint64)
    bs = np.zeros((M, E), dtype=np.uint64)
    iso = np.zeros(F, dtype=np.uint64)
    th, rp, se = np.zeros(F), np.zeros(F), np.zeros(F)
    ll, lp = np.zeros(E), np.zeros(E)
    it = np.zeros(E, dtype=np.uint32)
    co = _lib.CountOut(_ptr(nv), _ptr(bs), _ptr(iso))
    so = _lib.SolveOut(_ptr(th), _ptr(rp), _ptr(ll), _ptr(lp), _ptr(it), _ptr(se))
    check(full_ctx._lib.lsq_gather(full_ctx._c, shard_ctx._c, comm._c, root, C.byref(co), C.byref(so) if solved else None))
    count = {"n_valid": nv, "bases": bs, "iso_count": iso}
    solve = {"theta": th, "rpkm": rp, "loglik": ll, "ll_per_read": lp, "iters": it, "se": se} if solved else None
    return count, solve


def ars_length(read_type, read_length, seg_len):
    rt = READ_TYPES[read_type] if isinstance(read_type, str) else read_type
    a = np.ascontiguousarray(seg_len, dtype=np.uint64)
    return int(_lib.load().lsq_ars_length(rt, read_length, a.ctypes.data_as(_lib.c_u64p), len(a)))


def run_program(prog, args, precision=None, timeout=None, env_extra=None):
    """Run the drop-in `count` / `solve` executable (lesseq_b200/bin) with the reference's
    positional arguments; returns (returncode, stdout, stderr)."""
    exe = os.path.join(PKG, "bin", prog)
    if not os.path.exists(exe):
        raise ImportError("lesseq_b200: %s not built -- run `python -m lesseq_b200.build`" % exe)
    env = dict(os.environ)
    if precision is not None:
        env["LSQ_PRECISION"] = str(precision)
    env.update(env_extra or {})
    p = subprocess.run([exe] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=timeout)
    return p.returncode, p.stdout, p.stderr
