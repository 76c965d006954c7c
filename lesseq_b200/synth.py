"""Synthetic annotations and reads of the BASELINE.json shapes.

Event shapes follow the 8 types emitted by the reference's bin/Events.r:45-160 (all 2-form):
SE, RI, A5SS, A3SS, MXE, AFE, ALE, T3.  Annotations are produced as the reference's text formats
(LH_GENE_TXT forms + UCSC_GENE2ISOFORM map) so that they go through the real front end; reads are
produced directly in packed form with torch (on the GPU for the large configurations) and can be
written as MRF_SINGLE text for the small ones that also run through the reference binaries.
"""
import numpy as np
import torch

EVENT_TYPES = ("SE", "RI", "A5SS", "A3SS", "MXE", "AFE", "ALE", "T3")


class SynthAnnotation:
    def __init__(self):
        self.chrom_names = []
        self.chrom_len = []
        self.forms = []        # (name, chrom_idx, strand, [(s,e),...])
        self.map_rows = []     # (event_id, form_name)
        self.ev_gene = []      # gene index per event (2 consecutive forms per event)
        self.n_genes = 0

    # -- text formats ---------------------------------------------------------------------------
    def forms_text(self):
        out = []
        for name, c, st, ex in self.forms:
            out.append("%s\t%s\t%s\t%d\t%d\t%d\t%s\t%s\n" % (
                name, self.chrom_names[c], st, ex[0][0], ex[-1][1], len(ex),
                ",".join(str(s) for s, _ in ex), ",".join(str(e) for _, e in ex)))
        return "".join(out)

    def map_text(self):
        return "".join("%s\t%s\n" % r for r in self.map_rows)

    def write(self, forms_path, map_path):
        with open(forms_path, "w") as f:
            f.write(self.forms_text())
        with open(map_path, "w") as f:
            f.write(self.map_text())

    # -- tensors for read generation ------------------------------------------------------------
    def form_tensors(self, device):
        n = len(self.forms)
        ex = np.zeros((n, 3, 2), dtype=np.int64)
        ch = np.zeros(n, dtype=np.int64)
        for i, (_, c, _, e) in enumerate(self.forms):
            ch[i] = c
            for j, (s, t) in enumerate(e):
                ex[i, j] = (s, t)
            for j in range(len(e), 3):
                ex[i, j] = (e[-1][1], e[-1][1])  # zero-length padding
        return (torch.from_numpy(ex).to(device), torch.from_numpy(ch).to(device),
                torch.tensor(self.ev_gene, dtype=torch.int64, device=device))


def _event_forms(rng, typ, exons, strand, i):
    """Two forms of one event of type `typ` anchored at exon i of the gene (Events.r shapes)."""
    U = exons[i]
    if typ == "SE":
        X, D = exons[i + 1], exons[i + 2]
        return [U, X, D], [U, D]
    if typ == "RI":
        D = exons[i + 1]
        return [U, (U[1], D[0]), D], [U, D]
    if typ in ("A5SS", "A3SS"):
        D = exons[i + 1]
        intron = D[0] - U[1]
        ext = int(rng.integers(10, max(11, min(100, intron - 20))))
        near_up = (typ == "A5SS") == (strand == "+")
        X = (U[1], U[1] + ext) if near_up else (D[0] - ext, D[0])
        return [U, X, D], [U, D]
    if typ == "MXE":
        B, Cx, D = exons[i + 1], exons[i + 2], exons[i + 3]
        return [U, B, D], [U, Cx, D]
    if typ == "AFE":
        F2, D = exons[i + 1], exons[i + 2]
        return [U, D], [F2, D]
    if typ == "ALE":
        L1, L2 = exons[i + 1], exons[i + 2]
        return [U, L1], [U, L2]
    if typ == "T3":
        cut = int(rng.integers(max(1, (U[1] - U[0]) // 3), U[1] - U[0]))
        if strand == "+":
            return [U], [(U[0], U[0] + cut)]
        return [U], [(U[1] - cut, U[1])]
    raise ValueError(typ)


_NEED = {"SE": 3, "RI": 2, "A5SS": 2, "A3SS": 2, "MXE": 4, "AFE": 3, "ALE": 3, "T3": 1}


def make_annotation(n_genes, seed, n_chroms=1, types=("SE",), exons_per_gene=(3, 3), events_per_gene=(1, 1),
                    exon_len=(80, 300), intron_len=(200, 2000), gene_gap=(3000, 8000), numeric_ids=False):
    """C1: n_genes=500, types=("SE",).  C2/C3: n_genes=20000, n_chroms=24, all 8 types,
    exons_per_gene=(6,12), events_per_gene=(2,5).  C4: exon_len=(40,120)."""
    rng = np.random.default_rng(seed)
    A = SynthAnnotation()
    A.chrom_names = ["chr%d" % (c + 1) for c in range(n_chroms)]
    cursor = [int(rng.integers(1000, 5000)) for _ in range(n_chroms)]
    A.n_genes = n_genes
    eid = 0
    for g in range(n_genes):
        c = g % n_chroms
        strand = "+" if rng.random() < 0.5 else "-"
        nex = int(rng.integers(exons_per_gene[0], exons_per_gene[1] + 1))
        pos = cursor[c] + int(rng.integers(gene_gap[0], gene_gap[1] + 1))
        exons = []
        for _ in range(nex):
            ln = int(rng.integers(exon_len[0], exon_len[1] + 1))
            exons.append((pos, pos + ln))
            pos += ln + int(rng.integers(intron_len[0], intron_len[1] + 1))
        cursor[c] = exons[-1][1]
        nev = int(rng.integers(events_per_gene[0], events_per_gene[1] + 1))
        for _ in range(nev):
            typ = types[int(rng.integers(0, len(types)))]
            if _NEED[typ] > nex:
                typ = "T3"
            i = int(rng.integers(0, nex - _NEED[typ] + 1))
            f1, f2 = _event_forms(rng, typ, exons, strand, i)
            eid += 1
            ev_id = str(eid) if numeric_ids else "%s_%d" % (typ, eid)
            tag = typ if typ in ("AFE", "ALE", "T3") else "2"
            for k, f in enumerate((f1, f2)):
                fname = "G%d.%d|%s|%d" % (g, eid, tag, k + 1)
                A.forms.append((fname, c, strand, list(f)))
                A.map_rows.append((ev_id, fname))
            A.ev_gene.append(g)
    A.chrom_len = [x + 10000 for x in cursor]
    return A


def make_reads(A, n, read_len, seed, device="cpu", bg_frac=0.10, overhang_frac=0.0, sigma=1.5, chunk=1 << 24,
               events=None):
    """n reads of length `read_len`: (1-bg_frac) drawn from event forms (log-normal gene abundance,
    Beta(0.5,0.5) inclusion ratio, uniform start on the spliced form, blocks cut at exon ends with
    abutting pieces fused), bg_frac uniform single-block background reads.  `events` restricts the
    expressed events to a subset (a gene-range shard's read bucket).  Returns a dict of int32 torch
    tensors (uint32 bit patterns) in lsq_packed_reads layout, in random read order."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    ex, fchrom, ev_gene = A.form_tensors(dev)
    n_events = ev_gene.numel()
    cpu_gen = torch.Generator().manual_seed(seed + 1)
    abundance = torch.exp(sigma * torch.randn(A.n_genes, generator=cpu_gen, dtype=torch.float64)).to(dev)
    psi = torch.from_numpy(np.random.default_rng(seed + 2).beta(0.5, 0.5, ev_gene.numel())).to(dev)
    subset = None
    if events is not None:
        subset = torch.as_tensor(np.asarray(events), dtype=torch.int64, device=dev)
        cdf = torch.cumsum(abundance[ev_gene[subset]], 0)
        n_events = subset.numel()
    else:
        cdf = torch.cumsum(abundance[ev_gene], 0)
    chrom_len = torch.tensor(A.chrom_len, dtype=torch.float64, device=dev)
    chrom_cdf = torch.cumsum(chrom_len, 0)
    R = int(read_len)
    parts = {"nblk": [], "blk": [], "cs": []}
    done = 0
    while done < n:
        m = min(chunk, n - done)
        is_bg = torch.rand(m, generator=gen, device=dev) < bg_frac
        u = torch.rand(m, generator=gen, device=dev, dtype=torch.float64) * cdf[-1]
        ev = torch.searchsorted(cdf, u).clamp(max=n_events - 1)
        if subset is not None:
            ev = subset[ev]
        pick2 = torch.rand(m, generator=gen, device=dev, dtype=torch.float64) > psi[ev]
        form = ev * 2 + pick2.to(torch.int64)
        e = ex[form]                                   # [m,3,2]
        lens = e[:, :, 1] - e[:, :, 0]
        cum = torch.cumsum(lens, 1)
        T = cum[:, 2]
        max_start = (T - R).clamp(min=0)
        start = (torch.rand(m, generator=gen, device=dev, dtype=torch.float64) * (max_start + 1).to(torch.float64)).to(torch.int64)
        start = torch.minimum(start, max_start)
        end = torch.minimum(start + R, T)
        prev = cum - lens
        lo = torch.maximum(start[:, None], prev)
        hi = torch.minimum(end[:, None], cum)
        valid = hi > lo
        gs = e[:, :, 0] + (lo - prev)
        ge = e[:, :, 0] + (hi - prev)
        for j in (0, 1):                               # fuse pieces of abutting exons
            fuse = valid[:, j] & valid[:, j + 1] & (ge[:, j] == gs[:, j + 1])
            gs[:, j + 1] = torch.where(fuse, gs[:, j], gs[:, j + 1])
            valid[:, j] = valid[:, j] & ~fuse
        if overhang_frac > 0:                          # +-1 bp overhang at the first junction
            multi = valid.sum(1) > 1
            hit = multi & (torch.rand(m, generator=gen, device=dev) < overhang_frac)
            first = torch.argmax(valid.to(torch.int8), 1)
            delta = torch.where(torch.rand(m, generator=gen, device=dev) < 0.5, 1, -1)
            idx = torch.arange(m, device=dev)
            ge[idx, first] = torch.where(hit, ge[idx, first] + delta, ge[idx, first])
        chrom = fchrom[form]
        # background reads replace the event reads where is_bg
        bu = torch.rand(m, generator=gen, device=dev, dtype=torch.float64) * chrom_cdf[-1]
        bc = torch.searchsorted(chrom_cdf, bu).clamp(max=len(A.chrom_len) - 1)
        bstart = (torch.rand(m, generator=gen, device=dev, dtype=torch.float64) * (chrom_len[bc] - R - 1)).to(torch.int64)
        gs[:, 0] = torch.where(is_bg, bstart, gs[:, 0])
        ge[:, 0] = torch.where(is_bg, bstart + R, ge[:, 0])
        valid[:, 0] = valid[:, 0] | is_bg
        valid[:, 1] = valid[:, 1] & ~is_bg
        valid[:, 2] = valid[:, 2] & ~is_bg
        chrom = torch.where(is_bg, bc, chrom)
        strand = (torch.rand(m, generator=gen, device=dev) < 0.5).to(torch.int64)
        blk = torch.stack((gs, ge), 2)[valid].to(torch.int32)   # CSR order (row-major over valid)
        parts["nblk"].append(valid.sum(1).to(torch.int32))
        parts["blk"].append(blk)
        parts["cs"].append((chrom | (strand << 16)).to(torch.int32))
        done += m
    nblk = torch.cat(parts["nblk"])
    blk_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    blk_off[1:] = torch.cumsum(nblk, 0)
    out = {
        "blk_off": blk_off.to(torch.int32).contiguous(),
        "blk": torch.cat(parts["blk"]).contiguous().view(-1),
        "read_cs": torch.cat(parts["cs"]).contiguous(),
        "line_num": torch.arange(1, n + 1, dtype=torch.int64, device=dev).to(torch.int32).contiguous(),
        "chrom_names": list(A.chrom_names),
        "strand_names": ["-", "+"],
        "n_reads": n,
        "n_blocks": int(blk_off[-1].item()),
    }
    out["total_read_bases"] = float((out["blk"].view(-1, 2)[:, 1] - out["blk"].view(-1, 2)[:, 0]).sum().item())
    return out


def reads_to_mrf(reads, path):
    """Write packed reads as MRF_SINGLE text (header line + one line per read)."""
    off = reads["blk_off"].cpu().numpy().astype(np.int64)
    blk = reads["blk"].cpu().numpy().reshape(-1, 2)
    cs = reads["read_cs"].cpu().numpy().astype(np.int64)
    cn, sn = reads["chrom_names"], reads["strand_names"]
    with open(path, "w") as f:
        f.write("AlignmentBlocks\n")
        lines = []
        for r in range(len(cs)):
            c, st = cn[cs[r] & 0xFFFF], sn[(cs[r] >> 16) & 0xFF]
            q = 1
            parts = []
            for j in range(off[r], off[r + 1]):
                s, e = int(blk[j, 0]) + 1, int(blk[j, 1])
                parts.append("%s:%s:%d:%d:%d:%d" % (c, st, s, e, q, q + e - s))
                q += e - s + 1
            lines.append(",".join(parts))
            if len(lines) >= 100000:
                f.write("\n".join(lines) + "\n")
                lines = []
        if lines:
            f.write("\n".join(lines) + "\n")


CONFIGS = {
    # name: (annotation kwargs, read length, default reads)
    "C1": (dict(n_genes=500, seed=1, n_chroms=1, types=("SE",), exons_per_gene=(3, 3), events_per_gene=(1, 1),
                numeric_ids=True), 50, 1_000_000),
    "C2": (dict(n_genes=20000, seed=2, n_chroms=24, types=EVENT_TYPES, exons_per_gene=(6, 12),
                events_per_gene=(2, 5)), 75, 100_000_000),
    "C4": (dict(n_genes=20000, seed=4, n_chroms=24, types=EVENT_TYPES, exons_per_gene=(6, 12),
                events_per_gene=(2, 5), exon_len=(40, 120)), 100, 100_000_000),
}


def make_config(name, n_genes=None):
    kw, R, n = CONFIGS[name]
    kw = dict(kw)
    if n_genes is not None:
        kw["n_genes"] = n_genes
    return make_annotation(**kw), R, n


def make_solve_stress(n_events, seed, k_range=(2, 16)):
    """C5: events with K~U{2..16} forms over S~U{K..2K} segments (each form a random contiguous-run
    compatible subset) and C~U{K..3K} compatibility classes with log-normal multiplicities.
    Returns numpy arrays for tables_from_numpy + class tables for Context.set_classes."""
    rng = np.random.default_rng(seed)
    K = rng.integers(k_range[0], k_range[1] + 1, n_events)
    S = np.array([rng.integers(k, 2 * k + 1) for k in K])
    seg_off = np.zeros(n_events + 1, dtype=np.int64)
    seg_off[1:] = np.cumsum(S)
    form_off = np.zeros(n_events + 1, dtype=np.int64)
    form_off[1:] = np.cumsum(K)
    seg_start = np.zeros(seg_off[-1], dtype=np.int64)
    seg_end = np.zeros(seg_off[-1], dtype=np.int64)
    form_seg_off = [0]
    form_seg_idx = []
    cls_off = [0]
    cls_mask, cls_cnt, bases = [], [], []
    ev_gs = np.zeros(n_events, dtype=np.int64)
    ev_ge = np.zeros(n_events, dtype=np.int64)
    pos = 1000
    for e in range(n_events):
        lens = rng.integers(30, 400, S[e])
        gaps = rng.integers(0, 300, S[e]) * (rng.random(S[e]) < 0.7)
        p = pos
        for s in range(S[e]):
            p += int(gaps[s])
            seg_start[seg_off[e] + s] = p
            p += int(lens[s])
            seg_end[seg_off[e] + s] = p
        ev_gs[e], ev_ge[e] = seg_start[seg_off[e]], p
        pos = p + 2000
        for k in range(K[e]):
            member = rng.random(S[e]) < 0.6
            member[rng.integers(0, S[e])] = True
            idx = np.nonzero(member)[0]
            form_seg_idx.extend(idx.tolist())
            form_seg_off.append(len(form_seg_idx))
        C = int(rng.integers(K[e], 3 * K[e] + 1))
        masks = set()
        while len(masks) < min(C, (1 << int(K[e])) - 1):
            masks.add(int(rng.integers(1, 1 << int(K[e]))))
        masks = sorted(masks)
        cnts = np.maximum(1, np.exp(rng.normal(3.0, 2.0, len(masks)))).astype(np.int64)
        cls_mask.extend(masks)
        cls_cnt.extend(cnts.tolist())
        cls_off.append(len(cls_mask))
        bases.append(int(cnts.sum()) * 75)
    return {
        "n_events": n_events, "K": K, "S": S, "seg_off": seg_off, "form_off": form_off,
        "seg_start": seg_start, "seg_end": seg_end, "ev_gs": ev_gs, "ev_ge": ev_ge,
        "form_seg_off": np.array(form_seg_off, dtype=np.int64), "form_seg_idx": np.array(form_seg_idx, dtype=np.int64),
        "cls_off": np.array(cls_off, dtype=np.uint64), "cls_mask": np.array(cls_mask, dtype=np.uint64),
        "cls_cnt": np.array(cls_cnt, dtype=np.uint64), "bases": np.array(bases, dtype=np.uint64),
    }


def make_multiform_case(n_events, reads_per_event, read_len, seed, k_range=(2, 16)):
    """BASELINE config 5 materialised: events with K~U{k_range} forms over S~U{K..2K} segments, every form a
    random subset of the segments (written as one exon per member segment; 70 % of the gaps are introns, the others
    abut), and reads drawn from the forms (Dirichlet proportions, uniform start on the spliced form, blocks cut at
    segment ends with abutting pieces fused).  Returns (SynthAnnotation, packed reads dict of numpy arrays) -- small
    enough to go through the reference's own solve binary."""
    rng = np.random.default_rng(seed)
    A = SynthAnnotation()
    A.chrom_names = ["chr%d" % (c + 1) for c in range(4)]
    cursor = [1000] * 4
    blk, nblk, cs = [], [], []
    for e in range(n_events):
        c = e % 4
        K = int(rng.integers(k_range[0], k_range[1] + 1))
        S = int(rng.integers(K, 2 * K + 1))
        lens = rng.integers(30, 400, S)
        gaps = rng.integers(1, 300, S) * (rng.random(S) < 0.7)
        segs = []
        p = cursor[c] + 3000
        for j in range(S):
            p += int(gaps[j])
            segs.append((p, p + int(lens[j])))
            p += int(lens[j])
        cursor[c] = p
        forms = []
        for k in range(K):
            member = rng.random(S) < 0.6
            member[rng.integers(0, S)] = True
            forms.append([segs[j] for j in range(S) if member[j]])
        strand = "+" if rng.random() < 0.5 else "-"
        for k, f in enumerate(forms):
            name = "M%d|%d" % (e, k + 1)
            A.forms.append((name, c, strand, f))
            A.map_rows.append(("mf_%d" % e, name))
        A.ev_gene.append(e)
        theta = rng.dirichlet(np.full(K, 0.7))
        n = int(rng.integers(max(2, reads_per_event // 4), 2 * reads_per_event))
        for k in rng.choice(K, size=n, p=theta):
            f = forms[k]
            T = sum(b - a for a, b in f)
            if T < read_len:
                continue
            st = int(rng.integers(0, T - read_len + 1))
            pieces, off = [], 0
            for a, b in f:
                lo, hi = max(st, off), min(st + read_len, off + (b - a))
                if hi > lo:
                    g0, g1 = a + lo - off, a + hi - off
                    if pieces and pieces[-1][1] == g0:
                        pieces[-1][1] = g1
                    else:
                        pieces.append([g0, g1])
                off += b - a
            nblk.append(len(pieces))
            blk.extend(pieces)
            cs.append(c | (int(rng.integers(0, 2)) << 16))
    A.n_genes = n_events
    A.chrom_len = [x + 10000 for x in cursor]
    n = len(nblk)
    off = np.zeros(n + 1, dtype=np.uint32)
    off[1:] = np.cumsum(nblk)
    b = np.asarray(blk, dtype=np.int32).reshape(-1)
    reads = {"blk_off": off, "blk": b, "read_cs": np.asarray(cs, dtype=np.uint32), "line_num": np.arange(1, n + 1, dtype=np.uint32),
             "chrom_names": list(A.chrom_names), "strand_names": ["-", "+"], "n_reads": n, "n_blocks": int(off[-1]),
             "total_read_bases": float((b[1::2] - b[0::2]).sum())}
    return A, reads


def make_solve_stress_fast(n_events, seed, k_range=(2, 16)):
    """Vectorised variant of make_solve_stress for the full BASELINE config-5 size (500 k events):
    same shapes (K~U{2..16}, S~U{K..2K}, C~U{K..3K} classes, log-normal multiplicities); class masks
    are drawn independently (duplicates within an event are legal and simply act as two classes)."""
    rng = np.random.default_rng(seed)
    E = n_events
    K = rng.integers(k_range[0], k_range[1] + 1, E)
    S = K + (rng.random(E) * (K + 1)).astype(np.int64)
    seg_off = np.zeros(E + 1, dtype=np.int64)
    seg_off[1:] = np.cumsum(S)
    form_off = np.zeros(E + 1, dtype=np.int64)
    form_off[1:] = np.cumsum(K)
    nseg = int(seg_off[-1])
    ev_of_seg = np.repeat(np.arange(E), S)
    lens = rng.integers(30, 400, nseg)
    gaps = rng.integers(0, 300, nseg) * (rng.random(nseg) < 0.7)
    first = np.zeros(nseg, dtype=bool)
    first[seg_off[:-1]] = True
    step = lens + gaps
    step[first] += 2000                      # distance between events
    # events are laid out left to right on 16 chromosomes so coordinates stay far below 2^31
    n_chroms = 16
    ev_chrom = (np.arange(E) * n_chroms // max(E, 1)).astype(np.int64)
    cum = np.cumsum(step)
    chrom_first_seg = seg_off[np.searchsorted(ev_chrom, np.arange(n_chroms))]
    base_of_chrom = cum[chrom_first_seg] - step[chrom_first_seg]
    ends = cum - base_of_chrom[ev_chrom[ev_of_seg]] + 1000
    seg_end = ends
    seg_start = ends - lens
    ev_gs = seg_start[seg_off[:-1]]
    ev_ge = seg_end[seg_off[1:] - 1]
    # forms: membership of every (form, segment-of-its-event) pair, at least one segment per form
    nform = int(form_off[-1])
    ev_of_form = np.repeat(np.arange(E), K)
    s_per_form = S[ev_of_form]
    pair_off = np.zeros(nform + 1, dtype=np.int64)
    pair_off[1:] = np.cumsum(s_per_form)
    npair = int(pair_off[-1])
    member = rng.random(npair) < 0.6
    forced = pair_off[:-1] + (rng.random(nform) * s_per_form).astype(np.int64)
    member[forced] = True
    local_idx = np.arange(npair) - np.repeat(pair_off[:-1], s_per_form)
    form_seg_idx = local_idx[member]
    cnt_per_form = np.add.reduceat(member.astype(np.int64), pair_off[:-1])
    form_seg_off = np.zeros(nform + 1, dtype=np.int64)
    form_seg_off[1:] = np.cumsum(cnt_per_form)
    # classes
    C = K + (rng.random(E) * (2 * K + 1)).astype(np.int64)
    C = np.minimum(C, (1 << K) - 1)
    cls_off = np.zeros(E + 1, dtype=np.uint64)
    cls_off[1:] = np.cumsum(C)
    ncls = int(cls_off[-1])
    k_of_cls = np.repeat(K, C)
    cls_mask = (1 + (rng.random(ncls) * ((1 << k_of_cls) - 1)).astype(np.int64)).astype(np.uint64)
    cls_cnt = np.maximum(1, np.exp(rng.normal(3.0, 2.0, ncls))).astype(np.uint64)
    bases = np.add.reduceat(cls_cnt, cls_off[:-1].astype(np.int64)) * np.uint64(75)
    return {"n_events": E, "K": K, "S": S, "seg_off": seg_off, "form_off": form_off, "seg_start": seg_start,
            "seg_end": seg_end, "ev_gs": ev_gs, "ev_ge": ev_ge, "ev_chrom": ev_chrom, "n_chroms": n_chroms,
            "form_seg_off": form_seg_off,
            "form_seg_idx": form_seg_idx, "cls_off": cls_off, "cls_mask": cls_mask, "cls_cnt": cls_cnt,
            "bases": bases.astype(np.uint64)}
