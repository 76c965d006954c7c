"""Seeded adversarial generator of tiny annotation + MRF_SINGLE inputs.

Aims at the quirks listed in SURVEY.md section 3.3: lower_bound tie-break on
(start, end, strand, name), per-block covered-region filter with order-dependent merging,
truncating walk, the IEEE `> 0.98` rule, lexicographic gene order and range selection, comment /
header lines that still advance the read number, dropped unterminated last line, multiple
read sets, forms whose ARS length is 0.  Used to (a) build tests/golden/fuzz_*.json from the
reference binaries (tests/golden/make_golden.py) and (b) drive oracle-vs-CUDA parity tests.
"""
import random


def _rand_exons(rng, lo, n, min_len=1, max_len=120, max_gap=150, abut_p=0.25):
    exons = []
    pos = lo
    for _ in range(n):
        gap = 0 if (exons and rng.random() < abut_p) else rng.randint(1 if exons else 0, max_gap)
        s = pos + gap
        e = s + rng.randint(min_len, max_len)
        exons.append((s, e))
        pos = e
    return exons


def _derive_form(rng, base):
    """A variant of `base`: drop / shift / split / extend exons so boundaries interleave."""
    out = []
    for (s, e) in base:
        r = rng.random()
        if r < 0.2 and len(base) > 1:
            continue
        if r < 0.35:
            s = max(0, s + rng.randint(-15, 15))
        elif r < 0.5:
            e = e + rng.randint(-15, 15)
        elif r < 0.58 and e - s > 6:
            m = rng.randint(s + 2, e - 2)
            out.append((s, m))
            s = m + (0 if rng.random() < 0.5 else rng.randint(1, 3))
        if e <= s:
            e = s + rng.randint(1, 10)
        if out and s < out[-1][1]:
            s = out[-1][1] + (0 if rng.random() < 0.5 else rng.randint(0, 5))
            if e <= s:
                e = s + rng.randint(1, 10)
        out.append((s, e))
    if not out:
        out = [base[0]]
    return out


GNAME_POOL = ["1", "10", "100", "2", "21", "3", "z", "uc001", "read-3", "read-", "read-12", "G_a", "SE_7",
              "A", "b", "9", "read-2x", "r", "read", "~t"]


def _weird_form(rng, ex):
    """Annotation shapes a curated Events.r output never has but the reference accepts: exons out
    of order, zero-length exons, exact duplicates, overlapping exons inside one form."""
    ex = list(ex)
    r = rng.random()
    if r < 0.3 and len(ex) > 1:
        i = rng.randrange(len(ex) - 1)
        ex[i], ex[i + 1] = ex[i + 1], ex[i]               # out of order
    elif r < 0.5:
        s = rng.choice(ex)[rng.randint(0, 1)] + rng.randint(-3, 3)
        ex.insert(rng.randrange(len(ex) + 1), (max(0, s), max(0, s)))  # zero-length exon
    elif r < 0.7:
        ex.insert(rng.randrange(len(ex) + 1), rng.choice(ex))          # duplicate exon
    elif r < 0.9:
        s, e = rng.choice(ex)
        ex.append((max(0, s - rng.randint(0, 20)), e + rng.randint(0, 20)))  # overlapping, appended last
    return ex


def make_case(seed, n_events=None, n_reads=None, two_sets=None, weird=False):
    rng = random.Random(seed)
    n_events = n_events or rng.randint(1, 6)
    chroms = ["chr1", "chr2", "chrX"][: rng.randint(1, 3)]
    strands = ["+", "-"]
    gnames = rng.sample(GNAME_POOL, n_events)
    forms_lines, map_lines = [], []
    events = []
    cursor = {c: rng.randint(0, 50) for c in chroms}
    for g in gnames:
        c = rng.choice(chroms)
        # events may overlap the previous one on the chromosome
        if rng.random() < 0.3:
            cursor[c] = max(0, cursor[c] - rng.randint(50, 400))
        lo = cursor[c] + rng.randint(0, 300)
        base = _rand_exons(rng, lo, rng.randint(1, 5))
        k = rng.choice([1, 2, 2, 2, 3, 4])
        fl = [base] + [_derive_form(rng, base) for _ in range(k - 1)]
        st = rng.choice(strands)
        for i, ex in enumerate(fl):
            if weird and rng.random() < 0.5:
                ex = _weird_form(rng, ex)
            fname = "%s|%d" % (g, i + 1)
            fst = st if rng.random() < 0.9 else rng.choice(strands)
            forms_lines.append("\t".join([
                fname, c, fst, str(ex[0][0]), str(ex[-1][1]), str(len(ex)),
                ",".join(str(s) for s, _ in ex) + ("," if rng.random() < 0.5 else ""),
                ",".join(str(e) for _, e in ex) + ("," if rng.random() < 0.5 else "")]))
            map_lines.append("%s\t%s" % (g, fname))
        events.append((g, c, st, fl))
        cursor[c] = max(e for ex in fl for _, e in ex)
    if rng.random() < 0.3:
        rng.shuffle(map_lines)

    def rand_read(L):
        g, c, st, fl = rng.choice(events)
        ex = rng.choice(fl)
        r = rng.random()
        rs = rng.choice(strands)
        gs = min(s for f in fl for s, _ in f)
        ge = max(e for f in fl for _, e in f)
        if r < 0.08:   # exactly the event span (tie-break path)
            return [(c, rs, gs + 1, ge)]
        if r < 0.16:   # starts at the event start
            e0 = [e for f in fl for s, e in f if s == gs][0]
            return [(c, rs, gs + 1, rng.randint(gs + 1, e0))]
        if r < 0.22:   # background / other chromosome
            cc = rng.choice(chroms + ["chrUn"])
            s = rng.randint(1, 1500)
            return [(cc, rs, s, s + L - 1)]
        # walk L bases along the spliced form from a random offset
        tot = sum(e - s for s, e in ex)
        off = rng.randint(0, max(0, tot - 1))
        need = L if rng.random() < 0.8 else rng.randint(1, L)
        blocks = []
        acc = 0
        for (s, e) in ex:
            if off >= acc + (e - s):
                acc += e - s
                continue
            bs = s + max(0, off - acc)
            take = min(e - bs, need)
            if take <= 0:
                break
            blocks.append([bs + 1, bs + take])
            need -= take
            acc += e - s
            off = acc
            if need <= 0:
                break
        if not blocks:
            blocks = [[ex[0][0] + 1, ex[0][1]]]
        q = rng.random()
        if q < 0.10:    # overhang by a base or two on a junction
            b = rng.choice(blocks)
            b[rng.randint(0, 1)] += rng.choice([-2, -1, 1, 2])
        elif q < 0.15 and len(blocks) > 1:
            rng.shuffle(blocks)         # out-of-order blocks
        elif q < 0.20:                  # split one block into two abutting ones
            b = rng.choice(blocks)
            if b[1] - b[0] >= 2:
                m = rng.randint(b[0], b[1] - 1)
                i = blocks.index(b)
                nb = [[b[0], m], [m + 1, b[1]]]
                if rng.random() < 0.5:
                    nb.reverse()
                blocks[i:i + 1] = nb
        elif q < 0.23:                  # orphan 1-bp block far away
            blocks.append([blocks[-1][1] + rng.randint(5, 400), 0])
            blocks[-1][1] = blocks[-1][0]
        elif q < 0.25:                  # block on another chromosome
            blocks.append(None)
        out = []
        for b in blocks:
            if b is None:
                oc = rng.choice(chroms)
                s = rng.randint(1, 800)
                out.append((oc, rs, s, s + rng.randint(0, 30)))
            else:
                if b[1] < b[0] - 1:
                    b[1] = b[0]
                out.append((c, rs, max(1, b[0]), b[1]))
        return out

    two_sets = (rng.random() < 0.25) if two_sets is None else two_sets
    sets = []
    for m in range(2 if two_sets else 1):
        L = rng.choice([25, 36, 50, 75, 100])
        n = n_reads or rng.randint(5, 120)
        lines = ["AlignmentBlocks"]
        total_bases = 0
        for _ in range(n):
            r = rng.random()
            if r < 0.03:
                lines.append("# a comment line")
                continue
            if r < 0.05:
                lines.append("AlignmentBlocks")
                continue
            blocks = rand_read(L)
            q = 1
            parts = []
            for (c, st, s, e) in blocks:
                ln = e - s + 1
                parts.append("%s:%s:%d:%d:%d:%d" % (c, st, s, e, q, q + ln - 1))
                q += ln
                total_bases += max(ln, 0)
            lines.append(",".join(parts))
        text = "\n".join(lines) + ("\n" if rng.random() < 0.9 else "")
        rtype = rng.choice(["SHORT_READ", "SHORT_READ", "MEDIUM_READ"])
        sets.append({"type": rtype, "read_length": L, "mrf": text,
                     "total_read_bases": float(max(total_bases, 1))})
    ng = len(set(gnames))
    if rng.random() < 0.7:
        b, e = 0, 100
    else:
        b = rng.randint(0, ng - 1)
        e = rng.randint(b, ng + 1)
    forms_text = "\n".join(forms_lines) + "\n"
    map_text = "\n".join(map_lines) + ("\n" if rng.random() < 0.9 else "")
    return {"seed": seed, "forms": forms_text, "map": map_text, "sets": sets, "begin": b, "end": e}


def write_case(case, d):
    """Materialise a case in directory `d`; returns (count_args, solve_args)."""
    import os
    fp = os.path.join(d, "ev.interval")
    mp = os.path.join(d, "ev.map")
    open(fp, "w").write(case["forms"])
    open(mp, "w").write(case["map"])
    head = ["0", "p", os.path.join(d, "o_"), "LH_GENE_TXT", fp, "UCSC_GENE2ISOFORM", mp,
            str(case["begin"]), str(case["end"])]
    ca, sa = list(head), list(head)
    for i, s in enumerate(case["sets"]):
        rp = os.path.join(d, "r%d.mrf" % i)
        open(rp, "w").write(s["mrf"])
        ca += ["MRF_SINGLE", s["type"], str(s["read_length"]), rp]
        sa += ["MRF_SINGLE", s["type"], str(s["read_length"]), rp, repr(s["total_read_bases"])]
    return ca, sa
