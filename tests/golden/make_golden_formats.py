#!/usr/bin/env python
"""Generate tests/golden/formats.json: the other annotation formats `solve` reads (solve/solve.cpp:158-329),
from the REFERENCE'S OWN solve binary (run in the dev container; oracle/build_ref.sh makes it runnable).

Every case takes the LH_GENE_TXT annotation of an existing golden case, rewrites it as UCSC_GENE_TXT, UCSC_GFF,
WORMBASE_GFF2 or GENELETS_GFF3 (and the map as UCSC_GENE2ISOFORM or WORMBASE_GENE2ISOFORMS) and records what
/root/reference/bin/solve prints for it at 6 and at 17 significant digits.  The rewritten annotation is NOT
equivalent to the original in general -- the GFF loaders merge the exons of a form with interval_list semantics
and take chromosome/strand from the last exon line -- which is exactly what the fixtures pin.

    python tests/golden/make_golden_formats.py
"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import runner  # noqa: E402
import golden_util  # noqa: E402

FORM_FORMATS = ("UCSC_GENE_TXT", "UCSC_GFF", "WORMBASE_GFF2", "GENELETS_GFF3")


def parse_forms(text):
    rows = []
    for line in text.split("\n")[:-1]:
        t = line.split()
        if len(t) < 8:
            return None
        try:
            cnt = int(t[5])
            starts = [int(x) for x in t[6].split(",") if x]
            ends = [int(x) for x in t[7].split(",") if x]
            int(t[3]), int(t[4])
        except ValueError:
            return None
        if cnt < 0 or cnt > len(starts) or cnt > len(ends):
            return None
        rows.append(dict(name=t[0], chrom=t[1], strand=t[2], tx=(t[3], t[4]), cnt=cnt, starts=starts, ends=ends))
    return rows


def convert_forms(rows, fmt, variant):
    if fmt == "UCSC_GENE_TXT":
        return "".join("\t".join([r["name"], r["chrom"], r["strand"], r["tx"][0], r["tx"][1], r["tx"][0], r["tx"][1],
                                  str(r["cnt"]), ",".join(map(str, r["starts"])) + ",",
                                  ",".join(map(str, r["ends"])) + ","]) + "\n" for r in rows)
    if any(not r["chrom"].startswith("chr") for r in rows) and fmt != "UCSC_GFF":
        return None  # these loaders prepend "chr"
    if any(c in r["name"] for r in rows for c in ',;"') or any(not r["name"] for r in rows):
        return None
    out = []
    if fmt in ("UCSC_GFF", "GENELETS_GFF3"):
        out += ["##gff header line 1", "##gff header line 2"]
    if fmt == "GENELETS_GFF3" and variant % 2 == 1:
        # exons shared by several forms on one line: Parent=a,b
        shared = {}
        order = []
        for r in rows:
            for i in range(r["cnt"]):
                k = (r["chrom"], r["strand"], r["starts"][i], r["ends"][i])
                if k not in shared:
                    shared[k] = []
                    order.append(k)
                shared[k].append(r["name"])
        out.append("1\tsrc\tgene\t1\t2\t.\t+\t.\tID=ignored")  # not an exon line
        for n, k in enumerate(order):
            out.append("%s\tsrc\texon\t%d\t%d\t.\t%s\t.\tID=e%d;Parent=%s;Note=x" %
                       (k[0][3:], k[2] + 1, k[3], k[1], n, ",".join(shared[k])))
        return "\n".join(out) + "\n"
    for r in rows:
        idx = list(range(r["cnt"]))
        if variant % 3 == 2:
            idx.reverse()  # exon order matters to interval_list::add_interval
        for i in idx:
            s, e = r["starts"][i] + 1, r["ends"][i]
            if fmt == "UCSC_GFF":
                out.append('%s\tsrc\texon\t%d\t%d\t.\t%s\t.\t"%s"' % (r["chrom"], s, e, r["strand"], r["name"]))
            elif fmt == "WORMBASE_GFF2":
                out.append('%s\tsrc\texon\t%d\t%d\t.\t%s\t.\tTranscript\t"%s"' % (r["chrom"][3:], s, e, r["strand"], r["name"]))
            else:
                out.append("%s\tsrc\texon\t%d\t%d\t.\t%s\t.\tID=e%d;Parent=%s" % (r["chrom"][3:], s, e, r["strand"], i, r["name"]))
    return "\n".join(out) + "\n"


def convert_map(text, wormbase, variant):
    if not wormbase:
        return text
    rows = [ln.split() for ln in text.split("\n")[:-1]]
    if any(len(r) < 2 or ";" in r[1] for r in rows):
        return None
    out, cur = [], None
    for g, f in rows:
        if cur is not None and cur[0] == g and variant % 2 == 0:
            cur[1].append(f)  # one line per run of the same gene ...
        else:
            cur = [g, [f]]  # ... or one line per form (appending to the gene's list)
            out.append(cur)
    return "".join("%s\t%s\n" % (g, ";".join(fs) + (";" if variant % 4 == 0 else "")) for g, fs in out)


def write_case(case, d):
    fp, mp = os.path.join(d, "ev.forms"), os.path.join(d, "ev.map")
    open(fp, "w").write(case["forms"])
    open(mp, "w").write(case["map"])
    head = ["0", "p", os.path.join(d, "o_"), case["forms_format"], fp, case["map_format"], mp, str(case["begin"]), str(case["end"])]
    ca, sa = list(head), list(head)
    for i, s in enumerate(case["sets"]):
        rp = os.path.join(d, "r%d.mrf" % i)
        open(rp, "w").write(s["mrf"])
        ca += ["MRF_SINGLE", s["type"], str(s["read_length"]), rp]
        sa += ["MRF_SINGLE", s["type"], str(s["read_length"]), rp, repr(s["total_read_bases"])]
    return ca, sa


def main():
    src = golden_util.load("seed.json") + golden_util.load("fuzz.json")[:40] + golden_util.load("weird.json")[:12]
    out = []
    v = 0
    for case in src:
        if case["expect"]["solve_rc"] != 0 or not case["expect"]["solve17"]:
            continue
        rows = parse_forms(case["forms"])
        if rows is None:
            continue
        sets = []
        for s in case["sets"]:
            s2 = dict(s)
            if "mrf_repeat" in s2:
                s2["mrf"] = "AlignmentBlocks\n" + "".join((ln + "\n") * n for ln, n in s2.pop("mrf_repeat"))
            sets.append(s2)
        if sum(len(s["mrf"]) for s in sets) > 200000:
            continue
        for fmt in FORM_FORMATS:
            v += 1
            wormbase_map = ((v - 1) // 4 + (v - 1) % 4) % 2 == 1
            forms = convert_forms(rows, fmt, v)
            mp = convert_map(case["map"], wormbase_map, v)
            if forms is None or mp is None:
                continue
            c = {"name": "%s__%s%s" % (case["name"], fmt, "__WBMAP" if wormbase_map else ""), "forms_format": fmt,
                 "map_format": "WORMBASE_GENE2ISOFORMS" if wormbase_map else "UCSC_GENE2ISOFORM", "forms": forms, "map": mp,
                 "begin": case["begin"], "end": case["end"], "sets": sets}
            with tempfile.TemporaryDirectory() as d:
                _, sa = write_case(c, d)
                rc17, o17, _ = runner.run_ref("solve", sa, prec17=True, timeout=120)
                rc6, o6, _ = runner.run_ref("solve", sa, prec17=False, timeout=120)
            if rc17 != 0 or rc6 != 0 or not o17:
                print("skip %s: rc=%d" % (c["name"], rc17))
                continue
            c["expect"] = {"solve_rc": 0, "solve17": o17, "solve6": o6}
            out.append(c)
    json.dump(out, open(os.path.join(HERE, "formats.json"), "w"), indent=0)
    by = {}
    for c in out:
        by[c["forms_format"]] = by.get(c["forms_format"], 0) + 1
    print("wrote %d cases" % len(out), by, "wormbase maps:", sum(c["map_format"].startswith("W") for c in out))


if __name__ == "__main__":
    main()
