#!/usr/bin/env python
"""Native stand-in for `paftools.js mapeval` (example/run_ecoli.sh:27) and for
experiments/intersect_pafs.py: accuracy of a PAF against the truth encoded in the read names
(`S1_<n>!<contig>!<start>!<end>!<strand>`), and concordance of two PAFs.

  mapeval.py eval  <paf> [--n-reads N]     per-MAPQ table: mapped / wrong (overlap with truth < 10 % of the
                                           truth interval, or wrong contig), cumulative like paftools
  mapeval.py intersect <paf1> <paf2>       concordant = same contig and overlap / union-span > 0.1
"""
import sys


def parse(path):
    out = {}
    for ln in open(path):
        c = ln.rstrip("\n").split("\t")
        if len(c) >= 12:
            out[c[0]] = (c[5], int(c[7]), int(c[8]), c[4], int(c[11]))
    return out


def truth(name):
    p = name.split("!")
    return p[1], int(p[2]), int(p[3]), p[4]


def is_correct(name, rec, frac=0.1):
    contig, ts, te, _ = truth(name)
    rc, rs, re, _, _ = rec
    if rc != contig:
        return False
    return min(te, re) - max(ts, rs) >= frac * (te - ts)


def evaluate(path, n_reads=None):
    recs = parse(path)
    by_q = {}
    for name, rec in recs.items():
        q = rec[4]
        m, w = by_q.get(q, (0, 0))
        by_q[q] = (m + 1, w + (0 if is_correct(name, rec) else 1))
    rows, cm, cw = [], 0, 0
    for q in sorted(by_q, reverse=True):
        m, w = by_q[q]; cm += m; cw += w
        rows.append((q, m, w, cw / cm if cm else 0.0, cm))
    return {"mapped": len(recs), "unmapped": (n_reads - len(recs)) if n_reads else None, "rows": rows,
            "wrong": cw, "correct": cm - cw}


def intersect(p1, p2):
    a, b = parse(p1), parse(p2)
    conc = disc = diffchr = 0
    for name in a.keys() & b.keys():
        c1, s1, e1 = a[name][:3]; c2, s2, e2 = b[name][:3]
        if c1 != c2:
            diffchr += 1
        elif (min(e1, e2) - max(s1, s2)) / max(1, max(e1, e2) - min(s1, s2)) > 0.1:
            conc += 1
        else:
            disc += 1
    return {"common": len(a.keys() & b.keys()), "only_1": len(a.keys() - b.keys()), "only_2": len(b.keys() - a.keys()),
            "concordant": conc, "discordant": disc, "different_contig": diffchr}


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "eval":
        n = int(sys.argv[sys.argv.index("--n-reads") + 1]) if "--n-reads" in sys.argv else None
        r = evaluate(sys.argv[2], n)
        for q, m, w, cum_err, cm in r["rows"]:
            print(f"Q\t{q}\t{m}\t{w}\t{cum_err:.6f}\t{cm}")
        print(f"mapped {r['mapped']} correct {r['correct']} wrong {r['wrong']} unmapped {r['unmapped']}")
    elif len(sys.argv) >= 4 and sys.argv[1] == "intersect":
        print(intersect(sys.argv[2], sys.argv[3]))
    else:
        sys.stderr.write(__doc__); sys.exit(1)


if __name__ == "__main__":
    main()
