#!/usr/bin/env python
"""Opportunistic parity check against a REAL upstream mapquik binary.

The seeding arithmetic of upstream lives in an un-vendored, un-pinned crate (rust-seq2kminmers, Cargo.toml:30), so this
repo's S1/S2 spec is its own and parity with upstream binaries is unpinned.  If a `mapquik` binary built from
ekimb/mapquik is reachable -- baseline/_ref/{bin/,}mapquik or $PATH -- this script runs it exactly as the reference's
own example does (`mapquik <reads> --reference <ref> -p <prefix> --threads $(nproc)`, main.rs:168-272) on the config-1
fixture and on a small synthetic config, runs this repo's CLI on the same files, and diffs the sorted PAFs.
It prints one JSON line either way: {"upstream": {"found": false}} when there is nothing to compare with.

    python scripts/upstream_probe.py [--binary PATH] [--keep DIR]
"""
import argparse
import gzip
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OURS = os.path.join(ROOT, "host", "mapquik")


def find_upstream(explicit=None):
    cands = [explicit, os.path.join(ROOT, "baseline", "_ref", "bin", "mapquik"), os.path.join(ROOT, "baseline", "_ref", "mapquik"),
             shutil.which("mapquik")]
    for c in cands:
        if c and os.path.isfile(c) and os.access(c, os.X_OK) and os.path.realpath(c) != os.path.realpath(OURS):
            return c
    return None


def write_inputs(d):
    """config 1 (the reference's 100 reads against the scaffold stand-in of the missing genome) and a synthetic config"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as MG
    from mapquik_b200 import sim
    out = []
    names, seqs = MG.read_fasta_gz(os.path.join(ROOT, "tests", "golden", "nearperfect-ecoli.100.fa.gz"))
    g = MG.scaffold_genome(names, seqs)
    ref1 = os.path.join(d, "scaffold.genome.fa")
    with open(ref1, "wb") as f:
        f.write(b">chr000913\n" + g.tobytes() + b"\n")         # single-line FASTA: upstream accepts nothing else (README.md:36)
    reads1 = os.path.join(d, "nearperfect-ecoli.100.fa")
    open(reads1, "wb").write(gzip.open(os.path.join(ROOT, "tests", "golden", "nearperfect-ecoli.100.fa.gz"), "rb").read())
    out.append(("config1_default", ref1, reads1, []))
    out.append(("config1_script", ref1, reads1, ["-k", "8", "-d", "0.01", "-l", "16", "-g", "100"]))
    g2, go2, n2 = sim.genome(9, [3000000, 1200000])
    rb, ro, rn, _ = sim.reads(9, g2, go2, 5000, 10000, 1500, contig_names=n2)
    ref2, reads2 = os.path.join(d, "synth.fa"), os.path.join(d, "synth.reads.fa")
    with open(ref2, "wb") as f:
        for i, n in enumerate(n2):
            f.write(b">" + n.encode() + b"\n" + g2[int(go2[i]):int(go2[i + 1])].tobytes() + b"\n")
    with open(reads2, "wb") as f:
        for i, n in enumerate(rn):
            f.write(b">" + n.encode() + b"\n" + rb[int(ro[i]):int(ro[i + 1])].tobytes() + b"\n")
    out.append(("synthetic_4.2Mbp_5k_reads", ref2, reads2, []))
    return out


def run(binary, reads, ref, prefix, extra):
    cmd = [binary, reads, "--reference", ref, "-p", prefix, "--threads", str(os.cpu_count() or 1)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(prefix + ".paf"):
        return None, (r.stderr or r.stdout)[-500:]
    return sorted(open(prefix + ".paf").read().splitlines()), r.stdout[-800:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--binary")
    ap.add_argument("--keep")
    a = ap.parse_args()
    up = find_upstream(a.binary)
    if not up:
        print(json.dumps({"upstream": {"found": False, "paf_identical": None,
                                       "note": "no mapquik binary under baseline/_ref or on $PATH; parity with upstream stays unpinned"}}))
        return
    d = a.keep or tempfile.mkdtemp(prefix="mq_upstream_")
    os.makedirs(d, exist_ok=True)
    res = {"found": True, "path": up, "cases": {}}
    all_same = True
    for tag, ref, reads, extra in write_inputs(d):
        theirs, log_t = run(up, reads, ref, os.path.join(d, tag + ".upstream"), extra)
        ours, log_o = run(OURS, reads, ref, os.path.join(d, tag + ".b200"), extra)
        case = {"upstream_ran": theirs is not None, "b200_ran": ours is not None}
        if theirs is not None and ours is not None:
            same = theirs == ours
            case.update(paf_identical=same, upstream_lines=len(theirs), b200_lines=len(ours))
            if not same:
                ids_t = {ln.split("\t")[0]: ln for ln in theirs}; ids_o = {ln.split("\t")[0]: ln for ln in ours}
                diff = [i for i in ids_t if ids_o.get(i) != ids_t[i]] + [i for i in ids_o if i not in ids_t]
                case.update(differing_reads=len(diff), first_difference={"upstream": ids_t.get(diff[0]), "b200": ids_o.get(diff[0])} if diff else None)
            all_same &= same
        else:
            case["log"] = log_t if theirs is None else log_o
            all_same = False
        res["cases"][tag] = case
    res["paf_identical"] = all_same
    print(json.dumps({"upstream": res}))
    if not a.keep:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
