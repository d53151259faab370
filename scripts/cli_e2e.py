#!/usr/bin/env python
"""T_e of the drop-in CLI: writes the bench workload (BASELINE configs[1]) as FASTA files, runs
host/mapquik on them, checks the PAF against the library's own hits, prints timings as JSON."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_fasta(path, names, buf, offs):
    with open(path, "wb") as f:
        for i, n in enumerate(names):
            f.write(b">" + n.encode() + b"\n")
            f.write(buf[int(offs[i]):int(offs[i + 1])].tobytes())
            f.write(b"\n")


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    from mapquik_b200 import Index, Params, sim
    g, go, names = sim.genome(2, [4641652], names=["chr000913"])
    rb, ro, rn, _ = sim.reads(2, g, go, n_reads, 10000, 1500, 1000, 0.005, contig_names=names)
    d = tempfile.mkdtemp(prefix="mq_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ref, reads = os.path.join(d, "ref.fa"), os.path.join(d, "reads.fa")
    write_fasta(ref, names, g, go); write_fasta(reads, rn, rb, ro)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "host", "mapquik"), reads, "--reference", ref, "-p", os.path.join(d, "out")],
                       capture_output=True, text=True)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr
    lines = {ln.split(" in ")[0].split(":")[0]: ln for ln in r.stdout.splitlines()}
    ix = Index(Params()); ix.add_batch(names, g, go); ix.freeze()
    hits = ix.map_batch(rb, ro)
    exp = [ix.paf_line(rn[i], int(ro[i + 1] - ro[i]), hits[i]) for i in range(n_reads) if hits[i]["mapped"]]
    got = open(os.path.join(d, "out.paf")).read().splitlines()
    out = {"reads": n_reads, "read_bp": int(ro[-1]), "cli_wall_s": wall, "paf_identical": got == exp, "paf_lines": len(got),
           "stdout_tail": r.stdout.splitlines()[-5:], "reads_per_s_cli": n_reads / wall, "gbp_per_s_cli": float(ro[-1]) / wall / 1e9}
    if os.environ.get("MQ_CLI_TIMING"):
        out["timing_lines"] = [ln for ln in r.stderr.splitlines() if ln.startswith("[")]
    print(json.dumps(out))
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
