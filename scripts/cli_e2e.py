#!/usr/bin/env python
"""T_e of the drop-in CLI (host/mapquik): writes a workload as FASTA files, runs the CLI on them in several modes (packing
parser / --ascii / --gpus N), checks that every mode writes the byte-identical PAF and that it equals the library's own
hits, prints the timings as JSON.

    python scripts/cli_e2e.py [--config 2|3] [--reads N] [--gpus N] [--wrap COLUMNS]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_fasta(path, names, buf, offs, wrap=0):
    """one line per record, or `wrap` columns per line (what reference FASTA files look like)"""
    with open(path, "wb") as f:
        for i, n in enumerate(names):
            f.write(b">" + n.encode() + b"\n")
            s = buf[int(offs[i]):int(offs[i + 1])]
            if wrap and s.size:
                full = s.size // wrap * wrap
                if full:
                    lines = np.empty((full // wrap, wrap + 1), np.uint8)
                    lines[:, :wrap] = s[:full].reshape(-1, wrap); lines[:, wrap] = 10
                    f.write(lines.tobytes())
                if full < s.size:
                    f.write(s[full:].tobytes() + b"\n")
            else:
                f.write(s.tobytes())
                f.write(b"\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--reads", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--wrap", type=int, default=0, help="write the reference with this many columns per line (0: one line per record)")
    a = ap.parse_args()
    from mapquik_b200 import Index, Params, sim
    import bench
    cfg = bench.CONFIGS[a.config]
    n_reads = a.reads or (100000 if a.config == 2 else 200000)
    g, go, names = bench.make_genome(cfg)
    rb, ro, rn, _ = sim.reads(cfg["seed"], g, go, n_reads, cfg["mean"], cfg["sd"], cfg["min_len"], cfg["err"], contig_names=names)
    d = tempfile.mkdtemp(prefix="mq_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    ref, reads = os.path.join(d, "ref.fa"), os.path.join(d, "reads.fa")
    write_fasta(ref, names, g, go, a.wrap); write_fasta(reads, rn, rb, ro)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    ix = Index(Params()); ix.add_batch(names, g, go); ix.freeze()
    hits = ix.map_batch(rb, ro)
    exp = [ix.paf_line(rn[i], int(ro[i + 1] - ro[i]), hits[i]) for i in range(n_reads) if hits[i]["mapped"]]
    ix.close()
    modes = {"packed_1gpu": [], "ascii_1gpu": ["--ascii"]}
    if a.gpus > 1:
        modes[f"packed_{a.gpus}gpu"] = ["--gpus", str(a.gpus)]
        modes[f"ascii_{a.gpus}gpu"] = ["--gpus", str(a.gpus), "--ascii"]
    out = {"config": a.config, "reads": n_reads, "read_bp": int(ro[-1]), "genome_bp": int(go[-1]), "reference_columns": a.wrap or None,
           "reference_file_bytes": os.path.getsize(ref), "reads_file_bytes": os.path.getsize(reads), "host_cores": os.cpu_count(), "modes": {}}
    for tag, extra in modes.items():
        best = None
        for rep in range(2):                     # second run: files are in the page cache, CUDA modules loaded once more
            t0 = time.perf_counter()
            r = subprocess.run([os.path.join(ROOT, "host", "mapquik"), reads, "--reference", ref, "-p", os.path.join(d, tag)] + extra,
                               capture_output=True, text=True)
            wall = time.perf_counter() - t0
            assert r.returncode == 0, r.stderr
            if best is None or wall < best[0]:
                best = (wall, r.stdout)
        got = open(os.path.join(d, tag + ".paf")).read().splitlines()
        ln = {x.split(" in ")[0]: x for x in best[1].splitlines() if " in " in x}
        t_map = [float(x.split(" in ")[1].rstrip("s.")) for x in best[1].splitlines() if x.startswith("Mapped query sequences in")]
        t_idx = [float(x.split(" in ")[1].rstrip("s.")) for x in best[1].splitlines() if x.startswith("Indexed") and "unique" in x]
        out["modes"][tag] = {"cli_wall_s": best[0], "index_s": t_idx[0] if t_idx else None, "map_s": t_map[0] if t_map else None,
                             "gbp_per_s_map_phase": float(ro[-1]) / t_map[0] / 1e9 if t_map else None,
                             "paf_identical_to_library": got == exp, "paf_lines": len(got)}
    print(json.dumps(out))
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
