#!/usr/bin/env python
"""Throughput of the CLI's FASTX parser alone (host/mapquik --parse-only with MQ_CLI_NODIGEST=1: records are parsed and
packed exactly as for a mapping run, nothing is mapped): single-line reads and a 60-column reference, by batch size, with and
without MADV_POPULATE_READ, minimum of several runs.  No GPU involved.

    python scripts/parser_bench.py [--gb 2.0] [--reps 4]
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


def random_bases(rng, n):
    return np.frombuffer(b"ACGT", np.uint8)[np.frombuffer(rng.bytes(n), np.uint8) & 3]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    d = tempfile.mkdtemp(prefix="mq_parse_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    reads, ref, fq = os.path.join(d, "reads.fa"), os.path.join(d, "ref60.fa"), os.path.join(d, "reads.fastq")
    L = 24000
    n_reads = int(a.gb * 1e9 / L)
    with open(reads, "wb") as f, open(fq, "wb") as q:
        qual = b"I" * L
        for i in range(0, n_reads, 4000):
            m = min(4000, n_reads - i)
            s = random_bases(rng, m * L).reshape(m, L)
            for j in range(m):
                b = s[j].tobytes()
                f.write(b">read%d\n" % (i + j) + b + b"\n")
                if i + j < n_reads // 2:
                    q.write(b"@read%d\n" % (i + j) + b + b"\n+\n" + qual + b"\n")
    with open(ref, "wb") as f:
        for c in range(4):
            n = int(a.gb * 1e9 / 8) // 60 * 60
            lines = np.empty((n // 60, 61), np.uint8)
            lines[:, :60] = random_bases(rng, n).reshape(-1, 60); lines[:, 60] = 10
            f.write(b">chr%d\n" % c + lines.tobytes())
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    exe = os.path.join(ROOT, "host", "mapquik")
    out = {"host_cores": os.cpu_count(), "files": {}, "runs": []}
    for path in (reads, ref, fq):
        out["files"][os.path.basename(path)] = os.path.getsize(path)
    variants = [(reads, mb, pop) for mb in (16, 32, 64, 128, 256) for pop in (1, 0)] + [(ref, mb, 0) for mb in (128, 512)] + [(fq, 64, 0), (fq, 256, 0)]
    for path, mb, pop in variants:
        env = dict(os.environ, MQ_CLI_NODIGEST="1", MQ_CLI_PACK="1", MQ_CLI_BATCH=str(mb << 20))
        if pop:
            env["MQ_CLI_POPULATE"] = "1"
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            subprocess.run([exe, path, "--parse-only"], env=env, check=True, stdout=subprocess.DEVNULL)
            ts.append(time.perf_counter() - t0)
        out["runs"].append({"file": os.path.basename(path), "batch_mb": mb, "populate": bool(pop), "min_s": min(ts), "median_s": sorted(ts)[len(ts) // 2],
                            "gb_per_s": os.path.getsize(path) / min(ts) / 1e9})
    print(json.dumps(out))
    for f in os.listdir(d):
        os.unlink(os.path.join(d, f))
    os.rmdir(d)


if __name__ == "__main__":
    main()
