#!/usr/bin/env python
"""Regenerates the fixtures in tests/golden/ (run in the build container, where the upstream tree
is mounted read-only at /root/reference; the GPU box never sees that path).

  nearperfect-ecoli.100.fa.gz   the reference's own 100-read example fixture
                                (example/nearperfect-ecoli.100.fa, md5 1b7f89d3644d48476caf13c8f9e5b6cc),
                                gzip -9, bytes otherwise untouched
  config1_default.paf           CPU-oracle PAF of those reads against the scaffold stand-in genome
  config1_script.paf            (see scaffold_genome below), default parameters and the parameters of
                                example/run_ecoli.sh:26 (-k 8 -d 0.01 -l 16 -g 100)
  seeding_vectors.json          minimizers / k-min-mers of small fixed sequences computed by the
                                definitional Python restatement tests/pyref.py (NOT by the C oracle)

The genome of config 1 (example/ecoli.genome.fa) is listed in the reference's .MISSING_LARGE_BLOBS;
the stand-in is `chr000913`, 4,641,652 bp (example/ecoli.genome.fa.fai:1), seeded random bases with
every fixture read pasted at the start coordinate its name encodes (reverse-complemented for '-').
It is a stand-in, not E. coli: results on it say nothing about upstream accuracy.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

GENOME_NAME, GENOME_LEN = "chr000913", 4641652


def read_fasta_gz(path):
    names, seqs = [], []
    with gzip.open(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                names.append(line[1:].split()[0].decode()); seqs.append([])
            elif line:
                seqs[-1].append(line)
    return names, [np.frombuffer(b"".join(s).upper(), dtype=np.uint8) for s in seqs]


def scaffold_genome(names, seqs):
    from mapquik_b200 import sim
    from conftest import revcomp
    g, _, _ = sim.genome(1, [GENOME_LEN], names=[GENOME_NAME])
    g = g.copy()
    for n, s in zip(names, seqs):
        _, contig, start, end, strand = n.split("!")
        start = int(start)
        t = revcomp(s) if strand == "-" else s
        m = min(len(t), GENOME_LEN - start)
        g[start:start + m] = t[:m]
    return g


def oracle_paf(names, seqs, genome, **kw):
    from oracle import pyoracle as O
    ix = O.Index(O.params(**kw), 1 << 17)
    ix.add_ref(GENOME_NAME, genome)
    lines = []
    for n, s in zip(names, seqs):
        hit = ix.find_matches(s)
        if hit["mapped"]:
            lines.append(ix.paf_line(n, len(s), hit))
    return lines


def seeding_vectors():
    import pyref
    rng = np.random.default_rng(20260101)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    seqs = {
        "random_3000": acgt[rng.integers(0, 4, 3000)].tobytes().decode(),
        "runs_1200": np.repeat(acgt[rng.integers(0, 4, 400)], rng.integers(1, 6, 400)).tobytes().decode()[:1200],
        "with_N": (acgt[rng.integers(0, 4, 500)].tobytes() + b"NNNNN" + acgt[rng.integers(0, 4, 500)].tobytes()).decode(),
    }
    out = {"sequences": seqs, "cases": []}
    for name, s in seqs.items():
        for k, l, d, hpc in ((5, 31, 0.05, True), (3, 16, 0.1, True), (4, 21, 0.08, False)):
            mins = pyref.minimizers(s.encode(), l, d, hpc)
            kms = pyref.kminmers(s.encode(), k, l, d, hpc)
            out["cases"].append({"seq": name, "k": k, "l": l, "density": d, "hpc": hpc,
                                 "minimizers": [[p, str(h)] for p, h in mins],
                                 "kminmers": [[m.start, m.end, m.offset, int(m.rev), str(m.hash)] for m in kms]})
    return out


def main():
    src = "/root/reference/example/nearperfect-ecoli.100.fa"
    raw = open(src, "rb").read()
    assert hashlib.md5(raw).hexdigest() == "1b7f89d3644d48476caf13c8f9e5b6cc"
    gz = os.path.join(HERE, "nearperfect-ecoli.100.fa.gz")
    with open(gz, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", compresslevel=9, mtime=0) as z:
            z.write(raw)
    names, seqs = read_fasta_gz(gz)
    assert len(names) == 100
    g = scaffold_genome(names, seqs)
    for fn, kw in (("config1_default.paf", {}), ("config1_script.paf", dict(k=8, l=16, density=0.01, g=100))):
        lines = oracle_paf(names, seqs, g, **kw)
        open(os.path.join(HERE, fn), "w").write("\n".join(lines) + "\n")
        print(fn, len(lines), "lines")
    json.dump(seeding_vectors(), open(os.path.join(HERE, "seeding_vectors.json"), "w"))
    print("scaffold md5", hashlib.md5(g.tobytes()).hexdigest())


if __name__ == "__main__":
    main()
