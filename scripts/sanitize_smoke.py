#!/usr/bin/env python
"""Small end-to-end run meant to be executed under compute-sanitizer (memcheck / racecheck / initcheck):
adversarial minimizer inputs in both input formats, index build, mapping (ASCII, packed, one context over two
device slots), the overflow / redo path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from mapquik_b200 import Index, PackedSeqs, Params, sim
    import test_gpu_parity as T
    rng = np.random.default_rng(7)
    for fmt in ("ascii", "packed"):
        buf, offs = T.concat_raw(T.adversarial_seqs(rng))
        for p in (Params(), Params(l=16, density=0.2, use_hpc=False), Params(l=5, density=1.0)):
            T.check_minimizers(buf, offs, p, fmt)
    g, go, names = sim.genome(5, [120000, 40000])
    ix, oix = T.build_both(Params(), names, g, go)
    rb, ro, rn, _ = sim.reads(5, g, go, 200, 6000, 2000)
    T.compare_matches(ix, oix, rb, ro)
    hits = T.compare_hits(ix, oix, rb, ro, rn)
    assert ix.map_batch_packed(PackedSeqs(rb), ro).tobytes() == hits.tobytes()
    multi = Index(Params(), devices=[0, 0]); multi.add_batch(names, g, go); multi.freeze()
    assert multi.map_batch(rb, ro).tobytes() == hits.tobytes()
    ix.close(); multi.close()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
