#!/usr/bin/env python
"""Small end-to-end run meant to be executed under compute-sanitizer (memcheck / racecheck / initcheck):
adversarial minimizer inputs, index build, segment build, mapping, all scan kernel generations."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from mapquik_b200 import Index, Params, sim
    import test_gpu_parity as T
    rng = np.random.default_rng(7)
    for gen in ("v3", "v2", "v1"):
        os.environ["MQ_SCAN_V1"] = "1" if gen == "v1" else "0"
        os.environ["MQ_SCAN_V2"] = "1" if gen == "v2" else "0"
        buf, offs = T.concat_raw(T.adversarial_seqs(rng))
        for p in (Params(), Params(l=16, density=0.2, use_hpc=False), Params(l=5, density=1.0)):
            T.check_minimizers(buf, offs, p)
        g, go, names = sim.genome(5, [120000, 40000])
        ix, oix = T.build_both(Params(), names, g, go)
        rb, ro, rn, _ = sim.reads(5, g, go, 200, 6000, 2000)
        T.compare_matches(ix, oix, rb, ro)
        T.compare_hits(ix, oix, rb, ro, rn)
        ix.close()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
