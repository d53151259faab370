#!/usr/bin/env python
"""Small end-to-end run meant to be executed under compute-sanitizer (memcheck / racecheck / initcheck):
adversarial minimizer inputs in both input formats, index build, mapping (ASCII, packed, one context over two
device slots, ASCII with host threads packing sub-batches on the fly), the overflow / redo path."""
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
    # on-the-fly packing: small sub-batches so that this input is cut into ~10 of them and both routes are taken
    os.environ["MQ_SUB_BASES"] = str(1 << 17)
    hy = Index(Params()); hy.add_batch(names, g, go); hy.freeze(); hy.set_host_threads(3)
    rb2 = rb.copy(); rb2[1000:1040] = ord("N"); rb2[-5000:-4990] = ord("R")
    plain = Index(Params()); plain.add_batch(names, g, go); plain.freeze()
    want = plain.map_batch(rb2, ro)
    for _ in range(2):
        assert hy.map_batch(rb2, ro).tobytes() == want.tobytes() and hy.last_counter("host_packed_sub_batches") >= 1
    hy.close(); plain.close(); del os.environ["MQ_SUB_BASES"]
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
