#!/usr/bin/env python
"""Runs one of the BASELINE.json configs end to end on one GPU, checks the GPU path against the CPU
oracle (full index counts; hits of a read subsample byte for byte) and prints one JSON line.

  --config 3   synthetic CHM13-sized genome: 24 contigs, CHM13 length proportions, 3.1 Gbp, ~6 % tandem
               satellites, ~5 % segmental duplications, seed 3; reads 24 kb (sd 3 kb), 99.5 %
  --config 4   synthetic maize-like genome: 10 contigs, 2.2 Gbp, ~85 % from 300 repeat families, seed 4
  --config 2   the bench workload (E. coli-sized, 10 kb reads)
  --scale F    shrink the genome by F (quick runs); --reads N reads mapped on the GPU; --check N of them
               are also mapped by the oracle and compared.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reads", type=int, default=200000)
    ap.add_argument("--check", type=int, default=20000)
    ap.add_argument("--k", type=int, default=5); ap.add_argument("--l", type=int, default=31)
    ap.add_argument("--density", type=float, default=0.01)
    ap.add_argument("--no-oracle-index", action="store_true")
    ap.add_argument("--pinned", action="store_true", help="copy genome and reads into pinned host buffers first (full PCIe speed)")
    a = ap.parse_args()
    from mapquik_b200 import Index, Params, sim
    from oracle import pyoracle as O

    t0 = time.perf_counter()
    if a.config == 3:
        tot = 3.1e9 / a.scale
        lens = [int(tot * p / sum(sim.CHM13_PROPS)) for p in sim.CHM13_PROPS]
        g, go, names = sim.genome(3, lens, sat_frac=0.06, segdup_frac=0.05)
        mean, sd, seed = 24000, 3000, 3
    elif a.config == 4:
        tot = 2.2e9 / a.scale
        lens = [int(tot / 10)] * 10
        g, go, names = sim.genome(4, lens, repeat_frac=0.85, n_families=300)
        mean, sd, seed = 24000, 3000, 4
    else:
        g, go, names = sim.genome(2, [int(4641652 / a.scale)], names=["chr000913"])
        mean, sd, seed = 10000, 1500, 2
    t_gen = time.perf_counter() - t0
    rb, ro, rn, tr = sim.reads(seed, g, go, a.reads, mean, sd, 1000, 0.005, with_names=False)
    if a.pinned:
        import ctypes as C
        from mapquik_b200 import capi
        Lc = capi.lib()

        def pin(arr):
            ptr = Lc.mq_host_alloc(arr.nbytes + 64)
            v = np.frombuffer((C.c_uint8 * arr.nbytes).from_address(ptr), dtype=arr.dtype)
            v[:] = arr
            return v
        g, rb = pin(g), pin(rb)
    p = Params(k=a.k, l=a.l, density=a.density)
    out = {"pinned": bool(a.pinned), "config": a.config, "genome_bp": int(go[-1]), "contigs": len(names), "reads": a.reads, "read_bp": int(ro[-1]),
           "k": a.k, "l": a.l, "density": a.density, "gen_s": t_gen}

    ix = Index(p)
    t0 = time.perf_counter()
    nb = ix.add_batch(names, g, go)
    t_add = time.perf_counter() - t0
    add_ms = {s: ix.last_ms(s) for s in ("h2d", "scan", "scan_kernel", "gather")}
    t1 = time.perf_counter()
    n_unique = ix.freeze()
    t_freeze = time.perf_counter() - t1
    out.update(index_build_s=t_add + t_freeze, index_add_s=t_add, index_freeze_s=t_freeze, index_add_stage_ms=add_ms,
               index_insert_ms=ix.last_ms("insert"), n_kminmers=int(nb.sum()), n_unique=int(n_unique), n_keys=int(ix.n_keys),
               table_bytes=int(ix.table_bytes()))

    t0 = time.perf_counter()
    hits = ix.map_batch(rb, ro)
    t_map = time.perf_counter() - t0
    t0 = time.perf_counter()
    hits = ix.map_batch(rb, ro)
    t_map2 = time.perf_counter() - t0
    out.update(map_s_first=t_map, map_s=t_map2, reads_per_s=a.reads / t_map2, gbp_per_s=float(ro[-1]) / t_map2 / 1e9,
               map_stage_ms={s: ix.last_ms(s) for s in ("h2d", "scan", "gather", "probe", "chain", "d2h")},
               mapped=int(hits["mapped"].sum()), q60=int((hits["mapq"] == 60).sum()))
    ok = (hits["mapped"] == 1) & (hits["ref_idx"] == tr["contig"]) & (hits["rc"] == tr["strand"]) & \
        (np.minimum(hits["r_end"], tr["start"] + tr["len"]).astype(np.int64) -
         np.maximum(hits["r_start"], tr["start"]).astype(np.int64) > 0.1 * tr["len"])
    out.update(correct=int(ok.sum()), wrong_q60=int(((hits["mapq"] == 60) & ~ok).sum()))

    if not a.no_oracle_index:
        threads = O.lib().orc_max_threads()
        oix = O.Index(O.params(a.k, a.l, a.density), int(nb.sum()) + 1024)
        t0 = time.perf_counter()
        onb = oix.add_batch(names, g, go, threads=threads)
        o_unique = oix.count()
        t_oidx = time.perf_counter() - t0
        nchk = min(a.check, a.reads)
        cro = ro[:nchk + 1]; crb = rb[:int(cro[-1])]
        t0 = time.perf_counter()
        ohits = oix.map_batch(crb, cro, threads=threads)
        t_omap = time.perf_counter() - t0
        out.update(oracle_threads=threads, oracle_index_s=t_oidx, oracle_map_reads_per_s=nchk / t_omap,
                   parity_nb_mers=bool(np.array_equal(nb, onb)), parity_n_unique=bool(o_unique == n_unique),
                   parity_n_keys=bool(oix.slots() == ix.n_keys), parity_hits=bool(ohits.tobytes() == hits[:nchk].tobytes()),
                   checked_reads=nchk)
    print(json.dumps(out), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
