#!/usr/bin/env python
"""BASELINE.json configs[4]: figure-k-l style sweep (experiments/figure-k-l/{k,l,d}_perf.csv) of
k in {3..7}, l in {25..31}, density in {0.005, 0.01, 0.02} on the synthetic human-scale genome:
index size (k-min-mers, unique keys, table bytes, build time) versus probe/mapping throughput.
One GPU.  Prints CSV; --check N also runs the CPU oracle on a reduced genome for a few corner
combinations and asserts bit-identical hits."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--full-grid", action="store_true", help="all 105 combinations instead of the three 1-D sweeps")
    ap.add_argument("--parity-all", type=int, default=0, help="for EVERY combination also build the CPU-oracle index and compare the hits of this many reads")
    a = ap.parse_args()
    from mapquik_b200 import Index, Params, sim, capi
    tot = 3.1e9 / a.scale
    lens = [int(tot * p / sum(sim.CHM13_PROPS)) for p in sim.CHM13_PROPS]
    g, go, names = sim.genome(3, lens, sat_frac=0.06, segdup_frac=0.05)
    rb, ro, _, tr = sim.reads(3, g, go, a.reads, 24000, 3000, 1000, 0.005, with_names=False)
    L = capi.lib()
    ks, ls, ds = [3, 4, 5, 6, 7], [25, 26, 27, 28, 29, 30, 31], [0.005, 0.01, 0.02]
    if a.full_grid:
        combos = [(k, l, d) for k in ks for l in ls for d in ds]
    else:   # the reference's figures vary one parameter around the defaults (k=5, l=31, d=0.01)
        combos = [(k, 31, 0.01) for k in ks] + [(5, l, 0.01) for l in ls if l != 31] + [(5, 31, d) for d in ds if d != 0.01]
    print("k,l,density,n_kminmers,n_unique,n_keys,table_MB,index_gpu_ms,index_e2e_s,probe_ms,probes_per_s,map_kernels_ms,"
          "kernels_reads_per_s,e2e_reads_per_s,mapped,q60,wrong_q60" + (",parity" if a.parity_all else ""))
    if a.parity_all:
        from oracle import pyoracle as O
    for k, l, d in combos:
        ix = Index(Params(k=k, l=l, density=d))
        t0 = time.perf_counter()
        nb = ix.add_batch(names, g, go)
        gpu_ms = ix.last_ms("scan") + ix.last_ms("gather")
        ix.freeze()
        t_idx = time.perf_counter() - t0
        gpu_ms += ix.last_ms("insert")
        ix.map_batch(rb, ro)
        L.mq_minimizer_count(ix.handle, 1)
        t0 = time.perf_counter()
        hits = ix.map_batch(rb, ro)
        t_map = time.perf_counter() - t0
        st = {s: ix.last_ms(s) for s in ("scan", "gather", "probe", "chain")}
        n_min = L.mq_minimizer_count(ix.handle, 1)
        ok = (hits["mapped"] == 1) & (hits["ref_idx"] == tr["contig"]) & (hits["rc"] == tr["strand"]) & \
            (np.minimum(hits["r_end"], tr["start"] + tr["len"]).astype(np.int64) -
             np.maximum(hits["r_start"], tr["start"]).astype(np.int64) > 0.1 * tr["len"])
        kern = sum(st.values())
        nq = max(int(n_min), 1)          # minimizers of the timed map pass ~ index probes
        print(f"{k},{l},{d},{int(nb.sum())},{ix.n_unique},{ix.n_keys},{ix.table_bytes() / 1e6:.0f},{gpu_ms:.2f},{t_idx:.3f},"
              f"{st['probe']:.3f},{nq / (st['probe'] / 1e3):.3e},{kern:.2f},{a.reads / (kern / 1e3):.3e},{a.reads / t_map:.3e},"
              f"{int(hits['mapped'].sum())},{int((hits['mapq'] == 60).sum())},{int(((hits['mapq'] == 60) & ~ok).sum())}", end="", flush=True)
        if a.parity_all:
            oix = O.Index(O.params(k, l, d), int(nb.sum()) + 1024)
            onb = oix.add_batch(names, g, go, threads=os.cpu_count())
            n = min(a.parity_all, a.reads)
            oh = oix.map_batch(rb[:int(ro[n])], ro[:n + 1], threads=os.cpu_count())
            good = bool(np.array_equal(nb, onb) and oix.count() == ix.n_unique and oix.slots() == ix.n_keys and
                        oh.tobytes() == hits[:n].tobytes())
            print(f",{'ok' if good else 'MISMATCH'}", end="")
            del oix
        print(flush=True)
        ix.close()
    if a.check:
        from oracle import pyoracle as O
        lens2 = [max(int(x / 30), 50000) for x in lens]
        g2, go2, names2 = sim.genome(3, lens2, sat_frac=0.06, segdup_frac=0.05)
        rb2, ro2, _, _ = sim.reads(3, g2, go2, a.check, 24000, 3000, 1000, 0.005, with_names=False)
        for k, l, d in [(3, 25, 0.02), (7, 31, 0.005), (5, 28, 0.01), (7, 25, 0.02), (3, 31, 0.005)]:
            ix = Index(Params(k=k, l=l, density=d)); nb = ix.add_batch(names2, g2, go2); nu = ix.freeze()
            oix = O.Index(O.params(k, l, d), int(nb.sum()) + 1024); onb = oix.add_batch(names2, g2, go2)
            assert np.array_equal(nb, onb) and nu == oix.count() and ix.n_keys == oix.slots(), (k, l, d)
            assert ix.map_batch(rb2, ro2).tobytes() == oix.map_batch(rb2, ro2).tobytes(), (k, l, d)
            print(f"# parity ok k={k} l={l} d={d} on {int(go2[-1])} bp, {a.check} reads", flush=True)
            ix.close()


if __name__ == "__main__":
    main()
