#!/usr/bin/env python
"""mq_map_batch on ASCII in pinned host memory, config-3-like reads, by number of host threads given to the library
(mq_set_host_threads): reads/s, share of the bases that took the packed route, bytes over the link."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from mapquik_b200 import HIT_DTYPE, Index, Params, capi

ap = argparse.ArgumentParser(); ap.add_argument("--reads", type=int, default=400000); ap.add_argument("--threads", default="0,4,8,12,14,15,16,20,24")
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
cfg = bench.CONFIGS[3]; L = capi.lib(); p = Params()
g0, go, names = bench.make_genome(cfg)
g, pg = bench.pinned_array(L, g0.size + 64); g = g[:g0.size]; g[:] = g0; del g0
nb = bench.read_lengths_total(cfg, go, 0, a.reads)
h_seqs, p1 = bench.pinned_array(L, nb + 64)
rb, ro, truth = bench.make_reads(cfg, g, go, 0, a.reads, out=h_seqs)
h_offs, p2 = bench.pinned_array(L, (a.reads + 1) * 8, np.uint64); h_offs[:] = ro
h_hits, p3 = bench.pinned_array(L, a.reads * 48); hv = h_hits.view(HIT_DTYPE)
ix = Index(p); ix.add_batch(names, g, go); ix.freeze()
want = None; out = []
for t in [int(x) for x in a.threads.split(",")]:
    ix.set_host_threads(t)
    for _ in range(2):
        ix.map_batch(h_seqs[:nb], h_offs, out=hv)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ix.map_batch(h_seqs[:nb], h_offs, out=hv)
    dt = (time.perf_counter() - t0) / a.steps
    if want is None:
        want = hv.copy()
    rec = {"host_threads": t, "reads_per_s": a.reads / dt, "gbp_per_s": nb / dt / 1e9, "packed_fraction": ix.last_counter("host_packed_bases") / nb,
           "h2d_gb": ix.last_counter("h2d_bytes") / 1e9, "identical": bool(hv.tobytes() == want.tobytes()), "h2d_ms": ix.last_ms("h2d")}
    out.append(rec); print(json.dumps(rec), flush=True)
print(json.dumps({"host_cpus": os.cpu_count(), "reads": a.reads, "bases": int(nb), "sweep": out}))
