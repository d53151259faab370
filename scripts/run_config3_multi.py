#!/usr/bin/env python
"""BASELINE.json configs[2] at N GPUs (torchrun, one process per GPU): synthetic CHM13-sized genome (3.1 Gbp, 24
contigs), index built PARTITIONED BY REFERENCE CHUNK (rank r scans base range r of every contig), minimizer stores
all-gathered over NCCL, every rank freezes the union; 2,000,000 x 24 kb reads sharded in contiguous blocks, no
collective while mapping.  Rank 0 prints one JSON line; --check N compares N reads of rank 0 and all index counts
with the CPU oracle.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_config3_multi.py
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads-total", type=int, default=2000000)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--check", type=int, default=5000)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        w = torch.zeros(1, device=dev); dist.all_reduce(w); torch.cuda.synchronize()
    from mapquik_b200 import Index, Params, sim, shard, capi, HIT_DTYPE
    L = capi.lib()
    p = Params()
    tot = 3.1e9 / a.scale
    lens = [int(tot * x / sum(sim.CHM13_PROPS)) for x in sim.CHM13_PROPS]
    g, go, names = sim.genome(3, lens, sat_frac=0.06, segdup_frac=0.05)
    lo, hi = shard.read_shard(a.reads_total, rank, world)
    rb, ro, _, tr = sim.reads(3, g, go, hi - lo, 24000, 3000, 1000, 0.005, first=lo, with_names=False)

    def barrier():
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    # ---- index: partitioned scan + all-gather + replicated freeze --------------------------------------
    ix = Index(p, device=local)
    barrier(); t0 = time.perf_counter()
    for r in range(len(names)):
        seq = g[int(go[r]):int(go[r + 1])]
        s, own, data = shard.segment_for_rank(seq, rank, world, p.l)
        ix.add_segment(r, names[r], len(seq), s, own, data)
    t_scan = time.perf_counter() - t0
    if world > 1:
        d_pos, d_hash, n, directory = ix.store_export()

        class CAI:
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
        counts = [None] * world
        dist.all_gather_object(counts, (int(n), directory.tolist()))
        nmax = max(c[0] for c in counts)
        pos_in = torch.zeros(nmax * 4, dtype=torch.uint8, device=dev); hash_in = torch.zeros(nmax * 8, dtype=torch.uint8, device=dev)
        if n:
            pos_in[:n * 4] = torch.as_tensor(CAI(d_pos, n * 4), device=dev)
            hash_in[:n * 8] = torch.as_tensor(CAI(d_hash, n * 8), device=dev)
        pos_all = torch.empty(world * nmax * 4, dtype=torch.uint8, device=dev)
        hash_all = torch.empty(world * nmax * 8, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(); tg = time.perf_counter()
        dist.all_gather_into_tensor(pos_all, pos_in); dist.all_gather_into_tensor(hash_all, hash_in)
        torch.cuda.synchronize(); t_gather = time.perf_counter() - tg
        pos_m = torch.cat([pos_all[r * nmax * 4: r * nmax * 4 + counts[r][0] * 4] for r in range(world)]).contiguous()
        hash_m = torch.cat([hash_all[r * nmax * 8: r * nmax * 8 + counts[r][0] * 8] for r in range(world)]).contiguous()
        torch.cuda.synchronize()
        dirs = np.array([d for c in counts for d in c[1]], dtype=np.uint64).reshape(-1, 3)
        ix.store_import(pos_m.data_ptr(), hash_m.data_ptr(), sum(c[0] for c in counts), dirs)
        gathered_bytes = 12 * sum(c[0] for c in counts)
        del pos_all, hash_all, pos_in, hash_in
    else:
        t_gather, gathered_bytes = 0.0, 0
    n_unique = ix.freeze()
    barrier(); t_index = time.perf_counter() - t0

    # ---- mapping: my shard, pinned host buffers -----------------------------------------------------------
    def pin(arr):
        ptr = L.mq_host_alloc(arr.nbytes + 64)
        v = np.frombuffer((C.c_uint8 * arr.nbytes).from_address(ptr), dtype=np.uint8).view(arr.dtype)
        v[:] = arr
        return v
    prb, pro = pin(rb), pin(ro)
    hits = np.zeros(hi - lo, HIT_DTYPE)
    ix.map_batch(prb[:int(pro[2000])], pro[:2001])            # warm-up (allocations)
    barrier(); t0 = time.perf_counter()
    ix.map_batch(prb, pro, out=hits)
    barrier(); t_map = time.perf_counter() - t0
    ok = (hits["mapped"] == 1) & (hits["ref_idx"] == tr["contig"]) & (hits["rc"] == tr["strand"]) & \
        (np.minimum(hits["r_end"], tr["start"] + tr["len"]).astype(np.int64) -
         np.maximum(hits["r_start"], tr["start"]).astype(np.int64) > 0.1 * tr["len"])
    stats = np.array([float(hits["mapped"].sum()), float((hits["mapq"] == 60).sum()), float(ok.sum()),
                      float(((hits["mapq"] == 60) & ~ok).sum()), float(ro[-1]), t_map, t_index, t_scan], dtype=np.float64)
    if world > 1:
        st = torch.tensor(stats, device=dev)
        mx = st.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(st, op=dist.ReduceOp.SUM)
        stats = st.cpu().numpy(); mxs = mx.cpu().numpy()
        t_map, t_index, t_scan = float(mxs[5]), float(mxs[6]), float(mxs[7])
    if rank == 0:
        out = {"config": 3, "n_gpus": world, "genome_bp": int(go[-1]), "reads_total": a.reads_total, "read_bp_total": float(stats[4]),
               "index_build_s": t_index, "index_scan_s": t_scan, "index_allgather_s": t_gather, "index_allgather_bytes": gathered_bytes,
               "n_unique": int(n_unique), "n_keys": int(ix.n_keys), "map_s": t_map, "reads_per_s": a.reads_total / t_map,
               "gbp_per_s": float(stats[4]) / t_map / 1e9, "mapped": int(stats[0]), "q60": int(stats[1]), "correct": int(stats[2]),
               "wrong_q60": int(stats[3]), "map_stage_ms_rank0": {s: ix.last_ms(s) for s in ("h2d", "scan", "gather", "probe", "chain", "d2h")}}
        if a.check:
            from oracle import pyoracle as O
            th = os.cpu_count() or 8
            oix = O.Index(O.params(), 48000000)
            t0 = time.perf_counter(); onb = oix.add_batch(names, g, go, threads=th); t_o = time.perf_counter() - t0
            n = min(a.check, hi - lo)
            oh = oix.map_batch(rb[:int(ro[n])], ro[:n + 1], threads=th)
            out.update(oracle_index_s=t_o, parity_n_unique=bool(oix.count() == n_unique), parity_n_keys=bool(oix.slots() == ix.n_keys),
                       parity_nb_mers=bool(np.array_equal(onb, ix.nb_mers())), parity_hits=bool(oh.tobytes() == hits[:n].tobytes()),
                       checked_reads=n)
        print(json.dumps(out), flush=True)
    barrier()
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
