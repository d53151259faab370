#!/usr/bin/env python
"""bench.py -- headline benchmark of the mapquik seeding->chaining hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 3|2] [--reads R]

Default workload = BASELINE.json configs[2], the config the metric is quoted on: synthetic
CHM13-sized genome (3.1 Gbp, 24 contigs, 6 % satellites, 5 % segmental duplications, seed 3), 2,000,000 synthetic
HiFi-like reads, length N(24 kb, 3 kb), 99.5 % identity, seed 3 -> 48 Gbp per step.  STRONG scaling: the 2,000,000 reads
are sharded over the ranks (contiguous blocks, no data-path collective), the index is built partitioned by reference
base range, the per-rank minimizer stores are exchanged over NCCL (in place, exact sizes) and every rank freezes the
whole index.  A "step" is one pass of the whole hot path (S1 scan -> k-min-mers -> probe -> Match -> chain -> mq_hit)
over all reads.  `--config 2` is round 1's workload (E. coli-sized genome, 100,000 x 10 kb reads per rank, weak scaling),
`--config 4` the maize-like repetitive genome (2.2 Gbp, 85 % repeat families, 1,000,000 x 24 kb reads, strong scaling).

Printed JSON (one line, rank 0), see the contract in the task statement:
  value          reads/s with the reads resident in HBM in the library's packed input format (mq_map_batch_packed_device)
  value_ascii    the same with ASCII reads resident (mq_map_batch_device)
  e2e            through mq_map_batch with pinned HOST ASCII buffers, H2D of the sequences and D2H of the hits inside; the
                 library may use the rank's host threads (mq_set_host_threads, the reference's --threads) to pack part
                 of the batch on the fly, so fewer than 1 byte per base cross the link (bytes counted by the library);
                 on or off is calibrated on untimed steps (on a host that feeds several GPUs the packers can cost more
                 DRAM bandwidth than they save on the link)
  e2e_ascii_link_only  the same call with host threads off: every base crosses the link as one byte
  e2e_prepacked  through mq_map_batch_packed with pinned host buffers the caller's parser packed (what the CLI does)
  e2e_packed     ASCII in host memory, mq_pack on all host threads of the rank + mq_map_batch_packed, pipelined by chunk,
                 all of it inside the timed region
  roofline       the dominant kernel k_scan_minimizers<hpc, packed> of the `value` region
  parity         untimed: hits of the first 100,000 reads and every index count against the CPU oracle; all GPU paths
                 byte-identical on all reads
  cpu_baseline   the CPU oracle (a port, not the upstream Rust binary) on the box's host cores (N = 1 only)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    3: dict(workload="synthetic_chm13_3.1Gbp_24contigs_x_2M_hifi_reads_24kb_99.5pct (BASELINE configs[2])", seed=3, genome_bp=3.1e9,
            n_reads=2000000, mean=24000.0, sd=3000.0, min_len=1000, err=0.005, scaling="strong", sat=0.06, segdup=0.05),
    2: dict(workload="ecoli_4.64Mbp_x_100k_hifi_reads_10kb_99.5pct (BASELINE configs[1])", seed=2, genome_bp=4641652,
            n_reads=100000, mean=10000.0, sd=1500.0, min_len=1000, err=0.005, scaling="weak", sat=0.0, segdup=0.0),
    4: dict(workload="synthetic_maize_like_2.2Gbp_10contigs_85pct_repeat_families_x_1M_hifi_reads_24kb_99.5pct (BASELINE configs[3])", seed=4,
            genome_bp=2.2e9, n_reads=1000000, mean=24000.0, sd=3000.0, min_len=1000, err=0.005, scaling="strong", sat=0.0, segdup=0.0,
            repeat=0.85, families=300, contigs=10),
}
CPU_SAMPLE_READS = {3: 400000, 2: 100000, 4: 400000}       # reads per CPU-arm step (bounded sample of the workload)
PACK_CHUNK_BASES = 1 << 30


def host_threads():
    """all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def make_genome(cfg):
    from mapquik_b200 import sim
    if cfg.get("repeat"):
        return sim.genome(cfg["seed"], [int(cfg["genome_bp"] / cfg["contigs"])] * cfg["contigs"], repeat_frac=cfg["repeat"], n_families=cfg["families"])
    if cfg["genome_bp"] > 1e8:
        lens = [int(cfg["genome_bp"] * x / sum(sim.CHM13_PROPS)) for x in sim.CHM13_PROPS]
        return sim.genome(cfg["seed"], lens, sat_frac=cfg["sat"], segdup_frac=cfg["segdup"])
    return sim.genome(cfg["seed"], [int(cfg["genome_bp"])], names=["chr000913"])


def make_reads(cfg, g, go, first, count, out=None):
    from mapquik_b200 import sim
    rb, ro, _, tr = sim.reads(cfg["seed"], g, go, count, cfg["mean"], cfg["sd"], cfg["min_len"], cfg["err"], first=first,
                              with_names=False, out=out)
    return rb, ro, tr


def read_lengths_total(cfg, go, first, count):
    """bases of reads [first, first+count) without generating them (sizes the pinned buffer)"""
    from mapquik_b200 import sim
    L = sim.lib()
    rc = sim.ReadCfg(cfg["seed"], float(cfg["mean"]), float(cfg["sd"]), int(cfg["min_len"]), float(cfg["err"]))
    goffs = np.ascontiguousarray(go, dtype=np.uint64)
    tc = np.zeros(count, np.uint32); ts = np.zeros(count, np.uint64); tl = np.zeros(count, np.uint64)
    st = np.zeros(count, np.uint8); ol = np.zeros(count, np.uint64)
    L.mqsim_reads_plan(C.byref(rc), goffs.ctypes.data, goffs.size - 1, first, count, tc.ctypes.data, ts.ctypes.data, tl.ctypes.data,
                       st.ctypes.data, ol.ctypes.data)
    return int(ol.sum())


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # under load = samples in the upper half of the observed power range (the sampler also sees set-up phases)
            pw = np.array(power); lo, hi = pw.min(), pw.max()
            busy = np.array(sm)[pw >= lo + 0.5 * (hi - lo)] if hi > lo else np.array(sm)
            out.update(sm_mhz=float(np.median(busy)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(power)))
        return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def scan_profile():
    """numbers of the committed ncu captures of the scan kernel (profiles/scan_traffic.json): DRAM bytes and
    warp-instructions per base do not depend on timing, so they are carried over from the capture"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
    except Exception:
        return {}


def pinned_array(L, nbytes, dtype=np.uint8):
    ptr = L.mq_host_alloc(max(nbytes, 16))
    if not ptr:
        raise RuntimeError("mq_host_alloc failed")
    buf = (C.c_uint8 * max(nbytes, 16)).from_address(ptr)
    return np.frombuffer(buf, dtype=np.uint8, count=nbytes).view(dtype), ptr


# ---- the CPU arm ------------------------------------------------------------------------------------------------
def oracle_index(cfg, g, go, names, threads):
    from oracle import pyoracle as O
    ix = O.Index(O.params(), 48000000 if cfg["genome_bp"] > 1e8 else 1 << 17)
    t0 = time.perf_counter()
    nb = ix.add_batch(names, g, go, threads=threads)
    n_unique = ix.count()
    return ix, nb, n_unique, time.perf_counter() - t0


def run_reference(args, cfg, rank):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port: upstream cannot be built
    offline) with all host threads, on a bounded sample of the same workload per step.  Rank 0 alone works."""
    if rank != 0:
        return
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    g, go, names = make_genome(cfg)
    n_sample = min(args.reads or CPU_SAMPLE_READS[args.config], cfg["n_reads"])
    rb, ro, _ = make_reads(cfg, g, go, 0, n_sample)
    ix, _, n_unique, t_index = oracle_index(cfg, g, go, names, threads)
    for _ in range(max(args.warmup, 1)):
        ix.map_batch(rb, ro, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hits = ix.map_batch(rb, ro, threads=threads)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    rps = n_sample / dt
    line = {
        "impl": "reference", "metric": "reads/sec mapped (seeding->chaining hot path)", "value": rps, "unit": "reads/s",
        "n_gpus": args.gpus, "ranks_working": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": cfg["workload"], "reads_total": cfg["n_reads"], "k": 5, "l": 31, "density": 0.01, "hpc": True},
        "reads_per_step": n_sample,
        "gbp_per_s": float(ro[-1]) / dt / 1e9, "index_build_s": t_index,
        "cpu_baseline": {"value": rps, "unit": "reads/s", "cores": threads, "kind": "port",
                         "sample": f"first {n_sample} of the workload's {cfg['n_reads']} reads per step ({float(ro[-1]) / 1e9:.2f} Gbp), all host "
                                   "threads (OpenMP over reads); CPU restatement of mapquik (oracle/), not the upstream Rust binary "
                                   "(no Rust toolchain offline)"},
        "e2e": {"value": rps, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mapped_reads": int(hits["mapped"].sum()), "n_unique_kminmers": int(n_unique),
        "upstream": upstream_probe(),
    }
    print(json.dumps(line), flush=True)


def upstream_probe():
    """SURVEY 8c / BASELINE.md section 3: if a real `mapquik` binary is reachable (baseline/_ref or $PATH), say so; the PAF
    diff itself lives in scripts/upstream_probe.py (it needs FASTA files on disk).  Its absence is reported, not hidden."""
    import shutil
    cands = [os.path.join(ROOT, "baseline", "_ref", "bin", "mapquik"), os.path.join(ROOT, "baseline", "_ref", "mapquik"), shutil.which("mapquik")]
    for c in cands:
        if c and os.path.isfile(c) and os.access(c, os.X_OK) and os.path.realpath(c) != os.path.realpath(os.path.join(ROOT, "host", "mapquik")):
            return {"found": True, "path": c, "paf_identical": None, "note": "run scripts/upstream_probe.py for the PAF diff"}
    return {"found": False, "paf_identical": None}


# ---- multi-rank index build: partitioned scan + in-place NCCL exchange + replicated freeze ---------------------------
def build_index_ranks(ix, g, go, names, p, rank, world, dist, torch, device, L):
    from mapquik_b200 import shard
    t0 = time.perf_counter()
    for r in range(len(names)):
        seq = g[int(go[r]):int(go[r + 1])]
        s, own, data = shard.segment_for_rank(seq, rank, world, p.l)
        ix.add_segment(r, names[r], len(seq), s, own, data if len(data) else np.zeros(1, np.uint8))
    L.mq_sync(ix.handle)
    t_scan = time.perf_counter() - t0
    n, nseg = ix.store_info()
    _, _, _, directory = ix.store_export()
    counts = [None] * world
    dist.all_gather_object(counts, (int(n), directory.tolist()))
    total = sum(c[0] for c in counts)
    dp, dh = C.c_void_p(), C.c_void_p()
    assert L.mq_store_reserve(ix.handle, total, C.byref(dp), C.byref(dh)) == 0

    class CAI:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    pos_t = torch.as_tensor(CAI(dp.value, max(total, 1) * 4), device=device)
    hash_t = torch.as_tensor(CAI(dh.value, max(total, 1) * 8), device=device)
    # layout on this rank: [own share][the others, in rank order]; the directory follows the same order and
    # mq_index_freeze sorts segments by (reference, start), so every rank may keep its own physical order
    off, order = n, [rank]
    slot = {rank: 0}
    for q in range(world):
        if q != rank:
            slot[q] = off; off += counts[q][0]; order.append(q)
    torch.cuda.synchronize(); dist.barrier()
    tg = time.perf_counter()
    # one in-place broadcast per rank and array (exact sizes, no padding, no concatenation copies), all in flight at once.
    # (One uneven all_gather per array was tried instead: 6.0 ms against 2.6 ms at two GPUs.)
    works = []
    for q in range(world):
        cnt = counts[q][0]
        if cnt == 0:
            continue
        works.append(dist.broadcast(pos_t[slot[q] * 4:(slot[q] + cnt) * 4], src=q, async_op=True))
        works.append(dist.broadcast(hash_t[slot[q] * 8:(slot[q] + cnt) * 8], src=q, async_op=True))
    for w in works:
        w.wait()
    torch.cuda.synchronize()
    t_exchange = time.perf_counter() - tg
    dirs = np.array([d for q in order for d in counts[q][1]], dtype=np.uint64).reshape(-1, 3)
    assert L.mq_store_commit(ix.handle, total, dirs.ctypes.data, dirs.shape[0]) == 0
    tf = time.perf_counter()
    n_unique = ix.freeze()
    t_freeze = time.perf_counter() - tf
    return n_unique, {"scan_s": t_scan, "exchange_s": t_exchange, "exchange_bytes": int(12 * total), "freeze_s": t_freeze}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 2, 4])
    ap.add_argument("--reads", type=int, default=0, help="total reads per step (config 3) / per rank (config 2); default: the named workload")
    ap.add_argument("--check", type=int, default=100000, help="reads compared with the CPU oracle inside the run (untimed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-packed", action="store_true", help="skip the pack-inside-the-timed-region variant")
    ap.add_argument("--value-only", action="store_true", help="profiling runs (ncu launch lists): only the `value` region, no JSON line")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return
    args.warmup = max(args.warmup, 3)
    threads = max(1, host_threads() // world)
    os.environ["OMP_NUM_THREADS"] = str(threads)      # simulator / oracle threads of this rank (torchrun sets 1)

    dist = torch = device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        device = torch.device(f"cuda:{local_rank}")
        dist.init_process_group("nccl", device_id=device)
        warm = torch.zeros(1, device=device)
        dist.all_reduce(warm)                 # communicator set-up is not part of the index build
        torch.cuda.synchronize()

    from mapquik_b200 import HIT_DTYPE, Index, PackedSeqs, Params, capi
    L = capi.lib()
    p = Params()

    # ---- workload ---------------------------------------------------------------------------------------------------
    t_gen = time.perf_counter()
    g0, go, names = make_genome(cfg)
    g, pg = pinned_array(L, g0.size + 64)                      # the reference sits in pinned host memory, like the CLI's parser slots
    g = g[:g0.size]; g[:] = g0; del g0
    if cfg["scaling"] == "strong":
        total_reads = args.reads or cfg["n_reads"]
        lo, hi = total_reads * rank // world, total_reads * (rank + 1) // world
    else:
        per = args.reads or cfg["n_reads"]
        total_reads = per * world
        lo, hi = rank * per, (rank + 1) * per
    n_reads = hi - lo
    n_bases = read_lengths_total(cfg, go, lo, n_reads)
    h_seqs, p1 = pinned_array(L, n_bases + 64)                 # the reads are simulated straight into pinned memory
    rb, ro, truth = make_reads(cfg, g, go, lo, n_reads, out=h_seqs)
    assert int(ro[-1]) == n_bases
    h_offs, p2 = pinned_array(L, (n_reads + 1) * 8, np.uint64); h_offs[:] = ro
    h_hits, p3 = pinned_array(L, n_reads * 48); h_hits_v = h_hits.view(HIT_DTYPE)
    pk = PackedSeqs(rb, n_threads=threads, pinned=True)         # what a packing parser would have produced
    t_gen = time.perf_counter() - t_gen

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # one-off start-up costs of the process (CUDA module loading, first pinned / device allocations) are not the index build:
    # a toy index + a few reads first
    from mapquik_b200 import sim as _sim
    wg, wgo, wnames = _sim.genome(1, [200000]); wrb, wro, _, _ = _sim.reads(1, wg, wgo, 50, 5000, 1000)
    wix = Index(p, device=local_rank); wix.add_batch(wnames, wg, wgo); wix.freeze(); wix.map_batch(wrb, wro); wix.close()

    # ---- index build (timed end to end from host memory) ---------------------------------------------------------
    ix = Index(p, device=local_rank)
    h = ix.handle
    barrier()
    t0 = time.perf_counter()
    if world > 1:
        n_unique, ib = build_index_ranks(ix, g, go, names, p, rank, world, dist, torch, device, L)
    else:
        ta = time.perf_counter()
        ix.add_batch(names, g, go)
        tb = time.perf_counter()
        n_unique = ix.freeze()
        ib = {"scan_s": tb - ta, "exchange_s": 0.0, "exchange_bytes": 0, "freeze_s": time.perf_counter() - tb}
    barrier()
    index_build_s = time.perf_counter() - t0
    # the same build from the packed reference (what the CLI's packing parser hands over); second build: allocations are warm
    index_build_packed_s = None
    if world == 1:
        pgen = PackedSeqs(g, n_threads=threads, pinned=True)
        for rep_ in range(2):
            ix2 = Index(p, device=local_rank)
            t1 = time.perf_counter()
            ix2.add_batch_packed(names, pgen, go)
            nu2 = ix2.freeze()
            index_build_packed_s = time.perf_counter() - t1
            assert nu2 == n_unique and ix2.n_keys == ix.n_keys, "packed index build differs"
            ix2.close()
        pgen.close()

    # ---- device-resident inputs ---------------------------------------------------------------------------------
    d_hits = L.mq_dev_alloc(h, max(n_reads * 48, 16))
    d_seqs = L.mq_dev_alloc(h, n_bases + 512)
    d_w = L.mq_dev_alloc(h, pk.words.nbytes); d_f = L.mq_dev_alloc(h, pk.flags.nbytes); d_e = L.mq_dev_alloc(h, max(pk.exc.nbytes, 16))
    assert d_hits and d_seqs and d_w and d_f and d_e, "device allocation failed"
    L.mq_dev_memset(h, d_seqs, 0, n_bases + 512)
    assert L.mq_dev_upload(h, d_seqs, h_seqs.ctypes.data, n_bases) == 0
    assert L.mq_dev_upload(h, d_w, pk.words.ctypes.data, pk.words.nbytes) == 0
    assert L.mq_dev_upload(h, d_f, pk.flags.ctypes.data, pk.flags.nbytes) == 0
    if pk.exc.size:
        assert L.mq_dev_upload(h, d_e, pk.exc.ctypes.data, pk.exc.nbytes) == 0

    def step_packed_dev():
        ix.map_batch_packed_device(d_w, d_f, d_e, pk.exc.size, pk.n_bases, h_offs, d_hits)

    def step_ascii_dev():
        ix.map_batch_device(d_seqs, h_offs, d_hits)

    def timed(fn, steps, warmup, device_events=True):
        """W untimed steps, barrier, K timed steps bracketed by CUDA events on the ctx stream (and the wall clock), barrier"""
        for _ in range(warmup):
            fn()
        L.mq_sync(h)
        barrier()
        if device_events:
            L.mq_region_begin(h)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dev_ms = L.mq_region_end_ms(h) if device_events else 0.0
        L.mq_sync(h)
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        return max(dev_ms, wall_ms) if device_events else wall_ms

    clocks = ClockSampler(local_rank)     # sampled from the warm-up through all timed regions

    # value: packed reads resident in HBM
    for _ in range(args.warmup):
        step_packed_dev()
    L.mq_sync(h)
    launches0 = ix.launch_count(); sk0 = L.mq_scan_kernel_launches(h); L.mq_minimizer_count(h, 1)
    scan_ms0 = ix.total_ms("scan_kernel")
    t_dev = timed(step_packed_dev, args.steps, 0)
    scan_ms = ix.total_ms("scan_kernel") - scan_ms0
    launches = ix.launch_count() - launches0
    scan_launches = L.mq_scan_kernel_launches(h) - sk0
    n_min = L.mq_minimizer_count(h, 1) // max(args.steps, 1)
    stage_ms = {s: ix.last_ms(s) for s in ("scan", "scan_kernel", "gather", "probe", "chain")}
    hits = np.zeros(n_reads, HIT_DTYPE)
    assert L.mq_dev_download(h, hits.ctypes.data, d_hits, n_reads * 48) == 0

    if args.value_only:
        print(f"[value-only] {total_reads * args.steps / (t_dev / 1e3):.0f} reads/s, scan kernel {scan_ms / max(args.steps, 1):.3f} ms per step "
              f"of {t_dev / max(args.steps, 1):.3f}, {launches} launches, stages {stage_ms}", file=sys.stderr)
        clocks.stop()
        return

    # value_ascii: ASCII reads resident in HBM
    L.mq_dev_memset(h, d_hits, 0, n_reads * 48)
    for _ in range(args.warmup):
        step_ascii_dev()
    L.mq_sync(h)
    scan_ms0 = ix.total_ms("scan_kernel"); sk0 = L.mq_scan_kernel_launches(h)      # after the warm-up: the timed launches only
    t_dev_ascii = timed(step_ascii_dev, args.steps, 0)
    scan_ms_ascii = ix.total_ms("scan_kernel") - scan_ms0
    scan_launches_ascii = L.mq_scan_kernel_launches(h) - sk0
    hits_b = np.zeros(n_reads, HIT_DTYPE)
    assert L.mq_dev_download(h, hits_b.ctypes.data, d_hits, n_reads * 48) == 0
    path_diff = {"ascii_resident": int((hits_b != hits).sum())}

    # e2e_link_only: mq_map_batch on pinned HOST ASCII buffers, every base crosses the PCIe link as one byte
    h_hits[:] = 0
    ix.set_host_threads(0)
    t_e2e_link = timed(lambda: ix.map_batch(h_seqs[:n_bases], h_offs, out=h_hits_v), args.steps, args.warmup, device_events=False)
    e2e_link_stage = {s: ix.last_ms(s) for s in ("h2d", "scan", "gather", "probe", "chain", "d2h")}
    h2d_link = ix.last_counter("h2d_bytes")
    path_diff["e2e_ascii_link_only"] = int((h_hits_v != hits).sum())

    # e2e: the same call on the same buffers with this rank's host threads at the library's disposal (the reference's
    # --threads): sub-batches packed on the fly from the back of the batch while ASCII ones cross the link from the front.
    # Whether that pays depends on the host (with several GPUs on one host the packers compete with the DMA engines for
    # host DRAM), so the setting is calibrated like a user would: two untimed steps either way, the faster one (max over
    # ranks, the same decision on every rank) is what the timed region then runs with.
    def calib(nthr):
        ix.set_host_threads(nthr)
        t = timed(lambda: ix.map_batch(h_seqs[:n_bases], h_offs, out=h_hits_v), 2, 1, device_events=False)
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); t = float(tt.item())
        return t
    t_cal_off, t_cal_on = calib(0), calib(threads)
    e2e_threads = threads if t_cal_on < t_cal_off else 0
    h_hits[:] = 0
    ix.set_host_threads(e2e_threads)
    if e2e_threads:
        t_e2e = timed(lambda: ix.map_batch(h_seqs[:n_bases], h_offs, out=h_hits_v), args.steps, args.warmup, device_events=False)
        e2e_stage = {s: ix.last_ms(s) for s in ("h2d", "scan", "gather", "probe", "chain", "d2h")}
        h2d_e2e = ix.last_counter("h2d_bytes"); packed_bases_e2e = ix.last_counter("host_packed_bases")
        path_diff["e2e_ascii"] = int((h_hits_v != hits).sum())
    else:
        t_e2e, e2e_stage, h2d_e2e, packed_bases_e2e = t_e2e_link, e2e_link_stage, h2d_link, 0
        path_diff["e2e_ascii"] = path_diff["e2e_ascii_link_only"]
    ix.set_host_threads(0)

    # e2e_prepacked: pinned host buffers in the packed format
    h_hits[:] = 0
    t_e2e_pre = timed(lambda: ix.map_batch_packed(pk, h_offs, out=h_hits_v), args.steps, args.warmup, device_events=False)
    e2e_pre_stage = {s: ix.last_ms(s) for s in ("h2d", "scan", "gather", "probe", "chain", "d2h")}
    path_diff["e2e_prepacked"] = int((h_hits_v != hits).sum())

    # e2e_packed: ASCII in host memory -> mq_pack on this rank's host threads -> mq_map_batch_packed, chunk-pipelined
    t_e2e_pack = None; pack_gbs = None
    if not args.no_e2e_packed:
        cuts = [0]
        while cuts[-1] < n_reads:
            j = int(np.searchsorted(ro, ro[cuts[-1]] + PACK_CHUNK_BASES, side="right")) - 1
            cuts.append(min(n_reads, max(j, cuts[-1] + 1)))
        nch = len(cuts) - 1
        maxb = max(int(ro[cuts[i + 1]] - ro[cuts[i]]) for i in range(nch))
        bufs = []
        for _ in range(2):
            w, pw = pinned_array(L, int(L.mq_packed_words(maxb)) * 4, np.uint32)
            f, pf = pinned_array(L, int(L.mq_packed_flag_words(maxb)) * 4, np.uint32)
            e, pe = pinned_array(L, 16 * 65536, capi.EXC_DTYPE)
            bufs.append((w, f, e, (pw, pf, pe)))

        def pack_chunk(i, b):
            w, f, e, _ = bufs[b]
            s0, s1 = int(ro[cuts[i]]), int(ro[cuts[i + 1]])
            k = C.c_uint64()
            rc = L.mq_pack(h_seqs[s0:].ctypes.data, s1 - s0, w.ctypes.data, f.ctypes.data, e.ctypes.data, e.size, C.byref(k), threads, 0)
            assert rc == 0, "mq_pack"
            return capi.Packed(w.ctypes.data, f.ctypes.data, e.ctypes.data if k.value else None, k.value, s1 - s0)

        def step_pack_inside():
            nxt = [pack_chunk(0, 0)]
            for i in range(nch):
                cur = nxt[0]
                th = None
                if i + 1 < nch:
                    th = threading.Thread(target=lambda: nxt.__setitem__(0, pack_chunk(i + 1, (i + 1) & 1)))
                    th.start()              # ctypes releases the GIL: the next chunk is packed while this one is mapped
                offs = (ro[cuts[i]:cuts[i + 1] + 1] - ro[cuts[i]]).astype(np.uint64)
                rc = L.mq_map_batch_packed(h, C.byref(cur), offs.ctypes.data, offs.size - 1, h_hits_v[cuts[i]:].ctypes.data)
                assert rc == 0, L.mq_last_error(h)
                if th:
                    th.join()
        h_hits[:] = 0
        ps = min(args.steps, 5)
        t_e2e_pack = timed(step_pack_inside, ps, 1, device_events=False) * (args.steps / ps)
        path_diff["e2e_packed"] = int((h_hits_v != hits).sum())
        t0 = time.perf_counter(); pack_chunk(0, 0); pack_gbs = int(ro[cuts[1]] - ro[cuts[0]]) / (time.perf_counter() - t0) / 1e9
        for _, _, _, ptrs in bufs:
            for q in ptrs:
                L.mq_host_free(q)
    clk = clocks.stop()

    # ---- parity against the CPU oracle (untimed) + CPU baseline -----------------------------------------------------
    paths_identical = not any(path_diff.values())
    parity = {"paths_identical": bool(paths_identical), "reads_differing_from_packed_resident": path_diff}
    cpu_line = None
    if rank == 0 and (args.check or not args.no_cpu_baseline):
        all_threads = host_threads() if world == 1 else threads
        oix, onb, o_unique, t_oindex = oracle_index(cfg, g, go, names, all_threads)
        nchk = min(args.check, n_reads)
        oh = oix.map_batch(rb[:int(ro[nchk])], ro[:nchk + 1], threads=all_threads)
        bad = np.nonzero(oh != hits[:nchk])[0]
        if bad.size:
            print(f"[parity] {bad.size} of {nchk} reads differ from the oracle; first: read {bad[0]} len {int(ro[bad[0] + 1] - ro[bad[0]])} "
                  f"gpu {hits[bad[0]]} oracle {oh[bad[0]]} ascii-resident {hits_b[bad[0]]}", file=sys.stderr)
        parity.update(checked_reads=int(nchk), hits_identical=bool(oh.tobytes() == hits[:nchk].tobytes()), reads_differing_from_oracle=int(bad.size),
                      ascii_resident_differing_from_oracle=int((oh != hits_b[:nchk]).sum()),
                      index_identical=bool(o_unique == n_unique and oix.slots() == ix.n_keys and np.array_equal(onb, ix.nb_mers())),
                      oracle_index_s=t_oindex)
        if world == 1 and not args.no_cpu_baseline:
            ns = min(CPU_SAMPLE_READS[args.config], n_reads)
            srb, sro = rb[:int(ro[ns])], ro[:ns + 1]
            oix.map_batch(srb, sro, threads=all_threads)
            t0 = time.perf_counter()
            for _ in range(2):
                oix.map_batch(srb, sro, threads=all_threads)
            dt = (time.perf_counter() - t0) / 2
            cpu_line = {"value": ns / dt, "unit": "reads/s", "cores": all_threads, "kind": "port",
                        "sample": f"first {ns} reads ({float(sro[-1]) / 1e9:.2f} Gbp) x 2 timed passes after 1 warm-up, OpenMP over reads; "
                                  "CPU restatement (oracle/), not the upstream Rust binary",
                        "gbp_per_s": float(sro[-1]) / dt / 1e9, "index_build_s": t_oindex}
    ok = (hits["mapped"] == 1) & (hits["ref_idx"] == truth["contig"]) & (hits["rc"] == truth["strand"]) & \
        (np.minimum(hits["r_end"], truth["start"] + truth["len"]).astype(np.int64) -
         np.maximum(hits["r_start"], truth["start"]).astype(np.int64) > 0.1 * truth["len"])

    # ---- max / sum over ranks ----------------------------------------------------------------------------------------
    times = [t_dev, t_dev_ascii, t_e2e, t_e2e_pre, t_e2e_pack or 0.0, index_build_s, ib["scan_s"], ib["exchange_s"], ib["freeze_s"], t_e2e_link]
    sums = [float(hits["mapped"].sum()), float(n_bases), float(ok.sum()), float(((hits["mapq"] == 60) & ~ok).sum()), float(paths_identical),
            float(launches), float(n_min), float(scan_ms), float(scan_launches), float(scan_ms_ascii), float(scan_launches_ascii),
            float(h2d_e2e), float(packed_bases_e2e), float(h2d_link)]
    if world > 1:
        tt = torch.tensor(times, dtype=torch.float64, device=device); dist.all_reduce(tt, op=dist.ReduceOp.MAX); times = tt.tolist()
        ss = torch.tensor(sums, dtype=torch.float64, device=device); dist.all_reduce(ss, op=dist.ReduceOp.SUM); sums = ss.tolist()
    t_dev, t_dev_ascii, t_e2e, t_e2e_pre, t_e2e_pack_m, index_build_s, ib_scan, ib_exch, ib_freeze, t_e2e_link = times
    (mapped_total, bases_total, ok_total, wrong_q60, paths_ok, launches_t, n_min_t, scan_ms_t, scan_l_t, scan_ms_a, scan_l_a,
     h2d_e2e_t, packed_bases_t, h2d_link_t) = sums

    if rank == 0:
        peak, peak_src = peaks()
        prof = scan_profile()
        K = args.steps
        rps = lambda ms: total_reads * K / (ms / 1e3)
        gbps = lambda ms: bases_total * K / (ms / 1e3) / 1e9
        # roofline of the dominant kernel, summed over every launch of every rank in the timed region:
        # algorithmic bytes = packed codes in (1/4 byte per base) + (pos u32, hash u64) per emitted minimizer out
        algo_bytes = (bases_total / 4 + 12 * n_min_t) * K
        achieved = algo_bytes / (scan_ms_t / 1e3) / 1e9 if scan_ms_t > 0 else 0.0     # per GPU: bytes of all launches / their summed duration
        algo_ascii = (bases_total + 12 * n_min_t) * K
        achieved_ascii = algo_ascii / (scan_ms_a / 1e3) / 1e9 if scan_ms_a > 0 else 0.0
        parity["paths_identical"] = bool(paths_ok == world)
        issue = None
        wi = prof.get("packed", {}).get("warp_instructions_per_base")
        if wi and clk.get("sm_mhz") and scan_ms_t > 0:
            a = wi * bases_total * K / (scan_ms_t / 1e3)
            pk_issue = prof.get("issue_peak_warp_instr_per_clk_per_sm", 4.0) * 148 * clk["sm_mhz"] * 1e6
            issue = {"achieved_warp_instr_per_s": a, "peak_warp_instr_per_s": pk_issue, "frac": a / pk_issue,
                     "thread_instructions_per_base": wi * 32, "source": "profiles/scan_traffic.json (ncu smsp__inst_executed.sum of the same kernel)"}
        line = {
            "metric": "reads/sec mapped (seeding->chaining hot path)", "value": rps(t_dev), "unit": "reads/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": t_dev / K,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": cfg["workload"], "reads_total": int(total_reads), "bases_total": int(bases_total), "k": p.k, "l": p.l,
                       "density": p.density, "hpc": True,
                       "l2": f"inputs ({bases_total / world / 4e9:.1f} GB of packed reads per GPU and step) larger than L2; no flush needed",
                       "index": "replicated per GPU; built by reference base range + in-place NCCL exchange at N>1",
                       "input_format": "value: 2-bit packed reads resident in HBM; value_ascii: ASCII reads resident"},
            "gbp_per_s": gbps(t_dev),
            "value_ascii": rps(t_dev_ascii), "gbp_per_s_ascii": gbps(t_dev_ascii),
            "index_build_s": index_build_s, "index_build_packed_s": index_build_packed_s,
            "index_build": {"scan_s": ib_scan, "exchange_s": ib_exch, "exchange_bytes": ib["exchange_bytes"], "freeze_s": ib_freeze,
                            "exchange_gb_per_s": (ib["exchange_bytes"] / ib_exch / 1e9) if ib_exch > 0 else None},
            "n_unique_kminmers": int(n_unique), "table_bytes": int(ix.table_bytes()),
            "mapped_fraction": mapped_total / total_reads, "correct_fraction": ok_total / total_reads, "wrong_q60": int(wrong_q60),
            "parity": parity,
            "e2e": {"value": rps(t_e2e), "unit": "reads/s", "h2d_bytes_per_step": int(h2d_e2e_t),
                    "d2h_bytes_per_step": int(total_reads * 48), "gbp_per_s": gbps(t_e2e), "stage_ms_last_step_rank0": e2e_stage,
                    "host_threads_per_rank": e2e_threads, "host_threads_available_per_rank": threads,
                    "calibration_ms_per_step": {"host_threads_off": t_cal_off / 2, "host_threads_on": t_cal_on / 2},
                    "bases_packed_on_host_fraction": packed_bases_t / max(bases_total, 1),
                    "h2d_bytes_per_base": h2d_e2e_t / max(bases_total, 1),
                    "input": "upper-cased ASCII in pinned host memory through mq_map_batch; mq_set_host_threads(host threads of the rank, or 0) "
                             "as calibrated on untimed steps: with threads, sub-batches are packed on the fly from the back of the batch while "
                             "ASCII sub-batches cross the link from the front; h2d bytes counted by the library (last step)"},
            "e2e_ascii_link_only": {"value": rps(t_e2e_link), "unit": "reads/s", "h2d_bytes_per_step": int(h2d_link_t),
                                    "d2h_bytes_per_step": int(total_reads * 48), "gbp_per_s": gbps(t_e2e_link), "stage_ms_last_step_rank0": e2e_link_stage,
                                    "input": "the same call with mq_set_host_threads(0): every base crosses PCIe as one byte"},
            "e2e_prepacked": {"value": rps(t_e2e_pre), "unit": "reads/s", "h2d_bytes_per_step": int(pk.nbytes * world + (total_reads + world) * 8),
                              "d2h_bytes_per_step": int(total_reads * 48), "gbp_per_s": gbps(t_e2e_pre), "stage_ms_last_step_rank0": e2e_pre_stage,
                              "h2d_bytes_per_base": (pk.nbytes + (n_reads + 1) * 8) / max(n_bases, 1),
                              "input": "2-bit packed reads in pinned host memory, packed by the caller's parser (mq_map_batch_packed)"},
            "e2e_packed": None if t_e2e_pack is None else {
                "value": rps(t_e2e_pack_m), "unit": "reads/s", "gbp_per_s": gbps(t_e2e_pack_m), "pack_threads_per_rank": threads,
                "pack_gb_per_s_rank0": pack_gbs,
                "input": "ASCII in host memory; mq_pack on the rank's host threads + mq_map_batch_packed, chunk-pipelined, all inside the timed region"},
            "gpu_launches": int(launches_t),
            "roofline": {"bound": "hbm", "kernel": "k_scan_minimizers<hpc=1,packed=1>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": (prof["packed"]["dram_bytes_per_base"] * bases_total * K / max(scan_l_t, 1))
                         if prof.get("packed", {}).get("dram_bytes_per_base") else None,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes / max(scan_l_t, 1),
                         "avg_launch_ms": scan_ms_t / max(scan_l_t, 1), "launches": int(scan_l_t), "per": "GPU",
                         "issue": issue,
                         "ascii_kernel": {"kernel": "k_scan_minimizers<hpc=1,packed=0>", "achieved": achieved_ascii,
                                          "frac": achieved_ascii / peak if peak else None, "avg_launch_ms": scan_ms_a / max(scan_l_a, 1)},
                         "note": "the scan is bound by integer-instruction issue (two 64-bit ntHash rolls per symbol), not by HBM: "
                                 "`issue.frac` is the fraction of the issue-slot peak it sustains; see DESIGN.md section 5"},
            "stage_ms_last_step_rank0": stage_ms,
            "setup_s": {"generate_and_pack_workload": t_gen},
            "clocks": clk,
        }
        if cpu_line:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line), flush=True)

    for ptr in (p1, p2, p3, pg):
        L.mq_host_free(ptr)
    pk.close()
    for d in (d_seqs, d_hits, d_w, d_f, d_e):
        L.mq_dev_free(h, d)
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
