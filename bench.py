#!/usr/bin/env python
"""bench.py -- headline benchmark of the mapquik seeding->chaining hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): E. coli-sized genome, one contig of
4,641,652 bp (seed 2, uniform random), 100,000 synthetic HiFi-like reads, length N(10 kb, 1.5 kb)
clipped at 1 kb, 99.5 % identity (errors 1:1:1 sub/ins/del), seed 2 -> ~1 Gbp per step.  A "step" is
one pass of the whole hot path (S1 scan -> k-min-mers -> probe -> Match -> chain -> mq_hit) over all
reads of the rank.  At N > 1 every rank maps its own 100,000 reads (weak scaling, no data-path
collective); the index is built partitioned by reference chunk and replicated with one NCCL
all-gather of the minimizer store.

Printed JSON (one line, rank 0): see the contract in the task statement.  `value` = reads/s with the
inputs resident in HBM (mq_map_batch_device); `e2e` = the same through mq_map_batch with pinned HOST
buffers (H2D of the sequences and D2H of the hits inside the timed region); `roofline` is for the
dominant kernel k_scan_minimizers (algorithmic bytes = ASCII bases + 12 B per emitted minimizer);
`cpu_baseline` = the CPU oracle (a port, not the upstream Rust binary) on the box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GENOME_LEN = 4641652
N_READS = 100000
READ_MEAN, READ_SD, READ_MIN, READ_ERR = 10000.0, 1500.0, 1000, 0.005
SEED = 2
WORKLOAD = "ecoli_4.64Mbp_x_100k_hifi_reads_10kb_99.5pct (BASELINE configs[1])"


def make_workload(rank, n_reads):
    from mapquik_b200 import sim
    if os.environ.get("OMP_NUM_THREADS") == "1" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun pins OpenMP to one thread per rank; the (untimed) read simulator may use a fair share of the host
        os.environ["OMP_NUM_THREADS"] = str(max(1, host_threads() // int(os.environ["WORLD_SIZE"])))
    g, go, names = sim.genome(SEED, [GENOME_LEN], names=["chr000913"])
    rb, ro, _, _ = sim.reads(SEED, g, go, n_reads, READ_MEAN, READ_SD, READ_MIN, READ_ERR, first=rank * n_reads,
                             with_names=False)
    return g, go, names, rb, ro


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
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
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
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


def host_threads():
    """all host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def scan_traffic(n_reads):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_scan_minimizers_v3 launch on the default
    workload, from the committed `ncu --set full` capture (profiles/scan_traffic.json); null otherwise."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        return t["dram_bytes_per_launch"] if n_reads == t["reads_per_launch"] else None
    except Exception:
        return None


def scan_issue(n_reads, scan_avg_ms, n_sm, sm_mhz):
    """Second roofline of the dominant kernel: it is bound by integer-instruction issue, not HBM.  Warp-instructions per
    launch come from the committed ncu capture (they do not depend on timing), the per-SM issue peak from the committed
    microbenchmark, launch time and SM clock from this run."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        if n_reads != t["reads_per_launch"] or scan_avg_ms <= 0 or not sm_mhz:
            return None
        achieved = t["warp_instructions_per_launch"] / (scan_avg_ms / 1e3)
        peak = t["issue_peak_warp_instr_per_clk_per_sm"] * n_sm * sm_mhz * 1e6
        return {"achieved_warp_instr_per_s": achieved, "peak_warp_instr_per_s": peak, "frac": achieved / peak,
                "peak_source": t["issue_peak_source"]}
    except Exception:
        return None


def cpu_oracle_run(g, go, names, rb, ro, n_sample, steps, warmup, threads):
    """CPU oracle (port) on a bounded sample: returns (reads/s, bases/s, index_build_s)."""
    from oracle import pyoracle as O
    p = O.params()
    ix = O.Index(p, 1 << 17)
    t0 = time.perf_counter()
    ix.add_batch(names, g, go, threads=threads)
    n_unique = ix.count()
    t_index = time.perf_counter() - t0
    n_sample = min(n_sample, ro.size - 1)
    sro = ro[:n_sample + 1]
    srb = rb[:int(sro[-1])]
    for _ in range(warmup):
        ix.map_batch(srb, sro, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        hits = ix.map_batch(srb, sro, threads=threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_sample / dt, float(sro[-1]) / dt, t_index, dt, int(hits["mapped"].sum()), n_unique


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import pyoracle as O
    threads = host_threads()
    n_sample = 25000
    g, go, names, rb, ro = make_workload(0, n_sample)
    rps, bps, t_index, dt, mapped, n_unique = cpu_oracle_run(g, go, names, rb, ro, n_sample, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "reads/sec mapped (seeding->chaining hot path)", "value": rps, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "k": 5, "l": 31, "density": 0.01, "hpc": True},
        "gbp_per_s": bps / 1e9, "index_build_s": t_index,
        "cpu_baseline": {"value": rps, "unit": "reads/s", "cores": threads, "kind": "port",
                         "sample": f"first {n_sample} reads of the workload per step ({bps * dt / 1e6:.0f} Mbp), all host threads; "
                                   "CPU restatement of mapquik (oracle/), not the upstream Rust binary (no Rust toolchain offline)"},
        "e2e": {"value": rps, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mapped_reads": mapped, "n_unique_kminmers": n_unique,
    }
    print(json.dumps(line), flush=True)


def pinned_array(L, nbytes, dtype=np.uint8):
    ptr = L.mq_host_alloc(max(nbytes, 1))
    if not ptr:
        raise RuntimeError("mq_host_alloc failed")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(ptr)
    return np.frombuffer(buf, dtype=np.uint8, count=nbytes).view(dtype), ptr


def build_index_multi(ix, g, go, names, rank, world, dist, torch, device):
    """Index build partitioned by reference base-range chunk, then replicated: every rank scans its
    chunk, the minimizer stores are all-gathered over NCCL, every rank freezes the full store."""
    L = int(go[1] - go[0])
    lparam = ix.params.l
    cuts = [L * r // world for r in range(world + 1)]
    s, e = cuts[rank], cuts[rank + 1]
    lo = s - (1 if s > 0 else 0)
    # right halo: enough bytes to contain l-1 further run starts (checked on the host)
    hi = e
    need = lparam - 1
    while need > 0 and hi < L:
        nxt = min(L, hi + 4096)
        need -= int(np.count_nonzero(g[hi:nxt] != g[hi - 1:nxt - 1]))
        hi = nxt
    ix.add_segment(0, names[0], L, s, e - s, g[lo:hi])
    d_pos, d_hash, n, directory = ix.store_export()

    class CAI:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    counts = [None] * world
    dist.all_gather_object(counts, (int(n), directory.tolist()))
    nmax = max(c[0] for c in counts)
    pos_in = torch.zeros(nmax * 4, dtype=torch.uint8, device=device)
    hash_in = torch.zeros(nmax * 8, dtype=torch.uint8, device=device)
    if n:
        pos_in[:n * 4] = torch.as_tensor(CAI(d_pos, n * 4), device=device)
        hash_in[:n * 8] = torch.as_tensor(CAI(d_hash, n * 8), device=device)
    pos_all = torch.empty(world * nmax * 4, dtype=torch.uint8, device=device)
    hash_all = torch.empty(world * nmax * 8, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(pos_all, pos_in)
    dist.all_gather_into_tensor(hash_all, hash_in)
    pos_m = torch.cat([pos_all[r * nmax * 4: r * nmax * 4 + counts[r][0] * 4] for r in range(world)]).contiguous()
    hash_m = torch.cat([hash_all[r * nmax * 8: r * nmax * 8 + counts[r][0] * 8] for r in range(world)]).contiguous()
    torch.cuda.synchronize()
    dirs = np.array([d for c in counts for d in c[1]], dtype=np.uint64).reshape(-1, 3)
    ntot = sum(c[0] for c in counts)
    ix.store_import(pos_m.data_ptr(), hash_m.data_ptr(), ntot, dirs)
    return ix.freeze()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=N_READS, help="reads per rank per step (default: the named workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    dist = torch = None
    device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        device = torch.device(f"cuda:{local_rank}")
        dist.init_process_group("nccl", device_id=device)
        warm = torch.zeros(1, device=device)
        dist.all_reduce(warm)                 # communicator set-up is not part of the index build
        torch.cuda.synchronize()

    from mapquik_b200 import Index, Params, capi, HIT_DTYPE
    L = capi.lib()
    g, go, names, rb, ro = make_workload(rank, args.reads)
    n_reads = ro.size - 1
    n_bases = int(ro[-1])
    p = Params()

    # ---- index build (timed end to end from host memory) ---------------------------------------
    ix = Index(p, device=local_rank)
    t0 = time.perf_counter()
    if world > 1:
        n_unique = build_index_multi(ix, g, go, names, rank, world, dist, torch, device)
    else:
        ix.add_batch(names, g, go)
        n_unique = ix.freeze()
    index_build_s = time.perf_counter() - t0
    h = ix.handle

    # ---- device-resident inputs ----------------------------------------------------------------
    d_seqs = L.mq_dev_alloc(h, n_bases + 256); d_offs = L.mq_dev_alloc(h, (n_reads + 1) * 8)
    d_hits = L.mq_dev_alloc(h, n_reads * 48)
    assert d_seqs and d_offs and d_hits
    L.mq_dev_memset(h, d_seqs, 0, n_bases + 256)
    assert L.mq_dev_upload(h, d_seqs, rb.ctypes.data, n_bases) == 0
    assert L.mq_dev_upload(h, d_offs, ro.ctypes.data, (n_reads + 1) * 8) == 0

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_dev():
        ix.map_batch_device(d_seqs, d_offs, n_reads, n_bases, d_hits)

    clocks = ClockSampler(local_rank)     # sampled from the warm-up through both timed regions
    for _ in range(args.warmup):
        step_dev()
    L.mq_sync(h)
    launches0 = ix.launch_count(); sk0 = L.mq_scan_kernel_launches(h); L.mq_minimizer_count(h, 1)
    scan_ms = 0.0
    barrier()
    L.mq_region_begin(h)
    t0 = time.perf_counter()
    scan_ms0 = ix.total_ms("scan_kernel")
    for _ in range(args.steps):
        step_dev()
    dev_ms = L.mq_region_end_ms(h)
    scan_ms = ix.total_ms("scan_kernel") - scan_ms0
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    launches = ix.launch_count() - launches0
    scan_launches = L.mq_scan_kernel_launches(h) - sk0
    n_min = L.mq_minimizer_count(h, 1) // max(args.steps, 1)
    stage_ms = {s: ix.last_ms(s) for s in ("scan", "scan_kernel", "gather", "probe", "chain")}
    hits = np.zeros(n_reads, HIT_DTYPE)
    assert L.mq_dev_download(h, hits.ctypes.data, d_hits, n_reads * 48) == 0

    # ---- end to end through the C ABI with pinned HOST buffers ----------------------------------
    h_seqs, p1 = pinned_array(L, n_bases); h_offs, p2 = pinned_array(L, (n_reads + 1) * 8, np.uint64)
    h_hits, p3 = pinned_array(L, n_reads * 48)
    h_seqs[:] = rb; h_offs[:] = ro
    h_hits_v = h_hits.view(HIT_DTYPE)
    for _ in range(args.warmup):
        ix.map_batch(h_seqs, h_offs, out=h_hits_v)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ix.map_batch(h_seqs, h_offs, out=h_hits_v)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    clk = clocks.stop()
    e2e_stage = {s: ix.last_ms(s) for s in ("h2d", "scan", "gather", "probe", "chain", "d2h")}
    assert h_hits_v.tobytes() == hits.tobytes(), "e2e and device-resident results differ"

    # max over ranks
    t_dev, t_e2e = max(dev_ms, wall_ms), e2e_ms
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, index_build_s], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, index_build_s = (float(x) for x in tt.tolist())
        mm = torch.tensor([float(hits["mapped"].sum()), float(n_bases)], dtype=torch.float64, device=device)
        dist.all_reduce(mm, op=dist.ReduceOp.SUM)
        mapped_total, bases_total = (float(x) for x in mm.tolist())
    else:
        mapped_total, bases_total = float(hits["mapped"].sum()), float(n_bases)

    if rank == 0:
        peak, peak_src = peaks()
        reads_total = n_reads * world
        value = reads_total * args.steps / (t_dev / 1e3)
        e2e_value = reads_total * args.steps / (t_e2e / 1e3)
        algo_bytes = n_bases + 12 * n_min                 # per launch: ASCII in + (pos u32, hash u64) out
        scan_avg_ms = scan_ms / max(scan_launches, 1)
        achieved = algo_bytes / (scan_avg_ms / 1e3) / 1e9 if scan_avg_ms > 0 else 0.0
        line = {
            "metric": "reads/sec mapped (seeding->chaining hot path)", "value": value, "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reads_per_gpu": n_reads, "bases_per_gpu": n_bases, "k": p.k, "l": p.l,
                       "density": p.density, "hpc": True, "l2": "inputs (1 GB of reads) larger than L2; no flush needed",
                       "index": "replicated per GPU; built by reference chunk + NCCL all-gather at N>1"},
            "gbp_per_s": bases_total * args.steps / (t_dev / 1e3) / 1e9,
            "index_build_s": index_build_s, "n_unique_kminmers": int(n_unique),
            "mapped_fraction": mapped_total / reads_total,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(n_bases + (n_reads + 1) * 8),
                    "d2h_bytes_per_step": int(n_reads * 48), "gbp_per_s": bases_total * args.steps / (t_e2e / 1e3) / 1e9,
                    "stage_ms_last_step": e2e_stage},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_scan_minimizers_v3", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": scan_traffic(n_reads), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(algo_bytes), "avg_launch_ms": scan_avg_ms,
                         "launches": int(scan_launches),
                         "issue": scan_issue(n_reads, scan_avg_ms, 148, (clk or {}).get("sm_mhz")),   # B200: 148 SMs
                         "note": "integer-issue bound (64-bit ntHash roll per base), see DESIGN.md section 5: `issue` is "
                                 "the fraction of the measured instruction-issue peak the kernel sustains"},
            "stage_ms_last_step": stage_ms,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            n_sample = min(25000, n_reads)
            rps, bps, t_index, dt, mapped, _ = cpu_oracle_run(g, go, names, rb, ro, n_sample, 2, 1, threads)
            line["cpu_baseline"] = {"value": rps, "unit": "reads/s", "cores": threads, "kind": "port",
                                    "sample": f"first {n_sample} reads ({bps * dt / 1e6:.0f} Mbp) x 2 timed passes after 1 warm-up, "
                                              "OpenMP over reads; CPU restatement (oracle/), not the upstream Rust binary",
                                    "gbp_per_s": bps / 1e9, "index_build_s": t_index}
        print(json.dumps(line), flush=True)

    for ptr in (p1, p2, p3):
        L.mq_host_free(ptr)
    for d in (d_seqs, d_offs, d_hits):
        L.mq_dev_free(h, d)
    ix.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
