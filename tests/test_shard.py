"""Host-side logic of the multi-GPU path, on CPU: reference partitioning with halo, store merge and
read sharding -- exercised with world_size 2 over gloo.  The per-rank "scan" is done by the CPU
oracle here (checker standing in for the GPU kernel); the partition / merge code is the product's."""
import os
import socket

import numpy as np
import pytest

from conftest import random_dna
from mapquik_b200 import shard


def oracle_segment_minimizers(seq, seg_start, own_len, data, p):
    """what mq_index_add_segment computes: minimizers starting in [seg_start, seg_start+own_len)"""
    from oracle import pyoracle as O
    ctx = 1 if seg_start > 0 else 0
    # scanning the window with its left context byte reproduces the run starts of the whole record
    pos, hs = O.minimizers(data, p)
    if ctx and data[0] != data[1]:
        pass                      # context byte is its own run: minimizers starting on it are dropped below
    pos = pos.astype(np.int64) - ctx + seg_start
    keep = (pos >= seg_start) & (pos < seg_start + own_len)
    return pos[keep].astype(np.uint32), hs[keep]


def test_chunk_bounds_and_read_shards():
    assert shard.chunk_bounds(10, 3) == [0, 3, 6, 10]
    covered = []
    for r in range(4):
        lo, hi = shard.read_shard(1001, r, 4)
        covered += list(range(lo, hi))
    assert covered == list(range(1001))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partitioned_minimizers_equal_whole(world):
    from oracle import pyoracle as O
    rng = np.random.default_rng(world)
    seq = random_dna(rng, 60000).copy()
    seq[20000:27000] = ord("A")                      # a homopolymer that swallows a cut point
    seq[45000:45040] = ord("N")
    p = O.params()
    wpos, whs = O.minimizers(seq, p)
    parts = []
    for r in range(world):
        s, own, data = shard.segment_for_rank(seq, r, world, p.l)
        pos, hs = oracle_segment_minimizers(seq, s, own, data, p)
        parts.append((pos, hs, [(0, s, len(pos))]))
    pos, hs, d = shard.merge_stores(parts[::-1])     # arrival order must not matter
    assert np.array_equal(pos.astype(np.uint64), wpos) and np.array_equal(hs, whs)
    assert d[:, 1].tolist() == shard.chunk_bounds(len(seq), world)[:-1]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import pyoracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    seqs = [random_dna(rng, 50000), random_dna(rng, 21000)]
    p = O.params()
    mine_pos, mine_hs, mine_dir = [], [], []
    for ref_idx, seq in enumerate(seqs):
        s, own, data = shard.segment_for_rank(seq, rank, world, p.l)
        pos, hs = oracle_segment_minimizers(seq, s, own, data, p)
        mine_pos.append(pos); mine_hs.append(hs); mine_dir.append((ref_idx, s, len(pos)))
    gathered = [None] * world
    dist.all_gather_object(gathered, (np.concatenate(mine_pos), np.concatenate(mine_hs), mine_dir))
    pos, hs, d = shard.merge_stores(gathered)
    # every rank must now hold the same, complete, ordered store
    exp_pos = np.concatenate([O.minimizers(s, p)[0] for s in seqs]); exp_hs = np.concatenate([O.minimizers(s, p)[1] for s in seqs])
    ok = np.array_equal(pos.astype(np.uint64), exp_pos) and np.array_equal(hs, exp_hs)
    lo, hi = shard.read_shard(1000, rank, world)
    counts = [None] * world
    dist.all_gather_object(counts, hi - lo)
    ok = ok and sum(counts) == 1000
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok), d.tolist()))


def test_world_size_2_gloo_index_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert res[0][2] == res[1][2]                    # identical directories on both ranks
