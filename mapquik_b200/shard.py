"""Host-side partitioning logic of the multi-GPU path (one process per GPU).

  - index build: every reference record is cut into `world` base ranges; rank r scans range r
    (mq_index_add_segment, with one byte of left context and a right halo that holds l-1 further
    homopolymer-run starts); the per-rank minimizer stores are all-gathered and every rank freezes
    the union (mq_store_import + mq_index_freeze), so the unique-or-tombstone rule of
    index.rs:100-104 is applied genome-wide on every GPU;
  - read mapping: reads are sharded in contiguous blocks, no collective.

Only numpy here; the collective itself is torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
import numpy as np


def chunk_bounds(length, world):
    """`world`+1 cut points of [0, length) -- contiguous, near-equal base ranges."""
    return [length * r // world for r in range(world + 1)]


def halo_end(seq, end, l, use_hpc=True, step=4096):
    """Smallest (step-granular) hi >= end such that seq[end:hi] holds l-1 run starts, or len(seq)."""
    n = len(seq)
    hi, need = end, l - 1
    while need > 0 and hi < n:
        nxt = min(n, hi + step)
        if use_hpc and hi > 0:
            need -= int(np.count_nonzero(seq[hi:nxt] != seq[hi - 1:nxt - 1]))
        else:
            need -= nxt - hi
        hi = nxt
    return hi


def segment_for_rank(seq, rank, world, l, use_hpc=True):
    """(seg_start, own_len, data) for this rank's share of one reference record."""
    cuts = chunk_bounds(len(seq), world)
    s, e = cuts[rank], cuts[rank + 1]
    lo = s - (1 if s > 0 else 0)
    hi = halo_end(seq, e, l, use_hpc)
    return s, e - s, seq[lo:hi]


def merge_stores(parts):
    """parts: per rank (pos u32[], hash u64[], directory [(ref_idx, seg_start, count), ...]).
    Returns (pos, hash, directory) of the union, segments ordered by (ref_idx, seg_start)."""
    segs = []
    for pos, hs, d in parts:
        o = 0
        for ref_idx, seg_start, cnt in d:
            cnt = int(cnt)
            segs.append((int(ref_idx), int(seg_start), pos[o:o + cnt], hs[o:o + cnt]))
            o += cnt
    segs.sort(key=lambda t: (t[0], t[1]))
    pos = np.concatenate([s[2] for s in segs]) if segs else np.zeros(0, np.uint32)
    hs = np.concatenate([s[3] for s in segs]) if segs else np.zeros(0, np.uint64)
    d = np.array([(s[0], s[1], len(s[2])) for s in segs], dtype=np.uint64).reshape(-1, 3)
    return pos, hs, d


def read_shard(n_reads, rank, world):
    """contiguous block of reads for this rank: [lo, hi)"""
    return n_reads * rank // world, n_reads * (rank + 1) // world
