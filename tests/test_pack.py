"""Host side of the packed input format (mq_pack / mq_pack_at / mq_unpack): codes + block bitmap + exception intervals
must be EXACTLY the ASCII bytes -- everything the GPU path does with packed input rests on that."""
import ctypes as C
import threading

import numpy as np
import pytest

from conftest import random_dna
from mapquik_b200 import EXC_DTYPE, PackedSeqs, capi


def messy(rng, n):
    a = random_dna(rng, n).copy()
    if n > 100:
        for _ in range(max(1, n // 5000)):
            s = int(rng.integers(0, n - 50)); ln = int(rng.integers(1, 40))
            a[s:s + ln] = ord("N")
        idx = rng.integers(0, n, max(1, n // 2000))
        a[idx] = np.frombuffer(b"RYKMSWBDHVN*-\x00\xc1\xff", np.uint8)[rng.integers(0, 16, idx.size)]
    return a


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 2047, 2048, 2049, 100003, 1 << 20])
@pytest.mark.parametrize("threads", [1, 4])
def test_pack_roundtrip(n, threads):
    rng = np.random.default_rng(n + threads)
    a = messy(rng, n)
    p = PackedSeqs(a, n_threads=threads)
    assert np.array_equal(p.unpack(), a)
    # codes are (byte >> 1) & 3 for every base, exceptions included
    w = p.words
    codes = (w[np.arange(n) >> 4] >> (2 * (np.arange(n) & 15)).astype(np.uint32)) & 3 if n else np.zeros(0, np.uint32)
    assert np.array_equal(codes.astype(np.uint8), (a >> 1) & 3)
    # the block bitmap is exact
    bad = ~np.isin(a, np.frombuffer(b"ACGT", np.uint8))
    blocks = np.zeros((n + 63) // 64, bool)
    np.logical_or.at(blocks, np.arange(n)[bad] >> 6, True)
    got = np.array([(p.flags[b >> 5] >> (b & 31)) & 1 for b in range(blocks.size)], bool)
    assert np.array_equal(got, blocks)
    # intervals: sorted, disjoint, cover exactly the non-ACGT bytes
    e = p.exc
    assert np.all(e["start"][1:] >= e["start"][:-1] + e["len"][:-1])
    cover = np.zeros(n, bool)
    for s, ln, b in zip(e["start"], e["len"], e["byte"]):
        assert np.all(a[int(s):int(s) + int(ln)] == b)
        cover[int(s):int(s) + int(ln)] = True
    assert np.array_equal(cover, bad)
    # slack behind the last base reads as zero
    assert not w[(n + 15) // 16:].any()


def test_pack_fold_case():
    a = np.frombuffer(b"acgtnACGTNxyzacgtacgtacgtacgtacgtacgtacgtacgtRr", np.uint8)
    p = PackedSeqs(a, fold_case=True)
    assert bytes(p.unpack()) == bytes(a).upper()
    q = PackedSeqs(a, fold_case=False)
    assert bytes(q.unpack()) == bytes(a)          # lower case is NOT folded unless asked (the ABI contract)


@pytest.mark.parametrize("n", [64, 200, 100003])
def test_pack_fold_case_long(n):
    """the wide (64 bases per step) packer folds case through its lookup table; bytes >= 0x80 must never alias a letter"""
    rng = np.random.default_rng(n)
    a = messy(rng, n)
    low = rng.random(n) < 0.5
    a = np.where(low & (a >= 65) & (a <= 90), a + 32, a).astype(np.uint8)
    a[rng.integers(0, n, 5)] = np.frombuffer(b"\xc1\xe1\xc7\xd4\xf4", np.uint8)     # 'A' | 0x80, 'a' | 0x80, ...
    p = PackedSeqs(a, fold_case=True)
    assert bytes(p.unpack()) == bytes(a.tobytes().upper())
    q = PackedSeqs(a, fold_case=False)
    assert np.array_equal(q.unpack(), a)


def test_pack_giant_run_is_one_interval():
    a = np.concatenate([np.frombuffer(b"ACGT" * 10, np.uint8), np.full(3_000_001, ord("N"), np.uint8), np.frombuffer(b"TTGA", np.uint8)])
    p = PackedSeqs(a, n_threads=1)
    assert p.exc.size == 1 and int(p.exc[0]["start"]) == 40 and int(p.exc[0]["len"]) == 3_000_001
    p4 = PackedSeqs(a, n_threads=4)
    assert np.array_equal(p4.unpack(), a) and p4.exc.size <= 4


def test_pack_at_concurrent_ranges():
    """a parser's copy jobs pack disjoint ranges of one destination concurrently; edges share words"""
    L = capi.lib()
    rng = np.random.default_rng(5)
    n = 300001
    a = messy(rng, n)
    words = np.zeros(int(L.mq_packed_words(n)), np.uint32); flags = np.zeros(int(L.mq_packed_flag_words(n)), np.uint32)
    cuts = sorted(set([0, n] + [int(x) for x in rng.integers(0, n, 40)]))
    excs = [None] * (len(cuts) - 1)

    def job(i):
        s, e = cuts[i], cuts[i + 1]
        ex = np.zeros(e - s + 1, EXC_DTYPE); k = C.c_uint64()
        rc = L.mq_pack_at(a[s:e].ctypes.data, e - s, s, words.ctypes.data, flags.ctypes.data, ex.ctypes.data, ex.size, C.byref(k), 0)
        assert rc == 0
        excs[i] = ex[:k.value]
    th = [threading.Thread(target=job, args=(i,)) for i in range(len(cuts) - 1)]
    [t.start() for t in th]; [t.join() for t in th]
    exc = np.concatenate(excs)
    out = np.zeros(n, np.uint8)
    assert L.mq_unpack(words.ctypes.data, exc.ctypes.data, exc.size, 0, n, out.ctypes.data) == 0
    assert np.array_equal(out, a)
    ref = PackedSeqs(a, n_threads=1)
    assert np.array_equal(words, ref.words) and np.array_equal(flags, ref.flags)


def test_pack_at_aligned_chunks_need_no_zeroing():
    """what the CLI's parser relies on: a chunk that starts on a 32-base boundary and holds whole 32-base groups is written
    with plain stores (code words are overwritten, not OR-ed), so a destination full of stale bits needs no clearing; only
    the block bitmap and the words of a ragged tail must be zero beforehand"""
    L = capi.lib()
    rng = np.random.default_rng(17)
    n = 64 * 3000 + 37
    a = messy(rng, n)
    a[7000:7100] = ord("g")
    ref = PackedSeqs(a, n_threads=1, fold_case=True)
    words = np.full(int(L.mq_packed_words(n)), 0xFFFFFFFF, np.uint32); flags = np.zeros(int(L.mq_packed_flag_words(n)), np.uint32)
    tail = n & ~31
    words[tail >> 4:] = 0
    excs = []; at = 0
    for size in [64 * int(x) for x in rng.integers(1, 130, 400)]:
        size = min(size, n - at) if at + size >= tail else size
        if size <= 0:
            break
        ex = np.zeros(size + 1, EXC_DTYPE); k = C.c_uint64()
        assert L.mq_pack_at(a[at:at + size].ctypes.data, size, at, words.ctypes.data, flags.ctypes.data, ex.ctypes.data, ex.size, C.byref(k), 1) == 0
        excs.append(ex[:k.value]); at += size
    assert at == n
    assert np.array_equal(words, ref.words) and np.array_equal(flags, ref.flags)
    exc = np.concatenate(excs)
    out = np.zeros(n, np.uint8)
    assert L.mq_unpack(words.ctypes.data, exc.ctypes.data, exc.size, 0, n, out.ctypes.data) == 0
    assert np.array_equal(out, ref.unpack())


def test_pack_narrow_path_equals_wide_path():
    """the 64-bases-per-step (AVX-512 VBMI) and the 32-bases-per-step (AVX2) packers must write the same words, bitmap and
    intervals; which one runs is decided once per process, so the narrow one is exercised in a child (MQ_PACK_NO_AVX512=1)"""
    import os
    import subprocess
    import sys
    rng = np.random.default_rng(99)
    a = messy(rng, 300007)
    a[1000:1100] = ord("a")                               # lower case, folded on request
    here = PackedSeqs(a, n_threads=3, fold_case=True)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from mapquik_b200 import PackedSeqs;"
            "a = np.frombuffer(sys.stdin.buffer.read(), np.uint8); p = PackedSeqs(a, n_threads=3, fold_case=True);"
            "sys.stdout.buffer.write(p.words.tobytes() + p.flags.tobytes() + p.exc.tobytes())") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ); env["MQ_PACK_NO_AVX512"] = "1"
    r = subprocess.run([sys.executable, "-c", code], input=a.tobytes(), capture_output=True, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    assert r.stdout == here.words.tobytes() + here.flags.tobytes() + here.exc.tobytes()
