"""GPU parity: every stage of the CUDA path, through the C ABI, bit-exact against the CPU oracle."""
import numpy as np
import pytest

from conftest import random_dna, revcomp
from oracle import pyoracle as O
from mapquik_b200 import Index, PackedSeqs, Params, concat, sim

pytestmark = pytest.mark.gpu


def oparams(p):
    return O.params(p.k, p.l, p.density, p.use_hpc, p.c, p.s, p.g)


def oracle_minimizers(seqs, offs, p):
    pos, hs, so = [], [], [0]
    for i in range(len(offs) - 1):
        a, b = O.minimizers(seqs[int(offs[i]):int(offs[i + 1])], oparams(p))
        pos.append(a); hs.append(b); so.append(so[-1] + len(a))
    return np.array(so, np.uint64), np.concatenate(pos) if pos else np.zeros(0, np.uint64), \
        np.concatenate(hs) if hs else np.zeros(0, np.uint64)


@pytest.fixture(params=["ascii", "packed"])
def fmt(request):
    """both input formats of the S1 kernel: upper-cased ASCII and 2-bit packed codes (+ exception intervals)"""
    return request.param


def check_minimizers(seqs, offs, p, fmt="ascii"):
    ix = Index(p)
    if fmt == "packed":
        pk = PackedSeqs(seqs)
        assert np.array_equal(pk.unpack(), seqs)
        so, pos, hs = ix.minimizers_packed(pk, offs)
    else:
        so, pos, hs = ix.minimizers(seqs, offs)
    eso, epos, ehs = oracle_minimizers(seqs, offs, p)
    assert np.array_equal(so, eso), (so[:10], eso[:10])
    assert np.array_equal(pos.astype(np.uint64), epos)
    assert np.array_equal(hs, ehs)
    ix.close()
    return len(pos)


def adversarial_seqs(rng):
    seqs = []
    seqs.append(random_dna(rng, 50000))
    seqs.append(np.frombuffer(b"A" * 5000, np.uint8))                       # one homopolymer: a single symbol
    seqs.append(np.frombuffer(b"AC" * 4000, np.uint8))                      # dinucleotide repeat (period 2)
    x = random_dna(rng, 30000).copy(); x[10000:10040] = ord("N"); x[20000] = ord("N"); seqs.append(x)
    y = random_dna(rng, 40000).copy(); y[12000:32000] = ord("T"); seqs.append(y)   # 20 kb homopolymer inside
    seqs.append(random_dna(rng, 34))                                        # shorter than l+k-1 (35)
    seqs.append(random_dna(rng, 35))
    seqs.append(random_dna(rng, 31))
    seqs.append(np.zeros(0, np.uint8))                                      # empty record
    seqs.append(random_dna(rng, 8191)); seqs.append(random_dna(rng, 8192)); seqs.append(random_dna(rng, 8193))
    seqs.append(random_dna(rng, 16385)); seqs.append(random_dna(rng, 1)); seqs.append(random_dna(rng, 3))
    z = np.repeat(random_dna(rng, 3000), rng.integers(1, 12, 3000)); seqs.append(z)    # many long runs
    w = random_dna(rng, 9000).copy(); w[-2000:] = ord("G"); seqs.append(w)   # record ends in a long run
    v = random_dna(rng, 9000).copy(); v[:3000] = ord("C"); seqs.append(v)    # record starts with a long run
    seqs.append(np.frombuffer(b"acgtnACGT" * 500, np.uint8))                # lower case is NOT folded at the ABI
    # islands of a few symbols between runs longer than a lane chunk: most lanes of a tile hold no symbol at all, the
    # context of a lane comes from several sparse streams further right (v3: ballot walk over non-empty streams)
    parts = []
    for i in range(120):
        parts.append(np.full(int(rng.integers(100, 700)), ord("ACGTN"[i % 5]), np.uint8))
        parts.append(random_dna(rng, int(rng.integers(1, 40))))
    seqs.append(np.concatenate(parts))
    # IUPAC / garbage bytes, including ones that share their low bits with A, C, G, T (E, P, R, V, 0xC1, ...)
    q = random_dna(rng, 6000).copy()
    q[rng.integers(0, 6000, 300)] = np.frombuffer(b"EPRVBDHKMSWYU*-\xc1\xc3\xc7\xd4\x01\x00", np.uint8)[rng.integers(0, 21, 300)]
    seqs.append(q)
    return seqs


@pytest.mark.parametrize("l,density,hpc", [(31, 0.01, True), (16, 0.01, True), (31, 0.05, False), (5, 0.3, True),
                                           (32, 0.02, True), (2, 0.5, True), (25, 1.0, True),
                                           # full 128-symbol lane streams with the longest window: the least spare rows
                                           # for parked candidates (v3), first sparse, then every l-mer selected
                                           (32, 0.05, False), (32, 1.0, False),
                                           # l = 31 (the default) without HPC (full 128-symbol streams), densities right
                                           # under / over 1/64 (hash threshold around 2^58: twice the usual number of
                                           # parked candidates per lane), and a threshold far below the 32-bit pre-filter
                                           (31, 0.01, False), (31, 0.0156, True), (31, 0.0156, False), (31, 0.0157, True),
                                           (31, 0.0005, True)])
def test_minimizers_adversarial(l, density, hpc, fmt):
    rng = np.random.default_rng(7)
    buf, offs = concat_raw(adversarial_seqs(rng))
    check_minimizers(buf, offs, Params(k=5, l=l, density=density, use_hpc=hpc), fmt)


def concat_raw(seqs):
    offs = np.zeros(len(seqs) + 1, np.uint64); offs[1:] = np.cumsum([len(s) for s in seqs])
    return (np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)), offs


def test_minimizers_reads_and_genome(fmt):
    g, go, _ = sim.genome(11, [1500000, 700001, 123457])
    n = check_minimizers(g, go, Params(), fmt)
    assert n > 20000
    rb, ro, _, _ = sim.reads(11, g, go, 500, 10000, 3000)
    check_minimizers(rb, ro, Params(), fmt)
    check_minimizers(rb, ro, Params(l=16, k=8), fmt)


def test_minimizers_dense_overflow_pool(fmt):
    # density 1.0 selects every l-mer: every tile overflows its staged-event pool into the global pool, and the
    # minimizer buffers sized for the expected density are too small (status record -> redo with larger buffers)
    rng = np.random.default_rng(3)
    buf, offs = concat_raw([random_dna(rng, 100000), random_dna(rng, 20000)])
    check_minimizers(buf, offs, Params(l=7, density=1.0, use_hpc=False), fmt)
    big, boffs = concat_raw([random_dna(rng, 3000000)])
    check_minimizers(big, boffs, Params(l=9, density=1.0, use_hpc=False), fmt)


def test_giant_runs_stay_fast_and_exact(fmt):
    # a 6 Mbp N-gap and a 3 Mbp homopolymer inside one record: tiles inside a run must not walk to its end
    import time
    rng = np.random.default_rng(12)
    x = np.concatenate([random_dna(rng, 200000), np.full(6000000, ord("N"), np.uint8), random_dna(rng, 150000),
                        np.full(3000000, ord("A"), np.uint8), random_dna(rng, 100000)])
    buf, offs = concat_raw([x, random_dna(rng, 50000)])
    t0 = time.perf_counter()
    check_minimizers(buf, offs, Params(), fmt)
    assert time.perf_counter() - t0 < 60


def test_batch_split_invariance_and_idempotence():
    # size-independent properties at a larger size: mapping is a pure function of (read, index), so any
    # partition of the batch and any repetition must give the same bytes
    p = Params()
    g, go, names = sim.genome(71, [3000000, 1000000])
    ix = Index(p); ix.add_batch(names, g, go); ix.freeze()
    rb, ro, _, _ = sim.reads(71, g, go, 20000, 12000, 4000)
    whole = ix.map_batch(rb, ro)
    assert ix.map_batch(rb, ro).tobytes() == whole.tobytes()
    parts = []
    for lo, hi in ((0, 1), (1, 7001), (7001, 7002), (7002, 20000)):
        sub = ro[lo:hi + 1] - ro[lo]
        parts.append(ix.map_batch(rb[int(ro[lo]):int(ro[hi])], sub))
    assert np.concatenate(parts).tobytes() == whole.tobytes()
    # minimizer positions are strictly increasing inside every record, hashes below the bound
    so, pos, hs = ix.minimizers(rb[:int(ro[2000])], ro[:2001])
    for i in range(2000):
        q = pos[int(so[i]):int(so[i + 1])].astype(np.int64)
        assert np.all(np.diff(q) > 0)
    assert np.all(hs < 0x28f5c28f5c28f60)
    # reference order does not matter (ids are relabelled, everything else is identical)
    ix2 = Index(p)
    ix2.add_batch([names[1]], g[int(go[1]):], np.array([0, int(go[2] - go[1])], np.uint64), first_ref_idx=0)
    ix2.add_batch([names[0]], g[:int(go[1])], np.array([0, int(go[1])], np.uint64), first_ref_idx=1)
    ix2.freeze()
    assert ix2.n_unique == ix.n_unique and ix2.n_keys == ix.n_keys
    h2 = ix2.map_batch(rb, ro)
    assert np.array_equal(h2["ref_idx"][whole["mapped"] == 1], 1 - whole["ref_idx"][whole["mapped"] == 1])
    for f in ("mapped", "rc", "mapq", "q_start", "q_end", "r_start", "r_end", "score"):
        assert np.array_equal(h2[f], whole[f])
    ix.close(); ix2.close()


def test_kminmers():
    rng = np.random.default_rng(5)
    g, go, _ = sim.genome(12, [300000, 50000])
    seqs = [g[:300000], g[300000:], random_dna(rng, 200), random_dna(rng, 36), random_dna(rng, 20)]
    buf, offs = concat_raw(seqs)
    for p in (Params(), Params(k=8, l=16), Params(k=1, l=20, density=0.05), Params(k=3, l=31, use_hpc=False)):
        ix = Index(p)
        so, st, en, off, rev, hs = ix.kminmers(buf, offs)
        exp = [O.kminmers(s, oparams(p)) for s in seqs]
        eso = np.cumsum([0] + [len(e) for e in exp])
        assert np.array_equal(so, eso.astype(np.uint64))
        e = np.concatenate(exp)
        assert np.array_equal(st, e["start"]) and np.array_equal(en, e["end"])
        assert np.array_equal(off, e["offset"]) and np.array_equal(rev, e["rev"]) and np.array_equal(hs, e["hash"])
        ix.close()


def build_both(p, names, g, go):
    ix = Index(p)
    nb = ix.add_batch(names, g, go)
    n_unique = ix.freeze()
    oix = O.Index(oparams(p), 1 << 16)
    onb = oix.add_batch(names, g, go)
    assert np.array_equal(nb, onb)
    assert n_unique == oix.count()
    assert ix.n_keys == oix.slots()
    return ix, oix


def test_index_unique_or_tombstone():
    # a genome with exact duplications: duplicated k-min-mers must be tombstones, the rest unique
    rng = np.random.default_rng(9)
    a = random_dna(rng, 200000)
    b = np.concatenate([random_dna(rng, 50000), a[20000:90000], random_dna(rng, 30000), revcomp(a[100000:150000])])
    buf, offs = concat_raw([a, b])
    p = Params()
    ix, oix = build_both(p, ["a", "b"], buf, offs)
    assert ix.n_keys > ix.n_unique            # some tombstones exist
    # probe every reference k-min-mer and a batch of absent keys
    _, st, en, off, rev, hs = Index(p).kminmers(buf, offs)
    keys = np.concatenate([hs, rng.integers(0, 2**63, 1000).astype(np.uint64), np.array([2**64 - 1, 0], np.uint64)])
    f, rid, s, e, o, rc = ix.get(keys)
    for i, h in enumerate(keys):
        exp = oix.get(int(h))
        assert bool(f[i]) == (exp is not None)
        if exp is not None:
            assert (int(rid[i]), int(s[i]), int(e[i]), int(o[i]), int(rc[i])) == exp
    ix.close()


def compare_matches(ix, oix, rb, ro):
    mo, f = ix.matches(rb, ro)
    for i in range(len(ro) - 1):
        em = oix.chain_matches(rb[int(ro[i]):int(ro[i + 1])])
        got = f[int(mo[i]):int(mo[i + 1])]
        assert len(em) == len(got), (i, len(em), len(got))
        if len(em):
            exp = np.stack([em["q_start"], em["q_end"], em["r_start"], em["r_end"], em["count"],
                            (em["ref_id"].astype(np.uint64) << 1) | em["rc"]], axis=1).astype(np.uint32)
            assert np.array_equal(got, exp), (i, got[:5], exp[:5])


def compare_hits(ix, oix, rb, ro, names=None):
    hits = ix.map_batch(rb, ro)
    ohits = oix.map_batch(rb, ro)
    for f in ("mapped", "rc", "mapq", "ref_idx", "q_start", "q_end", "r_start", "r_end", "score"):
        bad = np.nonzero(hits[f] != ohits[f])[0]
        assert bad.size == 0, (f, bad[:10], hits[bad[:3]], ohits[bad[:3]])
    if names is not None:
        for i in range(min(len(names), 200)):
            ln = int(ro[i + 1] - ro[i])
            a = ix.paf_line(names[i], ln, hits[i])
            b = oix.paf_line(names[i], ln, ohits[i]) if ohits[i]["mapped"] else None
            assert a == b
    return hits


def test_matches_and_hits_ecoli_like():
    p = Params()
    g, go, names = sim.genome(21, [2000000, 1000000, 400000])
    ix, oix = build_both(p, names, g, go)
    rb, ro, rn, tr = sim.reads(21, g, go, 3000, 10000, 2500, contig_names=names)
    compare_matches(ix, oix, rb, ro)
    hits = compare_hits(ix, oix, rb, ro, rn)
    assert hits["mapped"].mean() > 0.95
    ok = (hits["ref_idx"] == tr["contig"]) & (hits["rc"] == tr["strand"]) & \
        (np.minimum(hits["r_end"], tr["start"] + tr["len"]).astype(np.int64) -
         np.maximum(hits["r_start"], tr["start"]).astype(np.int64) > 0.1 * tr["len"])
    assert (ok | (hits["mapped"] == 0)).mean() > 0.99
    ix.close()


def repeat_genome(rng, n_contigs=10, units=40, fam=5, fam_len=3000, div=0.003):
    """contigs made of near-identical copies of a few families (0.3 % substitutions) and unique spacers"""
    cons = [random_dna(rng, fam_len) for _ in range(fam)]
    contigs = []
    for _ in range(n_contigs):
        parts = []
        for _ in range(units):
            if rng.random() < 0.7:
                c = cons[rng.integers(fam)].copy()
                mut = rng.random(c.size) < div
                c[mut] = random_dna(rng, int(mut.sum()))
                parts.append(revcomp(c) if rng.random() < 0.5 else c)
            else:
                parts.append(random_dna(rng, int(rng.integers(500, 4000))))
        contigs.append(np.concatenate(parts))
    return concat_raw(contigs)


def test_hits_repetitive_many_refs():
    # repeat-rich multi-contig genome: tombstones, short Matches, cross-reference ties, the fwd-Match
    # `check` quirk (ref id / strand not compared) all get exercised
    p = Params(k=3, l=15, density=0.03, c=2, s=3, g=500)
    rng = np.random.default_rng(31)
    g, go = repeat_genome(rng)
    names = [f"ctg{i}" for i in range(len(go) - 1)]
    ix, oix = build_both(p, names, g, go)
    assert ix.n_keys > 1.02 * ix.n_unique             # tombstones (every repeat-family k-min-mer)
    rb, ro, rn, _ = sim.reads(31, g, go, 4000, 2500, 1000, min_len=300, error_rate=0.01, contig_names=names)
    compare_matches(ix, oix, rb, ro)
    hits = compare_hits(ix, oix, rb, ro, rn)
    assert 0 < hits["mapped"].sum() < len(hits)      # both mapped and unmapped (ties / no hits) occur
    ix.close()


@pytest.mark.parametrize("k,l,d,hpc", [(8, 16, 0.01, True), (5, 31, 0.01, False), (7, 25, 0.02, True), (2, 31, 0.005, True)])
def test_hits_param_sweep(k, l, d, hpc):
    p = Params(k=k, l=l, density=d, use_hpc=hpc, g=100 if k == 8 else 2000)
    g, go, names = sim.genome(41, [800000, 300000], segdup_frac=0.05)
    ix, oix = build_both(p, names, g, go)
    rb, ro, rn, _ = sim.reads(41, g, go, 1500, 9000, 3000, contig_names=names)
    compare_matches(ix, oix, rb, ro)
    compare_hits(ix, oix, rb, ro, rn)
    ix.close()


def test_edge_reads():
    p = Params()
    g, go, names = sim.genome(51, [500000])
    ix, oix = build_both(p, names, g, go)
    rng = np.random.default_rng(1)
    seqs = [np.zeros(0, np.uint8), random_dna(rng, 10), random_dna(rng, 34), random_dna(rng, 5000),   # unrelated
            g[1000:1035].copy(), g[0:3000].copy(), g[-3000:].copy(), revcomp(g[0:3000]), revcomp(g[-2500:]),
            g[250000:250000 + 40000].copy(),                       # longer than several tiles
            np.concatenate([g[1000:4000], g[400000:403000]]),        # chimeric: two Matches, gap filter
            np.concatenate([g[1000:4000], revcomp(g[5000:8000])])]   # strand switch
    buf, offs = concat_raw(seqs)
    compare_matches(ix, oix, buf, offs)
    compare_hits(ix, oix, buf, offs, [f"r{i}" for i in range(len(seqs))])
    # empty batch
    assert len(ix.map_batch(np.zeros(0, np.uint8), np.zeros(1, np.uint64))) == 0
    ix.close()


def test_segment_partitioned_index_equals_whole():
    # multi-GPU style build: the reference is cut into base-range segments (with halo) and the result
    # must equal the single-shot index
    p = Params()
    g, go, names = sim.genome(61, [700000, 350000])
    g = g.copy(); g[100000:100900] = ord("G"); g[800000:800050] = ord("T")     # long runs for the cuts below
    whole = Index(p); whole.add_batch(names, g, go); nu = whole.freeze()
    part = Index(p)
    rng = np.random.default_rng(2)
    order = []
    for r in range(2):
        a, b = int(go[r]), int(go[r + 1]); L = b - a
        inside_run = [100450, 100451] if r == 0 else [800020 - a]           # cut points inside homopolymer runs
        eq = np.nonzero(g[a + 1:b] == g[a:b - 1])[0][:40:13] + 1             # ... and inside ordinary 2-base runs
        cuts = sorted(set([0, L] + rng.integers(1, L, 5).tolist() + inside_run + eq.tolist()))
        for s, e in zip(cuts[:-1], cuts[1:]):
            order.append((r, L, s, e - s, a))
    for j in rng.permutation(len(order)):       # out of order on purpose
        r, L, s, own, a = order[j]
        lo = a + s - (1 if s > 0 else 0); hi = min(a + L, a + s + own + 4000)
        part.add_segment(r, names[r], L, s, own, g[lo:hi])
    nu2 = part.freeze()
    assert nu == nu2 and whole.n_keys == part.n_keys
    assert np.array_equal(whole.nb_mers(), part.nb_mers())
    rb, ro, _, _ = sim.reads(61, g, go, 800, 8000, 2000)
    h1 = whole.map_batch(rb, ro); h2 = part.map_batch(rb, ro)
    assert h1.tobytes() == h2.tobytes()
    whole.close(); part.close()


def test_errors_and_state():
    from mapquik_b200 import MqError
    with pytest.raises(MqError):
        Index(Params(l=40))
    ix = Index(Params())
    with pytest.raises(MqError):
        ix.map_batch(np.zeros(4, np.uint8), np.array([0, 4], np.uint64))     # not frozen
    ix.freeze({})
    with pytest.raises(MqError):
        ix.add_batch(["x"], np.zeros(4, np.uint8), np.array([0, 4], np.uint64))
    ix.close()


def test_index_save_load_roundtrip(tmp_path):
    # the frozen table written to disk and loaded into a fresh context maps identically
    from mapquik_b200 import MqError
    p = Params(k=4, l=21, density=0.03)
    g, go, names = sim.genome(81, [600000, 250000, 90000])
    ix = Index(p); nb = ix.add_batch(names, g, go); ix.freeze()
    rb, ro, _, _ = sim.reads(81, g, go, 1500, 8000, 2500)
    hits = ix.map_batch(rb, ro)
    path = tmp_path / "idx.mqi"
    ix.save(path)
    ix2 = Index.from_file(path)
    assert ix2.n_unique == ix.n_unique and ix2.n_keys == ix.n_keys
    assert ix2.ref_map == ix.ref_map and np.array_equal(ix2.nb_mers(), nb)
    assert ix2.map_batch(rb, ro).tobytes() == hits.tobytes()
    with pytest.raises(MqError):
        Index.from_file(path, params=Params(k=5, l=21, density=0.03))     # parameter mismatch is refused
    ix.close(); ix2.close()


def test_reads_and_reference_with_non_acgt():
    # N runs and IUPAC codes in both the reference and the reads go through the N-aware scan path
    p = Params()
    g, go, names = sim.genome(91, [500000, 200000])
    g = g.copy()
    rng = np.random.default_rng(91)
    for pos in rng.integers(1000, 690000, 40):
        g[pos:pos + int(rng.integers(1, 300))] = ord("N")
    g[rng.integers(0, 700000, 200)] = np.frombuffer(b"RYKMSWBDHV", np.uint8)[rng.integers(0, 10, 200)]
    ix, oix = build_both(p, names, g, go)
    rb, ro, rn, _ = sim.reads(91, g, go, 1200, 9000, 3000, contig_names=names)
    rb = rb.copy(); rb[rng.integers(0, rb.size, 3000)] = ord("N")
    compare_matches(ix, oix, rb, ro)
    hits = compare_hits(ix, oix, rb, ro, rn)
    assert hits["mapped"].mean() > 0.9
    ix.close()


def _fuzz_record(rng, n):
    kind = rng.integers(0, 6)
    if kind == 0:
        return random_dna(rng, n)
    if kind == 1:                                            # long runs
        base = random_dna(rng, max(1, n // 6 + 1))
        return np.repeat(base, rng.integers(1, 13, base.size))[:n]
    if kind == 2:                                            # short tandem repeat
        unit = random_dna(rng, int(rng.integers(1, 7)))
        return np.tile(unit, n // unit.size + 1)[:n]
    if kind == 3:                                            # sprinkled non-ACGT
        x = random_dna(rng, n).copy()
        if n:
            x[rng.integers(0, n, max(1, n // 50))] = np.frombuffer(b"NRYacgt-", np.uint8)[rng.integers(0, 8, max(1, n // 50))]
        return x
    if kind == 4:                                            # one symbol
        return np.full(n, ord("ACGTN"[int(rng.integers(0, 5))]), np.uint8)
    x = random_dna(rng, n).copy()                            # giant run in the middle
    if n > 10:
        a = int(rng.integers(0, n // 2)); x[a:a + int(rng.integers(1, n - a))] = ord("T")
    return x


@pytest.mark.parametrize("seed", range(24))
def test_minimizers_fuzz_tile_boundaries(seed, fmt):
    # record lengths around every granularity of the tiling (16-byte groups, 128-byte lane chunks, 4,096 / 8,192-base
    # tiles), random start alignment (records are packed back to back), random l / density / HPC
    rng = np.random.default_rng(1000 + seed)
    lens = []
    for _ in range(int(rng.integers(3, 40))):
        base = int(rng.choice([0, 16, 128, 256, 4096, 8192, 12288, 3 * 4096 + 128]))
        lens.append(max(0, base * int(rng.integers(0, 3)) + int(rng.integers(-20, 40))))
    seqs = [_fuzz_record(rng, n) for n in lens]
    buf, offs = concat_raw(seqs)
    l = int(rng.integers(2, 33))
    p = Params(k=int(rng.integers(1, 9)), l=l, density=float(rng.choice([0.005, 0.01, 0.05, 0.3, 1.0])), use_hpc=bool(rng.integers(0, 2)))
    check_minimizers(buf, offs, p, fmt)


@pytest.mark.parametrize("seed", range(8))
def test_hits_fuzz_small_genomes(seed):
    rng = np.random.default_rng(2000 + seed)
    k, l = int(rng.integers(2, 8)), int(rng.integers(8, 33))
    p = Params(k=k, l=l, density=float(rng.choice([0.01, 0.03, 0.08])), use_hpc=bool(rng.integers(0, 2)),
               c=int(rng.integers(0, 6)), s=int(rng.integers(0, 15)), g=int(rng.choice([50, 500, 2000, 100000])))
    if rng.integers(0, 2):
        g, go = repeat_genome(rng, n_contigs=int(rng.integers(1, 9)), units=int(rng.integers(5, 30)), fam=int(rng.integers(1, 5)),
                              fam_len=int(rng.integers(500, 4000)), div=float(rng.choice([0.0, 0.002, 0.01])))
        names = [f"c{i}" for i in range(len(go) - 1)]
    else:
        g, go, names = sim.genome(int(seed) + 300, [int(x) for x in rng.integers(20000, 400000, int(rng.integers(1, 6)))])
    ix, oix = build_both(p, names, g, go)
    rb, ro, rn, _ = sim.reads(int(seed) + 300, g, go, 600, int(rng.integers(800, 12000)), 1500, min_len=100,
                              error_rate=float(rng.choice([0.0, 0.005, 0.03])), contig_names=names)
    compare_matches(ix, oix, rb, ro)
    compare_hits(ix, oix, rb, ro, rn)
    ix.close()


def test_match_quirks_on_crafted_index():
    """The Match::check precedence quirk (match.rs:39-43) on the GPU: a forward Match extends on `offset + 1` alone,
    across reference ids and through reverse-oriented hits.  Real sequence cannot be made to hit that on demand, so
    the reference side is crafted: minimizer stores spliced from the read's own minimizers are imported with
    mq_store_import, and the oracle index gets the same k-min-mers tuple by tuple."""
    rng = np.random.default_rng(4242)
    p = Params(k=5, l=31, density=0.02)
    k, l = p.k, p.l
    read = random_dna(rng, 30000)
    ro = np.array([0, read.size], np.uint64)
    probe = Index(p)
    _, rpos, rh = probe.minimizers(read, ro)
    probe.close()
    M = len(rh)
    assert M > 120
    fill = lambda n: rng.integers(1, 2**62, n).astype(np.uint64)
    recs = []            # per reference record: (hash list, position list)
    # ref0: read minimizers [0, 20+k-1)  -> query windows 0..19 hit ref0 at offsets 0..19 (forward)
    # ref1: 20 fillers + read minimizers [20, 45+k-1) -> windows 20..44 hit ref1 at offsets 20..44: the forward Match
    #       of ref0 keeps extending although the reference id changes
    # ref2: 45 fillers + reversed(read minimizers [45, 45+k)) -> window 45 hits ref2 at offset 45 in REVERSE orientation:
    #       still extends the forward Match (strand is not compared either)
    # ref3: reversed(read minimizers [60, 90+k-1)) -> windows 60..89 hit with decreasing offsets: a genuine rc Match
    # ref4: 7 fillers + reversed(read minimizers [100, 100+k)) ; ref5: reversed(read minimizers [101, 101+k)) -> two rc
    #       hits with offsets 7 and 0 on different references: an rc Match must NOT extend across references
    recs.append(rh[0:20 + k - 1])
    recs.append(np.concatenate([fill(20), rh[20:45 + k - 1]]))
    recs.append(np.concatenate([fill(45), rh[45:45 + k][::-1]]))
    recs.append(rh[60:90 + k - 1][::-1].copy())
    recs.append(np.concatenate([fill(7), rh[100:100 + k][::-1]]))
    recs.append(rh[101:101 + k][::-1].copy())
    ix = Index(p)
    oix = O.Index(oparams(p))
    pos_all, hash_all, directory, lens = [], [], [], []
    for rid, hs in enumerate(recs):
        pos = (np.arange(len(hs), dtype=np.uint32) * 97 + 11)
        pos_all.append(pos); hash_all.append(hs); directory.append((rid, 0, len(hs)))
        lens.append(int(pos[-1]) + 5000)
        for j in range(len(hs) - k + 1):
            kh, rev = O.kminmer_hash(hs[j:j + k])
            oix.add_tuple(kh, rid, int(pos[j]), int(pos[j + k - 1]) + l, j, rev)
    pos_all = np.ascontiguousarray(np.concatenate(pos_all)); hash_all = np.ascontiguousarray(np.concatenate(hash_all))
    ix.store_import(pos_all.ctypes.data, hash_all.ctypes.data, len(pos_all), np.array(directory, np.uint64))
    ix.freeze({i: (f"ref{i}", lens[i]) for i in range(len(recs))})
    oix.ref_names = [f"ref{i}" for i in range(len(recs))]; oix.ref_lens = lens
    assert ix.n_unique == oix.count() and ix.n_keys == oix.slots()
    compare_matches(ix, oix, read, ro)
    compare_hits(ix, oix, read, ro, ["crafted"])
    mo, f = ix.matches(read, ro)
    counts = f[:, 4].tolist(); refrc = f[:, 5].tolist()
    assert 46 in counts                                   # windows 0..45 form ONE forward Match across ref0/ref1/ref2 (the quirk)
    assert refrc[counts.index(46)] == (0 << 1 | 0)        # filed under the head's reference id, forward
    assert 30 in counts and refrc[counts.index(30)] == (3 << 1 | 1)      # the genuine rc Match on ref3
    assert counts.count(1) >= 2                           # windows 100 and 101: rc hits on different references stay apart
    ix.close()


def test_segment_partition_tiny_and_odd_contigs():
    # the multi-GPU index build on awkward references: contigs shorter than l+k-1, shorter than the rank count, one
    # base, homopolymers -- cut into 8 ranges exactly like mapquik_b200.shard does -- must equal the single-shot index
    from mapquik_b200 import shard
    rng = np.random.default_rng(77)
    p = Params()
    contigs = [random_dna(rng, n) for n in (5, 1, 34, 35, 36, 200, 1000, 4097, 50000)]
    contigs.append(np.full(3000, ord("A"), np.uint8))
    contigs.append(np.concatenate([random_dna(rng, 700), np.full(900, ord("C"), np.uint8), random_dna(rng, 800)]))
    buf, offs = concat_raw(contigs)
    names = [f"c{i}" for i in range(len(contigs))]
    whole = Index(p); nb = whole.add_batch(names, buf, offs); whole.freeze()
    world = 8
    part = Index(p)
    for rank in range(world):
        for r, seq in enumerate(contigs):
            s, own, data = shard.segment_for_rank(seq, rank, world, p.l)
            part.add_segment(r, names[r], len(seq), s, own, data if len(data) else np.zeros(1, np.uint8))
    part.freeze()
    assert whole.n_unique == part.n_unique and whole.n_keys == part.n_keys
    assert np.array_equal(nb, part.nb_mers())
    rb, ro, _, _ = sim.reads(77, buf, offs, 300, 3000, 1000, min_len=200)
    assert whole.map_batch(rb, ro).tobytes() == part.map_batch(rb, ro).tobytes()
    whole.close(); part.close()


# ---- packed input: the same results as the ASCII path and the oracle, through every packed entry point ------------------
def test_packed_index_and_hits_equal_ascii_and_oracle():
    p = Params()
    g, go, names = sim.genome(101, [900000, 400000, 30])
    g = g.copy(); g[5000:5600] = ord("N"); g[700000] = ord("R")
    ix, oix = build_both(p, names, g, go)
    pg = PackedSeqs(g)
    ixp = Index(p)
    nbp = ixp.add_batch_packed(names, pg, go)
    assert ixp.freeze() == ix.n_unique and ixp.n_keys == ix.n_keys
    assert np.array_equal(nbp, ix.nb_mers())
    rb, ro, rn, _ = sim.reads(101, g, go, 2500, 9000, 3000, contig_names=names)
    rb = rb.copy(); rb[np.random.default_rng(1).integers(0, rb.size, 500)] = ord("N")
    pr = PackedSeqs(rb, pinned=True)
    hits = compare_hits(ix, oix, rb, ro, rn)
    for index in (ix, ixp):
        assert index.map_batch_packed(pr, ro).tobytes() == hits.tobytes()
    assert ixp.map_batch(rb, ro).tobytes() == hits.tobytes()
    pr.close(); ix.close(); ixp.close()


def test_packed_split_invariance_many_sub_batches():
    # more bases than one pipelined sub-batch (128 Mbases): slices that start off any 2048-base boundary, exception
    # intervals cut by slice edges, results independent of how the batch is cut
    p = Params()
    g, go, names = sim.genome(102, [4000000])
    ix = Index(p); ix.add_batch(names, g, go); ix.freeze()
    rb, ro, _, _ = sim.reads(102, g, go, 16000, 10000, 2000)
    rb = rb.copy()
    rng = np.random.default_rng(2)
    for s in rng.integers(0, rb.size - 3000, 300):
        rb[s:s + int(rng.integers(1, 2500))] = ord("N")
    assert rb.size > (128 << 20)
    pr = PackedSeqs(rb)
    whole = ix.map_batch_packed(pr, ro)
    assert ix.map_batch(rb, ro).tobytes() == whole.tobytes()
    lo, hi = 777, 9001
    sub = PackedSeqs(rb[int(ro[lo]):int(ro[hi])])
    assert ix.map_batch_packed(sub, ro[lo:hi + 1] - ro[lo]).tobytes() == whole[lo:hi].tobytes()
    from oracle import pyoracle as O2
    oix = O2.Index(oparams(p), 1 << 16); oix.add_batch(names, g, go)
    assert oix.map_batch(rb[:int(ro[3000])], ro[:3001]).tobytes() == whole[:3000].tobytes()
    ix.close()


def test_device_resident_entry_points_equal_host_ones():
    import ctypes as C
    from mapquik_b200 import capi, HIT_DTYPE
    L = capi.lib()
    p = Params()
    g, go, names = sim.genome(103, [1500000, 200000])
    ix = Index(p); ix.add_batch(names, g, go); ix.freeze()
    rb, ro, _, _ = sim.reads(103, g, go, 3000, 9000, 2500)
    rb = rb.copy(); rb[100:140] = ord("N"); rb[-5] = ord("Y")
    ref = ix.map_batch(rb, ro)
    h = ix.handle; n = len(ro) - 1
    d_hits = L.mq_dev_alloc(h, n * 48)
    # ASCII resident
    d_seqs = L.mq_dev_alloc(h, rb.size + 256)
    L.mq_dev_memset(h, d_seqs, 0, rb.size + 256); assert L.mq_dev_upload(h, d_seqs, rb.ctypes.data, rb.size) == 0
    ix.map_batch_device(d_seqs, ro, d_hits)
    out = np.zeros(n, HIT_DTYPE); assert L.mq_dev_download(h, out.ctypes.data, d_hits, n * 48) == 0
    assert out.tobytes() == ref.tobytes()
    # packed resident
    pr = PackedSeqs(rb)
    d_w = L.mq_dev_alloc(h, pr.words.nbytes); d_f = L.mq_dev_alloc(h, pr.flags.nbytes); d_e = L.mq_dev_alloc(h, max(pr.exc.nbytes, 16))
    assert L.mq_dev_upload(h, d_w, pr.words.ctypes.data, pr.words.nbytes) == 0
    assert L.mq_dev_upload(h, d_f, pr.flags.ctypes.data, pr.flags.nbytes) == 0
    assert L.mq_dev_upload(h, d_e, pr.exc.ctypes.data, pr.exc.nbytes) == 0
    L.mq_dev_memset(h, d_hits, 0, n * 48)
    ix.map_batch_packed_device(d_w, d_f, d_e, pr.exc.size, pr.n_bases, ro, d_hits)
    assert L.mq_dev_download(h, out.ctypes.data, d_hits, n * 48) == 0
    assert out.tobytes() == ref.tobytes()
    for d in (d_hits, d_seqs, d_w, d_f, d_e):
        L.mq_dev_free(h, d)
    ix.close()


# ---- one context over several GPUs (mq_create_multi) ---------------------------------------------------------------------
def _device_ids(n):
    import torch
    have = torch.cuda.device_count()
    return [i % have for i in range(n)]         # on a one-GPU box the same device serves twice: same code path, same answers


@pytest.mark.parametrize("n_dev", [2, 3])
def test_multi_gpu_context_equals_single(n_dev):
    p = Params()
    g, go, names = sim.genome(104, [1200000, 500000, 77, 20000])
    g = g.copy(); g[300000:300800] = ord("A"); g[900000:900050] = ord("N")
    one = Index(p); nb1 = one.add_batch(names, g, go); one.freeze()
    multi = Index(p, devices=_device_ids(n_dev))
    nbm = multi.add_batch(names, g, go)
    assert multi.freeze() == one.n_unique and multi.n_keys == one.n_keys
    assert np.array_equal(nb1, nbm) and np.array_equal(one.nb_mers(), multi.nb_mers())
    rb, ro, _, _ = sim.reads(104, g, go, 5000, 9000, 3000)
    h1 = one.map_batch(rb, ro)
    assert multi.map_batch(rb, ro).tobytes() == h1.tobytes()
    pr = PackedSeqs(rb)
    assert multi.map_batch_packed(pr, ro).tobytes() == h1.tobytes()
    # packed reference through the multi-GPU build as well
    multi2 = Index(p, devices=_device_ids(n_dev))
    multi2.add_batch_packed(names, PackedSeqs(g), go); multi2.freeze()
    assert multi2.n_unique == one.n_unique and multi2.map_batch(rb, ro).tobytes() == h1.tobytes()
    one.close(); multi.close(); multi2.close()


# ---- on-the-fly packing of ASCII host input (mq_set_host_threads) -------------------------------------------------
@pytest.mark.parametrize("threads,n_dev", [(1, 1), (3, 1), (16, 1), (5, 2)])
def test_host_threads_pack_on_the_fly_equals_plain(threads, n_dev, monkeypatch):
    """mq_map_batch on ASCII with host threads: sub-batches packed on the host from the back of the batch, ASCII ones
    uploaded from the front -- same hits as the plain path and as the oracle, whatever route a read took"""
    monkeypatch.setenv("MQ_SUB_BASES", str(1 << 20))          # 1 Mbase sub-batches: ~45 of them from a small input
    p = Params()
    g, go, names = sim.genome(211, [900000, 400000])
    g = g.copy(); g[200000:200300] = ord("N")
    ix, oix = build_both(p, names, g, go)
    rb, ro, rn, _ = sim.reads(211, g, go, 5000, 9000, 3000, contig_names=names)
    rb = rb.copy()
    rng = np.random.default_rng(7)
    for s0 in rng.integers(0, rb.size - 100, 300):          # N runs, IUPAC codes, lower case, bytes >= 0x80 inside reads
        rb[s0:s0 + int(rng.integers(1, 60))] = int(rng.choice(np.frombuffer(b"NNNRYacgt\xc1\xff", np.uint8)))
    want = compare_hits(ix, oix, rb, ro, rn)
    dev = Index(p, devices=_device_ids(n_dev)) if n_dev > 1 else ix
    if n_dev > 1:
        dev.add_batch(names, g, go); dev.freeze()
    dev.set_host_threads(threads)
    for _ in range(3):                                        # the split between the two routes differs from call to call
        got = dev.map_batch(rb, ro)
        assert got.tobytes() == want.tobytes()
        assert dev.last_counter("sub_batches") >= 40
        assert 1 <= dev.last_counter("host_packed_sub_batches") < dev.last_counter("sub_batches")
        assert dev.last_counter("h2d_bytes") < rb.size + 64 * ro.size       # some of it crossed the link at 2 bits per base
    dev.set_host_threads(0)
    assert dev.map_batch(rb, ro).tobytes() == want.tobytes() and dev.last_counter("host_packed_sub_batches") == 0
    assert dev.last_counter("h2d_bytes") >= rb.size
    if n_dev > 1:
        dev.close()
    ix.close()


# ---- BASELINE configs 2-5 at a scale the oracle finishes in seconds (the full sizes run in bench.py / scripts) -----------
def _config3_like(scale, seed=3):
    lens = [int(3.1e9 / scale * x / sum(sim.CHM13_PROPS)) for x in sim.CHM13_PROPS]
    return sim.genome(seed, lens, sat_frac=0.06, segdup_frac=0.05)


def test_config3_scaled_human_like_genome():
    p = Params()
    g, go, names = _config3_like(10)                        # 310 Mbp, 24 contigs, satellites + segmental duplications
    ix, oix = build_both(p, names, g, go)
    rb, ro, rn, tr = sim.reads(3, g, go, 20000, 24000, 3000, contig_names=names)
    hits = compare_hits(ix, oix, rb, ro, rn)
    assert ix.map_batch_packed(PackedSeqs(rb), ro).tobytes() == hits.tobytes()
    ok = (hits["mapped"] == 1) & (hits["ref_idx"] == tr["contig"]) & (hits["rc"] == tr["strand"])
    assert ok.mean() > 0.97 and ((hits["mapq"] == 60) & ~ok).sum() <= 2
    ix.close()


def test_config4_scaled_repeat_family_genome():
    p = Params()
    g, go, names = sim.genome(4, [22000000] * 10, repeat_frac=0.85, n_families=300)      # 220 Mbp, 85 % repeats
    ix, oix = build_both(p, names, g, go)
    assert ix.n_keys > 1.015 * ix.n_unique                   # many tombstones
    rb, ro, rn, _ = sim.reads(4, g, go, 10000, 24000, 3000, contig_names=names)
    hits = compare_hits(ix, oix, rb, ro, rn)
    assert ix.map_batch_packed(PackedSeqs(rb), ro).tobytes() == hits.tobytes()
    ix.close()


@pytest.mark.parametrize("k,l,d", [(3, 25, 0.02), (7, 31, 0.005), (5, 28, 0.01)])
def test_config5_grid_points(k, l, d):
    p = Params(k=k, l=l, density=d)
    g, go, names = _config3_like(40)                        # 77 Mbp
    ix, oix = build_both(p, names, g, go)
    rb, ro, rn, _ = sim.reads(5, g, go, 4000, 24000, 3000, contig_names=names)
    compare_hits(ix, oix, rb, ro, rn)
    ix.close()
