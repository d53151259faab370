"""CPU tests: the oracle against known answers, the reference's golden PAF line, a second
independent restatement (tests/pyref.py), and hand-built cases for every documented quirk."""
import numpy as np
import pytest

import pyref
from conftest import random_dna, revcomp
from oracle import pyoracle as O

GOLDEN_PAF = ("S1_1!chr1!224752794!224777027!+\t24299\t0\t24298\t+\tchr1\t248387328\t224752793\t224777027\t132\t"
              "248387328\t60")   # experiments/intersect_pafs.py:14 (whitespace = tabs, mers.rs:181)


# ---- S1 ---------------------------------------------------------------------------------------------
def test_nthash_known_answers():
    # vectors of the `nthash` crate's own test-suite (ntf64 / ntr64 / ntc64 on "TGCAG"); not present
    # in /root/reference -- they pin the seeds, the rotation directions and canonical = min
    L = O.lib()
    assert L.orc_nthash_fwd(b"TGCAG", 5) == 0x0bafa6728fc6dabf
    assert L.orc_nthash_rev(b"TGCAG", 5) == 0x8cf2d4072cca480e
    assert min(L.orc_nthash_fwd(b"ACGTC", 5), L.orc_nthash_rev(b"ACGTC", 5)) == 0x480202d54e8ebecd
    # forward hash of a sequence == reverse hash of its reverse complement
    rng = np.random.default_rng(0)
    for l in (2, 15, 31, 32, 64, 70):
        s = random_dna(rng, l)
        assert L.orc_nthash_fwd(s.tobytes(), l) == L.orc_nthash_rev(revcomp(s).tobytes(), l)
        assert L.orc_nthash_fwd(s.tobytes(), l) == pyref.ntf(list(s))
        assert L.orc_nthash_rev(s.tobytes(), l) == pyref.ntr(list(s))
    # N (and anything that is not ACGT) hashes as 0
    assert L.orc_nthash_fwd(b"N", 1) == 0 and L.orc_nthash_fwd(b"a", 1) == 0


def test_hash_bound():
    assert O.lib().orc_hash_bound(0.01) == 0x28f5c28f5c28f60 == pyref.hash_bound(0.01)
    assert O.lib().orc_hash_bound(0.0) == 0
    assert O.lib().orc_hash_bound(1.0) == 2**64 - 1          # Rust `as` saturates
    assert O.lib().orc_hash_bound(0.5) == 2**63


@pytest.mark.parametrize("l,density,hpc", [(31, 0.01, True), (16, 0.05, True), (5, 0.2, False), (2, 0.5, True), (40, 0.1, True)])
def test_minimizers_three_forms_agree(l, density, hpc):
    rng = np.random.default_rng(1)
    seqs = [random_dna(rng, 5000), np.frombuffer(b"A" * 300, np.uint8), np.frombuffer(b"AC" * 400, np.uint8),
            np.repeat(random_dna(rng, 400), rng.integers(1, 9, 400)), random_dna(rng, l), random_dna(rng, l - 1),
            np.zeros(0, np.uint8)]
    x = random_dna(rng, 2000).copy(); x[500:520] = ord("N"); x[900] = ord("n"); seqs.append(x)
    p = O.params(5, l, density, hpc)
    for s in seqs:
        a = O.minimizers(s, p)
        b = O.minimizers(s, p, slow=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        ref = pyref.minimizers(s.tobytes(), l, density, hpc)
        assert [int(v) for v in a[0]] == [r[0] for r in ref]
        assert [int(v) for v in a[1]] == [r[1] for r in ref]


def test_density_and_hpc_coordinates():
    rng = np.random.default_rng(2)
    s = random_dna(rng, 400000)
    pos, hs = O.minimizers(s, O.params())
    # canonical = min(fwd, rev) below density*max => P ~ 2d per compressed l-mer, 0.75 compression
    assert 0.0135 < len(pos) / len(s) < 0.0165
    assert np.all(np.diff(pos.astype(np.int64)) > 0)
    assert np.all(hs < 0x28f5c28f5c28f60)
    # positions are raw indices of run starts
    assert np.all((pos == 0) | (s[pos.astype(np.int64)] != s[pos.astype(np.int64) - 1]))


# ---- S2 ---------------------------------------------------------------------------------------------
def test_kminmers_against_pyref():
    rng = np.random.default_rng(3)
    s = random_dna(rng, 60000)
    for k, l, d in ((5, 31, 0.01), (8, 16, 0.02), (1, 20, 0.05), (3, 12, 0.1)):
        km = O.kminmers(s, O.params(k, l, d))
        ref = pyref.kminmers(s.tobytes(), k, l, d)
        assert len(km) == len(ref) > 0
        for a, b in zip(km, ref):
            assert (int(a["start"]), int(a["end"]), int(a["offset"]), int(a["rev"]), int(a["hash"])) == \
                (b.start, b.end, b.offset, int(b.rev), b.hash)
    # too-short record: mers.rs:18,44
    assert len(O.kminmers(random_dna(rng, 34), O.params())) == 0


def test_kminmer_canonical_orientation():
    h, rev = O.kminmer_hash([5, 1, 9])
    h2, rev2 = O.kminmer_hash([9, 1, 5])
    assert h == h2 and rev == 0 and rev2 == 1            # same canonical vector, opposite flags
    assert O.kminmer_hash([7, 3, 7]) == (pyref.kminmer_hash([7, 3, 7])[0], 0)   # palindrome => forward
    assert O.kminmer_hash([4, 8, 15, 16])[0] != O.kminmer_hash([4, 8, 16, 15])[0]
    # a sequence and its reverse complement give the same k-min-mer hashes with flipped rev (no HPC
    # so that run starts are strand-symmetric)
    rng = np.random.default_rng(4)
    s = random_dna(rng, 30000)
    p = O.params(5, 21, 0.02, use_hpc=False)
    a, b = O.kminmers(s, p), O.kminmers(revcomp(s), p)
    assert len(a) == len(b) > 50
    assert np.array_equal(a["hash"], b["hash"][::-1])
    nonpal = a["hash"] != 0
    assert np.array_equal(a["rev"][nonpal], 1 - b["rev"][::-1][nonpal])


# ---- index ------------------------------------------------------------------------------------------
def test_index_unique_or_tombstone_order_independent():
    p = O.params()
    tuples = [(11, 0, 10, 60, 0, 0), (22, 0, 70, 130, 1, 1), (11, 1, 500, 560, 7, 0), (33, 1, 5, 65, 0, 1),
              (22, 1, 9, 69, 3, 0), (22, 0, 300, 360, 9, 0), (44, 0, 900, 960, 12, 1)]
    rng = np.random.default_rng(5)
    states = set()
    for _ in range(6):
        ix = O.Index(p)
        for i in rng.permutation(len(tuples)):
            ix.add_tuple(*tuples[i])
        assert ix.count() == 2 and ix.slots() == 4          # 33 and 44 unique; 11, 22 tombstones (stay in the map)
        states.add((ix.get(11), ix.get(22), ix.get(33), ix.get(44), ix.get(55)))
    assert states == {(None, None, (1, 5, 65, 0, 1), (0, 900, 960, 12, 1), None)}


# ---- Match / Chain rules on crafted inputs ----------------------------------------------------------------
def km(start, end, offset, h, rev=0):
    a = np.zeros(1, O.KM_DTYPE)
    a[0] = (start, end, offset, h, rev, 0)
    return a[0]


def crafted_index(entries):
    ix = O.Index(O.params())
    for h, rid, s, e, off, rc in entries:
        ix.add_tuple(h, rid, s, e, off, rc)
    ix.ref_names = ["r0", "r1", "r2"]; ix.ref_lens = [100000, 100000, 100000]
    return ix


def pyref_index(entries):
    ix = pyref.Index()
    for h, rid, s, e, off, rc in entries:
        k = pyref.Kminmer(); k.hash, k.start, k.end, k.offset, k.rev = h, s, e, off, bool(rc)
        ix.add_with_mer(rid, k)
    return ix


def pyref_kms(kms):
    out = []
    for a in kms:
        k = pyref.Kminmer()
        k.start, k.end, k.offset, k.hash, k.rev = int(a["start"]), int(a["end"]), int(a["offset"]), int(a["hash"]), bool(a["rev"])
        out.append(k)
    return out


def both_chain(entries, kms):
    ix = crafted_index(entries)
    kms = np.array(kms, dtype=O.KM_DTYPE)
    got = ix.chain_matches_kms(kms)
    ref = pyref.chain_matches(pyref_kms(kms), pyref_index(entries))
    flat = sorted(((m.q_start, m.q_end, m.r_start, m.r_end, m.count, int(m.rc), rid) for rid, ms in ref.items() for m in ms))
    mine = sorted((int(m["q_start"]), int(m["q_end"]), int(m["r_start"]), int(m["r_end"]), int(m["count"]), int(m["rc"]),
                   int(m["ref_id"])) for m in got)
    assert flat == mine
    return got, ix, kms


def test_match_check_precedence_quirk():
    # match.rs:39-43: for a forward Match only `r.offset - p.offset == 1` is tested -- reference id
    # and strand are NOT.  Entry 2 sits on another reference and the other strand yet extends.
    entries = [(1, 0, 100, 160, 4, 0), (2, 1, 9000, 9060, 5, 1), (3, 0, 300, 360, 9, 0)]
    got, _, _ = both_chain(entries, [km(10, 70, 0, 1), km(40, 100, 1, 2), km(80, 140, 2, 3)])
    assert len(got) == 2
    assert tuple(int(got[0][f]) for f in ("q_start", "q_end", "r_start", "r_end", "count", "rc", "ref_id")) == \
        (10, 100, 100, 9060, 2, 0, 0)              # r_end taken from the foreign entry, filed under ref 0
    assert int(got[1]["count"]) == 1 and int(got[1]["r_start"]) == 300


def test_match_rc_extension_needs_same_ref_and_strand():
    # rc Match (q.rev != r.rc): extends only on same ref, rc hit, p.offset - r.offset == 1
    e = [(1, 0, 500, 560, 8, 1), (2, 0, 450, 510, 7, 1), (3, 1, 400, 460, 6, 1), (4, 0, 380, 440, 5, 0)]
    got, _, _ = both_chain(e, [km(0, 60, 0, 1), km(30, 90, 1, 2), km(60, 120, 2, 3), km(90, 150, 3, 4)])
    # 1+2 chain (rc, offsets 8 -> 7); 3 is on ref 1 => new Match; 4 has rc == False => new (forward) Match
    assert [int(m["count"]) for m in got] == [2, 1, 1]
    assert (int(got[0]["r_start"]), int(got[0]["r_end"]), int(got[0]["rc"])) == (450, 560, 1)


def test_miss_consumes_failed_check_reprobes():
    # match.rs:45-58: a miss after a Match is consumed; a hit that fails `check` is re-probed and
    # opens the next Match
    e = [(1, 0, 100, 160, 0, 0), (3, 0, 300, 360, 5, 0), (4, 0, 340, 400, 6, 0)]
    got, _, _ = both_chain(e, [km(0, 60, 0, 1), km(20, 80, 1, 99), km(40, 100, 2, 3), km(60, 120, 3, 4)])
    assert [(int(m["q_start"]), int(m["count"])) for m in got] == [(0, 1), (40, 2)]
    got, _, _ = both_chain(e, [km(0, 60, 0, 1), km(40, 100, 1, 3), km(60, 120, 2, 4)])
    assert [(int(m["q_start"]), int(m["count"])) for m in got] == [(0, 1), (40, 2)]


def test_offsets_compare_as_i32():
    # offsets are cast `as i32` before subtracting: 2^32 + 6 and 5 are "consecutive"
    e = [(1, 0, 100, 160, 5, 0), (2, 0, 140, 200, 2**32 + 6, 0)]
    got, _, _ = both_chain(e, [km(0, 60, 0, 1), km(30, 90, 1, 2)])
    assert len(got) == 1 and int(got[0]["count"]) == 2


def mk_match(qs, qe, rs, re, cnt, rc, ref):
    a = np.zeros(1, O.MATCH_DTYPE); a[0] = (qs, qe, rs, re, cnt, rc, ref); return a[0]


def both_best(matches, q_len, ref_lens, p=None, c=4, s=11, g=2000):
    p = p or O.params(c=c, s=s, g=g)
    hit = O.best_of_matches(np.array(matches, dtype=O.MATCH_DTYPE), q_len, ref_lens, p)
    per_ref = {}
    for m in matches:
        pm = pyref.Match.__new__(pyref.Match)
        pm.q_start, pm.q_end, pm.r_start, pm.r_end, pm.count = (int(m[f]) for f in ("q_start", "q_end", "r_start", "r_end", "count"))
        pm.rc = bool(m["rc"])
        per_ref.setdefault(int(m["ref_id"]), []).append(pm)
    allc = [(rid, pyref.get_match(ms, p.c, p.s, p.g)) for rid, ms in per_ref.items()]
    return hit, allc


def test_chain_filter_keeps_only_matches_compatible_with_largest():
    ms = [mk_match(0, 500, 1000, 1500, 3, 0, 0),          # compatible, before the largest
          mk_match(600, 3000, 1600, 4000, 20, 0, 0),      # largest
          mk_match(3100, 3500, 90000, 90400, 5, 0, 0),    # gap difference >> g: dropped
          mk_match(3600, 3900, 4600, 4900, 2, 1, 0),      # other strand: dropped
          mk_match(4000, 4400, 1200, 1600, 4, 0, 0)]      # r_start not increasing: dropped
    hit, allc = both_best(ms, 5000, [100000])
    assert int(hit["score"]) == 23 and int(hit["mapq"]) == 60 and int(hit["rc"]) == 0
    assert allc[0][1][5] == 23
    # PseudoChainCoords = first.q_start, last.q_end-1, first.r_start, last.r_end-1, then find_coords
    assert (int(hit["q_start"]), int(hit["q_end"]), int(hit["r_start"]), int(hit["r_end"])) == \
        O.find_coords(5000, 100000, 0, 0, 2999, 1000, 3999)


def test_chain_first_largest_wins_and_rc_coordinates():
    ms = [mk_match(0, 1000, 8000, 9000, 7, 1, 0), mk_match(1100, 2100, 6900, 7900, 7, 1, 0)]   # equal counts: first is "largest"
    hit, allc = both_best(ms, 2200, [50000])
    assert int(hit["score"]) == 14 and int(hit["rc"]) == 1
    # rc && len > 1: r_start = last.r_start, r_end = first.r_end - 1 (chain.rs:166)
    assert allc[0][1][1:5] == (0, 2099, 6900, 8999)


def test_mapq_rule():
    one = [mk_match(0, 900, 100, 1000, 10, 0, 0)]
    assert int(both_best(one, 1000, [5000])[0]["mapq"]) == 0                      # len 1 < c=4 and score 10 < s=11
    assert int(both_best([mk_match(0, 900, 100, 1000, 11, 0, 0)], 1000, [5000])[0]["mapq"]) == 60
    four = [mk_match(200 * i, 200 * i + 150, 1000 + 200 * i, 1150 + 200 * i, 1, 0, 0) for i in range(4)]
    assert int(both_best(four, 1000, [5000])[0]["mapq"]) == 60                    # len_f >= c
    assert int(both_best(four, 1000, [5000], c=0)[0]["mapq"]) == 0                # c == 0 disables (chain.rs:158)
    assert int(both_best(four, 1000, [5000], s=0)[0]["mapq"]) == 0


def test_tie_between_references_is_unmapped():
    a = mk_match(0, 900, 100, 1000, 12, 0, 0); b = mk_match(0, 900, 700, 1600, 12, 0, 1); c = mk_match(0, 900, 50, 950, 3, 0, 2)
    assert int(both_best([a, b, c], 1000, [5000, 5000, 5000])[0]["mapped"]) == 0      # mers.rs:106
    b2 = mk_match(0, 900, 700, 1600, 13, 0, 1)
    hit, _ = both_best([a, b2, c], 1000, [5000, 5000, 5000])
    assert int(hit["mapped"]) == 1 and int(hit["ref_idx"]) == 1
    assert int(both_best([c, a, c, b2], 1000, [5000, 5000, 5000])[0]["ref_idx"]) == 1   # order independent


def test_gap_filter_i32_truncation():
    # chain.rs:132-142 casts each coordinate `as i32` before subtracting: coordinates 2^32 apart look equal
    u = mk_match(0, 500, 1000, 1500, 9, 0, 0)
    v_far = mk_match(600, 1100, 1600 + 2**32, 2100 + 2**32, 3, 0, 0)
    hit, allc = both_best([u, v_far], 1200, [2**33])
    assert int(hit["score"]) == 12 == allc[0][1][5]
    assert pyref.gap_too_long(600, 500, 1600 + 2**32, 1500, 2000) is False
    assert pyref.gap_too_long(600, 500, 5000, 1500, 2000) is True


@pytest.mark.parametrize("rc", [0, 1])
def test_find_coords_clipping(rc):
    for (ql, rl, qs, qe, rs, re) in [(1000, 50000, 100, 899, 5000, 5799),     # interior
                                     (1000, 50000, 100, 899, 40, 839),        # clipped at reference start
                                     (1000, 50000, 100, 899, 49100, 49899),   # clipped at reference end
                                     (1000, 900, 0, 999, 0, 899),             # read longer than reference
                                     (1000, 50000, 0, 999, 7, 1006)]:
        got = O.find_coords(ql, rl, rc, qs, qe, rs, re)
        line = pyref.find_coords("q", ql, "r", rl, (bool(rc), qs, qe, rs, re, 1, 0)).split("\t")
        assert got == (int(line[2]), int(line[3]), int(line[7]), int(line[8]))
        assert got[3] <= rl - 1


def test_golden_paf_line():
    # the reference's only golden record: layout, inclusive ends, column 11 == r_len, MAPQ 60
    hit = np.zeros(1, O.HIT_DTYPE)[0]
    ql, rl = 24299, 248387328
    # a forward pseudo-chain 13 bases in from both read ends extends to the full read (mers.rs:142-158)
    fq_s, fq_e, fr_s, fr_e = O.find_coords(ql, rl, 0, 13, 24285, 224752806, 224777014)
    assert (fq_s, fq_e, fr_s, fr_e) == (0, 24298, 224752793, 224777027)
    ix = O.Index(O.params()); ix.ref_names = ["chr1"]; ix.ref_lens = [rl]
    hit["mapped"], hit["rc"], hit["mapq"], hit["ref_idx"] = 1, 0, 60, 0
    hit["q_start"], hit["q_end"], hit["r_start"], hit["r_end"], hit["score"] = fq_s, fq_e, fr_s, fr_e, 132
    assert ix.paf_line("S1_1!chr1!224752794!224777027!+", ql, hit) == GOLDEN_PAF
    assert pyref.find_coords("S1_1!chr1!224752794!224777027!+", ql, "chr1", rl,
                             (False, 13, 24285, 224752806, 224777014, 132, 60)) == GOLDEN_PAF


# ---- end to end: C oracle == Python restatement -----------------------------------------------------------
def end_to_end(seed, k, l, d, hpc, g, genome_fn, n_reads):
    from mapquik_b200 import sim
    gb, go, names = genome_fn(seed)
    p = O.params(k, l, d, hpc, g=g)
    ix = O.Index(p)
    pix = pyref.Index()
    ref_map = {}
    for i in range(len(names)):
        s = gb[int(go[i]):int(go[i + 1])]
        nb = ix.add_ref(names[i], s)
        kms = pyref.kminmers(s.tobytes(), k, l, d, hpc)
        assert nb == len(kms)
        for m in kms:
            pix.add_with_mer(i, m)
        ref_map[i] = (names[i], len(s))
    assert ix.count() == pix.get_count()
    rb, ro, rn, _ = sim.reads(seed, gb, go, n_reads, 3000, 1000, min_len=200, error_rate=0.01, contig_names=names)
    n_mapped = 0
    for i in range(n_reads):
        s = rb[int(ro[i]):int(ro[i + 1])]
        hit = ix.find_matches(s)
        line = ix.paf_line(rn[i], len(s), hit) if hit["mapped"] else None
        ref = pyref.find_matches(rn[i], s.tobytes(), ref_map, pix, k, l, d, hpc, p.c, p.s, g)
        assert line == ref, (i, line, ref)
        n_mapped += line is not None
    return n_mapped


def test_end_to_end_plain_genome():
    from mapquik_b200 import sim
    n = end_to_end(7, 5, 21, 0.03, True, 2000, lambda s: sim.genome(s, [60000, 30000]), 40)
    assert n >= 38


def test_end_to_end_repetitive_genome():
    import test_gpu_parity as T
    rng = np.random.default_rng(8)

    def gen(seed):
        g, go = T.repeat_genome(rng, n_contigs=4, units=12, fam=3, fam_len=1500)
        return g, go, [f"c{i}" for i in range(4)]
    n = end_to_end(8, 3, 15, 0.04, True, 500, gen, 60)
    assert 0 < n


def test_batch_threads_deterministic():
    from mapquik_b200 import sim
    g, go, names = sim.genome(9, [300000, 100000])
    rb, ro, _, _ = sim.reads(9, g, go, 300, 8000, 2000)
    a = O.Index(O.params()); a.add_batch(names, g, go, threads=4)
    b = O.Index(O.params()); b.add_batch(names, g, go, threads=1)
    assert a.count() == b.count() and a.slots() == b.slots()
    assert a.map_batch(rb, ro, threads=4).tobytes() == b.map_batch(rb, ro, threads=1).tobytes()
