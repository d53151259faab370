"""Exhaustive CPU checks of the bit tricks the scan kernel relies on (mapquik_b200/csrc/mq_scan_v2.cuh, mq_scan_v3.cuh).

The device code cannot run here; these are numpy restatements of the exact integer expressions, checked against the
plain definition over their whole input domain, so that an edit to a constant shows up before it reaches a GPU."""
import numpy as np

U32 = np.uint32
M32 = 0xFFFFFFFF


def words_of(b0, b1, b2, b3):
    return (np.asarray(b0, np.uint64) | (np.asarray(b1, np.uint64) << 8) | (np.asarray(b2, np.uint64) << 16)
            | (np.asarray(b3, np.uint64) << 24)).astype(np.uint64)


def acgt_diff(u):
    """v2_acgt_diff: 0 in every byte that is 'A', 'C', 'G' or 'T'"""
    m = (u >> 2) & ~(u >> 1) & 0x01010101
    return ((u & 0xF9F9F9F9) ^ ((m * 0x0F + 0x41414141) & M32)) & M32


def nonzero80(e):
    """0x80 in every non-zero byte"""
    return ((((e & 0x7F7F7F7F) + 0x7F7F7F7F) | e) & 0x80808080) & M32


def test_acgt_validity_every_byte_value_in_every_lane():
    valid = {ord(c) for c in "ACGT"}
    rng = np.random.default_rng(1)
    allb = np.arange(256, dtype=np.uint64)
    for lane in range(4):
        # the byte under test in `lane`, valid random letters elsewhere (neighbours must not leak into the verdict)
        for _ in range(8):
            others = [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 256)].astype(np.uint64) for _ in range(4)]
            others[lane] = allb
            u = words_of(*others)
            d = acgt_diff(u)
            bad = (d >> (8 * lane)) & 0xFF
            rest = d & ~(np.uint64(0xFF) << np.uint64(8 * lane)) & M32
            assert np.all(rest == 0)
            for b in range(256):
                assert (bad[b] == 0) == (b in valid), (lane, b, hex(int(bad[b])))


def test_symbol_codes_are_bits_1_and_2():
    # A=0 C=1 T=2 G=3 is the raw (c >> 1) & 3; the tables on the host are built in that order
    assert [(ord(c) >> 1) & 3 for c in "ACTG"] == [0, 1, 2, 3]


def test_run_start_mask_and_multiply_gather():
    rng = np.random.default_rng(2)
    u = rng.integers(0, 1 << 32, 200000, dtype=np.uint64)
    # make runs likely
    u = np.where(rng.random(u.size) < 0.5, (u & 0xFF) * 0x01010101 & M32, u)
    prev = rng.integers(0, 256, u.size, dtype=np.uint64)
    pv = ((u << 8) | prev) & M32
    run80 = nonzero80(u ^ pv)
    # definition: byte i starts a run iff it differs from the byte before it
    by = [(u >> (8 * i)) & 0xFF for i in range(4)]
    want = ((by[0] != prev).astype(np.uint64) | ((by[1] != by[0]).astype(np.uint64) << 1)
            | ((by[2] != by[1]).astype(np.uint64) << 2) | ((by[3] != by[2]).astype(np.uint64) << 3))
    p = ((run80 * 0x00204081) & M32) >> 28                 # v3: no pre-shift, bits 28..31
    assert np.array_equal(p, want)
    p4 = ((((run80 >> 7) * 0x04081020) & M32) >> 24)       # v2: 4 * p in the top byte
    assert np.array_equal(p4, want * 4)
    sel80 = (((want * 0x00204081) & 0x01010101) << 7) & M32  # expand the four bits back to 0x80 per byte
    assert np.array_equal(sel80, run80)


def test_prmt_compaction_selectors():
    # fill_tables_v2: selector p moves the run-start bytes to the front, zero fill behind (nibble 4 = byte 0 of operand b = 0)
    def prmt(a, b, sel):
        src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
        return sum(src[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))
    for p in range(16):
        sel, j = 0, 0
        for b in range(4):
            if p & (1 << b):
                sel |= b << (4 * j); j += 1
        for jj in range(j, 4):
            sel |= 4 << (4 * jj)
        x = 0x44332211
        got = prmt(x, 0, sel)
        want = 0
        k = 0
        for b in range(4):
            if p & (1 << b):
                want |= ((x >> (8 * b)) & 0xFF) << (8 * k); k += 1
        assert got == want, (p, hex(got), hex(want))


def test_pending_word_push_builds_the_byte_stream():
    # v3_push: append c8/8 bytes (zero above them) to a word-granular stream through a pending register
    rng = np.random.default_rng(3)
    for _ in range(200):
        P, n8, words, ref = 0, 0, [], []
        for _ in range(int(rng.integers(1, 60))):
            cnt = int(rng.integers(0, 5))
            bs = [int(x) for x in rng.integers(1, 256, cnt)]
            comp = sum(b << (8 * i) for i, b in enumerate(bs))
            ref += bs
            f8 = n8 & 24
            lo = (P | (comp << f8)) & M32
            hi = (comp >> (32 - f8)) if f8 else 0
            n8n = n8 + 8 * cnt
            if (n8n ^ n8) & 32:
                words.append(lo); P = hi
            else:
                P = lo
            n8 = n8n
        words.append(P)
        flat = [(w >> (8 * i)) & 0xFF for w in words for i in range(4)]
        assert flat[:len(ref)] == ref and all(b == 0 for b in flat[len(ref):])


def test_group_lookup_swar_and_select16():
    # v3_raw_offset: group = #(cum[g] <= o) - 1 with cum < 0x80 and 0x7F sentinels; then the (o - cum[g])-th set bit
    def select16(m, k):
        pos = 0
        c = bin(m & 0xFF).count("1")
        if k >= c: k -= c; pos += 8; m >>= 8
        c = bin(m & 0xF).count("1")
        if k >= c: k -= c; pos += 4; m >>= 4
        c = bin(m & 0x3).count("1")
        if k >= c: k -= c; pos += 2; m >>= 2
        if k >= (m & 1): pos += 1
        return pos
    rng = np.random.default_rng(4)
    for _ in range(300):
        gpl = int(rng.integers(1, 9))
        masks = [int(rng.integers(0, 1 << 16)) for _ in range(gpl)]
        cum, n = [], 0
        for m in masks:
            cum.append(n); n += bin(m).count("1")
        if n == 0:
            continue
        cumw = cum + [0x7F] * (8 - gpl)
        assert max(cum) < 0x80 and (gpl == 8 or n <= 112)
        for o in range(n):
            g = 0
            for w in range(2):
                cw = sum(cumw[4 * w + i] << (8 * i) for i in range(4))
                ob = o * 0x01010101
                g += bin((((ob | 0x80808080) - cw) & M32) & 0x80808080).count("1")
            g -= 1
            off = 16 * g + select16(masks[g], o - cum[g])
            # definition: position of the o-th set bit over the concatenated masks
            seen, want = -1, None
            for gi, m in enumerate(masks):
                for b in range(16):
                    if m >> b & 1:
                        seen += 1
                        if seen == o:
                            want = 16 * gi + b
            assert off == want


def test_phantom_window_is_a_fixed_point():
    # a window of l phantom 'A's stays itself when another phantom 'A' enters and one leaves: the warm-up may step over
    # zero padding (mq_scan_v3.cuh, phase 1)
    SEED_A, SEED_T = 0x3c8bfbb395c60474, 0x295549f54be24456
    M64 = (1 << 64) - 1
    rol = lambda x, r: ((x << (r % 64)) | (x >> ((64 - r) % 64))) & M64 if r % 64 else x
    for l in range(2, 33):
        F0 = R0 = 0
        for i in range(l):
            F0 ^= rol(SEED_A, l - 1 - i); R0 ^= rol(SEED_T, i)
        # step as the kernel does it (scanning right to left): F = ror(F,1) ^ rol(h(in),l-1) ^ ror(h(out),1)
        TF = rol(SEED_A, l - 1) ^ rol(SEED_A, 63)
        TR = SEED_T ^ rol(SEED_T, l)
        assert rol(F0, 63) ^ TF == F0
        assert rol(R0, 1) ^ TR == R0
