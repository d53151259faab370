"""Second, independent restatement of the hot path in plain Python (small cases only).

It deliberately mirrors the *shape* of the Rust sources (peekable iterator + recursive
Match.extend, HashMap<ref, Vec<Match>>, Chain filtering) rather than the streaming form
used by the C oracle and the CUDA kernels, so that agreement between the two pins the
oracle's control flow.  Citations: src/mers.rs, src/index.rs, src/match.rs, src/chain.rs.
The seeding stage follows DESIGN.md section 2 from the definition (no rolling hash).
"""
M64 = (1 << 64) - 1
SEED = {ord("A"): 0x3c8bfbb395c60474, ord("C"): 0x3193c18562a02b4c,
        ord("G"): 0x20323ed082572324, ord("T"): 0x295549f54be24456}
COMP = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A")}


def rol(x, r):
    r %= 64
    return ((x << r) | (x >> (64 - r))) & M64 if r else x


def ntf(s):
    l = len(s); v = 0
    for i, c in enumerate(s):
        v ^= rol(SEED.get(c, 0), l - 1 - i)
    return v


def ntr(s):
    v = 0
    for i, c in enumerate(s):
        v ^= rol(SEED.get(COMP.get(c, 0), 0), i)
    return v


def hash_bound(density):
    import struct
    b = density * float(M64)          # float(2^64-1) == 2^64
    if not b > 0:
        return 0
    if b >= 2.0 ** 64:
        return M64
    return int(b)


def minimizers(seq, l, density, use_hpc=True):
    s = list(bytes(seq))    # caller upper-cases (closures.rs:63,106)
    sym, raw = [], []
    for i, c in enumerate(s):
        if use_hpc and i > 0 and c == s[i - 1]:
            continue
        sym.append(c); raw.append(i)
    bound = hash_bound(density)
    out = []
    for i in range(0, len(sym) - l + 1):
        w = sym[i:i + l]
        h = min(ntf(w), ntr(w))
        if h < bound:
            out.append((raw[i], h))
    return out


def mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def kminmer_hash(w):
    k = len(w)
    r = list(reversed(w))
    rev = r < list(w)                 # lexicographic; palindrome => forward
    canon = r if rev else list(w)
    h = 0x9E3779B97F4A7C15 ^ k
    for x in canon:
        h = mix64(h ^ x)
    return h, rev


class Kminmer:
    __slots__ = ("start", "end", "offset", "rev", "hash")


def kminmers(seq, k, l, density, use_hpc=True):
    if len(seq) < l + k - 1:
        return []
    m = minimizers(seq, l, density, use_hpc)
    out = []
    for j in range(0, len(m) - k + 1):
        km = Kminmer()
        km.hash, km.rev = kminmer_hash([h for _, h in m[j:j + k]])
        km.start = m[j][0]; km.end = m[j + k - 1][0] + l; km.offset = j
        out.append(km)
    return out


class Entry:
    __slots__ = ("id", "start", "end", "offset", "rc")

    def __init__(self, id, start, end, offset, rc):
        self.id, self.start, self.end, self.offset, self.rc = id, start, end, offset, rc


class Index:                                   # index.rs
    def __init__(self):
        self.map = {}

    def add_with_mer(self, id, mer):           # index.rs:100-104
        old = self.map.get(mer.hash)
        self.map[mer.hash] = Entry(id, mer.start, mer.end, mer.offset, mer.rev)
        if old is not None:
            self.map[mer.hash] = Entry(0, 0, 0, 0, False)

    def get_count(self):
        return sum(1 for e in self.map.values() if e.end != 0)

    def get(self, h):                          # index.rs:118-126
        e = self.map.get(h)
        if e is not None and e.end != 0:
            return e
        return None


def i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def wrap_i32(x):
    return i32(x)


class Match:                                   # match.rs
    def __init__(self, q, r):
        self.q_start, self.q_end, self.r_start, self.r_end = q.start, q.end, r.start, r.end
        self.count = 1
        self.rc = (q.rev != r.rc)

    def key(self):
        return (self.q_start, self.q_end, self.r_start, self.r_end, self.count, self.rc)

    def update(self, q, r):
        if self.rc:
            self.r_start = r.start
        else:
            self.r_end = r.end
        self.q_end = q.end
        self.count += 1

    def check(self, q, r, p):                  # match.rs:39-43 -- Rust precedence: && binds tighter than ||
        return ((r.id == p.id) and ((q.rev != r.rc) == self.rc) and
                (self.rc and (wrap_i32(i32(p.offset) - i32(r.offset)) == 1))) or \
               ((not self.rc) and (wrap_i32(i32(r.offset) - i32(p.offset)) == 1))

    def extend(self, it, index, p):            # match.rs:45-58 (recursive, like the source)
        q = it.peek()
        if q is not None:
            r = index.get(q.hash)
            if r is not None:
                if self.check(q, r, p):
                    self.update(q, r)
                    it.next()
                    self.extend(it, index, r)
            else:
                it.next()
        else:
            it.next()


class Peekable:
    def __init__(self, xs):
        self.xs, self.i = xs, 0

    def peek(self):
        return self.xs[self.i] if self.i < len(self.xs) else None

    def next(self):
        if self.i < len(self.xs):
            self.i += 1
            return self.xs[self.i - 1]
        return None


def chain_matches(kms, index):                 # mers.rs:57-73
    import sys
    sys.setrecursionlimit(max(10000, len(kms) + 100))
    per_ref = {}
    it = Peekable(kms)
    while True:
        q = it.next()
        if q is None:
            break
        r = index.get(q.hash)
        if r is not None:
            h = Match(q, r)
            h.extend(it, index, r)
            per_ref.setdefault(r.id, []).append(h)
    return per_ref


def gap_too_long(a1, a0, b1, b0, g):           # chain.rs:132-142
    g1 = wrap_i32(i32(a1) - i32(a0)); g2 = wrap_i32(i32(b1) - i32(b0))
    d = wrap_i32(g1 - g2)
    ad = wrap_i32(-d) if d < 0 else d          # abs wraps at i32::MIN
    return (ad & M64) > g                      # `as usize` sign-extends


def compatible(h1, h2, g):                     # chain.rs:43-63
    if h1.key() == h2.key():
        return True
    if h1.rc != h2.rc:
        return False
    u, v = (h1, h2) if h1.q_start < h2.q_start else (h2, h1)
    if u.rc:
        if u.r_start <= v.r_start or gap_too_long(v.q_start, u.q_end, u.r_start, v.r_end, g):
            return False
    elif v.r_start <= u.r_start or gap_too_long(v.q_start, u.q_end, v.r_start, u.r_end, g):
        return False
    return True


def get_match(matches, c, s, g):               # chain.rs:147-169
    ms = list(matches)
    if len(ms) > 1:
        mx, mc = 0, 0
        for i, m in enumerate(ms):
            if m.count > mc:
                mx, mc = i, m.count
        h = ms[mx]
        ms = [m for m in ms if compatible(h, m, g)]
    if not ms:
        return None
    score = sum(m.count for m in ms)
    mapq = 60 if (s != 0 and c != 0) and (len(ms) >= c or score >= s) else 0
    first, last = ms[0], ms[-1]
    rc = first.rc
    if rc and len(ms) > 1:
        return (rc, first.q_start, (last.q_end - 1) & M64, last.r_start, (first.r_end - 1) & M64, score, mapq)
    return (rc, first.q_start, (last.q_end - 1) & M64, first.r_start, (last.r_end - 1) & M64, score, mapq)


def find_coords(q_id, q_len, r_id, r_len, coords):   # mers.rs:131-183
    rc, q_start, q_end, r_start, r_end, score, mapq = coords
    W = lambda x: x & M64
    tail = W(q_len - q_end - 1)
    if not rc:
        if r_start >= q_start:
            frs, exs = W(r_start - q_start), q_start
        else:
            frs, exs = 0, r_start
        if W(r_end + tail) <= W(r_len - 1):
            fre, exe = W(r_end + tail), tail
        else:
            fre, exe = W(r_len - 1), W(r_len - r_end - 1)
    else:
        if W(r_end + q_start) <= W(r_len - 1):
            fre, exs = W(r_end + q_start), q_start
        else:
            fre, exs = W(r_len - 1), W(r_len - r_end - 1)
        if r_start >= tail:
            frs, exe = W(r_start - tail), tail
        else:
            frs, exe = 0, r_start
    fqs, fqe = W(q_start - exs), W(q_end + exe)
    return "\t".join(str(x) for x in (q_id, q_len, fqs, fqe, "-" if rc else "+", r_id, r_len, frs, fre,
                                      score, r_len, mapq))


def find_matches(q_id, seq, ref_map, index, k, l, density, use_hpc, c, s, g):   # mers.rs:77-129
    kms = kminmers(seq, k, l, density, use_hpc)
    per_ref = chain_matches(kms, index)
    allc = []
    for rid, ms in per_ref.items():
        t = get_match(ms, c, s, g)
        if t is not None:
            allc.append((rid, t))
    if not allc:
        return None
    if len(allc) == 1:
        rid, t = allc[0]
        return find_coords(q_id, len(seq), ref_map[rid][0], ref_map[rid][1], t)
    mx = mxc = sec = secc = 0
    for i, (_, t) in enumerate(allc):           # find_largest_two_chains
        cnt = t[5]
        if cnt > mxc:
            sec, secc, mx, mxc = mx, mxc, i, cnt
        elif cnt > secc:
            sec, secc = i, cnt
    if mxc == secc:
        return None
    rid, t = allc[mx]
    return find_coords(q_id, len(seq), ref_map[rid][0], ref_map[rid][1], t)
