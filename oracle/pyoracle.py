"""ctypes binding of the CPU oracle (oracle/libmq_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under mapquik_b200/
imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmq_oracle.so")


class Params(C.Structure):
    _fields_ = [("k", C.c_uint32), ("l", C.c_uint32), ("density", C.c_double),
                ("use_hpc", C.c_uint32), ("c", C.c_uint32), ("s", C.c_uint32), ("g", C.c_uint32)]


class Kminmer(C.Structure):
    _fields_ = [("start", C.c_uint64), ("end", C.c_uint64), ("offset", C.c_uint64),
                ("hash", C.c_uint64), ("rev", C.c_uint32), ("pad_", C.c_uint32)]


class Match(C.Structure):
    _fields_ = [("q_start", C.c_uint64), ("q_end", C.c_uint64), ("r_start", C.c_uint64),
                ("r_end", C.c_uint64), ("count", C.c_uint64), ("rc", C.c_uint32), ("ref_id", C.c_uint32)]


class Hit(C.Structure):
    _fields_ = [("mapped", C.c_uint8), ("rc", C.c_uint8), ("mapq", C.c_uint8), ("pad_", C.c_uint8),
                ("ref_idx", C.c_uint32), ("q_start", C.c_uint64), ("q_end", C.c_uint64),
                ("r_start", C.c_uint64), ("r_end", C.c_uint64), ("score", C.c_uint64)]


HIT_DTYPE = np.dtype([("mapped", "u1"), ("rc", "u1"), ("mapq", "u1"), ("pad_", "u1"), ("ref_idx", "<u4"),
                      ("q_start", "<u8"), ("q_end", "<u8"), ("r_start", "<u8"), ("r_end", "<u8"),
                      ("score", "<u8")])
assert HIT_DTYPE.itemsize == C.sizeof(Hit)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "mq_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, u64p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        PP = C.POINTER(Params)
        L.orc_hash_bound.restype = C.c_uint64; L.orc_hash_bound.argtypes = [C.c_double]
        for f in (L.orc_nthash_fwd, L.orc_nthash_rev):
            f.restype = C.c_uint64; f.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_kminmer_hash.restype = C.c_uint64; L.orc_kminmer_hash.argtypes = [u64p, C.c_uint32, u32p]
        for f in (L.orc_minimizers, L.orc_minimizers_slow):
            f.restype = C.c_size_t; f.argtypes = [C.c_void_p, C.c_size_t, PP, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_kminmers.restype = C.c_size_t
        L.orc_kminmers.argtypes = [C.c_void_p, C.c_size_t, PP, C.c_void_p, C.c_size_t]
        L.orc_index_new.restype = C.c_void_p; L.orc_index_new.argtypes = [C.c_size_t]
        L.orc_index_free.restype = None; L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_ref_extract.restype = C.c_uint64
        L.orc_ref_extract.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, PP]
        L.orc_index_add.restype = None
        L.orc_index_add.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_index_count.restype = C.c_uint64; L.orc_index_count.argtypes = [C.c_void_p]
        L.orc_index_slots.restype = C.c_uint64; L.orc_index_slots.argtypes = [C.c_void_p]
        L.orc_index_get.restype = C.c_int
        L.orc_index_get.argtypes = [C.c_void_p, C.c_uint64, u32p, u64p, u64p, u64p, u32p]
        L.orc_chain_matches.restype = C.c_size_t
        L.orc_chain_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, PP, C.c_void_p, C.c_size_t]
        L.orc_chain_matches_kms.restype = C.c_size_t
        L.orc_chain_matches_kms.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_find_matches_kms.restype = C.c_int
        L.orc_find_matches_kms.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_uint32, PP, C.c_void_p]
        L.orc_best_of_matches.restype = C.c_int
        L.orc_best_of_matches.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_void_p, C.c_uint32, PP, C.c_void_p]
        L.orc_find_matches.restype = C.c_int
        L.orc_find_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, PP, C.c_void_p]
        L.orc_index_add_batch.restype = None
        L.orc_index_add_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, PP, C.c_void_p, C.c_int]
        L.orc_map_batch.restype = None
        L.orc_map_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, PP, C.c_void_p, C.c_int]
        L.orc_format_paf.restype = C.c_int
        L.orc_format_paf.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_void_p]
        L.orc_find_coords.restype = None
        L.orc_find_coords.argtypes = [C.c_uint64, C.c_uint64, C.c_int] + [C.c_uint64] * 4 + [u64p] * 4
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def params(k=5, l=31, density=0.01, use_hpc=True, c=4, s=11, g=2000):
    return Params(k, l, density, 1 if use_hpc else 0, c, s, g)


def _u8(seq):
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


def minimizers(seq, p, slow=False):
    a = _u8(seq); L = lib()
    f = L.orc_minimizers_slow if slow else L.orc_minimizers
    n = f(a.ctypes.data, a.size, C.byref(p), None, None, 0)
    pos = np.zeros(n, np.uint64); hs = np.zeros(n, np.uint64)
    if n:
        f(a.ctypes.data, a.size, C.byref(p), pos.ctypes.data, hs.ctypes.data, n)
    return pos, hs


KM_DTYPE = np.dtype([("start", "<u8"), ("end", "<u8"), ("offset", "<u8"), ("hash", "<u8"),
                     ("rev", "<u4"), ("pad_", "<u4")])
MATCH_DTYPE = np.dtype([("q_start", "<u8"), ("q_end", "<u8"), ("r_start", "<u8"), ("r_end", "<u8"),
                        ("count", "<u8"), ("rc", "<u4"), ("ref_id", "<u4")])


def kminmers(seq, p):
    a = _u8(seq); L = lib()
    n = L.orc_kminmers(a.ctypes.data, a.size, C.byref(p), None, 0)
    out = np.zeros(n, KM_DTYPE)
    if n:
        L.orc_kminmers(a.ctypes.data, a.size, C.byref(p), out.ctypes.data, n)
    return out


def kminmer_hash(mers):
    w = np.ascontiguousarray(mers, dtype=np.uint64); rev = C.c_uint32(0)
    h = lib().orc_kminmer_hash(w.ctypes.data_as(C.POINTER(C.c_uint64)), w.size, C.byref(rev))
    return h, rev.value


class Index:
    """index.rs Index + ReadOnlyIndex and the ref_map of closures.rs:29."""

    def __init__(self, p, capacity_hint=1 << 16):
        self.p = p
        self.h = lib().orc_index_new(capacity_hint)
        self.ref_names, self.ref_lens = [], []

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h); self.h = None

    def add_ref(self, name, seq):
        a = _u8(seq)
        idx = len(self.ref_names)
        nb = lib().orc_ref_extract(self.h, idx, a.ctypes.data, a.size, C.byref(self.p))
        self.ref_names.append(name); self.ref_lens.append(a.size)
        return nb

    def add_batch(self, names, seqs, offs, threads=0):
        seqs = _u8(seqs); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1; nb = np.zeros(n, np.uint64)
        lib().orc_index_add_batch(self.h, seqs.ctypes.data, offs.ctypes.data, n, len(self.ref_names),
                                  C.byref(self.p), nb.ctypes.data, threads)
        self.ref_names += list(names); self.ref_lens += [int(x) for x in np.diff(offs)]
        return nb

    def add_tuple(self, h, rid, start, end, offset, rc):
        lib().orc_index_add(self.h, h, rid, start, end, offset, rc)

    def count(self):
        return lib().orc_index_count(self.h)

    def slots(self):
        return lib().orc_index_slots(self.h)

    def get(self, h):
        rid, rc = C.c_uint32(), C.c_uint32(); s, e, o = C.c_uint64(), C.c_uint64(), C.c_uint64()
        ok = lib().orc_index_get(self.h, int(h), C.byref(rid), C.byref(s), C.byref(e), C.byref(o), C.byref(rc))
        return (rid.value, s.value, e.value, o.value, rc.value) if ok else None

    def chain_matches(self, seq):
        a = _u8(seq); L = lib()
        n = L.orc_chain_matches(self.h, a.ctypes.data, a.size, C.byref(self.p), None, 0)
        out = np.zeros(n, MATCH_DTYPE)
        if n:
            L.orc_chain_matches(self.h, a.ctypes.data, a.size, C.byref(self.p), out.ctypes.data, n)
        return out

    def chain_matches_kms(self, kms):
        kms = np.ascontiguousarray(kms, dtype=KM_DTYPE); L = lib()
        n = L.orc_chain_matches_kms(self.h, kms.ctypes.data, kms.size, None, 0)
        out = np.zeros(n, MATCH_DTYPE)
        if n:
            L.orc_chain_matches_kms(self.h, kms.ctypes.data, kms.size, out.ctypes.data, n)
        return out

    def find_matches_kms(self, kms, q_len):
        kms = np.ascontiguousarray(kms, dtype=KM_DTYPE); hit = np.zeros(1, HIT_DTYPE); lens = self._lens()
        lib().orc_find_matches_kms(self.h, kms.ctypes.data, kms.size, q_len, lens.ctypes.data, lens.size,
                                   C.byref(self.p), hit.ctypes.data)
        return hit[0]

    def _lens(self):
        return np.asarray(self.ref_lens, dtype=np.uint64)

    def find_matches(self, seq):
        a = _u8(seq); hit = np.zeros(1, HIT_DTYPE); lens = self._lens()
        lib().orc_find_matches(self.h, a.ctypes.data, a.size, lens.ctypes.data, lens.size, C.byref(self.p),
                               hit.ctypes.data)
        return hit[0]

    def map_batch(self, seqs, offs, threads=0):
        seqs = _u8(seqs); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1; hits = np.zeros(n, HIT_DTYPE); lens = self._lens()
        lib().orc_map_batch(self.h, seqs.ctypes.data, offs.ctypes.data, n, lens.ctypes.data, lens.size,
                            C.byref(self.p), hits.ctypes.data, threads)
        return hits

    def paf_line(self, q_id, q_len, hit):
        buf = C.create_string_buffer(1024)
        h = np.zeros(1, HIT_DTYPE); h[0] = hit
        rid = int(hit["ref_idx"])
        n = lib().orc_format_paf(buf, 1024, q_id.encode(), q_len, self.ref_names[rid].encode(),
                                 self.ref_lens[rid], h.ctypes.data)
        assert n > 0
        return buf.value.decode()


def best_of_matches(matches, q_len, ref_lens, p):
    ms = np.ascontiguousarray(matches, dtype=MATCH_DTYPE); lens = np.ascontiguousarray(ref_lens, dtype=np.uint64)
    hit = np.zeros(1, HIT_DTYPE)
    lib().orc_best_of_matches(ms.ctypes.data, ms.size, q_len, lens.ctypes.data, lens.size, C.byref(p), hit.ctypes.data)
    return hit[0]


def find_coords(q_len, r_len, rc, q_start, q_end, r_start, r_end):
    o = [C.c_uint64() for _ in range(4)]
    lib().orc_find_coords(q_len, r_len, int(rc), q_start, q_end, r_start, r_end, *[C.byref(x) for x in o])
    return tuple(x.value for x in o)
