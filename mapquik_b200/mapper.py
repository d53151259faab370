"""Host-side mirror of the reference's hot-path interface, over the C ABI.

Names follow the reference so parity tests read like its call sites:

    reference (Rust)                                   here
    -------------------------------------------------  ---------------------------------------
    Params{k,l,density,use_hpc,c,s,g} main.rs:33-47    Params
    Index::new()              closures.rs:24           Index(params, device)
    mers::ref_extract(...)    closures.rs:48           Index.ref_extract(ref_idx, seq_id, seq)
    index.get_count()         closures.rs:92           Index.freeze() -> n_unique
    ReadOnlyIndex::new(...)   closures.rs:94           (same call: freeze is the build->probe barrier)
    mers::find_matches(...)   closures.rs:102          Index.find_matches(q_id, seq) -> str | None

Everything that computes runs in libmapquik_b200.so on the GPU; this module only marshals
buffers and formats PAF text (mq_format_paf).  No CPU fallback exists.
"""
import ctypes as C
import numpy as np

from . import capi
from .capi import EXC_DTYPE, HIT_DTYPE, MqError, Packed as _CPacked, Params as _CParams


class Params:
    """main.rs:174-188 defaults."""

    def __init__(self, k=5, l=31, density=0.01, use_hpc=True, c=4, s=11, g=2000):
        self.k, self.l, self.density, self.use_hpc, self.c, self.s, self.g = k, l, density, use_hpc, c, s, g

    def c_struct(self):
        return _CParams(self.k, self.l, self.density, 1 if self.use_hpc else 0, self.c, self.s, self.g)


def to_upper_u8(seq):
    """closures.rs:63,106 `to_ascii_uppercase` -> contiguous uint8 array."""
    if isinstance(seq, (bytes, bytearray, str)):
        b = seq.encode() if isinstance(seq, str) else bytes(seq)
        return np.frombuffer(b.upper(), dtype=np.uint8)
    a = np.ascontiguousarray(seq, dtype=np.uint8)
    low = (a >= 97) & (a <= 122)
    if low.any():
        a = a.copy(); a[low] -= 32
    return a


def concat(seqs):
    arrs = [to_upper_u8(s) for s in seqs]
    offs = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offs[1:] = np.cumsum([a.size for a in arrs], dtype=np.uint64)
    buf = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    return np.ascontiguousarray(buf), offs


class PackedSeqs:
    """A sequence array in the packed input format (mq_pack): 2-bit codes + block bitmap + exception intervals.
    `pinned=True` keeps the three arrays in pinned host memory (mq_host_alloc) for full-speed uploads."""

    def __init__(self, ascii_u8, n_threads=0, fold_case=False, pinned=False):
        L = capi.lib()
        a = np.ascontiguousarray(ascii_u8, dtype=np.uint8)
        self.n_bases = int(a.size)
        nw, nf = int(L.mq_packed_words(a.size)), int(L.mq_packed_flag_words(a.size))
        self._pinned = []
        self._L = L
        self.words = self._alloc(L, nw, np.uint32, pinned); self.flags = self._alloc(L, nf, np.uint32, pinned)
        cap = 1024
        while True:
            exc = np.zeros(cap, EXC_DTYPE); n = C.c_uint64()
            rc = L.mq_pack(a.ctypes.data, a.size, self.words.ctypes.data, self.flags.ctypes.data, exc.ctypes.data, cap, C.byref(n),
                           n_threads, 1 if fold_case else 0)
            if rc == 0:
                break
            if rc == -5 and n.value > cap:
                cap = int(n.value) + 16
                continue
            raise MqError(f"mq_pack: {L.mq_strerror(rc).decode()}")
        self.exc = exc[:n.value].copy()
        if pinned and self.exc.size:
            e = self._alloc(L, self.exc.size, EXC_DTYPE, True); e[:] = self.exc; self.exc = e

    def _alloc(self, L, n, dtype, pinned):
        nbytes = max(int(n) * np.dtype(dtype).itemsize, 16)
        if not pinned:
            return np.zeros(n, dtype)
        ptr = L.mq_host_alloc(nbytes)
        if not ptr:
            raise MqError("mq_host_alloc failed")
        self._pinned.append(ptr)
        return np.frombuffer((C.c_uint8 * nbytes).from_address(ptr), dtype=np.uint8, count=int(n) * np.dtype(dtype).itemsize).view(dtype)

    def c_struct(self):
        return _CPacked(self.words.ctypes.data, self.flags.ctypes.data, self.exc.ctypes.data if self.exc.size else None,
                        self.exc.size, self.n_bases)

    @property
    def nbytes(self):
        return int(self.words.nbytes + self.flags.nbytes + self.exc.nbytes)

    def unpack(self, first=0, n=None):
        n = self.n_bases - first if n is None else n
        out = np.zeros(n, np.uint8)
        rc = self._L.mq_unpack(self.words.ctypes.data, self.exc.ctypes.data if self.exc.size else None, self.exc.size, first, n,
                               out.ctypes.data)
        if rc:
            raise MqError("mq_unpack")
        return out

    def close(self):
        for p in self._pinned:
            self._L.mq_host_free(p)
        self._pinned = []


class Index:
    def __init__(self, params=None, device=0, devices=None):
        """device: one GPU (mq_create); devices=[...]: one context over several GPUs of this process (mq_create_multi)"""
        self.params = params or Params()
        self._L = capi.lib()
        self._h = C.c_void_p()
        cp = self.params.c_struct()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = self._L.mq_create_multi(C.byref(self._h), C.byref(cp), arr, len(devices))
        else:
            rc = self._L.mq_create(C.byref(self._h), C.byref(cp), device)
        if rc != 0:
            self._h = None
            raise MqError(f"mq_create: {self._L.mq_strerror(rc).decode()}")
        self.ref_map = {}            # ref_idx -> (name, len)   (closures.rs:29)
        self.frozen = False
        self.n_unique = None
        self.n_keys = None

    # -- plumbing ---------------------------------------------------------------------------------
    def _ck(self, rc, what):
        if rc != 0:
            raise MqError(f"{what}: {self._L.mq_strerror(rc).decode()} ({self._L.mq_last_error(self._h).decode()})")

    def close(self):
        if getattr(self, "_h", None):
            self._L.mq_destroy(self._h); self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # -- index build ------------------------------------------------------------------------------
    def add_batch(self, names, seqs, offs, first_ref_idx=None):
        """≙ ref_extract for records first_ref_idx.. (seqs already upper-case uint8, offs uint64[n+1])."""
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1
        first = len(self.ref_map) if first_ref_idx is None else first_ref_idx
        nb = np.zeros(n, np.uint64)
        self._ck(self._L.mq_index_add(self._h, seqs.ctypes.data, offs.ctypes.data, n, first, nb.ctypes.data), "mq_index_add")
        for i in range(n):
            self.ref_map[first + i] = (names[i], int(offs[i + 1] - offs[i]))
        return nb

    def add_batch_packed(self, names, packed, offs, first_ref_idx=None):
        """≙ ref_extract on packed sequences (PackedSeqs)"""
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1
        first = len(self.ref_map) if first_ref_idx is None else first_ref_idx
        nb = np.zeros(n, np.uint64)
        cs = packed.c_struct()
        self._ck(self._L.mq_index_add_packed(self._h, C.byref(cs), offs.ctypes.data, n, first, nb.ctypes.data), "mq_index_add_packed")
        for i in range(n):
            self.ref_map[first + i] = (names[i], int(offs[i + 1] - offs[i]))
        return nb

    def ref_extract(self, ref_idx, seq_id, seq):
        a = to_upper_u8(seq)
        offs = np.array([0, a.size], dtype=np.uint64)
        return int(self.add_batch([seq_id], a, offs, ref_idx)[0])

    def add_segment(self, ref_idx, seq_id, ref_len, seg_start, own_len, data):
        """multi-GPU partitioning: scan [seg_start, seg_start+own_len) of record ref_idx; `data` starts one
        byte before seg_start (if seg_start > 0) and carries the right halo."""
        a = np.ascontiguousarray(data, dtype=np.uint8)
        self._ck(self._L.mq_index_add_segment(self._h, a.ctypes.data, a.size, ref_idx, ref_len, seg_start, own_len),
                 "mq_index_add_segment")
        self.ref_map[ref_idx] = (seq_id, int(ref_len))

    def store_info(self):
        n = C.c_uint64(); s = C.c_uint32()
        self._ck(self._L.mq_store_info(self._h, C.byref(n), C.byref(s)), "mq_store_info")
        return n.value, s.value

    def store_export(self):
        n, s = self.store_info()
        dp, dh = C.c_void_p(), C.c_void_p(); d = np.zeros((s, 3), np.uint64)
        self._ck(self._L.mq_store_export(self._h, C.byref(dp), C.byref(dh), d.ctypes.data), "mq_store_export")
        return dp.value, dh.value, n, d

    def store_import(self, d_pos, d_hash, n, directory):
        d = np.ascontiguousarray(directory, dtype=np.uint64).reshape(-1, 3)
        self._ck(self._L.mq_store_import(self._h, d_pos, d_hash, n, d.ctypes.data, d.shape[0]), "mq_store_import")

    def freeze(self, ref_map=None):
        if ref_map is not None:
            self.ref_map = dict(ref_map)
        n_refs = (max(self.ref_map) + 1) if self.ref_map else 0
        lens = np.zeros(n_refs, np.uint64)
        for i, (_, ln) in self.ref_map.items():
            lens[i] = ln
        u, kk = C.c_uint64(), C.c_uint64()
        self._ck(self._L.mq_index_freeze(self._h, lens.ctypes.data, n_refs, C.byref(u), C.byref(kk)), "mq_index_freeze")
        self.frozen = True; self.n_unique = u.value; self.n_keys = kk.value
        self._ref_lens = lens
        return u.value

    def save(self, path):
        """mq_index_save: frozen table + ref_map to one file."""
        n_refs = self._ref_lens.size
        blob = b"\0".join(self.ref_map.get(i, ("", 0))[0].encode() for i in range(n_refs)) + b"\0" if n_refs else b""
        self._ck(self._L.mq_index_save(self._h, str(path).encode(), blob, len(blob)), "mq_index_save")

    @classmethod
    def from_file(cls, path, params=None, device=0):
        import struct
        with open(path, "rb") as f:
            hdr = f.read(8 + 4 * 4 + 8 + 8 * 5)
        magic, ver, k, l, hpc, dens, slots, n_unique, n_keys, n_refs, names_bytes = struct.unpack("<8sIIIIdQQQQQ", hdr)
        if magic != b"MQB200IX":
            raise MqError("not a mapquik_b200 index file")
        p = params or Params(k=k, l=l, density=dens, use_hpc=bool(hpc))
        ix = cls(p, device)
        lens = np.zeros(n_refs, np.uint64); names = C.create_string_buffer(max(int(names_bytes), 1))
        n = C.c_uint32(); nb = C.c_uint64(); nu = C.c_uint64()
        ix._ck(ix._L.mq_index_load(ix._h, str(path).encode(), lens.ctypes.data, n_refs, C.byref(n), names, names_bytes,
                                   C.byref(nb), C.byref(nu)), "mq_index_load")
        parts = names.raw[:names_bytes].split(b"\0")
        ix.ref_map = {i: (parts[i].decode() if i < len(parts) else "", int(lens[i])) for i in range(n_refs)}
        ix._ref_lens = lens; ix.frozen = True; ix.n_unique = nu.value; ix.n_keys = n_keys
        return ix

    def get_count(self):
        return self.n_unique

    def nb_mers(self):
        n = self._ref_lens.size; nb = np.zeros(n, np.uint64)
        self._ck(self._L.mq_index_nb_mers(self._h, nb.ctypes.data, n), "mq_index_nb_mers")
        return nb

    # -- mapping ----------------------------------------------------------------------------------
    def map_batch(self, seqs, offs, out=None):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1
        hits = out if out is not None else np.zeros(n, HIT_DTYPE)
        self._ck(self._L.mq_map_batch(self._h, seqs.ctypes.data, offs.ctypes.data, n, hits.ctypes.data), "mq_map_batch")
        return hits

    def map_batch_packed(self, packed, offs, out=None):
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1
        hits = out if out is not None else np.zeros(n, HIT_DTYPE)
        cs = packed.c_struct()
        self._ck(self._L.mq_map_batch_packed(self._h, C.byref(cs), offs.ctypes.data, n, hits.ctypes.data), "mq_map_batch_packed")
        return hits

    def map_batch_device(self, d_seqs, offs, d_hits):
        """sequences and hits resident on the device; offs is a HOST uint64 array"""
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        self._ck(self._L.mq_map_batch_device(self._h, d_seqs, offs.ctypes.data, offs.size - 1, d_hits), "mq_map_batch_device")

    def map_batch_packed_device(self, d_words, d_flags, d_exc, n_exc, n_bases, offs, d_hits):
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        cs = _CPacked(d_words, d_flags, d_exc, n_exc, n_bases)
        self._ck(self._L.mq_map_batch_packed_device(self._h, C.byref(cs), offs.ctypes.data, offs.size - 1, d_hits),
                 "mq_map_batch_packed_device")

    def paf_line(self, q_id, q_len, hit):
        if not hit["mapped"]:
            return None
        name, rlen = self.ref_map[int(hit["ref_idx"])]
        h = np.zeros(1, HIT_DTYPE); h[0] = hit
        buf = C.create_string_buffer(len(q_id) + len(name) + 256)
        n = self._L.mq_format_paf(buf, len(buf), q_id.encode(), q_len, name.encode(), rlen, h.ctypes.data)
        if n < 0:
            raise MqError("mq_format_paf")
        return buf.value.decode()

    def find_matches(self, q_id, seq):
        """≙ mers::find_matches: PAF line or None."""
        a = to_upper_u8(seq)
        hits = self.map_batch(a, np.array([0, a.size], dtype=np.uint64))
        return self.paf_line(q_id, a.size, hits[0])

    # -- stage introspection (parity tests) -----------------------------------------------------------
    def minimizers(self, seqs, offs):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1; tot = C.c_uint64(); so = np.zeros(n + 1, np.uint64)
        self._ck(self._L.mq_minimizers(self._h, seqs.ctypes.data, offs.ctypes.data, n, so.ctypes.data, None, None, 0,
                                       C.byref(tot)), "mq_minimizers")
        pos = np.zeros(tot.value, np.uint32); hs = np.zeros(tot.value, np.uint64)
        self._ck(self._L.mq_minimizers(self._h, seqs.ctypes.data, offs.ctypes.data, n, so.ctypes.data, pos.ctypes.data,
                                       hs.ctypes.data, tot.value, C.byref(tot)), "mq_minimizers")
        return so, pos, hs

    def minimizers_packed(self, packed, offs):
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1; tot = C.c_uint64(); so = np.zeros(n + 1, np.uint64); cs = packed.c_struct()
        self._ck(self._L.mq_minimizers_packed(self._h, C.byref(cs), offs.ctypes.data, n, so.ctypes.data, None, None, 0,
                                              C.byref(tot)), "mq_minimizers_packed")
        pos = np.zeros(tot.value, np.uint32); hs = np.zeros(tot.value, np.uint64)
        self._ck(self._L.mq_minimizers_packed(self._h, C.byref(cs), offs.ctypes.data, n, so.ctypes.data, pos.ctypes.data,
                                              hs.ctypes.data, tot.value, C.byref(tot)), "mq_minimizers_packed")
        return so, pos, hs

    def kminmers(self, seqs, offs):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1; tot = C.c_uint64(); so = np.zeros(n + 1, np.uint64)
        self._ck(self._L.mq_kminmers(self._h, seqs.ctypes.data, offs.ctypes.data, n, so.ctypes.data, None, None, None,
                                     None, 0, C.byref(tot)), "mq_kminmers")
        q = tot.value
        st = np.zeros(q, np.uint32); en = np.zeros(q, np.uint32); orv = np.zeros(q, np.uint32); hs = np.zeros(q, np.uint64)
        self._ck(self._L.mq_kminmers(self._h, seqs.ctypes.data, offs.ctypes.data, n, so.ctypes.data, st.ctypes.data,
                                     en.ctypes.data, orv.ctypes.data, hs.ctypes.data, q, C.byref(tot)), "mq_kminmers")
        return so, st, en, orv >> 1, orv & 1, hs

    def get(self, hashes):
        h = np.ascontiguousarray(hashes, dtype=np.uint64); n = h.size
        f = np.zeros(n, np.uint8); rc = np.zeros(n, np.uint8)
        rid, st, en, off = (np.zeros(n, np.uint32) for _ in range(4))
        self._ck(self._L.mq_index_get(self._h, h.ctypes.data, n, f.ctypes.data, rid.ctypes.data, st.ctypes.data,
                                      en.ctypes.data, off.ctypes.data, rc.ctypes.data), "mq_index_get")
        return f.astype(bool), rid, st, en, off, rc

    def matches(self, seqs, offs):
        seqs = np.ascontiguousarray(seqs, dtype=np.uint8); offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = offs.size - 1; tot = C.c_uint64(); mo = np.zeros(n + 1, np.uint64)
        self._ck(self._L.mq_matches(self._h, seqs.ctypes.data, offs.ctypes.data, n, mo.ctypes.data, None, 0, C.byref(tot)),
                 "mq_matches")
        f = np.zeros((tot.value, 6), np.uint32)
        self._ck(self._L.mq_matches(self._h, seqs.ctypes.data, offs.ctypes.data, n, mo.ctypes.data, f.ctypes.data,
                                    tot.value, C.byref(tot)), "mq_matches")
        return mo, f

    def last_ms(self, stage="total"):
        return self._L.mq_last_ms(self._h, stage.encode())

    def total_ms(self, stage):
        return self._L.mq_total_ms(self._h, stage.encode())

    def launch_count(self):
        return self._L.mq_launch_count(self._h)

    def set_host_threads(self, n):
        """host threads the mapping calls may use to pack ASCII input on the fly (the reference's --threads, main.rs:138-141)"""
        rc = self._L.mq_set_host_threads(self._h, int(n))
        if rc:
            raise MapquikError(rc, self._L.mq_last_error(self._h))

    def last_counter(self, name):
        return int(self._L.mq_last_counter(self._h, name.encode()))

    def table_bytes(self):
        return self._L.mq_table_bytes(self._h)
