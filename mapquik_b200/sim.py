"""Seeded genome / read simulator (ctypes over csrc/mq_sim.cpp in libmq_host.so).

Replaces pbsim (example/simulate_pbsim.sh) for the synthetic configs of BASELINE.json; read names
use the fixture's truth encoding `S1_<n>!<contig>!<start>!<end>!<strand>`.
"""
import ctypes as C
import os
import numpy as np

from . import _build

_lib = None


class ReadCfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("mean_len", C.c_double), ("sd_len", C.c_double), ("min_len", C.c_uint64),
                ("error_rate", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.HOSTLIB):
            _build.build_host()
        L = C.CDLL(_build.HOSTLIB)
        vp = C.c_void_p
        L.mqsim_random_bases.restype = None; L.mqsim_random_bases.argtypes = [C.c_uint64, C.c_uint64, vp, C.c_uint64]
        L.mqsim_add_satellites.restype = None
        L.mqsim_add_satellites.argtypes = [C.c_uint64, C.c_uint64, vp, C.c_uint64, C.c_double]
        L.mqsim_add_segdups.restype = None; L.mqsim_add_segdups.argtypes = [C.c_uint64, vp, C.c_uint64, C.c_double]
        L.mqsim_add_repeat_families.restype = None
        L.mqsim_add_repeat_families.argtypes = [C.c_uint64, vp, C.c_uint64, C.c_double, C.c_uint32]
        L.mqsim_reads_plan.restype = None
        L.mqsim_reads_plan.argtypes = [C.POINTER(ReadCfg), vp, C.c_uint32, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp]
        L.mqsim_reads_fill.restype = None
        L.mqsim_reads_fill.argtypes = [C.POINTER(ReadCfg), vp, vp, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


# CHM13-like contig length proportions (chr1..22, X, Y), used for the human-scale synthetic genome
CHM13_PROPS = [248.4, 242.7, 201.1, 193.6, 182.0, 172.1, 160.6, 146.3, 150.6, 134.8, 135.1, 133.3, 113.6, 101.2,
               99.8, 96.3, 84.3, 80.5, 61.7, 66.2, 45.1, 51.3, 154.3, 62.5]


def genome(seed, contig_lens, sat_frac=0.0, segdup_frac=0.0, repeat_frac=0.0, n_families=300, names=None):
    """Returns (buf uint8[total], offs uint64[n+1], names)."""
    L = lib()
    lens = np.asarray(contig_lens, dtype=np.uint64)
    offs = np.zeros(lens.size + 1, np.uint64); offs[1:] = np.cumsum(lens)
    buf = np.empty(int(offs[-1]), np.uint8)
    for i in range(lens.size):
        view = buf[int(offs[i]):int(offs[i + 1])]
        L.mqsim_random_bases(seed, i, view.ctypes.data, view.size)
        if sat_frac > 0:
            L.mqsim_add_satellites(seed, i, view.ctypes.data, view.size, sat_frac)
    if repeat_frac > 0:
        L.mqsim_add_repeat_families(seed, buf.ctypes.data, buf.size, repeat_frac, n_families)
    if segdup_frac > 0:
        L.mqsim_add_segdups(seed, buf.ctypes.data, buf.size, segdup_frac)
    if names is None:
        names = [f"chr{i + 1}" for i in range(lens.size)]
    return buf, offs, list(names)


def reads(seed, gbuf, goffs, n_reads, mean_len, sd_len, min_len=1000, error_rate=0.005, first=0, contig_names=None,
          with_names=True, out=None):
    """Returns (buf uint8, offs uint64[n+1], names | None, truth dict of arrays)."""
    L = lib()
    cfg = ReadCfg(seed, float(mean_len), float(sd_len), int(min_len), float(error_rate))
    goffs = np.ascontiguousarray(goffs, dtype=np.uint64); nc = goffs.size - 1
    tc = np.zeros(n_reads, np.uint32); ts = np.zeros(n_reads, np.uint64); tl = np.zeros(n_reads, np.uint64)
    strand = np.zeros(n_reads, np.uint8); olen = np.zeros(n_reads, np.uint64)
    L.mqsim_reads_plan(C.byref(cfg), goffs.ctypes.data, nc, first, n_reads, tc.ctypes.data, ts.ctypes.data, tl.ctypes.data,
                       strand.ctypes.data, olen.ctypes.data)
    offs = np.zeros(n_reads + 1, np.uint64); offs[1:] = np.cumsum(olen)
    total = int(offs[-1])
    buf = out if out is not None else np.empty(total, np.uint8)
    assert buf.size >= total
    L.mqsim_reads_fill(C.byref(cfg), gbuf.ctypes.data, goffs.ctypes.data, first, n_reads, tc.ctypes.data, ts.ctypes.data,
                       tl.ctypes.data, strand.ctypes.data, offs.ctypes.data, buf.ctypes.data)
    names = None
    if with_names:
        cn = contig_names or [f"chr{i + 1}" for i in range(nc)]
        names = [f"S1_{first + i + 1}!{cn[tc[i]]}!{ts[i]}!{ts[i] + tl[i]}!{'-' if strand[i] else '+'}" for i in range(n_reads)]
    truth = {"contig": tc, "start": ts, "len": tl, "strand": strand}
    return buf[:total], offs, names, truth
