"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/*.h declares,
formats PAF like the reference, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mapquik_b200 import capi, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_PAF = ("S1_1!chr1!224752794!224777027!+\t24299\t0\t24298\t+\tchr1\t248387328\t224752793\t224777027\t132\t"
              "248387328\t60")   # experiments/intersect_pafs.py:14


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mapquik_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mq_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_build.LIB), "libmapquik_b200.so not built (python __graft_entry__.py)"
    L = C.CDLL(_build.LIB)
    syms = header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == syms          # the Python binding declares exactly the header's surface


def test_struct_layouts():
    assert C.sizeof(capi.Params) == 32
    assert capi.HIT_DTYPE.itemsize == 48
    assert [capi.HIT_DTYPE.fields[f][1] for f in ("mapped", "ref_idx", "q_start", "score")] == [0, 4, 8, 40]


def test_format_paf_golden_line():
    L = capi.lib()
    hit = np.zeros(1, capi.HIT_DTYPE)
    hit[0] = (1, 0, 60, 0, 0, 0, 24298, 224752793, 224777027, 132)
    buf = C.create_string_buffer(512)
    n = L.mq_format_paf(buf, 512, b"S1_1!chr1!224752794!224777027!+", 24299, b"chr1", 248387328, hit.ctypes.data)
    assert n == len(GOLDEN_PAF) and buf.value.decode() == GOLDEN_PAF
    hit[0]["rc"] = 1
    L.mq_format_paf(buf, 512, b"q", 10, b"r", 20, hit.ctypes.data)
    assert buf.value.decode().split("\t")[4] == "-"
    assert L.mq_format_paf(buf, 8, b"q", 10, b"r", 20, hit.ctypes.data) < 0      # buffer too small
    # every column at its extremes; the line fits a buffer of exactly its length + NUL and nothing smaller
    big = 2 ** 64 - 1
    for q_len, r_len, fields in ((0, 0, (1, 0, 0, 0, 0, 0, 0, 0, 0, 0)), (big, big, (1, 1, 255, 0, 2 ** 32 - 1, big, big, big, big, big))):
        hit[0] = fields
        n = L.mq_format_paf(buf, 512, b"read/1", q_len, b"chrUn_x", r_len, hit.ctypes.data)
        h = hit[0]
        want = "\t".join(str(x) for x in ("read/1", q_len, int(h["q_start"]), int(h["q_end"]), "-" if h["rc"] else "+", "chrUn_x", r_len,
                                          int(h["r_start"]), int(h["r_end"]), int(h["score"]), r_len, int(h["mapq"])))
        assert n == len(want) and buf.value.decode() == want
        assert L.mq_format_paf(buf, n + 1, b"read/1", q_len, b"chrUn_x", r_len, hit.ctypes.data) == n
        assert L.mq_format_paf(buf, n, b"read/1", q_len, b"chrUn_x", r_len, hit.ctypes.data) < 0


def test_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("GPU present: covered by the gpu tests")
    L = capi.lib()
    h = C.c_void_p()
    p = capi.Params(5, 31, 0.01, 1, 4, 11, 2000)
    rc = L.mq_create(C.byref(h), C.byref(p), 0)
    assert rc == -2 and not h.value                       # MQ_ERR_CUDA, loudly
    assert b"no CPU fallback" in L.mq_strerror(rc)
    from mapquik_b200 import Index, MqError
    with pytest.raises(MqError):
        Index()


def test_bad_params_rejected_before_any_device_work():
    L = capi.lib()
    h = C.c_void_p()
    for k, l in ((5, 1), (5, 33), (0, 31), (33, 31)):
        p = capi.Params(k, l, 0.01, 1, 4, 11, 2000)
        assert L.mq_create(C.byref(h), C.byref(p), 0) == -1
    assert L.mq_abi_version() == 2
    d = (C.c_int * 2)(0, 1)
    assert L.mq_create_multi(C.byref(h), C.byref(capi.Params(5, 33, 0.01, 1, 4, 11, 2000)), d, 2) == -1
    assert L.mq_create_multi(C.byref(h), C.byref(capi.Params(5, 31, 0.01, 1, 4, 11, 2000)), d, 0) == -1


def test_null_context_calls_fail_cleanly():
    """entry points that take a context reject NULL with a status code (nothing dereferenced, nothing thrown)"""
    L = capi.lib()
    assert L.mq_set_host_threads(None, 4) == -1
    assert L.mq_last_counter(None, b"h2d_bytes") == 0
    assert L.mq_device_count(None) == 0 and L.mq_launch_count(None) == 0
    assert L.mq_map_batch(None, None, None, 0, None) == -1


def test_upper_casing_mirror():
    from mapquik_b200 import to_upper_u8, concat
    assert to_upper_u8("acgtNn").tobytes() == b"ACGTNN"
    assert to_upper_u8(np.frombuffer(b"aCgT", np.uint8)).tobytes() == b"ACGT"
    buf, offs = concat([b"acg", b"", b"TT"])
    assert buf.tobytes() == b"ACGTT" and offs.tolist() == [0, 3, 3, 5]
