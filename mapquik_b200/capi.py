"""ctypes binding of the C ABI in include/mapquik_b200.h (libmapquik_b200.so).

The library is the product; this module only declares its entry points.  It raises if the
shared object is missing or cannot be loaded -- there is no CPU fallback.
"""
import ctypes as C
import os
import numpy as np

from . import _build

MQ_OK = 0


class MqError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("k", C.c_uint32), ("l", C.c_uint32), ("density", C.c_double), ("use_hpc", C.c_uint32),
                ("c", C.c_uint32), ("s", C.c_uint32), ("g", C.c_uint32)]


class Exc(C.Structure):
    _fields_ = [("start", C.c_uint64), ("len", C.c_uint32), ("byte", C.c_uint32)]


class Packed(C.Structure):
    _fields_ = [("words", C.c_void_p), ("flags", C.c_void_p), ("exc", C.c_void_p), ("n_exc", C.c_uint64),
                ("n_bases", C.c_uint64)]


EXC_DTYPE = np.dtype([("start", "<u8"), ("len", "<u4"), ("byte", "<u4")])
assert EXC_DTYPE.itemsize == 16

HIT_DTYPE = np.dtype([("mapped", "u1"), ("rc", "u1"), ("mapq", "u1"), ("pad_", "u1"), ("ref_idx", "<u4"),
                      ("q_start", "<u8"), ("q_end", "<u8"), ("r_start", "<u8"), ("r_end", "<u8"),
                      ("score", "<u8")])
assert HIT_DTYPE.itemsize == 48

EXPORTS = ["mq_create", "mq_destroy", "mq_strerror", "mq_last_error", "mq_abi_version", "mq_host_alloc",
           "mq_host_free", "mq_index_add", "mq_index_add_segment", "mq_store_info", "mq_store_export",
           "mq_store_import", "mq_index_freeze", "mq_index_nb_mers", "mq_map_batch", "mq_map_batch_device",
           "mq_format_paf", "mq_minimizers", "mq_kminmers", "mq_index_get", "mq_matches", "mq_last_ms",
           "mq_launch_count", "mq_stream", "mq_sync", "mq_table_bytes", "mq_table_slots", "mq_scan_kernel_launches",
           "mq_set_host_threads", "mq_last_counter",
           "mq_minimizer_count", "mq_dev_alloc", "mq_dev_free", "mq_dev_upload", "mq_dev_download", "mq_dev_memset",
           "mq_region_begin", "mq_region_end_ms", "mq_index_save", "mq_index_load", "mq_total_ms",
           "mq_create_multi", "mq_device_count", "mq_packed_words", "mq_packed_flag_words", "mq_pack", "mq_pack_at",
           "mq_unpack", "mq_index_add_packed", "mq_map_batch_packed", "mq_map_batch_packed_device", "mq_minimizers_packed",
           "mq_store_reserve", "mq_store_commit"]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MQ_LIB", _build.LIB)      # MQ_LIB: A/B builds of the same ABI (tuning sweeps)
    if not os.path.exists(path):
        raise MqError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(nvcc); there is no CPU fallback")
    L = C.CDLL(path)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.mq_create.restype = C.c_int; L.mq_create.argtypes = [C.POINTER(vp), C.POINTER(Params), C.c_int]
    L.mq_destroy.restype = None; L.mq_destroy.argtypes = [vp]
    L.mq_strerror.restype = C.c_char_p; L.mq_strerror.argtypes = [C.c_int]
    L.mq_last_error.restype = C.c_char_p; L.mq_last_error.argtypes = [vp]
    L.mq_abi_version.restype = C.c_int
    L.mq_host_alloc.restype = vp; L.mq_host_alloc.argtypes = [C.c_size_t]
    L.mq_host_free.restype = None; L.mq_host_free.argtypes = [vp]
    L.mq_index_add.restype = C.c_int; L.mq_index_add.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp]
    L.mq_index_add_segment.restype = C.c_int
    L.mq_index_add_segment.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64]
    L.mq_store_info.restype = C.c_int; L.mq_store_info.argtypes = [vp, u64p, C.POINTER(C.c_uint32)]
    L.mq_store_export.restype = C.c_int; L.mq_store_export.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp]
    L.mq_store_import.restype = C.c_int; L.mq_store_import.argtypes = [vp, vp, vp, C.c_uint64, vp, C.c_uint32]
    L.mq_index_freeze.restype = C.c_int; L.mq_index_freeze.argtypes = [vp, vp, C.c_uint32, u64p, u64p]
    L.mq_index_nb_mers.restype = C.c_int; L.mq_index_nb_mers.argtypes = [vp, vp, C.c_uint32]
    L.mq_map_batch.restype = C.c_int; L.mq_map_batch.argtypes = [vp, vp, vp, C.c_uint32, vp]
    L.mq_map_batch_device.restype = C.c_int
    L.mq_map_batch_device.argtypes = [vp, vp, vp, C.c_uint32, vp]
    pk = C.POINTER(Packed)
    L.mq_create_multi.restype = C.c_int; L.mq_create_multi.argtypes = [C.POINTER(vp), C.POINTER(Params), C.POINTER(C.c_int), C.c_int]
    L.mq_device_count.restype = C.c_int; L.mq_device_count.argtypes = [vp]
    L.mq_packed_words.restype = C.c_uint64; L.mq_packed_words.argtypes = [C.c_uint64]
    L.mq_packed_flag_words.restype = C.c_uint64; L.mq_packed_flag_words.argtypes = [C.c_uint64]
    L.mq_pack.restype = C.c_int; L.mq_pack.argtypes = [vp, C.c_uint64, vp, vp, vp, C.c_uint64, u64p, C.c_int, C.c_int]
    L.mq_pack_at.restype = C.c_int; L.mq_pack_at.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp, vp, C.c_uint64, u64p, C.c_int]
    L.mq_unpack.restype = C.c_int; L.mq_unpack.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
    L.mq_index_add_packed.restype = C.c_int; L.mq_index_add_packed.argtypes = [vp, pk, vp, C.c_uint32, C.c_uint32, vp]
    L.mq_map_batch_packed.restype = C.c_int; L.mq_map_batch_packed.argtypes = [vp, pk, vp, C.c_uint32, vp]
    L.mq_map_batch_packed_device.restype = C.c_int; L.mq_map_batch_packed_device.argtypes = [vp, pk, vp, C.c_uint32, vp]
    L.mq_minimizers_packed.restype = C.c_int
    L.mq_minimizers_packed.argtypes = [vp, pk, vp, C.c_uint32, vp, vp, vp, C.c_uint64, u64p]
    L.mq_store_reserve.restype = C.c_int; L.mq_store_reserve.argtypes = [vp, C.c_uint64, C.POINTER(vp), C.POINTER(vp)]
    L.mq_store_commit.restype = C.c_int; L.mq_store_commit.argtypes = [vp, C.c_uint64, vp, C.c_uint32]
    L.mq_format_paf.restype = C.c_int
    L.mq_format_paf.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, vp]
    L.mq_minimizers.restype = C.c_int
    L.mq_minimizers.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, vp, C.c_uint64, u64p]
    L.mq_kminmers.restype = C.c_int
    L.mq_kminmers.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, vp, vp, vp, C.c_uint64, u64p]
    L.mq_index_get.restype = C.c_int; L.mq_index_get.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, vp, vp, vp]
    L.mq_matches.restype = C.c_int; L.mq_matches.argtypes = [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint64, u64p]
    L.mq_last_ms.restype = C.c_double; L.mq_last_ms.argtypes = [vp, C.c_char_p]
    L.mq_launch_count.restype = C.c_uint64; L.mq_launch_count.argtypes = [vp]
    L.mq_stream.restype = vp; L.mq_stream.argtypes = [vp]
    L.mq_sync.restype = C.c_int; L.mq_sync.argtypes = [vp]
    L.mq_table_bytes.restype = C.c_uint64; L.mq_table_bytes.argtypes = [vp]
    L.mq_table_slots.restype = C.c_uint64; L.mq_table_slots.argtypes = [vp]
    L.mq_scan_kernel_launches.restype = C.c_uint64; L.mq_scan_kernel_launches.argtypes = [vp]
    L.mq_set_host_threads.restype = C.c_int; L.mq_set_host_threads.argtypes = [vp, C.c_int]
    L.mq_last_counter.restype = C.c_uint64; L.mq_last_counter.argtypes = [vp, C.c_char_p]
    L.mq_minimizer_count.restype = C.c_uint64; L.mq_minimizer_count.argtypes = [vp, C.c_int]
    L.mq_dev_alloc.restype = vp; L.mq_dev_alloc.argtypes = [vp, C.c_size_t]
    L.mq_dev_free.restype = None; L.mq_dev_free.argtypes = [vp, vp]
    L.mq_dev_upload.restype = C.c_int; L.mq_dev_upload.argtypes = [vp, vp, vp, C.c_size_t]
    L.mq_dev_download.restype = C.c_int; L.mq_dev_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.mq_dev_memset.restype = C.c_int; L.mq_dev_memset.argtypes = [vp, vp, C.c_int, C.c_size_t]
    L.mq_region_begin.restype = C.c_int; L.mq_region_begin.argtypes = [vp]
    L.mq_region_end_ms.restype = C.c_double; L.mq_region_end_ms.argtypes = [vp]
    L.mq_total_ms.restype = C.c_double; L.mq_total_ms.argtypes = [vp, C.c_char_p]
    L.mq_index_save.restype = C.c_int; L.mq_index_save.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_uint64]
    L.mq_index_load.restype = C.c_int
    L.mq_index_load.argtypes = [vp, C.c_char_p, vp, C.c_uint32, C.POINTER(C.c_uint32), vp, C.c_uint64, u64p, u64p]
    _lib = L
    return L
