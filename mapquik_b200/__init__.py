"""mapquik_b200 -- B200-native (sm_100a CUDA) implementation of mapquik's seeding->chaining hot path.

The product is libmapquik_b200.so (C ABI, include/mapquik_b200.h).  `Index`/`Params` mirror the
reference's call sites (closures.rs:48,94,102) over that ABI.
"""
from .mapper import Index, PackedSeqs, Params, concat, to_upper_u8  # noqa: F401
from .capi import EXC_DTYPE, HIT_DTYPE, MqError  # noqa: F401
