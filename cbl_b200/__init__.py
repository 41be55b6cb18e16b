"""cbl_b200 — B200-native (sm_100a) implementation of CBL's batched sequence path.

The product is the CUDA library ``cbl_b200/csrc/libcbl_gpu.so`` behind the C ABI of
``include/cbl_gpu.h``; this package is the thin host mirror of the reference's ``CBL<K,T,PREFIX_BITS>``.
There is no CPU fallback: importing fails if the CUDA library has not been built.
"""
from ._lib import LIB_PATH, lib
from .cbl import CBL, CBLError, OP_AND, OP_OR, OP_SUB, OP_XOR, concat_records, launch_count, profile_enable, profile_report, sort_fallback_count

lib()  # fail loudly at import time if the native library is missing

__all__ = ["CBL", "CBLError", "concat_records", "launch_count", "sort_fallback_count", "profile_enable", "profile_report", "lib", "LIB_PATH", "OP_OR", "OP_AND", "OP_SUB", "OP_XOR"]
