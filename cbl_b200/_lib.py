"""ctypes binding of libcbl_gpu.so (include/cbl_gpu.h).  There is NO fallback: if the CUDA library
is missing or fails to load, importing the product fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CBL_GPU_LIB: developer override to load an experimental build of the same library (still CUDA-only)
LIB_PATH = os.environ.get("CBL_GPU_LIB") or os.path.join(_HERE, "csrc", "libcbl_gpu.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
szp = C.POINTER(C.c_size_t)
i32p = C.POINTER(C.c_int32)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)

# every symbol include/cbl_gpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cbl_create": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, vpp]),
    "cbl_create_sharded": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, i32p, vpp]),
    "cbl_create_sharded_ex": (C.c_int32, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, i32p, u32p, vpp]),
    "cbl_sharded_splitters": (C.c_int32, [vp, u32p, C.c_size_t, szp]),
    "cbl_destroy": (C.c_int32, [vp]),
    "cbl_clone": (C.c_int32, [vp, vpp]),
    "cbl_last_error": (C.c_char_p, [vp]),
    "cbl_last_global_error": (C.c_char_p, []),
    "cbl_count": (C.c_int32, [vp, u64p]),
    "cbl_is_empty": (C.c_int32, [vp, i32p]),
    "cbl_is_canonical": (C.c_int32, [vp, i32p]),
    "cbl_num_buckets": (C.c_int32, [vp, u64p]),
    "cbl_insert_seq": (C.c_int32, [vp, vp, C.c_size_t]),
    "cbl_remove_seq": (C.c_int32, [vp, vp, C.c_size_t]),
    "cbl_contains_seq": (C.c_int32, [vp, vp, C.c_size_t, vp, szp]),
    "cbl_contains_all": (C.c_int32, [vp, vp, C.c_size_t, i32p]),
    "cbl_insert_seqs": (C.c_int32, [vp, vp, u64p, C.c_size_t]),
    "cbl_remove_seqs": (C.c_int32, [vp, vp, u64p, C.c_size_t]),
    "cbl_contains_seqs": (C.c_int32, [vp, vp, u64p, C.c_size_t, vp]),
    "cbl_insert_seqs_dev": (C.c_int32, [vp, vp, u64p, C.c_size_t]),
    "cbl_remove_seqs_dev": (C.c_int32, [vp, vp, u64p, C.c_size_t]),
    "cbl_contains_seqs_dev": (C.c_int32, [vp, vp, u64p, C.c_size_t, vp]),
    "cbl_count_kmers": (C.c_int32, [vp, u64p, C.c_size_t, u64p]),
    "cbl_last_kmer_count": (C.c_int32, [vp, u64p]),
    "cbl_contains_kmers": (C.c_int32, [vp, u64p, u64p, C.c_size_t, vp]),
    "cbl_insert_kmers": (C.c_int32, [vp, u64p, u64p, C.c_size_t, vp]),
    "cbl_remove_kmers": (C.c_int32, [vp, u64p, u64p, C.c_size_t, vp]),
    "cbl_setop": (C.c_int32, [C.c_int32, vp, vp, vpp]),
    "cbl_setop_assign": (C.c_int32, [C.c_int32, vp, vp]),
    "cbl_merge_many": (C.c_int32, [vpp, C.c_size_t, vpp]),
    "cbl_intersect_many": (C.c_int32, [vpp, C.c_size_t, vpp]),
    "cbl_export_words": (C.c_int32, [vp, C.c_uint64, u64p, u64p, C.c_size_t, szp]),
    "cbl_export_kmers": (C.c_int32, [vp, C.c_uint64, u64p, u64p, C.c_size_t, szp]),
    "cbl_bucket_sizes": (C.c_int32, [vp, u32p, u32p, C.c_size_t, szp]),
    "cbl_serialize_size": (C.c_int32, [vp, szp]),
    "cbl_serialize": (C.c_int32, [vp, vp, C.c_size_t, szp]),
    "cbl_deserialize": (C.c_int32, [vp, vp, C.c_size_t, vpp]),
    "cbl_deserialize_range": (C.c_int32, [vp, vp, C.c_size_t, C.c_uint64, C.c_uint64, vpp]),
    "cbl_save_to_file": (C.c_int32, [vp, C.c_char_p]),
    "cbl_load_from_file": (C.c_int32, [vp, C.c_char_p, vpp]),
    "cbl_seq_words_dev": (C.c_int32, [vp, vp, u64p, C.c_size_t, vp]),
    "cbl_words_op_dev": (C.c_int32, [vp, C.c_int32, vp, C.c_size_t, vp]),
    "cbl_words_op_segments_dev": (C.c_int32, [vp, C.c_int32, vpp, u64p, C.c_uint32]),
    "cbl_words_contains_segments_dev": (C.c_int32, [vp, vpp, u64p, vpp, C.c_uint32]),
    "cbl_export_words_dev": (C.c_int32, [vp, C.c_uint64, C.c_uint64, vp]),
    "cbl_route_words_dev": (C.c_int32, [vp, vp, C.c_size_t, u32p, C.c_uint32, vp, vp, u64p]),
    "cbl_gather_u8_dev": (C.c_int32, [vp, vp, vp, C.c_size_t, vp]),
    "cbl_peer_alloc": (C.c_int32, [vp, C.c_size_t, vpp, vp]),
    "cbl_peer_open": (C.c_int32, [vp, vp, vpp]),
    "cbl_peer_close": (C.c_int32, [vp, vp]),
    "cbl_peer_free": (C.c_int32, [vp, vp]),
    "cbl_route_counts_dev": (C.c_int32, [vp, vp, C.c_size_t, u32p, C.c_uint32, u64p]),
    "cbl_route_scatter_dev": (C.c_int32, [vp, vp, C.c_size_t, u32p, C.c_uint32, vpp, u64p, u64p, vp]),
    "cbl_seq_route_dev": (C.c_int32, [vp, vp, u64p, C.c_size_t, u32p, C.c_uint32, vpp, C.c_uint64, vp, u64p]),
    "cbl_seq_contains_fused_dev": (C.c_int32, [vp, vp, u64p, C.c_size_t, u32p, C.c_uint32, vpp, vpp, C.c_uint64, vp, vpp, vpp, vpp, C.c_uint32, u64p]),
    "cbl_peer_fill": (C.c_int32, [vp, vp, C.c_int32, C.c_size_t]),
    "cbl_peer_zero": (C.c_int32, [vp, vp, C.c_size_t]),
    "cbl_word_bytes": (C.c_int32, [vp, i32p]),
    "cbl_suffix_bits": (C.c_int32, [vp, i32p]),
    "cbl_seq_words": (C.c_int32, [vp, vp, u64p, C.c_size_t, u64p, u64p, C.c_int32]),
    "cbl_sync": (C.c_int32, [vp]),
    "cbl_stream": (vp, [vp]),
    "cbl_launch_count": (C.c_uint64, []),
    "cbl_sort_fallback_count": (C.c_uint64, []),
    "cbl_build_info": (C.c_char_p, []),
    "cbl_set_sort_concentration": (C.c_int32, [vp, C.c_double]),
    "cbl_mem_trim": (C.c_int32, [C.c_int32]),
    "cbl_mem_cached_bytes": (C.c_uint64, []),
    "cbl_profile_enable": (None, [C.c_int32]),
    "cbl_profile_report": (C.c_int32, [C.c_char_p, C.c_size_t]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA library first (`make -C cbl_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
