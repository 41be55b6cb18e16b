"""ctypes front-end of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module (see oracle/cbl_oracle.hpp).  The product package ``cbl_b200`` never does.

``load(prefer_ref=True)`` returns the library linked against the reference's own C++ half
(oracle/_ref/liboracle_ref.so) when it has been built, else the stand-alone restatement
(oracle/liboracle.so).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)


def _ptr(a: Optional[np.ndarray], ty):
    if a is None:
        return C.cast(None, ty)
    return a.ctypes.data_as(ty)


def lib_paths() -> Tuple[str, str]:
    return (os.path.join(_HERE, "_ref", "liboracle_ref.so"), os.path.join(_HERE, "liboracle.so"))


_LIBS = {}


def load(prefer_ref: bool = True) -> C.CDLL:
    ref, plain = lib_paths()
    order = [ref, plain] if prefer_ref else [plain, ref]
    for p in order:
        if os.path.exists(p):
            if p not in _LIBS:
                _LIBS[p] = _declare(C.CDLL(p))
            return _LIBS[p]
    raise RuntimeError("oracle library not built: run `make -C oracle` (and `make -C oracle ref`)")


def _declare(L: C.CDLL) -> C.CDLL:
    vp, i, sz, u64, u32, dbl = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64, C.c_uint32, C.c_double
    szp = C.POINTER(C.c_size_t)
    sig = {
        "orc_last_error": (C.c_char_p, []),
        "orc_uses_reference_cxx": (i, []),
        "orc_cbl_create": (i, [i, i, i, i, C.POINTER(vp)]),
        "orc_cbl_destroy": (None, [vp]),
        "orc_cbl_clone": (i, [vp, C.POINTER(vp)]),
        "orc_cbl_count": (u64, [vp]),
        "orc_cbl_is_empty": (i, [vp]),
        "orc_cbl_is_canonical": (i, [vp]),
        "orc_cbl_seq_words": (i, [vp, _u8p, sz, _u64p, _u64p, sz, szp]),
        "orc_cbl_insert_seq": (i, [vp, _u8p, sz]),
        "orc_cbl_remove_seq": (i, [vp, _u8p, sz]),
        "orc_cbl_contains_seq": (i, [vp, _u8p, sz, _u8p, sz, szp]),
        "orc_cbl_contains_all": (i, [vp, _u8p, sz, C.POINTER(i)]),
        "orc_cbl_insert": (i, [vp, u64, u64]),
        "orc_cbl_remove": (i, [vp, u64, u64]),
        "orc_cbl_contains": (i, [vp, u64, u64]),
        "orc_cbl_get_word": (None, [vp, u64, u64, _u64p, _u64p]),
        "orc_cbl_recover_kmer": (None, [vp, u64, u64, _u64p, _u64p]),
        "orc_cbl_iter_words": (i, [vp, i, _u64p, _u64p, sz, szp]),
        "orc_cbl_binary_op": (i, [i, vp, vp, C.POINTER(vp)]),
        "orc_cbl_assign_op": (i, [i, vp, vp]),
        "orc_cbl_merge_many": (i, [C.POINTER(vp), sz, i, C.POINTER(vp)]),
        "orc_cbl_serialize": (i, [vp, _u8p, sz, szp]),
        "orc_cbl_deserialize": (i, [vp, _u8p, sz, C.POINTER(vp)]),
        "orc_cbl_bucket_sizes": (i, [vp, _u64p, _u64p, sz, szp]),
        "orc_cbl_time_insert_seqs": (dbl, [vp, _u8p, _u64p, sz]),
        "orc_cbl_time_contains_seqs": (dbl, [vp, _u8p, _u64p, sz, _u64p]),
        "orc_revcomp_nucs": (i, [i, i, _u8p, _u8p]),
        "orc_kmer_from_nucs": (None, [i, _u8p, sz, _u64p, _u64p]),
        "orc_kmer_revcomp": (None, [i, i, u64, u64, _u64p, _u64p]),
        "orc_necklace_pos": (None, [i, u64, u64, _u64p, _u64p, _u64p]),
        "orc_revert_necklace_pos": (None, [i, u64, u64, u64, _u64p, _u64p]),
        "orc_necklace_queue_vs_brute": (u64, [i, i, i, _u64p, _u64p, sz]),
        "orc_queue_new": (vp, [i, i, i, i]),
        "orc_queue_free": (None, [vp]),
        "orc_queue_insert_full": (None, [vp, u64, u64]),
        "orc_queue_insert": (None, [vp, u64]),
        "orc_queue_insert2": (None, [vp, u64]),
        "orc_queue_get": (None, [vp, _u64p, _u64p, _u64p]),
        "orc_lexmin_new": (vp, [i]),
        "orc_lexmin_free": (None, [vp]),
        "orc_lexmin_insert_full": (None, [vp, _u64p, sz]),
        "orc_lexmin_insert": (None, [vp, u64]),
        "orc_lexmin_insert2": (None, [vp, u64, u64]),
        "orc_lexmin_min_pos": (sz, [vp, _u64p, sz]),
        "orc_bv_new": (vp, [i]),
        "orc_bv_free": (None, [vp]),
        "orc_bv_insert": (i, [vp, u64]),
        "orc_bv_remove": (i, [vp, u64]),
        "orc_bv_contains": (i, [vp, u64]),
        "orc_bv_rank": (u64, [vp, u64]),
        "orc_bv_count": (u64, [vp]),
        "orc_bv_iter": (sz, [vp, _u64p, sz]),
        "orc_bv_assign_op": (i, [vp, i, vp]),
        "orc_tiered_new": (vp, []),
        "orc_tiered_free": (None, [vp]),
        "orc_tiered_insert": (None, [vp, u64, u32]),
        "orc_tiered_remove": (None, [vp, u64]),
        "orc_tiered_get": (u32, [vp, u64]),
        "orc_tiered_len": (u64, [vp]),
        "orc_trie3_new": (vp, []),
        "orc_trie3_free": (None, [vp]),
        "orc_trie3_insert": (i, [vp, _u8p]),
        "orc_trie3_remove": (i, [vp, _u8p]),
        "orc_trie3_contains": (i, [vp, _u8p]),
        "orc_trie3_is_empty": (i, [vp]),
        "orc_trie3_count": (u64, [vp]),
        "orc_trie3_iter": (sz, [vp, _u8p, sz]),
        "orc_sliced3_roundtrip": (u64, [u64]),
        "orc_sliced3_cmp": (i, [u64, u64]),
        "orc_ws1_new": (vp, [i, i]),
        "orc_ws1_free": (None, [vp]),
        "orc_ws1_insert": (i, [vp, u64]),
        "orc_ws1_remove": (i, [vp, u64]),
        "orc_ws1_contains": (i, [vp, u64]),
        "orc_ws1_count": (u64, [vp]),
        "orc_ws1_is_empty": (i, [vp]),
        "orc_ws1_insert_batch": (None, [vp, _u64p, sz]),
        "orc_ws1_remove_batch": (None, [vp, _u64p, sz]),
        "orc_ws1_contains_batch": (None, [vp, _u64p, sz, _u8p]),
        "orc_ws1_iter": (sz, [vp, _u64p, sz]),
        "orc_ws1_binary_op": (vp, [i, vp, vp]),
        "orc_ws1_assign_op": (None, [i, vp, vp]),
        "orc_ws1_merge_many": (vp, [C.POINTER(vp), sz, i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


def _seq_arr(seq) -> np.ndarray:
    if isinstance(seq, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    a = np.ascontiguousarray(seq, dtype=np.uint8)
    return a


def words_to_int(lo: np.ndarray, hi: Optional[np.ndarray]) -> List[int]:
    if hi is None:
        return [int(x) for x in lo]
    return [int(l) | (int(h) << 64) for l, h in zip(lo, hi)]


class OracleError(RuntimeError):
    pass


class OracleCBL:
    """The restated reference ``CBL::<K, T, PREFIX_BITS>`` (src/cbl.rs) on the CPU."""

    OR, AND, SUB, XOR = 0, 1, 2, 3

    def __init__(self, k: int, t_bits: int, prefix_bits: int = 24, canonical: bool = False, *, lib=None, _handle=None):
        self.L = lib or load()
        self.k, self.t_bits, self.prefix_bits = k, t_bits, prefix_bits
        if _handle is not None:
            self.h = _handle
        else:
            h = C.c_void_p()
            if self.L.orc_cbl_create(k, t_bits, prefix_bits, int(canonical), C.byref(h)):
                raise OracleError(self.L.orc_last_error().decode())
            self.h = h

    def _wrap(self, h) -> "OracleCBL":
        return OracleCBL(self.k, self.t_bits, self.prefix_bits, lib=self.L, _handle=h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.orc_cbl_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise OracleError(self.L.orc_last_error().decode())

    def clone(self) -> "OracleCBL":
        h = C.c_void_p()
        self._chk(self.L.orc_cbl_clone(self.h, C.byref(h)))
        return self._wrap(h)

    def count(self) -> int:
        return int(self.L.orc_cbl_count(self.h))

    def is_empty(self) -> bool:
        return bool(self.L.orc_cbl_is_empty(self.h))

    def is_canonical(self) -> bool:
        return bool(self.L.orc_cbl_is_canonical(self.h))

    def seq_words(self, seq) -> Tuple[np.ndarray, np.ndarray]:
        s = _seq_arr(seq)
        cap = max(len(s), 1)
        lo = np.zeros(cap, dtype=np.uint64)
        hi = np.zeros(cap, dtype=np.uint64)
        n = C.c_size_t()
        self._chk(self.L.orc_cbl_seq_words(self.h, _ptr(s, _u8p), len(s), _ptr(lo, _u64p), _ptr(hi, _u64p), cap, C.byref(n)))
        return lo[: n.value].copy(), hi[: n.value].copy()

    def insert_seq(self, seq):
        s = _seq_arr(seq)
        self._chk(self.L.orc_cbl_insert_seq(self.h, _ptr(s, _u8p), len(s)))

    def remove_seq(self, seq):
        s = _seq_arr(seq)
        self._chk(self.L.orc_cbl_remove_seq(self.h, _ptr(s, _u8p), len(s)))

    def contains_seq(self, seq) -> np.ndarray:
        s = _seq_arr(seq)
        cap = max(len(s), 1)
        out = np.zeros(cap, dtype=np.uint8)
        n = C.c_size_t()
        self._chk(self.L.orc_cbl_contains_seq(self.h, _ptr(s, _u8p), len(s), _ptr(out, _u8p), cap, C.byref(n)))
        return out[: n.value].copy()

    def contains_all(self, seq) -> bool:
        s = _seq_arr(seq)
        r = C.c_int()
        self._chk(self.L.orc_cbl_contains_all(self.h, _ptr(s, _u8p), len(s), C.byref(r)))
        return bool(r.value)

    @staticmethod
    def _split(x: int) -> Tuple[int, int]:
        return x & 0xFFFFFFFFFFFFFFFF, (x >> 64) & 0xFFFFFFFFFFFFFFFF

    def insert(self, kmer: int) -> bool:
        return bool(self.L.orc_cbl_insert(self.h, *self._split(kmer)))

    def remove(self, kmer: int) -> bool:
        return bool(self.L.orc_cbl_remove(self.h, *self._split(kmer)))

    def contains(self, kmer: int) -> bool:
        return bool(self.L.orc_cbl_contains(self.h, *self._split(kmer)))

    def get_word(self, kmer: int) -> int:
        lo, hi = C.c_uint64(), C.c_uint64()
        self.L.orc_cbl_get_word(self.h, *self._split(kmer), C.byref(lo), C.byref(hi))
        return lo.value | (hi.value << 64)

    def recover_kmer(self, word: int) -> int:
        lo, hi = C.c_uint64(), C.c_uint64()
        self.L.orc_cbl_recover_kmer(self.h, *self._split(word), C.byref(lo), C.byref(hi))
        return lo.value | (hi.value << 64)

    def iter_words(self, sorted_: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        cap = max(self.count(), 1)
        lo = np.zeros(cap, dtype=np.uint64)
        hi = np.zeros(cap, dtype=np.uint64)
        n = C.c_size_t()
        self._chk(self.L.orc_cbl_iter_words(self.h, int(sorted_), _ptr(lo, _u64p), _ptr(hi, _u64p), cap, C.byref(n)))
        return lo[: n.value].copy(), hi[: n.value].copy()

    def binary_op(self, op: int, other: "OracleCBL") -> "OracleCBL":
        h = C.c_void_p()
        self._chk(self.L.orc_cbl_binary_op(op, self.h, other.h, C.byref(h)))
        return self._wrap(h)

    def assign_op(self, op: int, other: "OracleCBL"):
        self._chk(self.L.orc_cbl_assign_op(op, self.h, other.h))

    def __or__(self, o): return self.binary_op(self.OR, o)
    def __and__(self, o): return self.binary_op(self.AND, o)
    def __sub__(self, o): return self.binary_op(self.SUB, o)
    def __xor__(self, o): return self.binary_op(self.XOR, o)
    def __ior__(self, o): self.assign_op(self.OR, o); return self
    def __iand__(self, o): self.assign_op(self.AND, o); return self
    def __isub__(self, o): self.assign_op(self.SUB, o); return self
    def __ixor__(self, o): self.assign_op(self.XOR, o); return self

    @staticmethod
    def _many(sets: Sequence["OracleCBL"], intersect: bool) -> "OracleCBL":
        arr = (C.c_void_p * len(sets))(*[s.h for s in sets])
        h = C.c_void_p()
        sets[0]._chk(sets[0].L.orc_cbl_merge_many(arr, len(sets), int(intersect), C.byref(h)))
        return sets[0]._wrap(h)

    @staticmethod
    def merge(sets): return OracleCBL._many(sets, False)

    @staticmethod
    def intersect(sets): return OracleCBL._many(sets, True)

    def serialize(self) -> bytes:
        n = C.c_size_t()
        self._chk(self.L.orc_cbl_serialize(self.h, C.cast(None, _u8p), 0, C.byref(n)))
        buf = np.zeros(max(n.value, 1), dtype=np.uint8)
        self._chk(self.L.orc_cbl_serialize(self.h, _ptr(buf, _u8p), len(buf), C.byref(n)))
        return buf[: n.value].tobytes()

    def deserialize(self, data: bytes) -> "OracleCBL":
        a = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        self._chk(self.L.orc_cbl_deserialize(self.h, _ptr(a, _u8p), len(a), C.byref(h)))
        return self._wrap(h)

    def bucket_sizes(self) -> Tuple[np.ndarray, np.ndarray]:
        n = C.c_size_t()
        self._chk(self.L.orc_cbl_bucket_sizes(self.h, C.cast(None, _u64p), C.cast(None, _u64p), 0, C.byref(n)))
        p = np.zeros(max(n.value, 1), dtype=np.uint64)
        s = np.zeros(max(n.value, 1), dtype=np.uint64)
        self._chk(self.L.orc_cbl_bucket_sizes(self.h, _ptr(p, _u64p), _ptr(s, _u64p), len(p), C.byref(n)))
        return p[: n.value], s[: n.value]

    # timed legs for bench.py (1 host core, reference data structures on the timed path)
    def time_insert_seqs(self, buf: np.ndarray, offsets: np.ndarray) -> float:
        return float(self.L.orc_cbl_time_insert_seqs(self.h, _ptr(buf, _u8p), _ptr(offsets, _u64p), len(offsets) - 1))

    def time_contains_seqs(self, buf: np.ndarray, offsets: np.ndarray) -> Tuple[float, int]:
        npos = C.c_uint64()
        t = float(self.L.orc_cbl_time_contains_seqs(self.h, _ptr(buf, _u8p), _ptr(offsets, _u64p), len(offsets) - 1, C.byref(npos)))
        return t, int(npos.value)
