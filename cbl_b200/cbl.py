"""Host-side mirror of the reference's public type ``CBL<K, T, PREFIX_BITS>`` (src/cbl.rs:40-569) on
top of the C ABI (include/cbl_gpu.h).  Method names, argument meaning and error behaviour follow the
reference: where the Rust code panics this raises ``CBLError`` with the same message.

The Rust toolchain is absent from the build image, so this Python class (and the C++ facade
include/cbl.hpp) stand where the reference's ``src/cbl.rs`` would sit above the ABI; INTEGRATION.md
shows the Rust binding.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import u8p, u32p, u64p
from ._lib import vpp as vpp_t

Seq = Union[bytes, bytearray, memoryview, np.ndarray]

OP_OR, OP_AND, OP_SUB, OP_XOR = 0, 1, 2, 3


class CBLError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


def _as_u8(seq: Seq) -> np.ndarray:
    if isinstance(seq, np.ndarray):
        if seq.dtype != np.uint8:
            raise TypeError("sequence arrays must be uint8")
        return np.ascontiguousarray(seq)
    return np.frombuffer(bytes(seq) if not isinstance(seq, (bytes, bytearray)) else seq, dtype=np.uint8)


def _split_kmers(kmers: Iterable[int]) -> Tuple[np.ndarray, np.ndarray]:
    ks = list(kmers)
    lo = np.fromiter((k & 0xFFFFFFFFFFFFFFFF for k in ks), dtype=np.uint64, count=len(ks))
    hi = np.fromiter(((k >> 64) & 0xFFFFFFFFFFFFFFFF for k in ks), dtype=np.uint64, count=len(ks))
    return lo, hi


def concat_records(records: Sequence[Seq]) -> Tuple[np.ndarray, np.ndarray]:
    arrs = [_as_u8(r) for r in records]
    offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offsets[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
    buf = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
    return buf, offsets


class CBL:
    """A fully dynamic set of k-mers resident in the memory of one B200.

    ``CBL(k, t_bits, prefix_bits=24)`` ~ ``CBL::<K, T, PREFIX_BITS>::new()``;
    ``CBL.new_canonical(k, t_bits, prefix_bits)`` ~ ``::new_canonical()``.
    """

    def __init__(self, k: int, t_bits: int, prefix_bits: int = 24, canonical: bool = False, device: int = 0, *, _handle=None):
        self._L = _lib.lib()
        self.k, self.t_bits, self.prefix_bits, self.device = k, t_bits, prefix_bits, device
        if _handle is not None:
            self._h = _handle
            return
        h = C.c_void_p()
        rc = self._L.cbl_create(k, t_bits, prefix_bits, int(canonical), device, C.byref(h))
        if rc:
            raise CBLError(rc, self._L.cbl_last_global_error().decode())
        self._h = h

    @classmethod
    def new_canonical(cls, k: int, t_bits: int, prefix_bits: int = 24, device: int = 0) -> "CBL":
        return cls(k, t_bits, prefix_bits, True, device)

    @classmethod
    def sharded(cls, k: int, t_bits: int, prefix_bits: int = 24, canonical: bool = False, devices: Sequence[int] = (0,),
                splitters: Optional[Sequence[int]] = None) -> "CBL":
        """One set prefix-sharded over ``devices`` (GPUs of THIS process) behind the same handle type: every host-buffer
        method works unchanged (``cbl_create_sharded``); the ``*_dev`` methods raise."""
        L = _lib.lib()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        if splitters is None:
            rc = L.cbl_create_sharded(k, t_bits, prefix_bits, int(canonical), len(devices), devs, C.byref(h))
        else:
            sp = (C.c_uint32 * max(len(splitters), 1))(*[int(x) for x in splitters])
            rc = L.cbl_create_sharded_ex(k, t_bits, prefix_bits, int(canonical), len(devices), devs, sp, C.byref(h))
        if rc:
            raise CBLError(rc, L.cbl_last_global_error().decode())
        return cls(k, t_bits, prefix_bits, canonical, int(devices[0]), _handle=h)

    def shard_splitters(self) -> np.ndarray:
        n = C.c_size_t()
        self._chk(self._L.cbl_sharded_splitters(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), dtype=np.uint32)
        self._chk(self._L.cbl_sharded_splitters(self._h, out.ctypes.data_as(u32p), len(out), C.byref(n)))
        return out[: n.value]

    # -- plumbing --------------------------------------------------------------------------------
    def _wrap(self, h) -> "CBL":
        return CBL(self.k, self.t_bits, self.prefix_bits, device=self.device, _handle=h)

    def _chk(self, rc: int):
        if rc:
            raise CBLError(rc, self._L.cbl_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.cbl_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    # -- scalar queries (src/cbl.rs:162-177) -----------------------------------------------------
    def is_canonical(self) -> bool:
        v = C.c_int32()
        self._chk(self._L.cbl_is_canonical(self._h, C.byref(v)))
        return bool(v.value)

    def count(self) -> int:
        v = C.c_uint64()
        self._chk(self._L.cbl_count(self._h, C.byref(v)))
        return int(v.value)

    def __len__(self) -> int:
        return self.count()

    def is_empty(self) -> bool:
        v = C.c_int32()
        self._chk(self._L.cbl_is_empty(self._h, C.byref(v)))
        return bool(v.value)

    def num_buckets(self) -> int:
        v = C.c_uint64()
        self._chk(self._L.cbl_num_buckets(self._h, C.byref(v)))
        return int(v.value)

    def word_bytes(self) -> int:
        v = C.c_int32()
        self._chk(self._L.cbl_word_bytes(self._h, C.byref(v)))
        return int(v.value)

    def clone(self) -> "CBL":
        h = C.c_void_p()
        self._chk(self._L.cbl_clone(self._h, C.byref(h)))
        return self._wrap(h)

    # -- sequences (src/cbl.rs:293-354) ----------------------------------------------------------
    def insert_seq(self, seq: Seq) -> None:
        a = _as_u8(seq)
        self._chk(self._L.cbl_insert_seq(self._h, a.ctypes.data, len(a)))

    def remove_seq(self, seq: Seq) -> None:
        a = _as_u8(seq)
        self._chk(self._L.cbl_remove_seq(self._h, a.ctypes.data, len(a)))

    def contains_seq(self, seq: Seq) -> np.ndarray:
        a = _as_u8(seq)
        out = np.zeros(max(len(a) - self.k + 1, 1), dtype=np.uint8)
        n = C.c_size_t()
        self._chk(self._L.cbl_contains_seq(self._h, a.ctypes.data, len(a), out.ctypes.data, C.byref(n)))
        return out[: n.value].astype(bool)

    def contains_all(self, seq: Seq) -> bool:
        a = _as_u8(seq)
        v = C.c_int32()
        self._chk(self._L.cbl_contains_all(self._h, a.ctypes.data, len(a), C.byref(v)))
        return bool(v.value)

    # batches of records: one ABI call for the whole `for record in reader` loop of examples/cbl.rs
    def insert_seqs(self, buf: np.ndarray, offsets: np.ndarray) -> None:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._chk(self._L.cbl_insert_seqs(self._h, buf.ctypes.data, offsets.ctypes.data_as(u64p), len(offsets) - 1))

    def remove_seqs(self, buf: np.ndarray, offsets: np.ndarray) -> None:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._chk(self._L.cbl_remove_seqs(self._h, buf.ctypes.data, offsets.ctypes.data_as(u64p), len(offsets) - 1))

    def count_kmers(self, offsets: np.ndarray) -> int:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        v = C.c_uint64()
        self._chk(self._L.cbl_count_kmers(self._h, offsets.ctypes.data_as(u64p), len(offsets) - 1, C.byref(v)))
        return int(v.value)

    def last_kmer_count(self) -> int:
        """Words / answers produced by the last sequence call (fewer than count_kmers after non-ACGT bytes, F8)."""
        v = C.c_uint64()
        self._chk(self._L.cbl_last_kmer_count(self._h, C.byref(v)))
        return int(v.value)

    def contains_seqs(self, buf: np.ndarray, offsets: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = self.count_kmers(offsets)
        if out is None:
            out = np.zeros(max(n, 1), dtype=np.uint8)
        self._chk(self._L.cbl_contains_seqs(self._h, buf.ctypes.data, offsets.ctypes.data_as(u64p), len(offsets) - 1, out.ctypes.data))
        return out[: self.last_kmer_count()]

    # device-resident buffers: raw device pointers (e.g. torch_tensor.data_ptr())
    def insert_seqs_dev(self, d_buf: int, offsets: np.ndarray) -> None:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._chk(self._L.cbl_insert_seqs_dev(self._h, d_buf, offsets.ctypes.data_as(u64p), len(offsets) - 1))

    def remove_seqs_dev(self, d_buf: int, offsets: np.ndarray) -> None:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._chk(self._L.cbl_remove_seqs_dev(self._h, d_buf, offsets.ctypes.data_as(u64p), len(offsets) - 1))

    def contains_seqs_dev(self, d_buf: int, offsets: np.ndarray, d_out: int) -> None:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._chk(self._L.cbl_contains_seqs_dev(self._h, d_buf, offsets.ctypes.data_as(u64p), len(offsets) - 1, d_out))

    def seq_words_dev(self, d_buf: int, offsets: np.ndarray, d_words: int) -> None:
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self._chk(self._L.cbl_seq_words_dev(self._h, d_buf, offsets.ctypes.data_as(u64p), len(offsets) - 1, d_words))

    def words_op_dev(self, op: int, d_words: int, n: int, d_out: int = 0) -> None:
        self._chk(self._L.cbl_words_op_dev(self._h, op, d_words, n, d_out))

    def route_words_dev(self, d_words: int, n: int, splitters: np.ndarray, d_send: int, d_pos: int = 0) -> np.ndarray:
        """Stable partition of n device words by owner rank; returns the per-destination counts."""
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        counts = np.zeros(len(sp) + 1, dtype=np.uint64)
        self._chk(self._L.cbl_route_words_dev(self._h, d_words, n, sp.ctypes.data_as(u32p), len(sp), d_send, d_pos, counts.ctypes.data_as(u64p)))
        return counts

    # -- fused route + exchange over peer memory (include/cbl_gpu.h) ------------------------------
    IPC_HANDLE_BYTES = 64

    def peer_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte CUDA IPC handle) of a block other processes of the box can map."""
        p = C.c_void_p()
        hb = (C.c_uint8 * self.IPC_HANDLE_BYTES)()
        self._chk(self._L.cbl_peer_alloc(self._h, nbytes, C.byref(p), C.cast(hb, C.c_void_p)))
        return int(p.value), bytes(hb)

    def peer_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        hb = (C.c_uint8 * self.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        self._chk(self._L.cbl_peer_open(self._h, C.cast(hb, C.c_void_p), C.byref(p)))
        return int(p.value)

    def peer_close(self, ptr: int) -> None:
        self._chk(self._L.cbl_peer_close(self._h, ptr))

    def peer_free(self, ptr: int) -> None:
        self._chk(self._L.cbl_peer_free(self._h, ptr))

    def route_counts_dev(self, d_words: int, n: int, splitters: np.ndarray) -> np.ndarray:
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        counts = np.zeros(len(sp) + 1, dtype=np.uint64)
        self._chk(self._L.cbl_route_counts_dev(self._h, d_words, n, sp.ctypes.data_as(u32p), len(sp), counts.ctypes.data_as(u64p)))
        return counts

    def route_scatter_dev(self, d_words: int, n: int, splitters: np.ndarray, peer_recv, recv_offset, counts, d_pos: int = 0) -> None:
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        g = len(sp) + 1
        ptrs = (C.c_void_p * g)(*[int(x) for x in peer_recv])
        ro = np.ascontiguousarray(recv_offset, dtype=np.uint64)
        ct = np.ascontiguousarray(counts, dtype=np.uint64)
        self._chk(self._L.cbl_route_scatter_dev(self._h, d_words, n, sp.ctypes.data_as(u32p), len(sp), C.cast(ptrs, vpp_t), ro.ctypes.data_as(u64p),
                                                ct.ctypes.data_as(u64p), d_pos))

    def seq_route_dev(self, d_buf: int, offsets: np.ndarray, splitters: np.ndarray, peer_region, cap: int, d_pos: int = 0) -> np.ndarray:
        """Fused encode + necklace + route: the words of the records go straight into ``peer_region[d]`` (this rank's
        region, ``cap`` words, inside owner d's receive buffer).  Returns the per-owner counts (> cap = overflow)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        g = len(sp) + 1
        ptrs = (C.c_void_p * g)(*[int(x) for x in peer_region])
        counts = np.zeros(g, dtype=np.uint64)
        self._chk(self._L.cbl_seq_route_dev(self._h, d_buf, offsets.ctypes.data_as(u64p), len(offsets) - 1, sp.ctypes.data_as(u32p), len(sp),
                                            C.cast(ptrs, vpp_t), cap, d_pos, counts.ctypes.data_as(u64p)))
        return counts

    def peer_zero(self, ptr: int, nbytes: int) -> None:
        self._chk(self._L.cbl_peer_zero(self._h, ptr, nbytes))

    def peer_fill(self, ptr: int, byte: int, nbytes: int) -> None:
        self._chk(self._L.cbl_peer_fill(self._h, ptr, byte, nbytes))

    def seq_contains_fused_dev(self, d_buf: int, offsets: np.ndarray, splitters: np.ndarray, peer_region, peer_final, cap: int, d_pos: int,
                               recv_region, answer_region, final_counts, epoch: int) -> np.ndarray:
        """The fused sharded contains_seq of this rank: one kernel whose warps alternate between routing this rank's reads and
        probing the blocks the peers have completed here (include/cbl_gpu.h, csrc/shard_query.cuh).  Returns the per-owner
        word counts (> cap = overflow)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        g = len(sp) + 1

        def arr(xs):
            return C.cast((C.c_void_p * g)(*[int(x) for x in xs]), vpp_t)

        counts = np.zeros(g, dtype=np.uint64)
        self._chk(self._L.cbl_seq_contains_fused_dev(self._h, d_buf, offsets.ctypes.data_as(u64p), len(offsets) - 1, sp.ctypes.data_as(u32p), len(sp),
                                                     arr(peer_region), arr(peer_final), cap, d_pos, arr(recv_region), arr(answer_region),
                                                     arr(final_counts), epoch, counts.ctypes.data_as(u64p)))
        return counts

    def words_op_segments_dev(self, op: int, seg_ptrs, seg_n) -> None:
        g = len(seg_ptrs)
        ptrs = (C.c_void_p * g)(*[int(x) for x in seg_ptrs])
        n = np.ascontiguousarray(seg_n, dtype=np.uint64)
        self._chk(self._L.cbl_words_op_segments_dev(self._h, op, C.cast(ptrs, vpp_t), n.ctypes.data_as(u64p), g))

    def words_contains_segments_dev(self, seg_ptrs, seg_n, out_ptrs) -> None:
        """Membership of the words of several device segments in ONE launch; answers of segment i -> out_ptrs[i]."""
        g = len(seg_ptrs)
        ptrs = (C.c_void_p * g)(*[int(x) for x in seg_ptrs])
        outs = (C.c_void_p * g)(*[int(x) for x in out_ptrs])
        n = np.ascontiguousarray(seg_n, dtype=np.uint64)
        self._chk(self._L.cbl_words_contains_segments_dev(self._h, C.cast(ptrs, vpp_t), n.ctypes.data_as(u64p), C.cast(outs, vpp_t), g))

    def gather_u8_dev(self, d_src: int, d_pos: int, n: int, d_out: int) -> None:
        self._chk(self._L.cbl_gather_u8_dev(self._h, d_src, d_pos, n, d_out))

    def export_words_dev(self, start: int, count: int, d_out: int) -> None:
        self._chk(self._L.cbl_export_words_dev(self._h, start, count, d_out))

    # -- single k-mers (src/cbl.rs:219-235); k-mers are IntKmer integers -------------------------
    def _kmers(self, fn, kmers: Iterable[int], want: bool) -> np.ndarray:
        lo, hi = _split_kmers(kmers)
        out = np.zeros(max(len(lo), 1), dtype=np.uint8)
        self._chk(fn(self._h, lo.ctypes.data_as(u64p), hi.ctypes.data_as(u64p), len(lo), out.ctypes.data if want else None))
        return out[: len(lo)].astype(bool)

    def contains(self, kmer: int) -> bool:
        return bool(self._kmers(self._L.cbl_contains_kmers, [kmer], True)[0])

    def insert(self, kmer: int) -> bool:
        """Returns True if the k-mer was absent (src/cbl.rs:226-228)."""
        return not bool(self._kmers(self._L.cbl_insert_kmers, [kmer], True)[0])

    def remove(self, kmer: int) -> bool:
        """Returns True if the k-mer was present (src/cbl.rs:233-235)."""
        return bool(self._kmers(self._L.cbl_remove_kmers, [kmer], True)[0])

    def contains_kmers(self, kmers: Iterable[int]) -> np.ndarray:
        return self._kmers(self._L.cbl_contains_kmers, kmers, True)

    def insert_kmers(self, kmers: Iterable[int]) -> np.ndarray:
        """Inserts all; returns, per k-mer, whether it was in the set before the call."""
        return self._kmers(self._L.cbl_insert_kmers, kmers, True)

    def remove_kmers(self, kmers: Iterable[int]) -> np.ndarray:
        return self._kmers(self._L.cbl_remove_kmers, kmers, True)

    # -- iteration (src/cbl.rs:358-360): ascending word order ------------------------------------
    def _export(self, fn, chunk: int = 1 << 20) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        start, total = 0, self.count()
        lo = np.zeros(min(chunk, max(total, 1)), dtype=np.uint64)
        hi = np.zeros_like(lo)
        n = C.c_size_t()
        while start < total:
            self._chk(fn(self._h, start, lo.ctypes.data_as(u64p), hi.ctypes.data_as(u64p), len(lo), C.byref(n)))
            if n.value == 0:
                break
            yield lo[: n.value].copy(), hi[: n.value].copy()
            start += n.value

    def words(self) -> List[int]:
        """All stored words (necklace << POS_BITS | pos), ascending."""
        out: List[int] = []
        for lo, hi in self._export(self._L.cbl_export_words):
            out.extend(int(a) | (int(b) << 64) for a, b in zip(lo, hi))
        return out

    def words_arrays(self) -> Tuple[np.ndarray, np.ndarray]:
        parts = list(self._export(self._L.cbl_export_words))
        if not parts:
            return np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64)
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def iter(self) -> Iterator[int]:
        """The stored k-mers as IntKmer integers (in canonical mode: the even-parity representative)."""
        for lo, hi in self._export(self._L.cbl_export_kmers):
            for a, b in zip(lo, hi):
                yield int(a) | (int(b) << 64)

    __iter__ = iter

    def buckets_sizes(self) -> Tuple[np.ndarray, np.ndarray]:
        """(prefix, bucket size) pairs in ascending prefix order (src/cbl.rs:370-372)."""
        n = C.c_size_t()
        self._chk(self._L.cbl_bucket_sizes(self._h, None, None, 0, C.byref(n)))
        p = np.zeros(max(n.value, 1), dtype=np.uint32)
        s = np.zeros(max(n.value, 1), dtype=np.uint32)
        if n.value:
            self._chk(self._L.cbl_bucket_sizes(self._h, p.ctypes.data_as(u32p), s.ctypes.data_as(u32p), len(p), C.byref(n)))
        return p[: n.value], s[: n.value]

    def prefix_load(self) -> float:
        return self.num_buckets() / float(1 << self.prefix_bits)  # src/wordset/mod.rs:254-256

    def buckets_size_count(self) -> dict:
        """bucket size -> number of buckets of that size (src/cbl.rs:374-377, src/wordset/mod.rs:265-271)"""
        _, s = self.buckets_sizes()
        sizes, counts = np.unique(s, return_counts=True)
        return {int(a): int(b) for a, b in zip(sizes, counts)}

    def buckets_load_repartition(self) -> dict:
        """bucket size -> share of the stored k-mers held by buckets of that size (src/wordset/mod.rs:273-280)"""
        sc = self.buckets_size_count()
        total = float(sum(k * v for k, v in sc.items())) or 1.0
        return {k: k * v / total for k, v in sc.items()}

    TRIE_THRESHOLD = 1024   # src/wordset/mod.rs:34: a bucket becomes a byte trie above this many suffixes

    def buckets_nodes(self) -> Tuple[np.ndarray, np.ndarray]:
        """(prefix, node count) per bucket as the reference would report it (src/wordset/mod.rs:282-287): a Vec bucket
        counts its elements (src/trievec/mod.rs:37-42), a bucket above the trie threshold counts the nodes of the 256-ary byte
        trie over its big-endian suffixes (src/trie.rs:90-102) = 1 + the number of distinct d-byte heads for d < BYTES —
        a pure function of the bucket's contents, computed here from the sorted suffixes (the GPU keeps no trie)."""
        p, s = self.buckets_sizes()
        nodes = s.astype(np.uint64).copy()
        big = np.flatnonzero(s > self.TRIE_THRESHOLD)
        if len(big):
            kbits = 2 * self.k
            pos_bits = (kbits - 1).bit_length()
            suffix_bits = kbits + pos_bits - self.prefix_bits
            nbytes = (suffix_bits + 7) // 8
            starts = np.concatenate([[0], np.cumsum(s.astype(np.int64))])
            lo = np.zeros(int(s.max()), dtype=np.uint64)
            hi = np.zeros_like(lo)
            n = C.c_size_t()
            for b in big:
                self._chk(self._L.cbl_export_words(self._h, int(starts[b]), lo.ctypes.data_as(u64p), hi.ctypes.data_as(u64p), int(s[b]), C.byref(n)))
                suf = [(int(a) | (int(h) << 64)) & ((1 << suffix_bits) - 1) for a, h in zip(lo[: n.value], hi[: n.value])]
                total = 1
                for d in range(1, nbytes):
                    total += len({x >> (8 * (nbytes - d)) for x in suf})
                nodes[b] = total
        return p, nodes

    def buckets_node_count(self) -> dict:
        """node count -> number of buckets with that many nodes (src/cbl.rs:392-396)"""
        _, n = self.buckets_nodes()
        v, c = np.unique(n, return_counts=True)
        return {int(a): int(b) for a, b in zip(v, c)}

    # -- set operations (src/cbl.rs:411-569, 108-124) --------------------------------------------
    def _binary(self, op: int, other: "CBL") -> "CBL":
        h = C.c_void_p()
        self._chk(self._L.cbl_setop(op, self._h, other._h, C.byref(h)))
        return self._wrap(h)

    def _assign(self, op: int, other: "CBL") -> "CBL":
        self._chk(self._L.cbl_setop_assign(op, self._h, other._h))
        return self

    def __or__(self, o): return self._binary(OP_OR, o)
    def __and__(self, o): return self._binary(OP_AND, o)
    def __sub__(self, o): return self._binary(OP_SUB, o)
    def __xor__(self, o): return self._binary(OP_XOR, o)
    def __ior__(self, o): return self._assign(OP_OR, o)
    def __iand__(self, o): return self._assign(OP_AND, o)
    def __isub__(self, o): return self._assign(OP_SUB, o)
    def __ixor__(self, o): return self._assign(OP_XOR, o)

    @staticmethod
    def _many(fn_name: str, cbls: Sequence["CBL"]) -> "CBL":
        if not cbls:
            raise CBLError(1, "empty list of indexes")
        arr = (C.c_void_p * len(cbls))(*[c._h for c in cbls])
        h = C.c_void_p()
        first = cbls[0]
        first._chk(getattr(first._L, fn_name)(arr, len(cbls), C.byref(h)))
        return first._wrap(h)

    @staticmethod
    def merge(cbls: Sequence["CBL"]) -> "CBL":
        return CBL._many("cbl_merge_many", cbls)

    @staticmethod
    def intersect(cbls: Sequence["CBL"]) -> "CBL":
        return CBL._many("cbl_intersect_many", cbls)

    # -- serde (src/cbl.rs:127-160) --------------------------------------------------------------
    def serialize(self) -> bytes:
        n = C.c_size_t()
        self._chk(self._L.cbl_serialize_size(self._h, C.byref(n)))
        buf = np.zeros(max(n.value, 1), dtype=np.uint8)
        self._chk(self._L.cbl_serialize(self._h, buf.ctypes.data, len(buf), C.byref(n)))
        return buf[: n.value].tobytes()

    def deserialize(self, data: bytes) -> "CBL":
        a = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        self._chk(self._L.cbl_deserialize(self._h, a.ctypes.data, len(a), C.byref(h)))
        return self._wrap(h)

    def deserialize_range(self, data: bytes, prefix_lo: int, prefix_hi: int) -> "CBL":
        """Like ``deserialize`` but keeps only the buckets with prefix in [prefix_lo, prefix_hi) (one rank's range)."""
        a = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        self._chk(self._L.cbl_deserialize_range(self._h, a.ctypes.data, len(a), prefix_lo, prefix_hi, C.byref(h)))
        return self._wrap(h)

    def save_to_file(self, path: str) -> None:
        self._chk(self._L.cbl_save_to_file(self._h, path.encode()))

    def load_from_file(self, path: str) -> "CBL":
        """``CBL::<K,T,P>::load_from_file(path)``; K / T / PREFIX_BITS / device are taken from ``self``."""
        h = C.c_void_p()
        self._chk(self._L.cbl_load_from_file(self._h, path.encode(), C.byref(h)))
        return self._wrap(h)

    # -- diagnostics -----------------------------------------------------------------------------
    def seq_words(self, records: Sequence[Seq], brute: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        buf, offsets = concat_records(records)
        n = self.count_kmers(offsets)
        lo = np.zeros(max(n, 1), dtype=np.uint64)
        hi = np.zeros(max(n, 1), dtype=np.uint64)
        self._chk(self._L.cbl_seq_words(self._h, buf.ctypes.data, offsets.ctypes.data_as(u64p), len(offsets) - 1,
                                        lo.ctypes.data_as(u64p), hi.ctypes.data_as(u64p), int(brute)))
        n = self.last_kmer_count()
        return lo[:n], hi[:n]

    def sync(self) -> None:
        self._chk(self._L.cbl_sync(self._h))

    def set_sort_concentration(self, factor: float) -> None:
        """Planning hint of the batch sort: this handle's words cover ~1 / factor of the prefix mass (a shard of a sharded set)."""
        self._chk(self._L.cbl_set_sort_concentration(self._h, float(factor)))

    def stream_ptr(self) -> int:
        """The cudaStream_t every kernel of this handle is launched on (for CUDA-event timing)."""
        return int(self._L.cbl_stream(self._h) or 0)


def launch_count() -> int:
    return int(_lib.lib().cbl_launch_count())


def sort_fallback_count() -> int:
    """Batches the segment sort handed back to the plain LSD radix passes (heavily repeated words only)."""
    return int(_lib.lib().cbl_sort_fallback_count())


def profile_enable(on: bool) -> None:
    _lib.lib().cbl_profile_enable(int(on))


def profile_report() -> dict:
    """Per-kernel device time accumulated since the last report: {kernel: {"n": launches, "ms": total}}."""
    import json

    buf = C.create_string_buffer(1 << 16)
    rc = _lib.lib().cbl_profile_report(buf, len(buf))
    if rc:
        raise CBLError(rc, _lib.lib().cbl_last_global_error().decode())
    return json.loads(buf.value.decode())
