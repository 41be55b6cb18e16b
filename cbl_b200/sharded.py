"""Prefix-range sharding of one CBL set across the GPUs of a box: one process per GPU
(``torch.distributed``), one all-to-all per batch, everything else shard-local.

The reference has no multi-device notion; this is the scale-out design of
BASELINE.json's north star:

* the 2^PREFIX_BITS prefix space is cut into ``world`` contiguous ranges with EQUAL-MASS splitters
  (necklace prefixes are extremely skewed — SURVEY F4 — so equal-width ranges would put ~99 % of the
  words on rank 0);
* every rank turns ITS reads into words locally (fused encode + necklace kernel), routes each word
  to the rank that owns its prefix with a single all-to-all-v, and the owner applies the batch to
  its shard (sort / merge / probe) with no further communication;
* ``contains_seq`` answers come back with a second, one-byte-per-word all-to-all and are put back
  in the order of the local reads;
* shards hold disjoint ascending prefix ranges, so the set in global ascending-word order is the
  concatenation of the shards in rank order; ``count`` is a sum.

The device work is done by an *engine* (``cbl_b200.CBL`` on a GPU).  The routing logic only needs
``seq_words`` / ``words_op`` from it, which is what lets the world_size-2 ``gloo`` tests run this file
on CPU tensors with a stand-in engine.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def pos_bits(k: int) -> int:
    p = 0
    while (1 << p) < 2 * k:
        p += 1
    return p


def word_prefixes(words: torch.Tensor, suffix_bits: int, prefix_bits: int) -> torch.Tensor:
    """prefix = word >> SUFFIX_BITS (src/wordset/mod.rs:63-71) for words held as int64 tensors:
    shape (n,) for 64-bit words, (n, 2) = (lo, hi) for 128-bit words.  Returns int64."""
    mask = (1 << prefix_bits) - 1
    if words.dim() == 1:
        if suffix_bits == 0:
            return words & mask
        return (words >> suffix_bits) & ((1 << (64 - suffix_bits)) - 1) & mask
    lo, hi = words[:, 0], words[:, 1]
    if suffix_bits >= 64:
        s = suffix_bits - 64
        return ((hi >> s) & ((1 << (64 - s)) - 1) & mask) if s else (hi & mask)
    lo_part = (lo >> suffix_bits) & ((1 << (64 - suffix_bits)) - 1)
    return ((hi << (64 - suffix_bits)) | lo_part) & mask


def route(prefixes: torch.Tensor, splitters: torch.Tensor):
    """dest rank of every word, the permutation that groups words by dest (stable), per-dest counts."""
    world = splitters.numel() + 1
    dest = torch.bucketize(prefixes, splitters, right=True)
    order = torch.argsort(dest, stable=True)
    counts = torch.bincount(dest, minlength=world)
    return dest, order, counts


def exchange(send: torch.Tensor, send_counts: torch.Tensor, group=None):
    """all-to-all-v of rows of ``send`` (already grouped by destination).  Returns (recv, recv_counts)."""
    world = dist.get_world_size(group)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    recv = send.new_empty((int(sum(rc)),) + tuple(send.shape[1:]))
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return recv, recv_counts


# Relative probe cost of a word as a function of the mass quantile q of its prefix in the sorted sample, as (q, weight)
# knots of a piecewise-linear curve: flat except for the sparse tail of the prefix space (few suffixes per bucket, bucket
# tables spread over most of the prefix range).  Measured on B200, K=25: (a) word-level probe against the shard one rank
# of an 8-GPU run holds (scripts/exp_shard_probe.py: 500 M k-mers of 4 G in one eighth of the prefix mass, arrival order):
# 19.6, 19.8, 20.2, 20.9 and 23.7 ms per 1 G words in octiles 0, 2, 4, 6 and 7; (b) the fused query on 8 x B200 with
# equal-mass splitters: ranks 0-4 finish their share together, ranks 5 and 6 1-2 ms later, rank 7 7 ms later; with a
# tail weight of 1.25 rank 7 is 3.6 ms early: the knots below put it in between.  Before the correction bytes were scaled
# by bucket size (index_view.cuh sub_scale) measurement (a) rose from 20.3 to 30.3 ms: a plain signed byte saturates in
# the large, structured buckets of an 8-GPU set.
PROBE_COST_KNOTS = ((0.0, 1.0), (0.625, 1.0), (0.875, 1.05), (0.9375, 1.12), (1.0, 1.2))


def equal_mass_splitters(prefixes: torch.Tensor, world: int, tail_cost: Optional[float] = None, knee: float = 0.625,
                         knots: Optional[Sequence] = None) -> torch.Tensor:
    """world-1 splitters s.t. each range [s_{i-1}, s_i) holds ~1/world of the sample's COST: the splitters equalise the
    owner-side probe time, not the word count.

    ``knots``: (mass quantile, relative cost) points of a piecewise-linear cost curve over the sorted sample (default
    ``PROBE_COST_KNOTS``; ``((0, 1), (1, 1))`` = plain equal mass).  ``tail_cost`` (legacy two-piece model, what
    CBL_SPLIT_TAIL_COST sets): cost rising linearly from 1 at q = 0 to ``tail_cost`` at ``knee``, constant beyond."""
    if world == 1:
        return prefixes.new_empty(0)
    if tail_cost is not None:
        knots = ((0.0, 1.0), (knee, tail_cost), (1.0, tail_cost))
    elif knots is None:
        knots = PROBE_COST_KNOTS
    srt, _ = torch.sort(prefixes)
    n = srt.numel()
    grid = np.linspace(0.0, 1.0, 4097)
    w = np.interp(grid, [k[0] for k in knots], [k[1] for k in knots])
    cum = np.concatenate([[0.0], np.cumsum((w[1:] + w[:-1]) * 0.5 * np.diff(grid))])   # cumulative cost C(q)
    idx = []
    for i in range(1, world):
        q = float(np.interp(cum[-1] * i / world, cum, grid))
        idx.append(min(int(q * n), n - 1))
    sp = srt[torch.tensor(idx, device=srt.device)]
    # strictly increasing splitters keep every range non-empty in prefix space
    for i in range(1, sp.numel()):
        if sp[i] <= sp[i - 1]:
            sp[i] = sp[i - 1] + 1
    return sp


class GpuEngine:
    """Adapter: ``cbl_b200.CBL`` on the local GPU behind the interface the router needs."""

    def __init__(self, k, t_bits, prefix_bits, canonical, device, cbl=None):
        from .cbl import CBL

        self.cbl = cbl if cbl is not None else CBL(k, t_bits, prefix_bits, canonical, device)
        self.params = (k, t_bits, prefix_bits, canonical, device)
        self.device = torch.device("cuda", device)
        self.word_bytes = self.cbl.word_bytes()

    def _wrap(self, cbl) -> "GpuEngine":
        return GpuEngine(*self.params, cbl=cbl)

    # shard-local set algebra, clone, export, serde (what ShardedCBL composes rank by rank)
    def setop(self, op: int, other: "GpuEngine") -> "GpuEngine":
        return self._wrap(self.cbl._binary(op, other.cbl))

    def setop_assign(self, op: int, other: "GpuEngine") -> None:
        self.cbl._assign(op, other.cbl)

    def clone(self) -> "GpuEngine":
        return self._wrap(self.cbl.clone())

    def words_list(self) -> List[int]:
        return self.cbl.words()

    def kmers_list(self) -> List[int]:
        return list(self.cbl.iter())

    def serialize(self) -> bytes:
        return self.cbl.serialize()

    def deserialize_range(self, data: bytes, lo: int, hi: int) -> "GpuEngine":
        return self._wrap(self.cbl.deserialize_range(data, lo, hi))

    def seq_words(self, d_buf: int, offsets: np.ndarray) -> torch.Tensor:
        n = self.cbl.count_kmers(offsets)
        shape = (n,) if self.word_bytes == 8 else (n, 2)
        words = torch.empty(shape, dtype=torch.int64, device=self.device)
        if n:
            # the library launches on its own stream: everything torch queued must be done first
            torch.cuda.current_stream(self.device).synchronize()
            self.cbl.seq_words_dev(d_buf, offsets, words.data_ptr())  # returns after its stream is idle
        return words

    def words_op(self, op: int, words: torch.Tensor, want_flags: bool) -> Optional[torch.Tensor]:
        n = words.shape[0]
        flags = torch.empty(n, dtype=torch.uint8, device=self.device) if want_flags else None
        if n:
            words = words.contiguous()
            torch.cuda.current_stream(self.device).synchronize()
            self.cbl.words_op_dev(op, words.data_ptr(), n, flags.data_ptr() if want_flags else 0)
        return flags

    def count(self) -> int:
        return self.cbl.count()

    # hand-written router kernels (stable partition by owner rank + answer gather) instead of the
    # generic torch bucketize / argsort / index_select path
    def route(self, words: torch.Tensor, splitters: np.ndarray):
        n = words.shape[0]
        send = torch.empty_like(words)
        pos = torch.empty(n, dtype=torch.int32, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        counts = self.cbl.route_words_dev(words.data_ptr(), n, splitters, send.data_ptr(), pos.data_ptr())
        return send, pos, counts

    def gather(self, src: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        out = torch.empty(pos.shape[0], dtype=torch.uint8, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self.cbl.gather_u8_dev(src.data_ptr(), pos.data_ptr(), pos.shape[0], out.data_ptr())
        return out

    def sample_words(self, n_bases: int, seed: int) -> torch.Tensor:
        g = torch.Generator(device=self.device)
        g.manual_seed(seed)
        lut = torch.tensor(list(b"ACTG"), dtype=torch.uint8, device=self.device)
        seq = lut[torch.randint(0, 4, (n_bases,), generator=g, device=self.device, dtype=torch.int32)]
        return self.seq_words(seq.data_ptr(), np.array([0, n_bases], dtype=np.uint64))


class PeerExchange:
    """Receive / answer buffers of one rank, mapped into every other rank of the box with CUDA IPC, so that
    the fused encode + necklace + route kernel stores each word directly into its owner's memory over NVLink
    and the owner's probe kernel stores each answer directly back (no send buffer, no counting pass, no NCCL
    data path).  The process group only carries per-batch counts, the IPC handles and barriers, so it may
    be NCCL or gloo.

    Layout: the receive buffer of owner d is ``world`` regions of ``cap`` words, region s written by rank s
    only (slots reserved chunk by chunk with device-local atomics); the answer buffer of rank s is ``world``
    regions of ``cap`` bytes, region d written by owner d only, answer j of region d belonging to word j of
    region s at owner d."""


    def __init__(self, cbl, group, rank: int, world: int, device):
        self.cbl, self.group, self.rank, self.world, self.device = cbl, group, rank, world, device
        # buffer sets: sub-batch b of a pipelined query (CBL_PIPE > 1) uses set b & 1 (route of b + 1 overlaps the probe of b)
        self.SLOTS = 2 if int(os.environ.get("CBL_PIPE", "1")) > 1 else 1
        self.word_bytes = cbl.word_bytes()
        backend = dist.get_backend(group)
        self.ctrl = torch.device("cpu") if backend == "gloo" else device
        self.cap = 0        # words per region
        self.users = 0      # sets sharing these buffers
        self.own_recv = self.own_back = self.own_ctrl = 0
        self.peer_recv: List[int] = []
        self.peer_back: List[int] = []
        self.peer_ctrl: List[int] = []   # control blocks of the fused query: one final-count slot per source rank
        self.recv_dirty = True           # the receive buffer may hold something else than 0xFF (see clean_recv)
        self.epoch = 0                   # call counter of the fused query (same on every rank: the calls are collective)

    def barrier(self):
        if self.ctrl.type == "cpu":
            dist.barrier(group=self.group)
        else:
            dist.barrier(group=self.group, device_ids=[self.device.index])

    def all_counts(self, counts: np.ndarray) -> np.ndarray:
        """g x len(counts) matrix of every rank's vector; doubles as a barrier."""
        mine = torch.from_numpy(np.ascontiguousarray(counts).astype(np.int64)).to(self.ctrl)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return torch.stack(out).cpu().numpy().astype(np.uint64)

    def _release(self):
        for r, p in enumerate(self.peer_recv):
            if r != self.rank and p:
                self.cbl.peer_close(p)
        for r, p in enumerate(self.peer_back):
            if r != self.rank and p:
                self.cbl.peer_close(p)
        for r, p in enumerate(self.peer_ctrl):
            if r != self.rank and p:
                self.cbl.peer_close(p)
        self.peer_recv, self.peer_back, self.peer_ctrl = [], [], []
        self.barrier()  # nobody maps our blocks any more
        for p in (self.own_recv, self.own_back, self.own_ctrl):
            if p:
                self.cbl.peer_free(p)
        self.own_recv = self.own_back = self.own_ctrl = 0

    def ensure(self, cap_words: int):
        """Collective: every rank calls it with the same argument."""
        if cap_words <= self.cap and self.peer_recv:
            return
        self._release()
        self.cap = max((int(cap_words) + 2047) // 2048 * 2048, self.cap)
        if self.world * self.cap >= 1 << 32:
            raise ValueError("batch too large for one exchange: split the reads into smaller batches")
        self.own_recv, h_recv = self.cbl.peer_alloc(self.SLOTS * self.world * self.cap * self.word_bytes)
        self.own_back, h_back = self.cbl.peer_alloc(self.SLOTS * self.world * self.cap)
        self.own_ctrl, h_ctrl = self.cbl.peer_alloc(self.ctrl_bytes())
        handles = [None] * self.world
        dist.all_gather_object(handles, (h_recv, h_back, h_ctrl), group=self.group)
        self.peer_recv = [self.own_recv if r == self.rank else self.cbl.peer_open(handles[r][0]) for r in range(self.world)]
        self.peer_back = [self.own_back if r == self.rank else self.cbl.peer_open(handles[r][1]) for r in range(self.world)]
        self.peer_ctrl = [self.own_ctrl if r == self.rank else self.cbl.peer_open(handles[r][2]) for r in range(self.world)]
        self.cbl.peer_zero(self.own_ctrl, self.ctrl_bytes())
        self.recv_dirty = True
        self.barrier()

    # control block of the fused query (one per rank, peer-mapped): one u64 final-count slot per source rank, tagged with
    # the call's epoch by the source (csrc/shard_query.cuh), so it is zeroed once and never again
    def ctrl_bytes(self) -> int:
        return self.world * 8 + 64

    def final_slot(self, base: int, src: int) -> int:
        return base + src * 8

    def clean_recv(self):
        """Collective.  The fused query wants 0xFF in every slot of the receive buffers that holds no word (a word is its own
        arrival flag); it leaves them that way itself, so this runs after allocation, after the buffers served another path
        (mutations, the two-kernel query) and after an overflowed or failed call."""
        if self.recv_dirty:
            self.cbl.peer_fill(self.own_recv, 0xFF, self.SLOTS * self.world * self.cap * self.word_bytes)
            self.recv_dirty = False
            self.barrier()   # nobody stores into a buffer that is still being filled

    def next_epoch(self) -> int:
        self.epoch = self.epoch % 65535 + 1
        return self.epoch

    def my_regions(self, slot: int = 0) -> List[int]:
        """my region inside every owner's receive buffer (buffer set ``slot``)"""
        return [p + (slot * self.world + self.rank) * self.cap * self.word_bytes for p in self.peer_recv]

    def recv_region(self, src: int, slot: int = 0) -> int:
        """region of my receive buffer that rank ``src`` writes"""
        return self.own_recv + (slot * self.world + src) * self.cap * self.word_bytes

    def answer_region(self, src: int, slot: int = 0) -> int:
        """my region inside rank ``src``'s answer buffer"""
        return self.peer_back[src] + (slot * self.world + self.rank) * self.cap

    def answers(self, slot: int = 0) -> int:
        """my own answer buffer (what pos[] of seq_route_dev indexes)"""
        return self.own_back + slot * self.world * self.cap

    def close(self):
        if self.peer_recv or self.own_recv:
            self._release()
        self.cap = 0


_PEER_CACHE = {}


class ShardedCBL:
    """One CBL set sharded by prefix range over the ranks of ``group`` (default: the world)."""

    def __init__(self, k: int, t_bits: int, prefix_bits: int = 24, canonical: bool = False, device: int = 0,
                 engine=None, splitters: Optional[Sequence[int]] = None, group=None, sample_bases: int = 4_000_000):
        self.k, self.t_bits, self.prefix_bits, self.canonical = k, t_bits, prefix_bits, canonical
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.suffix_bits = 2 * k + pos_bits(k) - prefix_bits
        self.engine = engine if engine is not None else GpuEngine(k, t_bits, prefix_bits, canonical, device)
        self.device = self.engine.device
        if self.world > 1 and isinstance(self.engine, GpuEngine):
            # a shard's batches cover 1 / world of the prefix mass: the hybrid batch sort plans for heads that much more frequent
            # (without the hint the first batch of an 8-GPU shard overflows the segment tiles and is re-sorted by plain LSD passes)
            self.engine.cbl.set_sort_concentration(self.world)
        if splitters is not None:
            sp = torch.tensor(list(splitters), dtype=torch.int64, device=self.device)
        else:
            # equal-mass splitters for uniform random DNA from a fixed-seed sample pushed through the
            # real necklace kernel; rank 0 decides, everybody adopts (identical on every rank)
            sp = torch.zeros(max(self.world - 1, 0), dtype=torch.int64, device=self.device)
            if self.world > 1:
                if self.rank == 0:
                    w = self.engine.sample_words(sample_bases, seed=20240229)
                    tc = os.environ.get("CBL_SPLIT_TAIL_COST")   # developer knob: the legacy two-piece cost model
                    sp.copy_(equal_mass_splitters(word_prefixes(w, self.suffix_bits, prefix_bits), self.world,
                                                   tail_cost=float(tc) if tc else None))
                ctrl = sp.cpu() if dist.get_backend(group) == "gloo" else sp   # gloo: control traffic on CPU tensors
                dist.broadcast(ctrl, src=0, group=group)
                sp = ctrl.to(self.device)
        assert sp.numel() == self.world - 1
        self.splitters = sp
        self.splitters_u32 = sp.cpu().numpy().astype(np.uint32)
        # data path of the exchange: "peer" = fused route + NVLink stores into the owner's buffers (default on
        # GPUs), "nccl" = partition into a send buffer + all_to_all_single (also what CPU stand-ins use)
        mode = os.environ.get("CBL_EXCHANGE", "peer")
        self.peer = None
        if self.world > 1 and mode == "peer" and isinstance(self.engine, GpuEngine):
            # the exchange buffers (GBs, mapped into every peer: a collective allocation) are shared by all sets of this
            # process that live on the same group / device / word width: calls are collective and host-synchronous, so two
            # sets never use the buffers at the same time, and a fresh set (a set-op result, a clone, a scratch index) costs
            # no allocation, IPC handle exchange or barrier
            key = (id(group) if group is not None else 0, self.device.index, self.engine.word_bytes)
            px = _PEER_CACHE.get(key)
            if px is None:
                from .cbl import CBL

                # the exchange owns a small empty handle of its own (device context for the allocation / mapping calls), so it
                # does not pin the memory of whichever set happened to come first
                px = _PEER_CACHE[key] = PeerExchange(CBL(k, t_bits, prefix_bits, canonical, self.device.index), group, self.rank, self.world, self.device)
            px.users += 1
            self.peer = px

    # -- routing ---------------------------------------------------------------------------------
    def _route_words(self, words: torch.Tensor):
        """-> (words received by this rank, pos: slot of every local word in the send buffer,
        send counts, recv counts)."""
        if hasattr(self.engine, "route"):
            send, pos, counts_np = self.engine.route(words, self.splitters.cpu().numpy().astype(np.uint32))
            counts = torch.from_numpy(counts_np.astype(np.int64)).to(self.device)
        else:  # generic torch path (CPU tests / stand-in engines)
            pre = word_prefixes(words, self.suffix_bits, self.prefix_bits)
            _, order, counts = route(pre, self.splitters)
            send = words.index_select(0, order)
            pos = torch.empty_like(order)
            pos[order] = torch.arange(order.numel(), device=order.device)
        if self.world == 1:
            return send, pos, counts, counts
        recv, recv_counts = exchange(send, counts, self.group)
        return recv, pos, counts, recv_counts

    def _peer_route_seqs(self, d_buf: int, offsets: np.ndarray, want_pos: bool):
        """Fused encode + necklace + route + exchange of this rank's reads.  -> (C, pos): C[s][d] = words rank s
        stored into its region at owner d; pos (int32 device tensor or None) = answer slot of every local k-mer."""
        px, cbl = self.peer, self.engine.cbl
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = cbl.count_kmers(offsets)
        # doubles as the barrier "every rank is done with the buffers of the previous batch"
        n_max = int(px.all_counts(np.array([n], dtype=np.uint64)).max())
        cap = int(n_max / self.world * float(os.environ.get("CBL_ROUTE_SLACK", self.SLACK))) + 4096
        pos = torch.empty(n, dtype=torch.int32, device=self.device) if want_pos else None
        torch.cuda.current_stream(self.device).synchronize()
        while True:
            px.ensure(cap)
            px.recv_dirty = True
            counts = cbl.seq_route_dev(d_buf, offsets, self.splitters_u32, px.my_regions(), px.cap, pos.data_ptr() if want_pos else 0)
            C = px.all_counts(counts)                                # barrier: all words have landed
            if int(C.max()) <= px.cap:
                return C, pos
            cap = int(int(C.max()) * 1.1) + 4096                     # a region overflowed somewhere: everybody retries

    SLACK = 1.3    # region capacity over the even share (cost-weighted splitters give the first ranks a few % more words)

    def _mutate(self, op: int, d_buf: int, offsets: np.ndarray) -> None:
        if self.peer is not None:
            import time

            trace = os.environ.get("CBL_SHARD_TRACE") and self.rank == 0
            t0 = time.perf_counter()
            px, cbl = self.peer, self.engine.cbl
            C, _ = self._peer_route_seqs(d_buf, offsets, want_pos=False)
            t1 = time.perf_counter()
            col = C[:, self.rank]
            segs = [px.own_recv + s * px.cap * px.word_bytes for s in range(self.world)]
            if int(col.sum()):
                cbl.words_op_segments_dev(op, segs, col)
            if trace:
                print(f"[shard trace] mutate: route + exchanges {(t1 - t0) * 1e3:.2f} ms, owner-side sort + merge of {int(col.sum())} words "
                      f"{(time.perf_counter() - t1) * 1e3:.2f} ms", flush=True)
            return
        words = self.engine.seq_words(d_buf, np.ascontiguousarray(offsets, dtype=np.uint64))
        recv, _, _, _ = self._route_words(words)
        self.engine.words_op(op, recv, want_flags=False)

    def insert_seqs_dev(self, d_buf: int, offsets: np.ndarray) -> None:
        """Every rank passes ITS OWN reads; all k-mers of all ranks end up in the (global) set."""
        self._mutate(1, d_buf, offsets)

    def remove_seqs_dev(self, d_buf: int, offsets: np.ndarray) -> None:
        self._mutate(2, d_buf, offsets)

    def contains_words(self, words: torch.Tensor) -> torch.Tensor:
        recv, pos, counts, recv_counts = self._route_words(words)
        flags = self.engine.words_op(0, recv, want_flags=True)
        if self.world > 1:
            back = flags.new_empty(int(counts.sum()))
            dist.all_to_all_single(back, flags, output_split_sizes=counts.tolist(), input_split_sizes=recv_counts.tolist(), group=self.group)
            flags = back
        # answers in the order of the local reads: out[i] = flags[pos[i]]
        if hasattr(self.engine, "gather"):
            return self.engine.gather(flags, pos)
        return flags[pos.long()]

    def contains_seqs_dev(self, d_buf: int, offsets: np.ndarray) -> torch.Tensor:
        """Per-k-mer answers (uint8 device tensor) for this rank's reads, in the reference's order
        (src/cbl.rs:311-324)."""
        if self.peer is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
            if int(os.environ.get("CBL_FUSED", self.FUSED)):
                return self._peer_contains_fused(d_buf, offsets)
            return self._peer_contains(d_buf, offsets)
        words = self.engine.seq_words(d_buf, np.ascontiguousarray(offsets, dtype=np.uint64))
        return self.contains_words(words)

    # sub-batches of one contains_seqs call (same on every rank: the collectives must pair up).  Measured on 2 x B200
    # (bench.py, 1 Gbp per rank): 4 sub-batches 53.6 ms per step, 1 sub-batch 50.7 ms — the route and probe kernels contend
    # for the same SM resources when co-resident (each slows down by what the other takes), so the default is 1.
    PIPE = 1

    # The default query path is the fused kernel (csrc/shard_query.cuh): measured on 8 x B200 / 2 x B200 (bench.py, 1 Gbp per
    # rank) against route kernel -> count exchange -> one probe launch (CBL_FUSED=0), see DESIGN.md section 7.
    FUSED = 1

    def _peer_contains_fused(self, d_buf: int, offsets: np.ndarray) -> torch.Tensor:
        """contains_seq of this rank's reads as ONE kernel per rank (cbl_seq_contains_fused_dev, csrc/shard_query.cuh): every
        warp alternates between producing (encode + necklace + route of this rank's reads) and probing the blocks of words
        the peers have completed in this rank's buffer; words travel to their owners and answers back over NVLink peer
        memory while the kernel runs.  The process group carries ONE count exchange per call (region overflow check and
        "every answer has landed")."""
        px, cbl = self.peer, self.engine.cbl
        n = cbl.count_kmers(offsets)
        # Region sizing needs an exchange on the first call only (see _peer_contains); the previous call ended with "all
        # answers have landed", so nobody is still using the buffers
        if px.cap == 0:
            n_max = int(px.all_counts(np.array([n], dtype=np.uint64)).max())
            cap = int(n_max / self.world * float(os.environ.get("CBL_ROUTE_SLACK", self.SLACK))) + 4096
        else:
            cap = px.cap
        pos = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        out = torch.empty(n, dtype=torch.uint8, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        g = self.world
        while True:
            px.ensure(cap)
            px.clean_recv()
            epoch = px.next_epoch()
            px.recv_dirty = True           # until the call has come back clean
            # a rank whose kernel reports an error (reads with non-ACGT bytes: the fused kernel rejects them) has still run its
            # share of the protocol to the end; it tells the others through the count exchange, so that every rank raises
            # instead of the others waiting in a collective for ever
            failure = None
            try:
                counts = cbl.seq_contains_fused_dev(
                    d_buf, offsets, self.splitters_u32,
                    peer_region=px.my_regions(),
                    peer_final=[px.final_slot(px.peer_ctrl[d], self.rank) for d in range(g)],
                    cap=px.cap, d_pos=pos.data_ptr(),
                    recv_region=[px.recv_region(s_) for s_ in range(g)],
                    answer_region=[px.answer_region(s_) for s_ in range(g)],
                    final_counts=[px.final_slot(px.own_ctrl, s_) for s_ in range(g)],
                    epoch=epoch)
            except Exception as e:   # noqa: BLE001 (re-raised below, on every rank)
                failure, counts = e, np.zeros(g, dtype=np.uint64)
            full = px.all_counts(np.append(counts, np.uint64(1 if failure is not None else 0)))   # barrier: every consumer is done => my answers have landed
            C = full[:, :g]
            if full[:, g].any():
                if failure is not None:
                    raise failure
                bad = [int(r) for r in np.nonzero(full[:, g])[0]]
                raise RuntimeError(f"sharded contains_seq failed on rank(s) {bad} (see their error); nothing was answered")
            if os.environ.get("CBL_SHARD_TRACE"):
                print(f"[shard trace] rank {self.rank}: sent {int(C[self.rank].sum())} words, received {int(C[:, self.rank].sum())}", flush=True)
            if int(C.max()) <= px.cap:
                px.recv_dirty = False                                # every word that arrived was consumed and its slot reset
                self.last_route_counts = C                           # C[s][d] = words rank s sent to owner d (bench: NVLink traffic)
                break
            cap = int(int(C.max()) * 1.1) + 4096                     # a region overflowed somewhere: everybody refills and retries
        if n:
            cbl.gather_u8_dev(px.answers(), pos.data_ptr(), n, out.data_ptr())
        return out

    def _peer_contains(self, d_buf: int, offsets: np.ndarray) -> torch.Tensor:
        """Pipelined query over peer memory.  The reads are cut into PIPE sub-batches of whole records.  Sub-batch b:
        (1) fused encode + necklace + route on the ROUTER handle's stream (integer-pipe bound), words stored into the
        owners' buffer set b & 1; (2) count matrix exchange = "all words of b have landed"; (3) the owner-side probe of b
        (memory bound) runs on the index handle's stream in a worker thread WHILE the main thread routes b + 1 — the two
        kernels share the SMs the way the two halves of the single-GPU fused kernel do; (4) the answers of b are
        gathered into read order after the next exchange, which every rank enters only after its probe of b is
        complete.  A region overflow anywhere makes every rank start over with larger regions."""
        import threading

        import time

        trace = os.environ.get("CBL_SHARD_TRACE") and self.rank == 0
        t_last = [time.perf_counter()]

        def mark(what):
            if trace:
                now = time.perf_counter()
                print(f"[shard trace] {what}: {(now - t_last[0]) * 1e3:.2f} ms", flush=True)
                t_last[0] = now

        px, cbl, router = self.peer, self.engine.cbl, self._router()
        n_rec = len(offsets) - 1
        kpr = np.array([max(int(offsets[i + 1] - offsets[i]) - self.k + 1, 0) for i in range(n_rec)], dtype=np.int64)
        total = int(kpr.sum())
        pipe = max(1, int(os.environ.get("CBL_PIPE", self.PIPE))) if px.SLOTS > 1 else 1
        # record ranges with about the same number of k-mers (ranges may be empty on a rank with few records)
        csum = np.concatenate([[0], np.cumsum(kpr)])
        cuts = [int(np.searchsorted(csum, total * b / pipe, side="left")) for b in range(pipe)] + [n_rec]
        cuts = [min(max(c, 0), n_rec) for c in cuts]
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1])
        sub_n = [int(csum[cuts[b + 1]] - csum[cuts[b]]) for b in range(pipe)]
        # Region sizing needs one exchange on the first call only: later calls start from the capacity they find (a region
        # that turns out too small is detected by the count exchange after the route and everybody retries).  No barrier
        # is needed here: the previous call ended with "all answers have landed" and every rank gathered its own answers
        # before it could enter this call's first collective.
        if px.cap == 0:
            n_max = int(px.all_counts(np.array([max(sub_n)], dtype=np.uint64)).max())
            cap = int(n_max / self.world * float(os.environ.get("CBL_ROUTE_SLACK", self.SLACK))) + 4096
        else:
            cap = px.cap
        mark("plan + first count exchange")
        out = torch.empty(total, dtype=torch.uint8, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        mark("alloc out")

        def probe(C, slot):
            t_p = time.perf_counter()
            # region s of my receive buffer -> my region of rank s's answer buffer: ONE launch over the g regions
            cbl.words_contains_segments_dev([px.recv_region(s_, slot) for s_ in range(self.world)], C[:, self.rank],
                                            [px.answer_region(s_, slot) for s_ in range(self.world)])
            if os.environ.get("CBL_SHARD_TRACE"):
                print(f"[shard trace] rank {self.rank}: probe of {int(C[:, self.rank].sum())} words {(time.perf_counter() - t_p) * 1e3:.2f} ms, "
                      f"{cbl.num_buckets()} buckets, {cbl.count()} k-mers", flush=True)

        while True:
            px.ensure(cap)
            px.recv_dirty = True
            worker, prev, overflow = None, None, 0
            for b in range(pipe):
                slot = b & 1
                r0, r1 = cuts[b], cuts[b + 1]
                pos = torch.empty(sub_n[b], dtype=torch.int32, device=self.device)
                if sub_n[b]:
                    torch.cuda.current_stream(self.device).synchronize()   # pos is allocated before the router's stream writes it
                    counts = router.seq_route_dev(d_buf, offsets[r0:r1 + 1], self.splitters_u32, px.my_regions(slot), px.cap, pos.data_ptr())
                else:
                    counts = np.zeros(self.world, dtype=np.uint64)
                mark(f"route {b}")
                if worker is not None:
                    worker.join()
                mark(f"join probe {b - 1}")
                C = px.all_counts(counts)        # words of b have landed everywhere; so have the answers of b - 1
                mark(f"count exchange {b}")
                if prev is not None:
                    self._gather(router, px, prev, out)
                    prev = None
                if int(C.max()) > px.cap:
                    overflow = int(C.max())
                    break
                worker = threading.Thread(target=probe, args=(C, slot))
                worker.start()
                prev = (slot, pos, int(csum[r0]), sub_n[b])
            if worker is not None:
                worker.join()
            mark("join last probe")
            px.barrier()                         # the answers of the last sub-batch have landed
            mark("barrier")
            if not overflow:
                if prev is not None:
                    self._gather(router, px, prev, out)
                mark("gather")
                return out
            cap = int(overflow * 1.1) + 4096     # a region overflowed somewhere: everybody retries

    @staticmethod
    def _gather(router, px, prev, out):
        slot, pos, start, n = prev
        if n:
            router.gather_u8_dev(px.answers(slot), pos.data_ptr(), n, out.data_ptr() + start)

    def _router(self):
        """second handle (empty set, own CUDA stream) that runs the fused route kernel and the answer gather"""
        if getattr(self, "_router_cbl", None) is None:
            from .cbl import CBL

            os.environ["CBL_STREAM_HIGH_PRIORITY"] = "1"   # read by the handle's constructor only
            try:
                self._router_cbl = CBL(self.k, self.t_bits, self.prefix_bits, self.canonical, self.device.index)
            finally:
                del os.environ["CBL_STREAM_HIGH_PRIORITY"]
        return self._router_cbl

    # host-buffer front ends (what a user of the reference calls): copy, then the device path
    def insert_seqs(self, buf: np.ndarray, offsets: np.ndarray) -> None:
        t = torch.from_numpy(np.ascontiguousarray(buf)).to(self.device, non_blocking=True)
        self.insert_seqs_dev(t.data_ptr(), offsets)

    def remove_seqs(self, buf: np.ndarray, offsets: np.ndarray) -> None:
        t = torch.from_numpy(np.ascontiguousarray(buf)).to(self.device, non_blocking=True)
        self.remove_seqs_dev(t.data_ptr(), offsets)

    E2E_PIPE = 4   # sub-batches of a host-buffer query (CBL_E2E_PIPE); must be the same on every rank

    def contains_seqs(self, buf: np.ndarray, offsets: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host buffers in, answers on the host (what a user of the reference calls), software-pipelined: the reads are cut
        into sub-batches of whole records; while sub-batch b is routed / probed, the reads of b + 1 .. are on their way in
        and the answers of b - 1 on their way out (H2D and D2H on their own streams and copy engines; true DMA when the
        caller's buffers are pinned).  Every rank must use the same number of sub-batches (the collectives pair up)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n_rec = len(offsets) - 1
        kpr = np.array([max(int(offsets[i + 1] - offsets[i]) - self.k + 1, 0) for i in range(n_rec)], dtype=np.int64)
        total = int(kpr.sum())
        if out is None:
            out = np.empty(max(total, 1), dtype=np.uint8)
        if self.device.type != "cuda":      # stand-in engines (CPU tests): plain path
            t = torch.from_numpy(np.ascontiguousarray(buf))
            ans = self.contains_seqs_dev(t.data_ptr() if hasattr(t, "data_ptr") else t, offsets)
            torch.from_numpy(out)[: ans.numel()].copy_(ans)
            return out[: ans.numel()]
        pipe = max(1, int(os.environ.get("CBL_E2E_PIPE", self.E2E_PIPE)))
        csum = np.concatenate([[0], np.cumsum(kpr)])
        cuts = [int(np.searchsorted(csum, total * b / pipe, side="left")) for b in range(pipe)] + [n_rec]
        cuts = [min(max(c, 0), n_rec) for c in cuts]
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1])
        h_in, h_out = torch.from_numpy(np.ascontiguousarray(buf)), torch.from_numpy(out)
        if getattr(self, "_copy_streams", None) is None:
            self._copy_streams = (torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
        s_in, s_out = self._copy_streams
        cur = torch.cuda.current_stream(self.device)
        d_in, ev_in = [], []
        for b in range(pipe):
            b0, b1 = int(offsets[cuts[b]]), int(offsets[cuts[b + 1]])
            d_in.append(torch.empty(max(b1 - b0, 1), dtype=torch.uint8, device=self.device))
        s_in.wait_stream(cur)
        with torch.cuda.stream(s_in):
            for b in range(pipe):
                b0, b1 = int(offsets[cuts[b]]), int(offsets[cuts[b + 1]])
                if b1 > b0:
                    d_in[b][: b1 - b0].copy_(h_in[b0:b1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s_in)
                ev_in.append(ev)
        keep = []
        for b in range(pipe):
            ev_in[b].synchronize()
            sub_off = offsets[cuts[b] : cuts[b + 1] + 1] - offsets[cuts[b]]
            ans = self.contains_seqs_dev(d_in[b].data_ptr(), sub_off)        # host-synchronous: the answers are complete
            k0 = int(csum[cuts[b]])
            if ans.numel():
                with torch.cuda.stream(s_out):
                    h_out[k0 : k0 + ans.numel()].copy_(ans, non_blocking=True)
            keep.append(ans)
        s_out.synchronize()
        return out[:total]

    # -- set algebra, clone, iteration, serde: shard-local, composed in rank order (src/cbl.rs:411-569, 358-360, 127-160) ----
    def _derive(self, engine) -> "ShardedCBL":
        """a set with the same parameters, process group and splitters around another shard-local engine (collective
        only in so far as every rank does it; no communication)"""
        return ShardedCBL(self.k, self.t_bits, self.prefix_bits, self.canonical, device=self.device.index or 0, engine=engine,
                          splitters=[int(x) for x in self.splitters_u32], group=self.group)

    def _check(self, other: "ShardedCBL") -> None:
        if (other.k, other.t_bits, other.prefix_bits) != (self.k, self.t_bits, self.prefix_bits):
            raise ValueError("set operation between indexes with different K / T / PREFIX_BITS")
        if other.canonical != self.canonical:
            raise ValueError("One of the index is canonical while the other isn't")   # src/cbl.rs:422-425
        if other.world != self.world or list(other.splitters_u32) != list(self.splitters_u32):
            raise ValueError("set operation between indexes sharded differently (ranks / splitters)")

    def _binary(self, op: int, other: "ShardedCBL") -> "ShardedCBL":
        self._check(other)
        return self._derive(self.engine.setop(op, other.engine))   # prefix ranges coincide: the op is shard-local

    def _assign(self, op: int, other: "ShardedCBL") -> "ShardedCBL":
        self._check(other)
        self.engine.setop_assign(op, other.engine)
        return self

    def __or__(self, o): return self._binary(0, o)
    def __and__(self, o): return self._binary(1, o)
    def __sub__(self, o): return self._binary(2, o)
    def __xor__(self, o): return self._binary(3, o)
    def __ior__(self, o): return self._assign(0, o)
    def __iand__(self, o): return self._assign(1, o)
    def __isub__(self, o): return self._assign(2, o)
    def __ixor__(self, o): return self._assign(3, o)

    def clone(self) -> "ShardedCBL":
        return self._derive(self.engine.clone())

    def local_words(self) -> List[int]:
        """this rank's words, ascending: the slice [splitter[rank-1], splitter[rank]) of the global ascending set"""
        return self.engine.words_list()

    def _gather_lists(self, local):
        if self.world == 1:
            return [local]
        parts = [None] * self.world
        dist.all_gather_object(parts, local, group=self.group)
        return parts

    def words(self) -> List[int]:
        """the whole set in ascending word order on every rank = the shards concatenated in rank order (test scale: the
        lists travel as Python objects; at scale use local_words() / save_to_file())"""
        return [w for part in self._gather_lists(self.local_words()) for w in part]

    def iter(self):
        """CBL::iter (src/cbl.rs:358-360): the stored k-mers in ascending word order, gathered like words()"""
        return iter([x for part in self._gather_lists(self.engine.kmers_list()) for x in part])

    __iter__ = iter

    def is_empty(self) -> bool:
        return self.count() == 0

    def my_prefix_range(self):
        lo = int(self.splitters_u32[self.rank - 1]) if self.rank > 0 else 0
        hi = int(self.splitters_u32[self.rank]) if self.rank < self.world - 1 else 1 << self.prefix_bits
        return lo, hi

    @staticmethod
    def _varint(v: int) -> bytes:   # bincode 1.3 varint
        if v < 251:
            return bytes([v])
        if v <= 0xFFFF:
            return b"\xfb" + v.to_bytes(2, "little")
        if v <= 0xFFFFFFFF:
            return b"\xfc" + v.to_bytes(4, "little")
        return b"\xfd" + v.to_bytes(8, "little")

    @staticmethod
    def _read_varint(b: bytes, at: int):
        t = b[at]
        if t < 251:
            return t, at + 1
        n = {251: 2, 252: 4, 253: 8}[t]
        return int.from_bytes(b[at + 1 : at + 1 + n], "little"), at + 1 + n

    def save_to_file(self, path: str) -> None:
        """ONE file in the reference's layout (src/cbl.rs:127-141): [canonical][bucket count][entries by ascending prefix] —
        every rank serialises its shard, the bodies are written in rank order behind a header with the summed bucket count.
        Collective; the path must be visible to every rank (one box)."""
        data = self.engine.serialize()
        nb, at = self._read_varint(data, 1)
        body = data[at:]
        info = self._gather_lists((nb, len(body)))
        header = data[:1] + self._varint(sum(i[0] for i in info))
        off = len(header) + sum(i[1] for i in info[: self.rank])
        if self.rank == 0:
            with open(path, "wb") as f:
                f.write(header)
                f.truncate(len(header) + sum(i[1] for i in info))
        self._barrier()
        fd = os.open(path, os.O_WRONLY)
        try:
            os.pwrite(fd, body, off)
        finally:
            os.close(fd)
        self._barrier()

    def load_from_file(self, path: str) -> "ShardedCBL":
        """A set sharded like this one holding the file's contents: every rank keeps the buckets of its prefix range."""
        with open(path, "rb") as f:
            data = f.read()
        lo, hi = self.my_prefix_range()
        out = self._derive(self.engine.deserialize_range(data, lo, hi))
        out.canonical = bool(data[0])
        return out

    def _barrier(self):
        if self.world > 1:
            if dist.get_backend(self.group) == "gloo":
                dist.barrier(group=self.group)
            else:
                dist.barrier(group=self.group, device_ids=[self.device.index])

    # -- global scalars --------------------------------------------------------------------------
    def local_count(self) -> int:
        return self.engine.count()

    def _sum_over_ranks(self, v: int) -> int:
        if self.world == 1:
            return int(v)
        dev = torch.device("cpu") if dist.get_backend(self.group) == "gloo" else self.device
        c = torch.tensor([int(v)], dtype=torch.int64, device=dev)
        dist.all_reduce(c, group=self.group)
        return int(c.item())

    def count(self) -> int:
        return self._sum_over_ranks(self.engine.count())

    def num_buckets(self) -> int:
        nb = getattr(self.engine, "cbl", None)
        return self._sum_over_ranks(nb.num_buckets() if nb is not None else 0)

    def close(self) -> None:
        """Collective: the last set that shares the peer buffers unmaps / frees them."""
        if self.peer is not None:
            px, self.peer = self.peer, None
            px.users -= 1
            if px.users <= 0:
                for k_, v in list(_PEER_CACHE.items()):
                    if v is px:
                        del _PEER_CACHE[k_]
                px.close()

    def stream_ptr(self) -> int:
        return self.engine.cbl.stream_ptr()
