/*
 * cbl_gpu.h — C ABI of libcbl_gpu: a device-resident CBL k-mer set on NVIDIA B200 (sm_100a).
 *
 * This boundary REPLACES the reference's Rust->C++ FFI, src/ffi.rs:7-20 (autocxx bindings of
 * cxx/rank_bv.h `RankBV` and cxx/tiered_vec.h `TieredVec32`, called per prefix group / per element
 * from src/bitvector/mod.rs:18-62 and src/wordset/mod.rs:87-237).  Per-element calls into a GPU are
 * a non-starter, so the boundary moves UP to one call per sequence / batch / set operation — the
 * granularity of the public methods of `CBL<K, T, PREFIX_BITS>` in src/cbl.rs.  Each entry point
 * names the reference method it serves.  INTEGRATION.md shows the Rust `extern "C"` block and shim.
 *
 * Conventions
 *   - every function returns int32_t status (CBL_OK == 0); no exception or unwind crosses the ABI;
 *     cbl_last_error(h) / cbl_last_global_error() give the message (the reference panics instead:
 *     src/cbl.rs:87-91,294-299,422-425 — a host shim turns non-zero statuses into those panics);
 *   - handles are opaque, created/destroyed only by the library; buffers are caller-owned, plain
 *     host memory unless the name ends in _dev (then: device pointers on the handle's GPU), and are
 *     never retained after the call returns;
 *   - one handle is externally synchronised (like `&mut self`); distinct handles may be used from
 *     distinct threads;
 *   - k-mers and words cross as two parallel arrays (lo, hi) of 64-bit halves; hi may be NULL when
 *     the value fits 64 bits;
 *   - sequences are raw nucleotide bytes (what needletail hands to the reference), batches are a
 *     concatenation + n_seqs+1 byte offsets.  A record shorter than K => CBL_EINVAL (the reference
 *     panics).  Non-ACGT bytes are dropped chunk by chunk exactly like the reference does (see the note at
 *     cbl_last_kmer_count); only the fused multi-GPU kernels (cbl_seq_route_dev, cbl_seq_contains_fused_dev) reject them with
 *     CBL_EINVAL, and the sharded handle (cbl_create_sharded) then takes its exact slow path through the host.
 */
#ifndef CBL_GPU_H
#define CBL_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cbl_handle cbl_t;

enum { CBL_OK = 0, CBL_EINVAL = 1, CBL_ECUDA = 2, CBL_ENOMEM = 3, CBL_ENCCL = 4, CBL_EIO = 5 };
enum { CBL_OP_OR = 0, CBL_OP_AND = 1, CBL_OP_SUB = 2, CBL_OP_XOR = 3 };

/* ---- life cycle: CBL::new / new_canonical (src/cbl.rs:71-79), Clone, Drop ------------------- */
/* k: K; word_bits: bits of T (32/64/128, checked like src/cbl.rs:87-91); prefix_bits: PREFIX_BITS */
int32_t cbl_create(uint32_t k, uint32_t word_bits, uint32_t prefix_bits, int32_t canonical, int32_t device, cbl_t** out);
/* One set prefix-sharded over n_gpus GPUs of THIS process (the boundary the survey specified): the
 * 2^PREFIX_BITS prefix space is cut into n_gpus contiguous ranges, shard i lives on devices[i]; every host-buffer
 * entry point below works on the handle (records are split over the GPUs, words travel to their owners over NVLink
 * peer memory, answers come back in read order; | & - ^ run shard by shard between handles created with the same
 * devices; iteration / export / serialisation see ONE ascending set).  The *_dev entry points (device pointers of one
 * GPU) return CBL_EINVAL on such a handle.  _ex: explicit splitters (n_gpus - 1 ascending prefixes) instead of the
 * built-in sample-based ones; cbl_sharded_splitters reads them back (out may be NULL to query the count). */
int32_t cbl_create_sharded(uint32_t k, uint32_t word_bits, uint32_t prefix_bits, int32_t canonical, int32_t n_gpus, const int32_t* devices, cbl_t** out);
int32_t cbl_create_sharded_ex(uint32_t k, uint32_t word_bits, uint32_t prefix_bits, int32_t canonical, int32_t n_gpus, const int32_t* devices,
                              const uint32_t* splitters, cbl_t** out);
int32_t cbl_sharded_splitters(cbl_t* h, uint32_t* out, size_t cap, size_t* n_out);
int32_t cbl_destroy(cbl_t* h);
int32_t cbl_clone(cbl_t* h, cbl_t** out);
const char* cbl_last_error(const cbl_t* h);
const char* cbl_last_global_error(void);

/* ---- scalar queries: count / is_empty / is_canonical (src/cbl.rs:162-177) ------------------- */
int32_t cbl_count(const cbl_t* h, uint64_t* out);
int32_t cbl_is_empty(const cbl_t* h, int32_t* out);   /* reference semantics incl. the all-ones-prefix quirk */
int32_t cbl_is_canonical(const cbl_t* h, int32_t* out);
int32_t cbl_num_buckets(const cbl_t* h, uint64_t* out);

/* ---- one sequence: insert_seq / remove_seq / contains_seq / contains_all (src/cbl.rs:293-354) */
int32_t cbl_insert_seq(cbl_t* h, const uint8_t* seq, size_t len);
int32_t cbl_remove_seq(cbl_t* h, const uint8_t* seq, size_t len);
/* out: len-K+1 bytes (0/1) in the reference's order (canonical mode: per 2048-k-mer chunk the
 * forward-canonical k-mers first, then the reverse-complemented ones; src/cbl.rs:248-275) */
int32_t cbl_contains_seq(cbl_t* h, const uint8_t* seq, size_t len, uint8_t* out, size_t* n_out);
int32_t cbl_contains_all(cbl_t* h, const uint8_t* seq, size_t len, int32_t* out);

/* ---- batches of records (the CLI loops `for record { insert_seq(record) }`, examples/cbl.rs:160-163,
 *      216-228; one call here processes the whole loop) ------------------------------------- */
int32_t cbl_insert_seqs(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs);
int32_t cbl_remove_seqs(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs);
int32_t cbl_contains_seqs(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs, uint8_t* out);
/* same, with the bytes (and answers) already resident in the handle's GPU memory; offsets on the host */
int32_t cbl_insert_seqs_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs);
int32_t cbl_remove_seqs_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs);
int32_t cbl_contains_seqs_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, uint8_t* d_out);
/* number of k-mers (= answers) the records yield */
int32_t cbl_count_kmers(const cbl_t* h, const uint64_t* offsets, size_t n_seqs, uint64_t* out);
/* Non-ACGT bytes: like the reference (filter_map in src/kmer.rs:133-135 and src/cbl.rs:262,282 while chunking is
 * on raw byte offsets, src/cbl.rs:239-243) every 2048-k-mer chunk drops them: its first k-mer is built from the
 * valid bytes among its first K bytes and every later valid byte yields one more word.  Such input takes a slower
 * path; words / answers come out compacted, in the reference's order, at the front of the output buffer.
 * cbl_last_kmer_count = how many the handle's last sequence call produced (= cbl_count_kmers for clean reads). */
int32_t cbl_last_kmer_count(const cbl_t* h, uint64_t* out);

/* ---- single k-mers: contains / insert / remove (src/cbl.rs:219-235), batched.  k-mers are IntKmer
 *      integers (first base most significant, A=0 C=1 T=2 G=3).  out[i] (may be NULL) = whether
 *      k-mer i was in the set BEFORE the call (insert() returns !out, remove() returns out) ------- */
int32_t cbl_contains_kmers(cbl_t* h, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out);
int32_t cbl_insert_kmers(cbl_t* h, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out);
int32_t cbl_remove_kmers(cbl_t* h, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out);

/* ---- set operations: | & - ^ and |= &= -= ^= (src/cbl.rs:411-569), merge / intersect (:108-124) */
int32_t cbl_setop(int32_t op, cbl_t* a, cbl_t* b, cbl_t** out);
int32_t cbl_setop_assign(int32_t op, cbl_t* a, cbl_t* b);
int32_t cbl_merge_many(cbl_t** hs, size_t n, cbl_t** out);
int32_t cbl_intersect_many(cbl_t** hs, size_t n, cbl_t** out);

/* ---- iteration: CBL::iter (src/cbl.rs:358-360) in ascending word order; start = element rank --- */
int32_t cbl_export_words(cbl_t* h, uint64_t start, uint64_t* lo, uint64_t* hi, size_t cap, size_t* n_out);
int32_t cbl_export_kmers(cbl_t* h, uint64_t start, uint64_t* lo, uint64_t* hi, size_t cap, size_t* n_out);
/* stats: buckets_sizes (src/cbl.rs:370-372): pass NULL arrays to query the bucket count */
int32_t cbl_bucket_sizes(cbl_t* h, uint32_t* prefixes, uint32_t* sizes, size_t cap, size_t* n_out);

/* ---- serde: save_to_file / load_from_file (src/cbl.rs:127-160), the reference's bincode varint
 *      layout (always written with sorted Vec buckets; both bucket variants are read) ------------- */
int32_t cbl_serialize_size(cbl_t* h, size_t* out);
int32_t cbl_serialize(cbl_t* h, uint8_t* out, size_t cap, size_t* n_out);
/* proto supplies K / T / PREFIX_BITS / device (the reference fixes them at compile time) */
int32_t cbl_deserialize(const cbl_t* proto, const uint8_t* data, size_t len, cbl_t** out);
/* same, keeping only the buckets whose prefix lies in [prefix_lo, prefix_hi): one rank's share of a file when the set is
 * sharded one process per GPU (cbl_b200/sharded.py; a cbl_create_sharded handle needs no such call) */
int32_t cbl_deserialize_range(const cbl_t* proto, const uint8_t* data, size_t len, uint64_t prefix_lo, uint64_t prefix_hi, cbl_t** out);
int32_t cbl_save_to_file(cbl_t* h, const char* path);
int32_t cbl_load_from_file(const cbl_t* proto, const char* path, cbl_t** out);

/* ---- sharded (one process per GPU) building blocks: word-level entry points used after the
 *      all-to-all that routes each word to the GPU owning its prefix range ----------- */
/* words of the records, in the reference's order, left on the device: d_words = n_kmers * (8|16) bytes */
int32_t cbl_seq_words_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, void* d_words);
/* op: 0 contains (d_out required), 1 insert, 2 remove (d_out optional = membership before the call);
 * d_out may be peer memory (the sharded path answers straight into the asking rank's buffer).  The membership test accepts any
 * bit pattern (what is not the word of a k-mer of this set is simply absent); insert / remove require words made by this
 * library (cbl_seq_words_dev, cbl_export_words_dev, the route kernels): 2K + POS_BITS significant bits, never all ones. */
int32_t cbl_words_op_dev(cbl_t* h, int32_t op, const void* d_words, size_t n, uint8_t* d_out);
/* insert (1) / remove (2) the words of n_seg device segments as ONE batch (the per-source regions of a sharded receive
 * buffer): the shard is rewritten once, not once per segment */
int32_t cbl_words_op_segments_dev(cbl_t* h, int32_t op, const void* const* seg, const uint64_t* seg_n, uint32_t n_seg);
/* membership of the words of n_seg device segments in ONE kernel launch: answers of segment i (one byte per word) go to
 * seg_out[i], which may be peer memory (the owner-side probe of the sharded contains_seq, src/cbl.rs:311-324) */
int32_t cbl_words_contains_segments_dev(cbl_t* h, const void* const* seg, const uint64_t* seg_n, uint8_t* const* seg_out, uint32_t n_seg);
int32_t cbl_export_words_dev(cbl_t* h, uint64_t start, uint64_t count, void* d_out);
/* router: stable partition of n words by owner rank, dest = number of splitters <= prefix (contiguous
 * prefix ranges).  d_send: the words grouped by destination; d_pos (may be NULL): for every input word
 * the slot it went to; counts (host, n_splitters+1): words per destination.  n < 2^30 per call. */
int32_t cbl_route_words_dev(cbl_t* h, const void* d_words, size_t n, const uint32_t* splitters, uint32_t n_splitters, void* d_send,
                            uint32_t* d_pos, uint64_t* counts);
/* d_out[i] = d_src[d_pos[i]] — puts the answers that came back from the owners into read order */
int32_t cbl_gather_u8_dev(cbl_t* h, const uint8_t* d_src, const uint32_t* d_pos, size_t n, uint8_t* d_out);
/* ---- fused route + exchange over NVLink peer memory (one process per GPU, buffers shared with CUDA IPC).
 *      Replaces "partition into a send buffer, then NCCL all-to-all" by ONE kernel whose coalesced stores
 *      land directly in the owner's receive buffer; NCCL (or gloo) only carries the count matrix and barriers. */
#define CBL_IPC_HANDLE_BYTES 64
/* a cudaMalloc block other processes may map: handle receives the CUDA IPC handle (CBL_IPC_HANDLE_BYTES bytes) */
int32_t cbl_peer_alloc(cbl_t* h, size_t bytes, void** d_ptr, uint8_t* handle);
int32_t cbl_peer_open(cbl_t* h, const uint8_t* handle, void** d_ptr);   /* map a peer's block (lazy peer access) */
int32_t cbl_peer_close(cbl_t* h, void* d_ptr);                          /* unmap */
int32_t cbl_peer_free(cbl_t* h, void* d_ptr);                           /* free an own block (peers must have unmapped it) */
/* words per owner rank (host array of n_splitters + 1 counts) */
int32_t cbl_route_counts_dev(cbl_t* h, const void* d_words, size_t n, const uint32_t* splitters, uint32_t n_splitters, uint64_t* counts);
/* word i goes to peer_recv[d][recv_offset[d] + j], d = its owner, j = its rank among this rank's words for d
 * (stable); counts = this rank's cbl_route_counts_dev result; d_pos[i] (may be NULL) = slot of word i in
 * destination-major send order (what cbl_gather_u8_dev needs).  Returns when the stores are complete. */
int32_t cbl_route_scatter_dev(cbl_t* h, const void* d_words, size_t n, const uint32_t* splitters, uint32_t n_splitters,
                              void* const* peer_recv, const uint64_t* recv_offset, const uint64_t* counts, uint32_t* d_pos);
/* ONE kernel for "reads -> words -> owner": fused 2-bit encode + necklace + route (src/cbl.rs:247-289 plus the
 * all-to-all of the sharded design).  Every word of the records is stored straight into peer_region[d], THIS rank's
 * region (cap words, 16-byte aligned) inside owner d's receive buffer; space is reserved chunk by chunk, so the order
 * inside a region is arbitrary.  counts[d] (host, n_splitters+1) = words sent to d; counts[d] > cap means the region
 * overflowed (nothing was written past cap): retry the call with a larger cap.  d_pos (may be NULL): for k-mer i (the
 * reference's order) d * cap + its index inside region d, i.e. where its answer comes back when every owner
 * answers region s of its receive buffer into region (owner rank) of rank s's answer buffer (cbl_words_op_dev with
 * d_out = that peer pointer).  (n_splitters+1) * cap < 2^32.  Returns when the stores are complete. */
int32_t cbl_seq_route_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, const uint32_t* splitters,
                          uint32_t n_splitters, void* const* peer_region, uint64_t cap, uint32_t* d_pos, uint64_t* counts);
/* The fused sharded contains_seq of one rank (src/cbl.rs:311-324 for a prefix-sharded set) as ONE kernel per GPU
 * (cbl_b200/csrc/shard_query.cuh): every warp alternates between producing (2-bit encode + necklace + route of this
 * rank's reads, every word stored straight into its owner's receive buffer over NVLink) and consuming (membership probe
 * of the 1024-word blocks the peers have completed in THIS rank's receive buffer, every answer stored straight back into
 * the asking rank's answer buffer), so the integer work of the necklace hides under the memory stalls of the probe as it
 * does in the single-GPU kernel and no host round trip or collective separates routing from probing.
 * The receive buffers must hold 0xFF bytes wherever no word has been stored (cbl_peer_fill once after allocation and
 * after any other use of the buffer, e.g. cbl_seq_route_dev, or a failed / overflowed call): an all-ones word is never
 * valid, so a word is its own arrival flag, and the consumer puts the 0xFF back as it reads.
 * Arrays of g = n_splitters + 1 pointers, indexed by rank: peer_region / peer_final = THIS rank's region (cap words, cap
 * a multiple of 1024) and u64 final-count slot at owner d; recv_region / final_counts = the same objects of source s in
 * THIS rank's own buffers; answer_region = this rank's region (cap bytes) in source s's answer buffer.  epoch: 1..65535,
 * the same on every rank and different from the previous call's (the final-count slots carry it, so they need no
 * zeroing).  Every rank of the group must make the call (a rank without reads passes n_seqs = 0); a rank whose peers
 * never show up fails with CBL_ECUDA after 20 s.  d_pos and counts as for cbl_seq_route_dev (counts[d] > cap:
 * overflow, refill the buffers and retry with a larger cap).  Returns when the kernel of THIS rank is done, i.e. when it
 * has answered every block sent to it; a barrier over all ranks then guarantees every answer has landed. */
int32_t cbl_seq_contains_fused_dev(cbl_t* h, const uint8_t* d_buf, const uint64_t* offsets, size_t n_seqs, const uint32_t* splitters,
                                   uint32_t n_splitters, void* const* peer_region, void* const* peer_final, uint64_t cap, uint32_t* d_pos,
                                   void* const* recv_region, uint8_t* const* answer_region, const void* const* final_counts, uint32_t epoch,
                                   uint64_t* counts);
int32_t cbl_peer_fill(cbl_t* h, void* d_ptr, int32_t byte, size_t bytes);   /* memset a block of (own) device memory, synchronous */
int32_t cbl_peer_zero(cbl_t* h, void* d_ptr, size_t bytes);           /* = cbl_peer_fill(.., 0, ..) */
int32_t cbl_word_bytes(const cbl_t* h, int32_t* out);      /* 8 or 16: size of one device word */
int32_t cbl_suffix_bits(const cbl_t* h, int32_t* out);

/* ---- diagnostics ----------------------------------------------------------------------------- */
/* words of the records on the host (same order as cbl_seq_words_dev); brute != 0 uses the normative
 * brute-force necklace instead of the fast path */
int32_t cbl_seq_words(cbl_t* h, const uint8_t* buf, const uint64_t* offsets, size_t n_seqs, uint64_t* lo, uint64_t* hi, int32_t brute);
int32_t cbl_sync(cbl_t* h);
void* cbl_stream(const cbl_t* h);             /* the handle's cudaStream_t */
uint64_t cbl_launch_count(void);              /* kernels launched by this library so far */
/* batches whose segment sort (seg_sort.cuh) met a group of words too long for its tile and were re-sorted by the plain
 * LSD passes: 0 for k-mer data, > 0 only for heavily repeated words (results are identical either way) */
uint64_t cbl_sort_fallback_count(void);
const char* cbl_build_info(void);
/* Planning hint for the batch sort: the words this handle will be given cover about 1 / factor of the prefix mass (a shard of a
 * set sharded over `factor` GPUs by prefix range: every head of a necklace word is `factor` times as frequent as in a whole set,
 * so the hybrid sort needs an extra LSD pass before its segment sort).  Results never depend on it: a wrong plan costs one
 * re-sort by plain LSD passes (cbl_sort_fallback_count), after which the handle corrects the factor itself.  cbl_create_sharded
 * sets it on its shards; a host that shards across processes (cbl_b200/sharded.py) calls it on every shard handle. */
int32_t cbl_set_sort_concentration(cbl_t* h, double factor);
/* device memory: the library keeps the blocks it frees in a per-stream arena (no driver allocation in the steady
 * state; the reference relies on the Rust global allocator the same way).  cbl_mem_trim returns every cached block
 * of `device` to the driver (synchronises the device); cbl_mem_cached_bytes = bytes currently cached. */
int32_t cbl_mem_trim(int32_t device);
uint64_t cbl_mem_cached_bytes(void);
/* per-kernel device time (CUDA events around every launch); off by default.  report: JSON text
 * {"kernel": {"n": launches, "ms": total}, ...}, clears the accumulated records */
void cbl_profile_enable(int32_t on);
int32_t cbl_profile_report(char* out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* CBL_GPU_H */
