// Host-side orchestration of one device-resident CBL shard (the B200 counterpart of src/cbl.rs).
// One handle = one CUDA stream; every public operation enqueues its kernels on that stream and
// returns when the results the caller asked for are on the host.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

namespace cbl {

enum SetOp : int { SETOP_OR = 0, SETOP_AND = 1, SETOP_SUB = 2, SETOP_XOR = 3 };

struct Config {
    int k, word_bits, prefix_bits, canonical, device;
};

// Records of one batch, already cut into 2048-aligned pieces (host arrays).
struct PieceList {
    std::vector<uint64_t> byte_off, out_off, chunk0;
    std::vector<uint32_t> kmers;
    uint64_t n_kmers = 0, n_chunks = 0;
};

// pointers of one rank's fused sharded query (producer + consumer over peer memory); g = n_split + 1 ranks
struct FusedQuery {
    const uint32_t* splitters;
    uint32_t n_split;
    void* const* peer_region;                    // [g] this rank's region inside owner d's receive buffer
    unsigned long long* const* peer_final;       // [g] this rank's final-count slot at owner d
    uint64_t cap;                                // words per region (multiple of 1024)
    uint32_t* d_pos;                             // [k-mers] answer slot of every local k-mer (d * cap + index inside region d)
    void* const* recv_region;                    // [g] region of this rank's receive buffer written by source s (sentinel-filled)
    uint8_t* const* answer_region;               // [g] this rank's region inside source s's answer buffer
    const unsigned long long* const* final_;     // [g] local final-count slot of source s
    uint32_t epoch;                              // 1 .. 65535: same on every rank, different from the previous call's
    uint32_t grid_share = 1;                     // ranks whose kernels share THIS device (they must all be resident together)
};

class IIndex {
public:
    virtual ~IIndex() {}
    virtual const Config& config() const = 0;
    virtual const KParams& params() const = 0;
    virtual cudaStream_t stream() const = 0;
    virtual uint64_t count() const = 0;
    virtual uint32_t n_buckets() const = 0;
    virtual bool is_empty_reference_semantics() const = 0;
    virtual IIndex* clone() = 0;
    virtual IIndex* new_empty(int canonical = -1) = 0;   // (canonical < 0: same flag) an empty set with the same parameters (and, for a sharded set, devices and splitters)
    // sequences: `d_seq` device pointer to n_bytes bytes, host offsets[n_seqs + 1]
    virtual void insert_seqs_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs) = 0;
    virtual void remove_seqs_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs) = 0;
    virtual void contains_seqs_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, uint8_t* d_out) = 0;
    virtual void seq_words_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, void* d_words, bool brute) = 0;
    // host-buffer front ends (copies inside)
    virtual void insert_seqs(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, bool remove) = 0;
    virtual void contains_seqs(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, uint8_t* out) = 0;
    virtual void seq_words(const uint8_t* seq, const uint64_t* offsets, size_t n_seqs, uint64_t* lo, uint64_t* hi, bool brute) = 0;
    // k-mer integers
    virtual void kmers_op(int op /*0 contains,1 insert,2 remove*/, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) = 0;
    // words on the host: op 0 contains, 1 insert, 2 remove; out (may be null) = membership before the call
    virtual void words_op(int op, const uint64_t* lo, const uint64_t* hi, size_t n, uint8_t* out) = 0;
    // words (already transformed) — used by the sharded multi-GPU path after the all-to-all
    virtual void words_op_dev(int op, const void* d_words, uint64_t n, uint8_t* d_out) = 0;
    virtual void words_op_segments_dev(int op, const void* const* seg, const uint64_t* seg_n, uint32_t n_seg) = 0;
    virtual void words_contains_segments_dev(const void* const* seg, const uint64_t* seg_n, uint8_t* const* seg_out, uint32_t n_seg) = 0;
    // multi-GPU routing: stable partition of words by owner rank (dest = #splitters <= prefix)
    virtual void route_words_dev(const void* d_words, uint64_t n, const uint32_t* splitters, uint32_t n_split, void* d_send,
                                 uint32_t* d_pos, uint64_t* counts) = 0;
    virtual void gather_u8_dev(const uint8_t* d_src, const uint32_t* d_pos, uint64_t n, uint8_t* d_out) = 0;
    virtual void seq_contains_fused_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, const FusedQuery& q,
                                        uint64_t* counts) = 0;
    // fused route + exchange over peer memory (see include/cbl_gpu.h)
    virtual void route_counts_dev(const void* d_words, uint64_t n, const uint32_t* splitters, uint32_t n_split, uint64_t* counts) = 0;
    virtual void route_scatter_dev(const void* d_words, uint64_t n, const uint32_t* splitters, uint32_t n_split, void* const* peer_recv,
                                   const uint64_t* recv_offset, const uint64_t* counts, uint32_t* d_pos) = 0;
    virtual void seq_route_dev(const uint8_t* d_seq, uint64_t n_bytes, const uint64_t* offsets, size_t n_seqs, const uint32_t* splitters,
                               uint32_t n_split, void* const* peer_region, uint64_t cap, uint32_t* d_pos, uint64_t* counts) = 0;
    // set operations
    virtual IIndex* setop(int op, IIndex* other) = 0;
    virtual void setop_assign(int op, IIndex* other) = 0;
    // export
    virtual void export_words(uint64_t start, uint64_t cap, int to_kmers, uint64_t* lo, uint64_t* hi, uint64_t* n_out) = 0;
    virtual void export_words_dev(uint64_t start, uint64_t count, int to_kmers, void* d_out) = 0;
    virtual void bucket_sizes(uint32_t* prefixes, uint32_t* sizes, uint64_t cap, uint64_t* n_out) = 0;
    virtual void load_sorted_words(const uint64_t* lo, const uint64_t* hi, uint64_t n) = 0;
    virtual void sync() = 0;
    // the batches this handle sorts cover about 1 / factor of the prefix mass (a shard of a prefix-sharded set: factor = number of
    // shards): the hybrid sort then plans its LSD passes for groups that much denser (plan_sort)
    virtual void set_sort_concentration(double factor) = 0;
    std::string last_error;
    // words / answers produced by the last sequence call: sum of (len - K + 1) over the records, or fewer when the
    // reads held non-ACGT bytes and the reference's dropping behaviour was reproduced
    uint64_t last_produced = 0;
};

IIndex* make_index(const Config& cfg);
// one set prefix-sharded over several GPUs of this process (sharded_index.cu); splitters: n_gpus - 1 ascending prefixes or null
IIndex* make_sharded_index(const Config& cfg, const int* devices, int n_gpus, const uint32_t* splitters);
bool sharded_splitters(IIndex* ix, std::vector<uint32_t>& out);   // false when ix is not a sharded set
std::string prof_report();
uint64_t total_kmers_of(const Config& cfg, const uint64_t* offsets, size_t n_seqs);

}  // namespace cbl
