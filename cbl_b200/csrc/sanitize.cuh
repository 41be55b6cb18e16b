// Slow path for reads that contain bytes other than ACGT/acgt.
//
// The reference cuts a sequence into chunks on RAW byte offsets (src/cbl.rs:239-243) and, inside a chunk,
// drops every non-nucleotide byte with filter_map (src/kmer.rs:133-135, src/cbl.rs:262,282): the first k-mer is
// built from the valid bytes among the chunk's first K bytes (Kmer::from_nucs folds `extend` from zero, so a
// short first k-mer reads as if it were left-padded with 'A'), and every later valid byte yields one more word.
// A chunk with b valid bytes after its first K therefore behaves exactly like the clean string
//        'A' * (K - a)  ++  (the a valid bytes of the first K)  ++  (the b valid bytes of the rest)
// of length K + b, which produces 1 + b words.  These kernels build that string for every chunk of a batch
// (one CTA per chunk); the ordinary fused kernels then run on the cleaned records, one single-chunk piece per
// reference chunk, so word order (incl. the canonical per-chunk partition, F6) is the reference's.
#pragma once
#include "seq_words.cuh"

namespace cbl {

constexpr int SAN_THREADS = 256;
constexpr int SAN_BYTES = 9;   // bytes per thread: 256 * 9 = 2304 >= 2048 + 63 - 1 (K <= 63)

__device__ __forceinline__ bool is_nuc(uint8_t c) {   // src/kmer.rs:11-24
    c &= 0xDF;
    return c == 'A' || c == 'C' || c == 'G' || c == 'T';
}

// counts == nullptr: compaction (clean + out_byte given); otherwise count mode: counts[chunk] = b
static __global__ void __launch_bounds__(SAN_THREADS) sanitize_chunks_kernel(SeqBatch b, int k, uint32_t* __restrict__ counts,
                                                                             const uint64_t* __restrict__ out_byte, uint8_t* __restrict__ clean) {
    __shared__ uint32_t s_tmp[33];
    __shared__ uint32_t s_piece, s_a;
    for (uint64_t chunk = blockIdx.x; chunk < b.n_chunks; chunk += gridDim.x) {
        if (threadIdx.x == 0) s_piece = (uint32_t)(upper_bound_dev<uint64_t>(b.piece_chunk0, (uint64_t)b.n_pieces + 1, chunk) - 1);
        __syncthreads();
        const uint32_t piece = s_piece;
        const uint32_t ks = (uint32_t)((chunk - b.piece_chunk0[piece]) * CHUNK_KMERS);
        const int m = (int)min((uint32_t)CHUNK_KMERS, b.piece_kmers[piece] - ks);
        const int nbytes = m + k - 1;
        const uint8_t* src = b.seq + b.piece_byte[piece] + ks;
        const int i0 = threadIdx.x * SAN_BYTES;
        uint8_t c[SAN_BYTES];
        uint32_t valid = 0, first = 0;   // valid bytes of this thread, those among the chunk's first K bytes
#pragma unroll
        for (int j = 0; j < SAN_BYTES; j++) {
            const int i = i0 + j;
            c[j] = i < nbytes ? src[i] : (uint8_t)0;
            const bool ok = i < nbytes && is_nuc(c[j]);
            valid += ok;
            first += ok && i < k;
        }
        uint32_t total;
        const uint32_t before = block_excl_scan<uint32_t, SAN_THREADS>(valid, s_tmp, total);
        __syncthreads();
        const uint32_t a_part = warp_sum(first);
        if (threadIdx.x == 0) s_a = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0 && a_part) atomicAdd(&s_a, a_part);
        __syncthreads();
        const uint32_t a = s_a;
        if (counts != nullptr) {
            if (threadIdx.x == 0) counts[chunk] = total - a;
        } else {
            uint8_t* dst = clean + out_byte[chunk];
            const uint32_t pad = (uint32_t)k - a;
            for (uint32_t i = threadIdx.x; i < pad; i += SAN_THREADS) dst[i] = 'A';
            uint32_t r = pad + before;
#pragma unroll
            for (int j = 0; j < SAN_BYTES; j++) {
                const int i = i0 + j;
                if (i < nbytes && is_nuc(c[j])) dst[r++] = c[j];
            }
        }
        __syncthreads();
    }
}

}  // namespace cbl
