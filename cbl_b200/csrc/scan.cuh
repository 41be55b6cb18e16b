// Block / warp scan primitives and the decoupled look-back used by every single-pass kernel
// (unique, edit compaction, merge, rank directory, offsets scan).
#pragma once
#include "common.cuh"

namespace cbl {

constexpr uint64_t LB_FLAG_AGG = 1ull << 62;   // tile aggregate published
constexpr uint64_t LB_FLAG_INCL = 2ull << 62;  // inclusive prefix published
constexpr uint64_t LB_VAL_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1; }

template <class T> __device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, d);
        if ((int)lane_id() >= d) v += t;
    }
    return v;
}
template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Exclusive scan of one value per thread across the block; returns the exclusive prefix and the
// block total.  `tmp` must hold >= 33 T's of shared memory.  All threads of the block must call.
template <class T, int THREADS> __device__ __forceinline__ T block_excl_scan(T v, T* tmp, T& total) {
    constexpr int NW = THREADS / 32;
    T incl = warp_incl_scan(v);
    const int w = threadIdx.x >> 5;
    if (lane_id() == 31) tmp[w] = incl;
    __syncthreads();
    if (w == 0) {
        T x = (int)lane_id() < NW ? tmp[lane_id()] : T(0);
        T xi = warp_incl_scan(x);
        if ((int)lane_id() < NW) tmp[lane_id()] = xi - x;  // exclusive warp offsets
        if (lane_id() == 31) tmp[32] = xi;                 // block total
    }
    __syncthreads();
    T res = tmp[w] + incl - v;
    total = tmp[32];
    __syncthreads();  // tmp may be reused by the caller right away
    return res;
}

// Decoupled look-back over 64-bit status words (2 flag bits + 62-bit value), executed by warp 0.
// Tiles must be numbered by a ticket counter so every predecessor is already resident.
// Returns the exclusive prefix of `agg` over all earlier tiles (same value in every lane of warp 0).
// Split form: lookback_publish (one thread, as early as possible: successors can then add this tile's aggregate
// without waiting) and lookback_resolve (warp 0, as late as possible: everything that does not need the prefix can
// be done in between, by all warps).
__device__ __forceinline__ void lookback_publish(volatile uint64_t* status, uint32_t tile, uint64_t agg) {
    status[tile] = (tile == 0 ? LB_FLAG_INCL : LB_FLAG_AGG) | agg;
}
__device__ __forceinline__ uint64_t lookback_resolve(volatile uint64_t* status, uint32_t tile, uint64_t agg) {
    const unsigned lane = lane_id();
    if (tile == 0) return 0;
    uint64_t excl = 0;
    long long base = (long long)tile - 1;
    for (;;) {
        long long idx = base - (long long)lane;
        uint64_t s;
        if (idx >= 0) {
            do { s = status[idx]; } while ((s >> 62) == 0);
        } else {
            s = LB_FLAG_INCL;  // virtual tile before the first one: inclusive prefix 0
        }
        unsigned incl = __ballot_sync(0xffffffffu, (s >> 62) == 2);
        if (incl) {
            int first = __ffs(incl) - 1;
            uint64_t v = ((int)lane <= first) ? (s & LB_VAL_MASK) : 0;
            excl += warp_sum(v);
            break;
        }
        excl += warp_sum(s & LB_VAL_MASK);
        base -= 32;
    }
    if (lane == 0) status[tile] = LB_FLAG_INCL | (excl + agg);
    return excl;
}
__device__ __forceinline__ uint64_t lookback_warp(volatile uint64_t* status, uint32_t tile, uint64_t agg) {
    if (lane_id() == 0) lookback_publish(status, tile, agg);
    return lookback_resolve(status, tile, agg);
}

// Ticket + look-back for a whole block: returns the tile id (via *tile_out) and the exclusive prefix.
// Usage: tile = block_ticket(counter, smem_u32); ... excl = block_lookback(status, tile, agg, smem_u64)
__device__ __forceinline__ uint32_t block_ticket(uint32_t* counter, uint32_t* sh) {
    if (threadIdx.x == 0) *sh = atomicAdd(counter, 1u);
    __syncthreads();
    uint32_t t = *sh;
    __syncthreads();
    return t;
}
__device__ __forceinline__ uint64_t block_lookback(volatile uint64_t* status, uint32_t tile, uint64_t agg, uint64_t* sh) {
    if (threadIdx.x < 32) {
        uint64_t e = lookback_warp(status, tile, agg);
        if (threadIdx.x == 0) *sh = e;
    }
    __syncthreads();
    uint64_t e = *sh;
    __syncthreads();
    return e;
}

template <class T> __device__ __forceinline__ uint64_t lower_bound_dev(const T* a, uint64_t n, T v) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
template <class T> __device__ __forceinline__ uint64_t upper_bound_dev(const T* a, uint64_t n, T v) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace cbl
