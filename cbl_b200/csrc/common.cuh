// Host-side plumbing shared by the .cu files: error handling that never lets an exception cross the
// C ABI, a stream-ordered device buffer, and the kernel-launch counter bench.py reports.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <utility>

#include "../../include/cbl_gpu.h"  // status codes
#include "kmer_necklace.cuh"

namespace cbl {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %s (%s) at %s:%d: %s", cudaGetErrorName(e), what, file, line, cudaGetErrorString(e));
        throw Error(e == cudaErrorMemoryAllocation ? CBL_ENOMEM : CBL_ECUDA, buf);
    }
}
#define CUDA_CHECK(x) ::cbl::cuda_check((x), #x, __FILE__, __LINE__)

extern std::atomic<uint64_t> g_launches;  // every kernel launched by this library
extern std::atomic<uint64_t> g_sort_fallbacks;  // batches the segment sort handed back to the plain LSD passes

// Optional per-kernel device timing (CUDA events on the launching stream), off by default.
// bench.py turns it on for a separate, untimed pass to attribute time to kernels.
struct ProfRec { const char* tag; cudaEvent_t a, b; };
extern std::atomic<int> g_prof_on;
void prof_push(const char* tag, cudaEvent_t a, cudaEvent_t b);

#define CBL_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
    do {                                                                   \
        cudaEvent_t _e0 = nullptr, _e1 = nullptr;                          \
        const bool _prof = ::cbl::g_prof_on.load(std::memory_order_relaxed) != 0; \
        if (_prof) { cudaEventCreate(&_e0); cudaEventCreate(&_e1); cudaEventRecord(_e0, (stream)); } \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);        \
        if (_prof) { cudaEventRecord(_e1, (stream)); ::cbl::prof_push(#kernel, _e0, _e1); } \
        ::cbl::g_launches.fetch_add(1, std::memory_order_relaxed);         \
        CUDA_CHECK(cudaGetLastError());                                    \
    } while (0)

// Device memory arena.  cudaMallocAsync's pool maps fresh physical memory at ~25 GB/s on B200 (measured: 78 ms for
// the two 1 GB sort buffers of a batch) and, without a host sync between batches, cannot reuse a block whose free is
// still pending, so a build paid the driver for every batch.  The arena keeps every block it ever got (cudaMalloc,
// size classes with <= 12.5 % slack) in a free list PER STREAM: a block freed on stream s is handed out again only
// to work on s, where stream order makes the reuse safe without any synchronisation; blocks of a retired
// (synchronised, destroyed) stream move to an idle list any stream may take from.  Steady-state batches therefore
// perform no driver allocation at all.  CBL_ARENA=0 falls back to cudaMallocAsync.
namespace arena {
void* alloc(size_t bytes, cudaStream_t s);      // on the current device
void release(void* p, cudaStream_t s);          // p may still be in use by work already enqueued on s
void retire_stream(cudaStream_t s);             // s has been synchronised and is about to be destroyed
void trim();                                    // give every cached block of the current device back to the driver
uint64_t cached_bytes();                        // bytes held in free lists (all devices)
}  // namespace arena

template <class T>
class DevBuf {
    T* p_ = nullptr;
    size_t n_ = 0;
    cudaStream_t s_ = nullptr;

public:
    DevBuf() = default;
    DevBuf(size_t n, cudaStream_t s) { alloc(n, s); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p_(o.p_), n_(o.n_), s_(o.s_) { o.p_ = nullptr; o.n_ = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p_ = o.p_; n_ = o.n_; s_ = o.s_; o.p_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t n, cudaStream_t s) {
        release();
        s_ = s;
        n_ = n;
        p_ = static_cast<T*>(arena::alloc((n ? n : 1) * sizeof(T), s));
    }
    void release() {
        if (p_) { arena::release(p_, s_); p_ = nullptr; n_ = 0; }
    }
    void zero() { if (p_) CUDA_CHECK(cudaMemsetAsync(p_, 0, (n_ ? n_ : 1) * sizeof(T), s_)); }
    T* get() const { return p_; }
    size_t size() const { return n_; }
    size_t bytes() const { return n_ * sizeof(T); }
    void rebind(cudaStream_t s) { s_ = s; }  // later frees are ordered on `s` (all prior work must be complete)
    void swap(DevBuf& o) { std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(s_, o.s_); }
};

CBL_HD uint64_t div_up(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

}  // namespace cbl
