// Process-wide host services of the library (per-kernel timing records, the device memory arena) and the
// index factory; the index class itself lives in cbl_index_impl.cuh (one translation unit per instantiation).
#include "cbl_index.cuh"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>


namespace cbl {

std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_sort_fallbacks{0};
std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
void prof_push(const char* tag, cudaEvent_t a, cudaEvent_t b) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back({tag, a, b});
}
// JSON object {"kernel": {"n": launches, "ms": total device ms}, ...}; clears the records
std::string prof_report() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<uint64_t, double>> acc;
    for (auto& r : g_prof_recs) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        std::string t;
        for (const char* c = r.tag; *c; c++) if (*c != '(' && *c != ')' && *c != ' ') t.push_back(*c);
        auto& e = acc[t];
        e.first++;
        e.second += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : acc) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s\"%s\": {\"n\": %llu, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
                 (unsigned long long)kv.second.first, kv.second.second);
        out += buf;
        first = false;
    }
    return out + "}";
}

// ------------------------------------------------------------------------------------------------
// device memory arena (see common.cuh)
// ------------------------------------------------------------------------------------------------
namespace arena {
namespace {
struct Block { size_t bytes; int device; };
struct State {
    std::mutex mu;
    std::unordered_map<void*, Block> live;                                         // every block handed out
    std::map<std::pair<int, cudaStream_t>, std::multimap<size_t, void*>> by_stream; // free, reusable on that stream
    std::map<int, std::multimap<size_t, void*>> idle;                               // free, reusable anywhere on the device
    std::unordered_map<void*, Block> cached;                                       // blocks sitting in a free list
    uint64_t cached_bytes = 0;
    int mode = -1;                                                                 // 1 arena, 0 cudaMallocAsync
};
State& st() { static State* s = new State; return *s; }   // leaked on purpose: CUDA may be gone at static destruction
bool enabled(State& S) {
    if (S.mode < 0) { const char* e = getenv("CBL_ARENA"); S.mode = (e && e[0] == '0') ? 0 : 1; }
    return S.mode == 1;
}
size_t size_class(size_t b) {
    if (b < 512) return 512;
    const int top = 63 - __builtin_clzll(b);
    if (b <= (1u << 20)) return (b & (b - 1)) ? (size_t)1 << (top + 1) : b;        // power of two up to 1 MB
    const size_t step = (size_t)1 << (top - 3);                                    // 8 classes per octave above
    return (b + step - 1) / step * step;
}
void* take(std::multimap<size_t, void*>& m, size_t need, size_t limit) {
    auto it = m.lower_bound(need);
    if (it == m.end() || it->first > limit) return nullptr;
    void* p = it->second;
    m.erase(it);
    return p;
}
void trim_locked(State& S, int dev) {
    cudaDeviceSynchronize();
    for (auto it = S.by_stream.begin(); it != S.by_stream.end(); ++it)
        if (it->first.first == dev) { for (auto& kv : it->second) { S.cached_bytes -= kv.first; S.cached.erase(kv.second); cudaFree(kv.second); } it->second.clear(); }
    auto& idle = S.idle[dev];
    for (auto& kv : idle) { S.cached_bytes -= kv.first; S.cached.erase(kv.second); cudaFree(kv.second); }
    idle.clear();
}
}  // namespace

void* alloc(size_t bytes, cudaStream_t s) {
    State& S = st();
    std::lock_guard<std::mutex> lk(S.mu);
    if (!enabled(S)) {
        void* p = nullptr;
        CUDA_CHECK(cudaMallocAsync(&p, bytes, s));
        return p;
    }
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    const size_t need = size_class(bytes);
    const size_t limit = need <= (1u << 20) ? need * 2 : need + need / 2;
    void* p = take(S.by_stream[{dev, s}], need, limit);
    if (!p) p = take(S.idle[dev], need, limit);
    if (p) {
        auto it = S.cached.find(p);
        S.cached_bytes -= it->second.bytes;
        S.live[p] = it->second;
        S.cached.erase(it);
        return p;
    }
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaErrorMemoryAllocation) {  // give the cache back and try once more
        cudaGetLastError();
        trim_locked(S, dev);
        e = cudaMalloc(&p, need);
    }
    CUDA_CHECK(e);
    S.live[p] = Block{need, dev};
    return p;
}
void release(void* p, cudaStream_t s) {
    State& S = st();
    std::lock_guard<std::mutex> lk(S.mu);
    auto it = S.live.find(p);
    if (it == S.live.end()) { cudaFreeAsync(p, s); return; }   // allocated in cudaMallocAsync mode
    const Block b = it->second;
    S.live.erase(it);
    S.cached[p] = b;
    S.cached_bytes += b.bytes;
    S.by_stream[{b.device, s}].emplace(b.bytes, p);
}
void retire_stream(cudaStream_t s) {
    State& S = st();
    std::lock_guard<std::mutex> lk(S.mu);
    for (auto it = S.by_stream.begin(); it != S.by_stream.end();) {
        if (it->first.second == s) {
            auto& idle = S.idle[it->first.first];
            for (auto& kv : it->second) idle.emplace(kv.first, kv.second);
            it = S.by_stream.erase(it);
        } else ++it;
    }
}
void trim() {
    State& S = st();
    std::lock_guard<std::mutex> lk(S.mu);
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) trim_locked(S, dev);
}
uint64_t cached_bytes() {
    State& S = st();
    std::lock_guard<std::mutex> lk(S.mu);
    return S.cached_bytes;
}
}  // namespace arena


static int pos_bits_for(int kmer_bits) {  // src/cbl.rs:66
    int p = 0;
    while ((1 << p) < kmer_bits) p++;
    return p;
}
// one factory per instantiation (inst_*.cu)
IIndex* make_index_u64_u32(const Config& cfg);
IIndex* make_index_u64_u64(const Config& cfg);
IIndex* make_index_u128_u64(const Config& cfg);
IIndex* make_index_u128_u128(const Config& cfg);


uint64_t total_kmers_of(const Config& cfg, const uint64_t* offsets, size_t n_seqs) {
    uint64_t t = 0;
    for (size_t i = 0; i < n_seqs; i++) {
        uint64_t len = offsets[i + 1] - offsets[i];
        if (len >= (uint64_t)cfg.k) t += len - cfg.k + 1;
    }
    return t;
}

IIndex* make_index(const Config& cfg) {
    // parameter rules of the reference: build.rs:15-53, src/cbl.rs:87-91, src/wordset/mod.rs:37-41
    if (cfg.k < 1 || cfg.k > 59) throw Error(CBL_EINVAL, "K must be in 1..=59");
    const int bits = 2 * cfg.k, pb = pos_bits_for(bits), word = bits + pb;
    if (cfg.word_bits != 32 && cfg.word_bits != 64 && cfg.word_bits != 128) throw Error(CBL_EINVAL, "T must be u32, u64 or u128");
    if (word > cfg.word_bits)
        throw Error(CBL_EINVAL, "Cannot fit a " + std::to_string(cfg.k) + "-mer and its length in a " + std::to_string(cfg.word_bits) + "-bit integer");
    if (cfg.prefix_bits < 1 || cfg.prefix_bits > 31) throw Error(CBL_EINVAL, "PREFIX_BITS should be in 1..=31 on the GPU path (reference: <= 32)");
    const int sb = word - cfg.prefix_bits;
    if (sb <= 0) throw Error(CBL_EINVAL, "SUFFIX_BITS should be != 0");
    if (cfg.canonical && (cfg.k % 2 == 0)) throw Error(CBL_EINVAL, "canonical k-mers need an odd K (build.rs:22)");
    int ndev = 0;
    CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (cfg.device < 0 || cfg.device >= ndev) throw Error(CBL_EINVAL, "no such CUDA device");
    // device key = smallest of u64 / u128 that holds the word, independent of the host-side T
    if (word <= 64 && sb <= 32) return make_index_u64_u32(cfg);
#ifdef CBL_FAST_BUILD  // developer builds only (make FAST=1): one instantiation, quick to compile
    throw Error(CBL_EINVAL, "this is a FAST developer build: only u64 words with <= 32-bit suffixes are compiled in");
#else
    if (word <= 64) return make_index_u64_u64(cfg);
    if (sb <= 64) return make_index_u128_u64(cfg);
    return make_index_u128_u128(cfg);
#endif
}

}  // namespace cbl
