// Canary-band allocator behind every cudaMalloc / cudaFree of the library (see fct_common.cuh).
#define FCT_GUARD_IMPL
#include "fct_common.cuh"
#include "../../include/fctpdeco.h"

#include <stdlib.h>
#include <map>
#include <mutex>
#include <vector>

namespace {
const size_t kBand = 4096;
const unsigned char kCanary = 0xA5;
std::mutex g_mu;
std::map<void*, size_t> g_live;        // user pointer -> requested bytes
long long g_corrupt_at_free = 0;

int guard_mode() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FCT_GUARD"); v = e ? atoi(e) : 0; }
    return v;
}
size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

// number of canary bytes of one buffer that no longer hold the pattern
long long damaged(void* user, size_t bytes) {
    unsigned char* base = (unsigned char*)user - kBand;
    const size_t tail = padded(bytes) - bytes + kBand;
    std::vector<unsigned char> h(kBand + tail);
    if (cudaMemcpy(h.data(), base, kBand, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (cudaMemcpy(h.data() + kBand, (unsigned char*)user + bytes, tail, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    long long bad = 0;
    for (unsigned char c : h) bad += (c != kCanary);
    return bad;
}
}  // namespace

cudaError_t fct_guard_malloc(void** p, size_t bytes) {
    if (!guard_mode()) return cudaMalloc(p, bytes);
    unsigned char* base = nullptr;
    const size_t pb = padded(bytes);
    cudaError_t e = cudaMalloc((void**)&base, pb + 2 * kBand);
    if (e != cudaSuccess) return e;
    cudaMemset(base, kCanary, pb + 2 * kBand);
    cudaMemset(base + kBand, 0xFF, bytes);              // poison: a read before the first write shows up as NaN / -1
    *p = base + kBand;
    std::lock_guard<std::mutex> lk(g_mu);
    g_live[*p] = bytes;
    return cudaSuccess;
}

cudaError_t fct_guard_free(void* p) {
    if (!guard_mode() || !p) return cudaFree(p);
    size_t bytes = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_live.find(p);
        if (it == g_live.end()) return cudaFree(p);    // not ours (allocated before the mode was read, or foreign)
        bytes = it->second;
        g_live.erase(it);
    }
    cudaDeviceSynchronize();
    const long long bad = damaged(p, bytes);
    if (bad != 0) {
        fprintf(stderr, "libfctpdeco FCT_GUARD: buffer %p (%zu bytes) freed with %lld damaged canary bytes\n", p, bytes, bad);
        std::lock_guard<std::mutex> lk(g_mu);
        g_corrupt_at_free++;
    }
    return cudaFree((unsigned char*)p - kBand);
}

// corrupted: buffers (live now, or freed since the last call) whose canary bands were written to; live: buffers checked.
// Both are 0 when FCT_GUARD is off.
extern "C" int fct_guard_check(int64_t* corrupted, int64_t* live) {
    FCT_CHECK(corrupted && live, "fct_guard_check: null argument");
    *corrupted = 0; *live = 0;
    if (!guard_mode()) return 0;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_live) {
        const long long bad = damaged(kv.first, kv.second);
        if (bad != 0) {
            fprintf(stderr, "libfctpdeco FCT_GUARD: live buffer %p (%zu bytes) has %lld damaged canary bytes\n", kv.first,
                    kv.second, bad);
            (*corrupted)++;
        }
        (*live)++;
    }
    *corrupted += g_corrupt_at_free;
    g_corrupt_at_free = 0;
    return 0;
}
