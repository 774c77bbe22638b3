// host_util.hpp -- host-side plumbing shared by the translation units of libsigner_gibbs.so:
// error reporting, stream-ordered device allocations, a cache of pinned staging buffers and the
// transfer pool that moves pageable host memory to and from the device.
//
// Why a transfer pool: the arrays R (or NumPy) hands over are pageable.  A cudaMemcpy straight from /
// into them stages through the driver and blocks the calling thread, which starves the GPU of sweeps.
// Here worker threads own pinned staging buffers: device-to-host pieces are copied out at PCIe speed
// and then moved into the caller's array; host-to-device pieces are gathered into the pinned buffer,
// sent, and (optionally) consumed by a kernel on the worker's stream.  Several pieces are in flight.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/signer_gibbs.h"

namespace sgh {

extern thread_local std::string g_err;

inline int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return sgh::fail(e_ == cudaErrorMemoryAllocation ? SIGNER_ERR_NOMEM : SIGNER_ERR_CUDA,    \
                             std::string(#call) + ": " + cudaGetErrorString(e_));                     \
    } while (0)

#define ST(call)                                  \
    do {                                          \
        int s_ = (call);                          \
        if (s_ != SIGNER_OK) return s_;           \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Stream-ordered device memory from the device's default pool.  The pool's release threshold is
// raised to "never" once per device (ensure_pool), so a stateless call's ~40 allocations and frees
// are served from memory the previous call gave back -- no cudaMalloc / cudaFree (a device-wide
// synchronisation each) on the hot path of repeated calls.
// ------------------------------------------------------------------------------------------------
int ensure_pool(int device);

inline int dalloc(void **p, size_t bytes, cudaStream_t st)
{
    *p = nullptr;
    CU(cudaMallocAsync(p, bytes ? bytes : 8, st));
    return SIGNER_OK;
}
template <typename T>
inline int dalloc(T **p, size_t count, cudaStream_t st)
{
    return dalloc(reinterpret_cast<void **>(p), count * sizeof(T), st);
}
inline void dfree(void *p, cudaStream_t st)
{
    if (p) cudaFreeAsync(p, st);
}

// a temporary device allocation that is given back when the scope ends, also on error returns
struct DevTemp {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    DevTemp() = default;
    DevTemp(const DevTemp &) = delete;
    DevTemp &operator=(const DevTemp &) = delete;
    ~DevTemp() { dfree(p, st); }
    int get(size_t bytes, cudaStream_t s) { st = s; return dalloc(&p, bytes, s); }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------------
// Pinned host buffers are expensive to create (cudaHostAlloc pins pages): keep the ones a call
// returns for the next call.  Buffers are handed out whole; the cache holds at most kMaxCached bytes.
// ------------------------------------------------------------------------------------------------
class PinnedCache {
public:
    static constexpr size_t kMaxCached = size_t(1) << 30;
    static void *get(size_t bytes, size_t *got);
    static void put(void *p, size_t bytes);
    static void clear();
};

struct PinnedBuf {
    void *p = nullptr;
    size_t bytes = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { if (p) PinnedCache::put(p, bytes); }
    bool get(size_t n) { if (p && bytes >= n) return true; if (p) PinnedCache::put(p, bytes); p = PinnedCache::get(n, &bytes); return p != nullptr; }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------------
// Transfer pool: one per device, started on first use, never torn down.
// ------------------------------------------------------------------------------------------------
// A group of jobs the poster waits for; failures are reported per group (a failed copy of one call
// does not poison later calls).
struct TransferGroup {
    std::atomic<int> pending{0};
    std::atomic<bool> failed{false};
    std::string err;            // written under the pool's mutex
};
using GroupPtr = std::shared_ptr<TransferGroup>;
inline GroupPtr make_group() { return std::make_shared<TransferGroup>(); }

class TransferPool {
public:
    static constexpr size_t kStageBytes = size_t(32) << 20;
    static constexpr int kWorkers = 6;
    // what a host-to-device job runs on the worker's stream once its piece is in device memory
    // (dev = the worker's device scratch holding the piece); null: the piece is copied to `dst`
    using Hook = std::function<cudaError_t(const void *dev, size_t bytes, cudaStream_t st)>;
    struct Job {
        int kind;                // 0 device -> pageable host, 1 pageable host -> device
        const char *src; char *dst; size_t bytes;
        cudaEvent_t ready;       // waited for before the job touches device memory (may be null)
        GroupPtr group;
        Hook hook;
    };

    static TransferPool &of(int device);
    // copy [src, src+bytes) (device) to dst (pageable host) once `ready` has happened, in pieces
    void post_d2h(const void *src, void *dst, size_t bytes, cudaEvent_t ready, const GroupPtr &g, size_t piece = size_t(8) << 20);
    // copy [src, src+bytes) (pageable host) to dst (device), in pieces
    void post_h2d(const void *src, void *dst, size_t bytes, cudaEvent_t ready, const GroupPtr &g, size_t piece = size_t(4) << 20);
    // one piece (at most kStageBytes) of pageable host memory staged into device scratch and handed to `hook`
    void post_h2d_hook(const void *src, size_t bytes, cudaEvent_t ready, const GroupPtr &g, Hook hook);
    // block until every job of the group is done; false if one of them failed (text in g->err)
    bool wait(const GroupPtr &g);

private:
    explicit TransferPool(int device);
    void run();
    void push(Job &&j);
    int device_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
};

// Touch every page of a caller-allocated output range so that later copies land on resident pages.  The touch is an
// atomic no-op read-modify-write: a copy-out that has already landed on the page is left intact whatever the order.
void prefault(void *p, size_t bytes);

// 128-bit content hash of a host array (cohort cache key), multi-threaded for large arrays
void hash128(const void *p, size_t bytes, uint64_t out[2]);
// the same for two buffers of equal length in one fan-out of threads (M and W of a cohort)
void hash128_two(const void *a, const void *b, size_t bytes, uint64_t ha[2], uint64_t hb[2]);

} // namespace sgh
