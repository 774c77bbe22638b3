// host_util.cu -- see host_util.hpp.  No kernels here; compiled by nvcc only for the CUDA runtime headers.
#include "host_util.hpp"

#include <sys/mman.h>

#include <algorithm>
#include <cstdlib>

namespace sgh {

thread_local std::string g_err;

int ensure_pool(int device)
{
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) return fail(SIGNER_ERR_ARG, "device ordinal out of range");
    if (done[device]) return SIGNER_OK;
    int supported = 0;
    CU(cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device));
    if (!supported) return fail(SIGNER_ERR_CUDA, "device has no stream-ordered memory pools");
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    if (const char *e = std::getenv("SIGNER_POOL_KEEP_BYTES")) keep = std::strtoull(e, nullptr, 10);
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done[device] = true;
    return SIGNER_OK;
}

// ------------------------------------------------------------------------------------------------ pinned cache
namespace {
std::mutex g_pin_mu;
std::vector<std::pair<void *, size_t>> g_pin_free;
size_t g_pin_cached = 0;
}

void *PinnedCache::get(size_t bytes, size_t *got)
{
    bytes = std::max<size_t>(bytes, 4096);
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        int best = -1;
        for (int i = 0; i < (int)g_pin_free.size(); ++i)
            if (g_pin_free[i].second >= bytes && g_pin_free[i].second <= 4 * bytes && (best < 0 || g_pin_free[i].second < g_pin_free[best].second)) best = i;
        if (best >= 0) {
            void *p = g_pin_free[best].first;
            *got = g_pin_free[best].second;
            g_pin_cached -= *got;
            g_pin_free.erase(g_pin_free.begin() + best);
            return p;
        }
    }
    void *p = nullptr;
    const size_t rounded = (bytes + 65535) / 65536 * 65536;
    if (cudaHostAlloc(&p, rounded, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *got = rounded;
    return p;
}

void PinnedCache::put(void *p, size_t bytes)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (g_pin_cached + bytes <= kMaxCached) {
            g_pin_free.emplace_back(p, bytes);
            g_pin_cached += bytes;
            return;
        }
    }
    cudaFreeHost(p);
}

void PinnedCache::clear()
{
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (auto &b : g_pin_free) cudaFreeHost(b.first);
    g_pin_free.clear();
    g_pin_cached = 0;
}

// ------------------------------------------------------------------------------------------------ transfer pool
TransferPool &TransferPool::of(int device)
{
    static std::mutex mu;
    static TransferPool *pools[64] = {nullptr};
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[device]) pools[device] = new TransferPool(device);
    return *pools[device];
}

TransferPool::TransferPool(int device) : device_(device)
{
    for (int w = 0; w < kWorkers; ++w) std::thread([this] { run(); }).detach();
}

void TransferPool::push(Job &&j)
{
    j.group->pending.fetch_add(1);
    q_.push_back(std::move(j));
}

void TransferPool::post_d2h(const void *src, void *dst, size_t bytes, cudaEvent_t ready, const GroupPtr &g, size_t piece)
{
    piece = std::max<size_t>(4096, std::min(piece, kStageBytes));
    std::lock_guard<std::mutex> lk(mu_);
    for (size_t off = 0; off < bytes; off += piece)
        push(Job{0, static_cast<const char *>(src) + off, static_cast<char *>(dst) + off, std::min(piece, bytes - off), ready, g, nullptr});
    cv_.notify_all();
}

void TransferPool::post_h2d(const void *src, void *dst, size_t bytes, cudaEvent_t ready, const GroupPtr &g, size_t piece)
{
    piece = std::max<size_t>(4096, std::min(piece, kStageBytes));
    std::lock_guard<std::mutex> lk(mu_);
    for (size_t off = 0; off < bytes; off += piece)
        push(Job{1, static_cast<const char *>(src) + off, static_cast<char *>(dst) + off, std::min(piece, bytes - off), ready, g, nullptr});
    cv_.notify_all();
}

void TransferPool::post_h2d_hook(const void *src, size_t bytes, cudaEvent_t ready, const GroupPtr &g, Hook hook)
{
    std::lock_guard<std::mutex> lk(mu_);
    push(Job{1, static_cast<const char *>(src), nullptr, std::min(bytes, kStageBytes), ready, g, std::move(hook)});
    cv_.notify_all();
}

bool TransferPool::wait(const GroupPtr &g)
{
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return g->pending.load() == 0; });
    return !g->failed.load();
}

void TransferPool::run()
{
    void *stage = nullptr, *scratch = nullptr;
    cudaStream_t st = nullptr;
    cudaError_t init = cudaSetDevice(device_);
    if (init == cudaSuccess) init = cudaHostAlloc(&stage, kStageBytes, cudaHostAllocDefault);
    if (init == cudaSuccess) init = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return !q_.empty(); });
            j = std::move(q_.front()); q_.pop_front();
        }
        cudaError_t e = init;
        if (e == cudaSuccess && j.ready) e = cudaEventSynchronize(j.ready);
        if (e == cudaSuccess && j.kind == 0) {
            e = cudaMemcpyAsync(stage, j.src, j.bytes, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) std::memcpy(j.dst, stage, j.bytes);
        } else if (e == cudaSuccess) {
            std::memcpy(stage, j.src, j.bytes);
            if (j.hook) {
                if (!scratch) e = cudaMalloc(&scratch, kStageBytes);      // device scratch, created on first use, kept
                if (e == cudaSuccess) e = cudaMemcpyAsync(scratch, stage, j.bytes, cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = j.hook(scratch, j.bytes, st);
            } else {
                e = cudaMemcpyAsync(j.dst, stage, j.bytes, cudaMemcpyHostToDevice, st);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (e != cudaSuccess) {
                if (!j.group->failed.exchange(true)) j.group->err = cudaGetErrorString(e);
                cudaGetLastError();
            }
            j.group->pending.fetch_sub(1);
        }
        done_.notify_all();
    }
}

// ------------------------------------------------------------------------------------------------ prefault
void prefault(void *p, size_t bytes)
{
    if (!p || !bytes) return;
    static const bool no_hugepage = [] { const char *e = std::getenv("SIGNER_KEEP_HUGEPAGES"); return !(e && e[0] == '1'); }();
    char *b = static_cast<char *>(p);
    if (no_hugepage) {
        // A caller that asked for transparent huge pages (NumPy does for large arrays) makes every first touch wait
        // for a free 2 MB block -- direct compaction inside the fault, measured as sporadic stalls of up to a second
        // per call.  These pages are written once and read once: plain 4 KB faults, spread over the helper threads,
        // are faster and predictable.  (A hint only; R's vectors are not huge-page backed in the first place.)
        const uintptr_t a0 = (reinterpret_cast<uintptr_t>(p) + 4095) & ~uintptr_t(4095);
        const uintptr_t a1 = (reinterpret_cast<uintptr_t>(p) + bytes) & ~uintptr_t(4095);
        if (a1 > a0) madvise(reinterpret_cast<void *>(a0), a1 - a0, MADV_NOHUGEPAGE);
    }
    // x |= 0 as one atomic read-modify-write per page: faults the page in for writing and can never undo a store
    // another thread (a copy-out worker) has already made to that byte
    for (size_t off = 0; off < bytes; off += 4096) __atomic_fetch_or(b + off, (char)0, __ATOMIC_RELAXED);
    __atomic_fetch_or(b + bytes - 1, (char)0, __ATOMIC_RELAXED);
}

// ------------------------------------------------------------------------------------------------ hash
namespace {
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// four independent multiply-rotate lanes over 64-bit words (throughput-bound, ~10 GB/s per core)
void hash_chunk(const unsigned char *p, size_t bytes, uint64_t seed, uint64_t out[2])
{
    uint64_t a = seed ^ 0x9e3779b97f4a7c15ULL, b = seed ^ 0xc2b2ae3d27d4eb4fULL, c = seed ^ 0x165667b19e3779f9ULL, d = seed ^ 0x27d4eb2f165667c5ULL;
    const size_t n32 = bytes / 32;
    const uint64_t *w = reinterpret_cast<const uint64_t *>(p);
    for (size_t i = 0; i < n32; ++i) {
        uint64_t x0, x1, x2, x3;
        std::memcpy(&x0, w + 4 * i, 8); std::memcpy(&x1, w + 4 * i + 1, 8); std::memcpy(&x2, w + 4 * i + 2, 8); std::memcpy(&x3, w + 4 * i + 3, 8);
        a = rotl64(a ^ x0, 29) * 0x9fb21c651e98df25ULL; b = rotl64(b ^ x1, 31) * 0xa0761d6478bd642fULL;
        c = rotl64(c ^ x2, 27) * 0xe7037ed1a0b428dbULL; d = rotl64(d ^ x3, 33) * 0x8ebc6af09c88c6e3ULL;
    }
    uint64_t tail = bytes;
    for (size_t i = n32 * 32; i < bytes; ++i) tail = tail * 0x100000001b3ULL ^ p[i];
    out[0] = mix64(a ^ rotl64(b, 17) ^ mix64(c + tail));
    out[1] = mix64(d ^ rotl64(c, 23) ^ mix64(b + a + tail * 3));
}
}

void hash128_two(const void *pa, const void *pb, size_t bytes, uint64_t ha[2], uint64_t hb[2])
{
    const unsigned char *b[2] = {static_cast<const unsigned char *>(pa), static_cast<const unsigned char *>(pb)};
    const int nbuf = pb ? 2 : 1;
    const size_t kChunk = size_t(512) << 10;             // small enough that 2 x 7.7 MB keep eight threads busy
    const size_t nchunks = std::max<size_t>(1, (bytes + kChunk - 1) / kChunk);
    std::vector<uint64_t> h(2 * nchunks * nbuf);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nth = std::min<size_t>({nchunks * nbuf, (size_t)8, (size_t)hw});
    auto work = [&](size_t t) {
        for (size_t c = t; c < nchunks * nbuf; c += nth) {
            const size_t which = c / nchunks, cc = c - which * nchunks, off = cc * kChunk;
            hash_chunk(b[which] + off, std::min(kChunk, bytes - std::min(bytes, off)), (uint64_t)cc, &h[2 * c]);
        }
    };
    if (nth <= 1) work(0);
    else {
        std::vector<std::thread> th;
        for (size_t t = 1; t < nth; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
    }
    hash_chunk(reinterpret_cast<const unsigned char *>(h.data()), nchunks * 16, (uint64_t)bytes, ha);
    if (pb) hash_chunk(reinterpret_cast<const unsigned char *>(h.data() + 2 * nchunks), nchunks * 16, (uint64_t)bytes, hb);
}

void hash128(const void *p, size_t bytes, uint64_t out[2]) { hash128_two(p, nullptr, bytes, out, nullptr); }

} // namespace sgh
