// signer_gibbs.cu -- host side of libsigner_gibbs.so: device-resident chains, the sweep scheduler
// and the C ABI declared in include/signer_gibbs.h.
//
// Replaces the driver GibbsSamplerCpp (src/gibbs_2.cpp:253-424): flag logic (:265-267), the
// random scan (:303-318), sample storage (:320-334) and the returned list (:406-423).
// There is no CPU path in this file: every entry point that computes launches CUDA kernels.
#include "../../include/signer_gibbs.h"
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cstring>
#include <dlfcn.h>
#include <sys/mman.h>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace sg;

namespace {

thread_local std::string g_err;
thread_local double g_loop_ms = 0.0;      // wall time of the last sweep loop on this thread (sharded / stateless runs)

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? SIGNER_ERR_NOMEM : SIGNER_ERR_CUDA,         \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                          \
    } while (0)

#define ST(call)                                  \
    do {                                          \
        int s_ = (call);                          \
        if (s_ != SIGNER_OK) return s_;           \
    } while (0)

constexpr size_t kRingBytes = size_t(256) << 20;   // per ring set (two sets): small enough to pipeline D2H behind compute

// The data of one cohort on one device, shared by every chain on that device.
struct DeviceData {
    int device = 0, I = 0, J = 0, Jp = 0, n2 = 0;
    int col0 = 0, Jglobal = 0;  // column shard: first global genome, genomes in the whole cohort
    // z_alloc's work list: the non-zero cells sorted by count (largest first, ties genome-major)
    unsigned short *cell_row = nullptr;
    int *cell_col = nullptr, *cell_n = nullptr;
    int n_cells = 0;
    int *Mrow = nullptr;       // [I][Jp], original genome order (log-likelihood)
    double *W = nullptr;       // [I][Jp]
    double *lgm = nullptr;     // sum lgamma(M+1), one double
    ~DeviceData()
    {
        cudaSetDevice(device);
        cudaFree(cell_row); cudaFree(cell_col); cudaFree(cell_n); cudaFree(Mrow); cudaFree(W); cudaFree(lgm);
    }
};

// cells per batch of z_alloc_fused's work list for cells of count n (see upload_data); SIGNER_Z_WIDTHS="n:w,n:w,..."
// (descending n) overrides the built-in grading for experiments
const bool g_no_hugepage = [] { const char *e = std::getenv("SIGNER_KEEP_HUGEPAGES"); return !(e && e[0] == '1'); }();
const bool g_sort_rounds = [] { const char *e = std::getenv("SIGNER_Z_SORT_EXACT"); return !(e && e[0] == '1'); }();

int batch_width(int n)
{
    static const std::vector<std::pair<int, int>> grade = [] {
        std::vector<std::pair<int, int>> g;
        if (const char *e = std::getenv("SIGNER_Z_WIDTHS")) {
            g.clear();
            int a = 0, b = 0, used = 0;
            while (std::sscanf(e, "%d:%d%n", &a, &b, &used) == 2) {
                g.emplace_back(a, std::max(1, std::min(32, b)));
                e += used;
                if (*e == ',') ++e;
            }
        }
        return g;
    }();
    for (const auto &g : grade) if (n >= g.first) return g.second;
    return 32;
}

int upload_data(int device, int I, int J, const double *M, const double *W, std::shared_ptr<DeviceData> &out, int col0 = 0,
                int Jglobal = -1)
{
    if (Jglobal < 0) Jglobal = J;
    if (I <= 0 || J <= 0 || !M || !W) return fail(SIGNER_ERR_ARG, "problem: I, J must be positive and M, W non-null");
    if ((long long)I * J >= (1LL << 31)) return fail(SIGNER_ERR_ARG, "problem: I*J must be below 2^31");
    CU(cudaSetDevice(device));
    auto d = std::make_shared<DeviceData>();
    d->device = device; d->I = I; d->J = J; d->Jp = (J + 31) / 32 * 32;
    d->col0 = col0; d->Jglobal = Jglobal;
    d->n2 = 1; while (d->n2 < Jglobal) d->n2 <<= 1;
    if (I > 65535) return fail(SIGNER_ERR_ARG, "problem: I must be below 65536");
    const size_t n = (size_t)I * d->Jp;
    std::vector<int> m(n, 0);
    std::vector<double> w(n, 0.0);
    std::vector<int> cells;                  // element ids i + I*j of the non-zero cells, genome-major
    for (int j = 0; j < J; ++j)
        for (int i = 0; i < I; ++i) {
            const double mv = M[(size_t)i + (size_t)I * j], wv = W[(size_t)i + (size_t)I * j];
            if (!(mv >= 0.0) || !(mv <= 2147483647.0)) return fail(SIGNER_ERR_NONFINITE, "M has a negative, non-finite or too large entry");
            if (!(wv >= 0.0) || !(wv <= 1.7976931348623157e308)) return fail(SIGNER_ERR_NONFINITE, "W has a negative or non-finite entry");
            const int mi = (int)mv;                      // int size (src/gibbs_2.cpp:31)
            m[(size_t)i * d->Jp + j] = mi;
            w[(size_t)i * d->Jp + j] = wv;
            if (mi > 0) cells.push_back(i + I * j);
        }
    // z_alloc_fused's work list (a private order: results do not depend on it): the non-zero cells sorted by
    // count, largest first (ties keep the genome-major order), cut into batches of 32 list entries = one warp
    // round each.  batch_width() can give the heaviest cells narrower batches (the rest of the batch is padding,
    // count 0); measured on B200 every such grading lost to plain 32-cell batches, so the default is none.
    std::vector<unsigned short> crow;
    std::vector<int> ccol, cn;
    {
        const size_t nz = cells.size();
        std::vector<int> cnt(nz);
        for (size_t t = 0; t < nz; ++t) {
            const int i = cells[t] % I, j = cells[t] / I;
            cnt[t] = m[(size_t)i * d->Jp + j];
        }
        // sort key: counts above kNDirect by value; smaller ones by their number of direct-method rounds (four
        // draws per Philox block), so that cells that cost the same stay in genome-major order and a warp's lanes
        // read few distinct columns of E.  Counting sort (stable) when the keys are small enough.
        auto key = [&](size_t t) { return cnt[t] > kNDirect ? cnt[t] : (g_sort_rounds ? 4 * ((cnt[t] + 3) / 4) : cnt[t]); };
        std::vector<size_t> idx(nz);
        int kmax = 0;
        for (size_t t = 0; t < nz; ++t) kmax = std::max(kmax, key(t));
        if (kmax <= (1 << 22)) {
            std::vector<size_t> pos((size_t)kmax + 2, 0);
            for (size_t t = 0; t < nz; ++t) pos[(size_t)key(t)]++;
            size_t run = 0;
            for (int v = kmax; v >= 0; --v) { const size_t h = pos[(size_t)v]; pos[(size_t)v] = run; run += h; }
            for (size_t t = 0; t < nz; ++t) idx[pos[(size_t)key(t)]++] = t;
        } else {
            for (size_t t = 0; t < nz; ++t) idx[t] = t;
            std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return key(a) > key(b); });
        }
        crow.reserve(nz + 4096); ccol.reserve(nz + 4096); cn.reserve(nz + 4096);
        size_t t = 0;
        while (t < nz) {
            const int width = batch_width(cnt[idx[t]]);
            int filled = 0;
            for (; filled < width && t < nz; ++filled, ++t) {
                const size_t e = idx[t];
                crow.push_back((unsigned short)(cells[e] % I)); ccol.push_back(cells[e] / I); cn.push_back(cnt[e]);
            }
            if (t < nz) for (; filled < 32; ++filled) { crow.push_back(0); ccol.push_back(0); cn.push_back(0); }
        }
    }
    d->n_cells = (int)cn.size();
    if (cn.empty()) { crow.push_back(0); ccol.push_back(0); cn.push_back(0); }
    const size_t nc = cn.size();
    CU(cudaMalloc(&d->cell_row, nc * sizeof(unsigned short)));
    CU(cudaMalloc(&d->cell_col, nc * sizeof(int)));
    CU(cudaMalloc(&d->cell_n, nc * sizeof(int)));
    CU(cudaMalloc(&d->Mrow, n * sizeof(int)));
    CU(cudaMalloc(&d->W, n * sizeof(double)));
    CU(cudaMalloc(&d->lgm, sizeof(double)));
    CU(cudaMemcpy(d->cell_row, crow.data(), nc * sizeof(unsigned short), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d->cell_col, ccol.data(), nc * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d->cell_n, cn.data(), nc * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d->Mrow, m.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d->W, w.data(), n * sizeof(double), cudaMemcpyHostToDevice));
    // constant of the data: sum lgamma(M+1) (lgM in R/cpp_eBayesNMF.R:180); a column shard computes it with its peers
    if (Jglobal == J) {
        double *col = nullptr;
        CU(cudaMalloc(&col, (size_t)d->n2 * sizeof(double)));
        CU(cudaMemset(col, 0, (size_t)d->n2 * sizeof(double)));
        LLParams lp{I, J, d->Jp, 1, d->Mrow, d->W, nullptr, nullptr, col, 1, nullptr, 0, nullptr, nullptr, nullptr};
        loglik_cols<<<(J + 31) / 32, 32 * kLLWarps, loglik_smem_bytes(1)>>>(lp);
        tree_sum<<<1, 1024>>>(col, d->n2, nullptr, d->lgm);
        CU(cudaGetLastError());
        CU(cudaDeviceSynchronize());
        CU(cudaFree(col));
    }
    out = d;
    return SIGNER_OK;
}

// z_alloc_fused may need more than 48 KB of dynamic shared memory; raise the limit once per device to the
// architectural maximum so that concurrent chains with different K never race on the attribute.
int g_z_dyn_smem_max[64] = {0};

int ensure_device_setup(int device)
{
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) return fail(SIGNER_ERR_ARG, "device ordinal out of range");
    if (done[device]) return SIGNER_OK;
    CU(cudaSetDevice(device));
    int optin = 0;
    CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, z_alloc_fused));
    g_z_dyn_smem_max[device] = optin - (int)fa.sharedSizeBytes;
    CU(cudaFuncSetAttribute(z_alloc_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, g_z_dyn_smem_max[device]));
#ifdef SG_Z_CARVEOUT
    CU(cudaFuncSetAttribute(z_alloc_fused, cudaFuncAttributePreferredSharedMemoryCarveout, SG_Z_CARVEOUT));
#endif
    done[device] = true;
    return SIGNER_OK;
}

struct Ring {
    double *Ps = nullptr, *Es = nullptr, *Aps = nullptr, *Aes = nullptr, *Bps = nullptr, *Bes = nullptr;
    cudaEvent_t filled = nullptr;
    std::shared_ptr<std::atomic<int>> inflight = std::make_shared<std::atomic<int>>(0);   // copy-out jobs not yet done
};

// ------------------------------------------------------------------------------------------------
// Copy-out pool.  The caller's output arrays are pageable (R vectors, NumPy arrays); a device-to-host copy
// straight into them stages through the driver and blocks the enqueuing thread, which starves the GPU of
// sweeps.  Instead the samples leave in pieces of at most kStageBytes: a worker thread waits for the ring
// set to be filled, copies a piece into its own pinned buffer at PCIe speed and then moves it into the
// caller's array, several pieces in flight.  One pool per device, started on first use, never torn down.
// ------------------------------------------------------------------------------------------------
class CopyPool {
public:
    static constexpr size_t kStageBytes = size_t(32) << 20;
    static constexpr int kWorkers = 4;
    struct Job { const char *src; char *dst; size_t bytes; cudaEvent_t ready; std::shared_ptr<std::atomic<int>> pending; };

    static CopyPool &of(int device)
    {
        static std::mutex mu;
        static CopyPool *pools[64] = {nullptr};
        std::lock_guard<std::mutex> lk(mu);
        if (!pools[device]) pools[device] = new CopyPool(device);
        return *pools[device];
    }
    // copy [src, src+bytes) (device) to dst (pageable host) once `ready` has happened
    void post(const void *src, void *dst, size_t bytes, cudaEvent_t ready, const std::shared_ptr<std::atomic<int>> &pending)
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t off = 0; off < bytes; off += kStageBytes) {
            pending->fetch_add(1);
            q_.push_back(Job{static_cast<const char *>(src) + off, static_cast<char *>(dst) + off, std::min(kStageBytes, bytes - off), ready, pending});
        }
        cv_.notify_all();
    }
    // block until every job counted by `pending` is done; false if a worker met a CUDA error
    bool wait(const std::shared_ptr<std::atomic<int>> &pending)
    {
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending->load() == 0; });
        return !failed_;
    }
    const char *error() const { return err_; }

private:
    explicit CopyPool(int device) : device_(device)
    {
        for (int w = 0; w < kWorkers; ++w) std::thread([this] { run(); }).detach();
    }
    void run()
    {
        void *stage = nullptr;
        cudaStream_t st = nullptr;
        bool ok = cudaSetDevice(device_) == cudaSuccess && cudaHostAlloc(&stage, kStageBytes, cudaHostAllocDefault) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return !q_.empty(); });
                j = q_.front(); q_.pop_front();
            }
            cudaError_t e = ok ? cudaEventSynchronize(j.ready) : cudaErrorUnknown;
            if (e == cudaSuccess) e = cudaMemcpyAsync(stage, j.src, j.bytes, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e == cudaSuccess) std::memcpy(j.dst, stage, j.bytes);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (e != cudaSuccess) { failed_ = true; err_ = cudaGetErrorString(e); }
                j.pending->fetch_sub(1);
            }
            done_.notify_all();
        }
    }
    int device_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<Job> q_;
    bool failed_ = false;
    const char *err_ = "";
};

} // namespace

struct signer_chain {
    std::shared_ptr<DeviceData> data;
    int device = 0, I = 0, J = 0, Jp = 0, K = 0;
    size_t nP = 0, nE = 0;
    double *P = nullptr, *E = nullptr, *Ap = nullptr, *Bp = nullptr, *Ae = nullptr, *Be = nullptr;
    int *Zin[2] = {nullptr, nullptr}, *Znj[2] = {nullptr, nullptr};
    int zcur = 0;
    double *corrE_part = nullptr;
    int n_seg = 0;
    int shard = 0, n_shards = 1, col0 = 0;   // column sharding (signer_gibbs_run_sharded)
    double *corrE_all = nullptr;             // [n_shards][I*K]: every shard's ordered W E' sum, after the all-reduce
    double *col = nullptr, *ll_ring = nullptr;
    int ll_cap = 0;
    double *Zcube = nullptr;
    double *stats = nullptr;
    double *sumPs = nullptr, *sumEs = nullptr;   // full-length sample stores for the device-side summaries
    unsigned long long *draw_count = nullptr;    // [2] device counters of step-1 draws (allocated on request)
    double *acache_p = nullptr, *acache_e = nullptr;   // [3][I*K], [3][K*J]: A-only terms of the MH ratios (kernels.cuh, ACache)
    unsigned long long *tile_counter = nullptr;
    unsigned *ll_ticket = nullptr;               // block counter of loglik_cols' fused tree
    unsigned long long z_launches = 0;
    int z_grid = 0, z_warps = 8, n_batches = 0, n_chunks = 0, z_warp_smem = 0, z_zin_smem = 0;
    int ksplit = 1;
    Hyper h{};
    Key key{};
    uint32_t sweep0 = 0;
    int64_t sweeps_done = 0;
    int64_t launches = 0;
    int Pfixed = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // the log-likelihood of an evaluation sweep only reads P and E: it runs on its own stream next to the following
    // sweep's z_alloc_fused (which reads them too); the next kernel that writes P or E waits for it
    cudaStream_t ll_stream = nullptr;
    cudaEvent_t ll_ready = nullptr, ll_done = nullptr;
    bool ll_pending = false;
    Ring ring[2];
    int ring_cap = 0, ring_keep = 0;
    int last_eval = 0;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[4];

    ~signer_chain()
    {
        cudaSetDevice(device);
        if (own_stream) cudaStreamSynchronize(own_stream);
        if (ll_stream) cudaStreamSynchronize(ll_stream);
        for (double *p : {P, E, Ap, Bp, Ae, Be, corrE_part, corrE_all, col, ll_ring, Zcube, stats, sumPs, sumEs, acache_p, acache_e}) cudaFree(p);
        for (int b = 0; b < 2; ++b) {
            cudaFree(Zin[b]); cudaFree(Znj[b]);
            free_ring(ring[b]);
        }
        cudaFree(tile_counter); cudaFree(draw_count); cudaFree(ll_ticket);
        for (auto &v : ev) for (auto &pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        if (own_stream) cudaStreamDestroy(own_stream);
        if (ll_stream) cudaStreamDestroy(ll_stream);
        if (ll_ready) cudaEventDestroy(ll_ready);
        if (ll_done) cudaEventDestroy(ll_done);
    }
    static void free_ring(Ring &r)
    {
        for (double *p : {r.Ps, r.Es, r.Aps, r.Aes, r.Bps, r.Bes}) cudaFree(p);
        if (r.filled) cudaEventDestroy(r.filled);
        r = Ring{};
    }
};

namespace {

struct ProfScope {
    signer_chain *c; int cls; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(signer_chain *c_, int cls_, cudaStream_t st_ = nullptr) : c(c_), cls(cls_), st(st_ ? st_ : c_->stream)
    {
        if (!c->profiling) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
    ~ProfScope()
    {
        if (!c->profiling) return;
        cudaEventRecord(b, st);
        c->ev[cls].emplace_back(a, b);
    }
};

// a kernel that writes P or E is about to be enqueued: the pending log-likelihood must have read them
int wait_loglik(signer_chain *c)
{
    if (!c->ll_pending) return SIGNER_OK;
    CU(cudaStreamWaitEvent(c->stream, c->ll_done, 0));
    c->ll_pending = false;
    return SIGNER_OK;
}

void host_perm(uint64_t seed, uint32_t sweep, int order[7])
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const U4 a = philox4x32_10(0, 0, kStPerm, sweep, k0, k1), b = philox4x32_10(1, 0, kStPerm, sweep, k0, k1);
    const uint32_t w[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
    for (int i = 0; i < 7; ++i) order[i] = i + 1;
    for (int idx = 6; idx >= 1; --idx) {
        const uint32_t j = (uint32_t)(((uint64_t)w[6 - idx] * (uint64_t)(idx + 1)) >> 32);
        std::swap(order[idx], order[j]);
    }
}

int launch_z(signer_chain *c, uint32_t sweep, uint32_t stream_id, double *zcube)
{
    const int tgt = c->zcur ^ 1;
    ZParams p{};
    p.I = c->I; p.J = c->J; p.K = c->K;
    p.col0 = c->col0; p.Jglobal = c->data->Jglobal;
    p.cell_row = c->data->cell_row; p.cell_col = c->data->cell_col; p.cell_n = c->data->cell_n;
    p.n_cells = c->data->n_cells; p.n_batches = c->n_batches;
    p.P = c->P; p.E = c->E;
    p.Zin = c->Zin[tgt]; p.Znj = c->Znj[tgt];
    p.Zin_next = c->Zin[c->zcur]; p.Znj_next = c->Znj[c->zcur];
    p.Zcube = zcube;
    p.counter = c->tile_counter;
    p.base = c->z_launches * (unsigned long long)(c->n_chunks + c->z_grid * c->z_warps);   // every warp overshoots once
    p.warp_smem = c->z_warp_smem; p.zin_smem = c->z_zin_smem;
    p.draw_count = c->draw_count;
    p.key = c->key; p.sweep = sweep; p.stream = stream_id;
    if (zcube) CU(cudaMemsetAsync(zcube, 0, sizeof(double) * (size_t)c->I * c->J * c->K, c->stream));
    {
        ProfScope ps(c, 0);
        z_alloc_fused<<<c->z_grid, 32 * c->z_warps, (size_t)c->z_warp_smem * c->z_warps + (c->z_zin_smem ? c->nP * sizeof(int) : 0), c->stream>>>(p);
    }
    CU(cudaGetLastError());
    c->z_launches++;
    c->launches++;
    c->zcur = tgt;
    return SIGNER_OK;
}

int launch_p(signer_chain *c, uint32_t sweep, const Ops &ops, double *Ps, double *Aps, double *Bps)
{
    PParams p{};
    p.IK = (int)c->nP; p.P = c->P; p.Ap = c->Ap; p.Bp = c->Bp;
    p.Zin = c->Zin[c->zcur];
    if (c->n_shards > 1) { p.corrE_part = c->corrE_all; p.n_seg = c->n_shards; p.n_shards = c->n_shards; }
    else { p.corrE_part = c->corrE_part; p.n_seg = c->n_seg; p.n_shards = 1; }
    p.Ps_slot = Ps; p.Aps_slot = Aps; p.Bps_slot = Bps;
    p.acache = ACache{c->acache_p, c->acache_p + c->nP, c->acache_p + 2 * c->nP};
    p.h = c->h; p.key = c->key; p.sweep = sweep; p.ops = ops;
    ST(wait_loglik(c));
    {
        ProfScope ps(c, 1);
        p_family<<<(p.IK + 127) / 128, 128, 0, c->stream>>>(p);
    }
    c->launches++;
    CU(cudaGetLastError());
    return SIGNER_OK;
}

int launch_e(signer_chain *c, uint32_t sweep, const Ops &ops, int do_corrE, double *Es, double *Aes, double *Bes)
{
    EParams p{};
    p.I = c->I; p.J = c->J; p.Jp = c->Jp; p.K = c->K;
    p.col0 = c->col0;
    p.W = c->data->W; p.P = c->P; p.E = c->E; p.Ae = c->Ae; p.Be = c->Be;
    p.Znj = c->Znj[c->zcur]; p.corrE_part = c->corrE_part;
    p.do_corrE = do_corrE;
    p.Es_slot = Es; p.Aes_slot = Aes; p.Bes_slot = Bes;
    p.acache = ACache{c->acache_e, c->acache_e + c->nE, c->acache_e + 2 * c->nE};
    p.h = c->h; p.key = c->key; p.sweep = sweep; p.ops = ops;
    ST(wait_loglik(c));
    {
        ProfScope ps(c, 2);
        e_family<<<dim3(c->n_seg, c->ksplit), kEThreads, 0, c->stream>>>(p);
    }
    c->launches++;
    CU(cudaGetLastError());
    return SIGNER_OK;
}

// per-genome log-likelihood terms; with `out` the last block also runs the tree (unsharded chains)
int launch_loglik_cols(signer_chain *c, double *out = nullptr, cudaStream_t st = nullptr)
{
    LLParams lp{c->I, c->J, c->Jp, c->K, c->data->Mrow, c->data->W, c->P, c->E, c->col + c->col0, 0,
                out ? c->col : nullptr, c->data->n2, c->data->lgm, out, c->ll_ticket};
    loglik_cols<<<(c->J + 31) / 32, 32 * kLLWarps, loglik_smem_bytes(c->K), st ? st : c->stream>>>(lp);
    c->launches++;
    CU(cudaGetLastError());
    return SIGNER_OK;
}

// log-likelihood of the current state into out[0], on the side stream (see signer_chain::ll_stream)
int launch_loglik(signer_chain *c, double *out)
{
    ST(wait_loglik(c));                                    // one in flight at a time: they share col and the ticket
    CU(cudaEventRecord(c->ll_ready, c->stream));
    CU(cudaStreamWaitEvent(c->ll_stream, c->ll_ready, 0));
    {
        ProfScope ps(c, 3, c->ll_stream);
        ST(launch_loglik_cols(c, out, c->ll_stream));
    }
    CU(cudaEventRecord(c->ll_done, c->ll_stream));
    c->ll_pending = true;
    return SIGNER_OK;
}

// Touch every page of a caller-allocated output buffer (it is ours to overwrite) so that the later
// device-to-host copies land on resident pages; runs on helper threads while the GPU does the burn-in.
void prefault(double *p, size_t n_doubles)
{
    if (!p) return;
    volatile char *b = reinterpret_cast<volatile char *>(p);
    const size_t bytes = n_doubles * sizeof(double);
    if (g_no_hugepage) {
        // A caller that asked for transparent huge pages (NumPy does for large arrays) makes every first touch wait
        // for a free 2 MB block -- direct compaction inside the fault, measured as sporadic stalls of up to a second
        // per call.  These pages are written once and read once: plain 4 KB faults, spread over the helper threads,
        // are faster and predictable.  (A hint only; R's vectors are not huge-page backed in the first place.)
        const uintptr_t a0 = (reinterpret_cast<uintptr_t>(p) + 4095) & ~uintptr_t(4095);
        const uintptr_t a1 = (reinterpret_cast<uintptr_t>(p) + bytes) & ~uintptr_t(4095);
        if (a1 > a0) madvise(reinterpret_cast<void *>(a0), a1 - a0, MADV_NOHUGEPAGE);
    }
    for (size_t off = 0; off < bytes; off += 4096) b[off] = 0;
    if (bytes) b[bytes - 1] = 0;
}

int rebuild_a_cache(signer_chain *c)
{
    a_cache_init<<<(unsigned)((c->nP + 127) / 128), 128, 0, c->stream>>>(c->Ap, c->nP, c->h.var_ap, ACache{c->acache_p, c->acache_p + c->nP, c->acache_p + 2 * c->nP});
    a_cache_init<<<(unsigned)((c->nE + 127) / 128), 128, 0, c->stream>>>(c->Ae, c->nE, c->h.var_ae, ACache{c->acache_e, c->acache_e + c->nE, c->acache_e + 2 * c->nE});
    CU(cudaGetLastError());
    return SIGNER_OK;
}

struct Slots { double *Ps = nullptr, *Es = nullptr, *Aps = nullptr, *Aes = nullptr, *Bps = nullptr, *Bes = nullptr; };

// One sweep: the seven steps of `order` (0 = skip) regrouped into at most three launches.
struct SweepPlan {
    Ops pops{0, 0u}, eops{0, 0u};
    bool has3 = false;
    struct Item { int pos, kind; } items[3];
    int n = 0;
};

SweepPlan plan_sweep(const int order[7], bool Pfixed)
{
    SweepPlan pl;
    int pos1 = -1, posP = -1, posE = -1;
    bool has2 = false;
    for (int f = 0; f < 7; ++f) {
        const int s = order[f];
        if (s == 1) { if (pos1 < 0) pos1 = f; }
        else if (s == 2 || s == 4 || s == 6) {
            if (Pfixed) continue;                       // guards of src/gibbs_2.cpp:311,313,315
            if (pl.pops.n < 3) pl.pops.push(s);
            if (s == 2) { posP = f; has2 = true; } else if (!has2 && posP < 0) posP = f;
        } else if (s == 3 || s == 5 || s == 7) {
            if (pl.eops.n < 3) pl.eops.push(s);
            if (s == 3) { posE = f; pl.has3 = true; } else if (!pl.has3 && posE < 0) posE = f;
        }
    }
    if (pos1 >= 0) pl.items[pl.n++] = {pos1, 1};
    if (pl.pops.n) pl.items[pl.n++] = {posP, 2};
    if (pl.eops.n) pl.items[pl.n++] = {posE, 3};
    for (int a = 1; a < pl.n; ++a)                      // at most three items: insertion sort by position
        for (int b = a; b > 0 && pl.items[b].pos < pl.items[b - 1].pos; --b) std::swap(pl.items[b], pl.items[b - 1]);
    return pl;
}

int enqueue_sweep(signer_chain *c, uint32_t sweep, const int order[7], const Slots *slots, double *zcube)
{
    const SweepPlan pl = plan_sweep(order, c->Pfixed != 0);
    const Ops &pops = pl.pops, &eops = pl.eops;
    const bool has3 = pl.has3;
    const SweepPlan::Item *items = pl.items;
    const int n = pl.n;
    bool storedP = false, storedE = false;
    for (int t = 0; t < n; ++t) {
        if (items[t].kind == 1) ST(launch_z(c, sweep, kStZ, zcube));
        else if (items[t].kind == 2) {
            ST(launch_p(c, sweep, pops, slots ? slots->Ps : nullptr, slots ? slots->Aps : nullptr, slots ? slots->Bps : nullptr));
            storedP = true;
        } else {
            ST(launch_e(c, sweep, eops, (has3 && !c->Pfixed) ? 1 : 0, slots ? slots->Es : nullptr, slots ? slots->Aes : nullptr,
                        slots ? slots->Bes : nullptr));
            storedE = true;
        }
    }
    if (slots) {   // families not touched this sweep (Pfixed, forced orders): plain copies of the state
        if (!storedP) {
            if (slots->Ps) CU(cudaMemcpyAsync(slots->Ps, c->P, c->nP * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            if (slots->Aps) CU(cudaMemcpyAsync(slots->Aps, c->Ap, c->nP * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            if (slots->Bps) CU(cudaMemcpyAsync(slots->Bps, c->Bp, c->nP * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
        if (!storedE) {
            if (slots->Es) CU(cudaMemcpyAsync(slots->Es, c->E, c->nE * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            if (slots->Aes) CU(cudaMemcpyAsync(slots->Aes, c->Ae, c->nE * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            if (slots->Bes) CU(cudaMemcpyAsync(slots->Bes, c->Be, c->nE * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    return SIGNER_OK;
}

int ensure_rings(signer_chain *c, int eval, int keep_par)
{
    const size_t per_sweep = sizeof(double) * (c->nP + c->nE) * (keep_par ? 3 : 2);
    size_t ring_bytes = kRingBytes;
    if (const char *e = std::getenv("SIGNER_RING_BYTES")) {     // tests: force many small chunks through the copy-out path
        const long long v = std::atoll(e);
        if (v > 0) ring_bytes = (size_t)v;
    }
    int cap = (int)std::max<size_t>(1, std::min<size_t>((size_t)eval, ring_bytes / per_sweep));
    if (c->ring_cap >= cap && c->ring_keep >= keep_par) return SIGNER_OK;
    for (int b = 0; b < 2; ++b) {
        signer_chain::free_ring(c->ring[b]);
        Ring &r = c->ring[b];
        CU(cudaMalloc(&r.Ps, sizeof(double) * c->nP * cap));
        CU(cudaMalloc(&r.Es, sizeof(double) * c->nE * cap));
        CU(cudaMalloc(&r.Bps, sizeof(double) * c->nP * cap));
        CU(cudaMalloc(&r.Bes, sizeof(double) * c->nE * cap));
        if (keep_par) {
            CU(cudaMalloc(&r.Aps, sizeof(double) * c->nP * cap));
            CU(cudaMalloc(&r.Aes, sizeof(double) * c->nE * cap));
        }
        CU(cudaEventCreateWithFlags(&r.filled, cudaEventDisableTiming));
    }
    c->ring_cap = cap; c->ring_keep = keep_par;
    return SIGNER_OK;
}

// copy chunk `chunk` (sweeps r0 .. r0+cnt-1 of the eval phase) from its ring set to the host: posted to the
// device's copy-out pool, which waits for the ring set's `filled` event
int flush_chunk(signer_chain *c, int chunk, int r0, int cnt, int keep_par, signer_outputs *out)
{
    Ring &r = c->ring[chunk & 1];
    CopyPool &pool = CopyPool::of(c->device);
    auto cp = [&](double *dst, const double *src, size_t per) {
        if (!dst || !src) return;
        pool.post(src, dst + per * (size_t)r0, sizeof(double) * per * (size_t)cnt, r.filled, r.inflight);
    };
    if (!c->sumPs) { cp(out->Ps, r.Ps, c->nP); cp(out->Es, r.Es, c->nE); }   // else: from the full-length stores at the end
    cp(out->Bps, r.Bps, c->nP); cp(out->Bes, r.Bes, c->nE);
    if (keep_par) { cp(out->Aps, r.Aps, c->nP); cp(out->Aes, r.Aes, c->nE); }
    return SIGNER_OK;
}

// the ring set's samples have all reached the caller's arrays (it may be refilled)
int wait_ring(signer_chain *c, Ring &r)
{
    if (!CopyPool::of(c->device).wait(r.inflight)) return fail(SIGNER_ERR_CUDA, CopyPool::of(c->device).error());
    return SIGNER_OK;
}

// Phat / Ehat from the full-length sample stores (see include/signer_gibbs.h)
int device_summaries(signer_chain *c, int R, double *Phat, double *Ehat)
{
    double *s = nullptr, *T = nullptr, *med = nullptr;
    const size_t nmax = std::max(c->nP, c->nE);
    CU(cudaMalloc(&s, sizeof(double) * (size_t)R * c->K));
    CU(cudaMalloc(&T, sizeof(double) * nmax * (size_t)R));
    CU(cudaMalloc(&med, sizeof(double) * nmax));
    sig_sums<<<(R * c->K + 127) / 128, 128, 0, c->stream>>>(c->sumPs, c->I, c->K, R, s);
    for (int mode = 0; mode < 2; ++mode) {
        const size_t n = mode == 0 ? c->nP : c->nE;
        double *dst = mode == 0 ? Phat : Ehat;
        if (!dst) continue;
        norm_transpose<<<dim3((unsigned)((n + 31) / 32), (unsigned)((R + 31) / 32)), dim3(32, 8), 0, c->stream>>>(
            mode == 0 ? c->sumPs : c->sumEs, s, (int)n, R, c->I, c->K, mode, T);
        median_rows<<<(unsigned)((n + 7) / 8), 256, 0, c->stream>>>(T, (int)n, R, med);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(dst, med, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->launches += 2;
    }
    c->launches += 1;
    cudaFree(s); cudaFree(T); cudaFree(med);
    return SIGNER_OK;
}

int fetch_state(signer_chain *c, signer_outputs *out)
{
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    auto cp = [&](void *dst, const void *src, size_t bytes) -> int {
        if (!dst) return SIGNER_OK;
        CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
        return SIGNER_OK;
    };
    ST(cp(out->P, c->P, c->nP * 8)); ST(cp(out->Ap, c->Ap, c->nP * 8)); ST(cp(out->Bp, c->Bp, c->nP * 8));
    ST(cp(out->E, c->E, c->nE * 8)); ST(cp(out->Ae, c->Ae, c->nE * 8)); ST(cp(out->Be, c->Be, c->nE * 8));
    ST(cp(out->Zin, c->Zin[c->zcur], c->nP * 4)); ST(cp(out->Znj, c->Znj[c->zcur], c->nE * 4));
    return SIGNER_OK;
}

int chain_create(std::shared_ptr<DeviceData> data, int K, const signer_state *st, const signer_opts *o, signer_chain **out,
                 bool deferred_init = false)
{
    if (!st || !o || !out) return fail(SIGNER_ERR_ARG, "null argument");
    if (o->Zfixed || o->Thetafixed || o->Afixed)
        return fail(SIGNER_ERR_UNSUPPORTED, "Zfixed/Thetafixed/Afixed=TRUE are not reachable from signeR and not implemented");
    if (K <= 0 || K > 128) return fail(SIGNER_ERR_ARG, "K must be in 1..128");
    if ((long long)K * data->J >= (1LL << 31) || (long long)data->I * K >= (1LL << 31)) return fail(SIGNER_ERR_ARG, "K*J and I*K must be below 2^31");
    if (!st->P || !st->E || !st->Ap || !st->Bp || !st->Ae || !st->Be) return fail(SIGNER_ERR_ARG, "state: P,E,Ap,Bp,Ae,Be must be non-null");
    CU(cudaSetDevice(data->device));
    std::unique_ptr<signer_chain> c(new signer_chain());
    c->data = data; c->device = data->device; c->I = data->I; c->J = data->J; c->Jp = data->Jp; c->K = K;
    c->nP = (size_t)c->I * K; c->nE = (size_t)K * c->J;
    c->h = Hyper{o->ap, o->bp, o->ae, o->be, o->lp, o->le, o->var_ap, o->var_ae};
    c->key = Key{(uint32_t)o->seed, (uint32_t)(o->seed >> 32)};
    c->sweep0 = o->sweep0; c->Pfixed = o->Pfixed ? 1 : 0;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->ll_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ll_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ll_done, cudaEventDisableTiming));
    c->stream = c->own_stream;
    auto up = [&](double **dst, const double *src, size_t n) -> int {
        CU(cudaMalloc(dst, n * sizeof(double)));
        CU(cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        return SIGNER_OK;
    };
    ST(up(&c->P, st->P, c->nP)); ST(up(&c->Ap, st->Ap, c->nP)); ST(up(&c->Bp, st->Bp, c->nP));
    ST(up(&c->E, st->E, c->nE)); ST(up(&c->Ae, st->Ae, c->nE)); ST(up(&c->Be, st->Be, c->nE));
    for (int b = 0; b < 2; ++b) {
        CU(cudaMalloc(&c->Zin[b], c->nP * sizeof(int)));
        CU(cudaMalloc(&c->Znj[b], c->nE * sizeof(int)));
        CU(cudaMemsetAsync(c->Zin[b], 0, c->nP * sizeof(int), c->stream));
        CU(cudaMemsetAsync(c->Znj[b], 0, c->nE * sizeof(int), c->stream));
    }
    c->n_seg = (c->J + kSeg - 1) / kSeg;
    CU(cudaMalloc(&c->corrE_part, sizeof(double) * (size_t)c->n_seg * c->nP));
    CU(cudaMalloc(&c->col, sizeof(double) * (size_t)data->n2));
    CU(cudaMemsetAsync(c->col, 0, sizeof(double) * (size_t)data->n2, c->stream));
    CU(cudaMalloc(&c->tile_counter, sizeof(unsigned long long)));
    CU(cudaMemsetAsync(c->tile_counter, 0, sizeof(unsigned long long), c->stream));
    CU(cudaMalloc(&c->ll_ticket, sizeof(unsigned)));
    CU(cudaMemsetAsync(c->ll_ticket, 0, sizeof(unsigned), c->stream));
    CU(cudaMalloc(&c->stats, 8 * sizeof(double)));
    CU(cudaMalloc(&c->acache_p, 3 * c->nP * sizeof(double)));
    CU(cudaMalloc(&c->acache_e, 3 * c->nE * sizeof(double)));
    ST(rebuild_a_cache(c.get()));

    // launch geometry
    c->n_batches = (data->n_cells + 31) / 32;
    c->n_chunks = (c->n_batches + kZChunk - 1) / kZChunk;
    c->z_warp_smem = (int)z_smem_bytes_per_warp(K);
    c->z_zin_smem = (c->nP * sizeof(int) <= (size_t)kZinSmemMax) ? 1 : 0;
    ST(ensure_device_setup(c->device));
    // block size: the configuration that keeps the most warps resident per SM (at most kZMaxWarps: the register
    // file holds that many at the kernel's 80 registers); ties go to fewer, larger blocks (fewer Zin accumulators
    // in shared memory, which leaves more of the SM's 256 KB to L1)
    int sms = 0, best_w = 0, best_b = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    for (int w : {24, 12, 8, 6, 4, 3, 2, 1}) {
        const size_t zsm = (size_t)c->z_warp_smem * w + (c->z_zin_smem ? c->nP * sizeof(int) : 0);
        if (zsm > (size_t)g_z_dyn_smem_max[c->device]) continue;
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, z_alloc_fused, 32 * w, zsm));
        occ = std::min(occ, kZMaxWarps / w);
        if (occ * w > best_w * best_b) { best_w = w; best_b = occ; }
    }
    if (best_w == 0) return fail(SIGNER_ERR_ARG, "z_alloc_fused: K too large for shared memory");
    if (const char *e = std::getenv("SIGNER_Z_BLOCK")) {          // experiments: "warps:blocks_per_sm"
        int w = 0, b = 0;
        if (std::sscanf(e, "%d:%d", &w, &b) == 2 && w >= 1 && w <= kZMaxWarps && b >= 1) { best_w = w; best_b = b; }
    }
    c->z_warps = best_w;
    c->z_grid = std::max(1, std::min((c->n_chunks + best_w - 1) / best_w, best_b * sms));
    c->ksplit = (K + kEGroup - 1) / kEGroup;


    // initial sufficient statistics: reduce the given cube, or draw the first allocation on the device
    if (deferred_init) {                                  // a column shard: its peers take part (sharded_run)
        CU(cudaStreamSynchronize(c->stream));
        *out = c.release();
        return SIGNER_OK;
    }
    if (st->Z) {
        double *slab[2] = {nullptr, nullptr};
        const size_t ns = (size_t)c->I * c->J;
        CU(cudaMalloc(&slab[0], ns * sizeof(double)));
        CU(cudaMalloc(&slab[1], ns * sizeof(double)));
        for (int k = 0; k < K; ++k) {
            CU(cudaMemcpyAsync(slab[k & 1], st->Z + ns * k, ns * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            zslab_reduce<<<(c->J + 127) / 128, 128, 0, c->stream>>>(slab[k & 1], c->I, c->J, K, k, c->Zin[0], c->Znj[0]);
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaFree(slab[0])); CU(cudaFree(slab[1]));
        c->zcur = 0;
    } else {
        c->zcur = 0;
        ST(launch_z(c.get(), c->sweep0, kStZinit, nullptr));
    }
    if (!c->Pfixed) {   // W E' of the initial E, in case step 2 comes before step 3
        Ops none{0, 0u};
        ST(launch_e(c.get(), c->sweep0, none, 1, nullptr, nullptr, nullptr));
    }
    CU(cudaStreamSynchronize(c->stream));
    *out = c.release();
    return SIGNER_OK;
}

int chain_run(signer_chain *c, int burn, int eval, int keep_par, const int32_t *forced, signer_outputs *out)
{
    if (burn < 0 || eval < 0) return fail(SIGNER_ERR_ARG, "burn and eval must be non-negative");
    CU(cudaSetDevice(c->device));
    keep_par = keep_par ? 1 : 0;
    if (eval > 0) {
        ST(ensure_rings(c, eval, keep_par));
        if (c->ll_cap < eval) {
            cudaFree(c->ll_ring); c->ll_ring = nullptr; c->ll_cap = 0;
            CU(cudaMalloc(&c->ll_ring, sizeof(double) * (size_t)eval));
            c->ll_cap = eval;
        }
    }
    double *zcube = nullptr;
    if (keep_par && out && out->Zs && eval > 0) {
        if (!c->Zcube) CU(cudaMalloc(&c->Zcube, sizeof(double) * (size_t)c->I * c->J * c->K));
        zcube = c->Zcube;
    }
    const bool summ = out && eval > 0 && (out->Phat || out->Ehat);
    if (summ) {                                           // full-length stores for Ps and Es
        cudaFree(c->sumPs); cudaFree(c->sumEs); c->sumPs = c->sumEs = nullptr;
        CU(cudaMalloc(&c->sumPs, sizeof(double) * c->nP * (size_t)eval));
        CU(cudaMalloc(&c->sumEs, sizeof(double) * c->nE * (size_t)eval));
    }
    const int total = burn + eval, cap = std::max(1, c->ring_cap);
    int chunk_r0 = 0;
    std::vector<std::thread> faulters;
    auto join_faulters = [&]() { for (auto &t : faulters) if (t.joinable()) t.join(); faulters.clear(); };
    if (out && eval > 0) {
        const size_t R = (size_t)eval;
        double *bufs[7] = {out->Es, out->Bes, keep_par ? out->Aes : nullptr, out->Ps, out->Bps, keep_par ? out->Aps : nullptr, zcube ? out->Zs : nullptr};
        const size_t lens[7] = {c->nE * R, c->nE * R, c->nE * R, c->nP * R, c->nP * R, c->nP * R, (size_t)c->I * c->J * c->K};
        // large arrays are touched by several threads each (page faults scale across threads)
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        const size_t max_threads = std::min<size_t>(32, hw / 2);
        size_t big_bytes = 0;
        for (int b = 0; b < 7; ++b) if (bufs[b] && lens[b] * sizeof(double) >= (size_t(8) << 20)) big_bytes += lens[b] * sizeof(double);
        for (int b = 0; b < 7; ++b) {
            if (!bufs[b] || lens[b] * sizeof(double) < (size_t(8) << 20)) continue;
            const size_t parts = std::max<size_t>(1, std::min<size_t>(max_threads, max_threads * (lens[b] * sizeof(double)) / big_bytes));
            const size_t step = ((lens[b] + parts - 1) / parts + 511) / 512 * 512;       // whole pages
            for (size_t off = 0; off < lens[b]; off += step) faulters.emplace_back(prefault, bufs[b] + off, std::min(step, lens[b] - off));
        }
    }
    struct Joiner { std::function<void()> f; ~Joiner() { f(); } } joiner{join_faulters};
    // whatever way this call ends, no copy-out job may outlive it: they write into the caller's arrays
    Joiner drain{[c, out]() { if (out) { CopyPool &p = CopyPool::of(c->device); p.wait(c->ring[0].inflight); p.wait(c->ring[1].inflight); } }};
    for (int k = 0; k < total; ++k) {
        const uint32_t sweep = c->sweep0 + (uint32_t)c->sweeps_done;
        int order[7];
        if (forced) for (int f = 0; f < 7; ++f) order[f] = forced[(size_t)7 * k + f];
        else host_perm(((uint64_t)c->key.k1 << 32) | c->key.k0, sweep, order);
        const bool storing = k >= burn;
        if (!storing) {
            ST(enqueue_sweep(c, sweep, order, nullptr, nullptr));
        } else {
            const int r = k - burn, chunk = r / cap, slot = r % cap;
            Ring &rg = c->ring[chunk & 1];
            if (slot == 0 && chunk >= 2 && out) ST(wait_ring(c, rg));          // the set's previous chunk has left the device
            Slots s;
            s.Ps = rg.Ps + c->nP * slot; s.Es = rg.Es + c->nE * slot;
            s.Bps = rg.Bps + c->nP * slot; s.Bes = rg.Bes + c->nE * slot;
            if (keep_par) { s.Aps = rg.Aps + c->nP * slot; s.Aes = rg.Aes + c->nE * slot; }
            if (summ) { s.Ps = c->sumPs + c->nP * (size_t)r; s.Es = c->sumEs + c->nE * (size_t)r; }
            const bool last = (k == total - 1);
            // Zs keeps only the latest Z (src/gibbs_2.cpp:323): store the cube in the last sweep only
            ST(enqueue_sweep(c, sweep, order, &s, last ? zcube : nullptr));
            ST(launch_loglik(c, c->ll_ring + r));
            if (slot == cap - 1 || last) {
                CU(cudaEventRecord(rg.filled, c->stream));
                if (out) ST(flush_chunk(c, chunk, chunk_r0, slot + 1, keep_par, out));
                chunk_r0 += cap;
            }
        }
        c->sweeps_done++;
    }
    ST(wait_loglik(c));                                    // the launching stream ends after the last log-likelihood
    c->last_eval = eval;
    if (out && eval > 0) {
        CU(cudaStreamSynchronize(c->stream));
        ST(wait_ring(c, c->ring[0])); ST(wait_ring(c, c->ring[1]));
        join_faulters();
        if (zcube && out->Zs) {
            // step 1 may not have run in the last sweep only under a forced order; then Zs is undefined
            CU(cudaMemcpy(out->Zs, zcube, sizeof(double) * (size_t)c->I * c->J * c->K, cudaMemcpyDeviceToHost));
        }
        if (summ) {
            if (out->Ps) CU(cudaMemcpy(out->Ps, c->sumPs, sizeof(double) * c->nP * (size_t)eval, cudaMemcpyDeviceToHost));
            if (out->Es) CU(cudaMemcpy(out->Es, c->sumEs, sizeof(double) * c->nE * (size_t)eval, cudaMemcpyDeviceToHost));
            ST(device_summaries(c, eval, out->Phat, out->Ehat));
            cudaFree(c->sumPs); cudaFree(c->sumEs); c->sumPs = c->sumEs = nullptr;
        }
        if (out->loglik || out->bic) {
            std::vector<double> ll((size_t)eval);
            CU(cudaMemcpy(ll.data(), c->ll_ring, sizeof(double) * (size_t)eval, cudaMemcpyDeviceToHost));
            const double pen = (double)c->K * (double)(c->I + c->J) * std::log((double)c->J);   // R/cpp_eBayesNMF.R:186
            for (int r = 0; r < eval; ++r) {
                if (out->loglik) out->loglik[r] = ll[r];
                if (out->bic) out->bic[r] = 2.0 * ll[r] - pen;
            }
        }
    }
    if (out) ST(fetch_state(c, out));
    return SIGNER_OK;
}

// ------------------------------------------------------------------------------------------------
// Genome-column sharding: one chain over several GPUs.  Shard g owns the 128-genome
// segments [g*n_seg/N, (g+1)*n_seg/N) -- M, W, E, Ae, Be, Znj for those genomes; P, Ap, Bp are
// replicated and updated redundantly with identical Philox counters, so they stay identical.
// Per sweep two exchanges of I*K values: the integer sums Zin after step 1 and the shards' ordered
// W E' sums after step 3 (every shard fills only its slot of an [N][I*K] buffer, so the all-reduce
// adds zeros and is exact in any order); per evaluation sweep one all-reduce of the per-genome
// log-likelihood terms.  Results are bit-identical to the oracle run with nshards = N.
// NCCL (all-reduce over NVLink) is loaded lazily; shards placed on ONE device are emulated with a
// summing kernel so that the logic is testable on a single GPU.
// ------------------------------------------------------------------------------------------------
typedef void *nccl_comm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
    }
};
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0;   // nccl.h: ncclInt32, ncclFloat64, ncclSum

struct ShardSet {
    int n = 0;
    bool same_device = true;
    std::vector<signer_chain *> sh;
    std::vector<nccl_comm_t> comms;
    NcclApi nccl;
    void **ptr_tab = nullptr;        // device table of buffer pointers for the same-device emulation
    ~ShardSet()
    {
        for (auto *c : sh) delete c;
        for (auto cm : comms) if (cm && nccl.CommDestroy) nccl.CommDestroy(cm);
        if (ptr_tab) cudaFree(ptr_tab);
    }
};

// all-reduce (sum) of one buffer per shard, in place
template <typename T>
int shard_allreduce(ShardSet &S, const std::vector<T *> &bufs, size_t count)
{
    if (S.n == 1) return SIGNER_OK;
    if (S.same_device) {
        signer_chain *c0 = S.sh[0];
        CU(cudaSetDevice(c0->device));
        CU(cudaMemcpyAsync(S.ptr_tab, bufs.data(), sizeof(void *) * S.n, cudaMemcpyHostToDevice, c0->stream));
        sum_shards<T><<<(unsigned)std::min<size_t>((count + 255) / 256, 1024), 256, 0, c0->stream>>>(reinterpret_cast<T *const *>(S.ptr_tab), S.n, count);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c0->stream));    // the host-side pointer table is reused by the next call
        return SIGNER_OK;
    }
    const int dt = sizeof(T) == 4 ? kNcclInt32 : kNcclFloat64;
    int rc = S.nccl.GroupStart();
    for (int g = 0; g < S.n && rc == 0; ++g)
        rc = S.nccl.AllReduce(bufs[g], bufs[g], count, dt, kNcclSum, S.comms[g], S.sh[g]->stream);
    const int rc2 = S.nccl.GroupEnd();
    if (rc || rc2) return fail(SIGNER_ERR_NCCL, std::string("ncclAllReduce: ") + (S.nccl.GetErrorString ? S.nccl.GetErrorString(rc ? rc : rc2) : "error"));
    return SIGNER_OK;
}

int shard_launch_z(ShardSet &S, uint32_t sweep, uint32_t stream_id, bool cube)
{
    std::vector<int *> zin(S.n);
    for (int g = 0; g < S.n; ++g) {
        signer_chain *c = S.sh[g];
        CU(cudaSetDevice(c->device));
        ST(launch_z(c, sweep, stream_id, cube ? c->Zcube : nullptr));
        zin[g] = c->Zin[c->zcur];
    }
    return shard_allreduce<int>(S, zin, S.sh[0]->nP);
}

int shard_launch_e(ShardSet &S, uint32_t sweep, const Ops &ops, int do_corrE, const std::vector<Slots> *slots)
{
    std::vector<double *> all(S.n);
    for (int g = 0; g < S.n; ++g) {
        signer_chain *c = S.sh[g];
        CU(cudaSetDevice(c->device));
        const Slots *sl = slots ? &(*slots)[g] : nullptr;
        ST(launch_e(c, sweep, ops, do_corrE, sl ? sl->Es : nullptr, sl ? sl->Aes : nullptr, sl ? sl->Bes : nullptr));
        if (do_corrE) {
            CU(cudaMemsetAsync(c->corrE_all, 0, sizeof(double) * c->nP * S.n, c->stream));
            seg_reduce<<<((int)c->nP + 127) / 128, 128, 0, c->stream>>>(c->corrE_part, c->n_seg, (int)c->nP, c->corrE_all + c->nP * g);
            CU(cudaGetLastError());
        }
        all[g] = c->corrE_all;
    }
    if (do_corrE) return shard_allreduce<double>(S, all, S.sh[0]->nP * S.n);
    return SIGNER_OK;
}

// column terms of every shard into its part of the global array, all-reduce, tree on every shard
int shard_loglik(ShardSet &S, int mode_lgm, double *out0)
{
    std::vector<double *> cols(S.n);
    for (int g = 0; g < S.n; ++g) {
        signer_chain *c = S.sh[g];
        CU(cudaSetDevice(c->device));
        CU(cudaMemsetAsync(c->col, 0, sizeof(double) * (size_t)c->data->n2, c->stream));
        if (mode_lgm) {
            LLParams lp{c->I, c->J, c->Jp, 1, c->data->Mrow, c->data->W, nullptr, nullptr, c->col + c->col0, 1, nullptr, 0, nullptr, nullptr, nullptr};
            loglik_cols<<<(c->J + 31) / 32, 32 * kLLWarps, loglik_smem_bytes(1), c->stream>>>(lp);
            CU(cudaGetLastError());
        } else {
            ST(launch_loglik_cols(c));
        }
        cols[g] = c->col;
    }
    ST(shard_allreduce<double>(S, cols, (size_t)S.sh[0]->data->n2));
    for (int g = 0; g < S.n; ++g) {
        signer_chain *c = S.sh[g];
        CU(cudaSetDevice(c->device));
        if (mode_lgm) tree_sum<<<1, 1024, 0, c->stream>>>(c->col, c->data->n2, nullptr, c->data->lgm);
        else if (g == 0) tree_sum<<<1, 1024, 0, c->stream>>>(c->col, c->data->n2, c->data->lgm, out0);
        CU(cudaGetLastError());
    }
    return SIGNER_OK;
}

int shard_sync(ShardSet &S)
{
    for (auto *c : S.sh) { CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->stream)); }
    return SIGNER_OK;
}

int sharded_run(const signer_problem *pr, const signer_state *st, const signer_opts *o, signer_outputs *out,
                const int32_t *devices, int n_dev)
{
    const int I = pr->I, J = pr->J, K = pr->K;
    if (n_dev < 1 || n_dev > 64) return fail(SIGNER_ERR_ARG, "sharded run: 1..64 shards");
    const int n_seg = (J + kSeg - 1) / kSeg;
    if (n_dev > n_seg) return fail(SIGNER_ERR_ARG, "sharded run: more shards than 128-genome segments");
    ShardSet S;
    S.n = n_dev;
    for (int g = 1; g < n_dev; ++g) if (devices[g] != devices[0]) S.same_device = false;
    if (!S.same_device) {
        for (int a = 0; a < n_dev; ++a) for (int b = a + 1; b < n_dev; ++b)
            if (devices[a] == devices[b]) return fail(SIGNER_ERR_ARG, "sharded run: devices must be all distinct or all the same");
        if (!S.nccl.load()) return fail(SIGNER_ERR_NCCL, "cannot load libnccl.so.2");
        S.comms.assign(n_dev, nullptr);
        std::vector<int> devs(devices, devices + n_dev);
        const int rc = S.nccl.CommInitAll(S.comms.data(), n_dev, devs.data());
        if (rc) return fail(SIGNER_ERR_NCCL, std::string("ncclCommInitAll: ") + (S.nccl.GetErrorString ? S.nccl.GetErrorString(rc) : "error"));
    }
    // create the shards
    for (int g = 0; g < n_dev; ++g) {
        const int s0 = (int)((long long)g * n_seg / n_dev), s1 = (int)((long long)(g + 1) * n_seg / n_dev);
        const int c0 = s0 * kSeg, c1 = std::min(J, s1 * kSeg), Jg = c1 - c0;
        std::shared_ptr<DeviceData> d;
        ST(upload_data(devices[g], I, Jg, pr->M + (size_t)I * c0, pr->W + (size_t)I * c0, d, c0, J));
        signer_state sg = *st;
        sg.Z = nullptr;                                  // the cube (if any) is reduced below, per shard
        sg.E = st->E + (size_t)K * c0; sg.Ae = st->Ae + (size_t)K * c0; sg.Be = st->Be + (size_t)K * c0;
        signer_chain *c = nullptr;
        signer_opts og = *o;
        og.device = devices[g];
        ST(chain_create(d, K, &sg, &og, &c, /*deferred_init=*/true));
        S.sh.push_back(c);
        c->shard = g; c->n_shards = n_dev; c->col0 = c0;
        CU(cudaMalloc(&c->corrE_all, sizeof(double) * c->nP * n_dev));
        if (S.same_device && g > 0) { CU(cudaStreamSynchronize(c->stream)); c->stream = S.sh[0]->stream; }   // one stream: plain ordering
    }
    if (S.same_device) { CU(cudaSetDevice(devices[0])); CU(cudaMalloc(&S.ptr_tab, sizeof(void *) * n_dev)); }
    ST(shard_sync(S));
    ST(shard_loglik(S, 1, nullptr));                     // sum lgamma(M+1) over the whole cohort, on every shard
    // initial sufficient statistics and W E'
    if (st->Z) {
        for (int g = 0; g < n_dev; ++g) {
            signer_chain *c = S.sh[g];
            CU(cudaSetDevice(c->device));
            double *slab = nullptr;
            const size_t ns = (size_t)I * c->J;
            CU(cudaMalloc(&slab, ns * sizeof(double)));
            for (int k = 0; k < K; ++k) {
                CU(cudaMemcpyAsync(slab, st->Z + (size_t)I * J * k + (size_t)I * c->col0, ns * sizeof(double), cudaMemcpyHostToDevice, c->stream));
                zslab_reduce<<<(c->J + 127) / 128, 128, 0, c->stream>>>(slab, I, c->J, K, k, c->Zin[0], c->Znj[0]);
            }
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(c->stream));
            CU(cudaFree(slab));
            c->zcur = 0;
        }
        std::vector<int *> zin(n_dev);
        for (int g = 0; g < n_dev; ++g) zin[g] = S.sh[g]->Zin[0];
        ST(shard_allreduce<int>(S, zin, S.sh[0]->nP));
    } else {
        ST(shard_launch_z(S, o->sweep0, kStZinit, false));
    }
    if (!o->Pfixed) { Ops none{0, 0u}; ST(shard_launch_e(S, o->sweep0, none, 1, nullptr)); }

    const int burn = o->burn, eval = o->eval, keep_par = o->keep_par ? 1 : 0, total = burn + eval;
    if (burn < 0 || eval < 0) return fail(SIGNER_ERR_ARG, "burn and eval must be non-negative");
    int cap = 1;
    if (eval > 0) {
        for (auto *c : S.sh) { CU(cudaSetDevice(c->device)); ST(ensure_rings(c, eval, keep_par)); cap = c->ring_cap; }
        for (auto *c : S.sh) cap = std::min(cap, c->ring_cap);
        signer_chain *c0 = S.sh[0];
        CU(cudaSetDevice(c0->device));
        CU(cudaMalloc(&c0->ll_ring, sizeof(double) * (size_t)eval)); c0->ll_cap = eval;
    }
    const bool want_cube = keep_par && out->Zs && eval > 0;
    if (want_cube) for (auto *c : S.sh) { CU(cudaSetDevice(c->device)); CU(cudaMalloc(&c->Zcube, sizeof(double) * (size_t)I * c->J * K)); }

    auto flush = [&](int r0, int cnt) -> int {            // ring slots 0..cnt-1 hold eval sweeps r0..r0+cnt-1
        ST(shard_sync(S));
        signer_chain *c0 = S.sh[0];
        CU(cudaSetDevice(c0->device));
        auto cpP = [&](double *dst, const double *src) -> int {
            if (dst && src) CU(cudaMemcpy(dst + c0->nP * (size_t)r0, src, sizeof(double) * c0->nP * (size_t)cnt, cudaMemcpyDeviceToHost));
            return SIGNER_OK;
        };
        ST(cpP(out->Ps, c0->ring[0].Ps)); ST(cpP(out->Bps, c0->ring[0].Bps)); if (keep_par) ST(cpP(out->Aps, c0->ring[0].Aps));
        for (auto *c : S.sh) {
            CU(cudaSetDevice(c->device));
            auto cpE = [&](double *dst, const double *src) -> int {   // a shard's genomes are a contiguous block of every slice
                if (dst && src)
                    CU(cudaMemcpy2D(dst + (size_t)K * J * (size_t)r0 + (size_t)K * c->col0, sizeof(double) * (size_t)K * J, src,
                                    sizeof(double) * c->nE, sizeof(double) * c->nE, (size_t)cnt, cudaMemcpyDeviceToHost));
                return SIGNER_OK;
            };
            ST(cpE(out->Es, c->ring[0].Es)); ST(cpE(out->Bes, c->ring[0].Bes)); if (keep_par) ST(cpE(out->Aes, c->ring[0].Aes));
        }
        return SIGNER_OK;
    };

    ST(shard_sync(S));
    const auto t_loop = std::chrono::steady_clock::now();
    int r0 = 0;
    for (int k = 0; k < total; ++k) {
        const uint32_t sweep = o->sweep0 + (uint32_t)k;
        int order[7];
        if (o->forced_order) for (int f = 0; f < 7; ++f) order[f] = o->forced_order[(size_t)7 * k + f];
        else host_perm(o->seed, sweep, order);
        const SweepPlan pl = plan_sweep(order, o->Pfixed != 0);
        const bool storing = k >= burn, last = (k == total - 1);
        const int r = k - burn, slot = storing ? r % cap : 0;
        std::vector<Slots> slots(S.n);
        if (storing)
            for (int g = 0; g < S.n; ++g) {
                signer_chain *c = S.sh[g]; Ring &rg = c->ring[0];
                Slots &s = slots[g];
                s.Es = rg.Es + c->nE * slot; s.Bes = rg.Bes + c->nE * slot; if (keep_par) s.Aes = rg.Aes + c->nE * slot;
                if (g == 0) { s.Ps = rg.Ps + c->nP * slot; s.Bps = rg.Bps + c->nP * slot; if (keep_par) s.Aps = rg.Aps + c->nP * slot; }
            }
        bool storedP = false, storedE = false;
        for (int t = 0; t < pl.n; ++t) {
            if (pl.items[t].kind == 1) ST(shard_launch_z(S, sweep, kStZ, want_cube && last));
            else if (pl.items[t].kind == 2) {
                for (int g = 0; g < S.n; ++g) {
                    signer_chain *c = S.sh[g];
                    CU(cudaSetDevice(c->device));
                    const Slots *sl = storing ? &slots[g] : nullptr;
                    ST(launch_p(c, sweep, pl.pops, sl ? sl->Ps : nullptr, sl ? sl->Aps : nullptr, sl ? sl->Bps : nullptr));
                }
                storedP = true;
            } else {
                ST(shard_launch_e(S, sweep, pl.eops, (pl.has3 && !o->Pfixed) ? 1 : 0, storing ? &slots : nullptr));
                storedE = true;
            }
        }
        if (storing) {
            for (int g = 0; g < S.n; ++g) {
                signer_chain *c = S.sh[g]; const Slots &s = slots[g];
                CU(cudaSetDevice(c->device));
                if (!storedP && g == 0) {
                    CU(cudaMemcpyAsync(s.Ps, c->P, c->nP * 8, cudaMemcpyDeviceToDevice, c->stream));
                    CU(cudaMemcpyAsync(s.Bps, c->Bp, c->nP * 8, cudaMemcpyDeviceToDevice, c->stream));
                    if (s.Aps) CU(cudaMemcpyAsync(s.Aps, c->Ap, c->nP * 8, cudaMemcpyDeviceToDevice, c->stream));
                }
                if (!storedE) {
                    CU(cudaMemcpyAsync(s.Es, c->E, c->nE * 8, cudaMemcpyDeviceToDevice, c->stream));
                    CU(cudaMemcpyAsync(s.Bes, c->Be, c->nE * 8, cudaMemcpyDeviceToDevice, c->stream));
                    if (s.Aes) CU(cudaMemcpyAsync(s.Aes, c->Ae, c->nE * 8, cudaMemcpyDeviceToDevice, c->stream));
                }
            }
            ST(shard_loglik(S, 0, S.sh[0]->ll_ring + r));
            if (slot == cap - 1 || last) { ST(flush(r0, slot + 1)); r0 += slot + 1; }
        }
    }
    ST(shard_sync(S));
    g_loop_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_loop).count();
    signer_chain *c0 = S.sh[0];
    CU(cudaSetDevice(c0->device));
    if (eval > 0 && (out->loglik || out->bic)) {
        std::vector<double> ll((size_t)eval);
        CU(cudaMemcpy(ll.data(), c0->ll_ring, sizeof(double) * (size_t)eval, cudaMemcpyDeviceToHost));
        const double pen = (double)K * (double)(I + J) * std::log((double)J);
        for (int r = 0; r < eval; ++r) { if (out->loglik) out->loglik[r] = ll[r]; if (out->bic) out->bic[r] = 2.0 * ll[r] - pen; }
    }
    auto cp = [&](void *dst, const void *src, size_t bytes) -> int { if (dst) CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return SIGNER_OK; };
    ST(cp(out->P, c0->P, c0->nP * 8)); ST(cp(out->Ap, c0->Ap, c0->nP * 8)); ST(cp(out->Bp, c0->Bp, c0->nP * 8));
    ST(cp(out->Zin, c0->Zin[c0->zcur], c0->nP * 4));
    for (auto *c : S.sh) {
        CU(cudaSetDevice(c->device));
        const size_t off = (size_t)K * c->col0;
        ST(cp(out->E ? out->E + off : nullptr, c->E, c->nE * 8)); ST(cp(out->Ae ? out->Ae + off : nullptr, c->Ae, c->nE * 8));
        ST(cp(out->Be ? out->Be + off : nullptr, c->Be, c->nE * 8)); ST(cp(out->Znj ? out->Znj + off : nullptr, c->Znj[c->zcur], c->nE * 4));
        if (want_cube)
            for (int k = 0; k < K; ++k)
                CU(cudaMemcpy(out->Zs + (size_t)I * J * k + (size_t)I * c->col0, c->Zcube + (size_t)I * c->J * k, sizeof(double) * (size_t)I * c->J,
                              cudaMemcpyDeviceToHost));
    }
    return SIGNER_OK;
}

} // namespace

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" {

const char *signer_last_error(void) { return g_err.c_str(); }
double signer_last_loop_ms(void) { return g_loop_ms; }
int signer_version(void) { return SIGNER_ABI_VERSION; }
int signer_draw_spec_version(void) { return SIGNER_DRAW_SPEC_VERSION; }

int signer_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return -SIGNER_ERR_CUDA; }
    return n;
}

int signer_chain_create(const signer_problem *pr, const signer_state *st, const signer_opts *o, signer_chain **chain)
{
    g_err.clear();
    if (!pr || !st || !o || !chain) return fail(SIGNER_ERR_ARG, "null argument");
    std::shared_ptr<DeviceData> d;
    ST(upload_data(o->device, pr->I, pr->J, pr->M, pr->W, d));
    return chain_create(d, pr->K, st, o, chain);
}

int signer_chain_set_stream(signer_chain *c, void *cuda_stream)
{
    if (!c) return fail(SIGNER_ERR_ARG, "null chain");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return SIGNER_OK;
}

int signer_chain_set_hyper(signer_chain *c, const double h[8])
{
    if (!c || !h) return fail(SIGNER_ERR_ARG, "null argument");
    const bool var_changed = (h[6] != c->h.var_ap) || (h[7] != c->h.var_ae);
    c->h = Hyper{h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]};
    if (var_changed) { CU(cudaSetDevice(c->device)); ST(rebuild_a_cache(c)); }
    return SIGNER_OK;
}

int signer_chain_run(signer_chain *c, int32_t burn, int32_t eval, int32_t keep_par, const int32_t *forced, signer_outputs *out)
{
    g_err.clear();
    if (!c) return fail(SIGNER_ERR_ARG, "null chain");
    return chain_run(c, burn, eval, keep_par, forced, out);
}

int signer_chain_sync(signer_chain *c)
{
    if (!c) return fail(SIGNER_ERR_ARG, "null chain");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return SIGNER_OK;
}

int signer_chain_fetch_state(signer_chain *c, signer_outputs *out)
{
    if (!c || !out) return fail(SIGNER_ERR_ARG, "null argument");
    return fetch_state(c, out);
}

int signer_chain_fetch_stats(signer_chain *c, double stats[8])
{
    if (!c || !stats) return fail(SIGNER_ERR_ARG, "null argument");
    if (c->last_eval <= 0 || c->last_eval > c->ring_cap || !c->ring_keep)
        return fail(SIGNER_ERR_ARG, "fetch_stats needs a preceding run with keep_par and eval within the device ring");
    CU(cudaSetDevice(c->device));
    CU(cudaMemsetAsync(c->stats, 0, 8 * sizeof(double), c->stream));
    const Ring &r = c->ring[0];
    const size_t np = c->nP * (size_t)c->last_eval, ne = c->nE * (size_t)c->last_eval;
    stats_reduce<<<256, 256, 0, c->stream>>>(r.Aps, r.Bps, np, c->stats);
    stats_reduce<<<1024, 256, 0, c->stream>>>(r.Aes, r.Bes, ne, c->stats + 4);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(stats, c->stats, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    stats[3] = (double)np; stats[7] = (double)ne;
    return SIGNER_OK;
}

int signer_chain_profile(signer_chain *c, int32_t enable, double ms[4], int64_t launches[4])
{
    if (!c) return fail(SIGNER_ERR_ARG, "null chain");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 4; ++k) {
        double tot = 0;
        for (auto &pr : c->ev[k]) {
            float t = 0;
            cudaEventElapsedTime(&t, pr.first, pr.second);
            tot += t;
            cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
        }
        if (ms) ms[k] = tot;
        if (launches) launches[k] = (int64_t)c->ev[k].size();
        c->ev[k].clear();
    }
    c->profiling = enable != 0;
    return SIGNER_OK;
}

int64_t signer_chain_sweeps(const signer_chain *c) { return c ? c->sweeps_done : -1; }
int64_t signer_chain_launches(const signer_chain *c) { return c ? c->launches : -1; }

int signer_chain_draw_counts(signer_chain *c, int32_t enable, uint64_t counts[2])
{
    if (!c) return fail(SIGNER_ERR_ARG, "null chain");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (counts) { counts[0] = counts[1] = 0; }
    if (c->draw_count && counts) CU(cudaMemcpy(counts, c->draw_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (enable && !c->draw_count) CU(cudaMalloc(&c->draw_count, 2 * sizeof(unsigned long long)));
    if (c->draw_count) CU(cudaMemset(c->draw_count, 0, 2 * sizeof(unsigned long long)));
    if (!enable && c->draw_count) { cudaFree(c->draw_count); c->draw_count = nullptr; }
    return SIGNER_OK;
}

int signer_draw_ceiling(int32_t device, double *draws_per_s)
{
    g_err.clear();
    if (!draws_per_s) return fail(SIGNER_ERR_ARG, "null argument");
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    unsigned *sink = nullptr;
    CU(cudaMalloc(&sink, 2 * sizeof(unsigned)));
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
    const int iters = 2000, blocks = sms * 8, threads = 256;
    draw_ceiling<<<blocks, threads>>>(Key{12345u, 678u}, 200, 0.37, sink);      // warm-up
    CU(cudaEventRecord(a));
    draw_ceiling<<<blocks, threads>>>(Key{12345u, 678u}, iters, 0.37, sink);
    CU(cudaEventRecord(b));
    CU(cudaEventSynchronize(b));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, a, b));
    *draws_per_s = 4.0 * (double)iters * blocks * threads / (ms * 1e-3);
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    return SIGNER_OK;
}

void signer_chain_destroy(signer_chain *c) { delete c; }

int signer_gibbs_run(const signer_problem *pr, const signer_state *st, const signer_opts *o, signer_outputs *out)
{
    g_err.clear();
    if (!pr || !st || !o || !out) return fail(SIGNER_ERR_ARG, "null argument");
    signer_chain *c = nullptr;
    ST(signer_chain_create(pr, st, o, &c));
    const auto t_loop = std::chrono::steady_clock::now();
    const int rc = chain_run(c, o->burn, o->eval, o->keep_par, o->forced_order, out);
    g_loop_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_loop).count();
    const std::string keep = g_err;
    signer_chain_destroy(c);
    g_err = keep;
    return rc;
}

int signer_gibbs_batch(int32_t I, int32_t J, const double *M, const double *W, signer_task *tasks, int32_t n_tasks,
                       const int32_t *devices, int32_t n_devices)
{
    g_err.clear();
    if (!tasks || n_tasks < 0 || !devices || n_devices <= 0) return fail(SIGNER_ERR_ARG, "bad batch arguments");
    // longest-first (cost ~ K * sweeps) dealt greedily onto the least-loaded device
    std::vector<int> idx(n_tasks);
    for (int t = 0; t < n_tasks; ++t) idx[t] = t;
    // measured: a sweep costs about (12 + K) units (the step-1 part barely depends on K)
    auto cost = [&](int t) { return (12.0 + (double)tasks[t].K) * (double)(tasks[t].opts.burn + tasks[t].opts.eval); };
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return cost(a) > cost(b); });
    std::vector<std::vector<int>> mine(n_devices);
    std::vector<double> load(n_devices, 0.0);
    for (int t : idx) {
        const int d = (int)(std::min_element(load.begin(), load.end()) - load.begin());
        mine[d].push_back(t); load[d] += cost(t);
    }
    std::vector<std::string> errs(n_devices);
    std::vector<int> rcs(n_devices, SIGNER_OK);
    // small cohorts are launch-latency bound: run several chains per device concurrently (own streams)
    const long long cells = (long long)I * J;
    const int tpd = cells <= 250000 ? 4 : (cells <= 2000000 ? 2 : 1);
    std::vector<std::shared_ptr<DeviceData>> datas(n_devices);
    std::vector<std::atomic<int>> next(n_devices);
    std::vector<std::mutex> mus(n_devices);
    for (int d = 0; d < n_devices; ++d) {
        next[d] = 0;
        if (mine[d].empty()) continue;
        const int rc = upload_data(devices[d], I, J, M, W, datas[d]);
        if (rc != SIGNER_OK) { rcs[d] = rc; errs[d] = g_err; for (int t : mine[d]) tasks[t].status = rc; mine[d].clear(); }
    }
    std::vector<std::thread> th;
    for (int d = 0; d < n_devices; ++d) {
        const int nth = std::min<int>(tpd, (int)mine[d].size());
        for (int w = 0; w < nth; ++w) {
            th.emplace_back([&, d]() {
                for (;;) {
                    const int q = next[d].fetch_add(1);
                    if (q >= (int)mine[d].size()) break;
                    signer_task &tk = tasks[mine[d][q]];
                    signer_chain *c = nullptr;
                    int rc = chain_create(datas[d], tk.K, &tk.state, &tk.opts, &c);
                    if (rc == SIGNER_OK) {
                        signer_outputs out{};
                        out.loglik = tk.loglik; out.bic = tk.bic;
                        rc = chain_run(c, tk.opts.burn, tk.opts.eval, 0, tk.opts.forced_order, &out);
                    }
                    delete c;
                    tk.status = rc;
                    if (rc != SIGNER_OK) {
                        std::lock_guard<std::mutex> lk(mus[d]);
                        if (rcs[d] == SIGNER_OK) { rcs[d] = rc; errs[d] = g_err; }
                    }
                }
            });
        }
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < n_devices; ++d)
        if (rcs[d] != SIGNER_OK) return fail(rcs[d], errs[d]);
    return SIGNER_OK;
}

int signer_gibbs_run_sharded(const signer_problem *pr, const signer_state *st, const signer_opts *o, signer_outputs *out,
                             const int32_t *devices, int32_t n_devices)
{
    g_err.clear();
    if (!pr || !st || !o || !out || !devices) return fail(SIGNER_ERR_ARG, "null argument");
    if (o->Zfixed || o->Thetafixed || o->Afixed)
        return fail(SIGNER_ERR_UNSUPPORTED, "Zfixed/Thetafixed/Afixed=TRUE are not reachable from signeR and not implemented");
    return sharded_run(pr, st, o, out, devices, n_devices);
}

// ---- test hooks
static int test_setup(int device) { CU(cudaSetDevice(device)); return SIGNER_OK; }

int signer_test_math(int32_t fn, const double *x, double *y, int32_t n, int32_t device)
{
    g_err.clear();
    if (!x || !y || n < 0 || fn < 0 || fn > 3) return fail(SIGNER_ERR_ARG, "bad argument");
    ST(test_setup(device));
    double *dx = nullptr, *dy = nullptr;
    CU(cudaMalloc(&dx, sizeof(double) * (size_t)std::max(n, 1))); CU(cudaMalloc(&dy, sizeof(double) * (size_t)std::max(n, 1)));
    CU(cudaMemcpy(dx, x, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    if (n) test_math<<<(n + 127) / 128, 128>>>(fn, dx, dy, n);
    CU(cudaGetLastError());
    CU(cudaMemcpy(y, dy, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dy);
    return SIGNER_OK;
}

int signer_test_binomial(uint64_t seed, uint32_t sweep, uint32_t stream_id, int32_t kstep, const int32_t *n, const double *p,
                         int32_t *out, int32_t count, int32_t device)
{
    g_err.clear();
    if (!n || !p || !out || count < 0) return fail(SIGNER_ERR_ARG, "bad argument");
    ST(test_setup(device));
    int *dn = nullptr, *dout = nullptr; double *dp = nullptr;
    const size_t c = (size_t)std::max(count, 1);
    CU(cudaMalloc(&dn, 4 * c)); CU(cudaMalloc(&dout, 4 * c)); CU(cudaMalloc(&dp, 8 * c));
    CU(cudaMemcpy(dn, n, 4 * (size_t)count, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dp, p, 8 * (size_t)count, cudaMemcpyHostToDevice));
    if (count) test_binomial<<<(count + 127) / 128, 128>>>(Key{(uint32_t)seed, (uint32_t)(seed >> 32)}, sweep, stream_id, kstep, dn, dp, dout, count);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout, 4 * (size_t)count, cudaMemcpyDeviceToHost));
    cudaFree(dn); cudaFree(dout); cudaFree(dp);
    return SIGNER_OK;
}

int signer_test_gamma(uint64_t seed, uint32_t sweep, uint32_t stream_id, const double *shape, const double *rate, double *out,
                      int32_t count, int32_t device)
{
    g_err.clear();
    if (!shape || !rate || !out || count < 0) return fail(SIGNER_ERR_ARG, "bad argument");
    ST(test_setup(device));
    double *ds = nullptr, *dr = nullptr, *dout = nullptr;
    const size_t c = (size_t)std::max(count, 1);
    CU(cudaMalloc(&ds, 8 * c)); CU(cudaMalloc(&dr, 8 * c)); CU(cudaMalloc(&dout, 8 * c));
    CU(cudaMemcpy(ds, shape, 8 * (size_t)count, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dr, rate, 8 * (size_t)count, cudaMemcpyHostToDevice));
    if (count) test_gamma<<<(count + 127) / 128, 128>>>(Key{(uint32_t)seed, (uint32_t)(seed >> 32)}, sweep, stream_id, ds, dr, dout, count);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout, 8 * (size_t)count, cudaMemcpyDeviceToHost));
    cudaFree(ds); cudaFree(dr); cudaFree(dout);
    return SIGNER_OK;
}

void signer_test_perm(uint64_t seed, uint32_t sweep, int32_t order[7])
{
    int o[7];
    host_perm(seed, sweep, o);
    for (int i = 0; i < 7; ++i) order[i] = o[i];
}

void signer_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    const U4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

} // extern "C"

#ifdef SG_Z_TIMING
extern "C" int signer_debug_e_clocks(long long *out, int nblocks)
{
    return (int)cudaMemcpyFromSymbol(out, sg::g_e_clock, sizeof(long long) * 5 * (size_t)nblocks);
}
extern "C" int signer_debug_p_clocks(long long *out, int nblocks)
{
    return (int)cudaMemcpyFromSymbol(out, sg::g_p_clock, sizeof(long long) * 8 * (size_t)nblocks);
}
extern "C" int signer_debug_z_phases(unsigned long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, sg::g_z_phase, sizeof(unsigned long long) * 16);
}
extern "C" int signer_debug_batch_cycles(unsigned *out, int n)
{
    return (int)cudaMemcpyFromSymbol(out, sg::g_z_batch_cycles, sizeof(unsigned) * (size_t)n);
}
#endif
