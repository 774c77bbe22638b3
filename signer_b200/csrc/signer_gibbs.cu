// signer_gibbs.cu -- host side of libsigner_gibbs.so: device-resident chains, the sweep scheduler
// and the C ABI declared in include/signer_gibbs.h.
//
// Replaces the driver GibbsSamplerCpp (src/gibbs_2.cpp:253-424): flag logic (:265-267), the
// random scan (:303-318), sample storage (:320-334) and the returned list (:406-423).
// There is no CPU path in this file: every entry point that computes launches CUDA kernels.
#include "../../include/signer_gibbs.h"
#include "cohort.hpp"
#include "host_util.hpp"
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace sg;
using sgh::DeviceData;
using sgh::DevTemp;
using sgh::GroupPtr;
using sgh::PinnedBuf;
using sgh::TransferPool;
using sgh::dalloc;
using sgh::dfree;
using sgh::fail;
using sgh::g_err;

namespace {

thread_local double g_loop_ms = 0.0;      // wall time of the last sweep loop on this thread (sharded / stateless runs)

constexpr int kMaxRingSets = 8;                     // short runs: one set per chunk, the host never waits for a set to drain
constexpr size_t kRingBytes = size_t(64) << 20;    // per ring set: what is still on the device when the last sweep ends is the call's tail
constexpr size_t kRingTotalBytes = size_t(512) << 20;   // all ring sets of a chain: how far the sweeps may run ahead of the copy-out

// z_alloc_fused and loglik_cols may need more than 48 KB of dynamic shared memory; raise the limits once per device to
// the architectural maximum so that concurrent chains with different K never race on the attribute.
int g_z_dyn_smem_max[64] = {0};

int ensure_device_setup(int device)
{
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) return fail(SIGNER_ERR_ARG, "device ordinal out of range");
    if (done[device]) return SIGNER_OK;
    CU(cudaSetDevice(device));
    ST(sgh::ensure_pool(device));
    int optin = 0;
    CU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, z_alloc_fused<kZMaxWarps>));
    g_z_dyn_smem_max[device] = optin - (int)fa.sharedSizeBytes;
    CU(cudaFuncGetAttributes(&fa, z_alloc_fused<kZFewWarps>));
    g_z_dyn_smem_max[device] = std::min(g_z_dyn_smem_max[device], optin - (int)fa.sharedSizeBytes);
    CU(cudaFuncSetAttribute(z_alloc_fused<kZMaxWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_z_dyn_smem_max[device]));
    CU(cudaFuncSetAttribute(z_alloc_fused<kZFewWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_z_dyn_smem_max[device]));
    CU(cudaFuncGetAttributes(&fa, loglik_cols));          // 8 (129 K + 3264) bytes: above 48 KB from K = 23; 158 KB at K = 128, the largest accepted
    CU(cudaFuncSetAttribute(loglik_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    CU(cudaFuncGetAttributes(&fa, e_family<kEWarps, 3>));    // two W tiles: 51 KB
    CU(cudaFuncSetAttribute(e_family<kEWarps, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    CU(cudaFuncGetAttributes(&fa, e_family<kEWarps, 2>));
    CU(cudaFuncSetAttribute(e_family<kEWarps, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    CU(cudaFuncGetAttributes(&fa, e_family<kEWarpsWide, 1>));
    CU(cudaFuncSetAttribute(e_family<kEWarpsWide, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    CU(cudaFuncGetAttributes(&fa, p_family<2>));          // scratch of the ordered W E' sums grows with the number of genomes
    CU(cudaFuncSetAttribute(p_family<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    CU(cudaFuncGetAttributes(&fa, p_family<32>));
    CU(cudaFuncSetAttribute(p_family<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    CU(cudaFuncGetAttributes(&fa, seg_reduce));
    CU(cudaFuncSetAttribute(seg_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
#ifdef SG_Z_CARVEOUT
    CU(cudaFuncSetAttribute(z_alloc_fused<kZMaxWarps>, cudaFuncAttributePreferredSharedMemoryCarveout, SG_Z_CARVEOUT));
#endif
    done[device] = true;
    return SIGNER_OK;
}

// constant of the data: sum lgamma(M+1) (lgM in R/cpp_eBayesNMF.R:180), once per cohort; a column shard computes it
// with its peers (sharded_run)
std::mutex g_lgm_mu;
int ensure_lgm(DeviceData &d)
{
    std::lock_guard<std::mutex> lk(g_lgm_mu);
    if (d.lgm_ready || d.Jglobal != d.J) return SIGNER_OK;
    CU(cudaSetDevice(d.device));
    ST(ensure_device_setup(d.device));
    DevTemp col;
    ST(col.get((size_t)d.n2 * sizeof(double), nullptr));
    CU(cudaMemsetAsync(col.p, 0, (size_t)d.n2 * sizeof(double), nullptr));
    LLParams lp{d.I, d.J, d.Jp, 1, d.Mrow, d.W, nullptr, nullptr, col.as<double>(), 1, nullptr, 0, nullptr, nullptr, nullptr};
    loglik_cols<<<(d.J + 31) / 32, 32 * kLLWarps, loglik_smem_bytes(1)>>>(lp);
    tree_sum<<<1, 1024>>>(col.as<double>(), d.n2, nullptr, d.lgm);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(nullptr));
    d.lgm_ready = true;
    return SIGNER_OK;
}

// The TMA descriptor of W for e_family (kernels.cuh).  cuTensorMapEncodeTiled is a driver entry point: fetched through the
// runtime, so the library links against libcudart only.
std::mutex g_wmap_mu;
int ensure_wmap(DeviceData &d)
{
    std::lock_guard<std::mutex> lk(g_wmap_mu);
    if (d.wmap_ready) return SIGNER_OK;
    typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(SIGNER_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        encode = (encode_t)fn;
    }
    static_assert(sizeof(CUtensorMap) <= sizeof(d.wmap), "descriptor storage");
    CU(cudaSetDevice(d.device));
    const cuuint64_t dims[2] = {(cuuint64_t)d.Jp, (cuuint64_t)d.I};               // fastest first: genomes, contexts
    const cuuint64_t strides[1] = {(cuuint64_t)d.Jp * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)kEPitch, (cuuint32_t)kERows}, estr[2] = {1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap *>(d.wmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d.W, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SIGNER_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    d.wmap_ready = true;
    return SIGNER_OK;
}

int upload_data(int device, int I, int J, const double *M, const double *W, std::shared_ptr<DeviceData> &out, int col0 = 0,
                int Jglobal = -1)
{
    static_assert(kNDirect == 32, "cohort.cu sorts the work list with the same threshold (kListDirect)");
    ST(sgh::get_cohort(device, I, J, M, W, col0, Jglobal, out));
    return ensure_lgm(*out);
}

// SIGNER_SYNC_LAUNCHES=1 (debugging): wait for every kernel right after its launch and name the one that failed
const bool g_sync_launches = [] { const char *e = std::getenv("SIGNER_SYNC_LAUNCHES"); return e && e[0] == '1'; }();
int check_launch(cudaStream_t st, const char *kernel)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && g_sync_launches) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(SIGNER_ERR_CUDA, std::string(kernel) + ": " + cudaGetErrorString(e));
    return SIGNER_OK;
}

struct Ring {
    double *Ps = nullptr, *Es = nullptr, *Aps = nullptr, *Aes = nullptr, *Bps = nullptr, *Bes = nullptr;
    cudaEvent_t filled = nullptr;
    GroupPtr inflight = sgh::make_group();   // copy-out jobs not yet done
};

} // namespace

struct ShardLink;     // column sharding: this chain is one shard of a chain over several GPUs (see sharded_run)

struct signer_chain {
    std::shared_ptr<DeviceData> data;
    ShardLink *link = nullptr;
    int device = 0, I = 0, J = 0, Jp = 0, K = 0;
    size_t nP = 0, nE = 0, nPp = 0, nEp = 0;   // matrix sizes, and padded to 32 doubles inside the state block
    // one allocation [P | Ap | Bp | E | Ae | Be]: the state goes up and comes back with ONE copy through a pinned buffer
    double *state_block = nullptr;
    PinnedBuf hstate;
    double *P = nullptr, *E = nullptr, *Ap = nullptr, *Bp = nullptr, *Ae = nullptr, *Be = nullptr;
    int *Zin[2] = {nullptr, nullptr}, *Znj[2] = {nullptr, nullptr};
    int zcur = 0;
    double *corrE_part = nullptr;
    int n_seg = 0;
    int shard = 0, n_shards = 1, col0 = 0;   // column sharding (signer_gibbs_run_sharded)
    double *corrE_all = nullptr;             // [n_shards][I*K]: every shard's ordered W E' sum, after the exchange
    double *col = nullptr, *ll_ring = nullptr;
    int ll_cap = 0;
    double *Zcube = nullptr;
    double *stats = nullptr;                 // [8] sufficient statistics of the last keep_par run's samples (EM M-step)
    bool stats_valid = false;
    double stats_np = 0, stats_ne = 0;
    double *sumPs = nullptr, *sumEs = nullptr;   // full-length sample stores for the device-side summaries
    unsigned long long *draw_count = nullptr;    // [2] device counters of step-1 draws (allocated on request)
    double *acache_p = nullptr, *acache_e = nullptr;   // [3][I*K], [3][K*J]: A-only terms of the MH ratios (kernels.cuh, ACache)
    unsigned long long *tile_counter = nullptr;
    unsigned *ll_ticket = nullptr;               // block counter of loglik_cols' fused tree
    unsigned long long z_launches = 0;
    int z_grid = 0, z_warps = 8, z_budget = kZMaxWarps, n_batches = 0, n_chunks = 0, z_warp_smem = 0, z_zin_smem = 0;
    int e_grid = 1, e_warps = kEWarps, e_minb = 3, sms = 1;
    Hyper h{};
    Key key{};
    uint32_t sweep0 = 0;
    int64_t sweeps_done = 0;
    int64_t launches = 0;
    int Pfixed = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    bool borrowed_stream = false;            // `stream` belongs to a peer shard (same-device emulation): never synchronised here
    // the log-likelihood of an evaluation sweep only reads P and E: it runs on its own stream next to the following
    // sweep's z_alloc_fused (which reads them too); the next kernel that writes P or E waits for it
    cudaStream_t ll_stream = nullptr;
    cudaEvent_t ll_ready = nullptr, ll_done = nullptr, cube_done = nullptr;
    bool ll_pending = false;
    Ring ring[kMaxRingSets];
    int ring_cap = 0, ring_keep = 0, ring_sets = 0;
    int last_eval = 0;
    double last_enqueue_ms = 0.0;            // host time the last run spent enqueuing its sweeps (before waiting for the device)
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[4];

    ~signer_chain()
    {
        cudaSetDevice(device);
        if (own_stream) cudaStreamSynchronize(own_stream);
        if (stream && stream != own_stream && !borrowed_stream) cudaStreamSynchronize(stream);   // a caller's stream (set_stream) must outlive the chain
        if (ll_stream) cudaStreamSynchronize(ll_stream);
        cudaStream_t fs = own_stream;
        for (void *p : {(void *)state_block, (void *)corrE_part, (void *)corrE_all, (void *)col, (void *)ll_ring, (void *)Zcube, (void *)stats,
                        (void *)sumPs, (void *)sumEs, (void *)acache_p, (void *)acache_e, (void *)tile_counter, (void *)draw_count, (void *)ll_ticket})
            dfree(p, fs);
        for (int b = 0; b < 2; ++b) { dfree(Zin[b], fs); dfree(Znj[b], fs); }
        for (int b = 0; b < kMaxRingSets; ++b) free_ring(ring[b], fs);
        for (auto &v : ev) for (auto &pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
        if (own_stream) { cudaStreamSynchronize(own_stream); cudaStreamDestroy(own_stream); }
        if (ll_stream) cudaStreamDestroy(ll_stream);
        if (ll_ready) cudaEventDestroy(ll_ready);
        if (ll_done) cudaEventDestroy(ll_done);
        if (cube_done) cudaEventDestroy(cube_done);
    }
    static void free_ring(Ring &r, cudaStream_t fs)
    {
        for (double *p : {r.Ps, r.Es, r.Aps, r.Aes, r.Bps, r.Bes}) dfree(p, fs);
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

// the exchange steps of a column-sharded chain (defined with sharded_run)
template <typename T> int shard_sum(signer_chain *c, T *buf, size_t count);
int shard_gather_corrE(signer_chain *c);
int shard_loglik(signer_chain *c, int mode_lgm, double *out0);
bool is_sharded(const signer_chain *c);

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
        const size_t zsm = (size_t)c->z_warp_smem * c->z_warps + (c->z_zin_smem ? c->nP * sizeof(int) : 0);
        if (c->z_budget == kZFewWarps) z_alloc_fused<kZFewWarps><<<c->z_grid, 32 * c->z_warps, zsm, c->stream>>>(p);
        else z_alloc_fused<kZMaxWarps><<<c->z_grid, 32 * c->z_warps, zsm, c->stream>>>(p);
    }
    ST(check_launch(c->stream, "z_alloc_fused"));
    c->z_launches++;
    c->launches++;
    c->zcur = tgt;
    if (is_sharded(c)) ST(shard_sum<int>(c, c->Zin[tgt], c->nP));       // exchange 1: the context statistics of all shards
    return SIGNER_OK;
}

int launch_p(signer_chain *c, uint32_t sweep, const Ops &ops, double *Ps, double *Aps, double *Bps)
{
    PParams p{};
    p.IK = (int)c->nP; p.P = c->P; p.Ap = c->Ap; p.Bp = c->Bp;
    p.Zin = c->Zin[c->zcur];
    if (c->n_shards > 1) { p.corrE_part = c->corrE_all; p.n_part = c->n_shards; }      // the shards' ordered sums
    else { p.corrE_part = c->corrE_part; p.n_part = c->n_seg; }                          // the segment sums
    p.Ps_slot = Ps; p.Aps_slot = Aps; p.Bps_slot = Bps;
    p.acache = ACache{c->acache_p, c->acache_p + c->nP, c->acache_p + 2 * c->nP};
    p.h = c->h; p.key = c->key; p.sweep = sweep; p.ops = ops;
    ST(wait_loglik(c));
    {
        ProfScope ps(c, 1);
        if (p.IK >= kPWideFrom) p_family<32><<<(p.IK + 32 * kPWarps - 1) / (32 * kPWarps), 32 * kPWarps, p_family_smem_bytes(p.n_part), c->stream>>>(p);
        else p_family<2><<<(p.IK + 2 * kPWarps - 1) / (2 * kPWarps), 32 * kPWarps, p_family_smem_bytes(p.n_part), c->stream>>>(p);
    }
    c->launches++;
    ST(check_launch(c->stream, "p_family"));
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
    std::memcpy(&p.wmap, c->data->wmap, sizeof(CUtensorMap));
    ST(wait_loglik(c));
    {
        ProfScope ps(c, 2);
        if (c->I > kERows) e_family<kEWarpsWide, 1><<<c->e_grid, 32 * c->e_warps, e_smem_bytes(kEWarpsWide), c->stream>>>(p);
        else if (c->e_minb == 2) e_family<kEWarps, 2><<<c->e_grid, 32 * c->e_warps, e_smem_bytes(kEWarps), c->stream>>>(p);
        else e_family<kEWarps, 3><<<c->e_grid, 32 * c->e_warps, e_smem_bytes(kEWarps), c->stream>>>(p);
    }
    c->launches++;
    ST(check_launch(c->stream, "e_family"));
    if (do_corrE && is_sharded(c)) ST(shard_gather_corrE(c));           // exchange 2: every shard's ordered W E' sum
    return SIGNER_OK;
}

// per-genome log-likelihood terms; with `out` the last block also runs the tree (unsharded chains)
int launch_loglik_cols(signer_chain *c, double *out = nullptr, cudaStream_t st = nullptr)
{
    LLParams lp{c->I, c->J, c->Jp, c->K, c->data->Mrow, c->data->W, c->P, c->E, c->col + c->col0, 0,
                out ? c->col : nullptr, c->data->n2, c->data->lgm, out, c->ll_ticket};
    loglik_cols<<<(c->J + 31) / 32, 32 * kLLWarps, loglik_smem_bytes(c->K), st ? st : c->stream>>>(lp);
    c->launches++;
    ST(check_launch(st ? st : c->stream, "loglik_cols"));
    return SIGNER_OK;
}

// log-likelihood of the current state into out[0], on the side stream (see signer_chain::ll_stream)
int launch_loglik(signer_chain *c, double *out)
{
    if (is_sharded(c)) return shard_loglik(c, 0, out);     // on the launching stream: the shards meet in an all-reduce
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

// Sweeps per ring chunk: what fits a ring set, but at least eight chunks per run, so that also a short run (the EM loop's
// 100 evaluation sweeps, the 10 of a smoke-sized call) sends most of its samples home while it is still sweeping
size_t ring_bytes_limit()
{
    if (const char *e = std::getenv("SIGNER_RING_BYTES")) {     // tests: force many small chunks through the copy-out path
        const long long v = std::atoll(e);
        if (v > 0) return (size_t)v;
    }
    return kRingBytes;
}

int ring_chunk(const signer_chain *c, int eval, int keep_par)
{
    const size_t per_sweep = sizeof(double) * (c->nP + c->nE) * (keep_par ? 3 : 2);
    const size_t fit = std::max<size_t>(1, std::min<size_t>((size_t)eval, ring_bytes_limit() / per_sweep));
    return (int)std::max<size_t>(1, std::min<size_t>(fit, ((size_t)eval + 7) / 8));
}

// Ring sets: as many as kRingTotalBytes hold (two at least, kMaxRingSets at most, never more than the run has chunks).  With
// keep_par a config-3 sweep leaves 4.8 MB of samples, 24 GB/s at 0.2 ms per sweep -- about what the copy-out path sustains into
// fresh pageable memory -- so the EM loop's call runs ahead of its copy-out most of the time; with two sets its enqueue waited
// for drains and the device idled (measured: 100 - 107 ms per call against 91 with eight)
int ring_set_count(const signer_chain *c, int eval, int keep_par)
{
    const size_t per_sweep = sizeof(double) * (c->nP + c->nE) * (keep_par ? 3 : 2);
    const int cap = ring_chunk(c, eval, keep_par), chunks = (eval + cap - 1) / cap;
    const size_t total = std::getenv("SIGNER_RING_BYTES") ? 2 * ring_bytes_limit() : kRingTotalBytes;   // (tests: two small sets, reused many times)
    const size_t fit = total / std::max<size_t>(1, per_sweep * (size_t)cap);
    return (int)std::max<size_t>(2, std::min<size_t>({(size_t)kMaxRingSets, fit, (size_t)chunks}));
}

int ensure_rings(signer_chain *c, int eval, int keep_par)
{
    const int cap = ring_chunk(c, eval, keep_par), sets = ring_set_count(c, eval, keep_par);
    if (c->ring_cap >= cap && c->ring_keep >= keep_par && c->ring_sets >= sets) return SIGNER_OK;
    for (int b = 0; b < kMaxRingSets; ++b) signer_chain::free_ring(c->ring[b], c->stream);
    for (int b = 0; b < sets; ++b) {
        Ring &r = c->ring[b];
        ST(dalloc(&r.Ps, c->nP * cap, c->stream));
        ST(dalloc(&r.Es, c->nE * cap, c->stream));
        ST(dalloc(&r.Bps, c->nP * cap, c->stream));
        ST(dalloc(&r.Bes, c->nE * cap, c->stream));
        if (keep_par) {
            ST(dalloc(&r.Aps, c->nP * cap, c->stream));
            ST(dalloc(&r.Aes, c->nE * cap, c->stream));
        }
        CU(cudaEventCreateWithFlags(&r.filled, cudaEventDisableTiming));
    }
    c->ring_cap = cap; c->ring_keep = keep_par; c->ring_sets = sets;
    return SIGNER_OK;
}

// copy chunk `chunk` (sweeps r0 .. r0+cnt-1 of the eval phase) from its ring set to the host: posted to the
// device's transfer pool, which waits for the ring set's `filled` event
int flush_chunk(signer_chain *c, int chunk, int r0, int cnt, int keep_par, signer_outputs *out)
{
    Ring &r = c->ring[chunk % c->ring_sets];
    TransferPool &pool = TransferPool::of(c->device);
    auto cp = [&](double *dst, const double *src, size_t per) {
        if (!dst || !src) return;
        pool.post_d2h(src, dst + per * (size_t)r0, sizeof(double) * per * (size_t)cnt, r.filled, r.inflight);
    };
    // a column shard: its genomes are a contiguous block of every K x J slice of the caller's cube
    auto cp_shard = [&](double *dst, const double *src) {
        if (!dst || !src) return;
        const size_t slice = (size_t)c->K * c->data->Jglobal, off = (size_t)c->K * c->col0;
        for (int s = 0; s < cnt; ++s)
            pool.post_d2h(src + c->nE * (size_t)s, dst + slice * (size_t)(r0 + s) + off, sizeof(double) * c->nE, r.filled, r.inflight);
    };
    const bool shard = is_sharded(c);
    if (!c->sumPs) { cp(out->Ps, r.Ps, c->nP); if (shard) cp_shard(out->Es, r.Es); else cp(out->Es, r.Es, c->nE); }   // else: from the full-length stores at the end
    cp(out->Bps, r.Bps, c->nP);
    if (shard) cp_shard(out->Bes, r.Bes); else cp(out->Bes, r.Bes, c->nE);
    if (keep_par) { cp(out->Aps, r.Aps, c->nP); if (shard) cp_shard(out->Aes, r.Aes); else cp(out->Aes, r.Aes, c->nE); }
    return SIGNER_OK;
}

// every job of the group has finished; a failure is reported once and the group starts afresh
int wait_group(int device, GroupPtr &g, const char *what)
{
    if (TransferPool::of(device).wait(g)) return SIGNER_OK;
    const std::string err = g->err;
    g = sgh::make_group();
    return fail(SIGNER_ERR_CUDA, std::string(what) + ": " + err);
}

// the ring set's samples have all reached the caller's arrays (it may be refilled)
int wait_ring(signer_chain *c, Ring &r) { return wait_group(c->device, r.inflight, "sample copy-out"); }

// Phat / Ehat from the full-length sample stores (see include/signer_gibbs.h)
int device_summaries(signer_chain *c, int R, double *Phat, double *Ehat)
{
    DevTemp s, T, med;
    const size_t nmax = std::max(c->nP, c->nE);
    ST(s.get(sizeof(double) * (size_t)R * c->K, c->stream));
    ST(T.get(sizeof(double) * nmax * (size_t)R, c->stream));
    ST(med.get(sizeof(double) * nmax, c->stream));
    sig_sums<<<(R * c->K + 127) / 128, 128, 0, c->stream>>>(c->sumPs, c->I, c->K, R, s.as<double>());
    for (int mode = 0; mode < 2; ++mode) {
        const size_t n = mode == 0 ? c->nP : c->nE;
        double *dst = mode == 0 ? Phat : Ehat;
        if (!dst) continue;
        norm_transpose<<<dim3((unsigned)((n + 31) / 32), (unsigned)((R + 31) / 32)), dim3(32, 8), 0, c->stream>>>(
            mode == 0 ? c->sumPs : c->sumEs, s.as<double>(), (int)n, R, c->I, c->K, mode, T.as<double>());
        median_rows<<<(unsigned)((n + 7) / 8), 256, 0, c->stream>>>(T.as<double>(), (int)n, R, med.as<double>());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(dst, med.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->launches += 2;
    }
    c->launches += 1;
    return SIGNER_OK;
}

// the six state matrices come back with one copy of the state block through the chain's pinned buffer
// (two halves, so that a run can put the copy on the stream behind its last sweep and collect it after its samples have left)
int fetch_state_begin(signer_chain *c, const signer_outputs *out)
{
    CU(cudaSetDevice(c->device));
    const bool any_state = out->P || out->Ap || out->Bp || out->E || out->Ae || out->Be;
    const size_t sd = 3 * (c->nPp + c->nEp);
    const size_t need = sd * sizeof(double) + (c->nP + c->nE) * sizeof(int);
    if (!c->hstate.get(need)) return fail(SIGNER_ERR_NOMEM, "pinned staging buffer");
    double *hs = c->hstate.as<double>();
    int *hz = reinterpret_cast<int *>(hs + sd);
    if (any_state) CU(cudaMemcpyAsync(hs, c->state_block, sd * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (out->Zin) CU(cudaMemcpyAsync(hz, c->Zin[c->zcur], c->nP * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (out->Znj) CU(cudaMemcpyAsync(hz + c->nP, c->Znj[c->zcur], c->nE * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return SIGNER_OK;
}

int fetch_state_end(signer_chain *c, signer_outputs *out)
{
    const size_t sd = 3 * (c->nPp + c->nEp);
    const double *hs = c->hstate.as<double>();
    const int *hz = reinterpret_cast<const int *>(hs + sd);
    CU(cudaStreamSynchronize(c->stream));
    auto cp = [&](double *dst, size_t off, size_t n) { if (dst) std::memcpy(dst, hs + off, n * sizeof(double)); };
    cp(out->P, 0, c->nP); cp(out->Ap, c->nPp, c->nP); cp(out->Bp, 2 * c->nPp, c->nP);
    cp(out->E, 3 * c->nPp, c->nE); cp(out->Ae, 3 * c->nPp + c->nEp, c->nE); cp(out->Be, 3 * c->nPp + 2 * c->nEp, c->nE);
    if (out->Zin) std::memcpy(out->Zin, hz, c->nP * sizeof(int));
    if (out->Znj) std::memcpy(out->Znj, hz + c->nP, c->nE * sizeof(int));
    return SIGNER_OK;
}

int fetch_state(signer_chain *c, signer_outputs *out)
{
    ST(fetch_state_begin(c, out));
    return fetch_state_end(c, out);
}

// Reduce the given I x J x K cube (pageable host memory; columns [c0, c0+Jl) of every slice) to the chain's
// statistics Zin[0], Znj[0] (zeroed on c->stream before): pieces of at most 4 MB go through the transfer pool, each
// worker hands its piece to zslab_reduce on its own stream.
int reduce_cube(signer_chain *c, const double *Z, int Jfull, int c0)
{
    cudaEvent_t ready = nullptr;
    CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    struct EventGuard { cudaEvent_t e; ~EventGuard() { cudaEventDestroy(e); } } eg{ready};
    CU(cudaEventRecord(ready, c->stream));
    TransferPool &pool = TransferPool::of(c->device);
    GroupPtr g = sgh::make_group();
    const int I = c->I, Jl = c->J, K = c->K;
    const int nj_max = (int)std::max<size_t>(1, std::min<size_t>((size_t)Jl, (size_t(4) << 20) / ((size_t)I * sizeof(double))));
    int *Zin = c->Zin[0], *Znj = c->Znj[0];
    const int use_smem = ((size_t)I * sizeof(int) <= 32 * 1024) ? 1 : 0;
    for (int k = 0; k < K; ++k)
        for (int ja = 0; ja < Jl; ja += nj_max) {
            const int nj = std::min(nj_max, Jl - ja);
            const double *src = Z + (size_t)I * Jfull * k + (size_t)I * (c0 + ja);
            pool.post_h2d_hook(src, (size_t)I * nj * sizeof(double), ready, g,
                               [=](const void *dev, size_t, cudaStream_t st) -> cudaError_t {
                                   zslab_reduce<<<(nj + 63) / 64, 256, use_smem ? (size_t)I * sizeof(int) : 0, st>>>(
                                       static_cast<const double *>(dev), I, nj, ja, K, k, use_smem, Zin, Znj);
                                   return cudaGetLastError();
                               });
        }
    c->launches += (int64_t)K * ((Jl + nj_max - 1) / nj_max);
    return wait_group(c->device, g, "upload of the Z cube");
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
    ST(ensure_device_setup(data->device));
    ST(ensure_wmap(*data));
    std::unique_ptr<signer_chain> c(new signer_chain());
    c->data = data; c->device = data->device; c->I = data->I; c->J = data->J; c->Jp = data->Jp; c->K = K;
    c->nP = (size_t)c->I * K; c->nE = (size_t)K * c->J;
    c->nPp = (c->nP + 31) / 32 * 32; c->nEp = (c->nE + 31) / 32 * 32;
    c->h = Hyper{o->ap, o->bp, o->ae, o->be, o->lp, o->le, o->var_ap, o->var_ae};
    c->key = Key{(uint32_t)o->seed, (uint32_t)(o->seed >> 32)};
    c->sweep0 = o->sweep0; c->Pfixed = o->Pfixed ? 1 : 0;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->ll_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ll_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ll_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->cube_done, cudaEventDisableTiming));
    c->stream = c->own_stream;
    cudaStream_t s = c->stream;
    // the state: gathered into the pinned buffer, one copy up
    {
        const size_t sd = 3 * (c->nPp + c->nEp);
        if (!c->hstate.get(sd * sizeof(double) + (c->nP + c->nE) * sizeof(int))) return fail(SIGNER_ERR_NOMEM, "pinned staging buffer");
        double *hs = c->hstate.as<double>();
        std::memcpy(hs, st->P, c->nP * 8); std::memcpy(hs + c->nPp, st->Ap, c->nP * 8); std::memcpy(hs + 2 * c->nPp, st->Bp, c->nP * 8);
        std::memcpy(hs + 3 * c->nPp, st->E, c->nE * 8); std::memcpy(hs + 3 * c->nPp + c->nEp, st->Ae, c->nE * 8);
        std::memcpy(hs + 3 * c->nPp + 2 * c->nEp, st->Be, c->nE * 8);
        ST(dalloc(&c->state_block, sd, s));
        CU(cudaMemcpyAsync(c->state_block, hs, sd * sizeof(double), cudaMemcpyHostToDevice, s));
        c->P = c->state_block; c->Ap = c->P + c->nPp; c->Bp = c->Ap + c->nPp;
        c->E = c->Bp + c->nPp; c->Ae = c->E + c->nEp; c->Be = c->Ae + c->nEp;
    }
    for (int b = 0; b < 2; ++b) {
        ST(dalloc(&c->Zin[b], c->nP, s));
        ST(dalloc(&c->Znj[b], c->nE, s));
        CU(cudaMemsetAsync(c->Zin[b], 0, c->nP * sizeof(int), s));
        CU(cudaMemsetAsync(c->Znj[b], 0, c->nE * sizeof(int), s));
    }
    c->n_seg = (c->J + kSeg - 1) / kSeg;
    ST(dalloc(&c->corrE_part, (size_t)c->n_seg * c->nP, s));
    ST(dalloc(&c->col, (size_t)data->n2, s));
    CU(cudaMemsetAsync(c->col, 0, sizeof(double) * (size_t)data->n2, s));
    ST(dalloc(&c->tile_counter, 1, s));
    CU(cudaMemsetAsync(c->tile_counter, 0, sizeof(unsigned long long), s));
    ST(dalloc(&c->ll_ticket, 1, s));
    CU(cudaMemsetAsync(c->ll_ticket, 0, sizeof(unsigned), s));
    ST(dalloc(&c->stats, 8, s));
    ST(dalloc(&c->acache_p, 3 * c->nP, s));
    ST(dalloc(&c->acache_e, 3 * c->nE, s));
    ST(rebuild_a_cache(c.get()));

    // launch geometry
    c->n_batches = (data->n_cells + 31) / 32;
    c->n_chunks = (c->n_batches + kZChunk - 1) / kZChunk;
    c->z_warp_smem = (int)z_smem_bytes_per_warp(K);
    c->z_zin_smem = (c->nP * sizeof(int) <= (size_t)kZinSmemMax) ? 1 : 0;
    // block size: the configuration that keeps the most warps resident per SM (at most kZMaxWarps: the register
    // file holds that many at the kernel's 80 registers); ties go to fewer, larger blocks (fewer Zin accumulators
    // in shared memory, which leaves more of the SM's 256 KB to L1)
    int sms = 0, best_w = 0, best_b = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    // Block width and register budget: the candidate with the smallest estimated launch time.  Measured on configs 3 and 5
    // over K = 5..40 (profiles/r02_measurements.md section 5): at 80 registers a launch takes 1 + 0.23 (24 - warps) / 8 of its
    // time at 24 warps per SM; the 128-register build at 16 warps 1.046 of that; and what the cumulative sums leave of the SM's
    // 256 KB for L1 matters to the 80-register build, whose spills (128 bytes x 768 threads) live there next to the gathers of E
    // columns and of the work list -- the shared-memory carve-out steps are ..., 164, 196, 228 KB: 60 KB of L1 cost it ~7 %,
    // 28 KB ~20 %; the 128-register build (48 bytes of stack) showed no such step up to K = 40.
    auto l1_factor = [&](int w, int b) {
        const size_t per_block = (size_t)c->z_warp_smem * w + (c->z_zin_smem ? c->nP * sizeof(int) : 0) + 2048;   // + static + reserved
        const size_t kb = (per_block * (size_t)b + 1023) / 1024;
        size_t carve = 228;
        for (size_t step : {0, 8, 16, 32, 64, 100, 132, 164, 196, 228}) if (step >= kb) { carve = step; break; }
        const size_t l1 = 256 - carve;
        return l1 >= 92 ? 1.0 : l1 >= 60 ? 1.075 : 1.20;
    };
    int forced_budget = 0;
    if (const char *e = std::getenv("SIGNER_Z_BUDGET")) forced_budget = std::atoi(e);       // experiments: force the register budget
    double best_est = 1e30;
    for (int budget : {kZMaxWarps, kZFewWarps}) {
        if (forced_budget && budget != forced_budget) continue;
        for (int w : {24, 16, 12, 8, 6, 4, 3, 2, 1}) {
            if (w > budget) continue;
            const size_t zsm = (size_t)c->z_warp_smem * w + (c->z_zin_smem ? c->nP * sizeof(int) : 0);
            if (zsm > (size_t)g_z_dyn_smem_max[c->device]) continue;
            int occ = 0;
            if (budget == kZFewWarps) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, z_alloc_fused<kZFewWarps>, 32 * w, zsm));
            else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, z_alloc_fused<kZMaxWarps>, 32 * w, zsm));
            occ = std::min(occ, budget / w);
            if (occ < 1) continue;
            const double est = (1.0 + 0.23 * (budget - occ * w) / 8.0) * (budget == kZFewWarps ? 1.046 : l1_factor(w, occ));
            if (est < best_est - 1e-9) { best_est = est; best_w = w; best_b = occ; c->z_budget = budget; }   // ties: fewer, larger blocks
        }
    }
    if (best_w == 0) return fail(SIGNER_ERR_ARG, "z_alloc_fused: K too large for shared memory");
    if (const char *e = std::getenv("SIGNER_Z_BLOCK")) {          // experiments: "warps:blocks_per_sm"
        int w = 0, b = 0;
        if (std::sscanf(e, "%d:%d", &w, &b) == 2 && w >= 1 && w * b <= c->z_budget && b >= 1) { best_w = w; best_b = b; }
    }
    c->z_warps = best_w;
    c->z_grid = std::max(1, std::min((c->n_chunks + best_w - 1) / best_w, best_b * sms));
    // e_family: every block takes a contiguous range of warp tasks (column group x class pair); all blocks resident in one wave
    c->sms = sms;
    c->e_warps = e_block_warps(c->I, K);
    {
        int per_sm = 0;                                   // resident blocks of this width (two 26 KB tiles each)
        const int n_tasks = e_groups(c->J) * e_pairs(K);
        if (c->I > kERows) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, e_family<kEWarpsWide, 1>, 32 * c->e_warps, e_smem_bytes(kEWarpsWide)));
        else {
            // the register budget: rounds needed x the time of a full round, measured on config 3 (profiles/r02_measurements.md)
            int per3 = 0, per2 = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per3, e_family<kEWarps, 3>, 32 * c->e_warps, e_smem_bytes(kEWarps)));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per2, e_family<kEWarps, 2>, 32 * c->e_warps, e_smem_bytes(kEWarps)));
            const long long cap3 = (long long)std::max(1, per3) * sms * c->e_warps, cap2 = (long long)std::max(1, per2) * sms * c->e_warps;
            const double est3 = (double)((n_tasks + cap3 - 1) / cap3) * 64.0, est2 = (double)((n_tasks + cap2 - 1) / cap2) * 42.0;
            c->e_minb = est2 <= est3 ? 2 : 3;
            if (const char *e = std::getenv("SIGNER_E_MINB")) { const int v = std::atoi(e); if (v == 2 || v == 3) c->e_minb = v; }   // experiments
            per_sm = c->e_minb == 2 ? per2 : per3;
        }
        const int want = c->I > kERows ? e_groups(c->J) : (n_tasks + c->e_warps - 1) / c->e_warps;   // chunked contexts: whole column groups per block
        c->e_grid = std::max(1, std::min(std::max(1, per_sm) * sms, want));   // one round per block where they all fit
    }
    if (p_family_smem_bytes(c->n_seg) > (size_t)g_z_dyn_smem_max[c->device])
        return fail(SIGNER_ERR_ARG, "too many genomes for one device: the W E' sums do not fit the reduction's shared memory; shard the columns");

    // initial sufficient statistics: reduce the given cube, or draw the first allocation on the device
    if (deferred_init) {                                  // a column shard: its peers take part (sharded_run)
        CU(cudaStreamSynchronize(s));
        *out = c.release();
        return SIGNER_OK;
    }
    c->zcur = 0;
    if (st->Z) ST(reduce_cube(c.get(), st->Z, c->J, 0));
    else ST(launch_z(c.get(), c->sweep0, kStZinit, nullptr));
    if (!c->Pfixed) {   // W E' of the initial E, in case step 2 comes before step 3
        Ops none{0, 0u};
        ST(launch_e(c.get(), c->sweep0, none, 1, nullptr, nullptr, nullptr));
    }
    CU(cudaStreamSynchronize(s));                         // the pinned state buffer may be reused from here on
    *out = c.release();
    return SIGNER_OK;
}

int chain_run(signer_chain *c, int burn, int eval, int keep_par, const int32_t *forced, signer_outputs *out)
{
    if (burn < 0 || eval < 0) return fail(SIGNER_ERR_ARG, "burn and eval must be non-negative");
    CU(cudaSetDevice(c->device));
    keep_par = keep_par ? 1 : 0;
    static const bool trace = [] { const char *e = std::getenv("SIGNER_TRACE"); return e && e[0] == '2'; }();   // SIGNER_TRACE=2: phases of this function
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
    double tr[6] = {0, 0, 0, 0, 0, 0};
    // Sample rings: needed when samples leave for the host, and for a run that keeps everything on the device (out == NULL:
    // the EM loop reads their statistics).  A caller that asks only for log-likelihoods / BICs (the model-selection path,
    // signer_gibbs_batch) gets no ring stores at all.
    const bool want_rings = eval > 0 && (!out || out->Ps || out->Es || out->Bps || out->Bes || (keep_par && (out->Aps || out->Aes)));
    if (eval > 0) {
        if (want_rings) ST(ensure_rings(c, eval, keep_par));
        if (c->ll_cap < eval) {
            dfree(c->ll_ring, c->stream); c->ll_ring = nullptr; c->ll_cap = 0;
            ST(dalloc(&c->ll_ring, (size_t)eval, c->stream));
            c->ll_cap = eval;
        }
    }
    double *zcube = nullptr;
    if (keep_par && out && out->Zs && eval > 0) {
        if (!c->Zcube) ST(dalloc(&c->Zcube, (size_t)c->I * c->J * c->K, c->stream));
        zcube = c->Zcube;
    }
    const bool summ = out && eval > 0 && (out->Phat || out->Ehat);
    if (summ) {                                           // full-length stores for Ps and Es
        dfree(c->sumPs, c->stream); dfree(c->sumEs, c->stream); c->sumPs = c->sumEs = nullptr;
        ST(dalloc(&c->sumPs, c->nP * (size_t)eval, c->stream));
        ST(dalloc(&c->sumEs, c->nE * (size_t)eval, c->stream));
    }
    const bool do_stats = want_rings && keep_par;         // EM M-step statistics of (Aps, Bps), (Aes, Bes), accumulated per chunk
    if (eval > 0) {
        c->stats_valid = false;
        if (do_stats) CU(cudaMemsetAsync(c->stats, 0, 8 * sizeof(double), c->stream));
    }
    const int total = burn + eval, cap = std::max(1, want_rings ? ring_chunk(c, eval, keep_par) : eval);
    int chunk_r0 = 0;
    std::vector<std::thread> faulters;
    auto join_faulters = [&]() { for (auto &t : faulters) if (t.joinable()) t.join(); faulters.clear(); };
    if (out && eval > 0 && (!is_sharded(c) || c->shard == 0)) {      // (the first shard touches the caller's pages for all)
        const size_t R = (size_t)eval;
        double *bufs[7] = {out->Es, out->Bes, keep_par ? out->Aes : nullptr, out->Ps, out->Bps, keep_par ? out->Aps : nullptr, zcube ? out->Zs : nullptr};
        const size_t nEg = (size_t)c->K * c->data->Jglobal;                  // the caller's arrays hold the whole cohort
        const size_t lens[7] = {nEg * R, nEg * R, nEg * R, c->nP * R, c->nP * R, c->nP * R, (size_t)c->I * c->data->Jglobal * c->K};
        // large arrays are touched by several threads each (page faults scale across threads); the touch is an atomic
        // no-op, so it can never undo a copy-out that has already landed (host_util.cu)
        const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
        const size_t max_threads = std::min<size_t>(32, hw / 2);
        size_t big_bytes = 0;
        for (int b = 0; b < 7; ++b) if (bufs[b] && lens[b] * sizeof(double) >= (size_t(8) << 20)) big_bytes += lens[b] * sizeof(double);
        for (int b = 0; b < 7; ++b) {
            if (!bufs[b] || lens[b] * sizeof(double) < (size_t(8) << 20)) continue;
            const size_t parts = std::max<size_t>(1, std::min<size_t>(max_threads, max_threads * (lens[b] * sizeof(double)) / big_bytes));
            const size_t step = ((lens[b] + parts - 1) / parts + 511) / 512 * 512;       // whole pages
            for (size_t off = 0; off < lens[b]; off += step)
                faulters.emplace_back(sgh::prefault, (void *)(bufs[b] + off), std::min(step, lens[b] - off) * sizeof(double));
        }
    }
    struct Joiner { std::function<void()> f; ~Joiner() { f(); } } joiner{join_faulters};
    // whatever way this call ends, no copy-out job may outlive it: they write into the caller's arrays
    Joiner drain{[c, out]() { if (out) { TransferPool &p = TransferPool::of(c->device); for (int b = 0; b < kMaxRingSets; ++b) p.wait(c->ring[b].inflight); } }};
    GroupPtr cube_group = sgh::make_group();
    Joiner drain_cube{[c, &cube_group]() { TransferPool::of(c->device).wait(cube_group); }};
    const auto t_enqueue = std::chrono::steady_clock::now();
    tr[0] = since();
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
            Ring &rg = c->ring[want_rings ? chunk % c->ring_sets : 0];
            Slots s;
            if (want_rings) {
                if (slot == 0 && chunk >= c->ring_sets && out) ST(wait_ring(c, rg));      // the set's previous chunk has left the device
                s.Ps = rg.Ps + c->nP * slot; s.Es = rg.Es + c->nE * slot;
                s.Bps = rg.Bps + c->nP * slot; s.Bes = rg.Bes + c->nE * slot;
                if (keep_par) { s.Aps = rg.Aps + c->nP * slot; s.Aes = rg.Aes + c->nE * slot; }
            }
            if (summ) { s.Ps = c->sumPs + c->nP * (size_t)r; s.Es = c->sumEs + c->nE * (size_t)r; }
            const bool last = (k == total - 1);
            // Zs keeps only the latest Z (src/gibbs_2.cpp:323): store the cube in the last sweep only
            ST(enqueue_sweep(c, sweep, order, &s, last ? zcube : nullptr));
            ST(launch_loglik(c, c->ll_ring + r));
            if (want_rings && (slot == cap - 1 || last)) {
                if (do_stats) {
                    stats_reduce<<<256, 256, 0, c->stream>>>(rg.Aps, rg.Bps, c->nP * (size_t)(slot + 1), c->stats);
                    stats_reduce<<<1024, 256, 0, c->stream>>>(rg.Aes, rg.Bes, c->nE * (size_t)(slot + 1), c->stats + 4);
                    ST(check_launch(c->stream, "stats_reduce"));
                    c->launches += 2;
                }
                CU(cudaEventRecord(rg.filled, c->stream));
                if (out) ST(flush_chunk(c, chunk, chunk_r0, slot + 1, keep_par, out));
                chunk_r0 += cap;
            }
        }
        c->sweeps_done++;
    }
    ST(wait_loglik(c));                                    // the launching stream ends after the last log-likelihood
    tr[1] = since();
    c->last_enqueue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enqueue).count();
    c->last_eval = eval;
    if (eval > 0 && do_stats) { c->stats_valid = true; c->stats_np = (double)c->nP * eval; c->stats_ne = (double)c->nE * eval; }
    if (out) ST(fetch_state_begin(c, out));                // behind the last sweep on the stream; collected below
    if (out && eval > 0) {
        if (zcube && out->Zs) {
            // step 1 may not have run in the last sweep only under a forced order; then Zs is undefined.  The cube leaves
            // through the transfer pool like the samples (pageable destination).
            CU(cudaEventRecord(c->cube_done, c->stream));
            TransferPool &pool = TransferPool::of(c->device);
            if (!is_sharded(c)) pool.post_d2h(zcube, out->Zs, sizeof(double) * (size_t)c->I * c->J * c->K, c->cube_done, cube_group);
            else                                           // a shard's block of every I x J slice
                for (int k = 0; k < c->K; ++k)
                    pool.post_d2h(zcube + (size_t)c->I * c->J * k, out->Zs + (size_t)c->I * c->data->Jglobal * k + (size_t)c->I * c->col0,
                                  sizeof(double) * (size_t)c->I * c->J, c->cube_done, cube_group);
        }
        CU(cudaStreamSynchronize(c->stream));
        tr[2] = since();
        if (want_rings) for (int b = 0; b < c->ring_sets; ++b) ST(wait_ring(c, c->ring[b]));
        ST(wait_group(c->device, cube_group, "copy-out of the Z cube"));
        tr[3] = since();
        join_faulters();
        tr[4] = since();
        if (summ) {
            if (out->Ps) CU(cudaMemcpy(out->Ps, c->sumPs, sizeof(double) * c->nP * (size_t)eval, cudaMemcpyDeviceToHost));
            if (out->Es) CU(cudaMemcpy(out->Es, c->sumEs, sizeof(double) * c->nE * (size_t)eval, cudaMemcpyDeviceToHost));
            ST(device_summaries(c, eval, out->Phat, out->Ehat));
            dfree(c->sumPs, c->stream); dfree(c->sumEs, c->stream); c->sumPs = c->sumEs = nullptr;
        }
        if (out->loglik || out->bic) {
            std::vector<double> ll((size_t)eval);
            CU(cudaMemcpy(ll.data(), c->ll_ring, sizeof(double) * (size_t)eval, cudaMemcpyDeviceToHost));
            const double Jg = (double)c->data->Jglobal;                                          // (a shard: the whole cohort's genomes)
            const double pen = (double)c->K * ((double)c->I + Jg) * std::log(Jg);               // R/cpp_eBayesNMF.R:186
            for (int r = 0; r < eval; ++r) {
                if (out->loglik) out->loglik[r] = ll[r];
                if (out->bic) out->bic[r] = 2.0 * ll[r] - pen;
            }
        }
    }
    if (out) ST(fetch_state_end(c, out));
    tr[5] = since();
    if (trace)
        std::fprintf(stderr, "[chain_run] set-up %.2f ms, enqueue %.2f, device done +%.2f, copy-out drained +%.2f, page touchers joined +%.2f, "
                     "log-likelihoods + state back +%.2f (total %.2f ms)\n", tr[0], tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], tr[5]);
    return SIGNER_OK;
}

// ------------------------------------------------------------------------------------------------
// Genome-column sharding: one chain over several GPUs.  Shard g owns the 32-genome
// segments [g*n_seg/N, (g+1)*n_seg/N) -- M, W, E, Ae, Be, Znj for those genomes; P, Ap, Bp are
// replicated and updated redundantly with identical Philox counters, so they stay identical.
// Every shard is an ordinary chain with a ShardLink and runs the ordinary sweep loop (chain_run) on its
// OWN host thread, so the devices are fed in parallel; the path's real exchange steps hang off the
// kernels that produce their inputs (launch_z, launch_e, launch_loglik):
//   * after step 1: ncclAllReduce (sum, int32) of the I*K context statistics Zin -- exact in any order;
//   * after step 3: every shard adds its own W E' segment sums in the specified order (seg_reduce) and
//     one ncclAllGather hands every shard all N ordered sums (I*K doubles each), which p_family then adds
//     in shard order;
//   * per evaluation sweep: an all-reduce of the per-genome log-likelihood terms (disjoint supports),
//     then the tree on shard 0.
// Results are bit-identical to the oracle run with nshards = N.  NCCL is loaded lazily; shards placed on
// ONE device (tests on a single GPU) share one stream and emulate the collectives with a summing kernel
// between two host-side barriers.
// ------------------------------------------------------------------------------------------------
typedef void *nccl_comm_t;
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(nccl_comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
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
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllReduce && AllGather;
    }
};
constexpr int kNcclInt32 = 2, kNcclFloat64 = 8, kNcclSum = 0;   // nccl.h: ncclInt32, ncclFloat64, ncclSum

// the shards' host threads meet here (emulated collectives only); wait() returns false once any of them has failed
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n = 1, waiting = 0;
    unsigned long gen = 0;
    bool aborted = false;
    bool wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return false;
        const unsigned long g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return true; }
        cv.wait(lk, [&] { return gen != g || aborted; });
        return !aborted;
    }
    void abort()
    {
        std::lock_guard<std::mutex> lk(mu);
        aborted = true;
        cv.notify_all();
    }
};

struct ShardShared {
    int n = 0;
    bool emulated = true;
    NcclApi nccl;
    std::vector<nccl_comm_t> comms;
    HostBarrier bar;
    std::vector<void *> bufs;        // emulated collectives: this round's buffer of every shard
    void **ptr_tab = nullptr;        // ... and their device table
    bool keep_comms = false;         // the communicators belong to the process-wide cache
    HostBarrier start;               // the shards' threads line up before the sweep loop (timing)
    double loop_ms = 0.0;            // shard 0's sweep loop
    ~ShardShared()
    {
        if (!keep_comms) for (auto cm : comms) if (cm && nccl.CommDestroy) nccl.CommDestroy(cm);
        if (ptr_tab) cudaFree(ptr_tab);
    }
};

} // namespace

struct ShardLink {
    int rank = 0, n = 1;
    ShardShared *sh = nullptr;
};

namespace {

bool is_sharded(const signer_chain *c) { return c->link && c->link->n > 1; }

int nccl_fail(ShardShared *S, const char *what, int rc)
{
    return fail(SIGNER_ERR_NCCL, std::string(what) + ": " + (S->nccl.GetErrorString ? S->nccl.GetErrorString(rc) : "error"));
}

// every shard's buf <- sum over the shards, in place (the producers are already enqueued on the shards' streams)
template <typename T>
int shard_sum(signer_chain *c, T *buf, size_t count)
{
    ShardLink *L = c->link;
    if (!L || L->n == 1) return SIGNER_OK;
    ShardShared *S = L->sh;
    if (!S->emulated) {
        const int rc = S->nccl.AllReduce(buf, buf, count, sizeof(T) == 4 ? kNcclInt32 : kNcclFloat64, kNcclSum, S->comms[L->rank], c->stream);
        if (rc) return nccl_fail(S, "ncclAllReduce", rc);
        return SIGNER_OK;
    }
    S->bufs[L->rank] = buf;
    if (!S->bar.wait()) return fail(SIGNER_ERR_CUDA, "a peer shard failed");
    int rc = SIGNER_OK;
    if (L->rank == 0) {                                   // one stream carries all emulated shards: plain ordering
        cudaError_t e = cudaMemcpyAsync(S->ptr_tab, S->bufs.data(), sizeof(void *) * S->n, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) {
            sum_shards<T><<<(unsigned)std::min<size_t>((count + 255) / 256, 1024), 256, 0, c->stream>>>(reinterpret_cast<T *const *>(S->ptr_tab), S->n, count);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);    // the host-side pointer table is reused by the next round
        if (e != cudaSuccess) { rc = fail(SIGNER_ERR_CUDA, std::string("emulated all-reduce: ") + cudaGetErrorString(e)); S->bar.abort(); }
    }
    if (!S->bar.wait() && rc == SIGNER_OK) rc = fail(SIGNER_ERR_CUDA, "a peer shard failed");
    return rc;
}

// after step 3: this shard's ordered W E' sum into its slot; every shard ends up with all N slots
int shard_gather_corrE(signer_chain *c)
{
    ShardLink *L = c->link;
    ShardShared *S = L->sh;
    double *slot = c->corrE_all + c->nP * (size_t)L->rank;
    if (S->emulated) CU(cudaMemsetAsync(c->corrE_all, 0, sizeof(double) * c->nP * (size_t)L->n, c->stream));
    seg_reduce<<<(unsigned)((c->nP + kPElems - 1) / kPElems), 32 * kPWarps, blocked_sum_smem_bytes(c->n_seg), c->stream>>>(
        c->corrE_part, c->n_seg, (int)c->nP, slot);
    c->launches++;
    ST(check_launch(c->stream, "seg_reduce"));
    if (S->emulated) return shard_sum<double>(c, c->corrE_all, c->nP * (size_t)L->n);      // the other slots are zero: exact
    const int rc = S->nccl.AllGather(slot, c->corrE_all, c->nP, kNcclFloat64, S->comms[L->rank], c->stream);   // in place
    if (rc) return nccl_fail(S, "ncclAllGather", rc);
    return SIGNER_OK;
}

// per-genome terms of every shard into its part of the global array, all-reduce, tree on shard 0 (mode_lgm: the
// constant sum lgamma(M+1), kept by every shard)
int shard_loglik(signer_chain *c, int mode_lgm, double *out0)
{
    CU(cudaMemsetAsync(c->col, 0, sizeof(double) * (size_t)c->data->n2, c->stream));
    if (mode_lgm) {
        LLParams lp{c->I, c->J, c->Jp, 1, c->data->Mrow, c->data->W, nullptr, nullptr, c->col + c->col0, 1, nullptr, 0, nullptr, nullptr, nullptr};
        loglik_cols<<<(c->J + 31) / 32, 32 * kLLWarps, loglik_smem_bytes(1), c->stream>>>(lp);
        ST(check_launch(c->stream, "loglik_cols"));
    } else {
        ST(launch_loglik_cols(c));
    }
    ST(shard_sum<double>(c, c->col, (size_t)c->data->n2));
    if (mode_lgm) tree_sum<<<1, 1024, 0, c->stream>>>(c->col, c->data->n2, nullptr, c->data->lgm);
    else if (c->link->rank == 0) tree_sum<<<1, 1024, 0, c->stream>>>(c->col, c->data->n2, c->data->lgm, out0);
    return check_launch(c->stream, "tree_sum");
}

int sharded_run(const signer_problem *pr, const signer_state *st, const signer_opts *o, signer_outputs *out,
                const int32_t *devices, int n_dev)
{
    const int I = pr->I, J = pr->J, K = pr->K;
    if (n_dev < 1 || n_dev > 32) return fail(SIGNER_ERR_ARG, "sharded run: 1..32 shards");
    if (o->burn < 0 || o->eval < 0) return fail(SIGNER_ERR_ARG, "burn and eval must be non-negative");
    const int n_seg = (J + kSeg - 1) / kSeg;
    if (n_dev > n_seg) return fail(SIGNER_ERR_ARG, "sharded run: more shards than 32-genome segments");
    ShardShared S;
    S.n = n_dev; S.bar.n = n_dev; S.start.n = n_dev; S.bufs.assign(n_dev, nullptr);
    for (int g = 1; g < n_dev; ++g) if (devices[g] != devices[0]) S.emulated = false;
    std::unique_lock<std::mutex> comm_lock;              // communicators are kept per device list and used by one call at a time
    if (!S.emulated) {
        for (int a = 0; a < n_dev; ++a) for (int b = a + 1; b < n_dev; ++b)
            if (devices[a] == devices[b]) return fail(SIGNER_ERR_ARG, "sharded run: devices must be all distinct or all the same");
        if (!S.nccl.load()) return fail(SIGNER_ERR_NCCL, "cannot load libnccl.so.2");
        // creating communicators (and their first collective) costs tens of milliseconds: keep them for the process
        static std::mutex mu;
        static std::map<std::vector<int>, std::vector<nccl_comm_t>> cache;
        comm_lock = std::unique_lock<std::mutex>(mu);
        std::vector<int> devs(devices, devices + n_dev);
        auto it = cache.find(devs);
        if (it == cache.end()) {
            std::vector<nccl_comm_t> comms(n_dev, nullptr);
            const int rc = S.nccl.CommInitAll(comms.data(), n_dev, devs.data());
            if (rc) return nccl_fail(&S, "ncclCommInitAll", rc);
            it = cache.emplace(devs, comms).first;
        }
        S.comms = it->second;
        S.keep_comms = true;
    }
    // the shards (deleted in reverse order: emulated shards launch on shard 0's stream)
    std::vector<signer_chain *> sh;
    std::vector<ShardLink> links(n_dev);
    struct Cleanup { std::vector<signer_chain *> &v; ~Cleanup() { for (auto it = v.rbegin(); it != v.rend(); ++it) delete *it; } } cleanup{sh};
    for (int g = 0; g < n_dev; ++g) {
        const int s0 = (int)((long long)g * n_seg / n_dev), s1 = (int)((long long)(g + 1) * n_seg / n_dev);
        const int c0 = s0 * kSeg, c1 = std::min(J, s1 * kSeg), Jg = c1 - c0;
        std::shared_ptr<DeviceData> d;
        ST(sgh::get_cohort(devices[g], I, Jg, pr->M + (size_t)I * c0, pr->W + (size_t)I * c0, c0, J, d));
        signer_state sg = *st;
        sg.Z = nullptr;                                  // the cube (if any) is reduced by the shard's thread
        sg.E = st->E + (size_t)K * c0; sg.Ae = st->Ae + (size_t)K * c0; sg.Be = st->Be + (size_t)K * c0;
        signer_chain *c = nullptr;
        signer_opts og = *o;
        og.device = devices[g];
        ST(chain_create(d, K, &sg, &og, &c, /*deferred_init=*/true));
        sh.push_back(c);
        links[g].rank = g; links[g].n = n_dev; links[g].sh = &S;
        c->link = &links[g];
        c->shard = g; c->n_shards = n_dev; c->col0 = c0;
        ST(dalloc(&c->corrE_all, c->nP * (size_t)n_dev, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (S.emulated && g > 0) { c->stream = sh[0]->stream; c->borrowed_stream = true; }   // one stream: plain ordering
    }
    if (S.emulated) { CU(cudaSetDevice(devices[0])); CU(cudaMalloc(&S.ptr_tab, sizeof(void *) * n_dev)); }

    // one host thread per shard: initial statistics, then the ordinary sweep loop
    std::vector<int> rcs(n_dev, SIGNER_OK);
    std::vector<std::string> errs(n_dev);
    bool need_lgm = false;                               // a collective: all shards or none
    for (auto *c : sh) need_lgm |= !c->data->lgm_ready;
    auto body = [&](int g) -> int {
        signer_chain *c = sh[g];
        CU(cudaSetDevice(c->device));
        if (need_lgm) {                                  // sum lgamma(M+1) over the whole cohort, kept by every shard
            if (is_sharded(c)) ST(shard_loglik(c, 1, nullptr));
            else ST(ensure_lgm(*c->data));
        }
        c->zcur = 0;
        if (st->Z) {
            ST(reduce_cube(c, st->Z, J, c->col0));
            ST(shard_sum<int>(c, c->Zin[0], c->nP));
        } else {
            ST(launch_z(c, o->sweep0, kStZinit, nullptr));
        }
        if (!o->Pfixed) { Ops none{0, 0u}; ST(launch_e(c, o->sweep0, none, 1, nullptr, nullptr, nullptr)); }
        // this shard's view of the caller's outputs: its block of genomes; the context-side arrays come from shard 0
        signer_outputs v = *out;
        const size_t off = (size_t)K * c->col0;
        if (v.E) v.E += off; if (v.Ae) v.Ae += off; if (v.Be) v.Be += off; if (v.Znj) v.Znj += off;
        if (g != 0) { v.P = v.Ap = v.Bp = nullptr; v.Zin = nullptr; v.Ps = v.Aps = v.Bps = nullptr; v.loglik = v.bic = nullptr; }
        CU(cudaStreamSynchronize(c->stream));
        if (!S.start.wait()) return fail(SIGNER_ERR_CUDA, "a peer shard failed");
        static const bool trace = [] { const char *e = std::getenv("SIGNER_TRACE"); return e && e[0] == '1'; }();
        if (trace) c->profiling = true;                  // events around every kernel of this shard
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = chain_run(c, o->burn, o->eval, o->keep_par, o->forced_order, &v);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (g == 0) S.loop_ms = ms;
        if (trace && rc == SIGNER_OK) {                  // where a sharded sweep spends its time: device time per kernel class vs the loop's wall time
            double kms[4] = {0, 0, 0, 0};
            long long kn[4] = {0, 0, 0, 0};
            signer_chain_profile(c, 0, kms, (int64_t *)kn);
            const int nsw = std::max(1, o->burn + o->eval);
            std::fprintf(stderr, "[sharded] shard %d/%d (device %d, %d genomes): loop %.3f ms/sweep, host enqueue %.3f ms/sweep; kernels per sweep: "
                         "z_alloc_fused %.3f, p_family %.3f, e_family %.3f ms (sum %.3f); exchanges + waiting for peers + gaps %.3f ms\n",
                         g, n_dev, c->device, c->J, ms / nsw, c->last_enqueue_ms / nsw, kms[0] / nsw, kms[1] / nsw, kms[2] / nsw,
                         (kms[0] + kms[1] + kms[2]) / nsw, ms / nsw - (kms[0] + kms[1] + kms[2]) / nsw);
        }
        return rc;
    };
    std::vector<std::thread> th;
    for (int g = 0; g < n_dev; ++g)
        th.emplace_back([&, g]() {
            rcs[g] = body(g);
            if (rcs[g] != SIGNER_OK) { errs[g] = g_err; S.bar.abort(); S.start.abort(); }
        });
    for (auto &t : th) t.join();
    g_loop_ms = S.loop_ms;
    for (int g = 0; g < n_dev; ++g) {
        if (rcs[g] == SIGNER_OK) sh[g]->data->lgm_ready = true;   // cached shard cohorts keep their constant
        if (rcs[g] != SIGNER_OK) return fail(rcs[g], "shard " + std::to_string(g) + ": " + errs[g]);
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
    // accumulated on the device chunk by chunk while the last keep_par run stored its samples (chain_run), so the
    // M-step does not depend on the evaluation sweeps fitting into one ring set
    if (!c->stats_valid)
        return fail(SIGNER_ERR_ARG, "fetch_stats needs a preceding run with keep_par and eval > 0 that stored its samples");
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(stats, c->stats, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    stats[3] = c->stats_np; stats[7] = c->stats_ne;
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
    if (c->draw_count && counts) { CU(cudaMemcpyAsync(counts, c->draw_count, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream)); CU(cudaStreamSynchronize(c->stream)); }
    if (enable && !c->draw_count) ST(dalloc(&c->draw_count, 2, c->stream));
    if (c->draw_count) { CU(cudaMemsetAsync(c->draw_count, 0, 2 * sizeof(unsigned long long), c->stream)); CU(cudaStreamSynchronize(c->stream)); }
    if (!enable && c->draw_count) { dfree(c->draw_count, c->stream); c->draw_count = nullptr; }
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

void signer_release_caches(void)
{
    sgh::cohort_cache_clear();
    sgh::PinnedCache::clear();
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return; }
    for (int d = 0; d < n; ++d) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    cudaGetLastError();
}

int signer_gibbs_run(const signer_problem *pr, const signer_state *st, const signer_opts *o, signer_outputs *out)
{
    g_err.clear();
    if (!pr || !st || !o || !out) return fail(SIGNER_ERR_ARG, "null argument");
    static const bool trace = [] { const char *e = std::getenv("SIGNER_TRACE"); return e && (e[0] == '1' || e[0] == '2'); }();   // phase times on stderr
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
    std::shared_ptr<DeviceData> d;
    ST(upload_data(o->device, pr->I, pr->J, pr->M, pr->W, d));
    const double t_data = ms_since(t0);
    signer_chain *c = nullptr;
    ST(chain_create(d, pr->K, st, o, &c));
    const double t_create = ms_since(t0);
    const auto t_loop = std::chrono::steady_clock::now();
    const int rc = chain_run(c, o->burn, o->eval, o->keep_par, o->forced_order, out);
    g_loop_ms = ms_since(t_loop);
    const std::string keep = g_err;
    signer_chain_destroy(c);
    g_err = keep;
    if (trace)
        std::fprintf(stderr, "[signer_gibbs_run] cohort %.2f ms, chain create %.2f ms, sweeps + copy-out %.2f ms, destroy %.2f ms\n", t_data,
                     t_create - t_data, g_loop_ms, ms_since(t0) - t_create - g_loop_ms);
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
    const int tpd = cells <= 250000 ? 8 : (cells <= 2000000 ? 2 : 1);
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
    if (out->Phat || out->Ehat)
        return fail(SIGNER_ERR_UNSUPPORTED, "Phat/Ehat (device-side posterior medians) are not available on the column-sharded path");
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
    return (int)cudaMemcpyFromSymbol(out, sg::g_e_clock, sizeof(long long) * 8 * (size_t)nblocks);
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
