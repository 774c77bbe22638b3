// cohort.cu -- builds and caches the device copy of a cohort (cohort.hpp).
//
// The build runs on the device: M and W arrive as the caller holds them (column-major doubles, the memory
// Armadillo/R hand to the reference, src/RcppExports.cpp:35-36) and three small kernels plus one radix sort
// (CUB, set-up only -- not on the sweep path) produce
//   Mrow  int32 [I][Jp]   counts truncated like `int size` (src/gibbs_2.cpp:31), row-major, zero-padded to Jp
//   W     f64   [I][Jp]
//   cell list             the non-zero cells sorted by count, largest first; counts up to kListDirect by their
//                         number of four-draw rounds; ties keep the genome-major order (the sort is stable), so
//                         the 32 lanes of a warp read few distinct columns of E
// and validate the entries (negative / non-finite / too large -> SIGNER_ERR_NONFINITE).
#include "cohort.hpp"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdlib>
#include <list>

namespace sgh {

DeviceData::~DeviceData()
{
    cudaSetDevice(device);
    cudaFree(cell_row); cudaFree(cell_col); cudaFree(cell_n); cudaFree(Mrow); cudaFree(W); cudaFree(lgm);
}

namespace {

constexpr int kListDirect = 32;       // == sg::kNDirect (kernels.cuh; checked in signer_gibbs.cu)

__device__ __forceinline__ bool valid_m(double v) { return (v >= 0.0) && (v <= 2147483647.0); }
__device__ __forceinline__ bool valid_w(double v) { return (v >= 0.0) && (v <= 1.7976931348623157e308); }

// 32 x 32 tiles through shared memory: read along contexts (the fast index of the caller's matrices), write along
// genomes (the fast index of the device layouts).  blockIdx.x = genome tile, blockIdx.y = context tile.
__global__ void cohort_layout(const double *Mc, const double *Wc, int I, int J, int Jp, int *Mrow, double *Wrow, int *flags)
{
    __shared__ double tw[32][33];
    __shared__ int tm[32][33];
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    int bad = 0;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int i = i0 + threadIdx.x, j = j0 + jj;
        double mv = 0.0, wv = 0.0;
        if (i < I && j < J) {
            mv = Mc[(size_t)i + (size_t)I * j]; wv = Wc[(size_t)i + (size_t)I * j];
            if (!valid_m(mv)) { bad |= 1; mv = 0.0; }
            if (!valid_w(wv)) { bad |= 2; wv = 0.0; }
        }
        tm[jj][threadIdx.x] = (int)mv;                     // int size (src/gibbs_2.cpp:31)
        tw[jj][threadIdx.x] = wv;
    }
    if (bad) atomicOr(flags, bad);
    __syncthreads();
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
        const int i = i0 + ii, j = j0 + threadIdx.x;
        if (i < I && j < Jp) {
            Mrow[(size_t)i * Jp + j] = tm[threadIdx.x][ii];
            Wrow[(size_t)i * Jp + j] = tw[threadIdx.x][ii];
        }
    }
}

// sort key of every cell in genome-major order (element id i + I*j): counts above kListDirect by value, smaller ones
// by their number of direct-method rounds (four draws per Philox block), empty cells 0 (they sort last)
__global__ void cohort_keys(const double *Mc, size_t n, unsigned *keys, unsigned *ids)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double mv = Mc[e];
    const int c = valid_m(mv) ? (int)mv : 0;
    keys[e] = (unsigned)(c > kListDirect ? c : 4 * ((c + 3) / 4));
    ids[e] = (unsigned)e;
}

// number of listed cells = position of the last non-zero key of the descending sequence
__global__ void cohort_count(const unsigned *keys_sorted, size_t n, int *n_cells)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    if (keys_sorted[t] != 0u && (t + 1 == n || keys_sorted[t + 1] == 0u)) *n_cells = (int)(t + 1);
}

__global__ void cohort_expand(const unsigned *ids_sorted, const double *Mc, int I, int n_cells, unsigned short *cell_row,
                              int *cell_col, int *cell_n)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    const unsigned e = ids_sorted[t];
    cell_row[t] = (unsigned short)(e % (unsigned)I);
    cell_col[t] = (int)(e / (unsigned)I);
    cell_n[t] = (int)Mc[e];
}

int build_cohort(int device, int I, int J, const double *M, const double *W, int col0, int Jglobal, std::shared_ptr<DeviceData> &out)
{
    CU(cudaSetDevice(device));
    ST(ensure_pool(device));
    auto d = std::make_shared<DeviceData>();
    d->device = device; d->I = I; d->J = J; d->Jp = (J + 31) / 32 * 32;
    d->col0 = col0; d->Jglobal = Jglobal;
    d->n2 = 1; while (d->n2 < Jglobal) d->n2 <<= 1;
    const size_t n = (size_t)I * J, np = (size_t)I * d->Jp;
    cudaStream_t st = nullptr;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamSynchronize(s); cudaStreamDestroy(s); } } sguard{st};

    // M and W as the caller holds them, through the transfer pool (pageable source)
    DevTemp Mc, Wc, keys_in, keys_out, ids_in, ids_out, tmp, small;
    ST(Mc.get(n * sizeof(double), st)); ST(Wc.get(n * sizeof(double), st));
    ST(small.get(2 * sizeof(int), st));
    CU(cudaMemsetAsync(small.p, 0, 2 * sizeof(int), st));
    cudaEvent_t alloc_done = nullptr;
    CU(cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming));
    struct EventGuard { cudaEvent_t e; ~EventGuard() { cudaEventDestroy(e); } } eguard{alloc_done};
    CU(cudaEventRecord(alloc_done, st));
    TransferPool &pool = TransferPool::of(device);
    GroupPtr g = make_group();
    pool.post_h2d(M, Mc.p, n * sizeof(double), alloc_done, g);
    pool.post_h2d(W, Wc.p, n * sizeof(double), alloc_done, g);

    CU(cudaMalloc(&d->Mrow, np * sizeof(int)));
    CU(cudaMalloc(&d->W, np * sizeof(double)));
    CU(cudaMalloc(&d->lgm, sizeof(double)));
    ST(keys_in.get(n * 4, st)); ST(keys_out.get(n * 4, st)); ST(ids_in.get(n * 4, st)); ST(ids_out.get(n * 4, st));
    size_t tmp_bytes = 0;
    CU(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, keys_in.as<unsigned>(), keys_out.as<unsigned>(), ids_in.as<unsigned>(),
                                                 ids_out.as<unsigned>(), (int)n, 0, 32, st));
    ST(tmp.get(tmp_bytes, st));
    if (!pool.wait(g)) return fail(SIGNER_ERR_CUDA, "upload of M, W: " + g->err);

    int *flags = small.as<int>(), *n_cells_d = small.as<int>() + 1;
    cohort_layout<<<dim3((unsigned)(d->Jp / 32), (unsigned)((I + 31) / 32)), dim3(32, 8), 0, st>>>(Mc.as<double>(), Wc.as<double>(), I, J, d->Jp,
                                                                                                     d->Mrow, d->W, flags);
    cohort_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Mc.as<double>(), n, keys_in.as<unsigned>(), ids_in.as<unsigned>());
    CU(cudaGetLastError());
    CU(cub::DeviceRadixSort::SortPairsDescending(tmp.p, tmp_bytes, keys_in.as<unsigned>(), keys_out.as<unsigned>(), ids_in.as<unsigned>(),
                                                 ids_out.as<unsigned>(), (int)n, 0, 32, st));
    cohort_count<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys_out.as<unsigned>(), n, n_cells_d);
    CU(cudaGetLastError());
    int host_small[2] = {0, 0};
    CU(cudaMemcpyAsync(host_small, small.p, sizeof(host_small), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (host_small[0] & 1) return fail(SIGNER_ERR_NONFINITE, "M has a negative, non-finite or too large entry");
    if (host_small[0] & 2) return fail(SIGNER_ERR_NONFINITE, "W has a negative or non-finite entry");
    d->n_cells = host_small[1];
    const size_t nc = (size_t)std::max(1, d->n_cells);
    CU(cudaMalloc(&d->cell_row, nc * sizeof(unsigned short)));
    CU(cudaMalloc(&d->cell_col, nc * sizeof(int)));
    CU(cudaMalloc(&d->cell_n, nc * sizeof(int)));
    if (d->n_cells > 0) {
        cohort_expand<<<(unsigned)((d->n_cells + 255) / 256), 256, 0, st>>>(ids_out.as<unsigned>(), Mc.as<double>(), I, d->n_cells, d->cell_row,
                                                                           d->cell_col, d->cell_n);
        CU(cudaGetLastError());
    } else {
        CU(cudaMemsetAsync(d->cell_row, 0, sizeof(unsigned short), st));
        CU(cudaMemsetAsync(d->cell_col, 0, sizeof(int), st));
        CU(cudaMemsetAsync(d->cell_n, 0, sizeof(int), st));
    }
    CU(cudaStreamSynchronize(st));
    d->bytes = np * 12 + nc * 10 + 8;
    out = d;
    return SIGNER_OK;
}

struct CacheKey {
    int device, I, J, col0, Jglobal;
    uint64_t hm[2], hw[2];
    bool operator==(const CacheKey &o) const
    {
        return device == o.device && I == o.I && J == o.J && col0 == o.col0 && Jglobal == o.Jglobal && hm[0] == o.hm[0] && hm[1] == o.hm[1] &&
               hw[0] == o.hw[0] && hw[1] == o.hw[1];
    }
};

std::mutex g_cache_mu;
std::list<std::pair<CacheKey, std::shared_ptr<DeviceData>>> g_cache;

size_t cache_limit_bytes()
{
    static const size_t lim = [] {
        if (const char *e = std::getenv("SIGNER_COHORT_CACHE_MB")) return (size_t)std::max(0LL, std::atoll(e)) << 20;
        return size_t(16) << 30;
    }();
    return lim;
}
constexpr size_t kCacheEntries = 8;

} // namespace

int get_cohort(int device, int I, int J, const double *M, const double *W, int col0, int Jglobal, std::shared_ptr<DeviceData> &out)
{
    if (Jglobal < 0) Jglobal = J;
    if (I <= 0 || J <= 0 || !M || !W) return fail(SIGNER_ERR_ARG, "problem: I, J must be positive and M, W non-null");
    if ((long long)I * J >= (1LL << 31)) return fail(SIGNER_ERR_ARG, "problem: I*J must be below 2^31");
    if (I > 65535) return fail(SIGNER_ERR_ARG, "problem: I must be below 65536");
    if (device < 0 || device >= 64) return fail(SIGNER_ERR_ARG, "device ordinal out of range");
    const size_t limit = cache_limit_bytes();
    if (limit == 0) return build_cohort(device, I, J, M, W, col0, Jglobal, out);
    CacheKey key{device, I, J, col0, Jglobal, {0, 0}, {0, 0}};
    const size_t bytes = (size_t)I * J * sizeof(double);
    hash128_two(M, W, bytes, key.hm, key.hw);
    std::lock_guard<std::mutex> lk(g_cache_mu);        // builds are serialised: a second caller with the same content waits and hits
    for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
        if (it->first == key) {
            out = it->second;
            g_cache.splice(g_cache.begin(), g_cache, it);   // most recently used first
            return SIGNER_OK;
        }
    ST(build_cohort(device, I, J, M, W, col0, Jglobal, out));
    g_cache.emplace_front(key, out);
    size_t held = 0, cnt = 0;
    for (auto it = g_cache.begin(); it != g_cache.end();) {
        held += it->second->bytes; ++cnt;
        if (cnt > 1 && (held > limit || cnt > kCacheEntries)) { held -= it->second->bytes; it = g_cache.erase(it); }
        else ++it;
    }
    return SIGNER_OK;
}

void cohort_cache_clear()
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.clear();
}

} // namespace sgh
