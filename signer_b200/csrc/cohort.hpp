// cohort.hpp -- the data of one cohort (or one column shard of it) on one device, shared by every chain that
// runs on it: the layouts the kernels read and z_alloc_fused's work list.
//
// Replaces the by-value copies of M and W the reference makes on every call (src/RcppExports.cpp:35-36).  M and W
// never change while a model is fitted -- the EM loop calls the sampler up to 100 times on the same matrices
// (R/cpp_eBayesNMF.R:97-157) and model selection runs every candidate on them (R/Optimal_sigs.R:27) -- so the device
// copy is built once and kept in a small cache keyed by the content of M and W.
#pragma once
#include <memory>

#include "host_util.hpp"

namespace sgh {

struct DeviceData {
    int device = 0, I = 0, J = 0, Jp = 0, n2 = 0;
    int col0 = 0, Jglobal = 0;  // column shard: first global genome, genomes in the whole cohort
    // z_alloc's work list: the non-zero cells sorted by count (largest first, ties genome-major)
    unsigned short *cell_row = nullptr;
    int *cell_col = nullptr, *cell_n = nullptr;
    int n_cells = 0;
    int *Mrow = nullptr;       // [I][Jp], original genome order (log-likelihood)
    double *W = nullptr;       // [I][Jp]
    double *lgm = nullptr;     // sum lgamma(M+1), one double (filled by the caller: needs the log-likelihood kernel)
    bool lgm_ready = false;
    // TMA descriptor of W ([I][Jp] f64; box = 34 genomes x 96 contexts) for e_family's tile loads, made on first use
    alignas(64) unsigned char wmap[128] = {0};
    bool wmap_ready = false;
    size_t bytes = 0;          // device memory held
    ~DeviceData();
};

// The device copy of (M, W): from the cache when the same content was uploaded before, else built now --
// M and W go up as they are (column-major doubles, through the transfer pool) and kernels write the row-major
// padded layouts, validate the entries and sort the work list.  Host work per call: one hash of M and W.
// `fresh` tells the caller that lgm still has to be computed.
int get_cohort(int device, int I, int J, const double *M, const double *W, int col0, int Jglobal,
               std::shared_ptr<DeviceData> &out);

// drop every cached cohort (device memory is released when the last chain using it is destroyed)
void cohort_cache_clear();

} // namespace sgh
