/*
 * signer_gibbs.h -- C ABI of libsigner_gibbs.so: a B200 (sm_100a) implementation of signeR's
 * empirical-Bayes Poisson-NMF Gibbs sampler.
 *
 * This is the drop-in boundary for ONE path of TojalLab/signeR (v2.3.3):
 *     R: GibbsSamplerCpp(...)                      R/RcppExports.R:8-10
 *     -> .Call(`_signeR_GibbsSamplerCpp`, 24 SEXP) src/RcppExports.cpp:31-62, src/signeR_init.c:13
 *     -> GibbsSamplerCpp(...)                      src/gibbs_2.cpp:253-424
 * An Rcpp shim with the identical 24-argument signature (r_shim/) unmarshals the
 * SEXPs and calls signer_gibbs_run().  All pointers are HOST pointers to column-major doubles,
 * exactly the memory Armadillo/R hand to the reference:
 *     M, W   I x J        P, Ap, Bp  I x K        E, Ae, Be  K x J
 *     Z cube I x J x K    element (s,g,m) at s + I*(g + J*m)
 *     sample cubes (i,n,r) / (n,j,r): slice r is one full matrix   (src/gibbs_2.cpp:289-302)
 * Inputs are never written.  Outputs are caller-allocated; nothing the library keeps after a
 * stateless call returns refers to caller memory.  What it does keep, per process, is reusable device
 * state: the device copy of the last few cohorts (M, W: layouts and work list, keyed by a 128-bit hash
 * of their content -- the EM loop and the model selection call the sampler again and again on the same
 * matrices, R/cpp_eBayesNMF.R:97-157, R/Optimal_sigs.R:27), the stream-ordered memory pool and pinned
 * staging buffers.  signer_release_caches() gives all of it back.  No torch types, no R API, never
 * exit()/abort().
 *
 * There is NO CPU fallback: every entry that computes needs a CUDA device and returns
 * SIGNER_ERR_CUDA without one.
 */
#ifndef SIGNER_GIBBS_H
#define SIGNER_GIBBS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIGNER_ABI_VERSION 3
/* Version of the draw specification: the same seed gives the same chain only within one version. */
#define SIGNER_DRAW_SPEC_VERSION 7

/* status codes */
enum {
    SIGNER_OK = 0,
    SIGNER_ERR_ARG = 1,          /* bad argument / dimension mismatch */
    SIGNER_ERR_UNSUPPORTED = 2,  /* Zfixed/Thetafixed/Afixed=TRUE (src/gibbs_2.cpp:376-405): unreachable
                                    from the package (R/cpp_eBayesNMF.R:104-107,165-168), not implemented */
    SIGNER_ERR_CUDA = 3,
    SIGNER_ERR_NOMEM = 4,
    SIGNER_ERR_NCCL = 5,
    SIGNER_ERR_NONFINITE = 6     /* NaN/Inf/negative entries in M or W */
};

/* The data: replaces arguments M, W of GibbsSamplerCpp (src/gibbs_2.cpp:255). */
typedef struct {
    int32_t I, J, K;             /* contexts, genomes, signatures (Z.n_rows, n_cols, n_slices; :270-272) */
    const double *M;             /* I x J counts; truncated to int like `int size` (:31) */
    const double *W;             /* I x J opportunities */
} signer_problem;

/* The chain state: replaces arguments Z,P,E,Ap,Bp,Ae,Be (src/gibbs_2.cpp:255-256). */
typedef struct {
    const double *Z;             /* I x J x K cube or NULL.  The sampler only ever uses Z through its sums
                                    over j and over i (:112,:129) and step 1 overwrites it, so the cube is
                                    reduced to those sums on entry.  NULL: the initial allocation is drawn on
                                    the device (replaces the R loop R/cpp_eBayesNMF.R:79-89). */
    const double *P, *E, *Ap, *Bp, *Ae, *Be;
} signer_state;

/* Scalars and flags: replaces arguments ap..keep_par (src/gibbs_2.cpp:257-262). */
typedef struct {
    double ap, bp, ae, be, lp, le, var_ap, var_ae;
    int32_t burn, eval;
    int32_t Pfixed, Zfixed, Thetafixed, Afixed, keep_par;
    uint64_t seed;               /* replaces Rcpp::RNGScope (src/RcppExports.cpp:34): the shim derives it
                                    from R's RNG so set.seed() still controls the run */
    uint32_t sweep0;             /* first sweep index of the Philox streams (continue a chain: previous
                                    sweep0 + burn + eval) */
    int32_t device;              /* CUDA device ordinal */
    const int32_t *forced_order; /* test hook: 7 step ids per sweep (0 = skip) replacing the random scan
                                    (:303-306); NULL in production */
} signer_opts;

/* Caller-allocated outputs: replaces the returned Rcpp::List (src/gibbs_2.cpp:406-423).
 * Any pointer may be NULL (not wanted).  R = opts.eval. */
typedef struct {
    double *Zs;                  /* I x J x K x 1: the last Z (:279-282,323); filled only if keep_par */
    double *Ps, *Es;             /* I x K x R, K x J x R */
    double *Aps, *Aes;           /* filled only if keep_par (:293-296,327-330) */
    double *Bps, *Bes;           /* always (Bfixed is false, :265,299-302) */
    double *loglik;              /* R: per-sample Poisson log-likelihood (replaces R/cpp_eBayesNMF.R:180-185) */
    double *bic;                 /* R: 2*loglik - K*(I+J)*log(J)        (R/cpp_eBayesNMF.R:186) */
    double *P, *E, *Ap, *Bp, *Ae, *Be;   /* final state (what R re-reads from the last slice, :108-139) */
    int32_t *Zin, *Znj;          /* final sufficient statistics sum_j Z (I x K), sum_i Z (K x J) */
    double *Phat, *Ehat;         /* I x K, K x J: posterior medians of the NORMALISED samples, computed on the device
                                    (replaces Normalize + Median_sign/Median_exp, R/Definition_SignExp_object.R:116-150,
                                    :259,:297, called at R/cpp_eBayesNMF.R:187-189; signatures in sampler order, the
                                    re-ordering by total exposure :190-198 is left to the host).  Lets a final fit return
                                    its point estimates without copying the sample cubes; needs eval * 8 * K * (I+J)
                                    bytes of device memory twice. */
} signer_outputs;

/* Stateless drop-in for one GibbsSamplerCpp call. */
int signer_gibbs_run(const signer_problem *problem, const signer_state *state,
                     const signer_opts *opts, signer_outputs *out);

/* Device-resident chain: the EM loop (R/cpp_eBayesNMF.R:97-157) and long runs keep M, W, the
 * parameters and the Z statistics in HBM between calls.  opts.burn/eval are ignored by create. */
typedef struct signer_chain signer_chain;

int signer_chain_create(const signer_problem *problem, const signer_state *state,
                        const signer_opts *opts, signer_chain **chain);
/* Launch on an existing CUDA stream (a cudaStream_t) instead of the chain's own; pass NULL to
 * restore (so the legacy default stream, whose handle is NULL, cannot be selected).  Lets a host
 * framework time the chain with its own events.  The stream must outlive the chain (or be restored first). */
int signer_chain_set_stream(signer_chain *chain, void *cuda_stream);
/* Update the hyper-hyperparameters between EM iterations (ap,bp,ae,be,lp,le,var_ap,var_ae). */
int signer_chain_set_hyper(signer_chain *chain, const double hyper[8]);
/* Run `burn` sweeps without storage, then `eval` sweeps with storage into `out` (NULL allowed when
 * eval == 0).  With out == NULL and eval > 0 the eval sweeps still compute and keep, on the device,
 * everything an eval sweep produces (samples into the device ring, log-likelihood).  Asynchronous
 * when nothing has to be copied to the host; signer_chain_sync() waits. */
int signer_chain_run(signer_chain *chain, int32_t burn, int32_t eval, int32_t keep_par,
                     const int32_t *forced_order, signer_outputs *out);
int signer_chain_sync(signer_chain *chain);
/* Copy the current state to the host (any pointer of `out` may be NULL; sample fields ignored). */
int signer_chain_fetch_state(signer_chain *chain, signer_outputs *out);
/* Sufficient statistics of the stored eval samples for the EM M-step (R/EstimateParameters.R:1-13):
 * stats[0..3] = sum(Aps), sum(Bps), sum(log Bps), count ; stats[4..7] likewise for Aes/Bes. */
int signer_chain_fetch_stats(signer_chain *chain, double stats[8]);
/* Per-kernel device time of the sweeps run since the last call (CUDA events around every launch;
 * enable with profile != 0, which serialises nothing but adds two events per launch).
 * ms[0..3] = z_alloc_fused, p_family, e_family, loglik ; launches[0..3] likewise. */
int signer_chain_profile(signer_chain *chain, int32_t enable, double ms[4], int64_t launches[4]);
/* Number of sweeps run so far (next Philox sweep index = sweep0 + this). */
int64_t signer_chain_sweeps(const signer_chain *chain);
/* Number of kernels this chain has launched so far (memsets and copies not counted). */
int64_t signer_chain_launches(const signer_chain *chain);
/* Step-1 draw counters: counts[0] = categorical draws of the direct method, counts[1] = conditional binomials, since the
 * last call; enable != 0 switches counting on (one atomic per warp per launch), 0 off.  For the draw-rate report. */
int signer_chain_draw_counts(signer_chain *chain, int32_t enable, uint64_t counts[2]);
void signer_chain_destroy(signer_chain *chain);

/* Candidates x chains over several GPUs of one box (replaces the multisession futures of
 * R/Optimal_sigs.R:3-8,27,74): every task is an independent chain on the same (M, W); tasks are
 * dealt longest-first onto the devices, one host thread per device.  Per task only the BIC vector
 * (and log-likelihoods) return, as on the reference's testing path (R/signeR.R:169-172). */
typedef struct {
    int32_t K;
    signer_state state;          /* initial state for this K (Z may be NULL) */
    signer_opts opts;            /* device field ignored */
    double *loglik, *bic;        /* opts.eval doubles each; bic may be NULL */
    int32_t status;              /* out: per-task status */
} signer_task;

int signer_gibbs_batch(int32_t I, int32_t J, const double *M, const double *W,
                       signer_task *tasks, int32_t n_tasks,
                       const int32_t *devices, int32_t n_devices);

/* One chain over several GPUs, the genome columns sharded in 32-genome segments (very wide cohorts, BASELINE.json
 * config 4).  P, Ap, Bp are replicated; per sweep the shards exchange two sets of I*K values (NCCL all-reduce over
 * NVLink): the integer sums over genomes of Z and their ordered W E' sums.  Same arguments and outputs as
 * signer_gibbs_run (opts.device ignored).  devices: n_devices distinct CUDA ordinals, or the same ordinal repeated
 * (shards emulated on one GPU, for tests).  Results are identical for both placements and equal the oracle run with
 * nshards = n_devices; they differ from the unsharded run only in the rounding of W E' (a different, still fixed,
 * summation order). */
int signer_gibbs_run_sharded(const signer_problem *problem, const signer_state *state, const signer_opts *opts,
                             signer_outputs *out, const int32_t *devices, int32_t n_devices);

/* Test hooks: the deterministic primitives of the draw specification evaluated ON THE DEVICE, so
 * tests can compare them bit for bit with the oracle.  n values each; fn: 0 log, 1 exp, 2 lgamma,
 * 3 the ziggurat normal of stream (seed 12345, sweep 0, stream 3, element i), attempt (uint32) x[i]. */
int signer_test_math(int32_t fn, const double *x, double *y, int32_t n, int32_t device);
/* binomial(n[i], p[i]) drawn as conditional binomial number `kstep` of stream (seed, sweep, stream_id, element=i) */
int signer_test_binomial(uint64_t seed, uint32_t sweep, uint32_t stream_id, int32_t kstep, const int32_t *n,
                         const double *p, int32_t *out, int32_t count, int32_t device);
/* gamma(shape[i], rate[i]) drawn from stream (seed, sweep, stream_id, element=i) */
int signer_test_gamma(uint64_t seed, uint32_t sweep, uint32_t stream_id, const double *shape,
                      const double *rate, double *out, int32_t count, int32_t device);
/* host-side: scan order of one sweep (7 step ids) and raw Philox4x32-10 block */
void signer_test_perm(uint64_t seed, uint32_t sweep, int32_t order[7]);
void signer_test_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Measured ceiling of the uniform-draw rate on `device`: a kernel that only generates Philox4x32-10
 * blocks, turns each into four 32-bit uniforms and does one inversion step per uniform, no memory traffic. */
int signer_draw_ceiling(int32_t device, double *draws_per_s);

/* Drop the cached cohorts, trim the device memory pools and free the cached pinned buffers (see the header comment).
 * Environment: SIGNER_COHORT_CACHE_MB (default 16384; 0 disables the cohort cache). */
void signer_release_caches(void);

/* Thread-local text of the last error on this thread ("" if none). */
const char *signer_last_error(void);
/* Wall time (ms) of the sweep loop of the last signer_gibbs_run / signer_gibbs_run_sharded on this thread: from the
 * first sweep's launch to the last result on the host, excluding upload and set-up. */
double signer_last_loop_ms(void);
int signer_version(void);
int signer_draw_spec_version(void);
/* Number of CUDA devices visible, or a negative status. */
int signer_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SIGNER_GIBBS_H */
