// kernels.cuh -- the sm_100a kernels of the Gibbs sweep.
//
// One sweep of the reference (src/gibbs_2.cpp:305-318) is seven steps in a random order.  Steps
// {2,4,6} touch only (P,Ap,Bp), steps {5,7} only (E,Ae,Be); the only cross-family reads are
// step 1 <- (P,E), step 2 <- (E via W*E', sum_j Z), step 3 <- (P, sum_i Z).  So for any
// permutation the sweep is three launches ordered like steps 1,2,3 appear in it:
//   z_alloc_fused  (step 1 + the cube sums of src/gibbs_2.cpp:9-28; Z never written to HBM)
//   p_family       (the P-family steps of this sweep in permutation order, element-wise on I x K)
//   e_family       (the E-family steps in permutation order on K x J, with corrP = P'W in front
//                   and the corrE = W E' segment partials for the next step 2 behind)
// plus, in evaluation sweeps, the log-likelihood (R/cpp_eBayesNMF.R:180-185).
// Random numbers are counter-based (draws.cuh), so the regrouping cannot change any draw.
//
// Device layouts: P,Ap,Bp,Zin I x K and E,Ae,Be,Znj K x J column-major exactly as the reference
// (state and samples move with plain copies); M (int32) and W (fp64) row-major [I][Jp], Jp = J
// rounded up to 32 and zero-padded, so that a warp whose lanes are 32 genomes reads them with
// coalesced 128-bit loads.
#pragma once
#include "draws.cuh"

namespace sg {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kEThreads = 256;    // e_family: 128 columns x 2 k-group lanes
constexpr int kLLChunk = 32;      // rows per log-likelihood chunk

struct Hyper { double ap, bp, ae, be, lp, le, var_ap, var_ae; };

// ------------------------------------------------------------------------------------------------
// K1  z_alloc_fused : replaces gibbs_step1 (src/gibbs_2.cpp:85-103) + cube_sum_j/_i (:9-28)
//
// One thread per (context, genome) cell; a warp owns a unit of 32 genomes x 4 context rows and
// works alone (private shared memory, no block barrier; units are claimed heaviest-first from a
// global counter).  Per row the 32 lanes draw their cells' allocations with the per-count switch
// of the draw specification:
//   n <= 16 : direct method -- n categorical draws by binary search in the lane's cumulative
//             P[i,.] o E[.,j]; the Philox block of four draws is generated warp-converged;
//   n  > 16 : R's rmultinom structure -- the K classes in lockstep, one conditional binomial each.
// The class counts are reduced on the spot: over the 32 genomes with one REDUX (-> Zin, one
// integer atomic per row/class/warp); the per-genome sums Znj take one integer atomic per non-zero
// count (fire-and-forget reductions into L2).  Z itself is never written (except, on request, the final
// sweep's cube for the reference's `Zs`).
//
// Genomes are visited in a kernel-private order, sorted by mutation burden (col_id maps a
// position to the original genome): neighbouring lanes then hold similar counts, so they take the
// same sampler branch with similar trip counts.  Random streams, Znj and the cube are addressed
// by the ORIGINAL genome index, so the order cannot change any result.
// ------------------------------------------------------------------------------------------------
constexpr int kZUnitRows = 4;     // context rows per unit
constexpr int kZWarps = 2;        // independent warps per block
constexpr int kNDirect = 16;      // largest count drawn by the direct method

struct ZParams {
    int I, J, Jp, K;
    const int *M;               // [I][Jp], columns in burden order
    const int *col_id;          // [Jp] original genome of each position, -1 for padding
    const double *P, *E;
    int *Zin, *Znj;             // accumulated with integer atomics; zeroed by the previous launch
    int *Zin_next, *Znj_next;   // the other half of the ping-pong, zeroed by this launch
    double *Zcube;              // optional I x J x K (pre-zeroed) for the last Z (src/gibbs_2.cpp:323)
    unsigned long long *unit_counter;
    unsigned long long unit_base;
    int n_rc, n_units;          // row chunks per column tile, units in total
    int warp_smem;              // bytes of shared memory per warp
    Key key;
    uint32_t sweep, stream;
};

// per warp: E_w [K][33] f64, cum4 [ceil((K-1)/4)][32] f64, P_w [4][K] f64, zc [K][32] u8
inline int z_groups(int K) { return (K - 1 + 3) / 4; }
inline size_t z_smem_bytes_per_warp(int K)
{
    return (size_t)8 * ((size_t)K * 33 + (size_t)z_groups(K) * 32 + (size_t)kZUnitRows * K) + ((size_t)K * 32 + 7) / 8 * 8;
}

__global__ void __launch_bounds__(32 * kZWarps, 12) z_alloc_fused(const ZParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int I = p.I, J = p.J, K = p.K;
    unsigned char *base = smem_raw + (size_t)warp * p.warp_smem;
    double *E_w = reinterpret_cast<double *>(base);
    const int ng = (K - 1 + 3) / 4;                         // groups of four thresholds
    double *cum4 = E_w + K * 33;                            // cumulative P o E at the end of each group
    double *P_w = cum4 + ng * 32;
    unsigned char *zc = reinterpret_cast<unsigned char *>(P_w + kZUnitRows * K);

    // zero the other half of the ping-pong for the next launch
    {
        const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
        for (size_t i = gtid; i < (size_t)I * K; i += gsz) p.Zin_next[i] = 0;
        for (size_t i = gtid; i < (size_t)K * J; i += gsz) p.Znj_next[i] = 0;
    }
    for (int k = 0; k < K; ++k) zc[k * 32 + lane] = 0;      // invariant: all zero between cells
    int pow2 = 0;                                           // largest power of two <= ng
    if (ng > 0) { pow2 = 1; while (pow2 * 2 <= ng) pow2 *= 2; }

    for (;;) {
        int unit = 0;
        if (lane == 0) unit = (int)(atomicAdd(p.unit_counter, 1ULL) - p.unit_base);
        unit = __shfl_sync(kFull, unit, 0);
        if (unit >= p.n_units) break;
        const int ct = unit / p.n_rc, rc = unit - ct * p.n_rc;
        const int j0 = ct * 32, i0 = rc * kZUnitRows;
        const int jo = p.col_id[j0 + lane];
        // stage E[., 32 genomes] (gather by original genome) and P[4 rows, .]
        {
            int k = lane, jl = 0;
            while (k >= K) { k -= K; ++jl; }
            for (int idx = lane; idx < K * 32; idx += 32) {
                const int js = __shfl_sync(kFull, jo, jl);
                E_w[k * 33 + jl] = (js >= 0) ? p.E[(size_t)K * js + k] : 0.0;
                k += 32;
                while (k >= K) { k -= K; ++jl; }
            }
        }
        for (int idx = lane; idx < kZUnitRows * K; idx += 32) {
            const int r = idx & (kZUnitRows - 1), k = idx / kZUnitRows;
            P_w[r * K + k] = (i0 + r < I) ? p.P[(size_t)(i0 + r) + (size_t)I * k] : 0.0;
        }
        __syncwarp();

        for (int r = 0; r < kZUnitRows; ++r) {
            const int row = i0 + r;
            if (row >= I) break;
            const int n = (jo >= 0) ? p.M[(size_t)row * p.Jp + j0 + lane] : 0;
            if (!__any_sync(kFull, n > 0)) continue;
            const double *Pr = P_w + r * K;
            const bool maybe_small = __any_sync(kFull, n > 0 && n <= kNDirect);
            double S = 0.0, c = 0.0;
            for (int k = 0; k < K; ++k) {
                const double pe = Pr[k], ee = E_w[k * 33 + lane];
                S = fma(pe, ee, S);
                if (maybe_small && k < K - 1) {
                    c = c + pe * ee;
                    if ((k & 3) == 3 || k == K - 2) cum4[(k >> 2) * 32 + lane] = c;
                }
            }
            const bool ok = (S > 0.0 && S <= 1.7976931348623157e308);
            const bool small = n > 0 && ok && n <= kNDirect;
            int nrem = (n > kNDirect && ok) ? n : 0;
            const int nlast = (n > 0 && !ok) ? n : 0;       // degenerate probabilities: all to the last class
            UStream us;
            us.init(p.key, p.sweep, p.stream, (uint32_t)(row + I * jo));

            // ---- direct method for the small counts of this row
            if (__any_sync(kFull, small)) {
                const int tmax = __reduce_max_sync(kFull, small ? n : 0);
                unsigned touched = 0;
                U4 w4 = U4{0, 0, 0, 0};
                for (int t = 0; t < tmax; ++t) {
                    if ((t & 3) == 0) w4 = us.words((uint32_t)t >> 2);        // converged: once per four draws
                    if (small && t < n) {
                        const uint32_t wt = (t & 2) ? ((t & 1) ? w4.w : w4.z) : ((t & 1) ? w4.y : w4.x);
                        const double x = u32(wt) * S;
                        int g = 0;                                            // number of group ends <= x
                        for (int step = pow2; step; step >>= 1) {
                            const int np = g + step;
                            if (np <= ng && cum4[(np - 1) * 32 + lane] <= x) g = np;
                        }
                        int cls = K - 1;
                        if (g < ng) {                                         // x < cum4[g]: finish inside the group
                            double cc = g ? cum4[(g - 1) * 32 + lane] : 0.0;
                            cls = 4 * g;
                            for (;;) {
                                cc = cc + Pr[cls] * E_w[cls * 33 + lane];
                                if (x < cc || cls >= K - 2) break;
                                ++cls;
                            }
                        }
                        zc[cls * 32 + lane] += 1;
                        touched |= 1u << (cls & 31);
                    }
                }
                unsigned todo = (K <= 32) ? __reduce_or_sync(kFull, touched) : 0xffffffffu;
                for (int k = 0; k < K; ++k) {
                    if (K <= 32 && !((todo >> k) & 1u)) continue;
                    const int z = zc[k * 32 + lane];
                    const int zs = __reduce_add_sync(kFull, z);
                    if (zs) {
                        if (lane == 0) atomicAdd(p.Zin + (size_t)row + (size_t)I * k, zs);
                        if (z) {
                            zc[k * 32 + lane] = 0;
                            atomicAdd(p.Znj + (size_t)k + (size_t)K * jo, z);
                            if (p.Zcube) p.Zcube[(size_t)row + (size_t)I * ((size_t)jo + (size_t)J * k)] = (double)z;
                        }
                    }
                }
            }

            // ---- conditional-binomial chain for the large counts (and the degenerate cells)
            if (__any_sync(kFull, (nrem | nlast) != 0)) {
                double rem = S;
                U4 w4 = U4{0, 0, 0, 0};
                for (int k = 0; k <= K - 1; ++k) {
                    int z = 0;
                    if (k < K - 1) {
                        if (!__any_sync(kFull, nrem > 0)) { k = K - 2; continue; }   // jump to the remainder class
                        if ((k & 3) == 0) w4 = us.words((uint32_t)k >> 2);          // converged: once per four classes
                        if (nrem > 0) {
                            const double prob = Pr[k] * E_w[k * 33 + lane];
                            if (prob != 0.0) {
                                const double pp = prob / rem;
                                const uint32_t wk = (k & 2) ? ((k & 1) ? w4.w : w4.z) : ((k & 1) ? w4.y : w4.x);
                                z = (pp > 0.0 && pp < 1.0) ? binomial_draw(us, k, u32(wk), nrem, pp) : nrem;
                                nrem -= z;
                            }
                            rem = rem - prob;
                        }
                    } else {
                        z = nrem + nlast;
                    }
                    const int zs = __reduce_add_sync(kFull, z);
                    if (zs) {
                        if (lane == 0) atomicAdd(p.Zin + (size_t)row + (size_t)I * k, zs);
                        if (z) {
                            atomicAdd(p.Znj + (size_t)k + (size_t)K * jo, z);
                            if (p.Zcube) p.Zcube[(size_t)row + (size_t)I * ((size_t)jo + (size_t)J * k)] = (double)z;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// element-wise updates shared by the two families
// ------------------------------------------------------------------------------------------------
// steps 4/5 (src/gibbs_2.cpp:140-169)
SG_D double update_b(Key key, uint32_t sweep, uint32_t stream, uint32_t e, double A, double X, double a, double b)
{
    UStream us; us.init(key, sweep, stream, e);
    const double shape = A + a + 1;
    const double rate = X + b;
    const double g = gamma_draw(us, shape, rate);
    return (kTiny < g) ? g : kTiny;
}

// steps 6/7 (src/gibbs_2.cpp:172-251): gamma proposal, log acceptance ratio, uniform accept
SG_D double update_a(Key key, uint32_t sweep, uint32_t stream, uint32_t e, double A, double X, double B, double l, double var)
{
    UStream us; us.init(key, sweep, stream, e);
    const double shape = (A * A) / var;
    const double rate = A / var;
    const double g = gamma_draw(us, shape, rate);
    const double y = (kTiny < g) ? g : kTiny;
    const double lnp = (y - A) * (det_log(B) + det_log(X) - l)
                     + det_lgamma(A + 1) - det_lgamma(y + 1)
                     + det_lgamma((A * A) / var) - det_lgamma((y * y) / var)
                     + (((y * y) - (A * A)) / var) * det_log((A * y) / var)
                     + det_log(y) - det_log(A);
    double pch = det_exp(lnp);
    if (X == 0) pch = (y < A) ? 1.0 : 0.0;
    double ub, um;
    us.pair(kAuxBlock, ub, um);
    return (um <= pch) ? y : A;
}

// up to three step ids of one family, in permutation order, 4 bits each
struct Ops {
    int n; unsigned packed;
    __host__ __device__ int get(int o) const { return (int)((packed >> (4 * o)) & 15u); }
    __host__ void push(int s) { packed |= (unsigned)s << (4 * n); ++n; }
};

// ------------------------------------------------------------------------------------------------
// KP  p_family : gibbs_step2/4/6 (src/gibbs_2.cpp:106-120,140-153,172-210) on I x K
// ------------------------------------------------------------------------------------------------
struct PParams {
    int IK;
    double *P, *Ap, *Bp;
    const int *Zin;
    const double *corrE_part;   // [n_seg][I*K] segment partials of W E'
    int n_seg, n_shards;
    double *Ps_slot, *Aps_slot, *Bps_slot;   // sample slots or null (src/gibbs_2.cpp:325-333)
    Hyper h;
    Key key;
    uint32_t sweep;
    Ops ops;
};

__global__ void __launch_bounds__(128) p_family(const PParams p)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.IK) return;
    double P = p.P[e], Ap = p.Ap[e], Bp = p.Bp[e];
    for (int o = 0; o < p.ops.n; ++o) {
        const int op = p.ops.get(o);
        if (op == 2) {
            double tot = 0.0;
            for (int sh = 0; sh < p.n_shards; ++sh) {
                const int s0 = (int)((long long)sh * p.n_seg / p.n_shards), s1 = (int)((long long)(sh + 1) * p.n_seg / p.n_shards);
                double part = 0.0;
#pragma unroll 8
                for (int s = s0; s < s1; ++s) part = part + p.corrE_part[(size_t)s * p.IK + e];
                tot = tot + part;
            }
            UStream us; us.init(p.key, p.sweep, kStP, (uint32_t)e);
            const double shape = Ap + 1 + (double)p.Zin[e];
            const double rate = Bp + tot;
            P = gamma_draw(us, shape, rate);
        } else if (op == 4) {
            Bp = update_b(p.key, p.sweep, kStBp, (uint32_t)e, Ap, P, p.h.ap, p.h.bp);
        } else if (op == 6) {
            Ap = update_a(p.key, p.sweep, kStAp, (uint32_t)e, Ap, P, Bp, p.h.lp, p.h.var_ap);
        }
    }
    p.P[e] = P; p.Ap[e] = Ap; p.Bp[e] = Bp;
    if (p.Ps_slot) p.Ps_slot[e] = P;
    if (p.Aps_slot) p.Aps_slot[e] = Ap;
    if (p.Bps_slot) p.Bps_slot[e] = Bp;
}

// ------------------------------------------------------------------------------------------------
// KE  e_family : gibbs_step3/5/7 (src/gibbs_2.cpp:123-137,156-169,213-251) on K x J,
//     corrP = P' W (:128) in front, corrE = W E' (:111) segment partials behind.
// grid = (segments of 128 columns, k-splits); thread = (column, k-group lane)
// ------------------------------------------------------------------------------------------------
struct EParams {
    int I, J, Jp, K;
    const double *W;            // [I][Jp]
    const double *P;
    double *E, *Ae, *Be;
    const int *Znj;
    double *corrE_part;         // [n_seg][I*K]
    int kg_per_split;           // k-groups (of 4) per blockIdx.y
    int do_corrE;
    double *Es_slot, *Aes_slot, *Bes_slot;
    Hyper h;
    Key key;
    uint32_t sweep;
    Ops ops;
};

inline size_t e_smem_bytes(int kg_per_split) { return sizeof(double) * (size_t)kSeg * (size_t)(4 * kg_per_split); }

__global__ void __launch_bounds__(kEThreads) e_family(const EParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *E_s = reinterpret_cast<double *>(smem_raw);   // [128][Kloc]
    const int tid = threadIdx.x, jl = tid & (kSeg - 1), q = tid >> 7;
    const int I = p.I, J = p.J, K = p.K;
    const int seg = blockIdx.x, j0 = seg * kSeg, j = j0 + jl;
    const int kBegin = blockIdx.y * p.kg_per_split * 4;
    const int kEnd = min(K, kBegin + p.kg_per_split * 4);
    const int Kloc = p.kg_per_split * 4;
    bool has3 = false;
    for (int o = 0; o < p.ops.n; ++o) has3 |= (p.ops.get(o) == 3);

    if (j < J) {
        for (int k0 = kBegin + 4 * q; k0 < kEnd; k0 += 8) {
            const int nk = min(4, kEnd - k0);
            double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
            if (has3) {
                const double *Wc = p.W + j;
                const double *P0 = p.P + (size_t)I * k0;
                const double *P1 = p.P + (size_t)I * min(k0 + 1, K - 1);
                const double *P2 = p.P + (size_t)I * min(k0 + 2, K - 1);
                const double *P3 = p.P + (size_t)I * min(k0 + 3, K - 1);
#pragma unroll 4
                for (int i = 0; i < I; ++i) {
                    const double w = Wc[(size_t)i * p.Jp];
                    acc0 = fma(__ldg(P0 + i), w, acc0);
                    acc1 = fma(__ldg(P1 + i), w, acc1);
                    acc2 = fma(__ldg(P2 + i), w, acc2);
                    acc3 = fma(__ldg(P3 + i), w, acc3);
                }
            }
            for (int c = 0; c < nk; ++c) {
                const int k = k0 + c;
                const size_t e = (size_t)k + (size_t)K * j;
                double E = p.E[e], Ae = p.Ae[e], Be = p.Be[e];
                const double corrP = (c == 0) ? acc0 : (c == 1) ? acc1 : (c == 2) ? acc2 : acc3;
                for (int o = 0; o < p.ops.n; ++o) {
                    const int op = p.ops.get(o);
                    if (op == 3) {
                        UStream us; us.init(p.key, p.sweep, kStE, (uint32_t)e);
                        const double shape = Ae + 1 + (double)p.Znj[e];
                        const double rate = Be + corrP;
                        E = gamma_draw(us, shape, rate);
                    } else if (op == 5) {
                        Be = update_b(p.key, p.sweep, kStBe, (uint32_t)e, Ae, E, p.h.ae, p.h.be);
                    } else if (op == 7) {
                        Ae = update_a(p.key, p.sweep, kStAe, (uint32_t)e, Ae, E, Be, p.h.le, p.h.var_ae);
                    }
                }
                p.E[e] = E; p.Ae[e] = Ae; p.Be[e] = Be;
                if (p.Es_slot) p.Es_slot[e] = E;
                if (p.Aes_slot) p.Aes_slot[e] = Ae;
                if (p.Bes_slot) p.Bes_slot[e] = Be;
                E_s[jl * Kloc + (k - kBegin)] = E;
            }
        }
    }
    if (!p.do_corrE) return;
    __syncthreads();
    const int jn = min(kSeg, J - j0);
    const int Kn = kEnd - kBegin;
    for (int o = tid; o < I * Kn; o += kEThreads) {
        const int i = o / Kn, kk = o - i * Kn;
        const double *Wr = p.W + (size_t)i * p.Jp + j0;
        double acc = 0.0;
#pragma unroll 4
        for (int c = 0; c < jn; ++c) acc = fma(Wr[c], E_s[c * Kloc + kk], acc);
        p.corrE_part[(size_t)seg * ((size_t)I * K) + (size_t)i + (size_t)I * (kBegin + kk)] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// K8  log-likelihood (R/cpp_eBayesNMF.R:180-185): per-column sums in 32-row chunks, then a
//     stride-halving tree over the zero-padded column array.  mode 0: sum M log(lam) - lam,
//     mode 1: sum lgamma(M+1) (constant of the data, evaluated once per chain).
// ------------------------------------------------------------------------------------------------
struct LLParams {
    int I, J, Jp, K;
    const int *M;
    const double *W, *P, *E;
    double *col;                // [n2]
    int mode;
};

__global__ void __launch_bounds__(128) loglik_cols(const LLParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int I = p.I, J = p.J, K = p.K;
    const int nchunk = (I + kLLChunk - 1) / kLLChunk;
    double *E_s = reinterpret_cast<double *>(smem_raw);   // [K][33]
    double *part = E_s + K * 33;                          // [nchunk][32]
    const int tid = threadIdx.x, lane = tid & 31, cl = tid >> 5;
    const int j0 = blockIdx.x * 32, j = j0 + lane;
    if (p.mode == 0) {
        for (int idx = tid; idx < K * 32; idx += 128) {
            const int jj = idx / K, k = idx - jj * K;
            E_s[k * 33 + jj] = (j0 + jj < J) ? p.E[(size_t)K * j0 + idx] : 0.0;
        }
    }
    __syncthreads();
    for (int c = cl; c < nchunk; c += 4) {
        double a = 0.0;
        if (j < J) {
            const int i1 = min(I, (c + 1) * kLLChunk);
            for (int i = c * kLLChunk; i < i1; ++i) {
                const double m = (double)p.M[(size_t)i * p.Jp + j];
                if (p.mode == 0) {
                    double pe = 0.0;
                    for (int k = 0; k < K; ++k) pe = fma(__ldg(p.P + (size_t)i + (size_t)I * k), E_s[k * 33 + lane], pe);
                    const double lam = pe * p.W[(size_t)i * p.Jp + j];
                    a = a + (m * det_log(lam) - lam);
                } else {
                    a = a + det_lgamma(m + 1.0);
                }
            }
        }
        part[c * 32 + lane] = a;
    }
    __syncthreads();
    if (tid < 32 && j < J) {
        double a = 0.0;
        for (int c = 0; c < nchunk; ++c) a = a + part[c * 32 + tid];
        p.col[j] = a;
    }
}

// a[t] += a[t+s] for s = n2/2 .. 1 ; result written to out[0] = a[0] - sub (sub may be null)
__global__ void __launch_bounds__(1024) tree_sum(double *a, int n2, const double *sub, double *out)
{
    for (int s = n2 >> 1; s >= 1; s >>= 1) {
        for (int t = threadIdx.x; t < s; t += 1024) a[t] = a[t] + a[t + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sub ? a[0] - sub[0] : a[0];
}

// ------------------------------------------------------------------------------------------------
// entry: reduce one slab Z[:,:,k] of the reference's cube to its sums (src/gibbs_2.cpp:112,129)
// ------------------------------------------------------------------------------------------------
__global__ void zslab_reduce(const double *slab, int I, int J, int K, int k, int *Zin, int *Znj)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    int colsum = 0;
    for (int i = 0; i < I; ++i) {
        const int z = (int)slab[(size_t)i + (size_t)I * j];
        colsum += z;
        if (z) atomicAdd(Zin + (size_t)i + (size_t)I * k, z);
    }
    Znj[(size_t)k + (size_t)K * j] = colsum;
}

// sufficient statistics of stored samples for the EM M-step (R/EstimateParameters.R:1-13):
// out[0] += sum a, out[1] += sum b, out[2] += sum log b  (fp64 atomics; a tolerance quantity)
__global__ void stats_reduce(const double *a, const double *b, size_t n, double *out)
{
    double sa = 0, sb = 0, sl = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        sa += a[i]; sb += b[i]; sl += det_log(b[i]);
    }
    for (int o = 16; o; o >>= 1) {
        sa += __shfl_xor_sync(kFull, sa, o); sb += __shfl_xor_sync(kFull, sb, o); sl += __shfl_xor_sync(kFull, sl, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, sa); atomicAdd(out + 1, sb); atomicAdd(out + 2, sl); }
}

// ------------------------------------------------------------------------------------------------
// test hooks
// ------------------------------------------------------------------------------------------------
__global__ void test_math(int fn, const double *x, double *y, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    y[i] = fn == 0 ? det_log(v) : fn == 1 ? det_exp(v) : fn == 2 ? det_lgamma(v) : det_ppnd16(v);
}
__global__ void test_binomial(Key key, uint32_t sweep, uint32_t stream, int kstep, const int *n, const double *pp, int *out, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    UStream us; us.init(key, sweep, stream, (uint32_t)i);
    const U4 w = us.words((uint32_t)kstep >> 2);
    const uint32_t wk = (kstep & 2) ? ((kstep & 1) ? w.w : w.z) : ((kstep & 1) ? w.y : w.x);
    out[i] = binomial_draw(us, kstep, u32(wk), n[i], pp[i]);
}
__global__ void test_gamma(Key key, uint32_t sweep, uint32_t stream, const double *shape, const double *rate, double *out, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    UStream us; us.init(key, sweep, stream, (uint32_t)i);
    out[i] = gamma_draw(us, shape[i], rate[i]);
}

} // namespace sg
