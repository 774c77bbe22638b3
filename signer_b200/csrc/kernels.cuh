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
// (state and samples move with plain copies).  W (fp64) and M (int32, log-likelihood only) are
// row-major [I][Jp], Jp = J rounded up to 32 and zero-padded: warps whose lanes are genomes read
// coalesced rows, the W E' pass uses 128-bit loads.  z_alloc_fused does not read M: it walks a
// count-sorted list of the non-zero cells built once per cohort (see K1).
#pragma once
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched from the driver at run time)

#include "draws.cuh"

namespace sg {

constexpr unsigned kFull = 0xffffffffu;

constexpr int kLLChunk = 32;      // rows per log-likelihood chunk

struct Hyper { double ap, bp, ae, be, lp, le, var_ap, var_ae; };

// ------------------------------------------------------------------------------------------------
// K1  z_alloc_fused : replaces gibbs_step1 (src/gibbs_2.cpp:85-103) + cube_sum_j/_i (:9-28)
//
// The counts M never change during a chain, so the non-zero cells are listed ONCE, sorted by count
// (largest first; counts up to 32 by their number of four-draw rounds, ties in genome-major order), and
// every launch walks that list: a warp takes 32 consecutive cells -- cells of (nearly) the same count,
// so the lanes run the same sampler with the same trip count -- one thread per cell.  Each lane owns one
// column of the warp's shared memory: K doubles (padded to whole groups of four) and K order bytes.
//   * n <= 32 : direct method.  The lane stages the cumulative sums c_k of P[i,k]*E[k,j]; a draw is
//     x = u*S and a branch-free two-level search (count the group ends <= x -- kept in registers for the
//     cell -- then the thresholds of that group), four draws per Philox block searched together.
//   * n  > 32 : the lane stages the products themselves, sorts the classes by binary order of
//     magnitude (counting sort on the exponent distance to S: heavy classes first), runs R's
//     conditional-binomial chain in that order while more than 32 mutations are left (inversion
//     below n*p < 30, BTRS above), and the remainder goes to the unvisited classes by the direct
//     method on the cumulative sums in index order with the visited classes zeroed -- that part is
//     shared by the whole warp (the cells' Philox blocks are dealt round-robin to the lanes).
// Z is never written (except, on request, the final sweep's cube for the reference's `Zs`): every
// direct draw and every chain step adds to the sufficient statistics at once -- sum_j Z to a
// block-wide shared-memory Zin (integer atomics, flushed once per block), sum_i Z straight to L2
// integer atomics.  Batches are claimed heaviest-first from a global counter; a batch of small counts
// claims its successor and fetches its list entries while it draws.
// ------------------------------------------------------------------------------------------------
constexpr int kZMaxWarps = 24, kZFewWarps = 16;   // most warps per block (the two register budgets of the kernel); the host picks the block size per K (z_pick_block) -- the warps of a
                                 // block share only its Zin accumulator
constexpr int kZChunk = 1;        // batches of 32 cells per claim
constexpr int kNDirect = 32;      // largest count drawn by the direct method
constexpr int kZinSmemMax = 40 * 1024;
constexpr int kNTail = 32;          // a chain stops once this few mutations are left; they are dealt by the direct method
constexpr uint32_t kTailBlock0 = 0x800u;   // first Philox block of the direct draws that finish a chain

struct ZParams {
    int I, J, K;
    int col0, Jglobal;                // column shard: first global genome of this shard, genomes in the whole cohort
    const unsigned short *cell_row;   // [n_cells] context of each listed cell
    const int *cell_col;              // [n_cells] genome
    const int *cell_n;                // [n_cells] count, non-increasing
    int n_cells, n_batches;
    const double *P, *E;
    int *Zin, *Znj;             // accumulated with integer atomics; zeroed by the previous launch
    int *Zin_next, *Znj_next;   // the other half of the ping-pong, zeroed by this launch
    double *Zcube;              // optional I x J x K (pre-zeroed) for the last Z (src/gibbs_2.cpp:323)
    unsigned long long *counter;
    unsigned long long base;
    int warp_smem;              // bytes of shared memory per warp
    int zin_smem;               // 1: accumulate Zin in shared memory per block
    unsigned long long *draw_count;   // optional [2]: categorical draws (direct method), conditional binomials (chain)
    Key key;
    uint32_t sweep, stream;
};

// per warp: buf [R][32] f64 (cumulative sums or products; R = K rounded up so that the thresholds c_0..c_{K-2} are
// followed by at least one +inf pad and fill whole groups of four), ord [K][32] u8 ; per block: Zin [I*K] i32
__host__ __device__ inline int z_groups(int K) { return (K - 1) / 4 + 1; }
__host__ __device__ inline int z_rows(int K) { return 4 * z_groups(K); }
inline size_t z_smem_bytes_per_warp(int K)
{
    return (size_t)8 * z_rows(K) * 32 + ((size_t)K * 32 + 7) / 8 * 8;
}

SG_D void z_emit(const ZParams &p, int *zin_s, int row, int col, int k, int z)
{
    if (zin_s) atomicAdd(zin_s + row + p.I * k, z);
    else atomicAdd(p.Zin + (size_t)row + (size_t)p.I * k, z);
    atomicAdd(p.Znj + (size_t)k + (size_t)p.K * col, z);
    if (p.Zcube) atomicAdd(p.Zcube + (size_t)row + (size_t)p.I * ((size_t)col + (size_t)p.J * k), (double)z);   // the shard's own cube (small integers: exact in any order)
}

// Classes of four direct draws on one column c of cumulative sums: class = first k < K-1 with x < c_k, else K-1.
// The column holds c_0..c_{K-2} followed by +inf up to a multiple of four rows (so "else K-1" needs no special case).
// Branch-free two-level search: count the group ends (every fourth row) <= x -- loaded once for the four draws, or
// passed in registers (ge) when the caller keeps them across rounds -- then the first three rows of that group.
// NG > 0: the number of groups at compile time; NG = 0: run-time ng.
template <int NG>
SG_D void z_search4(const double *c, int ng, const double *ge, double x0, double x1, double x2, double x3,
                    int &c0, int &c1, int &c2, int &c3)
{
    int g0 = 0, g1 = 0, g2 = 0, g3 = 0;
    if (NG > 0) {
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) {
            const double e = ge ? ge[gg] : c[(4 * gg + 3) * 32];
            g0 += (e <= x0) ? 1 : 0; g1 += (e <= x1) ? 1 : 0; g2 += (e <= x2) ? 1 : 0; g3 += (e <= x3) ? 1 : 0;
        }
    } else {
        for (int gg = 0; gg < ng; ++gg) {
            const double e = c[(4 * gg + 3) * 32];
            g0 += (e <= x0) ? 1 : 0; g1 += (e <= x1) ? 1 : 0; g2 += (e <= x2) ? 1 : 0; g3 += (e <= x3) ? 1 : 0;
        }
    }
    const double *a0 = c + g0 * 128, *a1 = c + g1 * 128, *a2 = c + g2 * 128, *a3 = c + g3 * 128;
    const double v00 = a0[0], v01 = a0[32], v02 = a0[64], v10 = a1[0], v11 = a1[32], v12 = a1[64];
    const double v20 = a2[0], v21 = a2[32], v22 = a2[64], v30 = a3[0], v31 = a3[32], v32 = a3[64];
    c0 = 4 * g0 + ((v00 <= x0) ? 1 : 0) + ((v01 <= x0) ? 1 : 0) + ((v02 <= x0) ? 1 : 0);
    c1 = 4 * g1 + ((v10 <= x1) ? 1 : 0) + ((v11 <= x1) ? 1 : 0) + ((v12 <= x1) ? 1 : 0);
    c2 = 4 * g2 + ((v20 <= x2) ? 1 : 0) + ((v21 <= x2) ? 1 : 0) + ((v22 <= x2) ? 1 : 0);
    c3 = 4 * g3 + ((v30 <= x3) ? 1 : 0) + ((v31 <= x3) ? 1 : 0) + ((v32 <= x3) ? 1 : 0);
}

// the direct method for one batch: lane draws nn (0 for lanes that take no part) classes, four per Philox block;
// draw t uses word t&3 of block t>>2 of the lane's stream.  The group ends stay in registers for the whole cell.
// Every draw is added to the statistics at once (measured faster than counting per class and flushing: the flush
// walks all K classes with few lanes busy).
template <int NG>
SG_D void z_direct(const double *buf, const UStream &us, int ng, int nn, int tmax, double Sd,
                   const ZParams &p, int *zin_s, int row, int col)
{
    double ge[NG > 0 ? NG : 1];
    if (NG > 0) {
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) ge[gg] = buf[(4 * gg + 3) * 32];
    }
    for (int t0 = 0; t0 < tmax; t0 += 4) {
        const U4 w4 = us.words((uint32_t)t0 >> 2);                              // converged: once per four draws
        int c0, c1, c2, c3;
        z_search4<NG>(buf, ng, NG > 0 ? ge : nullptr, u32(w4.x) * Sd, u32(w4.y) * Sd, u32(w4.z) * Sd, u32(w4.w) * Sd, c0, c1, c2, c3);
        if (t0 < nn) z_emit(p, zin_s, row, col, c0, 1);
        if (t0 + 1 < nn) z_emit(p, zin_s, row, col, c1, 1);
        if (t0 + 2 < nn) z_emit(p, zin_s, row, col, c2, 1);
        if (t0 + 3 < nn) z_emit(p, zin_s, row, col, c3, 1);
    }
}

// the direct draws that finish the chains of one batch, shared by the warp: the cells' Philox blocks (four draws
// each) are dealt round-robin to the lanes, whoever owns the cell -- a lane reads the owner's column of shared
// memory and adds the draws to the statistics under the owner's row and column.  incl/excl: running totals of
// the lanes' block counts.
template <int NG>
SG_D void z_tail(const double *buf, const ZParams &p, int *zin_s, int lane, int ng, int incl, int excl, int utot, int nrem,
                 double Sr, int row, int col)
{
    const uint32_t elem = (uint32_t)(row + p.I * (col + p.col0));
    for (int ub = 0; ub < utot; ub += 32) {
        const int g = min(ub + lane, utot - 1);
        const bool act = ub + lane < utot;
        int o = 0;                                                             // owner: the first lane with incl > g
#pragma unroll
        for (int sft = 16; sft; sft >>= 1) { const int v = __shfl_sync(kFull, incl, o + sft - 1); if (v <= g) o += sft; }
        const int b = g - __shfl_sync(kFull, excl, o);
        const int nr_o = __shfl_sync(kFull, nrem, o);
        const double Sr_o = __shfl_sync(kFull, Sr, o);
        const uint32_t elem_o = __shfl_sync(kFull, elem, o);
        const U4 w = philox4x32_10(kTailBlock0 + (uint32_t)b, elem_o, p.stream, p.sweep, p.key.k0, p.key.k1);
        const int nd = act ? min(4, nr_o - 4 * b) : 0;                         // draws of this block
        const double *buf_o = buf + (o - lane);
        const int row_o = __shfl_sync(kFull, row, o), col_o = __shfl_sync(kFull, col, o);
        int c0, c1, c2, c3;
        z_search4<NG>(buf_o, ng, nullptr, u32(w.x) * Sr_o, u32(w.y) * Sr_o, u32(w.z) * Sr_o, u32(w.w) * Sr_o, c0, c1, c2, c3);
        if (nd > 0) z_emit(p, zin_s, row_o, col_o, c0, 1);
        if (nd > 1) z_emit(p, zin_s, row_o, col_o, c1, 1);
        if (nd > 2) z_emit(p, zin_s, row_o, col_o, c2, 1);
        if (nd > 3) z_emit(p, zin_s, row_o, col_o, c3, 1);
    }
}

// run F<NG>(args) with NG = ng when ng <= 8 (K <= 32), else the run-time version
#define SG_DISPATCH_NG(ng, F, ...)                                                                     \
    switch (ng) {                                                                                       \
    case 1: F<1>(__VA_ARGS__); break; case 2: F<2>(__VA_ARGS__); break; case 3: F<3>(__VA_ARGS__); break; \
    case 4: F<4>(__VA_ARGS__); break; case 5: F<5>(__VA_ARGS__); break; case 6: F<6>(__VA_ARGS__); break; \
    case 7: F<7>(__VA_ARGS__); break; case 8: F<8>(__VA_ARGS__); break; default: F<0>(__VA_ARGS__); break; \
    }

#ifdef SG_Z_TIMING      // experiments only: cycles each batch of the last launch took (its warp's wall clock)
__device__ unsigned g_z_batch_cycles[1 << 16];
__device__ unsigned long long g_z_phase[2][8];      // [small/large][phase] warp-cycles summed over batches (all launches)
#define SG_ZCLK(ph) { const long long t_now = clock64(); if (lane == 0) atomicAdd(&g_z_phase[t_large][ph], (unsigned long long)(t_now - t_last)); t_last = t_now; }
#else
#define SG_ZCLK(ph)
#endif

SG_D void sg_prefetch(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// The constant tables of the elementary functions (log: 2 KB, ziggurat: 2 KB) live in global memory; L1 does not keep
// them across kernels, and in a small latency-bound kernel every det_log would start with an L2 round trip in the middle
// of a dependent chain.  The first 33 threads of a block pull the lines into L1 while the block does something else.
SG_D void prefetch_tables(int t)
{
    if (t < 0) return;
    if (t < 16) sg_prefetch(reinterpret_cast<const char *>(g_log_tab) + 128 * t);
    else if (t < 25) sg_prefetch(reinterpret_cast<const char *>(g_zig_x) + 128 * (t - 16));
    else if (t < 33) sg_prefetch(reinterpret_cast<const char *>(g_zig_ratio) + 128 * (t - 25));
}

SG_D int z_expo(double v) { return (__double2hiint(v) >> 20) & 0x7ff; }

// MAXW: most warps per SM the register budget is set for -- 24 (80 registers) or 16 (128 registers: the kernel as the
// compiler wants it, 15 % faster per warp).  24 warps at 80 registers win by ~5 % where 24 fit; where shared memory holds
// fewer (96 contexts from K = 28: 10 KB of cumulative sums per warp) the 128-register build does; chain_create picks.
template <int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1) z_alloc_fused(const ZParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int I = p.I, J = p.J, K = p.K;
    unsigned char *base = smem_raw + (size_t)warp * p.warp_smem;
    double *buf = reinterpret_cast<double *>(base) + lane;              // this lane's column: buf[k * 32]
    const int ng = z_groups(K), R = 4 * ng;                             // groups of four thresholds, rows of buf
    unsigned char *ord = base + (size_t)8 * R * 32 + lane;              // ord[t * 32]: class visited at step t
    int *zin_s = p.zin_smem ? reinterpret_cast<int *>(smem_raw + (size_t)(blockDim.x >> 5) * p.warp_smem) : nullptr;

    if (zin_s) {
        for (int i = tid; i < I * K; i += blockDim.x) zin_s[i] = 0;
        __syncthreads();
    }
    const bool vec2 = (K & 1) == 0;
    unsigned n_direct = 0, n_binom = 0;                     // this lane's draw counts (reported only on request)

    // batches are claimed from a global counter.  While a batch of small counts runs, the claim for the next one is
    // already in flight; a batch of large counts claims its successor when it is done (nothing is held back from
    // idle warps while a long chain runs)
    const double kInf = __longlong_as_double(0x7ff0000000000000LL);
    int cur = 0, pend = 0;
    if (lane == 0) cur = (int)(atomicAdd(p.counter, 1ULL) - p.base);
    cur = __shfl_sync(kFull, cur, 0);
    int n = 0, row = 0, col = 0;                            // this lane's cell of the current batch
    if (cur < p.n_batches && cur * 32 + lane < p.n_cells) {
        const int cell = cur * 32 + lane;
        n = p.cell_n[cell]; row = (int)p.cell_row[cell]; col = p.cell_col[cell];
    }
    // zero the other half of the ping-pong for the next launch
    {
        const size_t gtid = (size_t)blockIdx.x * blockDim.x + tid, gsz = (size_t)gridDim.x * blockDim.x;
        for (size_t i = gtid; i < (size_t)I * K; i += gsz) p.Zin_next[i] = 0;
        for (size_t i = gtid; i < (size_t)K * J; i += gsz) p.Znj_next[i] = 0;
    }
    while (cur < p.n_batches) {
#ifdef SG_Z_TIMING
        const long long t_begin = clock64();
        const int t_batch = cur;
#endif
        const bool early = __shfl_sync(kFull, n, 0) <= kNDirect;               // lane 0 holds the batch's largest count
        if (early && lane == 0) pend = (int)(atomicAdd(p.counter, 1ULL) - p.base);
#ifdef SG_Z_TIMING
        long long t_last = t_begin;
        const int t_large = early ? 0 : 1;
#endif
        SG_ZCLK(0)                                          // cell list loads, claim
        const bool valid = n > 0;                           // listed cells have n >= 1; lanes past the end of the list have 0
        const bool bycount_small = n <= kNDirect;

        // stage P[i,.] o E[.,j]: S = fma chain; cumulative sums (direct method) or the products (chain)
        const double *Pr = p.P + row;
        const double *Ec = p.E + (size_t)K * col;
        double S = 0.0, c = 0.0;
        {
            int k = 0;
#pragma unroll 2
            for (; k + 4 <= K; k += 4) {                    // four classes per round, their loads issued together
                double e0, e1, e2, e3;
                if (vec2) {
                    const double2 a = *reinterpret_cast<const double2 *>(Ec + k), b = *reinterpret_cast<const double2 *>(Ec + k + 2);
                    e0 = a.x; e1 = a.y; e2 = b.x; e3 = b.y;
                } else {
                    e0 = Ec[k]; e1 = Ec[k + 1]; e2 = Ec[k + 2]; e3 = Ec[k + 3];
                }
                const double p0 = __ldg(Pr + (size_t)I * k), p1 = __ldg(Pr + (size_t)I * (k + 1));
                const double p2 = __ldg(Pr + (size_t)I * (k + 2)), p3 = __ldg(Pr + (size_t)I * (k + 3));
                S = fma(p0, e0, S); S = fma(p1, e1, S); S = fma(p2, e2, S); S = fma(p3, e3, S);
                const double q0 = p0 * e0, q1 = p1 * e1, q2 = p2 * e2, q3 = p3 * e3;
                const double c0 = c + q0, c1 = c0 + q1, c2 = c1 + q2;
                c = c2 + q3;
                buf[k * 32] = bycount_small ? c0 : q0;
                buf[(k + 1) * 32] = bycount_small ? c1 : q1;
                buf[(k + 2) * 32] = bycount_small ? c2 : q2;
                buf[(k + 3) * 32] = bycount_small ? c : q3;
            }
            for (; k < K; ++k) {
                const double ee = Ec[k], pe = __ldg(Pr + (size_t)I * k);
                S = fma(pe, ee, S);
                const double q0 = pe * ee;
                c = c + q0;
                buf[k * 32] = bycount_small ? c : q0;
            }
            // c_{K-1} is not a threshold.  A lane on the chain path keeps its K products but takes part in the warp's direct
            // draws with x < 0: its rows beyond K must not hold left-over shared memory (a garbage value <= x would send
            // the search past the lane's column)
            for (k = bycount_small ? K - 1 : K; k < R; ++k) buf[k * 32] = kInf;
        }
        // a batch of small counts: its claim has come back by now; fetch the next batch's list entries while this
        // one draws
        int nxt = 0, n_nxt = 0, row_nxt = 0, col_nxt = 0;
        if (early) {
            nxt = __shfl_sync(kFull, pend, 0);
            const int cell = nxt * 32 + lane;
            if (nxt < p.n_batches && cell < p.n_cells) {
                n_nxt = p.cell_n[cell]; row_nxt = (int)p.cell_row[cell]; col_nxt = p.cell_col[cell];
            }
        }
        SG_ZCLK(1)                                          // staging
        const bool ok = (S > 0.0 && S <= 1.7976931348623157e308);
        const bool small = valid && ok && bycount_small;
        int nrem = (valid && ok && !bycount_small) ? n : 0;
        const int nlast = (valid && !ok) ? n : 0;           // degenerate probabilities: all to the last class
        UStream us;
        us.init(p.key, p.sweep, p.stream, (uint32_t)(row + I * (col + p.col0)));

        // ---- direct method: draw t uses word t&3 of block t>>2
        if (__any_sync(kFull, small)) {
            const int tmax = __reduce_max_sync(kFull, small ? n : 0);
            const int nn = small ? n : 0;
            const double Sd = small ? S : -1.0;                                // other lanes search with x < 0: class 0, discarded
            SG_DISPATCH_NG(ng, z_direct, buf, us, ng, nn, tmax, Sd, p, zin_s, row, col)
            n_direct += (unsigned)nn;
        }

        SG_ZCLK(2)                                          // direct loop
        // ---- large counts
        if (__any_sync(kFull, nrem > 0)) {
            // visiting order: counting sort of the classes by d = min(7, expo(S) - expo(P*E)), ties by index
            if (nrem > 0) {
                const int eS = z_expo(S);
                unsigned lo = 0, hi = 0;                                       // eight byte-wide bucket counters
                for (int k = 0; k < K; ++k) {
                    const int d = max(0, min(7, eS - z_expo(buf[k * 32])));
                    if (d < 4) lo += 1u << (8 * d); else hi += 1u << (8 * (d - 4));
                }
                const unsigned tot_lo = (lo * 0x01010101u) >> 24;
                unsigned lo_off = lo * 0x01010100u;                            // byte b: number of classes in buckets below b
                unsigned hi_off = hi * 0x01010100u + tot_lo * 0x01010101u;
                for (int k = 0; k < K; ++k) {
                    const int d = max(0, min(7, eS - z_expo(buf[k * 32])));
                    unsigned pos;
                    if (d < 4) { pos = (lo_off >> (8 * d)) & 255u; lo_off += 1u << (8 * d); }
                    else { pos = (hi_off >> (8 * (d - 4))) & 255u; hi_off += 1u << (8 * (d - 4)); }
                    ord[pos * 32] = (unsigned char)k;
                }
            }
            SG_ZCLK(3)                                      // sort
            // conditional binomials in visiting order while more than kNTail mutations are left;
            // binomial t uses word t&3 of block t>>2
            double rem = S;
            int tl = 0;                                                        // classes this lane has visited
            U4 w4 = U4{0, 0, 0, 0};
            for (int t = 0; t < K - 1; ++t) {
                if (!__any_sync(kFull, nrem > kNTail)) break;
                if ((t & 3) == 0) w4 = us.words((uint32_t)t >> 2);             // converged: once per four binomials
                if (nrem > kNTail) {
                    const int cur = ord[t * 32];
                    const double best = buf[cur * 32];
                    if (best != 0.0) {
                        const double pp = best / rem;
                        const uint32_t wt = (t & 2) ? ((t & 1) ? w4.w : w4.z) : ((t & 1) ? w4.y : w4.x);
                        const int z = (pp > 0.0 && pp < 1.0) ? binomial_draw(us, t, u32(wt), nrem, pp) : nrem;
                        nrem -= z;
                        ++n_binom;
                        if (z) z_emit(p, zin_s, row, col, cur, z);
                    }
                    rem = rem - best;
                    tl = t + 1;
                }
            }
            SG_ZCLK(4)                                      // chain
            if (nrem > 0 && tl == K - 1) {                                      // only the class visited last is left
                z_emit(p, zin_s, row, col, ord[(K - 1) * 32], nrem);
                nrem = 0;
            }
            // the remainder goes to the unvisited classes by the direct method: the lane's column becomes the
            // cumulative sums in index order with the visited classes zeroed; draw d of a cell uses word d&3 of
            // block kTailBlock0 + (d>>2) of the cell's stream and the same search as the small counts.  The warp
            // shares this work: the cells' Philox blocks (four draws each) are dealt round-robin to the lanes,
            // whoever owns the cell -- a lane reads the owner's column of shared memory and adds the draws to the
            // statistics under the owner's context and genome.
            if (__any_sync(kFull, nrem > 0)) {
                double Sr = 0.0;
                if (nrem > 0) {
                    for (int t = 0; t < tl; ++t) buf[(int)ord[t * 32] * 32] = 0.0;
                    for (int k = 0; k < K; ++k) {
                        Sr = Sr + buf[k * 32];
                        buf[k * 32] = Sr;
                    }
                    for (int k = K - 1; k < R; ++k) buf[k * 32] = kInf;
                }
                const int units = (nrem + 3) >> 2;
                int incl = units;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
                const int excl = incl - units;
                const int utot = __shfl_sync(kFull, incl, 31);
                __syncwarp();
                SG_DISPATCH_NG(ng, z_tail, buf, p, zin_s, lane, ng, incl, excl, utot, nrem, Sr, row, col)
                __syncwarp();
                n_direct += (unsigned)nrem;
            }
        }

        SG_ZCLK(5)                                          // tail
        if (nlast) z_emit(p, zin_s, row, col, K - 1, nlast);                    // degenerate probabilities
        SG_ZCLK(6)                                          // degenerate cells
#ifdef SG_Z_TIMING
        if (lane == 0 && t_batch < (1 << 16)) g_z_batch_cycles[t_batch] = (unsigned)(clock64() - t_begin);
#endif
        if (!early) {
            if (lane == 0) pend = (int)(atomicAdd(p.counter, 1ULL) - p.base);
            nxt = __shfl_sync(kFull, pend, 0);
            const int cell = nxt * 32 + lane;
            if (nxt < p.n_batches && cell < p.n_cells) { n_nxt = p.cell_n[cell]; row_nxt = (int)p.cell_row[cell]; col_nxt = p.cell_col[cell]; }
        }
        cur = nxt; n = n_nxt; row = row_nxt; col = col_nxt;
    }
    if (p.draw_count) {
        const unsigned a = __reduce_add_sync(kFull, n_direct), b = __reduce_add_sync(kFull, n_binom);
        if (lane == 0) { atomicAdd(p.draw_count, (unsigned long long)a); atomicAdd(p.draw_count + 1, (unsigned long long)b); }
    }
    if (zin_s) {
        __syncthreads();
        for (int i = tid; i < I * K; i += blockDim.x) {
            const int v = zin_s[i];
            if (v) atomicAdd(p.Zin + i, v);
        }
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

// steps 6/7 (src/gibbs_2.cpp:172-251): gamma proposal, log acceptance ratio, uniform accept.
// The three terms of the ratio that depend only on the current A -- lgamma(A+1), lgamma(A^2/v), log A -- are kept
// in a per-element cache: A changes only when a proposal is accepted, and then its terms are exactly the
// y-terms just evaluated.  Same operations on the same operands, so the result is bit-identical to
// re-evaluating them (which is what the oracle does).
struct ACache { double *lg1, *lg2, *lg; };     // lgamma(A+1), lgamma(A*A/var), log(A): loaded and stored by the caller of update_a

SG_D double update_a(Key key, uint32_t sweep, uint32_t stream, uint32_t e, double A, double X, double B, double l, double var,
                                        double &c_lg1, double &c_lg2, double &c_lg)
{
    UStream us; us.init(key, sweep, stream, e);
    const double shape = (A * A) / var;
    const double rate = A / var;
    const double g = gamma_draw(us, shape, rate);
    const double y = (kTiny < g) ? g : kTiny;
    const double lgy1 = det_lgamma(y + 1), lgy2 = det_lgamma((y * y) / var), logy = det_log(y);
    const double lnp = (y - A) * (det_log(B) + det_log(X) - l)
                     + c_lg1 - lgy1
                     + c_lg2 - lgy2
                     + (((y * y) - (A * A)) / var) * det_log((A * y) / var)
                     + logy - c_lg;
    double pch = det_exp(lnp);
    if (X == 0) pch = (y < A) ? 1.0 : 0.0;
    double ub, um;
    us.pair(kAuxBlock, ub, um);
    if (um <= pch) {
        c_lg1 = lgy1; c_lg2 = lgy2; c_lg = logy;
        return y;
    }
    return A;
}

// (re)build the cache for a whole family: at chain creation and when var_ap / var_ae change
__global__ void a_cache_init(const double *A, size_t n, double var, ACache c)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double a = A[i];
    c.lg1[i] = det_lgamma(a + 1);
    c.lg2[i] = det_lgamma((a * a) / var);
    c.lg[i] = det_log(a);
}

// up to three step ids of one family, in permutation order, 4 bits each
struct Ops {
    int n; unsigned packed;
    __host__ __device__ int get(int o) const { return (int)((packed >> (4 * o)) & 15u); }
    __host__ void push(int s) { packed |= (unsigned)s << (4 * n); ++n; }
};

// ------------------------------------------------------------------------------------------------
// Ordered sum of a[0..n) for 32 consecutive elements e0 .. e0+31 (entry t of element e at a[t * stride + e]) by one
// block, draw specification v7: blocks of kSumBlk consecutive entries are added sequentially (from 0.0), the block
// sums form the next level, until one value is left.  Level 1 reads global memory: warp w takes the blocks w, w + 8,
// ..., lane = element, so every load is a coalesced row of 32 elements and a thread has 16 of them in flight.  The
// upper levels are a handful of adds per element in shared memory (lvl: [32][n1p], n1p = blocked_sum_pitch(n)).
// Call with all threads of the block; threads 0..31 get their element's total.
// ------------------------------------------------------------------------------------------------
constexpr int kSumBlk = 32;
constexpr int kPWarps = 16, kPElems = 32;   // p_family: 16 warps per block of 32 elements
__host__ __device__ inline int sum_blocks(int n) { return (n + kSumBlk - 1) / kSumBlk; }
__host__ __device__ inline int blocked_sum_pitch(int n) { return sum_blocks(n) | 1; }
__host__ __device__ inline size_t blocked_sum_smem_bytes(int n) { return sizeof(double) * kPElems * (size_t)blocked_sum_pitch(n); }

SG_D double blocked_sum_block(const double *a, int n, size_t stride, size_t e0, size_t n_elems, double *lvl)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n1 = sum_blocks(n), n1p = blocked_sum_pitch(n);
    const size_t e = e0 + lane;
    if (e < n_elems) {
        for (int b = warp; b < n1; b += kPWarps) {
            const int t0 = b * kSumBlk, cnt = min(kSumBlk, n - t0);
            const double *q = a + (size_t)t0 * stride + e;
            double v[16], sacc = 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (u < cnt) ? q[(size_t)u * stride] : 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) if (u < cnt) sacc = sacc + v[u];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (16 + u < cnt) ? q[(size_t)(16 + u) * stride] : 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) if (16 + u < cnt) sacc = sacc + v[u];
            lvl[lane * n1p + b] = sacc;
        }
    }
    __syncthreads();
    double tot = 0.0;
    if (tid < kPElems && e < n_elems && n > 0) {
        double *row = lvl + tid * n1p;
        int m = n1;
        while (m > 1) {                                    // in place: block b is written after everything it covers was read
            const int nb = sum_blocks(m);
            for (int b = 0; b < nb; ++b) {
                const int t0 = b * kSumBlk, cnt = min(kSumBlk, m - t0);
                double sacc = 0.0;
                for (int u = 0; u < cnt; ++u) sacc = sacc + row[t0 + u];
                row[b] = sacc;
            }
            m = nb;
        }
        tot = row[0];
    }
    return tot;
}

// ------------------------------------------------------------------------------------------------
// KP  p_family : gibbs_step2/4/6 (src/gibbs_2.cpp:106-120,140-153,172-210) on I x K.
// When step 2 runs, the whole block first adds the W E' segment sums of its elements in the specified order
// (blocked_sum_block, 32 elements at a time); then every warp applies the steps of this sweep to its elements.
// ------------------------------------------------------------------------------------------------
struct PParams {
    int IK;
    double *P, *Ap, *Bp;
    const int *Zin;
    const double *corrE_part;   // [n_part][I*K]: segment sums of W E' (unsharded) or the shards' ordered sums (sharded)
    int n_part;
    double *Ps_slot, *Aps_slot, *Bps_slot;   // sample slots or null (src/gibbs_2.cpp:325-333)
    ACache acache;
    Hyper h;
    Key key;
    uint32_t sweep;
    Ops ops;
};

// PER_WARP elements per warp, kPWarps * PER_WARP per block.
// The element updates are one long dependent chain per element, and a warp pays for every rare path (a rejected
// proposal, a ziggurat wedge, the logarithmic acceptance test) that ANY of its lanes takes.  With 32 elements in a warp
// that is all of them, every time; with two the chain stays near its expected length and the 16 warps of the block
// interleave on the four schedulers -- the right shape while I*K is a few thousand elements and the kernel is pure
// latency on the sweep's critical path.  Wide context sets (I*K in the tens of thousands) are throughput-bound
// instead and use full warps.
template <int PER_WARP>
__global__ void __launch_bounds__(32 * kPWarps) p_family(const PParams p)
{
    extern __shared__ __align__(16) double p_smem[];
    constexpr int kBlockElems = kPWarps * PER_WARP;
    __shared__ double tot_s[kBlockElems];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t e0 = (size_t)blockIdx.x * kBlockElems;
    bool has2 = false;
    for (int o = 0; o < p.ops.n; ++o) has2 |= (p.ops.get(o) == 2);
    const size_t eidx = e0 + (size_t)warp * PER_WARP + lane;
    const bool mine = lane < PER_WARP && eidx < (size_t)p.IK;
    const int e = (int)eidx;
    if (PER_WARP < 32) { if (lane >= PER_WARP) prefetch_tables((warp & 1) * 30 + lane - PER_WARP); }   // even warps: lines 0..29, odd warps: the rest
    else if (warp < 2) prefetch_tables(warp * 32 + lane);
    if (mine) {                                             // what the element updates will want, on its way before the sums
        sg_prefetch(p.P + e); sg_prefetch(p.Ap + e); sg_prefetch(p.Bp + e); sg_prefetch(p.Zin + e);
        sg_prefetch(p.acache.lg1 + e); sg_prefetch(p.acache.lg2 + e); sg_prefetch(p.acache.lg + e);
    }
    if (has2) {
        for (int sub = 0; sub < kBlockElems / kPElems; ++sub) {
            if (e0 + (size_t)sub * kPElems >= (size_t)p.IK) break;
            const double t = blocked_sum_block(p.corrE_part, p.n_part, (size_t)p.IK, e0 + (size_t)sub * kPElems, (size_t)p.IK, p_smem);
            if (tid < kPElems) tot_s[sub * kPElems + tid] = t;
            __syncthreads();                                // the level scratch is reused by the next 32 elements
        }
    }
    if (!mine) return;
    const double corrE = has2 ? tot_s[warp * PER_WARP + lane] : 0.0;
    double P = p.P[e], Ap = p.Ap[e], Bp = p.Bp[e];
    double lg1 = p.acache.lg1[e], lg2 = p.acache.lg2[e], lg = p.acache.lg[e];
    bool has6 = false;
    for (int o = 0; o < p.ops.n; ++o) {
        const int op = p.ops.get(o);
        if (op == 2) {
            UStream us; us.init(p.key, p.sweep, kStP, (uint32_t)e);
            const double shape = Ap + 1 + (double)p.Zin[e];
            const double rate = Bp + corrE;
            P = gamma_draw(us, shape, rate);
        } else if (op == 4) {
            Bp = update_b(p.key, p.sweep, kStBp, (uint32_t)e, Ap, P, p.h.ap, p.h.bp);
        } else if (op == 6) {
            Ap = update_a(p.key, p.sweep, kStAp, (uint32_t)e, Ap, P, Bp, p.h.lp, p.h.var_ap, lg1, lg2, lg);
            has6 = true;
        }
    }
    if (has6) { p.acache.lg1[e] = lg1; p.acache.lg2[e] = lg2; p.acache.lg[e] = lg; }
    p.P[e] = P; p.Ap[e] = Ap; p.Bp[e] = Bp;
    if (p.Ps_slot) p.Ps_slot[e] = P;
    if (p.Aps_slot) p.Aps_slot[e] = Ap;
    if (p.Bps_slot) p.Bps_slot[e] = Bp;
}
constexpr int kPWideFrom = 16384;   // I*K from which p_family uses full warps
inline size_t p_family_smem_bytes(int n_part) { return blocked_sum_smem_bytes(n_part); }

// column sharding: a shard's ordered sum of its own W E' segment sums, into its slot of the gathered buffer
__global__ void __launch_bounds__(32 * kPWarps) seg_reduce(const double *part, int n_seg, int IK, double *out_slot)
{
    extern __shared__ __align__(16) double p_smem[];
    const size_t e0 = (size_t)blockIdx.x * kPElems;
    const double tot = blocked_sum_block(part, n_seg, (size_t)IK, e0, (size_t)IK, p_smem);
    if (threadIdx.x < kPElems && e0 + threadIdx.x < (size_t)IK) out_slot[e0 + threadIdx.x] = tot;
}

// ------------------------------------------------------------------------------------------------
// KE  e_family : gibbs_step3/5/7 (src/gibbs_2.cpp:123-137,156-169,213-251) on K x J,
//     corrP = P' W (:128) in front, corrE = W E' (:111) segment sums behind.
//
// The unit of work is a warp task: one column group (32 consecutive genomes = one W E' segment of the draw
// specification, lane = genome) x one pair of classes, i.e. two elements per lane.  Tasks are numbered column group
// major and every block takes a CONTIGUOUS range of them (all blocks resident in one wave, the same number of tasks
// +-1 each), one task per warp per round; a round touches at most two column groups.  The W tile of a column group
// (up to kERows contexts x 32 genomes, 24 KB) is brought into shared memory ONCE per block by the bulk-copy engine
// (cp.async.bulk, one 256-byte row each, completion on an mbarrier) and serves both matrix passes:
//   pass 1  corrP[k,j] = fma chain over the contexts (lanes = genomes read a padded row: conflict-free),
//   ....    the element updates of this sweep's E-family steps, in permutation order,
//   pass 2  W E' of the segment for the task's two classes: lanes = contexts, the freshly drawn E values come
//           from the other lanes' registers by shuffle; the 32-term fma chain in genome order.
// A task never talks to another warp, so there is no block-wide barrier inside a round.  More than kERows contexts
// (pentanucleotide cohorts): the tile streams through the two slots in chunks of kERows rows, one column group per round.
// ------------------------------------------------------------------------------------------------
struct EParams {
    int I, J, Jp, K;
    int col0;                   // column shard: first global genome (random streams are addressed globally)
    const double *W;            // [I][Jp]
    const double *P;
    double *E, *Ae, *Be;
    const int *Znj;
    double *corrE_part;         // [n_seg][I*K]
    int do_corrE;
    double *Es_slot, *Aes_slot, *Bes_slot;
    ACache acache;
    Hyper h;
    Key key;
    uint32_t sweep;
    Ops ops;
    CUtensorMap wmap;           // W as a 2-D tensor [I][Jp], box = kEPitch genomes x kERows contexts
};

constexpr int kEWarps = 8;
constexpr int kEThreads = 32 * kEWarps;
constexpr int kERows = 96;        // contexts per staged chunk of W
constexpr int kEPitch = 34;       // doubles per staged row: 272 B keeps every row 16-byte aligned for the bulk copies; lanes
                                  // that walk down a column (pass 2) meet two-way bank conflicts only
__host__ __device__ constexpr int e_slots(int maxw) { return maxw == 8 ? 2 : 4; }
// the tile slots, a strip of 2 x kERows doubles per warp for the two classes' rows of P, slack for the 128-byte alignment
inline size_t e_smem_bytes(int maxw) { return sizeof(double) * ((size_t)e_slots(maxw) * kERows * kEPitch + (size_t)maxw * 2 * kERows) + 128; }

SG_D uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SG_D void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
SG_D void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SG_D void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA: one box of the 2-D tensor (x = first genome, y = first context) into shared memory, completion counted in bytes on
// the mbarrier; rows and columns beyond the tensor arrive as zeros (and count)
SG_D void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// The state of one element in registers: loaded and stored by the caller (a lane's two elements are neighbours in memory:
// one 16-byte access per array instead of two 8-byte ones when they are aligned).
struct EState { double E, Ae, Be, lg1, lg2, lg; int Znj; };

// inlined: the kernel's parameters are read from the constant bank (a call by reference would copy them to the stack,
// and every p.X inside would be a local-memory load -- 60 % of those missed L1 in the round-1 profile)
SG_D void e_update_element(const EParams &p, size_t e, double corrP, EState &s)
{
    const uint32_t eg = (uint32_t)(e + (size_t)p.K * p.col0);        // global element k + K*j of the streams
    for (int o = 0; o < p.ops.n; ++o) {
        const int op = p.ops.get(o);
        if (op == 3) {
            UStream us; us.init(p.key, p.sweep, kStE, eg);
            const double shape = s.Ae + 1 + (double)s.Znj;
            const double rate = s.Be + corrP;
            s.E = gamma_draw(us, shape, rate);
        } else if (op == 5) {
            s.Be = update_b(p.key, p.sweep, kStBe, eg, s.Ae, s.E, p.h.ae, p.h.be);
        } else if (op == 7) {
            s.Ae = update_a(p.key, p.sweep, kStAe, eg, s.Ae, s.E, s.Be, p.h.le, p.h.var_ae, s.lg1, s.lg2, s.lg);
        }
    }
}

// a lane's one or two elements (ea, and eb = ea + 1 when two): load
SG_D void e_load_pair(const EParams &p, size_t ea, bool two, bool vec, EState &a, EState &b)
{
    if (vec) {
        const double2 vE = *reinterpret_cast<const double2 *>(p.E + ea), vA = *reinterpret_cast<const double2 *>(p.Ae + ea);
        const double2 vB = *reinterpret_cast<const double2 *>(p.Be + ea);
        const double2 v1 = *reinterpret_cast<const double2 *>(p.acache.lg1 + ea), v2 = *reinterpret_cast<const double2 *>(p.acache.lg2 + ea);
        const double2 v3 = *reinterpret_cast<const double2 *>(p.acache.lg + ea);
        const int2 vz = *reinterpret_cast<const int2 *>(p.Znj + ea);
        a = EState{vE.x, vA.x, vB.x, v1.x, v2.x, v3.x, vz.x};
        b = EState{vE.y, vA.y, vB.y, v1.y, v2.y, v3.y, vz.y};
    } else {
        a = EState{p.E[ea], p.Ae[ea], p.Be[ea], p.acache.lg1[ea], p.acache.lg2[ea], p.acache.lg[ea], p.Znj[ea]};
        if (two) b = EState{p.E[ea + 1], p.Ae[ea + 1], p.Be[ea + 1], p.acache.lg1[ea + 1], p.acache.lg2[ea + 1], p.acache.lg[ea + 1], p.Znj[ea + 1]};
    }
}
SG_D void e_store2(double *base, size_t ea, bool two, bool vec, double x, double y)
{
    if (vec) *reinterpret_cast<double2 *>(base + ea) = make_double2(x, y);
    else { base[ea] = x; if (two) base[ea + 1] = y; }
}
SG_D void e_store_pair(const EParams &p, size_t ea, bool two, bool vec, bool has7, const EState &a, const EState &b)
{
    e_store2(p.E, ea, two, vec, a.E, b.E); e_store2(p.Ae, ea, two, vec, a.Ae, b.Ae); e_store2(p.Be, ea, two, vec, a.Be, b.Be);
    if (has7) {
        e_store2(p.acache.lg1, ea, two, vec, a.lg1, b.lg1); e_store2(p.acache.lg2, ea, two, vec, a.lg2, b.lg2);
        e_store2(p.acache.lg, ea, two, vec, a.lg, b.lg);
    }
    if (p.Es_slot) e_store2(p.Es_slot, ea, two, vec, a.E, b.E);
    if (p.Aes_slot) e_store2(p.Aes_slot, ea, two, vec, a.Ae, b.Ae);
    if (p.Bes_slot) e_store2(p.Bes_slot, ea, two, vec, a.Be, b.Be);
}

#ifdef SG_Z_TIMING      // experiments only: per-block phase clocks of the last e_family launch
__device__ long long g_e_clock[4096][8];
#define SG_ECLK(slot) if (threadIdx.x == 0 && blockIdx.x < 4096) g_e_clock[blockIdx.x][slot] = clock64();
#else
#define SG_ECLK(slot)
#endif

__host__ __device__ inline int e_pairs(int K) { return (K + 1) / 2; }
__host__ __device__ inline int e_groups(int J) { return (J + 31) / 32; }
// Warps per block.  A round takes one task per warp from at most two column groups (one when the contexts come in
// chunks).  Wherever a block's range starts inside a column group, pairs + 1 consecutive tasks always lie within two
// groups, so with at most that many warps a range of one task per warp is done in ONE round; few classes get narrower
// blocks (and more of them) instead of idle warps.
constexpr int kEWarpsWide = 16;
inline int e_block_warps(int I, int K)
{
    const int np = e_pairs(K);
    if (I > kERows) return std::max(1, std::min(kEWarpsWide, np));
    return std::max(1, std::min(kEWarps, np == 1 ? 2 : np + 1));
}

// launched with 32 * e_block_warps(I, K) threads.  MAXW = 8: contexts fit one tile (three blocks per SM); MAXW = 16: the
// contexts stream through in chunks, and one block takes ALL class pairs of a column group in one round (up to 16
// warps), so that the tile is streamed once per pass instead of once per pass and round
// MINB: resident blocks per SM the register budget is set for.  The narrow kernel comes in two budgets: 3 blocks (80
// registers, 24 warps per SM: config 3 at K = 19, 20 needs that many warp slots to stay in one round) and 2 blocks (128
// registers, nothing spilled: a task takes ~42 us instead of ~64 when the chain's tasks fit 16 warps per SM or need
// several rounds anyway); chain_create picks by the estimated rounds x round time.
template <int MAXW, int MINB>
__global__ void __launch_bounds__(32 * MAXW, MINB) e_family(const __grid_constant__ EParams p)
{
    constexpr int NS = e_slots(MAXW);                       // tile slots: two column groups, or a ring of context chunks
    extern __shared__ __align__(128) unsigned char e_smem[];
    __shared__ __align__(8) uint64_t mbar[NS];
    // [NS][kERows][kEPitch]; TMA wants its destination 128-byte aligned (a slot is 204 x 128 bytes)
    double *Ws = reinterpret_cast<double *>(e_smem + ((128u - (smem_u32(e_smem) & 127u)) & 127u));
    SG_ECLK(0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int I = p.I, J = p.J, K = p.K;
    const int npairs = e_pairs(K), n_tasks = e_groups(J) * npairs;
    const int nchunks = (I + kERows - 1) / kERows;
    // contiguous range of tasks; chunked contexts: whole column groups (a round cannot leave its group there)
    const int unit = nchunks == 1 ? 1 : npairs, n_units = n_tasks / unit;
    const int t_begin = (int)((long long)blockIdx.x * n_units / gridDim.x) * unit, t_end = (int)((long long)(blockIdx.x + 1) * n_units / gridDim.x) * unit;
    bool has3 = false;
    for (int o = 0; o < p.ops.n; ++o) has3 |= (p.ops.get(o) == 3);
    const bool need_w = has3 || p.do_corrE;
    const int groups_per_round = nchunks == 1 ? 2 : 1;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned nload[NS];                                     // tiles sent into each slot so far (uniform over the block)
#pragma unroll
    for (int s = 0; s < NS; ++s) nload[s] = 0;
    prefetch_tables(tid - 32);
    SG_ECLK(5)

    // thread 0 sends contexts [r0, r0 + kERows) of column group cg into `slot`: ONE request to the TMA unit (the box is
    // kEPitch = 34 genomes wide, so the tile lands with the padded pitch; a request per 256-byte row, as in the first
    // version of this kernel, kept the unit busy for ~20 k cycles per block: it is bound by requests, not bytes)
    auto send_tile = [&](int slot, int cg, int r0) {
        if (tid == 0) {
            uint64_t *bar = &mbar[slot];
            mbar_expect_tx(bar, (uint32_t)(kERows * kEPitch * sizeof(double)));
            tma_load_2d(Ws + (size_t)slot * kERows * kEPitch, &p.wmap, cg * 32, r0, bar);
        }
#pragma unroll
        for (int s = 0; s < NS; ++s) if (s == slot) ++nload[s];
    };
    auto wait_tile = [&](int slot) {
        unsigned n = 0;
#pragma unroll
        for (int s = 0; s < NS; ++s) if (s == slot) n = nload[s];
        mbar_wait(&mbar[slot], (n - 1u) & 1u);
    };
    // chunked contexts: the first NS chunks of column group cg go out
    auto start_stream = [&](int cg) {
        for (int c = 0; c < NS && c < nchunks; ++c) send_tile(c, cg, c * kERows);
    };

    for (int base = t_begin; base < t_end;) {
        const int cg0 = base / npairs;
        const int end = min(t_end, min(base + nwarps, (cg0 + groups_per_round) * npairs));
        const int task = base + warp;
        const bool active = task < end;
        const int cg = active ? task / npairs : cg0;
        const int pr = active ? task - cg * npairs : 0;
        const int ka = 2 * pr, kb = min(ka + 1, K - 1);
        const bool has_b = ka + 1 < K;
        const int j = cg * 32 + lane;
        const bool jin = active && j < J;
        const size_t ea = (size_t)ka + (size_t)K * min(j, J - 1), eb = (size_t)kb + (size_t)K * min(j, J - 1);
        if (base != t_begin) __syncthreads();               // every warp is done with the tiles of the previous round
        if (need_w) {
            if (nchunks == 1) {
                send_tile(0, cg0, 0);
                if ((end - 1) / npairs > cg0) send_tile(1, cg0 + 1, 0);
            } else {
                start_stream(cg0);
            }
        }
        // ---- pass 1: corrP = P'W for (ka, kb) x this lane's genome
        double acc0 = 0.0, acc1 = 0.0;
        if (has3) {
            for (int c = 0; c < nchunks; ++c) {
                const int slot = nchunks == 1 ? cg - cg0 : (c % NS);
                const int r0 = c * kERows, rows = min(kERows, I - r0);
                if (active) {
                    const double *Wt = Ws + (size_t)slot * kERows * kEPitch + lane;
                    const double *P0 = p.P + (size_t)I * ka + r0, *P1 = p.P + (size_t)I * kb + r0;
                    // The chunk's (up to) 96 rows of the two classes' columns of P: read coalesced (three rows per lane and
                    // class, on their way while the tile arrives) into the warp's own strip of shared memory, interleaved, so
                    // that a step of the chain is one broadcast 16-byte load for both classes next to the lane's W value.
                    // (Warp-uniform global loads inside the loop met a cold L1 in every launch -- two new lines per eight steps
                    // at L2 latency, 160 cycles per step; handing the values round by shuffle was bound by shuffle throughput,
                    // 96 shuffles per step and SM, 130 cycles per step.)
                    double *Pw = Ws + (size_t)NS * kERows * kEPitch + (size_t)warp * 2 * kERows;
                    __syncwarp();                           // the strip's previous chunk has been consumed by every lane
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int r = 32 * q + lane;
                        if (r < rows) *reinterpret_cast<double2 *>(Pw + 2 * r) = make_double2(__ldg(P0 + r), __ldg(P1 + r));
                    }
                    __syncwarp();
                    SG_ECLK(6)
                    wait_tile(slot);
                    SG_ECLK(7)
#pragma unroll 8
                    for (int r = 0; r < rows; ++r) {
                        const double w = Wt[r * kEPitch];
                        const double2 pp = *reinterpret_cast<const double2 *>(Pw + 2 * r);
                        acc0 = fma(pp.x, w, acc0);
                        acc1 = fma(pp.y, w, acc1);
                    }
                }
                if (nchunks > 1) {
                    __syncthreads();                        // the slot is free again
                    if (c + NS < nchunks) send_tile(slot, cg0, (c + NS) * kERows);
                }
            }
        }
        SG_ECLK(1)
        // ---- the element updates of this sweep, two elements per lane
        // the lane's element state (E, Ae, Be, sum_i Z, the three cached A terms)
        double Ea = 0.0, Eb = 0.0;
        if (jin) {
            EState sa{}, sb{};
            // neighbours, and 16-byte aligned in every array (the A-term caches and the sample slots are laid out K*J doubles
            // apart, so K*J must be even as well; always the case when K is even)
            const bool vec = has_b && ((ea & 1) == 0) && ((((size_t)K * J) & 1) == 0);
            e_load_pair(p, ea, has_b, vec, sa, sb);
            if (p.ops.n) {
                bool has7 = false;
                for (int o = 0; o < p.ops.n; ++o) has7 |= (p.ops.get(o) == 7);
#pragma unroll 1
                for (int el = 0; el < (has_b ? 2 : 1); ++el) {          // one copy of the inlined update code
                    EState s = el ? sb : sa;
                    e_update_element(p, el ? eb : ea, el ? acc1 : acc0, s);
                    if (el) sb = s; else sa = s;
                }
                e_store_pair(p, ea, has_b, vec, has7, sa, sb);
            }
            Ea = sa.E;
            if (has_b) Eb = sb.E;
        }
        SG_ECLK(2)
        if (!p.do_corrE) { base = end; continue; }
        // ---- pass 2: W E' of this segment for (ka, kb): lanes = contexts, E by shuffle, fma chain in genome order
        if (nchunks > 1 && has3) start_stream(cg0);         // restart the stream of chunks
        for (int c = 0; c < nchunks; ++c) {
            const int slot = nchunks == 1 ? cg - cg0 : (c % NS);
            const int r0 = c * kERows, rows = min(kERows, I - r0);
            if (active) {
                wait_tile(slot);
                const double *Wt = Ws + (size_t)slot * kERows * kEPitch;
                const int ra = lane, rb = lane + 32, rc = lane + 64;      // kERows == 96: three rows per lane
                const double *W0 = Wt + (size_t)min(ra, rows - 1) * kEPitch, *W1 = Wt + (size_t)min(rb, rows - 1) * kEPitch,
                             *W2 = Wt + (size_t)min(rc, rows - 1) * kEPitch;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
#pragma unroll 4
                for (int jj = 0; jj < 32; ++jj) {
                    const double xa = __shfl_sync(kFull, Ea, jj), xb = __shfl_sync(kFull, Eb, jj);
                    const double w0 = W0[jj], w1 = W1[jj], w2 = W2[jj];
                    a0 = fma(w0, xa, a0); a1 = fma(w1, xa, a1); a2 = fma(w2, xa, a2);
                    b0 = fma(w0, xb, b0); b1 = fma(w1, xb, b1); b2 = fma(w2, xb, b2);
                }
                double *da = p.corrE_part + (size_t)cg * ((size_t)I * K) + (size_t)I * ka + r0;
                double *db = p.corrE_part + (size_t)cg * ((size_t)I * K) + (size_t)I * kb + r0;
                if (ra < rows) da[ra] = a0;
                if (rb < rows) da[rb] = a1;
                if (rc < rows) da[rc] = a2;
                if (has_b) {
                    if (ra < rows) db[ra] = b0;
                    if (rb < rows) db[rb] = b1;
                    if (rc < rows) db[rc] = b2;
                }
            }
            if (nchunks > 1) {
                __syncthreads();
                if (c + NS < nchunks) send_tile(slot, cg0, (c + NS) * kERows);
            }
        }
        SG_ECLK(3)
        base = end;
    }
    SG_ECLK(4)
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
    // optional: the block that finishes last also runs the stride-halving tree over tree_a[0..n2) and writes
    // out[0] = tree_a[0] - sub[0] (one launch fewer per evaluation sweep); tree_a = null: tree_sum does it
    double *tree_a;
    int n2;
    const double *sub;
    double *out;
    unsigned *ticket;           // zero before the launch; reset by the last block
};

constexpr int kLLWarps = 8;       // warps per block
constexpr int kLLRound = 96;      // rows per round: three chunks; every warp evaluates 12 rows of a round

// 32 genomes per block, lanes = genomes.  A round covers 96 rows: warp w evaluates the terms of rows 12w .. 12w+11
// (four rows in flight: four independent fma chains over k) into shared memory; then the first three warps add the
// terms of one chunk each in row order, and the first warp adds the chunk sums to the genome's total in chunk order.
inline size_t loglik_smem_bytes(int K) { return sizeof(double) * ((size_t)K * 33 + (size_t)kLLRound * 33 + 3 * 32 + (size_t)K * kLLRound); }

__global__ void __launch_bounds__(32 * kLLWarps, 3) loglik_cols(const LLParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int I = p.I, J = p.J, K = p.K;
    double *E_s = reinterpret_cast<double *>(smem_raw);   // [K][33]
    double *term = E_s + K * 33;                          // [kLLRound][33]
    double *csum = term + kLLRound * 33;                  // [3][32]
    double *P_s = csum + 3 * 32;                          // [K][kLLRound]: this round's rows of P (warp-uniform global loads in
                                                          // the k loop met a cold L1 in every launch: one L2 round trip per class)
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int j0 = blockIdx.x * 32, j = j0 + lane, jc = min(j, J - 1);
    if (p.mode == 0) {
        for (int idx = tid; idx < K * 32; idx += 32 * kLLWarps) {
            const int jj = idx / K, k = idx - jj * K;
            E_s[k * 33 + jj] = (j0 + jj < J) ? p.E[(size_t)K * j0 + idx] : 0.0;
        }
    }
    double a = 0.0;
    for (int i0 = 0; i0 < I; i0 += kLLRound) {
        if (p.mode == 0) {
            if (i0) __syncthreads();                      // the previous round's rows of P have been used
            const int rows = min(kLLRound, I - i0);
            for (int idx = tid; idx < K * rows; idx += 32 * kLLWarps) {
                const int kk = idx / rows, r = idx - kk * rows;
                P_s[kk * kLLRound + r] = p.P[(size_t)I * kk + i0 + r];
            }
        }
        __syncthreads();                                  // E_s, P_s staged; term / csum of the previous round consumed
        for (int g = 0; g < 3; ++g) {
            const int rb = 12 * w + 4 * g, ib = i0 + rb;  // four rows ib .. ib+3 (clamped; rows beyond I are not added)
            if (ib >= I) break;
            const int ia = ib, ib1 = min(ib + 1, I - 1), ic = min(ib + 2, I - 1), id = min(ib + 3, I - 1);
            const double ma = (double)p.M[(size_t)ia * p.Jp + jc], mb = (double)p.M[(size_t)ib1 * p.Jp + jc];
            const double mc = (double)p.M[(size_t)ic * p.Jp + jc], md = (double)p.M[(size_t)id * p.Jp + jc];
            double ta, tb, tc, td;
            if (p.mode == 0) {
                const double wa = p.W[(size_t)ia * p.Jp + jc], wb = p.W[(size_t)ib1 * p.Jp + jc];
                const double wc = p.W[(size_t)ic * p.Jp + jc], wd = p.W[(size_t)id * p.Jp + jc];
                double pa = 0.0, pb = 0.0, pc = 0.0, pd = 0.0;
                const double *Pa = P_s + (ia - i0), *Pb = P_s + (ib1 - i0), *Pc = P_s + (ic - i0), *Pd = P_s + (id - i0);
#pragma unroll 4
                for (int k = 0; k < K; ++k) {                                 // four rows' chains side by side; P by broadcast from shared memory
                    const double e = E_s[k * 33 + lane];
                    pa = fma(Pa[k * kLLRound], e, pa); pb = fma(Pb[k * kLLRound], e, pb);
                    pc = fma(Pc[k * kLLRound], e, pc); pd = fma(Pd[k * kLLRound], e, pd);
                }
                const double la = pa * wa, lb = pb * wb, lc = pc * wc, ld = pd * wd;
                ta = ma * det_log(la) - la; tb = mb * det_log(lb) - lb;
                tc = mc * det_log(lc) - lc; td = md * det_log(ld) - ld;
            } else {
                ta = det_lgamma(ma + 1.0); tb = det_lgamma(mb + 1.0);
                tc = det_lgamma(mc + 1.0); td = det_lgamma(md + 1.0);
            }
            term[rb * 33 + lane] = ta; term[(rb + 1) * 33 + lane] = tb;
            term[(rb + 2) * 33 + lane] = tc; term[(rb + 3) * 33 + lane] = td;
        }
        __syncthreads();
        if (w < 3) {
            const int r0 = 32 * w, rows = min(kLLChunk, I - (i0 + r0));
            double ac = 0.0;
            for (int r = 0; r < rows; ++r) ac = ac + term[(r0 + r) * 33 + lane];
            csum[w * 32 + lane] = ac;
        }
        __syncthreads();
        if (w == 0) {
            const int nc = min(3, (I - i0 + kLLChunk - 1) / kLLChunk);
            for (int cc = 0; cc < nc; ++cc) a = a + csum[cc * 32 + lane];
        }
    }
    if (w == 0 && j < J) p.col[j] = a;
    if (!p.tree_a) return;
    // ---- last block: a[t] += a[t+s] for s = n2/2 .. 1 (the arithmetic of tree_sum), levels below 2048 in shared memory
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double *ta = p.tree_a;
    double *sh = reinterpret_cast<double *>(smem_raw);    // 2048 doubles fit: the term array alone is 96 * 33
    int sdist = p.n2 >> 1;
    for (; sdist >= 2048; sdist >>= 1) {                  // eight independent pairs in flight per thread
        for (int t0 = tid; t0 < sdist; t0 += 8 * 32 * kLLWarps) {
            double x[8], y[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * 32 * kLLWarps;
                if (t < sdist) { x[u] = ta[t]; y[u] = ta[t + sdist]; }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u * 32 * kLLWarps;
                if (t < sdist) ta[t] = x[u] + y[u];
            }
        }
        __syncthreads();
    }
    const int mlen = min(p.n2, 2048);
    for (int t = tid; t < mlen; t += 32 * kLLWarps) sh[t] = ta[t];
    __syncthreads();
    for (sdist = mlen >> 1; sdist >= 1; sdist >>= 1) {
        for (int t = tid; t < sdist; t += 32 * kLLWarps) sh[t] = sh[t] + sh[t + sdist];
        __syncthreads();
    }
    if (tid == 0) { p.out[0] = p.sub ? sh[0] - p.sub[0] : sh[0]; *p.ticket = 0u; }
}

// a[t] += a[t+s] for s = n2/2 .. 1 ; result written to out[0] = a[0] - sub (sub may be null).
// The last levels (s < 2048) run in shared memory; the arithmetic is the same.
__global__ void __launch_bounds__(1024) tree_sum(double *a, int n2, const double *sub, double *out)
{
    __shared__ double sh[4096];
    int s = n2 >> 1;
    for (; s >= 4096; s >>= 1) {
        for (int t = threadIdx.x; t < s; t += 1024) a[t] = a[t] + a[t + s];
        __syncthreads();
    }
    const int m = min(n2, 4096);
    for (int t = threadIdx.x; t < m; t += 1024) sh[t] = a[t];
    __syncthreads();
    for (s = m >> 1; s >= 1; s >>= 1) {
        for (int t = threadIdx.x; t < s; t += 1024) sh[t] = sh[t] + sh[t + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sub ? sh[0] - sub[0] : sh[0];
}

// ------------------------------------------------------------------------------------------------
// entry: reduce one piece of the reference's cube -- columns [ja, ja+nj) of slice k, column-major I x nj -- to its
// sums (the only use the sampler makes of its Z argument, src/gibbs_2.cpp:112,129).  One warp per column: coalesced
// reads along the contexts; the context sums go through a block-wide shared-memory accumulator when I fits.
// ------------------------------------------------------------------------------------------------
__global__ void zslab_reduce(const double *piece, int I, int nj, int ja, int K, int k, int use_smem, int *Zin, int *Znj)
{
    extern __shared__ int zin_blk[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (use_smem) {
        for (int i = threadIdx.x; i < I; i += blockDim.x) zin_blk[i] = 0;
        __syncthreads();
    }
    const int jb0 = blockIdx.x * 64, jb1 = min(nj, jb0 + 64);
    for (int jl = jb0 + warp; jl < jb1; jl += nw) {
        const double *col = piece + (size_t)I * jl;
        int colsum = 0;
        for (int i = lane; i < I; i += 32) {
            const int z = (int)col[i];
            colsum += z;
            if (z) {
                if (use_smem) atomicAdd(zin_blk + i, z);
                else atomicAdd(Zin + (size_t)i + (size_t)I * k, z);
            }
        }
        colsum = __reduce_add_sync(kFull, colsum);
        if (lane == 0) Znj[(size_t)k + (size_t)K * (ja + jl)] = colsum;
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < I; i += blockDim.x) {
            const int v = zin_blk[i];
            if (v) atomicAdd(Zin + (size_t)i + (size_t)I * k, v);
        }
    }
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
// posterior summaries on the device (R/Definition_SignExp_object.R:116-122 normalisation, :129-143 and
// :259,:297 medians; R/cpp_eBayesNMF.R:187-189): Phat = median_r P[i,k,r]/s_k[r], Ehat = median_r E[k,j,r]*s_k[r],
// s_k[r] = sum_i P[i,k,r].
// ------------------------------------------------------------------------------------------------
__global__ void sig_sums(const double *Ps, int I, int K, int R, double *s /* [R][K] */)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R * K) return;
    const int r = t / K, k = t - r * K;
    const double *col = Ps + (size_t)r * I * K + (size_t)I * k;
    double acc = 0.0;
    for (int i = 0; i < I; ++i) acc = acc + col[i];
    s[t] = acc;
}

// X [R][n] -> T [n][R], scaled by the sample's signature sum: mode 0 (signatures, element i + I*k): X / s ;
// mode 1 (exposures, element k + K*j): X * s.  32 x 32 tiles through shared memory: coalesced both ways.
__global__ void norm_transpose(const double *X, const double *s, int n, int R, int I, int K, int mode, double *T)
{
    __shared__ double tile[32][33];
    const int e0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {
        const int e = e0 + threadIdx.x, r = r0 + rr;
        double v = 0.0;
        if (e < n && r < R) {
            const int k = mode == 0 ? e / I : e % K;
            const double sk = s[(size_t)r * K + k];
            const double x = X[(size_t)r * n + e];
            v = mode == 0 ? x / sk : x * sk;
        }
        tile[rr][threadIdx.x] = v;
    }
    __syncthreads();
    for (int ee = threadIdx.y; ee < 32; ee += blockDim.y) {
        const int e = e0 + ee, r = r0 + threadIdx.x;
        if (e < n && r < R) T[(size_t)e * R + r] = tile[threadIdx.x][ee];
    }
}

// rank-th smallest (0-based) of R non-negative doubles by MSB-first radix selection on their bit patterns
__device__ unsigned long long radix_select(const double *v, int R, int rank, int *hist /* [256] shared, per warp */)
{
    const int lane = threadIdx.x & 31;
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = lane; b < 256; b += 32) hist[b] = 0;
        __syncwarp();
        for (int r = lane; r < R; r += 32) {
            const unsigned long long key = (unsigned long long)__double_as_longlong(v[r]);
            if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255ULL)], 1);
        }
        __syncwarp();
        int mine = 0;                                    // lane owns bins 8*lane .. 8*lane+7
        for (int b = 0; b < 8; ++b) mine += hist[8 * lane + b];
        int incl = mine;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
        const int excl = incl - mine;
        const bool here = rank >= excl && rank < incl;
        int bucket = 0, below = 0;
        if (here) {
            int cum = excl;
            for (int b = 0; b < 8; ++b) {
                const int h = hist[8 * lane + b];
                if (rank < cum + h) { bucket = 8 * lane + b; below = cum; break; }
                cum += h;
            }
        }
        const unsigned src = __ballot_sync(kFull, here);
        const int from = __ffs(src) - 1;
        bucket = __shfl_sync(kFull, bucket, from);
        below = __shfl_sync(kFull, below, from);
        rank -= below;
        prefix |= (unsigned long long)bucket << shift;
        mask |= 255ULL << shift;
        __syncwarp();
    }
    return prefix;
}

// one warp per element: median over the R samples of T[e][.] (mean of the two middle values when R is even, like R's median())
__global__ void median_rows(const double *T, int n, int R, double *out)
{
    __shared__ int hist[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e = blockIdx.x * 8 + warp;
    if (e >= n) return;
    const double *v = T + (size_t)e * R;
    const double a = __longlong_as_double((long long)radix_select(v, R, (R - 1) / 2, hist[warp]));
    const double b = (R & 1) ? a : __longlong_as_double((long long)radix_select(v, R, R / 2, hist[warp]));
    if (lane == 0) out[e] = (a + b) / 2.0;
}

// ------------------------------------------------------------------------------------------------
// same-device emulation of the all-reduce (shards sharing one GPU, for tests): every buffer <- sum of all
template <typename T>
__global__ void sum_shards(T *const *bufs, int n_shards, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T acc = bufs[0][i];
        for (int g = 1; g < n_shards; ++g) acc = acc + bufs[g][i];
        for (int g = 0; g < n_shards; ++g) bufs[g][i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// draw-rate ceiling: Philox4x32-10 block -> four 32-bit uniforms -> four inversion steps
// (compare against q^n, subtract), nothing read from memory; one store per thread keeps it alive
// ------------------------------------------------------------------------------------------------
__global__ void draw_ceiling(Key key, int iters, double r0, unsigned *sink)
{
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned acc = 0;
    double r = r0;
    for (int it = 0; it < iters; ++it) {
        const U4 w = philox4x32_10((uint32_t)it, e, 1u, 0u, key.k0, key.k1);
        const double u0 = u32(w.x), u1 = u32(w.y), u2 = u32(w.z), u3 = u32(w.w);
        acc += (u0 >= r) + (u1 >= r) + (u2 >= r) + (u3 >= r);
        r = r0 + (u0 - r) * 1e-9;                          // keep r data-dependent so nothing is hoisted
    }
    if (acc == 0xffffffffu) sink[0] = acc;
    if (e == 0) sink[1] = acc;
}

// ------------------------------------------------------------------------------------------------
// test hooks
// ------------------------------------------------------------------------------------------------
__global__ void test_math(int fn, const double *x, double *y, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    if (fn == 3) {                                       // ziggurat normal of element i, attempt (uint32) x[i], seed 12345
        UStream us; us.init(Key{12345u, 0u}, 0u, 3u, (uint32_t)i);
        const uint32_t t = (uint32_t)v;
        y[i] = normal_draw(us, t, us.words(t));
        return;
    }
    y[i] = fn == 0 ? det_log(v) : fn == 1 ? det_exp(v) : det_lgamma(v);
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
