// Posterior covariance slices through the low-rank structure of the GP prior.
//
// The reference builds K_k = (1-eps) SE(tau_k) + eps I per latent (funs/util.py:599-619) and inverts the dense
// qT x qT Hessian H = blkdiag(K_k^-1) + D, D = the per-bin q x q blocks W_t (funs/inference.py:50-65, :164-172).
// The squared-exponential part has a rapidly decaying spectrum: a pivoted Cholesky K_k - eps I = F_k F_k^T stops at
// rank r_k << T with residual below 1e-14.  With F = blkdiag(F_k) (qT x r, r = sum r_k) and the per-bin matrices
// P_t = (I + eps W_t)^-1, Dt_t = W_t P_t, the Woodbury identity gives EXACTLY (up to that residual)
//     Sigma = H^-1 = eps P + Y Y^T,   Y = P F L_b^-T,   L_b L_b^T = I_r + F^T Dt F        (r x r instead of qT x qT)
// so the slices the EM needs are
//     post_vsm[t]    = eps P_t + Y_(.,t) Y_(.,t)^T            (q x q per bin)
//     post_vsmGP[k]  = eps diag(P_t[k,k]) + Y_k Y_k^T         (T x T per latent)
// and the polishing Newton step is -Sigma g.  Neither K^-1 nor any qT x qT factorisation appears; cond(I + F^T Dt F)
// is ~1e2 where cond(H) ~ 1e3-1e5.  Work per trial at the headline shape (q=8, T=200, r ~ 220-380): ~0.15-0.3 GFLOP
// instead of 3.2 GFLOP.  The dense tiled path stays for priors whose numerical rank is not small (short timescales)
// and for the dense covariance output.
//
// The GEMM-shaped steps run on the FP64 tensor pipe (DMMA.8x8x4): F^T Dt F as ONE product per latent pair over all
// slots against a feature matrix (cap_features_kernel + gemm_nt_kernel with the scatter epilogue), F L_b^-T through the
// batched NT kernel (only the columns the triangular factor makes non-zero), the per-bin mixing and the per-bin slices
// as one-warp-per-bin DMMA sweeps over Y, and — what the timescale M-step actually consumes — the TRIAL-SUM of
// eps diag(P) + Y_k Y_k^T + m_k m_k^T (PautoSum, funs/learning.py:162-165) by the tile-pair SYRK kernels, so the per-trial
// T x T blocks are only written when a caller asks for post_vsmGP.  The r x r factorisation and triangular inverse are
// the tile kernels of factor.cu.
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "pgpfa_internal.h"

using namespace pgpfa;

namespace {

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

// ---------------------------------------------------------------------------------------------
// pivoted Cholesky of S_k = K_k - eps I, one CTA per latent:  F_k (T x T row-major [t][a], columns >= rank zero),
// Ft_k = F_k^T ([a][t]), rank[k].  Stops when the largest residual diagonal entry is <= delta.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pivchol_kernel(const double *__restrict__ K, int T, double eps, double delta,
                                                      double *__restrict__ F, double *__restrict__ Ft, int *__restrict__ rank) {
    extern __shared__ double sm[];
    double *dres = sm, *fj = sm + T;
    __shared__ double redv[8];
    __shared__ int redi[8];
    __shared__ double s_dmax;
    __shared__ int s_piv;
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *Kk = K + (size_t)k * T * T;
    double *Fk = F + (size_t)k * T * T, *Ftk = Ft + (size_t)k * T * T;
    for (int t = tid; t < T; t += blockDim.x) dres[t] = Kk[(size_t)t * T + t] - eps;
    __syncthreads();
    int r = 0;
    for (; r < T; r++) {
        double best = -1.0;
        int bi = T;
        for (int t = tid; t < T; t += blockDim.x)
            if (dres[t] > best) { best = dres[t]; bi = t; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double v2 = __shfl_xor_sync(0xffffffffu, best, o);
            const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
            if (v2 > best || (v2 == best && i2 < bi)) { best = v2; bi = i2; }
        }
        if (lane == 0) { redv[warp] = best; redi[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = redv[0];
            int i = redi[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); w++)
                if (redv[w] > b || (redv[w] == b && redi[w] < i)) { b = redv[w]; i = redi[w]; }
            s_dmax = b;
            s_piv = i;
        }
        __syncthreads();
        if (!(s_dmax > delta)) break;
        const int j = s_piv;
        for (int a = tid; a < r; a += blockDim.x) fj[a] = Fk[(size_t)j * T + a];
        __syncthreads();
        const double inv = 1.0 / sqrt(s_dmax);
        for (int t = tid; t < T; t += blockDim.x) {
            double s = Kk[(size_t)t * T + j] - (t == j ? eps : 0.0);
            const double *ft = Fk + (size_t)t * T;
            for (int a = 0; a < r; a++) s -= ft[a] * fj[a];
            const double c = s * inv;
            Fk[(size_t)t * T + r] = c;
            Ftk[(size_t)r * T + t] = c;
            if (t == j) dres[t] = -1.0;                       // pivoted: never selected again
            else if (dres[t] >= 0.0) dres[t] = fmax(dres[t] - c * c, 0.0);
        }
        __syncthreads();
    }
    if (tid == 0) rank[k] = r;
}

// ---------------------------------------------------------------------------------------------
// per (slot, bin): P = (I + eps W_t)^-1 and Dt = W_t P (both symmetric), layout (slot, q*q, T) like W
// ---------------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(128) lr_bins_kernel(const double *__restrict__ W, const int *__restrict__ act, int T,
                                                      double eps, double *__restrict__ Pm, double *__restrict__ Dm) {
    const int slot = blockIdx.y;
    const int trial = act ? act[slot] : slot;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double *Wt = W + (size_t)trial * Q * Q * T + t;
    double w[Q * (Q + 1) / 2], a[Q * (Q + 1) / 2];           // packed lower: (i,j) -> i(i+1)/2 + j
#pragma unroll
    for (int i = 0; i < Q; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            const double v = Wt[(size_t)(i * Q + j) * T];
            w[i * (i + 1) / 2 + j] = v;
            a[i * (i + 1) / 2 + j] = eps * v + (i == j ? 1.0 : 0.0);
        }
    // Cholesky a = L L^T in place
#pragma unroll
    for (int j = 0; j < Q; j++) {
        double dj = a[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < j; k++) dj -= a[j * (j + 1) / 2 + k] * a[j * (j + 1) / 2 + k];
        const double ljj = sqrt(dj), inv = 1.0 / ljj;
        a[j * (j + 1) / 2 + j] = ljj;
#pragma unroll
        for (int i = j + 1; i < Q; i++) {
            double s = a[i * (i + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; k++) s -= a[i * (i + 1) / 2 + k] * a[j * (j + 1) / 2 + k];
            a[i * (i + 1) / 2 + j] = s * inv;
        }
    }
    // a <- L^-1 in place (lower), column by column
#pragma unroll
    for (int j = 0; j < Q; j++) {
        a[j * (j + 1) / 2 + j] = 1.0 / a[j * (j + 1) / 2 + j];
    }
#pragma unroll
    for (int j = 0; j < Q; j++) {
#pragma unroll
        for (int i = j + 1; i < Q; i++) {
            // X(i,j) = -X(i,i) * sum_{k=j}^{i-1} L(i,k) X(k,j).  Columns are finished in increasing j and rows in increasing
            // i: row i still holds L(i,k) for k >= j, column j already holds X(k,j) for j <= k < i, the diagonal holds 1/L
            double s = 0.0;
#pragma unroll
            for (int k = j; k < i; k++) s += a[i * (i + 1) / 2 + k] * a[k * (k + 1) / 2 + j];     // L(i,k) X(k,j)
            a[i * (i + 1) / 2 + j] = -a[i * (i + 1) / 2 + i] * s;
        }
    }
    // P = X^T X (symmetric), Dt = W P
    double p[Q * (Q + 1) / 2];
#pragma unroll
    for (int i = 0; i < Q; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = i; k < Q; k++) s += a[k * (k + 1) / 2 + i] * a[k * (k + 1) / 2 + j];
            p[i * (i + 1) / 2 + j] = s;
        }
    double *Po = Pm + (size_t)slot * Q * Q * T + t, *Do = Dm + (size_t)slot * Q * Q * T + t;
#pragma unroll
    for (int i = 0; i < Q; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < Q; k++) {
                const double wik = (k <= i) ? w[i * (i + 1) / 2 + k] : w[k * (k + 1) / 2 + i];
                const double pkj = (k >= j) ? p[k * (k + 1) / 2 + j] : p[j * (j + 1) / 2 + k];
                s += wik * pkj;
            }
            const double pv = p[i * (i + 1) / 2 + j];
            Po[(size_t)(i * Q + j) * T] = pv;
            Po[(size_t)(j * Q + i) * T] = pv;
            Do[(size_t)(i * Q + j) * T] = s;
            Do[(size_t)(j * Q + i) * T] = s;
        }
}

// ---------------------------------------------------------------------------------------------
// batched C = A B^T on the FP64 tensor pipe.  A (M x K) and B (N x K) row-major along K.  One launch covers a
// table of problems (ragged blocks) times a batch.  CTA = 64 x 64 tile, 4 warps as 2 x 2 of 32 x 32; K in chunks of
// 16 through a three-stage ring of padded shared-memory stages filled by cp.async (the k-stride of 20 doubles keeps
// the 16 lanes of a half-warp on distinct 8-byte banks when they read their DMMA fragments).
// ---------------------------------------------------------------------------------------------
enum { GEMM_ADD_IDENTITY = 1, GEMM_SYMMETRIC = 2, GEMM_LOWER_ONLY = 4, GEMM_CAP_SCATTER = 8 };
struct GemmProb {
    long long a_off, b_off, c_off, s_off, d_off;   // element offsets into the batch's A, B, C, scale, dadd
    int M, N, K, flags;
    // GEMM_CAP_SCATTER (capacitance matrix as ONE product per latent pair over all slots): row m = a * cap_rl + b of the
    // product is entry (a, b) of the pair's block, column n is the slot: C[n * strideC + c_off + a * ldc + b]
    int cap_rl, pad;
};
struct GemmArgs {
    const double *A, *B;
    double *C;
    const double *scale;      // optional: B[n][k] is multiplied by scale[k]
    const double *dadd;       // optional: C[m][m] += dadd_alpha * dadd[m]
    long long strideA, strideB, strideC, strideS, strideD;
    int lda, ldb, ldc;
    const int *cmap;          // optional: C (and nothing else) is indexed by cmap[batch] instead of batch
    const GemmProb *probs;
    double dadd_alpha;
    int n_override = 0;       // when > 0 it replaces N of every problem (N = slots is only known at launch)
};

#define GN_STAGES 3
#define GN_LD 20                                   // padded k-stride of a staged row (doubles)
#define GN_STAGE_DOUBLES (2 * 64 * GN_LD + 16)     // A rows, B rows, 16 scale values
#define GN_SMEM (GN_STAGES * GN_STAGE_DOUBLES * 8)

// 8-byte asynchronous copy global -> shared (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one 16-wide k-chunk of a warp's MI x NJ live 8x8 blocks (compile-time shape: every issued MMA is a useful one)
template <int MI, int NJ>
__device__ __forceinline__ void gemm_nt_chunk(double (&acc)[4][4][2], const double *As, const double *Bs, const double *Ss,
                                              int wm, int wn, int fr, int fk) {
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
        double a[MI], bb[NJ];
        const double sv = Ss ? Ss[k4 + fk] : 1.0;
#pragma unroll
        for (int i = 0; i < MI; i++) a[i] = As[(wm * 32 + i * 8 + fr) * GN_LD + k4 + fk] * sv;
#pragma unroll
        for (int j = 0; j < NJ; j++) bb[j] = Bs[(wn * 32 + j * 8 + fr) * GN_LD + k4 + fk];
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], bb[j]);
    }
}

// K is consumed in chunks of 16 through a GN_STAGES-deep ring of shared-memory stages filled by cp.async, so the
// global loads of the next two chunks are in flight while the current one is multiplied.
__global__ void __launch_bounds__(128, 3) gemm_nt_kernel(GemmArgs g) {
    extern __shared__ __align__(16) double gsm[];
    GemmProb pr = g.probs[blockIdx.y];
    if (g.n_override > 0) pr.N = g.n_override;
    const int tm_n = (pr.M + 63) >> 6, tn_n = (pr.N + 63) >> 6;
    if ((int)blockIdx.x >= tm_n * tn_n) return;
    const int tm = blockIdx.x / tn_n, tn = blockIdx.x - tm * tn_n;
    if ((pr.flags & (GEMM_SYMMETRIC | GEMM_LOWER_ONLY)) && tn > tm) return;
    const int b = blockIdx.z;
    const double *A = g.A + (size_t)b * g.strideA + pr.a_off, *B = g.B + (size_t)b * g.strideB + pr.b_off;
    const double *sc = g.scale ? g.scale + (size_t)b * g.strideS + pr.s_off : nullptr;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    const int fr = lane >> 2, fk = lane & 3;
    const int m0 = tm * 64, n0 = tn * 64;
    const int lr = tid >> 4, lc = tid & 15;          // loader: rows lr + 8 j, column lc of the chunk
    // rows of this thread: lr + 8 j; out-of-range rows / columns are zero-filled by the copy (source clamped to row 0)
    const double *abase = A + (size_t)(m0 + lr) * g.lda, *bbase = B + (size_t)(n0 + lr) * g.ldb;
    const size_t astep = (size_t)8 * g.lda, bstep = (size_t)8 * g.ldb;
    auto stage_load = [&](int chunk, int st) {
        double *As = gsm + (size_t)st * GN_STAGE_DOUBLES, *Bs = As + 64 * GN_LD, *Ss = Bs + 64 * GN_LD;
        const int kk = chunk * 16 + lc;
        const bool kin = kk < pr.K;
        const int kc = kin ? kk : 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const bool am = kin && (m0 + lr + 8 * j < pr.M), bn = kin && (n0 + lr + 8 * j < pr.N);
            cp_async8(As + (lr + 8 * j) * GN_LD + lc, am ? abase + j * astep + kc : A, am ? 8 : 0);
            cp_async8(Bs + (lr + 8 * j) * GN_LD + lc, bn ? bbase + j * bstep + kc : B, bn ? 8 : 0);
        }
        if (sc && tid < 16) {
            const int ks = chunk * 16 + tid;
            cp_async8(Ss + tid, sc + (ks < pr.K ? ks : 0), ks < pr.K ? 8 : 0);
        }
    };
    double acc[4][4][2] = {};
    // 8-row / 8-column blocks of this warp's 32 x 32 patch that hold rows < M / columns < N (ragged edges are common:
    // T = 200 is 3 tiles + 8 rows, the rank blocks are 10-60 wide)
    const int mi = min(4, max(0, (pr.M - m0 - wm * 32 + 7) >> 3)), nj = min(4, max(0, (pr.N - n0 - wn * 32 + 7) >> 3));
    const int nchunks = (pr.K + 15) >> 4;
#pragma unroll
    for (int s0 = 0; s0 < GN_STAGES - 1; s0++) {
        if (s0 < nchunks) stage_load(s0, s0);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        cp_async_wait<GN_STAGES - 2>();
        __syncthreads();                    // chunk c has landed for everyone; stage (c-1) % GN_STAGES is free again
        if (c + GN_STAGES - 1 < nchunks) stage_load(c + GN_STAGES - 1, (c + GN_STAGES - 1) % GN_STAGES);
        cp_async_commit();
        const double *As = gsm + (size_t)(c % GN_STAGES) * GN_STAGE_DOUBLES, *Bs = As + 64 * GN_LD, *Ss = Bs + 64 * GN_LD;
        // The MMAs of a chunk for this warp's mi x nj live blocks, dispatched on the exact shape: a predicated-off
        // mma.sync still occupies the FP64 pipe (ncu: issued vs predicated-on DMMA counts), and ragged edges are the
        // normal case here (T = 200 is 3 tiles + 8 rows, rank blocks are 10-110 wide).
        switch (mi * 5 + nj) {
#define GN_CASE(MI_, NJ_) case MI_ * 5 + NJ_: gemm_nt_chunk<MI_, NJ_>(acc, As, Bs, sc ? Ss : nullptr, wm, wn, fr, fk); break;
#define GN_ROW(MI_) GN_CASE(MI_, 1) GN_CASE(MI_, 2) GN_CASE(MI_, 3) GN_CASE(MI_, 4)
            GN_ROW(1) GN_ROW(2) GN_ROW(3) GN_ROW(4)
#undef GN_ROW
#undef GN_CASE
            default: break;
        }
    }
    if (pr.flags & GEMM_CAP_SCATTER) {
        // transposed epilogue through shared memory: for a slot (column) the rows of the tile are runs of consecutive b,
        // i.e. contiguous pieces of a row of that slot's capacitance matrix -> coalesced stores
        __syncthreads();                                   // every warp is done with the operand stages
        double *Ts = gsm;                                  // [64 columns][65]
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int e = 0; e < 2; e++)
                    Ts[(wn * 32 + j * 8 + 2 * fk + e) * 65 + wm * 32 + i * 8 + fr] = acc[i][j][e];
        __syncthreads();
        for (int idx = tid; idx < 64 * 64; idx += 128) {
            const int nl = idx >> 6, ml = idx & 63;
            const int m = m0 + ml, n = n0 + nl;
            if (m >= pr.M || n >= pr.N) continue;
            const int ca = m / pr.cap_rl, cbb = m - ca * pr.cap_rl;
            double v = Ts[nl * 65 + ml];
            if ((pr.flags & GEMM_ADD_IDENTITY) && ca == cbb) v += 1.0;
            g.C[(size_t)n * g.strideC + pr.c_off + (size_t)ca * g.ldc + cbb] = v;
        }
        return;
    }
    const size_t cb = (size_t)(g.cmap ? g.cmap[b] : b) * g.strideC + pr.c_off;
    const double *dd = g.dadd ? g.dadd + (size_t)b * g.strideD + pr.d_off : nullptr;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int m = m0 + wm * 32 + i * 8 + fr, n = n0 + wn * 32 + j * 8 + 2 * fk + e;
                if (m >= pr.M || n >= pr.N) continue;
                double v = acc[i][j][e];
                if (m == n) {
                    if (pr.flags & GEMM_ADD_IDENTITY) v += 1.0;
                    if (dd) v += g.dadd_alpha * dd[m];
                }
                g.C[cb + (size_t)m * g.ldc + n] = v;
                if ((pr.flags & GEMM_SYMMETRIC) && tn < tm) g.C[cb + (size_t)n * g.ldc + m] = v;
            }
}

// ---------------------------------------------------------------------------------------------
// PautoSum through the low-rank factor (funs/learning.py:162-165): the timescale M-step only consumes
//     PautoSum[k] = sum_trials ( post_vsmGP[k] + m_k m_k^T ),   post_vsmGP[k] = eps diag(P_t[k,k]) + Y_k Y_k^T,
// so in the EM loop the per-trial T x T blocks never have to reach HBM: S_k = sum_slots Y_k Y_k^T is ONE symmetric
// product per latent with the reduction dimension (slot, c) of length slots * r.
//
// The lower block triangle of the T x T output (8 x 8 DMMA blocks) is cut into square tiles of tb = 8 blocks (64 x 64)
// plus a remainder strip.  A CTA = one tile pair x one latent x one part of the slots; its 8 warps form a 4 x 2 grid
// and every warp owns a FULL (tb/4) x (tb/2) rectangle of accumulator blocks, the same compile-time shape for the
// whole launch.  That matters twice (both measured with ncu on the first versions of this kernel): a predicated-off
// mma.sync still occupies the FP64 pipe for its 16 cycles (220 M DMMA issued for 128 M useful), and per-warp shapes
// through a switch over template instances thrash the instruction cache (stall "no instruction" 6 per issue).  In a
// diagonal tile the two warps whose rectangle lies above the diagonal only help with the loads.  The strip (T = 200:
// one block row of 25) goes through a generic variant of the same kernel with run-time rectangle bounds.
// Operands arrive through a 3-stage cp.async ring of 16-wide k-chunks that runs on across slot boundaries; one
// 16-byte shared-memory load per lane feeds two k4-steps.  Parts are sized by the pairs' block counts so that all CTAs
// carry the same work; every CTA writes its partial tile, a second kernel adds the parts in a fixed order
// (deterministic) together with the eps diag(P) and m m^T terms.
// ---------------------------------------------------------------------------------------------
#define SY_MAXPAIRS 40
#define SY_LD 24                                     // row pitch (doubles): 16-byte fragment loads of a quarter-warp hit 8 distinct 16-byte banks
#ifndef SY_STAGES
#define SY_STAGES 3
#endif
struct SyrkPair { int r0, nr, c0, nc, nparts, part0, diag; long long out_off; };
struct SyrkArgs {
    const double *Y;          // (slots, q*T, r)
    double *partial;          // per latent: concatenated [pair][part][nr*nc]
    long long strideY, partial_per_latent;
    int r, T, q, nslots, npairs;
    int first, count;         // pairs [first, first + count) belong to this launch
    int dbg;                  // development switch (PGPFA_SYRK_DBG): 1 = loads only, 2 = MMAs only
    int strip_pair, strip_r0, strip_nr;   // strip of <= 16 rows fused into the diagonal tiles' CTAs (strip_nr = 0: none)
    SyrkPair pairs[SY_MAXPAIRS];
};

// 16-byte asynchronous copy global -> shared (both addresses 16-byte aligned); src_bytes in {0, 8, 16}, the rest is zero-filled
__device__ __forceinline__ void cp_async16(double *dst_smem, const double *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes)
                 : "memory");
}

// every warp owns a full MI x NJ rectangle, CTA tile (4 MI) x (2 NJ) blocks (the pair's nr = 32 MI, nc = 16 NJ: full
// tiles, no row checks anywhere).  V16: r is even, so every row of Y starts 16-byte aligned and the copies move two
// columns at a time.
template <int MI, int NJ, bool V16>
__global__ void __launch_bounds__(256, 2) syrk_sum_kernel(const __grid_constant__ SyrkArgs a) {
    constexpr int ROWS = 32 * MI, COLS = 16 * NJ, SROWS = 8 * MI;       // tile operands + the fused strip rows
    constexpr int STAGE = (ROWS + COLS + SROWS) * SY_LD;
    extern __shared__ __align__(16) double ssm[];
    const int k = blockIdx.y;
    int pidx = 0;
    while (pidx + 1 < a.count && (int)blockIdx.x >= a.pairs[pidx + 1].part0) pidx++;
    // everything the loops need from the argument block lives in registers (indexed reads of the parameter space are
    // LDC instructions with a long scoreboard: they were the top stall of the first version)
    const int r0 = a.pairs[pidx].r0, c0 = a.pairs[pidx].c0, diag = a.pairs[pidx].diag, nparts = a.pairs[pidx].nparts;
    const int part = blockIdx.x - a.pairs[pidx].part0;
    const long long out_off = a.pairs[pidx].out_off;
    const int r = a.r;
    const long long strideY = a.strideY;
    const int s_begin = (int)((long long)a.nslots * part / nparts), s_end = (int)((long long)a.nslots * (part + 1) / nparts);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fr = lane >> 2, fk = lane & 3;
    // warp rectangle: 4 row groups x 2 column halves (a warp lives on SM sub-partition warp % 4 with its own FP64 pipe)
    // In a diagonal tile two of the eight rectangles are not needed; which pipes they idle rotates with the CTA index
    // so that co-resident CTAs do not starve the same pipes.
    const int wn = warp >> 2, g = ((warp & 3) + (diag ? (int)blockIdx.x : 0)) & 3;
    int base_m = g * MI, base_n = wn * NJ;
    // A rectangle entirely above the diagonal of a diagonal tile is not needed.  When the remainder strip of the output
    // (<= 8 MI rows below the last full tile) is fused in, these warps multiply the strip rows — staged behind the
    // tile operands — with this tile's rows instead: strip x tile j comes out of diagonal CTA (j, j) for free.
    const bool above = diag && base_m + MI - 1 < base_n;
    const bool fuse_strip = diag && a.strip_nr > 0;
    const bool strip_warp = above && fuse_strip;
    const bool active = !above || strip_warp;
    int aoff = 0, boff = diag ? 0 : ROWS * SY_LD;                 // operand regions inside a stage (doubles)
    if (strip_warp) { aoff = (ROWS + COLS) * SY_LD; base_n = g * NJ; base_m = 0; }       // g in {0, 1}: the two column halves
    const int nchunks = (r + 15) >> 4;
    const int total = (s_end - s_begin) * nchunks;
    // loader.  V16: thread (lr = tid / 8, lc2 = tid % 8) copies columns 2 lc2, 2 lc2 + 1 of rows lr + 32 u;
    // else: thread (lr = tid / 16, lc = tid % 16) copies column lc of rows lr + 16 u.
    const int lr = V16 ? tid >> 3 : tid >> 4, lc = V16 ? 2 * (tid & 7) : tid & 15;
    constexpr int RSTEP = V16 ? 32 : 16;
    const double *pa = a.Y + (size_t)s_begin * strideY + ((size_t)k * a.T + r0 + lr) * r + lc;
    const double *pb = a.Y + (size_t)s_begin * strideY + ((size_t)k * a.T + c0 + lr) * r + lc;
    const double *ps = a.Y + (size_t)s_begin * strideY + ((size_t)k * a.T + min(a.strip_r0 + lr, a.T - 1)) * r + lc;
    const size_t rstep = (size_t)RSTEP * r;
    int ld_chunk = 0;
    auto stage_load = [&](int st) {
        double *As = ssm + (size_t)st * STAGE + lr * SY_LD + lc, *Bs = As + ROWS * SY_LD;
        const int left = r - (ld_chunk * 16 + lc);                 // columns left from this thread's first one
        const int nb = V16 ? (left >= 2 ? 16 : (left == 1 ? 8 : 0)) : (left >= 1 ? 8 : 0);
        const double *sa = nb ? pa : a.Y, *sb = nb ? pb : a.Y;
#pragma unroll
        for (int u = 0; u < ROWS / RSTEP; u++) {
            if (V16) cp_async16(As + u * RSTEP * SY_LD, sa + (nb ? u * rstep : 0), nb);
            else cp_async8(As + u * RSTEP * SY_LD, sa + (nb ? u * rstep : 0), nb);
        }
        if (!diag) {
#pragma unroll
            for (int u = 0; u < COLS / RSTEP; u++) {
                if (V16) cp_async16(Bs + u * RSTEP * SY_LD, sb + (nb ? u * rstep : 0), nb);
                else cp_async8(Bs + u * RSTEP * SY_LD, sb + (nb ? u * rstep : 0), nb);
            }
        } else if (fuse_strip && lr < SROWS) {             // strip rows (zero-filled beyond strip_nr)
            const int nbs = lr < a.strip_nr ? nb : 0;
            double *Ss = ssm + (size_t)st * STAGE + (ROWS + COLS + lr) * SY_LD + lc;
            if (V16) cp_async16(Ss, nbs ? ps : a.Y, nbs);
            else cp_async8(Ss, nbs ? ps : a.Y, nbs);
        }
        pa += 16; pb += 16; ps += 16;
        if (++ld_chunk == nchunks) {
            ld_chunk = 0;
            pa += strideY - (size_t)nchunks * 16; pb += strideY - (size_t)nchunks * 16; ps += strideY - (size_t)nchunks * 16;
        }
    };
    double acc[MI][NJ][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
#pragma unroll
    for (int s0 = 0; s0 < SY_STAGES - 1; s0++) {
        if (s0 < total) stage_load(s0);
        cp_async_commit();
    }
    int st_cur = 0, st_load = SY_STAGES - 1;
    for (int it = 0; it < total; it++) {
        cp_async_wait<SY_STAGES - 2>();
        __syncthreads();
        if (it + SY_STAGES - 1 < total && a.dbg != 2) stage_load(st_load);
        cp_async_commit();
        if (active && a.dbg != 1) {
            const double *As = ssm + (size_t)st_cur * STAGE + aoff;
            const double *Bs = ssm + (size_t)st_cur * STAGE + boff;
            // one 16-byte load per lane and block feeds TWO k4-steps: lane (fr, fk) takes columns 2fk, 2fk+1 of an 8-wide
            // k-group, .x goes into the first MMA and .y into the second.  Both operands use the same permutation of k
            // inside the group, which a dot product does not see.
#pragma unroll
            for (int k8 = 0; k8 < 16; k8 += 8) {
                double2 av[MI], bv[NJ];
#pragma unroll
                for (int i = 0; i < MI; i++)
                    av[i] = *reinterpret_cast<const double2 *>(As + ((base_m + i) * 8 + fr) * SY_LD + k8 + 2 * fk);
#pragma unroll
                for (int j = 0; j < NJ; j++)
                    bv[j] = *reinterpret_cast<const double2 *>(Bs + ((base_n + j) * 8 + fr) * SY_LD + k8 + 2 * fk);
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i].x, bv[j].x);
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NJ; j++) dmma884(acc[i][j][0], acc[i][j][1], av[i].y, bv[j].y);
            }
        }
        st_cur = (st_cur + 1 == SY_STAGES) ? 0 : st_cur + 1;
        st_load = (st_load + 1 == SY_STAGES) ? 0 : st_load + 1;
    }
    if (!active) return;
    if (strip_warp) {
        // strip x this tile's columns -> the strip pair's area, part = this CTA's part (same count of parts by plan)
        const int snc = a.pairs[a.strip_pair].nc;
        double *out = a.partial + (size_t)k * a.partial_per_latent + a.pairs[a.strip_pair].out_off + (size_t)part * a.strip_nr * snc;
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NJ; j++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int row = i * 8 + fr, col = r0 + (base_n + j) * 8 + 2 * fk + e;
                    if (row < a.strip_nr) out[(size_t)row * snc + col] = acc[i][j][e];
                }
        return;
    }
    double *out = a.partial + (size_t)k * a.partial_per_latent + out_off + (size_t)part * ROWS * COLS;
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NJ; j++) {
            // (blocks above the diagonal inside a kept rectangle hold correct values too; the finishing kernel only reads
            // the lower triangle)
            const int row = (base_m + i) * 8 + fr, col = (base_n + j) * 8 + 2 * fk;
            *reinterpret_cast<double2 *>(out + (size_t)row * COLS + col) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
}

// Remainder strip of the PautoSum product (rows [s0, T) that no full tile covers; T = 200: one 8-row block against all
// 25 column blocks).  Too thin for the tile kernel (a CTA-wide barrier per 16-wide chunk for a handful of MMAs), so
// the reduction dimension is split over WARPS instead: every warp owns a subset of the slots and keeps the
// accumulators of one strip block row against a group of SYS_NB column blocks; its MMAs are fed straight from L2 (one
// 8-byte load per lane and fragment — the rows were just streamed by the tile kernel), the fragments of the next k4-step
// are in flight while the current one is multiplied.  No shared memory, no barriers; each warp writes its own partial
// strip, the finishing kernel adds them in order.
#define SYS_NB 13
__global__ void __launch_bounds__(256, 2) syrk_strip_kernel(const __grid_constant__ SyrkArgs a, int pair_index, int warps_total,
                                                            int ngroups, int jb0) {
    const int r0 = a.pairs[pair_index].r0, nr = a.pairs[pair_index].nr, nc = a.pairs[pair_index].nc, pc0 = a.pairs[pair_index].c0;
    const long long out_off = a.pairs[pair_index].out_off;
    const int k = blockIdx.y, sbrow = blockIdx.z / ngroups, grp = blockIdx.z - sbrow * ngroups;
    const int lane = threadIdx.x & 31, gw = blockIdx.x * 8 + (threadIdx.x >> 5);      // global warp = part
    const int fr = lane >> 2, fk = lane & 3;
    const int row0 = r0 + sbrow * 8;                              // first row of this strip block row
    const int nbc = (row0 >> 3) + 1;                              // column blocks 0 .. own diagonal block are needed
    const int jb = jb0 + grp * SYS_NB;                            // first column block of this group
    const int nb = max(0, min(SYS_NB, nbc - jb));
    if (nb == 0) return;
    const int T = a.T, r = a.r;
    const int s_begin = (int)((long long)a.nslots * gw / warps_total), s_end = (int)((long long)a.nslots * (gw + 1) / warps_total);
    double acc[SYS_NB][2];
#pragma unroll
    for (int j = 0; j < SYS_NB; j++) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
    const bool aok = row0 + fr < T;
    const int aoff = min(row0 + fr, T - 1) * r;               // T r < 2^31 (T <= 640)
    int boff[SYS_NB];
    unsigned bok = 0;
#pragma unroll
    for (int j = 0; j < SYS_NB; j++) {
        const int brow = (jb + j) * 8 + fr;
        if (j < nb && brow < T) bok |= 1u << j;
        boff[j] = min(brow, T - 1) * r;
    }
    const int nk4 = (r + 3) >> 2;
    for (int slot = s_begin; slot < s_end; slot++) {
        const double *Yk = a.Y + (size_t)slot * a.strideY + (size_t)k * T * r;
        double av, bv[SYS_NB];
        auto fetch = [&](int k4, double &fa, double *fb) {
            const int c = k4 * 4 + fk;
            const bool cok = c < r;
            const double *col = Yk + (cok ? c : 0);
            fa = (aok && cok) ? col[aoff] : 0.0;
#pragma unroll
            for (int j = 0; j < SYS_NB; j++) fb[j] = (cok && ((bok >> j) & 1)) ? col[boff[j]] : 0.0;
        };
        fetch(0, av, bv);
        for (int k4 = 0; k4 < nk4; k4++) {
            double an = 0.0, bn[SYS_NB];
            if (k4 + 1 < nk4) fetch(k4 + 1, an, bn);
#pragma unroll
            for (int j = 0; j < SYS_NB; j++) dmma884(acc[j][0], acc[j][1], av, bv[j]);       // full frame: no predicated-off MMA
            av = an;
#pragma unroll
            for (int j = 0; j < SYS_NB; j++) bv[j] = bn[j];
        }
    }
    // partial strip of this warp: [part = gw][nr rows][nc = T cols] inside the pair's area
    double *out = a.partial + (size_t)k * a.partial_per_latent + out_off + (size_t)gw * nr * nc;
#pragma unroll
    for (int j = 0; j < SYS_NB; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int row = sbrow * 8 + fr, col = (jb + j) * 8 + 2 * fk + e - pc0;     // column inside the pair
            if (j < nb && row < nr && col >= 0 && col < nc) out[(size_t)row * nc + col] = acc[j][e];
        }
}

// Corner of the PautoSum product when the strip is fused into the tile kernel: strip rows x strip rows (<= 8 x 8).
// CTA = one latent x one part of the slots; threads stride (slot, column), 36 running products each, block reduction.
__global__ void __launch_bounds__(256) syrk_corner_kernel(const __grid_constant__ SyrkArgs a, int pair_index) {
    __shared__ double red[32];
    const int k = blockIdx.y, part = blockIdx.x;
    const int r0 = a.pairs[pair_index].r0, nr = a.pairs[pair_index].nr, nparts = a.pairs[pair_index].nparts;
    const int s_begin = (int)((long long)a.nslots * part / nparts), s_end = (int)((long long)a.nslots * (part + 1) / nparts);
    double acc[36];
#pragma unroll
    for (int i = 0; i < 36; i++) acc[i] = 0.0;
    const long long items = (long long)(s_end - s_begin) * a.r;
    for (long long it = threadIdx.x; it < items; it += blockDim.x) {
        const int slot = s_begin + (int)(it / a.r), c = (int)(it % a.r);
        const double *col = a.Y + (size_t)slot * a.strideY + ((size_t)k * a.T + r0) * a.r + c;
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = (i < nr) ? col[(size_t)i * a.r] : 0.0;
        int idx = 0;
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) { acc[idx] = fma(v[i], v[j], acc[idx]); idx++; }
    }
    double *out = a.partial + (size_t)k * a.partial_per_latent + a.pairs[pair_index].out_off + (size_t)part * nr * nr;
    int idx = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            const double sacc = block_sum(acc[idx], red);
            idx++;
            if (threadIdx.x == 0 && i < nr) out[i * nr + j] = sacc;
        }
}

// dsum[k][t] = sum_slots P_t[k,k] (the eps diag(P) term of post_vsmGP summed over the slots), one CTA per (t, k)
__global__ void __launch_bounds__(128) syrk_pdiag_kernel(const double *__restrict__ Pm, int nslots, int q, int T,
                                                         double *__restrict__ dsum) {
    __shared__ double red[32];
    const int t = blockIdx.x, k = blockIdx.y;
    double sacc = 0.0;
    for (int sl = threadIdx.x; sl < nslots; sl += blockDim.x) sacc += Pm[((size_t)sl * q * q + k * q + k) * T + t];
    sacc = block_sum(sacc, red);
    if (threadIdx.x == 0) dsum[(size_t)k * T + t] = sacc;
}

// PautoSum[k][s][t] (+)= sum_parts partial + [s == t] eps sum_slots P_t[k,k] + sum_slots m[trial,k,s] m[trial,k,t]
// CTA = 16 x 16 output tile of one latent; the posterior means of 32 slots at a time go through shared memory (each
// value is then used 16 times), all sums in a fixed order.
__global__ void __launch_bounds__(256) syrk_finish_kernel(const __grid_constant__ SyrkArgs a, const double *__restrict__ dsum, const double *__restrict__ m,
                                                          const int *__restrict__ act, double eps, int accumulate,
                                                          double *__restrict__ Pout) {
    __shared__ double ms[32][17], mt[32][17];
    const int k = blockIdx.z, T = a.T, q = a.q;
    const int ts = threadIdx.x >> 4, tt = threadIdx.x & 15;
    const int s = blockIdx.y * 16 + ts, t = blockIdx.x * 16 + tt;
    const bool in = s < T && t < T;
    double v = 0.0;
    if (in) {
        const int rs = max(s, t), ct = min(s, t);
        for (int p = 0; p < a.npairs; p++) {
            const int r0 = a.pairs[p].r0, nr = a.pairs[p].nr, c0 = a.pairs[p].c0, nc = a.pairs[p].nc;
            if (rs < r0 || rs >= r0 + nr || ct < c0 || ct >= c0 + nc) continue;
            const double *src = a.partial + (size_t)k * a.partial_per_latent + a.pairs[p].out_off + (size_t)(rs - r0) * nc + (ct - c0);
            const int np = a.pairs[p].nparts;
            for (int j = 0; j < np; j++) v += src[(size_t)j * nr * nc];
            break;
        }
    }
    const size_t strideM = (size_t)q * T;
    const double *mp = m + (size_t)k * T;
    double d0 = 0.0;
    for (int sl0 = 0; sl0 < a.nslots; sl0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * 32 * 16; i += 256) {
            const int which = i >> 9, sl = (i >> 4) & 31, c = i & 15;
            const int slot = sl0 + sl;
            const int idx = (which ? blockIdx.x : blockIdx.y) * 16 + c;
            double val = 0.0;
            if (slot < a.nslots && idx < T) val = mp[(size_t)(act ? act[slot] : slot) * strideM + idx];
            (which ? mt : ms)[sl][c] = val;
        }
        __syncthreads();
#pragma unroll 8
        for (int sl = 0; sl < 32; sl++) d0 = fma(ms[sl][ts], mt[sl][tt], d0);
    }
    if (!in) return;
    v += d0;
    if (s == t) v += eps * dsum[(size_t)k * T + t];
    double *o = Pout + (size_t)k * T * T + (size_t)s * T + t;
    *o = (accumulate ? *o : 0.0) + v;
}

// Phi[(a,b)][t] = Ft_k[a][t] Ft_l[b][t] for the latent pair of this problem: the rows of the big capacitance product.
// grid = (row groups, pairs); the pair's problem entry carries the row offset (a_off / T) and shapes.
__global__ void __launch_bounds__(256) cap_features_kernel(const double *__restrict__ Ft, const GemmProb *__restrict__ probs,
                                                           const int2 *__restrict__ kl, int T, double *__restrict__ Phi) {
    const GemmProb pr = probs[blockIdx.y];
    const int k = kl[blockIdx.y].x, l = kl[blockIdx.y].y;
    const double *Fk = Ft + (size_t)k * T * T, *Fl = Ft + (size_t)l * T * T;
    double *out = Phi + pr.a_off;
    const long long total = (long long)pr.M * T;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(e / T), t = (int)(e - (long long)m * T);
        const int ca = m / pr.cap_rl, cbb = m - ca * pr.cap_rl;
        out[e] = Fk[(size_t)ca * T + t] * Fl[(size_t)cbb * T + t];
    }
}

// Zd[slot][c][a] = (L_b^-1)[c][a] from the packed-upper tiles ZT = L_b^-T (row-major r x r, lower triangular)
__global__ void zt_to_dense_lower_kernel(const double *__restrict__ ZT, int nb, int r, double *__restrict__ Zd) {
    const int slot = blockIdx.y;
    const double *Zs = ZT + (size_t)slot * ((size_t)nb * (nb + 1) / 2) * PGPFA_TILE;
    double *out = Zd + (size_t)slot * r * r;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)r * r; e += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(e / r), a = (int)(e - (size_t)c * r);
        out[e] = (a <= c) ? Zs[utile(a >> 6, c >> 6, nb) * PGPFA_TILE + tile_off(a & 63, c & 63)] : 0.0;
    }
}

// Y[(k,t), :] = sum_l P_t[k,l] Yh[(l,t), :]   in place; one CTA handles LR_TB consecutive bins of a slot.  The same pass
// takes this bin group's share of Y^T g (needed by the polishing Newton step): upart[slot][group][c] =
// sum_{k, t in group} Y[(k,t), c] g[(k,t)], so Y is not streamed a second time for it; lr_usum_kernel adds the groups in
// order.
#define LR_TB 8
struct LrOffs { int off[PGPFA_QMAX]; };      // first column of each latent in the rank-r factor
// One warp per bin.  Per 8-column chunk the mixing is the product  Y^T[c][k] = sum_l Yh^T[c][l] P_t[k][l]  on DMMA.8x8x4:
// the chunk of Yh is the A operand (row = column c of Y, loaded as 64-byte runs of the latents' rows), P_t sits in the B
// fragments for the whole row sweep, and the C fragment has lane (fr, fk) holding column c0 + fr of latents 2 fk, 2 fk + 1
// - so the eight lanes of an fk write 64 contiguous bytes of one row of Y.  (The scalar form spent 64 shared-memory
// reads and 64 DFMAs per output element; a row-major C fragment wrote 8-byte pieces at a 16-byte stride, 2.5x the sectors.)
template <int Q>
__global__ void __launch_bounds__(256) lr_mix_kernel(double *__restrict__ Y, const double *__restrict__ Pm, int T, int r,
                                                     const double *__restrict__ gvec, const int *__restrict__ act,
                                                     double *__restrict__ upart, const LrOffs offs) {
    constexpr int NB = (Q + 7) / 8, KS = (Q + 3) / 4;
    extern __shared__ double us[];            // [LR_TB][r] partial Y^T g of each bin (only with upart)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
    const int slot = blockIdx.y, t0 = blockIdx.x * LR_TB, t = t0 + warp;
    double *Ys = Y + (size_t)slot * Q * T * r;
    if (t < T) {
        double pb[NB][KS], g0[NB], g1[NB];
        const int trial = (upart && act) ? act[slot] : slot;
#pragma unroll
        for (int nb = 0; nb < NB; nb++) {
            const int k = nb * 8 + fr;                  // B fragment: B[kdim = l][n = k] = P_t[k][l]
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                const int l = ks * 4 + fk;
                pb[nb][ks] = (k < Q && l < Q) ? Pm[((size_t)slot * Q * Q + k * Q + l) * T + t] : 0.0;
            }
            const int kc = nb * 8 + 2 * fk;             // C fragment columns of this lane: latents kc, kc + 1
            g0[nb] = (upart && kc < Q) ? gvec[(size_t)trial * Q * T + (size_t)kc * T + t] : 0.0;
            g1[nb] = (upart && kc + 1 < Q) ? gvec[(size_t)trial * Q * T + (size_t)(kc + 1) * T + t] : 0.0;
        }
        int offl[KS];
        const double *arow[KS];
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            const int l = ks * 4 + fk;
            offl[ks] = l < Q ? offs.off[l] : r;                      // columns left of off_l are structural zeros
            arow[ks] = Ys + ((size_t)(l < Q ? l : 0) * T + t) * r;
        }
#pragma unroll 4
        for (int c0 = 0; c0 < r; c0 += 8) {
            const int c = c0 + fr;                       // column of this lane in the A and the C fragment
            double av[KS];
#pragma unroll
            for (int ks = 0; ks < KS; ks++) {
                av[ks] = 0.0;
                if (c >= offl[ks] && c < r) av[ks] = arow[ks][c];
            }
            double acc[NB][2];
#pragma unroll
            for (int nb = 0; nb < NB; nb++) {
                acc[nb][0] = 0.0; acc[nb][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ks++) dmma884(acc[nb][0], acc[nb][1], av[ks], pb[nb][ks]);
            }
            double s = 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; nb++) {
                const int kc = nb * 8 + 2 * fk;
                if (c < r) {
                    if (kc < Q) Ys[((size_t)kc * T + t) * r + c] = acc[nb][0];
                    if (kc + 1 < Q) Ys[((size_t)(kc + 1) * T + t) * r + c] = acc[nb][1];
                }
                s = fma(acc[nb][0], g0[nb], s);
                s = fma(acc[nb][1], g1[nb], s);
            }
            if (upart) {
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                if (fk == 0 && c < r) us[warp * r + c] = s;
            }
        }
    }
    if (!upart) return;
    __syncthreads();
    const int nt = min(LR_TB, T - t0);
    for (int c = threadIdx.x; c < r; c += blockDim.x) {
        double ua = 0.0;
        for (int bb = 0; bb < nt; bb++) ua += us[bb * r + c];
        upart[((size_t)slot * gridDim.x + blockIdx.x) * r + c] = ua;
    }
}

// u[slot][c] = sum over the bin groups of upart[slot][group][c]
__global__ void lr_usum_kernel(const double *__restrict__ upart, int ngroups, int r, double *__restrict__ u) {
    const int slot = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= r) return;
    double s = 0.0;
    for (int g = 0; g < ngroups; g++) s += upart[((size_t)slot * ngroups + g) * r + c];
    u[(size_t)slot * r + c] = s;
}

// post_vsm[trial][t][k][l] = eps P_t[k,l] + sum_c Y[(k,t),c] Y[(l,t),c]: the q x q Gram matrix of the q rows of Y that
// belong to bin t, one warp per bin on the FP64 tensor pipe.  For m8n8k4 the A fragment (row = lane/4, k = lane%4) of a
// matrix and the B fragment (k = lane%4, col = lane/4) of its transpose are the same register, so one 8-byte load per
// lane and k-step feeds the MMA; Y is streamed from HBM exactly once.
template <int Q, bool STEP>
__global__ void __launch_bounds__(256) lr_vsm_kernel(const double *__restrict__ Y, const double *__restrict__ Pm,
                                                     const int *__restrict__ act, int T, int r, double eps,
                                                     double *__restrict__ vsm, const double *__restrict__ u,
                                                     const double *__restrict__ gvec, double *__restrict__ dx) {
    constexpr int QB = (Q + 7) / 8;
    const int slot = blockIdx.y, trial = act ? act[slot] : slot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * 8 + warp;
    extern __shared__ double u_s[];                       // Y^T g of the slot (only with dx), zero-padded to 4
    if (STEP) {
        for (int c = threadIdx.x; c < ((r + 3) & ~3); c += blockDim.x) u_s[c] = c < r ? u[(size_t)slot * r + c] : 0.0;
        __syncthreads();
    }
    if (t >= T) return;
    const int fr = lane >> 2, fk = lane & 3;
    const double *rowp[QB];
    bool rowok[QB];
#pragma unroll
    for (int bi = 0; bi < QB; bi++) {
        const int row = bi * 8 + fr;
        rowok[bi] = row < Q;
        rowp[bi] = Y + ((size_t)slot * Q * T + (size_t)(rowok[bi] ? row : 0) * T + t) * r;
    }
    // With dx the same sweep over the rows of Y also takes the polishing Newton step of this bin,
    // dx[(k,t)] = -(eps (P_t g_t)_k + Y[(k,t),:] u): the rows are streamed once for both.
    double acc[QB][QB][2] = {};
    double su[QB] = {};
#pragma unroll 4
    for (int k0 = 0; k0 < r; k0 += 4) {
        double a[QB];
#pragma unroll
        for (int bi = 0; bi < QB; bi++) a[bi] = (rowok[bi] && k0 + fk < r) ? rowp[bi][k0 + fk] : 0.0;
        if (STEP) {
            const double uv = u_s[k0 + fk];
#pragma unroll
            for (int bi = 0; bi < QB; bi++) su[bi] = fma(a[bi], uv, su[bi]);
        }
#pragma unroll
        for (int bi = 0; bi < QB; bi++)
#pragma unroll
            for (int bj = 0; bj < QB; bj++) dmma884(acc[bi][bj][0], acc[bi][bj][1], a[bi], a[bj]);
    }
    if (STEP) {
#pragma unroll
        for (int bi = 0; bi < QB; bi++) {
            su[bi] += __shfl_xor_sync(0xffffffffu, su[bi], 1);
            su[bi] += __shfl_xor_sync(0xffffffffu, su[bi], 2);
            const int k = bi * 8 + fr;
            if (fk == 0 && k < Q) {
                double pg = 0.0;
#pragma unroll
                for (int l = 0; l < Q; l++)
                    pg += Pm[((size_t)slot * Q * Q + k * Q + l) * T + t] * gvec[(size_t)trial * Q * T + (size_t)l * T + t];
                dx[(size_t)trial * Q * T + (size_t)k * T + t] = -(eps * pg + su[bi]);
            }
        }
    }
    double *out = vsm + ((size_t)trial * T + t) * Q * Q;
#pragma unroll
    for (int bi = 0; bi < QB; bi++)
#pragma unroll
        for (int bj = 0; bj < QB; bj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int k = bi * 8 + fr, l = bj * 8 + 2 * fk + e;
                if (k < Q && l < Q) out[k * Q + l] = acc[bi][bj][e] + eps * Pm[((size_t)slot * Q * Q + k * Q + l) * T + t];
            }
}

// u[slot][c] = sum_row Y[row][c] g[trial][row]
__global__ void __launch_bounds__(128) lr_ytg_kernel(const double *__restrict__ Y, const double *__restrict__ gvec,
                                                     const int *__restrict__ act, int n, int r, double *__restrict__ u) {
    const int slot = blockIdx.y, trial = act ? act[slot] : slot;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= r) return;
    const double *Ys = Y + (size_t)slot * n * r + c;
    const double *gt = gvec + (size_t)trial * n;
    double s0 = 0.0, s1 = 0.0;
    int row = 0;
    for (; row + 1 < n; row += 2) {
        s0 += Ys[(size_t)row * r] * gt[row];
        s1 += Ys[(size_t)(row + 1) * r] * gt[row + 1];
    }
    if (row < n) s0 += Ys[(size_t)row * r] * gt[row];
    u[(size_t)slot * r + c] = s0 + s1;
}

// dx[trial][(k,t)] = -( eps sum_l P_t[k,l] g[(l,t)] + sum_c Y[(k,t),c] u[c] )       warp per row
template <int Q>
__global__ void __launch_bounds__(256) lr_step_kernel(const double *__restrict__ Y, const double *__restrict__ u,
                                                      const double *__restrict__ Pm, const double *__restrict__ gvec,
                                                      const int *__restrict__ act, int T, int r, double eps,
                                                      double *__restrict__ dx) {
    const int slot = blockIdx.y, trial = act ? act[slot] : slot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = Q * T;
    const int row = blockIdx.x * 8 + warp;
    if (row >= n) return;
    const int k = row / T, t = row - k * T;
    const double *yr = Y + ((size_t)slot * n + row) * r, *us = u + (size_t)slot * r;
    double s = 0.0;
    for (int c = lane; c < r; c += 32) s += yr[c] * us[c];
    s = warp_sum(s);
    if (lane == 0) {
        double pg = 0.0;
#pragma unroll
        for (int l = 0; l < Q; l++)
            pg += Pm[((size_t)slot * Q * Q + k * Q + l) * T + t] * gvec[(size_t)trial * n + (size_t)l * T + t];
        dx[(size_t)trial * n + row] = -(eps * pg + s);
    }
}

}  // namespace

// =============================================================================================
// host side
// =============================================================================================
extern "C" int pgpfa_prior_lowrank(const double *K, int q, int T, double eps, double delta, double *F, double *Ft,
                                   int *rank, cudaStream_t st) {
    if (!K || !F || !Ft || !rank || q <= 0 || T <= 0 || !(delta > 0.0)) return PGPFA_ERR_ARG;
    PGPFA_CUDA_TRY(cudaMemsetAsync(F, 0, (size_t)q * T * T * 8, st));
    PGPFA_CUDA_TRY(cudaMemsetAsync(Ft, 0, (size_t)q * T * T * 8, st));
    const size_t smem = (size_t)2 * T * sizeof(double);
    if (smem > 48 * 1024)
        PGPFA_CUDA_TRY(cudaFuncSetAttribute(pivchol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pivchol_kernel<<<q, 256, smem, st>>>(K, T, eps, delta, F, Ft, rank);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

size_t pgpfa_i_lowrank_bytes_per_slot(int q, int T, int r) {
    const int nbr = pgpfa_nb(r);
    const size_t tiles = (size_t)(2 * pgpfa_ltiles(nbr) + nbr) * PGPFA_TILE * 8;
    return align_up(tiles) + 2 * align_up((size_t)r * r * 8) + align_up((size_t)q * T * r * 8) +
           2 * align_up((size_t)q * q * T * 8) + align_up((size_t)r * 8) + 4096;
}

namespace {
template <int Q>
int lr_launch_bins(const double *W, const int *act, int T, double eps, double *Pm, double *Dm, int nslots, cudaStream_t st) {
    dim3 grid((T + 127) / 128, nslots);
    lr_bins_kernel<Q><<<grid, 128, 0, st>>>(W, act, T, eps, Pm, Dm);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
template <int Q>
int lr_launch_mix(double *Y, const double *Pm, int T, int r, int nslots, cudaStream_t st, const double *gvec, const int *act,
                  double *upart, const LrOffs &offs) {
    dim3 grid((T + LR_TB - 1) / LR_TB, nslots);
    const size_t smem = upart ? (size_t)LR_TB * r * sizeof(double) : 0;
    if (smem > 48 * 1024)
        PGPFA_CUDA_TRY(cudaFuncSetAttribute(lr_mix_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lr_mix_kernel<Q><<<grid, 256, smem, st>>>(Y, Pm, T, r, gvec, act, upart, offs);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
template <int Q>
int lr_launch_vsm(const double *Y, const double *Pm, const int *act, int T, int r, double eps, double *vsm, int nslots,
                  cudaStream_t st, const double *u, const double *gvec, double *dx) {
    dim3 grid((T + 7) / 8, nslots);
    if (dx) lr_vsm_kernel<Q, true><<<grid, 256, (size_t)((r + 3) & ~3) * sizeof(double), st>>>(Y, Pm, act, T, r, eps, vsm, u, gvec, dx);
    else lr_vsm_kernel<Q, false><<<grid, 256, 0, st>>>(Y, Pm, act, T, r, eps, vsm, u, gvec, dx);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
template <int Q>
int lr_launch_step(const double *Y, const double *u, const double *Pm, const double *g, const int *act, int T, int r,
                   double eps, double *dx, int nslots, cudaStream_t st) {
    dim3 grid((Q * T + 7) / 8, nslots);
    lr_step_kernel<Q><<<grid, 256, 0, st>>>(Y, u, Pm, g, act, T, r, eps, dx);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int launch_gemm(const GemmArgs &g, const GemmProb *dprobs, int nprobs, int max_tiles, int batch, cudaStream_t st) {
    if (nprobs == 0 || max_tiles == 0 || batch <= 0) return PGPFA_OK;
    GemmArgs a = g;
    a.probs = dprobs;
    dim3 grid(max_tiles, nprobs, batch);
    static bool attr_set = false;
    if (!attr_set) {
        PGPFA_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_SMEM));
        attr_set = true;
    }
    gemm_nt_kernel<<<grid, 128, GN_SMEM, st>>>(a);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

// the three problem tables of a posterior pass: capacitance blocks, Yh per latent, T x T block per latent
struct LrTables {
    std::vector<GemmProb> cap, yh, blk, capbig;
    std::vector<int2> capbig_kl;
    long long phi_rows;              // rows of the feature matrix of the big capacitance product
    int cap_tiles, yh_tiles, blk_tiles, capbig_mtiles;
};
int max_tiles(const std::vector<GemmProb> &v) {
    int t = 0;
    for (const GemmProb &p : v) t = std::max(t, ((p.M + 63) / 64) * ((p.N + 63) / 64));
    return t;
}
LrTables lr_tables(const PgpfaLowRank &lr, int q, int T) {
    LrTables tb;
    const int r = lr.r;
    for (int k = 0; k < q; k++)
        for (int l = 0; l <= k; l++) {
            if (lr.rank[k] == 0 || lr.rank[l] == 0) continue;
            GemmProb pb;
            pb.a_off = (long long)k * T * T; pb.b_off = (long long)l * T * T;
            pb.c_off = (long long)lr.off[k] * r + lr.off[l];
            pb.s_off = (long long)(k * q + l) * T; pb.d_off = 0;
            pb.M = lr.rank[k]; pb.N = lr.rank[l]; pb.K = T;
            pb.flags = (k == l) ? (GEMM_ADD_IDENTITY | GEMM_LOWER_ONLY) : 0;
            pb.cap_rl = 0; pb.pad = 0;
            tb.cap.push_back(pb);
        }
    // the same blocks as ONE product per latent pair over all slots: rows (a, b), columns = slots, K = T
    tb.phi_rows = 0;
    tb.capbig_mtiles = 0;
    for (int k = 0; k < q; k++)
        for (int l = 0; l <= k; l++) {
            if (lr.rank[k] == 0 || lr.rank[l] == 0) continue;
            GemmProb pb;
            pb.a_off = tb.phi_rows * T; pb.b_off = (long long)(k * q + l) * T;
            pb.c_off = (long long)lr.off[k] * r + lr.off[l];
            pb.s_off = 0; pb.d_off = 0;
            pb.M = lr.rank[k] * lr.rank[l]; pb.N = 0; pb.K = T;        // N = slots, filled in at launch
            pb.flags = GEMM_CAP_SCATTER | ((k == l) ? GEMM_ADD_IDENTITY : 0);
            pb.cap_rl = lr.rank[l]; pb.pad = 0;
            tb.capbig.push_back(pb);
            tb.capbig_kl.push_back(make_int2(k, l));
            tb.phi_rows += pb.M;
            tb.capbig_mtiles = std::max(tb.capbig_mtiles, (pb.M + 63) / 64);
        }
    for (int l = 0; l < q; l++) {
        GemmProb pb;
        // Z = L^-1 is lower triangular: Z[c][off_l + a] = 0 for c < off_l, so latent l has no columns c < off_l
        // (never computed, never stored; the mixing pass reads them as zeros)
        pb.a_off = (long long)l * T * T; pb.b_off = lr.off[l] + (long long)lr.off[l] * r;
        pb.c_off = (long long)l * T * r + lr.off[l];
        pb.s_off = 0; pb.d_off = 0;
        pb.M = T; pb.N = r - lr.off[l]; pb.K = lr.rank[l]; pb.flags = 0; pb.cap_rl = 0; pb.pad = 0;
        tb.yh.push_back(pb);
    }
    for (int k = 0; k < q; k++) {
        GemmProb pb;
        pb.a_off = (long long)k * T * r; pb.b_off = pb.a_off; pb.c_off = (long long)k * T * T;
        pb.s_off = 0; pb.d_off = (long long)(k * q + k) * T;
        pb.M = T; pb.N = T; pb.K = r; pb.flags = GEMM_SYMMETRIC; pb.cap_rl = 0; pb.pad = 0;
        tb.blk.push_back(pb);
    }
    tb.cap_tiles = max_tiles(tb.cap); tb.yh_tiles = max_tiles(tb.yh); tb.blk_tiles = max_tiles(tb.blk);
    return tb;
}
}  // namespace

// Tile pairs and parts of the PautoSum product for a T x T output and `nslots` slots.  Pairs [0, n_uniform) are
// full tb x tb tiles (uniform kernel), the rest covers the remainder strip (generic kernel).  Returns tb.
static int syrk_plan(SyrkArgs &a, int T, int nslots, int q, int &n_uniform, int &grid_uniform, int &grid_generic, int &jb0_out) {
    const int nblk = (T + 7) / 8;
    int tb = 0;
    // 8-block (64 x 64) tiles: with more, smaller tiles a smaller share of the issued MMAs falls into the half-empty
    // diagonal tiles (T = 200: 384 issued blocks for 300 useful; 12-block tiles: 432), measured 6.5 vs 7.4 ms
    tb = nblk >= 8 ? 8 : 0;
    const int nt = tb ? nblk / tb : 0;
    int weights[SY_MAXPAIRS];
    a.npairs = 0;
    long long wsum = 0;
    auto add = [&](int r0, int nr, int c0, int nc, int diag, int w) {
        SyrkPair &p = a.pairs[a.npairs];
        p.r0 = r0; p.nr = nr; p.c0 = c0; p.nc = nc; p.diag = diag;
        weights[a.npairs++] = w;
        wsum += w;
    };
    for (int ti = 0; ti < nt; ti++)
        for (int tj = 0; tj <= ti; tj++)
            add(ti * tb * 8, tb * 8, tj * tb * 8, tb * 8, ti == tj, 1);     // equal weights: a diagonal CTA's busiest
                                                                           // pipes carry a full tile's load per slot
    n_uniform = a.npairs;
    const int s0 = nt * tb * 8;                               // first row of the strip
    const int budget = std::max(std::max(n_uniform, 1), (2 * 148) / std::max(q, 1));    // tile CTAs per latent: about two per SM over all latents
    long long off = 0;
    int part0 = 0;
    // equal parts for all tile pairs: a diagonal CTA's busiest pipes carry a full tile's load per slot, and the strip
    // partials written by the diagonal CTAs (below) need the same part count in every column tile
    const int np_tile = n_uniform ? std::max(1, std::min(budget / n_uniform, nslots)) : 0;
    for (int p = 0; p < n_uniform; p++) {
        a.pairs[p].nparts = np_tile;
        a.pairs[p].part0 = part0;
        a.pairs[p].out_off = off;
        part0 += np_tile;
        off += (long long)np_tile * a.pairs[p].nr * a.pairs[p].nc;
    }
    (void)weights; (void)wsum;
    grid_uniform = part0;
    grid_generic = 0;
    a.strip_pair = -1; a.strip_r0 = 0; a.strip_nr = 0;
    jb0_out = 0;
    if (s0 < T) {
        const int sr = T - s0, sbrows = (sr + 7) / 8;
        const int fused = (nt >= 1 && sr <= 8) ? 1 : 0;       // one block row: the diagonal CTAs' spare warps take strip x tiles
        if (fused) {
            SyrkPair &pm = a.pairs[a.npairs];
            pm.r0 = s0; pm.nr = sr; pm.c0 = 0; pm.nc = s0; pm.diag = 0; pm.nparts = np_tile; pm.part0 = 0; pm.out_off = off;
            off += (long long)np_tile * sr * s0;
            a.strip_pair = a.npairs; a.strip_r0 = s0; a.strip_nr = sr;
            a.npairs++;
        }
        // strip kernel: the whole strip (rows [s0, T) x columns [0, T)), or only its corner [s0, T) x [s0, T) when the
        // rest is fused into the tile kernel; one part per warp (8 warps per CTA)
        SyrkPair &p = a.pairs[a.npairs];
        p.r0 = s0; p.nr = sr; p.c0 = fused ? s0 : 0; p.nc = T - p.c0; p.diag = 0;
        jb0_out = p.c0 / 8;
        const int ngroups = ((T + 7) / 8 - jb0_out + 12) / 13;                 // column groups of the strip kernel (SYS_NB)
        int ctas = std::max(1, (2 * 148) / std::max(q * sbrows * ngroups, 1));
        ctas = std::max(1, std::min(ctas, fused ? nslots : (nslots + 7) / 8));
        p.nparts = fused ? ctas : ctas * 8;             // corner kernel: one part per CTA; strip kernel: one per warp
        p.part0 = 0;
        p.out_off = off;
        off += (long long)p.nparts * p.nr * p.nc;
        grid_generic = ctas;
        a.npairs++;
    }
    a.partial_per_latent = (off + 1) & ~1LL;                 // even: 16-byte stores into the areas of later latents
    return tb;
}

size_t pgpfa_i_pautosum_partial_bytes(int q, int T) {
    SyrkArgs a;
    int nu, gu, gg, jb;
    syrk_plan(a, T, 1 << 20, q, nu, gu, gg, jb);
    return align_up((size_t)q * a.partial_per_latent * 8);
}

template <int MI, int NJ, bool V16>
static int syrk_launch1(const SyrkArgs &a, int grid_x, int q, cudaStream_t st) {
    constexpr int ROWS = 32 * MI, COLS = 16 * NJ, SROWS = 8 * MI;
    constexpr int smem = SY_STAGES * (ROWS + COLS + SROWS) * SY_LD * 8;
    static bool attr_set = false;
    if (!attr_set) {
        PGPFA_CUDA_TRY(cudaFuncSetAttribute(syrk_sum_kernel<MI, NJ, V16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    dim3 grid(grid_x, q);
    syrk_sum_kernel<MI, NJ, V16><<<grid, 256, smem, st>>>(a);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
template <int MI, int NJ>
static int syrk_launch(const SyrkArgs &a, int grid_x, int q, cudaStream_t st) {
    if (grid_x <= 0) return PGPFA_OK;
    // 16-byte copies need every row of Y 16-byte aligned: r even (the slot stride q T r is then even too) and an aligned base
    const bool v16 = (a.r % 2 == 0) && (reinterpret_cast<size_t>(a.Y) % 16 == 0);
    return v16 ? syrk_launch1<MI, NJ, true>(a, grid_x, q, st) : syrk_launch1<MI, NJ, false>(a, grid_x, q, st);
}

#define LR_DISPATCH(FN, ...)                                   \
    switch (q) {                                               \
        LR_CASES(FN, __VA_ARGS__)                              \
        default: return PGPFA_ERR_ARG;                         \
    }
#define LR_CASE(QQ, FN, ...) case QQ: PGPFA_TRY(FN<QQ>(__VA_ARGS__)); break;
#define LR_CASES(FN, ...)                                                                                           \
    LR_CASE(1, FN, __VA_ARGS__) LR_CASE(2, FN, __VA_ARGS__) LR_CASE(3, FN, __VA_ARGS__) LR_CASE(4, FN, __VA_ARGS__)  \
    LR_CASE(5, FN, __VA_ARGS__) LR_CASE(6, FN, __VA_ARGS__) LR_CASE(7, FN, __VA_ARGS__) LR_CASE(8, FN, __VA_ARGS__)  \
    LR_CASE(9, FN, __VA_ARGS__) LR_CASE(10, FN, __VA_ARGS__) LR_CASE(11, FN, __VA_ARGS__) LR_CASE(12, FN, __VA_ARGS__)

// Uploads the problem tables of the posterior pass (once per E-step; they depend only on the ranks).
int pgpfa_i_lowrank_prepare(pgpfa_handle_s *h, const PgpfaLowRank &lr, int q, int T, void *probs_dev, cudaStream_t st) {
    const LrTables tb = lr_tables(lr, q, T);
    const size_t qq = (size_t)q * (q + 1);
    const size_t table_bytes = 5 * qq * sizeof(GemmProb);        // 4 problem tables + the (k, l) list of the 4th
    if (table_bytes > PGPFA_LOWRANK_TABLE_BYTES || table_bytes > PGPFA_STAGE_BYTES) return PGPFA_ERR_WORKSPACE;
    // staged through the handle's pinned buffer: the copy is asynchronous and needs no stream synchronisation (the
    // buffer is rewritten at the earliest by the next E-step, whose predecessor has long consumed it: every E-step
    // ends phase A with a wait for a count produced after this copy)
    PGPFA_CUDA_TRY(cudaEventSynchronize(h->ev_stage));      // previous use of the staging buffer (long complete)
    GemmProb *hs = reinterpret_cast<GemmProb *>(h->stage_h);
    memset(hs, 0, table_bytes);
    std::copy(tb.cap.begin(), tb.cap.end(), hs);
    std::copy(tb.yh.begin(), tb.yh.end(), hs + qq);
    std::copy(tb.blk.begin(), tb.blk.end(), hs + 2 * qq);
    std::copy(tb.capbig.begin(), tb.capbig.end(), hs + 3 * qq);
    std::copy(tb.capbig_kl.begin(), tb.capbig_kl.end(), reinterpret_cast<int2 *>(hs + 4 * qq));
    PGPFA_CUDA_TRY(cudaMemcpyAsync(probs_dev, hs, table_bytes, cudaMemcpyHostToDevice, st));
    PGPFA_CUDA_TRY(cudaEventRecord(h->ev_stage, st));
    return PGPFA_OK;
}

// Posterior pass of one chunk of trials through the low-rank prior factor.  `area` is scratch of at least
// nslots * pgpfa_i_lowrank_bytes_per_slot(q, T, r) bytes; `probs_dev` was filled by pgpfa_i_lowrank_prepare.
// Order: per-bin matrices, capacitance matrix, its factorisation and triangular inverse, Y, polishing Newton step,
// time-diagonal blocks (event `ev_means` is recorded here: x / vsm final), then the T x T blocks of every latent.
int pgpfa_i_lowrank_posterior(pgpfa_handle_s *h, const PgpfaLowRank &lr, const double *W, const double *gvec, double *x,
                              double *dx, const int *act, int nslots, int q, int T, double tol, double *steplen,
                              double *vsm, double *vsmGP, void *area, size_t area_bytes, void *probs_dev, cudaStream_t st,
                              int *info, double *pautosum, int pauto_accumulate, const double *post_mean,
                              double *pauto_partial) {
    if (nslots <= 0) return PGPFA_OK;
    const int r = lr.r, n = q * T, nbr = pgpfa_nb(r);
    const long long ltr = pgpfa_ltiles(nbr);
    unsigned char *p = static_cast<unsigned char *>(area);
    auto take = [&](size_t bytes) { unsigned char *o = p; p += align_up(bytes); return o; };
    double *Lr = (double *)take((size_t)nslots * ltr * PGPFA_TILE * 8);
    double *Dr = (double *)take((size_t)nslots * nbr * PGPFA_TILE * 8);
    double *Zr = (double *)take((size_t)nslots * ltr * PGPFA_TILE * 8);
    double *G = (double *)take((size_t)nslots * r * r * 8);
    double *Zd = (double *)take((size_t)nslots * r * r * 8);
    double *Y = (double *)take((size_t)nslots * n * r * 8);
    double *Pm = (double *)take((size_t)nslots * q * q * T * 8);
    double *Dm = (double *)take((size_t)nslots * q * q * T * 8);
    double *u = (double *)take((size_t)nslots * r * 8);
    const GemmProb *dprobs = static_cast<const GemmProb *>(probs_dev);
    const size_t qq = (size_t)q * (q + 1);
    const LrTables tb = lr_tables(lr, q, T);
    // The capacitance blocks are ONE product per latent pair over all slots (rows = entries (a, b) of the block,
    // columns = slots, K = T) against a feature matrix Phi built per call in the slack of the area: full 64-row tiles
    // instead of ragged r_k x r_l ones.  The choice depends only on the room in the area, never on the slot count, so
    // chunked and unchunked solves stay bit-identical (PGPFA_CAP_BIG_MIN is a development switch).
    double *Phi = (double *)take((size_t)tb.phi_rows * T * 8);
    static const int cap_big_min = [] { const char *e = getenv("PGPFA_CAP_BIG_MIN"); return e ? atoi(e) : 1; }();
    const bool cap_big = nslots >= cap_big_min && (size_t)(p - static_cast<unsigned char *>(area)) <= area_bytes &&
                         !tb.capbig.empty();

    // ---- per-bin P, Dt; capacitance matrix G = I + F^T Dt F (lower block triangle, ragged blocks r_k x r_l)
    pgpfa_prof_begin(h, PGPFA_PROF_LOWRANK, st);
    LR_DISPATCH(lr_launch_bins, W, act, T, lr.eps, Pm, Dm, nslots, st)
    if (cap_big) {
        const GemmProb *pb = dprobs + 3 * qq;
        const int2 *kl = reinterpret_cast<const int2 *>(dprobs + 4 * qq);
        dim3 gf(32, (unsigned)tb.capbig.size());
        cap_features_kernel<<<gf, 256, 0, st>>>(lr.Ft, pb, kl, T, Phi);
        PGPFA_LAUNCH_CHECK();
        GemmArgs g;
        g.A = Phi; g.B = Dm; g.C = G; g.scale = nullptr; g.dadd = nullptr;
        g.strideA = 0; g.strideB = 0; g.strideC = (long long)r * r; g.strideS = 0; g.strideD = 0;
        g.lda = T; g.ldb = q * q * T; g.ldc = r; g.cmap = nullptr; g.probs = nullptr; g.dadd_alpha = 0.0;
        g.n_override = nslots;
        PGPFA_TRY(launch_gemm(g, pb, (int)tb.capbig.size(), tb.capbig_mtiles * ((nslots + 63) / 64), 1, st));
    } else {
        GemmArgs g;
        g.A = lr.Ft; g.B = lr.Ft; g.C = G; g.scale = Dm; g.dadd = nullptr;
        g.strideA = 0; g.strideB = 0; g.strideC = (long long)r * r; g.strideS = (long long)q * q * T; g.strideD = 0;
        g.lda = T; g.ldb = T; g.ldc = r; g.cmap = nullptr; g.probs = nullptr; g.dadd_alpha = 0.0;
        PGPFA_TRY(launch_gemm(g, dprobs, (int)tb.cap.size(), tb.cap_tiles, nslots, st));
    }
    pgpfa_prof_end(h, st);
    // ---- L_b L_b^T = G, Z_b = L_b^-1 (tile kernels of factor.cu on r x r matrices)
    PgpfaMatSrc ms;
    ms.Kinv = nullptr; ms.W = nullptr; ms.dense = G; ms.q = 1; ms.T = r; ms.n = r; ms.diag_scale = 1.0;
    pgpfa_prof_begin(h, PGPFA_PROF_FACTOR, st);
    // info is trial-indexed through `act`; the dense source G is slot-indexed (factor.cu mat_elem), so passing the
    // active list only routes a non-positive pivot of slot s to info[act[s]]
    PGPFA_TRY(pgpfa_i_factor(ms, Lr, Dr, Zr, info ? act : nullptr, info, nslots, st, h));
    pgpfa_prof_end(h, st);
    h->prof_work[PGPFA_PROF_FACTOR] += (double)nslots * r * (double)r * r / 3.0;
    pgpfa_prof_begin(h, PGPFA_PROF_TRTRI, st);
    PGPFA_TRY(pgpfa_i_trtri(Lr, Dr, Zr, r, nslots, st, h));
    pgpfa_prof_end(h, st);
    h->prof_work[PGPFA_PROF_TRTRI] += (double)nslots * r * (double)r * r / 3.0;
    pgpfa_prof_begin(h, PGPFA_PROF_LOWRANK, st);
    {
        dim3 grid(64, nslots);
        zt_to_dense_lower_kernel<<<grid, 256, 0, st>>>(Zr, nbr, r, Zd);
        PGPFA_LAUNCH_CHECK();
    }
    // ---- Yh[(l,t), c] = sum_a F_l[t][a] Z_b[c][off_l + a], then Y = P Yh per bin
    {
        GemmArgs g;
        g.A = lr.F; g.B = Zd; g.C = Y; g.scale = nullptr; g.dadd = nullptr;
        g.strideA = 0; g.strideB = (long long)r * r; g.strideC = (long long)n * r; g.strideS = 0; g.strideD = 0;
        g.lda = T; g.ldb = r; g.ldc = r; g.cmap = nullptr; g.probs = nullptr; g.dadd_alpha = 0.0;
        PGPFA_TRY(launch_gemm(g, dprobs + qq, (int)tb.yh.size(), tb.yh_tiles, nslots, st));
    }
    // ---- polishing Newton step  dx = -Sigma g = -(eps P g + Y (Y^T g)); Y^T g is taken inside the mixing pass (its
    // per-group partial sums live in G, which is dead once the capacitance matrix is factored)
    const int ngroups = (T + LR_TB - 1) / LR_TB;
    LrOffs offs;
    for (int l = 0; l < PGPFA_QMAX; l++) offs.off[l] = l < q ? lr.off[l] : 0;
    if (ngroups <= r) {
        LR_DISPATCH(lr_launch_mix, Y, Pm, T, r, nslots, st, gvec, act, G, offs)
        dim3 grid((r + 127) / 128, nslots);
        lr_usum_kernel<<<grid, 128, 0, st>>>(G, ngroups, r, u);
        PGPFA_LAUNCH_CHECK();
    } else {
        LR_DISPATCH(lr_launch_mix, Y, Pm, T, r, nslots, st, nullptr, nullptr, nullptr, offs)
        dim3 grid((r + 127) / 128, nslots);
        lr_ytg_kernel<<<grid, 128, 0, st>>>(Y, gvec, act, n, r, u);
        PGPFA_LAUNCH_CHECK();
    }
    // ---- the step and the per-bin q x q slices come out of ONE sweep over Y
    if (vsm) {
        LR_DISPATCH(lr_launch_vsm, Y, Pm, act, T, r, lr.eps, vsm, nslots, st, u, gvec, dx)
    } else {
        LR_DISPATCH(lr_launch_step, Y, u, Pm, gvec, act, T, r, lr.eps, dx, nslots, st)
    }
    PGPFA_TRY(pgpfa_i_polish(x, dx, act, n, 1e3 * tol, steplen, nslots, st));
    PGPFA_CUDA_TRY(cudaEventRecord(h->ev_means, st));      // pgpfa_stream_wait_means
    pgpfa_prof_end(h, st);
    if (vsmGP) {
        pgpfa_prof_begin(h, PGPFA_PROF_SLICES, st);
        GemmArgs g;
        g.A = Y; g.B = Y; g.C = vsmGP; g.scale = nullptr; g.dadd = Pm;
        g.strideA = (long long)n * r; g.strideB = g.strideA; g.strideC = (long long)q * T * T; g.strideS = 0;
        g.strideD = (long long)q * q * T;
        g.lda = r; g.ldb = r; g.ldc = T; g.cmap = act; g.probs = nullptr; g.dadd_alpha = lr.eps;
        PGPFA_TRY(launch_gemm(g, dprobs + 2 * qq, (int)tb.blk.size(), tb.blk_tiles, nslots, st));
        pgpfa_prof_end(h, st);
        // algorithmic work of the q symmetric T x T x r products: T (T+1) r flops each
        h->prof_work[PGPFA_PROF_SLICES] += (double)nslots * q * (double)T * (T + 1) * r;
    }
    if (pautosum) {
        // the trial-sum of the same products without the per-trial blocks ever reaching HBM
        if (T > 575 || !pauto_partial || !post_mean) return PGPFA_ERR_ARG;     // <= SY_MAXPAIRS tile pairs (8 x 9 / 2 + strip)
        pgpfa_prof_begin(h, PGPFA_PROF_SLICES, st);
        SyrkArgs a;
        a.Y = Y; a.partial = pauto_partial; a.strideY = (long long)n * r; a.r = r; a.T = T; a.q = q; a.nslots = nslots;
        int n_uniform = 0, grid_uniform = 0, grid_generic = 0, jb0 = 0;
        const int tb = syrk_plan(a, T, nslots, q, n_uniform, grid_uniform, grid_generic, jb0);
        a.first = 0; a.count = n_uniform;          // the tile kernel reads pairs [0, count)
        { const char *e = getenv("PGPFA_SYRK_DBG"); a.dbg = e ? atoi(e) : 0; }
        if (tb == 8) PGPFA_TRY((syrk_launch<2, 4>(a, grid_uniform, q, st)));
        if (grid_generic > 0) {
            const SyrkPair &sp = a.pairs[a.npairs - 1];
            if (a.strip_nr > 0) {                       // fused strip: only its 8 x 8 corner is left
                dim3 gc(grid_generic, q);
                syrk_corner_kernel<<<gc, 256, 0, st>>>(a, a.npairs - 1);
            } else {
                const int ngroups = ((T + 7) / 8 - jb0 + SYS_NB - 1) / SYS_NB;
                dim3 gs(grid_generic, q, ((sp.nr + 7) / 8) * ngroups);
                syrk_strip_kernel<<<gs, 256, 0, st>>>(a, a.npairs - 1, sp.nparts, ngroups, jb0);
            }
            PGPFA_LAUNCH_CHECK();
        }
        double *dsum = Dm;                                   // (q, T) doubles; Dm (slots, q*q, T) is dead after the capacitance GEMM
        dim3 gpd(T, q);
        syrk_pdiag_kernel<<<gpd, 128, 0, st>>>(Pm, nslots, q, T, dsum);
        PGPFA_LAUNCH_CHECK();
        dim3 gfin((T + 15) / 16, (T + 15) / 16, q);
        syrk_finish_kernel<<<gfin, 256, 0, st>>>(a, dsum, post_mean, act, lr.eps, pauto_accumulate, pautosum);
        PGPFA_LAUNCH_CHECK();
        pgpfa_prof_end(h, st);
        h->prof_work[PGPFA_PROF_SLICES] += (double)nslots * q * (double)T * (T + 1) * r;
    }
    return PGPFA_OK;
}
