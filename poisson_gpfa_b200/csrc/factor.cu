// Batched blocked FP64 Cholesky / triangular inverse / selected inverse of the qT x qT posterior
// systems on the FP64 tensor pipe (DMMA.8x8x4), tiles staged through shared memory by bulk async
// copies (UBLKCP).  Replaces the dense LAPACK calls of the reference:
//   np.linalg.inv(hess)            funs/inference.py:130-131   (posterior covariance)
//   Newton-CG's inner solves        funs/inference.py:119-126   (Newton step H d = -g)
//   np.linalg.inv / slogdet         funs/inference.py:82,190,207; funs/learning.py:191-192
// Algorithm (left-looking by block column j, 64x64 tiles, packed lower storage):
//   diag kernel : S = A(j,j) - sum_k L(j,k) L(j,k)^T ; L(j,j) = chol(S) ; Dinv_j = L(j,j)^-1
//   panel kernel: L(i,j) = (A(i,j) - sum_k L(i,k) L(j,k)^T) Dinv_j^T         for all i > j
// Triangular inverse ZT = L^-T (packed upper, by block row i of L):
//   ZT(j,i) = -(sum_{k=j}^{i-1} ZT(j,k) L(i,k)^T) Dinv_i^T                   for all j < i
// Selected inverse Sigma = ZT ZT^T: only the tiles that are consumed (diagonal T x T blocks for
// post_vsmGP, or everything for post_cov / K^-1); the time-diagonals (post_vsm) come from row Gram
// products of ZT.
#include "common.cuh"
#include "pgpfa_internal.h"

using namespace pgpfa;

namespace {

// ---------------------------------------------------------------------------------------------
// Matrix element source: either generated on the fly  H = blkdiag(Kinv_k) + scatter(W)  (never
// materialised) or read from a dense row-major batch.
// ---------------------------------------------------------------------------------------------
struct MatSource {
    const double *Kinv;   // (q,T,T) or nullptr
    const double *W;      // (trials, q*q, T)
    const double *dense;  // (slots, n, n) when Kinv == nullptr
    int q, T, n;
    double diag_scale;    // 1 (Laplace) or 1+1e-6 (variational, funs/inference.py:190)
};

struct FragIdx {
    int row[4], col[4];        // global matrix rows / first col of each pair
    int rk[4], rs[4];          // latent / bin of each row      (-1 latent => padding)
    int cl[4][2], ct[4][2];    // latent / bin of each column
};

__device__ __forceinline__ void frag_index(FragIdx &f, const MatSource &src, int ti, int tj, int wm, int wn, int lane) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = ti * PGPFA_NB + frag_row(wm, i, lane);
        f.row[i] = r;
        if (r < src.n) { f.rk[i] = r / src.T; f.rs[i] = r - f.rk[i] * src.T; } else { f.rk[i] = -1; f.rs[i] = 0; }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c0 = tj * PGPFA_NB + frag_col(wn, j, lane);
        f.col[j] = c0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int c = c0 + e;
            if (c < src.n) { f.cl[j][e] = c / src.T; f.ct[j][e] = c - f.cl[j][e] * src.T; } else { f.cl[j][e] = -1; f.ct[j][e] = 0; }
        }
    }
}

__device__ __forceinline__ double mat_elem(const MatSource &src, int trial, int slot, const FragIdx &f, int i, int j, int e) {
    const int r = f.row[i], c = f.col[j] + e;
    if (f.rk[i] < 0 || f.cl[j][e] < 0) return (r == c) ? 1.0 : 0.0;
    double v;
    if (src.Kinv == nullptr) {
        v = src.dense[((size_t)slot * src.n + r) * src.n + c];
    } else {
        v = 0.0;
        const int s = f.rs[i], t = f.ct[j][e];
        const int k = f.rk[i], l = f.cl[j][e];
        if (k == l) v = __ldg(&src.Kinv[((size_t)k * src.T + s) * src.T + t]);
        if (s == t) v += src.W[((size_t)trial * src.q * src.q + k * src.q + l) * src.T + t];
    }
    if (r == c) v *= src.diag_scale;
    return v;
}

struct FactorArgs {
    MatSource src;
    double *L;          // slots x ltiles x 4096
    double *Dinv;       // slots x nb x 4096
    double *ZT;         // slots x ltiles x 4096 (packed upper) or nullptr
    float *L32, *D32;   // optional FP32 mirrors of L / Dinv (inexact chord sweeps stream half the bytes)
    const int *act;     // slot -> trial, or nullptr (trial = slot)
    int *info;          // per trial: 0 ok, >0 first non-positive pivot (1-based)
    int nb;
    int step;
    int mode;           // panel kernel: 0 = Cholesky panel, 1 = triangular-inverse row
    int nslots, ntiles; // panel kernel, mode 0: 1-D grid decode (ntiles tile rows per slot)
    int fuse;           // panel kernel, mode 0: first-tile CTAs also factor diagonal tile step+1
};

__device__ __forceinline__ void zero_acc(double (&acc)[4][4][2]) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// ---------------------------------------------------------------------------------------------
// Diagonal-tile kernel: SYRK update + 64x64 Cholesky + 64x64 triangular inverse (in shared memory)
// ---------------------------------------------------------------------------------------------
// row index of the bi-th block in a row-major enumeration of a lower-triangular block grid (bi = rb(rb+1)/2 + cb)
__constant__ unsigned char c_tri_row[28] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6};

#define SLD 65
// One warp factors the 8x8 diagonal sub-block b of S in place (lower triangle) and writes its inverse to
// DI[b]: lane i (mod 8) owns row i in registers, pivots travel by shuffle.
__device__ __forceinline__ void diag8_factor(double *S, double *DI, int *bad_sm, int b, int jtile, int lane) {
    const int i = lane & 7;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[c] = (lane < 8 && c <= i) ? S[(8 * b + i) * SLD + 8 * b + c] : 0.0;   // lanes 8-31 only relay shuffles
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double piv = __shfl_sync(0xffffffffu, a[k], k);
        if (!(piv > 0.0)) { if (lane == 0 && *bad_sm == 0) *bad_sm = jtile * PGPFA_NB + 8 * b + k + 1; piv = 1.0; }
        const double dd = sqrt(piv), dinv = 1.0 / dd;
        a[k] = (i == k) ? dd : ((i > k) ? a[k] * dinv : a[k]);
#pragma unroll
        for (int c = k + 1; c < 8; c++) {
            const double lck = __shfl_sync(0xffffffffu, a[k], c);
            if (i >= c) a[c] -= a[k] * lck;
        }
    }
    // inverse of the sub-block: lane i owns column i of it (forward substitution)
    double xv[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const double lrr = __shfl_sync(0xffffffffu, a[r], r);
        double sacc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k < r) { const double lrk = __shfl_sync(0xffffffffu, a[k], r); sacc += lrk * xv[k]; }
        xv[r] = (r < i) ? 0.0 : ((r == i) ? 1.0 / lrr : -sacc / lrr);
    }
    if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; c++) if (c <= i) S[(8 * b + i) * SLD + 8 * b + c] = a[c];
#pragma unroll
        for (int r = 0; r < 8; r++) DI[b * 64 + r * 8 + i] = xv[r];
    }
}
// acc holds sum_k L(j,k) L(j,k)^T for the diagonal tile j of this slot: form S = A(j,j) - acc in shared
// memory, factor it, invert the factor, and write L(j,j), Dinv_j and (optionally) ZT(j,j).
// Must be entered by all threads with every earlier use of the shared staging area finished (__syncthreads).
__device__ __forceinline__ void diag_epilogue(const FactorArgs &a, unsigned char *smem_raw, double (&acc)[4][4][2],
                                              int trial, int slot, int j, double *Ls) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    const long long ltl = (long long)a.nb * (a.nb + 1) / 2;
    double *S = reinterpret_cast<double *>(smem_raw);      // [64][65]
    double *X = S + PGPFA_NB * SLD;                        // [64][65]
    {
        FragIdx f;
        frag_index(f, a.src, j, j, wm, wn, lane);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int jj = 0; jj < 4; jj++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int r = frag_row(wm, i, lane), c = frag_col(wn, jj, lane) + e;
                    if (c <= r) S[r * SLD + c] = mat_elem(a.src, trial, slot, f, i, jj, e) - acc[i][jj][e];
                }
    }
    // blocked right-looking Cholesky (8x8 sub-blocks) + blocked triangular inverse, all in shared memory
    double *DI = X + PGPFA_NB * SLD;          // [8][8][8] inverses of the diagonal sub-blocks
    double *TT = DI + 8 * 64;                 // [7][8][8] scratch for the inverse
    int *bad_sm = reinterpret_cast<int *>(TT + 7 * 64);
    if (tid == 0) *bad_sm = 0;
    // Software pipeline over the 8 sub-block columns: while warp 0 factors and inverts diagonal sub-block b+1
    // (a serial chain of shuffles, square roots and divisions), warps 1-3 apply the rank-8 trailing update of
    // step b to the rest of the tile.
    __syncthreads();
    if (warp == 0) diag8_factor(S, DI, bad_sm, 0, j, lane);
    for (int b = 0; b < 8; b++) {
        __syncthreads();                               // DI_b and the factored sub-block b are visible
        // panel below the sub-block: P = S_panel * Dinv_b^T (one thread per row)
        {
            const int r = 8 * (b + 1) + tid;
            if (r < PGPFA_NB) {
                double v[8], o[8];
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = S[r * SLD + 8 * b + k];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double acc2 = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; k++) if (k <= c) acc2 += v[k] * DI[b * 64 + c * 8 + k];
                    o[c] = acc2;
                }
#pragma unroll
                for (int c = 0; c < 8; c++) S[r * SLD + 8 * b + c] = o[c];
            }
        }
        __syncthreads();
        if (b == 7) break;
        const int m = 7 - b;                           // remaining block rows / columns
        if (warp == 0) {
            // next diagonal sub-block first (64 elements, 2 per lane), then its factorisation
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
                const int e = lane + 32 * h2, er = e >> 3, ec = e & 7;
                const int r = 8 * (b + 1) + er, c2 = 8 * (b + 1) + ec;
                double sum = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += S[r * SLD + 8 * b + k] * S[c2 * SLD + 8 * b + k];
                if (c2 <= r) S[r * SLD + c2] -= sum;
            }
            __syncwarp();
            diag8_factor(S, DI, bad_sm, b + 1, j, lane);
        } else {
            // remaining 8x8 blocks (rb >= cb >= 1, not (1,1)) of the m x m trailing grid, 3 warps, 2 elements per lane
            const int nblk = m * (m + 1) / 2;
            for (int bi = 1 + (warp - 1); bi < nblk; bi += 3) {
                const int rb = c_tri_row[bi], cb = bi - rb * (rb + 1) / 2;
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    const int e = lane + 32 * h2, er = e >> 3, ec = e & 7;
                    const int r = 8 * (b + 1 + rb) + er, c2 = 8 * (b + 1 + cb) + ec;
                    double sum = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; k++) sum += S[r * SLD + 8 * b + k] * S[c2 * SLD + 8 * b + k];
                    if (c2 <= r) S[r * SLD + c2] -= sum;
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0 && *bad_sm && a.info) atomicCAS(&a.info[trial], 0, *bad_sm);
    // X = L^-1 by block rows: X(b,b) = DI_b ; X(b,j) = -DI_b * sum_{k=j}^{b-1} L(b,k) X(k,j)
    for (int e = tid; e < PGPFA_NB * PGPFA_NB; e += PGPFA_GEMM_THREADS) {
        const int r = e >> 6, c = e & 63;
        X[r * SLD + c] = ((r >> 3) == (c >> 3)) ? DI[(r >> 3) * 64 + (r & 7) * 8 + (c & 7)] : 0.0;
    }
    for (int b = 1; b < 8; b++) {
        __syncthreads();
        for (int e = tid; e < 64 * b; e += PGPFA_GEMM_THREADS) {
            const int jb = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
            double sum = 0.0;
            for (int k = 8 * jb; k < 8 * b; k++) sum += S[(8 * b + rr) * SLD + k] * X[k * SLD + 8 * jb + cc];
            TT[e] = sum;
        }
        __syncthreads();
        for (int e = tid; e < 64 * b; e += PGPFA_GEMM_THREADS) {
            const int jb = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
            double sum = 0.0;
#pragma unroll
            for (int mm = 0; mm < 8; mm++) if (mm <= rr) sum += DI[b * 64 + rr * 8 + mm] * TT[jb * 64 + mm * 8 + cc];
            X[(8 * b + rr) * SLD + 8 * jb + cc] = -sum;
        }
    }
    __syncthreads();
    double *Lt = Ls + ltile(j, j) * PGPFA_TILE;
    double *Dt = a.Dinv + ((size_t)slot * a.nb + j) * PGPFA_TILE;
    double *Zt = a.ZT ? a.ZT + ((size_t)slot * ltl + utile(j, j, a.nb)) * PGPFA_TILE : nullptr;
    float *Lt32 = a.L32 ? a.L32 + ((size_t)slot * ltl + ltile(j, j)) * PGPFA_TILE : nullptr;
    float *Dt32 = a.D32 ? a.D32 + ((size_t)slot * a.nb + j) * PGPFA_TILE : nullptr;
    for (int off = tid; off < PGPFA_TILE; off += PGPFA_GEMM_THREADS) {
        const int r = (((off >> 6) & 7) << 3) + ((off >> 3) & 7);
        const int c = ((off >> 11) << 5) + (((off >> 9) & 3) << 3) + (off & 7);
        const double lv = (c <= r) ? S[r * SLD + c] : 0.0, dv = (c <= r) ? X[r * SLD + c] : 0.0;
        Lt[off] = lv;
        Dt[off] = dv;
        if (Lt32) Lt32[off] = (float)lv;
        if (Dt32) Dt32[off] = (float)dv;
        if (Zt) Zt[off] = (r <= c) ? X[c * SLD + r] : 0.0;
    }
}

__global__ void __launch_bounds__(PGPFA_GEMM_THREADS, PGPFA_GEMM_CTAS_PER_SM) chol_diag_kernel(FactorArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lpos = blockIdx.x;
    const int trial = a.act ? a.act[lpos] : lpos;
    const int slot = lpos;
    const int j = a.step;
    const long long ltl = (long long)a.nb * (a.nb + 1) / 2;
    double *Ls = a.L + (size_t)slot * ltl * PGPFA_TILE;
    GemmPipe pipe;
    pipe_setup(pipe, smem_raw);
    double acc[4][4][2];
    zero_acc(acc);
    const double *Arow = Ls + ltile(j, 0) * PGPFA_TILE;
    gemm_slabs<true>(acc, pipe, Arow, Arow, 2 * j);
    diag_epilogue(a, smem_raw, acc, trial, slot, j, Ls);
}

// ---------------------------------------------------------------------------------------------
// Panel kernel: out = (init - sum_k A_k B_k^T) * D^T   (Cholesky panel tile or inverse-row tile)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PGPFA_GEMM_THREADS, PGPFA_GEMM_CTAS_PER_SM) chol_panel_kernel(FactorArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // mode 0 (Cholesky panel of step j): 1-D slot-major grid (a slot's CTAs are neighbours and share the L(j,:)
    // operand in L2).  The CTA of tile row j+1 goes on, when a.fuse is set, to factor diagonal tile j+1 (which
    // only waits for this very tile); these long-running CTAs are spread evenly through the launch so that the
    // second resident CTA of each SM keeps the tensor pipe busy meanwhile.
    // mode 1 (triangular inverse, row i): 2-D grid (tile, slot).
    int slot, tile;
    if (a.mode == 0) {
        const int b = blockIdx.x;
        slot = b / a.ntiles;
        tile = b - slot * a.ntiles;
    } else {
        slot = blockIdx.y;
        tile = blockIdx.x;
    }
    const int trial = a.act ? a.act[slot] : slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    const long long ltl = (long long)a.nb * (a.nb + 1) / 2;
    double *Ls = a.L + (size_t)slot * ltl * PGPFA_TILE;
    double *Zs = a.ZT ? a.ZT + (size_t)slot * ltl * PGPFA_TILE : nullptr;

    const double *A, *B, *D;
    double *out;
    int nslab, ti = 0, tj = 0;
    if (a.mode == 0) {
        const int j = a.step, i = j + 1 + tile;
        A = Ls + ltile(i, 0) * PGPFA_TILE;
        B = Ls + ltile(j, 0) * PGPFA_TILE;
        nslab = 2 * j;
        D = a.Dinv + ((size_t)slot * a.nb + j) * PGPFA_TILE;
        out = Ls + ltile(i, j) * PGPFA_TILE;
        ti = i; tj = j;
    } else {
        const int i = a.step, j = tile;
        A = Zs + utile(j, j, a.nb) * PGPFA_TILE;
        B = Ls + ltile(i, j) * PGPFA_TILE;
        nslab = 2 * (i - j);
        D = a.Dinv + ((size_t)slot * a.nb + i) * PGPFA_TILE;
        out = Zs + utile(j, i, a.nb) * PGPFA_TILE;
    }
    GemmPipe pipe;
    pipe_setup(pipe, smem_raw);
    uint64_t *aux = pipe.full + PGPFA_STAGES;
    if (tid == 0) { mbar_init(aux, 1); mbar_fence_init(); }
    double acc[4][4][2];
    zero_acc(acc);
    gemm_slabs<false>(acc, pipe, A, B, nslab);
    __syncthreads();
    double *St = pipe.stages;                 // 32 KB: (init - acc) in operand layout
    double *Dt = pipe.stages + PGPFA_TILE;    // 32 KB: Dinv tile
    if (tid == 0) {
        mbar_expect_tx(aux, PGPFA_TILE * 8);
        bulk_g2s(Dt, D, PGPFA_TILE * 8, aux);
    }
    {
        double2 *S2 = reinterpret_cast<double2 *>(St);
        if (a.mode == 0) {
            FragIdx f;
            frag_index(f, a.src, ti, tj, wm, wn, lane);
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double2 v;
                    v.x = mat_elem(a.src, trial, slot, f, i, jj, 0) - acc[i][jj][0];
                    v.y = mat_elem(a.src, trial, slot, f, i, jj, 1) - acc[i][jj][1];
                    S2[frag_slot2(wm, wn, i, jj, lane)] = v;
                }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    double2 v;
                    v.x = -acc[i][jj][0];
                    v.y = -acc[i][jj][1];
                    S2[frag_slot2(wm, wn, i, jj, lane)] = v;
                }
        }
    }
    __syncthreads();
    mbar_wait(aux, 0);
    zero_acc(acc);
    slab_mma(acc, St, Dt, wm, wn, lane);
    if (wn == 1) slab_mma(acc, St + PGPFA_SLAB, Dt + PGPFA_SLAB, wm, wn, lane);   // D lower-triangular: k<=n
    double2 *O2 = reinterpret_cast<double2 *>(out);
    float2 *F2 = (a.mode == 0 && a.L32)
                     ? reinterpret_cast<float2 *>(a.L32 + ((size_t)slot * ltl + ltile(ti, tj)) * PGPFA_TILE) : nullptr;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            double2 v;
            v.x = acc[i][jj][0];
            v.y = acc[i][jj][1];
            O2[frag_slot2(wm, wn, i, jj, lane)] = v;
            if (F2) F2[frag_slot2(wm, wn, i, jj, lane)] = make_float2((float)v.x, (float)v.y);
        }
    if (a.mode != 0 || !a.fuse || tile != 0) return;
    // ---- fused look-ahead: this CTA just produced L(j+1, j), the last tile diagonal step j+1 was waiting for.
    // Keep it in shared memory (B halves of stages 0/1, untouched by the single-operand pipeline), run the SYRK
    // over the earlier tiles of block row j+1 from global memory, add the k = j term from shared memory, and
    // factor / invert diagonal tile j+1 while the other CTAs of this launch finish panel j.
    const int jn = a.step + 1;
    __syncthreads();                                   // everyone done reading St / Dt
    {
        double2 *N0 = reinterpret_cast<double2 *>(pipe.stages + PGPFA_SLAB);                // slab 0 of L(j+1,j)
        double2 *N1 = reinterpret_cast<double2 *>(pipe.stages + PGPFA_TILE + PGPFA_SLAB);   // slab 1
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                double2 v;
                v.x = acc[i][jj][0];
                v.y = acc[i][jj][1];
                // frag_slot2 = wn*1024 + jj*256 + (wm*4+i)*32 + lane (double2 units); wn selects the slab
                (wn == 0 ? N0 : N1)[jj * 256 + (wm * 4 + i) * 32 + lane] = v;
            }
    }
    __syncthreads();
    zero_acc(acc);
    const double *Arow = Ls + ltile(jn, 0) * PGPFA_TILE;
    gemm_slabs<true>(acc, pipe, Arow, Arow, 2 * a.step);
    slab_mma(acc, pipe.stages + PGPFA_SLAB, pipe.stages + PGPFA_SLAB, wm, wn, lane);
    slab_mma(acc, pipe.stages + PGPFA_TILE + PGPFA_SLAB, pipe.stages + PGPFA_TILE + PGPFA_SLAB, wm, wn, lane);
    __syncthreads();
    diag_epilogue(a, smem_raw, acc, trial, slot, jn, Ls);
}

// ---------------------------------------------------------------------------------------------
// Selected inverse tiles: Sigma(a,b) = sum_{m>=a} ZT(a,m) ZT(b,m)^T, scattered to the consumer layout
// ---------------------------------------------------------------------------------------------
struct LauumArgs {
    const double *ZT;
    const int2 *pairs;       // (a, b), a >= b
    const int *act;          // slot -> trial (output index), or nullptr
    double *vsmGP;           // (trials, q, T, T) or nullptr
    double *dense;           // (slots, n, n) or nullptr
    int nb, n, q, T;
};

__global__ void __launch_bounds__(PGPFA_GEMM_THREADS, PGPFA_GEMM_CTAS_PER_SM) lauum_tiles_kernel(LauumArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int slot = blockIdx.y;
    const int trial = a.act ? a.act[slot] : slot;
    const int2 pr = a.pairs[blockIdx.x];
    const int ta = pr.x, tb = pr.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    const long long ltl = (long long)a.nb * (a.nb + 1) / 2;
    const double *Zs = a.ZT + (size_t)slot * ltl * PGPFA_TILE;
    GemmPipe pipe;
    pipe_setup(pipe, smem_raw);
    double acc[4][4][2];
    zero_acc(acc);
    const double *A = Zs + utile(ta, ta, a.nb) * PGPFA_TILE;
    const double *B = Zs + utile(tb, ta, a.nb) * PGPFA_TILE;
    if (ta == tb) gemm_slabs<true>(acc, pipe, A, A, 2 * (a.nb - ta));
    else gemm_slabs<false>(acc, pipe, A, B, 2 * (a.nb - ta));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = ta * PGPFA_NB + frag_row(wm, i, lane);
        if (r >= a.n) continue;
        const int rk = r / a.T, rs = r - rk * a.T;
#pragma unroll
        for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int c = tb * PGPFA_NB + frag_col(wn, jj, lane) + e;
                if (c >= a.n) continue;
                const double v = acc[i][jj][e];
                if (a.dense) {
                    double *dd = a.dense + (size_t)slot * a.n * a.n;
                    dd[(size_t)r * a.n + c] = v;
                    dd[(size_t)c * a.n + r] = v;
                }
                if (a.vsmGP) {
                    const int ck = c / a.T;
                    if (ck == rk) {
                        const int cs = c - ck * a.T;
                        double *g = a.vsmGP + ((size_t)trial * a.q + rk) * a.T * a.T;
                        g[(size_t)rs * a.T + cs] = v;
                        g[(size_t)cs * a.T + rs] = v;
                    }
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Time-diagonal blocks: vsm[t][k][l] = Sigma[kT+t, lT+t] = <row kT+t of ZT, row lT+t of ZT>
// one warp per bin t; each lane streams 16-byte pieces of the q rows tile by tile
// ---------------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(128) cov_timediag_kernel(const double *__restrict__ ZT, const int *act,
                                                           double *__restrict__ vsm, int nb, int T) {
    const int slot = blockIdx.y;
    const int trial = act ? act[slot] : slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * 4 + warp;
    if (t >= T) return;
    const long long ltl = (long long)nb * (nb + 1) / 2;
    const double *Zs = ZT + (size_t)slot * ltl * PGPFA_TILE;
    int arow[Q], roff[Q];
#pragma unroll
    for (int k = 0; k < Q; k++) {
        const int rho = k * T + t;
        arow[k] = rho >> 6;
        const int r = rho & 63;
        roff[k] = ((r >> 3) << 6) + ((r & 7) << 3);
    }
    const int seg = lane >> 2;
    const int loff = ((seg >> 2) << 11) + ((seg & 3) << 9) + 2 * (lane & 3);
    double acc[Q * (Q + 1) / 2];
#pragma unroll
    for (int i = 0; i < Q * (Q + 1) / 2; i++) acc[i] = 0.0;
    for (int m = arow[0]; m < nb; m++) {
        double2 v[Q];
#pragma unroll
        for (int k = 0; k < Q; k++) {
            if (m >= arow[k]) {
                v[k] = *reinterpret_cast<const double2 *>(Zs + utile(arow[k], m, nb) * PGPFA_TILE + roff[k] + loff);
            } else {
                v[k].x = 0.0; v[k].y = 0.0;
            }
        }
        int idx = 0;
#pragma unroll
        for (int k = 0; k < Q; k++)
#pragma unroll
            for (int l = k; l < Q; l++) { acc[idx] += v[k].x * v[l].x + v[k].y * v[l].y; idx++; }
    }
    double *o = vsm + ((size_t)trial * T + t) * Q * Q;
    int idx = 0;
#pragma unroll
    for (int k = 0; k < Q; k++)
#pragma unroll
        for (int l = k; l < Q; l++) {
            const double s = warp_sum(acc[idx]);
            idx++;
            if (lane == 0) { o[k * Q + l] = s; o[l * Q + k] = s; }
        }
}

// log-determinant from the diagonal of the factor: 2 * sum log L_ii  (per slot)
__global__ void chol_logdet_kernel(const double *__restrict__ L, int nb, int n, double *__restrict__ out) {
    __shared__ double red[32];
    const int slot = blockIdx.x;
    const long long ltl = (long long)nb * (nb + 1) / 2;
    const double *Ls = L + (size_t)slot * ltl * PGPFA_TILE;
    double s = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int j = r >> 6, rr = r & 63;
        s += log(Ls[ltile(j, j) * PGPFA_TILE + tile_off(rr, rr)]);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[slot] = 2.0 * s;
}

// debug / test helper: packed-lower tiles -> dense row-major lower-triangular (n x n per slot)
__global__ void tiles_to_dense_kernel(const double *__restrict__ L, int nb, int n, int upper, double *__restrict__ out) {
    const int slot = blockIdx.y;
    const long long ltl = (long long)nb * (nb + 1) / 2;
    const double *Ls = L + (size_t)slot * ltl * PGPFA_TILE;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)n * n; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
        double v = 0.0;
        if (!upper && c <= r) v = Ls[ltile(r >> 6, c >> 6) * PGPFA_TILE + tile_off(r & 63, c & 63)];
        if (upper && r <= c) v = Ls[utile(r >> 6, c >> 6, nb) * PGPFA_TILE + tile_off(r & 63, c & 63)];
        out[(size_t)slot * n * n + e] = v;
    }
}

int set_smem_attrs() {
    static bool done = false;
    if (done) return PGPFA_OK;
    PGPFA_CUDA_TRY(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PGPFA_GEMM_SMEM));
    PGPFA_CUDA_TRY(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PGPFA_GEMM_SMEM));
    PGPFA_CUDA_TRY(cudaFuncSetAttribute(lauum_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PGPFA_GEMM_SMEM));
    done = true;
    return PGPFA_OK;
}

template <int Q>
void launch_timediag(const double *ZT, const int *act, double *vsm, int nb, int T, int nslots, cudaStream_t st) {
    dim3 grid((T + 3) / 4, nslots);
    cov_timediag_kernel<Q><<<grid, 128, 0, st>>>(ZT, act, vsm, nb, T);
}

}  // namespace

// =============================================================================================
// internal host API (used by laplace.cu / mstep.cu / the C-ABI wrappers in api.cu)
// =============================================================================================
static int factor_one(const PgpfaMatSrc &ms, double *L, double *Dinv, double *ZT, const int *act, int *info, int nslots,
                      cudaStream_t st, float *L32, float *D32) {
    if (nslots <= 0) return PGPFA_OK;
    PGPFA_TRY(set_smem_attrs());
    FactorArgs a;
    a.src.Kinv = ms.Kinv; a.src.W = ms.W; a.src.dense = ms.dense;
    a.src.q = ms.q; a.src.T = ms.T; a.src.n = ms.n; a.src.diag_scale = ms.diag_scale;
    a.L = L; a.Dinv = Dinv; a.ZT = ZT; a.act = act; a.info = info;
    a.L32 = L32; a.D32 = D32;
    a.nb = pgpfa_nb(ms.n);
    a.mode = 0;
    a.nslots = nslots;
    a.fuse = 1;
    const int nb = a.nb;
    // diag(0), then one launch per panel step j: its first-tile CTAs also factor diagonal tile j+1 (fused
    // look-ahead), so the latency-bound 64x64 factor/inverse never runs in front of the tensor-bound panel.
    a.step = 0;
    chol_diag_kernel<<<nslots, PGPFA_GEMM_THREADS, PGPFA_GEMM_SMEM, st>>>(a);
    PGPFA_LAUNCH_CHECK();
    for (int j = 0; j + 1 < nb; j++) {
        a.step = j;
        a.ntiles = nb - 1 - j;
        const long long blocks = (long long)nslots * a.ntiles;
        chol_panel_kernel<<<(unsigned)blocks, PGPFA_GEMM_THREADS, PGPFA_GEMM_SMEM, st>>>(a);
        PGPFA_LAUNCH_CHECK();
    }
    return PGPFA_OK;
}

static int trtri_one(const double *L, const double *Dinv, double *ZT, int n, int nslots, cudaStream_t st) {
    if (nslots <= 0) return PGPFA_OK;
    PGPFA_TRY(set_smem_attrs());
    FactorArgs a;
    a.src.Kinv = nullptr; a.src.W = nullptr; a.src.dense = nullptr; a.src.q = 1; a.src.T = n; a.src.n = n; a.src.diag_scale = 1.0;
    a.L = const_cast<double *>(L); a.Dinv = const_cast<double *>(Dinv); a.ZT = ZT; a.act = nullptr; a.info = nullptr;
    a.L32 = nullptr; a.D32 = nullptr;
    a.nb = pgpfa_nb(n);
    a.mode = 1;
    a.nslots = nslots; a.ntiles = 0; a.fuse = 0;
    for (int i = 1; i < a.nb; i++) {
        a.step = i;
        dim3 grid(i, nslots);
        chol_panel_kernel<<<grid, PGPFA_GEMM_THREADS, PGPFA_GEMM_SMEM, st>>>(a);
        PGPFA_LAUNCH_CHECK();
    }
    return PGPFA_OK;
}

int pgpfa_i_lauum(const double *ZT, const int2 *pairs, int npairs, const int *act, double *vsmGP, double *dense, int n,
                  int q, int T, int nslots, cudaStream_t st) {
    if (nslots <= 0 || npairs <= 0) return PGPFA_OK;
    PGPFA_TRY(set_smem_attrs());
    LauumArgs a;
    a.ZT = ZT; a.pairs = pairs; a.act = act; a.vsmGP = vsmGP; a.dense = dense;
    a.nb = pgpfa_nb(n); a.n = n; a.q = q; a.T = T;
    dim3 grid(npairs, nslots);
    lauum_tiles_kernel<<<grid, PGPFA_GEMM_THREADS, PGPFA_GEMM_SMEM, st>>>(a);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_timediag(const double *ZT, const int *act, double *vsm, int n, int q, int T, int nslots, cudaStream_t st) {
    if (nslots <= 0) return PGPFA_OK;
    const int nb = pgpfa_nb(n);
    switch (q) {
#define CASE_Q(QQ) case QQ: launch_timediag<QQ>(ZT, act, vsm, nb, T, nslots, st); break;
        PGPFA_FOR_EACH_Q(CASE_Q)
#undef CASE_Q
        default: return PGPFA_ERR_ARG;
    }
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_logdet(const double *L, int n, int nslots, double *out, cudaStream_t st) {
    if (nslots <= 0) return PGPFA_OK;
    chol_logdet_kernel<<<nslots, 256, 0, st>>>(L, pgpfa_nb(n), n, out);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_tiles_to_dense(const double *tiles, int n, int upper, int nslots, double *out, cudaStream_t st) {
    if (nslots <= 0) return PGPFA_OK;
    dim3 grid(64, nslots);
    tiles_to_dense_kernel<<<grid, 256, 0, st>>>(tiles, pgpfa_nb(n), n, upper, out);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

// Large batches are split into parts that run the same launch sequence on separate streams: the partial last
// wave (and the serial diagonal-tile work) of one part's launch is filled by CTAs of the others instead of leaving
// SMs idle between steps.  PGPFA_SPLIT=k overrides the number of parts (1 = off).
static int split_parts(pgpfa_handle_s *h, int nslots) {
    if (!h || !h->s_part[0]) return 1;
    static int forced = -1;
    if (forced < 0) { const char *e = getenv("PGPFA_SPLIT"); forced = e ? atoi(e) : 0; }
    int parts = forced > 0 ? forced : (nslots >= 512 ? 2 : nslots >= 64 ? 4 : 1);
    if (parts > PGPFA_MAX_PARTS) parts = PGPFA_MAX_PARTS;
    if (parts > nslots) parts = nslots > 0 ? nslots : 1;
    return parts;
}
static int fork_streams(pgpfa_handle_s *h, int parts, cudaStream_t st) {
    PGPFA_CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    for (int p = 0; p < parts; p++) PGPFA_CUDA_TRY(cudaStreamWaitEvent(h->s_part[p], h->ev_fork, 0));
    return PGPFA_OK;
}
static int join_streams(pgpfa_handle_s *h, int parts, cudaStream_t st) {
    for (int p = 0; p < parts; p++) {
        PGPFA_CUDA_TRY(cudaEventRecord(h->ev_join[p], h->s_part[p]));
        PGPFA_CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join[p], 0));
    }
    return PGPFA_OK;
}

int pgpfa_i_factor(const PgpfaMatSrc &ms, double *L, double *Dinv, double *ZT, const int *act, int *info, int nslots,
                   cudaStream_t st, pgpfa_handle_s *h, float *L32, float *D32) {
    const int parts = split_parts(h, nslots);
    if (parts <= 1) return factor_one(ms, L, Dinv, ZT, act, info, nslots, st, L32, D32);
    const int nb = pgpfa_nb(ms.n);
    const size_t lt = (size_t)pgpfa_ltiles(nb) * PGPFA_TILE, dt = (size_t)nb * PGPFA_TILE;
    PGPFA_TRY(fork_streams(h, parts, st));
    int rc = PGPFA_OK;
    for (int p = 0; p < parts; p++) {
        const int s0 = (int)((long long)nslots * p / parts), s1 = (int)((long long)nslots * (p + 1) / parts);
        // factor storage is slot-indexed: shift the bases.  With an active list the matrix source and info stay
        // trial-indexed through act; without one trial == slot, so they are shifted as well.
        PgpfaMatSrc m = ms;
        int *inf = info;
        if (m.dense) m.dense += (size_t)s0 * ms.n * ms.n;      // a dense source is always slot-indexed
        if (!act) {
            if (m.W) m.W += (size_t)s0 * ms.q * ms.q * ms.T;
            if (inf) inf += s0;
        }
        const int r = factor_one(m, L + s0 * lt, Dinv + s0 * dt, ZT ? ZT + s0 * lt : nullptr, act ? act + s0 : nullptr, inf,
                                 s1 - s0, h->s_part[p], L32 ? L32 + s0 * lt : nullptr, D32 ? D32 + s0 * dt : nullptr);
        if (rc == PGPFA_OK) rc = r;
    }
    PGPFA_TRY(join_streams(h, parts, st));
    return rc;
}

int pgpfa_i_trtri(const double *L, const double *Dinv, double *ZT, int n, int nslots, cudaStream_t st, pgpfa_handle_s *h) {
    const int parts = split_parts(h, nslots);
    if (parts <= 1) return trtri_one(L, Dinv, ZT, n, nslots, st);
    const int nb = pgpfa_nb(n);
    const size_t lt = (size_t)pgpfa_ltiles(nb) * PGPFA_TILE, dt = (size_t)nb * PGPFA_TILE;
    PGPFA_TRY(fork_streams(h, parts, st));
    int rc = PGPFA_OK;
    for (int p = 0; p < parts; p++) {
        const int s0 = (int)((long long)nslots * p / parts), s1 = (int)((long long)nslots * (p + 1) / parts);
        const int r = trtri_one(L + s0 * lt, Dinv + s0 * dt, ZT + s0 * lt, n, s1 - s0, h->s_part[p]);
        if (rc == PGPFA_OK) rc = r;
    }
    PGPFA_TRY(join_streams(h, parts, st));
    return rc;
}
