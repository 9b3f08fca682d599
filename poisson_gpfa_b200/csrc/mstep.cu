// M-step kernels.
//  * C,d: the reference minimises MStepObservationCost (funs/learning.py:20-91) with scipy; the cost
//    is separable over neurons and convex, so here every neuron runs its own (q+1)-dimensional damped
//    Newton.  One pass over (y, post_mean, post_vsm) produces per-neuron cost / gradient / Hessian
//    sums (HBM-bound streaming reduction, deterministic two-stage sum); the per-neuron update kernel
//    does accept/reject + the next Newton step.  Between the two the host all-reduces `stats` across
//    GPUs (trial sharding).
//  * tau: cost and gradient of MStepGPtimescaleCost (funs/learning.py:175-255, prior variant
//    :681-769) for all latents at once, on the batched SPD inverse.
#include "common.cuh"
#include "pgpfa_internal.h"
#include "tau_search.h"

using namespace pgpfa;

namespace {

#define CD_TT 16   // bins per work item

__device__ __forceinline__ void cd_cp_async8(double *dst_smem, const double *src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}

// Per-neuron cost / gradient / Hessian sums of MStepObservationCost (funs/learning.py:20-91) over the trials and bins of
// this rank: thread <-> neuron, a work item = 16 bins of one trial staged in shared memory (counts, posterior means,
// per-bin covariance blocks).  The next item's tile is copied (cp.async) while the current one is reduced, every
// update of an accumulator is ONE fused multiply-add (the first version left the compiler with a*b + c*d -> DMUL,
// DFMA, DADD: 249 FP64 instructions per sample, now ~190), three CTAs per SM.
template <int Q>
__global__ void __launch_bounds__(256, 1) mstep_cd_stats_kernel(const double *__restrict__ y, const double *__restrict__ m,
                                                                const double *__restrict__ vsm,
                                                                const double *__restrict__ theta, int R, int N, int T,
                                                                double *__restrict__ partial,
                                                                const int *__restrict__ skip_if_zero) {
    constexpr int P = Q + 1;
    constexpr int NS = 1 + P + P * (P + 1) / 2;
    if (skip_if_zero && *skip_if_zero == 0) return;       // device-driven Newton loop: every neuron has converged
    extern __shared__ double sm[];
    const int stage_doubles = N * (CD_TT + 1) + Q * CD_TT + CD_TT * Q * Q;
    const int nTT = (T + CD_TT - 1) / CD_TT;
    const long long items = (long long)R * nTT;
    auto stage_load = [&](long long item, int st) {
        double *ys = sm + (size_t)st * stage_doubles;     // N x (CD_TT+1)
        double *ms = ys + (size_t)N * (CD_TT + 1);        // Q x CD_TT
        double *vs = ms + Q * CD_TT;                      // CD_TT x Q*Q
        const int r = (int)(item / nTT);
        const int t0 = (int)(item - (long long)r * nTT) * CD_TT;
        const int tl = (T - t0) < CD_TT ? (T - t0) : CD_TT;
        for (int i = threadIdx.x; i < N * CD_TT; i += blockDim.x) {
            const int nn = i >> 4, tt = i & (CD_TT - 1);
            const bool ok = tt < tl;
            cd_cp_async8(ys + nn * (CD_TT + 1) + tt, ok ? y + ((size_t)r * N + nn) * T + t0 + tt : y, ok ? 8 : 0);
        }
        for (int i = threadIdx.x; i < Q * CD_TT; i += blockDim.x) {
            const int k = i >> 4, tt = i & (CD_TT - 1);
            const bool ok = tt < tl;
            cd_cp_async8(ms + i, ok ? m + ((size_t)r * Q + k) * T + t0 + tt : m, ok ? 8 : 0);
        }
        for (int i = threadIdx.x; i < CD_TT * Q * Q; i += blockDim.x) {
            const bool ok = (i / (Q * Q)) < tl;
            cd_cp_async8(vs + i, ok ? vsm + ((size_t)r * T + t0) * Q * Q + i : vsm, ok ? 8 : 0);
        }
    };
    const int nloc = threadIdx.x;             // neuron handled by this thread (block covers blockDim neurons per pass)
    for (int n0 = 0; n0 < N; n0 += blockDim.x) {
        const int n = n0 + nloc;
        const bool live = n < N;
        double c[Q], dd = 0.0;
#pragma unroll
        for (int k = 0; k < Q; k++) c[k] = live ? theta[(size_t)n * P + k] : 0.0;
        if (live) dd = theta[(size_t)n * P + Q];
        double st[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) st[i] = 0.0;
        long long item = blockIdx.x;
        __syncthreads();                                   // previous neuron pass done with both stages
        if (item < items) stage_load(item, 0);
        asm volatile("cp.async.commit_group;" ::: "memory");
        int cur = 0;
        for (; item < items; item += gridDim.x, cur ^= 1) {
            if (item + gridDim.x < items) stage_load(item + gridDim.x, cur ^ 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncthreads();
            const double *ys = sm + (size_t)cur * stage_doubles;
            const double *ms = ys + (size_t)N * (CD_TT + 1);
            const double *vs = ms + Q * CD_TT;
            const int t0 = (int)(item % nTT) * CD_TT;
            const int tl = (T - t0) < CD_TT ? (T - t0) : CD_TT;
            if (live) {
                for (int tt = 0; tt < tl; tt++) {
                    double u[Q];
                    double h = dd, s = 0.0;
                    const double *V = vs + tt * Q * Q;
#pragma unroll
                    for (int k = 0; k < Q; k++) {
                        double a = 0.0;
#pragma unroll
                        for (int l = 0; l < Q; l++) a = fma(V[k * Q + l], c[l], a);
                        s = fma(c[k], a, s);
                        const double mk = ms[k * CD_TT + tt];
                        h = fma(c[k], mk, h);
                        u[k] = mk + a;
                    }
                    const double yh = exp(fma(0.5, s, h));
                    const double yv = ys[n * (CD_TT + 1) + tt];
                    st[0] += fma(-yv, h, yh);
                    int idx = 1 + P;
#pragma unroll
                    for (int k = 0; k < Q; k++) {
                        const double yu = yh * u[k];
                        st[1 + k] += fma(-yv, ms[k * CD_TT + tt], yu);
#pragma unroll
                        for (int l = k; l < Q; l++) { st[idx] = fma(yu, u[l], fma(yh, V[k * Q + l], st[idx])); idx++; }
                        st[idx] += yu;   // H[k][d]
                        idx++;
                    }
                    st[1 + Q] += yh - yv;
                    st[idx] += yh;       // H[d][d]
                }
            }
            __syncthreads();                               // stage `cur` is free for the load after next
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (live) {
#pragma unroll
            for (int i = 0; i < NS; i++) partial[((size_t)blockIdx.x * NS + i) * N + n] = st[i];
        }
    }
}

__global__ void mstep_cd_reduce_kernel(const double *__restrict__ partial, int nblocks, int len, double *__restrict__ out,
                                       const int *__restrict__ skip_if_zero) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    if (skip_if_zero && *skip_if_zero == 0) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; b++) s += partial[(size_t)b * len + i];
    out[i] = s;
}

// per-neuron accept/reject and next Newton step.  State arrays are (N) or (N, q+1).
template <int Q>
__global__ void mstep_cd_update_kernel(const double *__restrict__ stats, double invR, double pw,
                                       const double *__restrict__ pmat, const double *__restrict__ theta0, double *__restrict__ theta_cur,
                                       double *__restrict__ theta_try, double *__restrict__ fcur,
                                       double *__restrict__ step, double *__restrict__ alpha,
                                       double *__restrict__ slope, int *__restrict__ done, int first, double tol, int N,
                                       int *__restrict__ n_open, int iter_index) {
    constexpr int P = Q + 1;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    if (done[n]) return;
    if (iter_index > 0) atomicMax(n_open + 1, iter_index);     // last iteration that still had an open neuron
    // prior: 0.5 pw |theta - theta0|^2 ('useDiag'), or 0.5 (theta-theta0)^T M_n (theta-theta0) with a per-neuron
    // (q+1)x(q+1) matrix M_n (packed upper, pmat[b*N+n]) for the accumulated-Hessian rule ('useHessian')
    double tt[P], t0[P], dl0[P], Md[P];
    double pen = 0.0;
#pragma unroll
    for (int k = 0; k < P; k++) {
        tt[k] = theta_try[(size_t)n * P + k];
        t0[k] = theta0[(size_t)n * P + k];
        dl0[k] = tt[k] - t0[k];
        Md[k] = pw * dl0[k];
    }
    if (pmat) {
        int idx = 0;
#pragma unroll
        for (int k = 0; k < P; k++) Md[k] = 0.0;
#pragma unroll
        for (int k = 0; k < P; k++)
#pragma unroll
            for (int l = k; l < P; l++) {
                const double m = pmat[(size_t)idx * N + n];
                Md[k] += m * dl0[l];
                if (l != k) Md[l] += m * dl0[k];
                idx++;
            }
    }
#pragma unroll
    for (int k = 0; k < P; k++) pen += dl0[k] * Md[k];
    const double ftry = stats[n] * invR + 0.5 * pen;
    bool accept = first != 0;
    if (!accept) {
        const double f0 = fcur[n], sl = slope[n], al = alpha[n];
        const bool tiny = fabs(sl) <= 1e-10 * (1.0 + fabs(f0));
        accept = isfinite(ftry) && (tiny || ftry <= f0 + 1e-4 * al * sl + 1e-14 * (1.0 + fabs(f0)));
        if (!accept && al < 1e-12) accept = isfinite(ftry);   // give up backtracking, keep going
    }
    if (!accept) {
        const double al = 0.5 * alpha[n];
        alpha[n] = al;
#pragma unroll
        for (int k = 0; k < P; k++) theta_try[(size_t)n * P + k] = theta_cur[(size_t)n * P + k] + al * step[(size_t)n * P + k];
        atomicAdd(n_open, 1);
        return;
    }
    // accepted: gradient / Hessian at theta_try, Cholesky solve of the (q+1) system
    double g[P], H[P][P];
#pragma unroll
    for (int k = 0; k < P; k++) g[k] = stats[(size_t)(1 + k) * N + n] * invR + Md[k];
    {
        int idx = 1 + P;
#pragma unroll
        for (int k = 0; k < P; k++)
#pragma unroll
            for (int l = k; l < P; l++) {
                const double v = stats[(size_t)idx * N + n] * invR +
                                 (pmat ? pmat[(size_t)(idx - 1 - P) * N + n] : ((k == l) ? pw : 0.0));
                H[k][l] = v;
                H[l][k] = v;
                idx++;
            }
    }
#pragma unroll
    for (int j = 0; j < P; j++) {
        double dj = H[j][j];
#pragma unroll
        for (int k = 0; k < P; k++) if (k < j) dj -= H[j][k] * H[j][k];
        dj = sqrt(fmax(dj, 1e-300));
        H[j][j] = dj;
#pragma unroll
        for (int i = 0; i < P; i++)
            if (i > j) {
                double v = H[i][j];
#pragma unroll
                for (int k = 0; k < P; k++) if (k < j) v -= H[i][k] * H[j][k];
                H[i][j] = v / dj;
            }
    }
    double z[P], dl[P];
#pragma unroll
    for (int i = 0; i < P; i++) {
        double v = -g[i];
#pragma unroll
        for (int k = 0; k < P; k++) if (k < i) v -= H[i][k] * z[k];
        z[i] = v / H[i][i];
    }
#pragma unroll
    for (int i = P - 1; i >= 0; i--) {
        double v = z[i];
#pragma unroll
        for (int k = 0; k < P; k++) if (k > i) v -= H[k][i] * dl[k];
        dl[i] = v / H[i][i];
    }
    double sl = 0.0, dmax = 0.0, tmax = 0.0;
#pragma unroll
    for (int k = 0; k < P; k++) { sl += g[k] * dl[k]; dmax = fmax(dmax, fabs(dl[k])); tmax = fmax(tmax, fabs(tt[k])); }
    // quadratic-convergence predictor: after a full Newton step of size p followed by one of size d, the error
    // left after applying d is ~ d^3 / p^2
    double prevmax = 0.0;
    if (!first && alpha[n] == 1.0) {
#pragma unroll
        for (int k = 0; k < P; k++) prevmax = fmax(prevmax, fabs(step[(size_t)n * P + k]));
    }
    fcur[n] = ftry;
    slope[n] = sl;
    alpha[n] = 1.0;
    bool conv = dmax <= tol * (1.0 + tmax);
    if (!conv && prevmax > 0.0 && dmax < 0.1 * prevmax && dmax * dmax * dmax / (prevmax * prevmax) <= 0.01 * tol * (1.0 + tmax))
        conv = true;
#pragma unroll
    for (int k = 0; k < P; k++) {
        step[(size_t)n * P + k] = dl[k];
        const double nx = tt[k] + dl[k];
        theta_cur[(size_t)n * P + k] = conv ? nx : tt[k];
        theta_try[(size_t)n * P + k] = nx;
    }
    if (conv) {
        done[n] = 1;
        fcur[n] = ftry + 0.5 * sl;     // cost at the final point theta + step (second-order model, error O(step^3))
    } else {
        atomicAdd(n_open, 1);
    }
}

// C = A * B for a batch of row-major n x n matrices on the FP64 tensor pipe (DMMA.8x8x4): CTA = 64x64 tile of C,
// 4 warps as 2x2 of 32x32, k in chunks of 16 through double-buffered padded shared memory (conflict-free fragment
// loads: A stride 20 and B stride 68 doubles put the 16 lanes of a half-warp on 16 distinct 8-byte banks); the next
// chunk's global loads are in flight while the current one is multiplied.
// Two problem sets in one launch (grid.z = 2 nb): batch b < nb is C[b] = A[b] B[b]; batch nb + b is
// C2[b] = A2[b % qmod] A[b] (the timescale gradient needs K^-1 dK and P K^-1: independent products, one launch).
__global__ void __launch_bounds__(128) small_gemm_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                         double *__restrict__ Cm, int n, const double *__restrict__ A2,
                                                         double *__restrict__ C2, int nb, int qmod) {
    __shared__ double As[2][64][20], Bs[2][16][68];
    const int b = blockIdx.z;
    const double *Ab, *Bb;
    double *Cb;
    if (b < nb) {
        Ab = A + (size_t)b * n * n; Bb = B + (size_t)b * n * n; Cb = Cm + (size_t)b * n * n;
    } else {
        const int b2 = b - nb;
        Ab = A2 + (size_t)(b2 % qmod) * n * n; Bb = A + (size_t)b2 * n * n; Cb = C2 + (size_t)b2 * n * n;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    const int fr = lane >> 2, fk = lane & 3;
    const int ar = tid >> 4, ac = tid & 15;        // A: rows ar + 8 j, column ac of the chunk
    const int bk = tid >> 6, bc = tid & 63;        // B: rows bk + 2 j, column bc of the chunk
    double ra[8], rb[8];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const int r = r0 + ar + 8 * jj, kk = k0 + bk + 2 * jj;
            ra[jj] = (r < n && k0 + ac < n) ? Ab[(size_t)r * n + k0 + ac] : 0.0;
            rb[jj] = (kk < n && c0 + bc < n) ? Bb[(size_t)kk * n + c0 + bc] : 0.0;
        }
    };
    double acc[4][4][2] = {};
    fetch(0);
    int buf = 0;
    for (int k0 = 0; k0 < n; k0 += 16, buf ^= 1) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            As[buf][ar + 8 * jj][ac] = ra[jj];
            Bs[buf][bk + 2 * jj][bc] = rb[jj];
        }
        __syncthreads();
        if (k0 + 16 < n) fetch(k0 + 16);
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            double a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = As[buf][wm * 32 + i * 8 + fr][k4 + fk];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) bb[jj] = Bs[buf][k4 + fk][wn * 32 + jj * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) dmma884(acc[i][jj][0], acc[i][jj][1], a[i], bb[jj]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jj = 0; jj < 4; jj++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int r = r0 + wm * 32 + i * 8 + fr, c = c0 + wn * 32 + jj * 8 + 2 * fk + e;
                if (r < n && c < n) Cb[(size_t)r * n + c] = acc[i][jj][e];
            }
}

// cost / gradient of the timescale objective from Kinv, logdet, dK, G = Kinv dK Kinv and PautoSum.
// Stage 1: TAU_PARTS CTAs per slot reduce slices of the three traces; stage 2 adds them in a fixed order.
#define TAU_PARTS 8
// G = P K^-1 and M1 = K^-1 dK (small_gemm_kernel): t3 = tr(K^-1 dK K^-1 P) = sum_e M1[e] G[e]
__global__ void __launch_bounds__(256) tau_trace_kernel(const double *__restrict__ Kinv, const double *__restrict__ dK,
                                                        const double *__restrict__ G, const double *__restrict__ P, int T,
                                                        double *__restrict__ part, int qmod,
                                                        const double *__restrict__ M1) {
    __shared__ double red[32];
    const int k = blockIdx.x;
    const size_t off = (size_t)k * T * T, poff = (size_t)(k % qmod) * T * T;
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < T * T; e += blockDim.x * TAU_PARTS) {
        const double ki = Kinv[off + e], pp = P[poff + e];
        t1 += ki * pp;
        t2 += ki * dK[off + e];
        t3 += G[off + e] * M1[off + e];
    }
    t1 = block_sum(t1, red);
    t2 = block_sum(t2, red);
    t3 = block_sum(t3, red);
    if (threadIdx.x == 0) {
        double *o = part + ((size_t)k * TAU_PARTS + blockIdx.y) * 3;
        o[0] = t1; o[1] = t2; o[2] = t3;
    }
}

__global__ void tau_reduce_kernel(const double *__restrict__ p, const double *__restrict__ part,
                                  const double *__restrict__ logdet, int nslots, double R, double pw,
                                  const double *__restrict__ tau_old, double bs, double *__restrict__ cost,
                                  double *__restrict__ grad, int qmod) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nslots) return;
    double t1 = 0.0, t2 = 0.0, t3 = 0.0;
    for (int i = 0; i < TAU_PARTS; i++) {
        const double *o = part + ((size_t)k * TAU_PARTS + i) * 3;
        t1 += o[0]; t2 += o[1]; t3 += o[2];
    }
    double c = 0.5 * R * logdet[k] + 0.5 * t1;
    const double dE = -0.5 * R * t2 + 0.5 * t3;
    double g = -dE * exp(p[k]);
    if (pw > 0.0) {
        const double tau = bs / 1000.0 * sqrt(1.0 / exp(p[k]));
        const double dt = tau - tau_old[k % qmod];
        c += 0.5 * dt * dt * pw;
        g += dt * pw;      // as written in funs/learning.py:734,769 (no chain-rule factor)
    }
    cost[k] = c;
    grad[k] = g;
}

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
inline int cd_blocks() { return 148 * 3; }
inline int cd_threads(int N) { int t = ((N + 31) / 32) * 32; return t > 256 ? 256 : t; }

template <int Q>
int launch_cd_stats(const double *y, const double *m, const double *vsm, const double *theta, int R, int N, int T,
                    double *partial, int nblocks, cudaStream_t st, const int *skip_if_zero) {
    const size_t smem = 2 * ((size_t)N * (CD_TT + 1) + Q * CD_TT + CD_TT * Q * Q) * sizeof(double);     // two stages
    if (smem > 48 * 1024)
        PGPFA_CUDA_TRY(cudaFuncSetAttribute(mstep_cd_stats_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mstep_cd_stats_kernel<Q><<<nblocks, cd_threads(N), smem, st>>>(y, m, vsm, theta, R, N, T, partial, skip_if_zero);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

}  // namespace

extern "C" int pgpfa_mstep_cd_nstats(int q) { return 1 + (q + 1) + (q + 1) * (q + 2) / 2; }

extern "C" long long pgpfa_mstep_cd_workspace_bytes(int q, int N) {
    if (q <= 0 || N <= 0) return -1;
    return (long long)align_up((size_t)cd_blocks() * pgpfa_mstep_cd_nstats(q) * N * 8) + 512;
}

static int cd_stats_impl(const double *y, const double *m, const double *vsm, const double *theta, int R, int q, int N,
                         int T, double *stats, void *workspace, long long ws_bytes, cudaStream_t st,
                         const int *skip_if_zero) {
    if (!stats || R < 0 || q <= 0 || q > PGPFA_QMAX || N <= 0 || T <= 0) return PGPFA_ERR_ARG;
    const int len = pgpfa_mstep_cd_nstats(q) * N;
    if (R == 0) {                  // a rank without trials (mini-batch smaller than the world) contributes zeros
        PGPFA_CUDA_TRY(cudaMemsetAsync(stats, 0, (size_t)len * 8, st));
        return PGPFA_OK;
    }
    if (!y || !m || !vsm || !theta || !workspace) return PGPFA_ERR_ARG;
    if (ws_bytes < pgpfa_mstep_cd_workspace_bytes(q, N)) return PGPFA_ERR_WORKSPACE;
    double *partial = reinterpret_cast<double *>(align_up(reinterpret_cast<size_t>(workspace)));
    const long long items = (long long)R * ((T + CD_TT - 1) / CD_TT);
    int nblocks = cd_blocks();
    if (items < nblocks) nblocks = (int)items;
    int rc = PGPFA_ERR_ARG;
    switch (q) {
#define CASE_Q(QQ) case QQ: rc = launch_cd_stats<QQ>(y, m, vsm, theta, R, N, T, partial, nblocks, st, skip_if_zero); break;
        PGPFA_FOR_EACH_Q(CASE_Q)
#undef CASE_Q
    }
    PGPFA_TRY(rc);
    mstep_cd_reduce_kernel<<<(len + 255) / 256, 256, 0, st>>>(partial, nblocks, len, stats, skip_if_zero);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

extern "C" int pgpfa_mstep_cd_stats(const double *y, const double *m, const double *vsm, const double *theta, int R,
                                    int q, int N, int T, double *stats, void *workspace, long long ws_bytes,
                                    cudaStream_t st) {
    return cd_stats_impl(y, m, vsm, theta, R, q, N, T, stats, workspace, ws_bytes, st, nullptr);
}

/* same, as one iteration of a device-driven Newton loop: when *n_open (device int, written by pgpfa_mstep_cd_update)
 * is zero the launch exits at once and leaves `stats` untouched */
extern "C" int pgpfa_mstep_cd_stats_gated(const double *y, const double *m, const double *vsm, const double *theta, int R,
                                          int q, int N, int T, double *stats, void *workspace, long long ws_bytes,
                                          const int *n_open, cudaStream_t st) {
    return cd_stats_impl(y, m, vsm, theta, R, q, N, T, stats, workspace, ws_bytes, st, n_open);
}

extern "C" int pgpfa_mstep_cd_update(const double *stats, double inv_R, double prior_w, const double *prior_mat,
                                     const double *theta0,
                                     double *theta_cur, double *theta_try, double *fcur, double *step, double *alpha,
                                     double *slope, int *done, int first, double tol, int N, int q, int *n_open,
                                     int iter_index, cudaStream_t st) {
    if (!stats || !theta0 || !theta_cur || !theta_try || !fcur || !step || !alpha || !slope || !done || !n_open)
        return PGPFA_ERR_ARG;
    PGPFA_CUDA_TRY(cudaMemsetAsync(n_open, 0, sizeof(int), st));
    const int blocks = (N + 63) / 64;
    switch (q) {
#define CASE_Q(QQ) case QQ: mstep_cd_update_kernel<QQ><<<blocks, 64, 0, st>>>(stats, inv_R, prior_w, prior_mat, theta0, theta_cur, theta_try, fcur, step, alpha, slope, done, first, tol, N, n_open, iter_index); break;
        PGPFA_FOR_EACH_Q(CASE_Q)
#undef CASE_Q
        default: return PGPFA_ERR_ARG;
    }
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

extern "C" long long pgpfa_tau_eval_workspace_bytes(int q, int T) {
    if (q <= 0 || T <= 0) return -1;
    const size_t mat = align_up((size_t)q * T * T * 8);
    return (long long)(5 * mat + align_up((size_t)q * 8) + align_up((size_t)q * 4) + align_up((size_t)q * TAU_PARTS * 24)) +
           pgpfa_spd_inverse_workspace_bytes(q, T) + 1024;
}

// C[b] = A[b] B[b] for `batch` n x n matrices on DMMA (internal: the CG preconditioner's N_k = K_k^-1 M_k^-1)
int pgpfa_i_small_gemm(const double *A, const double *B, double *C, int n, int batch, cudaStream_t st) {
    if (batch <= 0 || n <= 0) return PGPFA_OK;
    dim3 grid((n + 63) / 64, (n + 63) / 64, batch);
    small_gemm_kernel<<<grid, 128, 0, st>>>(A, B, C, n, nullptr, nullptr, batch, 1);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

static int tau_eval_impl(const double *p, const double *Psum, double numTrials, int q, int qmod, int T, double eps,
                         double prior_w, const double *tau_old, double bs, double *cost, double *grad, void *workspace,
                         long long ws_bytes, cudaStream_t st) {
    if (!p || !Psum || !cost || !grad || !workspace || q <= 0 || T <= 0) return PGPFA_ERR_ARG;
    if (prior_w > 0.0 && !tau_old) return PGPFA_ERR_ARG;
    if (ws_bytes < pgpfa_tau_eval_workspace_bytes(q, T)) return PGPFA_ERR_WORKSPACE;
    unsigned char *w = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(workspace)));
    const size_t mat = align_up((size_t)q * T * T * 8);
    double *K = (double *)w; w += mat;
    double *dK = (double *)w; w += mat;
    double *Kinv = (double *)w; w += mat;
    double *M1 = (double *)w; w += mat;
    double *G = (double *)w; w += mat;
    double *logdet = (double *)w; w += align_up((size_t)q * 8);
    int *info = (int *)w; w += align_up((size_t)q * 4);
    double *part = (double *)w; w += align_up((size_t)q * TAU_PARTS * 24);
    const long long inv_bytes = pgpfa_spd_inverse_workspace_bytes(q, T);
    PGPFA_TRY(pgpfa_make_K_gamma(p, q, T, eps, K, dK, st));
    PGPFA_TRY(pgpfa_spd_inverse_batched(K, q, T, Kinv, logdet, info, w, inv_bytes, st));
    // tr(K^-1 dK K^-1 P) = sum_ij (K^-1 dK)_ij (P K^-1)_ij: the two products are independent, one launch
    dim3 grid((T + 63) / 64, (T + 63) / 64, 2 * q);
    small_gemm_kernel<<<grid, 128, 0, st>>>(Kinv, dK, M1, T, Psum, G, q, qmod);
    PGPFA_LAUNCH_CHECK();
    dim3 gtr(q, TAU_PARTS);
    tau_trace_kernel<<<gtr, 256, 0, st>>>(Kinv, dK, G, Psum, T, part, qmod, M1);
    PGPFA_LAUNCH_CHECK();
    tau_reduce_kernel<<<(q + 127) / 128, 128, 0, st>>>(p, part, logdet, q, numTrials, prior_w, tau_old, bs, cost, grad, qmod);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

extern "C" int pgpfa_tau_eval(const double *p, const double *Psum, double numTrials, int q, int T, double eps,
                              double prior_w, const double *tau_old, double bs, double *cost, double *grad,
                              void *workspace, long long ws_bytes, cudaStream_t st) {
    return tau_eval_impl(p, Psum, numTrials, q, q, T, eps, prior_w, tau_old, bs, cost, grad, workspace, ws_bytes, st);
}

// ---------------------------------------------------------------------------------------------
// Device-driven C,d Newton loop and timescale search (no host reads inside; funs/learning.py:124-130, :283-288)
// ---------------------------------------------------------------------------------------------
/* n_iters Newton iterations on every neuron, starting with iteration index first_iter (1 = the first one of a
 * solve: theta_cur = theta_try = theta0, done = 0, n_open[0..3] = 0 set by the caller).  Each iteration is
 * pgpfa_mstep_cd_stats_gated + pgpfa_mstep_cd_update; once every neuron has converged the remaining iterations are
 * empty launches.  n_open (device int[4]): [0] neurons still open after the last iteration, [1] index of the last
 * iteration that had an open neuron.  Single-rank only (with trial sharding the statistics have to be all-reduced
 * between the two calls: poisson_gpfa_b200/core.py runs the same two entry points around the collective). */
extern "C" int pgpfa_mstep_cd_solve(const double *y, const double *post_mean, const double *vsm, int R, int q, int N, int T,
                                    double inv_R, double prior_w, const double *prior_mat, const double *theta0,
                                    double *theta_cur, double *theta_try, double *fcur, double *step, double *alpha,
                                    double *slope, int *done, int *n_open, double *stats, int first_iter, int n_iters,
                                    double tol, void *workspace, long long ws_bytes, cudaStream_t st) {
    if (first_iter < 1 || n_iters < 0) return PGPFA_ERR_ARG;
    for (int it = first_iter; it < first_iter + n_iters; it++) {
        PGPFA_TRY(cd_stats_impl(y, post_mean, vsm, theta_try, R, q, N, T, stats, workspace, ws_bytes, st,
                                it == 1 ? nullptr : n_open));
        PGPFA_TRY(pgpfa_mstep_cd_update(stats, inv_R, prior_w, prior_mat, theta0, theta_cur, theta_try, fcur, step, alpha,
                                        slope, done, it == 1 ? 1 : 0, tol, N, q, n_open, it, st));
    }
    return PGPFA_OK;
}

namespace {
// one thread per latent; phase 0: first candidates, phase 1: merge the round-0 evaluation and propose, phase >= 2:
// merge a later round and propose.  cands / cost / grad are (m, q) with slot = c*q + k.
// flags (device int[4]): [0] latents still open, [1] evaluations that had an open latent (nfev),
// [2] bit k set: a sign change of the gradient was bracketed for latent k, [3] bit k set: the search walked to the
// edge of the admissible range without one (the old timescale is kept).
__global__ void tau_ctrl_kernel(TauLatent *st, int q, int m, int phase, const double *__restrict__ tau_old, double bs,
                                double xtol, double *__restrict__ cands, const double *__restrict__ cost,
                                const double *__restrict__ grad, int *__restrict__ flags, double *__restrict__ tau_new,
                                double *__restrict__ details) {
    __shared__ int s_open, s_br, s_walk;
    const int k = threadIdx.x;
    if (k == 0) { s_open = 0; s_br = 0; s_walk = 0; }
    __syncthreads();
    double c[TAU_MAXC];
    if (k < q) {
        TauLatent &s = st[k];
        if (phase == 0) {
            tau_init(s, tau_old[k], bs, m, c);
            for (int i = 0; i < m; i++) cands[i * q + k] = c[i];
            atomicAdd(&s_open, 1);
        } else {
            const bool was_open = flags[0] > 0;
            if (was_open) {
                double g[TAU_MAXC], f[TAU_MAXC];
                for (int i = 0; i < m; i++) { c[i] = cands[i * q + k]; g[i] = grad[i * q + k]; f[i] = cost[i * q + k]; }
                tau_merge(s, m, c, g, f, phase == 1);
                const int dn = tau_next(s, m, xtol, c);
                for (int i = 0; i < m; i++) cands[i * q + k] = c[i];
                if (!dn) atomicAdd(&s_open, 1);
            }
            if (s.bracketed) atomicOr(&s_br, 1 << k);
            if (s.walked_out) atomicOr(&s_walk, 1 << k);
            double p_new, fun, gr;
            tau_result(s, p_new, fun, gr);
            tau_new[k] = sqrt(1.0 / exp(p_new)) * bs / 1000.0;
            details[0 * q + k] = p_new; details[1 * q + k] = s.p0; details[2 * q + k] = gr; details[3 * q + k] = fun;
            details[4 * q + k] = s.f0; details[5 * q + k] = s.g0;
        }
    }
    __syncthreads();
    if (k == 0) {
        if (phase == 0) { flags[0] = s_open; flags[1] = 1; flags[2] = 0; flags[3] = 0; }
        else {
            const bool was_open = flags[0] > 0;
            flags[0] = was_open ? s_open : 0;
            if (was_open && s_open > 0) flags[1] += 1;
            flags[2] = s_br; flags[3] = s_walk;
        }
    }
}
}  // namespace

extern "C" long long pgpfa_tau_solve_workspace_bytes(int q, int T, int m) {
    if (q <= 0 || q > PGPFA_QMAX || T <= 0 || m < 5 || m > TAU_MAXC || !(m & 1)) return -1;
    return pgpfa_tau_eval_workspace_bytes(q * m, T) + (long long)align_up((size_t)q * sizeof(TauLatent)) +
           3 * (long long)align_up((size_t)q * m * 8) + 1024;
}

/* GP-timescale M-step for all latents on the device (funs/learning.py:257-293; prior variant :771-830 with
 * prior_w = 1/step^2): rounds of [m candidate points per latent -> one batched cost/gradient evaluation -> bracket /
 * inverse-interpolation controller (tau_search.h)].  first_round = 0 starts a search (n_rounds evaluations follow);
 * a later call with first_round = the number of evaluations done so far continues it on the SAME workspace.
 * Nothing is read by the host: flags (device int[4]) = {latents still open, evaluations used, bracketed mask,
 * walked-out mask}; tau_new (q, seconds) and details (6, q) = {p, p0, grad, fun, fun0, grad0} are refreshed after
 * every round.  Rounds issued after every latent is done re-evaluate the final points and change nothing. */
extern "C" int pgpfa_mstep_tau_solve(const double *Psum, const double *tau_old, double numTrials, int q, int T, double eps,
                                     double prior_w, double bs, double xtol, int m, int first_round, int n_rounds,
                                     double *tau_new, double *details, int *flags, void *workspace, long long ws_bytes,
                                     cudaStream_t st) {
    if (!Psum || !tau_old || !tau_new || !details || !flags || !workspace || first_round < 0 || n_rounds < 0)
        return PGPFA_ERR_ARG;
    const long long need = pgpfa_tau_solve_workspace_bytes(q, T, m);
    if (need < 0) return PGPFA_ERR_ARG;
    if (ws_bytes < need) return PGPFA_ERR_WORKSPACE;
    unsigned char *w = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(workspace)));
    TauLatent *state = (TauLatent *)w; w += align_up((size_t)q * sizeof(TauLatent));
    double *cands = (double *)w; w += align_up((size_t)q * m * 8);
    double *cost = (double *)w; w += align_up((size_t)q * m * 8);
    double *grad = (double *)w; w += align_up((size_t)q * m * 8);
    const long long ev_bytes = pgpfa_tau_eval_workspace_bytes(q * m, T);
    if (first_round == 0) {
        tau_ctrl_kernel<<<1, 32, 0, st>>>(state, q, m, 0, tau_old, bs, xtol, cands, cost, grad, flags, tau_new, details);
        PGPFA_LAUNCH_CHECK();
    }
    for (int r = first_round; r < first_round + n_rounds; r++) {
        PGPFA_TRY(tau_eval_impl(cands, Psum, numTrials, q * m, q, T, eps, prior_w, tau_old, bs, cost, grad, w, ev_bytes, st));
        tau_ctrl_kernel<<<1, 32, 0, st>>>(state, q, m, r == 0 ? 1 : 2, tau_old, bs, xtol, cands, cost, grad, flags, tau_new,
                                          details);
        PGPFA_LAUNCH_CHECK();
    }
    return PGPFA_OK;
}
