// Laplace E-step on device: fused rate / objective / gradient / per-bin Hessian kernel, prior
// mat-vec, backtracking line search, active-trial compaction, and the host-side Newton driver.
// Reference: funs/inference.py:12-65 (objective, gradient, Hessian) and :67-185 (per-trial loop).
// The reference materialises C_big (qT x NT); here everything is per-bin (SURVEY.md §8a identities):
//   h[n,t] = sum_k C[n,k] x[k,t] + d[n]            lam = exp(h)
//   f      = sum lam - y h + 0.5 x^T Kinv x
//   g[k,t] = sum_n C[n,k] (lam - y)[n,t] + (Kinv_k x_k)[t]
//   W[k,l,t] = sum_n C[n,k] C[n,l] lam[n,t]        (H = blkdiag(Kinv) + scatter(W))
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <set>
#include <utility>
#include "common.cuh"
#include "pgpfa_internal.h"

using namespace pgpfa;

namespace {

// out[trial,k,s] = sum_t Kmat[k,s,t] v[trial,k,t]: per latent a (T x T) x (T x trials) product, on the FP64 tensor
// pipe (DMMA.8x8x4).  CTA = 64 rows (s) x 32 trials of one latent, 4 warps as 2x2 of 32x16; t in chunks of 16
// through double-buffered padded shared memory (strides 20 / 36 doubles keep the 16 lanes of a half-warp on distinct
// 8-byte banks), the next chunk's global loads are issued before the MMAs of the current one.  The matrices stay
// in L2 (q T^2 doubles); the vectors are gathered through the active list.
#define PA_TS 64
#define PA_TN 32
// With out2 the grid covers 2 q matrices: matrix m >= q is applied to the SAME vectors (latent m - q) and lands in out2
// (the CG loop's z = M^-1 r and K^-1 M^-1 r in one launch).
__global__ void __launch_bounds__(128) prior_apply_kernel(const double *__restrict__ Kmat, const double *__restrict__ v,
                                                          double *__restrict__ out, const int *act, int nslots, int q,
                                                          int T, const int *__restrict__ cnt, double *__restrict__ out2) {
    __shared__ double As[2][PA_TS][20];
    __shared__ double Bs[2][16][36];
    __shared__ int trial[PA_TN];
    const int m = blockIdx.x, k = m < q ? m : m - q, s0 = blockIdx.y * PA_TS, n0 = blockIdx.z * PA_TN;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    const int fr = lane >> 2, fk = lane & 3;
    if (m >= q) out = out2;
    if (cnt) {                                   // device-resident slot count (grid sized by a host upper bound)
        nslots = min(nslots, *cnt);
        if (n0 >= nslots) return;
    }
    if (tid < PA_TN) {
        const int slot = n0 + tid;
        trial[tid] = slot < nslots ? (act ? act[slot] : slot) : -1;
    }
    __syncthreads();
    const double *Kk = Kmat + (size_t)m * T * T;
    // this thread's share of a chunk: 8 elements of A (row ar + 8 j, column ac), 4 of B (trial bn + 8 j, bin bc)
    const int ar = tid >> 4, ac = tid & 15, bn = tid >> 4, bc = tid & 15;
    const double *bsrc[4];
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        const int tr = trial[bn + 8 * jj];
        bsrc[jj] = tr >= 0 ? v + ((size_t)tr * q + k) * T : nullptr;
    }
    double ra[8], rb[4];
    auto fetch = [&](int t0) {
        const bool tc = t0 + ac < T;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const int s = s0 + ar + 8 * jj;
            ra[jj] = (tc && s < T) ? Kk[(size_t)s * T + t0 + ac] : 0.0;
        }
#pragma unroll
        for (int jj = 0; jj < 4; jj++) rb[jj] = (bsrc[jj] && t0 + bc < T) ? bsrc[jj][t0 + bc] : 0.0;
    };
    double acc[4][2][2] = {};
    const int mblocks = min(4, (T - s0 - wm * 32 + 7) / 8);      // 8-row blocks of this warp that hold rows < T
    fetch(0);
    int buf = 0;
    for (int t0 = 0; t0 < T; t0 += 16, buf ^= 1) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) As[buf][ar + 8 * jj][ac] = ra[jj];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) Bs[buf][bc][bn + 8 * jj] = rb[jj];
        __syncthreads();
        if (t0 + 16 < T) fetch(t0 + 16);
#pragma unroll
        for (int k4 = 0; k4 < 16; k4 += 4) {
            double b0 = Bs[buf][k4 + fk][wn * 16 + fr], b1 = Bs[buf][k4 + fk][wn * 16 + 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (i < mblocks) {
                    const double a = As[buf][wm * 32 + i * 8 + fr][k4 + fk];
                    dmma884(acc[i][0][0], acc[i][0][1], a, b0);
                    dmma884(acc[i][1][0], acc[i][1][1], a, b1);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jn = 0; jn < 2; jn++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int s = s0 + wm * 32 + i * 8 + fr;
                const int tr = trial[wn * 16 + jn * 8 + 2 * fk + e];
                if (s < T && tr >= 0) out[((size_t)tr * q + k) * T + s] = acc[i][jn][e];
            }
}

// fused rates + objective + gradient + per-bin Hessian blocks; one CTA per trial, thread <-> bin.
// The kernel is bound by FP64 latency, not by its pipe (ncu r02c: pipe 30 % busy at one 8-warp CTA per SM, 166
// registers): BS = 224 threads cover T <= 224 bins with one bin per thread and leave room for TWO CTAs per SM
// (2 x 224 x 146 registers); HAS_OFF drops the variational path's offsets from the Laplace instantiation.
template <int Q, bool HAS_OFF, int BS, int MINB>
__global__ void __launch_bounds__(BS, MINB) laplace_eval_kernel(const double *__restrict__ x, const double *__restrict__ Kx,
                                                           const double *__restrict__ y, const double *__restrict__ C,
                                                           const double *__restrict__ d, const double *__restrict__ off,
                                                           const int *act, int N, int T,
                                                           double *__restrict__ f, double *__restrict__ g,
                                                           double *__restrict__ W, LooMap loo,
                                                           const int *__restrict__ cnt) {
    extern __shared__ double sm[];
    double *Cs = sm;             // N*Q
    double *ds = sm + N * Q;     // N
    __shared__ double red[32];
    if (cnt && (int)blockIdx.x >= *cnt) return;
    const int trial = act ? act[blockIdx.x] : blockIdx.x;
    for (int i = threadIdx.x; i < N * Q; i += blockDim.x) Cs[i] = C[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) ds[i] = d[i];
    __syncthreads();
    double fl = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double xk[Q], gk[Q], w[Q * (Q + 1) / 2];
#pragma unroll
        for (int k = 0; k < Q; k++) { xk[k] = x[((size_t)trial * Q + k) * T + t]; gk[k] = 0.0; }
#pragma unroll
        for (int i = 0; i < Q * (Q + 1) / 2; i++) w[i] = 0.0;
        // leave-one-neuron-out problems share the counts of their trial and skip one neuron
        const int yrow = loo.ymap ? loo.ymap[trial] : trial;
        const int skip = loo.excl ? loo.excl[trial] : -1;
        const double *yp = y + (size_t)yrow * N * T + t;
        const double *op = HAS_OFF ? off + (size_t)trial * N * T + t : nullptr;
        // neurons in groups of four: the four counts (and log-rate offsets) of the NEXT group are in flight while the
        // current one is computed; the row of C of a neuron is read once from shared memory into registers (every
        // thread reads the same address: broadcast)
        double yn[4], on[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            yn[u] = (u < N) ? yp[(size_t)u * T] : 0.0;
            on[u] = (HAS_OFF && u < N) ? op[(size_t)u * T] : 0.0;
        }
        for (int n0 = 0; n0 < N; n0 += 4) {
            double yv[4], ov[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                yv[u] = yn[u]; ov[u] = on[u];
                const int n = n0 + 4 + u;
                yn[u] = (n < N) ? yp[(size_t)n * T] : 0.0;
                on[u] = (HAS_OFF && n < N) ? op[(size_t)n * T] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int n = n0 + u;
                if (n >= N || n == skip) continue;
                double c[Q];
#pragma unroll
                for (int k = 0; k < Q; k++) c[k] = Cs[n * Q + k];
                double h = ds[n];
#pragma unroll
                for (int k = 0; k < Q; k++) h = fma(c[k], xk[k], h);
                // variational path: rates carry the extra log-offset s[n,t] = 0.5 c_n^T Sigma_tt c_n
                const double lam = exp(HAS_OFF ? h + ov[u] : h);
                fl += fma(-yv[u], h, lam);
                const double r = lam - yv[u];
                int idx = 0;
#pragma unroll
                for (int k = 0; k < Q; k++) {
                    gk[k] = fma(c[k], r, gk[k]);
                    const double cl = c[k] * lam;
#pragma unroll
                    for (int l = k; l < Q; l++) { w[idx] = fma(cl, c[l], w[idx]); idx++; }
                }
            }
        }
        int idx = 0;
#pragma unroll
        for (int k = 0; k < Q; k++) {
            const double kx = Kx[((size_t)trial * Q + k) * T + t];
            fl += 0.5 * xk[k] * kx;
            g[((size_t)trial * Q + k) * T + t] = gk[k] + kx;
#pragma unroll
            for (int l = k; l < Q; l++) {
                W[((size_t)trial * Q * Q + k * Q + l) * T + t] = w[idx];
                if (l != k) W[((size_t)trial * Q * Q + l * Q + k) * T + t] = w[idx];
                idx++;
            }
        }
    }
    fl = block_sum(fl, red);
    if (threadIdx.x == 0) f[trial] = fl;
}

// Armijo backtracking along the Newton direction; updates x in place and sets convergence flags.
template <int Q>
__global__ void __launch_bounds__(256) laplace_linesearch_kernel(
    double *__restrict__ x, const double *__restrict__ dx, const double *__restrict__ Kx, const double *__restrict__ Kd,
    const double *__restrict__ g, const double *__restrict__ y, const double *__restrict__ C,
    const double *__restrict__ d, const double *__restrict__ off, const int *act, int N, int T, double tol,
    double *__restrict__ fcur, int *__restrict__ conv, int *__restrict__ niter, double *__restrict__ steplen,
    int step_kind, LooMap loo, double *__restrict__ pcg_s, const int *__restrict__ cnt) {
    extern __shared__ double sm[];
    double *Cs = sm;
    double *ds = sm + N * Q;
    __shared__ double red[32];
    if (cnt && (int)blockIdx.x >= *cnt) return;
    const int trial = act ? act[blockIdx.x] : blockIdx.x;
    for (int i = threadIdx.x; i < N * Q; i += blockDim.x) Cs[i] = C[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) ds[i] = d[i];
    __syncthreads();
    const size_t base = (size_t)trial * Q * T;
    double slope = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0, dmax = 0.0;
    for (int i = threadIdx.x; i < Q * T; i += blockDim.x) {
        const double xv = x[base + i], dv = dx[base + i], kx = Kx[base + i], kd = Kd[base + i];
        slope += g[base + i] * dv;
        a0 += xv * kx;
        a1 += dv * kx;
        a2 += dv * kd;
        dmax = fmax(dmax, fabs(dv));
    }
    slope = block_sum(slope, red);
    a0 = block_sum(a0, red);
    a1 = block_sum(a1, red);
    a2 = block_sum(a2, red);
    dmax = block_max(dmax, red);
    const double f0 = fcur[trial];
    const bool tiny = fabs(slope) <= 1e-9 * (1.0 + fabs(f0));
    double alpha = 1.0, fnew = f0;
    for (int ls = 0; ls < 40; ls++) {
        double fl = 0.0;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            double xk[Q], dk[Q];
#pragma unroll
            for (int k = 0; k < Q; k++) { xk[k] = x[base + (size_t)k * T + t]; dk[k] = dx[base + (size_t)k * T + t]; }
            const int yrow = loo.ymap ? loo.ymap[trial] : trial;
            const int skip = loo.excl ? loo.excl[trial] : -1;
            const double *yp = y + (size_t)yrow * N * T + t;
            const double *op = off ? off + (size_t)trial * N * T + t : nullptr;
            double xa[Q];                                  // the trial point x + alpha dx of this bin
#pragma unroll
            for (int k = 0; k < Q; k++) xa[k] = fma(alpha, dk[k], xk[k]);
            for (int n0 = 0; n0 < N; n0 += 4) {            // four neurons' counts in flight (see laplace_eval_kernel)
                double yv[4], ov[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int n = n0 + u;
                    yv[u] = (n < N) ? yp[(size_t)n * T] : 0.0;
                    ov[u] = (op && n < N) ? op[(size_t)n * T] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int n = n0 + u;
                    if (n >= N || n == skip) continue;
                    double ha = ds[n];
#pragma unroll
                    for (int k = 0; k < Q; k++) ha = fma(Cs[n * Q + k], xa[k], ha);
                    fl += fma(-yv[u], ha, exp(ha + ov[u]));
                }
            }
        }
        fl = block_sum(fl, red);
        fnew = fl + 0.5 * a0 + alpha * a1 + 0.5 * alpha * alpha * a2;
        const bool ok = isfinite(fnew) && (tiny || fnew <= f0 + 1e-4 * alpha * slope + 1e-13 * (1.0 + fabs(f0)));
        if (ok) break;
        alpha *= 0.5;
    }
    double xmax = 0.0;
    for (int i = threadIdx.x; i < Q * T; i += blockDim.x) {
        const double xn = x[base + i] + alpha * dx[base + i];
        x[base + i] = xn;
        xmax = fmax(xmax, fabs(xn));
    }
    xmax = block_max(xmax, red);
    if (threadIdx.x == 0) {
        fcur[trial] = fnew;
        const double sl = alpha * dmax;
        // step_kind encodes the kind of step that was just taken:
        //   -1-k      exact Newton step k (fresh factor at this x).  Quadratic convergence: with
        //             c ~ sl_k / sl_{k-1}^2 the error after this step is ~ sl_k^3 / sl_{k-1}^2.
        //   1000+k    chord sweep k with the factor computed earlier in this call (at a point ~1e-2 from the mode):
        //             contraction ~ |H(x_f)^-1 (H(x_f) - H(x*))| ~ 1e-2 per sweep, replaces re-factorising.
        //   2000+k    inexact Newton step k (direction from PCG at relative residual eta).
        // states: 0 keep going with the same kind of step, 1 converged, 2 needs a fresh factorisation.
        int state;
        const double scale = 1.0 + xmax;
        const double prev = steplen[trial];
        if (step_kind < 0) {
            state = (sl <= tol * scale) ? 1 : 0;
            if (state == 0 && step_kind <= -2 && alpha == 1.0 && prev > 0.0 && sl < 0.1 * prev &&
                sl * sl * sl / (prev * prev) <= 0.1 * tol * scale)
                state = 1;
        } else if (step_kind >= 2000) {
            // error after this step ~ c sl^2 + 2 eta sl.  The next solve is asked for
            // eta' = 0.03 * (relative step just taken): superlinear overall.
            const double eta = pcg_s[trial * 4 + 2];
            const double rel = sl / scale;
            state = (0.2 * rel * rel + 2.0 * eta * rel <= 0.1 * tol || rel <= 0.01 * tol) ? 1 : 0;
            if (state == 0 && (alpha < 0.01 || step_kind % 100000 >= 2000 + 14)) state = 2;   // struggling: hand over to exact Newton
            // Forcing term of the next solve.  Default 0.03 * (relative step just taken): superlinear overall.  When the
            // NEXT step can be the last one (its quadratic term alone is below the aim of 0.002 tol, 50x under the exit
            // threshold), eta is set to what makes it the last one under the exit test's own model
            // (0.2 e^2 + 2 eta e = aim, e = the predicted size of that step) — tighter than the default for trials that
            // would otherwise miss the threshold by a little and need a whole extra Newton iteration, looser when the
            // step is already tiny (a 1e-8 step does not need a 1e-7 solve).  Measured: the model overestimates the
            // next step 2-5x; the point at which the covariances are taken (before the polishing step) stays ~1e-10 accurate.
            const double e_pred = 0.2 * rel * rel + 2.0 * eta * rel;
            const double aim = 0.002 * tol;
            double eta_next = 0.03 * rel;
            if (step_kind < 100000 && 0.2 * e_pred * e_pred < 0.5 * aim) eta_next = (aim - 0.2 * e_pred * e_pred) / (2.0 * e_pred);
            pcg_s[trial * 4 + 2] = fmin(1e-2, fmax(1e-9, eta_next));
        } else {
            const double rho = (prev > 0.0) ? sl / prev : 1.0;
            state = 0;
            if (rho < 0.5 && sl * rho / (1.0 - rho) <= 0.1 * tol * scale) state = 1;   // remaining error certified
            else if (sl <= 0.01 * tol * scale) state = 1;                               // at the rounding floor
            else if (rho > 0.3 || alpha < 1.0) state = 2;                               // contracts too slowly
        }
        steplen[trial] = sl;
        conv[trial] = state;
        niter[trial] += 1;
    }
}

// ordered compaction of the not-yet-converged trials (single CTA).  The input length is min(n_in, *n_in_dev) when a
// device-resident count is given (the host only knows an upper bound); the output count goes to *n_out and, when
// `prog` is given, to that word of mapped pinned host memory as well, where the host reads it WITHOUT synchronising
// the stream (it only throttles how far ahead of the device it enqueues, see laplace_solve_impl).
__global__ void __launch_bounds__(1024) compact_active_kernel(const int *__restrict__ act_in, int n_in,
                                                              const int *__restrict__ n_in_dev,
                                                              const int *__restrict__ conv, int keep_mask,
                                                              int *__restrict__ act_out, int *__restrict__ n_out,
                                                              volatile int *prog) {
    __shared__ int wsum[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n_in_dev) n_in = min(n_in, *n_in_dev);
    if (tid == 0) running = 0;
    __syncthreads();
    for (int b = 0; b < n_in; b += 1024) {
        const int i = b + tid;
        int trial = -1, keep = 0;
        if (i < n_in) { trial = act_in[i]; keep = (keep_mask >> conv[trial]) & 1; }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(m & ((1u << lane) - 1));
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int off = running;
        for (int w = 0; w < warp; w++) off += wsum[w];
        if (keep) act_out[off + wpre] = trial;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w = 0; w < 32; w++) tot += wsum[w]; running += tot; }
        __syncthreads();
    }
    if (tid == 0) {
        *n_out = running;
        if (prog) { *prog = running; __threadfence_system(); }
    }
}

// x <- x + dx for trials whose correction is small (it always is after convergence); records the step
__global__ void __launch_bounds__(256) polish_kernel(double *__restrict__ x, const double *__restrict__ dx,
                                                     const int *act, int n, double max_rel, double *__restrict__ steplen) {
    __shared__ double red[32];
    const int trial = act ? act[blockIdx.x] : blockIdx.x;
    double dmax = 0.0, xmax = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        dmax = fmax(dmax, fabs(dx[(size_t)trial * n + i]));
        xmax = fmax(xmax, fabs(x[(size_t)trial * n + i]));
    }
    dmax = block_max(dmax, red);
    xmax = block_max(xmax, red);
    if (dmax <= max_rel * (1.0 + xmax))
        for (int i = threadIdx.x; i < n; i += blockDim.x) x[(size_t)trial * n + i] += dx[(size_t)trial * n + i];
    if (threadIdx.x == 0) steplen[trial] = dmax;
}

// wbar[k] = mean over active trials and bins of W[trial,k,k,t], in two deterministic stages: WD_PARTS partial sums
// per latent, added in a fixed order by the consumer.
#define WD_PARTS 64
__global__ void __launch_bounds__(256) wdiag_mean_kernel(const double *__restrict__ W, const int *act, int n_act, int q, int T,
                                                         double *__restrict__ wpart) {
    __shared__ double red[32];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int a = blockIdx.y; a < n_act; a += WD_PARTS) {
        const double *Wk = W + ((size_t)act[a] * q * q + k * q + k) * T;
        for (int t = threadIdx.x; t < T; t += blockDim.x) s += Wk[t];
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) wpart[k * WD_PARTS + blockIdx.y] = s;
}

// M[k] = Kinv[k] + wbar[k] I
__global__ void shift_diag_kernel(const double *__restrict__ Kinv, const double *__restrict__ wpart, double inv_count, int T,
                                  double *__restrict__ M) {
    const int k = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * T) return;
    const int i = e / T, j = e - i * T;
    double v = Kinv[(size_t)k * T * T + e];
    if (i == j) {
        double sacc = 0.0;
        for (int pidx = 0; pidx < WD_PARTS; pidx++) sacc += wpart[k * WD_PARTS + pidx];
        v += sacc * inv_count;
    }
    M[(size_t)k * T * T + e] = v;
}

__global__ void scatter_slots_kernel(const int *act, int n, int *map) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) map[act[i]] = i;
}

__global__ void set_count_kernel(int *p, int v) { *p = v; }

// device-side generation of the tile-pair table of pgpfa_i_cov_pairs (same lexicographic order)
__global__ void gen_pairs_kernel(int2 *pairs, int q, int T, int all, int nb) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int idx = 0;
    for (int a = 0; a < nb; a++)
        for (int b = 0; b <= a; b++) {
            bool in = all != 0;
            for (int k = 0; k < q && !in; k++) {
                const int lo = (k * T) >> 6, hi = (k * T + T - 1) >> 6;
                in = (b >= lo && a <= hi);
            }
            if (in) pairs[idx++] = make_int2(a, b);
        }
}

__global__ void iota_kernel(int *p, int n, int start) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = start + i;
}

// PautoSum[k,s,t] (+)= sum_r vsmGP[r,k,s,t] + m[r,k,s] m[r,k,t]     (funs/learning.py:162-165)
__global__ void __launch_bounds__(256) pautosum_kernel(const double *__restrict__ vsmGP, const double *__restrict__ m,
                                                       int R, int q, int T, int accumulate, double *__restrict__ P) {
    const int k = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * T) return;
    const int s = e / T, t = e - s * T;
    double a0 = 0.0, a1 = 0.0;
    const size_t strideV = (size_t)q * T * T, strideM = (size_t)q * T;
    const double *vp = vsmGP + (size_t)k * T * T + e;
    const double *mp = m + (size_t)k * T;
    int r = 0;
    // eight trials' loads in flight per thread; the additions keep the even/odd order of the two-term loop below
    for (; r + 7 < R; r += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
            v[u] = vp[(size_t)(r + u) * strideV] + mp[(size_t)(r + u) * strideM + s] * mp[(size_t)(r + u) * strideM + t];
#pragma unroll
        for (int u = 0; u < 8; u += 2) { a0 += v[u]; a1 += v[u + 1]; }
    }
    for (; r + 1 < R; r += 2) {
        a0 += vp[(size_t)r * strideV] + mp[(size_t)r * strideM + s] * mp[(size_t)r * strideM + t];
        a1 += vp[(size_t)(r + 1) * strideV] + mp[(size_t)(r + 1) * strideM + s] * mp[(size_t)(r + 1) * strideM + t];
    }
    if (r < R) a0 += vp[(size_t)r * strideV] + mp[(size_t)r * strideM + s] * mp[(size_t)r * strideM + t];
    double *o = P + (size_t)k * T * T + e;
    *o = (accumulate ? *o : 0.0) + (a0 + a1);
}

// ---------------------------------------------------------------------------------------------
// Batched preconditioned conjugate gradients for the Newton systems H(x) delta = -g, H v = Kinv v + W v
// (per-bin q x q blocks); the preconditioner z = M^-1 r is applied by the caller (prior_apply with the shared
// per-latent T x T inverses).
// One CTA per trial; the per-trial scalars live in pcg_s[trial*4 + {rz, bnorm2, eta, unused}].
// ---------------------------------------------------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(256) pcg_init_kernel(const double *__restrict__ g, double *__restrict__ r,
                                                       double *__restrict__ delta, const int *act, int T, int first_outer,
                                                       double *__restrict__ pcg_s, int *__restrict__ conv,
                                                       const int *__restrict__ cnt) {
    __shared__ double red[32];
    if (cnt && (int)blockIdx.x >= *cnt) return;
    const int trial = act ? act[blockIdx.x] : blockIdx.x;
    const size_t base = (size_t)trial * Q * T;
    double bb = 0.0;
    for (int i = threadIdx.x; i < Q * T; i += blockDim.x) {
        const double gv = g[base + i];
        r[base + i] = -gv;
        delta[base + i] = 0.0;
        bb += gv * gv;
    }
    bb = block_sum(bb, red);
    if (threadIdx.x == 0) {
        pcg_s[trial * 4 + 0] = 0.0;
        pcg_s[trial * 4 + 1] = bb;
        if (first_outer) pcg_s[trial * 4 + 2] = 1e-2;
        conv[trial] = (bb == 0.0) ? 1 : 0;
    }
}

// The CG list is compacted by the LAST CTA of the step kernel to finish (ticket counter + __threadfence, the
// "threadFenceReduction" pattern) instead of by a launch of its own: one launch and ~7 us less per CG iteration.
struct CgTail {
    int *ticket;            // device counter, zero between launches; nullptr: no compaction here
    int *list_out, *cnt_out;
    int *prog;              // progress word in mapped pinned memory (or nullptr)
};

// One CG iteration per launch.  The preceding stacked prior_apply left z = M^-1 r and Nr = K^-1 M^-1 r (in Hp's buffer:
// it is read here before Hp is written, per trial, with a barrier in between).  M^-1 and K^-1 are functions of the same
// K_k, so K^-1 p follows the direction's own recurrence and is never multiplied out:
//   rz = r.z ; beta = rz / rz_old ; p = z + beta p ; Kp = Nr + beta Kp        (beta = 0 on the first iteration)
//   Hp = Kp + W p ; alpha = rz / p.Hp ; delta += alpha p ; r -= alpha Hp ; converged when |r| <= eta |b|
template <int Q>
__global__ void __launch_bounds__(256) pcg_step_kernel(double *__restrict__ p, double *__restrict__ Kp,
                                                       const double *__restrict__ W, double *__restrict__ Hp,
                                                       double *__restrict__ delta, double *__restrict__ r,
                                                       const int *act, int T, double *__restrict__ pcg_s,
                                                       int *__restrict__ conv, const int *__restrict__ cnt, CgTail tail,
                                                       const double *__restrict__ z, int first) {
    __shared__ double red[32];
    __shared__ int s_last, s_wsum[8], s_running;
    const int n_act = cnt ? *cnt : (int)gridDim.x;
    if ((int)blockIdx.x >= n_act) {
        if (blockIdx.x == 0 && tail.ticket && threadIdx.x == 0) {      // empty list: nobody takes a ticket
            *tail.cnt_out = 0;
            if (tail.prog) { *reinterpret_cast<volatile int *>(tail.prog) = 0; __threadfence_system(); }
        }
        return;
    }
    const int trial = act ? act[blockIdx.x] : blockIdx.x;
    const size_t base = (size_t)trial * Q * T;
    const double *Wt = W + (size_t)trial * Q * Q * T;
    const bool zero_rhs = pcg_s[trial * 4 + 1] == 0.0;   // zero gradient: delta = 0 is exact (pcg_init flagged it converged)
    if (zero_rhs && threadIdx.x == 0) conv[trial] = 1;
    if (!zero_rhs) {
    double rz = 0.0;
    for (int i = threadIdx.x; i < Q * T; i += blockDim.x) rz += r[base + i] * z[base + i];
    rz = block_sum(rz, red);
    const double beta = first ? 0.0 : rz / pcg_s[trial * 4 + 0];
    for (int i = threadIdx.x; i < Q * T; i += blockDim.x) {
        p[base + i] = z[base + i] + (first ? 0.0 : beta * p[base + i]);
        Kp[base + i] = Hp[base + i] + (first ? 0.0 : beta * Kp[base + i]);       // Hp's buffer holds Nr on entry
    }
    __syncthreads();
    if (threadIdx.x == 0) pcg_s[trial * 4 + 0] = rz;
    double pHp = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double pk[Q], hv[Q];
#pragma unroll
        for (int k = 0; k < Q; k++) { pk[k] = p[base + (size_t)k * T + t]; hv[k] = Kp[base + (size_t)k * T + t]; }
        // W_t is symmetric: only its upper triangle is read (36 of 64 values per bin at q = 8)
#pragma unroll
        for (int k = 0; k < Q; k++) {
            hv[k] = fma(Wt[(size_t)(k * Q + k) * T + t], pk[k], hv[k]);
#pragma unroll
            for (int l = k + 1; l < Q; l++) {
                const double wkl = Wt[(size_t)(k * Q + l) * T + t];
                hv[k] = fma(wkl, pk[l], hv[k]);
                hv[l] = fma(wkl, pk[k], hv[l]);
            }
        }
#pragma unroll
        for (int k = 0; k < Q; k++) {
            Hp[base + (size_t)k * T + t] = hv[k];
            pHp = fma(pk[k], hv[k], pHp);
        }
    }
    pHp = block_sum(pHp, red);
    const double alpha = rz / pHp;
    double rr = 0.0;
    for (int i = threadIdx.x; i < Q * T; i += blockDim.x) {
        delta[base + i] += alpha * p[base + i];
        const double rv = r[base + i] - alpha * Hp[base + i];
        r[base + i] = rv;
        rr += rv * rv;
    }
    rr = block_sum(rr, red);
    if (threadIdx.x == 0) {
        const double eta = pcg_s[trial * 4 + 2];
        conv[trial] = (rr <= eta * eta * pcg_s[trial * 4 + 1] || !(pHp > 0.0)) ? 1 : 0;
    }
    }
    if (!tail.ticket) return;
    // ---- last CTA to arrive compacts the list: trials with conv == 0 keep iterating, in list order (deterministic)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(tail.ticket, 1) == n_act - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_running = 0;
    __syncthreads();
    for (int b = 0; b < n_act; b += 256) {
        const int i = b + tid;
        int tr = -1, keep = 0;
        if (i < n_act) { tr = act ? act[i] : i; keep = (*reinterpret_cast<volatile int *>(conv + tr) == 0) ? 1 : 0; }
        const unsigned mk = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(mk & ((1u << lane) - 1));
        if (lane == 0) s_wsum[warp] = __popc(mk);
        __syncthreads();
        int off = s_running;
        for (int w2 = 0; w2 < warp; w2++) off += s_wsum[w2];
        if (keep) tail.list_out[off + wpre] = tr;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int w2 = 0; w2 < 8; w2++) tot += s_wsum[w2]; s_running += tot; }
        __syncthreads();
    }
    if (tid == 0) {
        *tail.cnt_out = s_running;
        *tail.ticket = 0;
        if (tail.prog) { *reinterpret_cast<volatile int *>(tail.prog) = s_running; __threadfence_system(); }
    }
}

template <int Q>
int launch_pcg_init(const double *g, double *r, double *delta, const int *act, int nslots, int T, int first_outer,
                    double *pcg_s, int *conv, cudaStream_t st, const int *cnt) {
    pcg_init_kernel<Q><<<nslots, 256, 0, st>>>(g, r, delta, act, T, first_outer, pcg_s, conv, cnt);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
template <int Q>
int launch_pcg_step(double *p, double *Kp, const double *W, double *Hp, double *delta, double *r, const int *act,
                    int nslots, int T, double *pcg_s, int *conv, cudaStream_t st, const int *cnt, CgTail tail, const double *z,
                    int first) {
    pcg_step_kernel<Q><<<nslots, 256, 0, st>>>(p, Kp, W, Hp, delta, r, act, T, pcg_s, conv, cnt, tail, z, first);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
#define PCG_DISPATCH(FN, ...)                                  \
    switch (q) {                                               \
        case 1: PGPFA_TRY(FN<1>(__VA_ARGS__)); break;          \
        case 2: PGPFA_TRY(FN<2>(__VA_ARGS__)); break;          \
        case 3: PGPFA_TRY(FN<3>(__VA_ARGS__)); break;          \
        case 4: PGPFA_TRY(FN<4>(__VA_ARGS__)); break;          \
        case 5: PGPFA_TRY(FN<5>(__VA_ARGS__)); break;          \
        case 6: PGPFA_TRY(FN<6>(__VA_ARGS__)); break;          \
        case 7: PGPFA_TRY(FN<7>(__VA_ARGS__)); break;          \
        case 8: PGPFA_TRY(FN<8>(__VA_ARGS__)); break;          \
        case 9: PGPFA_TRY(FN<9>(__VA_ARGS__)); break;          \
        case 10: PGPFA_TRY(FN<10>(__VA_ARGS__)); break;        \
        case 11: PGPFA_TRY(FN<11>(__VA_ARGS__)); break;        \
        case 12: PGPFA_TRY(FN<12>(__VA_ARGS__)); break;        \
        default: return PGPFA_ERR_ARG;                         \
    }

template <int Q>
int launch_eval(const double *x, const double *Kx, const double *y, const double *C, const double *d, const double *off,
                const int *act, int nslots, int N, int T, double *f, double *g, double *W, cudaStream_t st, LooMap loo,
                const int *cnt) {
    const size_t smem = (size_t)(N * Q + N) * sizeof(double);
#define EVAL_LAUNCH(OFF_, BS_, MINB_)                                                                                      \
    do {                                                                                                                   \
        if (smem > 48 * 1024)                                                                                              \
            PGPFA_CUDA_TRY(cudaFuncSetAttribute(laplace_eval_kernel<Q, OFF_, BS_, MINB_>,                                  \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                  \
        laplace_eval_kernel<Q, OFF_, BS_, MINB_><<<nslots, BS_, smem, st>>>(x, Kx, y, C, d, off, act, N, T, f, g, W, loo,  \
                                                                            cnt);                                          \
    } while (0)
    const bool small = 3 * smem <= 200 * 1024;
    if (off) { if (small) EVAL_LAUNCH(true, 128, 3); else EVAL_LAUNCH(true, 256, 1); }
    else     { if (small) EVAL_LAUNCH(false, 128, 3); else EVAL_LAUNCH(false, 256, 1); }
#undef EVAL_LAUNCH
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

template <int Q>
int launch_linesearch(double *x, const double *dx, const double *Kx, const double *Kd, const double *g, const double *y,
                      const double *C, const double *d, const double *off, const int *act, int nslots, int N, int T, double tol,
                      double *fcur, int *conv, int *niter, double *steplen, int step_kind, cudaStream_t st, LooMap loo,
                      double *pcg_s, const int *cnt) {
    const size_t smem = (size_t)(N * Q + N) * sizeof(double);
    if (smem > 48 * 1024)
        PGPFA_CUDA_TRY(cudaFuncSetAttribute(laplace_linesearch_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    laplace_linesearch_kernel<Q><<<nslots, 256, smem, st>>>(x, dx, Kx, Kd, g, y, C, d, off, act, N, T, tol, fcur, conv, niter, steplen, step_kind, loo, pcg_s, cnt);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

}  // namespace

int pgpfa_i_prior_apply(const double *Kmat, const double *v, double *out, const int *act, int nslots, int q, int T,
                        cudaStream_t st, const int *cnt, double *out2) {
    if (nslots <= 0) return PGPFA_OK;
    dim3 grid(out2 ? 2 * q : q, (T + PA_TS - 1) / PA_TS, (nslots + PA_TN - 1) / PA_TN);
    prior_apply_kernel<<<grid, 128, 0, st>>>(Kmat, v, out, act, nslots, q, T, cnt, out2);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_laplace_eval(const double *x, const double *Kx, const double *y, const double *C, const double *d,
                         const int *act, int nslots, int q, int N, int T, double *f, double *g, double *W,
                         cudaStream_t st, const double *off, LooMap loo, const int *cnt) {
    if (nslots <= 0) return PGPFA_OK;
    switch (q) {
#define CASE_Q(QQ) case QQ: return launch_eval<QQ>(x, Kx, y, C, d, off, act, nslots, N, T, f, g, W, st, loo, cnt);
        PGPFA_FOR_EACH_Q(CASE_Q)
#undef CASE_Q
    }
    return PGPFA_ERR_ARG;
}

int pgpfa_i_linesearch(double *x, const double *dx, const double *Kx, const double *Kd, const double *g,
                       const double *y, const double *C, const double *d, const int *act, int nslots, int q, int N,
                       int T, double tol, double *fcur, int *conv, int *niter, double *steplen, int step_kind,
                       cudaStream_t st, const double *off, LooMap loo, double *pcg_s, const int *cnt) {
    if (nslots <= 0) return PGPFA_OK;
    switch (q) {
#define CASE_Q(QQ) case QQ: return launch_linesearch<QQ>(x, dx, Kx, Kd, g, y, C, d, off, act, nslots, N, T, tol, fcur, conv, niter, steplen, step_kind, st, loo, pcg_s, cnt);
        PGPFA_FOR_EACH_Q(CASE_Q)
#undef CASE_Q
    }
    return PGPFA_ERR_ARG;
}

int pgpfa_i_iota(int *p, int n, int start, cudaStream_t st) {
    if (n <= 0) return PGPFA_OK;
    iota_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n, start);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

// act_out <- the entries of act_in whose state (conv[trial]) has its bit set in keep_mask; *n_out = how many
int pgpfa_i_compact(const int *act_in, int n_in, const int *conv, int keep_mask, int *act_out, int *n_out, cudaStream_t st,
                    const int *n_in_dev, int *prog) {
    compact_active_kernel<<<1, 1024, 0, st>>>(act_in, n_in, n_in_dev, conv, keep_mask, act_out, n_out, prog);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_pautosum(const double *vsmGP, const double *m, int R, int q, int T, int accumulate, double *P,
                     cudaStream_t st) {
    if (R <= 0) return PGPFA_OK;
    dim3 grid((T * T + 255) / 256, q);
    pautosum_kernel<<<grid, 256, 0, st>>>(vsmGP, m, R, q, T, accumulate, P);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_polish(double *x, const double *dx, const int *act, int n, double max_rel, double *steplen, int nslots,
                   cudaStream_t st) {
    if (nslots <= 0) return PGPFA_OK;
    polish_kernel<<<nslots, 256, 0, st>>>(x, dx, act, n, max_rel, steplen);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_num_pairs(int q, int T, bool all) { return (int)pgpfa_i_cov_pairs(q, T, all).size(); }

int pgpfa_i_gen_pairs(int2 *pairs_dev, int q, int T, bool all, cudaStream_t st) {
    gen_pairs_kernel<<<1, 32, 0, st>>>(pairs_dev, q, T, all ? 1 : 0, pgpfa_nb(q * T));
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

std::vector<int2> pgpfa_i_cov_pairs(int q, int T, bool all) {
    const int n = q * T, nb = pgpfa_nb(n);
    std::set<std::pair<int, int>> s;
    if (all) {
        for (int a = 0; a < nb; a++) for (int b = 0; b <= a; b++) s.insert({a, b});
    } else {
        for (int k = 0; k < q; k++) {
            const int lo = (k * T) >> 6, hi = (k * T + T - 1) >> 6;
            for (int a = lo; a <= hi; a++) for (int b = lo; b <= a; b++) s.insert({a, b});
        }
    }
    std::vector<int2> out;
    // heavy tiles (small a => long k-loop) first
    for (auto &p : s) out.push_back(make_int2(p.first, p.second));
    return out;
}

// ---------------------------------------------------------------------------------------------
// workspace layout of the Laplace driver
// ---------------------------------------------------------------------------------------------
namespace {
struct LapWs {
    double *Kx, *Kd, *g, *dx, *W, *fcur, *steplen, *pr, *pz, *pp, *pHp, *pcg_s;
    int *conv, *actA, *actB, *actC, *lslot, *cnt, *pinfo;
    double *pauto_partial;                  // partial tiles of the PautoSum product (lowrank.cu)
    double *Mk, *Minv, *wbar, *plogdet;     // shared CG preconditioner: (Kinv_k + wbar_k I)^-1, one T x T matrix per latent
    void *pws;
    long long pws_bytes;
    int2 *pairs;
    void *lr_tables;
    double *L, *Dinv, *ZT;
    float *L32, *D32;
    int chunk;
};
inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

size_t lap_fixed_bytes(int R, int q, int T, int npairs_max) {
    const size_t n = (size_t)q * T;
    size_t b = 0;
    b += 8 * align_up((size_t)R * n * 8) + align_up((size_t)R * 32);
    b += align_up((size_t)R * q * q * T * 8);
    b += 2 * align_up((size_t)R * 8);
    b += 5 * align_up((size_t)R * 4) + 256;
    b += align_up((size_t)q * T * T * 8) + align_up((size_t)2 * q * T * T * 8) + 2 * align_up((size_t)q * 8) + align_up((size_t)q * WD_PARTS * 8) + align_up((size_t)pgpfa_spd_inverse_workspace_bytes(q, T));
    b += align_up((size_t)npairs_max * sizeof(int2)) + align_up(PGPFA_LOWRANK_TABLE_BYTES);
    b += align_up(pgpfa_i_pautosum_partial_bytes(q, T));
    return b;
}
size_t lap_per_trial_bytes(int q, int T) {
    const int nb = pgpfa_nb(q * T);
    return (size_t)(2 * pgpfa_ltiles(nb) + nb) * PGPFA_TILE * 8 + (size_t)(pgpfa_ltiles(nb) + nb) * PGPFA_TILE * 4;
}
}  // namespace

extern "C" long long pgpfa_laplace_workspace_bytes(int R, int q, int T, int chunk) {
    if (R <= 0 || q <= 0 || T <= 0) return -1;
    if (chunk <= 0 || chunk > R) chunk = R;
    const int nb = pgpfa_nb(q * T);
    return (long long)(lap_fixed_bytes(R, q, T, (int)pgpfa_ltiles(nb)) + (size_t)chunk * lap_per_trial_bytes(q, T) + 4096);
}

static int laplace_solve_impl(pgpfa_handle_t h, const double *y, const double *C, const double *d,
                              const double *Kinv, double *x, int R, int q, int N, int T, double tol,
                              int max_newton, int flags, double *f_out, double *vsm, double *vsmGP,
                              double *cov_dense, int *niter, int *info, void *workspace, long long ws_bytes,
                              int *stats_out, cudaStream_t st, LooMap loo, bool posterior_pass,
                              const PgpfaLowRank *lr = nullptr, double *pautosum = nullptr) {
    if (!h || !y || !C || !d || !Kinv || !x || !f_out || !niter || !info || !workspace) return PGPFA_ERR_ARG;
    if (R <= 0 || q <= 0 || q > PGPFA_QMAX || N <= 0 || T <= 0 || max_newton <= 0) return PGPFA_ERR_ARG;
    const int n = q * T, nb = pgpfa_nb(n);
    const long long ltl = pgpfa_ltiles(nb);
    const size_t fixed = lap_fixed_bytes(R, q, T, (int)ltl);
    const size_t per = lap_per_trial_bytes(q, T);
    if ((size_t)ws_bytes < fixed + per) return PGPFA_ERR_WORKSPACE;
    long long chunk_ll = ((size_t)ws_bytes - fixed) / per;
    const int chunk = (int)(chunk_ll > R ? R : chunk_ll);

    unsigned char *p = static_cast<unsigned char *>(workspace);
    p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(p)));
    LapWs w;
    auto take = [&](size_t bytes) { unsigned char *r = p; p += align_up(bytes); return r; };
    const size_t vec = (size_t)R * n * 8;
    w.Kx = (double *)take(vec); w.Kd = (double *)take(vec); w.g = (double *)take(vec); w.dx = (double *)take(vec);
    w.pr = (double *)take(vec); w.pz = (double *)take(vec); w.pp = (double *)take(vec); w.pHp = (double *)take(vec);
    w.pcg_s = (double *)take((size_t)R * 32);
    w.W = (double *)take((size_t)R * q * q * T * 8);
    w.fcur = (double *)take((size_t)R * 8); w.steplen = (double *)take((size_t)R * 8);
    w.conv = (int *)take((size_t)R * 4); w.actA = (int *)take((size_t)R * 4); w.actB = (int *)take((size_t)R * 4);
    w.actC = (int *)take((size_t)R * 4); w.lslot = (int *)take((size_t)R * 4);
    w.cnt = (int *)take(256);
    w.pairs = (int2 *)take((size_t)ltl * sizeof(int2));
    w.lr_tables = take(PGPFA_LOWRANK_TABLE_BYTES);
    w.pauto_partial = (double *)take(pgpfa_i_pautosum_partial_bytes(q, T));
    w.Mk = (double *)take((size_t)q * T * T * 8); w.Minv = (double *)take((size_t)2 * q * T * T * 8);
    w.wbar = (double *)take((size_t)q * WD_PARTS * 8); w.plogdet = (double *)take((size_t)q * 8); w.pinfo = (int *)take((size_t)q * 8);
    w.pws_bytes = pgpfa_spd_inverse_workspace_bytes(q, T);
    w.pws = take((size_t)w.pws_bytes);
    w.L = (double *)take((size_t)chunk * ltl * PGPFA_TILE * 8);
    w.Dinv = (double *)take((size_t)chunk * nb * PGPFA_TILE * 8);
    w.ZT = (double *)take((size_t)chunk * ltl * PGPFA_TILE * 8);
    w.L32 = (float *)take((size_t)chunk * ltl * PGPFA_TILE * 4);
    w.D32 = (float *)take((size_t)chunk * nb * PGPFA_TILE * 4);


    PGPFA_CUDA_TRY(cudaMemsetAsync(niter, 0, (size_t)R * 4, st));
    PGPFA_CUDA_TRY(cudaMemsetAsync(info, 0, (size_t)R * 4, st));
    PGPFA_CUDA_TRY(cudaMemsetAsync(w.conv, 0, (size_t)R * 4, st));
    PGPFA_CUDA_TRY(cudaMemsetAsync(w.cnt, 0, 256, st));          // device counters, incl. the ticket of the CG compaction
    // low-rank posterior pass (lowrank.cu) when a prior factor is given, no dense covariance is wanted and its scratch
    // fits into the factor area of the workspace; otherwise the dense tiled path
    bool use_lr = lr && lr->r > 0 && posterior_pass && !cov_dense &&
                  pgpfa_i_lowrank_bytes_per_slot(q, T, lr->r) <= per;
    if (use_lr) PGPFA_TRY(pgpfa_i_lowrank_prepare(h, *lr, q, T, w.lr_tables, st));
    if (pautosum && !use_lr && !vsmGP) return PGPFA_ERR_ARG;     // the dense path sums the per-trial blocks it is given
    const int npairs = pgpfa_i_num_pairs(q, T, cov_dense != nullptr);
    bool pairs_ready = false;          // the tile-pair table of the selected inverse: generated only if the dense pass runs

    PgpfaMatSrc ms;
    ms.Kinv = Kinv; ms.W = w.W; ms.dense = nullptr; ms.q = q; ms.T = T; ms.n = n; ms.diag_scale = 1.0;
    const bool dbg = getenv("PGPFA_DEBUG") != nullptr;
    int total_factor_trials = 0, max_it_used = 0, not_converged = 0, inexact_its = 0, fallback_trials = 0, fresh_sweeps = 0, pcg_its = 0;
    const double solve_bytes = 2.0 * (double)(ltl + nb) * PGPFA_TILE * 8;
    auto read_count = [&](int &dst) -> int {
        PGPFA_CUDA_TRY(cudaMemcpyAsync(h->pinned, w.cnt + 8, sizeof(int), cudaMemcpyDeviceToHost, st));
        PGPFA_TRY(pgpfa_sync(h, st));
        dst = h->pinned[0];
        pgpfa_prof_resolve(h);
        return PGPFA_OK;
    };
    for (int c0 = 0; c0 < R; c0 += chunk) {
        const int cn = (R - c0) < chunk ? (R - c0) : chunk;
        iota_kernel<<<(cn + 255) / 256, 256, 0, st>>>(w.actA, cn, c0);
        PGPFA_LAUNCH_CHECK();
        int *act = w.actA, *act_next = w.actB;
        int n_act = cn;
        // ---- phase A: inexact Newton.  The Newton systems H(x_k) delta = -g are solved by conjugate gradients.
        // Preconditioner: M_k = Kinv_k + wbar_k I per latent (the bin- and trial-dependent diagonal W_kk replaced by
        // its mean, the cross-latent coupling dropped), ONE T x T matrix per latent shared by all trials: inverted
        // once per E-step (q small SPD inverses) and applied with the same batched mat-vec kernel as the prior.
        // cond(M^-1 H) ~ 6-10, so a Newton solve to 1e-5 takes ~15 CG iterations, each one fused H v product and
        // two batched T x T mat-vecs: no qT x qT factorisation before the one at the mode.
        if (flags & 1) {
            // Control flow: every kernel of this phase reads its slot count from device memory (`cnt` arguments) and
            // is launched on a grid sized by the host's last known upper bound, so nothing here waits for the stream.
            // The compaction kernels additionally publish their counts through the handle's progress ring (mapped
            // pinned memory); the host reads a count `depth` iterations late, only to stop enqueueing and to shrink
            // grids.  Iterations enqueued after convergence find a zero count and exit at once.  depth = 0 reproduces
            // the host-driven loop (every count is awaited before the next iteration is enqueued): the iterates are
            // bit-identical for every depth.
            const int depth = h->loop_depth;
            // PGPFA_FORCING=0 (development switch): plain 0.03 * rel forcing terms, no finish-next rule
            static const int forcing_off = [] { const char *e = getenv("PGPFA_FORCING"); return (e && atoi(e) == 0) ? 100000 : 0; }();
            int *cnt_act = w.cnt + 0, *cnt_act_next = w.cnt + 1, *cnt_cg = w.cnt + 2, *cnt_cg_next = w.cnt + 3;
            set_count_kernel<<<1, 1, 0, st>>>(cnt_act, cn);
            PGPFA_LAUNCH_CHECK();
            int ub_act = cn;
            bool newton_done = false;
            std::vector<unsigned long long> newton_seq;      // progress words of the Newton-level compactions
            std::vector<int> cg_in_newton;                   // CG iterations enqueued per Newton iteration
            std::vector<unsigned long long> cg_seq_all;
            for (int it = 0; it < 16 && !newton_done; it++) {
                if (it > 0) {
                    int nprev = -1;
                    if (depth == 0) PGPFA_TRY(pgpfa_prog_wait(h, newton_seq.back(), st, &nprev, true));
                    else if (!pgpfa_prog_peek(h, newton_seq.back(), &nprev)) nprev = -1;
                    if (nprev == 0) break;
                    if (nprev > 0) ub_act = nprev;
                }
                pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
                PGPFA_TRY(pgpfa_i_prior_apply(Kinv, x, w.Kx, act, ub_act, q, T, st, cnt_act));
                PGPFA_TRY(pgpfa_i_laplace_eval(x, w.Kx, y, C, d, act, ub_act, q, N, T, w.fcur, w.g, w.W, st, nullptr, loo, cnt_act));
                PCG_DISPATCH(launch_pcg_init, w.g, w.pr, w.dx, act, ub_act, T, it == 0, w.pcg_s, w.conv, st, cnt_act)
                pgpfa_prof_end(h, st);
                if (it == 0) {
                    pgpfa_prof_begin(h, PGPFA_PROF_BLOCKFACTOR, st);
                    dim3 gwd(q, WD_PARTS);
                    wdiag_mean_kernel<<<gwd, 256, 0, st>>>(w.W, act, cn, q, T, w.wbar);
                    PGPFA_LAUNCH_CHECK();
                    dim3 gsh((T * T + 255) / 256, q);
                    shift_diag_kernel<<<gsh, 256, 0, st>>>(Kinv, w.wbar, 1.0 / ((double)cn * T), T, w.Mk);
                    PGPFA_LAUNCH_CHECK();
                    PGPFA_TRY(pgpfa_spd_inverse_batched(w.Mk, q, T, w.Minv, w.plogdet, w.pinfo, w.pws, w.pws_bytes, st));
                    // N_k = K_k^-1 M_k^-1 behind the q inverses (both are symmetric functions of K_k, so N_k is symmetric)
                    PGPFA_TRY(pgpfa_i_small_gemm(Kinv, w.Minv, w.Minv + (size_t)q * T * T, T, q, st));
                    pgpfa_prof_end(h, st);
                }
                // PCG over the trials of this Newton iteration; converged trials drop out of `cg`
                int *cg = act_next, *cg_next = w.actC;
                PGPFA_CUDA_TRY(cudaMemcpyAsync(cg, act, (size_t)ub_act * sizeof(int), cudaMemcpyDeviceToDevice, st));
                PGPFA_CUDA_TRY(cudaMemcpyAsync(cnt_cg, cnt_act, sizeof(int), cudaMemcpyDeviceToDevice, st));
                int ub_cg = ub_act;
                std::vector<unsigned long long> cg_seq;
                for (int ci = 0; ci < 60; ci++) {
                    if (ci >= depth + (depth == 0 ? 1 : 0)) {
                        // count left after CG iteration ci - max(depth, 1); its arrival also implies that the previous
                        // Newton-level count has been written
                        int c = 0;
                        PGPFA_TRY(pgpfa_prog_wait(h, cg_seq[ci - (depth == 0 ? 1 : depth)], st, &c, depth == 0));
                        if (it > 0 && !newton_done) {
                            int nprev = 0;
                            PGPFA_TRY(pgpfa_prog_wait(h, newton_seq.back(), st, &nprev, false));
                            if (nprev == 0) { newton_done = true; break; }
                            if (nprev < ub_act) ub_act = nprev;
                        }
                        if (c == 0) break;
                        if (c < ub_cg) ub_cg = c;
                    }
                    pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
                    // z = M^-1 r -> pz and K^-1 M^-1 r -> pHp in one launch (matrices stacked in w.Minv)
                    PGPFA_TRY(pgpfa_i_prior_apply(w.Minv, w.pr, w.pz, cg, ub_cg, q, T, st, cnt_cg, w.pHp));
                    unsigned long long sq;
                    int *word;
                    PGPFA_TRY(pgpfa_prog_alloc(h, &sq, &word));
                    CgTail tail;
                    tail.ticket = w.cnt + 16; tail.list_out = cg_next; tail.cnt_out = cnt_cg_next; tail.prog = word;
                    PCG_DISPATCH(launch_pcg_step, w.pp, w.Kd, w.W, w.pHp, w.dx, w.pr, cg, ub_cg, T, w.pcg_s, w.conv, st, cnt_cg, tail,
                                 w.pz, ci == 0)
                    pgpfa_prof_end(h, st);
                    cg_seq.push_back(sq);
                    cg_seq_all.push_back(sq);
                    int *t3 = cg; cg = cg_next; cg_next = t3;
                    int *t4 = cnt_cg; cnt_cg = cnt_cg_next; cnt_cg_next = t4;
                }
                cg_in_newton.push_back((int)cg_seq.size());
                if (newton_done) break;
                pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
                PGPFA_TRY(pgpfa_i_prior_apply(Kinv, w.dx, w.Kd, act, ub_act, q, T, st, cnt_act));
                PGPFA_TRY(pgpfa_i_linesearch(x, w.dx, w.Kx, w.Kd, w.g, y, C, d, act, ub_act, q, N, T, tol, w.fcur, w.conv,
                                             niter, w.steplen, 2000 + it + forcing_off, st, nullptr, loo, w.pcg_s, cnt_act));
                pgpfa_prof_end(h, st);
                int *outp = (act == w.actA) ? w.actB : w.actA;
                unsigned long long sq;
                int *word;
                PGPFA_TRY(pgpfa_prog_alloc(h, &sq, &word));
                PGPFA_TRY(pgpfa_i_compact(act, ub_act, w.conv, 1, outp, cnt_act_next, st, cnt_act, word));
                newton_seq.push_back(sq);
                act = outp;
                act_next = (act == w.actA) ? w.actB : w.actA;
                int *t5 = cnt_act; cnt_act = cnt_act_next; cnt_act_next = t5;
            }
            // trials that struggled (state 2) or ran out of inexact-Newton iterations (state 0) go to exact Newton.
            // This count is the one value of the phase the host has to wait for (it decides whether phase B runs).
            iota_kernel<<<(cn + 255) / 256, 256, 0, st>>>(act_next, cn, c0);
            PGPFA_LAUNCH_CHECK();
            unsigned long long sq;
            int *word;
            PGPFA_TRY(pgpfa_prog_alloc(h, &sq, &word));
            PGPFA_TRY(pgpfa_i_compact(act_next, cn, w.conv, 5, act, w.cnt + 4, st, nullptr, word));
            PGPFA_TRY(pgpfa_prog_wait(h, sq, st, &n_act, true));
            pgpfa_prof_resolve(h);
            fallback_trials += n_act;
            // bookkeeping from the (now complete) progress words: Newton iterations / CG iterations that had work
            {
                int prev_n = cn;
                size_t cgpos = 0;
                for (size_t i2 = 0; i2 < cg_in_newton.size(); i2++) {
                    if (prev_n > 0) inexact_its = (int)i2 + 1;
                    int cprev = prev_n;
                    for (int c2 = 0; c2 < cg_in_newton[i2]; c2++, cgpos++) {
                        if (cprev > 0) pcg_its++;
                        int v = 0;
                        pgpfa_prog_peek(h, cg_seq_all[cgpos], &v);
                        cprev = v;
                    }
                    if (i2 < newton_seq.size()) { int v = 0; pgpfa_prog_peek(h, newton_seq[i2], &v); prev_n = v; }
                }
            }
        }
        // ---- phase B: exact Newton with fresh factorisations; every factor is re-used for a few chord sweeps
        // (4 ms each for 1024 trials) before anything is factorised again
        for (int it = 0; it < max_newton && n_act > 0; it++) {
            pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
            PGPFA_TRY(pgpfa_i_prior_apply(Kinv, x, w.Kx, act, n_act, q, T, st));
            PGPFA_TRY(pgpfa_i_laplace_eval(x, w.Kx, y, C, d, act, n_act, q, N, T, w.fcur, w.g, w.W, st, nullptr, loo));
            pgpfa_prof_end(h, st);
            pgpfa_prof_begin(h, PGPFA_PROF_FACTOR, st);
            PGPFA_TRY(pgpfa_i_factor(ms, w.L, w.Dinv, nullptr, act, info, n_act, st, h, w.L32, w.D32));
            pgpfa_prof_end(h, st);
            h->prof_work[PGPFA_PROF_FACTOR] += (double)n_act * n * (double)n * n / 3.0;
            scatter_slots_kernel<<<(n_act + 255) / 256, 256, 0, st>>>(act, n_act, w.lslot);
            PGPFA_LAUNCH_CHECK();
            pgpfa_prof_begin(h, PGPFA_PROF_SOLVE, st);
            PGPFA_TRY(pgpfa_i_solve(w.L, w.Dinv, w.g, w.dx, -1.0, act, n, n_act, st));
            pgpfa_prof_end(h, st);
            h->prof_work[PGPFA_PROF_SOLVE] += (double)n_act * solve_bytes;
            pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
            PGPFA_TRY(pgpfa_i_prior_apply(Kinv, w.dx, w.Kd, act, n_act, q, T, st));
            PGPFA_TRY(pgpfa_i_linesearch(x, w.dx, w.Kx, w.Kd, w.g, y, C, d, act, n_act, q, N, T, tol, w.fcur, w.conv,
                                         niter, w.steplen, -1 - it, st, nullptr, loo));
            pgpfa_prof_end(h, st);
            total_factor_trials += n_act;
            if (it + 1 > max_it_used) max_it_used = it + 1;
            // sweeps with the factor just computed; `act` keeps the list it was computed for
            int *swp = act_next, *swp_next = w.actC;
            PGPFA_TRY(pgpfa_i_compact(act, n_act, w.conv, 1, swp, w.cnt + 8, st));
            int n_swp = 0;
            PGPFA_TRY(read_count(n_swp));
            for (int cs = 0; cs < 8 && n_swp > 0; cs++) {
                pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
                PGPFA_TRY(pgpfa_i_prior_apply(Kinv, x, w.Kx, swp, n_swp, q, T, st));
                PGPFA_TRY(pgpfa_i_laplace_eval(x, w.Kx, y, C, d, swp, n_swp, q, N, T, w.fcur, w.g, w.W, st, nullptr, loo));
                pgpfa_prof_end(h, st);
                pgpfa_prof_begin(h, PGPFA_PROF_SOLVE, st);
                PGPFA_TRY(pgpfa_i_solve32(w.L32, w.D32, w.g, w.dx, -1.0, swp, n, n_swp, st, -1, w.lslot));
                pgpfa_prof_end(h, st);
                h->prof_work[PGPFA_PROF_SOLVE] += (double)n_swp * solve_bytes * 0.5;
                pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
                PGPFA_TRY(pgpfa_i_prior_apply(Kinv, w.dx, w.Kd, swp, n_swp, q, T, st));
                PGPFA_TRY(pgpfa_i_linesearch(x, w.dx, w.Kx, w.Kd, w.g, y, C, d, swp, n_swp, q, N, T, tol, w.fcur, w.conv,
                                             niter, w.steplen, 1000 + cs, st, nullptr, loo));
                pgpfa_prof_end(h, st);
                PGPFA_TRY(pgpfa_i_compact(swp, n_swp, w.conv, 1, swp_next, w.cnt + 8, st));
                PGPFA_TRY(read_count(n_swp));
                int *t2 = swp; swp = swp_next; swp_next = t2;
                fresh_sweeps++;
            }
            // whoever is not converged (state 0: sweeps exhausted, state 2: contraction too slow) is re-factorised
            int *outp = (act == w.actA) ? w.actB : w.actA;       // the sweep lists are dead by now
            PGPFA_TRY(pgpfa_i_compact(act, n_act, w.conv, 5, outp, w.cnt + 8, st));
            PGPFA_TRY(read_count(n_act));
            act = outp;
            act_next = (act == w.actA) ? w.actB : w.actA;
        }
        not_converged += n_act;
        // ---- posterior at the mode: objective, factor, one more (free) Newton correction, inverse slices
        iota_kernel<<<(cn + 255) / 256, 256, 0, st>>>(w.actA, cn, c0);
        PGPFA_LAUNCH_CHECK();
        pgpfa_prof_begin(h, PGPFA_PROF_EVAL, st);
        PGPFA_TRY(pgpfa_i_prior_apply(Kinv, x, w.Kx, w.actA, cn, q, T, st));
        PGPFA_TRY(pgpfa_i_laplace_eval(x, w.Kx, y, C, d, w.actA, cn, q, N, T, f_out, w.g, w.W, st, nullptr, loo));
        pgpfa_prof_end(h, st);
        if (posterior_pass && use_lr) {
            PGPFA_TRY(pgpfa_i_lowrank_posterior(h, *lr, w.W, w.g, x, w.dx, w.actA, cn, q, T, tol, w.steplen, vsm, vsmGP,
                                                w.L, (size_t)chunk * per, w.lr_tables, st, info, pautosum, c0 > 0 ? 1 : 0, x,
                                                w.pauto_partial));
            total_factor_trials += cn;
        } else if (posterior_pass) {
            pgpfa_prof_begin(h, PGPFA_PROF_FACTOR, st);
            PGPFA_TRY(pgpfa_i_factor(ms, w.L, w.Dinv, w.ZT, w.actA, info, cn, st, h, w.L32, w.D32));
            pgpfa_prof_end(h, st);
            h->prof_work[PGPFA_PROF_FACTOR] += (double)cn * n * (double)n * n / 3.0;
            total_factor_trials += cn;
            // polish: x <- x - H(x)^-1 g(x) with the factor just computed (the step is ~tol^2 for Newton-converged
            // trials and ~0.1 tol * contraction for chord-converged ones); covariance stays that of H(x) before it
            pgpfa_prof_begin(h, PGPFA_PROF_SOLVE, st);
            PGPFA_TRY(pgpfa_i_solve(w.L, w.Dinv, w.g, w.dx, -1.0, w.actA, n, cn, st));
            pgpfa_prof_end(h, st);
            h->prof_work[PGPFA_PROF_SOLVE] += (double)cn * solve_bytes;
            polish_kernel<<<cn, 256, 0, st>>>(x, w.dx, w.actA, n, 1e3 * tol, w.steplen);
            PGPFA_LAUNCH_CHECK();
            pgpfa_prof_begin(h, PGPFA_PROF_TRTRI, st);
            PGPFA_TRY(pgpfa_i_trtri(w.L, w.Dinv, w.ZT, n, cn, st, h));
            pgpfa_prof_end(h, st);
            h->prof_work[PGPFA_PROF_TRTRI] += (double)cn * n * (double)n * n / 3.0;
            pgpfa_prof_begin(h, PGPFA_PROF_SLICES, st);
            if (vsm) PGPFA_TRY(pgpfa_i_timediag(w.ZT, w.actA, vsm, n, q, T, cn, st));
            PGPFA_CUDA_TRY(cudaEventRecord(h->ev_means, st));      // pgpfa_stream_wait_means
            if ((vsmGP || cov_dense) && !pairs_ready) {
                PGPFA_TRY(pgpfa_i_gen_pairs(w.pairs, q, T, cov_dense != nullptr, st));
                pairs_ready = true;
            }
            if (vsmGP || cov_dense)
                PGPFA_TRY(pgpfa_i_lauum(w.ZT, w.pairs, npairs, w.actA, vsmGP,
                                        cov_dense ? cov_dense + (size_t)c0 * n * n : nullptr, n, q, T, cn, st));
            if (pautosum)
                PGPFA_TRY(pgpfa_i_pautosum(vsmGP + (size_t)c0 * q * T * T, x + (size_t)c0 * n, cn, q, T, c0 > 0 ? 1 : 0, pautosum, st));
            pgpfa_prof_end(h, st);
        }
    }
    if (stats_out) {
        stats_out[0] = total_factor_trials;
        stats_out[1] = max_it_used;
        stats_out[2] = not_converged;
        stats_out[3] = chunk;
        stats_out[4] = inexact_its;
        stats_out[5] = fallback_trials;
        stats_out[6] = use_lr ? lr->r : 0;       // rank of the prior factor when the low-rank posterior pass ran
        stats_out[7] = fresh_sweeps + 1000 * pcg_its;
    }
    return not_converged ? PGPFA_ERR_NOT_CONVERGED : PGPFA_OK;
}

extern "C" int pgpfa_laplace_solve(pgpfa_handle_t h, const double *y, const double *C, const double *d,
                                   const double *Kinv, double *x, int R, int q, int N, int T, double tol,
                                   int max_newton, int flags, double *f_out, double *vsm, double *vsmGP,
                                   double *cov_dense, int *niter, int *info, void *workspace, long long ws_bytes,
                                   int *stats_out, cudaStream_t st) {
    LooMap loo;
    loo.ymap = nullptr; loo.excl = nullptr;
    return laplace_solve_impl(h, y, C, d, Kinv, x, R, q, N, T, tol, max_newton, flags, f_out, vsm, vsmGP, cov_dense,
                              niter, info, workspace, ws_bytes, stats_out, st, loo, true);
}

extern "C" int pgpfa_laplace_solve_lowrank(pgpfa_handle_t h, const double *y, const double *C, const double *d,
                                           const double *Kinv, const double *F, const double *Ft, const int *rank_host,
                                           double eps, double *x, int R, int q, int N, int T, double tol, int max_newton,
                                           int flags, double *f_out, double *vsm, double *vsmGP, double *pautosum, int *niter,
                                           int *info, void *workspace, long long ws_bytes, int *stats_out, cudaStream_t st) {
    if (!F || !Ft || !rank_host || q <= 0 || q > PGPFA_QMAX || !(eps > 0.0)) return PGPFA_ERR_ARG;
    PgpfaLowRank lr;
    lr.F = F; lr.Ft = Ft; lr.eps = eps; lr.r = 0;
    for (int k = 0; k < q; k++) {
        if (rank_host[k] < 0 || rank_host[k] > T) return PGPFA_ERR_ARG;
        lr.rank[k] = rank_host[k];
        lr.off[k] = lr.r;
        lr.r += rank_host[k];
    }
    lr.off[q] = lr.r;
    LooMap loo;
    loo.ymap = nullptr; loo.excl = nullptr;
    return laplace_solve_impl(h, y, C, d, Kinv, x, R, q, N, T, tol, max_newton, flags, f_out, vsm, vsmGP, nullptr,
                              niter, info, workspace, ws_bytes, stats_out, st, loo, true, &lr, pautosum);
}

// y_pred[p][t] = exp(c_n . x_p[:,t] + d_n) for the left-out neuron n = excl[p]; err[p] = sum_t (y - y_pred)^2
template <int Q>
__global__ void __launch_bounds__(256) loo_predict_kernel(const double *__restrict__ x, const double *__restrict__ y,
                                                          const double *__restrict__ C, const double *__restrict__ d,
                                                          const int *__restrict__ ymap, const int *__restrict__ excl,
                                                          int N, int T, double *__restrict__ ypred, double *__restrict__ err) {
    __shared__ double red[32];
    const int p = blockIdx.x, n = excl[p];
    double e = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double h = d[n];
#pragma unroll
        for (int k = 0; k < Q; k++) h += C[n * Q + k] * x[((size_t)p * Q + k) * T + t];
        const double yp = exp(h);
        ypred[(size_t)p * T + t] = yp;
        const double r = y[((size_t)ymap[p] * N + n) * T + t] - yp;
        e += r * r;
    }
    e = block_sum(e, red);
    if (threadIdx.x == 0) err[p] = e;
}

// Leave-one-neuron-out prediction (funs/engine.py:599-644): problem p = (trial ymap[p], left-out neuron excl[p]).
// Finds the posterior mode without that neuron (x in/out, P x q x T, cold start = zeros) and predicts its rate.
extern "C" int pgpfa_loo_predict(pgpfa_handle_t h, const double *y, const double *C, const double *d, const double *Kinv,
                                 const int *ymap, const int *excl, double *x, int P, int q, int N, int T, double tol,
                                 int max_newton, double *ypred, double *err, int *niter, int *info, void *workspace,
                                 long long ws_bytes, int *stats_out, cudaStream_t st) {
    if (!ymap || !excl || !ypred || !err || !workspace) return PGPFA_ERR_ARG;
    LooMap loo;
    loo.ymap = ymap; loo.excl = excl;
    // f_out scratch: the first P doubles of the workspace tail are not needed afterwards -> use err as f_out
    int rc = laplace_solve_impl(h, y, C, d, Kinv, x, P, q, N, T, tol, max_newton, 1, err, nullptr, nullptr, nullptr, niter, info,
                                workspace, ws_bytes, stats_out, st, loo, false);
    if (rc != PGPFA_OK && rc != PGPFA_ERR_NOT_CONVERGED) return rc;
    switch (q) {
#define CASE_Q(QQ) case QQ: loo_predict_kernel<QQ><<<P, 256, 0, st>>>(x, y, C, d, ymap, excl, N, T, ypred, err); break;
        PGPFA_FOR_EACH_Q(CASE_Q)
#undef CASE_Q
        default: return PGPFA_ERR_ARG;
    }
    PGPFA_LAUNCH_CHECK();
    return rc;
}
