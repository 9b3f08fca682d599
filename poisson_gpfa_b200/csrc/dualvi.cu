// Dual variational E-step (funs/inference.py:188-432) on the Laplace machinery.
//
// The reference minimises the dual D(lambda) (:196-213) per trial with L-BFGS-B; every evaluation
// inverts the qT x qT precision and every gradient forms an NT x NT product.  The unique stationary
// point of D satisfies   log lambda = C m + d + s,   m = -K C_big (lambda - y),
//                        s[n,t] = 0.5 c_n^T Sigma_tt c_n,   Sigma = (P + 1e-6 diag(P))^-1,
//                        P = K^-1 + C_big diag(lambda) C_big^T           (:188-194, :215-219)
// i.e. m is the mode of a Laplace problem whose log-rates carry the offset s, and s is the posterior
// variance correction of that same Gaussian.  The driver iterates exactly that: one damped Newton
// step on m (fused rate/gradient/W kernel with offset, tiled Cholesky of the jittered precision) and
// one refresh of s from the time-diagonal blocks of the inverse, per sweep, for all trials at once.
// Function-level entry points (dual value / gradient at an arbitrary lambda) exist for parity tests.
#include <vector>
#include "common.cuh"
#include "pgpfa_internal.h"

using namespace pgpfa;

namespace {

// s[n,t] <- 0.5 c_n^T V_t c_n from the time-diagonal blocks; records max |s_new - s_old| per trial and
// folds it into the convergence flag written by the line search
template <int Q>
__global__ void __launch_bounds__(256) vi_s_update_kernel(const double *__restrict__ vsm, const double *__restrict__ C,
                                                          double *__restrict__ s, const int *act, int N, int T,
                                                          double tol, int *__restrict__ conv,
                                                          double *__restrict__ ds_out) {
    extern __shared__ double Cs[];
    __shared__ double red[32];
    const int trial = act ? act[blockIdx.x] : blockIdx.x;
    for (int i = threadIdx.x; i < N * Q; i += blockDim.x) Cs[i] = C[i];
    __syncthreads();
    double dmax = 0.0, smax = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double V[Q * Q];
#pragma unroll
        for (int i = 0; i < Q * Q; i++) V[i] = vsm[((size_t)trial * T + t) * Q * Q + i];
        double *sp = s + (size_t)trial * N * T + t;
        for (int n = 0; n < N; n++) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < Q; k++) {
                double a = 0.0;
#pragma unroll
                for (int l = 0; l < Q; l++) a += V[k * Q + l] * Cs[n * Q + l];
                acc += Cs[n * Q + k] * a;
            }
            const double sn = 0.5 * acc, so = sp[(size_t)n * T];
            dmax = fmax(dmax, fabs(sn - so));
            smax = fmax(smax, fabs(sn));
            sp[(size_t)n * T] = sn;
        }
    }
    dmax = block_max(dmax, red);
    smax = block_max(smax, red);
    if (threadIdx.x == 0) {
        if (ds_out) ds_out[trial] = dmax;
        if (conv && !(dmax <= tol * (1.0 + smax))) conv[trial] = 0;
    }
}

// From (m, s) or from a given lambda: rates, residual projections v = C^T (lambda - y), optional W, and
// the separable parts of the dual value:  sums[trial] = { sum d r, sum lambda (log lambda - 1) }.
// mode 0: lambda = exp(C m + d + s) (writes lam_out);  mode 1: lambda read from lam_in.
template <int Q>
__global__ void __launch_bounds__(256) vi_rates_kernel(const double *__restrict__ x, const double *__restrict__ s,
                                                       const double *__restrict__ lam_in, const double *__restrict__ y,
                                                       const double *__restrict__ C, const double *__restrict__ d,
                                                       const int *act, int N, int T, int mode,
                                                       double *__restrict__ lam_out, double *__restrict__ v,
                                                       double *__restrict__ W, double *__restrict__ sums) {
    extern __shared__ double sm[];
    double *Cs = sm, *ds = sm + N * Q;
    __shared__ double red[32];
    const int trial = act ? act[blockIdx.x] : blockIdx.x;
    for (int i = threadIdx.x; i < N * Q; i += blockDim.x) Cs[i] = C[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) ds[i] = d[i];
    __syncthreads();
    double sdr = 0.0, sent = 0.0;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double xk[Q], vk[Q], w[Q * (Q + 1) / 2];
#pragma unroll
        for (int k = 0; k < Q; k++) { xk[k] = (mode == 0) ? x[((size_t)trial * Q + k) * T + t] : 0.0; vk[k] = 0.0; }
#pragma unroll
        for (int i = 0; i < Q * (Q + 1) / 2; i++) w[i] = 0.0;
        const size_t base = (size_t)trial * N * T + t;
        for (int n = 0; n < N; n++) {
            double lam, loglam;
            if (mode == 0) {
                double h = ds[n] + s[base + (size_t)n * T];
#pragma unroll
                for (int k = 0; k < Q; k++) h += Cs[n * Q + k] * xk[k];
                loglam = h;
                lam = exp(h);
                if (lam_out) lam_out[base + (size_t)n * T] = lam;
            } else {
                lam = lam_in[base + (size_t)n * T];
                loglam = log(lam);
            }
            const double r = lam - y[base + (size_t)n * T];
            sdr += ds[n] * r;
            sent += lam * (loglam - 1.0);
            int idx = 0;
#pragma unroll
            for (int k = 0; k < Q; k++) {
                const double ck = Cs[n * Q + k];
                vk[k] += ck * r;
                const double cl = ck * lam;
#pragma unroll
                for (int l = k; l < Q; l++) { w[idx] += cl * Cs[n * Q + l]; idx++; }
            }
        }
        int idx = 0;
#pragma unroll
        for (int k = 0; k < Q; k++) {
            v[((size_t)trial * Q + k) * T + t] = vk[k];
#pragma unroll
            for (int l = k; l < Q; l++) {
                if (W) {
                    W[((size_t)trial * Q * Q + k * Q + l) * T + t] = w[idx];
                    if (l != k) W[((size_t)trial * Q * Q + l * Q + k) * T + t] = w[idx];
                }
                idx++;
            }
        }
    }
    sdr = block_sum(sdr, red);
    sent = block_sum(sent, red);
    if (threadIdx.x == 0) { sums[2 * trial] = sdr; sums[2 * trial + 1] = sent; }
}

// D = 0.5 v^T K v - sum d r - 0.5 logdet(P_jittered) + sum lambda (log lambda - 1);  mean = -K v
__global__ void __launch_bounds__(256) vi_dual_value_kernel(const double *__restrict__ v, const double *__restrict__ Kv,
                                                            const double *__restrict__ sums,
                                                            const double *__restrict__ logdetP, const int *act, int n,
                                                            double *__restrict__ D, double *__restrict__ mean) {
    __shared__ double red[32];
    const int slot = blockIdx.x;
    const int trial = act ? act[slot] : slot;
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double kv = Kv[(size_t)trial * n + i];
        a += v[(size_t)trial * n + i] * kv;
        if (mean) mean[(size_t)trial * n + i] = -kv;
    }
    a = block_sum(a, red);
    if (threadIdx.x == 0) D[trial] = 0.5 * a - sums[2 * trial] - 0.5 * logdetP[slot] + sums[2 * trial + 1];
}

// grad[n,t] = sum_k C[n,k] (K v)[k,t] - d_n + log lambda - 0.5 c_n^T V_t c_n      (funs/inference.py:218)
template <int Q>
__global__ void __launch_bounds__(256) vi_grad_kernel(const double *__restrict__ Kv, const double *__restrict__ vsm,
                                                      const double *__restrict__ lam, const double *__restrict__ C,
                                                      const double *__restrict__ d, int N, int T,
                                                      double *__restrict__ grad) {
    extern __shared__ double sm[];
    double *Cs = sm, *ds = sm + N * Q;
    const int trial = blockIdx.x;
    for (int i = threadIdx.x; i < N * Q; i += blockDim.x) Cs[i] = C[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) ds[i] = d[i];
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double V[Q * Q], kv[Q];
#pragma unroll
        for (int i = 0; i < Q * Q; i++) V[i] = vsm[((size_t)trial * T + t) * Q * Q + i];
#pragma unroll
        for (int k = 0; k < Q; k++) kv[k] = Kv[((size_t)trial * Q + k) * T + t];
        const size_t base = (size_t)trial * N * T + t;
        for (int n = 0; n < N; n++) {
            double lin = 0.0, quad = 0.0;
#pragma unroll
            for (int k = 0; k < Q; k++) {
                double a = 0.0;
#pragma unroll
                for (int l = 0; l < Q; l++) a += V[k * Q + l] * Cs[n * Q + l];
                quad += Cs[n * Q + k] * a;
                lin += Cs[n * Q + k] * kv[k];
            }
            grad[base + (size_t)n * T] = lin - ds[n] + log(lam[base + (size_t)n * T]) - 0.5 * quad;
        }
    }
}

// warm start: s = log lambda - (C m + d)
template <int Q>
__global__ void __launch_bounds__(256) vi_s_from_lambda_kernel(const double *__restrict__ x, const double *__restrict__ lam,
                                                               const double *__restrict__ C, const double *__restrict__ d,
                                                               int N, int T, double *__restrict__ s) {
    extern __shared__ double sm[];
    double *Cs = sm, *ds = sm + N * Q;
    const int trial = blockIdx.x;
    for (int i = threadIdx.x; i < N * Q; i += blockDim.x) Cs[i] = C[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) ds[i] = d[i];
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double xk[Q];
#pragma unroll
        for (int k = 0; k < Q; k++) xk[k] = x[((size_t)trial * Q + k) * T + t];
        const size_t base = (size_t)trial * N * T + t;
        for (int n = 0; n < N; n++) {
            double h = ds[n];
#pragma unroll
            for (int k = 0; k < Q; k++) h += Cs[n * Q + k] * xk[k];
            s[base + (size_t)n * T] = log(lam[base + (size_t)n * T]) - h;
        }
    }
}

__global__ void negate_kernel(double *p, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = -p[i];
}
inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
inline size_t smem_cd(int N, int q) { return (size_t)(N * q + N) * sizeof(double); }

#define VI_DISPATCH(KERNEL, GRID, SMEM, ...)                                                          \
    switch (q) {                                                                                      \
        case 1: KERNEL<1><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 2: KERNEL<2><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 3: KERNEL<3><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 4: KERNEL<4><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 5: KERNEL<5><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 6: KERNEL<6><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 7: KERNEL<7><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 8: KERNEL<8><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 9: KERNEL<9><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                               \
        case 10: KERNEL<10><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                             \
        case 11: KERNEL<11><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                             \
        case 12: KERNEL<12><<<GRID, 256, SMEM, st>>>(__VA_ARGS__); break;                             \
        default: return PGPFA_ERR_ARG;                                                                \
    }                                                                                                 \
    PGPFA_LAUNCH_CHECK();

struct ViWs {
    double *Kx, *Kd, *g, *dx, *W, *fcur, *steplen, *v, *Kv, *sums, *logdet, *dsmax;
    int *conv, *actA, *actB, *cnt, *niter_dummy;
    int2 *pairs;
    double *L, *Dinv, *ZT;
};
size_t vi_fixed_bytes(int R, int q, int T) {
    const size_t n = (size_t)q * T;
    const int nb = pgpfa_nb((int)n);
    size_t b = 6 * align_up((size_t)R * n * 8) + align_up((size_t)R * q * q * T * 8);
    b += 6 * align_up((size_t)R * 8 * 2) + 4 * align_up((size_t)R * 4) + 256;
    b += align_up((size_t)pgpfa_ltiles(nb) * sizeof(int2));
    return b;
}
size_t vi_per_trial_bytes(int q, int T) {
    const int nb = pgpfa_nb(q * T);
    return (size_t)(2 * pgpfa_ltiles(nb) + nb) * PGPFA_TILE * 8;
}
int vi_carve(void *workspace, long long ws_bytes, int R, int q, int T, ViWs &w, int &chunk) {
    const size_t n = (size_t)q * T;
    const int nb = pgpfa_nb((int)n);
    const long long ltl = pgpfa_ltiles(nb);
    const size_t fixed = vi_fixed_bytes(R, q, T), per = vi_per_trial_bytes(q, T);
    if (!workspace || (size_t)ws_bytes < fixed + per) return PGPFA_ERR_WORKSPACE;
    long long c = ((size_t)ws_bytes - fixed) / per;
    chunk = (int)(c > R ? R : c);
    unsigned char *p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(workspace)));
    auto take = [&](size_t bytes) { unsigned char *r = p; p += align_up(bytes); return r; };
    const size_t vec = (size_t)R * n * 8;
    w.Kx = (double *)take(vec); w.Kd = (double *)take(vec); w.g = (double *)take(vec); w.dx = (double *)take(vec);
    w.v = (double *)take(vec); w.Kv = (double *)take(vec);
    w.W = (double *)take((size_t)R * q * q * T * 8);
    w.fcur = (double *)take((size_t)R * 16); w.steplen = (double *)take((size_t)R * 16);
    w.sums = (double *)take((size_t)R * 16); w.logdet = (double *)take((size_t)R * 16);
    w.dsmax = (double *)take((size_t)R * 16); (void)take((size_t)R * 16);
    w.conv = (int *)take((size_t)R * 4); w.actA = (int *)take((size_t)R * 4); w.actB = (int *)take((size_t)R * 4);
    w.niter_dummy = (int *)take((size_t)R * 4);
    w.cnt = (int *)take(256);
    w.pairs = (int2 *)take((size_t)ltl * sizeof(int2));
    w.L = (double *)take((size_t)chunk * ltl * PGPFA_TILE * 8);
    w.Dinv = (double *)take((size_t)chunk * nb * PGPFA_TILE * 8);
    w.ZT = (double *)take((size_t)chunk * ltl * PGPFA_TILE * 8);
    return PGPFA_OK;
}

}  // namespace

extern "C" long long pgpfa_dualvi_workspace_bytes(int R, int q, int T, int chunk) {
    if (R <= 0 || q <= 0 || T <= 0) return -1;
    if (chunk <= 0 || chunk > R) chunk = R;
    return (long long)(vi_fixed_bytes(R, q, T) + (size_t)chunk * vi_per_trial_bytes(q, T) + 4096);
}

// Dual value and gradient at a given lambda for every trial (function-level parity with
// funs/inference.py:196-219; post_mean = VIPostMean, vsm from VIPostCov).  grad / mean / vsm may be NULL.
extern "C" int pgpfa_dualvi_eval(pgpfa_handle_t h, const double *lam, const double *y, const double *C, const double *d,
                                 const double *K, const double *Kinv, int R, int q, int N, int T, double *D,
                                 double *grad, double *mean, double *vsm, double *cov_dense, int *info, void *workspace,
                                 long long ws_bytes, cudaStream_t st) {
    if (!h || !lam || !y || !C || !d || !K || !Kinv || !D || R <= 0 || q <= 0 || q > PGPFA_QMAX) return PGPFA_ERR_ARG;
    ViWs w;
    int chunk = 0;
    PGPFA_TRY(vi_carve(workspace, ws_bytes, R, q, T, w, chunk));
    const int n = q * T;
    if (info) PGPFA_CUDA_TRY(cudaMemsetAsync(info, 0, (size_t)R * 4, st));
    const int npairs = pgpfa_i_num_pairs(q, T, true);
    if (cov_dense) PGPFA_TRY(pgpfa_i_gen_pairs(w.pairs, q, T, true, st));
    if ((grad || cov_dense) && !vsm) return PGPFA_ERR_ARG;   // the gradient needs the time-diagonal blocks
    VI_DISPATCH(vi_rates_kernel, R, smem_cd(N, q), nullptr, nullptr, lam, y, C, d, nullptr, N, T, 1, nullptr, w.v, w.W, w.sums)
    PGPFA_TRY(pgpfa_i_prior_apply(K, w.v, w.Kv, nullptr, R, q, T, st));
    PgpfaMatSrc ms;
    ms.Kinv = Kinv; ms.W = w.W; ms.dense = nullptr; ms.q = q; ms.T = T; ms.n = n; ms.diag_scale = 1.0 + 1e-6;
    for (int c0 = 0; c0 < R; c0 += chunk) {
        const int cn = (R - c0) < chunk ? (R - c0) : chunk;
        PGPFA_TRY(pgpfa_i_iota(w.actA, cn, c0, st));
        PGPFA_TRY(pgpfa_i_factor(ms, w.L, w.Dinv, w.ZT, w.actA, info, cn, st, h));
        PGPFA_TRY(pgpfa_i_logdet(w.L, n, cn, w.logdet, st));
        vi_dual_value_kernel<<<cn, 256, 0, st>>>(w.v, w.Kv, w.sums, w.logdet, w.actA, n, D, mean);
        PGPFA_LAUNCH_CHECK();
        if (grad || vsm || cov_dense) {
            PGPFA_TRY(pgpfa_i_trtri(w.L, w.Dinv, w.ZT, n, cn, st, h));
            if (vsm) PGPFA_TRY(pgpfa_i_timediag(w.ZT, w.actA, vsm, n, q, T, cn, st));
            if (cov_dense)
                PGPFA_TRY(pgpfa_i_lauum(w.ZT, w.pairs, npairs, w.actA, nullptr,
                                        cov_dense + (size_t)c0 * n * n, n, q, T, cn, st));
        }
    }
    if (grad) {
        VI_DISPATCH(vi_grad_kernel, R, smem_cd(N, q), w.Kv, vsm, lam, C, d, N, T, grad)
    }
    return PGPFA_OK;
}

// Fixed-point solve of the dual problem for all trials.  x (R,q,T) and s (R,N,T) are in/out (zeros for a
// cold start, or produced by pgpfa_dualvi_init_from_lambda for a warm start).
// Outputs: lam (R,N,T), mean (R,q,T) = VIPostMean(lam), D (R) dual value, f_out (R) =
// negLogPosteriorUnNorm at the mean, vsm, vsmGP, optional dense covariance.
extern "C" int pgpfa_dualvi_solve(pgpfa_handle_t h, const double *y, const double *C, const double *d, const double *K,
                                  const double *Kinv, double *x, double *s, int R, int q, int N, int T, double tol,
                                  int max_iter, double *lam, double *mean, double *D, double *f_out, double *vsm,
                                  double *vsmGP, double *cov_dense, int *niter, int *info, void *workspace,
                                  long long ws_bytes, int *stats_out, cudaStream_t st) {
    if (!h || !y || !C || !d || !K || !Kinv || !x || !s || !lam || !mean || !D || !f_out || !vsm || !niter || !info)
        return PGPFA_ERR_ARG;
    if (R <= 0 || q <= 0 || q > PGPFA_QMAX || N <= 0 || T <= 0 || max_iter <= 0) return PGPFA_ERR_ARG;
    ViWs w;
    int chunk = 0;
    PGPFA_TRY(vi_carve(workspace, ws_bytes, R, q, T, w, chunk));
    const int n = q * T;
    PGPFA_CUDA_TRY(cudaMemsetAsync(niter, 0, (size_t)R * 4, st));
    PGPFA_CUDA_TRY(cudaMemsetAsync(info, 0, (size_t)R * 4, st));
    const int npairs = pgpfa_i_num_pairs(q, T, cov_dense != nullptr);
    PGPFA_TRY(pgpfa_i_gen_pairs(w.pairs, q, T, cov_dense != nullptr, st));
    PgpfaMatSrc ms;
    ms.Kinv = Kinv; ms.W = w.W; ms.dense = nullptr; ms.q = q; ms.T = T; ms.n = n; ms.diag_scale = 1.0 + 1e-6;
    int sweeps = 0, not_converged = 0, total_factor = 0;
    for (int c0 = 0; c0 < R; c0 += chunk) {
        const int cn = (R - c0) < chunk ? (R - c0) : chunk;
        PGPFA_TRY(pgpfa_i_iota(w.actA, cn, c0, st));
        int *act = w.actA, *act_next = w.actB;
        int n_act = cn;
        for (int it = 0; it < max_iter && n_act > 0; it++) {
            PGPFA_TRY(pgpfa_i_prior_apply(Kinv, x, w.Kx, act, n_act, q, T, st));
            PGPFA_TRY(pgpfa_i_laplace_eval(x, w.Kx, y, C, d, act, n_act, q, N, T, w.fcur, w.g, w.W, st, s));
            pgpfa_prof_begin(h, PGPFA_PROF_FACTOR, st);
            PGPFA_TRY(pgpfa_i_factor(ms, w.L, w.Dinv, w.ZT, act, info, n_act, st, h));
            pgpfa_prof_end(h, st);
            h->prof_work[PGPFA_PROF_FACTOR] += (double)n_act * n * (double)n * n / 3.0;
            total_factor += n_act;
            PGPFA_TRY(pgpfa_i_solve(w.L, w.Dinv, w.g, w.dx, -1.0, act, n, n_act, st));
            PGPFA_TRY(pgpfa_i_prior_apply(Kinv, w.dx, w.Kd, act, n_act, q, T, st));
            PGPFA_TRY(pgpfa_i_linesearch(x, w.dx, w.Kx, w.Kd, w.g, y, C, d, act, n_act, q, N, T, tol, w.fcur, w.conv,
                                         niter, w.steplen, -1, st, s));
            pgpfa_prof_begin(h, PGPFA_PROF_TRTRI, st);
            PGPFA_TRY(pgpfa_i_trtri(w.L, w.Dinv, w.ZT, n, n_act, st, h));
            pgpfa_prof_end(h, st);
            h->prof_work[PGPFA_PROF_TRTRI] += (double)n_act * n * (double)n * n / 3.0;
            PGPFA_TRY(pgpfa_i_timediag(w.ZT, act, vsm, n, q, T, n_act, st));
            VI_DISPATCH(vi_s_update_kernel, n_act, (size_t)N * q * sizeof(double), vsm, C, s, act, N, T, tol, w.conv, w.dsmax)
            PGPFA_TRY(pgpfa_i_compact(act, n_act, w.conv, 1, act_next, w.cnt, st));
            PGPFA_CUDA_TRY(cudaMemcpyAsync(h->pinned, w.cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
            PGPFA_TRY(pgpfa_sync(h, st));
            n_act = h->pinned[0];
            pgpfa_prof_resolve(h);
            int *tmp = act; act = act_next; act_next = tmp;
            if (it + 1 > sweeps) sweeps = it + 1;
        }
        not_converged += n_act;
        // ---- outputs at the fixed point
        PGPFA_TRY(pgpfa_i_iota(w.actA, cn, c0, st));
        VI_DISPATCH(vi_rates_kernel, cn, smem_cd(N, q), x, s, nullptr, y, C, d, w.actA, N, T, 0, lam, w.v, w.W, w.sums)
        PGPFA_TRY(pgpfa_i_prior_apply(K, w.v, w.Kv, w.actA, cn, q, T, st));
        pgpfa_prof_begin(h, PGPFA_PROF_FACTOR, st);
        PGPFA_TRY(pgpfa_i_factor(ms, w.L, w.Dinv, w.ZT, w.actA, info, cn, st, h));
        pgpfa_prof_end(h, st);
        h->prof_work[PGPFA_PROF_FACTOR] += (double)cn * n * (double)n * n / 3.0;
        total_factor += cn;
        PGPFA_TRY(pgpfa_i_logdet(w.L, n, cn, w.logdet, st));
        vi_dual_value_kernel<<<cn, 256, 0, st>>>(w.v, w.Kv, w.sums, w.logdet, w.actA, n, D, mean);
        PGPFA_LAUNCH_CHECK();
        pgpfa_prof_begin(h, PGPFA_PROF_TRTRI, st);
        PGPFA_TRY(pgpfa_i_trtri(w.L, w.Dinv, w.ZT, n, cn, st, h));
        pgpfa_prof_end(h, st);
        h->prof_work[PGPFA_PROF_TRTRI] += (double)cn * n * (double)n * n / 3.0;
        PGPFA_TRY(pgpfa_i_timediag(w.ZT, w.actA, vsm, n, q, T, cn, st));
        if (vsmGP || cov_dense)
            PGPFA_TRY(pgpfa_i_lauum(w.ZT, w.pairs, npairs, w.actA, vsmGP,
                                    cov_dense ? cov_dense + (size_t)c0 * n * n : nullptr, n, q, T, cn, st));
        // post_lik term: the (offset-free) Laplace objective at the variational mean (funs/inference.py:333)
        PGPFA_TRY(pgpfa_i_prior_apply(Kinv, mean, w.Kx, w.actA, cn, q, T, st));
        PGPFA_TRY(pgpfa_i_laplace_eval(mean, w.Kx, y, C, d, w.actA, cn, q, N, T, f_out, w.g, w.W, st, nullptr));
    }
    if (stats_out) {
        stats_out[0] = total_factor;
        stats_out[1] = sweeps;
        stats_out[2] = not_converged;
        stats_out[3] = chunk;
    }
    return not_converged ? PGPFA_ERR_NOT_CONVERGED : PGPFA_OK;
}

// Warm start from a previous lambda (funs/inference.py:294-297): m = VIPostMean(lambda), s = log lambda - C m - d
extern "C" int pgpfa_dualvi_init_from_lambda(const double *lam, const double *y, const double *C, const double *d,
                                             const double *K, int R, int q, int N, int T, double *x, double *s,
                                             void *workspace, long long ws_bytes, cudaStream_t st) {
    if (!lam || !y || !C || !d || !K || !x || !s || !workspace || R <= 0 || q <= 0 || q > PGPFA_QMAX) return PGPFA_ERR_ARG;
    const size_t n = (size_t)q * T;
    const size_t need = 2 * align_up((size_t)R * n * 8) + align_up((size_t)R * 16) + 512;
    if ((size_t)ws_bytes < need) return PGPFA_ERR_WORKSPACE;
    unsigned char *p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(workspace)));
    double *v = (double *)p; p += align_up((size_t)R * n * 8);
    double *sums = (double *)p;
    VI_DISPATCH(vi_rates_kernel, R, smem_cd(N, q), nullptr, nullptr, lam, y, C, d, nullptr, N, T, 1, nullptr, v, nullptr, sums)
    PGPFA_TRY(pgpfa_i_prior_apply(K, v, x, nullptr, R, q, T, st));
    negate_kernel<<<(unsigned)((R * n + 255) / 256), 256, 0, st>>>(x, (size_t)R * n);
    PGPFA_LAUNCH_CHECK();
    VI_DISPATCH(vi_s_from_lambda_kernel, R, smem_cd(N, q), x, lam, C, d, N, T, s)
    return PGPFA_OK;
}

// W[r][k*q+l][t] = sum_n C[n,k] C[n,l] lambda[r,n,t]: the per-bin blocks of C_big diag(lambda) C_big^T
// (funs/inference.py:189); scratch = R*q*T + 2R doubles
extern "C" int pgpfa_rate_blocks(const double *lam, const double *C, int R, int q, int N, int T, double *W,
                                 double *scratch, cudaStream_t st) {
    if (!lam || !C || !W || !scratch || R <= 0 || q <= 0 || q > PGPFA_QMAX) return PGPFA_ERR_ARG;
    double *v = scratch, *sums = scratch + (size_t)R * q * T;
    // y and d are irrelevant for W: pass lambda as y (residual 0) and C's first column as a dummy d
    VI_DISPATCH(vi_rates_kernel, R, smem_cd(N, q), nullptr, nullptr, lam, lam, C, C, nullptr, N, T, 1, nullptr, v, W, sums)
    return PGPFA_OK;
}
