/* libpgpfa_b200 — C ABI of the B200-native Poisson-GPFA EM hot path.
 *
 * The reference (mackelab/poisson-gpfa) has no FFI: its boundary is the Python module API
 * (funs/engine.py, funs/inference.py, funs/learning.py, funs/util.py).  This header declares the
 * entry points a ctypes binding of that API needs; each one cites the reference code it replaces.
 * The Python mirror lives in poisson_gpfa_b200/{util,inference,learning,engine}.py and calls ONLY
 * these functions for arithmetic (no CPU fallback).
 *
 * Conventions: every array pointer is a DEVICE pointer (float64 unless stated, C-contiguous);
 * scalars by value; `stream` is a cudaStream_t passed as void*-compatible handle; every function
 * returns 0 on success or a PGPFA_ERR_* code and never throws.  The library owns no device memory:
 * callers pass workspaces sized by the *_workspace_bytes helpers (a handle owns a few KB of pinned HOST memory for
 * device-written progress words and table staging).  One handle per device/thread.
 *
 * Layouts: x[r][k][t] (latent-major, funs/inference.py:97 reshape), y[r][n][t], C[n][k], d[n],
 * K / Kinv [k][s][t], W[r][k*q+l][t], post_vsm[r][t][k][l] (= reference post_vsm[r], (T,q,q)),
 * post_vsmGP[r][k][s][t] (reference post_vsmGP[r] is (T,T,q): transpose of this), theta[n][q+1]=[C|d].
 * Factor tiles: 64x64 fragment-major tiles, packed lower (L) / packed upper (ZT = L^-T), see DESIGN.md.
 */
#ifndef PGPFA_B200_H
#define PGPFA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st *cudaStream_t;
#endif

typedef struct pgpfa_handle_s *pgpfa_handle_t;

#define PGPFA_ABI_VERSION 2

/* ---- lifecycle / errors -------------------------------------------------------------------- */
int pgpfa_abi_version(void);
int pgpfa_create(pgpfa_handle_t *out);                 /* fails with PGPFA_ERR_NO_DEVICE without a GPU */
int pgpfa_destroy(pgpfa_handle_t h);
const char *pgpfa_error_string(int code);
const char *pgpfa_last_cuda_error(void);
long long pgpfa_launch_count(void);                    /* kernels launched by this library so far */
/* Host <-> device synchronisation made by the library's drivers on this handle so far: cudaStreamSynchronize calls
 * plus blocking reads of a device-written progress word that had not arrived yet.  The iterative drivers (Newton /
 * CG of funs/inference.py:119-126, the M-step optimisers of funs/learning.py:124-130, :283-288) run from device-side
 * counters; pgpfa_set_loop_depth sets how many loop iterations they enqueue ahead of the last count the host has seen
 * (default 4; 0 = wait for every count, i.e. a host-driven loop: same iterates bit for bit, used by the tests). */
long long pgpfa_host_sync_count(pgpfa_handle_t h);
/* throttle waits so far: the host wanted a count from `depth` iterations back that the device had not produced yet;
 * unlike a synchronisation the device keeps working on the iterations queued behind it */
long long pgpfa_throttle_wait_count(pgpfa_handle_t h);
int pgpfa_set_loop_depth(pgpfa_handle_t h, int depth);
/* in-stream CUDA-event profiling of the kernel families inside the drivers (bench.py roofline):
 * slots: 0 batched Cholesky, 1 triangular solves, 2 eval/prior/line-search, 3 triangular inverse,
 * 4 covariance slices; ms_out/work_out/count_out hold 8 entries each (work = algorithmic flops or bytes) */
int pgpfa_set_profiling(pgpfa_handle_t h, int on);
int pgpfa_get_profile(pgpfa_handle_t h, double *ms_out, double *work_out, long long *count_out);

/* element-wise helper: op 0 out = exp(x), op 1 out = log(x) (lambda <-> rho of funs/inference.py:227,397),
 * op 2 out = x + a*y (scaled Newton step of funs/learning.py:890) */
int pgpfa_map(int op, long long n, const double *x, const double *y, double a, double *out, cudaStream_t stream);

/* Stevenson-style loader (funs/datamanager.py:8-54): spike times (CSR over (trial, neuron)) -> counts (R,N,T)
 * with numpy.histogram's binning rules; ptr is int64 */
int pgpfa_bin_spikes(const double *times, const long long *ptr, const double *t0, double dur, int R, int N, int T,
                     double *Y, cudaStream_t stream);

/* Device-side dataset sampling (funs/util.py:733-752; statistically, not stream-, equivalent to numpy):
 * standard normals, and Poisson counts y ~ Poisson(exp(C x + d)) */
int pgpfa_sample_normal(double *z, long long n, unsigned long long seed, cudaStream_t stream);
int pgpfa_sample_poisson(const double *x, const double *C, const double *d, int R, int q, int N, int T,
                         unsigned long long seed, double *y, cudaStream_t stream);

/* ---- (1) GP prior: funs/util.py:599-619 makeK_big, funs/inference.py:82 inv(K_big) ---------- */
int pgpfa_make_K(const double *tau_sec, int q, int T, double binSize_ms, double epsNoise, double *K, cudaStream_t stream);
int pgpfa_make_K_big(const double *K, int q, int T, double *K_big, cudaStream_t stream);
/* K(p), dK/dgamma for the timescale M-step, gamma = exp(p) in bins^-2: funs/learning.py:183-185 */
int pgpfa_make_K_gamma(const double *p, int q, int T, double epsNoise, double *K, double *dK, cudaStream_t stream);
/* batched SPD inverse + logdet (np.linalg.inv / slogdet replacement, funs/inference.py:82, funs/learning.py:186-192):
 * n <= 220 one register-resident CTA per matrix (symmetric sweep, pivots of LDL^T), larger n the tiled Cholesky chain.
 * info[b] = 0, or 1 + the index of the first non-positive pivot.  Ainv must not alias A. */
long long pgpfa_spd_inverse_workspace_bytes(int batch, int n);
int pgpfa_spd_inverse_batched(const double *A, int batch, int n, double *Ainv, double *logdet, int *info,
                              void *workspace, long long ws_bytes, cudaStream_t stream);

/* ---- factorisation primitives (exposed for tests and roofline micro-benchmarks) -------------- */
long long pgpfa_tiles_bytes(int n);        /* bytes of one packed-triangular tile set (L or ZT) */
long long pgpfa_dinv_bytes(int n);         /* bytes of one set of inverted diagonal tiles */
int pgpfa_potrf_dense(const double *A, int batch, int n, double *L_tiles, double *Dinv_tiles, double *ZT_tiles_or_null,
                      int *info, cudaStream_t stream);
/* factor H = blkdiag(Kinv_k) + scatter(W) (* diag_scale on the diagonal) without materialising it:
 * funs/inference.py:50-65 negLogPosteriorUnNorm_hess, :188-190 VIPostCov */
int pgpfa_potrf_posterior(const double *Kinv, const double *W, double diag_scale, int batch, int q, int T,
                          double *L_tiles, double *Dinv_tiles, double *ZT_tiles_or_null, int *info,
                          cudaStream_t stream);
int pgpfa_potrs(const double *L_tiles, const double *Dinv_tiles, const double *rhs, double scale, int batch, int n,
                double *out, cudaStream_t stream);
int pgpfa_trtri(const double *L_tiles, const double *Dinv_tiles, double *ZT_tiles, int batch, int n, cudaStream_t stream);
int pgpfa_potri_dense(const double *ZT_tiles, int batch, int n, double *Ainv, void *workspace, long long ws_bytes,
                      cudaStream_t stream);
int pgpfa_cov_slices(const double *ZT_tiles, int batch, int q, int T, double *vsm, double *vsmGP, void *workspace,
                     long long ws_bytes, cudaStream_t stream);
int pgpfa_logdet(const double *L_tiles, int batch, int n, double *logdet, cudaStream_t stream);
int pgpfa_tiles_to_dense(const double *tiles, int batch, int n, int upper, double *out, cudaStream_t stream);

/* ---- (2) Laplace E-step: funs/inference.py:12-185 -------------------------------------------- */
/* out[r][k][s] = sum_t Kmat[k][s][t] v[r][k][t]: one T x T matrix per latent (row-major, need not be symmetric)
 * applied to every trial (K^-1 x in funs/inference.py:16-17, and chol(K) z when sampling). */
int pgpfa_prior_apply(const double *Kmat, const double *v, int R, int q, int T, double *out, cudaStream_t stream);
/* f[r], g[r][k][t], W[r][kl][t] at given x; Kx_ws is an R*q*T scratch (receives Kinv x) */
int pgpfa_laplace_eval(const double *x, const double *y, const double *C, const double *d, const double *Kinv, int R,
                       int q, int N, int T, double *f, double *g, double *W, double *Kx_ws, cudaStream_t stream);
int pgpfa_hessian_dense(const double *Kinv, const double *W, double diag_scale, int R, int q, int T, double *H,
                        cudaStream_t stream);
long long pgpfa_laplace_workspace_bytes(int R, int q, int T, int chunk);
/* Batched Laplace E-step (funs/inference.py:67-185): posterior mode to ||step||_inf <= tol (1+||x||_inf) per trial,
 * then the posterior slices at the mode.
 * x: in = start (zeros or warm start, funs/inference.py:99-102), out = mode (post_mean).
 * flags bit 0: inexact Newton -- every Newton system is solved matrix-free by preconditioned conjugate gradients
 * (one shared T x T preconditioner per latent), so the only qT x qT factorisation is the one at the mode; trials
 * that struggle fall back per trial to exact Newton (fresh factorisation + chord sweeps), which is also what
 * flags = 0 runs for every trial.  Same fixed point either way.
 * vsm / vsmGP / cov_dense may be NULL (skipped).  stats_out[8] = {trial-factorisations, max exact-Newton
 * iterations, trials not converged, chunk size, inexact-Newton iterations, trials that fell back to exact Newton,
 * rank r of the prior factor if the low-rank posterior pass ran (else 0), chord sweeps + 1000 * CG iterations}. */
int pgpfa_laplace_solve(pgpfa_handle_t h, const double *y, const double *C, const double *d, const double *Kinv,
                        double *x, int R, int q, int N, int T, double tol, int max_newton, int flags,
                        double *f_out, double *vsm, double *vsmGP, double *cov_dense, int *niter, int *info,
                        void *workspace, long long ws_bytes, int *stats_out, cudaStream_t stream);
/* Makes `waiting_stream` wait for the point of the most recent pgpfa_laplace_solve on this handle after which x,
 * f_out, vsm, niter and info are final; the selected-inverse tiles (vsmGP / cov_dense) of the last chunk may still
 * be running on the solve's own stream.  Lets the C,d M-step (which needs only x and vsm, funs/learning.py:28-91)
 * run on a second stream underneath the tensor-bound selected inverse.  No-op before the first solve. */
int pgpfa_stream_wait_means(pgpfa_handle_t h, cudaStream_t waiting_stream);

/* Low-rank factor of the smooth part of the GP prior: K_k - eps I = F_k F_k^T by pivoted Cholesky, stopped when the
 * largest residual diagonal entry is <= delta (K from pgpfa_make_K with the same eps, funs/util.py:599-619).
 * F (q,T,T) row-major [k][t][a], Ft (q,T,T) = [k][a][t]; columns / rows >= rank[k] are zero.  rank: q ints (device). */
int pgpfa_prior_lowrank(const double *K, int q, int T, double eps, double delta, double *F, double *Ft, int *rank,
                        cudaStream_t stream);
/* pgpfa_laplace_solve with the posterior pass done through that factor: with Y = P F L_b^-T, P_t = (I + eps W_t)^-1 and
 * L_b L_b^T = I + F^T (W P) F (r x r, r = sum of the ranks), Sigma = eps P + Y Y^T exactly, so post_vsm / post_vsmGP
 * and the polishing Newton step need no qT x qT factorisation.  Same outputs as pgpfa_laplace_solve to the
 * truncation level delta (no dense covariance output).  rank_host: q ints on the HOST.  Falls back to the dense tiled
 * path when the scratch for rank r does not fit into the workspace's factor area.  stats_out[6] = r when it ran.
 * pautosum (q,T,T), optional: sum over the R trials of post_vsmGP[k] + m_k m_k^T (makePrecomp, funs/learning.py:162-165)
 * computed inside the pass as ONE symmetric product per latent over all trials; with vsmGP = NULL the per-trial T x T
 * blocks are then never written (the EM loop only consumes their sum). */
int pgpfa_laplace_solve_lowrank(pgpfa_handle_t h, const double *y, const double *C, const double *d, const double *Kinv,
                                const double *F, const double *Ft, const int *rank_host, double eps, double *x, int R,
                                int q, int N, int T, double tol, int max_newton, int flags, double *f_out, double *vsm,
                                double *vsmGP, double *pautosum, int *niter, int *info, void *workspace,
                                long long ws_bytes, int *stats_out, cudaStream_t stream);

/* Leave-one-neuron-out prediction, funs/engine.py:599-644: problem p = (trial ymap[p], left-out neuron excl[p]);
 * the posterior mode is found without that neuron (x: P x q x T, in = start, out = mode) and its rate predicted:
 * ypred[p][t] = exp(c_n . x_p[:,t] + d_n), err[p] = sum_t (y - ypred)^2.  Workspace: pgpfa_laplace_workspace_bytes(P,..) */
int pgpfa_loo_predict(pgpfa_handle_t h, const double *y, const double *C, const double *d, const double *Kinv,
                      const int *ymap, const int *excl, double *x, int P, int q, int N, int T, double tol, int max_newton,
                      double *ypred, double *err, int *niter, int *info, void *workspace, long long ws_bytes,
                      int *stats_out, cudaStream_t stream);

/* ---- (3) dual variational E-step: funs/inference.py:188-432 ------------------------------------ */
long long pgpfa_dualvi_workspace_bytes(int R, int q, int T, int chunk);
/* dualProblem / dualProblem_grad / VIPostMean / VIPostCov at a given lambda (R,N,T) for every trial:
 * D[r], grad (R,N,T), mean (R,q,T), vsm (R,T,q,q), optional dense covariance.  grad needs vsm. */
int pgpfa_dualvi_eval(pgpfa_handle_t h, const double *lam, const double *y, const double *C, const double *d,
                      const double *K, const double *Kinv, int R, int q, int N, int T, double *D, double *grad,
                      double *mean, double *vsm, double *cov_dense, int *info, void *workspace, long long ws_bytes,
                      cudaStream_t stream);
/* the dual optimum for every trial (stationary point of D; the reference approaches it with L-BFGS-B).
 * x (R,q,T), s (R,N,T) in/out: zeros (cold) or from pgpfa_dualvi_init_from_lambda (warm start).
 * stats_out[4] = {trial-factorisations, sweeps, trials not converged, chunk} */
int pgpfa_dualvi_solve(pgpfa_handle_t h, const double *y, const double *C, const double *d, const double *K,
                       const double *Kinv, double *x, double *s, int R, int q, int N, int T, double tol, int max_iter,
                       double *lam, double *mean, double *D, double *f_out, double *vsm, double *vsmGP,
                       double *cov_dense, int *niter, int *info, void *workspace, long long ws_bytes, int *stats_out,
                       cudaStream_t stream);
/* W[r][k*q+l][t] = sum_n C[n,k] C[n,l] lambda[r][n][t]  (scratch: R*q*T + 2R doubles) */
int pgpfa_rate_blocks(const double *lam, const double *C, int R, int q, int N, int T, double *W, double *scratch,
                      cudaStream_t stream);
int pgpfa_dualvi_init_from_lambda(const double *lam, const double *y, const double *C, const double *d, const double *K,
                                  int R, int q, int N, int T, double *x, double *s, void *workspace, long long ws_bytes,
                                  cudaStream_t stream);

/* ---- (4) M-step: funs/learning.py:20-309 (+ prior variants :445-534, :681-769) ---------------- */
/* PautoSum[k][s][t] (+)= sum_r vsmGP[r][k][s][t] + m[r][k][s] m[r][k][t]   funs/learning.py:162-165 */
int pgpfa_pautosum(const double *vsmGP, const double *post_mean, int R, int q, int T, int accumulate, double *P,
                   cudaStream_t stream);
int pgpfa_mstep_cd_nstats(int q);          /* 1 + (q+1) + (q+1)(q+2)/2 */
long long pgpfa_mstep_cd_workspace_bytes(int q, int N);
/* per-neuron un-normalised sums over local trials/bins of cost, gradient, Hessian of
 * MStepObservationCost at theta: stats[s][n] */
int pgpfa_mstep_cd_stats(const double *y, const double *post_mean, const double *vsm, const double *theta, int R, int q,
                         int N, int T, double *stats, void *workspace, long long ws_bytes, cudaStream_t stream);
/* one accept/reject + Newton-step update per neuron; see poisson_gpfa_b200/learning.py.
 * prior: 0.5*prior_w*|theta-theta0|^2, or (prior_mat != NULL) 0.5*(theta-theta0)^T M_n (theta-theta0) with
 * per-neuron packed-upper (q+1)x(q+1) blocks prior_mat[b][n] */
/* n_open: device int[4]; [0] receives the number of neurons still open, [1] the largest iter_index (>0) at which a
 * neuron was still open */
int pgpfa_mstep_cd_update(const double *stats, double inv_R, double prior_w, const double *prior_mat,
                          const double *theta0, double *theta_cur,
                          double *theta_try, double *fcur, double *step, double *alpha, double *slope, int *done,
                          int first, double tol, int N, int q, int *n_open, int iter_index, cudaStream_t stream);
/* pgpfa_mstep_cd_stats as one iteration of a device-driven loop: exits at once when *n_open == 0 */
int pgpfa_mstep_cd_stats_gated(const double *y, const double *post_mean, const double *vsm, const double *theta, int R,
                               int q, int N, int T, double *stats, void *workspace, long long ws_bytes,
                               const int *n_open, cudaStream_t stream);
/* learnLTparams' optimiser loop (funs/learning.py:124-130, scipy TNC/BFGS there) as n_iters per-neuron Newton
 * iterations enqueued without any host read: iteration = pgpfa_mstep_cd_stats_gated + pgpfa_mstep_cd_update on the
 * caller's state arrays (theta_cur = theta_try = theta0, done = 0, n_open = 0 before first_iter = 1).  Iterations
 * after convergence are empty launches.  Single rank; with trial sharding core.py runs the same two entry points
 * around the all-reduce of `stats`. */
int pgpfa_mstep_cd_solve(const double *y, const double *post_mean, const double *vsm, int R, int q, int N, int T,
                         double inv_R, double prior_w, const double *prior_mat, const double *theta0,
                         double *theta_cur, double *theta_try, double *fcur, double *step, double *alpha,
                         double *slope, int *done, int *n_open, double *stats, int first_iter, int n_iters,
                         double tol, void *workspace, long long ws_bytes, cudaStream_t stream);
long long pgpfa_tau_eval_workspace_bytes(int q, int T);
/* cost[k], grad[k] of MStepGPtimescaleCost(+WithPrior) at p[k]; prior_w = 1/step^2 or 0 */
int pgpfa_tau_eval(const double *p, const double *PautoSum, double numTrials, int q, int T, double epsNoise,
                   double prior_w, const double *tau_old_sec, double binSize_ms, double *cost, double *grad,
                   void *workspace, long long ws_bytes, cudaStream_t stream);

/* learnGPparams' optimiser loops (funs/learning.py:283-288; prior variant :819-825) for all latents on the device:
 * rounds of m candidate points per latent -> one batched cost/gradient evaluation -> bracket / inverse-interpolation
 * controller (csrc/tau_search.h).  first_round = 0 starts a search, a later call continues it on the same workspace.
 * No host read: flags (device int[4]) = {latents still open, evaluations used, bracketed mask, walked-out mask};
 * tau_new (q, s) and details (6,q) = {p, p0, grad, fun, fun0, grad0} are refreshed after every round.  m odd, 5..15. */
long long pgpfa_tau_solve_workspace_bytes(int q, int T, int m);
int pgpfa_mstep_tau_solve(const double *PautoSum, const double *tau_old_sec, double numTrials, int q, int T,
                          double epsNoise, double prior_w, double binSize_ms, double xtol, int m, int first_round,
                          int n_rounds, double *tau_new_sec, double *details, int *flags, void *workspace,
                          long long ws_bytes, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
