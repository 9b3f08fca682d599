// GP prior construction (squared-exponential kernel, one T x T block per latent) and the batched
// SPD inverse / log-determinant used for K^-1 (funs/util.py:599-619, funs/inference.py:82,
// funs/learning.py:183-192).  Also the C-ABI wrappers of the factorisation primitives.
#include <vector>
#include "common.cuh"
#include "pgpfa_internal.h"

using namespace pgpfa;

namespace {

// K[k,i,j] = (1-eps) exp(-0.5 ((i bs - j bs)^2 / (tau_k 1000)^2)) + eps delta_ij   — same operation
// order as funs/util.py:609-614 so that the only difference is the last ulp of exp().
__global__ void make_K_kernel(const double *__restrict__ tau, int q, int T, double bs, double eps, double *__restrict__ K) {
    const int k = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * T) return;
    const int i = e / T, j = e - i * T;
    const double dif = (double)i * bs - (double)j * bs;
    const double den = (tau[k] * 1000.0) * (tau[k] * 1000.0);
    double v = (1.0 - eps) * exp(-0.5 * ((dif * dif) / den));
    if (i == j) v += eps;
    K[(size_t)k * T * T + e] = v;
}

__global__ void make_K_big_kernel(const double *__restrict__ K, int q, int T, double *__restrict__ Kb) {
    const size_t n = (size_t)q * T;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n * n; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
        const int k = r / T, l = c / T;
        Kb[e] = (k == l) ? K[((size_t)k * T + (r - k * T)) * T + (c - l * T)] : 0.0;
    }
}

// temp = (1-eps) exp(-exp(p)/2 difSq); K = temp + eps I; dK = -0.5 temp difSq   (funs/learning.py:183-185)
__global__ void make_K_gamma_kernel(const double *__restrict__ p, int q, int T, double eps, double *__restrict__ K,
                                    double *__restrict__ dK) {
    const int k = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= T * T) return;
    const int i = e / T, j = e - i * T;
    const double difSq = (double)((i - j) * (i - j));
    const double temp = (1.0 - eps) * exp(-exp(p[k]) / 2.0 * difSq);
    K[(size_t)k * T * T + e] = temp + (i == j ? eps : 0.0);
    if (dK) dK[(size_t)k * T * T + e] = -0.5 * temp * difSq;
}

__global__ void hessian_dense_kernel(const double *__restrict__ Kinv, const double *__restrict__ W, double diag_scale,
                                     int q, int T, double *__restrict__ H) {
    const int r_ = blockIdx.y;
    const size_t n = (size_t)q * T;
    double *Hr = H + (size_t)r_ * n * n;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < n * n; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
        const int k = r / T, s = r - k * T, l = c / T, t = c - l * T;
        double v = 0.0;
        if (k == l) v = Kinv[((size_t)k * T + s) * T + t];
        if (s == t) v += W[((size_t)r_ * q * q + k * q + l) * T + t];
        if (r == c) v *= diag_scale;
        Hr[e] = v;
    }
}

// small element-wise helpers so that no arithmetic has to go through torch on the host side
__global__ void map_kernel(int op, size_t n, const double *__restrict__ x, const double *__restrict__ y, double a,
                           double *__restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v;
    if (op == 0) v = exp(x[i]);
    else if (op == 1) v = log(x[i]);
    else v = x[i] + a * y[i];
    out[i] = v;
}

// Spike times -> counts with numpy.histogram's edge rules (funs/datamanager.py:38-39: T equal bins over
// [t0, t0 + dur], right edge closed, out-of-range spikes ignored).  CSR input: spikes of (trial r, neuron n) are
// times[ptr[r*N+n] .. ptr[r*N+n+1]).
__global__ void bin_spikes_kernel(const double *__restrict__ times, const long long *__restrict__ ptr,
                                  const double *__restrict__ t0, double dur, int N, int T, double *__restrict__ Y) {
    const int cell = blockIdx.x;                    // r*N + n
    const int r = cell / N;
    const double first = t0[r], last = t0[r] + dur;
    const double step = (last - first) / (double)T;
    for (long long i = ptr[cell] + threadIdx.x; i < ptr[cell + 1]; i += blockDim.x) {
        const double x = times[i];
        if (!(x >= first && x <= last)) continue;
        int idx = (int)(((x - first) / (last - first)) * (double)T);
        if (idx == T) idx = T - 1;                  // x == last edge
        // numpy re-checks against the linspace edges (edge_k = first + k*step, last edge exact)
        const double lo = first + (double)idx * step;
        const double hi = (idx + 1 == T) ? last : first + (double)(idx + 1) * step;
        if (x < lo) idx--;
        else if (x >= hi && idx != T - 1) idx++;
        atomicAdd(&Y[(size_t)cell * T + idx], 1.0);
    }
}

struct InvWs { double *L, *Dinv, *ZT; int2 *pairs; };
inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" int pgpfa_make_K(const double *tau, int q, int T, double bs, double eps, double *K, cudaStream_t st) {
    if (!tau || !K || q <= 0 || T <= 0) return PGPFA_ERR_ARG;
    dim3 grid((T * T + 255) / 256, q);
    make_K_kernel<<<grid, 256, 0, st>>>(tau, q, T, bs, eps, K);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

extern "C" int pgpfa_make_K_big(const double *K, int q, int T, double *Kb, cudaStream_t st) {
    if (!K || !Kb || q <= 0 || T <= 0) return PGPFA_ERR_ARG;
    make_K_big_kernel<<<592, 256, 0, st>>>(K, q, T, Kb);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

extern "C" int pgpfa_make_K_gamma(const double *p, int q, int T, double eps, double *K, double *dK, cudaStream_t st) {
    if (!p || !K || q <= 0 || T <= 0) return PGPFA_ERR_ARG;
    dim3 grid((T * T + 255) / 256, q);
    make_K_gamma_kernel<<<grid, 256, 0, st>>>(p, q, T, eps, K, dK);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

extern "C" int pgpfa_hessian_dense(const double *Kinv, const double *W, double diag_scale, int R, int q, int T,
                                   double *H, cudaStream_t st) {
    if (!Kinv || !W || !H || R <= 0) return PGPFA_ERR_ARG;
    dim3 grid(148, R);
    hessian_dense_kernel<<<grid, 256, 0, st>>>(Kinv, W, diag_scale, q, T, H);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

extern "C" long long pgpfa_tiles_bytes(int n) { return n > 0 ? pgpfa_ltiles(pgpfa_nb(n)) * PGPFA_TILE * 8 : -1; }
extern "C" long long pgpfa_dinv_bytes(int n) { return n > 0 ? (long long)pgpfa_nb(n) * PGPFA_TILE * 8 : -1; }

namespace {
// ---------------------------------------------------------------------------------------------
// Register-resident SPD inverse for small matrices (n <= SW_NMAX): ONE CTA per matrix, no tile-kernel chain.
// The T x T prior blocks, the timescale search's candidate matrices and the CG preconditioner are 8-72 matrices of
// n = T ~ 200: far too few to fill the GPU, so the blocked factor / trtri / lauum chain (12+ dependent launches of
// latency-bound kernels) costs ~1 ms where the arithmetic is ~30 us of one SM's FP64 pipe.  Here the lower triangle lives
// in REGISTERS as SW_TS x SW_TS tiles (one thread per tile) and is inverted in place by the symmetric sweep operator
//     d = A_kk;  A_ij -= A_ik A_kj / d;  A_ik = A_ik / d;  A_kk = -1 / d        (k = 0 .. n-1  ->  -A^-1)
// whose pivots are those of the Cholesky / LDL^T factorisation (logdet = sum log d).  Per step the column k goes
// through shared memory (double-buffered: one __syncthreads per step) and every thread does SW_TS^2 DFMAs on registers; row
// and column k are folded into the same rank-1 form (u_k = 1 - 1/d, w_k = d - 1; A_kk = -1/d is then set by its owner) so
// there is no per-element select.  The folded entries c_i - (c_i / d)(d - 1) = c_i / d carry a relative rounding error of
// ~eps * max(1, d): nothing for the unit-diagonal prior blocks this is used on (every pivot <= 1), ~1e-13 for pivots of
// 1e3 (tests/test_sweep_algebra.py shows the effect and the Jacobi equilibration that removes it; in this kernel the
// equilibration costs more registers than it has: +13 % run time, so it is left to callers with badly scaled input).
// ---------------------------------------------------------------------------------------------
#define SW_TS 10                       // tile edge: 100 doubles of the matrix per thread
#define SW_THREADS 256                 // 2 warps per SM sub-partition -> 255 registers per thread
#define SW_NMAX 220                    // 22 * 23 / 2 = 253 tiles

__global__ void __launch_bounds__(SW_THREADS, 1) spd_sweep_kernel(const double *__restrict__ A, int n, double *__restrict__ Ainv,
                                                                  double *__restrict__ logdet, int *__restrict__ info) {
    constexpr int TS = SW_TS;
    __shared__ __align__(16) double cbuf[2][SW_NMAX + 4];      // column k, then 1 / d_k and d_k
    __shared__ double piv[SW_NMAX];
    __shared__ double red[SW_THREADS / 32];
    __shared__ int s_bad;
    const int nt = (n + TS - 1) / TS, ntiles = nt * (nt + 1) / 2, NP = nt * TS;
    const int t = threadIdx.x;
    const bool live = t < ntiles;
    int ti = 0, tj = 0;
    if (live) {
        ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= t) ti++;
        while (ti * (ti + 1) / 2 > t) ti--;
        tj = t - ti * (ti + 1) / 2;
    }
    const double *Ab = A + (size_t)blockIdx.x * n * n;
    if (t == 0) s_bad = 0;
    double a[TS][TS];
#pragma unroll
    for (int r = 0; r < TS; r++)
#pragma unroll
        for (int c = 0; c < TS; c++) {
            const int i = ti * TS + r, j = tj * TS + c;
            a[r][c] = (live && i < n && j < n) ? Ab[(size_t)i * n + j] : (i == j ? 1.0 : 0.0);   // identity padding
        }
    if (live && tj == 0) {
#pragma unroll
        for (int r = 0; r < TS; r++) cbuf[0][ti * TS + r] = a[r][0];
        if (ti == 0) { cbuf[0][NP] = 1.0 / a[0][0]; cbuf[0][NP + 1] = a[0][0]; piv[0] = a[0][0]; }
    }
    __syncthreads();
    for (int tk = 0; tk < nt; tk++) {
#pragma unroll
        for (int kk = 0; kk < TS; kk++) {
            const int k = tk * TS + kk;
            const int kn = (kk + 1) % TS;                       // in-tile index of column k + 1
            const int tkn = kk + 1 == TS ? tk + 1 : tk;         // and its tile
            const double *cc = cbuf[k & 1];
            double *cn = cbuf[(k + 1) & 1];
            if (live) {
                const double dinv = cc[NP], d = cc[NP + 1];
                double w[TS];
#pragma unroll
                for (int c = 0; c < TS; c++) w[c] = cc[tj * TS + c];
                if (tj == tk) w[kk] = d - 1.0;
                const bool rowk = ti == tk;
                // (1) the entries column k + 1 is read from go first, so that its publication (and the owner's
                //     reciprocal of the next pivot) overlaps the bulk of the update instead of preceding the barrier
#pragma unroll
                for (int r = 0; r < TS; r++) {
                    double u = cc[ti * TS + r] * dinv;
                    if (r == kk && rowk) u = 1.0 - dinv;
                    a[r][kn] = fma(-u, w[kn], a[r][kn]);
                    if (r == kn) {
#pragma unroll
                        for (int c = 0; c < TS; c++)
                            if (c != kn) a[r][c] = fma(-u, w[c], a[r][c]);
                    }
                }
                if (k + 1 < NP) {
                    if (tj == tkn) {
#pragma unroll
                        for (int r = 0; r < TS; r++) cn[ti * TS + r] = a[r][kn];
                        if (ti == tkn) {
                            const double dn = a[kn][kn];
                            cn[NP] = 1.0 / dn; cn[NP + 1] = dn; piv[k + 1] = dn;
                        }
                    } else if (ti == tkn) {
#pragma unroll
                        for (int c = 0; c < TS; c++) cn[tj * TS + c] = a[kn][c];
                    }
                }
                // (2) the rest of the tile
#pragma unroll
                for (int r = 0; r < TS; r++) {
                    if (r == kn) continue;
                    double u = cc[ti * TS + r] * dinv;
                    if (r == kk && rowk) u = 1.0 - dinv;
#pragma unroll
                    for (int c = 0; c < TS; c++)
                        if (c != kn) a[r][c] = fma(-u, w[c], a[r][c]);
                }
                if (rowk && tj == tk) a[kk][kk] = -dinv;            // exact (the folded form would cancel d - (d - 2 + 1/d))
            }
            __syncthreads();
        }
    }
    if (live) {
        double *Ob = Ainv + (size_t)blockIdx.x * n * n;
#pragma unroll
        for (int r = 0; r < TS; r++)
#pragma unroll
            for (int c = 0; c < TS; c++) {
                const int i = ti * TS + r, j = tj * TS + c;
                if (i < n && j < n && (ti != tj || r >= c)) {
                    Ob[(size_t)i * n + j] = -a[r][c];
                    Ob[(size_t)j * n + i] = -a[r][c];
                }
            }
    }
    // log-determinant and the first non-positive pivot from the stored pivots
    double ld = 0.0;
    for (int k = t; k < n; k += blockDim.x) {
        const double d = piv[k];
        if (!(d > 0.0)) atomicMax(&s_bad, n - k);          // smallest k wins
        ld += log(d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ld += __shfl_xor_sync(0xffffffffu, ld, o);
    if ((t & 31) == 0) red[t >> 5] = ld;
    __syncthreads();
    if (t == 0) {
        double tot = 0.0;
        for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) tot += red[wv];
        if (logdet) logdet[blockIdx.x] = tot;
        if (info) info[blockIdx.x] = s_bad ? n - s_bad + 1 : 0;
    }
}
}  // namespace

extern "C" long long pgpfa_spd_inverse_workspace_bytes(int batch, int n) {
    if (batch <= 0 || n <= 0) return -1;
    const int nb = pgpfa_nb(n);
    return (long long)(2 * align_up((size_t)batch * pgpfa_ltiles(nb) * PGPFA_TILE * 8) +
                       align_up((size_t)batch * nb * PGPFA_TILE * 8) + align_up(pgpfa_ltiles(nb) * sizeof(int2)) + 1024);
}

static int carve_inv_ws(void *workspace, long long ws_bytes, int batch, int n, InvWs &w) {
    if (!workspace || ws_bytes < pgpfa_spd_inverse_workspace_bytes(batch, n)) return PGPFA_ERR_WORKSPACE;
    const int nb = pgpfa_nb(n);
    unsigned char *p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<size_t>(workspace)));
    const size_t tb = align_up((size_t)batch * pgpfa_ltiles(nb) * PGPFA_TILE * 8);
    w.L = (double *)p; p += tb;
    w.ZT = (double *)p; p += tb;
    w.Dinv = (double *)p; p += align_up((size_t)batch * nb * PGPFA_TILE * 8);
    w.pairs = (int2 *)p;
    return PGPFA_OK;
}

extern "C" int pgpfa_spd_inverse_batched(const double *A, int batch, int n, double *Ainv, double *logdet, int *info,
                                         void *workspace, long long ws_bytes, cudaStream_t st) {
    if (!A || !Ainv || batch <= 0 || n <= 0) return PGPFA_ERR_ARG;
    InvWs w;
    PGPFA_TRY(carve_inv_ws(workspace, ws_bytes, batch, n, w));
    static const bool sweep_on = [] { const char *e = getenv("PGPFA_SPD_SWEEP"); return !e || atoi(e) != 0; }();
    if (sweep_on && n <= SW_NMAX) {                 // one register-resident CTA per matrix
        const int nt = (n + SW_TS - 1) / SW_TS, threads = ((nt * (nt + 1) / 2 + 31) / 32) * 32;
        spd_sweep_kernel<<<batch, threads, 0, st>>>(A, n, Ainv, logdet, info);
        PGPFA_LAUNCH_CHECK();
        return PGPFA_OK;
    }
    const int npairs = pgpfa_i_num_pairs(1, n, true);
    PGPFA_TRY(pgpfa_i_gen_pairs(w.pairs, 1, n, true, st));        // generated on the device: no host temporary, no sync
    if (info) PGPFA_CUDA_TRY(cudaMemsetAsync(info, 0, (size_t)batch * 4, st));
    PgpfaMatSrc ms;
    ms.Kinv = nullptr; ms.W = nullptr; ms.dense = A; ms.q = 1; ms.T = n; ms.n = n; ms.diag_scale = 1.0;
    PGPFA_TRY(pgpfa_i_factor(ms, w.L, w.Dinv, w.ZT, nullptr, info, batch, st));
    if (logdet) PGPFA_TRY(pgpfa_i_logdet(w.L, n, batch, logdet, st));
    PGPFA_TRY(pgpfa_i_trtri(w.L, w.Dinv, w.ZT, n, batch, st));
    PGPFA_TRY(pgpfa_i_lauum(w.ZT, w.pairs, npairs, nullptr, nullptr, Ainv, n, 1, n, batch, st));
    return PGPFA_OK;
}

extern "C" int pgpfa_potrf_dense(const double *A, int batch, int n, double *L, double *Dinv, double *ZT, int *info,
                                 cudaStream_t st) {
    if (!A || !L || !Dinv || batch <= 0 || n <= 0) return PGPFA_ERR_ARG;
    if (info) PGPFA_CUDA_TRY(cudaMemsetAsync(info, 0, (size_t)batch * 4, st));
    PgpfaMatSrc ms;
    ms.Kinv = nullptr; ms.W = nullptr; ms.dense = A; ms.q = 1; ms.T = n; ms.n = n; ms.diag_scale = 1.0;
    return pgpfa_i_factor(ms, L, Dinv, ZT, nullptr, info, batch, st);
}

extern "C" int pgpfa_potrf_posterior(const double *Kinv, const double *W, double diag_scale, int batch, int q, int T,
                                     double *L, double *Dinv, double *ZT, int *info, cudaStream_t st) {
    if (!Kinv || !W || !L || !Dinv || batch <= 0 || q <= 0 || T <= 0) return PGPFA_ERR_ARG;
    if (info) PGPFA_CUDA_TRY(cudaMemsetAsync(info, 0, (size_t)batch * 4, st));
    PgpfaMatSrc ms;
    ms.Kinv = Kinv; ms.W = W; ms.dense = nullptr; ms.q = q; ms.T = T; ms.n = q * T; ms.diag_scale = diag_scale;
    return pgpfa_i_factor(ms, L, Dinv, ZT, nullptr, info, batch, st);
}

extern "C" int pgpfa_potrs(const double *L, const double *Dinv, const double *rhs, double scale, int batch, int n,
                           double *out, cudaStream_t st) {
    if (!L || !Dinv || !rhs || !out || batch <= 0 || n <= 0) return PGPFA_ERR_ARG;
    return pgpfa_i_solve(L, Dinv, rhs, out, scale, nullptr, n, batch, st);
}

extern "C" int pgpfa_trtri(const double *L, const double *Dinv, double *ZT, int batch, int n, cudaStream_t st) {
    if (!L || !Dinv || !ZT || batch <= 0 || n <= 0) return PGPFA_ERR_ARG;
    return pgpfa_i_trtri(L, Dinv, ZT, n, batch, st);
}

extern "C" int pgpfa_potri_dense(const double *ZT, int batch, int n, double *Ainv, void *workspace, long long ws_bytes,
                                 cudaStream_t st) {
    if (!ZT || !Ainv || !workspace || batch <= 0 || n <= 0) return PGPFA_ERR_ARG;
    const int npairs = pgpfa_i_num_pairs(1, n, true);
    if ((size_t)ws_bytes < (size_t)npairs * sizeof(int2) + 256) return PGPFA_ERR_WORKSPACE;
    int2 *dp = reinterpret_cast<int2 *>(align_up(reinterpret_cast<size_t>(workspace)));
    PGPFA_TRY(pgpfa_i_gen_pairs(dp, 1, n, true, st));
    return pgpfa_i_lauum(ZT, dp, npairs, nullptr, nullptr, Ainv, n, 1, n, batch, st);
}

extern "C" int pgpfa_cov_slices(const double *ZT, int batch, int q, int T, double *vsm, double *vsmGP, void *workspace,
                                long long ws_bytes, cudaStream_t st) {
    if (!ZT || batch <= 0 || q <= 0 || q > PGPFA_QMAX || T <= 0) return PGPFA_ERR_ARG;
    const int n = q * T;
    if (vsm) PGPFA_TRY(pgpfa_i_timediag(ZT, nullptr, vsm, n, q, T, batch, st));
    if (vsmGP) {
        const int npairs = pgpfa_i_num_pairs(q, T, false);
        if (!workspace || (size_t)ws_bytes < (size_t)npairs * sizeof(int2) + 256) return PGPFA_ERR_WORKSPACE;
        int2 *dp = reinterpret_cast<int2 *>(align_up(reinterpret_cast<size_t>(workspace)));
        PGPFA_TRY(pgpfa_i_gen_pairs(dp, q, T, false, st));
        PGPFA_TRY(pgpfa_i_lauum(ZT, dp, npairs, nullptr, vsmGP, nullptr, n, q, T, batch, st));
    }
    return PGPFA_OK;
}

extern "C" int pgpfa_logdet(const double *L, int batch, int n, double *logdet, cudaStream_t st) {
    if (!L || !logdet || batch <= 0 || n <= 0) return PGPFA_ERR_ARG;
    return pgpfa_i_logdet(L, n, batch, logdet, st);
}

extern "C" int pgpfa_tiles_to_dense(const double *tiles, int batch, int n, int upper, double *out, cudaStream_t st) {
    if (!tiles || !out || batch <= 0 || n <= 0) return PGPFA_ERR_ARG;
    return pgpfa_i_tiles_to_dense(tiles, n, upper, batch, out, st);
}

extern "C" int pgpfa_prior_apply(const double *Kmat, const double *v, int R, int q, int T, double *out, cudaStream_t st) {
    if (!Kmat || !v || !out || R <= 0 || q <= 0 || T <= 0) return PGPFA_ERR_ARG;
    return pgpfa_i_prior_apply(Kmat, v, out, nullptr, R, q, T, st);
}

extern "C" int pgpfa_laplace_eval(const double *x, const double *y, const double *C, const double *d,
                                  const double *Kinv, int R, int q, int N, int T, double *f, double *g, double *W,
                                  double *Kx_ws, cudaStream_t st) {
    if (!x || !y || !C || !d || !Kinv || !f || !g || !W || !Kx_ws || R <= 0 || q <= 0 || q > PGPFA_QMAX) return PGPFA_ERR_ARG;
    PGPFA_TRY(pgpfa_i_prior_apply(Kinv, x, Kx_ws, nullptr, R, q, T, st));
    return pgpfa_i_laplace_eval(x, Kx_ws, y, C, d, nullptr, R, q, N, T, f, g, W, st);
}

extern "C" int pgpfa_pautosum(const double *vsmGP, const double *m, int R, int q, int T, int accumulate, double *P,
                              cudaStream_t st) {
    if (!vsmGP || !m || !P || R <= 0 || q <= 0 || T <= 0) return PGPFA_ERR_ARG;
    return pgpfa_i_pautosum(vsmGP, m, R, q, T, accumulate, P, st);
}

/* op 0: out = exp(x); op 1: out = log(x); op 2: out = x + a*y   (n doubles) */
extern "C" int pgpfa_map(int op, long long n, const double *x, const double *y, double a, double *out, cudaStream_t st) {
    if (n <= 0 || !x || !out || op < 0 || op > 2 || (op == 2 && !y)) return PGPFA_ERR_ARG;
    map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(op, (size_t)n, x, y, a, out);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

/* Spike times -> count matrix Y (R,N,T) with numpy.histogram semantics (funs/datamanager.py:38-39).
 * times: all spike times (s), ptr: (R*N+1) CSR offsets, t0: (R) trial start times, dur: trial duration (s). */
extern "C" int pgpfa_bin_spikes(const double *times, const long long *ptr, const double *t0, double dur, int R, int N,
                                int T, double *Y, cudaStream_t st) {
    if (!ptr || !t0 || !Y || R <= 0 || N <= 0 || T <= 0 || !(dur > 0.0)) return PGPFA_ERR_ARG;
    PGPFA_CUDA_TRY(cudaMemsetAsync(Y, 0, (size_t)R * N * T * sizeof(double), st));
    bin_spikes_kernel<<<R * N, 64, 0, st>>>(times, ptr, t0, dur, N, T, Y);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
