// Batched triangular solves with the tiled Cholesky factor:  out = scale * (L L^T)^-1 rhs.
// This is the Newton step of the Laplace E-step (replaces the inner CG of scipy's Newton-CG at
// funs/inference.py:119-126).  HBM-bound: each trial's factor (packed 64x64 tiles) is streamed
// twice (forward + backward), once per Newton iteration.  One CTA per trial, 8 warps; the diagonal
// blocks are applied through their explicit inverses (Dinv) produced by the factorisation.
#include "common.cuh"
#include "pgpfa_internal.h"

using namespace pgpfa;

namespace {

#define SOLVE_GROUPS 4                       // warp groups of 8 warps; group g takes tiles k = g (mod 4)
#define SOLVE_THREADS (SOLVE_GROUPS * 256)

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ double2 ld2(const float *p) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    return make_double2((double)v.x, (double)v.y);
}

template <typename TL>
__global__ void __launch_bounds__(SOLVE_THREADS) chol_solve_kernel(const TL *__restrict__ L, const TL *__restrict__ Dinv,
                                                                   const double *__restrict__ rhs, double *__restrict__ out,
                                                                   double scale, const int *act, int nb, int n, int lslot_base,
                                                                   const int *__restrict__ lslot_map) {
    extern __shared__ double sm[];
    double *z = sm;                         // nb*64
    double *tmp = sm + nb * PGPFA_NB;       // 64
    double *part = tmp + PGPFA_NB;          // SOLVE_GROUPS x 64 partial sums
    const int slot = blockIdx.x;
    const int trial = act ? act[slot] : slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int w = warp & 7, grp = warp >> 3;
    const long long ltl = (long long)nb * (nb + 1) / 2;
    // factor slot: position in the active list (fresh factorisation), trial - base (factor kept from an earlier
    // call, stored by local trial index) or lslot_map[trial] (factor computed for an earlier, longer active list)
    const int lslot = lslot_map ? lslot_map[trial] : (lslot_base >= 0 ? trial - lslot_base : slot);
    const TL *Ls = L + (size_t)lslot * ltl * PGPFA_TILE;
    const TL *Ds = Dinv + (size_t)lslot * nb * PGPFA_TILE;
    for (int i = tid; i < nb * PGPFA_NB; i += SOLVE_THREADS) z[i] = (i < n) ? rhs[(size_t)trial * n + i] : 0.0;
    __syncthreads();
    const int c2 = 2 * (lane & 3);
    // ---- forward: L y = rhs.  Block row j: warp (grp, w) sums L(j,k)[rows 8w..8w+7] z_k over its tiles k.
    for (int j = 0; j < nb; j++) {
        const TL *row = Ls + ltile(j, 0) * PGPFA_TILE + w * 64 + lane * 2;
        double acc = 0.0;
        for (int k = grp; k < j; k += SOLVE_GROUPS) {
            const TL *tp = row + (size_t)k * PGPFA_TILE;
            const double *zp = z + k * PGPFA_NB + c2;
#pragma unroll
            for (int sp = 0; sp < 8; sp++) {
                const double2 a = ld2(tp + (sp >> 2) * PGPFA_SLAB + (sp & 3) * 512);
                const double2 v = *reinterpret_cast<const double2 *>(zp + sp * 8);
                acc += a.x * v.x + a.y * v.y;
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if ((lane & 3) == 0) part[grp * PGPFA_NB + w * 8 + (lane >> 2)] = acc;
        __syncthreads();
        if (tid < PGPFA_NB) {
            double sacc = 0.0;
#pragma unroll
            for (int g = 0; g < SOLVE_GROUPS; g++) sacc += part[g * PGPFA_NB + tid];
            tmp[tid] = z[j * PGPFA_NB + tid] - sacc;
        }
        __syncthreads();
        if (grp == 0) {
            const TL *tp = Ds + (size_t)j * PGPFA_TILE + w * 64 + lane * 2;
            double a2 = 0.0;
#pragma unroll
            for (int sp = 0; sp < 8; sp++) {
                const double2 a = ld2(tp + (sp >> 2) * PGPFA_SLAB + (sp & 3) * 512);
                const double2 v = *reinterpret_cast<const double2 *>(tmp + sp * 8 + c2);
                a2 += a.x * v.x + a.y * v.y;
            }
            a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
            a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
            if ((lane & 3) == 0) z[j * PGPFA_NB + w * 8 + (lane >> 2)] = a2;
        }
        __syncthreads();
    }
    // ---- backward: L^T x = y   (right-looking over block rows, streaming each block row once); warp (grp, w)
    // owns column group w of the tiles k = grp (mod 4), so the updates of z_k never collide
    const int colbase = (w >> 2) * 32 + (w & 3) * 8;       // this warp's 8-column group inside a tile
    const int toff = (w >> 2) * PGPFA_SLAB + (w & 3) * 512 + lane * 2;
    for (int i = nb - 1; i >= 0; i--) {
        if (grp == 0) {
            const TL *tp = Ds + (size_t)i * PGPFA_TILE + toff;
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int rb = 0; rb < 8; rb++) {
                const double2 a = ld2(tp + rb * 64);
                const double v = z[i * PGPFA_NB + rb * 8 + (lane >> 2)];
                ax += a.x * v;
                ay += a.y * v;
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o);
            }
            if (lane < 4) { tmp[colbase + 2 * lane] = ax; tmp[colbase + 2 * lane + 1] = ay; }
        }
        __syncthreads();
        if (tid < PGPFA_NB) z[i * PGPFA_NB + tid] = tmp[tid];
        const TL *row = Ls + ltile(i, 0) * PGPFA_TILE + toff;
        double dv[8];
#pragma unroll
        for (int rb = 0; rb < 8; rb++) dv[rb] = tmp[rb * 8 + (lane >> 2)];
        for (int k = grp; k < i; k += SOLVE_GROUPS) {
            const TL *tp = row + (size_t)k * PGPFA_TILE;
            double ax = 0.0, ay = 0.0;
#pragma unroll
            for (int rb = 0; rb < 8; rb++) {
                const double2 a = ld2(tp + rb * 64);
                ax += a.x * dv[rb];
                ay += a.y * dv[rb];
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o);
            }
            if (lane < 4) {
                z[k * PGPFA_NB + colbase + 2 * lane] -= ax;
                z[k * PGPFA_NB + colbase + 2 * lane + 1] -= ay;
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += SOLVE_THREADS) out[(size_t)trial * n + i] = scale * z[i];
}

}  // namespace

template <typename TL>
static int launch_solve(const TL *L, const TL *Dinv, const double *rhs, double *out, double scale, const int *act, int n,
                        int nslots, cudaStream_t st, int lslot_base, const int *lslot_map) {
    if (nslots <= 0) return PGPFA_OK;
    const int nb = pgpfa_nb(n);
    const size_t smem = (size_t)(nb * PGPFA_NB + PGPFA_NB + SOLVE_GROUPS * PGPFA_NB) * sizeof(double);
    if (smem > 48 * 1024)
        PGPFA_CUDA_TRY(cudaFuncSetAttribute(chol_solve_kernel<TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chol_solve_kernel<TL><<<nslots, SOLVE_THREADS, smem, st>>>(L, Dinv, rhs, out, scale, act, nb, n, lslot_base, lslot_map);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

int pgpfa_i_solve(const double *L, const double *Dinv, const double *rhs, double *out, double scale, const int *act,
                  int n, int nslots, cudaStream_t st, int lslot_base, const int *lslot_map) {
    return launch_solve<double>(L, Dinv, rhs, out, scale, act, n, nslots, st, lslot_base, lslot_map);
}

int pgpfa_i_solve32(const float *L32, const float *D32, const double *rhs, double *out, double scale, const int *act,
                    int n, int nslots, cudaStream_t st, int lslot_base, const int *lslot_map) {
    return launch_solve<float>(L32, D32, rhs, out, scale, act, n, nslots, st, lslot_base, lslot_map);
}
