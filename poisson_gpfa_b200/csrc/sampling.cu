// Device-side sampling of synthetic Poisson-GPFA datasets (funs/util.py:733-752: X ~ N(0, K_big) per trial,
// Y ~ Poisson(exp(C X + d))).  Counter-based Philox4x32-10 generator, Box-Muller normals, Poisson by inversion
// (sequential search from the mode-free left end; exact for the rates that occur here, lambda < ~500).
// The numpy RNG stream of the reference cannot be reproduced on a GPU; parity tests keep using numpy-generated
// inputs, this path exists to build configs[4]-scale inputs in well under a second.
#include "common.cuh"
#include "pgpfa_internal.h"

namespace {

struct Philox {
    uint32_t key[2];
    __device__ __forceinline__ static uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
    __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) const {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        const uint32_t hi0 = mulhi(M0, c[0]), lo0 = M0 * c[0];
        const uint32_t hi1 = mulhi(M1, c[2]), lo1 = M1 * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ __forceinline__ void operator()(uint64_t ctr, uint32_t stream, uint32_t (&out)[4]) const {
        uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), stream, 0u};
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; r++) { round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = c[i];
    }
};

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {      // (0,1), 53 bits
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return ((double)(v >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// z[i] ~ N(0,1): two normals per Philox call
__global__ void normal_kernel(double *__restrict__ z, size_t n, uint32_t seed_lo, uint32_t seed_hi) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (2 * i >= n) return;
    Philox ph; ph.key[0] = seed_lo; ph.key[1] = seed_hi;
    uint32_t r[4];
    ph(i, 1u, r);
    const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z[2 * i] = rad * c;
    if (2 * i + 1 < n) z[2 * i + 1] = rad * s;
}

// y[r,n,t] ~ Poisson(exp(sum_k C[n,k] x[r,k,t] + d[n]))  by inversion
__global__ void poisson_kernel(const double *__restrict__ x, const double *__restrict__ C, const double *__restrict__ d,
                               int R, int q, int N, int T, uint32_t seed_lo, uint32_t seed_hi, double *__restrict__ y) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t total = (size_t)R * N * T;
    if (i >= total) return;
    const int t = (int)(i % T), n = (int)((i / T) % N), r = (int)(i / ((size_t)T * N));
    double h = d[n];
    for (int k = 0; k < q; k++) h += C[n * q + k] * x[((size_t)r * q + k) * T + t];
    const double lam = exp(h);
    Philox ph; ph.key[0] = seed_lo; ph.key[1] = seed_hi;
    uint32_t rr[4];
    ph(i, 2u, rr);
    const double u = u01(rr[0], rr[1]);
    double p = exp(-lam), cdf = p;
    int k = 0;
    if (!(p > 0.0)) {                       // lambda too large for exp(-lambda): normal approximation, rounded
        double s, c;
        const double u2 = u01(rr[2], rr[3]);
        sincospi(2.0 * u2, &s, &c);
        const double g = sqrt(-2.0 * log(u)) * c;
        y[i] = fmax(0.0, floor(lam + sqrt(lam) * g + 0.5));
        return;
    }
    while (u > cdf && k < 100000) { k++; p *= lam / (double)k; cdf += p; }
    y[i] = (double)k;
}

}  // namespace

/* z (n doubles) ~ N(0,1) */
extern "C" int pgpfa_sample_normal(double *z, long long n, unsigned long long seed, cudaStream_t st) {
    if (!z || n <= 0) return PGPFA_ERR_ARG;
    const size_t pairs = ((size_t)n + 1) / 2;
    normal_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, st>>>(z, (size_t)n, (uint32_t)seed, (uint32_t)(seed >> 32));
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}

/* y (R,N,T) ~ Poisson(exp(C x + d)) for latent trajectories x (R,q,T) */
extern "C" int pgpfa_sample_poisson(const double *x, const double *C, const double *d, int R, int q, int N, int T,
                                    unsigned long long seed, double *y, cudaStream_t st) {
    if (!x || !C || !d || !y || R <= 0 || q <= 0 || N <= 0 || T <= 0) return PGPFA_ERR_ARG;
    const size_t total = (size_t)R * N * T;
    poisson_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, C, d, R, q, N, T, (uint32_t)seed,
                                                                   (uint32_t)(seed >> 32), y);
    PGPFA_LAUNCH_CHECK();
    return PGPFA_OK;
}
