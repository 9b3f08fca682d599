// Shared device-side building blocks for libpgpfa_b200 (sm_100a only).
//
// Tile storage ("fragment-major"): every qT x qT posterior system is kept as 64x64
// tiles of doubles (32 KB, contiguous).  Inside a tile the element (r, c) lives at
//     (c>>5)*2048 + ((c>>3)&3)*512 + (r>>3)*64 + (r&7)*8 + (c&7)
// i.e. [k-slab 2][k-group 4][row-block 8][8x8 row-major block].  A 64x32 k-slab (16 KB) is one
// contiguous run, so a single cp.async.bulk (UBLKCP) stages it into shared memory, and a warp's
// DMMA.8x8x4 operand fragment for two consecutive k4-steps is one conflict-free 16-byte LDS per lane
// (lane = (r&7)*4 + ((c&7)>>1) holds columns c, c+1; the k index inside an 8-group is permuted
// identically for both operands, which a dot product does not see).  The accumulator fragment of
// DMMA.8x8x4 (row = lane/4, cols = 2*(lane%4)+{0,1}) maps onto the same 16-byte slot, so results are
// stored in operand layout with vector stores and no shuffle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PGPFA_NB 64
#define PGPFA_TILE 4096          // doubles per tile
#define PGPFA_SLAB 2048          // doubles per 64x32 k-slab
#ifndef PGPFA_STAGES
#define PGPFA_STAGES 2
#endif
#define PGPFA_GEMM_THREADS 128
#define PGPFA_GEMM_CTAS_PER_SM 3
// dynamic shared memory of the tile-GEMM kernels: the staging area (PGPFA_STAGES x (A slab + B slab), re-used
// by the 64x64 diagonal factorisation which needs 74,248 bytes) followed by the mbarriers; 3 CTAs per SM
#define PGPFA_BAR_OFF 74304
#define PGPFA_GEMM_SMEM (PGPFA_BAR_OFF + 64)

enum {
    PGPFA_OK = 0,
    PGPFA_ERR_CUDA = 1,
    PGPFA_ERR_ARG = 2,
    PGPFA_ERR_WORKSPACE = 3,
    PGPFA_ERR_NOT_SPD = 4,
    PGPFA_ERR_NOT_CONVERGED = 5,
    PGPFA_ERR_NO_DEVICE = 6
};

#define PGPFA_CUDA_TRY(expr)                                  \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) {                              \
            pgpfa_set_last_cuda_error(_e, __FILE__, __LINE__); \
            return PGPFA_ERR_CUDA;                            \
        }                                                     \
    } while (0)
#define PGPFA_LAUNCH_CHECK()                 \
    do {                                     \
        g_pgpfa_launches++;                  \
        PGPFA_CUDA_TRY(cudaGetLastError());  \
    } while (0)
#define PGPFA_TRY(expr)          \
    do {                         \
        int _r = (expr);         \
        if (_r != PGPFA_OK) return _r; \
    } while (0)

void pgpfa_set_last_cuda_error(cudaError_t e, const char *file, int line);
extern unsigned long long g_pgpfa_launches;   // kernels launched by this library (bench.py gpu_launches)

static inline int pgpfa_nb(int n) { return (n + PGPFA_NB - 1) / PGPFA_NB; }
static inline long long pgpfa_ltiles(int nb) { return (long long)nb * (nb + 1) / 2; }

namespace pgpfa {

__host__ __device__ __forceinline__ int tile_off(int r, int c) {
    return ((c >> 5) << 11) + (((c >> 3) & 3) << 9) + ((r >> 3) << 6) + ((r & 7) << 3) + (c & 7);
}
// packed lower-triangular tile index (row i >= col j); a block row is contiguous
__host__ __device__ __forceinline__ long long ltile(int i, int j) { return (long long)i * (i + 1) / 2 + j; }
// packed upper-triangular tile index (row j <= col k); a block row is contiguous
__host__ __device__ __forceinline__ long long utile(int j, int k, int nb) {
    return (long long)j * nb - (long long)j * (j - 1) / 2 + (k - j);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// Tile-GEMM core: acc(64x64, 4 warps as 2x2 of 32x32) += sum over k-slabs  A_slab * B_slab^T
// A and B are contiguous streams of 64x32 slabs in fragment-major layout (a run of tiles of one
// block row).  3-stage bulk-copy pipeline: thread 0 is the producer, one __syncthreads per slab
// releases the stage consumed in the previous iteration.
// ---------------------------------------------------------------------------------------------
struct GemmPipe {
    double *stages;     // PGPFA_STAGES x 2 x PGPFA_SLAB doubles
    uint64_t *full;     // PGPFA_STAGES barriers
    uint32_t it;        // slabs consumed so far (keeps stage/parity continuity across calls)
};

__device__ __forceinline__ void pipe_setup(GemmPipe &p, unsigned char *smem_raw) {
    p.stages = reinterpret_cast<double *>(smem_raw);
    p.full = reinterpret_cast<uint64_t *>(smem_raw + PGPFA_BAR_OFF);
    p.it = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < PGPFA_STAGES; s++) mbar_init(&p.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
}

__device__ __forceinline__ void slab_mma(double (&acc)[4][4][2], const double *As, const double *Bs, int wm, int wn,
                                         int lane) {
    const double2 *A2 = reinterpret_cast<const double2 *>(As);
    const double2 *B2 = reinterpret_cast<const double2 *>(Bs);
#pragma unroll
    for (int kp = 0; kp < 4; kp++) {
        double2 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = A2[(kp * 8 + wm * 4 + i) * 32 + lane];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = B2[(kp * 8 + wn * 4 + j) * 32 + lane];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
                dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
            }
    }
}

template <bool SAME>
__device__ __forceinline__ void gemm_slabs(double (&acc)[4][4][2], GemmPipe &p, const double *__restrict__ A,
                                           const double *__restrict__ B, int nslab) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const uint32_t bytes = SAME ? PGPFA_SLAB * 8 : 2 * PGPFA_SLAB * 8;
    const uint32_t base = p.it;
    if (tid == 0) {
        const int pre = nslab < (PGPFA_STAGES - 1) ? nslab : (PGPFA_STAGES - 1);
        for (int s = 0; s < pre; s++) {
            const uint32_t st = (base + s) % PGPFA_STAGES;
            double *dst = p.stages + st * 2 * PGPFA_SLAB;
            mbar_expect_tx(&p.full[st], bytes);
            bulk_g2s(dst, A + (size_t)s * PGPFA_SLAB, PGPFA_SLAB * 8, &p.full[st]);
            if (!SAME) bulk_g2s(dst + PGPFA_SLAB, B + (size_t)s * PGPFA_SLAB, PGPFA_SLAB * 8, &p.full[st]);
        }
    }
    for (int s = 0; s < nslab; s++) {
        if (tid == 0) {
            const int sn = s + PGPFA_STAGES - 1;
            if (sn < nslab) {
                const uint32_t st = (base + sn) % PGPFA_STAGES;
                double *dst = p.stages + st * 2 * PGPFA_SLAB;
                mbar_expect_tx(&p.full[st], bytes);
                bulk_g2s(dst, A + (size_t)sn * PGPFA_SLAB, PGPFA_SLAB * 8, &p.full[st]);
                if (!SAME) bulk_g2s(dst + PGPFA_SLAB, B + (size_t)sn * PGPFA_SLAB, PGPFA_SLAB * 8, &p.full[st]);
            }
        }
        const uint32_t g = base + s;
        const uint32_t st = g % PGPFA_STAGES;
        mbar_wait(&p.full[st], (g / PGPFA_STAGES) & 1);
        const double *As = p.stages + st * 2 * PGPFA_SLAB;
        const double *Bs = SAME ? As : As + PGPFA_SLAB;
        slab_mma(acc, As, Bs, wm, wn, lane);
        __syncthreads();
    }
    p.it = base + nslab;
}

// accumulator fragment (i, j) of this lane  <->  tile rows / cols
__device__ __forceinline__ int frag_row(int wm, int i, int lane) { return wm * 32 + i * 8 + (lane >> 2); }
__device__ __forceinline__ int frag_col(int wn, int j, int lane) { return wn * 32 + j * 8 + 2 * (lane & 3); }
// double2 slot of accumulator fragment (i, j) inside a fragment-major tile
__device__ __forceinline__ int frag_slot2(int wm, int wn, int i, int j, int lane) {
    return wn * 1024 + j * 256 + (wm * 4 + i) * 32 + lane;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum; `red` must hold >= 32 doubles of shared memory; result valid in all threads
__device__ __forceinline__ double block_sum(double v, double *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; w++) t += red[w];
    return t;
}
__device__ __forceinline__ double block_max(double v, double *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = red[0];
    for (int w = 1; w < nw; w++) t = fmax(t, red[w]);
    return t;
}

}  // namespace pgpfa
