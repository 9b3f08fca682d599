// FP64 peak probe for B200 (sm_100a): DMMA m8n8k4 vs DFMA register-only loops,
// plus smem-fed DMMA. Used once to choose the factorisation kernel's math pipe
// and to record the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 line).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_loop(double *out, int iters) {
    double c[NACC][2];
    #pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_loop(double *out, int iters) {
    double c[NACC];
    #pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = threadIdx.x * 1e-3 + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// smem-fed DMMA: warp computes a 32x32 tile; fragments read from shared memory each k-step
// in the fragment-major layout planned for the factorisation (LDS.128 per two k4 steps).
__global__ void dmma_smem_loop(double *out, int iters) {
    extern __shared__ double sm[];
    // 2 operands x 64 rows x 32 k  (fragment-major: [kpair 4][rb 8][lane 32][2])
    for (int i = threadIdx.x; i < 2 * 64 * 32; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 1, wn = (warp >> 1) & 1;
    double c[4][4][2];
    #pragma unroll
    for (int i = 0; i < 4; i++)
        #pragma unroll
        for (int j = 0; j < 4; j++) { c[i][j][0] = 0; c[i][j][1] = 0; }
    const double2 *A = reinterpret_cast<const double2 *>(sm);
    const double2 *B = reinterpret_cast<const double2 *>(sm + 64 * 32);
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int kp = 0; kp < 4; kp++) {
            double2 a[4], b[4];
            #pragma unroll
            for (int i = 0; i < 4; i++) a[i] = A[(kp * 8 + wm * 4 + i) * 32 + lane];
            #pragma unroll
            for (int j = 0; j < 4; j++) b[j] = B[(kp * 8 + wn * 4 + j) * 32 + lane];
            #pragma unroll
            for (int i = 0; i < 4; i++)
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    dmma884(c[i][j][0], c[i][j][1], a[i].x, b[j].x);
                    dmma884(c[i][j][0], c[i][j][1], a[i].y, b[j].y);
                }
        }
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 4; i++)
        #pragma unroll
        for (int j = 0; j < 4; j++) s += c[i][j][0] + c[i][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, nsm, p.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * nsm * 16 * 1024));
    const int iters = 20000;
    // DMMA register-only: sweep warps/SM and accumulators
    int wcfg[] = {4, 8, 16, 32};
    for (int wi = 0; wi < 4; wi++) {
        int warps = wcfg[wi];
        int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks = nsm * (warps * 32 / threads);
        float ms8 = timeit([&] { dmma_loop<8><<<blocks, threads>>>(out, iters); }, 3);
        float ms16 = timeit([&] { dmma_loop<16><<<blocks, threads>>>(out, iters); }, 3);
        double fl8 = 2.0 * 256 * 8 * (double)iters * warps * nsm, fl16 = 2.0 * fl8;
        printf(" \"dmma_tflops_w%d_acc8\": %.2f, \"dmma_tflops_w%d_acc16\": %.2f,\n", warps, fl8 / ms8 * 1e-9, warps, fl16 / ms16 * 1e-9);
    }
    for (int wi = 0; wi < 4; wi++) {
        int warps = wcfg[wi];
        int threads = warps * 32 > 1024 ? 1024 : warps * 32;
        int blocks = nsm * (warps * 32 / threads);
        float ms = timeit([&] { dfma_loop<16><<<blocks, threads>>>(out, iters); }, 3);
        double fl = 2.0 * 32 * 16 * (double)iters * warps * nsm;
        printf(" \"dfma_tflops_w%d\": %.2f,\n", warps, fl / ms * 1e-9);
    }
    CK(cudaFuncSetAttribute(dmma_smem_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 64 * 32 * 8));
    for (int cps = 1; cps <= 4; cps *= 2) {
        float ms = timeit([&] { dmma_smem_loop<<<nsm * cps, 128, 2 * 64 * 32 * 8>>>(out, 2000); }, 3);
        double fl = 2.0 * 64 * 64 * 32 * 2000.0 * nsm * cps;
        printf(" \"dmma_smem_tflops_cta%d\": %.2f,\n", cps, fl / ms * 1e-9);
    }
    printf(" \"done\": 1}\n");
    return 0;
}
