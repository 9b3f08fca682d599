// Handle lifecycle and error reporting of libpgpfa_b200.
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "pgpfa_internal.h"

static thread_local char g_last_cuda_error[512] = "";

void pgpfa_set_last_cuda_error(cudaError_t e, const char *file, int line) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s (%s) at %s:%d", cudaGetErrorName(e),
             cudaGetErrorString(e), file, line);
    cudaGetLastError();   // clear the sticky-less error state
}

extern "C" const char *pgpfa_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" int pgpfa_abi_version(void) { return PGPFA_ABI_VERSION; }

extern "C" const char *pgpfa_error_string(int code) {
    switch (code) {
        case PGPFA_OK: return "ok";
        case PGPFA_ERR_CUDA: return "CUDA runtime error (see pgpfa_last_cuda_error)";
        case PGPFA_ERR_ARG: return "invalid argument (null pointer, non-positive size, or unsupported latent dimension)";
        case PGPFA_ERR_WORKSPACE: return "workspace too small";
        case PGPFA_ERR_NOT_SPD: return "matrix not positive definite";
        case PGPFA_ERR_NOT_CONVERGED: return "Newton iteration limit reached before all trials converged";
        case PGPFA_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    }
    return "unknown error code";
}

extern "C" int pgpfa_create(pgpfa_handle_t *out) {
    if (!out) return PGPFA_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return PGPFA_ERR_NO_DEVICE;
    }
    pgpfa_handle_s *h = new pgpfa_handle_s();
    h->pinned = nullptr;
    PGPFA_CUDA_TRY(cudaGetDevice(&h->device));
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&h->pinned), 256, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        pgpfa_set_last_cuda_error(e, __FILE__, __LINE__);
        delete h;
        return PGPFA_ERR_CUDA;
    }
    *out = h;
    return PGPFA_OK;
}

extern "C" int pgpfa_destroy(pgpfa_handle_t h) {
    if (!h) return PGPFA_OK;
    if (h->pinned) cudaFreeHost(h->pinned);
    delete h;
    return PGPFA_OK;
}
