// Handle lifecycle and error reporting of libpgpfa_b200.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "pgpfa_internal.h"

static thread_local char g_last_cuda_error[512] = "";
unsigned long long g_pgpfa_launches = 0;

void pgpfa_set_last_cuda_error(cudaError_t e, const char *file, int line) {
    snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s (%s) at %s:%d", cudaGetErrorName(e),
             cudaGetErrorString(e), file, line);
    cudaGetLastError();   // clear the sticky-less error state
}

extern "C" const char *pgpfa_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" int pgpfa_abi_version(void) { return PGPFA_ABI_VERSION; }

extern "C" long long pgpfa_launch_count(void) { return (long long)g_pgpfa_launches; }

// ---- lightweight in-stream profiling (CUDA events around kernel families inside the drivers) ----
void pgpfa_prof_begin(pgpfa_handle_t h, int slot, cudaStream_t st) {
    if (!h || !h->profiling) return;
    PgpfaProfSpan sp;
    sp.slot = slot;
    cudaEventCreate(&sp.e0);
    cudaEventCreate(&sp.e1);
    cudaEventRecord(sp.e0, st);
    h->open_spans.push_back(sp);
}
void pgpfa_prof_end(pgpfa_handle_t h, cudaStream_t st) {
    if (!h || !h->profiling || h->open_spans.empty()) return;
    cudaEventRecord(h->open_spans.back().e1, st);
    h->spans.push_back(h->open_spans.back());
    h->open_spans.pop_back();
}
void pgpfa_prof_resolve(pgpfa_handle_t h) {   // call after a stream synchronise
    if (!h) return;
    for (auto &sp : h->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) { h->prof_ms[sp.slot] += ms; h->prof_cnt[sp.slot] += 1; }
        cudaEventDestroy(sp.e0);
        cudaEventDestroy(sp.e1);
    }
    h->spans.clear();
}

// ---- progress ring: device -> host counters without stream synchronisation -----------------------
int pgpfa_prog_alloc(pgpfa_handle_s *h, unsigned long long *seq, int **dev_word) {
    if (!h || !h->prog_h) return PGPFA_ERR_ARG;
    const unsigned long long s = h->prog_seq++;
    h->prog_h[s % PGPFA_PROG_RING] = -1;
    *seq = s;
    *dev_word = h->prog_d + (s % PGPFA_PROG_RING);
    return PGPFA_OK;
}
bool pgpfa_prog_peek(pgpfa_handle_s *h, unsigned long long seq, int *value) {
    const int v = h->prog_h[seq % PGPFA_PROG_RING];
    if (v < 0) return false;
    *value = v;
    return true;
}
int pgpfa_prog_wait(pgpfa_handle_s *h, unsigned long long seq, cudaStream_t st, int *value, bool newest) {
    // `newest`: nothing was enqueued behind the producer of this word, i.e. the wait empties the stream like a
    // cudaStreamSynchronize (counted as a host synchronisation).  Otherwise it is a throttle: the host is merely too
    // far ahead, the device still has later iterations queued while the host waits.
    if (newest) h->n_drain++;
    if (pgpfa_prog_peek(h, seq, value)) return PGPFA_OK;
    if (!newest) h->n_throttle++;
    unsigned long long spins = 0;
    for (;;) {
        if (pgpfa_prog_peek(h, seq, value)) return PGPFA_OK;
        if ((++spins & 0xfff) == 0) {
            // a faulted or finished stream would never write the word: check instead of spinning forever
            const cudaError_t e = cudaStreamQuery(st);
            if (e == cudaSuccess) {
                if (pgpfa_prog_peek(h, seq, value)) return PGPFA_OK;
                pgpfa_set_last_cuda_error(cudaErrorUnknown, __FILE__, __LINE__);
                return PGPFA_ERR_CUDA;
            }
            if (e != cudaErrorNotReady) { pgpfa_set_last_cuda_error(e, __FILE__, __LINE__); return PGPFA_ERR_CUDA; }
        }
    }
}
int pgpfa_sync(pgpfa_handle_s *h, cudaStream_t st) {
    if (h) h->n_sync++;
    PGPFA_CUDA_TRY(cudaStreamSynchronize(st));
    return PGPFA_OK;
}
extern "C" int pgpfa_set_loop_depth(pgpfa_handle_t h, int depth) {
    if (!h || depth < 0 || depth > 16) return PGPFA_ERR_ARG;
    h->loop_depth = depth;
    return PGPFA_OK;
}
extern "C" long long pgpfa_host_sync_count(pgpfa_handle_t h) { return h ? h->n_sync + h->n_drain : -1; }
extern "C" long long pgpfa_throttle_wait_count(pgpfa_handle_t h) { return h ? h->n_throttle : -1; }

extern "C" int pgpfa_stream_wait_means(pgpfa_handle_t h, cudaStream_t waiting_stream) {
    if (!h) return PGPFA_ERR_ARG;
    PGPFA_CUDA_TRY(cudaStreamWaitEvent(waiting_stream, h->ev_means, 0));
    return PGPFA_OK;
}

extern "C" int pgpfa_set_profiling(pgpfa_handle_t h, int on) {
    if (!h) return PGPFA_ERR_ARG;
    h->profiling = on != 0;
    for (int i = 0; i < PGPFA_PROF_SLOTS; i++) { h->prof_ms[i] = 0.0; h->prof_cnt[i] = 0; h->prof_work[i] = 0.0; }
    return PGPFA_OK;
}

extern "C" int pgpfa_get_profile(pgpfa_handle_t h, double *ms_out, double *work_out, long long *count_out) {
    if (!h || !ms_out) return PGPFA_ERR_ARG;
    cudaDeviceSynchronize();
    pgpfa_prof_resolve(h);
    for (int i = 0; i < PGPFA_PROF_SLOTS; i++) {
        ms_out[i] = h->prof_ms[i];
        if (work_out) work_out[i] = h->prof_work[i];
        if (count_out) count_out[i] = h->prof_cnt[i];
    }
    return PGPFA_OK;
}

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
    h->profiling = false;
    for (int i = 0; i < PGPFA_PROF_SLOTS; i++) { h->prof_ms[i] = 0.0; h->prof_cnt[i] = 0; h->prof_work[i] = 0.0; }
    PGPFA_CUDA_TRY(cudaGetDevice(&h->device));
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&h->pinned), 256, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        pgpfa_set_last_cuda_error(e, __FILE__, __LINE__);
        delete h;
        return PGPFA_ERR_CUDA;
    }
    h->prog_h = nullptr; h->prog_d = nullptr; h->prog_seq = 0; h->stage_h = nullptr; h->n_sync = 0; h->n_drain = 0; h->n_throttle = 0;
    h->loop_depth = 4;
    if (const char *e = getenv("PGPFA_LOOP_DEPTH")) { const int v = atoi(e); if (v >= 0 && v <= 16) h->loop_depth = v; }
    {
        int *ph = nullptr;
        PGPFA_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&ph), PGPFA_PROG_RING * sizeof(int), cudaHostAllocMapped));
        for (int i = 0; i < PGPFA_PROG_RING; i++) ph[i] = -1;
        h->prog_h = ph;
        PGPFA_CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->prog_d), ph, 0));
        PGPFA_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&h->stage_h), PGPFA_STAGE_BYTES, cudaHostAllocDefault));
    }
    for (int p = 0; p < PGPFA_MAX_PARTS; p++) h->s_part[p] = nullptr;
    PGPFA_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    PGPFA_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_means, cudaEventDisableTiming));
    PGPFA_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_stage, cudaEventDisableTiming));
    for (int p = 0; p < PGPFA_MAX_PARTS; p++) {
        PGPFA_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join[p], cudaEventDisableTiming));
        PGPFA_CUDA_TRY(cudaStreamCreateWithFlags(&h->s_part[p], cudaStreamNonBlocking));
    }
    *out = h;
    return PGPFA_OK;
}

extern "C" int pgpfa_destroy(pgpfa_handle_t h) {
    if (!h) return PGPFA_OK;
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->prog_h) cudaFreeHost(const_cast<int *>(h->prog_h));
    if (h->stage_h) cudaFreeHost(h->stage_h);
    if (h->s_part[0]) {
        cudaEventDestroy(h->ev_fork);
        cudaEventDestroy(h->ev_means);
        cudaEventDestroy(h->ev_stage);
        for (int p = 0; p < PGPFA_MAX_PARTS; p++) {
            if (h->s_part[p]) cudaStreamDestroy(h->s_part[p]);
            cudaEventDestroy(h->ev_join[p]);
        }
    }
    delete h;
    return PGPFA_OK;
}
