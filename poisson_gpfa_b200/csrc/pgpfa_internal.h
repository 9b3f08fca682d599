// Internal host-side interfaces between the translation units of libpgpfa_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

// latent dimensionalities with compiled kernels (accumulator counts depend on q)
#define PGPFA_FOR_EACH_Q(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12)
#define PGPFA_QMAX 12

struct PgpfaMatSrc {
    const double *Kinv;   // (q,T,T) generator mode; nullptr => dense mode
    const double *W;      // (trials, q*q, T)
    const double *dense;  // (slots, n, n)
    int q, T, n;
    double diag_scale;
};

struct pgpfa_handle_s;
int pgpfa_i_factor(const PgpfaMatSrc &ms, double *L, double *Dinv, double *ZT, const int *act, int *info, int nslots,
                   cudaStream_t st, pgpfa_handle_s *h = nullptr, float *L32 = nullptr, float *D32 = nullptr);
int pgpfa_i_trtri(const double *L, const double *Dinv, double *ZT, int n, int nslots, cudaStream_t st,
                  pgpfa_handle_s *h = nullptr);
int pgpfa_i_lauum(const double *ZT, const int2 *pairs, int npairs, const int *act, double *vsmGP, double *dense, int n,
                  int q, int T, int nslots, cudaStream_t st);
int pgpfa_i_timediag(const double *ZT, const int *act, double *vsm, int n, int q, int T, int nslots, cudaStream_t st);
int pgpfa_i_logdet(const double *L, int n, int nslots, double *out, cudaStream_t st);
int pgpfa_i_tiles_to_dense(const double *tiles, int n, int upper, int nslots, double *out, cudaStream_t st);
int pgpfa_i_solve(const double *L, const double *Dinv, const double *rhs, double *out, double scale, const int *act,
                  int n, int nslots, cudaStream_t st, int lslot_base = -1, const int *lslot_map = nullptr);
// same solve streaming the FP32 mirrors of the factor (inexact: for chord sweeps only)
int pgpfa_i_solve32(const float *L32, const float *D32, const double *rhs, double *out, double scale, const int *act,
                    int n, int nslots, cudaStream_t st, int lslot_base = -1, const int *lslot_map = nullptr);

#include <vector>
#include "../../include/pgpfa_b200.h"

// profiling slots (pgpfa_get_profile): time in ms, algorithmic work (flops or bytes), span count
enum {
    PGPFA_PROF_FACTOR = 0,    // batched Cholesky (diag + panel kernels); work = trial-factorisations * n^3/3 flops
    PGPFA_PROF_SOLVE = 1,     // triangular solves; work = bytes of factor streamed (2 passes)
    PGPFA_PROF_EVAL = 2,      // prior mat-vec + fused rates/gradient/W + line search
    PGPFA_PROF_TRTRI = 3,     // triangular inverse; work = trials * n^3/3 flops
    PGPFA_PROF_SLICES = 4,    // time-diagonals + selected inverse tiles
    PGPFA_PROF_BLOCKFACTOR = 5,  // CG preconditioner set-up (q shared T x T inverses per E-step)
    PGPFA_PROF_LOWRANK = 6,      // low-rank posterior pass: per-bin matrices, capacitance GEMM, Y, slices (lowrank.cu)
    PGPFA_PROF_SLOTS = 8
};
#define PGPFA_MAX_PARTS 4
struct PgpfaProfSpan { cudaEvent_t e0, e1; int slot; };

#define PGPFA_PROG_RING 4096          // progress words (ints) in mapped pinned memory
#define PGPFA_STAGE_BYTES 65536        // pinned staging for small host -> device tables (truly asynchronous copies)
struct pgpfa_handle_s {
    int *pinned;        // small pinned host scratch for device -> host counters
    // Device -> host progress without stream synchronisation: compaction kernels write the number of still-active
    // trials into a ring of words in mapped pinned memory (prog_h = host view, prog_d = device view); the host
    // drivers read them to stop enqueueing / shrink grids and never wait for an empty stream inside a loop.
    volatile int *prog_h;
    int *prog_d;
    unsigned long long prog_seq;
    cudaEvent_t ev_stage;                   // completion of the last copy out of stage_h
    unsigned char *stage_h;                 // pinned staging buffer (PGPFA_STAGE_BYTES)
    int loop_depth;                         // how many loop iterations the drivers enqueue ahead of the last count read
    long long n_sync;                       // cudaStreamSynchronize / cudaDeviceSynchronize calls made by the library
    long long n_drain;                      // blocking reads of the NEWEST progress word (the stream runs empty: a sync)
    long long n_throttle;                   // blocking reads of an older word (device still has queued work: no sync)
    int device;
    bool profiling;
    double prof_ms[PGPFA_PROF_SLOTS];
    double prof_work[PGPFA_PROF_SLOTS];
    long long prof_cnt[PGPFA_PROF_SLOTS];
    std::vector<PgpfaProfSpan> spans, open_spans;
    cudaStream_t s_part[PGPFA_MAX_PARTS];   // streams for split batches (factor.cu)
    cudaEvent_t ev_fork, ev_join[PGPFA_MAX_PARTS];
    cudaEvent_t ev_means;                   // recorded after the time-diagonal kernel of a Laplace solve
};
// progress ring (api.cu): allocate a word (reset to -1), blocking / non-blocking read
int pgpfa_prog_alloc(pgpfa_handle_s *h, unsigned long long *seq, int **dev_word);
int pgpfa_prog_wait(pgpfa_handle_s *h, unsigned long long seq, cudaStream_t st, int *value, bool newest);
bool pgpfa_prog_peek(pgpfa_handle_s *h, unsigned long long seq, int *value);
int pgpfa_sync(pgpfa_handle_s *h, cudaStream_t st);      // counted cudaStreamSynchronize
// device-generated tile-pair table of pgpfa_i_cov_pairs (same order), no host temporary, no synchronisation
int pgpfa_i_gen_pairs(int2 *pairs_dev, int q, int T, bool all, cudaStream_t st);
int pgpfa_i_num_pairs(int q, int T, bool all);
void pgpfa_prof_begin(pgpfa_handle_t h, int slot, cudaStream_t st);
void pgpfa_prof_end(pgpfa_handle_t h, cudaStream_t st);
void pgpfa_prof_resolve(pgpfa_handle_t h);

// leave-one-neuron-out problems: problem -> row of y, problem -> excluded neuron (nullptr = ordinary trials)
struct LooMap { const int *ymap; const int *excl; };
static inline LooMap pgpfa_no_loo() { LooMap l; l.ymap = nullptr; l.excl = nullptr; return l; }

// `cnt` (optional, device): the true number of slots; nslots is then only the host's upper bound that sizes the grid
int pgpfa_i_small_gemm(const double *A, const double *B, double *C, int n, int batch, cudaStream_t st);   // mstep.cu
// `out2` (optional): Kmat holds 2 q matrices; matrices q .. 2q-1 are applied to the same vectors and land in out2
int pgpfa_i_prior_apply(const double *Kmat, const double *v, double *out, const int *act, int nslots, int q, int T,
                        cudaStream_t st, const int *cnt = nullptr, double *out2 = nullptr);
int pgpfa_i_laplace_eval(const double *x, const double *Kx, const double *y, const double *C, const double *d,
                         const int *act, int nslots, int q, int N, int T, double *f, double *g, double *W,
                         cudaStream_t st, const double *off = nullptr, LooMap loo = pgpfa_no_loo(),
                         const int *cnt = nullptr);
int pgpfa_i_linesearch(double *x, const double *dx, const double *Kx, const double *Kd, const double *g,
                       const double *y, const double *C, const double *d, const int *act, int nslots, int q, int N,
                       int T, double tol, double *fcur, int *conv, int *niter, double *steplen, int step_kind,
                       cudaStream_t st, const double *off = nullptr, LooMap loo = pgpfa_no_loo(), double *pcg_s = nullptr,
                       const int *cnt = nullptr);
int pgpfa_i_pautosum(const double *vsmGP, const double *m, int R, int q, int T, int accumulate, double *P,
                     cudaStream_t st);
std::vector<int2> pgpfa_i_cov_pairs(int q, int T, bool all);
int pgpfa_i_polish(double *x, const double *dx, const int *act, int n, double max_rel, double *steplen, int nslots,
                   cudaStream_t st);

// low-rank factor of the smooth part of the prior, K_k - eps I = F_k F_k^T (lowrank.cu)
struct PgpfaLowRank {
    const double *F, *Ft;            // (q,T,T): [k][t][a] and [k][a][t], columns / rows >= rank[k] zero
    int rank[PGPFA_QMAX], off[PGPFA_QMAX + 1];
    int r;                           // sum of the ranks
    double eps;
};
#define PGPFA_LOWRANK_TABLE_BYTES 65536
size_t pgpfa_i_lowrank_bytes_per_slot(int q, int T, int r);
int pgpfa_i_lowrank_prepare(pgpfa_handle_s *h, const PgpfaLowRank &lr, int q, int T, void *probs_dev, cudaStream_t st);
int pgpfa_i_lowrank_posterior(pgpfa_handle_s *h, const PgpfaLowRank &lr, const double *W, const double *gvec, double *x,
                              double *dx, const int *act, int nslots, int q, int T, double tol, double *steplen,
                              double *vsm, double *vsmGP, void *area, size_t area_bytes, void *probs_dev, cudaStream_t st,
                              int *info = nullptr, double *pautosum = nullptr, int pauto_accumulate = 0,
                              const double *post_mean = nullptr, double *pauto_partial = nullptr);
size_t pgpfa_i_pautosum_partial_bytes(int q, int T);
int pgpfa_i_iota(int *p, int n, int start, cudaStream_t st);
int pgpfa_i_compact(const int *act_in, int n_in, const int *conv, int keep_mask, int *act_out, int *n_out, cudaStream_t st,
                    const int *n_in_dev = nullptr, int *prog = nullptr);
