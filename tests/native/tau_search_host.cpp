// Host build of the timescale-search controller (poisson_gpfa_b200/csrc/tau_search.h) for the CPU test-suite:
// tests/test_tau_search_host.py drives it with the oracle's cost / gradient and compares it with the host-driven
// search of round 1.  Built on the fly with g++ (no CUDA needed).
#include "../../poisson_gpfa_b200/csrc/tau_search.h"

extern "C" {
int c_tau_state_bytes() { return (int)sizeof(TauLatent); }
void c_tau_init(void *s, double tau_old, double bs, int m, double *cands) { tau_init(*(TauLatent *)s, tau_old, bs, m, cands); }
void c_tau_merge(void *s, int m, const double *cands, const double *g, const double *f, int first) {
    tau_merge(*(TauLatent *)s, m, cands, g, f, first != 0);
}
int c_tau_next(void *s, int m, double xtol, double *cands) { return tau_next(*(TauLatent *)s, m, xtol, cands); }
void c_tau_result(void *s, double *out4) {
    TauLatent &t = *(TauLatent *)s;
    tau_result(t, out4[0], out4[1], out4[2]);
    out4[3] = (double)t.bracketed;
}
}
