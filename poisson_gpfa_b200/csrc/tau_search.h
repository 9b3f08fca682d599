// Controller of the GP-timescale search (funs/learning.py:257-293, :771-830), shared by the device kernel
// (mstep.cu: tau_ctrl_kernel) and a host test harness (tests/tau_search_host.cpp).
//
// The reference hands each latent's scalar problem over p = log(1/tau_bins^2) to scipy (BFGS / TNC) started at the
// old timescale; the result is the first zero of the gradient in the descent direction from p0.  Here every latent
// keeps a sorted list of evaluated points (p, g, f); each round evaluates `m` candidate points per latent in ONE
// batched device evaluation (the T x T factorisations are latency-bound, so candidates are free):
//   round 0    p0 + {0, -0.1, 0.1, -0.25, 0.25, -0.5, 0.5, -1, 1}
//   later      if no sign change of g is bracketed yet: walk downhill with growing steps;
//              else: zero of the inverse interpolating polynomial through up to 4 neighbours of the bracket, with
//              the other candidates placed around it at distances matched to its error estimate.
// A latent is done when the error estimate or the bracket width is below xtol (1 + |p|).
//
// Everything here is plain arithmetic on a per-latent state; the same source compiles for the host, so the logic is
// unit-tested on the CPU against an oracle evaluation of the reference's cost / gradient.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PGPFA_HD __host__ __device__ __forceinline__
#else
#define PGPFA_HD inline
#endif

#define TAU_MAXC 16       // candidates per latent and round (m <= TAU_MAXC, odd, >= 5)
#define TAU_MAXP 240      // evaluated points kept per latent (15 rounds of 16)

struct TauLatent {
    double p[TAU_MAXP], g[TAU_MAXP], f[TAU_MAXP];    // ascending in p (ties: g, then f)
    double p0, p_star, g0, f0;
    int npts, done, bracketed, walked_out;
};

PGPFA_HD bool tau_less(double p1, double g1, double f1, double p2, double g2, double f2) {
    if (p1 != p2) return p1 < p2;
    if (g1 != g2) return g1 < g2;
    return f1 < f2;
}

// insert (p, g, f) keeping the list sorted
PGPFA_HD void tau_insert(TauLatent &s, double p, double g, double f) {
    if (s.npts >= TAU_MAXP) return;
    int i = s.npts;
    while (i > 0 && tau_less(p, g, f, s.p[i - 1], s.g[i - 1], s.f[i - 1])) {
        s.p[i] = s.p[i - 1]; s.g[i] = s.g[i - 1]; s.f[i] = s.f[i - 1];
        i--;
    }
    s.p[i] = p; s.g[i] = g; s.f[i] = f;
    s.npts++;
}

PGPFA_HD void tau_offsets(int m, double *offs) {
    const double base[9] = {0.0, -0.1, 0.1, -0.25, 0.25, -0.5, 0.5, -1.0, 1.0};
    for (int c = 0; c < m; c++) offs[c] = c < 9 ? base[c] : (c & 1 ? -1.0 : 1.0) * (1.0 + 0.5 * ((c - 7) / 2));
}

// round 0: candidates around the old timescale.  tau_old in seconds, binSize in ms.
PGPFA_HD void tau_init(TauLatent &s, double tau_old, double binSize, int m, double *cands) {
    const double oldTau_bins = tau_old * 1000.0 / binSize;
    s.p0 = log(1.0 / (oldTau_bins * oldTau_bins));
    s.p_star = s.p0;
    s.npts = 0; s.done = 0; s.bracketed = 0; s.walked_out = 0;
    s.g0 = 0.0; s.f0 = 0.0;
    double offs[TAU_MAXC];
    tau_offsets(m, offs);
    for (int c = 0; c < m; c++) cands[c] = s.p0 + offs[c];
}

// merge the evaluated candidates of one round (first = the round-0 candidates)
PGPFA_HD void tau_merge(TauLatent &s, int m, const double *cands, const double *g, const double *f, bool first) {
    if (first) { s.g0 = g[0]; s.f0 = f[0]; }
    if (s.done) return;
    // a candidate is dropped only when a point evaluated in an EARLIER round sits at exactly the same p (clipped
    // candidates of one round may coincide with each other; they are all kept, exactly like the round-1 host code)
    bool skip[TAU_MAXC];
    for (int c = 0; c < m; c++) {
        skip[c] = false;
        for (int i = 0; i < s.npts && !skip[c]; i++) skip[c] = (s.p[i] == cands[c]);
    }
    for (int c = 0; c < m; c++)
        if (!skip[c]) tau_insert(s, cands[c], g[c], f[c]);
}

// Index i with a sign change of g between points i and i+1 nearest to p0 on the descent side.
// kind: 0 exact zero at i, 1 bracket (i, i+1), 2 none (i = the end to walk from)
PGPFA_HD int tau_bracket_of(const TauLatent &s, int &i_out) {
    int idx0 = 0;
    double best = fabs(s.p[0] - s.p0);
    for (int i = 1; i < s.npts; i++) {
        const double dd = fabs(s.p[i] - s.p0);
        if (dd < best) { best = dd; idx0 = i; }
    }
    if (s.g[idx0] == 0.0) { i_out = idx0; return 0; }
    if (s.g[idx0] < 0.0) {
        for (int i = idx0; i < s.npts - 1; i++) {
            if (s.g[i] == 0.0) { i_out = i; return 0; }
            if (s.g[i] < 0.0 && 0.0 <= s.g[i + 1]) { i_out = i; return 1; }
        }
        i_out = s.npts - 1;
    } else {
        for (int i = idx0 - 1; i >= 0; i--) {
            if (s.g[i] == 0.0) { i_out = i; return 0; }
            if (s.g[i] < 0.0 && 0.0 <= s.g[i + 1]) { i_out = i; return 1; }
        }
        i_out = 0;
    }
    return 2;
}

// zero of g inside (p_i, p_{i+1}) by inverse polynomial interpolation through up to 4 neighbours
PGPFA_HD void tau_interpolate(const TauLatent &s, int i, double &c, double &a, double &b, double &err) {
    a = s.p[i]; b = s.p[i + 1];
    const double ga = s.g[i], gb = s.g[i + 1];
    const int lo = i - 1 > 0 ? i - 1 : 0, hi = i + 3 < s.npts ? i + 3 : s.npts;     // [lo, hi)
    const int ns = hi - lo;
    c = a - ga * (b - a) / (gb - ga);
    err = 0.5 * (b - a) * (b - a);
    bool mono = ns >= 3;
    for (int u = lo; u + 1 < hi && mono; u++) mono = (s.g[u + 1] - s.g[u]) > 0.0;
    if (mono) {
        double est = 0.0, pmax = s.p[lo], pmin = s.p[lo];
        for (int u = lo; u < hi; u++) {
            double wgt = 1.0;
            for (int v = lo; v < hi; v++)
                if (v != u) wgt *= (0.0 - s.g[v]) / (s.g[u] - s.g[v]);
            est += wgt * s.p[u];
            pmax = fmax(pmax, s.p[u]); pmin = fmin(pmin, s.p[u]);
        }
        if (a < est && est < b) {
            c = est;
            const double span = pmax - pmin;
            err = 0.25 * (b - a) * (b - a) * pow(span, (double)(ns - 2));
        }
    }
}

PGPFA_HD double tau_clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// next round's candidates for one latent (all equal to p_star when the latent is done); returns 1 when done
PGPFA_HD int tau_next(TauLatent &s, int m, double xtol, double *cands) {
    for (int c = 0; c < m; c++) cands[c] = s.p_star;
    if (s.done) return 1;
    int i = 0;
    const int kind = tau_bracket_of(s, i);
    if (kind == 0) {
        s.p_star = s.p[i]; s.done = 1; s.bracketed = 1;
        for (int c = 0; c < m; c++) cands[c] = s.p_star;
        return 1;
    }
    if (kind == 2) {                                  // walk further downhill with growing steps
        const double edge = s.p[i];
        const double span = fmax(0.5, fabs(edge - s.p0));
        const double sgn = s.g[i] < 0.0 ? 1.0 : -1.0;
        for (int c = 0; c < m; c++) cands[c] = tau_clip(edge + sgn * span * (0.5 * pow(1.7, (double)c)), -40.0, 20.0);
        if (fabs(edge) >= 20.0) { s.done = 1; s.walked_out = 1; }    // monotone cost: keep the old tau (flagged)
        return s.done;
    }
    s.bracketed = 1;
    double c0, a, b, err;
    tau_interpolate(s, i, c0, a, b, err);
    const double w = b - a;
    s.p_star = c0;
    if (err <= xtol * (1.0 + fabs(c0)) || w <= xtol * (1.0 + fabs(a))) {
        s.done = 1;
        for (int c = 0; c < m; c++) cands[c] = s.p_star;
        return 1;
    }
    const double h1 = fmin(fmax(2.0 * err, 4.0 * xtol * (1.0 + fabs(c0))), w / 16.0);
    const double h2 = fmin(fmax(4.0 * h1, 0.25 * w * w), w / 4.0);
    double hs[TAU_MAXC];
    int nh = 0;
    hs[nh++] = h1; hs[nh++] = h2;
    for (int e = 1; e < (m - 1) / 2 - 1; e++) hs[nh++] = fmin(h2 * pow(4.0, (double)e), w / 2.5);
    const double lo = a + 1e-3 * w, hi = b - 1e-3 * w;
    int nc = 0;
    cands[nc++] = tau_clip(c0, lo, hi);
    for (int e = 0; e < nh && nc < m; e++) {
        cands[nc++] = tau_clip(c0 - hs[e], lo, hi);
        if (nc < m) cands[nc++] = tau_clip(c0 + hs[e], lo, hi);
    }
    for (; nc < m; nc++) cands[nc] = tau_clip(c0, lo, hi);
    return 0;
}

// result of the search: p_new (the old p0 when no sign change was ever bracketed) and cost / gradient at the evaluated
// point nearest to it
PGPFA_HD void tau_result(const TauLatent &s, double &p_new, double &fun, double &grad) {
    p_new = s.bracketed ? s.p_star : s.p0;
    int bi = 0;
    double best = fabs(s.p[0] - p_new);
    for (int i = 1; i < s.npts; i++) {
        const double dd = fabs(s.p[i] - p_new);
        if (dd < best) { best = dd; bi = i; }
    }
    fun = s.f[bi];
    grad = s.g[bi];
}
