"""The small-matrix inverse on the device (csrc/gpprior.cu: spd_sweep_kernel) sweeps the symmetric matrix in place with
row and column k folded into the same rank-1 form as the rest of the update, so that no element needs a select:
    u = c / d,  u_k = 1 - 1/d;   w = c,  w_k = d - 1;   A -= u w^T;   A_kk = -1/d      (c = column k, d = A_kk).
This CPU test pins that algebra (the result is -A^-1, the pivots are those of LDL^T) independently of the GPU.  The folded
entries c_i - (c_i/d)(d - 1) = c_i/d carry a rounding error ~eps * max(1, d): nothing for the unit-diagonal prior blocks the
kernel is used on (every pivot <= 1); the second test shows the effect on a badly scaled matrix and the Jacobi
equilibration D^-1/2 A D^-1/2 that removes it (left to the caller: in the kernel it costs registers it does not have)."""
import numpy as np
import pytest


def folded_sweep(A):
    M = np.array(A, dtype=np.float64)
    n = M.shape[0]
    logdet = 0.0
    for k in range(n):
        c = M[:, k].copy()
        d = c[k]
        logdet += np.log(d)
        u = c / d
        w = c.copy()
        u[k] = 1.0 - 1.0 / d
        w[k] = d - 1.0
        M -= np.outer(u, w)
        M[k, k] = -1.0 / d          # set exactly by the owner of the diagonal entry (the folded form gives 2 - 1/d - 2)
    return -M, logdet


def equilibrated_sweep(A):
    s = 1.0 / np.sqrt(np.diag(A))
    B = A * s[:, None] * s[None, :]
    np.fill_diagonal(B, 1.0)
    inv, logdet = folded_sweep(B)
    return inv * s[:, None] * s[None, :], logdet + np.log(np.diag(A)).sum()


@pytest.mark.parametrize("n", [1, 2, 7, 30, 64])
def test_folded_sweep_is_the_inverse(n):
    rng = np.random.RandomState(n)
    X = rng.randn(n, n)
    A = X @ X.T + n * np.eye(n)
    A /= np.abs(np.diag(A)).max()          # pivots <= 1, like the prior blocks
    inv, logdet = folded_sweep(A)
    ref = np.linalg.inv(A)
    assert np.abs(inv - ref).max() <= 2e-14 * np.abs(ref).max()
    assert abs(logdet - np.linalg.slogdet(A)[1]) <= 1e-12 * max(1.0, abs(logdet))
    assert np.abs(inv - inv.T).max() <= 1e-14 * np.abs(inv).max()


def test_fold_without_equilibration_loses_digits_on_large_diagonals():
    rng = np.random.RandomState(64)
    X = rng.randn(64, 64)
    A = X @ X.T + 64 * np.eye(64)
    ref = np.linalg.inv(A)
    plain, _ = folded_sweep(A)
    equil, _ = equilibrated_sweep(A)
    e_plain = np.abs(plain - ref).max() / np.abs(ref).max()
    e_equil = np.abs(equil - ref).max() / np.abs(ref).max()
    assert e_equil <= 2e-14 and e_plain >= 10 * e_equil          # measured 5e-13 vs 1e-15


def test_folded_sweep_on_a_prior_block():
    # the matrices it is mostly used on: K = (1 - eps) SE + eps I, unit diagonal, cond ~ 1e5 (funs/util.py:599-619)
    T, eps, tau_bins = 120, 1e-3, 12.0
    i = np.arange(T)
    K = (1 - eps) * np.exp(-0.5 * (i[:, None] - i[None, :]) ** 2 / tau_bins ** 2) + eps * np.eye(T)
    inv, logdet = folded_sweep(K)
    assert np.abs(K @ inv - np.eye(T)).max() <= 1e-9
    assert abs(logdet - np.linalg.slogdet(K)[1]) <= 1e-10 * abs(logdet)
