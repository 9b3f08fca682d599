"""Mirror of the hot-path part of the reference's ``funs/util.py`` (same names, arguments, returns).

GP covariance construction runs on the device; vec(C,d) packing, trial sub-sampling and the
Poisson-PCA initialiser are host-side bookkeeping exactly as in the reference (they feed the hot
path but are not part of it).
"""
import copy

import numpy as np
import torch

from . import _lib, kernels as kn
from ._synth import Experiment, initializeParams, simulate  # noqa: F401  (pure numpy, no library import: _synth.py)


def makeK_big(params, trialDur, binSize, epsNoise=0.001):
    """(K_big (qT,qT), K (q,T,T)) — funs/util.py:599-619.  Like the reference it flattens
    ``params['tau']`` in place (:602)."""
    params['tau'] = np.ndarray.flatten(np.asarray(params['tau'], dtype=np.float64))
    T = int(trialDur / binSize)
    K = kn.make_K(_lib.dev_f64(params['tau']), T, binSize, epsNoise)
    K_big = kn.make_K_big(K)
    return K_big.cpu().numpy(), K.cpu().numpy()


def makeCd_big(params, T):
    """(C_big (qT,NT), d_big (NT,)) — funs/util.py:594-597.  Kept for API compatibility only: the
    rebuilt hot path never materialises C_big.  Pure data movement (a Kronecker
    product with the identity), done with device indexing."""
    C = _lib.dev_f64(params['C'])
    N, q = C.shape
    C_big = torch.zeros(q, T, N, T, dtype=torch.float64, device="cuda")
    idx = torch.arange(T, device="cuda")
    C_big[:, idx, :, idx] = C.T.reshape(1, q, N).expand(T, q, N)
    d_big = _lib.dev_f64(np.ravel(params['d'])).repeat_interleave(T)
    return C_big.reshape(q * T, N * T).cpu().numpy(), d_big.cpu().numpy()


def CdtoVecCd(C, d):
    """funs/util.py:560-574: vecCd[j*N+n] = C[n,j], vecCd[q*N+n] = d[n]."""
    C = np.asarray(C)
    return np.concatenate([C.T.reshape(-1), np.ravel(d)])


def vecCdtoCd(vecCd, xdim, ydim):
    """funs/util.py:576-592."""
    m = np.reshape(vecCd, [xdim + 1, ydim]).T
    return m[:, :xdim], m[:, xdim]


def seenTrials(experiment, seenIdx):
    """funs/util.py:449-457."""
    idx = np.asarray(seenIdx).flatten()
    out = copy.copy(experiment)
    out.data = [experiment.data[i] for i in idx]
    out.numTrials = len(out.data)
    # the copy must not inherit the parent's device caches or its stacked counts (they describe OTHER trials)
    for key in ('_pgpfa_dev', '_pgpfa_y', 'Y_all'):
        out.__dict__.pop(key, None)
    return out


def subsampleTrials(experiment, batchSize):
    """funs/util.py:459-473: one ``np.random.choice(numTrials, batchSize, replace=False)`` on the GLOBAL
    numpy RNG per call (RNG-stream parity with the reference's online EM)."""
    numTrials = len(experiment.data)
    batchTrIdx = np.random.choice(numTrials, batchSize, replace=False)
    out = copy.copy(experiment)
    out.data = [experiment.data[i] for i in batchTrIdx]
    out.numTrials = batchSize
    out.batchTrIdx = batchTrIdx
    out._pgpfa_parent = experiment      # lets the device layer gather the batch from the resident parent
    for key in ('_pgpfa_dev', '_pgpfa_y', 'Y_all'):
        out.__dict__.pop(key, None)
    return out


def JSLogdetDiv(X, Y):
    """funs/util.py:21-22 (post-fit diagnostic, host side)."""
    return np.log(np.linalg.det((X + Y) / 2)) - 1 / 2 * np.log(np.linalg.det(X.dot(Y)))


def getMeanCovYfromParams(params, experiment):
    """funs/util.py:24-39: model-implied mean / second moment of the spike counts (post-fit diagnostic, host side)."""
    rho = np.ravel(params['d'])
    lamb = np.dot(params['C'], np.asarray(params['C']).T)
    E_y = np.exp(1 / 2 * np.diag(lamb) + rho)
    E_yy = np.outer(E_y, E_y) * np.exp(lamb / 2)
    E_yy[np.diag_indices_from(E_yy)] = E_y + np.exp(np.diag(lamb) / 2) * E_y ** 2
    return E_y, E_yy


def simulate_on_device(seed, xdim, ydim, numTrials, T, binSize=10, dOffset=-1.0, tau=None):
    """Same model and parameter distributions as ``simulate`` with the latent trajectories and the counts drawn on
    the device (``pgpfa_sample_*``): configs[4]-scale inputs (16384 trials) in well under a second instead of half a
    minute.  Returns an Experiment whose counts are one (R,N,T) array (``Y_all``) plus per-trial views."""
    rng = np.random.RandomState(seed)
    C = rng.rand(ydim, xdim) - 0.5
    d = rng.rand(ydim) * (-2) + dOffset
    tau = np.linspace(0.05, 0.3, xdim) if tau is None else np.asarray(tau, dtype=np.float64)
    X, Y = kn.sample_dataset(_lib.dev_f64(C), _lib.dev_f64(d), _lib.dev_f64(tau), numTrials, T, binSize, seed)
    Yh, Xh = Y.cpu().numpy(), X.cpu().numpy()
    ex = Experiment([{'X': Xh[r], 'Y': Yh[r]} for r in range(numTrials)], T * binSize, binSize,
                    {'C': C, 'd': d, 'tau': tau.copy()})
    ex.Y_all = Yh
    return ex


