"""Pure-numpy pieces that feed the hot path but are not part of it: the Poisson-PCA
initialiser, a minimal experiment container and the synthetic-data generator.  This file imports nothing from the
package (no ctypes, no CUDA library), so bench.py can load it by path for BOTH arms — the reference arm then maps no
product code at all."""
import numpy as np


def initializeParams(xdim, ydim, experiment=None):
    """funs/util.py:505-558 (random, or Poisson-PCA moment matching when an experiment is given)."""
    if experiment is None:
        print('Initializing parameters randomly..')
        return {'C': np.random.rand(ydim, xdim) * 2 - 1,
                'd': np.random.randn(ydim) * 2 - 2,
                'tau': np.random.rand(xdim) * 0.5}
    print('Initializing parameters with Poisson-PCA..')
    spikes = np.concatenate([np.asarray(tr['Y'], dtype=np.float64) for tr in experiment.data], axis=1)
    meanY = np.mean(spikes, 1) + 1e-10
    covY = np.cov(spikes)
    lamb = np.log(np.abs(covY + np.outer(meanY, meanY) - np.diag(meanY))) - np.log(np.outer(meanY, meanY))
    gamma = np.log(meanY)
    evals, evecs = np.linalg.eig(lamb)
    order = np.argsort(evals)[::-1]
    evecs = evecs[:, order][:, :xdim]
    return {'C': evecs, 'd': gamma, 'tau': np.random.rand(xdim) * 0.5 + 0.1}


class Experiment:
    """Minimal duck-typed experiment (attributes read by the hot path: funs/engine.py:131-136)."""

    def __init__(self, data, trialDur, binSize, params=None):
        self.data = data
        self.trialDur = trialDur
        self.binSize = binSize
        self.T = int(trialDur / binSize)
        self.numTrials = len(data)
        self.ydim = np.shape(data[0]['Y'])[0]
        if params is not None:
            self.params = params
            self.xdim = np.shape(params['C'])[1]


def simulate(seed, xdim, ydim, numTrials, T, binSize=10, dOffset=-1.0, tau=None):
    """Synthetic Poisson-GPFA data with the reference generator's distributions (funs/util.py:707-750):
    C ~ U(-0.5,0.5), d ~ -2U(0,1)+dOffset, x_k ~ GP(0,K(tau_k)), y ~ Poisson(exp(Cx+d)).  Sampled per
    latent through a Cholesky factor (host side; input generation is outside the hot path)."""
    rng = np.random.RandomState(seed)
    C = rng.rand(ydim, xdim) - 0.5
    d = rng.rand(ydim) * (-2) + dOffset
    tau = np.linspace(0.05, 0.3, xdim) if tau is None else np.asarray(tau, dtype=np.float64)
    t_ms = np.arange(T) * float(binSize)
    dif2 = (t_ms[:, None] - t_ms[None, :]) ** 2
    Lk = np.stack([np.linalg.cholesky(0.999 * np.exp(-0.5 * dif2 / (tk * 1000) ** 2) + 0.001 * np.eye(T)) for tk in tau])
    data = []
    for _ in range(numTrials):
        X = np.einsum('kts,ks->kt', Lk, rng.randn(xdim, T))
        data.append({'X': X, 'Y': rng.poisson(np.exp(C @ X + d[:, None]))})
    return Experiment(data, T * binSize, binSize, {'C': C, 'd': d, 'tau': tau.copy()})
