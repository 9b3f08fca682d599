import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            a = a.detach().cpu().numpy()
        if isinstance(b, torch.Tensor):
            b = b.detach().cpu().numpy()
    except ImportError:
        pass
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


class Exp:
    """Duck-typed experiment built from a golden file."""

    def __init__(self, g, sel=None):
        Y = g['Y'] if sel is None else g['Y'][sel]
        self.data = [{'Y': Y[r]} for r in range(Y.shape[0])]
        self.trialDur = float(g['trialDur'])
        self.binSize = float(g['binSize'])
        self.T = Y.shape[2]
        self.ydim = Y.shape[1]
        self.numTrials = Y.shape[0]


def init_params(g):
    return {'C': g['init_C'].copy(), 'd': g['init_d'].copy(), 'tau': g['init_tau'].copy()}
