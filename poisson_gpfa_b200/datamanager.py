"""Mirror of the reference's ``funs/datamanager.py:8-54`` ``StevensonDataset`` (BASELINE.json configs[1]).

The loader walks the .mat structure ``matdat['Subject'][id]['Trial'][0][tr]['Time' | 'Neuron']`` exactly like the
reference (second half of the trials, ``ydim`` neurons, ``trialDur`` ms from the first time stamp), but collects the
spike times into one CSR array and bins them on the device (``pgpfa_bin_spikes``, numpy.histogram semantics) instead
of one ``np.histogram`` call per trial and neuron.  ``data/Stevenson_2011_e1.mat`` is not shipped with the reference
checkout used here (.MISSING_LARGE_BLOBS); ``synthetic_matdat`` builds a schema-compatible stand-in for tests.
"""
import os

import numpy as np
import torch

from . import _lib, kernels as kn


def synthetic_matdat(seed=0, numTrials=8, ydim=12, dur_s=1.6, rate_hz=20.0):
    """A nested structure that indexes like scipy.io.loadmat's output for the Stevenson file."""
    rng = np.random.RandomState(seed)
    trials = []
    t_start = 10.0
    for _ in range(numTrials):
        time = t_start + np.sort(rng.rand(50)) * dur_s
        neurons = []
        for _n in range(ydim):
            k = rng.poisson(rate_hz * dur_s)
            spk = np.sort(time[0] + rng.rand(k) * (time[-1] - time[0]))
            neurons.append([[spk.reshape(-1, 1)]])
        trials.append({'Time': [time.reshape(-1, 1)], 'Neuron': [neurons]})
        t_start += dur_s + rng.rand()
    return {'Subject': [{'Trial': [trials]}]}


class StevensonDataset():
    def __init__(self, subject_id=0, ydim=90, trialDur=1400, binSize=10, numTrials=100, ydimData=False, numTrData=True,
                 matfile='data/Stevenson_2011_e1.mat', matdat=None):
        T = int(trialDur / binSize)
        if matdat is None:
            if not os.path.exists(matfile):
                raise FileNotFoundError("%s not found (the recording is not part of the repository); pass matdat=..." % matfile)
            import scipy.io as sio
            matdat = sio.loadmat(matfile)
        self.matdat = matdat
        trials = matdat['Subject'][subject_id]['Trial'][0]
        if numTrData:
            numTrials = len(trials)
        if ydimData:
            ydim = len(trials[0]['Neuron'][0])
        self.trial_durs = []
        for trial_id in range(numTrials):
            tt = np.asarray(trials[trial_id]['Time'][0]).flatten()
            self.trial_durs.append(np.max(tt) - np.min(tt))
        used = list(range(int(numTrials / 2), numTrials))
        chunks, ptr_, t0, spike_time = [], [0], [], []
        for trial_id in used:
            tt = np.asarray(trials[trial_id]['Time'][0]).flatten()
            begin = np.min(tt)
            t0.append(begin)
            per_neuron = []
            for yd in range(ydim):
                spk = np.asarray(trials[trial_id]['Neuron'][0][yd][0][0], dtype=np.float64).flatten()
                chunks.append(spk)
                ptr_.append(ptr_[-1] + spk.size)
                rel = spk - begin
                per_neuron.append(rel[rel < trialDur / 1000])
            spike_time.append(per_neuron)
        R = len(used)
        times = np.concatenate(chunks) if chunks else np.zeros(0)
        Y = kn.bin_spikes(_lib.dev_f64(times), torch.as_tensor(np.asarray(ptr_, dtype=np.int64)).cuda(),
                          _lib.dev_f64(np.asarray(t0)), trialDur / 1000, R, ydim, T)
        Yh = Y.cpu().numpy().astype(np.int64)
        self.data = [{'Y': Yh[r], 'spike_time': spike_time[r]} for r in range(R)]
        self.__dict__['_pgpfa_y'] = (Y, self.data)  # counts are already resident on the device (cache tied to this list)
        self.trialDur = trialDur
        self.binSize = binSize
        self.numTrials = int(numTrials / 2)
        self.ydim = ydim
        self.T = T
        self.all_raster = np.concatenate([d['Y'] for d in self.data], axis=1).astype(np.float64)
        self.avgFR = self.all_raster.mean(1) / (binSize / 1000.0)
