"""Mirror of the reference's ``funs/engine.py``: ``PPGPFAfit`` — the constructor IS the fit
(funs/engine.py:107-481).  Same keyword arguments and result attributes; the E- and M-steps run on
the B200 kernels and stay device-resident between each other.  Trials are sharded over the ranks of
an initialised ``torch.distributed`` group (one process per GPU); with a single process nothing
changes.  Plot methods of the reference are visualisation and out of scope.
"""
import copy
import sys
import time

import numpy as np

from . import inference, learning, util
from .dist import Reducer


def _banner(rows):
    print('+-------------------- Fit Options --------------------+')
    for label, val in rows:
        print(('| ' + label).ljust(54 - len(str(val)) - 1) + str(val) + ' |')
    print('+-----------------------------------------------------+')


class PPGPFAfit():
    """Poisson-GPFA fit; see the reference docstring (funs/engine.py:27-105) for the attribute list."""

    def __init__(self,
                 experiment,
                 initParams=None,
                 xdim=2,
                 inferenceMethod='laplace',
                 maxEMiter=50,
                 optimLogLamb=False,
                 CdOptimMethod='TNC',
                 tauOptimMethod='TNC',
                 verbose=False,
                 EMmode='Online',
                 batchSize=5,
                 onlineParamUpdateMethod='diag',
                 hessTol=None,
                 stepPow=0.75,
                 updateCdJointly=True,
                 fullyUpdateTau=False,
                 extractAllTraj=False,
                 extractAllTraj_trueParams=False,
                 getPredictionErr=False,
                 CdMaxIter=None,
                 tauMaxIter=None,
                 quiet=False):
        self.experiment = experiment
        ydim, T = np.shape(experiment.data[0]['Y'])
        trialDur = experiment.trialDur
        numTrials = len(experiment.data)
        binSize = experiment.binSize
        reducer = Reducer()
        self._reducer = reducer
        say = (not quiet) and reducer.rank == 0

        if initParams is None:
            initParams = util.initializeParams(xdim, ydim, experiment)
        else:
            _, xdim = np.shape(initParams['C'])

        posteriorLikelihood, variationalLowerBound, learningDetails = [], [], []
        params = initParams
        paramSeq = [initParams]
        learningTime, inferenceTime = [], []
        infRes = None

        if say:
            rows = [('Dimensionality of Latent State: ', xdim),
                    ('Dimensionality of Observed State (# neurons): ', ydim),
                    ('EM mode: ', EMmode), ('Max EM iterations: ', maxEMiter),
                    ('Inference Method: ', inferenceMethod)]
            if EMmode == 'Online':
                rows += [('Online Param Update Method: ', '`%s`' % onlineParamUpdateMethod),
                         ('Batch size (trials): ', batchSize)]
            _banner(rows)

        def e_step(exp, prev):
            if inferenceMethod == 'laplace':
                res, nll, optim = inference.laplace(exp, params, prevOptimRes=prev, verbose=verbose, reducer=reducer)
                return res, nll, None, optim
            if inferenceMethod == 'variational':
                res, nll, vlb, optim = inference.dualVariational(exp, params, optimizeLogLambda=optimLogLamb,
                                                                 prevOptimRes=prev, verbose=verbose, reducer=reducer)
                return res, nll, vlb, optim
            raise ValueError("inferenceMethod must be 'laplace' or 'variational'")

        def progress(i, nll, vlb):
            if not say:
                return
            out = 'Iteration: %3d of %3d, nPLL: = %.4f' % (i + 1, maxEMiter, nll)
            if vlb is not None:
                out += ', VLB = %.4f' % vlb
            sys.stdout.write('\r\x1b[K' + out)
            sys.stdout.flush()

        # ---------------------------------------------------------------- batch EM, funs/engine.py:156-239
        if EMmode == 'Batch':
            prev = None
            for i in range(maxEMiter):
                before = time.time()
                infRes, nll, vlb, prev = e_step(experiment, prev)      # warm start from the previous modes (:192-196)
                posteriorLikelihood.append(nll)
                if vlb is not None:
                    variationalLowerBound.append(vlb)
                inferenceTime.append(time.time() - before)
                before = time.time()
                params, learnDet = learning.updateParams(params, infRes, experiment, CdOptimMethod=CdOptimMethod)
                learningTime.append(time.time() - before)
                learningDetails.append(learnDet)
                paramSeq.append(params)
                progress(i, nll, vlb)

        # ---------------------------------------------------------------- online EM, funs/engine.py:243-450
        if EMmode == 'Online':
            gamma = np.linspace(0, 1, maxEMiter)
            regularizer_stepsize_Cd = 1 / (np.arange(maxEMiter) + 1) ** (stepPow)
            regularizer_stepsize_tau = 1 / (np.arange(maxEMiter) + 1) ** (stepPow)
            grad_descent_stepsize = 1 / (np.arange(maxEMiter) + 1) ** stepPow
            dimCd = xdim * ydim + ydim if updateCdJointly else xdim * ydim
            self.invPriorCovs = [np.diag(np.ones(dimCd))]
            self.cumHess = [np.diag(np.ones(dimCd))]
            seenTrialIdx = []
            for n in range(maxEMiter):
                subsampledDat = util.subsampleTrials(experiment, batchSize)   # the only RNG draw per iteration
                seenTrialIdx.append(subsampledDat.batchTrIdx)
                before = time.time()
                infRes, nll, vlb, _ = e_step(subsampledDat, None)             # always cold start (:298-301)
                posteriorLikelihood.append(nll)
                if vlb is not None:
                    variationalLowerBound.append(vlb)
                inferenceTime.append(time.time() - before)
                before = time.time()
                if onlineParamUpdateMethod in ('balancingGamma', 'sequentialAverage', 'fullyUpdateAll'):
                    newParams, learnDet = learning.updateParams(params, infRes, subsampledDat, CdOptimMethod=CdOptimMethod,
                                                                CdMaxIter=CdMaxIter, tauMaxIter=None, verbose=verbose)
                    nextParams = newParams
                    if onlineParamUpdateMethod == 'balancingGamma':
                        for key in ('C', 'd', 'tau'):
                            nextParams[key] = gamma[n] * params[key] + (1 - gamma[n]) * newParams[key]
                    elif onlineParamUpdateMethod == 'sequentialAverage':
                        for key in ('C', 'd', 'tau'):
                            nextParams[key] = (params[key] + newParams[key]) / 2
                elif onlineParamUpdateMethod in ('hess', 'diag'):
                    newParams, learnDet, priorCov = learning.updateParamsWithPrior(
                        params, infRes, subsampledDat, CdOptimMethod, tauOptimMethod,
                        regularizer_stepsize_Cd[n], regularizer_stepsize_tau[n], self.invPriorCovs[-1],
                        covOpts='useHessian' if onlineParamUpdateMethod == 'hess' else 'useDiag',
                        verbose=verbose, updateCdJointly=updateCdJointly, hessTol=hessTol)
                    nextParams = newParams
                    self.invPriorCovs.append(priorCov)
                elif onlineParamUpdateMethod == 'grad':
                    newParams, learnDet, hess = learning.updateParamsWithGradDescent(
                        params, infRes, subsampledDat, grad_descent_stepsize[n], self.cumHess[-1],
                        regularizer_stepsize_tau[n], tauOptimMethod, verbose=verbose,
                        updateCdJointly=updateCdJointly, hessTol=hessTol)
                    self.cumHess.append(self.cumHess[-1] + hess)
                    nextParams = newParams
                else:
                    raise ValueError('unknown onlineParamUpdateMethod %r' % (onlineParamUpdateMethod,))
                learningTime.append(time.time() - before)
                if fullyUpdateTau:
                    nextParams['tau'] = newParams['tau']
                progress(n, nll, vlb)
                learningDetails.append(learnDet)
                params = nextParams
                paramSeq.append(params)
            self.onlineParamUpdateMethod = onlineParamUpdateMethod
            self.seenTrialIdx = seenTrialIdx
        if say:
            print()

        self.xdim, self.ydim, self.trialDur, self.numTrials = xdim, ydim, trialDur, numTrials
        self.binSize, self.T, self.maxEMiter, self.EMmode = binSize, T, maxEMiter, EMmode
        self.inferenceMethod, self.initParams, self.paramSeq = inferenceMethod, initParams, paramSeq
        self.posteriorLikelihood = posteriorLikelihood
        self.variationalLowerBound = variationalLowerBound
        self.learningDetails = learningDetails
        self.optimParams = params
        self.infRes = infRes              # of the last batch processed in online EM (this rank's shard)
        self.processParamResults()
        self.performSpikeCountAnalysis()
        self.learningTime = np.asarray(learningTime)
        self.inferenceTime = np.asarray(inferenceTime)
        self.CdOptimMethod = CdOptimMethod
        self.optimLogLamb = optimLogLamb
        if extractAllTraj:
            self.extractTrajectories(method=inferenceMethod)
        if extractAllTraj_trueParams:
            self.extractTrajWithTrueParams(method=inferenceMethod)
        if getPredictionErr:
            self.leaveOneOutPrediction()

    # -------------------------------------------------------------------- post-fit helpers (host side)
    def extractTrajectories(self, method='laplace'):
        """funs/engine.py:523-532."""
        if method == 'laplace':
            self.infRes, self.nll_all_traj, _ = inference.laplace(self.experiment, self.optimParams, reducer=self._reducer)
        else:
            self.infRes, self.nll_all_traj, self.vlb_all_traj, _ = inference.dualVariational(
                self.experiment, self.optimParams, optimizeLogLambda=self.optimLogLamb, reducer=self._reducer)

    def extractTrajWithTrueParams(self, method='laplace'):
        """funs/engine.py:534-543."""
        tp = copy.deepcopy(self.experiment.params)
        if method == 'laplace':
            self.infRes_trueParams, self.nll_trueParams_all_traj, _ = inference.laplace(self.experiment, tp, reducer=self._reducer)
        else:
            (self.infRes_trueParams, self.nll_trueParams_all_traj, self.vlb_trueParams_all_traj, _) = \
                inference.dualVariational(self.experiment, tp, optimizeLogLambda=self.optimLogLamb, reducer=self._reducer)

    def performSpikeCountAnalysis(self):
        """funs/engine.py:483-512 (post-fit diagnostics on the host: model-implied vs observed count moments)."""
        ex = self.experiment
        raster = np.concatenate([np.asarray(t['Y'], dtype=np.float64) for t in ex.data], axis=1)
        ex.all_raster = raster
        E_y_init, E_yy_init = util.getMeanCovYfromParams(self.initParams, ex)
        E_y_opt, E_yy_opt = util.getMeanCovYfromParams(self.optimParams, ex)
        E_y_obs, E_yy_obs = np.mean(raster, 1), np.cov(raster)
        nrm = np.linalg.norm
        with np.errstate(all='ignore'):
            if hasattr(ex, 'params'):
                E_y_true, E_yy_true = util.getMeanCovYfromParams(ex.params, ex)
                self.E_y_true_params, self.E_yy_true_params = E_y_true, E_yy_true
                v = np.var(E_y_true)
                self.mean_err_optim_true = np.dot(E_y_true - E_y_opt, E_y_true - E_y_opt) / v / self.numTrials
                self.mean_err_init_true = np.dot(E_y_true - E_y_init, E_y_true - E_y_init) / v / self.numTrials
                self.cov_err_optim_true = nrm(E_yy_true - E_yy_opt) / nrm(E_yy_obs)
                self.cov_err_init_true = nrm(E_yy_true - E_yy_init) / nrm(E_yy_obs)
                self.JSdiv_cov_optim_true = util.JSLogdetDiv(E_yy_opt, E_yy_true)
                self.JSdiv_cov_init_true = util.JSLogdetDiv(E_yy_init, E_yy_true)
            self.E_y_init_params, self.E_y_optim_params = E_y_init, E_y_opt
            self.E_yy_init_params, self.E_yy_optim_params = E_yy_init, E_yy_opt
            self.E_y_obs, self.E_yy_obs = E_y_obs, E_yy_obs
            v = np.var(E_y_obs)
            self.mean_err_optim_obs = np.dot(E_y_obs - E_y_opt, E_y_obs - E_y_opt) / v / self.numTrials
            self.mean_err_init_obs = np.dot(E_y_obs - E_y_init, E_y_obs - E_y_init) / v / self.numTrials
            self.cov_err_optim_obs = nrm(E_yy_obs - E_yy_opt) / nrm(E_yy_obs)
            self.cov_err_init_obs = nrm(E_yy_obs - E_yy_init) / nrm(E_yy_obs)
            self.JSdiv_cov_optim_obs = util.JSLogdetDiv(E_yy_opt, E_yy_obs)
            self.JSdiv_cov_init_obs = util.JSLogdetDiv(E_yy_init, E_yy_obs)

    def leaveOneOutPrediction(self):
        """funs/engine.py:599-644: for every trial and neuron, the posterior mode from the other N-1 neurons and the
        predicted rate of the left-out one.  The reference runs R*N scipy fmin_ncg solves in a Python double loop;
        here they are one batch of R*N Laplace problems (pgpfa_loo_predict) over this rank's trials.
        Sets y_pred_mode (R, N, T) and pred_err_mode (sum of squared errors, all ranks)."""
        import torch
        from . import kernels as kn
        trials = inference.device_trials(self.experiment, self._reducer)
        p = inference.device_params(self.optimParams, self.T, self.binSize)
        R, N = trials.R, trials.N
        ymap = torch.arange(R, device="cuda", dtype=torch.int32).repeat_interleave(N).contiguous()
        excl = torch.arange(N, device="cuda", dtype=torch.int32).repeat(R).contiguous()
        ypred, err, _, st = kn.loo_predict(trials.y, p.C, p.d, p.Kinv, ymap, excl)
        if st["not_converged"]:
            raise RuntimeError("leave-one-out prediction: %d problems did not converge" % st["not_converged"])
        self.y_pred_mode = ypred.reshape(R, N, self.T).cpu().numpy()
        self.pred_err_mode = self._reducer.sum_scalar(float(err.sum()))

    def processParamResults(self):
        """funs/engine.py:545-597 (host-side bookkeeping over paramSeq; diagnostics, not hot path)."""
        n_it = self.maxEMiter
        self.tauSeq = np.zeros([self.xdim, n_it])
        self.expectedSpikeCountsEst = np.zeros([self.ydim, n_it])
        self.expectedSpikeCountsEstVar = np.zeros(n_it)
        self.CabsoluteValue = np.zeros(n_it)
        for i in range(n_it):
            C, d = np.asarray(self.paramSeq[i]['C']), np.asarray(self.paramSeq[i]['d'])
            self.tauSeq[:, i] = self.paramSeq[i]['tau']
            self.expectedSpikeCountsEst[:, i] = self.T * np.exp(0.5 * np.einsum('nk,nk->n', C, C) + d)
            self.expectedSpikeCountsEstVar[i] = np.var(self.expectedSpikeCountsEst[:, i])
            self.CabsoluteValue[i] = C.flatten().dot(C.flatten())
        s = np.zeros(self.ydim)
        for tr in range(self.numTrials):
            s = s + np.sum(self.experiment.data[tr]['Y'], 1)
        self.sampleMeanSpikeCounts = s / self.numTrials
        self.sampleMeanSpikeCountsVar = np.var(self.sampleMeanSpikeCounts)
        with np.errstate(divide='ignore', invalid='ignore'):
            self.varESpkCountSampleMean_Ratios = self.expectedSpikeCountsEstVar / self.sampleMeanSpikeCountsVar
            self.meanSquaredErrorOverTrueVariance_SM = [
                1 / self.numTrials * np.dot(self.expectedSpikeCountsEst[:, i] - self.sampleMeanSpikeCounts,
                                            self.expectedSpikeCountsEst[:, i] - self.sampleMeanSpikeCounts)
                / self.sampleMeanSpikeCountsVar for i in range(n_it)]
        if hasattr(self.experiment, 'params'):
            Ct, dt = np.asarray(self.experiment.params['C']), np.asarray(self.experiment.params['d'])
            self.expectedSpikeCountsTrue = self.T * np.exp(0.5 * np.einsum('nk,nk->n', Ct, Ct) + dt)
            self.expectedSpikeCountsTrueVar = np.var(self.expectedSpikeCountsTrue)
            self.varESpkCountTrue_Ratios = self.expectedSpikeCountsEstVar / self.expectedSpikeCountsTrueVar
