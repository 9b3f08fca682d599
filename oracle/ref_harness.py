"""Reference harness — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Imports the UNMODIFIED reference from /root/reference (it exists only in the build
container, never on the GPU box) so that (a) the numpy restatement in
``oracle/pgpfa_oracle.py`` can be validated against the real thing and (b) golden
vectors for ``tests/golden/`` can be generated (``oracle/make_golden.py``).

Three harness-side shims are needed, none of which touches the
reference's arithmetic:
  1. matplotlib / mpl_toolkits are not installed -> stub modules
     (imports at funs/inference.py:7, funs/learning.py:12, funs/util.py:12-14,
     funs/engine.py:17-19).
  2. statsmodels is not installed -> stub with ``tools.numdiff._get_epsilon``
     (the only symbol executed: funs/util.py:419) restating statsmodels 0.6.1.
  3. scipy >= 1.15 dropped ``disp`` from ``fmin_l_bfgs_b``
     (call sites funs/inference.py:316-324, :391-396) -> wrapper drops it.
The same wrapper layer optionally injects TIGHT optimiser tolerances (the parity
protocol of DESIGN.md: both sides converged).
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import scipy.optimize as _sopt

REFERENCE_ROOT = os.environ.get("PGPFA_REFERENCE_ROOT", "/root/reference")

_EPS = np.finfo(float).eps


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "funs"))


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless callable/namespace."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _Anything(self.__name__ + "." + name)
        setattr(self, name, obj)
        return obj

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")

    def __iter__(self):
        return iter(())


def _get_epsilon(x, s, epsilon, n):
    # statsmodels 0.6.1 tools/numdiff.py::_get_epsilon (restated)
    if epsilon is None:
        h = _EPS ** (1.0 / s) * np.maximum(np.abs(x), 0.1)
    else:
        if np.isscalar(epsilon):
            h = np.empty(n)
            h.fill(epsilon)
        else:
            h = np.asarray(epsilon)
            if h.shape != np.shape(x):
                raise ValueError("If h is not a scalar it must have the same shape as x.")
    return h


def _install_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.pylab", "matplotlib.gridspec",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].pylab = sys.modules["matplotlib.pylab"]
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    sys.modules["mpl_toolkits"].mplot3d = sys.modules["mpl_toolkits.mplot3d"]
    sys.modules["mpl_toolkits"].axes_grid1 = sys.modules["mpl_toolkits.axes_grid1"]
    if "statsmodels" not in sys.modules:
        sm = types.ModuleType("statsmodels")
        tools = types.ModuleType("statsmodels.tools")
        nd = types.ModuleType("statsmodels.tools.numdiff")
        nd._get_epsilon = _get_epsilon
        nd.EPS = _EPS
        nd.approx_fprime = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
        nd.approx_hess = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
        nd.Jacobian = None
        sm.tools = tools
        tools.numdiff = nd
        sys.modules["statsmodels"] = sm
        sys.modules["statsmodels.tools"] = tools
        sys.modules["statsmodels.tools.numdiff"] = nd


# ----------------------------------------------------------------------------
# scipy.optimize wrappers (disp shim + optional tight tolerances)
# ----------------------------------------------------------------------------
_orig_minimize = _sopt.minimize
_orig_lbfgsb = _sopt.fmin_l_bfgs_b
_TIGHT = {"on": False}


def _minimize_wrapper(fun, x0, args=(), method=None, jac=None, hess=None, hessp=None,
                      bounds=None, constraints=(), tol=None, callback=None, options=None):
    options = dict(options or {})
    m = (method or "BFGS")
    if m == "TNC" and "maxiter" in options:
        # reference passes 'maxiter' to TNC (funs/learning.py:130); scipy only warns. Map it.
        mi = options.pop("maxiter")
        if mi is not None:
            options["maxfun"] = mi
    if options.get("maxiter", 0) is None:
        options.pop("maxiter")
    if _TIGHT["on"]:
        if m == "Newton-CG":
            options["xtol"] = 1e-14
        elif m == "TNC":
            options.update(gtol=1e-13, ftol=1e-16, xtol=1e-16, maxfun=200000)
        elif m == "BFGS":
            options.update(gtol=1e-11)
        elif m == "L-BFGS-B":
            options.update(gtol=1e-13, ftol=1e-16, maxfun=200000, maxiter=200000)
    return _orig_minimize(fun, x0, args=args, method=method, jac=jac, hess=hess, hessp=hessp,
                          bounds=bounds, constraints=constraints, tol=tol, callback=callback,
                          options=options)


def _lbfgsb_wrapper(func, x0, fprime=None, args=(), approx_grad=0, bounds=None, m=10,
                    factr=1e7, pgtol=1e-5, epsilon=1e-8, iprint=-1, maxfun=15000,
                    maxiter=15000, disp=None, callback=None, maxls=20):
    if _TIGHT["on"]:
        factr, pgtol, maxfun, maxiter = 10.0, 1e-12, 200000, 200000
    return _orig_lbfgsb(func, x0, fprime=fprime, args=args, approx_grad=approx_grad,
                        bounds=bounds, m=m, factr=factr, pgtol=pgtol, epsilon=epsilon,
                        maxfun=maxfun, maxiter=maxiter, callback=callback, maxls=maxls)


@contextlib.contextmanager
def tight_tolerances(on=True):
    """Parity protocol: run the reference's scipy optimisers to (near) machine convergence."""
    prev = _TIGHT["on"]
    _TIGHT["on"] = bool(on)
    try:
        yield
    finally:
        _TIGHT["on"] = prev


_ref = {}


def load_reference():
    """Returns a namespace with the reference modules (util, inference, learning, engine)."""
    if _ref:
        return types.SimpleNamespace(**_ref)
    if not reference_available():
        raise RuntimeError("reference tree not found at %s (it only exists in the build container)"
                           % REFERENCE_ROOT)
    _install_stubs()
    _sopt.minimize = _minimize_wrapper
    _sopt.fmin_l_bfgs_b = _lbfgsb_wrapper
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "funs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    cwd = os.getcwd()
    try:
        os.chdir(REFERENCE_ROOT)          # funs/__init__.py:9 appends cwd+'/funs/' to sys.path
        with contextlib.redirect_stdout(io.StringIO()):
            import funs  # noqa: F401
            import util, inference, learning, engine  # noqa: E401  (top-level copies, engine.py:9-11)
    finally:
        os.chdir(cwd)
    _ref.update(util=util, inference=inference, learning=learning, engine=engine)
    return types.SimpleNamespace(**_ref)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
