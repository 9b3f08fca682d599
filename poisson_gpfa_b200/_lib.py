"""ctypes binding of libpgpfa_b200.so (include/pgpfa_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every compute
entry point needs a CUDA device (``pgpfa_create`` fails with PGPFA_ERR_NO_DEVICE otherwise).
torch is used only to own device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PGPFA_LIB: development switch for A/B builds of the library (kernel variants under ncu); never set in tests / bench
LIB_PATH = os.environ.get("PGPFA_LIB") or os.path.join(_HERE, "csrc", "libpgpfa_b200.so")

c_int, c_ll, c_dbl, c_void_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p
P = c_void_p  # device pointers and streams travel as integers

# name -> (restype, argtypes); must list every symbol declared in include/pgpfa_b200.h
SIGNATURES = {
    "pgpfa_abi_version": (c_int, []),
    "pgpfa_create": (c_int, [ctypes.POINTER(c_void_p)]),
    "pgpfa_destroy": (c_int, [c_void_p]),
    "pgpfa_error_string": (ctypes.c_char_p, [c_int]),
    "pgpfa_last_cuda_error": (ctypes.c_char_p, []),
    "pgpfa_launch_count": (c_ll, []),
    "pgpfa_host_sync_count": (c_ll, [c_void_p]),
    "pgpfa_set_loop_depth": (c_int, [c_void_p, c_int]),
    "pgpfa_throttle_wait_count": (c_ll, [c_void_p]),
    "pgpfa_stream_wait_means": (c_int, [c_void_p, c_void_p]),
    "pgpfa_prior_lowrank": (c_int, [P, c_int, c_int, c_dbl, c_dbl, P, P, P, P]),
    "pgpfa_laplace_solve_lowrank": (c_int, [c_void_p, P, P, P, P, P, P, P, c_dbl, P, c_int, c_int, c_int, c_int, c_dbl, c_int,
                                            c_int, P, P, P, P, P, P, P, c_ll, P, P]),
    "pgpfa_set_profiling": (c_int, [c_void_p, c_int]),
    "pgpfa_get_profile": (c_int, [c_void_p, P, P, P]),
    "pgpfa_map": (c_int, [c_int, c_ll, P, P, c_dbl, P, P]),
    "pgpfa_bin_spikes": (c_int, [P, P, P, c_dbl, c_int, c_int, c_int, P, P]),
    "pgpfa_sample_normal": (c_int, [P, c_ll, ctypes.c_ulonglong, P]),
    "pgpfa_sample_poisson": (c_int, [P, P, P, c_int, c_int, c_int, c_int, ctypes.c_ulonglong, P, P]),
    "pgpfa_make_K": (c_int, [P, c_int, c_int, c_dbl, c_dbl, P, P]),
    "pgpfa_make_K_big": (c_int, [P, c_int, c_int, P, P]),
    "pgpfa_make_K_gamma": (c_int, [P, c_int, c_int, c_dbl, P, P, P]),
    "pgpfa_spd_inverse_workspace_bytes": (c_ll, [c_int, c_int]),
    "pgpfa_spd_inverse_batched": (c_int, [P, c_int, c_int, P, P, P, P, c_ll, P]),
    "pgpfa_tiles_bytes": (c_ll, [c_int]),
    "pgpfa_dinv_bytes": (c_ll, [c_int]),
    "pgpfa_potrf_dense": (c_int, [P, c_int, c_int, P, P, P, P, P]),
    "pgpfa_potrf_posterior": (c_int, [P, P, c_dbl, c_int, c_int, c_int, P, P, P, P, P]),
    "pgpfa_potrs": (c_int, [P, P, P, c_dbl, c_int, c_int, P, P]),
    "pgpfa_trtri": (c_int, [P, P, P, c_int, c_int, P]),
    "pgpfa_potri_dense": (c_int, [P, c_int, c_int, P, P, c_ll, P]),
    "pgpfa_cov_slices": (c_int, [P, c_int, c_int, c_int, P, P, P, c_ll, P]),
    "pgpfa_logdet": (c_int, [P, c_int, c_int, P, P]),
    "pgpfa_tiles_to_dense": (c_int, [P, c_int, c_int, c_int, P, P]),
    "pgpfa_prior_apply": (c_int, [P, P, c_int, c_int, c_int, P, P]),
    "pgpfa_laplace_eval": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P]),
    "pgpfa_hessian_dense": (c_int, [P, P, c_dbl, c_int, c_int, c_int, P, P]),
    "pgpfa_laplace_workspace_bytes": (c_ll, [c_int, c_int, c_int, c_int]),
    "pgpfa_laplace_solve": (c_int, [c_void_p, P, P, P, P, P, c_int, c_int, c_int, c_int, c_dbl, c_int, c_int,
                                    P, P, P, P, P, P, P, c_ll, P, P]),
    "pgpfa_loo_predict": (c_int, [c_void_p, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_dbl, c_int, P, P, P, P, P,
                                  c_ll, P, P]),
    "pgpfa_dualvi_workspace_bytes": (c_ll, [c_int, c_int, c_int, c_int]),
    "pgpfa_dualvi_eval": (c_int, [c_void_p, P, P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_ll, P]),
    "pgpfa_dualvi_solve": (c_int, [c_void_p, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_dbl, c_int,
                                   P, P, P, P, P, P, P, P, P, P, c_ll, P, P]),
    "pgpfa_rate_blocks": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "pgpfa_dualvi_init_from_lambda": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, c_ll, P]),
    "pgpfa_pautosum": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "pgpfa_mstep_cd_nstats": (c_int, [c_int]),
    "pgpfa_mstep_cd_workspace_bytes": (c_ll, [c_int, c_int]),
    "pgpfa_mstep_cd_stats": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, c_ll, P]),
    "pgpfa_mstep_cd_update": (c_int, [P, c_dbl, c_dbl, P, P, P, P, P, P, P, P, P, c_int, c_dbl, c_int, c_int, P, c_int, P]),
    "pgpfa_mstep_cd_stats_gated": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, c_ll, P, P]),
    "pgpfa_mstep_cd_solve": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_dbl, c_dbl, P, P, P, P, P, P, P, P, P, P, P,
                                     c_int, c_int, c_dbl, P, c_ll, P]),
    "pgpfa_tau_solve_workspace_bytes": (c_ll, [c_int, c_int, c_int]),
    "pgpfa_mstep_tau_solve": (c_int, [P, P, c_dbl, c_int, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_int, c_int, c_int, P, P, P,
                                      P, c_ll, P]),
    "pgpfa_tau_eval_workspace_bytes": (c_ll, [c_int, c_int]),
    "pgpfa_tau_eval": (c_int, [P, P, c_dbl, c_int, c_int, c_dbl, c_dbl, P, c_dbl, P, P, P, c_ll, P]),
}

ERR_NOT_CONVERGED = 5


class PgpfaError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib.pgpfa_error_string(code).decode()
        if code == 1:
            msg += ": " + lib.pgpfa_last_cuda_error().decode()
        super().__init__("%s failed: %s" % (where, msg))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libpgpfa_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C poisson_gpfa_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    dll = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(dll, name)       # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return dll


lib = _load()

_handles = {}


def handle():
    """One library handle per device (created lazily; needs a GPU)."""
    if not torch.cuda.is_available():
        raise RuntimeError("poisson_gpfa_b200 needs a CUDA device: the hot path has no CPU fallback")
    dev = torch.cuda.current_device()
    if dev not in _handles:
        h = c_void_p()
        rc = lib.pgpfa_create(ctypes.byref(h))
        if rc != 0:
            raise PgpfaError(rc, "pgpfa_create")
        _handles[dev] = h
    return _handles[dev]


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return t.data_ptr()


def call(name, *args, allow=()):
    rc = getattr(lib, name)(*args)
    if rc != 0 and rc not in allow:
        raise PgpfaError(rc, name)
    return rc


# host reads of device results made by the Python layer (each one waits for the stream): bench.py reports them per step
host_reads = [0]


def to_host(t):
    """Device tensor -> numpy, counted as one host synchronisation."""
    host_reads[0] += 1
    return t.detach().cpu().numpy()


def host_sync_count():
    """Synchronisations so far: inside the library's drivers (this device's handle) + host reads of the Python layer."""
    return int(lib.pgpfa_host_sync_count(handle())) + host_reads[0]


def dev_f64(a):
    """numpy / tensor -> contiguous float64 CUDA tensor."""
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    import numpy as np
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).cuda()


def empty(*shape, dtype=torch.float64):
    return torch.empty(*shape, dtype=dtype, device="cuda")


def workspace(nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
