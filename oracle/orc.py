"""ctypes mirror of oracle/oracle_config.h — TEST INFRASTRUCTURE ONLY.

Loads oracle/_ref/libmci_ref.so (the unmodified reference behind oracle/ref_harness.cpp; only exists where it was
built from /root/reference) and oracle/libmci_oracle.so (the plain-C restatement). Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this module.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))

MAXDIM, MAXOBS, MAXTYPES, MAXOBSDIM = 1024, 8, 8, 2048

PDF_NONE, PDF_GAUSS3D, PDF_GAUSS, PDF_EXP1D, PDF_EXPND, PDF_NORMLINE = range(6)
(OBS_XSQUARED, OBS_GAUSSXSQUARED, OBS_XYZSQUARED, OBS_X1D, OBS_XND, OBS_UPDXND, OBS_CONSTVAL, OBS_POLYNOM,
 OBS_X2SUM, OBS_X2, OBS_PARABOLA, OBS_NORMPARABOLA, OBS_DEPENDENT) = range(1, 14)
MOVE_ALL, MOVE_VEC, MOVE_MULTISTEP = range(3)
MOVE_USER_DRIFT = 3  # reference harness only: a user-defined TrialMoveInterface subclass (oracle/ref_harness.cpp: HarnessDriftMove), srrd_par = (drift,)
SRRD_UNIFORM, SRRD_GAUSSIAN = range(2)
EST_NOOP, EST_UNCORRELATED, EST_CORRELATED, EST_FCBLOCKER, EST_MJBLOCKER = range(5)
EST_BLOCK = 100
DOMAIN_UNBOUND, DOMAIN_ORTHO = range(2)

OBS_NOBS = {  # nobs as a function of ndim
    OBS_XSQUARED: lambda nd: 1, OBS_GAUSSXSQUARED: lambda nd: 1, OBS_XYZSQUARED: lambda nd: 3, OBS_X1D: lambda nd: 1,
    OBS_XND: lambda nd: nd, OBS_UPDXND: lambda nd: nd, OBS_CONSTVAL: lambda nd: 1, OBS_POLYNOM: lambda nd: 1,
    OBS_X2SUM: lambda nd: 1, OBS_X2: lambda nd: nd, OBS_PARABOLA: lambda nd: 1, OBS_NORMPARABOLA: lambda nd: 1, OBS_DEPENDENT: lambda nd: 2,
}


class Obs(C.Structure):
    _fields_ = [("obs_id", C.c_int32), ("blocksize", C.c_int32), ("nskip", C.c_int32), ("flag_equil", C.c_int32),
                ("estim_type", C.c_int32)]


class Config(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("seed", C.c_uint64), ("pdf_id", C.c_int32), ("move_type", C.c_int32), ("srrd", C.c_int32),
        ("veclen", C.c_int32), ("ntypes", C.c_int32), ("type_ends", C.c_int32 * MAXTYPES), ("steps", C.c_double * MAXTYPES),
        ("ms_nsteps", C.c_int32), ("ms_sub_pdf_id", C.c_int32), ("domain", C.c_int32), ("lb", C.c_double * MAXDIM),
        ("ub", C.c_double * MAXDIM), ("x0", C.c_double * MAXDIM), ("nobs", C.c_int32), ("obs", Obs * MAXOBS),
        ("nfind", C.c_int32), ("ndecorr", C.c_int64), ("target_acc", C.c_double), ("nmc", C.c_int64),
        ("do_find", C.c_int32), ("do_decorr", C.c_int32), ("nranks_for_minstat", C.c_int32),
        ("srrd_npar", C.c_int32), ("srrd_par", C.c_double * 2),
    ]


class Result(C.Structure):
    _fields_ = [("nobsdim", C.c_int32), ("avg", C.c_double * MAXOBSDIM), ("err", C.c_double * MAXOBSDIM),
                ("acc_rate", C.c_double), ("x_final", C.c_double * MAXDIM), ("steps_final", C.c_double * MAXTYPES),
                ("n_acc", C.c_int64), ("n_rej", C.c_int64)]


class Trace(C.Structure):
    _fields_ = [("cap_steps", C.c_int64), ("cap_draws", C.c_int64), ("accepted", C.POINTER(C.c_uint8)),
                ("draws", C.POINTER(C.c_double)), ("n_steps", C.c_int64), ("n_draws", C.c_int64)]


def default_estim(blocksize, flag_correlated=None):
    """addObservable(obs, blocksize, nskip) default mapping: include/mci/MCIntegrator.hpp:161-164, Factories.hpp:61-71."""
    if flag_correlated is None:
        flag_correlated = blocksize == 1
    if flag_correlated:
        return EST_CORRELATED
    return EST_UNCORRELATED if blocksize > 0 else EST_NOOP


def make_config(ndim, seed, pdf_id, obs, nmc, *, move_type=MOVE_ALL, srrd=SRRD_UNIFORM, veclen=0, ntypes=1, type_ends=None,
                steps=(0.05,), ms_nsteps=0, ms_sub_pdf_id=PDF_NONE, lb=None, ub=None, x0=None, nfind=-50, ndecorr=-10000,
                target_acc=0.5, do_find=False, do_decorr=False, nranks=1, srrd_par=()):
    """obs: list of (obs_id, blocksize, nskip[, flag_equil[, estim_type]]) tuples (defaults as MCIntegrator.hpp:161-164)."""
    c = Config()
    c.ndim, c.seed, c.pdf_id, c.move_type, c.srrd, c.veclen, c.ntypes = ndim, seed, pdf_id, move_type, srrd, veclen, ntypes
    for i, t in enumerate(type_ends or []):
        c.type_ends[i] = t
    steps = list(steps)
    for i in range(max(1, ntypes)):
        c.steps[i] = steps[i] if i < len(steps) else steps[-1]
    c.ms_nsteps, c.ms_sub_pdf_id = ms_nsteps, ms_sub_pdf_id
    if lb is not None:
        c.domain = DOMAIN_ORTHO
        for i in range(ndim):
            c.lb[i] = lb[i] if hasattr(lb, "__len__") else lb
            c.ub[i] = ub[i] if hasattr(ub, "__len__") else ub
    for i in range(ndim):
        c.x0[i] = 0.0 if x0 is None else x0[i]
    c.nobs = len(obs)
    for i, o in enumerate(obs):
        o = tuple(o)
        oid, bs, ns = o[0], o[1], o[2]
        bs, ns = max(0, bs), max(1, ns)
        fe = o[3] if len(o) > 3 else (bs > 0)
        et = o[4] if len(o) > 4 else default_estim(bs)
        c.obs[i] = Obs(oid, bs, ns, int(fe), et)
    c.nfind, c.ndecorr, c.target_acc, c.nmc = nfind, ndecorr, target_acc, nmc
    c.do_find, c.do_decorr, c.nranks_for_minstat = int(do_find), int(do_decorr), nranks
    c.srrd_npar = len(srrd_par)
    for i, v in enumerate(srrd_par):
        c.srrd_par[i] = v
    return c


def _load(path, prefix):
    lib = C.CDLL(path)
    run = getattr(lib, prefix + "_run")
    run.restype = C.c_int
    run.argtypes = [C.POINTER(Config), C.POINTER(Result), C.POINTER(Trace)]
    getattr(lib, prefix + "_last_error").restype = C.c_char_p
    est = getattr(lib, prefix + "_estimate")
    est.restype = C.c_int
    est.argtypes = [C.c_int, C.c_int64, C.c_int, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    tw = getattr(lib, prefix + "_testwalk")
    tw.restype = C.c_double
    tw.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
    ac = getattr(lib, prefix + "_accumulate")
    ac.restype = C.c_int64
    ac.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                   C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    return lib


REF_PATH = os.path.join(HERE, "_ref", "libmci_ref.so")
ORACLE_PATH = os.path.join(HERE, "libmci_oracle.so")


class Engine:
    """Uniform Python face of either the compiled reference (prefix mciref) or the C restatement (prefix mcio)."""

    def __init__(self, path, prefix):
        self.lib = _load(path, prefix)
        self.prefix = prefix

    def _f(self, name):
        return getattr(self.lib, self.prefix + "_" + name)

    def run(self, cfg, trace=False):
        import numpy as np
        res = Result()
        tr = None
        if trace:
            tr = Trace()
            ndraw_cap = int(cfg.nmc) * (3 * cfg.ndim + 2) * max(1, 1) + 16
            acc = np.zeros(max(1, int(cfg.nmc)), dtype=np.uint8)
            draws = np.zeros(ndraw_cap, dtype=np.float64)
            tr.cap_steps, tr.cap_draws = int(cfg.nmc), ndraw_cap
            tr.accepted = acc.ctypes.data_as(C.POINTER(C.c_uint8))
            tr.draws = draws.ctypes.data_as(C.POINTER(C.c_double))
        rc = self._f("run")(C.byref(cfg), C.byref(res), C.byref(tr) if tr is not None else None)
        if rc != 0:
            raise RuntimeError(self._f("last_error")().decode())
        n = res.nobsdim
        out = {"avg": list(res.avg[:n]), "err": list(res.err[:n]), "acc_rate": res.acc_rate,
               "x_final": list(res.x_final[:cfg.ndim]), "steps_final": list(res.steps_final[:max(1, cfg.ntypes)]),
               "n_acc": res.n_acc, "n_rej": res.n_rej}
        if trace:
            out["accepted"] = acc[:tr.n_steps].copy()
            out["draws"] = draws[:max(0, tr.n_draws)].copy()
            out["n_draws"] = tr.n_draws
        return out

    def estimate(self, estim_type, x, nblocks=0):
        import numpy as np
        x = np.ascontiguousarray(x, dtype=np.float64)
        n, ndim = (x.shape[0], 1) if x.ndim == 1 else x.shape
        avg = np.zeros(ndim)
        err = np.zeros(ndim)
        dp = C.POINTER(C.c_double)
        rc = self._f("estimate")(estim_type, n, ndim, x.ctypes.data_as(dp), nblocks, avg.ctypes.data_as(dp), err.ctypes.data_as(dp))
        if rc != 0:
            raise RuntimeError(self._f("last_error")().decode())
        return avg, err

    def testwalk(self, pdf, nmc, ndim, step, change_prob, seed):
        import numpy as np
        datax = np.zeros((nmc, ndim))
        datacc = np.zeros(nmc, dtype=np.uint8)
        nchanged = np.zeros(nmc, dtype=np.int32)
        cidx = np.zeros((nmc, ndim), dtype=np.int32)
        rate = self._f("testwalk")(pdf, nmc, ndim, step, change_prob, seed, datax.ctypes.data_as(C.POINTER(C.c_double)),
                                   datacc.ctypes.data_as(C.POINTER(C.c_uint8)), nchanged.ctypes.data_as(C.POINTER(C.c_int)),
                                   cidx.ctypes.data_as(C.POINTER(C.c_int)))
        return datax, datacc, nchanged, cidx, rate

    def accumulate(self, obs_id, ndim, blocksize, nskip, datax, datacc, nchanged, cidx):
        import numpy as np
        nmc = datax.shape[0]
        dp, up, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int)
        args = (obs_id, ndim, blocksize, nskip, nmc, datax.ctypes.data_as(dp), datacc.ctypes.data_as(up),
                nchanged.ctypes.data_as(ip), cidx.ctypes.data_as(ip))
        nstore = self._f("accumulate")(*args, None)
        if nstore < 0:
            raise RuntimeError(self._f("last_error")().decode())
        nobs = OBS_NOBS[obs_id](ndim)
        out = np.zeros((nstore, nobs))
        self._f("accumulate")(*args, out.ctypes.data_as(dp))
        return out


def have_ref():
    return os.path.exists(REF_PATH)


def ref_timing_path():
    """The timing build of the unmodified reference best matched to THIS host's CPU (oracle/Makefile: -O3 -march=x86-64-v4 / -v3 unity builds,
    the portable spelling of the reference's own -march=native), falling back to the parity build. Returns (path, flags description)."""
    flags = set()
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = set(line.split(":", 1)[1].split())
                break
    except OSError:
        pass
    v3 = {"avx2", "fma", "bmi2", "bmi1", "movbe", "f16c", "abm"} <= flags
    v4 = v3 and {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
    for ok, name, desc in ((v4, "libmci_ref_v4.so", "-O3 -march=x86-64-v4, all reference sources as one translation unit (whole-program, in place of -flto)"), (v3, "libmci_ref_v3.so", "-O3 -march=x86-64-v3, all reference sources as one translation unit (whole-program, in place of -flto)")):
        path = os.path.join(HERE, "_ref", name)
        if ok and os.path.exists(path):
            return path, desc
    return REF_PATH, "-O3 -ffp-contract=off (generic x86-64)"


def ref():
    return Engine(REF_PATH, "mciref")


def oracle():
    return Engine(ORACLE_PATH, "mcio")
