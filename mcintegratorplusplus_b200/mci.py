"""Python mirror of the reference's `mci::MCI` interface on top of the C-ABI (include/mcig.h).

Used by the tests and bench.py so that they read like the reference's own tests (test/ut2..ut5, benchmark/*):

    mci = MCI(3); mci.setSeed(5649871); mci.addSamplingFunction(ThreeDimGaussianPDF())
    mci.addObservable(XSquared(), 0, 1); mci.setMRT2Step(1.0); avg, err = mci.integrate(100000, False, False)

Method names, argument meaning, defaults and error behaviour follow include/mci/MCIntegrator.hpp:93-225 of the
reference; everything numeric happens in libmcig.so on the GPU (no fallback).
"""
import ctypes as C
from enum import IntEnum

import numpy as np

from . import _capi

_dp = C.POINTER(C.c_double)


class MoveType(IntEnum):  # include/mci/Factories.hpp:108-114
    All = 0
    Vec = 1
    MultiStep = 2


class SRRDType(IntEnum):  # include/mci/Factories.hpp:119-133
    Uniform = 0
    Gaussian = 1
    Student = 2
    Cauchy = 3
    Exponential = 4
    Gamma = 5
    Weibull = 6
    Lognormal = 7
    Chisq = 8
    Fisher = 9


class EstimatorType(IntEnum):  # include/mci/Factories.hpp:52-59
    Noop = 0
    Uncorrelated = 1
    Correlated = 2
    FCBlocker = 3
    MJBlocker = 4


class RngMode(IntEnum):
    Philox32 = 0
    Philox53 = 1
    Replay = 2


def selectEstimatorType(flag_correlated, flag_error=True):  # include/mci/Factories.hpp:61-71
    if flag_correlated:
        if not flag_error:
            raise ValueError("[selectEstimatorType] Error calculation is set off, but correlated error estimation is set on.")
        return EstimatorType.Correlated
    return EstimatorType.Uncorrelated if flag_error else EstimatorType.Noop


class Plugin:
    """A sampling function / observable living on the device: name of a registered functor + its parameters."""
    kind = None

    def __init__(self, name, par=()):
        self.name = name
        self.par = [float(v) for v in par]

    def plugin_id(self):
        pid = _capi.lib().mcig_lookup_plugin(self.kind, self.name.encode())
        if pid < 0:
            raise KeyError("plugin %r is not registered" % self.name)
        return pid


class SamplingFunction(Plugin):
    kind = 0


class Observable(Plugin):
    kind = 1


class StepCallback(Plugin):
    """MCI::setCallback as a device functor (see include/mcig.h: mcig_set_callback)."""
    kind = 2


class Move(Plugin):
    """A user-defined trial move as a device functor (see include/mcig.h: mcig_set_move_plugin)."""
    kind = 4


class Domain(Plugin):
    """A user-defined (separable) domain as a device functor (see include/mcig.h: mcig_set_domain_plugin)."""
    kind = 3


def register_plugin(kind, name, type_expr, source, ndim=0, nvalues=0, npar=0, has_update=False, elementwise=False, log_acceptance=False,
                    dependent=False, proto_element=False, sum_acceptance=False):
    """Register a user functor (CUDA C++ source, see csrc/device/mcig_functors.cuh for the contract).
    kind: 0 sampling function, 1 observable (dependent=True: DependentObservableInterface), 2 step callback."""
    flags = ((1 if has_update else 0) | (2 if elementwise else 0) | (4 if log_acceptance else 0) | (8 if dependent else 0) | (16 if proto_element else 0) |
             (32 if sum_acceptance else 0))
    pid = _capi.lib().mcig_register_plugin(kind, name.encode(), type_expr.encode(), (source or "").encode(), ndim, nvalues, npar, flags)
    if pid < 0:
        raise _capi.McigError(-pid, _capi.lib().mcig_last_error().decode())
    return pid


# the reference's fixtures (test/common/TestMCIFunctions.hpp, examples/common/ExampleFunctions.hpp), pre-registered
def ThreeDimGaussianPDF(): return SamplingFunction("ThreeDimGaussianPDF")
def Gauss(ndim): return SamplingFunction("Gauss")
def Exp1DPDF(): return SamplingFunction("Exp1DPDF")
def ExpNDPDF(ndim): return SamplingFunction("ExpNDPDF")
def NormalizedLine(): return SamplingFunction("NormalizedLine")
def XSquared(): return Observable("XSquared")
def GaussXSquared(): return Observable("GaussXSquared")
def XYZSquared(): return Observable("XYZSquared")
def X1D(): return Observable("X1D")
def XND(ndim): return Observable("XND")
def UpdateableXND(ndim): return Observable("UpdateableXND")
def Constval(ndim): return Observable("Constval")
def Polynom(ndim): return Observable("Polynom")
def X2Sum(ndim): return Observable("X2Sum")
def X2(ndim): return Observable("X2")
def Parabola(): return Observable("Parabola")
def NormalizedParabola(): return Observable("NormalizedParabola")


def _darr(values):
    a = np.ascontiguousarray(values, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class MCI:
    def __init__(self, ndim, device=None):
        self._lib = _capi.lib()
        self._ctx = self._lib.mcig_create(int(ndim))
        if not self._ctx:
            raise _capi.McigError(1, self._lib.mcig_last_error().decode())
        self._ndim = int(ndim)
        self._cb = None
        if device is not None:
            _capi.check(self._lib.mcig_set_device(self._ctx, int(device)))

    def __del__(self):
        try:
            if getattr(self, "_ctx", None):
                self._lib.mcig_destroy(self._ctx)
                self._ctx = None
        except Exception:
            pass

    # --- setters (include/mci/MCIntegrator.hpp:96-123)
    def setSeed(self, seed): _capi.check(self._lib.mcig_set_seed(self._ctx, int(seed)))

    def setWalkerSeeds(self, seeds):
        a = np.ascontiguousarray(seeds, dtype=np.uint64)
        _capi.check(self._lib.mcig_set_walker_seeds(self._ctx, a.ctypes.data_as(C.POINTER(C.c_uint64)), len(a)))

    def setRngMode(self, mode): _capi.check(self._lib.mcig_set_rng_mode(self._ctx, int(mode)))

    def setNWalkers(self, n, global_offset=0, total=None):
        _capi.check(self._lib.mcig_set_walkers(self._ctx, int(n), int(global_offset), int(total if total is not None else n + global_offset)))

    def getNWalkers(self): return self._lib.mcig_get_walkers(self._ctx)

    def setX(self, *args):
        if len(args) == 2:  # setX(i, val)
            x = self.getX()
            x[int(args[0])] = float(args[1])
        else:
            x = np.asarray(args[0], dtype=np.float64)
        a, p = _darr(x)
        _capi.check(self._lib.mcig_set_x(self._ctx, p))

    def setXWalkers(self, x):
        a, p = _darr(x)
        assert a.shape == (self.getNWalkers(), self._ndim)
        _capi.check(self._lib.mcig_set_x_walkers(self._ctx, p))

    def getX(self, i=None, walker=0):
        a = np.zeros(self._ndim)
        _capi.check(self._lib.mcig_get_x(self._ctx, int(walker), a.ctypes.data_as(_dp)))
        return a if i is None else float(a[i])

    def setMRT2Step(self, *args):
        n = self._lib.mcig_get_nsteps_sizes(self._ctx)
        if len(args) == 2:
            _capi.check(self._lib.mcig_set_step(self._ctx, int(args[0]), float(args[1])))
        elif np.isscalar(args[0]):
            for i in range(n):
                _capi.check(self._lib.mcig_set_step(self._ctx, i, float(args[0])))
        else:
            for i in range(n):
                _capi.check(self._lib.mcig_set_step(self._ctx, i, float(args[0][i])))

    def getMRT2Step(self, i=0): return self._lib.mcig_get_step(self._ctx, int(i))

    def setTargetAcceptanceRate(self, r):
        self._target = float(r)
        self._push_autotune()

    def setNfindMRT2Iterations(self, n):
        self._nfind = int(n)
        self._push_autotune()

    def setNdecorrelationSteps(self, n):
        self._ndecorr = int(n)
        self._push_autotune()

    _target, _nfind, _ndecorr = 0.5, -50, -10000  # src/MCIntegrator.cpp:637-639

    def _push_autotune(self):
        _capi.check(self._lib.mcig_set_autotune(self._ctx, self._nfind, self._ndecorr, self._target))

    def getTargetAcceptanceRate(self): return self._target
    def getNfindMRT2Iterations(self): return self._nfind
    def getNdecorrelationSteps(self): return self._ndecorr

    # --- domain (include/mci/MCIntegrator.hpp:129-137)
    def resetDomain(self): _capi.check(self._lib.mcig_set_domain_unbound(self._ctx))

    def setDomain(self, domain, sizes, volume):
        """setDomain(Domain functor, lengths of the ndim dimensions, volume (0 = infinite)): MCI::setDomain with a user-defined DomainInterface"""
        a, p = _darr(domain.par)
        sz = np.full(self._ndim, sizes, dtype=np.float64) if np.isscalar(sizes) else np.asarray(sizes, dtype=np.float64)
        b, pb = _darr(sz)
        _capi.check(self._lib.mcig_set_domain_plugin(self._ctx, domain.plugin_id(), p, len(domain.par), pb, float(volume)))

    def setIRange(self, lbound, ubound):
        lb = np.full(self._ndim, lbound, dtype=np.float64) if np.isscalar(lbound) else np.asarray(lbound, dtype=np.float64)
        ub = np.full(self._ndim, ubound, dtype=np.float64) if np.isscalar(ubound) else np.asarray(ubound, dtype=np.float64)
        a, pa = _darr(lb)
        b, pb = _darr(ub)
        _capi.check(self._lib.mcig_set_domain_ortho(self._ctx, pa, pb))

    # --- trial moves (include/mci/MCIntegrator.hpp:139-146)
    def setTrialMove(self, move, veclen=0, ntypes=1, typeEnds=None, nsteps=0, sub_pdfs=(), params=None):
        """setTrialMove(MoveType) or setTrialMove(SRRDType, veclen, ntypes, typeEnds). For MoveType.MultiStep, `nsteps` and
        `sub_pdfs` configure the move's own sub-sampling (MultiStepMove::setNSteps / addSamplingFunction). `params`: the parameters of a
        pre-made distribution handed to the move (reference: the `rdist` constructor argument, include/mci/SRRDAllMove.hpp:45-58), e.g.
        setTrialMove(SRRDType.Student, 0, params=(2.0,)) for test/ut5's StudentAllMove(ndim, 0.05, &student_t(2))."""
        te = None
        if typeEnds is not None:
            te_arr = np.ascontiguousarray(typeEnds, dtype=np.int32)
            te = te_arr.ctypes.data_as(C.POINTER(C.c_int))
        if isinstance(move, Move):  # setTrialMove(const TrialMoveInterface &) with a user-defined move functor
            a, p = _darr(move.par)
            _capi.check(self._lib.mcig_set_move_plugin(self._ctx, move.plugin_id(), p, len(move.par), ntypes, te))
            return
        srrd = 0
        if isinstance(move, MoveType):
            mt = int(move)
            vl = veclen if veclen > 0 else 1
        else:  # SRRDType: veclen 0 = all-move
            mt = int(MoveType.Vec) if veclen > 0 else int(MoveType.All)
            vl = max(1, veclen)
            srrd = int(move)
        _capi.check(self._lib.mcig_set_move(self._ctx, mt, srrd, vl, ntypes, te))
        if params is not None and len(params) > 0:
            a, p = _darr(list(params))
            _capi.check(self._lib.mcig_set_srrd_params(self._ctx, len(params), p))
        if mt == int(MoveType.MultiStep):
            _capi.check(self._lib.mcig_multistep_config(self._ctx, int(nsteps)))
            for pdf in sub_pdfs:
                a, p = _darr(pdf.par)
                _capi.check(self._lib.mcig_multistep_add_pdf(self._ctx, pdf.plugin_id(), p, len(pdf.par)))

    # --- observables / sampling functions (include/mci/MCIntegrator.hpp:149-184)
    def addObservable(self, obs, blocksize=1, nskip=1, flag_equil=None, estim=None):
        """addObservable(obs, blocksize=1, nskip=1[, flag_equil, flag_correlated | EstimatorType])"""
        blocksize = max(0, int(blocksize))
        nskip = max(1, int(nskip))
        if flag_equil is None:
            flag_equil = blocksize > 0
        if estim is None:
            estim = selectEstimatorType(blocksize == 1, blocksize > 0)
        elif isinstance(estim, bool):
            estim = selectEstimatorType(estim, blocksize > 0)
        a, p = _darr(obs.par)
        _capi.check(self._lib.mcig_add_obs(self._ctx, obs.plugin_id(), p, len(obs.par), blocksize, nskip, int(bool(flag_equil)), int(estim)))

    def setCallback(self, callback, buffer_doubles):
        """setCallback(StepCallback, size of its device buffer in doubles); the buffer is zeroed at every integrate call"""
        a, p = _darr(callback.par)
        _capi.check(self._lib.mcig_set_callback(self._ctx, callback.plugin_id(), p, len(callback.par), int(buffer_doubles)))
        self._cb_doubles = int(buffer_doubles)

    def clearCallback(self): _capi.check(self._lib.mcig_clear_callback(self._ctx))

    def getCallbackBuffer(self):
        out = np.zeros(self._cb_doubles)
        _capi.check(self._lib.mcig_get_callback_buffer(self._ctx, out.ctypes.data_as(_dp), self._cb_doubles))
        return out

    def popObservable(self): _capi.check(self._lib.mcig_pop_obs(self._ctx))
    def clearObservables(self): _capi.check(self._lib.mcig_clear_obs(self._ctx))

    def addSamplingFunction(self, pdf):
        a, p = _darr(pdf.par)
        _capi.check(self._lib.mcig_add_pdf(self._ctx, pdf.plugin_id(), p, len(pdf.par)))

    def popSamplingFunction(self): _capi.check(self._lib.mcig_pop_pdf(self._ctx))
    def clearSamplingFunctions(self): _capi.check(self._lib.mcig_clear_pdfs(self._ctx))

    # --- getters
    def getNDim(self): return self._ndim
    def getNObsDim(self): return self._lib.mcig_get_nobsdim(self._ctx)
    def getAcceptanceRate(self): return self._lib.mcig_get_acceptance_rate(self._ctx)

    # --- integrate (include/mci/MCIntegrator.hpp:225)
    def integrate(self, Nmc, doFindMRT2step=True, doDecorrelation=True):
        n = max(1, self.getNObsDim())
        avg = np.zeros(n)
        err = np.zeros(n)
        _capi.check(self._lib.mcig_integrate(self._ctx, int(Nmc), avg.ctypes.data_as(_dp), err.ctypes.data_as(_dp),
                                             int(bool(doFindMRT2step)), int(bool(doDecorrelation))))
        nod = self.getNObsDim()
        return avg[:nod], err[:nod]

    # --- beyond the reference
    def setAllreduce(self, fn):
        """fn(numpy array) sums the array in place over all processes (e.g. torch.distributed.all_reduce)."""
        if fn is None:
            self._cb = None
            _capi.check(self._lib.mcig_set_allreduce(self._ctx, _capi.ALLREDUCE_FN(), None))
            return

        def _tramp(buf, n, user):
            arr = np.ctypeslib.as_array(buf, shape=(n,))
            fn(arr)
        self._cb = _capi.ALLREDUCE_FN(_tramp)
        _capi.check(self._lib.mcig_set_allreduce(self._ctx, self._cb, None))

    def walkerResults(self):
        nod, W = self._lib.mcig_get_result_nobsdim(self._ctx), self.getNWalkers()
        avg = np.zeros((nod, W))
        err = np.zeros((nod, W))
        _capi.check(self._lib.mcig_get_walker_results(self._ctx, avg.ctypes.data_as(_dp), err.ctypes.data_as(_dp)))
        return avg, err

    def sums(self):
        nod = self._lib.mcig_get_result_nobsdim(self._ctx)  # of the last integrate: the observable list may have changed since
        s = np.zeros(3*max(1, nod))
        _capi.check(self._lib.mcig_get_sums(self._ctx, s.ctypes.data_as(_dp), len(s)))
        return s[:3*nod]

    def crossWalkerError(self):
        nod = self._lib.mcig_get_result_nobsdim(self._ctx)
        e = np.zeros(max(1, nod))
        _capi.check(self._lib.mcig_get_cross_walker_error(self._ctx, e.ctypes.data_as(_dp), len(e)))
        return e[:nod]

    def setKeepSamples(self, on=True):
        """Keep the stored series of Block / Full accumulators whose one-pass estimator runs inside the walk kernel (for obsData)."""
        _capi.check(self._lib.mcig_set_keep_samples(self._ctx, int(bool(on))))

    def attachComm(self, on=True):
        """This MCI is one rank's shard of a job over comm_size() processes: every all-reduce of the path runs as ncclAllReduce on the
        engine's stream (see parallel.init_comm)."""
        _capi.check(self._lib.mcig_attach_comm(self._ctx, int(bool(on))))

    def obsData(self, iobs, walker=0, nobs=1):
        nstore = self._lib.mcig_get_nstore(self._ctx, int(iobs))
        out = np.zeros((max(1, nstore), nobs))
        _capi.check(self._lib.mcig_get_obs_data(self._ctx, int(iobs), int(walker), out.ctypes.data_as(_dp)))
        return out

    def timings(self):
        w, e, t, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
        _capi.check(self._lib.mcig_get_timings(self._ctx, C.byref(w), C.byref(e), C.byref(t), C.byref(n)))
        f, d, j = C.c_double(), C.c_double(), C.c_double()
        _capi.check(self._lib.mcig_get_phase_timings(self._ctx, C.byref(f), C.byref(d), C.byref(j)))
        return {"walk_ms": w.value, "estim_ms": e.value, "total_ms": t.value, "launches": n.value, "find_ms": f.value, "decorr_ms": d.value, "jit_ms": j.value}

    def setBlockSize(self, n): _capi.check(self._lib.mcig_set_block_size(self._ctx, int(n)))
    def setStatePlacement(self, p): _capi.check(self._lib.mcig_set_state_placement(self._ctx, int(p)))
    def storeObservablesOnFile(self, path, freq): _capi.check(self._lib.mcig_store_on_file(self._ctx, 0, path.encode(), int(freq)))
    def storeWalkerPositionsOnFile(self, path, freq): _capi.check(self._lib.mcig_store_on_file(self._ctx, 1, path.encode(), int(freq)))
    def clearObservableFile(self): _capi.check(self._lib.mcig_store_on_file(self._ctx, 0, b"", 0))
    def clearWalkerFile(self): _capi.check(self._lib.mcig_store_on_file(self._ctx, 1, b"", 0))
    def setStreamPosition(self, group): _capi.check(self._lib.mcig_set_stream_position(self._ctx, int(group)))
    def getStreamPosition(self): return int(self._lib.mcig_get_stream_position(self._ctx))
    def setLazyAccumulation(self, on): _capi.check(self._lib.mcig_set_lazy_accumulation(self._ctx, int(on)))
    def setDeviceCalibration(self, on): _capi.check(self._lib.mcig_set_device_calibration(self._ctx, int(on)))
    def getCalibrationIterations(self): return self._lib.mcig_get_calibration_iterations(self._ctx)
    def getDecorrelationChunks(self): return self._lib.mcig_get_decorrelation_chunks(self._ctx)
    def getStagingChunks(self): return self._lib.mcig_get_staging_chunks(self._ctx)
    def setPhiloxRounds(self, rounds): _capi.check(self._lib.mcig_set_philox_rounds(self._ctx, int(rounds)))
    def setDynamicScheduling(self, mode): _capi.check(self._lib.mcig_set_dynamic_scheduling(self._ctx, int(mode)))
    def prebuild(self): _capi.check(self._lib.mcig_prebuild(self._ctx))

    def warmup(self, nmc, findMRT2Step=True, initialDecorr=True):
        """prebuild() plus an undone rehearsal of integrate(nmc, findMRT2Step, initialDecorr): buffers allocated, kernels loaded and run once"""
        _capi.check(self._lib.mcig_warmup(self._ctx, int(nmc), int(bool(findMRT2Step)), int(bool(initialDecorr))))

    def kernelSource(self):
        n = self._lib.mcig_get_kernel_source(self._ctx, None, 0)
        buf = C.create_string_buffer(int(n))
        self._lib.mcig_get_kernel_source(self._ctx, buf, n)
        return buf.value.decode()


def estimate(estim_type, x):
    """Device version of the reference's free estimator functions (include/mci/Estimators.hpp:9-45): x[n] or x[n][ndim]."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, ndim = (x.shape[0], 1) if x.ndim == 1 else x.shape
    avg = np.zeros(ndim)
    err = np.zeros(ndim)
    _capi.check(_capi.lib().mcig_estimate(int(estim_type), n, ndim, x.ctypes.data_as(_dp), avg.ctypes.data_as(_dp), err.ctypes.data_as(_dp)))
    return avg, err


def estimate_blocks(x, nblocks):
    """One/MultiDimBlockEstimator (include/mci/Estimators.hpp:12, :30): fixed number of blocks, then the uncorrelated estimator."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n, ndim = (x.shape[0], 1) if x.ndim == 1 else x.shape
    avg = np.zeros(ndim)
    err = np.zeros(ndim)
    _capi.check(_capi.lib().mcig_estimate_blocks(n, ndim, x.ctypes.data_as(_dp), int(nblocks), avg.ctypes.data_as(_dp), err.ctypes.data_as(_dp)))
    return avg, err


def measure_peaks(device=0):
    d, i = C.c_double(), C.c_double()
    _capi.check(_capi.lib().mcig_measure_peaks(int(device), C.byref(d), C.byref(i)))
    return d.value, i.value


def measure_philox_peak(device=0):
    """Philox4x32-10 blocks per second with nothing else in the loop: the RNG's own issue-rate bound (mcig_measure_philox_peak)."""
    b = C.c_double()
    _capi.check(_capi.lib().mcig_measure_philox_peak(int(device), C.byref(b)))
    return b.value
