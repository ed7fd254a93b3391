#!/usr/bin/env python
"""bench.py — Metropolis samples/sec of the sampling path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...     # the reference's own CPU implementation on the host cores

Headline workload ("bench_throughput_3G" of BASELINE.json configs[1], pinned in SURVEY.md §8(d) C2): ThreeDimGaussianPDF + XSquared,
uniform all-move step 1.0, SimpleAccumulator, 65536 walkers per GPU x 1e5 Metropolis steps. One bench "step" = one
mci.integrate(1e5, avg, err, false, false) over all walkers of the rank (6.5536e9 samples per GPU). Weak scaling: the walkers per GPU
are fixed, global walker ids rank*65536.., one all-reduce of [sum avg | sum err^2] per integrate — issued by the library itself as an
ncclAllReduce on the engine's stream (include/mcig.h: mcig_comm_*, mcig_attach_comm); torch.distributed only transports the NCCL id,
the barriers and the max-over-ranks of the timings.

value  : samples/s, inputs resident in HBM, device time (CUDA events on the engine's stream around the whole integrate call),
         max over ranks.
e2e    : the same metric through the public C-ABI call sequence with HOST buffers — set walker positions from host memory
         (H2D), integrate, read avg/err/acceptance back (D2H) — wall clock around the calls, max over ranks.
roofline: FP64 issue rate. achieved = 34 algorithmic FP64 instructions per Metropolis step x steps / walk-kernel
         time; peak = DFMA/s measured live on the same GPU (MEASURED_PEAKS.json has no FP64 figure); the nominal peak
         (64 FP64 lanes x 148 SMs x max SM clock) and the fraction against it are printed next to it.
secondary: the other BASELINE.json configs, measured OUTSIDE the timed headline region, each with its own roofline and CPU baseline:
         C3 (configs[2]: dimension sweep, vec / all / MultiStepMove, Block(20)), C4 (configs[3]: Full + MJBlocker, HBM-bound),
         C5 (configs[4]: mixed observables with automatic calibration + decorrelation; sharded over the ranks under torchrun) and the
         strong-scaling variant of the headline (65536 walkers in total).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WALKERS_PER_GPU = 65536
NMC = 100000
FP64_INSTR_PER_STEP = 34   # 4 uniform conversions + 6 proposal + 3 proto + 1 exponent + 17 exp + 1 compare + 1 obs + 1 accumulate
FLOP_PER_STEP = 53
WALK_DRAM_BYTES_PER_LAUNCH = 1586432  # ncu dram__bytes_read.sum (+ 0 written) of one mcig_walk_dyn launch, profiles/r02z_walk_ncu_raw.csv
METRIC = "metropolis_samples_per_sec"
WORKLOAD = "bench_throughput_3G: ThreeDimGaussianPDF ndim=3 + XSquared, uniform all-move step 1.0, SimpleAccumulator, 65536 walkers/GPU x 1e5 steps"
FP64_LANES_PER_SM, N_SM = 64, 148
REF_NMC = 15000000  # Metropolis steps per chain of one reference-arm bench step / of the cpu_baseline sample (~1.5 s per chain)


def bench_config(world):
    return {"workload": WORKLOAD, "rng": "Philox4x32-10, 32-bit uniforms", "walkers_per_gpu": WALKERS_PER_GPU, "nmc": NMC,
            "l2": "no HBM inputs to cache: walker state lives in registers; 1.5 MB of positions read once per step",
            "parallelism": "walkers sharded over %d GPU(s), 1 all-reduce of 2 doubles per integrate" % world}


def hbm_peak():
    """Measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region: one streaming `nvidia-smi -lms 100` child (B200_PROFILING.md)."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        mhz, reasons, mx, pw = [], set(), None, []
        for t, line in self.lines:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                mhz.append(float(f[0]))
                mx = float(f[1])
                pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(self.NAMES, f[3:]):
                if v.startswith("Active"):
                    reasons.add(n)
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz)//2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(mhz),
                "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------------------- reference on the host cores
_WORKER = r"""
import sys, json, time
sys.path.insert(0, %(orc_dir)r)
import numpy as np
import orc
e = orc.Engine(%(path)r, %(prefix)r)
cache = {}
print("ready", flush=True)
for line in sys.stdin:
    t = json.loads(line)
    if t.get("quit"):
        break
    if t["kind"] == "run":
        kw = dict(t["kw"])
        kw["obs"] = [tuple(o) for o in kw["obs"]]
        kw["seed"] = kw["seed"] + t["index"]
        cfg = orc.make_config(kw.pop("ndim"), kw.pop("seed"), kw.pop("pdf_id"), kw.pop("obs"), t["nmc"], **kw)
        while time.time() < t["t_start"]:
            pass
        t0 = time.time()
        r = e.run(cfg)
        t1 = time.time()
        print(json.dumps({"t0": t0, "t1": t1, "avg0": (r["avg"][0] if r["avg"] else 0.0), "acc": r["acc_rate"]}), flush=True)
    else:  # MJBlocker on an AR(1) series of n samples (benchmark/bench_estimators)
        n = t["n"]
        if n not in cache:
            rng = np.random.default_rng(7 + t["index"])
            x = rng.normal(size=n)
            for i in range(1, n):
                x[i] += 0.5*x[i - 1]
            cache = {n: x}
        x = cache[n]
        while time.time() < t["t_start"]:
            pass
        t0 = time.time()
        a, er = e.estimate(orc.EST_MJBLOCKER, x)
        t1 = time.time()
        print(json.dumps({"t0": t0, "t1": t1, "avg0": float(a[0]), "acc": float(er[0])}), flush=True)
"""


class CpuPool:
    """The reference's CPU implementation on ALL host cores: P persistent worker processes, each one independent chain (one `MCI` per
    core is the reference's MPI model, src/MPIMCI.cpp:83). The unmodified reference compiled with its own tuning flags
    (oracle/_ref/libmci_ref_v4|v3.so, see oracle/Makefile) when it was built, else the C restatement. A task starts on every worker at
    the same wall-clock instant and is timed INSIDE the workers: wall = last finish - common start, no process start-up or teardown."""

    def __init__(self, nproc=None):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc
        self.orc = orc
        self.nproc = nproc or os.cpu_count() or 1
        if orc.have_ref():
            path, self.flags = orc.ref_timing_path()
            self.kind, prefix = "reference", "mciref"
        else:
            if not os.path.exists(orc.ORACLE_PATH):
                subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
            path, self.flags, self.kind, prefix = orc.ORACLE_PATH, "gcc -O3 -ffp-contract=off (C restatement)", "port", "mcio"
        code = _WORKER % {"orc_dir": os.path.join(ROOT, "oracle"), "path": path, "prefix": prefix}
        self.procs = [subprocess.Popen([sys.executable, "-c", code], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True) for _ in range(self.nproc)]
        for p in self.procs:
            assert p.stdout.readline().strip() == "ready"

    def _go(self, task):
        t_start = time.time() + 0.05
        for i, p in enumerate(self.procs):
            p.stdin.write(json.dumps(dict(task, index=i, t_start=t_start)) + "\n")
            p.stdin.flush()
        outs = [json.loads(p.stdout.readline()) for p in self.procs]
        return max(o["t1"] for o in outs) - t_start, outs

    def run(self, kw, nmc):
        """P chains x nmc Metropolis steps of the configuration `kw` (oracle/orc.py:make_config keywords). Returns (steps/s, wall s, outs)."""
        wall, outs = self._go({"kind": "run", "kw": kw, "nmc": int(nmc)})
        return self.nproc*nmc/wall, wall, outs

    def mjblocker(self, n):
        """P independent series of n samples through the reference's MJBlocker. Returns (GB/s of series read once, wall s)."""
        self._go({"kind": "est", "n": int(n)})  # first call generates the data
        wall, outs = self._go({"kind": "est", "n": int(n)})
        return self.nproc*8.0*n/wall/1e9, wall

    def close(self):
        for p in self.procs:
            try:
                p.stdin.write('{"quit": 1}\n')
                p.stdin.flush()
                p.wait(timeout=5)
            except Exception:
                p.kill()


def _orc_ids():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    return orc


def c2_kw():
    orc = _orc_ids()
    return dict(ndim=3, seed=1000, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], steps=(1.0,))


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the headline workload, all host cores, one chain per core,
    REF_NMC steps per chain per bench step, timed inside the worker processes."""
    if rank != 0:
        return
    pool = CpuPool()
    kw = c2_kw()
    for _ in range(args.warmup):
        pool.run(kw, REF_NMC//10)
    walls = []
    for _ in range(args.steps):
        v, wall, outs = pool.run(kw, REF_NMC)
        assert all(0.3 < o["avg0"] < 0.7 for o in outs), outs
        walls.append(wall)
    pool.close()
    total = sum(walls)
    # The box's host cores are shared with other jobs (the same arm read 2.3e8 .. 3.3e8 samples/s on boxes of this pool within an hour): the value is the
    # FASTEST of the K steps, i.e. the reference at its least disturbed -- the conservative denominator for any GPU/CPU ratio; the mean over all K steps
    # is reported next to it
    best = min(walls)
    value = pool.nproc*REF_NMC/best
    sample = ("%d independent chains (one per host core) x %.1e Metropolis steps of the same integrand per bench step, fastest of %d steps (mean over all: %.4g samples/s); %s"
              % (pool.nproc, REF_NMC, args.steps, args.steps*pool.nproc*REF_NMC/total, pool.flags))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3*best, "mean_value": args.steps*pool.nproc*REF_NMC/total, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": bench_config(world),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": pool.nproc, "kind": pool.kind, "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------- GPU arms
def make_mci(m, rank, world, walkers=WALKERS_PER_GPU, device=None):
    mci = m.MCI(3, device=int(os.environ.get("LOCAL_RANK", 0)) if device is None else device)
    mci.setRngMode(m.RngMode.Philox32)
    mci.setSeed(1337)
    mci.setNWalkers(walkers, global_offset=rank*walkers, total=world*walkers)
    mci.addSamplingFunction(m.ThreeDimGaussianPDF())
    mci.addObservable(m.XSquared(), 0, 1)  # blocksize 0: SimpleAccumulator + Noop estimator
    mci.setMRT2Step(1.0)  # acceptance ~0.5 (benchmark/bench_integrate_mixed/main.cpp:44)
    if world > 1:
        mci.attachComm()
    return mci


C3_NDIMS = (1, 2, 4, 8, 16, 32, 64)


# DRAM bytes of the C4 estimator stage (65536 chains x 2^14 samples = 8.59 GB of series), summed over its kernels: profiles/r02bf_c4_launches_ncu.csv
C4_DRAM_BYTES = {"MJBlocker": 8591.4e6 + 179.8e6 + 184.6e6 + 4.0e6 + 16.3e6 + 0.7e6,   # mj_segments_tiled + mj_merge + mj_finish
                 "FCBlocker": 8590.0e6 + 632.3e6 + 2.1e6 + 662.8e6 + 6.0e6 + 47.3e6 + 2.7e6}  # fc_split_prefix + fc_seg_scan + fc_part_stats + fc_final


def c3_mci(m, move, nd, W, local):
    """BASELINE configs[2]: ExpNDPDF + XND with BlockAccumulator(20); `vec` is the reference's own single-particle sweep
    (benchmark/bench_throughput_ndim_single/main.cpp:26-50), `all` its all-move twin, `multistep` MultiStepMove(ndim sub-steps of a
    single-index uniform move) with main pdf Gauss and sub-pdf ExpNDPDF as in test/ut5/main.cpp:113-127."""
    mci = m.MCI(nd, device=local)
    mci.setRngMode(0)
    mci.setSeed(1337)
    mci.setNWalkers(W)
    if move == "vec":
        mci.setTrialMove(m.MoveType.Vec)
    elif move == "multistep":
        mci.setTrialMove(m.MoveType.MultiStep, 1, sub_pdfs=[m.ExpNDPDF(nd)] if nd > 1 else [])
    else:
        mci.setTrialMove(m.MoveType.All)
    mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)])
    mci.setMRT2Step(3.0 if move == "vec" else (0.5 if move == "multistep" else 3.0/nd**0.5))
    mci.addSamplingFunction(m.ExpNDPDF(nd) if move != "multistep" else m.Gauss(nd))
    mci.addObservable(m.XND(nd), 20, 1)  # BlockAccumulator(20) + Uncorrelated
    return mci


def c3_kw(move, nd):
    orc = _orc_ids()
    x0 = [0.1 if j % 2 == 0 else -0.05 for j in range(nd)]
    if move == "vec":
        return dict(ndim=nd, seed=2000, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], move_type=orc.MOVE_VEC, veclen=1, steps=(3.0,), x0=x0)
    if move == "all":
        return dict(ndim=nd, seed=2000, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], steps=(3.0/nd**0.5,), x0=x0)
    return dict(ndim=nd, seed=2000, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 20, 1)], move_type=orc.MOVE_MULTISTEP, veclen=1,
                ms_sub_pdf_id=orc.PDF_EXPND if nd > 1 else orc.PDF_NONE, steps=(0.5,), x0=x0)


def c3_work(move, nd):
    """Algorithmic FP64-pipe instructions and Philox4x32 blocks per Metropolis step (SURVEY.md §8d formulas)."""
    if move == "vec":    # 2 uniforms + proposal 2 + proto delta 1 + exp 17 + compare 1 = 23, + ndim accumulate; 3 draws = 1 block
        return 23 + nd, 1
    if move == "all":    # (ndim+1) uniforms + 2 ndim proposal + 2 ndim proto sums + 1 + exp 17 + compare 1 + ndim accumulate
        return (nd + 1) + 2*nd + 2*nd + 1 + 17 + 1 + nd, (nd + 1 + 3)//4
    # MultiStepMove: 23 per sub-step + outer test; Philox blocks: four sub-steps share one draw group of 12 values = 3 blocks, the last ndim % 4 sub-steps
    # and the outer accept uniform a group of 3 (ndim % 4) + 1 values (device/mcig_device.cuh: MCIG_MS_QUADS)
    return 28*nd + 64, 3*(nd//4) + (3*(nd % 4) + 1 + 3)//4


def secondary_c3(m, local, peaks, philox_peak, pool):
    out = {}
    W = 65536
    for move in ("vec", "all", "multistep"):
        rows = []
        for nd in C3_NDIMS:
            nmc = {"vec": 8000, "all": 4000 if nd <= 16 else 2000}.get(move, max(200, 8000//(2*nd)//20*20))
            mci = c3_mci(m, move, nd, W, local)
            mci.integrate(max(200, nmc//10//20*20), False, False)
            best = None
            for _ in range(2):
                avg, err = mci.integrate(nmc, False, False)
                t = mci.timings()
                best = t if best is None or t["walk_ms"] < best["walk_ms"] else best
            sps = W*nmc/(best["walk_ms"]*1e-3)
            instr, blocks = c3_work(move, nd)
            row = {"ndim": nd, "steps_per_s": sps, "walk_ms": best["walk_ms"], "estim_ms": best["estim_ms"], "nmc": nmc, "acceptance": mci.getAcceptanceRate(),
                   "roofline": {"bound": "fp64_issue", "fp64_instr_per_step": instr, "achieved": instr*sps/1e9, "peak": peaks[0]/1e9, "unit": "GFP64inst/s",
                                "frac": instr*sps/peaks[0], "philox_blocks_per_step": blocks, "rng_frac": blocks*sps/philox_peak}}
            if move == "multistep":
                row["roofline"]["note"] = ("fp64_instr_per_step is SURVEY.md §8d's ALGORITHMIC count of the reference's expression (one exp per sub-step, three exp and a "
                                           "division per outer step); the kernel decides sub-steps and the outer step in log form through the FP32 pre-filter and executes far "
                                           "fewer FP64 instructions, so this fraction overstates the pipe's load (it exceeds 1 at ndim 1): the bound that holds is rng_frac")
            if pool is not None:
                nmc_cpu = max(2000, int({"vec": 2.4e7/(1 + nd/6.0), "all": 2.4e7/(1 + nd/1.5)}.get(move, 2.4e7/(1 + 6.0*nd)))//20*20)  # ~0.2-1 s per chain
                v, wall, _ = pool.run(c3_kw(move, nd), nmc_cpu)
                row["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": pool.nproc, "kind": pool.kind,
                                       "sample": "%d chains x %d steps, %.2f s" % (pool.nproc, nmc_cpu, wall)}
            rows.append(row)
            del mci
        out[move] = rows
    return {"workload": "BASELINE configs[2]: ExpNDPDF/Gauss + XND, BlockAccumulator(20), 65536 walkers, ndim 1..64; moves: vec (single index, "
                        "the reference's bench_throughput_ndim_single), all, multistep (MultiStepMove, ndim sub-steps, sub-pdf ExpNDPDF)",
            "unit": "Metropolis steps/s (walk kernel, CUDA events)", **out}


def secondary_c4(m, local, pool):
    """BASELINE configs[3]: FullAccumulator + MJBlocker with device-side HBM staging. (a) 65536 chains x 2^14 samples (8.6 GB staged by the
    walk, one streaming read by the estimator); (b) the same integrand at 65536 chains x 2^20 = 550 GB of series, which does not fit HBM:
    staged and folded chunk by chunk (chunked Full staging)."""
    peak, src = hbm_peak()
    W, npow = 65536, 14
    nmc = 1 << npow
    out = {"workload": "BASELINE configs[3]: ThreeDimGaussianPDF + XSquared, FullAccumulator + MJBlocker, device-side HBM staging"}
    rows = []
    for label, est in (("MJBlocker", m.EstimatorType.MJBlocker), ("FCBlocker", m.EstimatorType.FCBlocker)):
        mci = m.MCI(3, device=local)
        mci.setRngMode(0)
        mci.setSeed(1337)
        mci.setNWalkers(W)
        mci.addSamplingFunction(m.ThreeDimGaussianPDF())
        mci.addObservable(m.XSquared(), 1, 1, False, est)
        mci.setMRT2Step(1.0)
        mci.integrate(nmc, False, False)
        ts = []
        for _ in range(5):
            avg, err = mci.integrate(nmc, False, False)
            ts.append(mci.timings())
        e_ms = min(q["estim_ms"] for q in ts)
        w_ms = min(q["walk_ms"] for q in ts)
        nbytes = 8.0*nmc*W
        rows.append({"estimator": label, "walkers": W, "n_per_chain": nmc, "series_GB": nbytes/1e9, "walk_ms": w_ms, "estim_ms": e_ms,
                     "walk_steps_per_s": W*nmc/(w_ms*1e-3), "walk_store_GBps": nbytes/(w_ms*1e-3)/1e9, "avg": float(avg[0]), "err": float(err[0]),
                     "roofline": {"bound": "hbm", "achieved": nbytes/(e_ms*1e-3)/1e9, "peak": peak, "unit": "GB/s", "frac": nbytes/(e_ms*1e-3)/1e9/peak,
                                  "peak_source": src, "algorithmic_bytes": "8 B per stored sample: one read of the series (the mean comes from the walk's running sum)",
                                  "traffic": C4_DRAM_BYTES[label],
                                  "traffic_note": "dram__bytes_read + dram__bytes_write of the estimator's kernels for this series, ncu launch list "
                                                  "profiles/r02bf_c4_launches_ncu.csv (not re-measured inside bench.py): the series is read exactly once (8.59 GB), "
                                                  "the rest is the per-segment partial results; the walk that stages it writes 8.57 GB and reads 2 MB"}})
        del mci
    out["in_hbm"] = rows
    # (b) series larger than HBM
    try:
        W2, npow2 = 65536, 20
        mci = m.MCI(3, device=local)
        mci.setRngMode(0)
        mci.setSeed(1337)
        mci.setNWalkers(W2)
        mci.addSamplingFunction(m.ThreeDimGaussianPDF())
        mci.addObservable(m.XSquared(), 1, 1, False, m.EstimatorType.MJBlocker)
        mci.setMRT2Step(1.0)
        mci.integrate(1 << npow2, False, False)
        t0 = time.perf_counter()
        avg, err = mci.integrate(1 << npow2, False, False)
        wall = time.perf_counter() - t0
        t = mci.timings()
        nbytes = 8.0*(1 << npow2)*W2
        out["beyond_hbm"] = {"walkers": W2, "n_per_chain": 1 << npow2, "series_GB": nbytes/1e9, "total_ms": t["total_ms"], "wall_ms": 1e3*wall,
                             "samples_per_s": W2*(1 << npow2)/(t["total_ms"]*1e-3), "staged_GBps_write_plus_read": 2*nbytes/(t["total_ms"]*1e-3)/1e9,
                             "avg": float(avg[0]), "err": float(err[0]),
                             "roofline": {"bound": "hbm", "achieved": 2*nbytes/(t["total_ms"]*1e-3)/1e9, "peak": peak, "unit": "GB/s",
                                          "frac": 2*nbytes/(t["total_ms"]*1e-3)/1e9/peak, "peak_source": src,
                                          "algorithmic_bytes": "16 B per stored sample: written once by the walk, read once by the estimator (SURVEY.md §8d)"}}
        del mci
    except Exception as ex:  # reported, not hidden: the in-HBM rows above stand on their own
        out["beyond_hbm"] = {"error": str(ex)[:300]}
    if pool is not None:
        gbs, wall = pool.mjblocker(1 << 21)
        out["cpu_baseline"] = {"value": gbs, "unit": "GB/s of series (read once)", "cores": pool.nproc, "kind": pool.kind,
                               "sample": "%d series x 2^21 samples through the reference's MJBlocker (src/MJBlocker.cpp), %.2f s" % (pool.nproc, wall)}
    return out


def c5_mci(m, local, W, rank=0, world=1):
    mci = m.MCI(3, device=local)
    mci.setRngMode(0)
    mci.setSeed(5649871)
    mci.setNWalkers(W, global_offset=rank*W, total=world*W)
    mci.addSamplingFunction(m.ThreeDimGaussianPDF())
    mci.addObservable(m.XND(3), 0, 1)        # Simple
    mci.addObservable(m.XSquared(), 1, 5)    # Full, nskip 5, Correlated
    mci.addObservable(m.XYZSquared(), 5, 2)  # Block(5), nskip 2, Uncorrelated
    mci.setMRT2Step(1.0)
    if world > 1:
        mci.attachComm()
    return mci


def secondary_c5(m, local, rank, world, peaks, pool, barrier, maxreduce):
    """BASELINE configs[4]: bench_integrate_mixed (benchmark/bench_integrate_mixed/main.cpp:28-46) with automatic step calibration and
    decorrelation, 65536 walkers per GPU sharded over the ranks; every all-reduce of the two control loops and the final one run inside
    the library (NCCL on the engine's stream)."""
    W, nmc = 65536, 100000
    mci = c5_mci(m, local, W, rank, world)
    for _ in range(2):  # the first automatic run loads the calibration / equilibration kernel variants
        mci.setMRT2Step(1.0)
        mci.integrate(nmc, True, True)
    runs = []
    for _ in range(3):
        mci.setMRT2Step(1.0)
        barrier()
        avg, err = mci.integrate(nmc, True, True)
        t = mci.timings()
        t["total_ms"] = maxreduce(t["total_ms"])
        runs.append((t, avg, err))
    t, avg, err = min(runs, key=lambda q: q[0]["total_ms"])
    sps = float(world)*W*nmc/(t["total_ms"]*1e-3)
    instr = 34 + 3 + 2.0/5 + 6.0/2
    out = {"workload": "BASELINE configs[4]: ThreeDimGaussianPDF; XND(3) Simple + XSquared Full nskip 5 (Correlated) + XYZSquared Block(5) nskip 2 (Uncorrelated), "
                       "integrate(1e5, findMRT2Step=true, decorrelation=true), 65536 walkers per GPU",
           "n_gpus": world, "samples_per_s": sps, "total_ms": t["total_ms"], "walk_ms": t["walk_ms"], "estim_ms": t["estim_ms"], "find_ms": t["find_ms"],
           "decorr_ms": t["decorr_ms"], "launches": t["launches"], "calibration_iterations": mci.getCalibrationIterations(),
           "decorrelation_chunks": mci.getDecorrelationChunks(), "step": mci.getMRT2Step(0), "acceptance": mci.getAcceptanceRate(),
           "avg": [float(v) for v in avg], "err": [float(v) for v in err],
           "roofline": {"bound": "fp64_issue", "fp64_instr_per_step": instr, "achieved": instr*sps/world/1e9, "peak": peaks[0]/1e9, "unit": "GFP64inst/s",
                        "frac": instr*sps/world/peaks[0],
                        "note": "34 (C2) + 3 (XND accumulate) + 2/5 (XSquared every 5th step) + 6/2 (XYZSquared every 2nd); whole integrate incl. both control loops and the estimators; "
                                "HBM: 1.6 B/step written + read for the Full series, nothing for the Block(5) leg (one-pass estimator fused into the walk)"}}
    if pool is not None and rank == 0:
        orc = _orc_ids()
        kw = dict(ndim=3, seed=5649871, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XND, 0, 1), (orc.OBS_XSQUARED, 1, 5), (orc.OBS_XYZSQUARED, 5, 2)], steps=(1.0,),
                  do_find=True, do_decorr=True)
        v, wall, _ = pool.run(kw, 10000000)
        out["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": pool.nproc, "kind": pool.kind,
                               "sample": "%d chains x 1e7 steps incl. calibration + decorrelation, %.2f s" % (pool.nproc, wall)}
    del mci
    return out


def secondary_strong(m, local, rank, world, barrier, maxreduce):
    """Fixed total work: 65536 walkers over N GPUs (thin grids: the dynamically scheduled kernel's case)."""
    w = WALKERS_PER_GPU//world
    mci = make_mci(m, rank, world, walkers=w, device=local)
    for _ in range(3):
        mci.integrate(NMC, False, False)
    barrier()
    ms = 0.0
    reps = 5
    for _ in range(reps):
        mci.integrate(NMC, False, False)
        ms += mci.timings()["total_ms"]
    ms = maxreduce(ms)
    del mci
    return {"workload": "headline integrand, 65536 walkers in TOTAL over %d GPU(s) (%d per GPU) x 1e5 steps" % (world, w), "scaling": "strong", "n_gpus": world,
            "samples_per_s": float(WALKERS_PER_GPU)*NMC*reps/(ms*1e-3), "ms_per_step": ms/reps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)  # 20 x 26 ms: long enough for the clock sampler to see the load
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import mcintegratorplusplus_b200 as m
    from mcintegratorplusplus_b200 import _capi, parallel
    if _capi.lib().mcig_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the sampling path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        saved = os.dup(1)  # NCCL announces its version on stdout at the first communicator: keep stdout for the one JSON line
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            parallel.init_comm(local)  # the library's own NCCL communicator (id broadcast through torch.distributed)
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    mci = make_mci(m, rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = m.measure_peaks(local)
    philox_peak = m.measure_philox_peak(local)  # Philox4x32-10 blocks/s with nothing else issued: one block per Metropolis step in this workload
    # ---- device-resident arm
    for _ in range(max(3, args.warmup)):
        mci.integrate(NMC, False, False)
    sampler = ClockSampler(local)
    time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    dev_ms, walk_ms, launches = 0.0, 0.0, 0
    avg = err = None
    for _ in range(args.steps):
        avg, err = mci.integrate(NMC, False, False)
        tm = mci.timings()
        dev_ms += tm["total_ms"]
        walk_ms += tm["walk_ms"]
        launches += tm["launches"]
    barrier()
    t1 = time.perf_counter()
    wall_ms = 1e3*(t1 - t0)
    time.sleep(0.15)
    sampler.stop()
    clocks = sampler.summary(t0, t1)
    dev_ms = maxreduce(dev_ms)
    wall_ms = maxreduce(wall_ms)
    walk_ms_max = maxreduce(walk_ms)
    samples = float(world)*WALKERS_PER_GPU*NMC*args.steps
    value = samples/(dev_ms*1e-3)

    # ---- end-to-end arm: host buffers in, host results out, every step
    x0 = np.zeros((WALKERS_PER_GPU, 3))
    h2d = x0.nbytes
    nod = mci.getNObsDim()
    d2h = 8*((5 if world > 1 else 3)*nod) + 8
    for _ in range(2):
        mci.setXWalkers(x0)
        mci.integrate(NMC, False, False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mci.setXWalkers(x0)                       # H2D of this step's inputs
        a, e = mci.integrate(NMC, False, False)   # kernel + reductions + (N>1) all-reduce + D2H of avg/err
        rate = mci.getAcceptanceRate()
    barrier()
    e2e_ms = maxreduce(1e3*(time.perf_counter() - t0))
    e2e_value = samples/(e2e_ms*1e-3)
    cw_local = float(mci.crossWalkerError()[0])

    # ---- context for the roofline fraction (N = 1 only, outside every timed region): the named workload's 65536 walkers are 3.46 warps
    # per scheduler on 148 SMs; the same kernel with every scheduler holding 8 warps shows what the instruction stream itself allows
    full = None
    if world == 1:
        wfull = 148*2048
        mf = m.MCI(3, device=local)
        mf.setRngMode(m.RngMode.Philox32)
        mf.setSeed(1337)
        mf.setNWalkers(wfull)
        mf.addSamplingFunction(m.ThreeDimGaussianPDF())
        mf.addObservable(m.XSquared(), 0, 1)
        mf.setMRT2Step(1.0)
        mf.setBlockSize(256)
        for _ in range(2):
            mf.integrate(NMC, False, False)
        full = (wfull, float(wfull)*NMC/(mf.timings()["walk_ms"]*1e-3))
        del mf

    # ---- secondary workloads (outside the timed headline region)
    secondary = None
    pool = None
    if not args.no_secondary:
        if rank == 0 and not args.no_cpu_baseline:
            pool = CpuPool()
        secondary = {}
        secondary["C5"] = secondary_c5(m, local, rank, world, peaks, pool, barrier, maxreduce)
        secondary["strong_scaling"] = secondary_strong(m, local, rank, world, barrier, maxreduce)
        if world == 1:
            secondary["C3"] = secondary_c3(m, local, peaks, philox_peak, pool)
            secondary["C4"] = secondary_c4(m, local, pool)

    if rank == 0:
        steps_per_s_kernel = float(WALKERS_PER_GPU)*NMC*args.steps/(walk_ms_max*1e-3)  # per GPU
        achieved = FP64_INSTR_PER_STEP*steps_per_s_kernel
        sm_max_mhz = clocks.get("sm_max_mhz") or 1965.0
        nominal = FP64_LANES_PER_SM*N_SM*sm_max_mhz*1e6
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": dev_ms/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": bench_config(world),
            "wall_ms_per_step": wall_ms/args.steps,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms/args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64_issue", "achieved": achieved/1e9, "peak": peaks[0]/1e9, "unit": "GFP64inst/s", "frac": achieved/peaks[0],
                         "peak_nominal": nominal/1e9, "frac_nominal": achieved/nominal,
                         "peak_nominal_source": "64 FP64 lanes/SM/clk x 148 SMs x %.0f MHz (clocks.max.sm); the measured DFMA microbenchmark reaches %.2f of it" % (sm_max_mhz, peaks[0]/nominal),
                         "traffic": WALK_DRAM_BYTES_PER_LAUNCH, "traffic_note": "dram__bytes_read + dram__bytes_write of one walk launch (ncu --set full, profiles/r02z_walk_ncu_raw.csv): the 1.5 MB of start positions, nothing written; the bound is FP64/ALU issue, not HBM",
                         "kernel": "mcig_walk (JIT-specialised Metropolis walk)", "kernel_ms_per_step": walk_ms_max/args.steps,
                         "fp64_instr_per_metropolis_step": FP64_INSTR_PER_STEP, "flop_per_metropolis_step": FLOP_PER_STEP,
                         "achieved_tflops": FLOP_PER_STEP*steps_per_s_kernel/1e12, "peak_tflops": 2*peaks[0]/1e12,
                         "peak_source": "DFMA/s measured live by mcig_measure_peaks on this GPU (MEASURED_PEAKS.json has no FP64 figure)",
                         "imad_peak_ginst": peaks[1]/1e9,
                         "ncu_pipes": {"source": "profiles/r02z_walk_ncu_raw.csv (ncu --set full of this kernel; not re-measured inside bench.py)",
                                       "fp64_pipe_active_pct": 15.6, "fma_heavy_pipe_active_pct": 58.0, "alu_pipe_active_pct": 50.7, "issue_slots_busy_pct": 57.3,
                                       "note": "the kernel EXECUTES ~12 FP64 instructions per step (FP32 pre-filter of the accept test, cached observable); the 34 above are the algorithmic count the fraction is quoted on"},
                         "rng_bound": {"philox4x32_10_blocks_per_s": philox_peak, "blocks_per_step": 1, "frac": steps_per_s_kernel/philox_peak,
                                       "note": "issue-rate bound of the counter RNG alone (20 IMAD.WIDE.U32 at a quarter of the FP32 rate per block), measured live; "
                                               "the walk loop cannot exceed it whatever its FP64 content: context for the FP64 fraction above"},
                         "full_occupancy": None if full is None else {
                             "walkers": full[0], "steps_per_s": full[1], "frac": FP64_INSTR_PER_STEP*full[1]/peaks[0],
                             "note": "same kernel, 8 warps per scheduler instead of the workload's 3.46; context only, not the bench value"}},
            "clocks": clocks,
            "result": {"avg": float(avg[0]), "err": float(err[0]), "acceptance": float(rate), "cross_walker_err_local": cw_local},
        }
        if pool is not None and world == 1:
            v, wall, outs = max((pool.run(c2_kw(), REF_NMC) for _ in range(2)), key=lambda r: r[0])  # the less disturbed of two samples (shared host cores)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": pool.nproc, "kind": pool.kind,
                                    "sample": "%d independent chains (one per host core) x %.1e Metropolis steps of the same integrand, %.1f s wall; %s" % (pool.nproc, REF_NMC, wall, pool.flags)}
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line))
    if pool is not None:
        pool.close()
    if dist is not None:
        dist.barrier()
        parallel.finalize_comm()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
