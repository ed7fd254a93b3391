#!/usr/bin/env python
"""bench.py — Metropolis samples/sec of the sampling path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...     # the reference's own CPU implementation on the host cores

Workload ("bench_throughput_3G" of BASELINE.json configs[1], pinned in SURVEY.md §8(d) C2): ThreeDimGaussianPDF + XSquared,
uniform all-move step 1.0, SimpleAccumulator, 65536 walkers per GPU x 1e5 Metropolis steps. One bench "step" = one
mci.integrate(1e5, avg, err, false, false) over all walkers of the rank (6.5536e9 samples per GPU). Weak scaling: the
walkers per GPU are fixed, global walker ids rank*65536.., one all-reduce of [sum avg | sum err^2] per integrate.

value  : samples/s, inputs resident in HBM, device time (CUDA events on the engine's stream around the whole integrate call),
         max over ranks.
e2e    : the same metric through the public C-ABI call sequence with HOST buffers — set walker positions from host memory
         (H2D), integrate, read avg/err/acceptance back (D2H) — wall clock around the calls, max over ranks.
roofline: FP64 issue rate. achieved = 34 algorithmic FP64 instructions per Metropolis step x steps / walk-kernel
         time; peak = DFMA/s measured live on the same GPU by the library's microbenchmark (MEASURED_PEAKS.json has no FP64 figure).
"""
import argparse
import ctypes
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
WALK_DRAM_BYTES_PER_LAUNCH = 1586176  # ncu dram__bytes_read.sum (+ 0 written) of one mcig_walk_dyn launch, profiles/r01_walk_r1f_dyn_ncu_raw.csv
METRIC = "metropolis_samples_per_sec"
WORKLOAD = "bench_throughput_3G: ThreeDimGaussianPDF ndim=3 + XSquared, uniform all-move step 1.0, SimpleAccumulator, 65536 walkers/GPU x 1e5 steps"


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

    def mark(self):
        return time.perf_counter()

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


def make_mci(m, rank, world):
    mci = m.MCI(3, device=int(os.environ.get("LOCAL_RANK", 0)))
    mci.setRngMode(m.RngMode.Philox32)
    mci.setSeed(1337)
    mci.setNWalkers(WALKERS_PER_GPU, global_offset=rank*WALKERS_PER_GPU, total=world*WALKERS_PER_GPU)
    mci.addSamplingFunction(m.ThreeDimGaussianPDF())
    mci.addObservable(m.XSquared(), 0, 1)  # blocksize 0: SimpleAccumulator + Noop estimator
    mci.setMRT2Step(1.0)  # acceptance ~0.5 (benchmark/bench_integrate_mixed/main.cpp:44)
    return mci


def cpu_reference_run(nmc, nproc):
    """The reference's CPU implementation of the same path on the host cores: P independent chains (one per core, the
    reference's MPI model, src/MPIMCI.cpp:83), each nmc steps. Uses oracle/_ref (the unmodified reference) when it was
    built, else the C port. Returns (samples/s, kind, wall seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    kind = "reference" if orc.have_ref() else "port"
    if kind == "port" and not os.path.exists(orc.ORACLE_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, capture_output=True)
    code = (
        "import sys,time; sys.path.insert(0,%r); import orc\n"
        "e = orc.ref() if %r=='reference' else orc.oracle()\n"
        "c = orc.make_config(3, int(sys.argv[1]), orc.PDF_GAUSS3D, [(orc.OBS_XSQUARED,0,1)], int(sys.argv[2]), steps=(1.0,))\n"
        "sys.stdin.readline(); t=time.perf_counter(); r=e.run(c); print(time.perf_counter()-t, r['avg'][0])\n"
    ) % (os.path.join(ROOT, "oracle"), kind)
    procs = [subprocess.Popen([sys.executable, "-c", code, str(1000 + p), str(nmc)], stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True)
             for p in range(nproc)]
    time.sleep(1.5)  # let every interpreter load the library before the start signal
    t0 = time.perf_counter()
    for p in procs:
        p.stdin.write("go\n")
        p.stdin.flush()
    outs = [p.communicate()[0] for p in procs]
    wall = time.perf_counter() - t0
    avgs = [float(o.split()[1]) for o in outs]
    assert all(0.3 < a < 0.7 for a in avgs), avgs
    return nproc*nmc/wall, kind, wall


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    nproc = os.cpu_count() or 1
    nmc = 4000000  # bounded sample: ~0.5 s per chain per step
    for _ in range(args.warmup):
        cpu_reference_run(max(100000, nmc//10), nproc)
    t = []
    kind = "port"
    for _ in range(args.steps):
        v, kind, wall = cpu_reference_run(nmc, nproc)
        t.append(wall)
    total = sum(t)
    value = args.steps*nproc*nmc/total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3*total/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": "%d host processes x %d steps per bench step" % (nproc, nmc)},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": nproc, "kind": kind,
                             "sample": "%d independent chains (one per core) x %d Metropolis steps of the same integrand" % (nproc, nmc)},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)  # 20 x 26 ms: long enough for the clock sampler to see the load
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import mcintegratorplusplus_b200 as m
    from mcintegratorplusplus_b200 import _capi
    if _capi.lib().mcig_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the sampling path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    mci = make_mci(m, rank, world)
    if dist is not None:
        red = torch.zeros(16, dtype=torch.float64, device="cuda")

        def allreduce(buf):  # the single collective of the path: [sum avg | sum err^2] over NVLink (src/MPIMCI.cpp:85-87)
            n = len(buf)
            red[:n].copy_(torch.from_numpy(buf))
            dist.all_reduce(red[:n])
            buf[:] = red[:n].cpu().numpy()
        mci.setAllreduce(allreduce)

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
    d2h = 8*(3*nod) + 8 + 8*2*nod
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

    if rank == 0:
        steps_per_s_kernel = float(WALKERS_PER_GPU)*NMC*args.steps/(walk_ms_max*1e-3)  # per GPU
        achieved = FP64_INSTR_PER_STEP*steps_per_s_kernel
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": dev_ms/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rng": "Philox4x32-10, 32-bit uniforms", "walkers_per_gpu": WALKERS_PER_GPU, "nmc": NMC,
                       "l2": "no HBM inputs to cache: walker state lives in registers; 1.5 MB of positions read once per step",
                       "parallelism": "walkers sharded over %d GPU(s), 1 all-reduce of 2 doubles per integrate" % world},
            "wall_ms_per_step": wall_ms/args.steps,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms/args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64_issue", "achieved": achieved/1e9, "peak": peaks[0]/1e9, "unit": "GFP64inst/s", "frac": achieved/peaks[0],
                         "traffic": WALK_DRAM_BYTES_PER_LAUNCH, "traffic_note": "dram__bytes_read + dram__bytes_write of one walk launch (ncu --set full, profiles/r01_walk_r1f_dyn_ncu_raw.csv): the 1.5 MB of start positions, nothing written; the bound is FP64/ALU issue, not HBM",
                         "kernel": "mcig_walk (JIT-specialised Metropolis walk)", "kernel_ms_per_step": walk_ms_max/args.steps,
                         "fp64_instr_per_metropolis_step": FP64_INSTR_PER_STEP, "flop_per_metropolis_step": FLOP_PER_STEP,
                         "achieved_tflops": FLOP_PER_STEP*steps_per_s_kernel/1e12, "peak_tflops": 2*peaks[0]/1e12,
                         "peak_source": "DFMA/s measured live by mcig_measure_peaks on this GPU (MEASURED_PEAKS.json has no FP64 figure)",
                         "imad_peak_ginst": peaks[1]/1e9,
                         "rng_bound": {"philox4x32_10_blocks_per_s": philox_peak, "blocks_per_step": 1, "frac": steps_per_s_kernel/philox_peak,
                                       "note": "issue-rate bound of the counter RNG alone (20 IMAD.WIDE.U32 at a quarter of the FP32 rate per block), measured live; "
                                               "the walk loop cannot exceed it whatever its FP64 content: context for the FP64 fraction above"},
                         "full_occupancy": None if full is None else {
                             "walkers": full[0], "steps_per_s": full[1], "frac": FP64_INSTR_PER_STEP*full[1]/peaks[0],
                             "note": "same kernel, 8 warps per scheduler instead of the workload's 3.46; context only, not the bench value"}},
            "clocks": clocks,
            "result": {"avg": float(avg[0]), "err": float(err[0]), "acceptance": float(rate), "cross_walker_err_local": float(mci.crossWalkerError()[0])},
        }
        if not args.no_cpu_baseline and world == 1:
            nproc = os.cpu_count() or 1
            v, kind, wall = cpu_reference_run(15000000, nproc)
            line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": nproc, "kind": kind,
                                    "sample": "%d independent chains (one per host core) x 1.5e7 Metropolis steps of the same integrand, %.1f s wall" % (nproc, wall)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
