"""Does a streaming `nvidia-smi -lms` child disturb the integrate calls it is meant to watch? total_ms / walk_ms of 20 headline integrates without a sampler,
with bench.py's query, and with reduced queries / longer periods."""
import sys, os, json, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import mcintegratorplusplus_b200 as m
mci = g._bench_mci(m)
for _ in range(3):
    mci.integrate(100000, False, False)
FULL = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
cases = [("none", None, 0), ("full/100ms", FULL, 100), ("none", None, 0), ("clocks.sm/100ms", "clocks.sm", 100), ("power/100ms", "power.draw", 100),
         ("reasons/100ms", "clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_power_cap", 100), ("full/500ms", FULL, 500), ("none", None, 0)]
for name, q, ms in cases:
    proc = None
    if q:
        proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", str(ms)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.5)
    tot = walk = 0.
    t0 = time.perf_counter()
    for _ in range(20):
        mci.integrate(100000, False, False)
        t = mci.timings()
        tot += t["total_ms"]; walk += t["walk_ms"]
    wall = 1e3*(time.perf_counter() - t0)
    if proc:
        proc.terminate(); proc.wait()
    print(json.dumps({"sampler": name, "total_ms": tot/20, "walk_ms": walk/20, "wall_ms": wall/20}), flush=True)
