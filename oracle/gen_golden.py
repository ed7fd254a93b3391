"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libmci_ref.so) — TEST INFRASTRUCTURE ONLY.

    make -C oracle ref && python oracle/gen_golden.py

Only runs where /root/reference was compiled (this container). The committed JSON travels to the GPU box, where the
C oracle and the CUDA path are checked against it. Floats are stored as C99 hex strings (bit-exact).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np  # noqa: E402

import configs  # noqa: E402
import orc  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")


def hx(v):
    return [float(x).hex() for x in v]


def main():
    ref = orc.ref()
    runs = {}
    for name in configs.RUNS:
        cfg = configs.make(name)
        want_trace = cfg.nmc <= 100000
        r = ref.run(cfg, trace=want_trace)
        entry = {"avg": hx(r["avg"]), "err": hx(r["err"]), "acc_rate": float(r["acc_rate"]).hex(), "x_final": hx(r["x_final"]),
                 "steps_final": hx(r["steps_final"]), "n_acc": int(r["n_acc"]), "n_rej": int(r["n_rej"])}
        if want_trace:
            acc = r["accepted"]
            entry["n_steps"] = int(len(acc))
            entry["accepted_head"] = "".join(str(int(b)) for b in acc[:256])
            entry["accepted_crc"] = int(np.frombuffer(np.packbits(acc).tobytes(), dtype=np.uint8).astype(np.uint64).dot(
                (np.arange(len(np.packbits(acc)), dtype=np.uint64) % 65521 + 1)) % (2**61 - 1))
            entry["n_draws"] = int(r["n_draws"])
            entry["draws_head"] = hx(r["draws"][:16])
        runs[name] = entry
        print(name, r["avg"][:2], r["err"][:2], r["acc_rate"])
    with open(os.path.join(OUT, "ref_runs.json"), "w") as f:
        json.dump(runs, f, indent=1, sort_keys=True)

    # periodic text dumps (storeObservablesOnFile / storeWalkerPositionsOnFile) of one small run
    import ctypes as C
    f = ref.lib.mciref_run_with_files
    f.restype = C.c_int
    f.argtypes = [C.POINTER(orc.Config), C.POINTER(orc.Result), C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    cfg = configs.make("dump_files")
    res = orc.Result()
    assert f(C.byref(cfg), C.byref(res), os.path.join(OUT, "dump_observables.txt").encode(), configs.DUMP_OBS_FREQ,
             os.path.join(OUT, "dump_walker.txt").encode(), configs.DUMP_WLK_FREQ) == 0
    print("dump_files", list(res.avg[:res.nobsdim]))

    # step callback (MCI::setCallback) folded into four sums by the harness
    fcb = ref.lib.mciref_run_callback
    fcb.restype = C.c_int
    fcb.argtypes = [C.POINTER(orc.Config), C.POINTER(orc.Result), C.POINTER(C.c_double)]
    cbs = {}
    for name in configs.CALLBACK_RUNS:
        cfg = configs.make(name)
        res = orc.Result()
        buf = (C.c_double*4)()
        assert fcb(C.byref(cfg), C.byref(res), buf) == 0
        cbs[name] = {"buf": hx(buf), "acc_rate": float(res.acc_rate).hex(), "avg": hx(res.avg[:res.nobsdim])}
        print("callback", name, list(buf))
    with open(os.path.join(OUT, "ref_callback.json"), "w") as f:
        json.dump(cbs, f, indent=1, sort_keys=True)

    est = {}
    for wname, (pdf, nmc, ndim, step, cp, seed) in configs.WALKS.items():
        datax, datacc, nchanged, cidx, rate = ref.testwalk(pdf, nmc, ndim, step, cp, seed)
        e = {"acc_rate": float(rate).hex(), "x_first": hx(datax[0]), "x_last": hx(datax[-1])}
        for ename, et in (("uncorrelated", orc.EST_UNCORRELATED), ("correlated", orc.EST_CORRELATED), ("fcblocker", orc.EST_FCBLOCKER),
                          ("mjblocker", orc.EST_MJBLOCKER), ("noop", orc.EST_NOOP)):
            if et == orc.EST_MJBLOCKER and (nmc & (nmc - 1)) != 0:
                continue
            avg, err = ref.estimate(et, datax if ndim > 1 else datax[:, 0])
            e[ename] = {"avg": hx(avg), "err": hx(err)}
        avg, err = ref.estimate(orc.EST_BLOCK, datax if ndim > 1 else datax[:, 0], nblocks=nmc // 16)
        e["block16"] = {"avg": hx(avg), "err": hx(err)}
        # accumulators driven by hand through the walk, as test/ut1/main.cpp:116-139
        if wname == "ut1_gauss":
            for label, obs_id, bs, ns in (("simple_xnd", orc.OBS_XND, 0, 1), ("block4_xnd", orc.OBS_XND, 4, 1), ("full_xnd", orc.OBS_XND, 1, 1),
                                          ("full_skip2_updxnd", orc.OBS_UPDXND, 1, 2), ("block4_skip2_x2", orc.OBS_X2, 4, 2)):
                d = ref.accumulate(obs_id, ndim, bs, ns, datax, datacc, nchanged, cidx)
                e["accu_" + label] = {"nstore": int(d.shape[0]), "col_sums": hx(d.sum(axis=0)), "first": hx(d[0]), "last": hx(d[-1])}
        est[wname] = e
        print(wname, rate)
    with open(os.path.join(OUT, "ref_estimators.json"), "w") as f:
        json.dump(est, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
