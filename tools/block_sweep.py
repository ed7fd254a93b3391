"""Block-size sweep of the state-memory kernels (the C3 shapes of bench.py): python tools/block_sweep.py MOVE NDIM [PLACEMENT] -> steps/s per block size.
Backs the block-size rules of the engine (pick_block / pick_block_rounds in csrc/host/mcig_engine.cu)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mcintegratorplusplus_b200 as m
from mcintegratorplusplus_b200._capi import McigError
move, nd = sys.argv[1], int(sys.argv[2])
placement = int(sys.argv[3]) if len(sys.argv) > 3 else None
nmc = {"vec": 4000, "all": max(200, 32000//nd), "multistep": max(200, 8000//(2*nd)//20*20)}[move]
for bs in (0, 32, 64, 96, 128, 160, 192, 224, 256, 320, 384, 448, 512):
    mci = bench.c3_mci(m, move, nd, 65536, None)
    if placement is not None:
        mci.setStatePlacement(placement)
    mci.setBlockSize(bs)
    try:
        mci.integrate(nmc//10*2 if move != "multistep" else 40, False, False)
        best = 1e30
        for _ in range(3):
            mci.integrate(nmc, False, False)
            best = min(best, mci.timings()["walk_ms"])
        print(json.dumps({"move": move, "ndim": nd, "block": bs, "steps_per_s": 65536*nmc/(best*1e-3), "walk_ms": best}), flush=True)
    except McigError as e:
        print(json.dumps({"move": move, "ndim": nd, "block": bs, "error": str(e)[:80]}), flush=True)
