"""Ad-hoc GPU probe: MSM timings per phase for several sizes / window sizes (not the bench)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import blaze_b200 as bz
from oracle import capi
from oracle.py import curves
from util import random_scalars, seed_points

CNAME = os.environ.get("PROBE_CURVE", "BLS12_381")
c = curves.CURVES[CNAME]
CENUM = {"BLS12_381": bz.Curve.BLS381, "BLS12_377": bz.Curve.BLS377, "BN254": bz.Curve.BN254}[CNAME]
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,22").split(",")]
cs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
check = os.environ.get("PROBE_CHECK", "1") == "1"
dc = bz.DriverClient("0")
print(dc.device_info(), flush=True)
out = []
for logn in sizes:
    n = 1 << logn
    m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, CENUM), dc)
    p0, q = seed_points(c, 71)
    t = time.time()
    m.generate_chain_points(p0 + q, 0, n, 0, 0)
    tg = time.time() - t
    sc = random_scalars(c, n, seed=72)
    exp = capi.chain_expected(CNAME, p0, q, sc, n) if check else None
    params = bz.MSMParams(n, (0, 0))
    for cb in cs:
        m.set_window_bits(cb)
        for rep in range(2):
            m.initialize(params)
            m.start_process()
            t = time.time()
            m.set_data(bz.MSMInput(None, sc, params))
            m.wait_result()
            wall = time.time() - t
            got = m.result().result
        rec = {"logn": logn, "gen_s": round(tg, 2), "plan": m.plan_info(), "phase_ms": m.phase_times(),
               "wall_s": round(wall, 4), "ok": (got == exp) if check else None,
               "mults_per_s": n / (m.phase_times()["total"] / 1e3)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    m.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "perf_probe.json"), "w"), indent=1)
