"""GPU probe: bucket-accumulation kernel time, XYZZ sweep vs fused batched-affine sweep, HBM-resident BLS12-381 points
with the window-merged table.  python scripts/ba_probe.py [log_n ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import blaze_b200 as bz                      # noqa: E402
from oracle import capi                      # noqa: E402
from oracle.py import curves                 # noqa: E402
from util import random_scalars, seed_points  # noqa: E402

logs = [int(a) for a in sys.argv[1:]] or [24]
c = curves.BLS12_381
dc = bz.DriverClient("0")
for log_n in logs:
    n = 1 << log_n
    p0, q = seed_points(c, 2026)
    m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, bz.Curve.BLS381), dc)
    m.generate_chain_points(p0 + q, 0, n, 0, 0)
    m.set_precompute(2)
    params = bz.MSMParams(n, (0, 0))
    sc = random_scalars(c, n, seed=2027)
    exp = capi.chain_expected("BLS12_381", p0, q, sc, n)
    for mode, rounds in ((0, -1), (2, 2), (2, 3), (2, 4), (2, 5), (0, -1), (2, -1)):
        m.set_accumulate_mode(mode, rounds)
        ts = []
        for it in range(3):
            m.initialize(params)
            m.start_process()
            m.set_data(bz.MSMInput(None, sc, params))
            m.wait_result()
            r = m.result().result
            ts.append(m.phase_times())
        ok = r == exp
        t = ts[-1]
        print("2^%d mode %d rounds %2d ok=%s total %.2f sort %.2f accumulate %.2f reduce %.2f  plan %s" %
              (log_n, mode, rounds, ok, t["total"], t["sort"], t["accumulate"], t["reduce"], m.plan_info()), flush=True)
    m.close()
dc.close()
