"""One accumulate mode, a few MSMs (for ncu captures): python scripts/ba_one.py log_n mode rounds [precompute_mode=2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import blaze_b200 as bz                      # noqa: E402
from oracle.py import curves                 # noqa: E402
from util import random_scalars, seed_points  # noqa: E402

log_n, mode, rounds = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
c = curves.BLS12_381
n = 1 << log_n
dc = bz.DriverClient("0")
p0, q = seed_points(c, 2026)
m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, bz.Curve.BLS381), dc)
m.generate_chain_points(p0 + q, 0, n, 0, 0)
m.set_precompute(int(sys.argv[4]) if len(sys.argv) > 4 else 2)
m.set_accumulate_mode(mode, rounds)
params = bz.MSMParams(n, (0, 0))
sc = random_scalars(c, n, seed=2027)
for it in range(3):
    m.initialize(params)
    m.start_process()
    m.set_data(bz.MSMInput(None, sc, params))
    m.wait_result()
    m.result()
    print(m.phase_times(), m.plan_info(), flush=True)
m.close()
dc.close()
