"""Ad-hoc GPU probe: NTT timings (not the bench)."""
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

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,24").split(",")]
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
dc = bz.DriverClient("0")
for log_n in sizes:
    n = 1 << log_n
    rng = np.random.default_rng(log_n)
    d = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    d[:, 31] &= 0x3f
    d = d.reshape(-1)
    f = bz.NTTClient.new_ex(dc, 2, log_n, False)
    i = bz.NTTClient.new_ex(dc, 2, log_n, True)
    f.initialize()
    i.initialize()
    f.set_data(bz.NTTInput(0, d))
    times = []
    for rep in range(3):
        f.set_data(bz.NTTInput(0, d))
        f.start_process(0)
        f.wait_result()
        times.append(f.phase_times()["total"])
    out = np.frombuffer(f.result(0), dtype=np.uint8)
    ok = None
    if log_n <= 24:
        exp = d.copy()
        t = time.time()
        capi.ntt("BLS12_381", exp, log_n)
        cpu_s = time.time() - t
        ok = bool(np.array_equal(out, exp))
    else:
        cpu_s = None
    i.set_data(bz.NTTInput(0, out))
    i.start_process(0)
    i.wait_result()
    back = np.frombuffer(i.result(0), dtype=np.uint8)
    rt = bool(np.array_equal(back, d))
    print(json.dumps({"log_n": log_n, "ms": times, "passes": f.phase_times()["passes"], "equals_oracle": ok,
                      "roundtrip": rt, "cpu_oracle_s": cpu_s, "inv_ms": i.phase_times()["total"]}), flush=True)
    f.close()
    i.close()
