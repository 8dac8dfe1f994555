"""Ad-hoc GPU probe: device-resident MSM steps, strictly serial vs two tasks in flight (BZ_MSM_TAIL=0/1 in the env
switches the tail stream).  usage: tail_probe.py "<log sizes>" [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import blaze_b200 as bz
from oracle import capi
from oracle.py import curves
from util import random_scalars, seed_points

c = curves.CURVES["BLS12_381"]
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "23").split(",")]
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dc = bz.DriverClient("0")
for logn in sizes:
    n = 1 << logn
    m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, bz.Curve.BLS381), dc)
    p0, q = seed_points(c, 71)
    m.generate_chain_points(p0 + q, 0, n, 0, 0)
    sc = random_scalars(c, n, seed=72)
    exp = capi.chain_expected("BLS12_381", p0, q, sc, n)
    params = bz.MSMParams(n, (0, 0))
    dev = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()

    def enqueue():
        m.initialize(params)
        m.start_process()
        m.set_scalars_device(dev.data_ptr(), params)

    def serial(k):
        ok = True
        for _ in range(k):
            enqueue()
            m.wait_result()
            ok &= m.result().result == exp
        return ok

    def pipe(k):
        ok = True
        enqueue()
        for i in range(k):
            if i + 1 < k:
                enqueue()
            m.wait_result()
            ok &= m.result().result == exp
        return ok

    pinned = torch.frombuffer(bytearray(sc), dtype=torch.uint8).pin_memory()

    def enqueue_host():
        m.initialize(params)
        m.start_process()
        m.set_data(bz.MSMInput(None, (pinned.data_ptr(), n * 32), params))

    def pipe_host(k):
        ok = True
        enqueue_host()
        for i in range(k):
            if i + 1 < k:
                enqueue_host()
            m.wait_result()
            ok &= m.result().result == exp
        return ok

    serial(4)
    for name, fn in (("serial", serial), ("pipelined", pipe), ("pipe_host", pipe_host), ("pipelined", pipe), ("pipe_host", pipe_host)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ok = fn(K)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / K * 1e3
        print("tail=%s 2^%d %-9s %.3f ms/step ok=%s plan c=%d phases %s" % (os.environ.get("BZ_MSM_TAIL", "1"), logn, name, dt, ok,
              m.plan_info()["c"], {k: round(v, 2) for k, v in m.phase_times().items()}), flush=True)
    m.close()
