"""torchrun helper: one process per GPU with a ranked DriverClient (bz_dclient_comm_init).

  * point-sharded MSM: every rank loads / generates its shard, streams its slice of the scalars, and result() returns the
    SAME full sum on every rank (NCCL all-gather of the projective partial records + combine kernel, device side);
  * four-step NTT with the handle exchange and both barriers through the client's communicator (bz_ntt_dist_run).
torch.distributed is used only to hand the NCCL unique id around and to compare outputs.
Usage: torchrun --nproc-per-node N tests/dist_msm_check.py [curve] [log_n_msm] [ntt_log_n ...]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import blaze_b200 as bz          # noqa: E402
from oracle import capi          # noqa: E402
from oracle.py import curves     # noqa: E402
from util import random_scalars, seed_points   # noqa: E402

cname = sys.argv[1] if len(sys.argv) > 1 else "BLS12_377"
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
ntt_logs = [int(a) for a in sys.argv[3:]] or [16, 21]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

dc = bz.DriverClient(str(local))
uid = [bz.DriverClient.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
dc.comm_init(rank, world, uid[0])
assert dc.comm_info() == (rank, world)

ok = True
c = curves.CURVES[cname]
curve = {"BLS12_381": bz.Curve.BLS381, "BLS12_377": bz.Curve.BLS377, "BN254": bz.Curve.BN254}[cname]
N = (1 << log_n) + 5 * world          # not a power of two; every rank the same share
per = N // world
first = rank * per
p0, q = seed_points(c, 77)
m = bz.MSMClient.new(bz.MSMInit(bz.PointMemoryType.HBM, False, curve), dc)
m.generate_chain_points(p0 + q, first, per, 0, 0)
params = bz.MSMParams(per, (0, 0))
full = [random_scalars(c, N, seed=500 + i) for i in range(3)]
exp = [capi.chain_expected(cname, p0, q, s, N) for s in full]
# two tasks in flight, then a third (window-merged table kicks in on the second use of the resident points)
for i in range(2):
    m.initialize(params)
    m.start_process()
    m.set_data(bz.MSMInput(None, full[i][first * 32:(first + per) * 32], params))
for i in range(3):
    if i == 2:
        m.start_process()
        m.set_data(bz.MSMInput(None, full[2][first * 32:(first + per) * 32], params))
    m.wait_result()
    r = m.result()
    ok &= r.result == exp[i] and r.result_label == i
m.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_MSM_OK" if int(flag) else "DIST_MSM_MISMATCH", cname, N, world, flush=True)
ok = bool(int(flag))

for ln in ntt_logs:
    n = 1 << ln
    d = random_scalars(curves.BLS12_381, n, seed=99)
    t = bz.DistributedNTT(dc, ln, rank, world)          # handles + barriers through the communicator
    t.set_input(d)
    t.run()
    out = np.zeros_like(d)
    t.get_output(out)
    fullv = torch.from_numpy(out).cuda()
    dist.all_reduce(fullv, op=dist.ReduceOp.SUM)        # every rank filled a disjoint part of a zero vector
    got = fullv.cpu().numpy()
    e = d.copy()
    capi.ntt("BLS12_381", e, ln)
    good = bool(np.array_equal(got, e))
    # a second transform of the same buffers (exercises the "exchange buffer is free again" barrier)
    t.set_input(d)
    t.run()
    out2 = np.zeros_like(d)
    t.get_output(out2)
    good &= bool(np.array_equal(out2, out))
    times = t.times()
    t.close()
    ok &= good
    if rank == 0:
        print("DIST_NTT_OK" if good else "DIST_NTT_MISMATCH", ln, world, times, flush=True)
dc.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
