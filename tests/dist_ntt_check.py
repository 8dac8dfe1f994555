"""torchrun helper: distributed four-step NTT across WORLD_SIZE GPUs vs the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blaze_b200 as bz          # noqa: E402
from oracle import capi          # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def exchange(h):
    out = [None] * world
    dist.all_gather_object(out, h)
    return out


n = 1 << log_n
rng = np.random.default_rng(99)
d = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
d[:, 31] &= 0x3f
d = d.reshape(-1)
dc = bz.DriverClient(str(local))
t = bz.DistributedNTT(dc, log_n, rank, world, exchange=exchange, barrier=dist.barrier)
t.set_input(d)
t.run()
out = np.zeros_like(d)
t.get_output(out)
full = torch.from_numpy(out).cuda()
dist.all_reduce(full, op=dist.ReduceOp.SUM)      # every rank filled a disjoint part of a zero vector
got = full.cpu().numpy()
exp = d.copy()
capi.ntt("BLS12_381", exp, log_n)
ok = bool(np.array_equal(got, exp))
times = t.times()
t.close()
dist.barrier()
if rank == 0:
    print("DIST_NTT_OK" if ok else "DIST_NTT_MISMATCH", log_n, world, times, flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
