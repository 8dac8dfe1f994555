"""Ad-hoc GPU probe: one NTT of 2^log_n (device-resident random data), prints the pass timings."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import blaze_b200 as bz
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dc = bz.DriverClient("0")
t = bz.NTTClient.new_ex(dc, 2, log_n, False)
t.initialize()
n = 1 << log_n
class V:
    def __init__(s, p, nb): s.__cuda_array_interface__ = {"shape": (nb,), "typestr": "|u1", "data": (p, False), "version": 2}
v = torch.as_tensor(V(t.slot_device_ptr(0), n * 32), device="cuda")
v.copy_(torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda"))
v.view(n, 32)[:, 31] &= 0x3f
torch.cuda.synchronize()
for i in range(reps):
    t.start_process(0); t.wait_result()
    print(log_n, t.phase_times(), flush=True)
t.close()
