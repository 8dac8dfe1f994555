"""One Poseidon TreeC tree (for ncu captures): python scripts/poseidon_one.py height"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blaze_b200 as bz   # noqa: E402

h = int(sys.argv[1]) if len(sys.argv) > 1 else 6
nbase = 8 ** (h - 1)
total = (8 ** h - 1) // 7
rng = np.random.default_rng(77)
raw = rng.integers(0, 256, size=(nbase * 11, 32), dtype=np.uint8)
raw[:, 31] &= 0x3f
dc = bz.DriverClient("0")
pc = bz.PoseidonClient.new(bz.Hash.Poseidon, dc)
for it in range(2):
    pc.initialize(bz.PoseidonInitializeParameters(h, bz.TreeMode.TreeC, ""))
    pc.set_data(raw.reshape(-1))
    n = pc.get_num_of_pending_results()
    pc.get_raw_results(n)
    print(n, total, pc.device_ms(), flush=True)
pc.close()
dc.close()
