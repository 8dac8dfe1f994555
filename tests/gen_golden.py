"""Generates tests/golden/*.json with the big-integer Python oracle (oracle/py).

The reference holds no golden vectors for this path (its tests draw from thread_rng and compare
with arkworks at run time, tests/msm/mod.rs:66-90; NTT/Poseidon golden files are external), and it
cannot be run here (Rust, FPGA).  These fixtures pin the C++ oracle -- and through it the CUDA
path -- to an independent, obviously-correct implementation of the same definitions.
Run from the repo root:  python tests/gen_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.py import curves, ec, ntt   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    for c in curves.CURVES.values():
        rng = random.Random(0xB1A2E0000 + c.code)
        G = (c.gx, c.gy)
        cases = []
        for n, factor in ((1, 1), (5, 1), (40, 1), (3, 8)):
            pts = [ec.scalar_mul(c, rng.randrange(1, c.r), G) for _ in range(n)]
            sc = [rng.randrange(c.r) for _ in range(n)]
            if n >= 5:
                sc[0], sc[1], sc[2] = 0, 1, c.r - 1
                pts[4] = pts[3]                       # duplicate point
            bases = b"".join(ec.encode_base(c, P, factor) for P in pts)
            scal = b"".join(ec.encode_scalar(s) for s in sc)
            res = ec.msm_naive(c, sc, pts)
            cases.append({"n": n, "factor": factor, "bases": bases.hex(), "scalars": scal.hex(),
                          "result": ec.encode_result(c, res).hex()})
        # generator multiples (known-answer style): k*G for a few k, and r*G = infinity
        kg = []
        for k in (1, 2, 3, 0xDEADBEEF, c.r - 1):
            kg.append({"k": hex(k), "point": ec.encode_point(c, ec.scalar_mul(c, k, G)).hex()})
        ntt_cases = []
        for log_n in (0, 1, 3, 6):
            v = [rng.randrange(c.r) for _ in range(1 << log_n)]
            ntt_cases.append({"log_n": log_n, "in": ntt.encode(v).hex(),
                              "out": ntt.encode(ntt.dft_definition(c, v)).hex()})
        json.dump({"curve": c.name, "msm": cases, "generator_multiples": kg, "ntt": ntt_cases},
                  open(os.path.join(OUT, "%s.json" % c.name.lower()), "w"), indent=1)
        print("wrote", c.name)


if __name__ == "__main__":
    main()
