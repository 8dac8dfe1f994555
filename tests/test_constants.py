"""CPU: the generated product constants (blaze_b200/csrc/field_constants.h, from
tools/gen_constants.py) agree with the oracle's independent copy of the curve parameters."""
import os
import re

from oracle.py import curves

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_arrays():
    src = open(os.path.join(ROOT, "blaze_b200", "csrc", "field_constants.h")).read()
    out = {}
    for name, body in re.findall(r"static const uint32_t (\w+)_H\[\d+\] = \{([^}]*)\};", src):
        limbs = [int(x.strip().rstrip("u"), 16) for x in body.split(",")]
        out[name] = sum(v << (32 * i) for i, v in enumerate(limbs)), len(limbs)
    return out, src


def test_moduli_and_montgomery_constants():
    arr, src = parse_arrays()
    pairs = {"FQ381": curves.BLS12_381.q, "FR381": curves.BLS12_381.r, "FQ377": curves.BLS12_377.q,
             "FR377": curves.BLS12_377.r, "FQ254": curves.BN254.q, "FR254": curves.BN254.r}
    for name, p in pairs.items():
        mod, n = arr[name + "_MOD"]
        assert mod == p
        R = 1 << (32 * n)
        assert arr[name + "_ONE"][0] == R % p
        assert arr[name + "_R2"][0] == R * R % p
        inv = int(re.search(r"struct F%s \{.*?INV = 0x([0-9a-f]+)u" % name[1:].lower(), src, re.S).group(1), 16)
        assert (inv * p + 1) % (1 << 32) == 0
    for name, c in (("FR381", curves.BLS12_381), ("FR377", curves.BLS12_377), ("FR254", curves.BN254)):
        n = arr[name + "_MOD"][1]
        R = 1 << (32 * n)
        root = curves.root_of_unity(c, c.fr_two_adicity)
        assert arr[name + "_ROOT"][0] == root * R % c.r
        assert arr[name + "_ROOT_INV"][0] == pow(root, -1, c.r) * R % c.r
