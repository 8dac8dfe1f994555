"""CPU: a big-integer model of the GPU NTT pass (same index maps, DIF rounds, digit reversal and
twiddle exponents as blaze_b200/csrc/ntt.cu + the planner in ntt_api.cu) against the oracle.
Keeps the kernel's algorithm checkable on a machine without a GPU."""
import subprocess
import sys
import os


def test_ntt_pass_model_matches_oracle():
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "ntt_model.py")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "13 [7, 6] ok" in out.stdout


def test_distributed_four_step_model_matches_oracle():
    """same for the multi-GPU plan of bz_ntt_dist_step1/step3 (column slabs, fused twiddle +
    exchange addressing, transposed intermediate), G ranks simulated in one process."""
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, os.path.join(here, "ntt_dist_model.py")], capture_output=True, text=True)
    assert "3 12 4 [3, 3] [3, 3] ok" in out.stdout, out.stdout + out.stderr
