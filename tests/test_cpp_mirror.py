"""The C++ host mirror (blaze_b200/host/blaze.hpp -- the host side above the C ABI, since the reference's own
language, Rust, has no toolchain in this image) and the C++ counterparts of the reference's integration tests
(tests/cpp/integration.cpp).  CPU: the mirror compiles and links against the product library; without a device the
constructor fails loudly with NoDevice (no CPU fallback).  GPU: the tests run and pass."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "cpp", "integration_test")


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib_dir, orc_dir = os.path.join(ROOT, "blaze_b200"), os.path.join(ROOT, "oracle")
    assert os.path.exists(os.path.join(lib_dir, "libblaze_b200.so")), "build the library first (__graft_entry__.build())"
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(HERE, "cpp", "integration.cpp"), "-o", EXE,
                           "-L" + lib_dir, "-lblaze_b200", "-L" + orc_dir, "-loracle",
                           "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + orc_dir])
    return EXE


def test_cpp_mirror_compiles_and_fails_loudly_without_a_device(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present: covered by the gpu test")
    r = subprocess.run([exe, "error_behaviour"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_cpp_integration_tests(exe):
    env = dict(os.environ, MSM_SIZE="8192", ID="0")
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "9 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
