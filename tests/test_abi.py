"""CPU: the C-ABI library loads, exports every symbol include/blaze_b200.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "blaze_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bz_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from blaze_b200._lib import LIB_PATH, SIGNATURES
    assert os.path.exists(LIB_PATH), "build the library first (python -c 'import __graft_entry__ as g; g.build()')"
    L = ctypes.CDLL(LIB_PATH)
    syms = declared_symbols()
    assert len(syms) > 30
    for s in syms:
        assert hasattr(L, s), "header declares %s but the library does not export it" % s
    # the Python binding covers the whole header, nothing more
    assert sorted(SIGNATURES) == syms


def test_status_codes_match_error_variants():
    import blaze_b200 as bz
    from blaze_b200.error import _BY_CODE
    # reference error.rs:6-32 has 8 variants; codes -1..-8 map onto them in declaration order
    names = [_BY_CODE[-i].variant for i in range(1, 9)]
    assert names == ["WriteError", "ReadError", "HBICAPNotReady", "InvalidPrimitiveParam", "CsvError", "LoadFailed",
                     "FileError", "Unknown"]
    assert issubclass(bz.error.InvalidPrimitiveParam, bz.error.DriverClientError)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import blaze_b200 as bz
    with pytest.raises(bz.error.NoDevice):
        bz.DriverClient("0", bz.DriverConfig.driver_client_cfg(bz.CardType.B200))


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (or any CPU fallback)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "blaze_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt, f


def test_image_parameter_word_roundtrip():
    from blaze_b200 import MSMImageParametrs
    word = (1 << 4) | (21 << 8) | (0xF << 16) | (2 << 20)
    p = MSMImageParametrs.parse_image_params(word)
    assert (p.hif2_cpu_c_curve, p.hif2_cpu_c_buckets_mem_addr_width, p.hif2_cpu_c_number_of_segments,
            p.hif2_cpu_c_number_of_ec_adders, p.hif2cpu_c_is_stub) == (2, 21, 1, 15, 0)
