"""ctypes binding of oracle/liboracle.so (the C++ CPU oracle).

TEST INFRASTRUCTURE — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  `build()` compiles it with the Makefile beside it.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CURVE_CODES = {"BLS12_377": 0, "BN254": 1, "BLS12_381": 2}
FQ_BYTES = {0: 48, 1: 32, 2: 48}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        u8p, u64, i = ctypes.c_char_p, ctypes.c_uint64, ctypes.c_int
        _LIB.orc_msm_naive.argtypes = [i, u8p, u8p, u64, i, i, u8p]
        _LIB.orc_msm_pippenger.argtypes = [i, u8p, u8p, u64, i, u8p]
        _LIB.orc_point_mul.argtypes = [i, u8p, u8p, u8p]
        _LIB.orc_point_add.argtypes = [i, u8p, u8p, u8p]
        _LIB.orc_on_curve.argtypes = [i, u8p]
        _LIB.orc_normalize_result.argtypes = [i, u8p, u8p]
        _LIB.orc_chain_points.argtypes = [i, u8p, u8p, u64, ctypes.c_void_p]
        _LIB.orc_chain_expected.argtypes = [i, u8p, u8p, ctypes.c_void_p, u64, u64, u8p]
        _LIB.orc_ntt.argtypes = [i, ctypes.c_void_p, i, i, i]
        _LIB.orc_ntt_eval.argtypes = [i, ctypes.c_void_p, i, i, ctypes.c_void_p, i, ctypes.c_void_p, i]
        _LIB.orc_fq_mul.argtypes = [i, u8p, u8p, u8p]
        _LIB.orc_fr_mul.argtypes = [i, u8p, u8p, u8p]
    return _LIB


def hw_threads():
    return lib().orc_hw_threads()


def _code(curve):
    return curve if isinstance(curve, int) else CURVE_CODES[curve]


def msm_naive(curve, bases: bytes, scalars: bytes, n: int, factor=1, threads=0) -> bytes:
    c = _code(curve)
    out = ctypes.create_string_buffer(3 * FQ_BYTES[c])
    rc = lib().orc_msm_naive(c, bytes(bases), bytes(scalars), n, factor, threads or hw_threads(), out)
    assert rc == 0
    return out.raw


def msm_pippenger(curve, bases, scalars, n: int, threads=0) -> bytes:
    """bases/scalars: bytes or objects exposing the buffer protocol (numpy arrays)."""
    c = _code(curve)
    out = ctypes.create_string_buffer(3 * FQ_BYTES[c])
    b = bases if isinstance(bases, bytes) else bytes(memoryview(bases))
    s = scalars if isinstance(scalars, bytes) else bytes(memoryview(scalars))
    rc = lib().orc_msm_pippenger(c, b, s, n, threads or hw_threads(), out)
    assert rc == 0
    return out.raw


def point_mul(curve, point: bytes, k: int):
    c = _code(curve)
    out = ctypes.create_string_buffer(2 * FQ_BYTES[c])
    rc = lib().orc_point_mul(c, point, int(k).to_bytes(32, "little"), out)
    return None if rc == 1 else out.raw


def point_add(curve, p: bytes, q: bytes):
    c = _code(curve)
    out = ctypes.create_string_buffer(2 * FQ_BYTES[c])
    rc = lib().orc_point_add(c, p, q, out)
    return None if rc == 1 else out.raw


def on_curve(curve, p: bytes) -> bool:
    return lib().orc_on_curve(_code(curve), p) == 1


def normalize_result(curve, rec: bytes) -> bytes:
    c = _code(curve)
    out = ctypes.create_string_buffer(3 * FQ_BYTES[c])
    assert lib().orc_normalize_result(c, bytes(rec), out) == 0
    return out.raw


def chain_points(curve, p0: bytes, q: bytes, n: int):
    """numpy uint8 array of n wire points P0 + i*Q."""
    import numpy as np
    c = _code(curve)
    out = np.empty(n * 2 * FQ_BYTES[c], dtype=np.uint8)
    assert lib().orc_chain_points(c, p0, q, n, out.ctypes.data) == 0
    return out


def chain_expected(curve, p0: bytes, q: bytes, scalars, n: int, index_base=0) -> bytes:
    """(sum s_i) P0 + (sum (index_base+i) s_i) Q as a canonical result record.
    scalars: numpy uint8 array (n*32 bytes)."""
    c = _code(curve)
    out = ctypes.create_string_buffer(3 * FQ_BYTES[c])
    assert lib().orc_chain_expected(c, p0, q, scalars.ctypes.data, n, index_base, out) == 0
    return out.raw


def ntt(curve, data, log_n: int, inverse=False, threads=0):
    """In-place NTT of a numpy uint8 array holding 2^log_n canonical 32-byte LE elements."""
    rc = lib().orc_ntt(_code(curve), data.ctypes.data, log_n, 1 if inverse else 0, threads or hw_threads())
    assert rc == 0, rc
    return data


def ntt_eval(curve, data, log_n: int, ks, inverse=False, threads=0):
    """Outputs k in `ks` of the size-2^log_n transform of `data` (numpy uint8, canonical 32-byte LE elements), straight
    from the definition out[k] = sum_j in[j] w^(jk) (Horner, O(n) per point).  Returns a list of ints."""
    import numpy as np
    kk = np.asarray(list(ks), dtype=np.uint64)
    out = np.zeros(32 * len(kk), dtype=np.uint8)
    rc = lib().orc_ntt_eval(_code(curve), data.ctypes.data, log_n, 1 if inverse else 0, kk.ctypes.data, len(kk),
                            out.ctypes.data, threads or hw_threads())
    assert rc == 0, rc
    return [int.from_bytes(bytes(out[32 * i:32 * i + 32]), "little") for i in range(len(kk))]


def fq_mul(curve, a: bytes, b: bytes) -> bytes:
    c = _code(curve)
    out = ctypes.create_string_buffer(FQ_BYTES[c])
    lib().orc_fq_mul(c, a, b, out)
    return out.raw


def fr_mul(curve, a: bytes, b: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    lib().orc_fr_mul(_code(curve), a, b, out)
    return out.raw
