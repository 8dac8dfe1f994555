"""MSM parity on the GPU, through the C ABI (blaze_b200 Python mirror -> libblaze_b200.so).

Mirrors the call order of the reference's integration tests
(/root/reference/tests/integration_msm.rs:150-207: initialize -> start_process -> set_data ->
wait_result -> result) and checks what they check (result equals the expected MSM value,
tests/msm/mod.rs:382-420) -- but bit-exactly against the oracle on seeded inputs.
"""
import numpy as np
import pytest

import blaze_b200 as bz
from blaze_b200 import Curve, MSMClient, MSMInit, MSMInput, MSMParams, PointMemoryType

from util import CURVE_BY_NAME, chain_points, precompute_bases, random_scalars, tile

pytestmark = pytest.mark.gpu

CURVES = [("BLS12_381", Curve.BLS381), ("BLS12_377", Curve.BLS377), ("BN254", Curve.BN254)]


def run_dma(dclient, curve, points, scalars, n, precompute=False, c=0):
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, precompute, curve), dclient)
    try:
        if c:
            m.set_window_bits(c)
        params = MSMParams(n, None)
        m.initialize(params)
        m.start_process()
        m.set_data(MSMInput(points, scalars, params))
        m.wait_result()
        r = m.result()
        return r.result, r.result_label, m.plan_info()
    finally:
        m.close()


@pytest.mark.parametrize("cname,curve", CURVES)
def test_field_selftest(dclient, oracle, cname, curve):
    c = CURVE_BY_NAME[cname]
    rng = np.random.default_rng(5)
    n = 64
    vals = [int.from_bytes(rng.bytes(c.fq_bytes), "little") % c.q for _ in range(2 * n)]
    vals[0:6] = [0, 1, c.q - 1, c.q - 2, 2, c.q >> 1]
    a = b"".join(v.to_bytes(c.fq_bytes, "little") for v in vals[:n])
    b = b"".join(v.to_bytes(c.fq_bytes, "little") for v in vals[n:])
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, False, curve), dclient)
    try:
        ops = {0: lambda x, y: x * y % c.q, 1: lambda x, y: (x + y) % c.q, 2: lambda x, y: (x - y) % c.q,
               3: lambda x, y: x * x % c.q, 5: lambda x, y: (-x) % c.q,
               4: lambda x, y: pow(x, -1, c.q) if x else 0}
        for op, fn in ops.items():
            out = m.field_selftest(a, b, n, op)
            for i in range(n):
                got = int.from_bytes(out[i * c.fq_bytes:(i + 1) * c.fq_bytes], "little")
                assert got == fn(vals[i], vals[n + i]), (cname, op, i)
    finally:
        m.close()


@pytest.mark.parametrize("cname,curve", CURVES)
@pytest.mark.parametrize("n", [1, 2, 33, 1000, 1 << 12])
def test_msm_dma_vs_oracle(dclient, oracle, cname, curve, n):
    c = CURVE_BY_NAME[cname]
    pts, _, _ = chain_points(c, n, seed=n)
    sc = random_scalars(c, n, seed=100 + n)
    got, label, plan = run_dma(dclient, curve, pts, sc, n)
    exp = oracle.msm_pippenger(cname, pts, sc, n)
    assert got == exp, (cname, n, plan)


@pytest.mark.parametrize("c_bits", [4, 7, 11, 16])
def test_msm_window_sizes(dclient, oracle, c_bits):
    c = CURVE_BY_NAME["BLS12_381"]
    n = 3000
    pts, _, _ = chain_points(c, n, seed=3)
    sc = random_scalars(c, n, seed=4)
    got, _, plan = run_dma(dclient, Curve.BLS381, pts, sc, n, c=c_bits)
    assert plan["c"] == c_bits
    assert got == oracle.msm_pippenger("BLS12_381", pts, sc, n)


def test_msm_edge_scalars(dclient, oracle):
    """0, 1, r-1, powers of two, all-ones digits: exercises digit carries and the top window."""
    c = CURVE_BY_NAME["BLS12_381"]
    special = [0, 1, 2, c.r - 1, c.r - 2, (1 << 254), (1 << 255) % c.r, (1 << 128) - 1, 0x8000, 0x7fff, 0xffff,
               (c.r - 1) // 2, (c.r + 1) // 2]
    n = len(special)
    pts, _, _ = chain_points(c, n, seed=9)
    sc = np.frombuffer(b"".join(int(s).to_bytes(32, "little") for s in special), dtype=np.uint8).copy()
    for cb in (0, 5, 16):
        got, _, _ = run_dma(dclient, Curve.BLS381, pts, sc, n, c=cb)
        assert got == oracle.msm_naive("BLS12_381", bytes(pts), bytes(sc), n, 1)


@pytest.mark.parametrize("kind", ["zeros_and_ones", "small_10bit", "all_equal", "sparse"])
def test_msm_skewed_scalar_distributions(dclient, oracle, kind):
    """Distributions real provers produce: mostly 0/1 witnesses, range-checked small values, one
    repeated scalar (a single bucket per window holds everything), mostly-zero vectors.  Exercises
    the zero-digit drop, the segment walk over huge buckets and the partial-merge tree."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 20000
    pts, _, _ = chain_points(c, n, seed=77)
    rng = np.random.default_rng(78)
    if kind == "zeros_and_ones":
        vals = rng.integers(0, 2, size=n).tolist()
    elif kind == "small_10bit":
        vals = rng.integers(0, 1024, size=n).tolist()
    elif kind == "all_equal":
        vals = [0x1d3c5b7a99f0e1d2c3b4a5968778695a4b3c2d1e0f] * n
    else:
        vals = [0] * n
        for i in rng.integers(0, n, size=50):
            vals[int(i)] = int.from_bytes(rng.bytes(31), "little")
    sc = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).copy()
    exp = oracle.msm_pippenger("BLS12_381", pts, sc, n)
    for cb in (0, 9, 13):
        got, _, plan = run_dma(dclient, Curve.BLS381, pts, sc, n, c=cb)
        assert got == exp, (kind, plan)


def test_msm_result_infinity(dclient):
    c = CURVE_BY_NAME["BLS12_381"]
    n = 4
    pts, _, _ = chain_points(c, n, seed=1)
    sc = np.zeros(n * 32, dtype=np.uint8)
    got, _, _ = run_dma(dclient, Curve.BLS381, pts, sc, n)
    assert got == (0).to_bytes(48, "little") + (1).to_bytes(48, "little") + (0).to_bytes(48, "little")


def test_msm_cancellation_and_duplicates(dclient, oracle):
    """P and -P and repeated points in one bucket: complete addition (doubling / inverse)."""
    c = CURVE_BY_NAME["BLS12_381"]
    base, _, _ = chain_points(c, 4, seed=2)
    ps = c.point_size
    p = bytes(base[:ps])
    x, y = p[:48], int.from_bytes(p[48:], "little")
    negp = x + ((c.q - y) % c.q).to_bytes(48, "little")
    pts = np.frombuffer(p + negp + p + p + p + bytes(base[ps:2 * ps]), dtype=np.uint8).copy()
    s = 0x1234567
    sc = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in [s, s, s, s, s, 5]), dtype=np.uint8).copy()
    got, _, _ = run_dma(dclient, Curve.BLS381, pts, sc, 6)
    assert got == oracle.msm_naive("BLS12_381", bytes(pts), bytes(sc), 6, 1)


@pytest.mark.parametrize("cname,curve", CURVES)
def test_msm_reference_tiled_distribution(dclient, oracle, cname, curve):
    """The reference's own large-input shape: 256 random pairs tiled to n (tests/msm/mod.rs:21-31,
    92-109): every bucket holds many copies of the same point."""
    c = CURVE_BY_NAME[cname]
    n = 256 * 37 + 19
    pts256, _, _ = chain_points(c, 256, seed=11)
    sc256 = random_scalars(c, 256, seed=12)
    pts = tile(pts256, c.point_size, 256, n)
    sc = tile(sc256, 32, 256, n)
    got, _, _ = run_dma(dclient, curve, pts, sc, n)
    assert got == oracle.msm_pippenger(cname, pts, sc, n)


@pytest.mark.parametrize("cname,curve", [("BLS12_381", Curve.BLS381), ("BN254", Curve.BN254)])
def test_msm_precompute_factor8(dclient, oracle, cname, curve):
    """x8 precomputed bases (tests/msm/mod.rs:360-380): the core pairs 32-bit scalar limbs with
    the stored 2^(32 i) P; the result must equal the plain MSM."""
    c = CURVE_BY_NAME[cname]
    n = 300
    pts, _, _ = chain_points(c, n, seed=21)
    sc = random_scalars(c, n, seed=22)
    bases8 = precompute_bases(c, pts, n, 8)
    got, _, plan = run_dma(dclient, curve, bases8, sc, n, precompute=True)
    assert got == oracle.msm_pippenger(cname, pts, sc, n)
    assert got == oracle.msm_naive(cname, bytes(bases8), bytes(sc), n, 8)


def test_msm_hbm_mode_and_labels(dclient, oracle):
    """integration_msm_hbm.rs:121-226 call order: load points once, then scalars-only tasks;
    get_data_from_hbm returns the bytes written; labels increase per task."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 2048 + 77
    pts, p0, q = chain_points(c, n, seed=31)
    addr, off = 0x0, 0x0
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        params = MSMParams(n, (addr, off))
        m.load_data_to_hbm(pts, addr, off)
        assert m.get_data_from_hbm(len(pts), addr, off) == bytes(pts)
        labels = []
        for it in range(3):
            sc = random_scalars(c, n, seed=40 + it)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
            m.wait_result()
            r = m.result()
            labels.append(r.result_label)
            assert r.result == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert labels == sorted(labels) and len(set(labels)) == 3
        assert m.nof_elements() == n
        assert m.is_msm_engine_ready() == 1
        # (points, hbm addr) mode: preload + compute in one call (msm_api.rs:203-216)
        sc = random_scalars(c, n, seed=50)
        m.start_process()
        m.set_data(MSMInput(pts, sc, MSMParams(n, (0x1000000, 0))))
        m.wait_result()
        assert m.result().result == oracle.chain_expected("BLS12_381", p0, q, sc, n)
    finally:
        m.close()


def test_msm_device_generator_matches_oracle(dclient, oracle):
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1000
    pts, p0, q = chain_points(c, n, seed=61)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.generate_chain_points(p0 + q, 0, n, 0x2000000, 0)
        assert m.get_data_from_hbm(n * 96, 0x2000000, 0) == bytes(pts)
    finally:
        m.close()


def test_msm_errors(dclient):
    c = CURVE_BY_NAME["BLS12_381"]
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.initialize(MSMParams(16, None))          # HBM without an address: the reference panics
        with pytest.raises(bz.error.NoResult):
            m.wait_result()
        pts, _, _ = chain_points(c, 4, seed=1)
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.set_data(MSMInput(pts, np.zeros(3 * 32, dtype=np.uint8), MSMParams(4, (0, 0))))   # short scalars
        # non-canonical scalar (>= r) is reported, not silently reduced
        bad = np.full(4 * 32, 0xff, dtype=np.uint8)
        m.initialize(MSMParams(4, (0, 0)))
        m.start_process()
        m.set_data(MSMInput(pts, bad, MSMParams(4, (0, 0))))
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.wait_result()
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.result()
    finally:
        m.close()


@pytest.mark.parametrize("log_n", [16, 20])
def test_msm_large_closed_form(dclient, oracle, log_n):
    """Config 1 size (2^16) vs the full oracle, and 2^20 vs the closed form of the chain workload."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << log_n
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        from util import seed_points
        p0, q = seed_points(c, 71)
        m.generate_chain_points(p0 + q, 0, n, 0, 0)
        sc = random_scalars(c, n, seed=72)
        params = MSMParams(n, (0, 0))
        m.initialize(params)
        m.start_process()
        m.set_data(MSMInput(None, sc, params))
        m.wait_result()
        got = m.result().result
        assert got == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        if log_n == 16:
            pts = np.frombuffer(m.get_data_from_hbm(n * 96, 0, 0), dtype=np.uint8)
            assert got == oracle.msm_pippenger("BLS12_381", pts, sc, n)
    finally:
        m.close()


def test_config3_bn254_2p24_dma_mode(dclient, oracle):
    """BASELINE.json configs[2]: BN254 MSM 2^24, DMA mode (points AND scalars streamed from host
    memory with the call).  Checked against the closed form of the chain workload (full-size,
    size-independent property) -- the full oracle Pippenger would take minutes."""
    c = CURVE_BY_NAME["BN254"]
    n = 1 << 24
    from util import seed_points
    p0, q = seed_points(c, 91)
    gen = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BN254), dclient)   # BN254 x HBM: todo!() in the reference
    try:
        gen.generate_chain_points(p0 + q, 0, n, 0, 0)
        pts = np.frombuffer(gen.get_data_from_hbm(n * 64, 0, 0), dtype=np.uint8)
    finally:
        gen.close()
    sc = random_scalars(c, n, seed=92)
    got, _, plan = run_dma(dclient, Curve.BN254, pts, sc, n)
    assert got == oracle.chain_expected("BN254", p0, q, sc, n), plan
    # spot-check the generated points against the oracle's own chain
    exp_pts, _, _ = chain_points(c, 64, seed=91)
    assert bytes(pts[:64 * 64]) == bytes(exp_pts)


def test_config5_bls12_377_hbm_closed_form(dclient, oracle):
    """BASELINE.json configs[4] curve (BLS12-377), HBM-resident points, 2^20 per GPU shard."""
    c = CURVE_BY_NAME["BLS12_377"]
    n = 1 << 20
    from util import seed_points
    p0, q = seed_points(c, 93)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS377), dclient)
    try:
        first = 12345                      # a shard that does not start at index 0
        m.generate_chain_points(p0 + q, first, n, 0x4000000, 0)
        sc = random_scalars(c, n, seed=94)
        params = MSMParams(n, (0x4000000, 0))
        m.initialize(params)
        m.start_process()
        m.set_data(MSMInput(None, sc, params))
        m.wait_result()
        assert m.result().result == oracle.chain_expected("BLS12_377", p0, q, sc, n, index_base=first)
    finally:
        m.close()


def test_msm_precompute_hbm_resident(dclient, oracle):
    """integration_msm_hbm.rs `hbm_msm_bls12_381_precomp_test` shape: x8 bases loaded once into HBM
    (mem_type DMA + an hbm address => HBM mode, msm_api.rs:75-95), then scalars-only tasks."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 2048 + 5
    pts, _, _ = chain_points(c, n, seed=95)
    bases8 = precompute_bases(c, pts, n, 8)
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, True, Curve.BLS381), dclient)
    try:
        addr = (0x8000000, 0x100)
        params = MSMParams(n, addr)
        m.load_data_to_hbm(bases8, *addr)
        for it in range(2):
            sc = random_scalars(c, n, seed=96 + it)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
            m.wait_result()
            assert m.result().result == oracle.msm_pippenger("BLS12_381", pts, sc, n)
    finally:
        m.close()


def test_msm_task_queue_two_in_flight(dclient, oracle):
    """Two tasks queued back to back (the copy of the second overlaps the kernels of the first);
    results pop in FIFO order with increasing labels (msm_api.rs:260-273)."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << 14
    pts, p0, q = chain_points(c, n, seed=97)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        params = MSMParams(n, (0, 0))
        m.load_data_to_hbm(pts, 0, 0)
        scs = [random_scalars(c, n, seed=98 + i) for i in range(3)]
        for sc in scs:
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
        labels = []
        for sc in scs:
            m.wait_result()
            r = m.result()
            labels.append(r.result_label)
            assert r.result == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert labels == [labels[0], labels[0] + 1, labels[0] + 2]
    finally:
        m.close()


# ---- window-merged table (HBM-resident precomputed points: entry (w, i) = 2^(c w) P_i, one bucket set)
def run_hbm(m, params, sc):
    m.initialize(params)
    m.start_process()
    m.set_data(MSMInput(None, sc, params))
    m.wait_result()
    return m.result().result


@pytest.mark.parametrize("cname,curve", CURVES)
@pytest.mark.parametrize("c_bits", [0, 6, 11])
def test_msm_merged_table_vs_oracle(dclient, oracle, cname, curve, c_bits):
    c = CURVE_BY_NAME[cname]
    n = 3000
    pts, _, _ = chain_points(c, n, seed=201)
    pts = pts.copy()
    ps = c.point_size
    pts[5 * ps:6 * ps] = 0                      # an identity padding record (0, 0)
    pts[7 * ps:8 * ps] = pts[6 * ps:7 * ps]     # a duplicate point
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, curve), dclient)
    try:
        m.set_precompute(2)
        if c_bits:
            m.set_window_bits(c_bits)
        params = MSMParams(n, (0x100, 0))
        m.load_data_to_hbm(pts, 0x100, 0)
        for it in range(2):
            sc = random_scalars(c, n, seed=202 + it)
            if it == 1:
                sc[7 * 32:8 * 32] = sc[6 * 32:7 * 32]   # same digit for the duplicate point: doubling inside a bucket
            got = run_hbm(m, params, sc)
            plan = m.plan_info()
            assert plan["merged_table"] and plan["bucket_sets"] == 1 and plan["windows"] > 1, plan
            if c_bits:
                assert plan["c"] == c_bits
            assert got == oracle.msm_pippenger(cname, pts, sc, n), (cname, plan)
    finally:
        m.close()


def test_msm_merged_table_built_on_reuse(dclient, oracle):
    """Default policy: the first MSM on a freshly loaded point set runs on the plain table, the table of
    window multiples is derived when the set is used again; new points invalidate it.  Same bytes either way."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 5000
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        params = MSMParams(n, (0, 0))
        for rnd in range(2):
            pts, p0, q = chain_points(c, n, seed=210 + rnd)
            m.load_data_to_hbm(pts, 0, 0)
            flags = []
            for it in range(3):
                sc = random_scalars(c, n, seed=220 + 10 * rnd + it)
                assert run_hbm(m, params, sc) == oracle.chain_expected("BLS12_381", p0, q, sc, n)
                flags.append(m.plan_info()["merged_table"])
            assert flags == [False, True, True], flags
        m.set_precompute(0)
        sc = random_scalars(c, n, seed=240)
        assert run_hbm(m, params, sc) == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert not m.plan_info()["merged_table"]
    finally:
        m.close()


def test_msm_merged_table_skewed_and_tiled(dclient, oracle):
    """Tiled inputs (tests/msm/mod.rs:92-109) and a constant scalar through the merged table: one bucket of
    the single bucket set receives entries from every window."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 256 * 20 + 3
    pts256, _, _ = chain_points(c, 256, seed=250)
    sc256 = random_scalars(c, 256, seed=251)
    pts = tile(pts256, c.point_size, 256, n)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.set_precompute(2)
        params = MSMParams(n, (0, 0))
        m.load_data_to_hbm(pts, 0, 0)
        sc = tile(sc256, 32, 256, n)
        assert run_hbm(m, params, sc) == oracle.msm_pippenger("BLS12_381", pts, sc, n)
        const = np.frombuffer(b"".join([int(0x0123456789abcdef0123456789abcdef0123456789abcdef).to_bytes(32, "little")] * n),
                              dtype=np.uint8).copy()
        assert run_hbm(m, params, const) == oracle.msm_pippenger("BLS12_381", pts, const, n)
    finally:
        m.close()


def test_msm_merged_table_2p20_closed_form(dclient, oracle):
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << 20
    from util import seed_points
    p0, q = seed_points(c, 260)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.set_precompute(2)
        m.generate_chain_points(p0 + q, 0, n, 0, 0)
        params = MSMParams(n, (0, 0))
        for it in range(2):
            sc = random_scalars(c, n, seed=261 + it)
            assert run_hbm(m, params, sc) == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert m.plan_info()["merged_table"]
    finally:
        m.close()


@pytest.mark.parametrize("mode,c_bits", [(0, 17), (0, 23), (2, 18), (2, 24), (2, 26)])   # 26: three partition levels
def test_msm_wide_windows_small_input(dclient, oracle, mode, c_bits):
    """Wide windows on a small input: two partition levels of the sort with (almost) empty parents, millions of
    empty buckets in the reduction, plain table (mode 0) and window-merged table (mode 2)."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 5000
    pts, p0, q = chain_points(c, n, seed=300 + c_bits)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.set_precompute(mode)
        m.set_window_bits(c_bits)
        params = MSMParams(n, (0, 0))
        m.load_data_to_hbm(pts, 0, 0)
        sc = random_scalars(c, n, seed=301)
        assert run_hbm(m, params, sc) == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        plan = m.plan_info()
        assert plan["c"] == c_bits and plan["merged_table"] == (mode == 2), plan
    finally:
        m.close()


@pytest.mark.parametrize("n", [1, 2, 33])
def test_msm_merged_table_tiny_inputs(dclient, oracle, n):
    c = CURVE_BY_NAME["BN254"]
    pts, _, _ = chain_points(c, n, seed=400 + n)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BN254), dclient)
    try:
        m.set_precompute(2)
        params = MSMParams(n, (0x40, 0))
        m.load_data_to_hbm(pts, 0x40, 0)
        for it in range(2):
            sc = random_scalars(c, n, seed=410 + it)
            assert run_hbm(m, params, sc) == oracle.msm_naive("BN254", bytes(pts), bytes(sc), n, 1)
        assert m.plan_info()["merged_table"]
    finally:
        m.close()


def test_msm_raw_projective_result_and_combine(dclient, oracle):
    """Shards of a point-sharded multi-GPU MSM keep their record projective (Z != 1: the reference's own result
    format, tests/msm/mod.rs:397-403); combine_results sums the shard records and normalises once."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 4000
    pts, p0, q = chain_points(c, n, seed=500)
    sc = random_scalars(c, n, seed=501)
    exp = oracle.chain_expected("BLS12_381", p0, q, sc, n)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.set_raw_result(True)
        half = n // 2
        recs = []
        for lo, cnt, addr in ((0, half, 0), (half, n - half, 0x1000000)):
            m.load_data_to_hbm(pts[lo * 96:(lo + cnt) * 96], addr, 0)
            params = MSMParams(cnt, (addr, 0))
            rec = run_hbm(m, params, sc[lo * 32:(lo + cnt) * 32])
            assert rec[:48] != (1).to_bytes(48, "little")          # not normalised
            recs.append(rec)
        assert oracle.normalize_result("BLS12_381", recs[0]) == oracle.chain_expected("BLS12_381", p0, q, sc[:half * 32], half)
        assert m.combine_results(b"".join(recs), 2) == exp
        m.set_raw_result(False)
        assert run_hbm(m, MSMParams(half, (0, 0)), sc[:half * 32]) == oracle.normalize_result("BLS12_381", recs[0])
    finally:
        m.close()
