"""GPU: the multi-device MSMClient (one client over a device list), the card address space (virtual-memory arena, range-
scoped invalidation), the task-queue corner cases the round-1 review flagged, the get_api() register file, published
known-answer values through the CUDA path, and the 2^27 transform of BASELINE.json's configs[3] checked against the
definition of its outputs.

A device list may name one device several times ("0,0,0"): every code path of the sharded client -- splitting
load_data_to_hbm / set_data, one pipeline per member, peer copies of the partial records, the device-side sum -- then
runs on a one-GPU box; with >= 2 GPUs the same tests also run over distinct devices.
"""
import json
import os

import numpy as np
import pytest

import blaze_b200 as bz
from blaze_b200 import Curve, DriverClient, MSMClient, MSMInit, MSMInput, MSMParams, PointMemoryType

from util import CURVE_BY_NAME, chain_points, precompute_bases, random_scalars, seed_points

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def device_lists():
    import torch
    n = torch.cuda.device_count()
    ids = ["0,0", "0,0,0"]
    if n >= 2:
        ids.append("0,1")
    if n >= 4:
        ids.append("0,1,2,3")
    return ids


@pytest.fixture(scope="module", params=["0,0", "0,0,0", "multi"])
def group_dc(request):
    import torch
    if request.param == "multi":
        n = torch.cuda.device_count()
        if n < 2:
            pytest.skip("needs >= 2 GPUs")
        ids = ",".join(str(i) for i in range(min(n, 8)))
    else:
        ids = request.param
    dc = DriverClient(ids, bz.DriverConfig.driver_client_cfg(bz.CardType.B200))
    assert dc.device_count() == len(ids.split(","))
    yield dc
    dc.close()


# ------------------------------------------------------------------------------------------ multi-device client
@pytest.mark.parametrize("cname,curve", [("BLS12_381", Curve.BLS381), ("BN254", Curve.BN254), ("BLS12_377", Curve.BLS377)])
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 4099])
def test_group_dma_vs_oracle(group_dc, oracle, cname, curve, n):
    """integration_msm.rs:150-207 call order on a device list: same bytes as the oracle (n < devices leaves members idle)."""
    c = CURVE_BY_NAME[cname]
    pts, _, _ = chain_points(c, n, seed=n + 1)
    sc = random_scalars(c, n, seed=200 + n)
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, False, curve), group_dc)
    try:
        params = MSMParams(n, None)
        m.initialize(params)
        m.start_process()
        m.set_data(MSMInput(pts, sc, params))
        m.wait_result()
        r = m.result()
        assert r.result == oracle.msm_pippenger(cname, pts, sc, n)
        assert r.result_label == 0
    finally:
        m.close()


def test_group_hbm_chunked_load_readback_and_tasks(group_dc, oracle):
    """integration_msm_hbm.rs:121-226 on a device list: bases loaded in chunks (the reference streams chunks), read back
    byte-exact, then scalars-only tasks with two in flight; labels in order."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 3000 + 17
    pts, p0, q = chain_points(c, n, seed=33)
    addr = 0x4000000
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), group_dc)
    try:
        params = MSMParams(n, (addr, 0))
        m.initialize(params)
        rec = c.point_size
        step = 777 * rec
        for off in range(0, n * rec, step):
            m.load_data_to_hbm(pts[off:off + step], addr, off)
        assert m.get_data_from_hbm(n * rec, addr, 0) == bytes(pts)
        assert m.get_data_from_hbm(10 * rec, addr, 1495 * rec) == bytes(pts[1495 * rec:1505 * rec])   # straddles members
        scs = [random_scalars(c, n, seed=60 + i) for i in range(3)]
        for i in range(2):
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, scs[i], params))
        m.wait_result()
        r0 = m.result()
        m.start_process()                          # task queued before its data (integration_msm.rs:186-193)
        m.set_data(MSMInput(None, scs[2], params))
        m.wait_result()
        r1 = m.result()
        m.wait_result()
        r2 = m.result()
        assert [r0.result_label, r1.result_label, r2.result_label] == [0, 1, 2]
        for r, sc in zip((r0, r1, r2), scs):
            assert r.result == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        with pytest.raises(bz.error.NoResult):
            m.result()
    finally:
        m.close()


def test_group_generated_points_merged_table_and_errors(group_dc, oracle):
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << 14
    p0, q = seed_points(c, 15)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), group_dc)
    try:
        m.generate_chain_points(p0 + q, 0, n, 0, 0)
        m.set_precompute(2)                        # every member derives its own window-merged table
        params = MSMParams(n, (0, 0))
        for it in range(2):
            sc = random_scalars(c, n, seed=80 + it)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
            m.wait_result()
            assert m.result().result == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert m.plan_info()["merged_table"]
        # a non-canonical scalar in the LAST member's shard is reported by the combined task
        bad = random_scalars(c, n, seed=90).copy()
        bad[(n - 1) * 32:] = 0xff
        m.start_process()
        m.set_data(MSMInput(None, bad, params))
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.wait_result()
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.result()
        sc = random_scalars(c, n, seed=91)
        m.start_process()
        m.set_data(MSMInput(None, sc, params))
        m.wait_result()
        assert m.result().result == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.set_scalars_device(16, params)
    finally:
        m.close()


def test_group_precompute_factor8(group_dc, oracle):
    c = CURVE_BY_NAME["BN254"]
    n = 203
    pts, _, _ = chain_points(c, n, seed=5)
    sc = random_scalars(c, n, seed=6)
    bases8 = precompute_bases(c, pts, n, 8)
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, True, Curve.BN254), group_dc)
    try:
        params = MSMParams(n, None)
        m.initialize(params)
        m.start_process()
        m.set_data(MSMInput(bases8, sc, params))
        m.wait_result()
        assert m.result().result == oracle.msm_pippenger("BN254", pts, sc, n)
    finally:
        m.close()


def test_ranked_msm_and_ntt_two_processes(oracle):
    """One process per GPU over NCCL (bz_dclient_comm_init): needs >= 2 GPUs (skipped on a 1-GPU box)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(HERE, "dist_msm_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_MSM_OK" in r.stdout and "DIST_NTT_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


# ------------------------------------------------------------------------------------------ review findings
def test_dma_input_replaced_by_hbm_input(dclient, oracle):
    """DMA set_data with no task pending, then an HBM-mode set_data, then start_process must run on the HBM
    point set (the deferred table build of the replaced DMA input must not overwrite it)."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 500
    pts_a, _, _ = chain_points(c, n, seed=1)
    pts_b, p0, q = chain_points(c, n, seed=2)
    sc_a = random_scalars(c, n, seed=3)
    sc_b = random_scalars(c, n, seed=4)
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, False, Curve.BLS381), dclient)
    try:
        m.set_data(MSMInput(pts_a, sc_a, MSMParams(n, None)))            # streamed points, nobody asked for a task yet
        m.load_data_to_hbm(pts_b, 0x8000000, 0)
        params = MSMParams(n, (0x8000000, 0))
        m.initialize(params)
        m.set_data(MSMInput(None, sc_b, params))                         # replaces the pending input
        m.start_process()
        m.wait_result()
        assert m.result().result == oracle.chain_expected("BLS12_381", p0, q, sc_b, n)
    finally:
        m.close()


def test_result_queue_overflow_keeps_labels_and_slots(dclient, oracle):
    """A 17th task is refused without burning a label, a result slot or an event; the 16 queued results stay intact."""
    c = CURVE_BY_NAME["BN254"]
    n = 64
    pts, p0, q = chain_points(c, n, seed=7)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BN254), dclient)
    try:
        m.load_data_to_hbm(pts, 0, 0)
        params = MSMParams(n, (0, 0))
        m.initialize(params)
        scs = [random_scalars(c, n, seed=300 + i) for i in range(18)]
        for i in range(16):
            m.start_process()
            m.set_data(MSMInput(None, scs[i], params))
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            m.start_process()
            m.set_data(MSMInput(None, scs[16], params))
        assert m.task_label() == 15
        for i in range(16):
            m.wait_result()
            r = m.result()
            assert r.result_label == i
            assert r.result == oracle.chain_expected("BN254", p0, q, scs[i], n), i
        m.start_process()                                   # the refused task's data is still there
        m.wait_result()
        r = m.result()
        assert r.result_label == 16 and r.result == oracle.chain_expected("BN254", p0, q, scs[16], n)
    finally:
        m.close()


def test_arena_sparse_addresses_and_range_invalidation(dclient, oracle):
    """Card address space: far-apart addresses do not cost the memory in between, unwritten HBM reads as zeros, and a
    write that does not overlap a client's bases leaves its derived tables alone (no rebuild)."""
    c = CURVE_BY_NAME["BLS12_381"]
    hi = 0x20_0000_0000          # 128 GiB: beyond anything a grow-by-copy arena could reach next to other tests
    blob = bytes(range(256)) * 16
    dclient.dma_write(hi, 64, blob)
    assert dclient.dma_read(hi, 64, len(blob)) == blob
    assert dclient.dma_read(hi, 0, 64) == bytes(64)
    assert dclient.dma_read(hi - 4096, 0, 4096) == bytes(4096)            # neighbouring chunk never mapped
    assert dclient.dma_read(0x30_0000_0000, 0, 1000) == bytes(1000)
    with pytest.raises(bz.error.ReadError):
        dclient.dma_read((1 << 64) - 8, 16, 64)                            # address overflow
    with pytest.raises(bz.error.WriteError):
        dclient.dma_write(1 << 40, 0, b"x")                                # the stream ports are not HBM
    n = 1 << 12
    p0, q = seed_points(c, 3)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        base = 0x1_0000_0000
        m.generate_chain_points(p0 + q, 0, n, base, 0)
        m.set_precompute(2)
        params = MSMParams(n, (base, 0))

        def run(seed):
            sc = random_scalars(c, n, seed=seed)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
            m.wait_result()
            assert m.result().result == oracle.chain_expected("BLS12_381", p0, q, sc, n)

        run(1)
        t0 = m.table_build_ms()
        assert t0 > 0 and m.plan_info()["merged_table"]
        dclient.dma_write(base + n * 96, 0, b"\x01" * 96)                  # right behind the bases: no overlap
        dclient.dma_write(hi, 0, blob)
        run(2)
        assert m.table_build_ms() == t0                                    # not rebuilt
        pts, _, _ = chain_points(c, n, seed=3)
        patch = bytes(pts[5 * 96:6 * 96])
        dclient.dma_write(base, 7 * 96, patch)                             # overlaps: point 7 := point 5
        sc = random_scalars(c, n, seed=9)
        m.initialize(params)
        m.start_process()
        m.set_data(MSMInput(None, sc, params))
        m.wait_result()
        pts2 = pts.copy()
        pts2[7 * 96:8 * 96] = pts[5 * 96:6 * 96]
        assert m.result().result == oracle.msm_pippenger("BLS12_381", pts2, sc, n)
        assert m.table_build_ms() != t0                                    # rebuilt from the new bytes
    finally:
        m.close()


def test_get_api_register_file(dclient, oracle):
    """msm_api.rs:324-330 get_api(): every INGO_MSM_ADDR register (msm_hw_code.rs:6-55) has a value."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 2048
    pts, p0, q = chain_points(c, n, seed=12)
    sc = random_scalars(c, n, seed=13)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.load_data_to_hbm(pts, 0x100000, 0)
        params = MSMParams(n, (0x100000, 0))
        m.initialize(params)
        regs = m.get_api()
        assert set(regs) == set(bz.ingo_msm.INGO_MSM_ADDR)
        assert regs["ADDR_CPU2HIF_C_NUMBER_OF_MSM_ELEMENTS"] == n and regs["ADDR_CPU2HIF_C_BASES_SOURCE"] == 1
        assert regs["ADDR_CPU2HIF_C_BASES_HBM_START_ADDRESS_LO"] == 0x100000
        assert regs["ADDR_HIF2CPU_C_RESULT_VALID"] == 0 and regs["ADDR_HIF2CPU_C_MSM_ENGINE_READY"] == 1
        m.start_process()
        assert m.get_api()["ADDR_HIF2CPU_C_NOF_PENDING_TASKS_IN_QUEUE"] == 1
        m.set_data(MSMInput(None, sc, params))
        m.wait_result()
        regs = m.get_api()
        assert regs["ADDR_HIF2CPU_C_RESULT_VALID"] == 1 and regs["ADDR_HIF2CPU_C_NOF_PENDING_RESULTS_IN_QUEUE"] == 1
        assert regs["ADDR_HIF2CPU_C_RESULT"] == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert regs["ADDR_HIF2CPU_C_RESULT_LABEL"] == 0
        tot = regs["ADDR_HIF2CPU_C_LAST_TASK_PHASE1_TOTAL_CLOCKS_LO"] | regs["ADDR_HIF2CPU_C_LAST_TASK_PHASE1_TOTAL_CLOCKS_HI"] << 32
        busy = regs["ADDR_HIF2CPU_C_LAST_TASK_PHASE1_BUSY_ECADDER_CLOCKS_LO"]
        assert 0 < busy <= tot
        assert regs["ADDR_HIF2CPU_E_BUCKET_ACCUMULATION_PHASE_COMPLETED"] == 1
        m.log_api_values()
        m.result()
        assert m.get_api()["ADDR_HIF2CPU_C_RESULT_VALID"] == 0
        p = bz.MSMImageParametrs.parse_image_params(regs["ADDR_HIF2CPU_C_IMAGE_PARAMTERS"])
        assert p.hif2_cpu_c_curve == int(Curve.BLS381) and p.hif2cpu_c_is_stub == 0
    finally:
        m.close()


# ------------------------------------------------------------------------------------------ published values, CUDA path
KATS = json.load(open(os.path.join(HERE, "golden", "external_kats.json")))


def one_point_msm(dclient, curve, point, k):
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, False, curve), dclient)
    try:
        params = MSMParams(1, None)
        m.initialize(params)
        m.start_process()
        m.set_data(MSMInput(point, int(k).to_bytes(32, "little"), params))
        m.wait_result()
        return m.result().result
    finally:
        m.close()


def test_external_kats_through_cuda(dclient):
    from oracle.py import curves, ec
    c = curves.BLS12_381
    G = ec.encode_point(c, (c.gx, c.gy))
    for v in KATS["bls12_381_g1_pubkeys"]["vectors"]:
        x, y = ec.decode_result(c, one_point_msm(dclient, Curve.BLS381, G, v["k"]))
        b = bytearray(x.to_bytes(48, "big"))
        b[0] |= 0x80 | (0x20 if y > (c.q - 1) // 2 else 0)
        assert b.hex() == v["compressed"]
    b254 = curves.BN254
    k = KATS["bn254_g1_double"]
    got = ec.decode_result(b254, one_point_msm(dclient, Curve.BN254, ec.encode_point(b254, (1, 2)), 2))
    assert got == (int(k["x"], 16), int(k["y"], 16))
    c377 = curves.BLS12_377
    g = KATS["bls12_377_g1_generator"]
    G377 = ec.encode_point(c377, (int(g["x"]), int(g["y"])))
    rec = one_point_msm(dclient, Curve.BLS377, G377, c377.r - 1)          # (r-1) G = -G
    assert ec.decode_result(c377, rec) == (int(g["x"]), c377.q - int(g["y"]))


def test_ntt_root_is_the_published_one(dclient):
    """NTT(delta_1)[k] = w^k with w derived from the PUBLISHED arkworks two-adic roots (external_kats.json)."""
    from oracle.py import curves, ntt as pyntt
    from blaze_b200 import NTTClient, NTTInput
    for name, k in KATS["two_adic_roots"].items():
        if name == "source":
            continue
        c = curves.CURVES[name]
        root = int(k["root_hex"], 16) if "root_hex" in k else int(k["root_dec"])
        log_n = 12
        n = 1 << log_n
        w = pow(root, 1 << (k["two_adicity"] - log_n), c.r)
        d = np.frombuffer(pyntt.encode([0, 1] + [0] * (n - 2)), dtype=np.uint8).copy()
        t = NTTClient.new_ex(dclient, field=c.code, log_size=log_n)
        try:
            t.initialize()
            t.set_data(NTTInput(0, d))
            t.start_process(0)
            t.wait_result()
            out = pyntt.decode(bytes(t.result(0)))
        finally:
            t.close()
        exp, acc = [], 1
        for _ in range(n):
            exp.append(acc)
            acc = acc * w % c.r
        assert out == exp, name


def test_ntt_rejects_non_canonical_input(dclient):
    from blaze_b200 import NTTClient, NTTInput
    log_n = 10
    d = np.zeros((1 << log_n) * 32, dtype=np.uint8)
    d[5 * 32:6 * 32] = 0xff
    t = NTTClient.new_ex(dclient, field=2, log_size=log_n)
    try:
        t.initialize()
        t.set_data(NTTInput(0, d))
        t.start_process(0)
        with pytest.raises(bz.error.InvalidPrimitiveParam):
            t.wait_result()
        d[5 * 32:6 * 32] = 0
        t.set_data(NTTInput(0, d))
        t.start_process(0)
        t.wait_result()
        assert bytes(t.result(0)) == bytes(len(d))
    finally:
        t.close()


def test_ntt_2p27_reference_size_against_definition(dclient, oracle):
    """BASELINE.json configs[3] size through the reference's constructor (NTTClient::new: fixed 2^27, ntt_data.rs:65-66),
    call order of integration_ntt.rs:6-60.  The reference compares with an external golden file; here outputs are
    checked against their DEFINITION out[k] = sum_j in[j] w^(jk) (oracle Horner, O(n) per point) at 12 positions
    that hit every digit of the 9+9+9 pass plan, plus linearity in one input element."""
    from blaze_b200 import NTT, NTTClient, NTTInput
    log_n = 27
    n = 1 << log_n
    c = CURVE_BY_NAME["BLS12_381"]
    d = random_scalars(c, n, seed=2027)
    t = NTTClient.new(NTT.Ntt, dclient)
    try:
        t.initialize()
        t.set_data(NTTInput(0, d))
        t.start_process(0)
        t.wait_result()
        out = np.frombuffer(t.result(0), dtype=np.uint8)
        ks = [0, 1, n - 1, n // 2, 511, 512, (1 << 18) - 1, 1 << 18, 0x2AAAAAA, 0x5555555 % n, 123456789 % n, (1 << 26) + (1 << 9) + 1]
        exp = oracle.ntt_eval("BLS12_381", d, log_n, ks)
        got = [int.from_bytes(bytes(out[32 * k:32 * k + 32]), "little") for k in ks]
        assert got == exp
        # linearity: in[j0] += 1  =>  out[k] += w^(j0 k)
        from oracle.py import curves
        w = curves.root_of_unity(curves.BLS12_381, log_n)
        j0 = 98765432
        d2 = d.copy()
        v = (int.from_bytes(bytes(d2[32 * j0:32 * j0 + 32]), "little") + 1) % c.r
        d2[32 * j0:32 * j0 + 32] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
        t.set_data(NTTInput(1, d2))
        t.start_process(1)
        t.wait_result()
        out2 = np.frombuffer(t.result(1), dtype=np.uint8)
        for k, e in zip(ks, exp):
            g2 = int.from_bytes(bytes(out2[32 * k:32 * k + 32]), "little")
            assert g2 == (e + pow(w, j0 * k, c.r)) % c.r
    finally:
        t.close()


def test_precompute_x8_hbm_mode_at_scale(dclient, oracle):
    """The reference's precomputed wire format in HBM mode (integration_msm_hbm.rs:13-119: MSMInit{mem_type: DMA,
    is_precompute: true} + Some((addr, off))) at 2^20 bases = 768 MiB of x8 records: records derived on the device
    (sampled against the oracle's 2^(32 i) P), read back through get_data_from_hbm, MSM result vs the closed form."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << 20
    p0, q = seed_points(c, 44)
    base_addr, x8_addr = 0x8_0000_0000, 0x9_0000_0000
    rec = c.point_size * 8
    gen = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        gen.generate_chain_points(p0 + q, 0, n, base_addr, 0)
        gen.expand_precompute(base_addr, n, x8_addr)
        for k in (0, 12345, n - 1):
            got = gen.get_data_from_hbm(rec, x8_addr, k * rec)
            assert got[:c.point_size] == gen.get_data_from_hbm(c.point_size, base_addr, k * c.point_size)
            for i in range(8):
                assert got[i * c.point_size:(i + 1) * c.point_size] == oracle.point_mul("BLS12_381", got[:c.point_size], pow(2, 32 * i, c.r))
    finally:
        gen.close()
    m = MSMClient.new(MSMInit(PointMemoryType.DMA, True, Curve.BLS381), dclient)     # the reference's own combination
    try:
        params = MSMParams(n, (x8_addr, 0))
        for seed in (1, 2):
            sc = random_scalars(c, n, seed=seed)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
            m.wait_result()
            assert m.result().result == oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert m.plan_info()["windows"] >= 1
    finally:
        m.close()


# ------------------------------------------------------------------------------------------ batched-affine sweep
@pytest.mark.parametrize("cname,curve", [("BLS12_381", Curve.BLS381), ("BN254", Curve.BN254), ("BLS12_377", Curve.BLS377)])
@pytest.mark.parametrize("rounds", [1, 3, 5])
def test_batched_affine_sweep_vs_oracle(dclient, oracle, cname, curve, rounds):
    """The fused batched-affine accumulation (msm_ba2.cuh) forced on, against the oracle: uniform scalars at two window
    sizes (long and short bucket runs), the reference's tiled distribution (every first-round pair is a doubling),
    and skewed scalars (one giant bucket per window; mostly empty buckets)."""
    from util import tile
    c = CURVE_BY_NAME[cname]
    n = 6000
    pts, _, _ = chain_points(c, n, seed=17)
    sc = random_scalars(c, n, seed=18)
    exp = oracle.msm_pippenger(cname, pts, sc, n)
    for cb in (5, 9):
        m = MSMClient.new(MSMInit(PointMemoryType.DMA, False, curve), dclient)
        try:
            m.set_window_bits(cb)
            m.set_accumulate_mode(2, rounds)
            params = MSMParams(n, None)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(pts, sc, params))
            m.wait_result()
            assert m.result().result == exp
            assert m.plan_info()["accumulate"] == "batched-affine" and m.plan_info()["ba_rounds"] == rounds
        finally:
            m.close()
    pts256, _, _ = chain_points(c, 256, seed=11)
    sc256 = random_scalars(c, 256, seed=12)
    nt = 256 * 41 + 7
    tp, ts = tile(pts256, c.point_size, 256, nt), tile(sc256, 32, 256, nt)
    same = np.frombuffer((0x1d3c5b7a99f0e1d2c3b4a5968778695a4b3c2d1e0f).to_bytes(32, "little") * n, dtype=np.uint8).copy()
    few = np.zeros(n * 32, dtype=np.uint8)
    few[32 * 100:32 * 100 + 31] = 0xAB
    few[32 * 4000:32 * 4000 + 8] = 0x77
    for p_, s_, k in ((tp, ts, nt), (pts, same, n), (pts, few, n)):
        m = MSMClient.new(MSMInit(PointMemoryType.DMA, False, curve), dclient)
        try:
            m.set_window_bits(7)
            m.set_accumulate_mode(2, rounds)
            params = MSMParams(k, None)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(p_, s_, params))
            m.wait_result()
            assert m.result().result == oracle.msm_pippenger(cname, p_, s_, k)
        finally:
            m.close()


def test_batched_affine_merged_table_2p20_matches_xyzz(dclient, oracle):
    """HBM-resident points with the window-merged table at 2^20: automatic mode picks the batched-affine sweep; result
    equals the closed form and the XYZZ sweep's bytes."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << 20
    p0, q = seed_points(c, 21)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.generate_chain_points(p0 + q, 0, n, 0x40_0000_0000, 0)
        m.set_precompute(2)
        params = MSMParams(n, (0x40_0000_0000, 0))
        sc = random_scalars(c, n, seed=22)
        res = {}
        for mode in (-1, 0, 2):
            m.set_accumulate_mode(mode, -1)
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
            m.wait_result()
            res[mode] = (m.result().result, m.plan_info())
        exp = oracle.chain_expected("BLS12_381", p0, q, sc, n)
        assert res[-1][0] == exp and res[0][0] == exp and res[2][0] == exp
        assert res[0][1]["accumulate"] == "xyzz" and res[2][1]["accumulate"] == "batched-affine"
    finally:
        m.close()


def test_producer_consumer_threads(dclient, oracle):
    """One thread queues tasks (start_process + set_data), another drains them (wait_result + result), like the
    reference's two-thread Poseidon test (integration_poseidon.rs:60-121): wait_result must not hold the client while it blocks."""
    import threading
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << 15
    p0, q = seed_points(c, 8)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.generate_chain_points(p0 + q, 0, n, 0x50_0000_0000, 0)
        params = MSMParams(n, (0x50_0000_0000, 0))
        m.initialize(params)
        K = 10
        scs = [random_scalars(c, n, seed=700 + i) for i in range(K)]
        exp = [oracle.chain_expected("BLS12_381", p0, q, s, n) for s in scs]
        got, errors = [], []
        queued = threading.Semaphore(0)

        def producer():
            try:
                for s in scs:
                    m.start_process()
                    m.set_data(MSMInput(None, s, params))
                    queued.release()
            except Exception as e:      # pragma: no cover
                errors.append(e)
                queued.release()

        def consumer():
            try:
                for _ in range(K):
                    queued.acquire()
                    m.wait_result()
                    got.append(m.result())
            except Exception as e:      # pragma: no cover
                errors.append(e)

        tp, tc = threading.Thread(target=producer), threading.Thread(target=consumer)
        tp.start(); tc.start(); tp.join(120); tc.join(120)
        assert not errors, errors
        assert [r.result_label for r in got] == list(range(K))
        assert [r.result for r in got] == exp
    finally:
        m.close()


def test_tail_stream_many_tasks_in_flight(dclient, oracle):
    """Eight tasks queued back to back on one resident point set (window-merged table from the second on): the end of
    every task -- upper reduction levels, window combine, result copy -- runs on the client's tail stream while the
    next task's sort / accumulation already occupies the work stream.  Every task must return ITS sum, and the
    non-canonical scalar of task 3 must be reported by task 3 only (the error word is read back before the next
    task's windowing can clear it)."""
    c = CURVE_BY_NAME["BLS12_381"]
    n = 1 << 18
    p0, q = seed_points(c, 81)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS381), dclient)
    try:
        m.generate_chain_points(p0 + q, 0, n, 0x60_0000_0000, 0)
        params = MSMParams(n, (0x60_0000_0000, 0))
        m.initialize(params)
        K = 8
        scs = [np.array(random_scalars(c, n, seed=900 + i), dtype=np.uint8, copy=True) for i in range(K)]
        scs[3][32 * 77:32 * 78] = 0xff                      # >= r
        exp = [oracle.chain_expected("BLS12_381", p0, q, s, n) if i != 3 else None for i, s in enumerate(scs)]
        for s in scs:
            m.start_process()
            m.set_data(MSMInput(None, s, params))
        for i in range(K):
            if i == 3:
                with pytest.raises(bz.error.InvalidPrimitiveParam):
                    m.wait_result()
                with pytest.raises(bz.error.InvalidPrimitiveParam):
                    m.result()
                continue
            m.wait_result()
            r = m.result()
            assert r.result_label == i
            assert r.result == exp[i], "task %d" % i
        assert m.plan_info()["merged_table"]
    finally:
        m.close()


def test_config5_size_bls12_377_2p26_one_gpu(dclient, oracle):
    """The reference's 2^26 "max" size on BLS12-377 (integration_msm.rs:386-468 msm_bls12_377_max_test; BASELINE.json
    configs[4] runs it across 8 GPUs, bench.py `config5`): HBM-resident points, two tasks (the second on the window-merged
    table), bit-exact against the closed form of the chain workload."""
    c = CURVE_BY_NAME["BLS12_377"]
    n = 1 << 26
    p0, q = seed_points(c, 377)
    m = MSMClient.new(MSMInit(PointMemoryType.HBM, False, Curve.BLS377), dclient)
    try:
        base = 0x60_0000_0000
        m.generate_chain_points(p0 + q, 0, n, base, 0)
        params = MSMParams(n, (base, 0))
        sc = random_scalars(c, n, seed=378)
        exp = oracle.chain_expected("BLS12_377", p0, q, sc, n)
        for it in range(2):
            m.initialize(params)
            m.start_process()
            m.set_data(MSMInput(None, sc, params))
            m.wait_result()
            r = m.result()
            assert r.result == exp and r.result_label == it
        assert m.plan_info()["merged_table"] and m.plan_info()["windows"] == 11
    finally:
        m.close()
