"""NTT parity on the GPU through the C ABI, against the oracle (arkworks radix-2 FFT semantics).

Call order follows /root/reference/tests/integration_ntt.rs:6-60 (set_data -> initialize ->
start_process(buf) -> wait_result -> result(buf)) and :103-136 (double-buffer pipeline).  The
reference compares with external golden files that are not in its repository; parity here is with
the oracle's definition (oracle/py/ntt.py), i.e. UNPINNED against the reference's own files."""
import numpy as np
import pytest

from blaze_b200 import NTT, NTTClient, NTTInput, NttInit

pytestmark = pytest.mark.gpu

FIELDS = {"BLS12_381": 2, "BLS12_377": 0, "BN254": 1}
R = {"BLS12_381": 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001,
     "BLS12_377": 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001,
     "BN254": 21888242871839275222246405745257275088548364400416034343698204186575808495617}


def rand_elems(name, n, seed):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    keep = R[name].bit_length() - 1 - 8 * 31
    raw[:, 31] &= (1 << keep) - 1
    return raw.reshape(-1)


def run(dclient, name, log_n, data, inverse=False, buf=0):
    t = NTTClient.new_ex(dclient, FIELDS[name], log_n, inverse)
    try:
        t.set_data(NTTInput(buf, data))
        t.initialize(NttInit())
        t.start_process(buf)
        t.wait_result()
        return bytes(t.result(buf))
    finally:
        t.close()


@pytest.mark.parametrize("name", ["BLS12_381", "BLS12_377", "BN254"])
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8, 9, 10, 13, 16, 19])
def test_ntt_vs_oracle(dclient, oracle, name, log_n):
    d = rand_elems(name, 1 << log_n, seed=log_n)
    exp = d.copy()
    oracle.ntt(name, exp, log_n)
    assert run(dclient, name, log_n, d) == bytes(exp)
    # inverse (with 1/n scaling) brings it back
    assert run(dclient, name, log_n, exp, inverse=True) == bytes(d)


def test_ntt_edge_values(dclient, oracle):
    """all zero, all r-1, delta: carries through every limb of the butterflies."""
    name, log_n = "BLS12_381", 11
    n = 1 << log_n
    for kind in range(3):
        vals = [0] * n
        if kind == 1:
            vals = [R[name] - 1] * n
        if kind == 2:
            vals[1] = 1
        d = np.frombuffer(b"".join(v.to_bytes(32, "little") for v in vals), dtype=np.uint8).copy()
        exp = d.copy()
        oracle.ntt(name, exp, log_n)
        assert run(dclient, name, log_n, d) == bytes(exp)


def test_ntt_double_buffer_pipeline(dclient, oracle):
    """integration_ntt.rs:103-136: compute slot 1-h while slot h is read back / refilled."""
    name, log_n = "BLS12_381", 14
    t = NTTClient.new_ex(dclient, FIELDS[name], log_n)
    try:
        ins = [rand_elems(name, 1 << log_n, seed=100 + i) for i in range(4)]
        exps = []
        for d in ins:
            e = d.copy()
            oracle.ntt(name, e, log_n)
            exps.append(bytes(e))
        t.initialize(NttInit())
        h = 0
        t.set_data(NTTInput(h, ins[0]))
        outs = []
        for i in range(4):
            t.start_process(h)                 # kernel works on slot h
            if i + 1 < 4:
                t.set_data(NTTInput(1 - h, ins[i + 1]))   # host fills the other slot meanwhile
            t.wait_result()
            outs.append(bytes(t.result(h)))
            h = 1 - h
        assert outs == exps
    finally:
        t.close()


def test_ntt_pipeline_reference_order_overlapped_copies(dclient, oracle):
    """The reference's own pipelined cycle (integration_ntt.rs:103-136) at a size where the copies take as long
    as the transform: start_process(1-h); result(h); set_data(h); wait_result().  The copies run on their own
    streams here, so every ordering hazard (read-out vs next transform, refill vs read-out) is exercised."""
    name, log_n = "BLS12_381", 20
    n = 1 << log_n
    t = NTTClient.new_ex(dclient, FIELDS[name], log_n)
    try:
        rounds = 6
        ins = [rand_elems(name, n, seed=500 + i) for i in range(rounds)]
        exps = []
        for d in ins:
            e = d.copy()
            oracle.ntt(name, e, log_n)
            exps.append(bytes(e))
        t.initialize(NttInit())
        outs = []
        h = 0
        t.set_data(NTTInput(0, ins[0]))
        t.start_process(0)
        t.set_data(NTTInput(1, ins[1]))
        t.wait_result()
        for i in range(1, rounds):
            t.start_process(1 - h)                       # transform input i on the other slot
            outs.append(bytes(t.result(h)))              # read output i-1 meanwhile
            if i + 1 < rounds:
                t.set_data(NTTInput(h, ins[i + 1]))      # refill the slot just read
            t.wait_result()
            h = 1 - h
        outs.append(bytes(t.result(h)))
        assert outs == exps
    finally:
        t.close()


def test_ntt_feeder_and_drainer_threads(dclient, oracle):
    """Three host threads on one NTTClient (the client is Send + Sync like the reference's): a feeder calls set_data, the
    main thread start_process / wait_result, a drainer result.  The blocking calls do not hold the client lock, so the
    copy into one slot and the read-out of the other overlap; a slot is refilled only after its previous result was read."""
    import threading
    name, log_n = "BLS12_381", 20
    n = 1 << log_n
    t = NTTClient.new_ex(dclient, FIELDS[name], log_n)
    try:
        rounds = 8
        ins = [rand_elems(name, n, seed=900 + i) for i in range(rounds)]
        exps = []
        for d in ins:
            e = d.copy()
            oracle.ntt(name, e, log_n)
            exps.append(bytes(e))
        t.initialize(NttInit())
        in_ready = [threading.Semaphore(0) for _ in range(rounds)]
        cmp_done = [threading.Semaphore(0) for _ in range(rounds)]
        out_done = [threading.Semaphore(0) for _ in range(rounds)]
        outs, errs = [None] * rounds, []

        def feeder():
            try:
                for i in range(rounds):
                    if i >= 2:
                        assert out_done[i - 2].acquire(timeout=60)
                    t.set_data(NTTInput(i % 2, ins[i]))
                    in_ready[i].release()
            except Exception as ex:      # pragma: no cover
                errs.append(ex)

        def drainer():
            try:
                for i in range(rounds):
                    assert cmp_done[i].acquire(timeout=60)
                    outs[i] = bytes(t.result(i % 2))
                    out_done[i].release()
            except Exception as ex:      # pragma: no cover
                errs.append(ex)
        tf, td = threading.Thread(target=feeder), threading.Thread(target=drainer)
        tf.start(); td.start()
        for i in range(rounds):
            assert in_ready[i].acquire(timeout=60)
            t.start_process(i % 2)
            t.wait_result()
            cmp_done[i].release()
        tf.join(120); td.join(120)
        assert not errs, errs
        assert outs == exps
    finally:
        t.close()


def test_ntt_reference_constructor_is_2p27(dclient):
    t = NTTClient.new(NTT.Ntt, dclient)
    try:
        assert t.log_size == 27 and t.loaded_binary_parameters()[1] == 27
    finally:
        t.close()


def test_ntt_large_linearity(dclient, oracle):
    """2^22: too slow to pin fully in Python, so: oracle C++ (multi-thread) equality, and the
    size-independent properties NTT(delta_1)[k] = w^k spot checks + inverse round trip."""
    name, log_n = "BLS12_381", 22
    d = rand_elems(name, 1 << log_n, seed=7)
    out = run(dclient, name, log_n, d)
    exp = d.copy()
    oracle.ntt(name, exp, log_n)
    assert out == bytes(exp)
    assert run(dclient, name, log_n, np.frombuffer(out, dtype=np.uint8), inverse=True) == bytes(d)


def test_distributed_ntt_single_rank(dclient, oracle):
    """The multi-GPU four-step code path with world = 1 (the exchange buffer's only peer is the rank
    itself): column passes, fused twiddle + exchange store, row passes, strided host I/O."""
    from blaze_b200 import DistributedNTT
    for log_n in (2, 7, 12, 15, 20):
        d = rand_elems("BLS12_381", 1 << log_n, seed=300 + log_n)
        exp = d.copy()
        oracle.ntt("BLS12_381", exp, log_n)
        t = DistributedNTT(dclient, log_n)
        try:
            t.set_input(d)
            t.run()
            out = np.zeros_like(d)
            t.get_output(out)
            assert bytes(out) == bytes(exp), log_n
        finally:
            t.close()


def test_distributed_ntt_two_ranks_if_available(oracle):
    """Real peer stores over NVLink: needs >= 2 GPUs (skipped on a 1-GPU box)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    for log_n in ("16", "21"):      # 8 + 8 (one pass each) and 9 + 12 (unbalanced split: 1 + 2 passes)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(here, "dist_ntt_check.py"), log_n],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "DIST_NTT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
