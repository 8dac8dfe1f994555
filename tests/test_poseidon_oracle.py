"""CPU: structural checks of the Poseidon oracle (parameter generation, record format)."""
from oracle.py import poseidon as P


def test_parameter_shapes_and_mds():
    for t in (9, 12):
        rc, mds = P.params(t)
        assert len(rc) == (P.R_F + P.R_P[t]) * t and all(0 <= x < P.R_ for x in rc)
        assert len(set(rc)) == len(rc)
        for i in range(t):
            for j in range(t):
                assert mds[i][j] * (i + t + j) % P.R_ == 1        # Cauchy matrix


def test_permutation_is_a_bijection_sample():
    a = P.permute(list(range(9)))
    b = P.permute([1] + list(range(1, 9)))
    assert a != b and len(set(a)) == 9


def test_record_roundtrip_matches_reference_parser():
    # poseidon_api.rs:50-61: hash_id = LE32(meta[0..4]) & 0x3fffffff, layer_id = LE32(meta[3..5],0,0) >> 6
    for hid, lid in ((0, 0), (511, 0), (7, 3), (0x3fffffff, 9), (12345678, 1)):
        h, i, l = P.parse_record(P.record(42, hid, lid))
        assert (i, l) == (hid, lid) and int.from_bytes(h, "little") == 42


def test_tree_sizes():
    assert sum(P.tree_sizes(4)) == 585                               # integration_poseidon.rs:23
    assert P.tree_sizes(4)[0] == 512


def test_optimized_evaluation_equals_plain_rounds():
    """The form the CUDA kernel evaluates (constants pushed forward, sparse partial-round matrices) is the same
    permutation as round = add constants, S-box, MDS."""
    import random
    rng = random.Random(11)
    for t, mode in ((3, P.MDS_GRAIN), (3, P.MDS_CAUCHY), (9, P.MDS_CAUCHY), (12, P.MDS_CAUCHY)):
        for _ in range(2):
            s = [rng.randrange(P.R_) for _ in range(t)]
            assert P.permute_optimized(s, mode) == P.permute(s, mode)


def test_library_host_constants_match_oracle_derivation():
    """bz_poseidon_optimized_constants (host-only, no device needed): the C++ preprocessing of the product library
    produces exactly the constants the big-integer derivation gives, for the client's widths and the KAT's."""
    import ctypes
    from blaze_b200._lib import lib
    L = lib()
    for t, mode in ((3, 1), (3, 0), (9, 0), (12, 0)):
        n = ctypes.c_size_t()
        assert L.bz_poseidon_optimized_constants(t, mode, None, 0, ctypes.byref(n)) == 0
        buf = ctypes.create_string_buffer(n.value)
        assert L.bz_poseidon_optimized_constants(t, mode, buf, n.value, ctypes.byref(n)) == 0
        vals = [int.from_bytes(buf.raw[i:i + 32], "little") for i in range(0, n.value, 32)]
        o = P.optimized_params(t, mode)
        exp = [x for r in o["rc_full_first"] for x in r] + [x for r in o["mds"] for x in r] + [x for r in o["pre_sparse"] for x in r]
        for c0, (row0, col0) in zip(o["partial_c0"], o["sparse"]):
            exp += [c0] + row0 + col0
        exp += [x for r in o["rc_full_second"] for x in r]
        assert vals == exp, (t, mode)
    assert L.bz_poseidon_optimized_constants(5, 0, None, 0, ctypes.byref(n)) != 0
