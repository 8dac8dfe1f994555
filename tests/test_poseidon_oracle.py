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
