"""CPU: the oracle's restatement of libbsc's BWT stage (oracle/bsc_oracle.c) against the
reference's own bsc_bwt_encode = divbwt (oracle/_ref/libref_bsc.so, built from
cuda-bsc/libbsc without CUDA), and the reference's bsc_bwt_decode as the inverse."""
import numpy as np
import pytest

import oracle_lib as O

import os

needs_ref = pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["bsc_zipf", "bsc_binary"])
def test_oracle_matches_golden_from_reference_divbwt(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))       # written by tools/make_golden.py from the reference
    u, p, idx = O.bsc_oracle_bwt_encode(g["data"])
    assert p == int(g["primary"]) and np.array_equal(u, g["U"]) and np.array_equal(idx, g["indexes"])


def _cases():
    rng = np.random.default_rng(5)
    text = np.frombuffer((b"the quick brown fox jumps over the lazy dog. " * 3000), np.uint8)
    return {
        "n2": np.array([7, 7], np.uint8),
        "n3": np.array([2, 1, 2], np.uint8),
        "n9_binary": rng.integers(0, 2, 9, dtype=np.uint8),
        "zeros_1000": np.zeros(1000, np.uint8),
        "periodic": np.tile(np.array([5, 0, 9], np.uint8), 4000),
        "random_70001": rng.integers(0, 256, 70001, dtype=np.uint8),
        "zipf_200k": O.zipf_bytes(200000, 1.3, seed=3),
        "text_135k": text.copy(),
        "quant_1m": O.quant_codes(1 << 20),
    }


@needs_ref
@pytest.mark.parametrize("name", list(_cases().keys()))
def test_oracle_equals_reference_divbwt(name):
    data = _cases()[name]
    ru, rp, ri = O.bsc_ref_bwt_encode(data)
    ou, op, oi = O.bsc_oracle_bwt_encode(data)
    assert op == rp
    assert np.array_equal(ou, ru)
    assert np.array_equal(oi, ri)
    # and the reference's inverse reproduces the input from it
    back = ou.copy()
    idx = np.zeros(256, np.int32)
    idx[: oi.size] = oi
    assert O.ref_bsc().bsc_bwt_decode(back, back.size, op, oi.size, idx, 0) == 0
    assert np.array_equal(back, data)


# ------------------------------------------------------------------------------ Sort Transform ST5-8
def _st_cases():
    rng = np.random.default_rng(3)
    return {
        "text": np.frombuffer((b"the quick brown fox jumps over the lazy dog. " * 3000)[:100000], np.uint8).copy(),
        "rand4": rng.integers(0, 4, 50000, dtype=np.uint8),
        "rand": rng.integers(0, 256, 70001, dtype=np.uint8),
        "zeros": np.zeros(4097, np.uint8),
        "quant": O.quant_codes(1 << 16),
        "tiny2": np.array([5, 5], np.uint8),
        "tiny3": np.array([3, 1, 2], np.uint8),
        "tiny9": rng.integers(0, 3, 9, dtype=np.uint8),
        "period7": (np.arange(7000) % 7).astype(np.uint8),
    }


@pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
@pytest.mark.parametrize("name", list(_st_cases().keys()))
def test_st_oracle_is_pinned_to_the_reference(name):
    """oracle/bsc_oracle.c::bsc_oracle_st_encode == the reference's CPU bsc_st_encode for k = 5, 6
    (st/st.cpp:1005-1027; k = 7, 8 exist on the reference's GPU path only) and is inverted by the
    reference's CPU bsc_st_decode for k = 5..8."""
    data = _st_cases()[name]
    for k in (5, 6, 7, 8):
        out, idx = O.bsc_oracle_st_encode(data, k)
        if k <= 6:
            ref, ri = O.bsc_ref_st_encode(data, k)
            assert ri == idx and np.array_equal(ref, out), k
        rc, back = O.bsc_ref_st_decode(out, k, idx)
        assert rc == 0 and np.array_equal(back, data), k
