"""CPU: pins oracle/cudpp_oracle.c against the reference testrig's gold code
(oracle/_ref/libref_cudpp.so: computeSaGold, computeBwtGold, computeMtfGold,
huffman_build_tree_cpu, computeCompressGold) and the committed golden blocks."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tree(hist):
    arrs = [np.zeros(513, np.int32) for _ in range(4)]
    head = O.oracle().cudpp_oracle_tree(np.ascontiguousarray(hist, np.uint32), *arrs)
    return head, arrs


@pytest.mark.parametrize("name", ["cudpp_zipf.npz", "cudpp_text.npz"])
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLD, name))
    data = g["data"]
    bwt, idx = O.cudpp_oracle_bwt(data)
    assert idx == int(g["bwt_index"]) and np.array_equal(bwt, g["bwt"])
    mtf = O.cudpp_oracle_mtf(bwt)
    assert np.array_equal(mtf, g["mtf"])
    rc, hist, offs, words = O.cudpp_oracle_huffman(mtf)
    assert rc == 0 and np.array_equal(offs, g["offsets"]) and np.array_equal(words, g["words"])
    head, arrs = _tree(hist)
    assert head == int(g["tree_head"])
    assert np.array_equal(arrs[0], g["tree_left"]) and np.array_equal(arrs[1], g["tree_right"])
    assert np.array_equal(arrs[3], g["tree_value"])
    drc, back = O.cudpp_oracle_decompress(data.size, idx, hist, offs, words)
    assert drc == 0 and np.array_equal(back, data)


@pytest.mark.skipif(not O.have_ref("cudpp"), reason="oracle/_ref/libref_cudpp.so not built")
@pytest.mark.parametrize("kind,n", [("rand", 4099), ("zipf", 70000), ("markov", 131072), ("text", 50000)])
def test_oracle_matches_reference_golds(kind, n):
    ref = O.ref_cudpp()
    data = O.cudpp_block(n, kind, seed=n % 11)
    bwt, idx = O.cudpp_oracle_bwt(data)
    rb = np.zeros(n, np.uint8)
    ridx = C.c_int(-1)
    ref.ref_cudpp_bwt(data, rb, C.byref(ridx), n)
    assert ridx.value == idx and np.array_equal(rb, bwt)
    mtf = O.cudpp_oracle_mtf(bwt)
    rm = np.zeros(n, np.uint8)
    ref.ref_cudpp_mtf(bwt, rm, n)
    assert np.array_equal(rm, mtf)
    hist = np.bincount(mtf, minlength=256).astype(np.uint32)
    head, arrs = _tree(hist)
    rarrs = [np.zeros(513, np.int32) for _ in range(4)]
    rhead = C.c_int(-1)
    ref.ref_cudpp_tree(hist, rarrs[0], rarrs[1], rarrs[2], rarrs[3], C.byref(rhead))
    assert rhead.value == head
    for a, b in zip(arrs, rarrs):
        assert np.array_equal(a, b)


@pytest.mark.skipif(not O.have_ref("cudpp"), reason="oracle/_ref/libref_cudpp.so not built")
def test_reference_decoder_accepts_oracle_stream():
    """The reference pins the compressed words only through its decoder
    (test_compress.cpp:744-797): computeCompressGold must rebuild the input, n = 1,048,576."""
    n = 1 << 20
    data = O.cudpp_block(n, "zipf", seed=4)
    rc, idx, hist, offs, words = O.cudpp_oracle_compress(data)
    assert rc == 0
    out = np.zeros(n, np.uint8)
    h257 = np.zeros(257, np.uint32)
    h257[:256] = hist
    O.ref_cudpp().ref_cudpp_decompress(out, idx, h257, offs.copy(), words.size, words.copy(), n)
    assert np.array_equal(out, data)


def test_huffman_stream_rules():
    # one symbol only: alphabet {sym, EOF} -> 1-bit codes, 4096 symbols = 128 words per block
    mtf = np.zeros(8192, np.uint8)
    rc, hist, offs, words = O.cudpp_oracle_huffman(mtf)
    assert rc == 0 and hist[0] == 8192 and list(offs) == [0, 129] and words[0] == 128 and words.size == 258
    # capacity: 257 equiprobable-ish symbols need > 8 bits but stay below 12 bits/symbol
    rng = np.random.default_rng(0)
    rc, _, _, w = O.cudpp_oracle_huffman(rng.integers(0, 256, 1 << 16, dtype=np.uint8))
    assert rc == 0 and w.size <= 16 * 1537
