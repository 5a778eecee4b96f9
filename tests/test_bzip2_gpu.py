"""GPU parity for the cuda-bzip2 row: libb200lc.so's gpuBlockSort against
the CPU oracle and against the reference's own gpuBWTSort.cu (oracle/_ref/libref_bzip2.so, run
on this GPU), and the drop-in link: the reference's libbz2 produces byte-identical .bz2 streams
with its own GPU sort and with libb200lc.so's (oracle/_ref/libref_bzip2_b200.so)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
from pkg import b200lc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def norun_bytes(n, seed, alphabet=256):
    """Bytes without equal neighbours: bzip2's RLE1 stage (bzlib.c:336-367) leaves them unchanged."""
    rng = np.random.default_rng(seed)
    steps = rng.integers(1, alphabet, n)
    return (np.cumsum(steps) % alphabet).astype(np.uint8)


def texty(n, seed):
    rng = np.random.default_rng(seed)
    words = [b"block", b"sort", b"rotate", b"suffix", b"burrows", b"wheeler", b"gpu", b"merge"]
    idx = rng.integers(0, len(words), n // 4 + 8)
    return np.frombuffer(b" ".join(words[i] for i in idx)[:n], np.uint8).copy()


def _mine(block):
    L = b200lc.lib()
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    L.b200lc_bzip2_block_sort.restype = C.c_int
    L.b200lc_bzip2_block_sort.argtypes = [u8p, u32p, u32p, u32p, C.c_int, C.POINTER(C.c_int)]
    n = block.size
    first = np.zeros(n, np.uint32)
    second = np.zeros(n, np.uint32)
    rank = np.zeros(n, np.uint32)
    depth = C.c_int(-1)
    f = L.b200lc_bzip2_block_sort(np.ascontiguousarray(block), first, second, rank, n, C.byref(depth))
    assert f > 0
    return f, first[:f], second[: n - f], rank, depth.value


@pytest.mark.parametrize("n", [30000, 30001, 30002, 299981])
@pytest.mark.parametrize("kind", ["norun", "text", "small_alphabet"])
def test_block_sort_arrays(n, kind):
    block = {"norun": lambda: norun_bytes(n, n), "text": lambda: texty(n, n),
             "small_alphabet": lambda: norun_bytes(n, n + 1, alphabet=5)}[kind]()
    f, first, second, rank, depth = _mine(block)
    wf, wfirst, wsecond, wrank = O.bzip2_oracle_block_sort(block)
    assert f == wf
    assert np.array_equal(first, wfirst) and np.array_equal(second, wsecond) and np.array_equal(rank, wrank)
    if O.have_ref("bzip2"):
        ref = O.ref_bzip2()
        rfirst = np.zeros(n, np.uint32)
        rsecond = np.zeros(n, np.uint32)
        rrank = np.zeros(n, np.uint32)
        rdepth = C.c_int(-1)
        rf = ref.ref_bzip2_gpuBlockSort(np.ascontiguousarray(block), rfirst, rsecond, rrank, n, C.byref(rdepth))
        assert rf == f
        assert np.array_equal(rfirst[:f], first) and np.array_equal(rsecond[: n - f], second)
        assert np.array_equal(rrank, rank)
        assert rdepth.value == depth


def test_rotation_order_entry_point():
    L = b200lc.lib()
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
    L.b200lc_bzip2_rotation_order.restype = C.c_int
    L.b200lc_bzip2_rotation_order.argtypes = [u8p, C.c_int, u32p, C.POINTER(C.c_int)]
    for n in (1, 2, 3, 1000, 123457):
        block = texty(n, n) if n > 10 else norun_bytes(n, n)
        ptr = np.zeros(n, np.uint32)
        orig = C.c_int(-1)
        assert L.b200lc_bzip2_rotation_order(block, n, ptr, C.byref(orig)) == 0
        want = np.zeros(n, np.uint32)
        O.oracle().bzip2_oracle_rotation_order(block, n, want)
        assert np.array_equal(ptr, want) and ptr[orig.value] == 0


@pytest.mark.skipif(not (O.have_ref("bzip2") and O.have_ref("bzip2_b200")), reason="oracle/_ref bzip2 libs not built")
def test_reference_libbz2_links_against_libb200lc_and_output_is_identical():
    data = np.concatenate([norun_bytes(250000, 1, alphabet=16), texty(199905, 2)])
    bs = 100000 - 19
    # the reference's merge leaves origPtr unset when rotation 0 is taken on its "first bytes
    # differ" path (compress.c:636-639 `continue`) and then aborts (AssertH 1003); skip inputs
    # that would hit that bug
    for lo in range(0, data.size, bs):
        blk = np.ascontiguousarray(data[lo:lo + bs])
        f, a, b, r = O.bzip2_oracle_block_sort(blk)
        _, orig = O.bzip2_oracle_merge(blk, f, a, b, r)
        if orig < 0:
            pytest.skip("input would trip the reference's origPtr bug")
    ref = O.bzip2_ref_compress(data, 1, 0, "")           # reference GPU sort, all 5 blocks on the GPU
    mine = O.bzip2_ref_compress(data, 1, 0, "_b200")     # same reference objects + libb200lc.so
    assert ref.size == mine.size and np.array_equal(ref, mine)
    assert mine[:4].tobytes() == b"BZh1" and mine.size < data.size // 2
    # Not decoded here: through this entry path the reference resets every block CRC
    # (BZ_INITIALISE_CRC at the top of blocksort_wrapper, compress.c:716) and emits zero CRCs, which
    # any bzip2 decoder -- its own included -- rejects.  The block contents are covered by the
    # array-level parity tests above.


# ------------------------------------------------------------------------------ MTF + RLE stage (row N2)
from test_oracle_bzip2 import _mtf_cases  # noqa: E402


@pytest.mark.parametrize("name", list(_mtf_cases().keys()))
def test_mtf_rle_matches_oracle_and_reference(name):
    block = _mtf_cases()[name]
    n = block.size
    ptr = np.zeros(n, np.uint32)
    O.oracle().bzip2_oracle_rotation_order(block, n, ptr)
    gm, gf, gu = b200lc.bzip2_mtf_rle(block, ptr)
    om, of, ou = O.bzip2_oracle_mtf_rle(block, ptr)
    assert gu == ou and gm.size == om.size
    assert np.array_equal(gm, om) and np.array_equal(gf, of)
    if O.have_ref("bzip2_mtf"):
        rm, rf, ru = O.bzip2_ref_mtf_rle(block, ptr)
        assert np.array_equal(gm, rm) and np.array_equal(gf, rf) and gu == ru


def test_mtf_rle_full_900k_block_on_gpu_rotation_order():
    # a full -9 block: rotation order from the GPU sorter, MTF + RLE on the GPU, against the oracle
    n = 900000 - 19
    block = np.concatenate([texty(n // 2, 5), norun_bytes(n - n // 2, 6, alphabet=40)])
    ptr = np.zeros(n, np.uint32)
    orig = C.c_int(-1)
    L = b200lc.lib()
    L.b200lc_bzip2_rotation_order.restype = C.c_int
    L.b200lc_bzip2_rotation_order.argtypes = [np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_int,
                                              np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), C.POINTER(C.c_int)]
    block = np.ascontiguousarray(block)
    assert L.b200lc_bzip2_rotation_order(block, n, ptr, C.byref(orig)) == 0
    gm, gf, gu = b200lc.bzip2_mtf_rle(block, ptr)
    om, of, ou = O.bzip2_oracle_mtf_rle(block, ptr)
    assert gu == ou and np.array_equal(gm, om) and np.array_equal(gf, of)


@pytest.mark.skipif(not (O.have_ref("bzip2") and O.have_ref("bzip2_b200mtf")), reason="oracle/_ref bzip2 libs not built")
def test_reference_libbz2_with_gpu_sort_and_gpu_mtf_rle_is_byte_identical():
    # the reference library with BOTH its block sort and its generateMTFValues replaced by
    # libb200lc.so (oracle/_ref/libref_bzip2_b200mtf.so) writes the same .bz2 stream
    data = np.concatenate([norun_bytes(250000, 1, alphabet=16), texty(199905, 2)])
    bs = 100000 - 19
    for lo in range(0, data.size, bs):
        blk = np.ascontiguousarray(data[lo:lo + bs])
        f, a, b, r = O.bzip2_oracle_block_sort(blk)
        _, orig = O.bzip2_oracle_merge(blk, f, a, b, r)
        if orig < 0:
            pytest.skip("input would trip the reference's origPtr bug")
    ref = O.bzip2_ref_compress(data, 1, 0, "")
    mine = O.bzip2_ref_compress(data, 1, 0, "_b200mtf")
    assert ref.size == mine.size and np.array_equal(ref, mine)


# ------------------------------------------------------------------------------ Huffman stage (row N2)
@pytest.mark.parametrize("name", list(_mtf_cases().keys()))
def test_send_mtf_values_matches_oracle_and_reference(name):
    block = _mtf_cases()[name]
    n = block.size
    ptr = np.zeros(n, np.uint32)
    O.oracle().bzip2_oracle_rotation_order(block, n, ptr)
    mtfv, freq, used = O.bzip2_oracle_mtf_rle(block, ptr)
    in_use = O.bzip2_in_use(block)
    gb, gn, gl, gsel = b200lc.bzip2_send_mtf_values(mtfv, freq, in_use, used)
    ob, on, ol, osel, og = O.bzip2_oracle_send_mtf(mtfv, freq, in_use, used)
    assert gn == on and np.array_equal(gsel, osel)
    assert np.array_equal(gl[:og, : used + 2], ol[:og, : used + 2])
    assert np.array_equal(gb, ob)
    if O.have_ref("bzip2_mtf"):
        rb, rn, rl, rsel = O.bzip2_ref_send_mtf(mtfv, freq, in_use, used)
        assert gn == rn and np.array_equal(gb, rb)


def test_back_end_of_a_full_900k_block_on_the_gpu():
    # sort -> MTF + RLE -> Huffman stage of one -9 block, every stage on the GPU, against the oracle
    n = 900000 - 19
    block = np.ascontiguousarray(np.concatenate([texty(n // 2, 5), norun_bytes(n - n // 2, 6, alphabet=40)]))
    L = b200lc.lib()
    L.b200lc_bzip2_rotation_order.restype = C.c_int
    L.b200lc_bzip2_rotation_order.argtypes = [np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_int,
                                              np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), C.POINTER(C.c_int)]
    ptr = np.zeros(n, np.uint32)
    orig = C.c_int(-1)
    assert L.b200lc_bzip2_rotation_order(block, n, ptr, C.byref(orig)) == 0
    mtfv, freq, used = b200lc.bzip2_mtf_rle(block, ptr)
    in_use = O.bzip2_in_use(block)
    gb, gn, gl, gsel = b200lc.bzip2_send_mtf_values(mtfv, freq, in_use, used)
    ob, on, ol, osel, og = O.bzip2_oracle_send_mtf(mtfv, freq, in_use, used)
    assert og == 6 and gn == on and np.array_equal(gsel, osel) and np.array_equal(gb, ob)
    assert gn < 8 * n // 2


@pytest.mark.skipif(not (O.have_ref("bzip2") and O.have_ref("bzip2_b200full")), reason="oracle/_ref bzip2 libs not built")
def test_reference_libbz2_with_whole_gpu_back_end_is_byte_identical():
    # sort, generateMTFValues and sendMTFValues all from libb200lc.so (oracle/_ref/libref_bzip2_b200full.so)
    data = np.concatenate([norun_bytes(250000, 1, alphabet=16), texty(199905, 2)])
    bs = 100000 - 19
    for lo in range(0, data.size, bs):
        blk = np.ascontiguousarray(data[lo:lo + bs])
        f, a, b, r = O.bzip2_oracle_block_sort(blk)
        _, orig = O.bzip2_oracle_merge(blk, f, a, b, r)
        if orig < 0:
            pytest.skip("input would trip the reference's origPtr bug")
    ref = O.bzip2_ref_compress(data, 1, 0, "")
    mine = O.bzip2_ref_compress(data, 1, 0, "_b200full")
    assert ref.size == mine.size and np.array_equal(ref, mine)
