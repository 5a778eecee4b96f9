"""CPU: the cuda-bzip2 block-sort contract restated in oracle/bzip2_oracle.c is self-consistent:
merging the three gpuBlockSort arrays with the restated merge_two_sort_arrays
(compress.c:609-710) gives exactly the rotation order, for every n mod 3."""
import numpy as np
import pytest

import oracle_lib as O


@pytest.mark.parametrize("n", [10, 11, 12, 1000, 1001, 1002, 30000])
@pytest.mark.parametrize("kind", ["rand4", "text"])
def test_merge_of_block_sort_arrays_is_rotation_order(n, kind):
    rng = np.random.default_rng(n)
    if kind == "rand4":
        block = rng.integers(0, 4, n, dtype=np.uint8)        # small alphabet: deep ties
    else:
        block = np.frombuffer((b"it was the best of times, it was the worst of times, " * 700)[:n], np.uint8).copy()
        block[rng.integers(0, n, 3)] = rng.integers(0, 255, 3)   # break exact periodicity
    ptr = np.zeros(n, np.uint32)
    O.oracle().bzip2_oracle_rotation_order(block, n, ptr)
    # brute-force check of the rotation order on the small cases
    if n <= 1002:
        dbl = np.concatenate([block, block])
        keys = [bytes(dbl[i:i + n]) for i in range(n)]
        assert [int(x) for x in ptr] == sorted(range(n), key=lambda i: (keys[i], -i))
    f, first, second, rank = O.bzip2_oracle_block_sort(block)
    assert f == 2 * ((n - 1) // 3) + (n - 1) % 3 + (1 if n % 3 == 1 else 0)
    order, orig = O.bzip2_oracle_merge(block, f, first, second, rank)
    assert np.array_equal(order, ptr)
    assert orig == -1 or ptr[orig] == 0


def _mtf_cases():
    rng = np.random.default_rng(9)
    text = np.frombuffer((b"it was the best of times, it was the worst of times, " * 2000), np.uint8)
    return {
        "single_symbol": np.full(5000, 65, np.uint8),                  # one long zero run
        "two_symbols": rng.integers(7, 9, 3000, dtype=np.uint8),
        "binary_runs": np.repeat(rng.integers(0, 2, 400, dtype=np.uint8), rng.integers(1, 40, 400)),
        "text_100k": text[:100000].copy(),
        "random_50k": rng.integers(0, 256, 50000, dtype=np.uint8),
        "zipf_120k": O.zipf_bytes(120000, 1.5, seed=4),
        "n1": np.array([200], np.uint8),
    }


@pytest.mark.skipif(not O.have_ref("bzip2_mtf"), reason="oracle/_ref/libref_bzip2_mtf.so not built")
@pytest.mark.parametrize("name", list(_mtf_cases().keys()))
def test_mtf_rle_oracle_equals_reference_generateMTFValues(name):
    block = _mtf_cases()[name]
    n = block.size
    ptr = np.zeros(n, np.uint32)
    O.oracle().bzip2_oracle_rotation_order(block, n, ptr)
    om, of, ou = O.bzip2_oracle_mtf_rle(block, ptr)
    rm, rf, ru = O.bzip2_ref_mtf_rle(block, ptr)
    assert ou == ru and om.size == rm.size
    assert np.array_equal(om, rm) and np.array_equal(of, rf)
    assert om[-1] == ou + 1 and int(of.sum()) == om.size       # EOB last; every symbol counted once


@pytest.mark.skipif(not O.have_ref("bzip2_mtf"), reason="oracle/_ref/libref_bzip2_mtf.so not built")
@pytest.mark.parametrize("name", list(_mtf_cases().keys()))
def test_send_mtf_oracle_equals_reference_sendMTFValues(name):
    block = _mtf_cases()[name]
    n = block.size
    ptr = np.zeros(n, np.uint32)
    O.oracle().bzip2_oracle_rotation_order(block, n, ptr)
    mtfv, freq, used = O.bzip2_oracle_mtf_rle(block, ptr)
    in_use = O.bzip2_in_use(block)
    ob, on, ol, osel, og = O.bzip2_oracle_send_mtf(mtfv, freq, in_use, used)
    rb, rn, rl, rsel = O.bzip2_ref_send_mtf(mtfv, freq, in_use, used)
    assert on == rn
    assert np.array_equal(osel, rsel)
    assert np.array_equal(ol[:og, : used + 2], rl[:og, : used + 2])
    assert np.array_equal(ob, rb)


@pytest.mark.skipif(not O.have_ref("bzip2"), reason="oracle/_ref/libref_bzip2.so not built")
def test_code_lengths_oracle_equals_reference_hbMakeCodeLengths():
    import ctypes as C
    import os
    lib = C.CDLL(os.path.join(O.ORACLE_DIR, "_ref", "libref_bzip2.so"))
    ref = getattr(lib, "_Z21BZ2_hbMakeCodeLengthsPhPiii")        # compiled as C++ (Makefile:18)
    ref.restype = None
    rng = np.random.default_rng(3)
    for trial in range(60):
        alpha = int(rng.integers(3, 259))
        kind = trial % 4
        if kind == 0:
            freq = rng.integers(0, 50, alpha)
        elif kind == 1:
            freq = (rng.pareto(0.7, alpha) * 10).astype(np.int64) % 2000000      # heavy tail: limit 17 gets hit
        elif kind == 2:
            freq = np.zeros(alpha, np.int64)
            freq[rng.integers(0, alpha, 3)] = rng.integers(1, 1000, 3)
        else:
            fib = [1, 1]
            while len(fib) < alpha:
                fib.append(min(fib[-1] + fib[-2], 4000000 // alpha))     # (freq << 8) must stay inside int32
            freq = np.array(fib[:alpha])
        freq = np.ascontiguousarray(freq, dtype=np.int32)
        a = np.zeros(alpha, np.uint8)
        b = np.zeros(alpha, np.uint8)
        O.oracle().bzip2_oracle_code_lengths(a, freq, alpha, 17)
        ref(b.ctypes.data_as(C.c_void_p), freq.ctypes.data_as(C.c_void_p), alpha, 17)
        assert np.array_equal(a, b), (trial, alpha)
        assert a.max() <= 17 and a.min() >= 1


@pytest.mark.parametrize("name", ["bzip2_text", "bzip2_zipf"])
def test_back_end_oracle_matches_golden_from_reference(name):
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))   # tools/make_golden.py
    block, ptr = g["block"], g["ptr"]
    mtfv, freq, used = O.bzip2_oracle_mtf_rle(block, ptr)
    assert used == int(g["n_in_use"]) and np.array_equal(mtfv, g["mtfv"]) and np.array_equal(freq, g["freq"])
    bits, nbits, lens, sel, groups = O.bzip2_oracle_send_mtf(mtfv, freq, O.bzip2_in_use(block), used)
    assert nbits == int(g["nbits"]) and np.array_equal(bits, g["bits"]) and np.array_equal(sel, g["selector"])
    assert np.array_equal(lens[:groups, : used + 2], g["lens"][:groups, : used + 2])
