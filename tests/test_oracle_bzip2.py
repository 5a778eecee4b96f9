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
