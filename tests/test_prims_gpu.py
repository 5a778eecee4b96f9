"""GPU parity for the device-wide primitives under the BWT paths (csrc/devprims.cu): segmented
stable LSD radix sort of (key, value) pairs and single-pass scans, against numpy.  These replace
the CUB / Thrust / moderngpu calls of the reference (sa_app.cu:125-298, gpuBWTSort.cu:290-418).
Bar: bit-exact, including the order of equal keys (stability is what the BWT relies on)."""
import numpy as np
import pytest
import torch

from pkg import b200lc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _want_sorted(keys, seg_len, begin_bit, end_bit):
    n = keys.size
    width = end_bit - begin_bit
    field = (keys >> np.uint64(begin_bit)) & np.uint64((1 << width) - 1 if width < 64 else 0xFFFFFFFFFFFFFFFF)
    order = np.empty(n, np.int64)
    seg = seg_len if seg_len and seg_len < n else n
    for lo in range(0, n, seg):
        hi = min(n, lo + seg)
        order[lo:hi] = lo + np.argsort(field[lo:hi], kind="stable")
    return order


@pytest.mark.parametrize("wide", [True, False])
@pytest.mark.parametrize("n,seg_len,begin_bit,end_bit,keyspace", [
    (1, 0, 0, 8, 256),
    (4095, 0, 0, 16, 1 << 16),
    (4096, 0, 0, 32, 1 << 32),
    (4097, 0, 3, 27, 1 << 32),
    (100003, 0, 0, 32, 1 << 32),
    (100003, 0, 0, 32, 7),              # few distinct keys: long runs of equal digits
    (1 << 20, 1 << 18, 0, 24, 1 << 24),  # aligned segments (TMA path)
    (300000, 70001, 0, 20, 1 << 20),     # ragged segments (guarded loads), short last segment
    (50000, 4096, 8, 16, 1 << 16),       # one tile per segment
    (50000, 1000, 0, 8, 256),            # segments smaller than a tile
])
def test_sort_pairs_matches_stable_argsort(wide, n, seg_len, begin_bit, end_bit, keyspace):
    rng = np.random.default_rng(n * 31 + seg_len + begin_bit)
    keys = rng.integers(0, keyspace, n, dtype=np.uint64)
    if wide:
        keys |= rng.integers(0, 1 << 20, n, dtype=np.uint64) << np.uint64(44)   # bits outside the sorted range
        if end_bit == 32:
            end_bit = 64
    else:
        keys &= np.uint64(0xFFFFFFFF)
    vals = np.arange(n, dtype=np.uint32)
    order = _want_sorted(keys, seg_len, begin_bit, end_bit)
    dk = _dev(keys.view(np.int64) if wide else keys.astype(np.uint32).view(np.int32))
    dv = _dev(vals.view(np.int32))
    sk, sv = b200lc.sort_pairs(dk, dv, seg_len, begin_bit, end_bit)
    got_v = sv.cpu().numpy().view(np.uint32)
    got_k = sk.cpu().numpy()
    got_k = got_k.view(np.uint64) if wide else got_k.view(np.uint32).astype(np.uint64)
    assert np.array_equal(got_v, order.astype(np.uint32))
    assert np.array_equal(got_k, keys[order])


def test_sort_pairs_large_skewed():
    # 16 Mi pairs, head-position style keys (sorted high bits, random low bits): look-back chains
    # across thousands of tiles and warp-uniform digits in the upper passes
    n = 1 << 24
    rng = np.random.default_rng(7)
    hi = np.sort(rng.integers(0, 1 << 27, n, dtype=np.uint64))
    keys = (hi << np.uint64(21)) | rng.integers(0, 1 << 21, n, dtype=np.uint64)
    perm = rng.permutation(n)
    keys = keys[perm]
    order = np.argsort(keys, kind="stable")
    sk, sv = b200lc.sort_pairs(_dev(keys.view(np.int64)), _dev(np.arange(n, dtype=np.int32)), 0, 0, 48)
    assert np.array_equal(sv.cpu().numpy().view(np.uint32), order.astype(np.uint32))
    assert np.array_equal(sk.cpu().numpy().view(np.uint64), keys[order])


@pytest.mark.parametrize("n", [1, 31, 4095, 4096, 4097, 1 << 20, (1 << 22) + 12345])
@pytest.mark.parametrize("kind", ["exclusive_sum", "inclusive_max"])
def test_scans(n, kind):
    rng = np.random.default_rng(n)
    if kind == "exclusive_sum":
        x = rng.integers(0, 2, n, dtype=np.uint32)
        want = np.concatenate([[0], np.cumsum(x[:-1], dtype=np.uint64)]).astype(np.uint32)
    else:
        # "heads" pattern of the BWT: mostly zeros, position index at group starts
        x = np.where(rng.random(n) < 0.3, np.arange(n, dtype=np.uint32), 0).astype(np.uint32)
        want = np.maximum.accumulate(x)
    d = _dev(x.view(np.int32))
    got = b200lc.scan_u32(d, kind).cpu().numpy().view(np.uint32)
    assert np.array_equal(got, want)
    # in place
    got2 = b200lc.scan_u32(d, kind, out=d).cpu().numpy().view(np.uint32)
    assert np.array_equal(got2, want)
