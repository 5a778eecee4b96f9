"""GPU parity for hot path 1 (cudppCompress = BWT -> MTF -> Huffman) against the CPU oracle
(oracle/cudpp_oracle.c, pinned to the reference testrig golds) and, where it applies, against
the reference's own gold code in oracle/_ref/libref_cudpp.so.  Sizes / inputs follow the
reference tests (test_compress.cpp:375-377,439-444,515,552-556,687-692).  Bar: bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle_lib as O
from pkg import b200lc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MIB = 1 << 20


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# ------------------------------------------------------------------------------------------ MTF
MTF_SIZES = [39, 128, 256, 1023, 1024, 1025, 2047, 2048, 2049, 4095, 4096, 4097, 65535, 65536,
             65537, 1048575, 1048581]


@pytest.mark.parametrize("n", MTF_SIZES)
def test_mtf_reference_test_sizes(n):
    # reference test: srand(95835), bytes rand() % 255 + 1 (test_compress.cpp:439-441)
    rng = np.random.Generator(np.random.MT19937(95835 + n))
    data = rng.integers(1, 256, n, dtype=np.uint8)
    got = b200lc.mtf_batch(_dev(data), 1, n).cpu().numpy()
    assert np.array_equal(got, O.cudpp_oracle_mtf(data))


@pytest.mark.parametrize("kind", ["zipf", "markov", "text", "zeros"])
def test_mtf_batched_blocks(kind):
    n, nb = 70001, 5
    blocks = [np.zeros(n, np.uint8) if kind == "zeros" else O.cudpp_block(n, kind, seed=b) for b in range(nb)]
    got = b200lc.mtf_batch(_dev(np.concatenate(blocks)), nb, n).cpu().numpy()
    for b in range(nb):
        assert np.array_equal(got[b * n:(b + 1) * n], O.cudpp_oracle_mtf(blocks[b])), b


# ------------------------------------------------------------------------------------------ BWT
def test_bwt_reference_test_vector():
    # n = 1,048,576, glibc srand(95835), rand() % 255 + 1 (test_compress.cpp:552-556)
    data = O.cudpp_test_vector(MIB, sentinel=False)
    want, widx = O.cudpp_oracle_bwt(data)
    got, idx = b200lc.bwt_batch(_dev(data), 1, MIB)
    assert int(idx.cpu()[0]) == widx
    assert np.array_equal(got.cpu().numpy(), want)
    if O.have_ref("cudpp"):
        rb = np.zeros(MIB, np.uint8)
        ridx = C.c_int(-1)
        O.ref_cudpp().ref_cudpp_bwt(data, rb, C.byref(ridx), MIB)
        assert ridx.value == widx and np.array_equal(rb, want)


@pytest.mark.parametrize("kind,n,nb", [("zipf", 65536, 4), ("markov", 100000, 3), ("text", 262144, 2),
                                       ("rand", 4097, 7), ("zeros", 5000, 2), ("rand", 1, 3)])
def test_bwt_batched_blocks(kind, n, nb):
    blocks = [np.zeros(n, np.uint8) if kind == "zeros" else O.cudpp_block(n, kind, seed=b) for b in range(nb)]
    got, idx = b200lc.bwt_batch(_dev(np.concatenate(blocks)), nb, n)
    got, idx = got.cpu().numpy(), idx.cpu().numpy()
    for b in range(nb):
        want, widx = O.cudpp_oracle_bwt(blocks[b])
        assert idx[b] == widx, b
        assert np.array_equal(got[b * n:(b + 1) * n], want), b


@pytest.mark.parametrize("n", [2, 5, 6, 7, 13, 100, 4099])
@pytest.mark.parametrize("alphabet", [1, 2, 3])
def test_bwt_short_suffixes_and_zero_bytes(n, alphabet):
    # bytes from {0..alphabet-1}: many suffixes tie on their first six bytes with zero padding,
    # which is where the first-round key relies on sort stability instead of a length field
    nb = 9
    rng = np.random.default_rng(n * 10 + alphabet)
    blocks = [rng.integers(0, alphabet, n, dtype=np.uint8) for _ in range(nb)]
    blocks[0][:] = 0
    got, idx = b200lc.bwt_batch(_dev(np.concatenate(blocks)), nb, n)
    got, idx = got.cpu().numpy(), idx.cpu().numpy()
    for b in range(nb):
        want, widx = O.cudpp_oracle_bwt(blocks[b])
        assert idx[b] == widx, b
        assert np.array_equal(got[b * n:(b + 1) * n], want), b


def test_suffix_array_matches_gold():
    # test_sa.cpp:124-126: bytes rand() % 128 + 1
    n = 200001
    rng = np.random.Generator(np.random.MT19937(5))
    data = rng.integers(1, 129, n, dtype=np.uint8)
    sa = b200lc.suffix_array_batch(_dev(data), 1, n).cpu().numpy().astype(np.uint32)
    want = np.zeros(n, np.uint32)
    O.oracle().cudpp_oracle_sa(data, n, want)
    assert np.array_equal(sa, want)
    if O.have_ref("cudpp"):
        ref = np.zeros(n + 3, np.uint32)
        O.ref_cudpp().ref_cudpp_sa(data, ref, n)
        assert np.array_equal(ref[:n], want)


# ------------------------------------------------------------------------------------------ compress
def _check_block(res, b, data, nhb):
    rc, widx, whist, woffs, wwords = O.cudpp_oracle_compress(data)
    assert rc == 0
    stride = res.stride
    tw = int(res.total_words.cpu()[b])
    assert int(res.bwt_index.cpu()[b]) == widx
    assert np.array_equal(res.hist.cpu().numpy()[b * 256:(b + 1) * 256].astype(np.uint32), whist)
    assert np.array_equal(res.offsets.cpu().numpy()[b * nhb:(b + 1) * nhb].astype(np.uint32), woffs)
    assert tw == wwords.size
    words = res.words.cpu().numpy()[b * stride: b * stride + tw].view(np.uint32)
    assert np.array_equal(words, wwords)
    return widx, whist, woffs, wwords


def test_compress_reference_test_vector_and_reference_decoder():
    # test_compress.cpp:687-692: rand() % 255 + 1 with a trailing 0 sentinel, n = 1,048,576
    data = O.cudpp_test_vector(MIB, sentinel=True)
    res = b200lc.cudpp_compress_batch(_dev(data), 1, MIB)
    assert int(res.error.cpu()[0]) == 0
    idx, hist, offs, words = _check_block(res, 0, data, 256)
    if O.have_ref("cudpp"):
        # the reference's own decoder (computeCompressGold) must reproduce the input
        out = np.zeros(MIB, np.uint8)
        h257 = np.zeros(257, np.uint32)
        h257[:256] = hist
        O.ref_cudpp().ref_cudpp_decompress(out, idx, h257, offs.copy(), words.size, words.copy(), MIB)
        assert np.array_equal(out, data)


@pytest.mark.parametrize("kind", ["zipf", "markov", "text"])
def test_compress_batch_blocks(kind):
    n, nb = MIB, 3
    blocks = [O.cudpp_block(n, kind, seed=10 + b) for b in range(nb)]
    res = b200lc.cudpp_compress_batch(_dev(np.concatenate(blocks)), nb, n)
    assert int(res.error.cpu()[0]) == 0
    for b in range(nb):
        idx, hist, offs, words = _check_block(res, b, blocks[b], 256)
        rc, back = O.cudpp_oracle_decompress(n, idx, hist, offs, words)
        assert rc == 0 and np.array_equal(back, blocks[b])


def test_compress_small_blocks():
    n, nb = 8192, 6
    blocks = [O.cudpp_block(n, "zipf", seed=30 + b) for b in range(nb)]
    res = b200lc.cudpp_compress_batch(_dev(np.concatenate(blocks)), nb, n)
    for b in range(nb):
        _check_block(res, b, blocks[b], 2)


def test_cudpp_named_entry_points():
    """cudppCreate / cudppPlan / cudppCompress / ... with the reference's call sequence
    (test_compress.cpp:650-747) and its error codes (cudpp.cpp:776-805)."""
    L = b200lc.lib()

    class Config(C.Structure):
        _fields_ = [("algorithm", C.c_int), ("op", C.c_int), ("datatype", C.c_int),
                    ("options", C.c_uint), ("bucket_mapper", C.c_int)]
    CUDPP_UCHAR, CUDPP_UINT, CUDPP_COMPRESS, CUDPP_BWT, CUDPP_MTF, CUDPP_SCAN = 1, 5, 10, 12, 13, 0
    vp, sz = C.c_void_p, C.c_size_t
    L.cudppCreate.argtypes = [C.POINTER(sz)]
    L.cudppPlan.argtypes = [sz, C.POINTER(sz), Config, sz, sz, sz]
    L.cudppCompress.argtypes = [sz, vp, vp, vp, vp, vp, vp, vp, sz]
    L.cudppBurrowsWheelerTransform.argtypes = [sz, vp, vp, vp, sz]
    L.cudppMoveToFrontTransform.argtypes = [sz, vp, vp, sz]
    L.cudppDestroyPlan.argtypes = [sz]
    L.cudppDestroy.argtypes = [sz]
    mgr = sz(0)
    assert L.cudppCreate(C.byref(mgr)) == 0
    plan = sz(0)
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_SCAN, 0, CUDPP_UINT, 0, 0), MIB, 1, 0) == 2
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_COMPRESS, 0, CUDPP_UINT, 0, 0), MIB, 1, 0) == 2
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_COMPRESS, 0, CUDPP_UCHAR, 0, 0), MIB, 1, 0) == 0

    data = O.cudpp_block(MIB, "zipf", seed=77)
    d_in = _dev(data)
    d_idx = torch.zeros(1, dtype=torch.int32, device=DEV)
    d_hist = torch.zeros(256, dtype=torch.int32, device=DEV)
    d_off = torch.zeros(256, dtype=torch.int32, device=DEV)
    d_size = torch.zeros(1, dtype=torch.int32, device=DEV)
    d_comp = torch.zeros(256 * 1537, dtype=torch.int32, device=DEV)
    assert L.cudppCompress(0, d_in.data_ptr(), d_idx.data_ptr(), None, d_hist.data_ptr(), d_off.data_ptr(),
                           d_size.data_ptr(), d_comp.data_ptr(), MIB) == 1          # invalid handle
    assert L.cudppBurrowsWheelerTransform(plan, d_in.data_ptr(), d_in.data_ptr(), d_idx.data_ptr(), MIB) == 3
    assert L.cudppCompress(plan, d_in.data_ptr(), d_idx.data_ptr(), None, d_hist.data_ptr(),
                           d_off.data_ptr(), d_size.data_ptr(), d_comp.data_ptr(), MIB) == 0
    torch.cuda.synchronize()
    rc, widx, whist, woffs, wwords = O.cudpp_oracle_compress(data)
    assert int(d_idx.cpu()[0]) == widx and int(d_size.cpu()[0]) == wwords.size
    assert np.array_equal(d_comp.cpu().numpy()[: wwords.size].view(np.uint32), wwords)
    assert np.array_equal(d_off.cpu().numpy().astype(np.uint32), woffs)
    assert L.cudppDestroyPlan(plan) == 0

    # standalone BWT and MTF plans (test_compress.cpp:453,588)
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_BWT, 0, CUDPP_UCHAR, 0, 0), MIB, 1, 0) == 0
    d_out = torch.zeros(MIB, dtype=torch.uint8, device=DEV)
    assert L.cudppBurrowsWheelerTransform(plan, d_in.data_ptr(), d_out.data_ptr(), d_idx.data_ptr(), MIB) == 0
    torch.cuda.synchronize()
    wb, wi = O.cudpp_oracle_bwt(data)
    assert int(d_idx.cpu()[0]) == wi and np.array_equal(d_out.cpu().numpy(), wb)
    assert L.cudppDestroyPlan(plan) == 0
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_MTF, 0, CUDPP_UCHAR, 0, 0), MIB, 1, 0) == 0
    n = 1048571      # plans are sized for max n; any smaller n is accepted (test_compress.cpp:375-377)
    assert L.cudppMoveToFrontTransform(plan, d_in.data_ptr(), d_out.data_ptr(), n) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy()[:n], O.cudpp_oracle_mtf(data[:n]))
    assert L.cudppDestroyPlan(plan) == 0
    assert L.cudppDestroy(mgr) == 0


@pytest.mark.skipif(not O.have_ref("cudpp_gpu"), reason="oracle/_ref/libref_cudpp_gpu.so not built")
@pytest.mark.parametrize("kind,seed", [("zipf", 21), ("text", 22), ("markov", 23), ("vector", 95835)])
def test_reference_huffman_kernels_write_the_same_words(kind, seed):
    """The reference's own huffman_build_histogram / _build_tree / huffman_kernel_en /
    huffman_datapack kernels (compress_kernel.cuh:2037-2750, compiled for sm_100a, launched as
    compress_app.cu:65-125 does) on the MTF bytes of a 1 MiB block: histogram, block offsets, size
    and every compressed word equal the product's and the oracle's."""
    data = O.cudpp_test_vector() if kind == "vector" else O.cudpp_block(MIB, kind, seed=seed)
    bwt, _ = O.cudpp_oracle_bwt(data)
    mtf = O.cudpp_oracle_mtf(bwt)
    d_mtf = _dev(mtf)
    d_hist = torch.zeros(256, dtype=torch.int32, device=DEV)
    d_off = torch.zeros(256, dtype=torch.int32, device=DEV)
    d_size = torch.zeros(1, dtype=torch.int32, device=DEV)
    d_comp = torch.zeros(256 * 1537, dtype=torch.int32, device=DEV)
    torch.cuda.synchronize()
    assert O.ref_cudpp_gpu().ref_cudpp_huffman_gpu(d_mtf.data_ptr(), MIB, d_hist.data_ptr(), d_off.data_ptr(),
                                                   d_size.data_ptr(), d_comp.data_ptr()) == 0
    rc, whist, woffs, wwords = O.cudpp_oracle_huffman(mtf)
    assert rc == 0
    nw = int(d_size.cpu()[0])
    assert nw == wwords.size
    assert np.array_equal(d_hist.cpu().numpy().view(np.uint32), whist)
    assert np.array_equal(d_off.cpu().numpy().view(np.uint32), woffs)
    assert np.array_equal(d_comp.cpu().numpy()[:nw].view(np.uint32), wwords)
    res = b200lc.cudpp_compress_batch(_dev(data), 1, MIB)
    assert int(res.total_words.cpu()[0]) == nw
    assert torch.equal(res.words[:nw], d_comp[:nw]) and torch.equal(res.offsets.view(-1)[:256], d_off)


@pytest.mark.parametrize("n", [1, 1000, MIB, MIB + 5])
def test_cudpp_radix_sort_is_the_sort_of_the_reference_tests_decoder(n):
    """cudppRadixSort as apps/cudpp_testrig/test_compress.cpp:318-344 calls it: plan
    {CUDPP_SORT_RADIX, CUDPP_UCHAR, KEY_VALUE_PAIRS}, keys = the BWT bytes, values = 0..n-1; the
    sorted values are the permutation the test walks from the BWT index.  Stable, in place."""
    L = b200lc.lib()

    class Config(C.Structure):
        _fields_ = [("algorithm", C.c_int), ("op", C.c_int), ("datatype", C.c_int),
                    ("options", C.c_uint), ("bucket_mapper", C.c_int)]
    CUDPP_UCHAR, CUDPP_UINT, CUDPP_FLOAT, CUDPP_SORT_RADIX, KEYS_ONLY, PAIRS = 1, 5, 6, 4, 0x20, 0x40
    vp, sz = C.c_void_p, C.c_size_t
    L.cudppCreate.argtypes = [C.POINTER(sz)]
    L.cudppPlan.argtypes = [sz, C.POINTER(sz), Config, sz, sz, sz]
    L.cudppRadixSort.argtypes = [sz, vp, vp, sz]
    L.cudppDestroyPlan.argtypes = [sz]
    L.cudppDestroy.argtypes = [sz]
    mgr, plan = sz(0), sz(0)
    assert L.cudppCreate(C.byref(mgr)) == 0
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_SORT_RADIX, 0, CUDPP_FLOAT, PAIRS, 0), n, 1, 0) == 2
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_SORT_RADIX, 0, CUDPP_UCHAR, PAIRS, 0), n, 1, 0) == 0
    data = O.cudpp_block(n, "zipf", seed=n % 71)
    bwt, idx = O.cudpp_oracle_bwt(data)
    d_keys = _dev(bwt)
    d_vals = torch.arange(n, dtype=torch.int32, device=DEV)
    assert L.cudppRadixSort(0, d_keys.data_ptr(), d_vals.data_ptr(), n) == 1           # invalid handle
    assert L.cudppRadixSort(plan, d_keys.data_ptr(), d_vals.data_ptr(), n + 1) == 2     # larger than the plan
    assert L.cudppRadixSort(plan, d_keys.data_ptr(), d_vals.data_ptr(), n) == 0
    torch.cuda.synchronize()
    order = np.argsort(bwt, kind="stable")
    assert np.array_equal(d_keys.cpu().numpy(), bwt[order])
    assert np.array_equal(d_vals.cpu().numpy().view(np.uint32), order.astype(np.uint32))
    # the walk of test_compress.cpp:346-362 over the sorted values reproduces the input
    if n <= 1000:
        vals = d_vals.cpu().numpy()
        at, out = int(idx), np.zeros(n, np.uint8)
        for i in range(n):
            at = int(vals[at])
            out[i] = bwt[at]
        assert np.array_equal(out, data)
    assert L.cudppDestroyPlan(plan) == 0
    # unsigned int keys, keys only
    assert L.cudppPlan(mgr, C.byref(plan), Config(CUDPP_SORT_RADIX, 0, CUDPP_UINT, KEYS_ONLY, 0), n, 1, 0) == 0
    k = np.random.default_rng(n).integers(0, 1 << 32, n, dtype=np.uint32)
    d_k = _dev(k.view(np.int32))
    assert L.cudppRadixSort(plan, d_k.data_ptr(), None, n) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_k.cpu().numpy().view(np.uint32), np.sort(k))
    assert L.cudppDestroyPlan(plan) == 0
    assert L.cudppDestroy(mgr) == 0


# ------------------------------------------------------------------------------------------ decoder (N3)
def _np_inverse_mtf(r):
    lst = list(range(256))
    out = np.empty(r.size, np.uint8)
    for i, x in enumerate(r.tolist()):
        c = lst.pop(x)
        out[i] = c
        lst.insert(0, c)
    return out


@pytest.mark.parametrize("n", [1, 39, 2047, 2048, 2049, 4097, 65537, 300001])
def test_inverse_mtf_sizes(n):
    rng = np.random.Generator(np.random.MT19937(7 + n))
    data = rng.integers(0, 256, n, dtype=np.uint8) if n % 2 else O.cudpp_block(n, "zipf", seed=n)
    ranks = O.cudpp_oracle_mtf(data)
    got = b200lc.inverse_mtf_batch(_dev(ranks), 1, n).cpu().numpy()
    assert np.array_equal(got, data)
    if n <= 5000:
        assert np.array_equal(_np_inverse_mtf(ranks), data)   # the oracle's MTF is what we invert


@pytest.mark.parametrize("kind", ["zipf", "markov", "text", "rand", "zeros"])
def test_inverse_mtf_batched(kind):
    n, nb = 70001, 5
    blocks = [np.zeros(n, np.uint8) if kind == "zeros" else O.cudpp_block(n, kind, seed=b) for b in range(nb)]
    ranks = np.concatenate([O.cudpp_oracle_mtf(b) for b in blocks])
    got = b200lc.inverse_mtf_batch(_dev(ranks), nb, n).cpu().numpy()
    assert np.array_equal(got, np.concatenate(blocks))


def _ibwt_case(blocks):
    n, nb = blocks[0].size, len(blocks)
    pairs = [O.cudpp_oracle_bwt(b) for b in blocks]
    bwt = np.concatenate([p[0] for p in pairs])
    idx = np.array([p[1] for p in pairs], np.int32)
    out, err = b200lc.inverse_bwt_batch(_dev(bwt), _dev(idx), nb, n)
    assert int(err.cpu()[0]) == 0
    assert np.array_equal(out.cpu().numpy(), np.concatenate(blocks))


@pytest.mark.parametrize("kind,n,nb", [("zipf", 65536, 4), ("markov", 100000, 3), ("text", 262144, 2),
                                       ("rand", 1000, 7), ("rand", 1, 3), ("rand", 2, 2), ("zipf", 255, 1),
                                       ("zipf", 257, 2), ("text", MIB, 2), ("zipf", (1 << 21) - 1, 1)])
def test_inverse_bwt_blocks(kind, n, nb):
    _ibwt_case([O.cudpp_block(n, kind, seed=40 + b) for b in range(nb)])


@pytest.mark.parametrize("period,n", [(1, 5000), (2, 4096), (3, 3000), (7, 7 * 1024), (256, 65536), (5, 5)])
def test_inverse_bwt_periodic_blocks(period, n):
    """A periodic block makes the walk close its cycle before n steps (T is not one cycle); no
    trailing sentinel here on purpose."""
    rng = np.random.Generator(np.random.MT19937(period))
    unit = rng.permutation(256)[:period].astype(np.uint8)
    _ibwt_case([np.tile(unit, n // period)[:n].copy(), np.tile(unit[::-1], n // period)[:n].copy()])


def test_inverse_bwt_rejects_bad_index():
    n = 4096
    blk = O.cudpp_block(n, "zipf", seed=1)
    bwt, _ = O.cudpp_oracle_bwt(blk)
    _, err = b200lc.inverse_bwt_batch(_dev(bwt), _dev(np.array([n], np.int32)), 1, n)
    assert int(err.cpu()[0]) == 6


def test_decompress_reference_test_vector():
    data = O.cudpp_test_vector(MIB, sentinel=True)
    res = b200lc.cudpp_compress_batch(_dev(data), 1, MIB)
    out, err = b200lc.cudpp_decompress_batch(res, 1, MIB)
    assert int(err.cpu()[0]) == 0
    assert np.array_equal(out.cpu().numpy(), data)


@pytest.mark.parametrize("kind,n,nb", [("zipf", MIB, 3), ("markov", MIB, 2), ("text", MIB, 2), ("rand", 8192, 6),
                                       ("zipf", 100000, 3), ("zipf", 4095, 2), ("markov", 4097, 5),
                                       ("zipf", 1, 2), ("text", 600000, 1)])
def test_decompress_round_trip(kind, n, nb):
    blocks = [O.cudpp_block(n, kind, seed=50 + b) for b in range(nb)]
    data = np.concatenate(blocks)
    res = b200lc.cudpp_compress_batch(_dev(data), nb, n)
    assert int(res.error.cpu()[0]) == 0
    out, err = b200lc.cudpp_decompress_batch(res, nb, n)
    assert int(err.cpu()[0]) == 0
    assert np.array_equal(out.cpu().numpy(), data)


def test_decompress_oracle_stream():
    """Streams produced by the CPU oracle (pinned to the reference golds) decode on the GPU, and
    the GPU decoder agrees with the oracle's decoder."""
    n, nb = 200000, 3
    nhb = (n + 4095) // 4096
    stride = nhb * 1537
    blocks = [O.cudpp_block(n, k, seed=60 + i) for i, k in enumerate(["zipf", "markov", "text"])]
    idx = np.zeros(nb, np.int32)
    hist = np.zeros(nb * 256, np.uint32)
    offs = np.zeros(nb * nhb, np.uint32)
    words = np.zeros(nb * stride, np.uint32)
    tw = np.zeros(nb, np.uint32)
    for b, blk in enumerate(blocks):
        rc, i, h, o, w = O.cudpp_oracle_compress(blk)
        assert rc == 0
        idx[b], hist[b * 256:(b + 1) * 256], offs[b * nhb:(b + 1) * nhb] = i, h, o
        words[b * stride:b * stride + w.size] = w
        tw[b] = w.size
        rc, back = O.cudpp_oracle_decompress(n, i, h, o, w)
        assert rc == 0 and np.array_equal(back, blk)
    comp = b200lc.CudppCompressed(_dev(idx), _dev(hist.view(np.int32)), _dev(offs.view(np.int32)),
                                  _dev(tw.view(np.int32)), _dev(words.view(np.int32)), stride,
                                  torch.zeros(1, dtype=torch.int32, device=DEV))
    out, err = b200lc.cudpp_decompress_batch(comp, nb, n)
    assert int(err.cpu()[0]) == 0
    assert np.array_equal(out.cpu().numpy(), np.concatenate(blocks))


def test_decompress_flags_corrupt_stream():
    n = 65536
    data = O.cudpp_block(n, "zipf", seed=70)
    res = b200lc.cudpp_compress_batch(_dev(data), 1, n)
    res.words[5:40] = -1          # all-ones words: run of the longest code -> symbols no longer add up
    res.offsets[3] = 10 ** 9      # outside the stream
    _, err = b200lc.cudpp_decompress_batch(res, 1, n)
    assert int(err.cpu()[0]) in (4, 5)
