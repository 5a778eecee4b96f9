"""libbsc's block container behind the reference's entry points (SURVEY.md 8b libbsc row):
bsc_init / bsc_store / bsc_block_info / bsc_decompress / bsc_compress exported by libb200lc.so
(include/libbsc_gpu.h) against the reference's own library (oracle/_ref/libref_bsc.so, built from
cuda-bsc/libbsc/libbsc/libbsc.cpp).  CPU part: stored blocks, header validation and error codes need
no GPU.  GPU part: bsc_compress with the block sort on the GPU and the reference's CPU stages (LZP,
QLFC, inverse BWT) registered through b200lc_bsc_set_stages -- byte-identical blocks."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from pkg import b200lc
from test_ref_bsc_cpu import synthetic_largefile

HEADER = 28
NOT_SUPPORTED, UNEXPECTED_EOB, DATA_CORRUPT, BAD_PARAMETER = -4, -5, -6, -1
u8p = C.POINTER(C.c_ubyte)


class Stages(C.Structure):
    _fields_ = [("coder_compress", C.c_void_p), ("coder_decompress", C.c_void_p), ("lzp_compress", C.c_void_p),
                ("lzp_decompress", C.c_void_p), ("bwt_decode", C.c_void_p), ("st_decode", C.c_void_p)]


def _protos(lib):
    lib.bsc_init.argtypes = [C.c_int]
    lib.bsc_store.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.bsc_compress.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6
    lib.bsc_block_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]
    lib.bsc_decompress.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    for f in (lib.bsc_init, lib.bsc_store, lib.bsc_compress, lib.bsc_block_info, lib.bsc_decompress):
        f.restype = C.c_int
    return lib


def ours():
    lib = _protos(b200lc.lib())
    lib.b200lc_bsc_set_stages.argtypes = [C.c_void_p]
    lib.b200lc_bsc_set_stages.restype = None
    return lib


def ref():
    return _protos(O.ref_bsc())


def ref_stages():
    r = O.ref_bsc()
    addr = lambda name: C.cast(getattr(r, name), C.c_void_p).value
    return Stages(addr("bsc_coder_compress"), addr("bsc_coder_decompress"), addr("bsc_lzp_compress"),
                  addr("bsc_lzp_decompress"), addr("bsc_bwt_decode"), addr("bsc_st_decode"))


def _store(lib, data):
    out = np.zeros(data.size + HEADER, np.uint8)
    n = lib.bsc_store(data.ctypes.data, out.ctypes.data, data.size, 0)
    assert n == data.size + HEADER
    return out


def _info(lib, block, size=None):
    bs, ds = C.c_int(-1), C.c_int(-1)
    rc = lib.bsc_block_info(block.ctypes.data, block.size if size is None else size, C.byref(bs), C.byref(ds), 0)
    return rc, bs.value, ds.value


def _decompress(lib, block, n):
    out = np.zeros(max(1, n), np.uint8)
    rc = lib.bsc_decompress(block.ctypes.data, block.size, out.ctypes.data, n, 0)
    return rc, out[:n]


@pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
@pytest.mark.parametrize("n", [0, 1, 27, 28, 29, 5553, 100000])
def test_stored_block_equals_reference(n):
    data = np.random.default_rng(n).integers(0, 256, n, dtype=np.uint8)
    a, b = _store(ours(), data), _store(ref(), data)
    assert np.array_equal(a, b)
    assert _info(ours(), a) == _info(ref(), a) == (0, n + HEADER, n)
    for lib in (ours(), ref()):
        rc, out = _decompress(lib, a, n)
        assert rc == 0 and np.array_equal(out, data)


def test_init_and_allocator_rules():
    lib = ours()
    assert lib.bsc_init(0) == 0 and lib.bsc_init(8 | 2 | 1) == 0
    lib.bsc_init_full.restype = C.c_int
    lib.bsc_init_full.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    libc = C.CDLL(None)
    m = C.cast(libc.malloc, C.c_void_p).value
    assert lib.bsc_init_full(0, m, None, None) == BAD_PARAMETER      # all three or none
    assert lib.bsc_init_full(0, None, None, None) == 0


@pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
def test_header_validation_matches_reference():
    data = np.frombuffer(synthetic_largefile(4000, seed=3), np.uint8)
    good = _store(ours(), data)
    cases = []
    for off, val in [(0, 5), (4, 1), (8, 0x21), (8, 2), (8, 0x10 << 16 | 3 << 8 | 0x21), (12, 99999), (24, 0), (0, 1 << 20)]:
        b = good.copy()
        b[off:off + 4] = np.frombuffer(np.int32(val).tobytes(), np.uint8)
        if off != 24:      # keep the header checksum valid so that the field checks are reached
            import zlib
            b[24:28] = np.frombuffer(np.uint32(zlib.adler32(b[:24].tobytes())).tobytes(), np.uint8)
        cases.append(b)
    for b in cases:
        rc = _info(ours(), b)[0]
        assert rc == _info(ref(), b)[0]
        # (a header that passes but announces a coded payload makes the reference decode garbage)
        if rc != 0:
            assert _decompress(ours(), b, data.size)[0] == _decompress(ref(), b, data.size)[0] == rc
    assert _info(ours(), good, 27)[0] == _info(ref(), good, 27)[0] == UNEXPECTED_EOB
    short = good[:-1].copy()
    assert _decompress(ours(), short, data.size)[0] == _decompress(ref(), short, data.size)[0] == UNEXPECTED_EOB
    bad = good.copy()
    bad[100] ^= 1
    assert _decompress(ours(), bad, data.size)[0] == _decompress(ref(), bad, data.size)[0] == DATA_CORRUPT


@pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
def test_compressed_block_without_stages_is_not_supported_and_with_reference_stages_decodes():
    data = np.frombuffer(synthetic_largefile(200000, seed=5), np.uint8)
    block = np.zeros(data.size + HEADER, np.uint8)
    n = ref().bsc_compress(data.ctypes.data, block.ctypes.data, data.size, 16, 128, 1, 1, 0)
    assert HEADER < n < data.size
    block = block[:n].copy()
    lib = ours()
    lib.b200lc_bsc_set_stages(None)
    assert _info(lib, block) == (0, n, data.size)
    assert _decompress(lib, block, data.size)[0] == NOT_SUPPORTED
    st = ref_stages()
    lib.b200lc_bsc_set_stages(C.byref(st))
    try:
        rc, out = _decompress(lib, block, data.size)
        assert rc == 0 and np.array_equal(out, data)
        # in place, as bsc.cpp:595 calls it
        buf = np.zeros(data.size + HEADER, np.uint8)
        buf[:n] = block
        assert lib.bsc_decompress(buf.ctypes.data, n, buf.ctypes.data, data.size, 0) == 0
        assert np.array_equal(buf[:data.size], data)
    finally:
        lib.b200lc_bsc_set_stages(None)


def test_compress_argument_errors_before_any_cuda_call():
    lib = ours()
    d = np.zeros(100, np.uint8)
    o = np.zeros(200, np.uint8)
    assert lib.bsc_compress(d.ctypes.data, o.ctypes.data, 100, 0, 0, 3, 1, 0) == BAD_PARAMETER     # ST3 / ST4: CPU-only, not built
    assert lib.bsc_compress(d.ctypes.data, o.ctypes.data, 100, 0, 0, 9, 1, 0) == BAD_PARAMETER
    assert lib.bsc_compress(d.ctypes.data, o.ctypes.data, 100, 0, 0, 1, 3, 0) == BAD_PARAMETER     # unknown coder
    assert lib.bsc_compress(d.ctypes.data, o.ctypes.data, 100, 16, 3, 1, 1, 0) == BAD_PARAMETER    # lzpMinLen < 4
    assert lib.bsc_compress(d.ctypes.data, o.ctypes.data, 100, 9, 128, 1, 1, 0) == BAD_PARAMETER   # lzpHashSize < 10
    assert lib.bsc_compress(d.ctypes.data, o.ctypes.data, -1, 0, 0, 1, 1, 0) == BAD_PARAMETER
    assert lib.bsc_compress(None, o.ctypes.data, 100, 0, 0, 1, 1, 0) == BAD_PARAMETER
    # no coder registered: stored block, no GPU needed
    lib.b200lc_bsc_set_stages(None)
    assert lib.bsc_compress(d.ctypes.data, o.ctypes.data, 100, 0, 0, 1, 1, 0) == 128
    assert _info(lib, o[:128].copy()) == (0, 128, 100)


@pytest.mark.gpu
@pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
@pytest.mark.parametrize("n,lzp_hash,lzp_min,coder", [(3569598, 16, 128, 1), (1 << 20, 0, 0, 2), (70000, 15, 32, 1),
                                                      (65535, 0, 0, 1), (40, 16, 128, 1)])
def test_bsc_compress_on_gpu_bwt_equals_reference_block(n, lzp_hash, lzp_min, coder):
    data = np.frombuffer(synthetic_largefile(n, seed=n % 89), np.uint8)
    want = np.zeros(n + HEADER, np.uint8)
    wn = ref().bsc_compress(data.ctypes.data, want.ctypes.data, n, lzp_hash, lzp_min, 1, coder, 0)
    lib = ours()
    st = ref_stages()
    lib.b200lc_bsc_set_stages(C.byref(st))
    try:
        got = np.zeros(n + HEADER, np.uint8)
        gn = lib.bsc_compress(data.ctypes.data, got.ctypes.data, n, lzp_hash, lzp_min, 1, coder, 0)
        assert gn == wn and np.array_equal(got[:gn], want[:wn])
        # in place (bsc.cpp:363 compresses the block inside its own buffer)
        buf = np.zeros(n + HEADER, np.uint8)
        buf[:n] = data
        assert lib.bsc_compress(buf.ctypes.data, buf.ctypes.data, n, lzp_hash, lzp_min, 1, coder, 0) == wn
        assert np.array_equal(buf[:wn], want[:wn])
        rc, out = _decompress(lib, got[:gn].copy(), n)
        assert rc == 0 and np.array_equal(out, data)
        rc, out = _decompress(ref(), got[:gn].copy(), n)
        assert rc == 0 and np.array_equal(out, data)
        # without a registered bwt_decode the inverse BWT runs on the GPU (bsc_bwt_decode of the library)
        st2 = ref_stages()
        st2.bwt_decode = None
        lib.b200lc_bsc_set_stages(C.byref(st2))
        rc, out = _decompress(lib, got[:gn].copy(), n)
        assert rc == 0 and np.array_equal(out, data)
    finally:
        lib.b200lc_bsc_set_stages(None)


@pytest.mark.gpu
@pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
@pytest.mark.parametrize("n,lzp_hash,lzp_min,coder,sorter", [(3569598, 16, 128, 1, 5), (1 << 20, 0, 0, 2, 6),
                                                             (70000, 15, 32, 1, 7), (1 << 20, 16, 128, 1, 8),
                                                             (40, 0, 0, 1, 6)])
def test_bsc_compress_with_sort_transform_blocks(n, lzp_hash, lzp_min, coder, sorter):
    """blockSorter = LIBBSC_BLOCKSORTER_ST5..ST8 (libbsc.h:70-73): ST on the GPU (bsc_st_encode_cuda).
    ST5 / ST6 blocks equal the reference library's (its CPU transform); ST7 / ST8 have no CPU encoder
    in the reference (st.cpp:1026) -- the reference library decompresses them back to the input."""
    data = np.frombuffer(synthetic_largefile(n, seed=n % 83 + sorter), np.uint8)
    lib = ours()
    st = ref_stages()
    lib.b200lc_bsc_set_stages(C.byref(st))
    try:
        got = np.zeros(n + HEADER, np.uint8)
        gn = lib.bsc_compress(data.ctypes.data, got.ctypes.data, n, lzp_hash, lzp_min, sorter, coder, 0)
        assert gn > 0
        if sorter <= 6:
            want = np.zeros(n + HEADER, np.uint8)
            wn = ref().bsc_compress(data.ctypes.data, want.ctypes.data, n, lzp_hash, lzp_min, sorter, coder, 0)
            assert gn == wn and np.array_equal(got[:gn], want[:wn])
        if n > 64:
            assert got[8] & 0x1f == sorter                      # mode word: the block sorter
        rc, out = _decompress(lib, got[:gn].copy(), n)
        assert rc == 0 and np.array_equal(out, data)
        rc, out = _decompress(ref(), got[:gn].copy(), n)
        assert rc == 0 and np.array_equal(out, data)
        assert lib.bsc_compress(data.ctypes.data, got.ctypes.data, n, lzp_hash, lzp_min, 3, coder, 0) == BAD_PARAMETER
        assert lib.bsc_compress(data.ctypes.data, got.ctypes.data, n, lzp_hash, lzp_min, 4, coder, 0) == BAD_PARAMETER
    finally:
        lib.b200lc_bsc_set_stages(None)


@pytest.mark.gpu
def test_bsc_compress_incompressible_block_is_stored():
    data = np.random.default_rng(1).integers(0, 256, 300000, dtype=np.uint8)
    lib = ours()
    if O.have_ref("bsc"):
        st = ref_stages()
        lib.b200lc_bsc_set_stages(C.byref(st))
    try:
        out = np.zeros(data.size + HEADER, np.uint8)
        n = lib.bsc_compress(data.ctypes.data, out.ctypes.data, data.size, 0, 0, 1, 1, 0)
        assert n == data.size + HEADER and np.array_equal(out[HEADER:], data)
        assert np.array_equal(out, _store(lib, data))
    finally:
        lib.b200lc_bsc_set_stages(None)
