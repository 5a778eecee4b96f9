"""GPU parity for libbsc's BWT stage: bsc_bwt_encode exported by
libb200lc.so against the oracle restatement (oracle/bsc_oracle.c), the reference's own divbwt
(oracle/_ref/libref_bsc.so) and, end to end, the reference's bsc program linked against the GPU
implementation (oracle/_ref/bsc_b200) against the all-CPU reference program (oracle/_ref/bsc).
Bar: bit-exact arrays, byte-identical .bsc files."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from pkg import b200lc
from test_oracle_bsc import _cases
from test_ref_bsc_cpu import synthetic_largefile

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(O.ORACLE_DIR, "_ref")


@pytest.mark.parametrize("name", list(_cases().keys()))
def test_bwt_encode_matches_oracle_and_reference(name):
    data = _cases()[name]
    gu, gp, gi = b200lc.bsc_bwt_encode(data)
    ou, op, oi = O.bsc_oracle_bwt_encode(data)
    assert gp == op and np.array_equal(gu, ou) and np.array_equal(gi, oi)
    if O.have_ref("bsc"):
        ru, rp, ri = O.bsc_ref_bwt_encode(data)
        assert gp == rp and np.array_equal(gu, ru) and np.array_equal(gi, ri)


def test_bwt_encode_trivial_sizes():
    lib = b200lc.lib()
    assert lib.bsc_bwt_encode(None, 5, None, None, 0) == -1          # LIBBSC_BAD_PARAMETER
    one = np.array([9], np.uint8)
    assert lib.bsc_bwt_encode(one.ctypes.data, 1, None, None, 0) == 1 and one[0] == 9
    assert lib.bsc_bwt_encode(one.ctypes.data, 0, None, None, 0) == 0


@pytest.mark.skipif(not O.have_ref("bsc"), reason="oracle/_ref/libref_bsc.so not built")
def test_bwt_encode_default_block_size_25mb():
    # bsc's default block (-b25): 25 MiB of word-structured text and of quantisation codes
    n = 25 << 20
    text = np.frombuffer(synthetic_largefile(n, seed=11), np.uint8)
    for data in (text, O.quant_codes(n)):
        gu, gp, gi = b200lc.bsc_bwt_encode(data)
        ru, rp, ri = O.bsc_ref_bwt_encode(data)
        assert gp == rp and np.array_equal(gi, ri)
        assert np.array_equal(gu, ru)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "bsc_b200")), reason="oracle/_ref/bsc_b200 not built")
@pytest.mark.parametrize("n,args,exact", [(3569598, [], True), (60 << 20, ["-t"], True), (60 << 20, [], False),
                                          (8 << 20, ["-b1", "-t"], True), (8 << 20, ["-b1"], False),
                                          (5 << 20, ["-b2", "-p", "-t"], True)])
def test_reference_bsc_program_on_gpu_bwt_is_byte_identical(tmp_path, n, args, exact):
    # Without -t the reference compresses blocks from OpenMP threads (bsc.cpp:206) and appends
    # them to the file in completion order, so two runs of the SAME binary may order the blocks
    # differently: those cases exercise the serialised GPU work area and are checked by size and
    # round trip; with -t (blocks one after another) the files must be byte-identical.
    # -b1: many small blocks; -p: no LZP stage in front of the BWT.
    data = synthetic_largefile(n, seed=n % 97)
    src = tmp_path / "in"
    src.write_bytes(data)
    outs = {}
    for exe in ("bsc", "bsc_b200"):
        comp = tmp_path / (exe + ".bsc")
        r = subprocess.run([os.path.join(REF_DIR, exe), "e", str(src), str(comp)] + args, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[exe] = comp.read_bytes()
    assert len(outs["bsc"]) == len(outs["bsc_b200"])
    if exact:
        assert outs["bsc"] == outs["bsc_b200"]
    back = tmp_path / "back"
    r = subprocess.run([os.path.join(REF_DIR, "bsc"), "d", str(tmp_path / "bsc_b200.bsc"), str(back)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and back.read_bytes() == data
    # and the other way round: the program on the library (bsc_bwt_decode on the GPU) decodes the
    # all-CPU program's file
    back2 = tmp_path / "back2"
    r = subprocess.run([os.path.join(REF_DIR, "bsc_b200"), "d", str(tmp_path / "bsc.bsc"), str(back2)] +
                       [a for a in args if a == "-t"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and back2.read_bytes() == data, r.stdout + r.stderr


# ------------------------------------------------------------------------------ Sort Transform ST5-8
def _st_gpu(data, k):
    lib = b200lc.lib()
    lib.bsc_st_encode_cuda.restype = C.c_int
    lib.bsc_st_encode_cuda.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    t = np.ascontiguousarray(data).copy()
    idx = lib.bsc_st_encode_cuda(t.ctypes.data if t.size else None, t.size, k, 0)
    return t, idx


def _st_cases():
    rng = np.random.default_rng(3)
    return {
        "text": np.frombuffer((b"the quick brown fox jumps over the lazy dog. " * 3000)[:100000], np.uint8).copy(),
        "rand4": rng.integers(0, 4, 50000, dtype=np.uint8),
        "rand": rng.integers(0, 256, 70001, dtype=np.uint8),
        "zeros": np.zeros(4097, np.uint8),
        "quant": O.quant_codes(1 << 18),
        "tiny2": np.array([5, 5], np.uint8),
        "tiny3": np.array([3, 1, 2], np.uint8),
        "tiny9": rng.integers(0, 3, 9, dtype=np.uint8),
        "period7": (np.arange(7000) % 7).astype(np.uint8),
    }


@pytest.mark.parametrize("k", [5, 6, 7, 8])
@pytest.mark.parametrize("name", list(_st_cases().keys()))
def test_st_encode_matches_oracle_and_reference(name, k):
    """bsc_st_encode_cuda (st.cuh:64-72) == the oracle's restatement of the transform; k = 5, 6 also ==
    the reference's CPU bsc_st_encode; every k inverted by the reference's CPU bsc_st_decode."""
    data = _st_cases()[name]
    got, gi = _st_gpu(data, k)
    want, wi = O.bsc_oracle_st_encode(data, k)
    assert gi == wi and np.array_equal(got, want)
    if O.have_ref("bsc"):
        if k <= 6:
            ref, ri = O.bsc_ref_st_encode(data, k)
            assert gi == ri and np.array_equal(got, ref)
        rc, back = O.bsc_ref_st_decode(got, k, gi)
        assert rc == 0 and np.array_equal(back, data)


def test_st_encode_arguments_and_default_block():
    lib = b200lc.lib()
    lib.bsc_st_encode_cuda.restype = C.c_int
    lib.bsc_st_encode_cuda.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    one = np.array([7], np.uint8)
    assert lib.bsc_st_encode_cuda(None, 5, 5, 0) == -1                  # LIBBSC_BAD_PARAMETER
    assert lib.bsc_st_encode_cuda(one.ctypes.data, 1, 4, 0) == -1       # k outside 5..8 (st2.cu:372)
    assert lib.bsc_st_encode_cuda(one.ctypes.data, 1, 9, 0) == -1
    assert lib.bsc_st_encode_cuda(one.ctypes.data, 1, 5, 0) == 0 and one[0] == 7
    assert lib.bsc_st_cuda_init(0) == 0
    if not O.have_ref("bsc"):
        return
    # bsc's default block size, 25 MiB: ST6 against the reference CPU encoder, ST8 through its decoder
    n = 25 << 20
    data = np.frombuffer(synthetic_largefile(n, seed=5), np.uint8)
    got, gi = _st_gpu(data, 6)
    ref, ri = O.bsc_ref_st_encode(data, 6)
    assert gi == ri and np.array_equal(got, ref)
    got8, gi8 = _st_gpu(data, 8)
    rc, back = O.bsc_ref_st_decode(got8, 8, gi8)
    assert rc == 0 and np.array_equal(back, data)
    lib.b200lc_bsc_st_release()


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DIR, "bsc_b200")), reason="oracle/_ref/bsc_b200 not built")
@pytest.mark.parametrize("k", [5, 6, 7, 8])
def test_reference_bsc_program_with_sort_transform_on_the_gpu(tmp_path, k):
    """The reference's bsc program built with its own LIBBSC_CUDA_SUPPORT flag and linked against
    libb200lc.so instead of st2.cu + b40c: `bsc e -m<k> -G` runs ST-k through bsc_st_encode_cuda.
    ST5 / ST6: the file equals the all-CPU reference program's (same transform on its CPU path);
    ST7 / ST8 exist on the GPU only (st.cpp:1016,1026: the CPU program refuses them) -- the all-CPU
    reference program decodes the file back to the input."""
    n = 6 << 20
    data = synthetic_largefile(n, seed=k)
    src = tmp_path / "in"
    src.write_bytes(data)
    gpu = tmp_path / "gpu.bsc"
    r = subprocess.run([os.path.join(REF_DIR, "bsc_b200"), "e", str(src), str(gpu), "-m%d" % k, "-G", "-t", "-b2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and gpu.exists(), r.stdout + r.stderr
    assert "error" not in (r.stdout + r.stderr).lower(), r.stdout + r.stderr
    if k <= 6:
        cpu = tmp_path / "cpu.bsc"
        r = subprocess.run([os.path.join(REF_DIR, "bsc"), "e", str(src), str(cpu), "-m%d" % k, "-t", "-b2"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert cpu.read_bytes() == gpu.read_bytes()
    back = tmp_path / "back"
    r = subprocess.run([os.path.join(REF_DIR, "bsc"), "d", str(gpu), str(back)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and back.read_bytes() == data, r.stdout + r.stderr


# ------------------------------------------------------------------------------ inverse BWT
@pytest.mark.parametrize("name", list(_cases().keys()))
def test_bwt_decode_inverts_encode_and_matches_reference(name):
    """bsc_bwt_decode (bwt.h:61) on the GPU: inverts the GPU encoder's and the reference encoder's
    output, and equals the reference's CPU bsc_bwt_decode."""
    data = _cases()[name]
    gu, gp, _ = b200lc.bsc_bwt_encode(data)
    assert np.array_equal(b200lc.bsc_bwt_decode(gu, gp), data)
    if O.have_ref("bsc"):
        ru, rp, ri = O.bsc_ref_bwt_encode(data)
        assert np.array_equal(b200lc.bsc_bwt_decode(ru, rp), data)
        t = ru.copy()
        idx = np.zeros(256, np.int32)
        assert O.ref_bsc().bsc_bwt_decode(t, t.size, rp, 0, idx, 0) == 0 and np.array_equal(t, data)


def test_bwt_decode_arguments_and_blocks_beyond_24_bit_rows():
    lib = b200lc.lib()
    one = np.array([9, 8, 7], np.uint8)
    assert lib.bsc_bwt_decode(None, 3, 1, 0, None, 0) == -1            # LIBBSC_BAD_PARAMETER (bwt.cpp:361)
    assert lib.bsc_bwt_decode(one.ctypes.data, 3, 0, 0, None, 0) == -1
    assert lib.bsc_bwt_decode(one.ctypes.data, 3, 4, 0, None, 0) == -1
    assert lib.bsc_bwt_decode(one.ctypes.data, 1, 1, 0, None, 0) == 0 and one[0] == 9
    # bsc's default block (25 MiB) and one just past 2^24 rows: the 64-bit row entries
    for n, seed in ((25 << 20, 11), ((1 << 24) + 12345, 12)):
        data = np.frombuffer(synthetic_largefile(n, seed=seed), np.uint8)
        gu, gp, _ = b200lc.bsc_bwt_encode(data)
        assert np.array_equal(b200lc.bsc_bwt_decode(gu, gp), data)
    periodic = np.tile(np.frombuffer(b"abracadabra-", np.uint8), (1 << 21) // 12 + 1)[: 1 << 21]
    gu, gp, _ = b200lc.bsc_bwt_encode(periodic)
    assert np.array_equal(b200lc.bsc_bwt_decode(gu, gp), periodic)
