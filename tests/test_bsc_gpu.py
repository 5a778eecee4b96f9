"""GPU parity for libbsc's BWT stage: bsc_bwt_encode exported by
libb200lc.so against the oracle restatement (oracle/bsc_oracle.c), the reference's own divbwt
(oracle/_ref/libref_bsc.so) and, end to end, the reference's bsc program linked against the GPU
implementation (oracle/_ref/bsc_b200) against the all-CPU reference program (oracle/_ref/bsc).
Bar: bit-exact arrays, byte-identical .bsc files."""
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
