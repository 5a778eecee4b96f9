"""GPU parity: b200lc_cuhd_decode vs the CPU oracle (oracle/cuhd_oracle.c) on the same streams.

The reference's own check is demo.cc:176-178 (decode on GPU, compare with the original bytes);
here every case is additionally compared with the oracle's serial LUT decode of the same stream.
Bar: bit-exact.
"""
import numpy as np
import pytest
import torch

import oracle_lib as O
from pkg import b200lc

pytestmark = pytest.mark.gpu


def _run(data, max_len=11, use_ref=True, drop_pad=False, misalign=0):
    code, length, lut, units = O.cuhd_make_case(data, max_len, use_ref)
    if drop_pad:
        units = units[:-1]
    expect, got = O.cuhd_oracle_decode(units, lut, data.size, max_len)
    assert got == data.size
    assert np.array_equal(expect, data)
    dev = torch.device("cuda:0")
    if misalign:
        buf = torch.zeros(units.size + 4, dtype=torch.int32, device=dev)
        d_units = buf[misalign:misalign + units.size]
        d_units.copy_(torch.from_numpy(units.view(np.int32)))
    else:
        d_units = torch.from_numpy(units.view(np.int32)).to(dev)
    d_lut = torch.from_numpy(np.ascontiguousarray(lut)).to(dev)
    out = b200lc.cuhd_decode(d_units, data.size, d_lut, max_len)
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    if not np.array_equal(res, expect):
        bad = np.flatnonzero(res != expect)
        raise AssertionError("mismatch at %d positions, first %s" % (bad.size, bad[:8]))


@pytest.mark.parametrize("n", [2, 3, 17, 100, 4097, 65536, (1 << 20) + 5, 1 << 24])
def test_zipf_sizes(n):
    _run(O.zipf_bytes(n, 1.1, seed=n))


@pytest.mark.parametrize("alpha", [0.0, 0.5, 2.0, 4.0])
def test_zipf_skew(alpha):
    _run(O.zipf_bytes(1 << 20, alpha, seed=7))


def test_uniform_fixed_8bit_codes():
    rng = np.random.default_rng(1)
    _run(rng.integers(0, 256, 1 << 20, dtype=np.uint8), use_ref=False)


def test_two_symbols_one_bit_codes():
    # 1-bit codes: 128 symbols per 16-byte subsequence -> staging window overflow path
    rng = np.random.default_rng(2)
    _run(rng.integers(0, 2, 1 << 20, dtype=np.uint8), use_ref=False)


def test_single_symbol():
    _run(np.full(100000, 65, np.uint8), use_ref=False)


def test_fixed_3bit_codes_never_resynchronise():
    # 8 equiprobable symbols -> all codes 3 bits; 128 % 3 != 0 so entry states rotate and no
    # speculative path ever merges: exercises the sequential slow paths.
    rng = np.random.default_rng(3)
    _run(rng.integers(0, 8, 20000, dtype=np.uint8), use_ref=False)


def test_runs_of_rare_symbols():
    rng = np.random.default_rng(4)
    d = O.zipf_bytes(1 << 18, 1.5, seed=5)
    d[1000:9000] = 255           # long codes back to back
    d[50000:50100] = rng.integers(200, 256, 100, dtype=np.uint8)
    _run(d)


@pytest.mark.parametrize("max_len", [9, 12, 13])
def test_other_table_widths(max_len):
    _run(O.zipf_bytes(300000, 1.1, seed=max_len), max_len=max_len, use_ref=False)


def test_without_pad_unit_and_unaligned_base():
    d = O.zipf_bytes(500000, 1.1, seed=11)
    _run(d, drop_pad=True)
    _run(d, misalign=1)
    _run(d, misalign=3, drop_pad=True)


def test_reference_demo_program_unmodified_against_compat_headers(tmp_path):
    """oracle/_ref/cuhd_demo_b200 = cuhd-icpp/src/demo.cc, unmodified, compiled against
    include/cuhd_compat/{cuhd,llhuff}.h and linked with libb200lc.so (oracle/Makefile).  The demo
    encodes a file, decodes it on the GPU and prints "mismatch" if the bytes differ
    (demo.cc:176-178)."""
    import os
    import subprocess
    exe = os.path.join(O.ORACLE_DIR, "_ref", "cuhd_demo_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/cuhd_demo_b200 not built")
    data = O.zipf_bytes(3_000_001, 1.1, seed=99)
    src, dst = tmp_path / "in.bin", tmp_path / "out.huf"
    src.write_bytes(data.tobytes())
    r = subprocess.run([exe, "0", str(src), str(dst)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatch" not in r.stdout and "decoding.." in r.stdout
    # the file the demo wrote is the packed stream: decode it with the oracle
    hist = np.bincount(data, minlength=256)
    code, length, lut = b200lc.cuhd_build_table(hist)
    units = np.frombuffer(dst.read_bytes(), np.uint32)
    want, _ = O.cuhd_oracle_encode(data, code, length)
    assert np.array_equal(units, want)
