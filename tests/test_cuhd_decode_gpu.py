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


def test_table_width_order_13_11_13():
    # the dynamic shared-memory limit is a kernel attribute: a narrower table after a wider one
    # must not lower it (round-1 ADVICE: 13 -> 11 -> 13 failed with a cached occupancy)
    d = O.zipf_bytes(200000, 1.1, seed=21)
    for max_len in (13, 11, 13, 9, 13):
        _run(d, max_len=max_len, use_ref=False)


def test_zero_runs_with_two_bit_code_do_not_resynchronise():
    # symbol 0 dominates and gets a short all-zero codeword: long runs of it keep a wrong entry
    # state alive, so guesses fail and the repair rounds (segment + piece level) must run
    rng = np.random.default_rng(8)
    d = np.zeros(3 << 20, np.uint8)
    idx = rng.integers(0, d.size, d.size // 7)
    d[idx] = rng.integers(1, 7, idx.size, dtype=np.uint8)
    d[1 << 20:(1 << 20) + 400000] = 0
    _run(d, use_ref=False)


def test_fixed_3bit_codes_large_never_resynchronise():
    # as test_fixed_3bit_codes_never_resynchronise but over many pieces: every guess of a segment
    # or piece entry state is wrong two times out of three
    rng = np.random.default_rng(9)
    _run(rng.integers(0, 8, 3 << 20, dtype=np.uint8), use_ref=False)


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


def test_piece_ranges_equal_one_shot_decode():
    """b200lc_cuhd_decode_pieces over consecutive ranges == one b200lc_cuhd_decode call, and the
    progress word counts exactly the symbols of the finished pieces."""
    import ctypes as C
    n = 3 * (1 << 20) + 12345
    data = O.zipf_bytes(n, 1.1, seed=21)
    code, length, lut, units = O.cuhd_make_case(data, 11, True)
    dev = torch.device("cuda:0")
    L = b200lc.lib()
    L.b200lc_cuhd_decode_piece_units.restype = C.c_size_t
    vp, sz = C.c_void_p, C.c_size_t
    L.b200lc_cuhd_decode_pieces.argtypes = [vp, sz, vp, sz, vp, C.c_int, vp, sz, sz, sz, vp]
    L.b200lc_cuhd_decode_progress_async.argtypes = [vp, sz, vp, vp]
    pu = L.b200lc_cuhd_decode_piece_units()
    pieces = (units.size + pu - 1) // pu
    assert pieces >= 5
    d_units = torch.from_numpy(units.view(np.int32)).to(dev)
    d_lut = torch.from_numpy(np.ascontiguousarray(lut)).to(dev)
    out = torch.zeros(n, dtype=torch.uint8, device=dev)
    scr = torch.empty(L.b200lc_cuhd_decode_scratch_bytes(units.size) + 256, dtype=torch.uint8, device=dev)
    prog = torch.zeros(4, dtype=torch.int64).pin_memory()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    cuts = [0, 1, 3, pieces - 1, pieces]
    done = []
    for i in range(len(cuts) - 1):
        b200lc.check(L.b200lc_cuhd_decode_pieces(d_units.data_ptr(), units.size, out.data_ptr(), n,
                                                 d_lut.data_ptr(), 11, scr.data_ptr(), scr.numel() - 256,
                                                 cuts[i], cuts[i + 1], sp), "pieces")
        b200lc.check(L.b200lc_cuhd_decode_progress_async(scr.data_ptr(), cuts[i + 1],
                                                         prog.data_ptr() + 8 * i, sp), "progress")
        torch.cuda.synchronize()
        ready = min(int(prog[i]) & ((1 << 56) - 1), n)
        done.append(ready)
        assert np.array_equal(out[:ready].cpu().numpy(), data[:ready]), i
    assert done == sorted(done) and done[-1] == n and 0 < done[0] < done[1] < n
    assert np.array_equal(out.cpu().numpy(), data)


@pytest.mark.parametrize("n,pinned", [(1000, True), (5 * (1 << 20) + 7, False), (40 * (1 << 20) + 3, True)])
def test_session_round_trip_host_buffers(n, pinned):
    """b200lc_cuhd_session_encode / _decode (demo.cc:90-168 as two calls); 40 MiB spans several
    H2D chunks of the pipelined decode."""
    data = O.zipf_bytes(n, 1.1, seed=n & 0xff)
    mk = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    h_in = mk(torch.from_numpy(data.copy()))
    cap = (n * 11 + 31) // 32 + 2
    h_units = mk(torch.zeros(cap, dtype=torch.int32))
    h_out = mk(torch.zeros(n, dtype=torch.uint8))
    h_code = torch.zeros(256, dtype=torch.int32)
    h_len = torch.zeros(256, dtype=torch.uint8)
    h_lut = torch.zeros((1 << 11, 2), dtype=torch.uint8)
    sess = b200lc.CuhdSession(n)
    nu = sess.encode(h_in, h_units, h_code, h_len, h_lut, 11)
    # the session's stream is what the oracle's packer produces with the session's table
    want, _ = O.cuhd_oracle_encode(data, h_code.numpy().view(np.uint32), h_len.numpy())
    assert nu == want.size
    assert np.array_equal(h_units.numpy().view(np.uint32)[:nu], want)
    sess.decode(h_units, nu + 1, h_lut, h_out, 11)
    sess.decode(h_units, nu + 1, h_lut, h_out, 11)   # scratch reuse across calls
    sess.close()
    assert np.array_equal(h_out.numpy(), data)


# ------------------------------------------------------------------------------------------ batch
DEV = "cuda:0"


def _encode_blocks(data, sizes, align_units=4):
    """Blocks of `data` packed separately with ONE shared table; returns (units tensor, streams
    array, lut tensor).  Each stream starts at a multiple of align_units units."""
    d = torch.from_numpy(data).to(DEV)
    hist = b200lc.histogram_u8(d).cpu().numpy()
    hist = np.maximum(hist, 1)                      # every symbol gets a code: blocks share the table
    code, length, lut = b200lc.cuhd_build_table(hist)
    d_code = torch.from_numpy(code.view(np.int32)).to(DEV)
    d_len = torch.from_numpy(length).to(DEV)
    parts, streams, uoff, ooff = [], [], 0, 0
    for n in sizes:
        if n == 0:
            streams.append((uoff, 0, ooff, 0))
            continue
        enc = b200lc.cuhd_encode(d[ooff:ooff + n], d_code, d_len)
        nu = enc.n_units
        parts.append(enc.units[:nu])
        pad = (-nu) % align_units if align_units > 1 else 0
        if pad:
            parts.append(torch.zeros(pad, dtype=torch.int32, device=DEV))
        streams.append((uoff, nu, ooff, n))
        uoff += nu + pad
        ooff += n
    units = torch.cat(parts + [torch.zeros(8, dtype=torch.int32, device=DEV)])
    return units, np.array(streams, np.uint64), torch.from_numpy(lut).to(DEV), code, length, lut


@pytest.mark.parametrize("sizes,align", [
    ([1 << 20] * 8, 4),                                  # equal blocks (config 5 shape)
    ([65536] * 64, 4),
    ([1, 2, 3, 100, 4097, 0, 70001, 1 << 20, 5, 300000], 4),   # ragged, an empty stream in the middle
    ([12345, 54321, 99999, 1 << 18], 1),                 # streams at arbitrary unit offsets (no TMA)
    ([3 << 20, 1000, 1 << 21], 4),                       # long + short mixed: look-back stops per stream
])
def test_batch_decode_equals_per_stream_oracle(sizes, align):
    total = sum(sizes)
    data = O.zipf_bytes(total, 1.1, seed=total % 1000)
    units, streams, d_lut, code, length, lut = _encode_blocks(data, sizes, align)
    out = torch.full((total + 64,), 0xEE, dtype=torch.uint8, device=DEV)
    b200lc.cuhd_decode_batch(units, out, streams, d_lut)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got[:total], data)
    assert (got[total:] == 0xEE).all()                   # nothing written past the last stream
    # oracle decode of two of the streams from their own units
    hu = units.cpu().numpy().view(np.uint32)
    for (uo, nu, oo, n) in streams[[0, len(streams) - 1]]:
        if n:
            exp, cnt = O.cuhd_oracle_decode(hu[int(uo):int(uo + nu) + 1].copy(), lut, int(n))
            assert cnt == n and np.array_equal(exp, data[int(oo):int(oo + n)])


def test_batch_decode_many_small_streams_one_launch():
    # 2048 streams of 16 KiB: far more streams than CTAs, one piece each
    sizes = [16384] * 2048
    total = sum(sizes)
    data = O.zipf_bytes(total, 1.3, seed=11)
    units, streams, d_lut, *_ = _encode_blocks(data, sizes, 4)
    out = torch.empty(total, dtype=torch.uint8, device=DEV)
    b200lc.cuhd_decode_batch(units, out, streams, d_lut)
    assert np.array_equal(out.cpu().numpy(), data)
