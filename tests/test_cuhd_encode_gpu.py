"""GPU parity: b200lc_cuhd_encode / histogram / table builder vs the reference encoder and oracle.

Bit-exactness bar: with the REFERENCE's dictionary as input, every unit the reference defines
(all full units) is identical, and the last partial unit equals the
oracle's zero-filled flush.  Round trip through b200lc_cuhd_decode is byte-identical.
"""
import numpy as np
import pytest
import torch

import oracle_lib as O
from pkg import b200lc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _encode_gpu(data, code, length, **kw):
    d = torch.from_numpy(data).to(DEV)
    c = torch.from_numpy(code.view(np.int32)).to(DEV)
    l = torch.from_numpy(length).to(DEV)
    enc = b200lc.cuhd_encode(d, c, l, **kw)
    torch.cuda.synchronize()
    return enc


@pytest.mark.parametrize("n", [2, 3, 100, 8191, 8192, 8193, 100000, (1 << 22) + 77])
def test_encode_matches_reference_and_oracle(n):
    data = O.zipf_bytes(n, 1.1, seed=n)
    if O.have_ref("cuhd"):
        code, length, lut, ref_units = O.cuhd_ref_encode(data)
    else:
        code, length, lut, _ = O.cuhd_make_case(data, use_ref=False)
        ref_units = None
    want, defined = O.cuhd_oracle_encode(data, code, length)
    enc = _encode_gpu(data, code, length)
    got = enc.units.cpu().numpy().view(np.uint32)
    assert enc.n_units == want.size
    assert np.array_equal(got[: want.size], want)
    assert got[want.size] == 0          # pad unit (cuhd_input_buffer.cc:20-27)
    if ref_units is not None:
        assert ref_units.size == want.size
        assert np.array_equal(got[:defined], ref_units[:defined])


@pytest.mark.parametrize("kind", ["uniform", "two", "single", "skew4", "long_codes"])
def test_encode_roundtrip_distributions(kind):
    rng = np.random.default_rng(5)
    n = 3_000_001
    if kind == "uniform":
        data = rng.integers(0, 256, n, dtype=np.uint8)
    elif kind == "two":
        data = rng.integers(0, 2, n, dtype=np.uint8)
    elif kind == "single":
        data = np.full(n, 7, np.uint8)
    elif kind == "skew4":
        data = O.zipf_bytes(n, 4.0, seed=1)
    else:
        data = O.zipf_bytes(n, 1.1, seed=2)
        data[100000:400000] = 255
    d = torch.from_numpy(data).to(DEV)
    hist = b200lc.histogram_u8(d)
    torch.cuda.synchronize()
    h = hist.cpu().numpy()
    assert np.array_equal(h, np.bincount(data, minlength=256))
    code, length, lut = b200lc.cuhd_build_table(h)
    want, _ = O.cuhd_oracle_encode(data, code, length)
    enc = _encode_gpu(data, code, length)
    got = enc.units.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[: want.size], want)
    out = b200lc.cuhd_decode(enc.units, n, torch.from_numpy(lut).to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(out, d)


def test_encode_capacity_overflow_is_reported():
    data = O.zipf_bytes(200000, 1.1, seed=3)
    code, length, lut, _ = O.cuhd_make_case(data)
    with pytest.raises(b200lc.B200LCError):
        _encode_gpu(data, code, length, units_cap=1000)


def test_encode_unaligned_input_base():
    data = O.zipf_bytes(100001, 1.1, seed=4)
    code, length, lut, _ = O.cuhd_make_case(data)
    want, _ = O.cuhd_oracle_encode(data, code, length)
    buf = torch.zeros(data.size + 16, dtype=torch.uint8, device=DEV)
    view = buf[3:3 + data.size]
    view.copy_(torch.from_numpy(data))
    c = torch.from_numpy(code.view(np.int32)).to(DEV)
    l = torch.from_numpy(length).to(DEV)
    enc = b200lc.cuhd_encode(view, c, l)
    got = enc.units.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[: want.size], want)


# ------------------------------------------------------------------------------------------ blocks
@pytest.mark.parametrize("n,block", [(1 << 22, 1 << 20), (1 << 21, 65536), (300000, 65536), (1000, 4096),
                                     (5 * 131072 + 77, 131072), (3 * 200000, 200000), (7 * 50001, 50001)])
def test_encode_blocks_equals_per_block_encode_and_round_trips_through_batch_decode(n, block):
    DEVB = "cuda:0"
    data = O.zipf_bytes(n, 1.1, seed=n % 997)
    d = torch.from_numpy(data).to(DEVB)
    hist = np.maximum(np.bincount(data, minlength=256), 1)
    code, length, lut = b200lc.cuhd_build_table(hist)
    d_code = torch.from_numpy(code.view(np.int32)).to(DEVB)
    d_len = torch.from_numpy(length).to(DEVB)
    units, bits, stride = b200lc.cuhd_encode_blocks(d, block, d_code, d_len)
    torch.cuda.synchronize()
    hb = bits.cpu().numpy()
    hu = units.cpu().numpy().view(np.uint32)
    nblocks = (n + block - 1) // block
    streams = np.zeros((nblocks, 4), np.uint64)
    for b in range(nblocks):
        part = data[b * block:(b + 1) * block]
        want, _ = O.cuhd_oracle_encode(part, code, length)
        wbits = int(length[part].astype(np.int64).sum())
        assert int(hb[b]) == wbits, b
        nu = (wbits + 31) // 32
        assert np.array_equal(hu[b * stride: b * stride + nu], want[:nu]), b
        assert hu[b * stride + nu] == 0                      # the reference's pad unit
        streams[b] = (b * stride, nu, b * block, part.size)
    out = torch.empty(n, dtype=torch.uint8, device=DEVB)
    b200lc.cuhd_decode_batch(units, out, streams, torch.from_numpy(lut).to(DEVB))
    assert np.array_equal(out.cpu().numpy(), data)


# ------------------------------------------------------------------------------------------ planned
@pytest.mark.parametrize("n", [2, 100, 8193, 131072, 131073, 5 * 131072, (1 << 22) + 77, (1 << 24) + 12345])
@pytest.mark.parametrize("alpha", [1.1, 3.0])
def test_planned_encode_is_identical_to_the_two_pass_encode(n, alpha):
    data = O.zipf_bytes(n, alpha, seed=n % 991)
    d = torch.from_numpy(data).to(DEV)
    hist, ph = b200lc.histogram_u8_pieces(d)
    assert np.array_equal(hist.cpu().numpy(), np.bincount(data, minlength=256))
    code, length, lut = b200lc.cuhd_build_table(hist.cpu().numpy())
    d_code = torch.from_numpy(code.view(np.int32)).to(DEV)
    d_len = torch.from_numpy(length).to(DEV)
    a = b200lc.cuhd_encode(d, d_code, d_len)
    b = b200lc.cuhd_encode(d, d_code, d_len, piece_hist=ph)
    assert a.bits == b.bits and a.n_units == b.n_units
    assert torch.equal(a.units[: a.n_units + 1], b.units[: b.n_units + 1])       # incl. the pad unit
    out = b200lc.cuhd_decode(b.units, n, torch.from_numpy(lut).to(DEV))
    assert np.array_equal(out.cpu().numpy(), data)


@pytest.mark.parametrize("n", [1, 131071, 131072, 3 * 131072 + 5, (1 << 23) + 4099])
def test_planned_encode_equals_oracle_and_reference_encoder(n):
    """The packer bench.py times, against the checker directly: oracle restatement of
    llhuffman_encoder.cc:200-238, and the reference encoder itself with its own dictionary."""
    data = O.zipf_bytes(n, 1.1, seed=n % 983)
    d = torch.from_numpy(data).to(DEV)
    _, ph = b200lc.histogram_u8_pieces(d)
    cases = []
    hist = np.bincount(data, minlength=256)
    code, length, _ = b200lc.cuhd_build_table(hist)
    want, defined = O.cuhd_oracle_encode(data, code, length)
    cases.append((code, length, want, defined))
    if O.have_ref("cuhd") and np.unique(data).size > 1:
        rcode, rlen, _, runits = O.cuhd_ref_encode(data)
        # the reference loses the codeword that straddles into its last unit: that
        # unit is not comparable
        cases.append((rcode, rlen, runits, runits.size - 1))
    for code, length, want, defined in cases:
        enc = b200lc.cuhd_encode(d, torch.from_numpy(code.view(np.int32)).to(DEV),
                                 torch.from_numpy(length).to(DEV), piece_hist=ph)
        bits = int(length[data].astype(np.int64).sum())
        assert enc.bits == bits and enc.n_units == want.size
        got = enc.units[: enc.n_units].cpu().numpy().view(np.uint32)
        assert np.array_equal(got[:defined], want[:defined])


def test_planned_encode_reports_overflow():
    n = 1 << 20
    data = O.zipf_bytes(n, 1.1, seed=5)
    d = torch.from_numpy(data).to(DEV)
    hist, ph = b200lc.histogram_u8_pieces(d)
    code, length, _ = b200lc.cuhd_build_table(hist.cpu().numpy())
    with pytest.raises(b200lc.B200LCError):
        b200lc.cuhd_encode(d, torch.from_numpy(code.view(np.int32)).to(DEV), torch.from_numpy(length).to(DEV),
                           units_cap=1000, piece_hist=ph)
