"""Parity at BASELINE.json's full sizes through size-independent properties: encode -> decode
round trips byte-identical, exact bit/byte accounting against the histogram, and oracle spot
checks on the first / last units of the big buffers (the oracle is too slow for the whole of
them).  Inputs are the synthetic generators of bench.py / tools/bench_paths.py."""
import os
import sys

import numpy as np
import pytest
import torch

import oracle_lib as O
from pkg import b200lc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MIB = 1 << 20


def test_c2_cuhd_one_gib_zipf_round_trip():
    import bench as B
    n = 1 << 30
    data = B.gen_zipf_gpu(n, torch.device(DEV), B.SEED)
    hist = b200lc.histogram_u8(data).cpu().numpy()
    assert int(hist.sum()) == n
    code, length, lut = b200lc.cuhd_build_table(hist)
    enc = b200lc.cuhd_encode(data, torch.from_numpy(code.view(np.int32)).to(DEV), torch.from_numpy(length).to(DEV))
    # exact accounting: stream bits = sum over symbols of count * code length
    assert enc.bits == int((hist.astype(np.int64) * length.astype(np.int64)).sum())
    out = b200lc.cuhd_decode(enc.units, n, torch.from_numpy(lut).to(DEV))
    assert torch.equal(out, data)
    # the stream is a concatenation of codewords: its first units equal the oracle's encoding of
    # the first MiB of symbols (all but the last, possibly shared, unit)
    head = data[:MIB].cpu().numpy()
    want, _ = O.cuhd_oracle_encode(head, code, length)
    got = enc.units[: want.size].cpu().numpy().view(np.uint32)
    assert np.array_equal(got[: want.size - 1], want[: want.size - 1])


def test_c3_culzss_four_gib_quant_codes_round_trip():
    from bench_paths import quant_codes_gpu
    n, buf = 4 << 30, MIB
    nbuf = n // buf
    data = quant_codes_gpu(n, torch.device(DEV), itemsize=4)
    out, clen = b200lc.culzss_encode(data, buf)
    cl = clen.cpu().numpy().astype(np.int64)
    assert (cl > 0).all() and (cl < buf).all()           # every buffer of this stream compresses
    stride = b200lc.culzss_out_stride(buf)
    # oracle spot checks: first, middle and last buffer byte for byte
    for b in (0, nbuf // 2, nbuf - 1):
        src = data[b * buf:(b + 1) * buf].cpu().numpy()
        ok, want = O.culzss_oracle_compress(src)
        got = out[b * stride: b * stride + int(cl[b])].cpu().numpy()
        assert ok and np.array_equal(got, want), b
    # pack back to back and decode everything
    offs = np.zeros(nbuf + 1, np.int64)
    offs[1:] = np.cumsum(cl)
    rows = out.view(nbuf, stride)
    comp = torch.empty(int(offs[-1]), dtype=torch.uint8, device=DEV)
    d_offs = torch.from_numpy(offs).to(DEV)
    # gather the ragged rows with one masked copy per 256 buffers (keeps the temporary small)
    col = torch.arange(stride, device=DEV)
    d_cl = torch.from_numpy(cl).to(DEV)
    for lo in range(0, nbuf, 256):
        hi = min(nbuf, lo + 256)
        mask = col[None, :] < d_cl[lo:hi, None]
        comp[int(offs[lo]):int(offs[hi])] = rows[lo:hi][mask]
    del out, rows
    dec = b200lc.culzss_decode(comp, d_offs, buf)
    assert torch.equal(dec, data)
    assert 2.0 < n / float(offs[-1]) < 3.0


def test_c4_cudpp_1024_blocks_round_trip():
    from bench_paths import cudpp_blocks_gpu
    n, nblocks, batch = MIB, 1024, 128          # 1024 blocks = one GPU's share of config 4
    L = b200lc.lib()
    scratch = torch.empty(max(L.b200lc_cudpp_compress_scratch_bytes(batch, n),
                              L.b200lc_cudpp_decompress_scratch_bytes(batch, n)) + 256,
                          dtype=torch.uint8, device=DEV)
    back = torch.empty(batch * n, dtype=torch.uint8, device=DEV)
    res = None
    total_words = 0
    for g, kind in enumerate(["zipf", "markov", "rand", "zipf", "markov", "rand", "zipf", "markov"]):
        data = cudpp_blocks_gpu(batch, n, torch.device(DEV), kind, seed=95835 + g)
        res = b200lc.cudpp_compress_batch(data, batch, n, scratch=scratch, out=res)
        assert int(res.error.item()) == 0
        tw = res.total_words.cpu().numpy().astype(np.int64)
        # per block: sum of the histogram = n, offsets are the running sum of block sizes
        hist = res.hist.cpu().numpy().astype(np.int64).reshape(batch, 256)
        assert (hist.sum(1) == n).all()
        offs = res.offsets.cpu().numpy().astype(np.int64).reshape(batch, n // 4096)
        assert (offs[:, 0] == 0).all() and (np.diff(offs, axis=1) > 0).all() and (offs[:, -1] < tw).all()
        total_words += int(tw.sum())
        if g == 0:
            blk = data[:n].cpu().numpy()
            rc, widx, whist, woffs, wwords = O.cudpp_oracle_compress(blk)
            assert rc == 0 and int(res.bwt_index[0].item()) == widx
            assert np.array_equal(res.words[: wwords.size].cpu().numpy().view(np.uint32), wwords)
        _, derr = b200lc.cudpp_decompress_batch(res, batch, n, scratch=scratch, out=back)
        assert int(derr.item()) == 0
        assert torch.equal(back, data), kind
    assert total_words * 4 < nblocks * n
