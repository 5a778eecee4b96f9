"""Parity at BASELINE.json's full sizes.  Round-trip and accounting properties on the synthetic
generators of bench.py / tools/bench_paths.py, and -- round 2 -- the WHOLE of each full-size
output against the checker: C2 on SURVEY.md 8(d)'s input with the reference's own dictionary and
stream (oracle/_ref llhuff), every one of C3's 4096 buffers against the oracle on all host cores,
every block of one C4 batch against the oracle."""
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


# ------------------------------------------------------------------ whole outputs against the checker
def _pool(fn, items):
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:      # the ctypes calls release the GIL
        return list(ex.map(fn, items))


@pytest.mark.skipif(not O.have_ref("cuhd"), reason="oracle/_ref/libref_cuhd.so not built")
def test_c2_one_gib_reference_input_dictionary_and_stream():
    """SURVEY.md 8(d) C2 as specified: zipf_bytes(2^30, 1.1, seed 12345), table AND stream from the
    reference encoder (llhuffman_encoder.cc:18-238).  The packer bench.py times
    (b200lc_cuhd_encode_planned), given the reference's dictionary, must write the reference's
    stream unit for unit, and the decoder must turn the reference's stream back into the input."""
    n = 1 << 30
    data = O.zipf_bytes(n, 1.1, seed=12345)
    code, length, lut, ref_units = O.cuhd_ref_encode(data)
    assert int(length.max()) == 11
    d = torch.from_numpy(data).to(DEV)
    hist, ph = b200lc.histogram_u8_pieces(d)
    assert np.array_equal(hist.cpu().numpy(), np.bincount(data, minlength=256))
    d_code = torch.from_numpy(code.view(np.int32)).to(DEV)
    d_len = torch.from_numpy(length).to(DEV)
    enc = b200lc.cuhd_encode(d, d_code, d_len, piece_hist=ph)
    bits = int((hist.cpu().numpy().astype(np.int64) * length.astype(np.int64)).sum())
    assert enc.bits == bits and enc.n_units == ref_units.size == (bits + 31) // 32
    want = torch.from_numpy(ref_units.view(np.int32))
    got = enc.units[: enc.n_units].cpu()
    # every unit but the last: the reference loses the codeword that straddles into its final unit
    # (SURVEY.md R3), so that unit is not comparable
    assert torch.equal(got[:-1], want[:-1])
    # two-pass packer too
    enc2 = b200lc.cuhd_encode(d, d_code, d_len)
    assert torch.equal(enc2.units[: enc2.n_units], enc.units[: enc.n_units])
    del enc2
    # the reference's stream (+ its pad unit) through the decoder
    d_lut = torch.from_numpy(np.ascontiguousarray(lut)).to(DEV)
    stream = torch.cat([want, torch.zeros(1, dtype=torch.int32)]).to(DEV)
    out = b200lc.cuhd_decode(stream, n, d_lut)
    # the reference drops the tail of a codeword split across the last unit: all but the final symbol
    assert torch.equal(out[:-1], d[:-1])
    out = b200lc.cuhd_decode(enc.units, n, d_lut)
    assert torch.equal(out, d)


def test_c3_every_buffer_of_four_gib_against_the_oracle():
    from bench_paths import quant_codes_gpu
    n, buf = 4 << 30, MIB
    nbuf = n // buf
    data = quant_codes_gpu(n, torch.device(DEV), itemsize=4)
    out, clen = b200lc.culzss_encode(data, buf)
    cl = clen.cpu().numpy().astype(np.int64)
    stride = b200lc.culzss_out_stride(buf)
    rows = out.view(nbuf, stride)
    bad = []
    for lo in range(0, nbuf, 512):
        h_in = data[lo * buf:(lo + 512) * buf].cpu().numpy().reshape(512, buf)
        h_out = rows[lo:lo + 512].cpu().numpy()

        def check(i):
            ok, want = O.culzss_oracle_compress(h_in[i])
            return ok and int(cl[lo + i]) == want.size and np.array_equal(h_out[i, : want.size], want)

        bad += [lo + i for i, good in enumerate(_pool(check, range(512))) if not good]
    assert not bad, bad[:10]


def test_c4_every_block_of_a_batch_against_the_oracle():
    from bench_paths import cudpp_blocks_gpu
    n, batch = MIB, 128
    for kind, seed in (("zipf", 95835), ("markov", 95836)):
        data = cudpp_blocks_gpu(batch, n, torch.device(DEV), kind, seed=seed)
        res = b200lc.cudpp_compress_batch(data, batch, n)
        assert int(res.error.item()) == 0
        h = data.cpu().numpy().reshape(batch, n)
        idx = res.bwt_index.cpu().numpy()
        hist = res.hist.cpu().numpy().reshape(batch, 256)
        offs = res.offsets.cpu().numpy().reshape(batch, n // 4096)
        tw = res.total_words.cpu().numpy()
        words = res.words.cpu().numpy().view(np.uint32).reshape(batch, -1)

        def check(b):
            rc, widx, whist, woffs, wwords = O.cudpp_oracle_compress(h[b])
            return (rc == 0 and int(idx[b]) == widx and np.array_equal(hist[b], whist)
                    and np.array_equal(offs[b], woffs) and int(tw[b]) == wwords.size
                    and np.array_equal(words[b, : wwords.size], wwords))

        bad = [b for b, good in enumerate(_pool(check, range(batch))) if not good]
        assert not bad, (kind, bad[:10])
