#!/usr/bin/env python
"""Per-path device-resident timings (CUDA events, best of N) for BASELINE.md section 4.
Usage: python tools/bench_paths.py [cuhd|culzss|cudpp|bsc|bzip2|cpu|all] [--mib M]   (one JSON line per measurement)"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("gpu-lossless-compression_b200")
PEAK = 6557.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def quant_codes_gpu(n_bytes, dev, seed=2024, itemsize=4):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = n_bytes // itemsize
    out = torch.empty(n, dtype=torch.int32 if itemsize == 4 else torch.int16, device=dev)
    chunk = 1 << 26
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        u = torch.rand(m, generator=g, device=dev) - 0.5
        lap = -2.0 * torch.sign(u) * torch.log1p(-2.0 * u.abs())
        out[lo:lo + m] = (512 + torch.round(lap)).clamp_(0, 1023).to(out.dtype)
    return out.view(torch.uint8)


def bench_culzss(mib, dev, kind="quant32"):
    n = mib << 20
    buf_len = 1 << 20
    if kind == "quant32":
        data = quant_codes_gpu(n, dev, itemsize=4)
    elif kind == "quant16":
        data = quant_codes_gpu(n, dev, itemsize=2)
    elif kind == "text":
        t = torch.frombuffer(bytearray(b"the quick brown fox jumps over the lazy dog. " * 23302), dtype=torch.uint8)[:buf_len]
        data = t.to(dev).repeat(mib)
    else:
        data = torch.randint(0, 256, (n,), dtype=torch.uint8, device=dev)
    nbuf = n // buf_len
    L = pkg.lib()
    stride = pkg.culzss_out_stride(buf_len)
    out = torch.empty(nbuf * stride, dtype=torch.uint8, device=dev)
    clen = torch.empty(nbuf, dtype=torch.int32, device=dev)
    scratch = torch.empty(L.b200lc_culzss_encode_scratch_bytes(nbuf, buf_len), dtype=torch.uint8, device=dev)
    fast = {}
    for depth in ("lane", 1, 2, 4):
        ms = timeit(lambda: pkg.culzss_encode(data, buf_len, out, clen, scratch, fast=depth), iters=3, warm=1)
        cf = clen.cpu().numpy().astype(np.int64)
        cbytes = int(np.where(cf == 0, buf_len, cf).sum())
        fast["lane" if depth == "lane" else "depth%d" % depth] = {"encode_ms": ms, "encode_gbs": n / ms / 1e6, "ratio": n / cbytes,
                                   "encode_hbm_frac": (n + cbytes) / ms / 1e6 / PEAK}
    cta_ms = timeit(lambda: pkg.culzss_encode(data, buf_len, out, clen, scratch, kernel=pkg.CULZSS_KERNEL_CTA), iters=2, warm=1)
    enc_ms = timeit(lambda: pkg.culzss_encode(data, buf_len, out, clen, scratch))
    cl = clen.cpu().numpy().astype(np.int64)
    raw = int((cl == 0).sum())
    sizes = np.where(cl == 0, buf_len, cl)
    offs = np.zeros(nbuf + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    comp = torch.empty(int(offs[-1]), dtype=torch.uint8, device=dev)
    for b in range(nbuf):      # pack the per-buffer outputs back to back (container layout)
        src = out[b * stride: b * stride + cl[b]] if cl[b] else data[b * buf_len:(b + 1) * buf_len]
        comp[offs[b]:offs[b + 1]] = src
    d_offs = torch.from_numpy(offs).to(dev)
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    dscratch = torch.empty(L.b200lc_culzss_decode_scratch_bytes(nbuf, buf_len), dtype=torch.uint8, device=dev)
    dec_ms = timeit(lambda: pkg.culzss_decode(comp, d_offs, buf_len, dec, dscratch))
    assert torch.equal(dec, data), "CULZSS round trip mismatch"
    C = int(offs[-1])
    print(json.dumps({"path": "culzss", "data": kind, "mib": mib, "ratio": n / C, "raw_buffers": raw,
                      "encode_ms": enc_ms, "decode_ms": dec_ms,
                      "encode_gbs": n / enc_ms / 1e6, "decode_gbs": n / dec_ms / 1e6,
                      "encode_hbm_frac": (n + C) / enc_ms / 1e6 / PEAK,
                      "decode_hbm_frac": (n + C) / dec_ms / 1e6 / PEAK,
                      "encode_cta_kernel_gbs": n / cta_ms / 1e6, "fast_mode_non_parity": fast}))


def cudpp_blocks_gpu(nblocks, n, dev, kind="zipf", seed=95835):
    """Synthetic C4 input on the GPU: bytes 1..255 (Zipf(1.3) or order-1 Markov), last byte 0."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    N = nblocks * n
    if kind == "zipf":
        p = 1.0 / torch.arange(1, 256, dtype=torch.float64) ** 1.3
        cdf = torch.cumsum(p / p.sum(), 0).to(device=dev, dtype=torch.float32)
        out = torch.empty(N, dtype=torch.uint8, device=dev)
        chunk = 1 << 26
        for lo in range(0, N, chunk):
            m = min(chunk, N - lo)
            u = torch.rand(m, generator=g, device=dev)
            out[lo:lo + m] = (torch.searchsorted(cdf, u).clamp_(max=254) + 1).to(torch.uint8)
    elif kind == "markov":
        steps = torch.randint(-2, 3, (N,), generator=g, device=dev, dtype=torch.int32)
        out = (torch.cumsum(steps, 0) % 200 + 1).to(torch.uint8)
    elif kind == "text":
        # word-structured text (long repeated substrings: many prefix-doubling rounds); 8 distinct
        # blocks from the CPU generator of tests/test_ref_bsc_cpu.py, repeated
        from test_ref_bsc_cpu import synthetic_largefile
        distinct = [torch.frombuffer(bytearray(synthetic_largefile(n, seed=seed + b)), dtype=torch.uint8)
                    for b in range(min(8, nblocks))]
        out = torch.cat([distinct[b % len(distinct)] for b in range(nblocks)]).to(dev)
    else:
        out = torch.randint(1, 256, (N,), generator=g, device=dev, dtype=torch.uint8)
    out.view(nblocks, n)[:, -1] = 0
    return out


def bench_cudpp(mib, dev, kind="zipf", batch=128):
    n = 1 << 20
    nblocks = mib
    data = cudpp_blocks_gpu(nblocks, n, dev, kind)
    L = pkg.lib()
    nb = min(batch, nblocks)
    nhb = n // 4096
    stride = nhb * 1537
    scratch = torch.empty(L.b200lc_cudpp_compress_scratch_bytes(nb, n) + 256, dtype=torch.uint8, device=dev)
    res = None
    stage = {}

    def run_all():
        nonlocal res
        for lo in range(0, nblocks, nb):
            res = pkg.cudpp_compress_batch(data[lo * n:(lo + nb) * n], nb, n, scratch=scratch, out=res)

    total_ms = timeit(run_all, iters=3, warm=1)
    # stage split on one batch
    d0 = data[: nb * n]
    bscr = torch.empty(L.b200lc_bwt_scratch_bytes(nb, n) + 256, dtype=torch.uint8, device=dev)
    bout = torch.empty(nb * n, dtype=torch.uint8, device=dev)
    bidx = torch.empty(nb, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    stage["bwt_ms"] = timeit(lambda: pkg.check(L.b200lc_bwt_batch(d0.data_ptr(), nb, n, bout.data_ptr(), bidx.data_ptr(), bscr.data_ptr(), bscr.numel(), sp), "bwt"), iters=3, warm=1)
    mout = torch.empty(nb * n, dtype=torch.uint8, device=dev)
    stage["mtf_ms"] = timeit(lambda: pkg.check(L.b200lc_mtf_batch(bout.data_ptr(), nb, n, mout.data_ptr(), bscr.data_ptr(), bscr.numel(), sp), "mtf"), iters=3, warm=1)
    hist = torch.empty(nb * 256, dtype=torch.int32, device=dev)
    offs = torch.empty(nb * nhb, dtype=torch.int32, device=dev)
    tw = torch.empty(nb, dtype=torch.int32, device=dev)
    words = torch.empty(nb * stride, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    stage["huffman_ms"] = timeit(lambda: pkg.check(L.b200lc_cudpp_huffman_batch(mout.data_ptr(), nb, n, hist.data_ptr(), offs.data_ptr(), tw.data_ptr(), words.data_ptr(), stride, err.data_ptr(), bscr.data_ptr(), bscr.numel(), sp), "huff"), iters=3, warm=1)
    comp_words = int(tw.sum().item())
    N = nblocks * n
    # decoder (N3) on the last batch + its stages
    del bscr
    dscr = torch.empty(L.b200lc_cudpp_decompress_scratch_bytes(nb, n) + 256, dtype=torch.uint8, device=dev)
    back = torch.empty(nb * n, dtype=torch.uint8, device=dev)
    derr = None

    def run_dec():
        nonlocal derr
        _, derr = pkg.cudpp_decompress_batch(res, nb, n, scratch=dscr, out=back)

    dec_ms = timeit(run_dec, iters=3, warm=1)
    ok = bool(torch.equal(back, data[(nblocks - nb) * n:])) and int(derr.item()) == 0
    dstage = {}
    dstage["imtf_ms"] = timeit(lambda: pkg.check(L.b200lc_inverse_mtf_batch(mout.data_ptr(), nb, n, words.data_ptr(), dscr.data_ptr(), dscr.numel(), sp), "imtf"), iters=3, warm=1)
    dstage["ibwt_ms"] = timeit(lambda: pkg.check(L.b200lc_inverse_bwt_batch(bout.data_ptr(), bidx.data_ptr(), nb, n, back.data_ptr(), err.data_ptr(), dscr.data_ptr(), dscr.numel(), sp), "ibwt"), iters=3, warm=1)
    dstage["huffman_ms"] = max(0.0, dec_ms - dstage["imtf_ms"] - dstage["ibwt_ms"])
    print(json.dumps({"path": "cudpp_compress", "data": kind, "mib": mib, "batch_blocks": nb,
                      "ratio": nb * n / (4.0 * comp_words), "error": int(err.item()),
                      "encode_ms": total_ms, "encode_gbs": N / total_ms / 1e6,
                      "stage_ms_per_batch": stage,
                      "stage_gbs": {k.replace("_ms", ""): nb * n / v / 1e6 for k, v in stage.items()},
                      "decode_ok": ok, "decode_ms_per_batch": dec_ms, "decode_gbs": nb * n / dec_ms / 1e6,
                      "decode_stage_ms_per_batch": dstage}))


def bench_cuhd(mib, dev):
    import numpy as np
    n = mib << 20
    sys.path.insert(0, ROOT)
    import bench as B
    data = B.gen_zipf_gpu(n, dev, 12345)
    hist = pkg.histogram_u8(data).cpu().numpy()
    code, length, lut = pkg.cuhd_build_table(hist)
    d_code = torch.from_numpy(code.view(np.int32)).to(dev)
    d_len = torch.from_numpy(length).to(dev)
    d_lut = torch.from_numpy(lut).to(dev)
    enc = pkg.cuhd_encode(data, d_code, d_len)
    L = pkg.lib()
    units = torch.empty((n * 11 + 31) // 32 + 2, dtype=torch.int32, device=dev)
    bits = torch.zeros(1, dtype=torch.int64, device=dev)
    escr = torch.empty(L.b200lc_cuhd_encode_scratch_bytes(n), dtype=torch.uint8, device=dev)
    enc_ms = timeit(lambda: pkg.cuhd_encode(data, d_code, d_len, units=units, total_bits=bits, scratch=escr, sync=False))
    hist_ms = timeit(lambda: pkg.histogram_u8(data))
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    dscr = torch.empty(L.b200lc_cuhd_decode_scratch_bytes(enc.units.numel()), dtype=torch.uint8, device=dev)
    dec_ms = timeit(lambda: pkg.cuhd_decode(enc.units, n, d_lut, 11, out=out, scratch=dscr))
    assert torch.equal(out, data)
    C = enc.n_units * 4
    print(json.dumps({"path": "cuhd", "data": "zipf1.1", "mib": mib, "ratio": n / C,
                      "hist_ms": hist_ms, "encode_ms": enc_ms, "decode_ms": dec_ms,
                      "hist_gbs": n / hist_ms / 1e6, "encode_gbs": n / enc_ms / 1e6,
                      "decode_gbs": n / dec_ms / 1e6,
                      "hist_hbm_frac": n / hist_ms / 1e6 / PEAK,
                      "encode_hbm_frac": (n + C) / enc_ms / 1e6 / PEAK,
                      "decode_hbm_frac": (n + C) / dec_ms / 1e6 / PEAK}))


def _pool_time(fn, items, threads):
    """Wall time of fn over items on `threads` host threads (ctypes calls release the GIL)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    if threads == 1:
        out = [fn(x) for x in items]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            out = list(ex.map(fn, items))
    return time.perf_counter() - t0, out


def bench_cpu(what):
    """CPU paths timed beside the GPU numbers on the box's own host cores, on a bounded sample of the
    same synthetic inputs: 1 thread and one block/buffer per core.  CULZSS: the oracle port
    (oracle/culzss_oracle.c) of EncodeKernel + aftercomp / DecodeKernel -- the reference has no CPU
    encoder.  cudppCompress: the reference's own CPU golds (computeBwtGold = DC3 sa_gold,
    computeMtfGold, decoder computeCompressGold; oracle/_ref/libref_cudpp.so) with the oracle port
    of the Huffman stage (the reference has no CPU Huffman encoder for this stream)."""
    import ctypes as C
    import oracle_lib as O
    cores = os.cpu_count() or 1
    buf = 1 << 20
    if what in ("culzss", "all"):
        for kind, gen in (("quant32", lambda b: O.quant_codes(buf, seed=2024 + b)),
                          ("text", lambda b: np.frombuffer((b"the quick brown fox jumps over the lazy dog. " * 23302)[:buf], np.uint8).copy())):
            bufs = [gen(b) for b in range(cores)]
            enc1, r1 = _pool_time(lambda x: O.culzss_oracle_compress(x), bufs[:2], 1)
            encn, rn = _pool_time(lambda x: O.culzss_oracle_compress(x), bufs, cores)
            comps = [c for ok, c in rn if ok]
            dec1, _ = _pool_time(lambda c: O.culzss_oracle_decompress(c, buf), comps[:2], 1)
            decn, back = _pool_time(lambda c: O.culzss_oracle_decompress(c, buf), comps, cores)
            print(json.dumps({"path": "culzss_cpu", "kind": "port", "data": kind, "cores": cores,
                              "sample": "%d buffers of 1 MiB (2 for the 1-thread figure)" % len(bufs),
                              "encode_gbs_1thread": 2 * buf / enc1 / 1e9, "encode_gbs_all_cores": len(bufs) * buf / encn / 1e9,
                              "decode_gbs_1thread": min(2, len(comps)) * buf / dec1 / 1e9 if comps else None,
                              "decode_gbs_all_cores": len(comps) * buf / decn / 1e9 if comps else None,
                              "round_trip": all(bool(np.array_equal(b[1], x)) for b, x in zip(back, bufs))}))
    if what in ("cudpp", "all") and O.have_ref("cudpp"):
        R = O.ref_cudpp()
        for kind in ("zipf", "markov"):
            blocks = [O.cudpp_block(buf, kind, seed=b) for b in range(cores)]

            def encode(data):
                bw = np.zeros(buf, np.uint8)
                idx = C.c_int(-1)
                R.ref_cudpp_bwt(data, bw, C.byref(idx), buf)
                mt = np.zeros(buf, np.uint8)
                R.ref_cudpp_mtf(bw, mt, buf)
                rc, hist, offs, words = O.cudpp_oracle_huffman(mt)
                return idx.value, hist, offs, words

            def decode(enc):
                idx, hist, offs, words = enc
                out = np.zeros(buf, np.uint8)
                h257 = np.zeros(257, np.uint32)
                h257[:256] = hist
                R.ref_cudpp_decompress(out, idx, h257, offs.copy(), words.size, words.copy(), buf)
                return out

            enc1, _ = _pool_time(encode, blocks[:1], 1)
            encn, encs = _pool_time(encode, blocks, cores)
            dec1, _ = _pool_time(decode, encs[:1], 1)
            decn, back = _pool_time(decode, encs, cores)
            print(json.dumps({"path": "cudpp_cpu", "kind": "reference golds + oracle Huffman port", "data": kind,
                              "cores": cores, "sample": "%d blocks of 1 MiB (1 for the 1-thread figure)" % len(blocks),
                              "encode_gbs_1thread": buf / enc1 / 1e9, "encode_gbs_all_cores": len(blocks) * buf / encn / 1e9,
                              "decode_gbs_1thread": buf / dec1 / 1e9, "decode_gbs_all_cores": len(blocks) * buf / decn / 1e9,
                              "round_trip": all(bool(np.array_equal(b, x)) for b, x in zip(back, blocks))}))


def bench_bzip2():
    """cuda-bzip2 back end of one -9 block (900 kB), HOST pointers like the reference's gpuBlockSort:
    GPU sort (b200lc_bzip2_rotation_order), MTF + RUNA/RUNB (b200lc_bzip2_mtf_rle), Huffman stage
    (b200lc_bzip2_send_mtf_values) next to the reference's own generateMTFValues + sendMTFValues
    on one host core (oracle/_ref/libref_bzip2_mtf.so: the "serial huffman.c stage" baseline)."""
    import ctypes as C
    import time
    import oracle_lib as O
    from test_ref_bsc_cpu import synthetic_largefile
    n = 900000 - 19
    L = pkg.lib()
    L.b200lc_bzip2_rotation_order.restype = C.c_int
    L.b200lc_bzip2_rotation_order.argtypes = [np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS"), C.c_int,
                                              np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"), C.POINTER(C.c_int)]
    for kind, block in (("text", np.frombuffer(synthetic_largefile(n, seed=3), np.uint8).copy()),
                        ("quant32", O.quant_codes(n + 3)[:n].copy())):
        ptr = np.zeros(n, np.uint32)
        orig = C.c_int(-1)
        in_use = O.bzip2_in_use(block)

        def best(fn, reps=3):
            out, t = None, 1e30
            for _ in range(reps):
                t0 = time.perf_counter()
                out = fn()
                t = min(t, time.perf_counter() - t0)
            return t, out

        L.b200lc_bzip2_rotation_order(block, n, ptr, C.byref(orig))      # warm-up (context, work areas)
        t_sort, _ = best(lambda: L.b200lc_bzip2_rotation_order(block, n, ptr, C.byref(orig)))
        # raw C-ABI calls on preallocated host arrays (the Python wrappers add np.unique etc.)
        mtfv_buf = np.zeros(n + 1, np.uint16)
        freq_buf = np.zeros(258, np.int32)
        n_mtf, used_c = C.c_int(0), C.c_int(0)
        call_mtf = lambda: L.b200lc_bzip2_mtf_rle(block.ctypes.data, ptr.ctypes.data, n, in_use.ctypes.data,  # noqa: E731
                                                  mtfv_buf.ctypes.data, C.byref(n_mtf), freq_buf.ctypes.data,
                                                  C.byref(used_c))
        assert call_mtf() == 0
        t_mtf, _ = best(call_mtf)
        mtfv, freq, used = mtfv_buf[: n_mtf.value].copy(), freq_buf.copy(), used_c.value
        cap = n * 17 // 8 + n // 50 + 8192
        bits_buf = np.zeros(cap, np.uint8)
        nbits = C.c_ulonglong(0)
        call_huf = lambda: L.b200lc_bzip2_send_mtf_values(mtfv.ctypes.data, mtfv.size, freq.ctypes.data,  # noqa: E731
                                                          in_use.ctypes.data, used, bits_buf.ctypes.data, cap,
                                                          C.byref(nbits), None, None)
        assert call_huf() == 0
        t_huf, _ = best(call_huf)
        gn = nbits.value
        gb = bits_buf[: (gn + 7) // 8].copy()
        freq = freq[: used + 2]
        rec = {"path": "bzip2_block_back_end", "data": kind, "block_bytes": n, "ratio": 8.0 * n / gn,
               "gpu_sort_ms": t_sort * 1e3, "gpu_mtf_rle_ms": t_mtf * 1e3, "gpu_huffman_ms": t_huf * 1e3,
               "gpu_gbs_host_buffers": n / (t_sort + t_mtf + t_huf) / 1e9}
        if O.have_ref("bzip2_mtf"):
            c_mtf, (rm, rf, ru) = best(lambda: O.bzip2_ref_mtf_rle(block, ptr), 2)
            c_huf, (rb, rn, _, _) = best(lambda: O.bzip2_ref_send_mtf(rm, rf, in_use, ru), 2)
            rec.update({"cpu_generateMTFValues_ms": c_mtf * 1e3, "cpu_sendMTFValues_ms": c_huf * 1e3, "cpu_cores": 1,
                        "cpu_gbs_mtf_huffman": n / (c_mtf + c_huf) / 1e9,
                        "identical": bool(rn == gn and np.array_equal(rb, gb) and np.array_equal(rm, mtfv))})
        print(json.dumps(rec))


def bench_bsc(mib):
    """libbsc BWT stage (row N4), HOST buffers: GPU bsc_bwt_encode (H2D + suffix sort + D2H inside)
    next to the reference's divbwt on the host cores (oracle/_ref/libref_bsc.so)."""
    import time
    import oracle_lib as O
    from test_ref_bsc_cpu import synthetic_largefile
    n = mib << 20
    for kind, data in (("text", np.frombuffer(synthetic_largefile(n, seed=11), np.uint8)),
                       ("quant32", O.quant_codes(n))):
        pkg.bsc_bwt_encode(data[: 1 << 20])          # warm-up: context, work area
        pkg.bsc_bwt_encode(data)
        t0 = time.perf_counter()
        gu, gp, gi = pkg.bsc_bwt_encode(data)
        gpu_s = time.perf_counter() - t0
        rec = {"path": "bsc_bwt_encode", "data": kind, "mib": mib, "gpu_ms": gpu_s * 1e3,
               "gpu_gbs_host_buffers": n / gpu_s / 1e9}
        if O.have_ref("bsc"):
            for feat, label in ((0, "cpu_1thread"), (2, "cpu_openmp")):
                t = data.copy()
                num = np.zeros(1, np.uint8)
                idx = np.zeros(256, np.int32)
                t0 = time.perf_counter()
                rp = O.ref_bsc().bsc_bwt_encode(t, n, num, idx, feat)
                s_ = time.perf_counter() - t0
                rec[label + "_ms"] = s_ * 1e3
                rec[label + "_gbs"] = n / s_ / 1e9
                rec["identical"] = bool(rp == gp and np.array_equal(t, gu))
            rec["host_cores"] = os.cpu_count()
        # the inverse: GPU bsc_bwt_decode next to the reference's CPU one (serial; its parallel mode
        # needs the secondary indexes)
        pkg.bsc_bwt_decode(gu, gp)
        t0 = time.perf_counter()
        back = pkg.bsc_bwt_decode(gu, gp)
        dec_s = time.perf_counter() - t0
        rec["decode_gpu_ms"] = dec_s * 1e3
        rec["decode_gpu_gbs_host_buffers"] = n / dec_s / 1e9
        rec["decode_ok"] = bool(np.array_equal(back, data))
        if O.have_ref("bsc"):
            t = gu.copy()
            idx = np.zeros(256, np.int32)
            t0 = time.perf_counter()
            O.ref_bsc().bsc_bwt_decode(t, n, gp, 0, idx, 0)
            s_ = time.perf_counter() - t0
            rec["decode_cpu_1thread_ms"] = s_ * 1e3
            rec["decode_cpu_1thread_gbs"] = n / s_ / 1e9
        print(json.dumps(rec))
    # Sort Transform ST5..ST8 (bsc_st_encode_cuda, row N4), HOST buffers; the reference's CPU path
    # covers k = 5, 6 only (st.cpp:1021-1026)
    import ctypes as C
    L = pkg.lib()
    L.bsc_st_encode_cuda.restype = C.c_int
    L.bsc_st_encode_cuda.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    data = np.frombuffer(synthetic_largefile(n, seed=11), np.uint8)
    for k in (5, 6, 7, 8):
        t = data.copy()
        L.bsc_st_encode_cuda(t.ctypes.data, n, k, 0)          # warm-up: work area
        t = data.copy()
        t0 = time.perf_counter()
        gi = L.bsc_st_encode_cuda(t.ctypes.data, n, k, 0)
        gpu_s = time.perf_counter() - t0
        rec = {"path": "bsc_st_encode", "k": k, "data": "text", "mib": mib, "gpu_ms": gpu_s * 1e3,
               "gpu_gbs_host_buffers": n / gpu_s / 1e9, "index": int(gi)}
        if O.have_ref("bsc") and k <= 6:
            t0 = time.perf_counter()
            ref, ri = O.bsc_ref_st_encode(data, k)
            s_ = time.perf_counter() - t0
            rec.update({"cpu_1thread_ms": s_ * 1e3, "cpu_1thread_gbs": n / s_ / 1e9,
                        "identical": bool(ri == gi and np.array_equal(ref, t))})
        print(json.dumps(rec))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--mib", type=int, default=1024)
    args = ap.parse_args()
    if args.what == "cpu":
        bench_cpu("all")
        return
    dev = torch.device("cuda:0")
    if args.what in ("cuhd", "all"):
        bench_cuhd(args.mib, dev)
    if args.what in ("culzss", "all"):
        for kind in ("quant32", "quant16", "text", "random"):
            bench_culzss(args.mib, dev, kind)
    if args.what in ("cudpp", "all"):
        for kind in ("zipf", "markov", "rand"):
            bench_cudpp(min(args.mib, 256), dev, kind)
    if args.what in ("bsc", "all"):
        bench_bsc(25)
    if args.what in ("bzip2", "all"):
        bench_bzip2()
    if args.what in ("cpu", "all"):
        bench_cpu("all")


if __name__ == "__main__":
    main()
