#!/usr/bin/env python
"""Per-path device-resident timings (CUDA events, best of N) for BASELINE.md section 4.
Usage: python tools/bench_paths.py [culzss|all] [--mib M]   (one JSON line per measurement)"""
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
PEAK = 6538.3
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
                      "decode_hbm_frac": (n + C) / dec_ms / 1e6 / PEAK}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--mib", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.what in ("culzss", "all"):
        for kind in ("quant32", "quant16", "text", "random"):
            bench_culzss(args.mib, dev, kind)


if __name__ == "__main__":
    main()
