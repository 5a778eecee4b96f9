#!/usr/bin/env python
"""One BWT batch per data kind (128 x 1 MiB by default) for `ncu --metrics gpu__time_duration.sum`
launch lists.  Usage: python tools/bwt_profile.py [kind ...] [--blocks B]"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

pkg = importlib.import_module("gpu-lossless-compression_b200")
from bench_paths import cudpp_blocks_gpu, timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kinds", nargs="*", default=["zipf", "markov", "rand"])
    ap.add_argument("--blocks", type=int, default=128)
    ap.add_argument("--iters", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n = 1 << 20
    L = pkg.lib()
    nb = args.blocks
    scr = torch.empty(L.b200lc_bwt_scratch_bytes(nb, n) + 256, dtype=torch.uint8, device=dev)
    out = torch.empty(nb * n, dtype=torch.uint8, device=dev)
    idx = torch.empty(nb, dtype=torch.int32, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    for kind in args.kinds:
        data = cudpp_blocks_gpu(nb, n, dev, kind)
        ms = timeit(lambda: pkg.check(L.b200lc_bwt_batch(data.data_ptr(), nb, n, out.data_ptr(), idx.data_ptr(),
                                                         scr.data_ptr(), scr.numel(), sp), "bwt"),
                    iters=args.iters, warm=0)
        print(kind, "bwt ms", ms, "GB/s", nb * n / ms / 1e6)


if __name__ == "__main__":
    main()
