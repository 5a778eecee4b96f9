#!/usr/bin/env python
"""Inverse BWT time per block as a function of the batch size (is the walk L2- or DRAM-bound?)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
from bench_paths import cudpp_blocks_gpu, timeit  # noqa: E402

pkg = importlib.import_module("gpu-lossless-compression_b200")
L = pkg.lib()
dev = torch.device("cuda:0")
n = 1 << 20
data = cudpp_blocks_gpu(128, n, dev, "zipf")
scr = torch.empty(L.b200lc_bwt_scratch_bytes(128, n) + 256, dtype=torch.uint8, device=dev)
bwt = torch.empty(128 * n, dtype=torch.uint8, device=dev)
idx = torch.empty(128, dtype=torch.int32, device=dev)
sp = torch.cuda.current_stream().cuda_stream
pkg.check(L.b200lc_bwt_batch(data.data_ptr(), 128, n, bwt.data_ptr(), idx.data_ptr(), scr.data_ptr(), scr.numel(), sp), "bwt")
del scr
back = torch.empty(128 * n, dtype=torch.uint8, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
for nb in (4, 8, 16, 32, 64, 128):
    dscr = torch.empty(L.b200lc_cudpp_decompress_scratch_bytes(nb, n) + 256, dtype=torch.uint8, device=dev)

    def run():
        for lo in range(0, 128, nb):
            pkg.check(L.b200lc_inverse_bwt_batch(bwt.data_ptr() + lo * n, idx.data_ptr() + 4 * lo, nb, n,
                                                 back.data_ptr() + lo * n, err.data_ptr(), dscr.data_ptr(), dscr.numel(), sp), "ibwt")
    ms = timeit(run, iters=3, warm=1)
    ok = bool(torch.equal(back, data))
    print("batch %3d blocks: %.2f ms per 128 blocks (%.1f GB/s) ok=%s" % (nb, ms, 128 * n / ms / 1e6, ok))
    del dscr
