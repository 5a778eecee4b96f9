#!/usr/bin/env python
"""One run of the planned CUHD packer on C2 data (1 GiB Zipf(1.1)) for an ncu capture:
ncu --set full --clock-control none --import-source on -k regex:cuhd_encode_kernel -c 1 \
    -o gpurun_out/x python tools/ncu_planned_encode.py [--mib 1024]"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("gpu-lossless-compression_b200")
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n = a.mib << 20
    data = B.gen_zipf_gpu(n, dev, 12345)
    hist, ph = pkg.histogram_u8_pieces(data)
    code, length, _ = pkg.cuhd_build_table(hist.cpu().numpy())
    d_code = torch.from_numpy(code.view(np.int32)).to(dev)
    d_len = torch.from_numpy(length).to(dev)
    enc = pkg.cuhd_encode(data, d_code, d_len, piece_hist=ph)
    torch.cuda.synchronize()
    print("units", enc.n_units, "ratio", n / (4.0 * enc.n_units))


if __name__ == "__main__":
    main()
