#!/usr/bin/env python
"""The reference's own GPU code, compiled for sm_100a (oracle/_ref), timed on this B200 next to the
library on the same inputs and through the same call protocol -- "reference GPU kernels on the B200".

  CUHD   the reference demo program (cuhd-icpp/src/demo.cc, unmodified) with the reference's decoder
         (oracle/_ref/cuhd_demo_ref) and with libb200lc.so behind the compat headers
         (oracle/_ref/cuhd_demo_b200): both print their own stage timers ("decoding", memcpys)
  CULZSS compression_kernel_wrapper + onestream_finish_GPU + aftercompression_wrapper and
         decompression_kernel_wrapper per 1 MiB buffer (host buffers, the reference's protocol,
         culzss.c:108,170,176 / deculzss.c:98) from oracle/_ref/libref_culzss.so and from libb200lc.so
One JSON line per comparison.  Usage: python tools/ref_gpu_compare.py [--mib 256]"""
import argparse
import ctypes as C
import importlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib as O  # noqa: E402

pkg = importlib.import_module("gpu-lossless-compression_b200")
MIB = 1 << 20


def demo_timers(exe, src, dst):
    r = subprocess.run([exe, "0", src, dst], capture_output=True, text=True, timeout=900)
    t = {m.group(1).strip(): int(m.group(2)) for m in re.finditer(r"^(.*?)\.\. (\d+)", r.stdout, re.M)}
    return t, ("mismatch" in r.stdout), r.returncode


def cuhd(mib):
    data = O.zipf_bytes(mib * MIB, 1.1, seed=12345)
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "in.bin")
        data.tofile(src)
        rec = {"path": "cuhd_demo", "mib": mib, "data": "zipf1.1"}
        for tag, exe in (("reference_gpu", "cuhd_demo_ref"), ("b200lc", "cuhd_demo_b200")):
            path = os.path.join(O.ORACLE_DIR, "_ref", exe)
            if not os.path.exists(path):
                continue
            best = None
            for _ in range(2):
                t, mismatch, rc = demo_timers(path, src, os.path.join(td, tag + ".out"))
                rec[tag + "_self_check"] = "mismatch" if mismatch else ("ok" if rc == 0 else "rc=%d" % rc)
                if rc == 0 and "decoding" in t and (best is None or t["decoding"] < best["decoding"]):
                    best = t
            if best:
                rec[tag + "_decoding_us"] = best.get("decoding")
                rec[tag + "_memcpy_dth_us"] = best.get("GPU memcpy DtH")
                rec[tag + "_memcpy_htd_us"] = best.get("GPU memcpy HtD")
                # the reference does not synchronise after its last kernel: count decode + DtH
                rec[tag + "_decode_plus_dth_gbs"] = mib * MIB / ((best.get("decoding", 0) + best.get("GPU memcpy DtH", 0)) * 1e-6) / 1e9
        print(json.dumps(rec))


def culzss_protocol(lib, bufs, index_base):
    vp = C.c_void_p
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    lib.initGPUmem.restype = vp
    lib.initGPUmem.argtypes = [C.c_int]
    lib.deleteGPUmem.argtypes = [vp]
    lib.compression_kernel_wrapper.restype = C.c_int
    lib.compression_kernel_wrapper.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.aftercompression_wrapper.restype = C.c_int
    lib.aftercompression_wrapper.argtypes = [u8p, C.c_int, u8p, C.POINTER(C.c_int)]
    lib.decompression_kernel_wrapper.restype = C.c_int
    lib.decompression_kernel_wrapper.argtypes = [u8p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
    lib.onestream_finish_GPU.argtypes = [C.c_int]
    lib.initGPU()
    in_d, out_d = lib.initGPUmem(MIB), lib.initGPUmem(2 * MIB)
    work = [np.zeros(2 * MIB + 4096, np.uint8) for _ in bufs]
    tok = np.zeros(2 * MIB, np.uint8)
    for w, b in zip(work, bufs):
        w[:MIB] = b
    clens = []
    t0 = time.perf_counter()
    for w in work:
        lib.compression_kernel_wrapper(w, MIB, tok, 0, 0, 128, 0, index_base, in_d, out_d)
        lib.onestream_finish_GPU(index_base)
        clen = C.c_int(0)
        lib.aftercompression_wrapper(w, MIB, tok, C.byref(clen))
        clens.append(clen.value)
    enc_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ok = True
    for w, cl, b in zip(work, clens, bufs):
        dlen = C.c_int(0)
        lib.decompression_kernel_wrapper(w, cl, C.byref(dlen), 0, 0, 1)
        ok = ok and dlen.value == MIB
    dec_s = time.perf_counter() - t0
    ok = ok and all(np.array_equal(w[:MIB], b) for w, b in zip(work, bufs))
    lib.deleteGPUmem(in_d)
    lib.deleteGPUmem(out_d)
    return enc_s, dec_s, sum(clens), ok


def culzss(nbuf):
    bufs = [O.quant_codes(MIB, seed=2024 + b) for b in range(nbuf)]
    rec = {"path": "culzss_wrappers", "buffers": nbuf, "data": "quant32", "protocol": "per 1 MiB host buffer"}
    for tag, lib in (("reference_gpu", O.ref_culzss() if O.have_ref("culzss") else None), ("b200lc", pkg.lib())):
        if lib is None:
            continue
        culzss_protocol(lib, bufs[:2], 0)          # warm-up
        enc_s, dec_s, comp, ok = culzss_protocol(lib, bufs, 0)
        rec[tag + "_encode_gbs"] = nbuf * MIB / enc_s / 1e9
        rec[tag + "_decode_gbs"] = nbuf * MIB / dec_s / 1e9
        rec[tag + "_round_trip"] = ok
        rec[tag + "_compressed"] = comp
    print(json.dumps(rec))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--buffers", type=int, default=32)
    a = ap.parse_args()
    cuhd(a.mib)
    culzss(a.buffers)
