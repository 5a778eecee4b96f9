#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container: python tools/make_golden.py

  cuhd_zipf.npz    : llhuff tables + packed stream for 20000 Zipf(1.1) symbols
  cuhd_binom.npz   : same for the reference demo's binomial byte distribution (demo.cc.ori:54-63)
  culzss_quant.npz : reference aftercompression_wrapper output for a 64 KiB quant-code buffer
  culzss_text.npz  : same for 64 KiB of the bundled pg1661.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def cuhd(name, data):
    code, length, lut, units = O.cuhd_ref_encode(data)
    _, defined = O.cuhd_oracle_encode(data, code, length)
    np.savez_compressed(os.path.join(OUT, name), data=data, code=code, length=length, lut=lut,
                        units=units, defined_units=np.int64(defined))


def culzss(name, data):
    tokens = O.culzss_oracle_tokens(data)
    ok, comp = O.culzss_ref_aftercomp(tokens, data)
    assert ok == 1
    np.savez_compressed(os.path.join(OUT, name), data=data, comp=comp,
                        token_lens=tokens[0::2].copy(), token_offs=tokens[1::2].copy())


def main():
    os.makedirs(OUT, exist_ok=True)
    cuhd("cuhd_zipf.npz", O.zipf_bytes(20000, 1.1, seed=12345))
    rng = np.random.Generator(np.random.MT19937(7))
    cuhd("cuhd_binom.npz", rng.binomial(255, 0.5, 20000).astype(np.uint8))
    culzss("culzss_quant.npz", O.quant_codes(1 << 16, seed=2024))
    ref_txt = "/root/reference/cuda-lzss-unknown/pg1661.txt"
    text = np.frombuffer(open(ref_txt, "rb").read()[4096:4096 + (1 << 16)], np.uint8).copy()
    culzss("culzss_text.npz", text)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
