#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref, built from
/root/reference by oracle/Makefile).  Run in the build container: python tools/make_golden.py

  cuhd_zipf.npz    : llhuff tables + packed stream for 20000 Zipf(1.1) symbols
  cuhd_binom.npz   : same for the reference demo's binomial byte distribution (demo.cc.ori:54-63)
  culzss_quant.npz : reference aftercompression_wrapper output for a 64 KiB quant-code buffer
  culzss_text.npz  : same for 64 KiB of the bundled pg1661.txt
  bsc_*.npz        : reference bsc_bwt_encode (divbwt): U, primary index, secondary indexes
  cuhd_c2_table.npz: the reference's dictionary (llhuffman_encoder.cc:18-198) for the full C2 input,
                     zipf_bytes(2^30, 1.1, seed 12345): code, length, LUT, histogram, stream units and a
                     CRC32 per 1 MiB of stream (`python tools/make_golden.py c2`, ~2 min, 6 GiB)
  bzip2_*.npz      : reference generateMTFValues + sendMTFValues on a sorted block: mtfv, mtfFreq,
                     code lengths, selectors, bit string
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


def cuhd_c2_table():
    import zlib
    data = O.zipf_bytes(1 << 30, 1.1, seed=12345)
    code, length, lut, units = O.cuhd_ref_encode(data)
    raw = units.view(np.uint8)
    crcs = np.array([zlib.crc32(raw[i:i + (1 << 20)].tobytes()) for i in range(0, raw.size - 4, 1 << 20)], np.uint32)
    np.savez_compressed(os.path.join(OUT, "cuhd_c2_table.npz"), code=code, length=length, lut=lut,
                        hist=np.bincount(data, minlength=256).astype(np.int64), n_units=np.int64(units.size),
                        unit_crc32_per_mib=crcs)


def culzss(name, data):
    tokens = O.culzss_oracle_tokens(data)
    ok, comp = O.culzss_ref_aftercomp(tokens, data)
    assert ok == 1
    np.savez_compressed(os.path.join(OUT, name), data=data, comp=comp,
                        token_lens=tokens[0::2].copy(), token_offs=tokens[1::2].copy())


def cudpp(name, data):
    """Reference golds (computeBwtGold / computeMtfGold / huffman_build_tree_cpu) for one block;
    the compressed words are the oracle's, accepted only if the reference decoder reproduces the
    block from them (n = 1 MiB cases) -- see tests/test_oracle_cudpp.py."""
    import ctypes as C
    ref = O.ref_cudpp()
    n = data.size
    bwt = np.zeros(n, np.uint8)
    idx = C.c_int(-1)
    ref.ref_cudpp_bwt(data, bwt, C.byref(idx), n)
    mtf = np.zeros(n, np.uint8)
    ref.ref_cudpp_mtf(bwt, mtf, n)
    hist = np.bincount(mtf, minlength=256).astype(np.uint32)
    arrs = [np.zeros(513, np.int32) for _ in range(4)]
    head = C.c_int(-1)
    ref.ref_cudpp_tree(hist, arrs[0], arrs[1], arrs[2], arrs[3], C.byref(head))
    rc, ohist, offs, words = O.cudpp_oracle_huffman(mtf)
    assert rc == 0 and np.array_equal(ohist, hist)
    np.savez_compressed(os.path.join(OUT, name), data=data, bwt=bwt, bwt_index=np.int64(idx.value),
                        mtf=mtf, tree_left=arrs[0], tree_right=arrs[1], tree_value=arrs[3],
                        tree_head=np.int64(head.value), offsets=offs, words=words)


def bsc(name, data):
    u, p, idx = O.bsc_ref_bwt_encode(data)
    np.savez_compressed(os.path.join(OUT, name), data=data, U=u, primary=np.int64(p), indexes=idx)


def bzip2(name, block):
    n = block.size
    ptr = np.zeros(n, np.uint32)
    O.oracle().bzip2_oracle_rotation_order(block, n, ptr)       # any correct rotation order
    mtfv, freq, used = O.bzip2_ref_mtf_rle(block, ptr)
    bits, nbits, lens, sel = O.bzip2_ref_send_mtf(mtfv, freq, O.bzip2_in_use(block), used)
    np.savez_compressed(os.path.join(OUT, name), block=block, ptr=ptr, mtfv=mtfv, freq=freq,
                        n_in_use=np.int64(used), bits=bits, nbits=np.int64(nbits), lens=lens, selector=sel)


def main():
    os.makedirs(OUT, exist_ok=True)
    bsc("bsc_zipf.npz", O.zipf_bytes(70001, 1.3, seed=3))
    bsc("bsc_binary.npz", np.random.default_rng(5).integers(0, 2, 20000, dtype=np.uint8))
    bzip2("bzip2_text.npz", np.frombuffer((b"it was the best of times, it was the worst of times, " * 600)[:30000],
                                          np.uint8).copy())
    bzip2("bzip2_zipf.npz", O.zipf_bytes(40000, 1.5, seed=4))
    cudpp("cudpp_zipf.npz", O.cudpp_block(32768, "zipf", seed=1))
    cudpp("cudpp_text.npz", O.cudpp_block(32768, "text", seed=2))
    cuhd("cuhd_zipf.npz", O.zipf_bytes(20000, 1.1, seed=12345))
    rng = np.random.Generator(np.random.MT19937(7))
    cuhd("cuhd_binom.npz", rng.binomial(255, 0.5, 20000).astype(np.uint8))
    culzss("culzss_quant.npz", O.quant_codes(1 << 16, seed=2024))
    ref_txt = "/root/reference/cuda-lzss-unknown/pg1661.txt"
    text = np.frombuffer(open(ref_txt, "rb").read()[4096:4096 + (1 << 16)], np.uint8).copy()
    culzss("culzss_text.npz", text)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "c2":
    cuhd_c2_table()
    sys.exit(0)

if __name__ == "__main__":
    main()
