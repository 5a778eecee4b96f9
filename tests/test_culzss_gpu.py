"""GPU parity for hot path 2 (CULZSS): product kernels vs the CPU oracle and vs the reference's
own kernels (oracle/_ref/libref_culzss.so = gpu_compress.cu + gpu_decompress.cu for sm_100a)
executed on the same GPU.  Bar: bit-exact compressed buffers and byte-identical round trips."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle_lib as O
from pkg import b200lc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MIB = 1 << 20


def _cases():
    rng = np.random.default_rng(0)
    text = np.frombuffer((b"the quick brown fox jumps over the lazy dog. " * 30000)[:MIB], np.uint8).copy()
    mix = np.concatenate([O.quant_codes(MIB // 2, seed=3), rng.integers(0, 4, MIB // 2, dtype=np.uint8)])
    return {
        "quant32": O.quant_codes(MIB),
        "quant16": O.quant_codes(MIB, dtype=np.uint16),
        "spaces": np.full(MIB, 0x20, np.uint8),
        "zeros": np.zeros(MIB, np.uint8),
        "text": text,
        "mix": mix,
        "random": rng.integers(0, 256, MIB, dtype=np.uint8),     # expands -> stored raw
        "carets": np.full(MIB, ord("^"), np.uint8),              # '^' is the last-chunk filler
        "ramp": (np.arange(MIB) % 97).astype(np.uint8),
    }


def _gpu_encode(data, buf_length=MIB):
    d = torch.from_numpy(data).to(DEV)
    out, clen = b200lc.culzss_encode(d, buf_length)
    torch.cuda.synchronize()
    stride = b200lc.culzss_out_stride(buf_length)
    out = out.cpu().numpy()
    clen = clen.cpu().numpy()
    return [out[b * stride: b * stride + clen[b]] for b in range(clen.size)], clen


@pytest.mark.parametrize("name", list(_cases().keys()))
def test_encode_matches_oracle(name):
    data = _cases()[name]
    ok, want = O.culzss_oracle_compress(data)
    bufs, clen = _gpu_encode(data)
    if not ok:
        assert clen[0] == 0
        return
    assert clen[0] == want.size
    assert np.array_equal(bufs[0], want)


def test_encode_small_and_multi_buffer():
    # 64 KiB buffers (16 packets) and a 5-buffer batch with one expanding buffer in the middle
    rng = np.random.default_rng(4)
    parts = [O.quant_codes(1 << 16, seed=s) for s in (1, 2)] + \
            [rng.integers(0, 256, 1 << 16, dtype=np.uint8)] + \
            [np.zeros(1 << 16, np.uint8), O.quant_codes(1 << 16, seed=9, dtype=np.uint16)]
    data = np.concatenate(parts)
    bufs, clen = _gpu_encode(data, 1 << 16)
    for b, part in enumerate(parts):
        ok, want = O.culzss_oracle_compress(part)
        if ok:
            assert np.array_equal(bufs[b], want), b
        else:
            assert clen[b] == 0


def _decode_gpu(comp_list, buf_length):
    offs = np.zeros(len(comp_list) + 1, np.int64)
    offs[1:] = np.cumsum([c.size for c in comp_list])
    comp = torch.from_numpy(np.concatenate(comp_list)).to(DEV)
    out = b200lc.culzss_decode(comp, torch.from_numpy(offs).to(DEV), buf_length)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("name", ["quant32", "quant16", "spaces", "zeros", "text", "mix", "ramp"])
def test_decode_matches_oracle_and_roundtrips(name):
    data = _cases()[name]
    ok, comp = O.culzss_oracle_compress(data)
    assert ok
    dok, want = O.culzss_oracle_decompress(comp, data.size)
    assert dok and np.array_equal(want, data)
    got = _decode_gpu([comp], MIB)
    assert np.array_equal(got, data)


def test_decode_batch_with_raw_buffer_and_unaligned_offsets():
    rng = np.random.default_rng(5)
    parts = [O.quant_codes(1 << 16, seed=11), rng.integers(0, 256, 1 << 16, dtype=np.uint8),
             np.full(1 << 16, 7, np.uint8)]
    comps = []
    for part in parts:
        ok, comp = O.culzss_oracle_compress(part)
        comps.append(comp if ok else part)       # raw buffer stored as is
    got = _decode_gpu(comps, 1 << 16)
    assert np.array_equal(got, np.concatenate(parts))


def test_decode_hostile_overlapping_and_long_matches():
    # hand-made packets: matches whose source overlaps the bytes being written and lengths up
    # to 255 -- the reference reads the whole string from the old window first
    rng = np.random.default_rng(6)
    npk = 16
    packets, sizes = [], []
    for p in range(npk):
        body = bytearray()
        produced = 0
        while produced < 4096:
            flags = 0
            group = bytearray()
            for bit in range(8):
                if produced >= 4096:
                    break
                if rng.random() < 0.5:
                    flags |= 1 << bit
                    group.append(int(rng.integers(0, 256)))
                    produced += 1
                else:
                    ln = int(min(rng.integers(3, 256), 4096 - produced))
                    if ln < 3:
                        flags |= 1 << bit
                        group.append(0x41)
                        produced += 1
                        continue
                    group.append(ln)
                    group.append(int(rng.integers(0, 256)))
                    produced += ln
            body.append(flags)
            body += group
        packets.append(bytes(body))
        sizes.append(len(body))
    trailer = b"".join(int(s).to_bytes(2, "big") for s in sizes) + (npk * 4096).to_bytes(4, "big") + b"\0\0"
    comp = np.frombuffer(b"".join(packets) + trailer, np.uint8).copy()
    dok, want = O.culzss_oracle_decompress(comp, npk * 4096)
    assert dok
    got = _decode_gpu([comp], npk * 4096)
    assert np.array_equal(got, want)


@pytest.mark.skipif(not O.have_ref("culzss"), reason="oracle/_ref/libref_culzss.so not built")
def test_reference_kernels_agree_with_oracle_and_product():
    """Runs the reference's own EncodeKernel / aftercomp / DecodeKernel on this GPU."""
    ref = O.ref_culzss()
    ref.initGPU()
    in_d = ref.initGPUmem(MIB)
    out_d = ref.initGPUmem(2 * MIB)
    try:
        for name in ("quant32", "text", "spaces", "mix"):
            data = _cases()[name]
            buf = np.zeros(MIB + MIB // 8 + 1024, np.uint8)
            buf[:MIB] = data
            tokens = np.zeros(2 * MIB, np.uint8)
            assert ref.compression_kernel_wrapper(buf, MIB, tokens, 0, 0, 128, 0, 0, in_d, out_d) == 1
            ref.onestream_finish_GPU(0)
            torch.cuda.synchronize()
            assert np.array_equal(tokens, O.culzss_oracle_tokens(data)), name
            clen = C.c_int(0)
            assert ref.aftercompression_wrapper(buf, MIB, tokens, C.byref(clen)) == 1
            ref_comp = buf[: clen.value].copy()
            bufs, _ = _gpu_encode(data)
            assert np.array_equal(bufs[0], ref_comp), name
            # reference decoder on the product's output, in place
            work = np.zeros(MIB + MIB // 8 + 1024, np.uint8)
            work[: ref_comp.size] = bufs[0]
            dlen = C.c_int(0)
            assert ref.decompression_kernel_wrapper(work, int(ref_comp.size), C.byref(dlen), 0, 0, 1) == 1
            assert dlen.value == MIB and np.array_equal(work[:MIB], data), name
    finally:
        ref.deleteGPUmem(in_d)
        ref.deleteGPUmem(out_d)


def test_reference_named_wrappers_drop_in():
    """The call protocol of culzss.c:108,170,176 / deculzss.c:98 against libb200lc.so."""
    L = b200lc.lib()
    vp = C.c_void_p
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    L.initGPUmem.restype = vp
    L.initGPUmem.argtypes = [C.c_int]
    L.deleteGPUmem.argtypes = [vp]
    L.compression_kernel_wrapper.restype = C.c_int
    L.compression_kernel_wrapper.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, vp, vp]
    L.aftercompression_wrapper.restype = C.c_int
    L.aftercompression_wrapper.argtypes = [u8p, C.c_int, u8p, C.POINTER(C.c_int)]
    L.decompression_kernel_wrapper.restype = C.c_int
    L.decompression_kernel_wrapper.argtypes = [u8p, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
    L.onestream_finish_GPU.argtypes = [C.c_int]
    L.initGPU()
    in_d, out_d = L.initGPUmem(MIB), L.initGPUmem(2 * MIB)
    try:
        for index, name in enumerate(["quant32", "text", "random", "zeros"]):
            data = _cases()[name]
            buf = np.zeros(2 * MIB, np.uint8)
            buf[:MIB] = data
            bufout = np.zeros(2 * MIB, np.uint8)
            assert L.compression_kernel_wrapper(buf, MIB, bufout, 0, 0, 128, 0, index, in_d, out_d) == 1
            assert L.onestream_finish_GPU(index) == 1
            clen = C.c_int(0)
            rc = L.aftercompression_wrapper(buf, MIB, bufout, C.byref(clen))
            ok, want = O.culzss_oracle_compress(data)
            assert rc == ok
            if not ok:
                continue
            assert clen.value == want.size and np.array_equal(buf[: want.size], want)
            dlen = C.c_int(0)
            assert L.decompression_kernel_wrapper(buf, clen.value, C.byref(dlen), 0, 0, 1) == 1
            assert dlen.value == MIB and np.array_equal(buf[:MIB], data)
    finally:
        L.deleteGPUmem(in_d)
        L.deleteGPUmem(out_d)
