"""CPU: the lane-serial CULZSS fast-mode encoder (csrc/culzss_lane.cuh, the code a GPU lane runs in
culzss_encode_lane_kernel) compiled for the host with STRIDE = 1.  NON-PARITY mode: what is pinned
is the FORMAT -- every packet decodes to its input with the oracle's restatement of the reference
DecodeKernel (gpu_decompress.cu:164-242) -- plus the format's limits (3 <= length <= 124 <= 127,
packet <= 4608 bytes) and a floor on the compression ratio per data kind."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKT, SLOT = 4096, 4608


@pytest.fixture(scope="module")
def lane(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lane") / "liblane.so")
    r = subprocess.run(["g++", "-O2", "-std=c++14", "-shared", "-fPIC",
                        "-I", os.path.join(ROOT, "gpu-lossless-compression_b200", "csrc"),
                        os.path.join(ROOT, "tests", "c", "culzss_lane_host.cc"), "-o", so],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(so)
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
    lib.lane_encode_packets.argtypes = [u8p, C.c_uint, u8p, u16p, u8p]
    lib.lane_encode_packets.restype = None
    return lib


def _encode(lane, data):
    npk = data.size // PKT
    out = np.zeros(npk * SLOT, np.uint8)
    sizes = np.zeros(npk, np.uint16)
    last = np.zeros(npk, np.uint8)
    lane.lane_encode_packets(data, npk, out, sizes, last)
    return out, sizes.astype(np.int64), last.astype(np.int64)


def _check_tokens(body):
    """walks one packet: match lengths within the limits, 4096 bytes produced; returns the offset of
    the last flag byte"""
    i, produced, last_flag = 0, 0, 0
    while i < len(body):
        last_flag = i
        flags = body[i]
        i += 1
        for bit in range(8):
            if i >= len(body):
                assert flags >> bit == 0
                break
            if flags >> bit & 1:
                i += 1
                produced += 1
            else:
                assert 3 <= body[i] <= 124 and body[i + 1] < 128
                produced += body[i]
                i += 2
    assert produced == PKT
    return last_flag


CASES = {
    "quant32": (lambda n: O.quant_codes(n), 1.85),
    "quant16": (lambda n: O.quant_codes(n, dtype=np.uint16), 1.30),
    "text": (lambda n: np.frombuffer((b"the quick brown fox jumps over the lazy dog. " * (n // 45 + 1))[:n], np.uint8).copy(), 25.0),
    "random": (lambda n: np.random.default_rng(1).integers(0, 256, n, dtype=np.uint8), 0.88),
    "zeros": (lambda n: np.zeros(n, np.uint8), 35.0),
    "spaces": (lambda n: np.full(n, 0x20, np.uint8), 40.0),
    "ramp": (lambda n: (np.arange(n) % 97).astype(np.uint8), 15.0),
    "period3": (lambda n: (np.arange(n) % 3).astype(np.uint8), 30.0),
    "zipf": (lambda n: O.zipf_bytes(n, 1.1, 12345)[:n].copy(), 0.88),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_lane_encoder_output_decodes_with_the_oracle_decoder(lane, name):
    gen, floor = CASES[name]
    n = 64 * PKT
    data = gen(n)
    out, sizes, last = _encode(lane, data)
    orc = O.oracle()
    for k in range(n // PKT):
        sz = int(sizes[k])
        assert 0 < sz <= SLOT
        comp = np.ascontiguousarray(out[k * SLOT: k * SLOT + sz])
        dec = np.zeros(PKT, np.uint8)
        assert orc.culzss_oracle_decode_packet(comp, sz, dec, PKT) == PKT
        assert np.array_equal(dec, data[k * PKT: (k + 1) * PKT]), (name, k)
        assert sz - _check_tokens(comp.tolist()) == last[k]
    assert n / (sizes.sum() + 2 * (n // PKT)) >= floor


def test_lane_encoder_packets_are_independent(lane):
    """a packet's bytes depend on that packet only (the window restarts with spaces)"""
    a = O.quant_codes(8 * PKT, seed=7)
    out, sizes, _ = _encode(lane, a)
    b = np.ascontiguousarray(a[3 * PKT: 5 * PKT])
    out_b, sizes_b, _ = _encode(lane, b)
    for j in range(2):
        assert sizes_b[j] == sizes[3 + j]
        assert np.array_equal(out_b[j * SLOT: j * SLOT + sizes_b[j]], out[(3 + j) * SLOT: (3 + j) * SLOT + sizes[3 + j]])
