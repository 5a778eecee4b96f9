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


def _gpu_encode(data, buf_length=MIB, kernel=0):
    d = torch.from_numpy(data).to(DEV)
    out, clen = b200lc.culzss_encode(d, buf_length, kernel=kernel)
    torch.cuda.synchronize()
    stride = b200lc.culzss_out_stride(buf_length)
    out = out.cpu().numpy()
    clen = clen.cpu().numpy()
    return [out[b * stride: b * stride + clen[b]] for b in range(clen.size)], clen


# both parity kernels (a CTA per packet / a packet per lane, b200lc_culzss_encode_batch_ex) must write
# the reference's bytes
KERNELS = [b200lc.CULZSS_KERNEL_CTA, b200lc.CULZSS_KERNEL_LANE]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", list(_cases().keys()))
def test_encode_matches_oracle(name, kernel):
    data = _cases()[name]
    ok, want = O.culzss_oracle_compress(data)
    bufs, clen = _gpu_encode(data, kernel=kernel)
    if not ok:
        assert clen[0] == 0
        return
    assert clen[0] == want.size
    assert np.array_equal(bufs[0], want)


def test_encode_kernels_agree_on_a_large_batch_and_auto_picks_the_lane_kernel():
    """168 MiB (43008 packets: above the AUTO threshold of 40960) of mixed data: CTA kernel == lane
    kernel == b200lc_culzss_encode_batch byte for byte; three buffers checked against the oracle."""
    rng = np.random.default_rng(21)
    nbuf = 168
    distinct = []
    for b in range(12):
        kind = b % 4
        if kind == 0:
            distinct.append(O.quant_codes(MIB, seed=100 + b))
        elif kind == 1:
            distinct.append(O.quant_codes(MIB, seed=100 + b, dtype=np.uint16))
        elif kind == 2:
            distinct.append(rng.integers(0, 3, MIB, dtype=np.uint8))
        else:
            distinct.append(np.frombuffer((b"lorem ipsum dolor sit amet %d " % b) * 40000, np.uint8)[:MIB].copy())
    parts = [distinct[b % 12] for b in range(nbuf)]
    data = np.concatenate(parts)
    d = torch.from_numpy(data).to(DEV)
    outs = []
    for kernel in (b200lc.CULZSS_KERNEL_CTA, b200lc.CULZSS_KERNEL_LANE, b200lc.CULZSS_KERNEL_AUTO):
        out, clen = b200lc.culzss_encode(d, MIB, kernel=kernel)
        torch.cuda.synchronize()
        outs.append((out.cpu().numpy(), clen.cpu().numpy()))
    stride = b200lc.culzss_out_stride(MIB)
    for out, clen in outs[1:]:
        assert np.array_equal(clen, outs[0][1])
        for b in range(nbuf):
            assert np.array_equal(out[b * stride: b * stride + clen[b]], outs[0][0][b * stride: b * stride + clen[b]]), b
    for b in (0, 1, 3):
        ok, want = O.culzss_oracle_compress(parts[b])
        assert ok and np.array_equal(outs[1][0][b * stride: b * stride + outs[1][1][b]], want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_encode_small_and_multi_buffer(kernel):
    # 64 KiB buffers (16 packets) and a 5-buffer batch with one expanding buffer in the middle
    rng = np.random.default_rng(4)
    parts = [O.quant_codes(1 << 16, seed=s) for s in (1, 2)] + \
            [rng.integers(0, 256, 1 << 16, dtype=np.uint8)] + \
            [np.zeros(1 << 16, np.uint8), O.quant_codes(1 << 16, seed=9, dtype=np.uint16)]
    data = np.concatenate(parts)
    bufs, clen = _gpu_encode(data, 1 << 16, kernel=kernel)
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


def _container_api():
    L = b200lc.lib()
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    L.b200lc_culzss_container_bound.restype = C.c_size_t
    L.b200lc_culzss_container_bound.argtypes = [C.c_size_t]
    for f in (L.b200lc_culzss_compress_container, L.b200lc_culzss_decompress_container):
        f.restype = C.c_int
        f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
    return L


def test_file_container_matches_reference_cli(tmp_path):
    """b200lc_culzss_compress_container vs the reference's own command-line program
    (oracle/_ref/culzss_main = main.c + culzss.c + deculzss.c + decompression.c + its kernels,
    sm_100a) on a 5 MiB file with one incompressible (raw) buffer in the middle."""
    import os
    import subprocess
    rng = np.random.default_rng(8)
    data = np.concatenate([O.quant_codes(2 * MIB, seed=1), rng.integers(0, 256, MIB, dtype=np.uint8),
                           _cases()["text"], O.quant_codes(MIB, seed=2, dtype=np.uint16)])
    L = _container_api()
    cap = L.b200lc_culzss_container_bound(data.size)
    out = np.zeros(cap, np.uint8)
    olen = C.c_size_t(0)
    assert L.b200lc_culzss_compress_container(data, data.size, out, cap, C.byref(olen)) == 0
    mine = out[: olen.value].copy()
    hdr = mine[: 8 + 4 * 5].view(np.uint32)
    assert hdr[0] == 5 and hdr[1] == 0
    sizes = np.diff(np.concatenate([[0], hdr[2:7]]))
    assert sizes[2] == MIB                       # the random buffer is stored raw
    back = np.zeros(data.size, np.uint8)
    blen = C.c_size_t(0)
    assert L.b200lc_culzss_decompress_container(mine, mine.size, back, back.size, C.byref(blen)) == 0
    assert blen.value == data.size and np.array_equal(back, data)

    # the container built from the oracle's per-buffer output is the same
    want = [np.array([5, 0], np.uint32).view(np.uint8)]
    bufs, cum = [], 0
    for b in range(5):
        part = data[b * MIB:(b + 1) * MIB]
        ok, comp = O.culzss_oracle_compress(part)
        bufs.append(comp if ok else part)
        cum += bufs[-1].size
        want.append(np.array([cum], np.uint32).view(np.uint8))
    assert np.array_equal(mine, np.concatenate(want + bufs))

    exe = os.path.join(O.ORACLE_DIR, "_ref", "culzss_main")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/culzss_main not built")
    # The reference CLI shares ONE device input/output buffer between its four in-flight queue
    # slots (culzss.c:108 passes fifo->in_d / fifo->out_d for every slot) while each slot runs on
    # its own streams, so with more than one buffer in flight its output is racy (on this B200 it
    # "compresses" the random buffer to 277 KB).  A single-buffer file has no such race: compare
    # that container byte for byte, and let the reference CLI decode our 5-buffer container.
    one = np.ascontiguousarray(data[:MIB])
    src, ref_out, mine_file, mine_back = (tmp_path / n for n in ("in.bin", "ref.lz", "mine.lz", "mine.back"))
    src.write_bytes(one.tobytes())
    r = subprocess.run([exe, "-i", str(src), "-o", str(ref_out)], capture_output=True, text=True, timeout=300)
    assert ref_out.exists(), r.stdout + r.stderr
    ref = np.frombuffer(ref_out.read_bytes(), np.uint8)
    out1 = np.zeros(cap, np.uint8)
    assert L.b200lc_culzss_compress_container(one, one.size, out1, cap, C.byref(olen)) == 0
    assert ref.size == olen.value and np.array_equal(ref, out1[: olen.value])   # byte-identical file
    mine_file.write_bytes(mine.tobytes())
    subprocess.run([exe, "-d", "1", "-i", str(mine_file), "-o", str(mine_back)], capture_output=True,
                   text=True, timeout=300)
    assert mine_back.read_bytes() == data.tobytes()


def test_file_container_ragged_size_round_trips():
    L = _container_api()
    data = O.quant_codes(3 * MIB + 12345 * 4, seed=6)
    cap = L.b200lc_culzss_container_bound(data.size)
    out = np.zeros(cap, np.uint8)
    olen = C.c_size_t(0)
    assert L.b200lc_culzss_compress_container(data, data.size, out, cap, C.byref(olen)) == 0
    hdr = out[:8].view(np.uint32)
    assert hdr[0] == 4 and hdr[1] == MIB - 12345 * 4
    back = np.zeros(data.size, np.uint8)
    blen = C.c_size_t(0)
    assert L.b200lc_culzss_decompress_container(out[: olen.value].copy(), olen.value, back, back.size,
                                                C.byref(blen)) == 0
    assert blen.value == data.size and np.array_equal(back, data)
    small = np.zeros(1000, np.uint8)
    assert L.b200lc_culzss_compress_container(small, small.size, out, cap, C.byref(olen)) == b200lc.ERR_UNSUPPORTED


# ------------------------------------------------------------------------------------------ fast mode
# NON-PARITY: b200lc_culzss_encode_fast_batch writes the reference's format with other matches.
# What is pinned: every stream decodes to the input with the oracle's decoder, the product decoder
# and the reference's own DecodeKernel; tokens obey the format's limits; the size is within a
# stated factor of the parity encoder's.
def _fast_encode(data, depth, buf_length=MIB):
    d = torch.from_numpy(data).to(DEV)
    out, clen = b200lc.culzss_encode(d, buf_length, fast=depth)
    torch.cuda.synchronize()
    stride = b200lc.culzss_out_stride(buf_length)
    out = out.cpu().numpy()
    clen = clen.cpu().numpy()
    return [out[b * stride: b * stride + clen[b]] for b in range(clen.size)], clen


def _walk_tokens(comp, n):
    """Parses one compressed buffer: every match has 3 <= len <= 127 and packets produce exactly
    4096 bytes.  Returns the number of matches."""
    npk = n // 4096
    sizes = [int.from_bytes(comp[comp.size - 6 - 2 * npk + 2 * i: comp.size - 6 - 2 * npk + 2 * i + 2].tobytes(), "big")
             for i in range(npk)]
    assert int.from_bytes(comp[-6:-2].tobytes(), "big") == n
    at, matches = 0, 0
    for sz in sizes:
        body = comp[at: at + sz].tolist()
        at += sz
        i, produced = 0, 0
        while i < len(body):
            flags = body[i]
            i += 1
            for bit in range(8):
                if i >= len(body):
                    break
                if flags >> bit & 1:
                    i += 1
                    produced += 1
                else:
                    ln = body[i]
                    assert 3 <= ln <= 127
                    i += 2
                    produced += ln
                    matches += 1
        assert produced == 4096
    assert at == comp.size - 6 - 2 * npk
    return matches


@pytest.mark.parametrize("depth", [1, 2, 4, "lane"])
@pytest.mark.parametrize("name", ["quant32", "quant16", "spaces", "zeros", "text", "mix", "carets", "ramp"])
def test_fast_mode_streams_decode_everywhere(name, depth):
    data = _cases()[name]
    bufs, clen = _fast_encode(data, depth)
    assert clen[0] > 0
    comp = bufs[0]
    dok, back = O.culzss_oracle_decompress(comp, data.size)
    assert dok and np.array_equal(back, data)
    assert np.array_equal(_decode_gpu([comp], MIB), data)
    ok, parity = O.culzss_oracle_compress(data)
    assert ok and comp.size <= 1.6 * parity.size + 4096, (comp.size, parity.size)
    if name in ("quant32", "text"):
        assert _walk_tokens(comp[: comp.size], MIB) > 0


def test_fast_mode_incompressible_buffer_is_stored_raw_and_small_buffers():
    rng = np.random.default_rng(12)
    parts = [O.quant_codes(1 << 16, seed=5), rng.integers(0, 256, 1 << 16, dtype=np.uint8), np.full(1 << 16, 0x20, np.uint8)]
    data = np.concatenate(parts)
    bufs, clen = _fast_encode(data, 2, 1 << 16)
    assert clen[1] == 0 and clen[0] > 0 and clen[2] > 0
    comps = [bufs[0], parts[1], bufs[2]]
    assert np.array_equal(_decode_gpu(comps, 1 << 16), data)
    with pytest.raises(b200lc.B200LCError):
        b200lc.culzss_encode(torch.from_numpy(data).to(DEV), 1 << 16, fast=3)      # unsupported depth


@pytest.mark.skipif(not O.have_ref("culzss"), reason="oracle/_ref/libref_culzss.so not built")
@pytest.mark.parametrize("depth", [1, 2, 4, "lane"])
def test_fast_mode_streams_decode_with_the_reference_kernel(depth):
    """The reference's own DecodeKernel (gpu_decompress.cu:164-242, sm_100a build) on fast-mode output."""
    ref = O.ref_culzss()
    ref.initGPU()
    for name in ("quant32", "text", "mix", "spaces"):
        data = _cases()[name]
        bufs, _ = _fast_encode(data, depth)
        work = np.zeros(MIB + MIB // 8 + 1024, np.uint8)
        work[: bufs[0].size] = bufs[0]
        dlen = C.c_int(0)
        assert ref.decompression_kernel_wrapper(work, int(bufs[0].size), C.byref(dlen), 0, 0, 1) == 1
        assert dlen.value == MIB and np.array_equal(work[:MIB], data), name


def test_fast_lane_mode_behind_the_container_by_environment():
    """B200LC_CULZSS_FAST=lane switches the container writer (and the reference-named wrappers) to
    the NON-PARITY lane encoder: the file still decodes -- with the product decoder and, per buffer,
    with the oracle's restatement of the reference decoder -- but is not the parity file.  The
    switch is read once per process, hence the child process."""
    import os
    import subprocess
    import sys
    import textwrap
    code = textwrap.dedent("""
        import ctypes as C, sys, numpy as np
        sys.path.insert(0, %r)
        import oracle_lib as O
        from pkg import b200lc
        L = b200lc.lib()
        u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        L.b200lc_culzss_container_bound.restype = C.c_size_t
        L.b200lc_culzss_container_bound.argtypes = [C.c_size_t]
        for f in (L.b200lc_culzss_compress_container, L.b200lc_culzss_decompress_container):
            f.restype = C.c_int
            f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        MIB = 1 << 20
        data = np.concatenate([O.quant_codes(2 * MIB, seed=4), np.zeros(MIB, np.uint8)])
        cap = L.b200lc_culzss_container_bound(data.size)
        out = np.zeros(cap, np.uint8); olen = C.c_size_t(0)
        assert L.b200lc_culzss_compress_container(data, data.size, out, cap, C.byref(olen)) == 0
        back = np.zeros(data.size, np.uint8); blen = C.c_size_t(0)
        assert L.b200lc_culzss_decompress_container(out[: olen.value].copy(), olen.value, back, back.size, C.byref(blen)) == 0
        assert blen.value == data.size and np.array_equal(back, data)
        hdr = out[: 8 + 12].view(np.uint32)
        ends = hdr[2:5].astype(np.int64)
        first = out[20: 20 + ends[0]].copy()
        ok, dec = O.culzss_oracle_decompress(first, MIB)
        assert ok and np.array_equal(dec, data[:MIB])
        print("SIZE", olen.value)
    """ % os.path.dirname(os.path.abspath(__file__)))
    sizes = {}
    for mode in ("", "lane"):
        env = dict(os.environ)
        env.pop("B200LC_CULZSS_FAST", None)
        if mode:
            env["B200LC_CULZSS_FAST"] = mode
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        sizes[mode] = int(r.stdout.split("SIZE")[1])
    assert sizes["lane"] != sizes[""] and sizes["lane"] < 1.4 * sizes[""]
