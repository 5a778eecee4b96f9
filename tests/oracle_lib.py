"""ctypes bindings for the CPU checkers under oracle/ (test infrastructure only).

`oracle()`  -> oracle/liboracle.so   (plain-C restatements, built by `make -C oracle oracle`)
`ref_cuhd()` etc. -> oracle/_ref/libref_*.so (the reference's own CPU sources; prebuilt here, they
travel to the GPU box with the snapshot).  Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")

_cache = {}


def build_oracle():
    """(Re)build oracle/liboracle.so and, when /root/reference is present, oracle/_ref/."""
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)


def oracle():
    if "oracle" in _cache:
        return _cache["oracle"]
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    lib.cuhd_oracle_decode.restype = C.c_size_t
    lib.cuhd_oracle_decode.argtypes = [_u32p, C.c_size_t, _u8p, C.c_int, _u8p, C.c_size_t]
    lib.cuhd_oracle_canonical.restype = None
    lib.cuhd_oracle_canonical.argtypes = [_u8p, _u8p, C.c_size_t, _u32p, _u8p]
    lib.cuhd_oracle_build_lut.restype = None
    lib.cuhd_oracle_build_lut.argtypes = [_u32p, _u8p, C.c_int, _u8p]
    lib.cuhd_oracle_compressed_units.restype = C.c_size_t
    lib.cuhd_oracle_compressed_units.argtypes = [_u64p, _u8p]
    lib.cuhd_oracle_encode.restype = C.c_size_t
    lib.cuhd_oracle_encode.argtypes = [_u8p, C.c_size_t, _u32p, _u8p, _u32p, C.c_size_t,
                                       C.POINTER(C.c_size_t)]
    lib.culzss_oracle_packet_tokens.restype = None
    lib.culzss_oracle_packet_tokens.argtypes = [_u8p, _u8p]
    lib.culzss_oracle_buffer_tokens.restype = None
    lib.culzss_oracle_buffer_tokens.argtypes = [_u8p, C.c_int, _u8p]
    lib.culzss_oracle_aftercomp.restype = C.c_int
    lib.culzss_oracle_aftercomp.argtypes = [_u8p, C.c_int, _u8p, C.POINTER(C.c_int)]
    lib.culzss_oracle_decode_packet.restype = C.c_int
    lib.culzss_oracle_decode_packet.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
    lib.culzss_oracle_decode_buffer.restype = C.c_int
    lib.culzss_oracle_decode_buffer.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.POINTER(C.c_int)]
    _i32a = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    lib.cudpp_oracle_sa.restype = None
    lib.cudpp_oracle_sa.argtypes = [_u8p, C.c_uint32, _u32p]
    lib.cudpp_oracle_bwt.restype = None
    lib.cudpp_oracle_bwt.argtypes = [_u8p, C.c_uint32, _u8p, C.POINTER(C.c_int)]
    lib.cudpp_oracle_mtf.restype = None
    lib.cudpp_oracle_mtf.argtypes = [_u8p, C.c_uint32, _u8p]
    lib.cudpp_oracle_tree.restype = C.c_int
    lib.cudpp_oracle_tree.argtypes = [_u32p, _i32a, _i32a, _i32a, _i32a]
    lib.cudpp_oracle_codes.restype = None
    lib.cudpp_oracle_codes.argtypes = [_i32a, _i32a, _i32a, _i32a, C.c_int, _u64p, _u8p]
    lib.cudpp_oracle_huffman.restype = C.c_int
    lib.cudpp_oracle_huffman.argtypes = [_u8p, C.c_uint32, _u32p, _u32p, C.POINTER(C.c_uint32), _u32p,
                                         C.c_uint32]
    lib.cudpp_oracle_compress.restype = C.c_int
    lib.cudpp_oracle_compress.argtypes = [_u8p, C.c_uint32, C.POINTER(C.c_int), _u32p, _u32p,
                                          C.POINTER(C.c_uint32), _u32p, C.c_uint32]
    lib.cudpp_oracle_decompress.restype = C.c_int
    lib.cudpp_oracle_decompress.argtypes = [_u8p, C.c_uint32, C.c_int, _u32p, _u32p, _u32p]
    lib.bzip2_oracle_rotation_order.restype = None
    lib.bzip2_oracle_rotation_order.argtypes = [_u8p, C.c_uint32, _u32p]
    lib.bzip2_oracle_block_sort.restype = C.c_int
    lib.bzip2_oracle_block_sort.argtypes = [_u8p, C.c_uint32, _u32p, _u32p, _u32p]
    lib.bzip2_oracle_merge.restype = C.c_int
    lib.bzip2_oracle_merge.argtypes = [_u8p, C.c_int, C.c_int, _u32p, _u32p, _u32p, _u32p]
    _u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
    lib.bzip2_oracle_mtf_rle.restype = C.c_int
    lib.bzip2_oracle_mtf_rle.argtypes = [_u8p, _u32p, C.c_int, _u8p, _u16p, _i32p, C.POINTER(C.c_int)]
    lib.bzip2_oracle_send_mtf.restype = C.c_int
    lib.bzip2_oracle_send_mtf.argtypes = [_u16p, C.c_int, _i32p, _u8p, C.c_int, _u8p, C.POINTER(C.c_uint64), _u8p,
                                          _u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.bzip2_oracle_code_lengths.restype = None
    lib.bzip2_oracle_code_lengths.argtypes = [_u8p, _i32p, C.c_int, C.c_int]
    lib.bsc_oracle_bwt_encode.restype = C.c_int
    lib.bsc_oracle_bwt_encode.argtypes = [_u8p, C.c_int, _u8p, _u8p, _i32p]
    _cache["oracle"] = lib
    return lib


def have_ref(name):
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_%s.so" % name))


def ref_cuhd():
    if "ref_cuhd" in _cache:
        return _cache["ref_cuhd"]
    lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_cuhd.so"))
    lib.ref_cuhd_encode.restype = C.c_int
    lib.ref_cuhd_encode.argtypes = [_u8p, C.c_size_t, _u32p, _u8p, _u8p, _u32p, C.c_size_t,
                                    C.POINTER(C.c_size_t)]
    lib.ref_cuhd_encode_with_table.restype = C.c_int
    lib.ref_cuhd_encode_with_table.argtypes = [_u8p, C.c_size_t, _u32p, _u8p, _u32p, C.c_size_t]
    lib.ref_cuhd_max_codeword_length.restype = C.c_int
    _cache["ref_cuhd"] = lib
    return lib


# ---------------------------------------------------------------------------- CUHD helpers
def cuhd_ref_encode(data):
    """Reference llhuff on `data` -> (code[256] u32, len[256] u8, lut[2048,2] u8, units u32)."""
    lib = ref_cuhd()
    n = data.size
    code = np.zeros(256, np.uint32)
    length = np.zeros(256, np.uint8)
    lut = np.zeros((1 << 11) * 2, np.uint8)
    cap = n + 16  # codes are <= 11 bits -> never more than n units
    units = np.zeros(cap, np.uint32)
    nu = C.c_size_t(0)
    rc = lib.ref_cuhd_encode(np.ascontiguousarray(data), n, code, length, lut, units, cap,
                             C.byref(nu))
    if rc != 0:
        raise RuntimeError("reference llhuff refused input (rc=%d)" % rc)
    return code, length, lut.reshape(-1, 2), units[: nu.value].copy()


def cuhd_oracle_encode(data, code, length):
    lib = oracle()
    hist = np.bincount(data, minlength=256).astype(np.uint64)
    nu = lib.cuhd_oracle_compressed_units(hist, length)
    units = np.zeros(max(nu, 1), np.uint32)
    defined = C.c_size_t(0)
    wrote = lib.cuhd_oracle_encode(np.ascontiguousarray(data), data.size, code, length, units,
                                   units.size, C.byref(defined))
    assert wrote == nu, (wrote, nu)
    return units[:nu], defined.value


def cuhd_oracle_decode(units, lut, n_out, max_len=11):
    lib = oracle()
    out = np.zeros(n_out, np.uint8)
    got = lib.cuhd_oracle_decode(np.ascontiguousarray(units), units.size,
                                 np.ascontiguousarray(lut).reshape(-1), max_len, out, n_out)
    return out, got


def cuhd_oracle_lut(code, length, max_len=11):
    lib = oracle()
    lut = np.zeros((1 << max_len) * 2, np.uint8)
    lib.cuhd_oracle_build_lut(code, length, max_len, lut)
    return lut.reshape(-1, 2)


# ---------------------------------------------------------------------------- generators
def zipf_bytes(n, alpha=1.1, seed=12345, nsym=256):
    """Synthetic C2 input: symbol k with P ~ 1/(k+1)^alpha, inverse-CDF sampling:
    out[i] = min(searchsorted(cdf, u[i]), nsym - 1), u = Generator(MT19937(seed)).random(n).
    Evaluated in chunks through a 65536-bucket table of the inverse CDF (a bucket that lies inside
    one symbol's interval needs no search) -- same values as the plain formula, 10x faster, so that
    the full 1 GiB input can be produced inside a test."""
    rng = np.random.Generator(np.random.MT19937(seed))
    p = 1.0 / np.arange(1, nsym + 1, dtype=np.float64) ** alpha
    cdf = np.cumsum(p / p.sum())
    edges = np.arange(65537, dtype=np.float64) / 65536.0
    lo = np.searchsorted(cdf, edges[:-1])
    hi = np.searchsorted(cdf, np.nextafter(edges[1:], 0.0))
    direct = np.minimum(lo, nsym - 1).astype(np.uint8)
    ambiguous = lo != hi
    out = np.empty(n, np.uint8)
    step = 1 << 24
    for at in range(0, n, step):
        u = rng.random(min(step, n - at))
        b = (u * 65536.0).astype(np.int32)
        part = direct[b]
        m = ambiguous[b]
        part[m] = np.minimum(np.searchsorted(cdf, u[m]), nsym - 1).astype(np.uint8)
        out[at:at + part.size] = part
    return out


def limited_lengths(hist, max_len=11):
    """Test-side length-limited Huffman lengths (heap Huffman + Kraft repair), used only to make
    decodable streams when oracle/_ref is unavailable or when a test wants a specific table."""
    import heapq
    syms = [s for s in range(len(hist)) if hist[s] > 0]
    length = np.zeros(256, np.uint8)
    if len(syms) == 1:
        length[syms[0]] = 1
        return length
    heap = [(int(hist[s]), i, [s]) for i, s in enumerate(syms)]
    heapq.heapify(heap)
    depth = {s: 0 for s in syms}
    tie = len(heap)
    while len(heap) > 1:
        a = heapq.heappop(heap)
        b = heapq.heappop(heap)
        for s in a[2] + b[2]:
            depth[s] += 1
        heapq.heappush(heap, (a[0] + b[0], tie, a[2] + b[2]))
        tie += 1
    lens = {s: min(d, max_len) for s, d in depth.items()}
    kraft = sum(2 ** (max_len - l) for l in lens.values())
    order = sorted(syms, key=lambda s: (hist[s], s))
    while kraft > 2 ** max_len:          # lengthen the rarest symbols that still can grow
        for s in order:
            if lens[s] < max_len:
                kraft -= 2 ** (max_len - lens[s] - 1)
                lens[s] += 1
                break
    for s, l in lens.items():
        length[s] = l
    return length


def canonical_from_lengths(length):
    """Canonical codes in (length, symbol) order via the oracle's restatement of
    llhuffman_encoder.cc:183-195."""
    syms = np.array(sorted([s for s in range(256) if length[s]], key=lambda s: (length[s], s)),
                    np.uint8)
    lens = np.ascontiguousarray(length[syms])
    code = np.zeros(256, np.uint32)
    out_len = np.zeros(256, np.uint8)
    oracle().cuhd_oracle_canonical(syms, lens, syms.size, code, out_len)
    return code, out_len


def cuhd_make_case(data, max_len=11, use_ref=True):
    """-> (code, len, lut[1<<max_len, 2], units incl. one zero pad unit) for `data`."""
    if use_ref and max_len == 11 and have_ref("cuhd") and np.unique(data).size > 1:
        code, length, lut, _ = cuhd_ref_encode(data)
    else:
        hist = np.bincount(data, minlength=256)
        length = limited_lengths(hist, max_len)
        code, length = canonical_from_lengths(length)
        lut = cuhd_oracle_lut(code, length, max_len)
    units, _ = cuhd_oracle_encode(data, code, length)
    units = np.concatenate([units, np.zeros(1, np.uint32)])
    return code, length, lut, units


# ---------------------------------------------------------------------------- CULZSS helpers
CULZSS_PACKET = 4096
CULZSS_BUFFER = 1 << 20


def ref_culzss():
    """The reference's gpu_compress.cu + gpu_decompress.cu (oracle/_ref/libref_culzss.so).
    aftercompression_wrapper is pure CPU code; the kernel wrappers need a GPU."""
    if "ref_culzss" in _cache:
        return _cache["ref_culzss"]
    lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_culzss.so"))
    vp = C.c_void_p
    lib.aftercompression_wrapper.restype = C.c_int
    lib.aftercompression_wrapper.argtypes = [_u8p, C.c_int, _u8p, C.POINTER(C.c_int)]
    lib.compression_kernel_wrapper.restype = C.c_int
    lib.compression_kernel_wrapper.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, vp, vp]
    lib.decompression_kernel_wrapper.restype = C.c_int
    lib.decompression_kernel_wrapper.argtypes = [_u8p, C.c_int, C.POINTER(C.c_int), C.c_int,
                                                 C.c_int, C.c_int]
    lib.initGPUmem.restype = vp
    lib.initGPUmem.argtypes = [C.c_int]
    lib.deleteGPUmem.argtypes = [vp]
    lib.initGPU.restype = None
    lib.onestream_finish_GPU.restype = C.c_int
    lib.onestream_finish_GPU.argtypes = [C.c_int]
    _cache["ref_culzss"] = lib
    return lib


def culzss_oracle_tokens(buf):
    out = np.zeros(buf.size * 2, np.uint8)
    oracle().culzss_oracle_buffer_tokens(np.ascontiguousarray(buf), buf.size, out)
    return out


def culzss_oracle_aftercomp(tokens, buf_length):
    """-> (ok, compressed bytes incl. trailer)"""
    out = np.zeros(buf_length + buf_length // 8 + 1024, np.uint8)
    n = C.c_int(0)
    ok = oracle().culzss_oracle_aftercomp(tokens, buf_length, out, C.byref(n))
    return ok, out[: n.value].copy()


def culzss_ref_aftercomp(tokens, data):
    """Reference aftercompression_wrapper (CPU) -> (ok, compressed bytes incl. trailer)."""
    n = data.size
    buf = np.zeros(n + n // 8 + 1024, np.uint8)   # the reference writes in place, may overrun
    buf[:n] = data
    clen = C.c_int(0)
    ok = ref_culzss().aftercompression_wrapper(buf, n, np.ascontiguousarray(tokens), C.byref(clen))
    return ok, buf[: clen.value].copy()


def culzss_oracle_compress(buf):
    tokens = culzss_oracle_tokens(buf)
    return culzss_oracle_aftercomp(tokens, buf.size)


def culzss_oracle_decompress(comp, out_cap=CULZSS_BUFFER):
    out = np.zeros(out_cap, np.uint8)
    n = C.c_int(0)
    ok = oracle().culzss_oracle_decode_buffer(np.ascontiguousarray(comp), comp.size, out, out_cap,
                                              C.byref(n))
    return ok, out[: n.value].copy()


def quant_codes(n_bytes, seed=2024, dtype=np.int32):
    """Synthetic C3 input: cuSZ-like quantisation codes 512 + round(Laplace(b=2)),
    clipped to [0, 1023], little-endian int32 (or uint16)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    n = n_bytes // np.dtype(dtype).itemsize
    codes = np.clip(512 + np.rint(rng.laplace(0.0, 2.0, n)), 0, 1023).astype(dtype)
    return codes.view(np.uint8)[:n_bytes].copy()


# ---------------------------------------------------------------------------- cudppCompress helpers
def ref_cudpp():
    if "ref_cudpp" in _cache:
        return _cache["ref_cudpp"]
    lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_cudpp.so"))
    i32a = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    lib.ref_cudpp_sa.argtypes = [_u8p, _u32p, C.c_size_t]
    lib.ref_cudpp_bwt.argtypes = [_u8p, _u8p, C.POINTER(C.c_int), C.c_uint]
    lib.ref_cudpp_mtf.argtypes = [_u8p, _u8p, C.c_uint]
    lib.ref_cudpp_tree.argtypes = [_u32p, i32a, i32a, i32a, i32a, C.POINTER(C.c_int)]
    lib.ref_cudpp_decompress.argtypes = [_u8p, C.c_int, _u32p, _u32p, C.c_size_t, _u32p, C.c_size_t]
    for f in (lib.ref_cudpp_sa, lib.ref_cudpp_bwt, lib.ref_cudpp_mtf, lib.ref_cudpp_tree,
              lib.ref_cudpp_decompress):
        f.restype = None
    _cache["ref_cudpp"] = lib
    return lib


def ref_cudpp_gpu():
    """The reference's own Huffman kernels for sm_100a (oracle/_ref/libref_cudpp_gpu.so); needs a GPU."""
    if "ref_cudpp_gpu" in _cache:
        return _cache["ref_cudpp_gpu"]
    lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_cudpp_gpu.so"))
    lib.ref_cudpp_huffman_gpu.restype = C.c_int
    lib.ref_cudpp_huffman_gpu.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _cache["ref_cudpp_gpu"] = lib
    return lib


def cudpp_test_vector(n=1 << 20, seed=95835, lo=1, span=255, sentinel=True):
    """The reference test input: srand(seed); bytes rand() % span + lo; for the compress test
    the last byte is a 0 sentinel (test_compress.cpp:552-556,687-692).  glibc rand()."""
    libc = C.CDLL("libc.so.6")
    libc.srand(seed)
    out = np.empty(n, np.uint8)
    for i in range(n):
        out[i] = libc.rand() % span + lo
    if sentinel:
        out[n - 1] = 0
    return out


def cudpp_block(n, kind="zipf", seed=0):
    """Synthetic C4 block: bytes in 1..255 with the final byte 0."""
    rng = np.random.Generator(np.random.MT19937(95835 + seed))
    if kind == "zipf":
        d = zipf_bytes(n, 1.3, seed=95835 + seed, nsym=255) + 1
    elif kind == "markov":
        # order-1 source: next symbol = previous + small step -> long repeated contexts
        steps = rng.integers(-2, 3, n)
        d = (np.cumsum(steps) % 200 + 1).astype(np.uint8)
    elif kind == "text":
        words = [b"alpha", b"beta", b"gamma", b"delta", b"epsilon", b"zeta", b"eta", b"theta"]
        idx = rng.integers(0, len(words), n // 4 + 8)
        d = np.frombuffer(b" ".join(words[i] for i in idx)[:n], np.uint8).copy()
    else:
        d = rng.integers(1, 256, n, dtype=np.uint8)
    d = d.astype(np.uint8)
    d[n - 1] = 0
    return d


def cudpp_oracle_bwt(data):
    out = np.zeros(data.size, np.uint8)
    idx = C.c_int(-1)
    oracle().cudpp_oracle_bwt(np.ascontiguousarray(data), data.size, out, C.byref(idx))
    return out, idx.value


def cudpp_oracle_mtf(data):
    out = np.zeros(data.size, np.uint8)
    oracle().cudpp_oracle_mtf(np.ascontiguousarray(data), data.size, out)
    return out


def cudpp_oracle_huffman(mtf):
    n = mtf.size
    nb = (n + 4095) // 4096
    hist = np.zeros(256, np.uint32)
    offs = np.zeros(nb, np.uint32)
    cap = nb * 1537
    out = np.zeros(cap, np.uint32)
    tw = C.c_uint32(0)
    rc = oracle().cudpp_oracle_huffman(np.ascontiguousarray(mtf), n, hist, offs, C.byref(tw), out, cap)
    return rc, hist, offs, out[: tw.value].copy()


def cudpp_oracle_compress(data):
    """-> (rc, bwt_index, hist[256], offsets[nblocks], words)"""
    b, idx = cudpp_oracle_bwt(data)
    m = cudpp_oracle_mtf(b)
    rc, hist, offs, words = cudpp_oracle_huffman(m)
    return rc, idx, hist, offs, words


def cudpp_oracle_decompress(n, idx, hist, offs, words):
    out = np.zeros(n, np.uint8)
    rc = oracle().cudpp_oracle_decompress(out, n, idx, np.ascontiguousarray(hist),
                                          np.ascontiguousarray(offs), np.ascontiguousarray(words))
    return rc, out


# ---------------------------------------------------------------------------- cuda-bzip2 helpers
def ref_bzip2(flavour=""):
    """flavour "" = reference libbz2 with its own gpuBWTSort.cu; "_b200" = the same reference
    objects with gpuBlockSort resolved from libb200lc.so."""
    key = "ref_bzip2" + flavour
    if key in _cache:
        return _cache[key]
    lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_bzip2%s.so" % flavour))
    lib.ref_bzip2_decompress.restype = C.c_int
    lib.ref_bzip2_decompress.argtypes = [_u8p, C.POINTER(C.c_uint), _u8p, C.c_uint]
    lib.ref_bzip2_gpuBlockSort.restype = C.c_int
    lib.ref_bzip2_gpuBlockSort.argtypes = [_u8p, _u32p, _u32p, _u32p, C.c_int, C.POINTER(C.c_int)]
    _cache[key] = lib
    return lib


def bzip2_ref_compress(data, block100k=9, num_threads=0, flavour=""):
    """Runs the reference libbz2 in a CHILD PROCESS: its handle_compress calls exit(1) once the
    stream has been written to strm->handle (bzlib.c:606), so the call never returns."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src, dst = os.path.join(td, "in.bin"), os.path.join(td, "out.bz2")
        np.ascontiguousarray(data).tofile(src)
        code = (
            "import ctypes as C, numpy as np\n"
            "lib = C.CDLL(%r)\n"
            "d = np.fromfile(%r, np.uint8)\n"
            "lib.ref_bzip2_compress_to_file.argtypes = [C.c_char_p, C.c_void_p, C.c_uint, C.c_int, C.c_int]\n"
            "rc = lib.ref_bzip2_compress_to_file(%r, d.ctypes.data, d.size, %d, %d)\n"
            "raise SystemExit(100 + abs(rc))\n"
        ) % (os.path.join(ORACLE_DIR, "_ref", "libref_bzip2%s.so" % flavour), src, dst.encode(),
             block100k, num_threads)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert r.returncode == 1, "reference libbz2 did not reach its exit(1): rc=%d\n%s" % (
            r.returncode, (r.stdout + r.stderr)[-800:])
        return np.fromfile(dst, np.uint8)


def bzip2_ref_decompress(comp, n, flavour=""):
    lib = ref_bzip2(flavour)
    out = np.zeros(n + 16, np.uint8)
    olen = C.c_uint(n + 16)
    rc = lib.ref_bzip2_decompress(out, C.byref(olen), np.ascontiguousarray(comp), comp.size)
    assert rc == 0, rc
    return out[: olen.value].copy()


def bzip2_in_use(block):
    u = np.zeros(256, np.uint8)
    u[np.unique(block)] = 1
    return u


def bzip2_oracle_mtf_rle(block, ptr):
    """-> (mtfv[nMTF], freq[nInUse + 2], nInUse): oracle restatement of generateMTFValues."""
    n = block.size
    mtfv = np.zeros(n + 1, np.uint16)
    freq = np.zeros(258, np.int32)
    used = C.c_int(0)
    k = oracle().bzip2_oracle_mtf_rle(np.ascontiguousarray(block), np.ascontiguousarray(ptr, dtype=np.uint32), n,
                                      bzip2_in_use(block), mtfv, freq, C.byref(used))
    return mtfv[:k].copy(), freq[: used.value + 2].copy(), used.value


def bzip2_ref_mtf_rle(block, ptr):
    """The reference's own (static) generateMTFValues through oracle/_ref/libref_bzip2_mtf.so."""
    if "ref_bzip2_mtf" not in _cache:
        lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_bzip2_mtf.so"))
        lib.ref_bzip2_generate_mtf.restype = C.c_int
        lib.ref_bzip2_generate_mtf.argtypes = [_u8p, _u32p, C.c_int, _u8p, np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS"),
                                               _i32p, C.POINTER(C.c_int)]
        _cache["ref_bzip2_mtf"] = lib
    n = block.size
    mtfv = np.zeros(n + 1, np.uint16)
    freq = np.zeros(258, np.int32)
    used = C.c_int(0)
    k = _cache["ref_bzip2_mtf"].ref_bzip2_generate_mtf(np.ascontiguousarray(block),
                                                       np.ascontiguousarray(ptr, dtype=np.uint32).copy(), n,
                                                       bzip2_in_use(block), mtfv, freq, C.byref(used))
    return mtfv[:k].copy(), freq[: used.value + 2].copy(), used.value


def bzip2_oracle_send_mtf(mtfv, freq, in_use, n_in_use):
    """Oracle restatement of sendMTFValues -> (bits bytes, nbits, len[6][258], selector[nsel], nGroups)."""
    n = mtfv.size
    bits = np.zeros(n * 3 + 4096, np.uint8)
    nbits = C.c_uint64(0)
    lens = np.zeros((6, 258), np.uint8)
    sel = np.zeros((n + 49) // 50 + 1, np.uint8)
    ng, ns = C.c_int(0), C.c_int(0)
    f = np.zeros(258, np.int32)
    f[: freq.size] = freq
    rc = oracle().bzip2_oracle_send_mtf(np.ascontiguousarray(mtfv), n, f, np.ascontiguousarray(in_use), n_in_use,
                                        bits, C.byref(nbits), lens, sel, C.byref(ng), C.byref(ns))
    assert rc == 0
    return bits[: (nbits.value + 7) // 8].copy(), nbits.value, lens, sel[: ns.value].copy(), ng.value


def bzip2_ref_send_mtf(mtfv, freq, in_use, n_in_use):
    """The reference's own (static) sendMTFValues through oracle/_ref/libref_bzip2_mtf.so."""
    bzip2_ref_mtf_rle(np.array([1], np.uint8), np.array([0], np.uint32))      # loads the library
    lib = _cache["ref_bzip2_mtf"]
    _u16 = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
    lib.ref_bzip2_send_mtf.restype = C.c_int
    lib.ref_bzip2_send_mtf.argtypes = [_u16, C.c_int, _i32p, _u8p, C.c_int, _u8p, C.POINTER(C.c_uint64), _u8p, _u8p]
    n = mtfv.size
    bits = np.zeros(n * 3 + 4096, np.uint8)
    nbits = C.c_uint64(0)
    lens = np.zeros((6, 258), np.uint8)
    sel = np.zeros((n + 49) // 50 + 1, np.uint8)
    f = np.zeros(258, np.int32)
    f[: freq.size] = freq
    rc = lib.ref_bzip2_send_mtf(np.ascontiguousarray(mtfv).copy(), n, f, np.ascontiguousarray(in_use), n_in_use, bits,
                                C.byref(nbits), lens, sel)
    assert rc == 0
    return bits[: (nbits.value + 7) // 8].copy(), nbits.value, lens, sel[: (n + 49) // 50].copy()


def bzip2_oracle_block_sort(block):
    n = block.size
    first = np.zeros(n, np.uint32)
    second = np.zeros(n, np.uint32)
    rank = np.zeros(n, np.uint32)
    f = oracle().bzip2_oracle_block_sort(np.ascontiguousarray(block), n, first, second, rank)
    return f, first[:f].copy(), second[: n - f].copy(), rank


def bzip2_oracle_merge(block, f, first, second, rank):
    n = block.size
    order = np.zeros(n, np.uint32)
    f1 = np.zeros(n, np.uint32); f1[:f] = first
    s1 = np.zeros(n, np.uint32); s1[: n - f] = second
    orig = oracle().bzip2_oracle_merge(np.ascontiguousarray(block), n, f, f1, s1, rank, order)
    return order, orig


# ---------------------------------------------------------------------------- libbsc helpers
def ref_bsc():
    """The reference's libbsc (CPU build, oracle/_ref/libref_bsc.so)."""
    if "ref_bsc" in _cache:
        return _cache["ref_bsc"]
    lib = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libref_bsc.so"))
    lib.bsc_bwt_encode.restype = C.c_int
    lib.bsc_bwt_encode.argtypes = [_u8p, C.c_int, _u8p, _i32p, C.c_int]
    lib.bsc_bwt_decode.restype = C.c_int
    lib.bsc_bwt_decode.argtypes = [_u8p, C.c_int, C.c_int, C.c_ubyte, _i32p, C.c_int]
    lib.bsc_init.restype = C.c_int
    lib.bsc_init.argtypes = [C.c_int]
    lib.bsc_st_encode.restype = C.c_int
    lib.bsc_st_encode.argtypes = [_u8p, C.c_int, C.c_int, C.c_int]
    lib.bsc_st_decode.restype = C.c_int
    lib.bsc_st_decode.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.bsc_init(0)
    _cache["ref_bsc"] = lib
    return lib


def bsc_ref_st_encode(data, k):
    """Reference CPU bsc_st_encode (k = 3..6; 7 and 8 exist on its GPU path only) -> (bytes, index)."""
    n = int(np.asarray(data).size)
    t = np.zeros(n + 4096, np.uint8)          # the reference reads / writes a few bytes past n
    t[:n] = data
    idx = ref_bsc().bsc_st_encode(t, n, k, 0)
    return t[:n].copy(), idx


def bsc_ref_st_decode(data, k, index):
    n = int(np.asarray(data).size)
    t = np.zeros(n + 4096, np.uint8)
    t[:n] = data
    rc = ref_bsc().bsc_st_decode(t, n, k, index, 0)
    return rc, t[:n].copy()


def bsc_oracle_st_encode(data, k):
    lib = oracle()
    lib.bsc_oracle_st_encode.restype = C.c_int
    lib.bsc_oracle_st_encode.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
    t = np.ascontiguousarray(data)
    out = np.zeros(max(t.size, 1), np.uint8)
    idx = lib.bsc_oracle_st_encode(t, t.size, k, out)
    return out[: t.size].copy(), idx


def bsc_ref_bwt_encode(data):
    """Reference bsc_bwt_encode (divbwt) -> (U, primary index, indexes[num_indexes])."""
    t = np.ascontiguousarray(data).copy()
    num = np.zeros(1, np.uint8)
    idx = np.zeros(256, np.int32)
    p = ref_bsc().bsc_bwt_encode(t, t.size, num, idx, 0)
    return t, p, idx[: int(num[0])].copy()


def bsc_oracle_bwt_encode(data):
    t = np.ascontiguousarray(data)
    u = np.zeros(max(1, t.size), np.uint8)
    num = np.zeros(1, np.uint8)
    idx = np.zeros(256, np.int32)
    p = oracle().bsc_oracle_bwt_encode(t, t.size, u, num, idx)
    return u[: t.size], p, idx[: int(num[0])].copy()
