"""b200lc -- Blackwell (sm_100a) lossless block-compression kernels behind the C ABI in include/.

The product is lib/libb200lc.so (C ABI, no torch types).  This package is only the thin
host-side door used by tests/ and bench.py: it loads the library with ctypes and passes raw
device pointers of torch tensors.  There is no CPU fallback: if the library has not been built
(`python gpu-lossless-compression_b200/build.py` or `__graft_entry__.build()`), importing the
bindings raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200lc.so")

OK = 0
ERR_ARG, ERR_CUDA, ERR_SCRATCH, ERR_UNSUPPORTED, ERR_OVERFLOW = -1, -2, -3, -4, -5

_lib = None


class B200LCError(RuntimeError):
    pass


def lib():
    """The loaded libb200lc.so (ctypes.CDLL) with argtypes set; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LCError(
            "libb200lc.so is not built (%s missing): run __graft_entry__.build(); "
            "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    L.b200lc_version.restype = C.c_char_p
    L.b200lc_cuhd_decode_scratch_bytes.restype = sz
    L.b200lc_cuhd_decode_scratch_bytes.argtypes = [sz]
    L.b200lc_cuhd_decode.restype = i32
    L.b200lc_cuhd_decode.argtypes = [vp, sz, vp, sz, vp, i32, vp, sz, vp]
    L.b200lc_cuhd_decode_batch_scratch_bytes.restype = sz
    L.b200lc_cuhd_decode_batch_scratch_bytes.argtypes = [vp, sz]
    L.b200lc_cuhd_decode_batch.restype = i32
    L.b200lc_cuhd_decode_batch.argtypes = [vp, vp, vp, sz, vp, i32, vp, sz, vp]
    L.b200lc_histogram_u8.restype = i32
    L.b200lc_histogram_u8.argtypes = [vp, sz, vp, vp]
    L.b200lc_cuhd_build_table.restype = i32
    L.b200lc_cuhd_build_table.argtypes = [vp, i32, vp, vp, vp]
    L.b200lc_cuhd_compressed_units.restype = sz
    L.b200lc_cuhd_compressed_units.argtypes = [vp, vp]
    L.b200lc_cuhd_encode_scratch_bytes.restype = sz
    L.b200lc_cuhd_encode_scratch_bytes.argtypes = [sz]
    L.b200lc_cuhd_encode.restype = i32
    L.b200lc_cuhd_encode.argtypes = [vp, sz, vp, vp, vp, sz, vp, vp, sz, vp]
    L.b200lc_cuhd_encode_blocks_scratch_bytes.restype = sz
    L.b200lc_cuhd_encode_blocks_scratch_bytes.argtypes = [sz, sz]
    L.b200lc_cuhd_encode_blocks.restype = i32
    L.b200lc_cuhd_encode_blocks.argtypes = [vp, sz, sz, vp, vp, vp, sz, vp, vp, sz, vp]
    L.b200lc_cuhd_piece_hist_bytes.restype = sz
    L.b200lc_cuhd_piece_hist_bytes.argtypes = [sz]
    L.b200lc_histogram_u8_pieces.restype = i32
    L.b200lc_histogram_u8_pieces.argtypes = [vp, sz, vp, vp, vp]
    L.b200lc_cuhd_encode_planned.restype = i32
    L.b200lc_cuhd_encode_planned.argtypes = [vp, sz, vp, vp, vp, vp, sz, vp, vp, sz, vp]
    L.b200lc_cuhd_encode_overflowed.restype = i32
    L.b200lc_cuhd_encode_overflowed.argtypes = [vp, vp]
    L.b200lc_cuhd_session_create.restype = i32
    L.b200lc_cuhd_session_create.argtypes = [sz, C.POINTER(vp)]
    L.b200lc_cuhd_session_destroy.restype = i32
    L.b200lc_cuhd_session_destroy.argtypes = [vp]
    L.b200lc_cuhd_session_encode.restype = i32
    L.b200lc_cuhd_session_encode.argtypes = [vp, vp, sz, i32, vp, sz, C.POINTER(sz), vp, vp, vp]
    L.b200lc_cuhd_session_decode.restype = i32
    L.b200lc_cuhd_session_decode.argtypes = [vp, vp, sz, vp, i32, vp, sz]
    L.b200lc_culzss_encode_scratch_bytes.restype = sz
    L.b200lc_culzss_encode_scratch_bytes.argtypes = [sz, sz]
    L.b200lc_culzss_encode_batch.restype = i32
    L.b200lc_culzss_encode_batch.argtypes = [vp, sz, sz, vp, sz, vp, vp, sz, vp]
    L.b200lc_culzss_encode_batch_ex.restype = i32
    L.b200lc_culzss_encode_batch_ex.argtypes = [vp, sz, sz, vp, sz, vp, vp, sz, i32, vp]
    L.b200lc_culzss_encode_fast_batch.restype = i32
    L.b200lc_culzss_encode_fast_batch.argtypes = [vp, sz, sz, vp, sz, vp, vp, sz, i32, vp]
    L.b200lc_culzss_decode_scratch_bytes.restype = sz
    L.b200lc_culzss_decode_scratch_bytes.argtypes = [sz, sz]
    L.b200lc_culzss_decode_batch.restype = i32
    L.b200lc_culzss_decode_batch.argtypes = [vp, vp, sz, sz, vp, vp, sz, vp]
    L.b200lc_bwt_scratch_bytes.restype = sz
    L.b200lc_bwt_scratch_bytes.argtypes = [sz, sz]
    L.b200lc_bwt_batch.restype = i32
    L.b200lc_bwt_batch.argtypes = [vp, sz, sz, vp, vp, vp, sz, vp]
    L.b200lc_suffix_array_batch.restype = i32
    L.b200lc_suffix_array_batch.argtypes = [vp, sz, sz, vp, vp, sz, vp]
    L.b200lc_mtf_scratch_bytes.restype = sz
    L.b200lc_mtf_scratch_bytes.argtypes = [sz, sz]
    L.b200lc_mtf_batch.restype = i32
    L.b200lc_mtf_batch.argtypes = [vp, sz, sz, vp, vp, sz, vp]
    L.b200lc_cudpp_huffman_scratch_bytes.restype = sz
    L.b200lc_cudpp_huffman_scratch_bytes.argtypes = [sz, sz]
    L.b200lc_cudpp_huffman_batch.restype = i32
    L.b200lc_cudpp_huffman_batch.argtypes = [vp, sz, sz, vp, vp, vp, vp, sz, vp, vp, sz, vp]
    L.b200lc_cudpp_compress_scratch_bytes.restype = sz
    L.b200lc_cudpp_compress_scratch_bytes.argtypes = [sz, sz]
    L.b200lc_cudpp_compress_batch.restype = i32
    L.b200lc_cudpp_compress_batch.argtypes = [vp, sz, sz, vp, vp, vp, vp, vp, sz, vp, vp, sz, vp]
    for name in ("inverse_mtf", "inverse_bwt", "cudpp_decompress"):
        f = getattr(L, "b200lc_%s_scratch_bytes" % name)
        f.restype, f.argtypes = sz, [sz, sz]
    L.b200lc_inverse_mtf_batch.restype = i32
    L.b200lc_inverse_mtf_batch.argtypes = [vp, sz, sz, vp, vp, sz, vp]
    L.b200lc_inverse_bwt_batch.restype = i32
    L.b200lc_inverse_bwt_batch.argtypes = [vp, vp, sz, sz, vp, vp, vp, sz, vp]
    L.b200lc_cudpp_decompress_batch.restype = i32
    L.b200lc_cudpp_decompress_batch.argtypes = [vp, vp, vp, vp, sz, sz, sz, vp, vp, vp, sz, vp]
    L.b200lc_sort_scratch_bytes.restype = sz
    L.b200lc_sort_scratch_bytes.argtypes = [sz, sz]
    for name in ("b200lc_sort_pairs_u64", "b200lc_sort_pairs_u32"):
        f = getattr(L, name)
        f.restype, f.argtypes = i32, [vp, vp, vp, vp, sz, sz, i32, i32, vp, sz, vp, C.POINTER(i32)]
    L.b200lc_scan_scratch_bytes.restype = sz
    L.b200lc_scan_scratch_bytes.argtypes = [sz]
    for name in ("b200lc_exclusive_sum_u32", "b200lc_inclusive_max_u32"):
        f = getattr(L, name)
        f.restype, f.argtypes = i32, [vp, vp, sz, vp, sz, vp]
    L.b200lc_bzip2_mtf_rle.restype = i32
    L.b200lc_bzip2_mtf_rle.argtypes = [vp, vp, i32, vp, vp, C.POINTER(i32), vp, C.POINTER(i32)]
    L.b200lc_bzip2_send_mtf_values.restype = i32
    L.b200lc_bzip2_send_mtf_values.argtypes = [vp, i32, vp, vp, i32, vp, sz, C.POINTER(C.c_ulonglong), vp, vp]
    L.bsc_bwt_encode.restype = i32
    L.bsc_bwt_encode.argtypes = [vp, i32, vp, vp, i32]
    L.bsc_bwt_decode.restype = i32
    L.bsc_bwt_decode.argtypes = [vp, i32, i32, C.c_ubyte, vp, i32]
    L.b200lc_bsc_release.restype = None
    _lib = L
    return L


def check(rc, what):
    if rc != OK:
        raise B200LCError("%s failed with code %d" % (what, rc))


def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


# ------------------------------------------------------------------------------- CUHD
def cuhd_decode(units, n_out, lut, max_len=11, out=None, scratch=None, stream=None):
    """Decode a CUHD stream resident on the GPU.

    units: cuda int32/uint32 tensor of stream units; lut: cuda uint8 tensor [(1<<max_len), 2]
    of {num_bits, symbol}; returns a cuda uint8 tensor of n_out symbols.  Asynchronous.
    """
    import torch
    assert units.is_cuda and lut.is_cuda and units.is_contiguous() and lut.is_contiguous()
    L = lib()
    n_units = units.numel()
    if out is None:
        out = torch.empty(n_out, dtype=torch.uint8, device=units.device)
    need = L.b200lc_cuhd_decode_scratch_bytes(n_units)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=units.device)
    rc = L.b200lc_cuhd_decode(units.data_ptr(), n_units, out.data_ptr(), n_out, lut.data_ptr(),
                              max_len, scratch.data_ptr(), scratch.numel(), _stream_ptr(stream))
    check(rc, "b200lc_cuhd_decode")
    return out


def cuhd_decode_batch(units, out, streams, lut, max_len=11, scratch=None, stream=None):
    """Decode many independent streams sharing one table with one launch.  units: cuda int32 tensor
    holding all streams, out: cuda uint8 tensor, streams: numpy uint64 array [n, 4] of
    (unit_offset, n_units, out_offset, n_out).  Asynchronous."""
    import numpy as np
    import torch
    st = np.ascontiguousarray(streams, dtype=np.uint64)
    assert st.ndim == 2 and st.shape[1] == 4
    L = lib()
    need = L.b200lc_cuhd_decode_batch_scratch_bytes(st.ctypes.data, st.shape[0])
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=units.device)
    check(L.b200lc_cuhd_decode_batch(units.data_ptr(), out.data_ptr(), st.ctypes.data, st.shape[0],
                                     lut.data_ptr(), max_len, scratch.data_ptr(), scratch.numel(),
                                     _stream_ptr(stream)), "b200lc_cuhd_decode_batch")
    return out


def histogram_u8(data, stream=None):
    """256-bin histogram (cuda int64 tensor) of a cuda uint8 tensor.  Asynchronous."""
    import torch
    assert data.is_cuda and data.dtype == torch.uint8 and data.is_contiguous()
    hist = torch.empty(256, dtype=torch.int64, device=data.device)
    check(lib().b200lc_histogram_u8(data.data_ptr(), data.numel(), hist.data_ptr(),
                                    _stream_ptr(stream)), "b200lc_histogram_u8")
    return hist


def cuhd_build_table(hist, max_len=11):
    """Host: histogram (array-like of 256 counts) -> (code u32[256], len u8[256], lut u8[1<<L, 2])
    as numpy arrays."""
    import numpy as np
    h = np.ascontiguousarray(np.asarray(hist, dtype=np.uint64))
    assert h.size == 256
    code = np.zeros(256, np.uint32)
    length = np.zeros(256, np.uint8)
    lut = np.zeros((1 << max_len, 2), np.uint8)
    check(lib().b200lc_cuhd_build_table(h.ctypes.data, max_len, code.ctypes.data,
                                        length.ctypes.data, lut.ctypes.data),
          "b200lc_cuhd_build_table")
    return code, length, lut


class CuhdEncoded:
    """Result of cuhd_encode: `units` (cuda int32, incl. one zero pad unit), `n_units`, `bits`."""

    def __init__(self, units, n_units, bits):
        self.units, self.n_units, self.bits = units, n_units, bits


def cuhd_encode(data, code, length, units_cap=None, stream=None, scratch=None, units=None,
                total_bits=None, sync=True, piece_hist=None):
    """Pack a cuda uint8 tensor with the dictionary (code, length: cuda int32[256] / uint8[256]).

    With sync=True (default) the bit count is read back and the unit tensor is trimmed to
    ceil(bits/32) + 1 pad unit; with sync=False the untrimmed buffer and the device bit counter
    are returned without synchronising.
    """
    import torch
    assert data.is_cuda and data.dtype == torch.uint8 and data.is_contiguous()
    L = lib()
    n = data.numel()
    if units_cap is None:
        units_cap = (n * 13 + 31) // 32 + 2 if units is None else units.numel()
    if units is None:
        units = torch.empty(units_cap, dtype=torch.int32, device=data.device)
    if total_bits is None:
        total_bits = torch.zeros(1, dtype=torch.int64, device=data.device)
    need = L.b200lc_cuhd_encode_scratch_bytes(n)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=data.device)
    if piece_hist is not None:      # one-pass packer on the piece histograms of histogram_u8_pieces
        check(L.b200lc_cuhd_encode_planned(data.data_ptr(), n, code.data_ptr(), length.data_ptr(),
                                           piece_hist.data_ptr(), units.data_ptr(), units_cap,
                                           total_bits.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                           _stream_ptr(stream)), "b200lc_cuhd_encode_planned")
    else:
        check(L.b200lc_cuhd_encode(data.data_ptr(), n, code.data_ptr(), length.data_ptr(),
                                   units.data_ptr(), units_cap, total_bits.data_ptr(),
                                   scratch.data_ptr(), scratch.numel(), _stream_ptr(stream)),
              "b200lc_cuhd_encode")
    if not sync:
        return CuhdEncoded(units, None, total_bits)
    check(L.b200lc_cuhd_encode_overflowed(scratch.data_ptr(), _stream_ptr(stream)),
          "b200lc_cuhd_encode (capacity)")
    bits = int(total_bits.item())
    n_units = (bits + 31) // 32
    return CuhdEncoded(units[: min(n_units + 1, units_cap)], n_units, bits)


def histogram_u8_pieces(data, piece_hist=None, stream=None):
    """-> (hist int64[256], piece_hist int32[pieces * 256]) for cuhd_encode(..., piece_hist=...)."""
    import torch
    L = lib()
    n = data.numel()
    hist = torch.empty(256, dtype=torch.int64, device=data.device)
    if piece_hist is None:
        piece_hist = torch.empty(max(1, L.b200lc_cuhd_piece_hist_bytes(n) // 4), dtype=torch.int32, device=data.device)
    check(L.b200lc_histogram_u8_pieces(data.data_ptr(), n, hist.data_ptr(), piece_hist.data_ptr(),
                                       _stream_ptr(stream)), "b200lc_histogram_u8_pieces")
    return hist, piece_hist


def cuhd_encode_blocks(data, block, code, length, unit_stride=None, units=None, block_bits=None,
                       scratch=None, stream=None):
    """Pack blocks of `block` symbols of a cuda uint8 tensor as independent streams with one
    dictionary in one launch.  Returns (units int32 [nblocks * unit_stride], block_bits int64
    [nblocks], unit_stride).  Asynchronous."""
    import torch
    L = lib()
    n = data.numel()
    nblocks = (n + block - 1) // block
    if unit_stride is None:
        unit_stride = ((block * 13 + 31) // 32 + 2 + 3) // 4 * 4
    if units is None:
        units = torch.empty(nblocks * unit_stride, dtype=torch.int32, device=data.device)
    if block_bits is None:
        block_bits = torch.zeros(nblocks, dtype=torch.int64, device=data.device)
    need = L.b200lc_cuhd_encode_blocks_scratch_bytes(n, block)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=data.device)
    check(L.b200lc_cuhd_encode_blocks(data.data_ptr(), n, block, code.data_ptr(), length.data_ptr(),
                                      units.data_ptr(), unit_stride, block_bits.data_ptr(),
                                      scratch.data_ptr(), scratch.numel(), _stream_ptr(stream)),
          "b200lc_cuhd_encode_blocks")
    return units, block_bits, unit_stride


class CuhdSession:
    """Host-buffer CUHD encode/decode (b200lc_cuhd_session_*).  Arguments are host tensors
    (pinned for full PCIe speed); every call is synchronous and includes H2D + D2H."""

    def __init__(self, max_symbols):
        self._h = C.c_void_p()
        check(lib().b200lc_cuhd_session_create(max_symbols, C.byref(self._h)),
              "b200lc_cuhd_session_create")

    def close(self):
        if self._h:
            lib().b200lc_cuhd_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def encode(self, h_in, h_units, h_code, h_len, h_lut, max_len=11):
        """-> number of stream units written to h_units (a pad unit follows them)."""
        n_units = C.c_size_t(0)
        check(lib().b200lc_cuhd_session_encode(self._h, h_in.data_ptr(), h_in.numel(), max_len,
                                               h_units.data_ptr(), h_units.numel(),
                                               C.byref(n_units), h_code.data_ptr(),
                                               h_len.data_ptr(), h_lut.data_ptr()),
              "b200lc_cuhd_session_encode")
        return n_units.value

    def decode(self, h_units, n_units, h_lut, h_out, max_len=11):
        check(lib().b200lc_cuhd_session_decode(self._h, h_units.data_ptr(), n_units,
                                               h_lut.data_ptr(), max_len, h_out.data_ptr(),
                                               h_out.numel()),
              "b200lc_cuhd_session_decode")


# ------------------------------------------------------------------------------- CULZSS
CULZSS_FAST_LANE = -1      # include/b200lc.h B200LC_CULZSS_FAST_LANE


def culzss_out_stride(buf_length):
    return (buf_length + buf_length // 8 + 1024 + 15) // 16 * 16


CULZSS_KERNEL_AUTO, CULZSS_KERNEL_CTA, CULZSS_KERNEL_LANE = 0, 1, 2


def culzss_encode(data, buf_length=1 << 20, out=None, comp_len=None, scratch=None, stream=None, fast=0,
                  kernel=CULZSS_KERNEL_AUTO):
    """LZSS-encode a cuda uint8 tensor of nbuf * buf_length bytes.  Returns (out, comp_len):
    out[b * stride : b * stride + comp_len[b]] is buffer b incl. trailer; comp_len[b] == 0 means
    "store raw".  Asynchronous.  fast = 1, 2 or 4: the NON-PARITY fast mode
    (b200lc_culzss_encode_fast_batch, hash-chain depth), same format, different matches;
    fast = "lane" (CULZSS_FAST_LANE = -1): its packet-per-lane formulation, also NON-PARITY.
    kernel (parity mode only): CULZSS_KERNEL_CTA / _LANE force one of the two bit-identical kernels
    (b200lc_culzss_encode_batch_ex)."""
    import torch
    assert data.is_cuda and data.dtype == torch.uint8 and data.is_contiguous()
    assert data.numel() % buf_length == 0
    L = lib()
    nbuf = data.numel() // buf_length
    stride = culzss_out_stride(buf_length)
    if out is None:
        out = torch.empty(nbuf * stride, dtype=torch.uint8, device=data.device)
    if comp_len is None:
        comp_len = torch.empty(nbuf, dtype=torch.int32, device=data.device)
    need = L.b200lc_culzss_encode_scratch_bytes(nbuf, buf_length)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=data.device)
    if fast == "lane":
        fast = CULZSS_FAST_LANE
    if fast:
        check(L.b200lc_culzss_encode_fast_batch(data.data_ptr(), nbuf, buf_length, out.data_ptr(), stride,
                                                comp_len.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                                int(fast), _stream_ptr(stream)), "b200lc_culzss_encode_fast_batch")
    elif kernel != CULZSS_KERNEL_AUTO:
        check(L.b200lc_culzss_encode_batch_ex(data.data_ptr(), nbuf, buf_length, out.data_ptr(), stride,
                                              comp_len.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                              int(kernel), _stream_ptr(stream)), "b200lc_culzss_encode_batch_ex")
    else:
        check(L.b200lc_culzss_encode_batch(data.data_ptr(), nbuf, buf_length, out.data_ptr(), stride,
                                           comp_len.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                           _stream_ptr(stream)), "b200lc_culzss_encode_batch")
    return out, comp_len


def culzss_decode(comp, offsets, buf_length=1 << 20, out=None, scratch=None, stream=None):
    """Decode nbuf compressed buffers: comp (cuda uint8), offsets (cuda int64[nbuf + 1])."""
    import torch
    assert comp.is_cuda and offsets.is_cuda and offsets.dtype == torch.int64
    L = lib()
    nbuf = offsets.numel() - 1
    if out is None:
        out = torch.empty(nbuf * buf_length, dtype=torch.uint8, device=comp.device)
    need = L.b200lc_culzss_decode_scratch_bytes(nbuf, buf_length)
    if scratch is None or scratch.numel() < need:
        scratch = torch.empty(need, dtype=torch.uint8, device=comp.device)
    check(L.b200lc_culzss_decode_batch(comp.data_ptr(), offsets.data_ptr(), nbuf, buf_length,
                                       out.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                       _stream_ptr(stream)), "b200lc_culzss_decode_batch")
    return out


# ------------------------------------------------------------------------------- cudppCompress path
def _scratch(nbytes, device):
    import torch
    return torch.empty(nbytes + 256, dtype=torch.uint8, device=device)


def bwt_batch(data, nblocks, n, stream=None):
    """BWT of nblocks blocks of n bytes (cuda uint8).  Returns (bwt, index int32[nblocks])."""
    import torch
    L = lib()
    out = torch.empty_like(data)
    index = torch.empty(nblocks, dtype=torch.int32, device=data.device)
    sc = _scratch(L.b200lc_bwt_scratch_bytes(nblocks, n), data.device)
    check(L.b200lc_bwt_batch(data.data_ptr(), nblocks, n, out.data_ptr(), index.data_ptr(),
                             sc.data_ptr(), sc.numel(), _stream_ptr(stream)), "b200lc_bwt_batch")
    return out, index


def suffix_array_batch(data, nblocks, n, stream=None):
    import torch
    L = lib()
    sa = torch.empty(nblocks * n, dtype=torch.int32, device=data.device)
    sc = _scratch(L.b200lc_bwt_scratch_bytes(nblocks, n), data.device)
    check(L.b200lc_suffix_array_batch(data.data_ptr(), nblocks, n, sa.data_ptr(), sc.data_ptr(),
                                      sc.numel(), _stream_ptr(stream)), "b200lc_suffix_array_batch")
    return sa


def mtf_batch(data, nblocks, n, stream=None):
    import torch
    L = lib()
    out = torch.empty_like(data)
    sc = _scratch(L.b200lc_mtf_scratch_bytes(nblocks, n), data.device)
    check(L.b200lc_mtf_batch(data.data_ptr(), nblocks, n, out.data_ptr(), sc.data_ptr(), sc.numel(),
                             _stream_ptr(stream)), "b200lc_mtf_batch")
    return out


class CudppCompressed:
    def __init__(self, bwt_index, hist, offsets, total_words, words, stride, error):
        self.bwt_index, self.hist, self.offsets = bwt_index, hist, offsets
        self.total_words, self.words, self.stride, self.error = total_words, words, stride, error


def cudpp_compress_batch(data, nblocks, n, stream=None, scratch=None, out=None):
    """cudppCompress of nblocks blocks of n bytes.  Synchronises (BWT stage)."""
    import torch
    L = lib()
    dev = data.device
    nhb = (n + 4095) // 4096
    stride = nhb * 1537
    if out is None:
        out = CudppCompressed(torch.empty(nblocks, dtype=torch.int32, device=dev),
                              torch.empty(nblocks * 256, dtype=torch.int32, device=dev),
                              torch.empty(nblocks * nhb, dtype=torch.int32, device=dev),
                              torch.empty(nblocks, dtype=torch.int32, device=dev),
                              torch.empty(nblocks * stride, dtype=torch.int32, device=dev), stride,
                              torch.zeros(1, dtype=torch.int32, device=dev))
    if scratch is None:
        scratch = _scratch(L.b200lc_cudpp_compress_scratch_bytes(nblocks, n), dev)
    check(L.b200lc_cudpp_compress_batch(data.data_ptr(), nblocks, n, out.bwt_index.data_ptr(),
                                        out.hist.data_ptr(), out.offsets.data_ptr(),
                                        out.total_words.data_ptr(), out.words.data_ptr(), stride,
                                        out.error.data_ptr(), scratch.data_ptr(), scratch.numel(),
                                        _stream_ptr(stream)), "b200lc_cudpp_compress_batch")
    return out


def inverse_mtf_batch(ranks, nblocks, n, stream=None):
    """Inverse move-to-front of nblocks blocks of n ranks (initial list 0..255)."""
    import torch
    L = lib()
    out = torch.empty_like(ranks)
    sc = _scratch(L.b200lc_inverse_mtf_scratch_bytes(nblocks, n), ranks.device)
    check(L.b200lc_inverse_mtf_batch(ranks.data_ptr(), nblocks, n, out.data_ptr(), sc.data_ptr(),
                                     sc.numel(), _stream_ptr(stream)), "b200lc_inverse_mtf_batch")
    return out


def inverse_bwt_batch(bwt, bwt_index, nblocks, n, stream=None):
    """Inverse BWT of nblocks blocks (index convention of cudppBurrowsWheelerTransform)."""
    import torch
    L = lib()
    out = torch.empty_like(bwt)
    err = torch.zeros(1, dtype=torch.int32, device=bwt.device)
    sc = _scratch(L.b200lc_inverse_bwt_scratch_bytes(nblocks, n), bwt.device)
    check(L.b200lc_inverse_bwt_batch(bwt.data_ptr(), bwt_index.data_ptr(), nblocks, n, out.data_ptr(),
                                     err.data_ptr(), sc.data_ptr(), sc.numel(), _stream_ptr(stream)),
          "b200lc_inverse_bwt_batch")
    return out, err


def cudpp_decompress_batch(comp, nblocks, n, stream=None, scratch=None, out=None):
    """Inverse of cudpp_compress_batch: CudppCompressed -> (bytes, error word)."""
    import torch
    L = lib()
    dev = comp.words.device
    if out is None:
        out = torch.empty(nblocks * n, dtype=torch.uint8, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    if scratch is None:
        scratch = _scratch(L.b200lc_cudpp_decompress_scratch_bytes(nblocks, n), dev)
    check(L.b200lc_cudpp_decompress_batch(comp.bwt_index.data_ptr(), comp.hist.data_ptr(),
                                          comp.offsets.data_ptr(), comp.words.data_ptr(), comp.stride,
                                          nblocks, n, out.data_ptr(), err.data_ptr(), scratch.data_ptr(),
                                          scratch.numel(), _stream_ptr(stream)),
          "b200lc_cudpp_decompress_batch")
    return out, err


# ------------------------------------------------------------------------------- primitives
def sort_pairs(keys, vals, seg_len=0, begin_bit=0, end_bit=None, stream=None):
    """Stable segmented LSD radix sort (csrc/devprims.cu).  keys: cuda int64 (u64 bit pattern) or
    int32 (u32 bit pattern) tensor; vals: cuda int32 tensor.  Returns (sorted_keys, sorted_vals);
    the inputs are used as one of the ping-pong buffers and are clobbered."""
    import torch
    assert keys.is_cuda and vals.is_cuda and keys.numel() == vals.numel()
    L = lib()
    wide = keys.element_size() == 8
    if end_bit is None:
        end_bit = 64 if wide else 32
    n = keys.numel()
    kb, vb = torch.empty_like(keys), torch.empty_like(vals)
    nbytes = L.b200lc_sort_scratch_bytes(n, seg_len)
    scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=keys.device)
    where = C.c_int(0)
    fn = L.b200lc_sort_pairs_u64 if wide else L.b200lc_sort_pairs_u32
    check(fn(keys.data_ptr(), kb.data_ptr(), vals.data_ptr(), vb.data_ptr(), n, seg_len, begin_bit,
             end_bit, scratch.data_ptr(), scratch.numel(), _stream_ptr(stream), C.byref(where)), "sort_pairs")
    (stream if stream is not None else torch.cuda.current_stream()).synchronize()   # scratch dies here
    return (kb, vb) if where.value else (keys, vals)


def scan_u32(x, kind="exclusive_sum", out=None, stream=None):
    """kind: 'exclusive_sum' or 'inclusive_max' over a cuda int32 tensor (u32 bit pattern)."""
    import torch
    L = lib()
    n = x.numel()
    if out is None:
        out = torch.empty_like(x)
    scratch = torch.empty(L.b200lc_scan_scratch_bytes(n) + 256, dtype=torch.uint8, device=x.device)
    fn = L.b200lc_exclusive_sum_u32 if kind == "exclusive_sum" else L.b200lc_inclusive_max_u32
    check(fn(x.data_ptr(), out.data_ptr(), n, scratch.data_ptr(), scratch.numel(), _stream_ptr(stream)), kind)
    (stream if stream is not None else torch.cuda.current_stream()).synchronize()   # scratch dies here
    return out


# ------------------------------------------------------------------------------- libbsc BWT stage
def bsc_bwt_encode(data):
    """libbsc's bsc_bwt_encode on the GPU (include/libbsc_gpu.h).  data: numpy uint8 (host).
    Returns (U, primary_index, indexes) like the reference call does through its out-parameters."""
    import numpy as np
    t = np.ascontiguousarray(data, dtype=np.uint8).copy()
    num = np.zeros(1, np.uint8)
    idx = np.zeros(256, np.int32)
    p = lib().bsc_bwt_encode(t.ctypes.data, t.size, num.ctypes.data, idx.ctypes.data, 0)
    if p < 0:
        raise B200LCError("bsc_bwt_encode failed with LIBBSC code %d" % p)
    return t, p, idx[: int(num[0])].copy()


def bsc_bwt_decode(u, index):
    """libbsc's bsc_bwt_decode on the GPU: (U, primary index) as bsc_bwt_encode returns them -> block."""
    import numpy as np
    t = np.ascontiguousarray(u, dtype=np.uint8).copy()
    rc = lib().bsc_bwt_decode(t.ctypes.data, t.size, int(index), 0, None, 0)
    if rc < 0:
        raise B200LCError("bsc_bwt_decode failed with LIBBSC code %d" % rc)
    return t


# ------------------------------------------------------------------------------- bzip2 MTF + RLE stage
def bzip2_mtf_rle(block, ptr):
    """bzip2's generateMTFValues on the GPU (include/bzip2_gpu.h).  block: numpy uint8, ptr: numpy
    uint32 sorted rotation order (host).  Returns (mtfv[nMTF], freq[nInUse + 2], nInUse)."""
    import numpy as np
    block = np.ascontiguousarray(block, dtype=np.uint8)
    ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
    n = block.size
    in_use = np.zeros(256, np.uint8)
    in_use[np.unique(block)] = 1
    mtfv = np.zeros(n + 1, np.uint16)
    freq = np.zeros(258, np.int32)
    n_mtf, used = C.c_int(0), C.c_int(0)
    check(lib().b200lc_bzip2_mtf_rle(block.ctypes.data, ptr.ctypes.data, n, in_use.ctypes.data,
                                     mtfv.ctypes.data, C.byref(n_mtf), freq.ctypes.data, C.byref(used)),
          "b200lc_bzip2_mtf_rle")
    return mtfv[: n_mtf.value].copy(), freq[: used.value + 2].copy(), used.value


def bzip2_send_mtf_values(mtfv, freq, in_use, n_in_use):
    """bzip2's sendMTFValues on the GPU (include/bzip2_gpu.h), host numpy arrays.
    Returns (bits bytes, nbits, len[6][258], selector[ceil(nMTF / 50)])."""
    import numpy as np
    mtfv = np.ascontiguousarray(mtfv, dtype=np.uint16)
    n = mtfv.size
    f = np.zeros(258, np.int32)
    f[: len(freq)] = freq
    iu = np.ascontiguousarray(in_use, dtype=np.uint8)
    cap = n * 17 // 8 + n // 50 + 8192
    bits = np.zeros(cap, np.uint8)
    nbits = C.c_ulonglong(0)
    lens = np.zeros((6, 258), np.uint8)
    sel = np.zeros((n + 49) // 50, np.uint8)
    check(lib().b200lc_bzip2_send_mtf_values(mtfv.ctypes.data, n, f.ctypes.data, iu.ctypes.data, n_in_use,
                                             bits.ctypes.data, cap, C.byref(nbits), lens.ctypes.data,
                                             sel.ctypes.data), "b200lc_bzip2_send_mtf_values")
    return bits[: (nbits.value + 7) // 8].copy(), nbits.value, lens, sel
