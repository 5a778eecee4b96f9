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
