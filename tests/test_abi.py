"""CPU: the C-ABI library loads and exports every symbol the public headers declare
(no compute calls here -- there is no GPU in the build container)."""
import ctypes
import glob
import os
import re
import subprocess

from pkg import b200lc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b([A-Za-z_]\w*)\s*\(", re.M)


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
        for m in DECL.finditer(text):
            name = m.group(1)
            if name not in ("defined", "sizeof", "if", "while", "return"):
                names.add(name)
    return names


def test_library_exports_every_declared_symbol():
    lib = b200lc.lib()
    names = declared_symbols()
    assert len(names) >= 10
    # gpuBlockSort / gpuSetDevice keep the reference's C++ linkage (include/bzip2_gpu.h): look for
    # their Itanium-mangled names (_Z<len><name>...) in the dynamic symbol table
    dyn = subprocess.run(["nm", "-D", "--defined-only", b200lc.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    missing = []
    for n in sorted(names):
        try:
            getattr(lib, n)
        except AttributeError:
            if not re.search(r"\b_Z%d%s\w*" % (len(n), n), dyn):
                missing.append(n)
    assert not missing, "declared in include/ but not exported: %s" % missing


def test_version_string():
    v = b200lc.lib().b200lc_version().decode()
    assert v.startswith("b200lc ") and v.endswith("sm_100a")


def test_errors_without_gpu_are_loud():
    lib = b200lc.lib()
    # null pointers are rejected before any CUDA call
    assert lib.b200lc_cuhd_decode(None, 10, None, 10, None, 11, None, 0, None) == b200lc.ERR_ARG
    assert lib.b200lc_cuhd_decode(None, 10, None, 10, None, 20, None, 0, None) == b200lc.ERR_UNSUPPORTED
    assert lib.b200lc_culzss_encode_batch(None, 1, 4096, None, 0, None, None, 0, None) == b200lc.ERR_ARG


def test_new_entry_points_reject_bad_arguments_before_touching_cuda():
    lib = b200lc.lib()
    E = b200lc
    where = ctypes.c_int(0)
    assert lib.b200lc_sort_pairs_u64(None, None, None, None, 10, 0, 0, 48, None, 0, None, ctypes.byref(where)) == E.ERR_ARG
    assert lib.b200lc_sort_pairs_u32(None, None, None, None, 0, 0, 0, 32, None, 0, None, ctypes.byref(where)) == E.OK   # nothing to sort
    assert lib.b200lc_exclusive_sum_u32(None, None, 10, None, 0, None) == E.ERR_ARG
    assert lib.b200lc_cuhd_decode_batch(None, None, None, 3, None, 11, None, 0, None) == E.ERR_ARG
    assert lib.b200lc_cuhd_decode_batch(None, None, None, 3, None, 14, None, 0, None) == E.ERR_UNSUPPORTED
    assert lib.b200lc_cuhd_encode_blocks(None, 100, 0, None, None, None, 0, None, None, 0, None) == E.ERR_ARG
    assert lib.b200lc_bzip2_mtf_rle(None, None, 10, None, None, None, None, None) == E.ERR_ARG
    nb = ctypes.c_ulonglong(0)
    assert lib.b200lc_bzip2_send_mtf_values(None, 10, None, None, 4, None, 0, ctypes.byref(nb), None, None) == E.ERR_ARG
    assert lib.bsc_bwt_encode(None, 10, None, None, 0) == -1          # LIBBSC_BAD_PARAMETER (libbsc.h:52)
    assert lib.b200lc_sort_scratch_bytes(1 << 20, 1 << 18) > (1 << 20) // 4096 * 1024
    # scratch sizes are pure host arithmetic
    st = (ctypes.c_uint64 * 8)(0, 1000, 0, 4000, 1000, 2000, 4000, 8000)
    assert lib.b200lc_cuhd_decode_batch_scratch_bytes(st, 2) >= 128 + 2 * 64
    assert lib.b200lc_cuhd_encode_blocks_scratch_bytes(1 << 20, 1 << 16) >= 256 + 16 * 16


def test_round2_entry_points_reject_bad_arguments_before_touching_cuda():
    """The entry points added in round 2 (include/b200lc.h, include/libbsc_gpu.h): argument errors
    and trivial sizes come back before any CUDA call, so this runs without a GPU."""
    lib = b200lc.lib()
    E = b200lc
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    lib.b200lc_culzss_encode_batch_ex.argtypes = [vp, sz, sz, vp, sz, vp, vp, sz, i32, vp]
    assert lib.b200lc_culzss_encode_batch_ex(None, 1, 4096, None, 0, None, None, 0, 3, None) == E.ERR_ARG    # unknown kernel
    assert lib.b200lc_culzss_encode_batch_ex(None, 1, 4096, None, 0, None, None, 0, E.CULZSS_KERNEL_LANE, None) == E.ERR_ARG
    assert lib.b200lc_culzss_encode_batch_ex(None, 0, 4096, None, 0, None, None, 0, E.CULZSS_KERNEL_LANE, None) == E.OK
    assert lib.b200lc_culzss_encode_fast_batch(None, 1, 4096, None, 0, None, None, 0, 0, None) == E.ERR_ARG   # depth 0
    lib.b200lc_inverse_bwt_primary.restype = i32
    lib.b200lc_inverse_bwt_primary.argtypes = [vp, sz, i32, vp, vp, vp, sz, vp]
    lib.b200lc_inverse_bwt_primary_scratch_bytes.restype = sz
    lib.b200lc_inverse_bwt_primary_scratch_bytes.argtypes = [sz]
    assert lib.b200lc_inverse_bwt_primary(None, 0, 1, None, None, None, 0, None) == E.OK
    assert lib.b200lc_inverse_bwt_primary(None, 10, 1, None, None, None, 0, None) == E.ERR_ARG
    assert lib.b200lc_inverse_bwt_primary_scratch_bytes(1 << 20) > (1 << 20) * 16
    assert lib.b200lc_inverse_bwt_primary_scratch_bytes(0) == 256
    one = (ctypes.c_ubyte * 4)(9, 8, 7, 6)
    lib.bsc_st_encode_cuda.restype = i32
    lib.bsc_st_encode_cuda.argtypes = [vp, i32, i32, i32]
    assert lib.bsc_st_encode_cuda(None, 5, 5, 0) == -1                    # LIBBSC_BAD_PARAMETER (st2.cu:371)
    assert lib.bsc_st_encode_cuda(one, 4, 4, 0) == -1 and lib.bsc_st_encode_cuda(one, 4, 9, 0) == -1
    assert lib.bsc_st_encode_cuda(one, 1, 5, 0) == 0 and lib.bsc_st_cuda_init(0) == 0
    assert lib.bsc_bwt_decode(None, 4, 1, 0, None, 0) == -1               # bwt.cpp:361-364
    assert lib.bsc_bwt_decode(one, 4, 0, 0, None, 0) == -1 and lib.bsc_bwt_decode(one, 4, 5, 0, None, 0) == -1
    assert lib.bsc_bwt_decode(one, 1, 1, 0, None, 0) == 0 and one[0] == 9


def test_culzss_container_header_is_validated_before_touching_cuda():
    """u32 nblocks, u32 padding, u32 cumulative_end[nblocks] (cuda-lzss-cluster/culzss.c:220,243-264):
    non-increasing ends, a buffer larger than the worst case, a truncated payload and an output
    buffer that is too small are all rejected on the host."""
    import numpy as np
    lib = b200lc.lib()
    lib.b200lc_culzss_decompress_container.restype = ctypes.c_int
    lib.b200lc_culzss_decompress_container.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                       ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    out = np.zeros(4 << 20, dtype=np.uint8)
    got = ctypes.c_size_t(0)

    def run(words, payload_bytes, cap=out.nbytes):
        buf = np.concatenate([np.asarray(words, dtype=np.uint32).view(np.uint8),
                              np.zeros(payload_bytes, dtype=np.uint8)])
        return lib.b200lc_culzss_decompress_container(buf.ctypes.data, buf.nbytes, out.ctypes.data, cap,
                                                      ctypes.byref(got))

    assert run([2, 0, 1000, 1000], 1000) == b200lc.ERR_ARG          # second buffer is empty
    assert run([2, 0, 2000, 1000], 2000) == b200lc.ERR_ARG          # ends go backwards
    assert run([1, 0, (1 << 20) + 4096], (1 << 20) + 4096) == b200lc.ERR_ARG   # larger than any stored buffer
    assert run([2, 0, 1000, 2000], 1500) == b200lc.ERR_ARG          # payload truncated
    assert run([0, 0], 0) == b200lc.ERR_ARG                         # no buffers
    assert run([1, 1 << 20, 1000], 1000) == b200lc.ERR_ARG          # padding of a whole buffer
    assert run([2, 0, 1000, 2000], 2000, cap=1 << 20) == b200lc.ERR_OVERFLOW


def test_public_headers_compile_standalone():
    """Every header under include/ is self-contained: C99 for the C-ABI ones, C++ for the one that
    keeps the reference's C++ linkage (bzip2_gpu.h) and for the CUHD class mirror."""
    inc = os.path.join(ROOT, "include")
    for h in sorted(glob.glob(os.path.join(inc, "*.h"))):
        name = os.path.basename(h)
        for lang, std in (("c", "-std=c99"), ("c++", "-std=c++11")):
            r = subprocess.run(["gcc", std, "-Wall", "-Werror", "-I", inc, "-x", lang, "-fsyntax-only", "-"],
                               input='#include "%s"\n' % name, capture_output=True, text=True)
            assert r.returncode == 0, "%s as %s:\n%s" % (name, lang, r.stderr)
    for h in sorted(glob.glob(os.path.join(inc, "cuhd_compat", "*.h"))):
        r = subprocess.run(["g++", "-std=c++14", "-I", inc, "-I", os.path.join(inc, "cuhd_compat"),
                            "-I", "/usr/local/cuda/include", "-x", "c++", "-fsyntax-only", "-"],
                           input='#include "%s"\n' % os.path.basename(h), capture_output=True, text=True)
        assert r.returncode == 0, "%s:\n%s" % (h, r.stderr)


def test_cudpp_header_matches_the_reference_abi(tmp_path):
    """Enumerator values, CUDPPConfiguration layout and the handle type of include/cudpp.h equal what
    the reference's cudpp-inpar/include/cudpp.h gives a C++ caller (golden: tools/make_abi_golden.sh)."""
    exe = str(tmp_path / "dump")
    r = subprocess.run(["g++", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "cudpp_enum_dump.cc"), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ours = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert ours == open(os.path.join(ROOT, "tests", "golden", "cudpp_abi.txt")).read()
