"""CPU: the CUHD decoder's per-thread walks (csrc/cuhd_walks.cuh, the code the kernel inlines)
compiled for the host and checked against a bit-serial decode through the flat LUT -- the decode
contract of cuhd-icpp/src/cuhd_gpu_decoder.cu:16-143 -- on random prefix codes (complete and
incomplete, L = 1..13) and random / all-zero / all-one streams, every entry state."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("walks") / "cuhd_walks_host")
    r = subprocess.run(["g++", "-O2", "-std=c++14", "-Wno-unknown-pragmas",
                        "-I", os.path.join(ROOT, "gpu-lossless-compression_b200", "csrc"),
                        os.path.join(ROOT, "tests", "c", "cuhd_walks_host.cc"), "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_walks_match_bit_serial_decode(exe, seed):
    r = subprocess.run([exe, str(seed), "200"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("ok:")
