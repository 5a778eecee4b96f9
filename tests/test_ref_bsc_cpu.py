"""CPU: BASELINE.json config 1 -- the reference's libbsc CPU path (oracle/_ref/bsc, built from
/root/reference/cuda-bsc without CUDA: its default -m0 BWT never uses the GPU)
round-trips a single block bit-exactly.  This binary is the CPU baseline for the BWT stage."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

BSC = os.path.join(O.ORACLE_DIR, "_ref", "bsc")


def synthetic_largefile(n=3569598, seed=7):
    """Stand-in for the README's testdata/largefile (3,569,598 bytes, not shipped):
    word-structured text."""
    rng = np.random.default_rng(seed)
    words = [b"lossless", b"block", b"sorting", b"compression", b"gpu", b"the", b"of", b"and",
             b"transform", b"huffman", b"suffix", b"array", b"window", b"entropy"]
    idx = rng.integers(0, len(words), n // 5 + 8)
    return b" ".join(words[i] for i in idx)[:n]


@pytest.mark.skipif(not os.path.exists(BSC), reason="oracle/_ref/bsc not built")
def test_bsc_cpu_round_trip(tmp_path):
    data = synthetic_largefile()
    src, comp, back = tmp_path / "in", tmp_path / "out.bsc", tmp_path / "back"
    src.write_bytes(data)
    r = subprocess.run([BSC, "e", str(src), str(comp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "compressed 3569598 into" in r.stdout
    r = subprocess.run([BSC, "d", str(comp), str(back)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0
    assert back.read_bytes() == data
    assert comp.stat().st_size < len(data) // 3
