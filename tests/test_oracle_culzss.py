"""CPU: pins oracle/culzss_oracle.c against the reference's CPU stage (aftercompression_wrapper
in oracle/_ref/libref_culzss.so) and the committed golden buffers."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["culzss_quant.npz", "culzss_text.npz"])
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLD, name))
    data, comp = g["data"], g["comp"]
    tokens = O.culzss_oracle_tokens(data)
    assert np.array_equal(tokens[0::2], g["token_lens"]) and np.array_equal(tokens[1::2], g["token_offs"])
    ok, mine = O.culzss_oracle_aftercomp(tokens, data.size)
    assert ok == 1 and np.array_equal(mine, comp)
    dok, back = O.culzss_oracle_decompress(comp, data.size)
    assert dok == 1 and np.array_equal(back, data)


@pytest.mark.skipif(not O.have_ref("culzss"), reason="oracle/_ref/libref_culzss.so not built")
@pytest.mark.parametrize("kind", ["quant32", "quant16", "zeros", "random", "text"])
def test_aftercomp_matches_reference_cpu_stage(kind):
    rng = np.random.default_rng(1)
    n = 1 << 17
    data = {"quant32": O.quant_codes(n), "quant16": O.quant_codes(n, dtype=np.uint16),
            "zeros": np.zeros(n, np.uint8), "random": rng.integers(0, 256, n, dtype=np.uint8),
            "text": np.frombuffer((b"lorem ipsum dolor sit amet " * 6000)[:n], np.uint8).copy()}[kind]
    tokens = O.culzss_oracle_tokens(data)
    ok, mine = O.culzss_oracle_aftercomp(tokens, n)
    rok, ref = O.culzss_ref_aftercomp(tokens, data)
    assert ok == rok
    if ok:
        assert np.array_equal(mine, ref)


def test_token_rules():
    # all spaces: the window is pre-filled with ' ' so position 0 already matches 127 bytes
    data = np.full(4096, 0x20, np.uint8)
    tok = O.culzss_oracle_tokens(data)
    assert tok[0] == 127 and tok[1] == 0
    # last chunk: match lengths shrink with the distance to the packet end (gpu_compress.cu:313-317)
    assert tok[2 * 4095] == 1 and tok[2 * 4000] <= 128 - (4000 - 3968)
    # no byte repeats within 128 positions: everything literal
    data = (np.arange(4096) % 251).astype(np.uint8)
    tok = O.culzss_oracle_tokens(data)
    assert np.all(tok[0::2] == 1) and np.array_equal(tok[1::2], data)
