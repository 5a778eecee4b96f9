#!/bin/sh
# Round-2 starter: parity + timing of the opt-in CUHD decode switches on a B200.
#   B200LC_CUHD_PASSA=multi   pass A advances over every whole codeword of the window per lookup
#                             (csrc/cuhd_walks.cuh walk_record_multi; CPU-checked by
#                             tests/test_cuhd_walks_cpu.py, never run on a GPU in round 1)
# Usage (from the repo root, e.g. under gpurun):  sh tools/try_decode_variants.sh > gpurun_out/variants.log 2>&1
set -x
for v in "" multi; do
    echo "=== B200LC_CUHD_PASSA='$v'"
    B200LC_CUHD_PASSA=$v timeout 300 python -m pytest tests/test_cuhd_decode_gpu.py -q -m gpu -x 2>&1 | tail -3
    B200LC_CUHD_PASSA=$v timeout 120 python tools/bench_paths.py cuhd --mib 1024
    B200LC_CUHD_PASSA=$v timeout 120 python tools/bench_paths.py cuhd --mib 64
done
