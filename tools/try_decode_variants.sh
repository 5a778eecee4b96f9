#!/bin/sh
# Round-2 starter: parity + timing of the opt-in CUHD decode switches on a B200.
#   B200LC_CUHD_PASSA=multi   pass A advances over every whole codeword of the window per lookup
#                             (csrc/cuhd_walks.cuh walk_record_multi)
#   B200LC_CUHD_WRITE=2       write-table layout 2 + running store pointer (walk_write2)
# Both are CPU-checked by tests/test_cuhd_walks_cpu.py and were never run on a GPU in round 1.
# Each alone compiles to 56 registers like the default; both together to 67 (3 CTAs/SM).
# Usage (from the repo root, e.g. under gpurun):  sh tools/try_decode_variants.sh > gpurun_out/variants.log 2>&1
for a in "" multi; do
    for w in "" 2; do
        echo "=== B200LC_CUHD_PASSA='$a' B200LC_CUHD_WRITE='$w'"
        export B200LC_CUHD_PASSA=$a B200LC_CUHD_WRITE=$w
        timeout 300 python -m pytest tests/test_cuhd_decode_gpu.py -q -m gpu -x 2>&1 | tail -3
        timeout 120 python tools/bench_paths.py cuhd --mib 1024
        timeout 120 python tools/bench_paths.py cuhd --mib 64
    done
done
