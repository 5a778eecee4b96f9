#!/bin/sh
# Round 2: parity + timing of the warp-autonomous CUHD decoder on a B200, with its tuning knobs.
#   B200LC_CUHD_MINB=5         register bound for 5 CTAs/SM (48 regs) instead of 4 (64 regs)
#   B200LC_CUHD_MULTI_BITS=n   window of the counting table
#   B200LC_CUHD_VARIANT=i      pins K (0: 16, 1: 8, 5: 32 warp-steps per segment)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_cuhd_decode_gpu.py -q -m gpu -x 2>&1 | tail -15
for cfg in "" "B200LC_CUHD_MULTI_BITS=14" "B200LC_CUHD_VARIANT=1" "B200LC_CUHD_MULTI_BITS=12 B200LC_CUHD_MINB=5"; do
    echo "=== $cfg"
    env $cfg timeout 40 python tools/bench_paths.py cuhd --mib 1024
    env $cfg timeout 40 python tools/bench_paths.py cuhd --mib 64
done
