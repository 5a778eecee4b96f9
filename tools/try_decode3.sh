#!/bin/sh
# Round 2: parity + timing of the CUHD decoder (packed write pass) on a B200, with its tuning knobs.
#   B200LC_CUHD_MULTI_BITS=n   window of the counting table and of the write table
#   B200LC_CUHD_VARIANT=i      pins K (0: 8, 1: 4, 2: 2, 3: 1, 4: 16, 5: 32 warp-steps per segment)
TAG=${1:-d3}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuhd_decode_gpu.py tests/test_cuhd_encode_gpu.py -q -m gpu -x 2>&1 | tail -15
for cfg in "" "B200LC_CUHD_MULTI_BITS=12" "B200LC_CUHD_MULTI_BITS=14" "B200LC_CUHD_VARIANT=4"; do
    echo "=== $cfg"
    env $cfg timeout 60 python tools/bench_paths.py cuhd --mib 1024 | grep -o '"decode_ms": [0-9.]*'
    env $cfg timeout 60 python tools/bench_paths.py cuhd --mib 64 | grep -o '"decode_ms": [0-9.]*'
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cuhd_decode_kernel -c 1 -f -o gpurun_out/${TAG}_dec \
    python tools/bench_paths.py cuhd --mib 1024 > gpurun_out/${TAG}_ncu.log 2>&1
