#!/bin/sh
# Round 2: one GPU call for the state that goes into profiles/ -- full GPU test suite, bench line,
# per-path timings, the ncu launch list of bench.py and ncu --set full of the kernels DESIGN.md quotes.
#   gpurun --timeout 1800 -- 'sh tools/round2_capture.sh'
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -6 > gpurun_out/r2_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python bench.py --table reference --no-paths --no-strong --no-cpu > gpurun_out/r2_bench_reftable.json 2>> gpurun_out/r2_bench.err
python tools/bench_paths.py all > gpurun_out/r2_paths.jsonl 2> gpurun_out/r2_paths.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 4 --input philox --culzss-mib 1024 --cudpp-blocks 256 --strong-blocks 256 \
    > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cuhd_decode_kernel -c 1 -f -o gpurun_out/r2_dec \
    python tools/bench_paths.py cuhd --mib 1024 > gpurun_out/r2_ncu_dec.log 2>&1
# lane kernels of tools/bench_paths.py culzss: launches 0-3 = fast mode (non-parity), 4.. = parity mode
ncu --set full --clock-control none --import-source on -k regex:culzss_encode_lane_kernel -c 1 -f -o gpurun_out/r2_lzenc_lane_fast \
    python tools/bench_paths.py culzss --mib 1024 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:culzss_encode_lane_kernel -s 5 -c 1 -f -o gpurun_out/r2_lzenc_lane_parity \
    python tools/bench_paths.py culzss --mib 1024 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:culzss_decode_kernel -c 1 -f -o gpurun_out/r2_lzdec \
    python tools/bench_paths.py culzss --mib 1024 > /dev/null 2>&1
tail -2 gpurun_out/r2_pytest.log; cat gpurun_out/r2_smoke.log | tail -1; cut -c1-400 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
