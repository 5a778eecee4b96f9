#!/bin/sh
# One GPU call that re-establishes the baseline of a round and tries the opt-in switches:
#   gpurun --timeout 900 -- 'sh tools/round_start.sh'
# Writes everything under gpurun_out/rs_* (read here afterwards; copy summaries into profiles/).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/rs_pytest.log
python bench.py > gpurun_out/rs_bench.json 2> gpurun_out/rs_bench.err
python tools/bench_paths.py all > gpurun_out/rs_paths.jsonl 2> gpurun_out/rs_paths.err
sh tools/try_decode_variants.sh > gpurun_out/rs_variants.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rs_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/rs_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cuhd_decode_kernel -c 1 -f -o gpurun_out/rs_dec \
    python tools/bench_paths.py cuhd --mib 1024 > gpurun_out/rs_ncu_dec.log 2>&1
B200LC_CUHD_PASSA=multi ncu --set full --clock-control none --import-source on -k regex:cuhd_decode_kernel -c 1 -f \
    -o gpurun_out/rs_dec_multi python tools/bench_paths.py cuhd --mib 1024 > gpurun_out/rs_ncu_dec_multi.log 2>&1
tail -2 gpurun_out/rs_pytest.log; cat gpurun_out/rs_bench.json; grep -h -e "===" -e "\"path\": \"cuhd\"" gpurun_out/rs_variants.log
