#!/bin/bash
# compute-sanitizer (racecheck + memcheck) over small GPU parity cases; logs under gpurun_out/.
# Usage (on a GPU box): bash tools/sanitize.sh
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool, tag, pytest args...
    local tool=$1 tag=$2; shift 2
    timeout 420 $CS --tool $tool --print-limit 5 --error-exitcode 9 \
        python -m pytest -x -q -m gpu -p no:cacheprovider "$@" > gpurun_out/san_${tool}_${tag}.log 2>&1
    echo "$tool $tag rc=$? :: $(grep -E 'passed|failed|error' gpurun_out/san_${tool}_${tag}.log | tail -1) :: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/san_${tool}_${tag}.log | tail -1)"
}
for tool in racecheck memcheck; do
    run $tool cuhd_dec tests/test_cuhd_decode_gpu.py -k "(zipf_sizes and (4097 or 65536 or 100)) or two_symbols or fixed_3bit or without_pad or (batch_decode_equals and (sizes2 or sizes3))"
    run $tool cuhd_enc tests/test_cuhd_encode_gpu.py -k "(reference_and_oracle and (8193 or 100000 or 100)) or unaligned or overflow or (encode_blocks and (300000 or 1000-4096 or 350007))"
    run $tool culzss tests/test_culzss_gpu.py -k "small_and_multi or hostile or unaligned_offsets"
    run $tool cudpp tests/test_cudpp_gpu.py -k "compress_small_blocks or (inverse_mtf_sizes and (2049 or 4097)) or periodic or (round_trip and (4095 or 4097 or 8192))"
    run $tool bzip2 tests/test_bzip2_gpu.py -k "rotation_order or (block_sort_arrays and 30001) or ((mtf_rle_matches or send_mtf_values_matches) and (two_symbols or binary_runs or random_50k or n1))"
    run $tool prims tests/test_prims_gpu.py -k "(sort_pairs_matches and (4095 or 4097 or 50000 or 300000)) or (scans and (4097 or 31-))"
    run $tool bsc tests/test_bsc_gpu.py -k "matches_oracle and (n9 or periodic or random_70001)"
done
