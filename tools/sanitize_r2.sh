#!/bin/bash
# compute-sanitizer (racecheck + memcheck) over small parity cases of the kernels added in round 2:
# the CULZSS lane encoders (parity and fast), the probe / selection path, the container gather, the
# libbsc sort transform and the inverse BWT with 32- and 64-bit row entries.  Logs under gpurun_out/.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # tool, tag, pytest args...
    local tool=$1 tag=$2; shift 2
    timeout 600 $CS --tool $tool --print-limit 5 --error-exitcode 9 \
        python -m pytest -x -q -m gpu -p no:cacheprovider "$@" > gpurun_out/san2_${tool}_${tag}.log 2>&1
    echo "$tool $tag rc=$? :: $(grep -E 'passed|failed|error' gpurun_out/san2_${tool}_${tag}.log | tail -1) :: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/san2_${tool}_${tag}.log | tail -1)"
}
for tool in memcheck racecheck; do
    run $tool lz_lane tests/test_culzss_gpu.py -k "(encode_matches_oracle and (quant32 or text or carets or ramp or random)) or small_and_multi or (fast_mode_streams_decode_everywhere and lane and (quant16 or zeros or mix)) or ragged"
    run $tool bsc_st tests/test_bsc_gpu.py -k "(st_encode_matches and (tiny or rand4 or period7 or zeros)) or (bwt_decode_inverts and (n2 or n3 or n9 or periodic or random_70001 or zeros))"
done
