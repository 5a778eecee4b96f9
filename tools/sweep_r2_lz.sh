#!/bin/sh
# CULZSS rows of the configs[4] sweep again (after the probe-based kernel choice and the last lane-kernel changes)
N=${1:-1}
ARGS="--mib 256 --paths culzss,culzss_lane --entropies 2,4,6,8 --blocks 262144,1048576,4194304"
if [ "$N" = "1" ]; then
    python tools/sweep.py $ARGS > gpurun_out/r02_sweeplz_n1.jsonl 2> gpurun_out/r02_sweeplz_n1.err
else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sweep.py $ARGS > gpurun_out/r02_sweeplz_n$N.jsonl 2> gpurun_out/r02_sweeplz_n$N.err
fi
tail -2 gpurun_out/r02_sweeplz_n$N.err; grep -c '^{' gpurun_out/r02_sweeplz_n$N.jsonl
