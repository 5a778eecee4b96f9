#!/bin/sh
# BASELINE.json configs[4] at N GPUs (round 2, reduced grid so that 1/2/4/8 GPUs fit the GPU budget):
#   gpurun --gpus N -- 'sh tools/sweep_r2.sh N'
N=${1:-1}
ARGS="--mib 256 --paths cuhd_batch,culzss,culzss_lane,cudpp --entropies 2,4,6,8 --blocks 262144,1048576,4194304"
if [ "$N" = "1" ]; then
    python tools/sweep.py $ARGS --md gpurun_out/r02_sweep_n1.md > gpurun_out/r02_sweep_n1.jsonl 2> gpurun_out/r02_sweep_n1.err
else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sweep.py $ARGS --md gpurun_out/r02_sweep_n$N.md > gpurun_out/r02_sweep_n$N.jsonl 2> gpurun_out/r02_sweep_n$N.err
fi
tail -2 gpurun_out/r02_sweep_n$N.err
tail -5 gpurun_out/r02_sweep_n$N.md
