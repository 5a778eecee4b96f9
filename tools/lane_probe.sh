#!/bin/sh
# quick GPU check of the CULZSS fast modes: tests + per-kind throughput
python -m pytest tests/test_culzss_gpu.py -q -m gpu -x 2>&1 | tail -3
python tools/bench_paths.py culzss --mib ${1:-1024} > gpurun_out/lane_culzss.jsonl 2> gpurun_out/lane_culzss.err
python - <<P
import json
for l in open("gpurun_out/lane_culzss.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); print(d["data"], "parity %.1f (cta %.1f) GB/s r=%.2f dec %.0f" % (d["encode_gbs"], d["encode_cta_kernel_gbs"], d["ratio"], d["decode_gbs"]), {k:(round(v["encode_gbs"],1), round(v["ratio"],3)) for k,v in d["fast_mode_non_parity"].items()})
P
tail -3 gpurun_out/lane_culzss.err
