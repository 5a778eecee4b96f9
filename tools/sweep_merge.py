#!/usr/bin/env python
"""Merges the per-N outputs of tools/sweep_r2.sh into one table: whole-job GB/s at 1/2/4/8 GPUs.
  python tools/sweep_merge.py gpurun_out/r02_sweep_n{1,2,4,8}.jsonl > profiles/r02_sweep.md"""
import json
import sys

rows = {}
ns = []
for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith("{"):
            continue
        r = json.loads(line)
        n = r["n_gpus"]
        if n not in ns:
            ns.append(n)
        rows.setdefault((r["path"], r["H"], r["block"]), {})[n] = r
ns.sort()
print("# BASELINE.json configs[4]: block size x order-0 entropy sweep at %s GPUs of one box (round 2)" % "/".join(map(str, ns)))
print("# tools/sweep_r2.sh N: 256 MiB per GPU and point (cudpp: 128 MiB), device-resident, CUDA events, best of 3, every rank its own")
print("# data (weak scaling: blocks are independent, no data-path collective), times reduced with MAX and sizes with SUM over NCCL;")
print("# GB/s = whole-job uncompressed bytes / time.  culzss = bit-exact parity mode (lane kernel from 160 MiB per call), culzss_lane =")
print("# NON-PARITY fast mode; cuhd_batch = one shared table, one launch per direction.  round trip checked on every rank.")
print()
hdr = "| path | H | block | " + " | ".join("enc N=%d" % n for n in ns) + " | " + " | ".join("dec N=%d" % n for n in ns) + \
      " | ratio | enc eff N=%d | dec eff N=%d | ok |" % (ns[-1], ns[-1])
print(hdr)
print("|" + "---|" * (hdr.count("|") - 1))
for key in rows:
    r = rows[key]
    f = lambda v: "-" if v is None else "%.1f" % v  # noqa: E731
    base, top = r[ns[0]], r[ns[-1]]
    eff = lambda k: "-" if not base.get(k) or not top.get(k) else "%.2f" % (top[k] / base[k] / (ns[-1] / ns[0]))  # noqa: E731
    print("| %s | %g | %d KiB | " % (key[0], key[1], key[2] // 1024) +
          " | ".join(f(r[n]["encode_gbs"]) if n in r else "-" for n in ns) + " | " +
          " | ".join(f(r[n]["decode_gbs"]) if n in r else "-" for n in ns) +
          " | %.3f | %s | %s | %s |" % (base["ratio"] or 0, eff("encode_gbs"), eff("decode_gbs"),
                                        "ok" if all(v["round_trip"] for v in r.values()) else "FAIL"))
