#!/usr/bin/env python
"""bench.py -- headline benchmark: CUHD-format Huffman encode+decode of 1 GiB of Zipf(1.1) bytes
per GPU (BASELINE.json configs[1]), GB/s of uncompressed data, with the roofline of the decode
kernel and a CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mib M]

One "step" = one pass of the hot path over one batch: histogram -> code table -> bit-pack
(encode) -> self-synchronising decode of the whole buffer.  `value` is measured with the input
resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` is the same step
through the host-buffer C ABI (b200lc_cuhd_session_encode/_decode) with pinned host buffers and
the H2D/D2H copies inside the timed region.  N > 1: one process per GPU (torchrun), every rank
works on its own independent buffer (weak scaling, no data-path collective; one all_gather of
the per-rank compressed sizes = the "block offsets" exchange).

--impl reference times the reference's own CPU code for this path (oracle/_ref: llhuff
encode_memory; decode = the oracle's serial LUT walk, the reference ships no CPU decoder) on all
host cores, on a bounded sample of the same workload.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MIB = 1 << 20
ZIPF_ALPHA = 1.1
SEED = 12345
MAX_LEN = 11


def zipf_cdf(nsym=256, alpha=ZIPF_ALPHA):
    import numpy as np
    p = 1.0 / np.arange(1, nsym + 1, dtype=np.float64) ** alpha
    return np.cumsum(p / p.sum())


def gen_zipf_gpu(n, device, seed):
    """n Zipf(1.1) bytes on the GPU (inverse-CDF sampling, Philox generator, fixed seed)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cdf = torch.from_numpy(zipf_cdf()).to(device=device, dtype=torch.float32)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    chunk = 1 << 26
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        u = torch.rand(m, generator=g, device=device)
        out[lo:lo + m] = torch.searchsorted(cdf, u).clamp_(max=255).to(torch.uint8)
    return out


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def profiled_traffic(mib):
    """DRAM bytes per launch of the decode kernel (dram__bytes_read.sum + dram__bytes_write.sum) from
    the committed `ncu --set full` capture of the same workload (profiles/, 1 GiB only), else None."""
    if mib != 1024:
        return None, None
    path = os.path.join(ROOT, "profiles", "r01_cuhd_decode_ncu_full_summary.txt")
    try:
        total = 0.0
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[-1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[-2]]
        return (total or None), "profiles/r01_cuhd_decode_ncu_full_summary.txt"
    except Exception:
        return None, None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------ CPU legs
def cpu_encode_decode(data_np, threads):
    """Reference CPU path on `data_np` split into `threads` independent chunks: table build +
    llhuff encode_memory (oracle/_ref when present, else the oracle port) + serial LUT decode.
    Returns (seconds, kind)."""
    import numpy as np
    import oracle_lib as O
    use_ref = O.have_ref("cuhd")
    chunks = np.array_split(data_np, threads)
    ok = [True] * threads

    def work(i):
        d = np.ascontiguousarray(chunks[i])
        if use_ref:
            code, length, lut, units = O.cuhd_ref_encode(d)
        else:
            code, length, lut, _ = O.cuhd_make_case(d, use_ref=False)
            units, _ = O.cuhd_oracle_encode(d, code, length)
        units = np.concatenate([units, np.zeros(1, np.uint32)])
        out, got = O.cuhd_oracle_decode(units, lut, d.size)
        # the reference drops the tail of a codeword split across the last unit:
        # compare all but the final symbol
        ok[i] = got == d.size and bool(np.array_equal(out[:-1], d[:-1]))

    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    if not all(ok):
        raise RuntimeError("CPU baseline round trip failed")
    return dt, ("reference" if use_ref else "port")


def run_reference(args, rank, world):
    import numpy as np
    import oracle_lib as O
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_mib = args.ref_mib
    n = sample_mib * MIB
    data = O.zipf_bytes(n, ZIPF_ALPHA, seed=SEED)
    for _ in range(args.warmup):
        cpu_encode_decode(data[: n // 8], cores)
    times = []
    kind = "port"
    for _ in range(args.steps):
        dt, kind = cpu_encode_decode(data, cores)
        times.append(dt)
    total = sum(times)
    gbs = n * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": "encode+decode GB/s (uncompressed)", "value": gbs,
        "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "CUHD Huffman encode+decode, Zipf(1.1) bytes, 256-symbol "
                               "length-limited (11 bit) canonical code",
                   "bytes_per_step": n, "sample": "%d MiB of the 1 GiB workload per step" % sample_mib},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": "%d MiB Zipf(1.1) split in %d independent chunks; llhuff "
                                   "table+encode_memory (reference) + serial LUT decode (oracle port)"
                                   % (sample_mib, cores)},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("gpu-lossless-compression_b200")
    L = pkg.lib()

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n = args.mib * MIB
    data = gen_zipf_gpu(n, dev, SEED + rank)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    # persistent buffers (nothing is allocated inside the timed region)
    units_cap = (n * MAX_LEN + 31) // 32 + 2
    units = torch.empty(units_cap, dtype=torch.int32, device=dev)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    hist = torch.empty(256, dtype=torch.int64, device=dev)
    h_hist = torch.empty(256, dtype=torch.int64).pin_memory()
    d_code = torch.empty(256, dtype=torch.int32, device=dev)
    d_len = torch.empty(256, dtype=torch.uint8, device=dev)
    d_lut = torch.empty((1 << MAX_LEN, 2), dtype=torch.uint8, device=dev)
    h_code = torch.empty(256, dtype=torch.int32).pin_memory()
    h_len = torch.empty(256, dtype=torch.uint8).pin_memory()
    h_lut = torch.empty((1 << MAX_LEN, 2), dtype=torch.uint8).pin_memory()
    bits = torch.zeros(1, dtype=torch.int64, device=dev)
    enc_scratch = torch.empty(L.b200lc_cuhd_encode_scratch_bytes(n), dtype=torch.uint8, device=dev)
    dec_scratch = torch.empty(L.b200lc_cuhd_decode_scratch_bytes(units_cap), dtype=torch.uint8,
                              device=dev)
    piece_hist = torch.empty(max(1, L.b200lc_cuhd_piece_hist_bytes(n) // 4), dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    state = {}

    def step(timed):
        """hist (+ per-piece histograms) -> table -> one-pass pack -> decode, all on `stream`;
        returns (enc_ms, dec_ms) if timed."""
        if timed:
            ev[0].record(stream)
        pkg.check(L.b200lc_histogram_u8_pieces(data.data_ptr(), n, hist.data_ptr(), piece_hist.data_ptr(), sp),
                  "hist")
        h_hist.copy_(hist, non_blocking=True)
        stream.synchronize()                      # the code table is built on the host
        pkg.check(L.b200lc_cuhd_build_table(h_hist.data_ptr(), MAX_LEN, h_code.data_ptr(),
                                            h_len.data_ptr(), h_lut.data_ptr()), "table")
        d_code.copy_(h_code, non_blocking=True)
        d_len.copy_(h_len, non_blocking=True)
        d_lut.copy_(h_lut, non_blocking=True)
        pkg.check(L.b200lc_cuhd_encode_planned(data.data_ptr(), n, d_code.data_ptr(), d_len.data_ptr(),
                                               piece_hist.data_ptr(), units.data_ptr(), units_cap,
                                               bits.data_ptr(), enc_scratch.data_ptr(), enc_scratch.numel(),
                                               sp), "encode")
        n_units = L.b200lc_cuhd_compressed_units(h_hist.data_ptr(), h_len.data_ptr()) + 1
        if timed:
            ev[1].record(stream)
        pkg.check(L.b200lc_cuhd_decode(units.data_ptr(), n_units, out.data_ptr(), n,
                                       d_lut.data_ptr(), MAX_LEN, dec_scratch.data_ptr(),
                                       dec_scratch.numel(), sp), "decode")
        state["n_units"] = n_units
        if timed:
            ev[2].record(stream)
            ev[2].synchronize()
            return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        return None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False)
    torch.cuda.synchronize()
    assert torch.equal(out, data), "round trip mismatch"
    assert int(bits.item()) > 0

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    enc_ms, dec_ms = [], []
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    for _ in range(args.steps):
        e, d = step(True)
        enc_ms.append(e)
        dec_ms.append(d)
    t_all1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_all0.elapsed_time(t_all1)
    assert torch.equal(out, data), "round trip mismatch after timed region"

    # ---------------------------------------------------------------- e2e: host buffers, C ABI
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.copy_(data)
    h_units = torch.empty(units_cap, dtype=torch.int32).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    sess = pkg.CuhdSession(n)
    nu = sess.encode(h_in, h_units, h_code, h_len, h_lut, MAX_LEN)   # warm-up
    sess.decode(h_units, nu + 1, h_lut, h_out, MAX_LEN)
    assert torch.equal(h_out, h_in), "e2e round trip mismatch"
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nu = sess.encode(h_in, h_units, h_code, h_len, h_lut, MAX_LEN)
        sess.decode(h_units, nu + 1, h_lut, h_out, MAX_LEN)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    sess.close()
    h2d = n + (nu + 1) * 4 + 256 * 5 + (2 << MAX_LEN)
    d2h = n + (nu + 1) * 4 + 256 * 8 + 8

    # ---------------------------------------------------------------- reduce over ranks
    stats = torch.tensor([total_ms, e2e_s, statistics.mean(dec_ms), statistics.mean(enc_ms),
                          float(state["n_units"])], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([state["n_units"]], dtype=torch.int64, device=dev))
        offsets = torch.cumsum(torch.cat(sizes), 0)   # block offsets of the concatenated stream
        assert int(offsets[-1]) > 0
        total_ms, e2e_s = float(mx[0]), float(mx[1])
        dec_mean, enc_mean = float(mx[2]), float(mx[3])
    else:
        dec_mean, enc_mean = float(stats[2]), float(stats[3])
    if rank != 0:
        return

    peak, peak_src = measured_peak()
    traffic, traffic_src = profiled_traffic(args.mib)
    n_units = state["n_units"]
    dec_bytes = 4 * n_units + n + (2 << MAX_LEN)          # algorithmic bytes of one decode launch
    dec_gbs = dec_bytes / (dec_mean * 1e-3) / 1e9
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle_lib as O
        sample = O.zipf_bytes(args.cpu_mib * MIB, ZIPF_ALPHA, seed=SEED)
        dt, kind = cpu_encode_decode(sample, 1)
        cpu = {"value": sample.size / dt / 1e9, "unit": "GB/s", "cores": 1, "kind": kind,
               "sample": "%d MiB Zipf(1.1): llhuff table + encode_memory (oracle/_ref) and serial "
                         "LUT decode (oracle port), 1 thread" % args.cpu_mib}

    line = {
        "metric": "encode+decode GB/s (uncompressed)", "value": value, "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "CUHD Huffman encode+decode, %d MiB Zipf(1.1) bytes per GPU, "
                               "256-symbol length-limited (11 bit) canonical code" % args.mib,
                   "bytes_per_step_per_gpu": n, "compressed_bytes": 4 * n_units,
                   "ratio": n / (4.0 * n_units), "parallelism": "independent buffer per GPU",
                   "l2": "inputs (%d MiB) exceed the 126 MB L2; no explicit flush" % args.mib},
        "encode_gbs": world * n / (enc_mean * 1e-3) / 1e9,
        "decode_gbs": world * n / (dec_mean * 1e-3) / 1e9,
        "roofline": {"kernel": "cuhd_decode_kernel", "bound": "hbm", "achieved": dec_gbs,
                     "peak": peak, "unit": "GB/s", "frac": dec_gbs / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes": dec_bytes, "launch_ms": dec_mean},
        "e2e": {"value": world * n / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "b200lc_cuhd_session_encode + b200lc_cuhd_session_decode, pinned host buffers"},
        "gpu_launches": 6 * args.steps,   # piece histograms, their reduction, piece bits, plan, pack, decode
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mib", type=int, default=1024, help="uncompressed MiB per GPU per step")
    ap.add_argument("--cpu-mib", type=int, default=128, help="CPU baseline sample size")
    ap.add_argument("--ref-mib", type=int, default=256, help="--impl reference sample per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
