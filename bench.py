#!/usr/bin/env python
"""bench.py -- headline benchmark: CUHD-format Huffman encode+decode of 1 GiB of Zipf(1.1) bytes
per GPU (BASELINE.json configs[1]), GB/s of uncompressed data, with the roofline of the decode
kernel and a CPU baseline; plus, in the same JSON line, the other two hot paths at BASELINE.json's
sizes (`paths`: CULZSS 4 GiB, cudppCompress 1024 x 1 MiB blocks per GPU) and the sharded
configs[3] run (`c4_strong`: 8192 x 1 MiB blocks split over the ranks, one gather of block sizes).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mib M]
                  [--table own|reference] [--input mt19937|philox] [--no-paths] [--no-strong]

One "step" = one pass of the hot path over one batch: per-piece histograms -> code table ->
one-pass bit-pack (encode) -> self-synchronising decode of the whole buffer.  `value` is measured
with the input resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` is the
same step through the host-buffer C ABI (b200lc_cuhd_session_encode/_decode) with pinned host
buffers and the H2D/D2H copies inside the timed region, two sessions on two host threads so that
the encode of step k + 1 shares the bus with the decode of step k (both PCIe directions busy).
N > 1: one process per GPU (torchrun), every rank works on its own independent buffer (weak
scaling, no data-path collective).  `c4_strong` is the strong-scaling leg: the same 8192 blocks
whatever N, contiguous block ranges per rank (shard.plan_blocks), per-block compressed sizes
gathered with ONE all_gather over NCCL (shard.gather_offsets) and checked against a scan on rank 0.

--table reference takes the dictionary from tests/golden/cuhd_c2_table.npz -- the reference
encoder's own code table for SURVEY.md 8(d)'s C2 input (llhuffman_encoder.cc:18-198, made by
tools/make_golden.py c2) -- instead of building one inside the step; the line says which was used.

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


def gen_zipf_mt19937(n, device, seed):
    """SURVEY.md 8(d)'s C2 input: out[i] = min(searchsorted(cdf, u[i]), 255) with
    u = numpy Generator(MT19937(seed)).random(n) -- the generator of tests/oracle_lib.zipf_bytes,
    restated here (bucket table of the inverse CDF, chunks uploaded as they are made)."""
    import numpy as np
    import torch
    rng = np.random.Generator(np.random.MT19937(seed))
    cdf = zipf_cdf()
    edges = np.arange(65537, dtype=np.float64) / 65536.0
    lo = np.searchsorted(cdf, edges[:-1])
    hi = np.searchsorted(cdf, np.nextafter(edges[1:], 0.0))
    direct = np.minimum(lo, 255).astype(np.uint8)
    ambiguous = lo != hi
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 24
    for at in range(0, n, step):
        u = rng.random(min(step, n - at))
        b = (u * 65536.0).astype(np.int32)
        part = direct[b]
        m = ambiguous[b]
        part[m] = np.minimum(np.searchsorted(cdf, u[m]), 255).astype(np.uint8)
        out[at:at + part.size] = torch.from_numpy(part).to(device, non_blocking=False)
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
    the newest committed `ncu --set full` capture of the same workload (profiles/, 1 GiB only), else None."""
    if mib != 1024:
        return None, None
    for name in ("r02_cuhd_decode_ncu_full_summary.txt", "r01_cuhd_decode_ncu_full_summary.txt"):
        path = os.path.join(ROOT, "profiles", name)
        try:
            total = 0.0
            for line in open(path):
                f = line.split()
                if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(f[-1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[-2]]
            if total:
                return total, "profiles/" + name
        except Exception:
            continue
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
def _events(n):
    import torch
    return [torch.cuda.Event(enable_timing=True) for _ in range(n)]


def _best_ms(fn, iters, warm=1):
    """Best CUDA-event time of fn() on the current stream."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        a, b = _events(2)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def _reduce_max(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def _reduce_sum(vals, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def reference_dictionary():
    """The reference encoder's dictionary for C2 (tests/golden/cuhd_c2_table.npz)."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "cuhd_c2_table.npz"))
    return z["code"].astype(np.uint32), z["length"].astype(np.uint8), np.ascontiguousarray(z["lut"]).astype(np.uint8)


def path_culzss(pkg, dev, rank, world, mib, peak, with_cpu):
    """C3: CULZSS encode + decode of `mib` MiB of cuSZ-like quantisation codes in 1 MiB buffers,
    device-resident through b200lc_culzss_encode_batch / _decode_batch; e2e through the host-buffer
    container call on a bounded sample."""
    import ctypes as C
    import numpy as np
    import torch
    from bench_paths import quant_codes_gpu
    L = pkg.lib()
    n, buf = mib * MIB, MIB
    nbuf = n // buf
    data = quant_codes_gpu(n, dev, seed=2024 + rank, itemsize=4)
    stride = pkg.culzss_out_stride(buf)
    out = torch.empty(nbuf * stride, dtype=torch.uint8, device=dev)
    clen = torch.empty(nbuf, dtype=torch.int32, device=dev)
    scratch = torch.empty(L.b200lc_culzss_encode_scratch_bytes(nbuf, buf), dtype=torch.uint8, device=dev)
    # fast mode (NON-PARITY: the reference's format, not the reference encoder's bytes), lane formulation
    fast_ms = _best_ms(lambda: pkg.culzss_encode(data, buf, out, clen, scratch, fast="lane"), iters=3, warm=1)
    fcl = clen.cpu().numpy().astype(np.int64)
    fast_bytes = int(np.where(fcl == 0, buf, fcl).sum())
    enc_ms = _best_ms(lambda: pkg.culzss_encode(data, buf, out, clen, scratch), iters=2, warm=1)
    cl = clen.cpu().numpy().astype(np.int64)
    sizes = np.where(cl == 0, buf, cl)
    offs = np.zeros(nbuf + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    comp = torch.empty(int(offs[-1]), dtype=torch.uint8, device=dev)
    rows = out.view(nbuf, stride)
    col = torch.arange(stride, device=dev)
    d_cl = torch.from_numpy(cl).to(dev)
    for lo in range(0, nbuf, 256):              # pack the buffers back to back (container layout)
        hi = min(nbuf, lo + 256)
        mask = col[None, :] < d_cl[lo:hi, None]
        comp[int(offs[lo]):int(offs[hi])] = rows[lo:hi][mask]
    assert int((cl == 0).sum()) == 0, "quantisation codes always compress"
    del out, rows, mask
    d_offs = torch.from_numpy(offs).to(dev)
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    dscratch = torch.empty(L.b200lc_culzss_decode_scratch_bytes(nbuf, buf), dtype=torch.uint8, device=dev)
    dec_ms = _best_ms(lambda: pkg.culzss_decode(comp, d_offs, buf, dec, dscratch), iters=5, warm=1)
    assert torch.equal(dec, data), "CULZSS round trip mismatch"
    cbytes = int(offs[-1])

    # e2e: host buffers through the container call (what the reference CLI does around its kernels)
    e2e_mib = min(mib, 1024 if world == 1 else 512)      # pinned host memory is shared by the ranks
    h_in_t = torch.empty(e2e_mib * MIB, dtype=torch.uint8).pin_memory()      # pinned host buffers
    h_in_t.copy_(data[: e2e_mib * MIB])
    h_in = h_in_t.numpy()
    u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
    L.b200lc_culzss_container_bound.restype = C.c_size_t
    L.b200lc_culzss_container_bound.argtypes = [C.c_size_t]
    for f in (L.b200lc_culzss_compress_container, L.b200lc_culzss_decompress_container):
        f.restype = C.c_int
        f.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
    cap = L.b200lc_culzss_container_bound(h_in.size)
    h_comp_t = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    h_back_t = torch.zeros(h_in.size, dtype=torch.uint8).pin_memory()
    h_comp, h_back = h_comp_t.numpy(), h_back_t.numpy()
    olen, blen = C.c_size_t(0), C.c_size_t(0)
    del data, dec, comp
    torch.cuda.empty_cache()
    e2e_s = 1e30
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pkg.check(L.b200lc_culzss_compress_container(h_in, h_in.size, h_comp, cap, C.byref(olen)), "container")
        pkg.check(L.b200lc_culzss_decompress_container(h_comp, olen.value, h_back, h_back.size, C.byref(blen)), "container")
        e2e_s = min(e2e_s, time.perf_counter() - t0)
    assert blen.value == h_in.size and np.array_equal(h_back, h_in), "CULZSS container round trip mismatch"
    e2e_bytes = int(h_in.size)
    # the CPU baseline below works on the first buffers of the same sample: keep a pageable copy
    cpu_sample = h_in[: min(e2e_mib, 2 * (os.cpu_count() or 1)) * MIB].copy() if with_cpu else None
    del h_in, h_comp, h_back, h_in_t, h_comp_t, h_back_t
    if hasattr(torch._C, "_host_emptyCache"):
        torch._C._host_emptyCache()

    enc_max, dec_max, e2e_max, fast_max = _reduce_max([enc_ms, dec_ms, e2e_s, fast_ms], dev, world)
    csum, fsum = _reduce_sum([float(cbytes), float(fast_bytes)], dev, world)
    res = {
        "workload": "CULZSS encode+decode, %d MiB of cuSZ-like int32 quantisation codes per GPU, 1 MiB buffers, "
                    "4 KiB packets, W = 128 (bit-exact parity mode)" % mib,
        "encode_gbs": world * n / enc_max / 1e6, "decode_gbs": world * n / dec_max / 1e6,
        "value": world * n / (enc_max + dec_max) / 1e6, "unit": "GB/s",
        "encode_ms": enc_max, "decode_ms": dec_max, "ratio": world * n / csum,
        "roofline": {"kernel": "culzss_decode_kernel", "bound": "hbm", "achieved": (n + cbytes) / dec_ms / 1e6,
                     "peak": peak, "unit": "GB/s", "frac": (n + cbytes) / dec_ms / 1e6 / peak,
                     "algorithmic_bytes": n + cbytes},
        "encode_roofline": {"kernel": "culzss_encode_lane_kernel<parity> (packet per lane; batches under 160 MiB: culzss_encode_kernel<0>)", "bound": "integer issue (parity mode), reported against hbm",
                            "achieved": (n + cbytes) / enc_ms / 1e6, "peak": peak, "unit": "GB/s",
                            "frac": (n + cbytes) / enc_ms / 1e6 / peak},
        "fast_mode_non_parity": {"encode_gbs": world * n / fast_max / 1e6, "encode_ms": fast_max, "ratio": world * n / fsum,
                                 "mode": "lane", "frac": (n + fast_bytes) / fast_ms / 1e6 / peak,
                                 "note": "b200lc_culzss_encode_fast_batch(depth = B200LC_CULZSS_FAST_LANE): same buffer/"
                                         "token format (decodes with the reference DecodeKernel), one packet per lane, "
                                         "greedy parse through a lane-private hash; NOT bit-exact with the reference "
                                         "encoder"},
        "e2e": {"value": world * e2e_bytes / e2e_max / 1e9, "unit": "GB/s", "sample_mib": e2e_mib,
                "h2d_bytes_per_step": int(e2e_bytes + olen.value), "d2h_bytes_per_step": int(e2e_bytes + olen.value),
                "api": "b200lc_culzss_compress_container + _decompress_container, pinned host buffers, "
                       "best of 3 calls (the work area is kept between calls)"},
    }
    if with_cpu:
        import oracle_lib as O
        cores = os.cpu_count() or 1
        bufs = [cpu_sample[i * MIB:(i + 1) * MIB] for i in range(cpu_sample.size // MIB)]

        def one(b):
            ok, c = O.culzss_oracle_compress(b)
            back = O.culzss_oracle_decompress(c)[1] if ok else b
            return bool(np.array_equal(back[:MIB], b))
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            good = list(ex.map(one, bufs))
        dt = time.perf_counter() - t0
        assert all(good)
        res["cpu_baseline"] = {"value": len(bufs) * MIB / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                               "sample": "%d buffers of 1 MiB: oracle/culzss_oracle.c encode + decode" % len(bufs)}
    return res


# Blocks per cudppCompress call.  The kernels are batched over blocks: 64 / 128 / 256 / 512 blocks per
# call give 7.3 / 8.2 / 8.6 / 8.9 GB/s encode and 7.1 / 10.4 / 12.9 / 15.1 GB/s decode on Zipf(1.3)
# (tools/bench_paths.py) -- the inverse-BWT walks are latency-bound and want walkers.
CUDPP_BATCH = 512


def _cudpp_run(pkg, dev, first_block, nblocks, batch, iters, seed0=95835, check=True):
    """Compress blocks [first_block, first_block + nblocks) of the C4 input (block b is generated
    from seed0 + b // batch ... on the GPU) in batches; returns (encode_ms, decode_ms, sizes
    int64[nblocks] in words, compressed words)."""
    import torch
    from bench_paths import cudpp_blocks_gpu
    L = pkg.lib()
    n = MIB
    scratch = torch.empty(max(L.b200lc_cudpp_compress_scratch_bytes(batch, n),
                              L.b200lc_cudpp_decompress_scratch_bytes(batch, n)) + 256, dtype=torch.uint8, device=dev)
    back = torch.empty(batch * n, dtype=torch.uint8, device=dev)
    sizes = torch.zeros(nblocks, dtype=torch.int64, device=dev)
    res = None
    enc_ms = dec_ms = 0.0
    kinds = ["zipf", "markov"]
    e = _events(3)
    for lo in range(0, nblocks, batch):
        nb = min(batch, nblocks - lo)
        g = (first_block + lo) // batch
        data = cudpp_blocks_gpu(nb, n, dev, kinds[g % 2], seed=seed0 + g)
        best_e = best_d = 1e30
        for _ in range(iters):
            e[0].record()
            res = pkg.cudpp_compress_batch(data, nb, n, scratch=scratch, out=res if nb == batch else None)
            e[1].record()
            _, derr = pkg.cudpp_decompress_batch(res, nb, n, scratch=scratch, out=back[: nb * n])
            e[2].record()
            e[2].synchronize()
            best_e = min(best_e, e[0].elapsed_time(e[1]))
            best_d = min(best_d, e[1].elapsed_time(e[2]))
        enc_ms += best_e
        dec_ms += best_d
        sizes[lo:lo + nb] = res.total_words[:nb].to(torch.int64)
        if check:
            assert int(res.error.item()) == 0 and int(derr.item()) == 0
            assert torch.equal(back[: nb * n], data), "cudppCompress round trip mismatch"
    return enc_ms, dec_ms, sizes


def path_cudpp(pkg, dev, rank, world, blocks, peak, with_cpu):
    """C4, one GPU's share: `blocks` independent 1 MiB blocks through BWT + MTF + Huffman
    (b200lc_cudpp_compress_batch) and back (b200lc_cudpp_decompress_batch), CUDPP_BATCH blocks per call."""
    import numpy as np
    import torch
    n = MIB
    enc_ms, dec_ms, sizes = _cudpp_run(pkg, dev, rank * blocks, blocks, CUDPP_BATCH, iters=1)
    words = int(sizes.sum().item())
    enc_max, dec_max = _reduce_max([enc_ms, dec_ms], dev, world)
    (wsum,) = _reduce_sum([float(words)], dev, world)
    N = blocks * n
    alg = N + 4 * words + blocks * 4 * (256 + 256 + 2)
    res = {
        "workload": "cudppCompress (BWT + MTF + Huffman) and its inverse, %d blocks of 1 MiB per GPU "
                    "(Zipf(1.3) / Markov bytes in 1..255, last byte 0), %d blocks per call" % (blocks, CUDPP_BATCH),
        "encode_gbs": world * N / enc_max / 1e6, "decode_gbs": world * N / dec_max / 1e6,
        "value": world * N / (enc_max + dec_max) / 1e6, "unit": "GB/s",
        "encode_ms": enc_max, "decode_ms": dec_max, "ratio": world * N / (4.0 * wsum),
        "roofline": {"kernel": "cudppCompress pipeline (sort-bound BWT: prefix doubling, 6-pass one-sweep radix "
                               "sorts; see DESIGN.md 3.4)", "bound": "hbm", "achieved": alg / enc_ms / 1e6,
                     "peak": peak, "unit": "GB/s", "frac": alg / enc_ms / 1e6 / peak, "algorithmic_bytes": alg},
    }
    if with_cpu:
        import oracle_lib as O
        from bench_paths import cudpp_blocks_gpu
        cores = os.cpu_count() or 1
        sample = cudpp_blocks_gpu(cores, n, dev, "zipf", seed=95835).cpu().numpy().reshape(cores, n)

        def one(b):
            rc, idx, hist, offs, w = O.cudpp_oracle_compress(sample[b])
            return rc == 0
        from concurrent.futures import ThreadPoolExecutor
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            good = list(ex.map(one, range(cores)))
        dt = time.perf_counter() - t0
        assert all(good)
        res["cpu_baseline"] = {"value": cores * n / dt / 1e9, "unit": "GB/s", "cores": cores, "kind": "port",
                               "sample": "%d blocks of 1 MiB: oracle/cudpp_oracle.c BWT + MTF + Huffman (encode only)" % cores}
    return res


def c4_strong(pkg, dev, rank, world, total_blocks):
    """BASELINE.json configs[3]: `total_blocks` x 1 MiB blocks sharded over the ranks
    (shard.plan_blocks), per-block compressed sizes exchanged with one all_gather
    (shard.gather_offsets; the reference's per-block outputs: cudpp.cpp:733-748), offsets checked
    against an exclusive scan of the gathered sizes on rank 0.  Total work is fixed: strong scaling."""
    import numpy as np
    import torch
    import torch.distributed as dist
    shard = importlib.import_module("gpu-lossless-compression_b200.shard")
    lo, hi = shard.plan_blocks(total_blocks, world)[rank]
    torch.cuda.synchronize()
    if world > 1:
        # the first all_gather of a communicator sets up its channels (10 ms): not part of the step
        shard.gather_offsets(torch.ones(hi - lo, dtype=torch.int64, device=dev), total_blocks)
        dist.barrier()
    t = _events(2)
    t[0].record()
    enc_ms, dec_ms, sizes = _cudpp_run(pkg, dev, lo, hi - lo, CUDPP_BATCH, iters=1, check=False)
    g = _events(2)
    g[0].record()
    if world > 1:
        offsets, all_sizes = shard.gather_offsets(sizes * 4, total_blocks)
    else:
        all_sizes = sizes * 4
        offsets = torch.zeros(total_blocks + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(all_sizes, 0)
    g[1].record()
    t[1].record()
    t[1].synchronize()
    gather_ms = g[0].elapsed_time(g[1])
    wall_ms = t[0].elapsed_time(t[1])          # includes input generation between the batches
    ok = True
    if rank == 0:
        h = all_sizes.cpu().numpy()
        want = np.zeros(total_blocks + 1, np.int64)
        want[1:] = np.cumsum(h)
        ok = bool(np.array_equal(offsets.cpu().numpy(), want)) and bool((h > 0).all())
    # A rank that finishes early WAITS in the all_gather for the slowest one, so its gather_ms holds the
    # imbalance that enc_max + dec_max already contain: the step is the max over ranks of each rank's
    # own encode + decode + gather, and the collective itself is what the LAST rank to arrive sees
    # (the minimum over ranks).
    step_ms = enc_ms + dec_ms + gather_ms
    enc_max, dec_max, wall_max, gather_max, step_max, neg_gather_min = _reduce_max(
        [enc_ms, dec_ms, wall_ms, gather_ms, step_ms, -gather_ms], dev, world)
    N = total_blocks * MIB
    return {"workload": "cudppCompress encode+decode, %d x 1 MiB blocks sharded over %d GPU(s), contiguous "
                        "block ranges, one all_gather of %d block sizes" % (total_blocks, world, total_blocks),
            "blocks": total_blocks, "blocks_per_rank": hi - lo, "scaling": "strong",
            "value": N / step_max / 1e6, "unit": "GB/s",
            "encode_gbs": N / enc_max / 1e6, "decode_gbs": N / dec_max / 1e6,
            "step_ms": step_max, "encode_ms": enc_max, "decode_ms": dec_max,
            "gather_ms": -neg_gather_min, "gather_ms_incl_wait_for_slowest_rank": gather_max,
            "wall_ms_with_input_generation": wall_max,
            "compressed_bytes": int(offsets[-1].item()), "offsets_ok": ok,
            "collective": "all_gather(int64[%d]) over %s" % (-(-total_blocks // world), "nccl" if world > 1 else "none (1 rank)")}


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("gpu-lossless-compression_b200")
    L = pkg.lib()
    sys.path.insert(0, os.path.join(ROOT, "tools"))

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n = args.mib * MIB
    if args.input == "mt19937":
        data = gen_zipf_mt19937(n, dev, SEED + rank)
        data_desc = "synthetic: SURVEY.md 8(d) generator (inverse-CDF Zipf(1.1) over numpy MT19937(%d + rank))" % SEED
    else:
        data = gen_zipf_gpu(n, dev, SEED + rank)
        data_desc = "synthetic: inverse-CDF Zipf(1.1) over torch Philox(%d + rank), generated on the GPU" % SEED
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
    ev = _events(3)
    state = {}
    ref_table = args.table == "reference"
    if ref_table:
        rcode, rlen, rlut = reference_dictionary()
        h_code.copy_(torch.from_numpy(rcode.view(np.int32)))
        h_len.copy_(torch.from_numpy(rlen))
        h_lut.copy_(torch.from_numpy(rlut))
        dictionary = ("reference: llhuff get_symbol_lengths/canonical codes for the C2 input "
                      "(tests/golden/cuhd_c2_table.npz), an input of the step")
    else:
        dictionary = "own: b200lc_cuhd_build_table on the step's histogram, built inside the step"

    def step(timed):
        """per-piece histograms -> table -> one-pass pack -> decode, all on `stream`;
        returns (enc_ms, dec_ms) if timed."""
        if timed:
            ev[0].record(stream)
        pkg.check(L.b200lc_histogram_u8_pieces(data.data_ptr(), n, hist.data_ptr(), piece_hist.data_ptr(), sp),
                  "hist")
        h_hist.copy_(hist, non_blocking=True)
        stream.synchronize()                      # the code table is built (or the stream sized) on the host
        if not ref_table:
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
    t_all0, t_all1 = _events(2)
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
    # The copy engines serve each direction in issue order and a step is a chain (symbols up ->
    # stream down -> stream up -> symbols down), so ONE encode and ONE decode call in flight leave
    # both directions idle half of the time (measured: encode 34.7 ms + decode 26.0 ms alone, 58.1 ms
    # when started together, tools/pcie_probe.py).  Two encode sessions and two decode sessions, each
    # on its own host thread, free-running over the steps, keep both directions busy.  Every step's
    # input goes up and every step's result comes down inside the timed region, pipeline fill and
    # drain included.
    e2e_steps = max(4, args.e2e_steps)             # its own step count: fill and drain amortise over it
    NE = int(os.environ.get("B200LC_E2E_ENC", "2"))     # encode sessions
    ND = int(os.environ.get("B200LC_E2E_DEC", "2"))     # decode sessions
    NB = NE + ND                                   # stream buffers in flight
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.copy_(data)
    h_units = [torch.empty(units_cap, dtype=torch.int32).pin_memory() for _ in range(NB)]
    h_luts = [torch.empty((1 << MAX_LEN, 2), dtype=torch.uint8).pin_memory() for _ in range(NB)]
    h_codes = [torch.empty(256, dtype=torch.int32).pin_memory() for _ in range(NE)]
    h_lens = [torch.empty(256, dtype=torch.uint8).pin_memory() for _ in range(NE)]
    h_outs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(ND)]
    s_enc = [pkg.CuhdSession(n) for _ in range(NE)]
    s_dec = [pkg.CuhdSession(n) for _ in range(ND)]
    for i in range(max(NE, ND)):                   # warm-up of all sessions
        e, d = i % NE, i % ND
        nu0 = s_enc[e].encode(h_in, h_units[e], h_codes[e], h_lens[e], h_luts[e], MAX_LEN)
        s_dec[d].decode(h_units[e], nu0 + 1, h_luts[e], h_outs[d], MAX_LEN)
        assert torch.equal(h_outs[d], h_in), "e2e round trip mismatch"
        h_outs[d].zero_()
    nus = [0] * e2e_steps
    enc_done = [threading.Event() for _ in range(e2e_steps)]
    dec_done = [threading.Event() for _ in range(e2e_steps)]
    errors = []

    def enc_worker(e):
        try:
            torch.cuda.set_device(dev)
            for k in range(e, e2e_steps, NE):
                if k >= NB:
                    dec_done[k - NB].wait()        # its stream buffer is free again
                nus[k] = s_enc[e].encode(h_in, h_units[k % NB], h_codes[e], h_lens[e], h_luts[k % NB], MAX_LEN)
                enc_done[k].set()
        except Exception as ex:                    # surfaced after the join
            errors.append(ex)
            for ev_ in enc_done:
                ev_.set()

    def dec_worker(d):
        try:
            torch.cuda.set_device(dev)
            for k in range(d, e2e_steps, ND):
                enc_done[k].wait()
                if errors:
                    break
                s_dec[d].decode(h_units[k % NB], nus[k] + 1, h_luts[k % NB], h_outs[d], MAX_LEN)
                dec_done[k].set()
        except Exception as ex:
            errors.append(ex)
        finally:
            for ev_ in dec_done:
                ev_.set()

    barrier()
    workers = [threading.Thread(target=enc_worker, args=(e,)) for e in range(NE)] + \
              [threading.Thread(target=dec_worker, args=(d,)) for d in range(ND)]
    t0 = time.perf_counter()
    for t_ in workers:
        t_.start()
    for t_ in workers:
        t_.join()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if errors:
        raise errors[0]
    assert all(torch.equal(h_, h_in) for h_ in h_outs[: min(ND, e2e_steps)]), "e2e round trip mismatch after the timed region"
    nu = nus[-1]
    # serial figure: one session, encode then decode, step after step (round 1's e2e).  With many
    # ranks on one host the copies of ONE step per rank already saturate the host's memory path
    # and four sessions per rank only add contention (N = 8: 58.9 GB/s serial, 37.7 pipelined), so
    # both are measured and `e2e.value` is the better of the two, named in `e2e.mode`.
    serial_steps = max(2, e2e_steps // 4)
    barrier()
    t0 = time.perf_counter()
    for _ in range(serial_steps):
        nu1 = s_enc[0].encode(h_in, h_units[0], h_codes[0], h_lens[0], h_luts[0], MAX_LEN)
        s_enc[0].decode(h_units[0], nu1 + 1, h_luts[0], h_outs[0], MAX_LEN)
    torch.cuda.synchronize()
    e2e_serial_s = (time.perf_counter() - t0) / serial_steps
    assert torch.equal(h_outs[0], h_in), "e2e round trip mismatch (serial)"
    for s_ in s_enc + s_dec:
        s_.close()
    del h_units, h_outs, h_luts, h_codes, h_lens
    del h_in
    if hasattr(torch._C, "_host_emptyCache"):      # give the pinned staging buffers back to the host
        torch._C._host_emptyCache()
    h2d = n + (nu + 1) * 4 + 256 * 5 + (2 << MAX_LEN)
    d2h = n + (nu + 1) * 4 + 256 * 8 + 8

    # ---------------------------------------------------------------- reduce over ranks
    total_ms, e2e_s, e2e_serial_s, dec_mean, enc_mean = _reduce_max(
        [total_ms, e2e_s, e2e_serial_s, statistics.mean(dec_ms), statistics.mean(enc_ms)], dev, world)

    peak, peak_src = measured_peak()
    # free the C2 buffers before the other paths
    n_units = state["n_units"]
    del units, out, enc_scratch, dec_scratch, piece_hist, data
    torch.cuda.empty_cache()

    paths = None
    if not args.no_paths:
        paths = {"culzss": path_culzss(pkg, dev, rank, world, args.culzss_mib, peak, world == 1 and not args.no_cpu),
                 "cudpp": path_cudpp(pkg, dev, rank, world, args.cudpp_blocks, peak, world == 1 and not args.no_cpu)}
        torch.cuda.empty_cache()
    strong = None
    if not args.no_strong:
        strong = c4_strong(pkg, dev, rank, world, args.strong_blocks)
    if rank != 0:
        return

    traffic, traffic_src = profiled_traffic(args.mib)
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
        "vs_baseline": None, "dtype": "u8", "data": data_desc,
        "config": {"workload": "CUHD Huffman encode+decode, %d MiB Zipf(1.1) bytes per GPU, "
                               "256-symbol length-limited (11 bit) canonical code" % args.mib,
                   "bytes_per_step_per_gpu": n, "compressed_bytes": 4 * n_units,
                   "ratio": n / (4.0 * n_units), "parallelism": "independent buffer per GPU",
                   "dictionary": dictionary,
                   "l2": "inputs (%d MiB) exceed the 126 MB L2; no explicit flush" % args.mib},
        "encode_gbs": world * n / (enc_mean * 1e-3) / 1e9,
        "decode_gbs": world * n / (dec_mean * 1e-3) / 1e9,
        "roofline": {"kernel": "cuhd_decode_kernel", "bound": "hbm", "achieved": dec_gbs,
                     "peak": peak, "unit": "GB/s", "frac": dec_gbs / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes": dec_bytes, "launch_ms": dec_mean},
        "e2e": {"value": world * n / min(e2e_s, e2e_serial_s) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "mode": "pipelined" if e2e_s <= e2e_serial_s else "serial",
                "pipelined_value": world * n / e2e_s / 1e9,
                "serial_value": world * n / e2e_serial_s / 1e9,
                "sessions": [NE, ND],
                "api": "b200lc_cuhd_session_encode + b200lc_cuhd_session_decode, pinned host buffers; %d "
                       "encode and %d decode sessions, one host thread each, free-running over `steps` steps "
                       "(fill and drain inside the timed region); serial_value = one session, encode then "
                       "decode" % (NE, ND)},
        "gpu_launches": 6 * args.steps,   # piece histograms, their reduction, piece bits, plan, pack, decode
        "clocks": clocks,
    }
    if HOST_PLACEMENT is not None:
        line["host_placement"] = HOST_PLACEMENT      # rank 0's; every rank pins to its own GPU's node
    if cpu:
        line["cpu_baseline"] = cpu
    if paths:
        line["paths"] = paths
    if strong:
        line["c4_strong"] = strong
    print(json.dumps(line))


HOST_PLACEMENT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mib", type=int, default=1024, help="uncompressed MiB per GPU per step")
    ap.add_argument("--cpu-mib", type=int, default=128, help="CPU baseline sample size")
    ap.add_argument("--ref-mib", type=int, default=256, help="--impl reference sample per step")
    ap.add_argument("--e2e-steps", type=int, default=48,
                    help="steps of the end-to-end (host buffer) pipeline; independent of --steps")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--table", default="own", choices=["own", "reference"],
                    help="C2 dictionary: built by b200lc_cuhd_build_table inside the step, or the reference encoder's (fixture)")
    ap.add_argument("--input", default="mt19937", choices=["mt19937", "philox"],
                    help="C2 input generator: SURVEY 8(d)'s (CPU, ~40 s per GiB) or torch Philox on the GPU")
    ap.add_argument("--no-paths", action="store_true", help="skip the CULZSS / cudppCompress legs")
    ap.add_argument("--no-strong", action="store_true", help="skip the sharded 8192-block leg")
    ap.add_argument("--culzss-mib", type=int, default=4096)
    ap.add_argument("--cudpp-blocks", type=int, default=1024, help="1 MiB blocks per GPU")
    ap.add_argument("--strong-blocks", type=int, default=8192, help="1 MiB blocks in total (configs[3])")
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
    # host placement: CPUs and memory of the GPU's NUMA node, before any pinned buffer exists
    global HOST_PLACEMENT
    if os.environ.get("B200LC_NO_NUMA_PIN", "0") != "1":
        import importlib
        HOST_PLACEMENT = importlib.import_module("gpu-lossless-compression_b200.hostpin").pin_to_gpu(local_rank)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
