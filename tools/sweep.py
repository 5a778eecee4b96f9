#!/usr/bin/env python
"""BASELINE.json configs[4]: block-size x entropy sweep, encode + decode GB/s and ratio for the
three paths on one GPU (device-resident, CUDA events, best of 3).  One JSON line per point.

  python tools/sweep.py [--mib 256] [--paths cuhd,cuhd_batch,culzss,culzss_lane,cudpp] [--md out.md]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/sweep.py ...      # N GPUs: every rank sweeps its own data, rank 0 prints whole-job GB/s

Input: order-0 bytes with Zipf exponent solved for the target entropy H (H = 8: uniform).
Block size means: CULZSS buffer length (the reference fixes 1 MiB, main.c:62 -- other sizes are
extensions of the same format), cudppCompress block length (the reference validates 1 MiB only;
nblocks * n < 2^30 here), CUHD: length of one independent stream (own table); the streams are
issued round-robin on 8 CUDA streams through the explicit-stream C ABI.
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("gpu-lossless-compression_b200")
from bench_paths import PEAK, timeit  # noqa: E402

KIB, MIB = 1 << 10, 1 << 20


def zipf_probs_for_entropy(H, nsym=256):
    """Zipf(s) over nsym symbols with Shannon entropy H bits (bisection on s)."""
    def ent(s):
        p = 1.0 / np.arange(1, nsym + 1, dtype=np.float64) ** s
        p /= p.sum()
        return float(-(p * np.log2(p)).sum()), p
    if H >= np.log2(nsym) - 1e-9:
        return ent(0.0)[1]
    lo, hi = 0.0, 20.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if ent(mid)[0] > H:
            lo = mid
        else:
            hi = mid
    return ent(0.5 * (lo + hi))[1]


def gen(n, H, dev, seed, lo_sym=0):
    """n bytes with order-0 entropy H; symbols lo_sym .. (lo_sym = 1 keeps byte 0 out)."""
    nsym = 256 - lo_sym
    p = zipf_probs_for_entropy(min(H, np.log2(nsym)), nsym)
    cdf = torch.from_numpy(np.cumsum(p)).to(device=dev, dtype=torch.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    for lo in range(0, n, 1 << 26):
        m = min(1 << 26, n - lo)
        u = torch.rand(m, generator=g, device=dev)
        out[lo:lo + m] = (torch.searchsorted(cdf, u).clamp_(max=nsym - 1) + lo_sym).to(torch.uint8)
    return out


NSTREAMS = 8


def point_cuhd(data, block, dev):
    """Independent streams of `block` symbols, each with its own table.  The C ABI takes an explicit
    CUDA stream, so the per-stream calls are issued round-robin on NSTREAMS CUDA streams (each with
    its own scratch / output buffers) and overlap on the device."""
    n = data.numel()
    nstreams = n // block
    L = pkg.lib()
    encs, tabs = [], []
    for s in range(nstreams):
        part = data[s * block:(s + 1) * block]
        hist = pkg.histogram_u8(part).cpu().numpy()
        code, length, lut = pkg.cuhd_build_table(hist)
        tabs.append((torch.from_numpy(code.view(np.int32)).to(dev), torch.from_numpy(length).to(dev),
                     torch.from_numpy(lut).to(dev)))
        encs.append(pkg.cuhd_encode(part, tabs[-1][0], tabs[-1][1]))
    comp = sum(e.n_units * 4 for e in encs)
    ucap = (block * 11 + 31) // 32 + 2
    lanes = min(NSTREAMS, nstreams)
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
    units = [torch.empty(ucap, dtype=torch.int32, device=dev) for _ in range(lanes)]
    bits = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(lanes)]
    escr = [torch.empty(L.b200lc_cuhd_encode_scratch_bytes(block), dtype=torch.uint8, device=dev) for _ in range(lanes)]
    outs = [torch.empty(block, dtype=torch.uint8, device=dev) for _ in range(lanes)]
    dscr = [torch.empty(L.b200lc_cuhd_decode_scratch_bytes(ucap), dtype=torch.uint8, device=dev) for _ in range(lanes)]

    def fan_out(body):
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        for st in streams:
            st.wait_event(start)
        for s in range(nstreams):
            k = s % lanes
            with torch.cuda.stream(streams[k]):
                body(s, k, streams[k])
        for st in streams:
            main.wait_stream(st)

    def enc_all():
        fan_out(lambda s, k, st: (pkg.histogram_u8(data[s * block:(s + 1) * block], stream=st),
                                  pkg.cuhd_encode(data[s * block:(s + 1) * block], tabs[s][0], tabs[s][1],
                                                  units=units[k], total_bits=bits[k], scratch=escr[k],
                                                  sync=False, stream=st)))

    def dec_all():
        fan_out(lambda s, k, st: pkg.cuhd_decode(encs[s].units, block, tabs[s][2], 11, out=outs[k],
                                                 scratch=dscr[k], stream=st))

    enc_ms = timeit(enc_all, iters=3, warm=1)
    dec_ms = timeit(dec_all, iters=3, warm=1)
    torch.cuda.synchronize()
    last = nstreams - 1
    ok = bool(torch.equal(outs[last % lanes], data[last * block:]))
    return enc_ms, dec_ms, comp, ok


def point_cuhd_batch(data, block, dev):
    """Blocks of `block` symbols packed as independent streams with ONE shared code table by one
    launch of b200lc_cuhd_encode_blocks and decoded by one launch of b200lc_cuhd_decode_batch
    (look-backs stop at block boundaries)."""
    n = data.numel()
    nstreams = n // block
    L = pkg.lib()
    hist = np.maximum(pkg.histogram_u8(data).cpu().numpy(), 1)
    code, length, lut = pkg.cuhd_build_table(hist)
    d_code, d_len = torch.from_numpy(code.view(np.int32)).to(dev), torch.from_numpy(length).to(dev)
    d_lut = torch.from_numpy(lut).to(dev)
    units, bits, stride = pkg.cuhd_encode_blocks(data, block, d_code, d_len)
    escr = torch.empty(L.b200lc_cuhd_encode_blocks_scratch_bytes(n, block), dtype=torch.uint8, device=dev)

    def enc_all():
        pkg.histogram_u8(data)
        pkg.cuhd_encode_blocks(data, block, d_code, d_len, unit_stride=stride, units=units, block_bits=bits,
                               scratch=escr)

    enc_ms = timeit(enc_all, iters=3, warm=1)
    nu = (bits.cpu().numpy().astype(np.int64) + 31) // 32
    streams = np.zeros((nstreams, 4), np.uint64)
    streams[:, 0] = np.arange(nstreams, dtype=np.uint64) * np.uint64(stride)
    streams[:, 1] = nu.astype(np.uint64)
    streams[:, 2] = np.arange(nstreams, dtype=np.uint64) * np.uint64(block)
    streams[:, 3] = block
    comp = int(nu.sum()) * 4
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    dscr = torch.empty(L.b200lc_cuhd_decode_batch_scratch_bytes(streams.ctypes.data, nstreams) + 256,
                       dtype=torch.uint8, device=dev)
    dec_ms = timeit(lambda: pkg.cuhd_decode_batch(units, out, streams, d_lut, scratch=dscr), iters=3, warm=1)
    torch.cuda.synchronize()
    return enc_ms, dec_ms, comp, bool(torch.equal(out, data))


def point_culzss_lane(data, block, dev):
    """NON-PARITY fast mode (packet-per-lane formulation), same format and decoder."""
    return point_culzss(data, block, dev, fast="lane")


def point_culzss(data, block, dev, fast=0):
    n = data.numel()
    nbuf = n // block
    L = pkg.lib()
    stride = pkg.culzss_out_stride(block)
    out = torch.empty(nbuf * stride, dtype=torch.uint8, device=dev)
    clen = torch.empty(nbuf, dtype=torch.int32, device=dev)
    scr = torch.empty(L.b200lc_culzss_encode_scratch_bytes(nbuf, block), dtype=torch.uint8, device=dev)
    enc_ms = timeit(lambda: pkg.culzss_encode(data, block, out, clen, scr, fast=fast), iters=3, warm=1)
    cl = clen.cpu().numpy().astype(np.int64)
    sizes = np.where(cl == 0, block, cl)
    offs = np.zeros(nbuf + 1, np.int64)
    offs[1:] = np.cumsum(sizes)
    comp = torch.empty(int(offs[-1]), dtype=torch.uint8, device=dev)
    rows = out.view(nbuf, stride)
    d_sz = torch.from_numpy(sizes).to(dev)
    raw = torch.from_numpy(cl == 0).to(dev)
    col = torch.arange(stride, device=dev)
    step = max(1, (64 * MIB) // stride)
    for lo in range(0, nbuf, step):
        hi = min(nbuf, lo + step)
        src = rows[lo:hi].clone()
        if bool(raw[lo:hi].any()):
            src[:, :block][raw[lo:hi]] = data.view(nbuf, block)[lo:hi][raw[lo:hi]]
        comp[int(offs[lo]):int(offs[hi])] = src[col[None, :] < d_sz[lo:hi, None]]
    d_offs = torch.from_numpy(offs).to(dev)
    dec = torch.empty(n, dtype=torch.uint8, device=dev)
    dscr = torch.empty(L.b200lc_culzss_decode_scratch_bytes(nbuf, block), dtype=torch.uint8, device=dev)
    dec_ms = timeit(lambda: pkg.culzss_decode(comp, d_offs, block, dec, dscr), iters=3, warm=1)
    return enc_ms, dec_ms, int(offs[-1]), bool(torch.equal(dec, data))


def point_cudpp(data, block, dev):
    n = data.numel()
    nblocks = n // block
    data = data.clone()
    data.view(nblocks, block)[:, -1] = 0       # the reference's validated domain: 1..255 + final 0
    L = pkg.lib()
    scr = torch.empty(max(L.b200lc_cudpp_compress_scratch_bytes(nblocks, block),
                          L.b200lc_cudpp_decompress_scratch_bytes(nblocks, block)) + 256,
                      dtype=torch.uint8, device=dev)
    res = [None]

    def enc():
        res[0] = pkg.cudpp_compress_batch(data, nblocks, block, scratch=scr, out=res[0])

    enc_ms = timeit(enc, iters=3, warm=1)
    if int(res[0].error.item()) != 0:
        return enc_ms, None, None, False
    comp = int(res[0].total_words.sum().item()) * 4
    back = torch.empty(n, dtype=torch.uint8, device=dev)
    err = [None]

    def dec():
        _, err[0] = pkg.cudpp_decompress_batch(res[0], nblocks, block, scratch=scr, out=back)

    dec_ms = timeit(dec, iters=3, warm=1)
    return enc_ms, dec_ms, comp, bool(torch.equal(back, data)) and int(err[0].item()) == 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--paths", default="cuhd,cuhd_batch,culzss,cudpp")
    ap.add_argument("--blocks", default="65536,262144,1048576,4194304")
    ap.add_argument("--entropies", default="1,2,3,4,5,6,7,8")
    ap.add_argument("--md", default=None)
    args = ap.parse_args()
    # under torchrun: every rank sweeps its own data (weak scaling, blocks are independent), the times
    # are reduced with MAX and the compressed sizes with SUM over NCCL, rank 0 reports the aggregate
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda:%d" % int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    rows = []
    for path in args.paths.split(","):
        fn = {"cuhd": point_cuhd, "cuhd_batch": point_cuhd_batch, "culzss": point_culzss,
              "culzss_lane": point_culzss_lane, "cudpp": point_cudpp}[path]
        total = (args.mib if path != "cudpp" else min(args.mib, 128)) * MIB
        for H in [float(x) for x in args.entropies.split(",")]:
            data = gen(total, H, dev, seed=1000 + int(H * 10) + 7919 * rank, lo_sym=1 if path == "cudpp" else 0)
            for block in [int(x) for x in args.blocks.split(",")]:
                rec = {"path": path, "H": H, "block": block, "mib": total // MIB, "n_gpus": world}
                enc_ms, dec_ms, comp, ok = fn(data, block, dev)
                if world > 1:
                    t = torch.tensor([enc_ms, dec_ms or 0.0, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    c = torch.tensor([float(comp or 0)], dtype=torch.float64, device=dev)
                    dist.all_reduce(c, op=dist.ReduceOp.SUM)
                    enc_ms, dec_ms, ok = float(t[0]), (float(t[1]) or None), float(t[2]) == 0.0
                    comp = int(c[0]) or None
                job = total * world        # bytes of the whole job; per-GPU fractions below
                rec.update({"encode_gbs": job / enc_ms / 1e6,
                            "decode_gbs": job / dec_ms / 1e6 if dec_ms else None,
                            "ratio": job / comp if comp else None, "round_trip": ok,
                            "encode_hbm_frac": (job + comp) / world / enc_ms / 1e6 / PEAK if comp else None,
                            "decode_hbm_frac": (job + comp) / world / dec_ms / 1e6 / PEAK if comp and dec_ms else None})
                rows.append(rec)
                if rank == 0:
                    print(json.dumps(rec), flush=True)
            del data
            torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if args.md and rank == 0:
        with open(args.md, "w") as f:
            f.write("| path | H (bits/byte) | block | encode GB/s | decode GB/s | ratio | %HBM enc / dec | round trip |\n")
            f.write("|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                if "skipped" in r:
                    f.write("| %s | %g | %d KiB | - | - | - | - | skipped: %s |\n" % (r["path"], r["H"], r["block"] // KIB, r["skipped"]))
                    continue
                fmt = lambda v, s="%.1f": "-" if v is None else s % v  # noqa: E731
                f.write("| %s | %g | %d KiB | %s | %s | %s | %s / %s | %s |\n" % (
                    r["path"], r["H"], r["block"] // KIB, fmt(r["encode_gbs"]), fmt(r["decode_gbs"]),
                    fmt(r["ratio"], "%.3f"), fmt(r["encode_hbm_frac"] and 100 * r["encode_hbm_frac"], "%.2f"),
                    fmt(r["decode_hbm_frac"] and 100 * r["decode_hbm_frac"], "%.2f"), "ok" if r["round_trip"] else "FAIL"))


if __name__ == "__main__":
    main()
