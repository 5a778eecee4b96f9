#!/usr/bin/env python
"""PCIe probe for the e2e figure: pinned H2D / D2H alone and together (two streams), and the CUHD
session encode / decode alone and concurrently on two host threads.  One JSON line.
Under torchrun every rank probes its own GPU at the same time (what the 8-rank e2e sees)."""
import importlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("gpu-lossless-compression_b200")


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    placement = None
    if "--pin" in sys.argv:      # CPUs / memory of the GPU's NUMA node before the pinned buffers exist
        placement = importlib.import_module("gpu-lossless-compression_b200.hostpin").pin_to_gpu(local)
    n = 1 << 30
    h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            sync_all()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    def up():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    def both():
        up()
        down()

    res = {"rank": rank, "world": world}
    res["h2d_gbs"] = n / timed(up) / 1e9
    res["d2h_gbs"] = n / timed(down) / 1e9
    res["duplex_each_gbs"] = n / timed(both) / 1e9

    # the sessions
    import bench as B
    data = B.gen_zipf_gpu(n, dev, 12345 + rank)
    h_a.copy_(data)
    del data, d_a, d_b
    cap = (n * 11 + 31) // 32 + 2
    h_units = [torch.empty(cap, dtype=torch.int32).pin_memory() for _ in range(2)]
    h_code = torch.empty(256, dtype=torch.int32).pin_memory()
    h_len = torch.empty(256, dtype=torch.uint8).pin_memory()
    h_lut = torch.empty((2048, 2), dtype=torch.uint8).pin_memory()
    se, sd = pkg.CuhdSession(n), pkg.CuhdSession(n)
    nu = [0]

    def enc(k=0):
        nu[0] = se.encode(h_a, h_units[k], h_code, h_len, h_lut, 11)

    def dec(k=0):
        sd.decode(h_units[k], nu[0] + 1, h_lut, h_b, 11)

    enc(0); dec(0); enc(1)
    res["encode_ms"] = 1e3 * timed(lambda: enc(0))
    res["decode_ms"] = 1e3 * timed(lambda: dec(0))

    def together():
        a = threading.Thread(target=enc, args=(1,))
        b = threading.Thread(target=dec, args=(0,))
        a.start(); b.start(); a.join(); b.join()

    res["encode_and_decode_concurrent_ms"] = 1e3 * timed(together)
    assert torch.equal(h_b, h_a)
    res["host_placement"] = placement
    print(json.dumps(res), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
