"""CPU, world size 2 over gloo: the multi-GPU host logic (unit ranges + offsets exchange)."""
import importlib
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_units, q):
    sys.path.insert(0, ROOT)
    shard = importlib.import_module("gpu-lossless-compression_b200.shard")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.plan_blocks(n_units, world)[rank]
        local = torch.tensor([1000 + 7 * i for i in range(lo, hi)], dtype=torch.int64)
        offsets, sizes = shard.gather_offsets(local, n_units)
        q.put((rank, offsets.tolist(), sizes.tolist()))
    finally:
        dist.destroy_process_group()


def _run(n_units, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res)


def test_plan_blocks_is_a_partition():
    shard = importlib.import_module("gpu-lossless-compression_b200.shard")
    for n in (0, 1, 7, 8, 8192, 8193):
        for w in (1, 2, 3, 8):
            r = shard.plan_blocks(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_offsets_exchange_world2_even_and_ragged():
    for n_units, port in ((8, 29611), (7, 29612), (1, 29613)):
        (r0, off0, sz0), (r1, off1, sz1) = _run(n_units, port)
        want = [1000 + 7 * i for i in range(n_units)]
        assert sz0 == want and sz1 == want
        assert off0 == off1 and off0[0] == 0 and off0[-1] == sum(want)
        assert all(off0[i + 1] - off0[i] == want[i] for i in range(n_units))
