"""Block sharding across GPUs of one node.

Every hot path works on independent units (CUHD streams, 1 MiB LZSS buffers, 1 MiB BWT blocks),
so ranks never exchange payload on the data path: rank r owns the contiguous unit range
plan_blocks(n_units, world)[r].  The only communication is one all_gather of the per-unit
compressed sizes, after which every rank computes the same exclusive scan = the offset of
every unit in the concatenated output (the "broadcast/gather of block offsets" of the north
star).  Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def plan_blocks(n_units, world):
    """Contiguous, balanced unit ranges: [(lo, hi)] per rank; the first n_units % world ranks get
    one extra unit."""
    base, extra = divmod(n_units, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def gather_offsets(local_sizes, n_units, group=None):
    """local_sizes: int64 tensor with the compressed sizes of this rank's units (on the device
    the backend needs).  Returns (offsets[n_units + 1], all_sizes[n_units]) identical on every
    rank; offsets[i] = start of unit i in the concatenated stream."""
    world = dist.get_world_size(group)
    ranges = plan_blocks(n_units, world)
    longest = max(hi - lo for lo, hi in ranges)
    padded = torch.zeros(longest, dtype=torch.int64, device=local_sizes.device)
    padded[: local_sizes.numel()] = local_sizes
    parts = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    sizes = torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(ranges)])
    offsets = torch.zeros(n_units + 1, dtype=torch.int64, device=local_sizes.device)
    offsets[1:] = torch.cumsum(sizes, 0)
    return offsets, sizes
