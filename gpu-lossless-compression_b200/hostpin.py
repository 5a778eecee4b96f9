"""Host-side placement for the end-to-end paths: run the calling process on the CPUs of the NUMA node
its GPU hangs off and prefer that node's memory, so that pinned staging buffers (cudaHostAlloc /
torch pin_memory, allocated after this call) are local to the GPU's PCIe root.  Best effort: every
step that the container does not allow (no sysfs NUMA information, CPUs of the node not in the
allowed set, set_mempolicy refused) is skipped and reported in the returned dict.

The reference has no multi-GPU host layer; this belongs to the N-GPU end-to-end
measurement of bench.py."""
import ctypes
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of CUDA device `device_index` from sysfs, or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index)
        pci = "%04x:%02x:%02x.0" % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
    except Exception:
        return None
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % pci) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def pin_to_gpu(device_index):
    """-> dict describing what was done (goes into the bench line)."""
    info = {"node": None, "cpus_pinned": 0, "mempolicy": False}
    node = gpu_numa_node(device_index)
    if node is None:
        info["note"] = "no NUMA node reported for the GPU"
        return info
    info["node"] = node
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        target = cpus & os.sched_getaffinity(0)
        if target:
            os.sched_setaffinity(0, target)
            info["cpus_pinned"] = len(target)
        else:
            info["note"] = "none of the node's CPUs is in the allowed set"
    except Exception as e:      # noqa: BLE001
        info["note"] = "affinity: %s" % e
    try:
        # set_mempolicy(MPOL_PREFERRED = 1, nodemask, maxnode): x86_64 syscall 238
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        if libc.syscall(238, 1, mask, 16 * 64 + 1) == 0:
            info["mempolicy"] = True
    except Exception:
        pass
    return info
