"""CPU: the host placement helper of the end-to-end bench (gpu-lossless-compression_b200/hostpin.py)
degrades to a report when there is nothing to pin (no GPU / one NUMA node)."""
import importlib
import os

hostpin = importlib.import_module("gpu-lossless-compression_b200.hostpin")


def test_cpulist_parser():
    assert hostpin._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostpin._parse_cpulist("") == set()
    assert hostpin._parse_cpulist("5") == {5}


def test_pin_without_numa_information_changes_nothing():
    before = os.sched_getaffinity(0)
    info = hostpin.pin_to_gpu(0)
    assert set(info) >= {"node", "cpus_pinned", "mempolicy"}
    if info["node"] is None:
        assert info["cpus_pinned"] == 0 and os.sched_getaffinity(0) == before
