"""Host-side placement for the pinned transfer buffers.

`cudaMallocHost` pages land on the NUMA node of the allocating thread.  On an 8-GPU box the
GPUs hang off two sockets: a rank whose pinned buffers live on the far socket pays the
inter-socket link on every H2D/D2H copy, and eight ranks allocating from one node saturate that
node's memory controllers (round 1: 8 ranks reached 113 GB/s aggregate host traffic vs 56 GB/s for
one).  `bind_to_gpu_numa_node` pins the calling process to the CPUs that are local to its GPU
BEFORE the pinned buffers are allocated, so that first-touch places them on the GPU's own node.
Pure sysfs + sched_setaffinity: no dependency, and a no-op (with the reason reported) when the
topology is not exposed.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]"""
    cpus: List[int] = []
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def pci_sysfs_dir(domain: int, bus: int, device: int, root: str = "/sys/bus/pci/devices") -> str:
    return os.path.join(root, "%04x:%02x:%02x.0" % (domain, bus, device))


def gpu_locality(domain: int, bus: int, device: int, root: str = "/sys/bus/pci/devices") -> Dict[str, object]:
    """{'numa_node': int or None, 'cpus': [..]} of a PCI function, from sysfs."""
    d = pci_sysfs_dir(domain, bus, device, root)
    out: Dict[str, object] = {"numa_node": None, "cpus": [], "sysfs": d}
    try:
        with open(os.path.join(d, "numa_node")) as f:
            node = int(f.read().strip())
            out["numa_node"] = node if node >= 0 else None
        with open(os.path.join(d, "local_cpulist")) as f:
            out["cpus"] = parse_cpulist(f.read())
    except (OSError, ValueError):
        pass
    return out


def bind_to_gpu_numa_node(device_index: int, root: str = "/sys/bus/pci/devices") -> Dict[str, object]:
    """Restricts this process to the CPUs local to CUDA device `device_index`; returns what was done."""
    info: Dict[str, object] = {"bound": False, "numa_node": None, "cpus": 0, "why": ""}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        dom, bus, dev = int(getattr(p, "pci_domain_id", 0)), int(p.pci_bus_id), int(p.pci_device_id)
    except Exception as e:                                   # no CUDA / old torch: nothing to bind to
        info["why"] = "no PCI address: %s" % str(e)[:80]
        return info
    loc = gpu_locality(dom, bus, dev, root)
    info["numa_node"] = loc["numa_node"]
    allowed = set(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else set()
    cpus = [c for c in loc["cpus"] if c in allowed]
    if not cpus:
        info["why"] = "no local_cpulist for %s (or none of it allowed)" % loc["sysfs"]
        return info
    if len(cpus) == len(allowed):
        info["why"] = "single NUMA node: every allowed CPU is local"
        info["cpus"] = len(cpus)
        return info
    try:
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus=len(cpus))
    except OSError as e:
        info["why"] = "sched_setaffinity failed: %s" % e
    return info
