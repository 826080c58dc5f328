"""Pin a rank's host threads (and, through first-touch, its pinned staging buffers) to the NUMA node its GPU hangs off.

Part of the frame loader: with one process per GPU on an 8-GPU box, host buffers allocated on the far
socket make every H2D/D2H of the end-to-end path cross the inter-socket link.  Pure host plumbing: sysfs + sched_setaffinity,
no dependency beyond the standard library (NVML / torch are only used to find the GPU's PCI address).
"""
from __future__ import annotations

import os
from typing import Optional


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def pci_address(cuda_index: int) -> Optional[str]:
    """'0000:1b:00.0'-style sysfs name of CUDA device `cuda_index` of this process, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(cuda_index)
        if hasattr(p, "pci_bus_id"):
            return "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, getattr(p, "pci_device_id", 0))
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[cuda_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else cuda_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        return "%04x:%s" % (int(dom, 16), rest.lower())
    except Exception:
        return None


def numa_node_of(pci: Optional[str], sysfs: str = "/sys") -> Optional[int]:
    if not pci:
        return None
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", pci, "numa_node")).read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def cpus_of_node(node: int, sysfs: str = "/sys"):
    try:
        return _parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
    except Exception:
        return set()


def bind_to_gpu(cuda_index: int, sysfs: str = "/sys") -> dict:
    """Restricts this process to the CPUs of the GPU's NUMA node (intersected with the CPUs it may already use).
    Returns what was done, for the bench report; never raises (no NUMA information = no binding)."""
    info = {"pci": pci_address(cuda_index), "node": None, "cpus": None, "bound": False}
    node = numa_node_of(info["pci"], sysfs)
    info["node"] = node
    if node is None or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        allowed = os.sched_getaffinity(0)
        cpus = cpus_of_node(node, sysfs) & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
        info["cpus"] = len(cpus) if cpus else len(allowed)
    except Exception:
        pass
    return info
