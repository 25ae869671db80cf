"""Host-side placement for the end-to-end (host buffers in, host results out) path.

A B200 box has two CPU sockets; a pinned buffer that lives on the socket the GPU is NOT attached to is copied
across the socket interconnect, and eight ranks pulling ~50 GB/s each from one socket's memory saturate it
(round 1 measured 51 -> 23 GB/s per GPU going from 1 to 8 ranks).  ``bind_to_gpu_numa_node`` pins the calling
process to the CPUs of the GPU's NUMA node BEFORE the pinned buffers are allocated and first touched, so that
every rank stages through its own socket's memory.  Pure host logic: no CUDA call besides the PCI bus id query.
"""
from __future__ import annotations

import os
from typing import List, Optional


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of /sys/devices/system/node/node*/cpulist)."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(pci_bus_id: str, sysfs: str = "/sys") -> Optional[int]:
    """NUMA node of the PCI device ('0000:1b:00.0'), or None if the platform does not say (-1 / no file)."""
    bus = pci_bus_id.lower()
    if len(bus.split(":")[0]) == 8:                      # CUDA prints an 8-digit domain, sysfs uses 4
        bus = bus[4:]
    try:
        with open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")) as fh:
            node = int(fh.read().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node: int, sysfs: str = "/sys") -> List[int]:
    try:
        with open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")) as fh:
            return parse_cpulist(fh.read())
    except OSError:
        return []


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Restrict this process to the CPUs of the NUMA node ``cuda:device_index`` hangs off.  Returns what was done
    ({'node': n, 'cpus': count} or {'node': None, 'reason': ...}); never raises for a platform without NUMA data."""
    bus = None
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        if all(hasattr(pr, a) for a in ("pci_bus_id", "pci_device_id", "pci_domain_id")):
            bus = f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:                                     # noqa: BLE001
        bus = None
    if bus is None:
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device_index]) if vis else device_index
            raw = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
            bus = raw.decode() if isinstance(raw, bytes) else raw
        except Exception as exc:                          # noqa: BLE001
            return {"node": None, "reason": f"no PCI bus id: {exc}"}
    node = gpu_numa_node(bus)
    if node is None:
        return {"node": None, "reason": f"no NUMA node for {bus}"}
    cpus = node_cpus(node)
    allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
    if not allowed:
        return {"node": node, "reason": "no allowed CPU on that node"}
    os.sched_setaffinity(0, allowed)
    return {"node": node, "cpus": len(allowed), "pci": bus}
