"""Host placement for the ingest path of a multi-GPU box: one process per GPU, each bound to the CPUs of the NUMA node
its GPU hangs off (sysfs: /sys/bus/pci/devices/<gpu>/local_cpulist), so that the pinned staging buffers it allocates
afterwards are first-touched on that node and the gather threads and the PCIe DMA stay on one socket.  Nothing here
touches the feature path; without sysfs access (or with a cgroup that excludes the node's CPUs) it does nothing."""
from __future__ import annotations

import os
from typing import Optional


def _parse_cpulist(s: str) -> set:
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def gpu_numa_info(device_index: int) -> dict:
    """{'pci': '0000:xx:00.0', 'numa_node': int or None, 'local_cpus': set} of a visible CUDA device."""
    import torch
    pr = torch.cuda.get_device_properties(device_index)
    addr = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    info = {"pci": addr, "numa_node": None, "local_cpus": set()}
    try:
        info["numa_node"] = int(open(f"/sys/bus/pci/devices/{addr}/numa_node").read().strip())
        info["local_cpus"] = _parse_cpulist(open(f"/sys/bus/pci/devices/{addr}/local_cpulist").read())
    except (OSError, ValueError):
        pass
    return info


def bind_to_gpu_numa_node(device_index: int) -> Optional[dict]:
    """Restrict the calling process to the CPUs local to the GPU (intersected with what it may use).  Returns what was
    done, or None when nothing could be done."""
    if not hasattr(os, "sched_setaffinity"):
        return None
    info = gpu_numa_info(device_index)
    allowed = os.sched_getaffinity(0)
    use = info["local_cpus"] & allowed
    if not use or len(use) == len(allowed):
        return {"bound": False, "numa_node": info["numa_node"], "cpus": len(allowed)}
    os.sched_setaffinity(0, use)
    return {"bound": True, "numa_node": info["numa_node"], "cpus": len(use)}


def restore_affinity(cpus) -> None:
    """Give every thread of this process the CPU set `cpus` again (sched_setaffinity acts on one thread; worker pools
    created while the process was bound have inherited the narrow mask)."""
    if not hasattr(os, "sched_setaffinity"):
        return
    try:
        tids = [int(t) for t in os.listdir("/proc/self/task")]
    except OSError:
        tids = [0]
    for tid in tids:
        try:
            os.sched_setaffinity(tid, cpus)
        except OSError:
            pass


def ingest_threads(local_world_size: int = 1) -> int:
    """Gather threads per rank: the CPUs this process may use, shared among the ranks of the box, 2..32."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return max(2, min(32, n // max(1, local_world_size)))
