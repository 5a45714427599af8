"""Host topology probe for the e2e leg: NUMA nodes, the GPU's node, and pinned-memory H2D bandwidth per node."""
import glob
import os
import time

import torch


def parse_cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def main():
    aff0 = os.sched_getaffinity(0)
    print("affinity", len(aff0), sorted(aff0)[:4], "...", sorted(aff0)[-4:])
    nodes = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        n = int(p.rsplit("node", 1)[1])
        nodes[n] = parse_cpulist(open(p + "/cpulist").read())
        print("node", n, "cpus", len(nodes[n]), "usable", len(nodes[n] & aff0))
    pr = torch.cuda.get_device_properties(0)
    addr = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    try:
        print("gpu", addr, "numa_node", open(f"/sys/bus/pci/devices/{addr}/numa_node").read().strip())
    except OSError as e:
        print("gpu", addr, "numa_node unreadable:", e)
    n = 264 * 1000 * 1000 // 4
    dst = torch.empty(n, device="cuda:0")
    for node, cpus in nodes.items():
        use = cpus & aff0
        if not use:
            continue
        os.sched_setaffinity(0, use)
        time.sleep(0.01)
        host = torch.empty(n).pin_memory()
        host.fill_(1.0)
        for _ in range(2):
            dst.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            dst.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"pinned on node {node}: {n * 4 / ms / 1e6:.1f} GB/s")
        del host
        os.sched_setaffinity(0, aff0)


if __name__ == "__main__":
    main()
