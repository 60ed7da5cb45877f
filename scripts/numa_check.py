"""Where do the arena's pinned buffers live?  Prints, per device, the GPU's NUMA node (sysfs) and the node
distribution of the pages behind raftgpu_host_alloc memory and the staging sets' pinned buffers, read from
/proc/self/numa_maps (the check the round-1 review asked for with numastat; numastat is not in the image).
    python scripts/numa_check.py [device ...]"""
import importlib, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
B = importlib.import_module("raft-rs_b200").binding


def node_pages(addr):
    """{node: pages} of the mapping that contains addr."""
    best = None
    for line in open("/proc/self/numa_maps"):
        parts = line.split()
        start = int(parts[0], 16)
        if start <= addr:
            best = (start, line)
        else:
            break
    if not best:
        return {}
    return {int(m.group(1)): int(m.group(2)) for m in re.finditer(r"N(\d+)=(\d+)", best[1])}


for dev in ([int(x) for x in sys.argv[1:]] or range(torch.cuda.device_count())):
    pr = torch.cuda.get_device_properties(dev)
    bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    try:
        node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
    except OSError:
        node = "?"
    a = B.Arena(200_000, device=dev, n_rings=4)
    a.group_alloc_range(200_000)
    buf = a.host_alloc_bytes(64 << 20)
    buf[:] = 1
    s = B.Synth(200_000, 1)
    a.load_columns(s.initial)
    a.step_begin_records(s.next_round().copy(), B.STEP_READ_COMMITTED)   # touches the staging set's pinned stream
    a.step_wait()
    bm, com = a.step_results(200_000)
    print(f"device {dev} ({bdf}) numa_node {node}: raftgpu_host_alloc pages {node_pages(buf.ctypes.data)}, "
          f"step results (pinned) pages {node_pages(com.ctypes.data)}")
    a.close()
