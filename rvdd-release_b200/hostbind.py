"""Pin the calling process to the CPU cores (and thereby the memory) of the NUMA node its GPU hangs off.

The end-to-end path moves ~1 GB per 29-pair batch between pinned host buffers and the GPU.  With one process per GPU
on a two-socket box, a process that runs -- and first-touches its pinned buffers -- on the other socket pushes all of
that over the inter-socket link, which all such processes share.  Call ``bind_to_gpu(local_rank)`` BEFORE allocating
pinned memory (bench.py and the precompute driver do).  Best effort: any failure leaves the affinity untouched.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index):
    """NUMA node of CUDA device `index` (as torch numbers it), or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        addr = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % addr) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu(index):
    """-> dict describing what was done (goes into the bench's JSON line)."""
    info = {"gpu": int(index), "numa_node": None, "cpus": None, "bound": False}
    try:
        node = gpu_numa_node(index)
        info["numa_node"] = node
        if node is None:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return info
        os.sched_setaffinity(0, cpus)
        info["cpus"], info["bound"] = len(cpus), True
        try:        # prefer the node's memory for everything allocated from now on (pinned buffers included)
            import ctypes
            numa = ctypes.CDLL("libnuma.so.1")
            if numa.numa_available() >= 0:
                numa.numa_set_preferred(node)
                info["mem_policy"] = "preferred"
        except OSError:
            pass
    except Exception as e:      # never fatal
        info["error"] = repr(e)
    return info
