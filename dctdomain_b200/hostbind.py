"""Host-side placement for one process per GPU: run on the CPUs of the NUMA node the GPU hangs off.

Staging host embeddings (fingerprint.quantize_batch, src/make_db.py:36-51 in the reference's flow) is bound by the
host: the gather kernel reads the pinned arrays over PCIe at ~51 GB/s per GPU.  With several ranks on a two-socket box
a rank whose pinned memory sits on the other socket pulls every byte over the socket interconnect as well.  Linux
places pages on the node of the thread that first touches them (cudaHostAlloc included), so binding the rank's threads
BEFORE it allocates is enough; nothing here touches device code.
"""
import os


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(device_index: int):
    """(numa_node, set of CPUs) of a CUDA device, from sysfs; (None, None) when the platform does not say."""
    import torch
    pr = torch.cuda.get_device_properties(device_index)
    bdf = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
    try:
        with open(f'/sys/bus/pci/devices/{bdf}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            return None, None
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            return node, _parse_cpulist(f.read())
    except (OSError, ValueError):
        return None, None


def bind_host_to_gpu(device_index: int) -> dict:
    """Restricts the calling thread (and every thread it starts afterwards) to the CPUs next to the GPU.  Returns
    {'numa_node', 'cpus', 'previous'}; 'cpus' is None when nothing was changed (single-node host, no sysfs entry, or the
    node's CPUs are outside this process's cgroup).  Undo with ``os.sched_setaffinity(0, info['previous'])``."""
    prev = os.sched_getaffinity(0)
    node, cpus = gpu_numa_cpus(device_index)
    info = {'numa_node': node, 'cpus': None, 'previous': prev, 'n_nodes': None}
    try:
        info['n_nodes'] = len([d for d in os.listdir('/sys/devices/system/node') if d.startswith('node') and d[4:].isdigit()])
    except OSError:
        pass
    if cpus is None:
        return info
    want = cpus & prev
    if not want or want == prev:
        return info
    os.sched_setaffinity(0, want)
    info['cpus'] = sorted(want)
    return info
