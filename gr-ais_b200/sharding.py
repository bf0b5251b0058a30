"""Multi-GPU plumbing: channels are independent, so the only exchange is a one-off
broadcast of the preamble template.  One process per GPU,
torch.distributed for the collective (NCCL on the GPU box, gloo in CPU tests).
"""
import numpy as np


def partition(channels, world_size, rank):
    """Contiguous split of [0, channels) over ranks: rank r owns [lo, hi)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    lo = channels * rank // world_size
    hi = channels * (rank + 1) // world_size
    return lo, hi


def broadcast_template(template, src=0, device=None):
    """Every rank returns rank `src`'s complex64 template (length first, then the taps).

    template may be None on the other ranks.  Without an initialised process group
    (single process) the template is returned unchanged."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.ascontiguousarray(template, dtype=np.complex64)
    rank = dist.get_rank()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
            else torch.device("cpu")
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        template = np.ascontiguousarray(template, dtype=np.complex64)
        n[0] = len(template)
    dist.broadcast(n, src=src)
    buf = torch.zeros((int(n.item()), 2), dtype=torch.float32, device=device)
    if rank == src:
        buf.copy_(torch.from_numpy(template.view(np.float32).reshape(-1, 2).copy()))
    dist.broadcast(buf, src=src)
    return buf.cpu().numpy().reshape(-1).view(np.complex64).copy()


def max_over_ranks(value, device=None):
    """MAX all-reduce of a scalar (timings are reported as the max over ranks)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" \
            else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs local to `device_index` (NVML's ideal affinity), so
    that the pinned staging buffers it allocates next live on the GPU's own NUMA node (each rank
    streams 6 GB per step through them).  On the single-NUMA-node boxes this was measured on it
    changes nothing (8 GPUs: 448 k channels/s end to end with and without).  Returns the CPU set,
    or None when NVML is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = None
        try:
            import torch
            pr = torch.cuda.get_device_properties(device_index)
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        except Exception:
            bus = None
        h = (pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()) if bus
             else pynvml.nvmlDeviceGetHandleByIndex(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        return None
    return None
