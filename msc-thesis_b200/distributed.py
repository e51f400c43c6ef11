"""Multi-GPU plumbing: episodes shard trivially (one process per GPU, no data-path collective); the only
exchange is the all-reduce of the per-batch statistics vector, which replaces the gather of
(success, rms_score) tuples over multiprocessing pipes in the reference's main.py:67-76."""
import torch


def shard_range(n_total, rank, world_size):
    """Contiguous block partition of the global episode index: rank r owns [lo, hi).  Seeds and
    hyper-parameters are functions of the GLOBAL index, so results do not depend on world_size."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs that are local to the GPU (NVML's ideal CPU affinity) so that the pinned host
    buffers it allocates next land on that NUMA node: with one process per GPU streaming batches from host memory, remote
    buffers halve the per-GPU copy rate at 8 GPUs.  Returns the CPU list, or None if NVML / the affinity call is
    unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber the devices: map through the PCI bus id
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        h = None
        if bus is not None:
            for i in range(pynvml.nvmlDeviceGetCount()):
                cand = pynvml.nvmlDeviceGetHandleByIndex(i)
                pci = pynvml.nvmlDeviceGetPciInfo(cand)
                if int(pci.bus) == int(bus):
                    h = cand
                    break
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c < n_cpu]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def allreduce_stats(stats, group=None):
    """In-place SUM all-reduce of a statistics vector (NCCL for CUDA tensors, gloo for CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats):
    """Named view of an (all-reduced) statistics vector; `success` as main.py:71-76 counts it."""
    from ._lib import STATS_FIELDS
    v = stats.detach().cpu().tolist() if torch.is_tensor(stats) else list(stats)
    d = dict(zip(STATS_FIELDS, v))
    n = max(d["n"], 1.0)
    d["success_rate"] = d["n_alive"] / n
    d["mean_return_lon"] = d["sum_r_lon"] / n
    d["mean_return_col"] = d["sum_r_col"] / n
    return d
