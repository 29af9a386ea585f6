"""Multi-GPU plumbing: envs are independent, so the batch is sharded by contiguous index ranges
(one process per GPU, tables replicated) and the ONLY collective on the path is a SUM all-reduce
of the episode statistics {episodes, successes, env_steps, reserved} — the quantities the
reference accumulates with util.add_stat (trainers/imitation.py:131-134,197-198).
NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors (tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises torch.distributed from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).
    Returns (rank, world_size, local_rank).  Single-process runs need no initialisation."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        # NCCL's banner/debug lines go to stdout by default; callers print JSON there
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
            # opt-in: a measured no-op on single-NUMA-node hosts like this pool's (profiles/README.md)
            if os.environ.get("PSK_BIND_CPUS", "0") == "1":
                bind_to_gpu_cpus(local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def bind_to_gpu_cpus(device_index):
    """Restricts this process to the CPUs that NVML reports as local to its GPU (same NUMA node /
    PCIe root), so that pinned host buffers allocated afterwards — the host-buffer path moves
    112 MB per step and GPU through them — land in memory next to the GPU instead of behind the
    socket interconnect.  Best effort: returns the CPU list, or None when NVML / the affinity call
    is unavailable or the box exposes a single node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:        # CUDA and NVML may number the devices differently (CUDA_VISIBLE_DEVICES)
            pr = torch.cuda.get_device_properties(int(device_index))
            bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:  # noqa: BLE001
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed or len(allowed) == len(os.sched_getaffinity(0)):
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001
        return None


def shard_range(n_total, rank, world):
    """Contiguous slice [lo, hi) of the env index range owned by ``rank`` (sizes differ by <= 1)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_stats(stats):
    """In-place SUM over ranks of a stats tensor (int64[4] as kept by VecCraft.stats)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def allreduce_max(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
