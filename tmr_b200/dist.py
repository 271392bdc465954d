"""Multi-GPU glue: one process per GPU (torchrun), NCCL inside the CUDA layer.

`init_from_torch()` ships the NCCL unique id through an already initialised
torch.distributed process group (nccl or gloo) and attaches this process to
the forest communicator -- the equivalent of passing an MPI communicator to
TMROctForest in the reference (src/TMROctForest.h:51).
"""
import ctypes

ID_BYTES = 128


def make_unique_id(lib):
    buf = ctypes.create_string_buffer(ID_BYTES)
    if lib.tmrgpu_comm_unique_id(buf, ID_BYTES) != 0:
        raise RuntimeError("tmr_b200: could not create a communicator id")
    return buf.raw


def init_world(lib, rank, size, unique_id):
    """Attach this process (or thread) as `rank` of `size`.  Collective."""
    lib.tmr_b200_init_world.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
    if lib.tmr_b200_init_world(rank, size, unique_id) != 0:
        raise RuntimeError("tmr_b200: communicator initialisation failed")


def partition_counts(total, size, max_rank=-1):
    """Equal-count SFC re-split used by repartition() (reference
    src/TMROctForest.cpp:1949-1964): rank k < max_rank gets total//max_rank
    octants, +1 for the first total % max_rank ranks; later ranks get none."""
    if max_rank <= 0 or max_rank > size:
        max_rank = size
    avg, rem = divmod(total, max_rank)
    return [avg + (1 if k < rem else 0) if k < max_rank else 0 for k in range(size)]


def init_from_torch(lib=None):
    """Use torch.distributed (already initialised) to distribute the id."""
    import torch
    import torch.distributed as dist

    if lib is None:
        from . import load_library

        lib = load_library()
    rank, size = dist.get_rank(), dist.get_world_size()
    if size == 1:
        return rank, size
    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(ID_BYTES, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_unique_id(lib)
        t.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    init_world(lib, rank, size, bytes(t.cpu().tolist()))
    return rank, size


def bind_to_gpu_numa(device_index, local_world=1):
    """Pin this process to CPU cores local to its GPU (NVML's ideal CPU
    affinity), so that page-locked host mirrors are allocated on the GPU's own
    NUMA node: with one process per GPU, node-array read-backs then do not
    cross the inter-socket link.  When several of the `local_world` GPUs of
    this job share one affinity set, each process takes its own contiguous
    slice of it, so that the host threads that rebuild the node arrays
    (ops_nodes.h: dep_expand_host) of one rank do not run on another rank's
    cores.  Returns the number of CPUs bound to, or 0 when NVML is unavailable
    (nothing changed)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        words = (os.cpu_count() + 63) // 64
        allowed = set(os.sched_getaffinity(0))

        def cpus_of(index):
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            return [64 * w + b for w, m in enumerate(mask) for b in range(64)
                    if (m >> b) & 1 and (64 * w + b) in allowed]

        cpus = cpus_of(device_index)
        if local_world > 1 and cpus:
            sharing = [d for d in range(local_world) if cpus_of(d) == cpus]
            k, n = sharing.index(device_index), len(sharing)
            if len(cpus) >= 2 * n:
                cpus = cpus[len(cpus) * k // n:len(cpus) * (k + 1) // n]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001
        return 0
