"""Frame-level data parallelism: frames are independent units, so a batch is split
contiguously over the ranks of a torch.distributed job (one process per GPU) with no data-path
collective; only the small, variable-length results are gathered on rank 0 when asked for."""
import numpy as np


def frame_range(n_frames, rank, world):
    """Contiguous split of n_frames over `world` ranks; the first n_frames % world ranks get one more."""
    base, extra = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(local_results, lo):
    """Gather per-frame results (any picklable objects) of all ranks on rank 0, in frame order.
    Returns the full list on rank 0 and None elsewhere.  Works with any initialised backend
    (nccl on the GPU box, gloo in the CPU tests); a single process needs no process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_results)
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = (int(lo), list(local_results))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    out = []
    for lo_r, res in sorted(gathered, key=lambda t: t[0]):
        assert lo_r == len(out), "ranks returned overlapping or missing frame ranges"
        out.extend(res)
    return out


def run_sharded(fn, frames):
    """Apply `fn(frames_slice) -> list of per-frame results` to this rank's share of `frames`
    (an array whose first axis is the frame index) and gather on rank 0."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = frame_range(len(frames), rank, world)
    local = fn(np.asarray(frames[lo:hi])) if hi > lo else []
    return gather_results(local, lo)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_cpus(pci_bus_id, sysfs="/sys"):
    """CPUs of the NUMA node the GPU with this PCI id ("0000:1b:00.0", any case, 8-digit domains accepted)
    hangs off, or None when the platform does not say (numa_node < 0, no sysfs)."""
    import os
    bdf = pci_bus_id.strip().lower()
    dom, _, rest = bdf.partition(":")
    if len(dom) == 8:                      # nvidia-smi prints 8-digit domains, sysfs uses 4
        bdf = dom[-4:] + ":" + rest
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")).read())
        if node < 0:
            return None
        return _parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read()) or None
    except (OSError, ValueError):
        return None


def bind_to_gpu_numa(device_index, pci_bus_id=None):
    """Host pipelines that feed one GPU should run (and first-touch their pinned buffers) on the GPU's own
    NUMA node: DMA from the far socket costs about a third of the link rate.  Restricts this process to the
    node's CPUs (intersected with the current mask).  Returns (previous mask, new mask or None if unknown)."""
    import os
    import subprocess
    prev = os.sched_getaffinity(0)
    if pci_bus_id is None:
        try:
            pci_bus_id = subprocess.run(["nvidia-smi", "-i", str(device_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                        capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0]
        except (OSError, IndexError, subprocess.SubprocessError):
            return prev, None
    cpus = gpu_numa_cpus(pci_bus_id)
    if not cpus or not (cpus & prev):
        return prev, None
    os.sched_setaffinity(0, cpus & prev)
    return prev, cpus & prev
