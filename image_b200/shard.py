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
