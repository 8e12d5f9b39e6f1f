"""Data-parallel plumbing: one process per GPU, batch sharding, ONE all-reduce of the flat gradient buffer per step.

The reference is single-GPU (train.py:29,233); the path shards naturally on batch (SURVEY.md section 8 e): every frame is
independent except train-mode BN statistics (kept per replica, like the reference's batch-32 behaviour) and the `mean`
of the two losses (equal shards => averaged gradients == gradient of the global mean).  The collective is a plain
NCCL sum over NVLink (gloo in CPU tests); the 1/world factor is folded into the fused Adam kernel (grad_scale).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None, device=None):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            # the gradient all-reduce overlaps the tail of backward: keep NCCL to the SMs the trainer leaves free for it
            # (FusedTrainer lowers the conv grids by AWR_B200_NCCL_SMS while a bucket is in flight)
            os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("AWR_B200_NCCL_SMS", "16"))
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, local, world


def shard_slice(rank, per_rank_batch):
    """Frames [rank*B, (rank+1)*B) of the global batch belong to `rank`."""
    return slice(rank * per_rank_batch, (rank + 1) * per_rank_batch)


def allreduce_sum_(flat, group=None):
    """In-place sum of the flat gradient buffer over all replicas (no-op for a single process)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_(tensors, src=0, group=None):
    """DDP-style start: every replica takes `src`'s parameters / buffers."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src=src, group=group)


def max_over_ranks(value, device):
    """Timing rule: a multi-GPU duration is the MAX over ranks."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(value)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()
    return float(value)


def allreduce_busbw(flat, reps=10, group=None):
    """Time `reps` sum all-reduces of `flat` alone (every rank must call this).  Returns (ms per all-reduce, max over ranks; bus bandwidth
    in GB/s = 2 (n-1)/n * bytes / t, the figure NCCL quotes against the 900 GB/s per direction of NVLink 5; SURVEY.md section 8 d).
    CUDA tensors are timed with CUDA events on the current stream, CPU tensors (gloo tests) with the wall clock.  The buffer is summed
    `reps` times, so pass a scratch copy."""
    import time
    n = dist.get_world_size(group) if dist.is_initialized() else 1
    if n <= 1:
        return 0.0, 0.0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)          # warm-up (communicator / buffers)
    if flat.is_cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(group)
        e0.record()
        for _ in range(reps):
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    else:
        dist.barrier(group)
        t0 = time.perf_counter()
        for _ in range(reps):
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        ms = (time.perf_counter() - t0) * 1e3 / reps
    ms = max_over_ranks(ms, flat.device)
    nbytes = flat.numel() * flat.element_size()
    return ms, (2.0 * (n - 1) / n * nbytes / (ms * 1e-3) / 1e9) if ms > 0 else 0.0

