"""Stream-parallel replicas: the only multi-GPU structure of the path.

Independent speech streams share nothing but read-only weights, so stream `s` is owned by rank
`s mod world` (one process per GPU, model replicated) and the hot path runs no collective.  The
reference's precedent is one GPU per process under SLURM arrays (scripts/infer/infinisst.sh:5-13).
`torch.distributed` is used only around the timed region: a barrier, and a max/sum of a few floats.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def env_world() -> Dict[str, int]:
    return {"rank": int(os.environ.get("RANK", "0")), "local_rank": int(os.environ.get("LOCAL_RANK", "0")),
            "world": int(os.environ.get("WORLD_SIZE", "1"))}


def shard_streams(n_streams: int, world: int, rank: int) -> List[int]:
    """Global stream ids owned by `rank`: round-robin, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_streams, world))


def init(backend: str, device: int = -1) -> None:
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    if env_world()["world"] > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": torch.device(f"cuda:{device}")} if (backend == "nccl" and device >= 0) else {}
        dist.init_process_group(backend=backend, **kw)


def shutdown() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_stats(local_ms: float, local_units: float, device: str = "cpu") -> Dict[str, float]:
    """Whole-job view of one timed region: time = max over ranks, units = sum over ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return {"ms": local_ms, "units": local_units, "world": 1}
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    u = torch.tensor([local_units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return {"ms": float(t.item()), "units": float(u.item()), "world": dist.get_world_size()}


def gather_floats(vals: Sequence[float], device: str = "cpu") -> List[float]:
    """All ranks' values concatenated (per-chunk latencies for whole-job p50/p99)."""
    if not (dist.is_available() and dist.is_initialized()):
        return list(vals)
    out: List[List[float]] = [None] * dist.get_world_size()          # type: ignore[list-item]
    dist.all_gather_object(out, list(vals))
    return [v for part in out for v in part]
