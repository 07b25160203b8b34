"""Multi-GPU: static sharding of independent trajectories / lattice realisations across ranks
(one process per GPU) and the single collective the path needs — a SUM all-reduce of the tally
vector (NCCL over NVLink on GPUs, gloo on CPUs for the host-logic tests).

The reference has no parallelism at all: a ``Lattice`` shares nothing with
another one, so the data path needs no exchange.  Philox stream ids are *global* trajectory and
realisation indices, hence the summed tallies do not depend on the number of ranks.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from ._native import HF_N_TALLIES, TALLY_NAMES


@dataclass(frozen=True)
class Shard:
    """Block partition of ``n_trajectories`` trajectories over ``n_realisations`` realisations."""
    rank: int
    world: int
    first_trajectory: int
    n_trajectories: int
    first_realisation: int
    n_realisations: int


def shard(n_trajectories: int, n_realisations: int, rank: int, world: int) -> Shard:
    """Rank ``rank`` of ``world`` gets a contiguous block of whole realisations when there are at
    least ``world`` of them, otherwise a contiguous block of the trajectories of one shared
    realisation set (every rank then holds all realisations it touches)."""
    if n_trajectories % n_realisations:
        raise ValueError("n_realisations must divide n_trajectories")
    per_real = n_trajectories // n_realisations
    if n_realisations >= world:
        if n_realisations % world:
            raise ValueError("world size must divide n_realisations")
        r = n_realisations // world
        return Shard(rank, world, rank * r * per_real, r * per_real, rank * r, r)
    if n_realisations != 1:
        raise ValueError("with fewer realisations than ranks, use exactly one shared realisation")
    if n_trajectories % world:
        raise ValueError("world size must divide n_trajectories")
    t = n_trajectories // world
    return Shard(rank, world, rank * t, t, 0, 1)


def env_rank_world() -> tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_process_group(backend: str | None = None):
    """Initialise ``torch.distributed`` from the torchrun environment; returns (rank, local_rank, world)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def all_reduce_vector(vec: np.ndarray) -> np.ndarray:
    """SUM all-reduce of a host int64/uint64 vector over the default process group (identity when
    the group is not initialised).  Used by the CPU (gloo) tests and as the generic path."""
    import torch
    import torch.distributed as dist
    out = np.ascontiguousarray(vec).astype(np.int64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.from_numpy(out)
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out = t.cpu().numpy()
    return out


def all_reduce_tallies(sim) -> dict:
    """Box totals over all ranks.  On GPUs the local totals are reduced on the device into a
    torch tensor (``hf_reduce_tallies_device``) and all-reduced with NCCL without a host round
    trip; the result is read once."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl":
        t = torch.zeros(HF_N_TALLIES, dtype=torch.int64, device="cuda")
        sim.reduce_tallies_device(t.data_ptr())          # on the handle's own stream, which it keeps
        sim.sync()                                       # ... finished before NCCL reads t on torch's stream
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        vec = t.cpu().numpy()
    else:
        local = sim.tallies()
        vec = all_reduce_vector(np.array([local[k] for k in TALLY_NAMES] + [0] * (HF_N_TALLIES - len(TALLY_NAMES))))
    return {k: int(vec[i]) for i, k in enumerate(TALLY_NAMES)}


def all_reduce_named(local: dict) -> dict:
    """SUM all-reduce of a dict of integer tallies (``Sim.x_tallies()``, ``Sim.xd_tallies()``: emission
    and quenching channels of the exciton extensions) over the default process group: the one collective of
    those paths, like ``all_reduce_tallies`` for the reference's counters.  Key order must agree on all ranks
    (it is the order of the C-ABI's HF_XT_* / HF_XDT_* enums)."""
    keys = list(local)
    vec = all_reduce_vector(np.array([int(local[k]) for k in keys], dtype=np.int64))
    return {k: int(v) for k, v in zip(keys, vec)}
