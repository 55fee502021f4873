"""Multi-GPU host logic: one process per GPU, Markov chains sharded contiguously, parameters replicated.

What crosses ranks (and nothing else): per SR iteration ONE all-reduce(sum) of [sum O (P complex) | sum O conj(h) (P complex) |
sum |O|^2 (P real) | sum h (complex) | sum |h|^2] = 5P+3 doubles, and per CG iteration ONE all-reduce(sum) of the P complex
partial products O_loc^H (O_loc v).  Every scalar product of the CG runs on replicated vectors, so all ranks take identical
decisions (NCCL all-reduce returns the same bits on every rank).  The internal RNG is keyed by the GLOBAL chain id, so the
Markov chains do not depend on how many GPUs they are spread over.

The NCCL communicator lives inside libnqs_b200.so (the all-reduces are enqueued on the engine's stream between its kernels);
torch.distributed is only the bootstrap channel for the 128-byte NCCL unique id.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class ShardPlan:
    """Contiguous split of K_total chains over `world` ranks: rank g owns [offset, offset + n_local)."""
    n_total: int
    world: int
    rank: int

    def __post_init__(self):
        if not (self.world >= 1 and 0 <= self.rank < self.world):
            raise ValueError("bad rank/world")
        if self.n_total < self.world:
            raise ValueError("fewer chains than ranks")

    @property
    def n_local(self) -> int:
        base, rem = divmod(self.n_total, self.world)
        return base + (1 if self.rank < rem else 0)

    @property
    def offset(self) -> int:
        base, rem = divmod(self.n_total, self.world)
        return self.rank * base + min(self.rank, rem)

    def engine_kwargs(self) -> dict:
        return {"n_chains": self.n_local, "n_chains_total": self.n_total, "chain_offset": self.offset}


def sr_allreduce_layout(P: int) -> dict:
    """Offsets (in doubles) inside the single SR-setup all-reduce buffer (mirrors csrc/sr_kernels.cuh:setup_finalize_kernel)."""
    return {"sum_O_re": 0, "sum_O_im": P, "sum_Ohc_re": 2 * P, "sum_Ohc_im": 3 * P, "sum_O2": 4 * P,
            "sum_h_re": 5 * P, "sum_h_im": 5 * P + 1, "sum_h2": 5 * P + 2, "count": 5 * P + 3}


def bootstrap_comm(engine, world: int, rank: int, group=None, p2p: bool = True) -> None:
    """Create the engine's NCCL communicator: rank 0 draws the unique id, torch.distributed broadcasts it."""
    if world == 1:
        return
    import torch.distributed as dist
    from .engine import Engine
    ids = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=group)
    engine.comm_init(world, rank, ids[0])
    if p2p:
        enable_p2p(engine, world, group)


def enable_p2p(engine, world: int, group=None) -> bool:
    """Map every rank's CG exchange buffer into every peer (cudaIpc over NVLink) so that the per-iteration all-reduce runs
    inside the CG kernel.  All ranks must agree: if any rank cannot map its peers, every rank stays on ncclAllReduce."""
    import torch.distributed as dist
    handles = [None] * world
    dist.all_gather_object(handles, engine.comm_p2p_export(), group=group)
    try:
        ok = engine.comm_p2p_import(handles)
    except Exception:
        ok = False
    oks = [None] * world
    dist.all_gather_object(oks, bool(ok), group=group)
    if not all(oks):
        engine.comm_p2p_disable()
        return False
    return True


def make_sharded_engine(model: str, n_inputs: int, n_hiddens: int, n_chains_total: int, h: float, J: float, alpha: float,
                        world: int, rank: int, device: Optional[int] = None, **kw):
    """Engine for this rank's shard with the communicator set up (call under an initialised torch.distributed group)."""
    from .engine import Engine
    plan = ShardPlan(n_chains_total, world, rank)
    e = Engine(model, n_inputs, n_hiddens, h=h, J=J, alpha=alpha, device=rank if device is None else device,
               **plan.engine_kwargs(), **kw)
    bootstrap_comm(e, world, rank)
    return e
