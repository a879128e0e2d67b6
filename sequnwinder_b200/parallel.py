"""Multi-GPU placement of the ADMM row blocks: one process per GPU, launched by torchrun.

The reference's only parallelism is consensus ADMM over T row blocks on T Java threads
(framework/LOne.java:337-357, 395-417, relative to
/root/reference/src/org/seqcode/projects/sequnwinder/). T stays the numerics knob; the number of
GPUs only decides where a block lives: block t is solved by rank ``t % world``. The per-iteration
consensus sums are an ``ncclAllReduce`` issued by the library itself on its own communicator;
``torch.distributed`` is only the plumbing that hands rank 0's ``ncclUniqueId`` to the others.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

from .optimizer import java_block_order


@dataclass
class RankInfo:
    rank: int
    world: int
    local_rank: int


def rank_info_from_env() -> RankInfo:
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                    int(os.environ.get("LOCAL_RANK", "0")))


def local_blocks(T: int, rank: int, world: int) -> List[int]:
    """ADMM blocks owned by ``rank``, in the order the library stores and sums them
    (java.util.HashMap order of "Thread<t>", LOne.java:399,411,427)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    blocks = [t for t in java_block_order(T) if t % world == rank]
    if not blocks:
        raise ValueError(f"rank {rank} owns no ADMM block (T={T}, world={world})")
    return blocks


def block_rows(n_rows: int, T: int, t: int):
    """Rows of block t: i = t + T*s for s < floor(N/T); the last N mod T rows belong to no block
    (LOne.java:337-357)."""
    nb = n_rows // T
    return [t + T * s for s in range(nb)]


def local_row_indices(n_rows: int, T: int, rank: int, world: int):
    """Global indices of the rows that live on ``rank`` (the rows of its blocks), ascending: the
    rows a rank featurises itself when path 1 is sharded by rows, and the row order the
    ``rows_local`` form of ``sqw_admm_set_problem_counts_dev`` expects."""
    import numpy as np

    nb = n_rows // T
    i = np.arange(nb * T, dtype=np.int64)
    return i[(i % T) % world == rank]


def local_ridges(ridges, rank: int, world: int):
    """Regularisation sweep (BASELINE config 5) on several GPUs: the solves are independent, so
    they are dealt out as replicas, no collective: value i runs on rank ``i % world``. Returns
    [(index, lambda)] for ``rank``; every rank then calls ``LOneCuda.sweep`` on its own values."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("invalid rank/world")
    return [(i, float(v)) for i, v in enumerate(ridges) if i % world == rank]


def broadcast_bytes(payload: Optional[bytes], nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string over the default torch.distributed group (any backend)."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def share_nccl_id() -> bytes:
    """Rank 0 creates the library's ncclUniqueId; everyone returns the same 128 bytes."""
    import torch.distributed as dist
    from .optimizer import nccl_unique_id

    payload = nccl_unique_id() if dist.get_rank() == 0 else None
    return broadcast_bytes(payload, 128, src=0)
