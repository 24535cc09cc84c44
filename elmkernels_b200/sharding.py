"""Multi-GPU decomposition of the column timestep: contiguous column ranges per rank, no exchange on the step
(columns are independent: no function of the path reads a neighbouring column).  The
only collective is the optional global balance diagnostic - the counterpart of the reference's
ELMKokkos::min_max_sum (src/utils/kokkos_utils.hh:13-58, parallel_reduce + MPI_Reduce)."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the columns owned by `rank`: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_diagnostics(local: np.ndarray, device=None) -> np.ndarray:
    """All-reduce of one rank's elmk_diag_reduce output (8 sums, 8 minima, 8 maxima) over the default
    torch.distributed process group (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(np.asarray(local, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    s, lo, hi = t[:8].clone(), t[8:16].clone(), t[16:24].clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return torch.cat([s, lo, hi]).cpu().numpy()
