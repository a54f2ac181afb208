"""Cell-sharded data parallelism: one process per GPU, torch.distributed (NCCL over NVLink) for
the four small exchanges of the path.  The same helpers run on CPU tensors over
gloo, which is how the host logic is tested without GPUs.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as td


@dataclass
class ShardPlan:
    """Contiguous row ranges per rank, aligned to ``align`` rows (the reference's chunk size, so that
    the chunk-local quirks of smoothen_dists fall inside one shard whenever possible)."""
    starts: list
    stops: list

    @staticmethod
    def make(n_rows: int, world: int, align: int = 1000) -> "ShardPlan":
        n_blocks = (n_rows + align - 1) // align
        base, extra = divmod(n_blocks, world)
        starts, stops, b = [], [], 0
        for r in range(world):
            nb = base + (1 if r < extra else 0)
            starts.append(min(b * align, n_rows))
            stops.append(min((b + nb) * align, n_rows))
            b += nb
        return ShardPlan(starts, stops)

    def rows(self, rank: int):
        return self.starts[rank], self.stops[rank]


class Comm:
    """Thin wrapper: world-1 (no process group) turns every collective into a no-op."""

    def __init__(self, group=None):
        self.enabled = td.is_available() and td.is_initialized()
        self.group = group
        self.rank = td.get_rank(group) if self.enabled else 0
        self.world = td.get_world_size(group) if self.enabled else 1
        # gloo moves host memory: device tensors are staged through the host (used to run several ranks on ONE GPU in
        # the tests, where NCCL refuses two ranks per device; the production backend is NCCL)
        self.staged = self.enabled and td.get_backend(group) == "gloo"

    def _allreduce_(self, t, op):
        if self.world > 1:
            if self.staged and t.is_cuda:
                h = t.cpu()
                td.all_reduce(h, op=op, group=self.group)
                t.copy_(h)
            else:
                td.all_reduce(t, op=op, group=self.group)
        return t

    def allreduce_sum_(self, t):
        return self._allreduce_(t, td.ReduceOp.SUM)

    def allreduce_min_(self, t):
        return self._allreduce_(t, td.ReduceOp.MIN)

    def allreduce_max_(self, t):
        return self._allreduce_(t, td.ReduceOp.MAX)

    def _allgather_into(self, out, t):
        if self.staged and t.is_cuda:
            ho = torch.empty(out.shape, dtype=out.dtype)
            td.all_gather_into_tensor(ho, t.cpu().contiguous(), group=self.group)
            out.copy_(ho)
        else:
            td.all_gather_into_tensor(out, t.contiguous(), group=self.group)

    def allgather_counts(self, n: int, device) -> list:
        if self.world == 1:
            return [n]
        mine = torch.tensor([n], dtype=torch.int64, device=device)
        out = torch.empty(self.world, dtype=torch.int64, device=device)
        self._allgather_into(out, mine)
        return [int(x) for x in out.tolist()]

    def allgather_rows(self, t, counts=None):
        """Concatenates the row blocks of all ranks (uneven counts allowed) -> [sum(counts), ...]."""
        if self.world == 1:
            return t
        if counts is None:
            counts = self.allgather_counts(int(t.shape[0]), t.device)
        mx = max(counts)
        if all(c == mx for c in counts):
            out = torch.empty((self.world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            self._allgather_into(out, t)
            return out
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        buf = torch.empty((self.world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self._allgather_into(buf, pad)
        return torch.cat([buf[r * mx: r * mx + c] for r, c in enumerate(counts)], dim=0)

    def barrier(self):
        if self.world > 1:
            td.barrier(group=self.group)
