"""The only collective of the algorithm: reduction of per-frame GMM statistics across the GPUs of one box.

One process per GPU (torchrun), frames sharded by k; `StatsComm` wraps a torch.distributed process group (NCCL over
NVLink on the B200 box, gloo in the CPU tests).  Payloads are O(C (D+3)) floats per EM step (about 1 KB at C = 50):
latency-bound, so everything that can travel together is fused into one buffer per reduction.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


class StatsComm:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    # -- plain reductions -----------------------------------------------------------------------------------
    def sum(self, t):
        t = t.contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def max(self, t):
        t = t.contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def broadcast(self, t, src=0):
        t = t.contiguous()
        dist.broadcast(t, src=src, group=self.group)
        return t

    # -- log-domain merges -----------------------------------------------------------------------------------
    def merge_colstats(self, stats):
        """stats (C, D+3) = [m (log2 exponent), S0, B (D), A] of the local points -> statistics of all points:
        MAX-reduce the exponents, rescale the local sums to the common exponent, SUM-reduce."""
        m_loc = stats[:, 0].clone()
        m_glob = self.max(m_loc.clone())
        scaled = stats[:, 1:] * torch.exp2(m_loc - m_glob)[:, None]
        scaled = self.sum(scaled)
        return torch.cat((m_glob[:, None], scaled), dim=1)

    def merge_colstats_ref(self, stats, m_ref, extra=None):
        """ONE all-reduce (SUM) for the column statistics AND whatever plain sums ride along.
        stats (C, D+3) = [m, S0, B, A] of the local points; m_ref (C,): an exponent every rank agrees on bit for bit (the log2
        column mass of the previous EM step, see GaussianMixtureUnif); extra: 1-D tensor of per-rank partial sums.
        Every rank rescales its sums to 2^m_ref -- no MAX round -- and appends a flag (1 if one of its exponents is so far
        above m_ref that the rescaled sums could overflow).  Returns (merged stats with exponent m_ref, reduced extra,
        flag sum); the caller falls back to `merge_colstats` when flag > 0 or a merged S0 vanished (all ranks see the same
        numbers, so they take the same decision)."""
        d = stats[:, 0] - m_ref
        scaled = stats[:, 1:] * torch.exp2(torch.clamp(d, max=120.0))[:, None]
        flag = (d > 100.0).any().to(stats.dtype).reshape(1)
        parts = [scaled.reshape(-1), flag]
        if extra is not None:
            parts.append(extra.to(stats.dtype).reshape(-1))
        buf = self.sum(torch.cat(parts))
        n = scaled.numel()
        merged = torch.cat((m_ref[:, None], buf[:n].view_as(scaled)), dim=1)
        return merged, (buf[n + 1:] if extra is not None else None), buf[n]

    def logsumexp_pair(self, a, b):
        """Global logsumexp of two per-rank log-sums (outlier log-odds update, core/GMM.py:290)."""
        v = torch.stack((a, b))
        m = self.max(v.clone())
        s = self.sum(torch.exp(v - m))
        out = m + torch.log(s)
        return out[0], out[1]


def shard_frames(K, rank, world, weights=None):
    """Indices of the frames owned by `rank`: k mod world by default; with per-frame weights (point counts) a greedy
    longest-processing-time assignment that balances the registration work."""
    if weights is None:
        return [k for k in range(K) if k % world == rank]
    order = sorted(range(K), key=lambda k: -weights[k])
    load = [0.0] * world
    owner = [0] * K
    for k in order:
        r = min(range(world), key=lambda i: (load[i], i))
        owner[k] = r
        load[r] += weights[k]
    return [k for k in range(K) if owner[k] == rank]
