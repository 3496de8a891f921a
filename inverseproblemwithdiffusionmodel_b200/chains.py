"""Multi-GPU layer of the path: independent reconstruction chains sharded over ranks, and the one
collective the path has -- the final per-pixel mean / standard deviation over all chains.

The reference has no multi-GPU code; its "mean of 105 reconstructions" is numpy
mean/std over a stack of results (helpers/visualizations.py:93-95,117-142).  Here every rank owns the
chains `i % world == rank`, accumulates sum|x|, sum|x|^2, sum(angle x), sum(angle x)^2 with
`ipdm_chain_stats_accumulate` (float64), and a single all-reduce (NCCL on GPUs, gloo in CPU tests)
merges them.  A rank passes its global chain indices to the sampler (`chain_ids=chain_partition(...)`, one seed on
every rank): the in-kernel noise of chain i is Philox(seed, counter = (pixel, i, step)), so chain i is the same chain
on any rank of any world size and at any slot of the rank's batch (tests: test_chain_noise_is_keyed_by_global_chain_id).
A rank that owns no chain (more ranks than chains) skips sampling but must still call `all_reduce`.
"""
import os

import torch
import torch.distributed as dist

from . import _lib


def chain_partition(n_chains, world, rank):
    """Global chain indices owned by `rank` (round-robin: 105 chains over 8 ranks -> 14,13,...,13)."""
    return list(range(rank, n_chains, world))


def init_distributed(backend=None):
    """One process per GPU from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); no-op for 1 rank."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


class PosteriorStats:
    """Sufficient statistics of magnitude and phase over chains; float64 [4][H*W] + count."""

    def __init__(self, hw, device):
        self.hw = hw
        self.acc = torch.zeros(4, hw, dtype=torch.float64, device=device)
        self.count = torch.zeros(1, dtype=torch.float64, device=device)

    def add(self, x):
        """x: complex64 CUDA tensor (chains, ..., H, W) with prod(...)*H*W == hw, on the accumulator's device."""
        _lib.require_cuda(x)
        x = x.to(torch.complex64).contiguous()
        chains = x.shape[0]
        if x[0].numel() != self.hw:
            raise ValueError("image size mismatch")
        _lib.check(_lib.lib().ipdm_chain_stats_accumulate(x.data_ptr(), self.acc.data_ptr(), chains, self.hw, _lib.stream()),
                   "chain_stats_accumulate")
        self.count += chains

    def add_sums(self, sums, chains):
        """Merge already-reduced statistics (float64 [4][hw]) of `chains` chains, e.g. from another accumulator."""
        self.acc += sums.to(self.acc.device, torch.float64).reshape(4, self.hw)
        self.count += chains

    def all_reduce(self):
        """Sum the statistics over all ranks (the path's only collective)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            flat = torch.cat([self.acc.reshape(-1), self.count])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            self.acc = flat[:-1].reshape(4, self.hw)
            self.count = flat[-1:].clone()
        return self

    def finalize(self, shape):
        """mean / population std of |x| and angle(x) per pixel, float32 tensors of `shape`."""
        n = self.count.item()
        m_mag, m_ph = self.acc[0] / n, self.acc[2] / n
        v_mag = (self.acc[1] / n - m_mag ** 2).clamp_min(0)
        v_ph = (self.acc[3] / n - m_ph ** 2).clamp_min(0)
        f = lambda t: t.float().reshape(shape)
        return {"mag_mean": f(m_mag), "mag_std": f(v_mag.sqrt()), "phase_mean": f(m_ph), "phase_std": f(v_ph.sqrt()), "n": int(n)}
