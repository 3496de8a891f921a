"""Multi-rank host logic on CPU: chain partition + the posterior-statistics all-reduce with gloo,
world_size 2 (the GPU path uses the same code with NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.fixture_inputs import crandn
from oracle import ald as OALD


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sums(x):
    mag, ang = x.abs().reshape(x.shape[0], -1).double(), torch.angle(x).reshape(x.shape[0], -1).double()
    return torch.stack([mag.sum(0), (mag * mag).sum(0), ang.sum(0), (ang * ang).sum(0)])


def _worker(rank, world, port, n_chains, outdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from inverseproblemwithdiffusionmodel_b200 import chains as CH
    r, local, w = CH.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    mine = CH.chain_partition(n_chains, world, rank)
    allx = crandn(99, n_chains, 1, 8, 8)            # chain i is the same tensor on every rank (keyed by global index)
    st = CH.PosteriorStats(64, torch.device("cpu"))
    if mine:                                         # a rank that owns no chain adds nothing but still joins the collective
        st.add_sums(_sums(allx[mine]), len(mine))
    st.all_reduce()
    out = st.finalize((8, 8))
    torch.save({"mine": mine, "out": out}, os.path.join(outdir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_partition():
    from inverseproblemwithdiffusionmodel_b200 import chains as CH
    parts = [CH.chain_partition(105, 8, r) for r in range(8)]
    assert [len(p) for p in parts] == [14] + [13] * 7
    assert sorted(sum(parts, [])) == list(range(105))
    assert CH.chain_partition(3, 8, 5) == []


import pytest


@pytest.mark.parametrize("n_chains", [7, 1])        # 1 chain over 2 ranks: rank 1 owns nothing (more ranks than chains)
def test_posterior_allreduce_gloo(tmp_path, n_chains):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_chains, str(tmp_path)), nprocs=world, join=True)
    ref = OALD.posterior_stats(crandn(99, n_chains, 1, 8, 8))
    seen = []
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        seen += d["mine"]
        assert d["out"]["n"] == n_chains
        for k in ("mag_mean", "mag_std", "phase_mean", "phase_std"):
            assert torch.allclose(d["out"][k], ref[k].reshape(8, 8), atol=1e-5), k
    assert sorted(seen) == list(range(n_chains))
