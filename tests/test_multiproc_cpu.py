"""CPU, world_size 2 over gloo: the multi-GPU host logic of predict (interval sharding + the single final gather)
and of data-parallel training (flat gradient all-reduce + identical optimizer state)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker_predict(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mural_b200.predict import gather_rows, shard_bounds
    lo, hi = shard_bounds(n, world, rank)
    # stand-in for the per-site network: any pure function of the site index
    idx = torch.arange(lo, hi, dtype=torch.float64)
    local = torch.stack([idx, idx * 2 + 1, torch.sin(idx), idx ** 2], 1)
    full = gather_rows(local, n, world, rank)
    if rank == 0:
        torch.save(full, out)
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [0, 1, 7, 1000, 1001])
def test_predict_sharding_gather_world2(tmp_path, n):
    from mural_b200.predict import shard_bounds
    for world in (1, 2, 3, 8):                      # bounds tile [0, n) exactly, sizes differ by at most 1
        b = [shard_bounds(n, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    out = str(tmp_path / "full.pt")
    mp.spawn(_worker_predict, args=(2, _free_port(), n, out), nprocs=2, join=True)
    full = torch.load(out)
    idx = torch.arange(0, n, dtype=torch.float64)
    exp = torch.stack([idx, idx * 2 + 1, torch.sin(idx), idx ** 2], 1) if n else torch.empty(0, 4, dtype=torch.float64)
    assert full.shape == exp.shape and torch.equal(full, exp)


def _worker_steps(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mural_b200.training import _steps_this_epoch

    class G:
        device = torch.device("cpu")

    class DS:
        genome = G()
        batch_sizes = np.array([100, 3, 250, 1, 64, 64, 7])
    mine, steps = _steps_this_epoch(DS(), np.arange(7), 64, world, rank)
    import json
    with open(out % rank, "w") as f:
        json.dump([[int(v) for v in mine], int(steps)], f)
    dist.destroy_process_group()


def test_data_parallel_step_counts_equal_world2(tmp_path):
    """ranks that draw different numbers of batches would deadlock in the gradient all-reduce; the per-epoch step
    count is the minimum over ranks (segments dealt round-robin)."""
    out = str(tmp_path / "steps%d.json")
    mp.spawn(_worker_steps, args=(2, _free_port(), out), nprocs=2, join=True)
    import json
    (m0, s0), (m1, s1) = json.load(open(out % 0)), json.load(open(out % 1))
    assert m0 == [0, 2, 4, 6] and m1 == [1, 3, 5]
    # rank 0: 421 sites -> 6 full + tail 37; rank 1: 68 sites -> 1 full + tail 4  => both run 2 steps
    assert s0 == s1 == 2
