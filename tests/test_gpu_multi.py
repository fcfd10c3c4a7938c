"""GPU, >= 2 devices (skipped otherwise): NCCL paths — interval-sharded predict == single-rank predict, and
data-parallel training keeps replicas bit-identical with one flat all-reduce per step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _setup(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _load(tag):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import GOLD, load_snv_golden
    from test_gpu_snv_forward import build_model
    z, cfg, state = load_snv_golden(tag)
    k = np.load(os.path.join(GOLD, "encode_kat.npz"))
    genome = {str(n): str(s) for n, s in zip(k["genome_names"], k["genome_seqs"])}
    return z, cfg, state, genome, build_model


def _worker_predict(rank, world, port, out):
    _setup(rank, world, port)
    from mural_b200 import PackedGenome, SiteBatch, pack_meta
    from mural_b200.predict import gather_rows, shard_bounds
    z, cfg, state, genome, build_model = _load("hs_AT")
    pg = PackedGenome(genome)
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    n = len(z["start"])
    lo, hi = shard_bounds(n, world, rank)
    meta = pack_meta(z["strand"], 0 * z["strand"], z["chrom"])
    sb = SiteBatch(torch.from_numpy(z["start"][lo:hi].astype(np.int32)).cuda(), torch.from_numpy(meta[lo:hi]).cuda(), pg)
    with torch.no_grad():
        local = m.forward(None, sb)
    full = gather_rows(local, n, world, rank)                 # NCCL gather of device tensors
    if rank == 0:
        sb_all = SiteBatch(torch.from_numpy(z["start"].astype(np.int32)).cuda(), torch.from_numpy(meta).cuda(), pg)
        with torch.no_grad():
            single = m.forward(None, sb_all).cpu()
        torch.save({"full": full.cpu(), "single": single}, out)
    dist.barrier()
    dist.destroy_process_group()


def _worker_train(rank, world, port, out):
    _setup(rank, world, port)
    from mural_b200 import PackedGenome, SiteBatch, pack_meta
    from mural_b200.training import TrainState
    z, cfg, state, genome, build_model = _load("ex_ckpt6")
    pg = PackedGenome(genome)
    m = build_model(cfg, state, int(z["n_cat"]))
    st = TrainState(m, "Adam", lr=1e-3, seed=3)
    st.set_dropout(0, 0, 0)
    m.train()
    n = 64
    lo, hi = rank * n, (rank + 1) * n                       # each rank draws its own batch
    labels = (z["start"][lo:hi] % 4).astype(np.int64)
    sb = SiteBatch(torch.from_numpy(z["start"][lo:hi].astype(np.int32)).cuda(),
                   torch.from_numpy(pack_meta(z["strand"][lo:hi], labels, z["chrom"][lo:hi])).cuda(), pg)
    # gradient exchange: all-reduced buffer == sum of the per-rank gradients
    logp = st.forward(sb)
    from mural_b200 import _lib
    dlogp = torch.empty_like(logp)
    _lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(logp), _lib.ptr(sb.meta), n, 4, None, _lib.ptr(dlogp), _lib.current_stream()))
    g_local = st.backward(dlogp).clone()
    gathered = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    st.all_reduce_grads()
    ok_sum = bool((st.grads - sum(gathered)).abs().max().item() <= 1e-5 * max(1.0, st.grads.abs().max().item()))
    for _ in range(3):
        st.step(sb)
    blobs = [torch.empty_like(st.blob[:st.n_trainable]) for _ in range(world)]
    dist.all_gather(blobs, st.blob[:st.n_trainable].contiguous())
    if rank == 0:
        torch.save({"ok_sum": ok_sum, "same": bool(all(torch.equal(blobs[0], b) for b in blobs[1:]))}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_predict_equals_single_rank(tmp_path):
    out = str(tmp_path / "p.pt")
    mp.spawn(_worker_predict, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.equal(r["full"], r["single"])      # per-site results do not depend on how sites are sharded


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_training_replicas_identical(tmp_path):
    out = str(tmp_path / "t.pt")
    mp.spawn(_worker_train, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok_sum"] and r["same"]
