"""GPU: MuRaL-indel training step (train-mode UNet_Small forward, CE(sum) on the Softplus outputs, full backward, fused
optimizer) vs fp64 autograd of the oracle (MuRaL/training.py:404-452, model_indel.py:6-176)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import network_t as NT
from test_gpu_indel import _model

pytestmark = pytest.mark.gpu


def _onehot(rng, B, L):
    idx = rng.integers(0, 4, (B, L))
    x = np.zeros((B, 4, L), np.float32)
    for b in range(B):
        x[b, idx[b], np.arange(L)] = 1
    x[:, :, rng.integers(0, L, 5)] = 0.25                       # a few N columns
    return x


@pytest.mark.parametrize("tag,R,B", [("hs_ins", 500, 6), ("hs_del_start", 1000, 4)])
def test_indel_train_forward_backward_vs_autograd(tag, R, B):
    from mural_b200 import _lib
    from mural_b200.training import IndelTrainState
    z = np.load(os.path.join(GOLD, "indel_%s.npz" % tag))
    m, state = _model(z)
    down, use_rev = [int(v) for v in z["down"]], bool(z["use_reverse"])
    NC = state["out_fc.2.weight"].shape[0]
    rng = np.random.default_rng(3)
    x = _onehot(rng, B, 2 * R)
    labels = rng.integers(0, NC, B)
    st = IndelTrainState(m, R, "Adam", lr=1e-3)
    st.set_dropout(0.0)
    m.train()
    out = st.forward(torch.from_numpy(x).cuda())
    sd = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=("running" not in k)) for k, v in state.items()
          if "num_batches" not in k}
    rec = NT._BNStats()
    ref = NT.unet_small_forward(sd, x, down, use_rev, torch.float64, train=True, rec=rec)
    assert np.abs(out.cpu().numpy() - ref.detach().numpy()).max() < 2e-4
    loss = NT.ce_sum(ref, labels)
    loss.backward()
    meta = torch.from_numpy((labels << 1).astype(np.int32)).cuda()
    dout = torch.empty_like(out)
    _lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(out), _lib.ptr(meta), B, NC, _lib.ptr(st.loss_dev), _lib.ptr(dout), _lib.current_stream()))
    assert abs(float(st.loss_dev.item()) - float(loss.detach())) < 1e-3 * max(1.0, abs(float(loss.detach())))
    st.backward(dout)
    gmax = max(float(v.grad.abs().max()) for k, v in sd.items() if v.grad is not None)
    worst = 0.0
    for p, gv in zip(st.params, st.grad_views):
        name = [k for k, v in m.named_parameters() if v is p][0]
        g_ref = sd[name].grad.numpy()
        g = gv.cpu().numpy()
        if np.abs(g_ref).max() < 1e-9 * gmax:     # conv biases in front of a batch-statistic BatchNorm: the true gradient is exactly 0
            assert np.abs(g).max() < 1e-4 * gmax, name
            continue
        err = np.abs(g - g_ref).max() / max(1e-4 * gmax, np.abs(g_ref).max())
        worst = max(worst, err)
        assert err < 5e-3, (name, err)
    print(tag, "worst relative gradient error %.2e" % worst)
    # running statistics after one training forward (the reverse-stem BatchNorm is applied twice: checked through eval below)
    sdm = m.state_dict()
    for bn, (mean, var_unb) in rec.stats.items():
        if bn == "conv.1":
            continue
        exp_m = 0.9 * np.asarray(state[bn + ".running_mean"]) + 0.1 * mean.numpy()
        exp_v = 0.9 * np.asarray(state[bn + ".running_var"]) + 0.1 * var_unb.numpy()
        assert np.abs(sdm[bn + ".running_mean"].cpu().numpy() - exp_m).max() < 1e-4 * max(1, np.abs(exp_m).max()), bn
        assert np.abs(sdm[bn + ".running_var"].cpu().numpy() - exp_v).max() < 1e-3 * max(1, np.abs(exp_v).max()), bn


def test_indel_fused_step_and_dropin_loop(kat, cuda_genome):
    """Sites from the packed genome: SiteBatch path == tensor path; the reference's loop body with a torch optimizer on the
    drop-in module equals IndelTrainState.step (SGD, see test_dropin_loop_equals_fused_step)."""
    from mural_b200 import SiteBatch, pack_meta
    from mural_b200.training import IndelTrainState
    z = np.load(os.path.join(GOLD, "indel_ex_indel9.npz"))
    Rd = 500
    n = 12
    _, genome = kat
    pos = torch.from_numpy(z["start"][:n].astype(np.int32)).cuda()
    labels = (z["start"][:n] % 8).astype(np.int64)
    meta = torch.from_numpy(pack_meta(z["strand"][:n], labels, z["chrom"][:n])).cuda()
    sb = SiteBatch(pos, meta, cuda_genome)
    y = torch.from_numpy(labels).cuda()
    crit = torch.nn.CrossEntropyLoss(reduction="sum")
    ma, _ = _model(z)
    sta = IndelTrainState(ma, Rd, "SGD", lr=1e-4, weight_decay=1e-4)
    sta.set_dropout(0.0)
    ma.train()
    a = ma.forward(sb, distal_radius=Rd)
    b = ma.forward(cuda_genome.encode_onehot(pos, meta, Rd, "indel"))
    assert torch.allclose(a, b, atol=1e-6)
    ma2, _ = _model(z)
    st2 = IndelTrainState(ma2, Rd, "SGD", lr=1e-4, weight_decay=1e-4)
    st2.set_dropout(0.0)
    ma2.train()
    opt = torch.optim.SGD(ma2.parameters(), lr=1e-4, weight_decay=1e-4, momentum=0.98, nesterov=True)
    losses = []
    for _ in range(3):
        preds = ma2.forward(sb, distal_radius=Rd)
        loss = crit(preds, y)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ma2.parameters(), max_norm=10, error_if_nonfinite=False)
        opt.step()
        losses.append(float(loss.detach()))
    mb, _ = _model(z)
    stb = IndelTrainState(mb, Rd, "SGD", lr=1e-4, weight_decay=1e-4)
    stb.set_dropout(0.0)
    mb.train()
    for _ in range(3):
        stb.step(sb)
    d = (st2.blob - stb.blob).abs().max().item()
    assert d < 5e-6 * max(1.0, stb.blob.abs().max().item()), d
    assert abs(float(stb.loss_dev.item()) - sum(losses)) < 1e-3 * abs(sum(losses))
    assert np.isfinite(losses).all()
    mb.eval()
    with torch.no_grad():
        assert torch.isfinite(mb.forward(sb, distal_radius=Rd)).all()


def test_indel_train_epochs_with_validation_metrics(kat, cuda_genome):
    """train_epochs on a MuRaL-indel dataset: the native training tape per batch, then per epoch the validation half of
    training.py:454-520 with model_type 'indel' — prediction through UNet_Small, CE(sum) on the Softplus outputs, Evaluator with
    kmer_list [2, 4, 6].  Loss and k-mer correlations equal the numpy oracle's on the same predictions."""
    from oracle import encode_np as E
    from oracle import evaluation_np as EN
    from mural_b200 import PackedSiteDataset, SiteBatch, SiteTable
    from mural_b200.training import train_epochs
    z = np.load(os.path.join(GOLD, "indel_hs_ins.npz"))
    m, state = _model(z)
    _, genome = kat
    names = list(genome)
    rng = np.random.default_rng(11)
    n, R, r_loc = 900, 500, 5
    ch = rng.integers(0, len(names), n)
    st = np.array([rng.integers(R + 10, len(genome[names[c]]) - R - 10) for c in ch])
    o = np.lexsort((st, ch))
    ch, st = ch[o], st[o]
    sd = rng.integers(0, 2, n)
    lab = rng.choice(8, n, p=[.65] + [.05] * 7)
    ds = PackedSiteDataset(SiteTable(names, ch, st, st + 1, sd, lab), cuda_genome, 2000, r_loc, 1, R, "indel")
    segs = np.arange(len(ds))
    valid, train = segs[::3], np.setdiff1d(segs, segs[::3])
    hist, lines = [], []
    losses = train_epochs(m, ds, 2, 32, sampled_segments=4, lr=1e-5, segment_indices=train, valid_indices=valid, history=hist,
                          printer=lambda *a: lines.append(a))
    assert len(losses) == 2 and len(hist) == 2 and all(np.isfinite(losses))
    h = hist[-1]
    assert {"valid_loss", "kmer2", "kmer4", "kmer6", "score"} <= set(h), sorted(h)
    rows = np.concatenate([np.arange(ds.batch_offsets[i], ds.batch_offsets[i + 1]) for i in valid])
    assert h["valid_size"] == len(rows)
    m.eval()
    with torch.no_grad():
        pred = m.forward(SiteBatch(torch.from_numpy(ds.pos[rows]).cuda(), torch.from_numpy(ds.meta[rows]).cuda(), cuda_genome), distal_radius=R)
    labels = ds.label[rows].astype(np.int64)
    loss = float(torch.nn.functional.cross_entropy(pred.double(), torch.from_numpy(labels).cuda(), reduction="sum").item())
    assert abs(loss / len(rows) - h["valid_loss"]) < 1e-5 * max(1.0, abs(h["valid_loss"]))
    prob = torch.softmax(pred, 1).cpu().numpy().astype(np.float64)
    flank = np.empty((len(rows), 2 * r_loc), np.int64)
    for c in range(len(names)):
        msk = ds.chrom[rows] == cuda_genome.chrom_index[names[c]]
        if msk.any():
            flank[msk] = E.kmer_windows(E.seq_to_symbols(genome[names[c]]), ds.pos[rows][msk].astype(np.int64), ds.strand[rows][msk], r_loc, 1, "indel")
    flank = np.concatenate([flank[:, :r_loc], np.zeros((len(rows), 1), np.int64), flank[:, r_loc:]], 1)   # dummy centre column
    for k in (2, 4):
        ref = EN.freq_kmer_comp_multi(flank, labels, prob, k, 8)
        assert np.allclose(h["kmer%d" % k], ref, rtol=0, atol=2e-5, equal_nan=True), (k, h["kmer%d" % k], ref)


def test_indel_graph_step_equals_eager_step(kat, cuda_genome):
    """IndelTrainState.step replayed from CUDA graphs (forward + CE + backward | clip + optimizer, weight gradients on their side
    stream inside the capture) == the eager step, with the out_fc dropout ON (its stream position lives in device memory and is
    advanced inside the graph) and a learning-rate change between steps."""
    from mural_b200 import SiteBatch, pack_meta
    from mural_b200.training import IndelTrainState
    z = np.load(os.path.join(GOLD, "indel_ex_indel9.npz"))
    Rd, n = 500, 16
    pos = torch.from_numpy(z["start"][:2 * n].astype(np.int32)).cuda()
    labels = (z["start"][:2 * n] % 8).astype(np.int64)
    meta = torch.from_numpy(pack_meta(z["strand"][:2 * n], labels, z["chrom"][:2 * n])).cuda()
    batches = [SiteBatch(pos[:n], meta[:n], cuda_genome), SiteBatch(pos[n:], meta[n:], cuda_genome)]
    blobs, losses, used = [], [], []
    for use_graph in (False, False, True):
        m, _ = _model(z)
        st = IndelTrainState(m, Rd, "Adam", lr=1e-4, weight_decay=1e-5, seed=7, use_graph=use_graph)
        st.set_dropout(0.1, seed=7)
        m.train()
        for i in range(6):
            if i == 3:
                st.lr = 5e-5
            st.step(batches[i % 2])
        blobs.append(st.blob.clone()); losses.append(float(st.loss_dev.item())); used.append(len(st._graphs))
    assert used == [0, 0, 1]
    scale = max(1.0, blobs[0].abs().max().item())
    d_ee = (blobs[0] - blobs[1]).abs().max().item()          # two eager runs: the float atomics of the weight gradients reorder
    d_ge = (blobs[0] - blobs[2]).abs().max().item()
    print("eager vs eager %.2e, graph vs eager %.2e (scale %.2f); losses %r" % (d_ee, d_ge, scale, losses))
    assert d_ge <= max(10 * d_ee, 2e-5 * scale), (d_ee, d_ge)
    assert abs(losses[0] - losses[2]) <= 1e-4 * abs(losses[0])
