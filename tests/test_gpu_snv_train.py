"""GPU: Network2 training step (train-mode forward, backward, clip + optimizer) vs torch autograd on the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_snv_golden
from oracle import encode_np as E
from oracle import network_t as NT
from test_gpu_snv_forward import build_model

pytestmark = pytest.mark.gpu


def _oracle_inputs(z, cfg, genome, n):
    names = list(genome)
    ch, st, sd = z["chrom"][:n], z["start"][:n], z["strand"][:n]
    cat = np.empty((n, int(z["n_cat"])), np.int64)
    oh = np.empty((n, 4, 2 * cfg["distal_radius"] + 1), np.float32)
    for c in range(len(names)):
        m = ch == c
        if m.any():
            sym = E.seq_to_symbols(genome[names[c]])
            cat[m] = E.kmer_windows(sym, st[m], sd[m], cfg["local_radius"], cfg["local_order"])
            oh[m] = E.onehot_windows(sym, st[m], sd[m], cfg["distal_radius"])
    return cat, oh


def _check_grads(layout, g, sd64, what):
    """Gradient gate vs fp64 autograd of the oracle: |g - ref| <= 5e-3 * max(1e-3, max|ref|) per tensor.

    ReLU kinks: with ~2e6 pre-activations per step, one of them lands within fp32 rounding of zero every few runs; the fp64
    reference and an fp32 implementation then take different sides of the kink, ONE channel of that layer's bias / weight
    gradient moves by that element's gradient (measured: a single channel off by 13 % of its value while the other 31 agree
    to 1e-7) and the tensors upstream of it in the same branch move by up to ~1e-2 of their scale.  Such an event is not an
    arithmetic error (torch's own fp32 and fp64 autograd differ the same way), so the gate is: at least 90 % of the tensors
    within 5e-3, every tensor within 2e-2, and the whole gradient vector within 5e-3 in relative L2.  A wrong tap, mask or
    scale moves whole tensors by O(1) and fails all three; a precision loss in the conv kernels is caught at 2e-6 by
    tests/test_gpu_conv_mma.py."""
    errs, num, den = [], 0.0, 0.0
    for name, off, cnt, is_buf in layout:
        if is_buf:
            continue
        ref_g = sd64[name].grad.numpy().reshape(-1)
        got = g[off:off + cnt]
        errs.append((float(np.abs(got - ref_g).max() / max(1e-3, np.abs(ref_g).max())), name))
        num += float(((got - ref_g) ** 2).sum()); den += float((ref_g ** 2).sum())
    rel = (num / max(den, 1e-300)) ** 0.5
    errs.sort(reverse=True)
    n_bad = sum(e >= 5e-3 for e, _ in errs)
    assert errs[0][0] < 2e-2, (what, errs[:4])
    assert n_bad <= len(errs) // 10, (what, n_bad, errs[:8])
    assert rel < 5e-3, (what, rel)
    return errs[n_bad][0] if n_bad < len(errs) else errs[-1][0], rel


def _batch(z, genome_dev, n, labels):
    from mural_b200 import SiteBatch, pack_meta
    pos = torch.from_numpy(z["start"][:n].astype(np.int32)).cuda()
    meta = torch.from_numpy(pack_meta(z["strand"][:n], labels, z["chrom"][:n])).cuda()
    return SiteBatch(pos, meta, genome_dev)


@pytest.mark.parametrize("tag", ["ex_ckpt6", "hs_AT"])
def test_train_forward_backward_vs_autograd(kat, cuda_genome, tag):
    from mural_b200 import _lib
    from mural_b200.training import TrainState
    z, cfg, state = load_snv_golden(tag)
    _, genome = kat
    n = 48
    labels = (z["start"][:n] % 4).astype(np.int64)
    m = build_model(cfg, state, int(z["n_cat"]))
    st = TrainState(m, "Adam", lr=1e-3)
    st.set_dropout(0, 0, 0)
    m.train()
    sb = _batch(z, cuda_genome, n, labels)
    logp = st.forward(sb)
    # ---- oracle: fp64 autograd through the restated network in train mode (batch-stat BN, dropout off)
    cat, oh = _oracle_inputs(z, cfg, genome, n)
    sd64 = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=("running" not in k))
            for k, v in state.items() if "num_batches" not in k}
    rec = NT._BNStats()
    ref = NT.network2_forward(sd64, cat, oh, torch.float64, train=True, rec=rec)
    assert np.abs(logp.cpu().numpy() - ref.detach().numpy()).max() < 2e-4
    loss = NT.ce_sum(ref, labels)
    loss.backward()
    # fused loss + gradient of the loss w.r.t. log-probs, then backward
    dlogp = torch.empty_like(logp)
    _lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(logp), _lib.ptr(sb.meta), n, 4, _lib.ptr(st.loss_dev), _lib.ptr(dlogp),
                                            _lib.current_stream()))
    assert abs(float(st.loss_dev.item()) - float(loss.detach())) < 1e-3 * max(1.0, abs(float(loss.detach())))
    g = st.backward(dlogp).cpu().numpy()
    worst, rel = _check_grads(m.native_layout(), g, sd64, tag)
    print(tag, "worst relative gradient error %.2e (tensors without a ReLU-kink event), whole-gradient relative L2 %.2e" % (worst, rel))
    # running statistics after one training forward (momentum 0.1, unbiased variance)
    sdm = m.state_dict()
    for bn, (mean, var_unb) in rec.stats.items():
        exp_m = 0.9 * np.asarray(state[bn + ".running_mean"]) + 0.1 * mean.numpy()
        exp_v = 0.9 * np.asarray(state[bn + ".running_var"]) + 0.1 * var_unb.numpy()
        assert np.abs(sdm[bn + ".running_mean"].cpu().numpy() - exp_m).max() < 1e-4 * max(1, np.abs(exp_m).max()), bn
        assert np.abs(sdm[bn + ".running_var"].cpu().numpy() - exp_v).max() < 1e-3 * max(1, np.abs(exp_v).max()), bn
    assert int(sdm["conv1.0.num_batches_tracked"]) == int(state["conv1.0.num_batches_tracked"]) + 1


@pytest.mark.parametrize("R_d", [100, 5000])
def test_transfer_window_sweep_train_step(kat, cuda_genome, R_d):
    """a checkpoint trained at R_d=1000 fine-tuned at another distal radius (the network is fully
    convolutional; run_train_TL_raytune.py:138-170 copies every other hyper-parameter from the pretrained config).
    Train-mode forward and all gradients at L=201 and L=10001 vs fp64 autograd on the oracle."""
    from mural_b200 import _lib
    from mural_b200.training import TrainState, load_pretrained
    z, cfg, state = load_snv_golden("hs_AT")
    cfg = dict(cfg, distal_radius=R_d)
    _, genome = kat
    n = 12
    labels = (z["start"][:n] % 4).astype(np.int64)
    pretrained = {k: v.cpu().clone() for k, v in build_model(dict(cfg, distal_radius=1000), state, int(z["n_cat"])).state_dict().items()}
    from mural_b200 import model_choice, weights_init
    common = dict(emb_dims=[(65, 2)] * int(z["n_cat"]), n_cont=0, n_class=cfg["n_class"], distal_order=1, in_channels=4)
    m = model_choice(2, cfg, common, "snv").to("cuda")
    m.apply(weights_init)
    load_pretrained(m, pretrained, train_all=True, init_fc_with_pretrained=True)
    assert all(p.requires_grad for p in m.parameters())
    st = TrainState(m, "Adam", lr=1e-3)
    st.set_dropout(0, 0, 0)
    m.train()
    sb = _batch(z, cuda_genome, n, labels)
    logp = st.forward(sb)
    cat, oh = _oracle_inputs(z, cfg, genome, n)
    sd64 = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=("running" not in k))
            for k, v in state.items() if "num_batches" not in k}
    ref = NT.network2_forward(sd64, cat, oh, torch.float64, train=True)
    assert np.abs(logp.cpu().numpy() - ref.detach().numpy()).max() < 2e-4
    NT.ce_sum(ref, labels).backward()
    dlogp = torch.empty_like(logp)
    _lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(logp), _lib.ptr(sb.meta), n, 4, _lib.ptr(st.loss_dev), _lib.ptr(dlogp),
                                            _lib.current_stream()))
    g = st.backward(dlogp).cpu().numpy()
    _check_grads(m.native_layout(), g, sd64, R_d)


@pytest.mark.parametrize("optim", ["Adam", "AdamW", "SGD"])
def test_fused_optimizer_matches_torch(optim):
    """clip_grad_norm_(10) + optimizer.step() over the flat buffer == torch.optim on the same gradients, 3 steps."""
    from mural_b200 import _lib
    torch.manual_seed(1)
    n = 10007
    p0 = torch.randn(n, device="cuda")
    ref_p = torch.nn.Parameter(p0.clone())
    if optim == "Adam":
        opt, kind = torch.optim.Adam([ref_p], lr=1e-2, weight_decay=1e-3), 0
    elif optim == "AdamW":
        opt, kind = torch.optim.AdamW([ref_p], lr=1e-2, weight_decay=1e-2, amsgrad=True), 1
    else:
        opt, kind = torch.optim.SGD([ref_p], lr=1e-3, weight_decay=1e-3, momentum=0.98, nesterov=True), 2
    wd = opt.param_groups[0]["weight_decay"]
    lr = opt.param_groups[0]["lr"]
    p = p0.clone()
    m = torch.zeros_like(p); v = torch.zeros_like(p); vm = torch.zeros_like(p)
    scratch = torch.zeros(2, dtype=torch.float64, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * (3.0 if step == 2 else 0.05)     # step 2 exceeds max_norm -> clipped
        ref_p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], max_norm=10, error_if_nonfinite=False)
        opt.step()
        _lib.check(_lib.lib().mural_optimizer_step(kind, _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.ptr(vm), n, lr, wd,
                                                   step, 10.0, 1.0, _lib.ptr(scratch), _lib.current_stream()))
        assert (p - ref_p.detach()).abs().max().item() < 2e-6 * max(1.0, ref_p.detach().abs().max().item()), (optim, step)


def test_dropin_loop_equals_fused_step(kat, cuda_genome):
    """The reference's loop body with torch's optimizer (training.py:424-436) == TrainState.step().
    SGD is used for the comparison because it is linear in the gradient: several biases sit in front of a
    batch-statistic BatchNorm, their true gradient is exactly zero, and Adam turns the ~1e-7 summation-order noise of
    such gradients into O(lr) updates in *any* implementation."""
    from mural_b200.training import TrainState
    z, cfg, state = load_snv_golden("ex_ckpt6")
    n = 64
    labels = (z["start"][:n] % 4).astype(np.int64)
    sb = _batch(z, cuda_genome, n, labels)
    y = torch.from_numpy(labels).cuda()
    crit = torch.nn.CrossEntropyLoss(reduction="sum")
    # (a) drop-in
    ma = build_model(cfg, state, int(z["n_cat"]))
    sta = TrainState(ma, "SGD", lr=1e-4, weight_decay=1e-4)
    sta.set_dropout(0, 0, 0)
    ma.train()
    opt = torch.optim.SGD(ma.parameters(), lr=1e-4, weight_decay=1e-4, momentum=0.98, nesterov=True)
    for _ in range(2):
        preds = ma.forward(None, sb)
        loss = crit(preds, y)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(ma.parameters(), max_norm=10, error_if_nonfinite=False)
        opt.step()
    # (b) fused
    mb = build_model(cfg, state, int(z["n_cat"]))
    stb = TrainState(mb, "SGD", lr=1e-4, weight_decay=1e-4)
    stb.set_dropout(0, 0, 0)
    mb.train()
    for _ in range(2):
        stb.step(sb)
    d = (sta.blob - stb.blob).abs().max().item()
    assert d < 2e-6 * max(1.0, sta.blob.abs().max().item()), d
    moved = (stb.blob[:stb.n_trainable] - torch.from_numpy(mb.flat_blob()[:stb.n_trainable]).cuda()).abs().max().item()
    assert moved == 0.0                      # model parameters ARE the flat buffer (views), nothing to copy back
    # eval after training re-folds the updated weights; the loss on the training batch went down
    mb.eval()
    with torch.no_grad():
        after = float(crit(mb.forward(None, sb), y))
    m0 = build_model(cfg, state, int(z["n_cat"]))
    with torch.no_grad():
        before = float(crit(m0.forward(None, sb), y))
    assert after < before, (before, after)


def test_dropout_statistics(kat, cuda_genome):
    from mural_b200.training import TrainState
    z, cfg, state = load_snv_golden("ex_ckpt6")
    n = 128
    sb = _batch(z, cuda_genome, n, np.zeros(n, np.int64))
    m = build_model(cfg, state, int(z["n_cat"]))
    st = TrainState(m, seed=5)
    m.train()
    st.set_dropout(0.0, 0.0, 0.0, 5)
    base = st.forward(sb)
    st.set_dropout(0.1, 0.1, 0.25, 5)
    a = st.forward(sb)
    b = st.forward(sb)
    assert not torch.equal(a, b)                       # different step -> different masks
    assert (a - base).abs().max().item() > 1e-4        # dropout does something
    assert torch.isfinite(a).all()
    m2 = build_model(cfg, state, int(z["n_cat"]))       # determinism: same seed, same step counter -> same result
    st2 = TrainState(m2, seed=5)
    m2.train()
    st2.set_dropout(0.0, 0.0, 0.0, 5)
    st2.forward(sb)
    st2.set_dropout(0.1, 0.1, 0.25, 5)
    assert torch.equal(st2.forward(sb), a)


def test_graph_step_equals_eager_step(kat, cuda_genome):
    """TrainState.step replays CUDA graphs from the second batch of a size on.  With dropout ON the masks are a function of
    (seed, forward counter, layer, element): the counter lives on the device and is bumped inside the graph, so graph and
    eager runs draw the same masks step by step and stay equal — a stale counter, lr or optimizer step would show up here.
    lr changes between steps (StepLR) and reaches the captured optimizer through device memory."""
    from mural_b200.training import TrainState
    z, cfg, state = load_snv_golden("ex_ckpt6")
    n = 64
    labels = (z["start"][:n] % 4).astype(np.int64)
    sb = _batch(z, cuda_genome, n, labels)
    runs = {}
    for key, use_graph in (("eager", False), ("graph", True)):
        m = build_model(cfg, state, int(z["n_cat"]))
        st = TrainState(m, "SGD", lr=1e-4, weight_decay=1e-4, seed=5, use_graph=use_graph)
        m.train()
        traj = []
        for i in range(5):
            st.lr = 1e-4 * (0.5 ** (i // 2))
            traj.append(st.step(sb).clone())
        assert bool(st._graphs) == use_graph
        assert int(st._opt_step_dev.item()) == 5 == st.opt_step and st.n_forward == 5
        runs[key] = (traj, st.blob.clone(), float(st.loss_dev.item()))
    for a, b in zip(runs["eager"][0], runs["graph"][0]):
        assert (a - b).abs().max().item() < 1e-4
    assert not torch.equal(runs["graph"][0][2], runs["graph"][0][3])
    d = (runs["eager"][1] - runs["graph"][1]).abs().max().item()
    assert d < 5e-6 * max(1.0, runs["eager"][1].abs().max().item()), d
    assert abs(runs["eager"][2] - runs["graph"][2]) < 1e-3 * abs(runs["eager"][2])
    # a batch of another size falls back to the eager path (captured on its second appearance) and keeps the counters in step
    sb2 = _batch(z, cuda_genome, 40, labels[:40])
    st.step(sb2)
    assert int(st._opt_step_dev.item()) == 6 == st.opt_step and set(st._graphs) == {64}
    st.step(sb2)
    assert set(st._graphs) == {64, 40}
    # a LARGER batch makes the library reallocate its workspace: graphs (they hold pointers into the old one) are dropped
    n3 = 96
    sb3 = _batch(z, cuda_genome, n3, (z["start"][:n3] % 4).astype(np.int64))
    st.step(sb3)
    assert not st._graphs
    for _ in range(3):
        out = st.step(sb)
    assert set(st._graphs) == {64} and torch.isfinite(out).all() and int(st._opt_step_dev.item()) == 11 == st.opt_step


def test_train_epochs_with_validation_metrics(kat, cuda_genome):
    """train_epochs with a held-out set of segments: every epoch ends with the device-side Evaluator (training.py:454-520).
    The reported validation loss and k-mer correlations equal the oracle's on the same predictions."""
    from oracle import evaluation_np as EN
    from mural_b200 import PackedSiteDataset, SiteTable, generate_site_batches, model_predict_m
    from mural_b200.training import train_epochs
    z, cfg, state = load_snv_golden("ex_ckpt6")
    _, genome = kat
    names = list(genome)
    rng = np.random.default_rng(4)
    n = 6000
    ch = rng.integers(0, len(names), n)
    st = np.array([rng.integers(300, len(genome[names[c]]) - 300) for c in ch])
    o = np.lexsort((st, ch))
    ch, st = ch[o], st[o]
    sd = rng.integers(0, 2, n)
    lab = rng.choice(4, n, p=[.7, .1, .1, .1])
    t = SiteTable(names, ch, st, st + 1, sd, lab)
    ds = PackedSiteDataset(t, cuda_genome, 2000, cfg["local_radius"], cfg["local_order"], cfg["distal_radius"])
    segs = np.arange(len(ds))
    valid, train = segs[::4], np.setdiff1d(segs, segs[::4])
    m = build_model(cfg, state, int(z["n_cat"]))
    hist, lines = [], []
    losses = train_epochs(m, ds, 2, 128, sampled_segments=4, lr=1e-4, segment_indices=train, valid_indices=valid, history=hist,
                          printer=lambda *a: lines.append(a))
    assert len(losses) == 2 and len(hist) == 2 and all(np.isfinite(losses))
    h = hist[-1]
    assert {"valid_loss", "kmer3", "kmer5", "kmer7", "score", "corr_list", "window100000", "window500000"} <= set(h)
    # same model, same held-out sites through the plain prediction loop + the numpy oracle
    pred, loss = model_predict_m(m, generate_site_batches(ds, 1 << 30, 512, segment_indices=valid), None, torch.device("cuda"), 4)
    rows = np.concatenate([np.arange(ds.batch_offsets[i], ds.batch_offsets[i + 1]) for i in valid])
    assert abs(loss / len(rows) - h["valid_loss"]) < 1e-6 * max(1.0, abs(h["valid_loss"])) and h["valid_size"] == len(rows)
    prob = torch.softmax(pred, 1).cpu().numpy().astype(np.float64)
    flank = np.empty((len(rows), 2 * cfg["local_radius"] + 1), np.int64)
    for c in range(len(names)):
        msk = ds.chrom[rows] == cuda_genome.chrom_index[names[c]]
        if msk.any():
            flank[msk] = E.kmer_windows(E.seq_to_symbols(genome[names[c]]), ds.pos[rows][msk].astype(np.int64), ds.strand[rows][msk], cfg["local_radius"], 1)
    for k in (3, 5):
        ref = EN.freq_kmer_comp_multi(flank, ds.label[rows].astype(np.int64), prob, k, 4)
        assert np.allclose(h["kmer%d" % k], ref, rtol=0, atol=2e-5, equal_nan=True), (k, h["kmer%d" % k], ref)
    assert any(isinstance(a[0], str) and a[0].startswith("Validation Loss") for a in lines)


def test_training_state_survives_module_moves_and_config(kat, cuda_genome):
    """(a) .to('cpu') / .to('cuda') (what load_pretrained does) after a TrainState exists must not detach the model's
    tensors from the state's flat buffer; (b) train_epochs honours optim / scheduler arguments even when a train-mode forward
    auto-created an Adam state; (c) out-of-range labels are refused."""
    from mural_b200 import PackedSiteDataset, SiteTable
    from mural_b200.training import OPTIMIZERS, TrainState, load_pretrained, train_epochs
    z, cfg, state = load_snv_golden("ex_ckpt6")
    _, genome = kat
    m = build_model(cfg, state, int(z["n_cat"]))
    labels = (z["start"][:64] % 4).astype(np.int64)
    sb = _batch(z, cuda_genome, 64, labels)
    m.train()
    st = TrainState(m, "Adam", lr=1e-3, use_graph=False)
    st.step(sb)
    # (a) round trip through the CPU with a state-dict load in between
    saved = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    m.to("cpu")
    with pytest.raises(RuntimeError):
        st.step(sb)
    load_pretrained(m.to("cuda"), saved)                   # cpu -> load -> cuda inside
    w = m.local_fc[0].weight
    assert w.data_ptr() >= st.blob.data_ptr() and w.data_ptr() < st.blob.data_ptr() + st.blob.numel() * 4
    before = w.detach().clone()
    st.step(sb)
    assert not torch.equal(before, m.local_fc[0].weight.detach())          # the step updates what state_dict() reads
    assert torch.equal(m.state_dict()["local_fc.0.weight"], st.blob[st.model.native_layout()[[n for n, *_ in st.model.native_layout()].index("local_fc.0.weight")][1]:][:w.numel()].view(w.shape))
    # (b) optimizer / scheduler arguments win over an auto-created state
    names = list(genome)
    rng = np.random.default_rng(1)
    n = 1200
    stt = np.sort(rng.integers(300, len(genome[names[0]]) - 300, n))
    t = SiteTable(names, np.zeros(n, int), stt, stt + 1, rng.integers(0, 2, n), rng.integers(0, 4, n))
    ds = PackedSiteDataset(t, cuda_genome, 2000, cfg["local_radius"], cfg["local_order"], cfg["distal_radius"])
    m2 = build_model(cfg, state, int(z["n_cat"]))
    m2.train()
    m2.forward(None, sb)                                                     # auto-creates an Adam state
    assert m2._train_state.kind == OPTIMIZERS["Adam"]
    for sched in ("StepLR2", "ROP"):
        losses = train_epochs(m2, ds, 2, 128, optim="SGD", lr=1e-4, lr_scheduler=sched, min_lr=1e-6, restart_lr=1e-4,
                              valid_indices=np.arange(0, len(ds), 5), printer=lambda *a: None)
        assert m2._train_state.kind == OPTIMIZERS["SGD"] and all(np.isfinite(losses))
    # (c) label validation
    bad = SiteTable(names, np.zeros(4, int), stt[:4], stt[:4] + 1, np.zeros(4, int), np.array([0, 1, 2, 9]))
    ds_bad = PackedSiteDataset(bad, cuda_genome, 2000, cfg["local_radius"], cfg["local_order"], cfg["distal_radius"])
    with pytest.raises(ValueError):
        train_epochs(m2, ds_bad, 1, 2)
