"""TEST INFRASTRUCTURE ONLY (oracle/): pins the TRAIN-mode arithmetic of the oracle against the unmodified reference.

Runs only in the build container (needs /root/reference).  For one MuRaL-snv and two MuRaL-indel checkpoints it puts the
reference `Network2` / `UNet_Small` (MuRaL/model/model_snv.py:290-525, model_indel.py:21-176) in train() mode in float64 with
every Dropout probability set to 0, runs one batch through the loop body of MuRaL/training.py:424-427 (forward,
CrossEntropyLoss(sum), backward) and stores inputs, output, loss, every parameter gradient and the updated BatchNorm running
statistics in tests/golden/train_kat.npz — after asserting that oracle/network_t.py in train mode reproduces all of them.
The CUDA training paths are tested against the oracle's autograd (tests/test_gpu_snv_train.py, test_gpu_indel_train.py,
test_indel_train_emu.py); this fixture closes the chain oracle == reference for train mode
(tests/test_oracle_golden.py::test_train_mode_oracle_matches_reference_goldens).

    python -m oracle.make_golden_train
"""
import os
import pickle
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import network_t as NT  # noqa: E402
from oracle import ref_import as R  # noqa: E402
from oracle.make_golden import INDEL_CKPTS, SNV_CKPTS, clean_state  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = R.REF_ROOT


def no_dropout(model):
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0


def onehot(rng, B, L):
    idx = rng.integers(0, 4, (B, L))
    x = np.zeros((B, 4, L), np.float64)
    for b in range(B):
        x[b, idx[b], np.arange(L)] = 1
    x[:, :, rng.integers(0, L, 4)] = 0.25
    return x


def compare(tag, model, sd_np, ref_out, ref_loss, oracle_out, oracle_loss, sd64, rec, out):
    assert np.abs(ref_out - oracle_out.detach().numpy()).max() < 1e-9, tag
    assert abs(ref_loss - float(oracle_loss.detach())) < 1e-9 * max(1.0, abs(ref_loss)), tag
    worst = 0.0
    for k, p in model.named_parameters():
        if ".layer." in k:
            continue
        if p.grad is None:          # declared but unused by forward(): no gradient in the reference either
            assert k not in sd64 or sd64[k].grad is None or float(sd64[k].grad.abs().max()) == 0.0, (tag, k)
            print("   ", tag, "parameter without gradient in the reference:", k)
            continue
        g_ref = p.grad.numpy()
        g = sd64[k].grad.numpy()
        worst = max(worst, float(np.abs(g - g_ref).max() / max(1e-12, np.abs(g_ref).max(), 1e-6 * 1.0)))
        out["%s:g:%s" % (tag, k)] = g_ref
    new = model.state_dict()
    for bn, (mean, var_unb) in rec.stats.items():
        if tag.startswith("indel") and bn == "conv.1":   # applied twice per forward; the oracle records its last call only
            continue
        em = 0.9 * sd_np[bn + ".running_mean"] + 0.1 * mean.numpy()
        ev = 0.9 * sd_np[bn + ".running_var"] + 0.1 * var_unb.numpy()
        assert np.abs(new[bn + ".running_mean"].numpy() - em).max() < 1e-9 * max(1, np.abs(em).max()), (tag, bn)
        assert np.abs(new[bn + ".running_var"].numpy() - ev).max() < 1e-9 * max(1, np.abs(ev).max()), (tag, bn)
        out["%s:rm:%s" % (tag, bn)] = new[bn + ".running_mean"].numpy()
        out["%s:rv:%s" % (tag, bn)] = new[bn + ".running_var"].numpy()
    print("%-22s reference(train, fp64) == oracle: out, loss, %d gradients (worst rel %.1e), %d BatchNorm updates"
          % (tag, sum(1 for k, _ in model.named_parameters() if ".layer." not in k), worst, len(rec.stats)))
    assert worst < 1e-7, (tag, worst)


def main():
    assert R.available(), "needs /root/reference"
    pre, snv, indel, nnu = R.import_reference()
    crit = torch.nn.CrossEntropyLoss(reduction="sum")
    rng = np.random.default_rng(31)
    out = {}
    # ---- MuRaL-snv: example checkpoint (R_d = 200)
    tag, rel = "snv_ex_ckpt6", SNV_CKPTS["ex_ckpt6"]
    cfg = pickle.load(open(os.path.join(REF, rel, "model.config.pkl"), "rb"))
    common = dict(emb_dims=cfg["emb_dims"], n_cont=0, n_class=cfg["n_class"], distal_order=1, in_channels=4)
    model = nnu.model_choice(cfg["model_no"], cfg, common, "snv")
    sd = torch.load(os.path.join(REF, rel, "model"), map_location="cpu")
    model.load_state_dict(sd)
    model.double().train()
    no_dropout(model)
    B, L, n_cat = 24, 2 * cfg["distal_radius"] + 1, len(cfg["emb_dims"])
    cat = rng.integers(0, 65, (B, n_cat))
    x = onehot(rng, B, L)
    y = rng.integers(0, cfg["n_class"], B)
    sd_np = {k: v.astype(np.float64) if v.dtype.kind == "f" else v for k, v in clean_state(sd).items()}
    preds = model.forward((torch.zeros(B, 1, dtype=torch.float64), torch.from_numpy(cat)), torch.from_numpy(x))
    loss = crit(preds, torch.from_numpy(y))
    loss.backward()
    sd64 = {k: torch.tensor(v, dtype=torch.float64, requires_grad=("running" not in k)) for k, v in sd_np.items() if "num_batches" not in k}
    rec = NT._BNStats()
    o = NT.network2_forward(sd64, cat, x, torch.float64, train=True, rec=rec)
    lo = NT.ce_sum(o, y)
    lo.backward()
    out.update({tag + ":cat": cat.astype(np.int16), tag + ":x": x.astype(np.float32), tag + ":y": y.astype(np.int8),
                tag + ":out": preds.detach().numpy(), tag + ":loss": np.array(float(loss.detach()))})
    compare(tag, model, sd_np, preds.detach().numpy(), float(loss.detach()), o, lo, sd64, rec, out)
    # ---- MuRaL-indel: with and without the reverse-strand stem, R = 500
    for key in ("hs_ins", "hs_del_start"):
        tag, rel = "indel_" + key, INDEL_CKPTS[key]
        cfg = pickle.load(open(os.path.join(REF, rel, "model.config.pkl"), "rb"))
        sd = torch.load(os.path.join(REF, rel, "model"), map_location="cpu")
        use_rev = any(k.startswith("conv.0") for k in sd)
        n_class, ch8, ks = sd["out_fc.2.weight"].shape[0], sd["uplblocks.0.0.weight"].shape[0], sd["uplblocks.0.0.weight"].shape[2]
        down = cfg["down_list"]
        model = indel.UNet_Small(n_class, ch8, ks, down, use_reverse=use_rev)
        model.load_state_dict(sd)
        model.double().train()
        no_dropout(model)
        B, L = 5, 1000
        x = onehot(rng, B, L)
        y = rng.integers(0, n_class, B)
        sd_np = {k: v.astype(np.float64) if v.dtype.kind == "f" else v for k, v in clean_state(sd).items()}
        preds = model.forward(torch.from_numpy(x))
        loss = crit(preds, torch.from_numpy(y))
        loss.backward()
        sd64 = {k: torch.tensor(v, dtype=torch.float64, requires_grad=("running" not in k)) for k, v in sd_np.items() if "num_batches" not in k}
        rec = NT._BNStats()
        o = NT.unet_small_forward(sd64, x, down, use_rev, torch.float64, train=True, rec=rec)
        lo = NT.ce_sum(o, y)
        lo.backward()
        out.update({tag + ":x": x.astype(np.float32), tag + ":y": y.astype(np.int8), tag + ":out": preds.detach().numpy(),
                    tag + ":loss": np.array(float(loss.detach()))})
        compare(tag, model, sd_np, preds.detach().numpy(), float(loss.detach()), o, lo, sd64, rec, out)
    path = os.path.join(GOLD, "train_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
