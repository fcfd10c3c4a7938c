"""Container check (no GPU): the indel training tape, compiled for the host (libindel_emu.so), against fp64 autograd of the
oracle UNet_Small in train mode — output, CE(sum) loss, every parameter gradient, running statistics."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import network_t as NT

lib = C.CDLL(os.path.join(HERE, "libindel_emu.so"))


def run(tag, R, B, seed=0):
    z = np.load(os.path.join(ROOT, "tests", "golden", "indel_%s.npz" % tag))
    state = {k[2:]: np.asarray(z[k]) for k in z.files if k.startswith("w:") and "num_batches" not in k}
    down = [int(v) for v in z["down"]]
    use_rev = bool(z["use_reverse"])
    Cc = state["uplblocks.0.0.weight"].shape[0]
    ks = state["uplblocks.0.0.weight"].shape[2]
    NC = state["out_fc.2.weight"].shape[0]
    rng = np.random.default_rng(seed)
    L = 2 * R
    idx = rng.integers(0, 4, (B, L))
    x = np.zeros((B, 4, L), np.float32)
    for b in range(B):
        x[b, idx[b], np.arange(L)] = 1
    x[:, :, rng.integers(0, L, 5)] = 0.25                       # a few N columns
    labels = rng.integers(0, NC, B).astype(np.int32)
    names = list(state)
    offs, o = [], 0
    for k in names:
        offs.append(o); o += state[k].size
    blob = np.concatenate([state[k].reshape(-1) for k in names]).astype(np.float32)
    blob0 = blob.copy()
    out = np.zeros((B, NC), np.float32)
    grads = np.zeros_like(blob)
    loss = C.c_double(0)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    n_launch = lib.indel_train_emu_step(R, Cc, ks, NC, (C.c_int * 6)(*down), int(use_rev), len(names), arr,
                                        (C.c_int64 * len(names))(*offs), blob.ctypes.data_as(C.c_void_p), C.c_int64(blob.size),
                                        x.ctypes.data_as(C.c_void_p), labels.ctypes.data_as(C.c_void_p), C.c_int64(B), C.c_float(1.0),
                                        out.ctypes.data_as(C.c_void_p), grads.ctypes.data_as(C.c_void_p), C.byref(loss))
    # oracle
    sd = {k: torch.tensor(v, dtype=torch.float64, requires_grad=("running" not in k)) for k, v in state.items()}
    rec = NT._BNStats()
    ref = NT.unet_small_forward(sd, x, down, use_rev, torch.float64, train=True, rec=rec)
    l_ref = NT.ce_sum(ref, labels.astype(np.int64))
    l_ref.backward()
    d_out = np.abs(out - ref.detach().numpy()).max()
    worst, worst_name = 0.0, ""
    gmax = max(float(sd[k].grad.abs().max()) for k in names if "running" not in k)
    for k, off in zip(names, offs):
        if "running" in k:
            continue
        g_ref = sd[k].grad.numpy().reshape(-1)
        g = grads[off:off + g_ref.size]
        # relative to the tensor's own gradient scale, floored at 1e-4 of the largest gradient of the model: conv biases in front
        # of a batch-statistic BatchNorm have a true gradient of exactly 0 and only carry fp32 summation noise
        if np.abs(g_ref).max() < 1e-9 * gmax:
            assert np.abs(g).max() < 1e-4 * gmax, (k, np.abs(g).max(), gmax)
            continue
        e = np.abs(g - g_ref).max() / max(1e-4 * gmax, np.abs(g_ref).max())
        if e > worst:
            worst, worst_name = e, k
    # running statistics: the reverse-stem BatchNorm runs twice per forward, the oracle records only the last call -> skip it
    rs = 0.0
    for bn, (mean, var_unb) in rec.stats.items():
        if bn == "conv.1":
            continue
        i = names.index(bn + ".running_mean"); j = names.index(bn + ".running_var")
        em = 0.9 * blob0[offs[i]:offs[i] + mean.numel()] + 0.1 * mean.numpy()
        ev = 0.9 * blob0[offs[j]:offs[j] + mean.numel()] + 0.1 * var_unb.numpy()
        rs = max(rs, np.abs(blob[offs[i]:offs[i] + mean.numel()] - em).max() / max(1, np.abs(em).max()),
                 np.abs(blob[offs[j]:offs[j] + mean.numel()] - ev).max() / max(1, np.abs(ev).max()))
    print("%s R=%d B=%d use_reverse=%d: %d launches | out max|d| %.2e | loss %.6f vs %.6f | worst grad rel err %.2e (%s) | running stats %.1e"
          % (tag, R, B, use_rev, n_launch, d_out, loss.value, float(l_ref.detach()), worst, worst_name, rs))
    assert d_out < 2e-4 and abs(loss.value - float(l_ref)) < 1e-3 * max(1, abs(float(l_ref))) and worst < 5e-3 and rs < 1e-4


if __name__ == "__main__":
    run("hs_ins", 500, 5)
    run("hs_del_start", 500, 4, seed=1)
    run("ex_indel9", 1000, 3, seed=2)
