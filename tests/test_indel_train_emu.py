"""CPU: the MuRaL-indel training tape (mural_b200/csrc/indel_train_{core,engine}.cuh) compiled for the host with g++
(tests/emu/indel_train_emu.cpp: every kernel becomes a serial loop over its work items) against fp64 autograd of the oracle
UNet_Small in train mode — output, CE(sum) loss, every parameter gradient, running statistics.  The same functors are what
nvcc compiles into the CUDA kernels; tests/test_gpu_indel_train.py repeats the comparison on the device."""
import ctypes as C
import os
import sys

import subprocess

import numpy as np
import pytest
import torch

from oracle import network_t as NT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libindel_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-DINDEL_EMU", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "emu", "indel_train_emu.cpp")], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("tag,R,B,seed", [("hs_ins", 500, 5, 0), ("hs_del_start", 500, 4, 1), ("ex_indel9", 1000, 3, 2)])
def test_indel_training_tape_matches_autograd(lib, tag, R, B, seed):
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




@pytest.mark.parametrize("tag", ["hs_ins", "hs_del_start"])
def test_indel_training_tape_matches_reference_gradients(lib, tag):
    """The same tape against the gradients of the UNMODIFIED reference UNet_Small in train mode (tests/golden/train_kat.npz,
    float64 reference run): closes tape == reference without the oracle in between."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "train_kat.npz"))
    w = np.load(os.path.join(ROOT, "tests", "golden", "indel_%s.npz" % tag))
    state = {k[2:]: np.asarray(w[k]) for k in w.files if k.startswith("w:") and "num_batches" not in k}
    down, use_rev = [int(v) for v in w["down"]], bool(w["use_reverse"])
    x = np.ascontiguousarray(z["indel_%s:x" % tag].astype(np.float32))
    labels = z["indel_%s:y" % tag].astype(np.int32)
    B, _, L = x.shape
    names = list(state)
    offs, o = [], 0
    for k in names:
        offs.append(o); o += state[k].size
    blob = np.concatenate([state[k].reshape(-1) for k in names]).astype(np.float32)
    NC = state["out_fc.2.weight"].shape[0]
    out = np.zeros((B, NC), np.float32)
    grads = np.zeros_like(blob)
    loss = C.c_double(0)
    arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
    lib.indel_train_emu_step(L // 2, state["uplblocks.0.0.weight"].shape[0], state["uplblocks.0.0.weight"].shape[2], NC,
                             (C.c_int * 6)(*down), int(use_rev), len(names), arr, (C.c_int64 * len(names))(*offs),
                             blob.ctypes.data_as(C.c_void_p), C.c_int64(blob.size), x.ctypes.data_as(C.c_void_p),
                             labels.ctypes.data_as(C.c_void_p), C.c_int64(B), C.c_float(1.0), out.ctypes.data_as(C.c_void_p),
                             grads.ctypes.data_as(C.c_void_p), C.byref(loss))
    assert np.abs(out - z["indel_%s:out" % tag]).max() < 2e-4
    assert abs(loss.value - float(z["indel_%s:loss" % tag])) < 1e-4 * max(1.0, abs(float(z["indel_%s:loss" % tag])))
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("indel_%s:g:" % tag))
    checked = 0
    for k, off in zip(names, offs):
        key = "indel_%s:g:%s" % (tag, k)
        if key not in z.files:
            continue
        g_ref = z[key].reshape(-1)
        g = grads[off:off + g_ref.size]
        if np.abs(g_ref).max() < 1e-9 * gmax:
            assert np.abs(g).max() < 1e-4 * gmax, k
            continue
        assert np.abs(g - g_ref).max() / max(1e-4 * gmax, np.abs(g_ref).max()) < 5e-3, k
        checked += 1
    assert checked > 80
