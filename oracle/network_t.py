"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference networks' arithmetic.

The reference networks are sequences of torch library ops (Conv1d, BatchNorm1d, MaxPool1d,
Embedding, Linear, softmax; MuRaL/model/model_snv.py:290-525, model_indel.py:6-176), so the
floating-point oracle is written with torch.nn.functional on CPU tensors (fp32 by default, fp64 on
request) directly from a plain {name: array} state dict — no nn.Module, no reference import.
It is pinned against the real reference classes by oracle/make_golden.py (max |diff| recorded in
tests/golden/MANIFEST.json) and re-checked against the committed logits in tests/test_oracle_golden.py.

Nothing under mural_b200/ may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm1d default


def _t(sd, k, dtype):
    v = sd[k]
    return (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))).to(dtype)


class _BNStats:
    """Collects train-mode batch statistics so tests can check running-stat updates."""
    def __init__(self):
        self.stats = {}


def _bn(x, sd, p, dtype, train=False, rec=None):
    """nn.BatchNorm1d: eval -> running stats; train -> biased batch var (torch semantics)."""
    w, b = _t(sd, p + ".weight", dtype), _t(sd, p + ".bias", dtype)
    if not train:
        m, v = _t(sd, p + ".running_mean", dtype), _t(sd, p + ".running_var", dtype)
    else:
        dims = (0,) if x.dim() == 2 else (0, 2)
        m = x.mean(dims)
        v = x.var(dims, unbiased=False)
        if rec is not None:
            n = x.numel() // x.shape[1]
            rec.stats[p] = (m.detach().clone(), (v * n / max(n - 1, 1)).detach().clone())
    sh = (1, -1) if x.dim() == 2 else (1, -1, 1)
    return (x - m.view(sh)) / torch.sqrt(v.view(sh) + EPS) * w.view(sh) + b.view(sh)


def _conv(x, sd, p, dtype, stride=1, bias=True):
    w = _t(sd, p + ".weight", dtype)
    b = _t(sd, p + ".bias", dtype) if bias and (p + ".bias") in sd else None
    return F.conv1d(x, w, b, stride=stride, padding=(w.shape[2] - 1) // 2)


def _resblock(x, sd, p, dtype, train, rec):
    """ResBlock (model_snv.py:794-812): x + conv2(bn2(relu(conv1(bn1(relu(x))))))."""
    o = _conv(_bn(F.relu(x), sd, p + ".bn1", dtype, train, rec), sd, p + ".conv1", dtype)
    o = _conv(_bn(F.relu(o), sd, p + ".bn2", dtype, train, rec), sd, p + ".conv2", dtype)
    return x + o


def _branch(x, sd, sfx, pools, dtype, train, rec, taps=None):
    """One CNN branch of Network2 (model_snv.py:474-489 / 496-511).  sfx '' or '_2'."""
    o = _conv(_bn(x, sd, "conv1%s.0" % sfx, dtype, train, rec), sd, "conv1%s.1" % sfx, dtype)
    if taps is not None: taps["conv1" + sfx] = o
    j = o = F.max_pool1d(o, pools[0][0], pools[0][1], pools[0][2])
    if taps is not None: taps["pool1" + sfx] = o
    for i in range(2):
        o = _resblock(o, sd, "RBs1%s.%d" % (sfx, i), dtype, train, rec)
    o = o + j
    if taps is not None: taps["rb1" + sfx] = o
    o = F.max_pool1d(o, pools[1][0], pools[1][1], pools[1][2])
    j = o = _conv(_bn(o, sd, "conv2%s.0" % sfx, dtype, train, rec), sd, "conv2%s.1" % sfx, dtype)
    if taps is not None: taps["conv2" + sfx] = o
    for i in range(2):
        o = _resblock(o, sd, "RBs2%s.%d" % (sfx, i), dtype, train, rec)
    o = o + j
    if taps is not None: taps["rb2" + sfx] = o
    o = F.max_pool1d(o, pools[2][0], pools[2][1], pools[2][2])
    o = F.relu(_conv(_bn(o, sd, "conv3%s.0" % sfx, dtype, train, rec), sd, "conv3%s.1" % sfx, dtype))
    o = o.max(dim=2)[0]
    if taps is not None: taps["gmax" + sfx] = o
    fc = "distal_fc1" if sfx == "" else "distal_fc2"
    o = _bn(o, sd, fc + ".0", dtype, train, rec)
    return F.linear(o, _t(sd, fc + ".2.weight", dtype), _t(sd, fc + ".2.bias", dtype))


def network2_forward(sd, cat_x, distal_x, dtype=torch.float32, train=False, rec=None, taps=None, cont_x=None):
    """Network2.forward (model_snv.py:439-525), dropout disabled (eval, or train with p=0).
    sd: state dict; cat_x int64 [B,n_cat]; distal_x float [B,4,L]; cont_x float [B,n_cont] for models with continuous
    features (first_bn_layer, :457-463).  Returns log-probs [B,n_class]."""
    cat_x = torch.as_tensor(cat_x).long()
    x = torch.as_tensor(distal_x).to(dtype)
    emb = _t(sd, "emb_layer.weight", dtype)
    lo = emb[cat_x].reshape(cat_x.shape[0], -1)                    # :452-454
    fb = sd.get("first_bn_layer.weight")
    if fb is not None and (fb.numel() if torch.is_tensor(fb) else np.asarray(fb).size) > 0:    # :457-463
        lo = torch.cat([lo, _bn(torch.as_tensor(cont_x).to(dtype), sd, "first_bn_layer", dtype, train, rec)], dim=1)
    i = 0
    while ("lin_layers.%d.weight" % i) in sd:                       # :465-468
        lo = F.relu(F.linear(lo, _t(sd, "lin_layers.%d.weight" % i, dtype), _t(sd, "lin_layers.%d.bias" % i, dtype)))
        lo = _bn(lo, sd, "bn_layers.%d" % i, dtype, train, rec)
        i += 1
    lo = F.linear(lo, _t(sd, "local_fc.0.weight", dtype), _t(sd, "local_fc.0.bias", dtype))   # :492
    L = x.shape[2]
    assert L > 200                                                  # :470
    mid = x[:, :, L // 2 - 100: L // 2 + 101]                       # :473
    d1 = _branch(mid, sd, "", ((3, 3, 1), (3, 3, 1), (3, 3, 1)), dtype, train, rec, taps)
    d2 = _branch(x, sd, "_2", ((15, 15, 7), (7, 7, 3), (3, 3, 1)), dtype, train, rec, taps)
    if taps is not None:
        taps["logit_local"], taps["logit_mid"], taps["logit_large"] = lo, d1, d2
    dist = (F.softmax(d1, 1) + F.softmax(d2, 1)) / 2                # :515
    return torch.log(torch.clamp((F.softmax(lo, 1) + dist) / 2, min=1e-9))   # :516,523


def ce_sum(logp, y):
    """CrossEntropyLoss(reduction='sum') on the returned log-probs (training.py:327,425)."""
    return F.cross_entropy(logp, torch.as_tensor(y).long(), reduction="sum")


# ------------------------------------------------------------------ INDEL: UNet_Small
def _convblock(x, sd, p, dtype, train, rec):
    """ConvBlock (model_indel.py:6-19): x + BN(Conv1x1(SiLU(BN(Conv5(x)))))."""
    o = _bn(_conv(x, sd, p + ".conv.0", dtype, bias=False), sd, p + ".conv.1", dtype, train, rec)
    o = _bn(_conv(F.silu(o), sd, p + ".conv.3", dtype, bias=False), sd, p + ".conv.4", dtype, train, rec)
    return x + o


def unet_small_forward(sd, distal_x, downsize, use_reverse, dtype=torch.float32, train=False, rec=None, taps=None):
    """UNet_Small.forward (model_indel.py:151-176), dropout disabled."""
    o = torch.as_tensor(distal_x).to(dtype)
    if use_reverse:                                                 # :154-155
        f = lambda z: _bn(_conv(z, sd, "conv.0", dtype), sd, "conv.1", dtype, train, rec)
        o = f(o) + f(o.flip([1, 2])).flip([2])
    enc = []
    for i in range(6):                                              # :158-163
        lo = _bn(_conv(o, sd, "uplblocks.%d.0" % i, dtype, stride=downsize[i]), sd, "uplblocks.%d.1" % i, dtype, train, rec)
        o = _convblock(lo, sd, "upblocks.%d.0" % i, dtype, train, rec)
        enc.append(o)
        if taps is not None: taps["enc%d" % i] = o
    for i in range(5):                                              # :165-170
        o = F.interpolate(o, scale_factor=float(downsize[5 - i]), mode="nearest")
        lo = _bn(_conv(o, sd, "downlblocks.%d.1" % i, dtype), sd, "downlblocks.%d.2" % i, dtype, train, rec)
        o = _convblock(lo, sd, "downblocks.%d.0" % i, dtype, train, rec)
        o = enc[4 - i] + o
        if taps is not None: taps["dec%d" % i] = o
    o = F.relu(_bn(_conv(o, sd, "out_conv.0", dtype), sd, "out_conv.1", dtype, train, rec))   # :172
    o = F.softplus(_conv(o, sd, "out_conv.3", dtype))
    o = o.max(dim=2)[0]                                             # :173
    o = _bn(o, sd, "out_fc.0", dtype, train, rec)                   # :174
    return F.softplus(F.linear(o, _t(sd, "out_fc.2.weight", dtype), _t(sd, "out_fc.2.bias", dtype)))


# ------------------------------------------------------------------ calibration (host epilogue)
def dirichlet_apply(weights, prob_f32):
    """FullDirichletCalibrator.predict_proba (dirichlet_python/dirichletcal/calib/fulldirichlet.py:78-80,
    calib/multinomial.py:60-64,235-244, utils.py:5-7): softmax([log clip(p,tiny,1-tiny), 1] @ W.T).
    The clip/log run in the dtype of the input (float32 in run_predict.py:214-221), the product in fp64."""
    p = np.asarray(prob_f32)
    eps = np.finfo(p.dtype).tiny
    s = np.log(np.clip(p, eps, 1 - eps))
    s1 = np.hstack((s, np.ones((len(s), 1))))
    mul = np.dot(s1, np.asarray(weights).transpose())
    sh = mul - np.max(mul, axis=1).reshape(-1, 1)
    e = np.exp(sh)
    return e / np.sum(e, axis=1).reshape(-1, 1)


def poisson_calibrate(prob):
    """poisson_calibrate (MuRaL/model/calibration.py:10-23) on an [n, k] array."""
    prob = np.asarray(prob)
    p0 = np.clip(prob[:, 0], 1e-10, 1.0)
    lam = -np.log(p0)
    out = prob.copy()
    for c in range(1, prob.shape[1]):
        out[:, c] = lam * prob[:, c] / (1 - p0)
    out[:, 0] = 1 - lam
    return out
