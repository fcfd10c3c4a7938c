#!/usr/bin/env python
"""bench.py — MuRaL-snv genome-wide predict (BASELINE.json configs[1]) on N B200s.

Workload: synthetic 100 Mb genome (4 x 25 Mb, iid ACGT, seed 1234, N runs at chromosome ends and 20 random
5 kb N runs per chromosome, seed 1235), sites = every A ('+') / T ('-') in genomic order,
Network2 with the shipped Homo_sapiens/SNV/AT weights (tests/golden/snv_hs_AT.npz; local ±7 bp 3-mers,
expanded ±1 kb, C=32), one *step* = the hot path (gather -> network -> log-probs) over one batch of
`--sites-per-step` consecutive sites.  Each rank owns a contiguous genomic interval (no collectives on the
data path; weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|bf16] [--impl reference]

`value`  : sites/s with the site records already in HBM (device-timed with CUDA events per step, max over ranks)
`e2e`    : sites/s through the C-ABI host entry point (pinned host buffers, H2D + D2H inside the timed region)
`roofline`: dominant kernel (the Conv1d stack) — algorithmic FLOPs per launch / mean launch duration from the
            library's own CUDA-event profile of a second, identical pass
`cpu_baseline` / `--impl reference`: the oracle port of the reference CPU path (numpy encoders + torch CPU fp32
            network, all host threads) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHROM_LEN = 25_000_000
N_CHROM = 4
FLOP_STACK = 6_352_896 + 43_112       # stage-S forward FLOPs/site at L=2001
FLOP_CONV_ONLY = 6_352_896
FLOP_CONV3 = 92_160                   # conv3 of both branches (8 + 7 rows x 6144): runs in k_tail, not in the stage kernels


# ------------------------------------------------------------------------------------------ workload
def synth_chromosome(ci):
    rng = np.random.default_rng(1234 + ci)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, CHROM_LEN, dtype=np.uint8)]
    seq = seq.copy()
    seq[:10_000] = ord("N"); seq[-10_000:] = ord("N")
    r2 = np.random.default_rng(1235 + ci)
    for s in r2.integers(20_000, CHROM_LEN - 30_000, 20):
        seq[s:s + 5_000] = ord("N")
    return seq


def rank_sites(chroms, rank, world, need):
    """First `need` A/T sites of this rank's genomic interval -> (pos int32, meta int32)."""
    from mural_b200.data import pack_meta
    N_CHROM = len(chroms)
    total = CHROM_LEN * N_CHROM
    lo = total * rank // world
    pos_l, meta_l, got = [], [], 0
    g = lo
    while got < need:
        ci, off = (g // CHROM_LEN) % N_CHROM, g % CHROM_LEN
        span = min(CHROM_LEN - off, int((need - got) * 2.2) + 100_000)
        sl = chroms[ci][off:off + span]
        isA, isT = sl == ord("A"), sl == ord("T")
        idx = np.flatnonzero(isA | isT)[: need - got]
        pos_l.append((idx + off).astype(np.int32))
        meta_l.append(pack_meta(isT[idx].astype(np.int64), np.zeros(len(idx), np.int64), np.full(len(idx), ci)))
        got += len(idx)
        g = (g + span) % total
    return np.concatenate(pos_l), np.concatenate(meta_l)


def load_weights():
    z = np.load(os.path.join(ROOT, "tests", "golden", "snv_hs_AT.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    return cfg, state, int(z["n_cat"])


def build_model(cfg, state, n_cat, mode):
    import torch
    from mural_b200 import model_choice
    common = dict(emb_dims=[(65, 2)] * n_cat, n_cont=0, n_class=cfg["n_class"], distal_order=1, in_channels=4)
    m = model_choice(2, cfg, common, "snv")
    sd = m.state_dict()
    alias = {"1": "bn1", "2": "conv1", "4": "bn2", "5": "conv2"}
    for k in sd:
        src = k
        if ".layer." in k:
            head, rest = k.split(".layer.")
            src = head + "." + alias[rest.split(".")[0]] + "." + rest.split(".", 1)[1]
        sd[k] = torch.from_numpy(np.asarray(state[src]))
    m.load_state_dict(sd, strict=True)
    m.to("cuda").eval()
    m.compute_mode = mode
    return m



# ------------------------------------------------------------------------------------------ per-leg roofline / CPU baselines
def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def leg_roofline(flops_per_site, sites_per_s, n_gpus, what):
    """Whole-leg tensor roofline: algorithmic FLOPs per site x sites/s per GPU over the measured sustained bf16
    peak.  The legs below are timed as whole steps (no per-kernel split), so this is the step's fraction, not a kernel's."""
    peaks, src = _peaks()
    peak = peaks["bf16_tflops_sustained"]
    ach = flops_per_site * sites_per_s / max(n_gpus, 1) / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "flops_per_site": flops_per_site, "what": what, "peak_source": src + " (bf16 sustained)"}


def snv_train_flops(L=2001):
    """fwd + dgrad + wgrad ~= 3 x stage-S + 2 x stage-G wgrad: 19.2 M + 3.4 M FLOP/site at L = 2001."""
    L1 = (L - 1) // 15 + 1; L2 = (L1 - 1) // 7 + 1; L3 = (L2 - 1) // 3 + 1
    large = 24576 * L1 + 30720 * L2 + 6144 * L3
    return 3 * (2402304 + large + 43112) + 2 * (768 * L + 768 * 201)


def unet_flops(C, ks, down, L, use_reverse=True):
    """Dense conv FLOPs (multiply-add = 2) of UNet_Small.forward per site (model_indel.py:151-176): 113 379 328 for the shipped
    human configuration (C=8, k=7, strides 1,4,5,5,5,2, L=8000)."""
    ch = [C * (i + 1) for i in range(6)]
    ln, x = [], L
    for s in down:
        x = (x - 1) // s + 1
        ln.append(x)
    f = 0
    for i in range(6):
        cin = ch[i - 1] if i else 4
        f += 2 * ln[i] * (ks * cin * ch[i] + 5 * ch[i] * 2 * ch[i] + 2 * ch[i] * ch[i])
    for i in range(5):
        c = ch[4 - i]
        f += 2 * ln[4 - i] * (ks * ch[5 - i] * c + 5 * c * 2 * c + 2 * c * c)
    f += 2 * ln[0] * 2 * C * C
    if use_reverse:                                          # strand-symmetric stem: the 4 -> 4 conv applied twice (:154-155)
        f += 2 * 2 * L * ks * 4 * 4
    return f


def cpu_unet(state, cfg, Rd, genome_sym, pos, train=False, labels=None, budget_s=12.0, batch=8):
    """Oracle port of the reference CPU path for MuRaL-indel (numpy one-hot windows + torch CPU fp32 UNet_Small; with train=True
    one forward + CE(sum) + backward per batch, autograd of the oracle) on as many `batch`-site batches as fit `budget_s`."""
    import torch
    from oracle import encode_np as E
    from oracle import network_t as NT
    torch.set_num_threads(os.cpu_count())
    sd = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=(train and "running" not in k and "num_batches" not in k))
          for k, v in state.items()}
    done, outs, t0 = 0, [], time.perf_counter()
    while done + batch <= len(pos) and (done == 0 or time.perf_counter() - t0 < budget_s):
        p = pos[done:done + batch]
        oh = E.onehot_windows(genome_sym, p, np.zeros(len(p), np.int64), Rd, "indel")
        if train:
            out = NT.unet_small_forward(sd, oh, cfg["down_list"], cfg["use_reverse"], torch.float32, train=True)
            NT.ce_sum(out, labels[done:done + batch]).backward()
        else:
            with torch.no_grad():
                outs.append(NT.unet_small_forward(sd, oh, cfg["down_list"], cfg["use_reverse"], torch.float32).numpy())
        done += batch
    dt = time.perf_counter() - t0
    return done / dt, dt, done, (np.concatenate(outs) if outs else None)


def cpu_snv_train(chroms, pos, meta_lab, cfg, state, batch=128, budget_s=12.0):
    """Oracle port of the reference CPU training step for MuRaL-snv: numpy encoders + train-mode Network2 (batch-statistic
    BatchNorm, dropout off) + CE(sum) + backward (torch autograd on CPU, fp32); optimizer step not included."""
    import torch
    from oracle import encode_np as E
    from oracle import network_t as NT
    torch.set_num_threads(os.cpu_count())
    sd = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=("running" not in k and "num_batches" not in k))
          for k, v in state.items()}
    sym = E._ASCII2SYM[chroms[0]]
    done, t0 = 0, time.perf_counter()
    while done + batch <= len(pos) and (done == 0 or time.perf_counter() - t0 < budget_s):
        p, mt = pos[done:done + batch], meta_lab[done:done + batch]
        cat = E.kmer_windows(sym, p, mt & 1, cfg["local_radius"], cfg["local_order"])
        oh = E.onehot_windows(sym, p, mt & 1, cfg["distal_radius"])
        out = NT.network2_forward(sd, cat, oh, torch.float32, train=True)
        NT.ce_sum(out, (mt >> 1) & 0x7f).backward()
        done += batch
    dt = time.perf_counter() - t0
    return done / dt, dt, done


# ------------------------------------------------------------------------------------------ train leg
def train_leg(genome, pos, meta, world, rank, dist, steps=30, warmup=5, chroms=None, cpu_legs=False):
    """BASELINE configs[2] on the same genome: MuRaL-snv training from scratch (local 10 bp 3-mers, expanded 1 Kb,
    Adam lr 1e-3), fused step = forward + CE(sum) + backward + flat-gradient all-reduce (NCCL, world > 1) + global-norm
    clip + Adam, one batch per GPU per step.  Sites: a seeded random subset of this rank's A/T sites (sorted), labels
    iid Categorical(0.952381, 0.0140095, 0.0198, 0.0138095) (training.py:332).  Returns sites/s per batch size."""
    import torch
    from mural_b200 import SiteBatch, model_choice, pack_meta, weights_init
    from mural_b200.training import TrainState
    cfg = {"local_radius": 10, "local_order": 3, "local_hidden1_size": 150, "local_hidden2_size": 75, "distal_radius": 1000,
           "emb_dropout": .1, "local_dropout": .1, "CNN_kernel_size": 3, "CNN_out_channels": 32, "distal_fc_dropout": .25,
           "n_class": 4, "model_no": 2}
    n_cat = 2 * 10 + 1 - 2
    torch.manual_seed(0)
    model = model_choice(2, cfg, dict(emb_dims=[(65, 2)] * n_cat, n_cont=0, n_class=4, distal_order=1, in_channels=4), "snv")
    model.apply(weights_init)
    model.to("cuda").train()
    ts = TrainState(model, "Adam", lr=1e-3, weight_decay=1e-5, seed=rank)
    rng = np.random.default_rng(4321 + rank)
    out = {}
    for B in (128, 4096):
        need = B * (steps + warmup)
        sel = np.sort(rng.choice(len(pos), size=min(need, len(pos)), replace=False))
        lab = rng.choice(4, size=len(sel), p=[0.952381, 0.0140095, 0.0198, 0.0138095])
        mt = pack_meta(meta[sel] & 1, lab, meta[sel] >> 8)
        order = rng.permutation(len(sel))                      # batches are drawn from a shuffled pool (training.py:241)
        d_pos = torch.from_numpy(pos[sel][order]).cuda(); d_meta = torch.from_numpy(mt[order]).cuda()
        k_tot = len(sel) // B
        w = min(warmup, k_tot - 1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for i in range(k_tot):
            if i == w:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                ev0.record()
            ts.step(SiteBatch(d_pos[i * B:(i + 1) * B], d_meta[i * B:(i + 1) * B], genome))
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out["batch_%d" % B] = {"sites_per_s": world * B * (k_tot - w) / (ms * 1e-3), "ms_per_step": ms / (k_tot - w), "steps": k_tot - w}
    loss = float(ts.loss_dev.item())
    res = {"metric": "sites/sec (train: fwd+bwd+clip+Adam, fused step)", "value": out["batch_128"]["sites_per_s"], "unit": "sites/s",
           "batch_per_gpu": 128, "large_batch": out["batch_4096"], "small_batch": out["batch_128"], "dtype": "f32",
           "config": "MuRaL-snv from scratch, local 10bp 3-mers + expanded 1Kb, Adam lr 1e-3, batch per GPU as stated; "
                     "gradient all-reduce over NCCL when n_gpus > 1", "loss_sum_finite": bool(np.isfinite(loss))}
    fl = snv_train_flops(2001)
    res["roofline"] = leg_roofline(fl, out["batch_4096"]["sites_per_s"], world, "whole training step at batch 4096 per GPU (fp32-equivalent "
                                   "arithmetic: fp32 FMA forward, split-bf16 mma.sync dgrad / wgrad)")
    res["roofline"]["batch_128"] = leg_roofline(fl, out["batch_128"]["sites_per_s"], world, "batch 128 per GPU")["frac"]
    if rank == 0 and cpu_legs:
        sd0 = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        on0 = np.flatnonzero((meta >> 8) == 0)                  # the CPU sample stays on the first chromosome
        sel = np.sort(rng.choice(on0, size=min(16384, len(on0)), replace=False))   # ~6 s of CPU work
        lab = rng.choice(4, size=len(sel), p=[0.952381, 0.0140095, 0.0198, 0.0138095])
        v, dt, done = cpu_snv_train(chroms, pos[sel], pack_meta(meta[sel] & 1, lab, meta[sel] >> 8), cfg, sd0)
        res["cpu_baseline"] = {"value": v, "unit": "sites/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d sites in batches of 128: numpy encoders + train-mode Network2 forward + CE(sum) + backward "
                                         "(torch CPU fp32 autograd of the oracle, no optimizer step), %.1f s" % (done, dt)}
    return res


# ------------------------------------------------------------------------------------------ transfer sweep leg
def sweep_leg(genome, pos, meta, cfg0, state, n_cat, radii=(100, 200, 500, 1000, 2000, 5000), n_predict=262144, B=128, steps=20):
    """SURVEY 8(d) config 5 (transfer sweep): one checkpoint at distal radii 100 ... 5000 (L = 201 ... 10 001; every tensor of
    Network2 is radius-independent).  Per radius: predict sites/s on the first n_predict sites of this rank's interval (bf16
    path, CUDA events) and fine-tuning sites/s of the fused step at batch B (all parameters trained, graph replay).  The
    shipped Homo_sapiens/SNV/AT weights stand in for the Macaca ones (same architecture).  Rank-local, no collective."""
    import torch
    from mural_b200 import SiteBatch, pack_meta
    from mural_b200.training import TrainState
    rng = np.random.default_rng(55)
    d_pos = torch.from_numpy(pos[:n_predict]).cuda()
    d_meta = torch.from_numpy(meta[:n_predict]).cuda()
    same_chrom = int(np.searchsorted(meta[:n_predict] >> 8, (meta[0] >> 8) + 1))      # keep to the first chromosome of the slice
    d_pos, d_meta = d_pos[:same_chrom], d_meta[:same_chrom]
    lab = rng.choice(4, size=B * (steps + 4), p=[0.952381, 0.0140095, 0.0198, 0.0138095])
    sel = np.sort(rng.choice(same_chrom, size=len(lab), replace=False))
    t_pos = torch.from_numpy(pos[sel]).cuda()
    t_meta = torch.from_numpy(pack_meta(meta[sel] & 1, lab, meta[sel] >> 8)).cuda()
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for R in radii:
        cfg = dict(cfg0, distal_radius=int(R))
        m = build_model(cfg, state, n_cat, "bf16")
        sb = SiteBatch(d_pos, d_meta, genome)
        with torch.no_grad():
            m.forward(None, sb)
            ev0.record()
            for _ in range(3):
                lp = m.forward(None, sb)
            ev1.record()
        torch.cuda.synchronize()
        pred = 3 * len(sb) / (ev0.elapsed_time(ev1) * 1e-3)
        m.train()
        ts = TrainState(m, "Adam", lr=1e-4, weight_decay=1e-5, seed=1)
        for i in range(steps + 4):
            if i == 4:
                ev0.record()
            ts.step(SiteBatch(t_pos[i * B:(i + 1) * B], t_meta[i * B:(i + 1) * B], genome))
        ev1.record()
        torch.cuda.synchronize()
        out[str(R)] = {"L": 2 * int(R) + 1, "predict_sites_per_s": pred, "train_sites_per_s": B * steps / (ev0.elapsed_time(ev1) * 1e-3),
                       "finite": bool(torch.isfinite(lp).all().item() and np.isfinite(float(ts.loss_dev.item())))}
        del ts, m
    return {"metric": "sites/sec per distal radius (transfer sweep, config 5)", "unit": "sites/s", "predict_sites": int(same_chrom),
            "train_batch": B, "dtype": "bf16 predict / f32 train", "radii": out}


# ------------------------------------------------------------------------------------------ indel leg
def indel_leg(genome, world, rank, dist, batch=2048, steps=5, warmup=2, chroms=None, cpu_legs=False):
    """BASELINE configs[3] (predict half): MuRaL-indel UNet_Small with the shipped Homo_sapiens/INDEL/insertion weights
    (tests/golden/indel_hs_ins.npz), one site every 50 bp on the '+' strand of this rank's interval, expanded radius as
    in the checkpoint's config.  Default kernels: the fused tensor-core level kernels (csrc/indel_tc.cuh)."""
    import torch
    from mural_b200 import SiteBatch, model_choice, pack_meta
    z = np.load(os.path.join(ROOT, "tests", "golden", "indel_hs_ins.npz"))
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    cfg = {"CNN_out_channels": state["uplblocks.0.0.weight"].shape[0], "CNN_kernel_size": state["uplblocks.0.0.weight"].shape[2],
           "down_list": [int(v) for v in z["down"]], "use_reverse": bool(z["use_reverse"]), "n_class": state["out_fc.2.weight"].shape[0]}
    m = model_choice(0, cfg, {"n_class": cfg["n_class"]}, "indel")
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
    m.to("cuda").eval()
    Rd = int(z["distal_radius"])
    n = batch * (steps + warmup)
    lo = 20_000 + (CHROM_LEN - 40_000) * rank // max(world, 1) // 50 * 50
    pos = (lo + 50 * np.arange(n)) % (CHROM_LEN - 40_000) + 20_000
    d_pos = torch.from_numpy(pos.astype(np.int32)).cuda()
    d_meta = torch.from_numpy(pack_meta(np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64))).cuda()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for i in range(steps + warmup):
            if i == warmup:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                ev0.record()
            out = m.forward(SiteBatch(d_pos[i * batch:(i + 1) * batch], d_meta[i * batch:(i + 1) * batch], genome), distal_radius=Rd)
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    from mural_b200 import _lib
    tc = _lib.lib().mural_indel_tc_available(m._handle(Rd)) == 1
    sps = world * batch * steps / (ms * 1e-3)
    res = {"metric": "sites/sec (MuRaL-indel predict)", "value": sps, "unit": "sites/s", "batch_per_gpu": batch,
           "ms_per_step": ms / steps, "dtype": "bf16 x2 split (fp32-equivalent), fp32 accumulate" if tc else "f32",
           "kernels": "fused tensor-core level kernels (k_unet_level, mma.sync)" if tc else "fp32 CUDA-core kernels",
           "finite": bool(torch.isfinite(out).all().item()),
           "config": "UNet_Small, Homo_sapiens/INDEL/insertion weights, expanded radius %d (L=%d), sites every 50 bp" % (Rd, 2 * Rd)}
    fl = unet_flops(cfg["CNN_out_channels"], cfg["CNN_kernel_size"], cfg["down_list"], 2 * Rd, cfg["use_reverse"])
    res["roofline"] = leg_roofline(fl, sps, world, "whole predict step (stem + 11 level kernels + head); every product is evaluated as three "
                                   "bf16 MMAs (hi*hi + hi*lo + lo*hi), so the executed tensor work is 3x the algorithmic count used here")
    m.compute_mode = "fp32"                                  # the CUDA-core kernels on the same sites, for the record
    with torch.no_grad():
        m.forward(SiteBatch(d_pos[:batch], d_meta[:batch], genome), distal_radius=Rd)
        torch.cuda.synchronize(); ev0.record()
        o32 = m.forward(SiteBatch(d_pos[:batch], d_meta[:batch], genome), distal_radius=Rd)
        ev1.record(); torch.cuda.synchronize()
    res["fp32_kernels_sites_per_s"] = batch / (ev0.elapsed_time(ev1) * 1e-3)
    m.compute_mode = "auto"
    if rank == 0 and cpu_legs:
        from oracle import encode_np as E
        v, dt, done, ref = cpu_unet(state, cfg, Rd, E._ASCII2SYM[chroms[0]], pos[:2048].astype(np.int64), budget_s=6.0)
        with torch.no_grad():
            got = m.forward(SiteBatch(d_pos[:done], d_meta[:done], genome), distal_radius=Rd).cpu().numpy()
        res["cpu_baseline"] = {"value": v, "unit": "sites/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "first %d sites in batches of 8: numpy one-hot windows + torch CPU fp32 UNet_Small (oracle), %.1f s" % (done, dt)}
        scale = float(max(1.0, np.abs(ref).max()))
        res["parity_spot_check"] = {"sites": int(done), "max_abs_diff": float(np.abs(got - ref).max()), "scale": scale,
                                    "tolerance": 1e-3 * scale, "what": "GPU outputs vs the CPU oracle on the same sites (fp32-equivalent gate)"}
        assert res["parity_spot_check"]["max_abs_diff"] <= res["parity_spot_check"]["tolerance"], res["parity_spot_check"]
    return res


def eval_leg(genome, pos, meta, logp, cfg, n=1_000_000, reps=5):
    """Per-epoch validation metrics of the reference's Evaluator (training.py:488-520) on the first n sites of a
    step: 3/5/7-mer correlations, the regional score (3- and 5-mer tables per 10 000-site region) and the 100 kb / 500 kb window
    correlations — device reductions (csrc/metrics.cu) + host Pearson over the small tables.  Rank-local, no collective."""
    import torch
    from mural_b200.calibration import calibrate
    from mural_b200.evaluation import EvalData, Evaluator
    n = int(min(n, len(pos), logp.shape[0]))
    rng = np.random.default_rng(99)
    labels = rng.choice(4, n, p=[0.952381, 0.0140095, 0.0198, 0.0138095])
    m = (np.asarray(meta[:n]).astype(np.int64) & ~0xfe) | (labels << 1)
    d_meta = torch.from_numpy(m.astype(np.int32)).cuda()
    d_pos = torch.from_numpy(np.asarray(pos[:n]).astype(np.int32)).cuda()
    flank = genome.encode_local(d_pos, d_meta, cfg["local_radius"], 1)
    prob = calibrate(logp[:n].contiguous())
    ed = EvalData(flank, d_meta, prob, f32=True, start=d_pos)
    lines = []

    def epoch_metrics():
        E = Evaluator(ed, None, cfg["n_class"], printer=lambda *a: lines.append(a))
        E.evaluate_kmer([3, 5, 7])
        E.evaluate_regional_score(n, [3, 5])
        return E
    epoch_metrics()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):                                    # wall clock per pass (host Pearson included); median: the host side shares
        t0 = time.perf_counter()                             # the box's cores with whatever else the bench has left running
        E = epoch_metrics()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    # device time of each reduction kernel alone (library CUDA-event profile); algorithmic bytes per site: the 2d flank codes
    # (int64) + meta + n_class fp64 probabilities for the k-mer tables; pos + meta (read twice) + probabilities for the windows
    from mural_b200 import _lib
    from mural_b200.evaluation import kmer_group_table, window_table
    L = _lib.lib()
    kern = {}
    for name, fn, bytes_site in (("kmer3", lambda: kmer_group_table(ed, 3), 2 * 8 + 4 + 8 * cfg["n_class"]),
                                 ("kmer5", lambda: kmer_group_table(ed, 5), 4 * 8 + 4 + 8 * cfg["n_class"]),
                                 ("kmer7", lambda: kmer_group_table(ed, 7), 6 * 8 + 4 + 8 * cfg["n_class"]),
                                 ("kmer5_regions", lambda: kmer_group_table(ed, 5, 10000), 4 * 8 + 4 + 8 * cfg["n_class"]),
                                 ("window100k", lambda: window_table(ed, 100000), 2 * 8 + 8 * cfg["n_class"])):
        fn()
        torch.cuda.synchronize()
        L.mural_profile_begin()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        buf = C.create_string_buffer(1 << 16)
        L.mural_profile_end(buf, len(buf))
        prof = json.loads(buf.value.decode())
        us = sum(v["ms"] for v in prof.values()) * 1e3 / reps
        kern[name] = {"us": round(us, 1), "algorithmic_GBps": round(n * bytes_site / (us * 1e-6) / 1e9, 1),
                      "kernels": sorted(prof)}
    return {"metric": "validation sites/sec (Evaluator: 3/5/7-mer correlations + regional score)", "value": n / dt, "unit": "sites/s",
            "sites": n, "ms_per_epoch_metrics": dt * 1e3, "kernels": kern, "bound": "hbm",
            "kmer3_corr": [round(float(c), 4) for c in E.metrics["kmer3"]], "regional_score": float(E.metrics["score"]),
            "note": "includes the host sync + table copy of every launch; labels synthetic (class proportions of training.py:332)"}


def indel_train_leg(genome, world, rank, dist, batch=32, steps=20, warmup=3, chroms=None, cpu_legs=False):
    """BASELINE configs[3] (training half): MuRaL-indel UNet_Small fine-tuned from the shipped Homo_sapiens/INDEL/insertion
    weights at its own radius (L = 8000), one site every 50 bp on the '+' strand, labels iid Categorical(0.907, 0.0133 x 7),
    fused step = train-mode forward + CE(sum) + backward + flat-gradient all-reduce (NCCL, world > 1) + clip + Adam.
    Tape of shared-memory-tiled fp32 kernels (csrc/indel_train.cu, indel_train_tiled.cuh)."""
    import torch
    from mural_b200 import SiteBatch, model_choice, pack_meta
    from mural_b200.training import IndelTrainState
    z = np.load(os.path.join(ROOT, "tests", "golden", "indel_hs_ins.npz"))
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    cfg = {"CNN_out_channels": state["uplblocks.0.0.weight"].shape[0], "CNN_kernel_size": state["uplblocks.0.0.weight"].shape[2],
           "down_list": [int(v) for v in z["down"]], "use_reverse": bool(z["use_reverse"]), "n_class": state["out_fc.2.weight"].shape[0]}
    m = model_choice(0, cfg, {"n_class": cfg["n_class"]}, "indel")
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
    m.to("cuda").train()
    Rd = int(z["distal_radius"])
    ts = IndelTrainState(m, Rd, "Adam", lr=1e-4, weight_decay=1e-5, seed=rank)
    n = batch * (steps + warmup)
    rng = np.random.default_rng(777 + rank)
    lo = 20_000 + (CHROM_LEN - 40_000) * rank // max(world, 1) // 50 * 50
    pos = (lo + 50 * np.arange(n)) % (CHROM_LEN - 40_000) + 20_000
    lab = rng.choice(8, size=n, p=[0.907] + [0.093 / 7] * 7)
    d_pos = torch.from_numpy(pos.astype(np.int32)).cuda()
    d_meta = torch.from_numpy(pack_meta(np.zeros(n, np.int64), lab, np.zeros(n, np.int64))).cuda()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(steps + warmup):
        if i == warmup:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0.record()
        ts.step(SiteBatch(d_pos[i * batch:(i + 1) * batch], d_meta[i * batch:(i + 1) * batch], genome))
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sps = world * batch * steps / (ms * 1e-3)
    res = {"metric": "sites/sec (MuRaL-indel train: fwd+bwd+clip+Adam)", "value": sps, "unit": "sites/s",
           "batch_per_gpu": batch, "ms_per_step": ms / steps, "dtype": "f32", "loss_sum_finite": bool(np.isfinite(float(ts.loss_dev.item()))),
           "config": "UNet_Small, Homo_sapiens/INDEL/insertion weights, expanded radius %d (L=%d), Adam lr 1e-4" % (Rd, 2 * Rd)}
    fl = 3 * unet_flops(cfg["CNN_out_channels"], cfg["CNN_kernel_size"], cfg["down_list"], 2 * Rd, cfg["use_reverse"])
    res["roofline"] = leg_roofline(fl, sps, world, "whole training step at batch %d per GPU (fwd + dgrad + wgrad ~= 3 x forward FLOPs; fp32 CUDA-core "
                                   "tiled kernels)" % batch)
    if rank == 0 and cpu_legs:
        from oracle import encode_np as E
        v, dt, done, _ = cpu_unet(state, cfg, Rd, E._ASCII2SYM[chroms[0]], pos[:512].astype(np.int64), train=True, labels=lab[:512], batch=8,
                                  budget_s=6.0)
        res["cpu_baseline"] = {"value": v, "unit": "sites/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "%d sites in batches of 8: numpy one-hot windows + train-mode UNet_Small forward + CE(sum) + backward "
                                         "(torch CPU fp32 autograd of the oracle, no optimizer step), %.1f s" % (done, dt)}
    return res



# ------------------------------------------------------------------------------------------ context: torch eager on this GPU
def torch_eager_leg(genome, pos, meta, cfg, state, n_cat):
    """Opt-in context numbers (BASELINE.md par. 4, item 5): the reference networks' arithmetic as plain PyTorch library calls on
    the SAME GPU (cuDNN / cuBLAS, eager, TF32 allowed as the reference's GPU environment would run it) — the oracle's
    torch-functional restatement fed with device tensors; inputs are the reference's own tensors (one-hot [B,4,L] + int64 k-mer
    indices), produced here by the device encoders and NOT timed.  Stated beside the product numbers, never as a target."""
    import torch
    from mural_b200 import model_choice, pack_meta, weights_init
    from oracle import network_t as NT
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    res = {"note": "torch %s eager, TF32 allowed, inputs pre-encoded on the device (not timed)" % torch.__version__}
    dev = torch.device("cuda")
    sd = {k: torch.as_tensor(np.asarray(v)).to(dev) for k, v in state.items()}

    def timeit(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    for B in (128, 4096):
        d_pos = torch.from_numpy(pos[:B]).cuda(); d_meta = torch.from_numpy(meta[:B]).cuda()
        cat = genome.encode_local(d_pos, d_meta, cfg["local_radius"], cfg["local_order"])
        oh = genome.encode_onehot(d_pos, d_meta, cfg["distal_radius"])
        with torch.no_grad():
            t = timeit(lambda: NT.network2_forward(sd, cat, oh, torch.float32), 5)
        res["snv_predict_batch_%d" % B] = {"sites_per_s": B / t, "ms": t * 1e3}
        sdt = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
        lab = torch.randint(0, 4, (B,), device=dev)

        def train_step():
            for v in sdt.values():
                if v.requires_grad:
                    v.grad = None
            NT.ce_sum(NT.network2_forward(sdt, cat, oh, torch.float32, train=True), lab).backward()
        t = timeit(train_step, 5)
        res["snv_train_fwd_bwd_batch_%d" % B] = {"sites_per_s": B / t, "ms": t * 1e3}
    z = np.load(os.path.join(ROOT, "tests", "golden", "indel_hs_ins.npz"))
    ist = {k[2:]: torch.as_tensor(np.asarray(z[k])).to(dev) for k in z.files if k.startswith("w:")}
    down, rev, Rd = [int(v) for v in z["down"]], bool(z["use_reverse"]), int(z["distal_radius"])
    for B, tr in ((256, False), (32, True)):
        p = torch.from_numpy((20000 + 50 * np.arange(B)).astype(np.int32)).cuda()
        mt = torch.from_numpy(pack_meta(np.zeros(B, np.int64), np.zeros(B, np.int64), np.zeros(B, np.int64))).cuda()
        oh = genome.encode_onehot(p, mt, Rd, "indel")
        if not tr:
            with torch.no_grad():
                t = timeit(lambda: NT.unet_small_forward(ist, oh, down, rev, torch.float32), 3)
            res["indel_predict_batch_%d" % B] = {"sites_per_s": B / t, "ms": t * 1e3}
        else:
            sdt = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in ist.items()}
            lab = torch.randint(0, 8, (B,), device=dev)

            def istep():
                for v in sdt.values():
                    if v.requires_grad:
                        v.grad = None
                NT.ce_sum(NT.unet_small_forward(sdt, oh, down, rev, torch.float32, train=True), lab).backward()
            t = timeit(istep, 3)
            res["indel_train_fwd_bwd_batch_%d" % B] = {"sites_per_s": B / t, "ms": t * 1e3}
    return res


# ------------------------------------------------------------------------------------------ sparse-site leg
def sparse_leg(genome, chroms, model, cfg, state, mode, n=262144, steps=5, n_check=2048):
    """Sparse-site predict (BED subsets, validation sets, every training-set predict): a seeded 1-in-100 sample of the A/T
    sites of ALL chromosomes, in (chromosome, position) order, so the windows of neighbouring sites hardly overlap and a call
    spans several chromosomes -> the per-site kernels (stem gather per site, RB4/C_RB4 on per-site rows), not the dense
    lattice.  Device-timed with CUDA events, inputs resident; parity spot check of the first sites against the CPU oracle."""
    import torch
    from mural_b200 import SiteBatch, pack_meta
    from oracle import encode_np as E
    from oracle import network_t as NT
    rng = np.random.default_rng(2718)
    per = n // len(chroms)
    pos_l, meta_l = [], []
    for ci, c in enumerate(chroms):
        span = c[: per * 220]                                   # ~ per*110 A/T sites -> keep 1 in ~100
        idx = np.flatnonzero((span == ord("A")) | (span == ord("T")))
        sel = np.sort(rng.choice(idx, size=per, replace=False))
        pos_l.append(sel.astype(np.int32))
        meta_l.append(pack_meta((span[sel] == ord("T")).astype(np.int64), np.zeros(per, np.int64), np.full(per, ci)))
    pos, meta = np.concatenate(pos_l), np.concatenate(meta_l)
    sb = SiteBatch(torch.from_numpy(pos).cuda(), torch.from_numpy(meta).cuda(), genome)
    model.compute_mode = mode
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        lp = model.forward(None, sb)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(steps):
            lp = model.forward(None, sb)
        ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    got = lp[:n_check].cpu().numpy()
    sym = E._ASCII2SYM[chroms[0]]
    cat = E.kmer_windows(sym, pos[:n_check], meta[:n_check] & 1, cfg["local_radius"], cfg["local_order"])
    oh = E.onehot_windows(sym, pos[:n_check], meta[:n_check] & 1, cfg["distal_radius"])
    with torch.no_grad():
        ref = NT.network2_forward(state, cat, oh, torch.float32).numpy()
    sm = lambda x: np.exp(x - x.max(1, keepdims=True)) / np.exp(x - x.max(1, keepdims=True)).sum(1, keepdims=True)
    return {"metric": "sites/sec (predict, sparse sites)", "value": len(pos) / (ms * 1e-3), "unit": "sites/s", "sites": int(len(pos)),
            "ms_per_call": ms, "mode": mode, "mean_site_spacing_bp": float(np.mean(np.diff(pos_l[0]))),
            "parity_spot_check": {"sites": n_check, "max_abs_dp": float(np.abs(sm(got) - sm(ref)).max()), "tolerance": 1e-3 if mode == "fp32" else 5e-3},
            "config": "1-in-100 A/T sites of all %d chromosomes in one call (per-site kernels)" % len(chroms)}


# ------------------------------------------------------------------------------------------ full pipeline leg
def pipeline_leg(chroms, cfg, state, n_cat, world, rank, dist, mode, n_sites=4_000_000):
    """mural_b200.predict.run_predict end to end (MuRaL/scripts/run_predict.py:34-239): FASTA + BED + checkpoint files on disk ->
    packed genome, site order, sharded compute, calibration, ONE NCCL gather, '%.4g' TSV on rank 0.  Chromosome 1 of the
    config-2 genome (25 Mb) and its first n_sites A/T sites; wall seconds per stage, max over ranks for the shared stages."""
    import pickle
    import shutil
    import tempfile
    import torch
    from mural_b200.predict import run_predict
    tmp = tempfile.mkdtemp(prefix="mural_pipe_") if rank == 0 else None
    if world > 1:
        box = [tmp]
        dist.broadcast_object_list(box, src=0)
        tmp = box[0]
    fa, bed = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "sites.bed")
    try:
        if rank == 0:
            c = chroms[0]
            with open(fa, "wb") as f:
                f.write(b">chr1\n")
                rows = c[: len(c) // 100 * 100].reshape(-1, 100)
                f.write(np.concatenate([rows, np.full((len(rows), 1), 10, np.uint8)], 1).tobytes())
            idx = np.flatnonzero((c == ord("A")) | (c == ord("T")))[:n_sites]
            strand = np.where(c[idx] == ord("T"), "-", "+")
            import pandas as pd
            pd.DataFrame({"c": "chr1", "s": idx, "e": idx + 1, "n": ".", "l": 0, "d": strand}).to_csv(bed, sep="\t", header=False, index=False)
            m = build_model(cfg, state, n_cat, "fp32")
            torch.save({k: v.cpu() for k, v in m.state_dict().items()}, os.path.join(tmp, "model"))
            pickle.dump(dict(cfg, emb_dims=[(65, 2)] * n_cat, segment_center=300000), open(os.path.join(tmp, "model.config.pkl"), "wb"))
            del m
        if world > 1:
            dist.barrier()
        tm = {}
        t0 = time.perf_counter()
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):           # stdout carries the one JSON line only
            run_predict(bed, fa, os.path.join(tmp, "model"), os.path.join(tmp, "model.config.pkl"), "", os.path.join(tmp, "pred.tsv"),
                        compute_mode=mode, return_frame=False, timings=tm)
        if world > 1:
            dist.barrier()
        total = time.perf_counter() - t0
        size = os.path.getsize(os.path.join(tmp, "pred.tsv")) if rank == 0 else 0
        keys = ["fasta_ingest_s", "bed_ingest_s", "site_order_s", "model_load_s", "compute_s", "calibrate_s", "gather_s"]
        v = torch.tensor([tm.get(k, 0.0) for k in keys], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        stages = {k: round(float(x), 4) for k, x in zip(keys, v.tolist())}
        stages["tsv_write_s"] = round(tm.get("tsv_write_s", 0.0), 4)
        n = int(tm.get("sites", n_sites))
        return {"metric": "sites/sec (run_predict: BED + FASTA -> calibrated TSV, wall clock)", "value": n / total, "unit": "sites/s", "sites": n,
                "total_s": round(total, 3), "stages": stages, "tsv_bytes": int(size), "mode": mode, "n_gpus": world,
                "config": "chr1 (25 Mb) of the config-2 genome, first %d A/T sites, segment_center 300000, one gather of fp64 [n,4]" % n}
    finally:
        if world > 1:
            dist.barrier()
        if rank == 0:
            shutil.rmtree(tmp, ignore_errors=True)


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_port_sites_per_sec(chroms, pos, meta, cfg, state, batch=1024):
    """Oracle port of the reference CPU path: numpy window/k-mer encoders + torch CPU fp32 Network2."""
    import torch
    from oracle import encode_np as E
    from oracle import network_t as NT
    torch.set_num_threads(os.cpu_count())
    comp_lut = E._ASCII2SYM
    syms = {}
    t0 = time.perf_counter()
    n = len(pos)
    out = []
    with torch.no_grad():
        for b0 in range(0, n, batch):
            p, mt = pos[b0:b0 + batch], meta[b0:b0 + batch]
            ch = mt >> 8
            cat = np.empty((len(p), 2 * cfg["local_radius"] + 1 - (cfg["local_order"] - 1)), np.int64)
            oh = np.empty((len(p), 4, 2 * cfg["distal_radius"] + 1), np.float32)
            for c in np.unique(ch):
                if c not in syms:
                    syms[c] = comp_lut[chroms[c]]
                m = ch == c
                cat[m] = E.kmer_windows(syms[c], p[m], mt[m] & 1, cfg["local_radius"], cfg["local_order"])
                oh[m] = E.onehot_windows(syms[c], p[m], mt[m] & 1, cfg["distal_radius"])
            out.append(NT.network2_forward(state, cat, oh, torch.float32))
    dt = time.perf_counter() - t0
    cpu_port_sites_per_sec.last_logp = torch.cat(out).numpy() if out else None   # kept for the GPU-vs-oracle spot check
    return n / dt, dt


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mural_b200", choices=["mural_b200", "reference"])
    ap.add_argument("--mode", default=os.environ.get("MURAL_BENCH_MODE", "auto"), choices=["auto", "fp32", "bf16"],
                    help="auto = the product default (MURAL_MODE_AUTO: bf16 tensor-core path + fp32-equivalent recompute of the "
                         "exception windows, decided on the device); bf16 / fp32 force one path")
    ap.add_argument("--no-sparse", action="store_true", help="skip the sparse-site predict leg (per-site kernels)")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the BED+FASTA -> TSV pipeline leg (run_predict)")
    ap.add_argument("--sites-per-step", type=int, default=1048576)
    ap.add_argument("--cpu-sample", type=int, default=131072)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training leg (BASELINE configs[2])")
    ap.add_argument("--no-indel", action="store_true", help="skip the MuRaL-indel predict leg (BASELINE configs[3])")
    ap.add_argument("--no-sweep", action="store_true", help="skip the distal-radius sweep leg (SURVEY 8d config 5)")
    ap.add_argument("--torch-eager", action="store_true", help="opt-in context leg: the oracle's torch-functional networks as eager "
                    "PyTorch (cuDNN/cuBLAS, TF32) on this GPU, for predict and fwd+bwd (BASELINE.md par. 4 item 5)")
    ap.add_argument("--no-eval", action="store_true", help="skip the validation-metrics leg (Evaluator, SURVEY 8f N3)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg, state, n_cat = load_weights()
    workload = ("MuRaL-snv genome-wide predict, Homo_sapiens/SNV/AT weights, synthetic 100 Mb genome (4x25 Mb), "
                "every A(+)/T(-) site, local 7bp 3-mers + expanded 1 Kb (L=2001), C=32, n_class=4")
    base_cfg = {"workload": workload, "sites_per_step_per_gpu": a.sites_per_step, "parallelism": "interval-sharded x%d" % world}

    # -------------------------------------------------------------------- reference arm (CPU oracle port)
    if a.impl == "reference":
        if rank != 0:
            return 0
        chroms = [synth_chromosome(0)]                # the bounded sample lives on the first chromosome
        per_step = min(a.cpu_sample, a.sites_per_step, 32768)      # bounded: K steps of this stay within a few minutes on 8-16 cores
        pos, meta = rank_sites(chroms, 0, 1, per_step * (a.steps + a.warmup))
        for w in range(a.warmup):
            cpu_port_sites_per_sec(chroms, pos[w * per_step:(w + 1) * per_step][:2048], meta[w * per_step:(w + 1) * per_step][:2048], cfg, state)
        t = 0.0
        for s in range(a.warmup, a.warmup + a.steps):
            _, dt = cpu_port_sites_per_sec(chroms, pos[s * per_step:(s + 1) * per_step], meta[s * per_step:(s + 1) * per_step], cfg, state)
            t += dt
        v = per_step * a.steps / t
        line = {"impl": "reference", "metric": "sites/sec (predict)", "value": v, "unit": "sites/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": dict(base_cfg, sites_per_step_per_gpu=per_step),
                "cpu_baseline": {"value": v, "unit": "sites/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": "%d consecutive A/T sites of chr0 per step, batch 1024, numpy encoders + torch CPU fp32" % per_step},
                "e2e": {"value": v, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # -------------------------------------------------------------------- mural_b200 arm
    import torch
    import torch.distributed as dist
    from mural_b200 import PackedGenome, SiteBatch, _lib
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()
    mode = a.mode
    chroms = [synth_chromosome(ci) for ci in range(N_CHROM)]
    genome = PackedGenome({"chr%d" % (i + 1): c.tobytes() for i, c in enumerate(chroms)})
    S, K, W = a.sites_per_step, a.steps, a.warmup
    pos, meta = rank_sites(chroms, rank, world, S * (K + W))
    model = build_model(cfg, state, n_cat, mode)
    model.refresh()
    if mode != "fp32" and L.mural_snv_tc_available(model._h) != 1:
        raise RuntimeError("the tcgen05 path is not available for this model shape; run with --mode fp32")
    d_pos, d_meta = torch.from_numpy(pos).cuda(), torch.from_numpy(meta).cuda()
    out = torch.empty((S, cfg["n_class"]), dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    # a batch that straddles two chromosomes is issued as one call per chromosome (as mural_b200.predict.predict_sites does):
    # the dense-site path works on single-chromosome chunks
    chrom_of = meta >> 8
    cuts = [[0] + [int(c) + 1 for c in np.flatnonzero(np.diff(chrom_of[i * S:(i + 1) * S]))] + [S] for i in range(K + W)]
    NCb = cfg["n_class"]

    auto_sites = [0]

    def step(i, md=None):
        for a0, a1 in zip(cuts[i][:-1], cuts[i][1:]):
            _lib.check(L.mural_snv_forward(model._h, genome.handle, C.c_void_p(d_pos.data_ptr() + 4 * (i * S + a0)),
                                           C.c_void_p(d_meta.data_ptr() + 4 * (i * S + a0)), a1 - a0, _lib.MODES[md or mode],
                                           C.c_void_p(out.data_ptr() + 4 * NCb * a0), C.c_void_p(stream.cuda_stream)))
            if (md or mode) == "auto":
                auto_sites[0] += int(L.mural_snv_last_auto_sites(model._h))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    model.refresh()
    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    L.mural_reset_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_wall = time.perf_counter()
    for i in range(K):
        flush.fill_(i & 0xff)                       # L2 flush (256 MiB write), outside the per-step event pair
        ev[i][0].record(stream)
        step(W + i)
        ev[i][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(L.mural_launch_count())
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * S * K / (dev_ms * 1e-3)
    n_auto = auto_sites[0] / max(1, W + K)            # exception-window sites recomputed per step (auto mode)

    # ---- the same K steps with the bf16 path alone (no exception-window recompute): reported beside the headline
    bf16_only = None
    if mode == "auto":
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.fill_(i & 0xff)
            ev2[i][0].record(stream)
            step(W + i, "bf16")
            ev2[i][1].record(stream)
        barrier()
        t2 = torch.tensor([sum(s_.elapsed_time(e_) for s_, e_ in ev2)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        bf16_only = {"value": world * S * K / (float(t2.item()) * 1e-3), "unit": "sites/s", "ms_per_step": float(t2.item()) / K,
                     "note": "MURAL_MODE_BF16 on the same steps: every site through the tcgen05 path only"}

    # ---- e2e: host buffers through the C-ABI host entry point
    h_pos = torch.from_numpy(pos).pin_memory(); h_meta = torch.from_numpy(meta).pin_memory()
    h_out = torch.empty((S, cfg["n_class"]), dtype=torch.float32).pin_memory()

    def step_host(i):
        for a0, a1 in zip(cuts[i][:-1], cuts[i][1:]):
            _lib.check(L.mural_snv_predict_host(model._h, genome.handle, C.c_void_p(h_pos.data_ptr() + 4 * (i * S + a0)),
                                                C.c_void_p(h_meta.data_ptr() + 4 * (i * S + a0)), a1 - a0, _lib.MODES[mode],
                                                C.c_void_p(h_out.data_ptr() + 4 * NCb * a0), C.c_void_p(stream.cuda_stream)))
    for i in range(min(W, 2)):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        step_host(W + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()          # sampled during the timed region and the end-to-end region (same kernels under load)
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * S * K / float(te.item())

    sparse = None
    if rank == 0 and not a.no_sparse:
        try:
            sparse = sparse_leg(genome, chroms, model, cfg, state, mode)
        except Exception as e:
            sparse = {"error": "%s: %s" % (type(e).__name__, e)}

    pipe = None
    if not a.no_pipeline:
        try:
            pipe = pipeline_leg(chroms, cfg, state, n_cat, world, rank, dist, mode)
        except Exception as e:
            pipe = {"error": "%s: %s" % (type(e).__name__, e)}

    cpu_legs = world == 1 and not a.no_cpu_baseline   # CPU-port baselines of the secondary legs: rank 0 at N = 1 only
    # ---- training leg (fwd + bwd + optimizer), all ranks (it contains the gradient all-reduce)
    train = None
    if not a.no_train:
        try:
            train = train_leg(genome, pos, meta, world, rank, dist, chroms=chroms, cpu_legs=cpu_legs)
        except Exception as e:  # the predict line must survive a failure here
            train = {"error": "%s: %s" % (type(e).__name__, e)}

    indel = None
    if not a.no_indel:
        try:
            indel = indel_leg(genome, world, rank, dist, chroms=chroms, cpu_legs=cpu_legs)
        except Exception as e:
            indel = {"error": "%s: %s" % (type(e).__name__, e)}

    indel_train = None
    if not a.no_indel:
        try:
            indel_train = indel_train_leg(genome, world, rank, dist, chroms=chroms, cpu_legs=cpu_legs)
        except Exception as e:
            indel_train = {"error": "%s: %s" % (type(e).__name__, e)}

    sweep = None
    if world == 1 and not a.no_sweep:   # single-GPU leg: its fine-tuning steps would otherwise enter the gradient all-reduce on one rank only
        try:
            sweep = sweep_leg(genome, pos, meta, cfg, state, n_cat)
        except Exception as e:
            sweep = {"error": "%s: %s" % (type(e).__name__, e)}

    eager = None
    if rank == 0 and world == 1 and a.torch_eager:
        try:
            eager = torch_eager_leg(genome, pos, meta, cfg, state, n_cat)
        except Exception as e:
            eager = {"error": "%s: %s" % (type(e).__name__, e)}

    evalm = None
    if rank == 0 and not a.no_eval:
        try:
            step(0)
            evalm = eval_leg(genome, pos, meta, out, cfg)
        except Exception as e:
            evalm = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- roofline of the dominant kernel: library-side CUDA-event profile over an identical pass
    roof = None
    if rank == 0:
        roof = kernel_roofline(L, step, W, K, S, mode, cfg, torch.cuda.synchronize, pos)   # rank-local: no collective here
    line = {"metric": "sites/sec (predict)", "value": value, "unit": "sites/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if mode == "fp32" else "bf16", "data": "synthetic",
            "config": dict(base_cfg, mode=mode, l2="flushed between steps (256 MiB write outside the per-step CUDA-event pairs); "
                           "per-step activation workspace > L2", wall_s_timed_region=t_wall,
                           exception_window_sites_per_step=n_auto if mode == "auto" else None, bf16_only=bf16_only),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "sites/s", "h2d_bytes_per_step": 8 * S, "d2h_bytes_per_step": 4 * cfg["n_class"] * S},
            "roofline": roof, "sparse_predict": sparse, "pipeline": pipe, "train": train, "indel": indel, "indel_train": indel_train,
            "eval_metrics": evalm, "transfer_sweep": sweep, "torch_eager_same_gpu": eager}
    if rank == 0:
        if not a.no_cpu_baseline and world == 1:
            n_s = a.cpu_sample
            v, dt = cpu_port_sites_per_sec(chroms, pos[:n_s], meta[:n_s], cfg, state)
            line["cpu_baseline"] = {"value": v, "unit": "sites/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "first %d sites of the step-0 batch, batch 1024, numpy encoders + torch CPU fp32 (%.1f s)" % (n_s, dt)}
            # parity at bench scale: the same sites through the GPU path (as part of a full-size call) vs the oracle's log-probs
            ref_lp = getattr(cpu_port_sites_per_sec, "last_logp", None)
            if ref_lp is not None:
                step(0)
                torch.cuda.synchronize()
                got = out[:n_s].cpu().numpy()
                pg = np.exp(got - got.max(1, keepdims=True)); pg /= pg.sum(1, keepdims=True)
                pr = np.exp(ref_lp - ref_lp.max(1, keepdims=True)); pr /= pr.sum(1, keepdims=True)
                line["parity_spot_check"] = {"sites": int(n_s), "max_abs_dp": float(np.abs(pg - pr).max()), "tolerance": 1e-3 if mode == "fp32" else 5e-3,
                                             "what": "GPU probabilities of the first sites of a full-size step vs the CPU oracle"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def executed_conv_flops(pos, S, K, W, R):
    """FLOPs the stage kernels actually execute on the dense-site path (bf16 mode): stage 1 runs once per genomic
    position and strand on the lattice plus a 19-row edge pseudo-site per site; stages 2 and 3 run per site."""
    chunk = int(os.environ.get("MURAL_TC_CHUNK", "524288"))
    pools = {0: ((3, 3, 1), (3, 3, 1), (3, 3, 1)), 1: ((15, 15, 7), (7, 7, 3), (3, 3, 1))}
    L0 = {0: 201, 1: 2 * R + 1}
    rl = 0
    for i in range(W, W + K):
        p = pos[i * S:(i + 1) * S]
        for c0 in range(0, S, chunk):
            pc = p[c0:c0 + chunk]
            ns = len(pc)
            n_pos = int(pc.max()) - int(pc.min()) + 2 * R + 64
            for br in (0, 1):
                Ls = [L0[br]]
                for (k, s_, pd) in pools[br]:
                    Ls.append((Ls[-1] + 2 * pd - k) // s_ + 1)
                ps = pools[br][0][1]
                M = -(-n_pos // ps)
                rl += 4 * (2 * ps * (M + 1) + 1) + 4 * (ns * 19 + 1) + 5 * (ns * (Ls[2] + 1) + 1)   # stage 3 runs in k_tail
    return rl * 2.0 * 32 * 32 * 3


def kernel_roofline(L, step, W, K, S, mode, cfg, barrier, pos=None):
    if not hasattr(L, "mural_profile_begin"):
        return None
    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    src = "measured"
    if os.path.exists(p):
        peaks = json.load(open(p))
    else:
        peaks, src = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"
    L.mural_profile_begin()
    for i in range(K):
        step(W + i)
    barrier()
    buf = C.create_string_buffer(1 << 16)
    L.mural_profile_end(buf, len(buf))
    prof = json.loads(buf.value.decode() or "{}")
    conv = {k: v for k, v in prof.items() if ("stage_tc" in k if mode != "fp32" else ("conv" in k or "site_chain" in k))}
    if not conv:
        return None
    ms = sum(v["ms"] for v in conv.values()); n = sum(v["count"] for v in conv.values())
    total_ms = sum(v["ms"] for v in prof.values())
    alg = FLOP_CONV_ONLY - (FLOP_CONV3 if any("k_tail" in k for k in prof) else 0)   # FLOPs of the layers these kernels run
    flops_per_launch = alg * S * K / n
    achieved = flops_per_launch / (ms / n * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    executed = None
    if pos is not None and mode != "fp32" and any("lattice" in k for k in conv):
        ex = executed_conv_flops(pos, S, K, W, cfg["distal_radius"])
        executed = {"flops_per_launch": ex / n, "tflops": ex / (ms * 1e-3) / 1e12, "frac_of_peak": ex / (ms * 1e-3) / 1e12 / peak,
                    "share_of_algorithmic": ex / (alg * S * K),
                    "note": "dense-site reuse: stage 1 is evaluated once per genomic position and strand (lattice) plus 19 edge rows per "
                            "site, so fewer FLOPs are executed than the per-site algorithmic count that `achieved` uses (SURVEY 8d)"}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_traffic.json")          # ncu capture of the current round's stage kernels ...
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")      # ... else the round-1 capture
    if os.path.exists(tp) and mode != "fp32":
        tj = json.load(open(tp))      # dram__bytes_read+write of the stage kernels from the committed ncu --set full capture,
        traffic = tj["dram_bytes_per_site"] * S * K / n     # per site of a dense chunk, scaled to this run's launches
    return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "executed": executed,
            "kernel": "+".join(sorted(conv)), "launches": n, "profile_count": {k: v["count"] for k, v in prof.items()}, "mean_launch_ms": ms / n, "share_of_step": ms / total_ms,
            "peak_source": src + " (bf16 sustained, kernel timed inside a long step)",
            "flops_per_launch": flops_per_launch, "profile_ms": {k: round(v["ms"], 3) for k, v in prof.items()}}


if __name__ == "__main__":
    sys.exit(main())
