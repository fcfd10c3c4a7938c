"""Genome-wide prediction pipeline: BED + FASTA + checkpoint -> calibrated per-site probabilities (TSV).

Mirrors MuRaL/scripts/run_predict.py:34-239 (`run_predict_pipline`): same inputs (model, model.config.pkl,
model.fdiri_cal.pkl), same sample order, same output columns / sort order / '%.4g' formatting.  Differences:
windows are gathered on the GPU from the packed genome, and with WORLD_SIZE > 1 (torchrun) the site list is
split into contiguous genomic intervals, one per rank, with a single final gather (no collective on the data
path).
"""
import os
import pickle
import sys
import time

import numpy as np
import torch

from .calibration import calibrate, load_calibrator_weights
from .data import PackedSiteDataset, SiteBatch, SiteTable
from .genome import PackedGenome
from .nn_utils import model_choice


def shard_bounds(n, world, rank):
    """Contiguous, balanced split of n sites (in emission order = genomic order per strand batch)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_total, world, rank, group=None):
    """Concatenate the per-rank [n_r, k] tensors on rank 0 in rank order — the only collective of predict.  One
    `gather` of equal-size device buffers over NCCL/NVLink (shards differ by at most one row, so each rank pads to the largest
    shard); CPU tensors go the same way over gloo.  Returns None on the other ranks."""
    if world == 1:
        return local
    import torch.distributed as dist
    rows = [hi - lo for lo, hi in (shard_bounds(n_total, world, r) for r in range(world))]
    assert local.shape[0] == rows[rank], (local.shape, rows, rank)
    cap = max(rows)
    buf = local.new_zeros((cap,) + tuple(local.shape[1:]))
    buf[:rows[rank]] = local
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, parts, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.cat([p[:r] for p, r in zip(parts, rows)], dim=0)
    assert out.shape[0] == n_total
    return out


def load_config(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def build_model_from_files(model_path, config, device, n_cont=0, model_type="snv"):
    common = {"emb_dims": config.get("emb_dims"), "n_cont": n_cont, "n_class": config["n_class"], "distal_order": 1,
              "in_channels": 4 + n_cont}
    model = model_choice(config["model_no"], config, common, model_type)
    state = torch.load(model_path, map_location="cpu")
    model.load_state_dict(state)
    return model.to(device).eval()


def predict_sites(model, dataset, lo, hi, batch_sites=1 << 20):
    """log-probs [hi-lo, n_class] (CUDA) for the dataset's sites [lo, hi) in emission order."""
    dev = dataset.genome.device
    outs = []
    # batches never straddle two chromosomes: the dense-site path (stem tables, stage-1 lattice) works on single-chromosome chunks
    cuts = [lo] + [lo + int(c) + 1 for c in np.flatnonzero(np.diff(dataset.chrom[lo:hi]))] + [hi]
    spans = [(a, min(b0, a + batch_sites)) for a0, b0 in zip(cuts[:-1], cuts[1:]) for a in range(a0, b0, batch_sites)]
    with torch.no_grad():
        for a, b in spans:
            sb = SiteBatch(torch.from_numpy(dataset.pos[a:b]).to(dev), torch.from_numpy(dataset.meta[a:b]).to(dev), dataset.genome)
            if dataset.model_type == "indel":        # UNet_Small.forward(distal) (model_indel.py:151); batches bound the U-Net workspace
                outs.extend(model.forward(SiteBatch(sb.pos[c:c + 4096], sb.meta[c:c + 4096], sb.genome), distal_radius=dataset.distal_radius)
                            for c in range(0, b - a, 4096))
            else:
                outs.append(model.forward(None, sb))
    logp = torch.cat(outs) if outs else torch.empty((0, model.n_class), device=dev)
    if dataset.model_type == "snv" and getattr(model, "compute_mode", "fp32") == "auto_bf16":
        # precision policy of compute_mode="auto": the bf16 tensor-core path everywhere, and the fp32 kernels again for the
        # (rare) sites whose expanded window contains N / IUPAC symbols or overhangs the chromosome — inputs far from what
        # the network was trained on can drive activations (and logits) an order of magnitude up, where bf16's relative
        # rounding no longer stays inside the 5e-3 gate on probabilities
        mask = dataset.genome.windows_with_exceptions(dataset.chrom[lo:hi], dataset.pos[lo:hi], dataset.distal_radius)
        idx = np.flatnonzero(mask)
        if len(idx):
            model.compute_mode = "fp32"
            try:
                with torch.no_grad():
                    for c in range(0, len(idx), 1 << 16):
                        sel = idx[c:c + (1 << 16)] + lo
                        sb = SiteBatch(torch.from_numpy(dataset.pos[sel]).to(dev), torch.from_numpy(dataset.meta[sel]).to(dev), dataset.genome)
                        logp[torch.from_numpy(sel - lo).to(dev)] = model.forward(None, sb)
            finally:
                model.compute_mode = "auto_bf16"
    return logp


def format_predictions(chrom, start, end, strand, mut_type, prob):
    """DataFrame in the reference's output shape (run_predict.py:228-238): sorted by (chrom, start)."""
    import pandas as pd
    cols = {"chrom": chrom, "start": start, "end": end, "strand": strand, "mut_type": np.asarray(mut_type, dtype=np.float64)}
    for i in range(prob.shape[1]):
        cols["prob%d" % i] = prob[:, i]
    df = pd.DataFrame(cols)
    df.sort_values(["chrom", "start"], inplace=True)
    df.reset_index(drop=True, inplace=True)
    return df


def write_tsv(path, chrom_names, start, end, strand, mut_type, prob, n_threads=None):
    """`pred_df.sort_values(['chrom','start']).to_csv(path, sep='\\t', float_format='%.4g', index=False)` (run_predict.py:237-239)
    without pandas: stable sort by (chromosome name, start) here, formatting + writing in the C library (threaded)."""
    import ctypes as C
    import os
    from . import _lib
    names = np.asarray(chrom_names, dtype=object)
    uniq, inv = np.unique(names.astype(str), return_inverse=True)          # lexicographic, like pandas on strings
    order = np.lexsort((np.asarray(start), inv))                            # stable: ties keep emission order
    idx = np.ascontiguousarray(inv[order].astype(np.int32))
    st = np.ascontiguousarray(np.asarray(start, dtype=np.int64)[order])
    en = np.ascontiguousarray(np.asarray(end, dtype=np.int64)[order])
    sd = np.ascontiguousarray(np.asarray([ord(c[0]) for c in ("+", "-")], dtype=np.uint8)[(np.asarray(strand) != "+").astype(np.int64)][order])
    mt = np.ascontiguousarray(np.asarray(mut_type, dtype=np.float64)[order])
    pr = np.ascontiguousarray(np.asarray(prob, dtype=np.float64)[order])
    arr = (C.c_char_p * len(uniq))(*[u.encode() for u in uniq])
    _lib.check(_lib.lib().mural_write_tsv(str(path).encode(), len(st), pr.shape[1], C.cast(arr, C.c_void_p), _lib.ptr(idx), _lib.ptr(st),
                                          _lib.ptr(en), _lib.ptr(sd), _lib.ptr(mt), _lib.ptr(pr), n_threads or min(32, os.cpu_count() or 1)))


def run_predict(test_data, ref_genome, model_path, model_config_path, calibrator_path="", pred_file=None, segment_center=None,
                poisson_calib=False, compute_mode="auto", genome=None, model_type="snv", return_frame=True):
    """Returns the prediction DataFrame on rank 0 (None elsewhere); writes `pred_file` when given."""
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    rank = torch.distributed.get_rank() if dist_on else 0
    world = torch.distributed.get_world_size() if dist_on else 1
    t0 = time.time()
    config = load_config(model_config_path)
    if not segment_center:
        segment_center = config.get("segment_center", 300000)          # run_predict.py:88-92 / commands/predict.py:84
    dev = torch.device("cuda", torch.cuda.current_device())
    if genome is None:
        genome = PackedGenome.from_fasta(ref_genome, dev)
    sites = SiteTable.from_bed(test_data)
    ds = PackedSiteDataset(sites, genome, segment_center, config["local_radius"], config["local_order"], config["distal_radius"],
                           model_type=model_type)
    model = build_model_from_files(model_path, config, dev, model_type=model_type)
    if model_type == "snv":
        model.compute_mode = "auto_bf16" if compute_mode == "auto" else compute_mode
    poisson_calib = bool(poisson_calib) or model_type == "indel"          # run_predict.py:224
    n = len(ds.pos)
    lo, hi = shard_bounds(n, world, rank)
    logp = predict_sites(model, ds, lo, hi)
    weights = load_calibrator_weights(calibrator_path) if calibrator_path else None
    prob = gather_rows(calibrate(logp, weights, poisson_calib), n, world, rank)     # fp64 [n_r, k] on the device, NCCL gather
    if rank != 0:
        return None
    prob = prob.cpu()
    names, start, end, strand = ds.position_info()
    if pred_file:
        write_tsv(pred_file, names, start, end, strand, ds.label, prob.numpy())   # run_predict.py:237-239
    df = format_predictions(names, start, end, strand, ds.label, prob.numpy()) if return_frame else None
    print("predicted %d sites in %.2f s (%d rank(s))" % (n, time.time() - t0, world))
    sys.stdout.flush()
    return df


def run_predict_pipline(args, model_type="snv"):
    """Same entry point / argument names as the reference CLI dispatcher expects (run_predict.py:34)."""
    if model_type not in ("snv", "indel"):
        raise ValueError("model_type must be 'snv' or 'indel'")
    if getattr(args, "cpu_only", False):
        raise RuntimeError("mural_b200 has no CPU path; drop --cpu_only")
    return run_predict(args.test_data, args.ref_genome, args.model_path, args.model_config_path,
                       getattr(args, "calibrator_path", ""), args.pred_file, getattr(args, "segment_center", None),
                       getattr(args, "poisson_calib", False), model_type=model_type)
