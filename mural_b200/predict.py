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


def write_tsv(path, chrom_names, start, end, strand, mut_type, prob, n_threads=None, chrom_codes=None):
    """`pred_df.sort_values(['chrom','start']).to_csv(path, sep='\\t', float_format='%.4g', index=False)` (run_predict.py:237-239)
    without pandas: stable sort by (chromosome name, start) here, formatting + writing in the C library (threaded).
    `chrom_names`: per-site names, or — with `chrom_codes` (int per site) — the list the codes index (no per-site strings)."""
    import ctypes as C
    import os
    from . import _lib
    if chrom_codes is None:
        uniq, inv = np.unique(np.asarray(chrom_names, dtype=object).astype(str), return_inverse=True)   # lexicographic, like pandas on strings
    else:
        names = np.asarray([str(c) for c in chrom_names])
        uniq, back = np.unique(names, return_inverse=True)
        inv = back[np.asarray(chrom_codes, dtype=np.int64)]
    start = np.asarray(start, dtype=np.int64)
    order = np.lexsort((start, inv))                                         # stable: ties keep emission order
    idx = np.ascontiguousarray(inv[order].astype(np.int32))
    st = np.ascontiguousarray(start[order])
    en = np.ascontiguousarray(np.asarray(end, dtype=np.int64)[order])
    strand = np.asarray(strand)
    minus = (strand != 0) if strand.dtype.kind in "iub" else (strand != "+")
    sd = np.ascontiguousarray(np.where(minus, np.uint8(ord("-")), np.uint8(ord("+")))[order].astype(np.uint8))
    mt = np.ascontiguousarray(np.asarray(mut_type, dtype=np.float64)[order])
    pr = np.ascontiguousarray(np.asarray(prob, dtype=np.float64)[order])
    arr = (C.c_char_p * len(uniq))(*[u.encode() for u in uniq])
    _lib.check(_lib.lib().mural_write_tsv(str(path).encode(), len(st), pr.shape[1], C.cast(arr, C.c_void_p), _lib.ptr(idx), _lib.ptr(st),
                                          _lib.ptr(en), _lib.ptr(sd), _lib.ptr(mt), _lib.ptr(pr), n_threads or min(32, os.cpu_count() or 1)))


def run_predict(test_data, ref_genome, model_path, model_config_path, calibrator_path="", pred_file=None, segment_center=None,
                poisson_calib=False, compute_mode="auto", genome=None, model_type="snv", return_frame=True, timings=None):
    """Returns the prediction DataFrame on rank 0 (None elsewhere); writes `pred_file` when given.  `timings` (a dict) receives
    the wall seconds of every stage (device work synchronised at the stage boundaries only when it is passed)."""
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    rank = torch.distributed.get_rank() if dist_on else 0
    world = torch.distributed.get_world_size() if dist_on else 1
    t0 = t_last = time.time()

    def lap(name):
        nonlocal t_last
        if timings is not None:
            torch.cuda.synchronize()
            now = time.time()
            timings[name] = timings.get(name, 0.0) + now - t_last
            t_last = now
    config = load_config(model_config_path)
    if not segment_center:
        segment_center = config.get("segment_center", 300000)          # run_predict.py:88-92 / commands/predict.py:84
    dev = torch.device("cuda", torch.cuda.current_device())
    if genome is None:
        genome = PackedGenome.from_fasta(ref_genome, dev)
    lap("fasta_ingest_s")
    sites = SiteTable.from_bed(test_data)
    lap("bed_ingest_s")
    ds = PackedSiteDataset(sites, genome, segment_center, config["local_radius"], config["local_order"], config["distal_radius"],
                           model_type=model_type)
    lap("site_order_s")
    model = build_model_from_files(model_path, config, dev, model_type=model_type)
    if model_type == "snv":
        model.compute_mode = compute_mode       # "auto" (MURAL_MODE_AUTO: the exception-window policy runs inside the library)
    lap("model_load_s")
    poisson_calib = bool(poisson_calib) or model_type == "indel"          # run_predict.py:224
    n = len(ds.pos)
    lo, hi = shard_bounds(n, world, rank)
    logp = predict_sites(model, ds, lo, hi)
    lap("compute_s")
    weights = load_calibrator_weights(calibrator_path) if calibrator_path else None
    prob = calibrate(logp, weights, poisson_calib)                         # fp64 [n_r, k] on the device
    lap("calibrate_s")
    prob = gather_rows(prob, n, world, rank)                               # one NCCL gather
    lap("gather_s")
    if rank != 0:
        return None
    prob = prob.cpu()
    if pred_file:                                                              # run_predict.py:237-239
        write_tsv(pred_file, ds.sites.chrom_names, ds.sites.start[ds.perm], ds.sites.end[ds.perm], ds.strand, ds.label, prob.numpy(),
                  chrom_codes=ds.sites.chrom[ds.perm])
    lap("tsv_write_s")
    df = format_predictions(*ds.position_info(), ds.label, prob.numpy()) if return_frame else None
    if timings is not None:
        timings["total_s"] = time.time() - t0
        timings["sites"] = n
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
