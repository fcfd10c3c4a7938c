"""ORACLE — TEST INFRASTRUCTURE ONLY.  Generates tests/golden/* by running the UNMODIFIED reference
(imported from /root/reference, see oracle/ref_import.py) and, in the same pass, pins the restatement
in oracle/encode_np.py and oracle/network_t.py against it (asserts; the deviations are written to
tests/golden/MANIFEST.json).

Run in the build container only:   python -m oracle.make_golden
The GPU box never runs this (no /root/reference there); tests read the committed fixtures.
"""
import json
import os
import pickle
import sys

import numpy as np
import torch

from oracle import encode_np as E
from oracle import network_t as NT
from oracle import ref_import as R

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF = R.REF_ROOT


# ---------------------------------------------------------------------------------------------
def make_kat_genome(seed=7):
    """Small 3-chromosome genome exercising every encoder edge the reference handles:
    lowercase, N runs (interior + chromosome ends), every IUPAC code, a tiny chromosome shorter than
    one expanded window."""
    rng = np.random.default_rng(seed)
    def rnd(n):
        return "".join(np.array(list("ACGT"))[rng.integers(0, 4, n)])
    c1 = list(rnd(30000))
    for s, l in ((0, 300), (5000, 700), (12000, 1), (29800, 200)):      # N runs incl. both ends
        c1[s:s + l] = "N" * l
    for i, ch in enumerate("RYMSWKBDHVN"):                              # one of each IUPAC code
        c1[7000 + 37 * i] = ch
        c1[7003 + 37 * i] = ch.lower()
    c1[9000:9800] = [c.lower() for c in c1[9000:9800]]                  # soft-masked block
    c2 = list(rnd(9000))
    c2[100:140] = "N" * 40
    c3 = list(rnd(1500))                                                # shorter than 2001 window
    return {"chrA": "".join(c1), "chrB": "".join(c2), "chrC": "".join(c3)}


def pick_sites(genome, rng, per_chrom, base_plus="A", base_minus="T", dense_every=None):
    """Sorted BED-like site table whose focal base is base_plus on '+' and base_minus on '-'
    (the reference aborts otherwise: preprocessing.py:479-484)."""
    chroms, starts, strands = [], [], []
    for ci, (name, seq) in enumerate(genome.items()):
        up = np.frombuffer(seq.upper().encode(), dtype=np.uint8)
        cand_p = np.flatnonzero(up == ord(base_plus))
        cand_m = np.flatnonzero(up == ord(base_minus))
        n = min(per_chrom, len(cand_p), len(cand_m))
        p = rng.choice(cand_p, n, replace=False)
        m = rng.choice(cand_m, n, replace=False)
        # force chromosome-edge sites (left/right imputation paths)
        p = np.union1d(p, cand_p[[0, 1, -1]]); m = np.union1d(m, cand_m[[0, -2, -1]])
        s = np.r_[p, m]; st = np.r_[np.zeros(len(p), int), np.ones(len(m), int)]
        o = np.argsort(s, kind="stable")
        chroms += [ci] * len(s); starts += list(s[o]); strands += list(st[o])
    return np.array(chroms), np.array(starts, dtype=np.int64), np.array(strands, dtype=np.int64)


def to_regions(names, chroms, starts, strands, labels=None):
    bt = sys.modules["pybedtools"].BedTool()
    for i in range(len(starts)):
        bt.append(R.Region(names[chroms[i]], starts[i], starts[i] + 1, "+-"[strands[i]],
                           0 if labels is None else int(labels[i])))
    return bt


def ref_encode(pre, genome, bt, central, R_l, order, R_d, model_type):
    recs = {k: R.SeqRec(v) for k, v in genome.items()}
    dig, y = pre.local_digitalized_seqs_by_region(bt, recs, central, R_l, local_order=order, model_type=model_type)
    cat = np.concatenate([d.values for d in dig]).astype(np.int64)
    seqs_list, shapes = pre.get_distal_seqs_by_region(bt, recs, R_d, central, model_type)
    oh = np.concatenate([pre.distal_encoding_by_region(iter(s), n, R_d, recs, model_type=model_type)
                         for s, n in zip(seqs_list, shapes)])
    return cat, oh, np.array(shapes)


def oracle_encode(genome, chroms, starts, strands, central, R_l, order, R_d, model_type):
    names = list(genome)
    perm, sizes = E.order_sites(chroms, starts, strands, central)
    syms = [E.seq_to_symbols(genome[n]) for n in names]
    cat = np.empty((len(perm), E.window_length(R_l, model_type) - (order - 1)), dtype=np.int64)
    oh = np.empty((len(perm), 4, E.window_length(R_d, model_type)), dtype=np.float32)
    c_p, s_p, st_p = chroms[perm], starts[perm], strands[perm]
    for ci in range(len(names)):
        m = c_p == ci
        cat[m] = E.kmer_windows(syms[ci], s_p[m], st_p[m], R_l, order, model_type)
        oh[m] = E.onehot_windows(syms[ci], s_p[m], st_p[m], R_d, model_type)
    return perm, sizes, cat, oh


# ---------------------------------------------------------------------------------------------
SNV_CKPTS = {
    "hs_AT": "models/Homo_sapiens/SNV/AT", "hs_CpG": "models/Homo_sapiens/SNV/CpG",
    "hs_nonCpG": "models/Homo_sapiens/SNV/nonCpG", "mm_AT": "models/Macaca_mulatta/SNV/AT",
    "dm_CG": "models/Drosophila_melanogaster/SNV/CG", "at_AT": "models/Arabidopsis_thaliana/SNV/AT",
    "ex_ckpt6": "examples/snv/models/checkpoint_6",
}
INDEL_CKPTS = {
    "hs_ins": "models/Homo_sapiens/INDEL/insertion", "hs_del_start": "models/Homo_sapiens/INDEL/deletion_start",
    "ex_indel9": "examples/indel/models/checkpoint_9",
}


def clean_state(sd):
    """Drop the aliased '.layer.N.' duplicates (model_snv.py:799-804); keep everything else as numpy."""
    return {k: v.detach().cpu().numpy() for k, v in sd.items() if ".layer." not in k}


def cal_weights(path):
    R.install_stubs()
    with open(path, "rb") as f:
        c = pickle.load(f)
    return np.asarray(c.calibrator_.weights_, dtype=np.float64)


def main():
    assert R.available(), "reference tree not mounted"
    pre, snv, indel, nnu = R.import_reference()
    os.makedirs(GOLD, exist_ok=True)
    manifest = {"generator": "oracle/make_golden.py", "reference": "CaiLiLab/MuRaL v1.2.0 @ /root/reference",
                "torch": torch.__version__, "encode": {}, "snv": {}, "indel": {}}
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---------------- (i) encoder KATs
    genome = make_kat_genome()
    names = list(genome)
    rng = np.random.default_rng(11)
    chroms, starts, strands = pick_sites(genome, rng, per_chrom=160)
    cases = [("snv", 7, 3, 1000, 5000), ("snv", 10, 3, 100, 300000), ("snv", 7, 1, 200, 2000),
             ("snv", 10, 1, 150, 777), ("indel", 5, 3, 400, 5000), ("indel", 5, 1, 130, 300000)]
    enc_out = {"genome_names": np.array(names), "genome_seqs": np.array([genome[n] for n in names]),
               "chrom": chroms, "start": starts, "strand": strands}
    for ci, (mt, R_l, order, R_d, central) in enumerate(cases):
        bt = to_regions(names, chroms, starts, strands)
        if mt == "indel":            # reference indel pipeline is '+'-only in its examples but the code is generic
            pass
        cat_ref, oh_ref, shapes = ref_encode(pre, genome, bt, central, R_l, order, R_d, mt)
        perm, sizes, cat_o, oh_o = oracle_encode(genome, chroms, starts, strands, central, R_l, order, R_d, mt)
        assert np.array_equal(sizes, shapes), (ci, sizes[:5], shapes[:5])
        # literal state machine == vectorised order
        lit = E.bed_batches(chroms, starts, strands, central)
        assert np.array_equal(np.concatenate([np.array(b[0]) for b in lit]), perm)
        assert np.array_equal(cat_ref, cat_o), ("kmer mismatch", ci)
        assert oh_ref.dtype == np.float32 and np.array_equal(oh_ref.view(np.uint32), oh_o.view(np.uint32)), ("onehot mismatch", ci)
        # store: k-mer table in full; one-hot as packed per-position symbol planes would lose the
        # float bits, so keep the raw fp32 for a strided subset of sites (compresses ~100x)
        sub = np.arange(0, len(perm), max(1, len(perm) // 48))
        enc_out["case%d_cfg" % ci] = np.array([0 if mt == "snv" else 1, R_l, order, R_d, central])
        enc_out["case%d_perm" % ci] = perm
        enc_out["case%d_sizes" % ci] = sizes
        enc_out["case%d_cat" % ci] = cat_ref
        enc_out["case%d_oh_rows" % ci] = sub
        enc_out["case%d_oh" % ci] = oh_ref[sub]
        manifest["encode"]["case%d" % ci] = {"model_type": mt, "local_radius": R_l, "local_order": order,
                                             "distal_radius": R_d, "segment_center": central,
                                             "n_sites": int(len(perm)), "kmer_equal": True, "onehot_bits_equal": True}
    np.savez_compressed(os.path.join(GOLD, "encode_kat.npz"), **enc_out)
    print("encoder KATs: reference == oracle on", len(cases), "cases,", len(starts), "sites")

    # ---------------- (ii)/(iii) SNV checkpoints: weights + logits on KAT-genome sites
    rng = np.random.default_rng(21)
    n_eval = 192
    for tag, rel in SNV_CKPTS.items():
        cfg = pickle.load(open(os.path.join(REF, rel, "model.config.pkl"), "rb"))
        common = dict(emb_dims=cfg["emb_dims"], n_cont=0, n_class=cfg["n_class"], distal_order=1, in_channels=4)
        model = nnu.model_choice(cfg["model_no"], cfg, common, "snv")
        sd = torch.load(os.path.join(REF, rel, "model"), map_location="cpu")
        model.load_state_dict(sd)
        model.eval()
        manifest.setdefault("state_dict_keys", {})[tag] = [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()]
        bp, bm = ("A", "T") if "AT" in tag or tag.startswith("ex") else ("C", "G")
        ch, stt, sd_ = pick_sites(genome, rng, per_chrom=n_eval // 3, base_plus=bp, base_minus=bm)
        central = int(cfg.get("segment_center", 300000))
        perm, sizes, cat, oh = oracle_encode(genome, ch, stt, sd_, central, cfg["local_radius"], cfg["local_order"],
                                             cfg["distal_radius"], "snv")
        with torch.no_grad():
            ref_lp = model.forward((torch.zeros(len(perm), 1), torch.from_numpy(cat)), torch.from_numpy(oh)).numpy()
            state = clean_state(sd)
            taps = {}
            o32 = NT.network2_forward(state, cat, oh, torch.float32, taps=taps).numpy()
            o64 = NT.network2_forward(state, cat, oh, torch.float64).numpy()
        d32 = float(np.abs(o32 - ref_lp).max()); d64 = float(np.abs(o64 - ref_lp).max())
        assert d32 < 2e-5, (tag, d32)
        calw = cal_weights(os.path.join(REF, rel, "model.fdiri_cal.pkl"))
        prob = torch.softmax(torch.from_numpy(ref_lp), 1).numpy()
        out = {"cfg_json": np.array(json.dumps({k: (v if not isinstance(v, list) else None) for k, v in cfg.items()
                                                 if k != "emb_dims"}, default=float)),
               "n_cat": np.array(len(cfg["emb_dims"])),
               "chrom": ch[perm], "start": stt[perm], "strand": sd_[perm],
               "ref_logp": ref_lp, "oracle64_logp": o64.astype(np.float64),
               "cal_weights": calw, "cal_prob": NT.dirichlet_apply(calw, prob)}
        for k in ("pool1", "pool1_2", "rb1_2", "conv2_2", "rb2_2", "gmax", "gmax_2", "logit_local", "logit_mid", "logit_large"):
            out["tap_" + k] = taps[k].numpy()[:16]
        for k, v in state.items():
            out["w:" + k] = v
        np.savez_compressed(os.path.join(GOLD, "snv_%s.npz" % tag), **out)
        manifest["snv"][tag] = {"checkpoint": rel, "n_sites": int(len(perm)), "oracle_fp32_vs_ref_maxabs": d32,
                                "oracle_fp64_vs_ref_maxabs": d64, "distal_radius": int(cfg["distal_radius"]),
                                "local_radius": int(cfg["local_radius"])}
        print("snv", tag, "ref vs oracle fp32 %.2e fp64 %.2e" % (d32, d64), "p0 mean %.3f" % prob[:, 0].mean())

    # ---------------- INDEL checkpoints
    for tag, rel in INDEL_CKPTS.items():
        cfgp = os.path.join(REF, rel, "model.config.pkl")
        if os.path.exists(cfgp):
            cfg = pickle.load(open(cfgp, "rb"))
        else:   # examples/indel ships no config pkl; same hyper-parameters as examples.sh -> use human insertion cfg
            cfg = pickle.load(open(os.path.join(REF, "models/Homo_sapiens/INDEL/insertion/model.config.pkl"), "rb"))
        sd = torch.load(os.path.join(REF, rel, "model"), map_location="cpu")
        use_rev = any(k.startswith("conv.0") for k in sd)
        n_class = sd["out_fc.2.weight"].shape[0]
        ch8 = sd["uplblocks.0.0.weight"].shape[0]
        ks = sd["uplblocks.0.0.weight"].shape[2]
        down = cfg["down_list"]
        model = indel.UNet_Small(n_class, ch8, ks, down, use_reverse=use_rev)
        model.load_state_dict(sd); model.eval()
        manifest.setdefault("state_dict_keys", {})[tag] = [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()]
        Rd = int(cfg["distal_radius"])
        ch, stt, sd_ = pick_sites(genome, rng, per_chrom=8)
        sd_[:] = 0                                   # reference indel data are '+' only
        perm, sizes, cat, oh = oracle_encode(genome, ch, stt, sd_, 300000, cfg["local_radius"], cfg["local_order"], Rd, "indel")
        with torch.no_grad():
            ref_o = model.forward(torch.from_numpy(oh)).numpy()
            state = clean_state(sd)
            o32 = NT.unet_small_forward(state, oh, down, use_rev, torch.float32).numpy()
        d32 = float(np.abs(o32 - ref_o).max())
        assert d32 < 1e-4, (tag, d32)
        out = {"down": np.array(down), "use_reverse": np.array(use_rev), "distal_radius": np.array(Rd),
               "chrom": ch[perm], "start": stt[perm], "strand": sd_[perm], "ref_out": ref_o}
        for k, v in state.items():
            out["w:" + k] = v
        np.savez_compressed(os.path.join(GOLD, "indel_%s.npz" % tag), **out)
        manifest["indel"][tag] = {"checkpoint": rel, "n_sites": int(len(perm)), "oracle_fp32_vs_ref_maxabs": d32,
                                  "use_reverse": bool(use_rev), "n_class": int(n_class)}
        print("indel", tag, "ref vs oracle fp32 %.2e" % d32)

    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
