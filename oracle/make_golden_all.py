"""ORACLE — TEST INFRASTRUCTURE ONLY.  Fixtures for the shipped checkpoints that oracle/make_golden.py does not cover, so that
every model under /root/reference/models (11 MuRaL-snv + 12 MuRaL-indel) has reference outputs: same file format
as make_golden.py (weights, sites on the KAT genome, the UNMODIFIED reference's outputs, calibrator weights), its own seeded
site sets, and the same in-pass pin of the oracle against the reference.  Existing fixtures are left untouched (other
goldens hash their sites).

Run in the build container only:   python -m oracle.make_golden_all
"""
import json
import os
import pickle

import numpy as np
import torch

from oracle import network_t as NT
from oracle import ref_import as R
from oracle.make_golden import GOLD, REF, cal_weights, clean_state, make_kat_genome, oracle_encode, pick_sites

SNV_MORE = {
    "at_CpG": "models/Arabidopsis_thaliana/SNV/CpG", "at_nonCpG": "models/Arabidopsis_thaliana/SNV/nonCpG",
    "dm_AT": "models/Drosophila_melanogaster/SNV/AT", "mm_CpG": "models/Macaca_mulatta/SNV/CpG",
    "mm_nonCpG": "models/Macaca_mulatta/SNV/nonCpG",
}
INDEL_MORE = {"%s_%s" % (sp, k): "models/%s/INDEL/%s" % (full, d)
              for sp, full in (("hs", "Homo_sapiens"), ("at", "Arabidopsis_thaliana"), ("dm", "Drosophila_melanogaster"), ("mm", "Macaca_mulatta"))
              for k, d in (("ins", "insertion"), ("del_start", "deletion_start"), ("del_end", "deletion_end"))
              if (sp, k) not in (("hs", "ins"), ("hs", "del_start"))}


def main():
    assert R.available(), "reference tree not mounted"
    pre, snv, indel, nnu = R.import_reference()
    manifest = json.load(open(os.path.join(GOLD, "MANIFEST.json")))
    torch.manual_seed(0)
    torch.set_num_threads(8)
    genome = make_kat_genome()
    rng = np.random.default_rng(2024)
    n_eval = 192
    for tag, rel in SNV_MORE.items():
        cfg = pickle.load(open(os.path.join(REF, rel, "model.config.pkl"), "rb"))
        common = dict(emb_dims=cfg["emb_dims"], n_cont=0, n_class=cfg["n_class"], distal_order=1, in_channels=4)
        model = nnu.model_choice(cfg["model_no"], cfg, common, "snv")
        sd = torch.load(os.path.join(REF, rel, "model"), map_location="cpu")
        model.load_state_dict(sd)
        model.eval()
        manifest.setdefault("state_dict_keys", {})[tag] = [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()]
        bp, bm = ("A", "T") if "AT" in tag else ("C", "G")
        ch, stt, sd_ = pick_sites(genome, rng, per_chrom=n_eval // 3, base_plus=bp, base_minus=bm)
        central = int(cfg.get("segment_center", 300000))
        perm, sizes, cat, oh = oracle_encode(genome, ch, stt, sd_, central, cfg["local_radius"], cfg["local_order"], cfg["distal_radius"], "snv")
        with torch.no_grad():
            ref_lp = model.forward((torch.zeros(len(perm), 1), torch.from_numpy(cat)), torch.from_numpy(oh)).numpy()
            state = clean_state(sd)
            o32 = NT.network2_forward(state, cat, oh, torch.float32).numpy()
        d32 = float(np.abs(o32 - ref_lp).max())
        assert d32 < 2e-5, (tag, d32)
        calw = cal_weights(os.path.join(REF, rel, "model.fdiri_cal.pkl"))
        prob = torch.softmax(torch.from_numpy(ref_lp), 1).numpy()
        out = {"cfg_json": np.array(json.dumps({k: (v if not isinstance(v, list) else None) for k, v in cfg.items() if k != "emb_dims"}, default=float)),
               "n_cat": np.array(len(cfg["emb_dims"])), "chrom": ch[perm], "start": stt[perm], "strand": sd_[perm],
               "ref_logp": ref_lp, "cal_weights": calw, "cal_prob": NT.dirichlet_apply(calw, prob)}
        for k, v in state.items():
            out["w:" + k] = v
        np.savez_compressed(os.path.join(GOLD, "snv_%s.npz" % tag), **out)
        manifest["snv"][tag] = {"checkpoint": rel, "n_sites": int(len(perm)), "oracle_fp32_vs_ref_maxabs": d32,
                                "distal_radius": int(cfg["distal_radius"]), "local_radius": int(cfg["local_radius"]), "generator": "oracle/make_golden_all.py"}
        print("snv", tag, "ref vs oracle fp32 %.2e" % d32, "R_d", cfg["distal_radius"], "R_l", cfg["local_radius"], "p0 mean %.3f" % prob[:, 0].mean())
    for tag, rel in INDEL_MORE.items():
        cfg = pickle.load(open(os.path.join(REF, rel, "model.config.pkl"), "rb"))
        sd = torch.load(os.path.join(REF, rel, "model"), map_location="cpu")
        use_rev = any(k.startswith("conv.0") for k in sd)
        n_class = sd["out_fc.2.weight"].shape[0]
        ch8, ks = sd["uplblocks.0.0.weight"].shape[0], sd["uplblocks.0.0.weight"].shape[2]
        down = cfg["down_list"]
        model = indel.UNet_Small(n_class, ch8, ks, down, use_reverse=use_rev)
        model.load_state_dict(sd)
        model.eval()
        manifest.setdefault("state_dict_keys", {})[tag] = [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()]
        Rd = int(cfg["distal_radius"])
        ch, stt, sd_ = pick_sites(genome, rng, per_chrom=4)
        sd_[:] = 0
        perm, sizes, cat, oh = oracle_encode(genome, ch, stt, sd_, 300000, cfg["local_radius"], cfg["local_order"], Rd, "indel")
        with torch.no_grad():
            ref_o = model.forward(torch.from_numpy(oh)).numpy()
            state = clean_state(sd)
            o32 = NT.unet_small_forward(state, oh, down, use_rev, torch.float32).numpy()
        d32 = float(np.abs(o32 - ref_o).max())
        assert d32 < 1e-4 * max(1.0, float(np.abs(ref_o).max())), (tag, d32)
        out = {"down": np.array(down), "use_reverse": np.array(use_rev), "distal_radius": np.array(Rd),
               "chrom": ch[perm], "start": stt[perm], "strand": sd_[perm], "ref_out": ref_o}
        for k, v in state.items():
            out["w:" + k] = v
        np.savez_compressed(os.path.join(GOLD, "indel_%s.npz" % tag), **out)
        manifest["indel"][tag] = {"checkpoint": rel, "n_sites": int(len(perm)), "oracle_fp32_vs_ref_maxabs": d32, "use_reverse": bool(use_rev),
                                  "n_class": int(n_class), "distal_radius": Rd, "down": [int(v) for v in down], "generator": "oracle/make_golden_all.py"}
        print("indel", tag, "ref vs oracle fp32 %.2e" % d32, "R", Rd, "down", down, "rev", use_rev, "n_class", n_class)
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
