"""Records sha256 of the bf16-path log-probs on fixed dense / sparse site sets (tests/golden/bf16_bits.json).
Run on a GPU box with the build whose arithmetic is the reference point; tests/test_gpu_snv_tc.py::test_bf16_bits_unchanged
then holds every later kernel restructuring to bit-identical outputs."""
import hashlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLD, load_snv_golden
from test_gpu_snv_forward import build_model
from mural_b200 import PackedGenome, SiteBatch, pack_meta

def site_sets(genome):
    names = list(genome)
    out = {}
    rng = np.random.default_rng(21)
    n = 7000
    st = np.sort(rng.integers(0, 30000, n)).astype(np.int32); sd = rng.integers(0, 2, n)
    out["dense"] = (st, sd, np.zeros(n, np.int64))
    ch = rng.integers(0, 2, n)
    st = np.array([rng.integers(0, len(genome[names[c]])) for c in ch]).astype(np.int32); sd = rng.integers(0, 2, n)
    out["sparse"] = (st, sd, ch)
    return out

def main():
    k = np.load(os.path.join(GOLD, "encode_kat.npz"))
    genome = {str(n): str(s) for n, s in zip(k["genome_names"], k["genome_seqs"])}
    pg = PackedGenome(genome)
    res = {}
    for tag in ("hs_AT", "ex_ckpt6"):
        z, cfg, state = load_snv_golden(tag)
        m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
        for name, (st, sd, ch) in site_sets(genome).items():
            sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, ch)).cuda(), pg)
            with torch.no_grad():
                lp = m.forward(None, sb).cpu().numpy()
            res["%s/%s" % (tag, name)] = {"sha256": hashlib.sha256(lp.tobytes()).hexdigest(), "head": lp[:4].astype(float).tolist()}
    json.dump(res, open(os.path.join(GOLD, "bf16_bits.json") if len(sys.argv) < 2 else sys.argv[1], "w"), indent=1)
    print(json.dumps({k: v["sha256"][:16] for k, v in res.items()}))
main()
