import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def manifest():
    with open(os.path.join(GOLD, "MANIFEST.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kat():
    """Encoder known-answer fixture generated from the real reference (oracle/make_golden.py)."""
    z = np.load(os.path.join(GOLD, "encode_kat.npz"))
    genome = {str(n): str(s) for n, s in zip(z["genome_names"], z["genome_seqs"])}
    return z, genome


def load_snv_golden(tag):
    z = np.load(os.path.join(GOLD, "snv_%s.npz" % tag))
    cfg = json.loads(str(z["cfg_json"]))
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    return z, cfg, state


# every checkpoint shipped under /root/reference/models (11 MuRaL-snv, 12 MuRaL-indel) plus the two example checkpoints
# (oracle/make_golden.py wrote the first group of each list, oracle/make_golden_all.py the rest)
SNV_TAGS = ["hs_AT", "hs_CpG", "hs_nonCpG", "mm_AT", "dm_CG", "at_AT", "ex_ckpt6",
            "at_CpG", "at_nonCpG", "dm_AT", "mm_CpG", "mm_nonCpG"]
INDEL_TAGS = ["hs_ins", "hs_del_start", "ex_indel9",
              "hs_del_end", "at_ins", "at_del_start", "at_del_end", "dm_ins", "dm_del_start", "dm_del_end", "mm_ins", "mm_del_start",
              "mm_del_end"]


@pytest.fixture(scope="session")
def cuda_genome(kat):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mural_b200 import PackedGenome
    return PackedGenome(kat[1])
