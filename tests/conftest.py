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


SNV_TAGS = ["hs_AT", "hs_CpG", "hs_nonCpG", "mm_AT", "dm_CG", "at_AT", "ex_ckpt6"]
INDEL_TAGS = ["hs_ins", "hs_del_start", "ex_indel9"]


@pytest.fixture(scope="session")
def cuda_genome(kat):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mural_b200 import PackedGenome
    return PackedGenome(kat[1])
