"""mural_b200 — B200-native implementation of MuRaL's per-site predict/train hot path.

Layout: csrc/ (CUDA kernels + C ABI, include/mural_b200.h), genome.py (packed genome), data.py
(site records in the reference's sample order), model_snv.py / model_indel.py (drop-in model
classes), nn_utils.py (model_choice, model_predict_m), training.py (train step), predict.py
(genome-wide prediction, interval sharding, calibration, TSV output).
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from .data import (PackedSiteDataset, ReferenceTupleDataset, SiteBatch, SiteTable, generate_data_batches,  # noqa: F401
                   generate_site_batches, get_local_header, pack_meta, prepare_dataset_np, segment_order)
from .genome import PackedGenome, read_fasta  # noqa: F401
from .model_snv import Network2  # noqa: F401
from .nn_utils import model_choice, model_predict_m, weights_init  # noqa: F401
