"""model_choice / model_predict_m / weights_init with the reference's signatures
(MuRaL/model/nn_utils.py:14-76, 186-231) on top of the B200 kernels."""
import inspect
import sys
import time

import torch
import torch.nn as nn

from . import _lib
from .data import SiteBatch
from .model_snv import Network2


def weights_init(m):
    """Same initialisation rules as the reference (nn_utils.py:14-35): xavier-uniform conv weights,
    kaiming-normal linear weights, zero biases."""
    name = m.__class__.__name__
    if "Conv1d" in name or "Conv2d" in name:
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif "Linear" in name:
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)


def _registry():
    reg = {"snv": {2: Network2}, "indel": {}}
    try:
        from .model_indel import UNet_Small
        reg["indel"][0] = UNet_Small
    except ImportError:
        pass
    return reg


def model_choice(model_no, config, common_model_config, model_type):
    """Factory with the reference's name-based constructor-argument resolution (nn_utils.py:186-231).
    Only the models on the hot path are registered: snv -> {2: Network2}, indel -> {0: UNet_Small}."""
    model_config = {**config, **common_model_config}
    if model_type == "snv":
        rules = {"lin_layer_sizes": [config["local_hidden1_size"], config["local_hidden2_size"]],
                 "lin_layer_dropouts": [config["local_dropout"], config["local_dropout"]],
                 "emb_padding_idx": 4 ** config["local_order"],
                 "out_channels": config["CNN_out_channels"], "kernel_size": config["CNN_kernel_size"],
                 "no_of_cont": common_model_config["n_cont"]}
    else:
        rules = {"out_channels": config["CNN_out_channels"], "kernel_size": config["CNN_kernel_size"],
                 "downsize": config["down_list"], "use_reverse": config.get("use_reverse", False)}
    reg = _registry()
    if model_type not in reg:
        raise ValueError(f"model_type must be one of {list(reg.keys())}, got {model_type}")
    cls = reg[model_type].get(model_no)
    if cls is None:
        raise ValueError(f"model_no for {model_type} must be one of {list(reg[model_type].keys())}, got {model_no}")
    names = [p for p in inspect.signature(cls.__init__).parameters if p != "self"]
    return cls(**{p: (rules[p] if p in rules else model_config[p]) for p in names})


def model_predict_m(model, dataloader, criterion, device, n_class, distal=True, model_type="snv"):
    """Prediction loop (nn_utils.py:37-76).  Accepts the reference's (y, cont_x, cat_x, distal_x) tuples or
    SiteBatch-es.  Differences that do not change results: outputs are collected and concatenated once
    (the reference re-allocates pred_y every batch), and the summed cross-entropy is accumulated on the
    device and read back once (the reference syncs with loss.item() every batch)."""
    model.to(device)
    model.eval()
    outs = []
    loss_dev = torch.zeros(1, dtype=torch.float64, device=device)
    total_loss_host = 0.0
    batch_count = 0
    t0 = time.time()
    L = _lib.lib()
    with torch.no_grad():
        for item in dataloader:
            batch_count += 1
            if isinstance(item, SiteBatch):
                if model_type == "indel":                            # UNet_Small.forward(distal_x) (nn_utils.py:59-61)
                    preds = model.forward(item, distal_radius=getattr(dataloader, "distal_radius", None) or getattr(model, "_hR", None))
                else:
                    preds = model.forward(None, item)
                with torch.cuda.device(preds.device):
                    _lib.check(L.mural_ce_sum(_lib.ptr(preds), _lib.ptr(item.meta), len(item), n_class, _lib.ptr(loss_dev),
                                              _lib.current_stream()))
            else:
                y, cont_x, cat_x, distal_x = item
                cat_x, cont_x, distal_x, y = cat_x.to(device), cont_x.to(device), distal_x.to(device), y.to(device)
                preds = model.forward((cont_x, cat_x), distal_x) if model_type == "snv" else model.forward(distal_x)
                if criterion is not None:
                    total_loss_host += float(criterion(preds, y.long().squeeze(1)).item())
            outs.append(preds)
    pred_y = torch.cat(outs, dim=0) if outs else torch.empty(0, n_class, device=device)
    total_loss = total_loss_host + float(loss_dev.item())
    print(f"Batch Number: {batch_count}; prediction Time of {batch_count} batch: {(time.time() - t0) / 60} min")
    sys.stdout.flush()
    return pred_y, total_loss
