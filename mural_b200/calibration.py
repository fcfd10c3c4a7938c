"""Calibration epilogue of prediction: FullDirichlet apply + Poisson calibration on the GPU.

Reference: MuRaL/scripts/run_predict.py:214-225; dirichlet_python/dirichletcal/calib/fulldirichlet.py:78-80,
calib/multinomial.py:60-64,235-244 (predict_proba = softmax([log clip(p), 1] @ W.T), fp64);
MuRaL/model/calibration.py:10-23 (poisson_calibrate).  The calibrator *fit* (JAX Newton) is out of scope; only
the fitted weight matrix (k x (k+1), fp64) stored in model.fdiri_cal.pkl is used.
"""
import io
import pickle

import numpy as np
import torch

from . import _lib


class _Captured:
    """Stand-in for dirichletcal classes when unpickling model.fdiri_cal.pkl (only attributes are kept)."""
    def __setstate__(self, state):
        self.__dict__.update(state)


class _CalUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("dirichletcal"):
            return type(name, (_Captured,), {})
        return super().find_class(module, name)


def load_calibrator_weights(path):
    """-> fp64 [k, k+1] weight matrix of a pickled FullDirichletCalibrator (calibrator_.weights_), or None if
    the pickle holds no fitted calibrator."""
    with open(path, "rb") as f:
        obj = _CalUnpickler(io.BytesIO(f.read())).load()
    cal = getattr(obj, "calibrator_", None)
    w = getattr(cal, "weights_", None) if cal is not None else getattr(obj, "weights_", None)
    return None if w is None else np.ascontiguousarray(np.asarray(w, dtype=np.float64))


def calibrate(logp, weights=None, poisson=False):
    """logp: float32 CUDA tensor [n, k] (Network2 output).  Returns fp64 CUDA tensor [n, k] =
    poisson?(dirichlet?(softmax(logp)))."""
    n, k = logp.shape
    out = torch.empty((n, k), dtype=torch.float64, device=logp.device)
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.shape != (k, k + 1):
            raise ValueError("calibrator weights must be [k, k+1], got %s" % (w.shape,))
    with torch.cuda.device(logp.device):
        _lib.check(_lib.lib().mural_calibrate(_lib.ptr(logp.contiguous()), n, k, _lib.ptr(w), int(bool(poisson)), _lib.ptr(out),
                                              _lib.current_stream()))
    return out
