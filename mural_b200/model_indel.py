"""Drop-in MuRaL-indel `UNet_Small` (MuRaL/model/model_indel.py:21-176) on the B200 kernels (eval forward).

Same constructor signature and `state_dict()` keys as the reference; the torch modules only hold parameters.
`forward(distal_input)` accepts the reference's one-hot tensor [B,4,2R] or a `SiteBatch` (fast path: windows are
gathered on the GPU from the packed genome).  In train() mode forward is differentiable and runs the native training tape
(csrc/indel_train.cu through mural_b200.training.IndelTrainState).
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .data import SiteBatch


class ConvBlock(nn.Module):
    def __init__(self, inp, oup, expand_ratio=2, fused=True):
        super().__init__()
        hidden = round(inp * expand_ratio)
        self.conv = nn.Sequential(nn.Conv1d(inp, hidden, 5, 1, padding=2, bias=False), nn.BatchNorm1d(hidden), nn.SiLU(inplace=False),
                                  nn.Conv1d(hidden, oup, 1, 1, 0, bias=False), nn.BatchNorm1d(oup))


class UNet_Small(nn.Module):
    def __init__(self, n_class, out_channels, kernel_size, downsize, use_reverse=None):
        super().__init__()
        self.use_reverse = use_reverse
        self.n_class = n_class
        ks, pad = kernel_size, (kernel_size - 1) // 2
        if use_reverse:
            self.conv = nn.Sequential(nn.Conv1d(4, 4, kernel_size=ks, padding=pad), nn.BatchNorm1d(4))
        ch = self.channels = [out_channels * (i + 1) for i in range(6)]
        cin = [4] + ch[:-1]
        self.uplblocks = nn.ModuleList([nn.Sequential(nn.Conv1d(cin[i], ch[i], stride=downsize[i], kernel_size=ks, padding=pad),
                                                      nn.BatchNorm1d(ch[i])) for i in range(6)])
        self.upblocks = nn.ModuleList([nn.Sequential(ConvBlock(ch[i], ch[i])) for i in range(6)])
        self.downlblocks = nn.ModuleList([nn.Sequential(nn.Upsample(scale_factor=downsize[5 - i]),
                                                        nn.Conv1d(ch[5 - i], ch[4 - i], kernel_size=ks, padding=pad),
                                                        nn.BatchNorm1d(ch[4 - i])) for i in range(5)])
        self.downblocks = nn.ModuleList([nn.Sequential(ConvBlock(ch[4 - i], ch[4 - i])) for i in range(5)])
        self.out_conv = nn.Sequential(nn.Conv1d(ch[0], ch[0], kernel_size=1), nn.BatchNorm1d(ch[0]), nn.ReLU(inplace=True),
                                      nn.Conv1d(ch[0], ch[0], kernel_size=1), nn.Softplus())
        self.out_fc = nn.Sequential(nn.BatchNorm1d(ch[0]), nn.Dropout(0.1), nn.Linear(ch[0], n_class), nn.Softplus())
        self._ks, self._down, self._C = kernel_size, list(downsize), out_channels
        self._h, self._hR, self._dirty = None, None, True
        # "auto": fused tensor-core level kernels (fp32-equivalent split-bf16 products) when the shapes allow it,
        # "fp32": the fp32 CUDA-core kernels
        self.compute_mode = "auto"
        self.register_load_state_dict_post_hook(lambda mod, keys: mod.mark_dirty())

    def mark_dirty(self):
        self._dirty = True

    def train(self, mode=True):
        self._dirty = True
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._dirty = True
        out = super()._apply(fn, *a, **k)
        st = getattr(self, "_train_state", None)
        if st is not None:
            st.rebind()                  # the module's tensors were replaced: re-attach them to the training state's flat buffer
        return out

    def _device_index(self):
        dev = self.out_fc[2].weight.device
        if dev.type != "cuda":
            raise RuntimeError("mural_b200.UNet_Small runs on CUDA only (no CPU fallback)")
        return dev.index if dev.index is not None else torch.cuda.current_device()

    def _handle(self, radius):
        L = _lib.lib()
        if self._h is None or self._hR != radius:
            if self._h is not None:
                L.mural_indel_model_destroy(self._h)
            cfg = _lib.IndelConfig(radius, self._C, self._ks, self.n_class, (C.c_int32 * 6)(*self._down), int(bool(self.use_reverse)))
            h = C.c_void_p()
            _lib.check(L.mural_indel_model_create(C.byref(cfg), self._device_index(), C.byref(h)))
            self._h, self._hR, self._dirty = h, radius, True
        if self._dirty:
            sd = self.state_dict()
            blob = np.empty(int(L.mural_indel_model_n_params(self._h)), dtype=np.float32)
            for i in range(L.mural_indel_model_n_tensors(self._h)):
                name, off, num, buf = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
                _lib.check(L.mural_indel_model_tensor(self._h, i, C.byref(name), C.byref(off), C.byref(num), C.byref(buf)))
                blob[off.value:off.value + num.value] = sd[name.value.decode()].detach().reshape(-1).to("cpu", torch.float32).numpy()
            with torch.cuda.device(self._device_index()):
                _lib.check(L.mural_indel_model_load(self._h, _lib.ptr(blob), blob.size))
            self._dirty = False
        if self.compute_mode not in ("auto", "fp32"):
            raise ValueError("compute_mode must be 'auto' or 'fp32'")
        _lib.check(L.mural_indel_set_mode(self._h, 1 if self.compute_mode == "fp32" else 0))
        return self._h

    def __del__(self):
        try:
            if self._h is not None:
                _lib.lib().mural_indel_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def forward(self, distal_input, distal_radius=None):
        if self.training:
            from .training import unet_train_forward
            return unet_train_forward(self, distal_input, distal_radius)
        L = _lib.lib()
        with torch.cuda.device(self._device_index()):
            if isinstance(distal_input, SiteBatch):
                if distal_radius is None:
                    raise ValueError("distal_radius is required with a SiteBatch")
                h = self._handle(int(distal_radius))
                n = len(distal_input)
                out = torch.empty((n, self.n_class), dtype=torch.float32, device=distal_input.pos.device)
                _lib.check(L.mural_indel_forward(h, distal_input.genome.handle, _lib.ptr(distal_input.pos), _lib.ptr(distal_input.meta), n,
                                                 _lib.ptr(out), _lib.current_stream()))
                return out
            x = distal_input.to(torch.float32).contiguous()
            h = self._handle(x.shape[2] // 2)
            out = torch.empty((x.shape[0], self.n_class), dtype=torch.float32, device=x.device)
            _lib.check(L.mural_indel_forward_tensors(h, _lib.ptr(x), x.shape[0], x.shape[2], _lib.ptr(out), _lib.current_stream()))
            return out
