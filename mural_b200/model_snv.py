"""Drop-in MuRaL-snv `Network2` whose forward runs on the sm_100a kernels behind the C ABI.

Contract kept from the reference (MuRaL/model/model_snv.py:290-525):
  * same constructor parameter names (model_choice resolves them by name, nn_utils.py:228-229);
  * identical `state_dict()` keys/shapes, including the aliased `RBs*.N.layer.{1,2,4,5}.*` entries
    (ResBlock registers bn/conv twice, model_snv.py:799-804), the zero-sized `first_bn_layer.*` when
    there are no continuous features, and `num_batches_tracked`;
  * `forward((cont_x, cat_x), distal_x) -> [B, n_class]` float32 log-probabilities.
The torch modules below only *hold* the parameters; no torch op computes the network.

Fast path: pass a `mural_b200.data.SiteBatch` as `distal_input` — windows are then gathered on the GPU
from the packed genome and no one-hot tensor ever exists.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .data import SiteBatch


class ResBlock(nn.Module):
    """Parameter container with the reference ResBlock's registration order (model_snv.py:794-804)."""

    def __init__(self, in_channels=32, kernel_size=3, stride=1, padding=0, dilation=1):
        super().__init__()
        self.bn1 = nn.BatchNorm1d(in_channels)
        self.conv1 = nn.Conv1d(in_channels, in_channels, kernel_size, stride=stride, padding=padding, dilation=dilation)
        self.bn2 = nn.BatchNorm1d(in_channels)
        self.conv2 = nn.Conv1d(in_channels, in_channels, kernel_size, stride=stride, padding=padding, dilation=dilation)
        # same objects registered a second time -> the '.layer.N.' aliases of the checkpoint format
        self.layer = nn.Sequential(nn.ReLU(), self.bn1, self.conv1, nn.ReLU(), self.bn2, self.conv2)


def _bn_conv(cin, cout, ks, relu=False):
    mods = [nn.BatchNorm1d(cin), nn.Conv1d(cin, cout, ks, 1, (ks - 1) // 2)]
    if relu:
        mods.append(nn.ReLU())
    return nn.Sequential(*mods)


class Network2(nn.Module):
    """Combined local (embedding + feed-forward) and two-scale ResNet model, B200 kernels underneath."""

    def __init__(self, emb_dims, no_of_cont, lin_layer_sizes, emb_dropout, lin_layer_dropouts, in_channels, out_channels,
                 kernel_size, distal_radius, distal_order, distal_fc_dropout, n_class, emb_padding_idx=None):
        super().__init__()
        # continuous (bigWig mean) features enter the LOCAL branch (first_bn_layer + wider first Linear, model_snv.py:326-334,
        # 457-463).  The expanded-window convs only ever see the 4 sequence channels: forward slices distal_input[:, 0:in_channels]
        # and CombinedDatasetNP never adds track channels to it (preprocessing.py:934-942), so in_channels > 4 can only be fed by
        # the reference's HDF5 datasets (out of scope).
        if in_channels not in (4, 4 + no_of_cont):
            raise ValueError("in_channels must be 4 (+ n_cont)")
        if in_channels != 4:
            raise NotImplementedError("mural_b200.Network2: bigWig tracks as extra channels of the expanded window (in_channels > 4) "
                                      "exist only in the reference's HDF5 data path; pass without_bw_distal (in_channels = 4)")
        if len(lin_layer_sizes) != 2:
            raise NotImplementedError("mural_b200.Network2 expects two hidden local layers (model_choice always passes two)")
        self.n_class = n_class
        self.in_channels = in_channels
        self.no_of_cat = len(emb_dims)
        self.emb_layer = nn.Embedding(emb_padding_idx + 1, 5)
        self.no_of_embs = len(emb_dims) * 5
        self.no_of_cont = no_of_cont
        sizes = [self.no_of_embs + no_of_cont] + list(lin_layer_sizes)
        self.lin_layers = nn.ModuleList([nn.Linear(sizes[i], sizes[i + 1]) for i in range(len(lin_layer_sizes))])
        self.first_bn_layer = nn.BatchNorm1d(no_of_cont)
        self.bn_layers = nn.ModuleList([nn.BatchNorm1d(s) for s in lin_layer_sizes])
        self.emb_dropout_layer = nn.Dropout(emb_dropout)
        self.droput_layers = nn.ModuleList([nn.Dropout(p) for p in lin_layer_dropouts])
        self.kernel_size = kernel_size
        self.distal_radius = distal_radius
        self.seq_len = distal_radius * 2 + 1 - (distal_order - 1)
        C_ = out_channels
        for sfx, pools in (("", ((3, 3, 1), (3, 3, 1), (3, 3, 1))), ("_2", ((15, 15, 7), (7, 7, 3), (3, 3, 1)))):
            setattr(self, "conv1" + sfx, _bn_conv(in_channels, C_, kernel_size))
            setattr(self, "maxpool1" + sfx, nn.MaxPool1d(*pools[0]))
            setattr(self, "RBs1" + sfx, nn.Sequential(*[ResBlock(C_, 3, 1, 1, 1) for _ in range(2)]))
            setattr(self, "maxpool2" + sfx, nn.MaxPool1d(*pools[1]))
            setattr(self, "conv2" + sfx, _bn_conv(C_, C_, kernel_size))
            setattr(self, "RBs2" + sfx, nn.Sequential(*[ResBlock(C_, 3, 1, 1, 1) for _ in range(2)]))
            setattr(self, "maxpool3" + sfx, nn.MaxPool1d(*pools[2]))
            setattr(self, "conv3" + sfx, _bn_conv(C_, C_, kernel_size, relu=True))
            setattr(self, "distal_fc1" if sfx == "" else "distal_fc2",
                    nn.Sequential(nn.BatchNorm1d(C_), nn.Dropout(distal_fc_dropout), nn.Linear(C_, n_class)))
        self.local_fc = nn.Sequential(nn.Linear(lin_layer_sizes[-1], n_class))
        # ---- native side
        k = 0
        while 4 ** k < emb_padding_idx:
            k += 1
        assert 4 ** k == emb_padding_idx, "emb_padding_idx must be 4**local_order (nn_utils.py:196)"
        self.local_order = k
        self.local_radius = (self.no_of_cat + (k - 1) - 1) // 2
        self._cfg = _lib.SnvConfig(self.local_radius, k, distal_radius, lin_layer_sizes[0], lin_layer_sizes[1], out_channels,
                                   kernel_size, n_class, no_of_cont)
        self._h = None
        self._dirty = True
        # "fp32" (fp32-equivalent, gate 1e-3) | "bf16" (tcgen05, gate 5e-3) | "auto" (bf16 everywhere + the fp32-equivalent
        # path again for windows with non-ACGT symbols / chromosome overhang; the default of run_predict)
        self.compute_mode = "fp32"             # models with continuous features stay in this mode
        self.register_load_state_dict_post_hook(lambda mod, keys: mod.mark_dirty())

    # ------------------------------------------------------------------ native handle management
    def mark_dirty(self):
        self._dirty = True

    def train(self, mode=True):
        self._dirty = True            # weights may change while training; re-fold on the next eval forward
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self._dirty = True
        out = super()._apply(fn, *a, **k)
        st = getattr(self, "_train_state", None)
        if st is not None:
            st.rebind()                  # the module's tensors were replaced: re-attach them to the training state's flat buffer
        return out

    def _device_index(self):
        dev = self.emb_layer.weight.device
        if dev.type != "cuda":
            raise RuntimeError("mural_b200.Network2 runs on CUDA only (no CPU fallback); call model.to('cuda')")
        return dev.index if dev.index is not None else torch.cuda.current_device()

    def native_layout(self):
        """[(state_dict key, offset, numel, is_buffer)] of the flat parameter blob (C ABI layout)."""
        L = _lib.lib()
        h = self._ensure_handle()
        out = []
        for i in range(L.mural_snv_model_n_tensors(h)):
            name, off, num, buf = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
            _lib.check(L.mural_snv_model_tensor(h, i, C.byref(name), C.byref(off), C.byref(num), C.byref(buf)))
            out.append((name.value.decode(), off.value, num.value, buf.value))
        return out

    def _ensure_handle(self):
        if self._h is None:
            h = C.c_void_p()
            _lib.check(_lib.lib().mural_snv_model_create(C.byref(self._cfg), self._device_index(), C.byref(h)))
            self._h = h
        return self._h

    def flat_blob(self):
        """All parameters + BN running stats as one float32 numpy vector in the C ABI layout."""
        sd = self.state_dict()
        n = int(_lib.lib().mural_snv_model_n_params(self._ensure_handle()))
        blob = np.empty(n, dtype=np.float32)
        for name, off, num, _ in self.native_layout():
            t = sd[name]
            if t.numel() != num:
                raise RuntimeError("size mismatch for %s: %d vs %d" % (name, t.numel(), num))
            blob[off:off + num] = t.detach().reshape(-1).to("cpu", torch.float32).numpy()
        return blob

    def refresh(self):
        """Re-fold the current parameters into the device-resident eval weights."""
        blob = self.flat_blob()
        with torch.cuda.device(self._device_index()):
            _lib.check(_lib.lib().mural_snv_model_load(self._ensure_handle(), _lib.ptr(blob), blob.size))
        self._dirty = False

    def __del__(self):
        try:
            if self._h is not None:
                _lib.lib().mural_snv_model_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def state_dict(self, *args, **kwargs):
        st = getattr(self, "_train_state", None)
        if st is not None:
            st.sync_counters()              # num_batches_tracked follows the kernels' training forwards
        return super().state_dict(*args, **kwargs)

    # ------------------------------------------------------------------ forward
    def forward(self, local_input, distal_input):
        if self.training:
            from .training import network2_train_forward
            return network2_train_forward(self, local_input, distal_input)
        if self._dirty:
            self.refresh()
        L = _lib.lib()
        mode = _lib.MODES["fp32" if self.no_of_cont else self.compute_mode]
        dev = self._device_index()
        with torch.cuda.device(dev):
            cont = None
            if self.no_of_cont:                 # cont_x [n, n_cont] (model_snv.py:448): local_input[0], or SiteBatch.cont on the fast path
                cont = distal_input.cont if isinstance(distal_input, SiteBatch) else local_input[0]
                if cont is None or cont.shape[1] != self.no_of_cont:
                    raise RuntimeError("cont_x with %d continuous features per site is required" % self.no_of_cont)
                cont = cont.to(device=self.emb_layer.weight.device, dtype=torch.float32).contiguous()
                _lib.check(L.mural_snv_set_cont(self._h, _lib.ptr(cont)))
            if isinstance(distal_input, SiteBatch):
                n = len(distal_input)
                out = torch.empty((n, self.n_class), dtype=torch.float32, device=distal_input.pos.device)
                _lib.check(L.mural_snv_forward(self._h, distal_input.genome.handle, _lib.ptr(distal_input.pos),
                                               _lib.ptr(distal_input.meta), n, mode, _lib.ptr(out), _lib.current_stream()))
                return out
            cont_data, cat_data = local_input
            assert distal_input.shape[2] > 200, "Error: distal seq len must be >200"      # model_snv.py:470
            x = distal_input[:, 0:self.in_channels, :].to(torch.float32).contiguous()
            cat = cat_data.to(torch.int64).contiguous()
            if cat.shape[1] != self.no_of_cat:
                raise RuntimeError("cat_x has %d columns, model expects %d" % (cat.shape[1], self.no_of_cat))
            out = torch.empty((x.shape[0], self.n_class), dtype=torch.float32, device=x.device)
            _lib.check(L.mural_snv_forward_tensors(self._h, _lib.ptr(cat), _lib.ptr(x), x.shape[0], x.shape[2], mode,
                                                   _lib.ptr(out), _lib.current_stream()))
            return out

    def predict_host(self, genome, pos, meta, out=None):
        """End-to-end call on HOST site arrays (int32 numpy / pinned tensors): H2D + kernels + D2H."""
        if self._dirty:
            self.refresh()
        n = len(pos)
        if out is None:
            out = np.empty((n, self.n_class), dtype=np.float32)
        with torch.cuda.device(self._device_index()):
            _lib.check(_lib.lib().mural_snv_predict_host(self._h, genome.handle, _lib.ptr(pos), _lib.ptr(meta), n,
                                                         _lib.MODES[self.compute_mode], _lib.ptr(out), _lib.current_stream()))
        return out

    # ------------------------------------------------------------------ parity helpers
    def set_debug(self, on=True, chunk=0):
        _lib.check(_lib.lib().mural_snv_set_debug(self._ensure_handle(), int(on)))
        _lib.check(_lib.lib().mural_snv_set_chunk(self._ensure_handle(), int(chunk)))

    def debug_tap(self, name):
        n = C.c_int64()
        _lib.check(_lib.lib().mural_snv_debug_tap(self._h, name.encode(), None, 0, C.byref(n)))
        buf = np.empty(n.value, dtype=np.float32)
        _lib.check(_lib.lib().mural_snv_debug_tap(self._h, name.encode(), _lib.ptr(buf), n.value, C.byref(n)))
        return buf
