"""Training step of MuRaL-snv on the B200 kernels (MuRaL/training.py:404-452).

Two ways in:

* drop-in — `model.train(); preds = model.forward((cont_x, cat_x), site_batch); loss = criterion(preds, y);
  loss.backward(); clip_grad_norm_(model.parameters(), 10); optimizer.step()` keeps working with torch optimizers:
  in train mode `Network2.forward` routes here and returns a differentiable tensor (custom autograd.Function around
  mural_snv_train_forward / mural_snv_train_backward).
* fused — `TrainState.step(batch)`: forward, CE(sum)+gradient, backward, (NCCL all-reduce of ONE flat gradient buffer
  when torch.distributed is initialised), global-norm clip + Adam / AdamW(amsgrad) / SGD(nesterov) in one kernel, no
  host synchronisation inside the step (the reference syncs with loss.item() every batch).

Parameters and BatchNorm buffers of the model are re-pointed at views of one flat device tensor (the C-ABI blob
layout), so `state_dict()` / checkpoints stay byte-compatible with the reference while the kernels see one buffer.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .data import SiteBatch

OPTIMIZERS = {"Adam": 0, "AdamW": 1, "AdamW2": 1, "SGD": 2}     # training.py:346-357


class StepLR:
    """torch.optim.lr_scheduler.StepLR as used by the reference: step_size=(5000*128)//batch_size, gamma=LR_gamma,
    stepped every batch, with the min_lr -> restart_lr rule (training.py:366, 444-450)."""

    def __init__(self, lr, step_size, gamma, min_lr=None, restart_lr=None):
        self.base, self.step_size, self.gamma, self.min_lr, self.restart_lr = lr, max(1, int(step_size)), gamma, min_lr, restart_lr
        self.lr, self.k = lr, 0

    def step(self):
        self.k += 1
        if self.k % self.step_size == 0:
            self.lr *= self.gamma
        if self.min_lr is not None and self.lr < self.min_lr:
            self.lr = self.restart_lr
        return self.lr


class ReduceLROnPlateau:
    """torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min', factor=0.2, patience=1, threshold=1e-4 (rel), min_lr=1e-7) as the
    reference configures it (training.py:370); stepped once per epoch with the validation loss."""

    def __init__(self, lr, factor=0.2, patience=1, threshold=1e-4, min_lr=1e-7, eps=1e-8):
        self.lr, self.factor, self.patience, self.threshold, self.min_lr, self.eps = lr, factor, patience, threshold, min_lr, eps
        self.best, self.bad = float("inf"), 0

    def step(self, metric):
        if metric < self.best * (1.0 - self.threshold):
            self.best, self.bad = metric, 0
        else:
            self.bad += 1
        if self.bad > self.patience:
            new = max(self.lr * self.factor, self.min_lr)
            if self.lr - new > self.eps:
                self.lr = new
            self.bad = 0
        return self.lr


def auto_weight_decay(weight_decay_auto, batch_size, epochs, train_size):
    """--weight_decay_auto (training.py:339-344): weight_decay = 1 - wda ** (batch_size / (epochs * train_size))."""
    if not 0 < weight_decay_auto < 1:
        raise ValueError("Please set a value smaller than 1 for --weight_decay_auto.")
    return 1 - weight_decay_auto ** (batch_size / (epochs * train_size))


def make_scheduler(config, train_size):
    """The three schedulers of training.py:364-371.  StepLR / StepLR2 are stepped every batch (with the min_lr -> restart_lr
    rule, :444-450), ROP once per epoch on the validation loss."""
    kind, lr = config["lr_scheduler"], config["learning_rate"]
    if kind == "StepLR":
        return StepLR(lr, (5000 * 128) // config["batch_size"], config["LR_gamma"], config.get("min_lr"), config.get("restart_lr"))
    if kind == "StepLR2":
        gamma = (config["min_lr"] / config["restart_lr"]) ** (1 / (train_size // config["batch_size"]))
        return StepLR(lr, 1, gamma, config.get("min_lr"), config.get("restart_lr"))
    if kind == "ROP":
        return ReduceLROnPlateau(lr)
    raise ValueError("unsupported lr_scheduler %s" % kind)


class TrainState:
    MAX_GRAPHS = 4   # distinct batch sizes kept as captured graphs (the full batch and the odd tail sizes that recur)

    def __init__(self, model, optim="Adam", lr=1e-3, weight_decay=0.0, max_norm=10.0, seed=0, grad_average=False, use_graph=True):
        """use_graph: capture the step of the first full-size batch in CUDA graphs and replay them for every batch of that size
        (the step is ~200 small launches and launch-bound at the reference's batch of 128); other batch sizes run eagerly."""
        L = _lib.lib()
        self.model = model
        self.device = model.emb_layer.weight.device
        if self.device.type != "cuda":
            raise RuntimeError("mural_b200 training runs on CUDA only")
        if optim not in OPTIMIZERS:
            raise ValueError("Error: unsupported optimization method %s" % optim)         # training.py:359-361
        self.n_cont = int(getattr(model, "no_of_cont", 0) or 0)   # bigWig window means behind the embeddings (model_snv.py:457-463)
        self.kind, self.lr, self.weight_decay, self.max_norm = OPTIMIZERS[optim], float(lr), float(weight_decay), float(max_norm)
        layout = model.native_layout()
        self.n_blob = int(L.mural_snv_model_n_params(model._ensure_handle()))
        self.n_trainable = int(L.mural_snv_model_n_trainable(model._ensure_handle()))
        self.blob = torch.empty(self.n_blob, dtype=torch.float32, device=self.device)
        sd = dict(model.named_parameters())
        sd.update(dict(model.named_buffers()))
        self.params, self.grad_views, seen = [], [], set()
        self.grads = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        for name, off, num, is_buf in layout:
            t = sd[name]
            view = self.blob[off:off + num].view(t.shape)
            view.copy_(t.detach())
            t.data = view                                     # parameter/buffer now lives inside the flat blob
            if not is_buf and id(t) not in seen:
                seen.add(id(t))
                self.params.append(t)
                self.grad_views.append(self.grads[off:off + num].view(t.shape))
        self.m = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        self.v = torch.zeros_like(self.m) if self.kind != 2 else None
        self.vmax = torch.zeros_like(self.m) if self.kind == 1 else None
        self.scratch = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.loss_dev = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.opt_step = 0
        # per-step hyper-parameters on the device (read by the optimizer kernel) so that a captured graph never goes stale
        self._lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
        self._lr_on_dev = float(lr)
        self._opt_step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.use_graph = bool(use_graph)
        self._graphs = {}             # batch size -> dict(n, genome, pos, meta, logp, dlogp, fb, opt); at most MAX_GRAPHS sizes
        self._graph_warm = {}         # batch size -> eager steps seen (one eager step sizes the tape before capture)
        self._tape_n = 0              # largest batch the native workspace has been sized for
        self.n_forward = 0
        self._tracked_synced = 0
        self.grad_average = grad_average
        h = C.c_void_p()
        _lib.check(L.mural_snv_train_create(model._h, C.byref(h)))
        self._h = h
        p_emb = float(model.emb_dropout_layer.p)
        p_loc = float(model.droput_layers[0].p) if len(model.droput_layers) else 0.0
        p_fc = float(model.distal_fc1[1].p)
        _lib.check(L.mural_snv_train_set_dropout(self._h, p_emb, p_loc, p_fc, int(seed)))
        model._train_state = self

    def set_dropout(self, p_emb, p_local, p_fc, seed=0):
        _lib.check(_lib.lib().mural_snv_train_set_dropout(self._h, float(p_emb), float(p_local), float(p_fc), int(seed)))

    def rebind(self):
        """Called by Network2._apply after .to() / .float() / ...: the module's tensors were replaced, so copy their
        current values into the flat blob and point them at its views again (optimizer moments are kept).  While the model sits
        on another device or dtype the state is detached and step() refuses to run."""
        _rebind_views(self, self.model.native_layout())

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().mural_snv_train_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _note_batch(self, n):
        """The native workspace (tape) is reallocated when a larger batch than any before arrives: captured graphs hold
        pointers into the old one and are dropped."""
        if n > self._tape_n:
            self._tape_n = n
            self._graphs.clear()
            self._graph_warm.clear()

    # ---- pieces
    def _cont_of(self, batch):
        """cont_x [n, n_cont] of a batch as a contiguous fp32 device tensor (None for models without continuous features)."""
        if not self.n_cont:
            return None
        cont = getattr(batch, "cont", None)
        if cont is None or cont.shape[0] != len(batch) or cont.shape[1] != self.n_cont:
            raise RuntimeError("cont_x with %d continuous features per site is required (SiteBatch.cont)" % self.n_cont)
        return cont.to(device=self.device, dtype=torch.float32).contiguous()

    def forward(self, batch):
        n = len(batch)
        _check_attached(self)
        self._note_batch(n)
        logp = torch.empty((n, self.model.n_class), dtype=torch.float32, device=self.device)
        cont = self._cont_of(batch)
        self._cont_keep = cont                               # read again by the backward of first_bn_layer
        with torch.cuda.device(self.device):
            if cont is not None:
                _lib.check(_lib.lib().mural_snv_set_cont(self.model._ensure_handle(), _lib.ptr(cont)))
            _lib.check(_lib.lib().mural_snv_train_forward(self._h, batch.genome.handle, _lib.ptr(batch.pos), _lib.ptr(batch.meta), n,
                                                          _lib.ptr(self.blob), _lib.ptr(logp), _lib.current_stream()))
        self.n_forward += 1
        self.model.mark_dirty()
        return logp

    def backward(self, dlogp):
        dlogp = dlogp.contiguous().to(torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_snv_train_backward(self._h, _lib.ptr(self.blob), _lib.ptr(dlogp), _lib.ptr(self.grads),
                                                           _lib.current_stream()))
        return self.grads

    def all_reduce_grads(self):
        """The only collective of a data-parallel step: one flat fp32 all-reduce (sum) over NVLink."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(self.grads, op=torch.distributed.ReduceOp.SUM)
            return torch.distributed.get_world_size()
        return 1

    def apply(self, world=1):
        self.opt_step += 1
        if self._lr_on_dev != self.lr:
            self._lr_dev.fill_(self.lr)
            self._lr_on_dev = self.lr
        self._launch_optimizer(world)
        self.model.mark_dirty()

    def _launch_optimizer(self, world):
        scale = 1.0 / world if self.grad_average else 1.0   # reference loss is SUM-reduced: summing == one big batch
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_optimizer_step_dev(self.kind, _lib.ptr(self.blob), _lib.ptr(self.grads), _lib.ptr(self.m),
                                                           _lib.ptr(self.v), _lib.ptr(self.vmax), self.n_trainable, _lib.ptr(self._lr_dev),
                                                           self.weight_decay, _lib.ptr(self._opt_step_dev), self.max_norm, scale,
                                                           _lib.ptr(self.scratch), _lib.current_stream()))

    def _launch_forward_backward(self, genome, pos, meta, n, logp, dlogp, cont=None):
        L = _lib.lib()
        with torch.cuda.device(self.device):
            if cont is not None:
                _lib.check(L.mural_snv_set_cont(self.model._ensure_handle(), _lib.ptr(cont)))
            _lib.check(L.mural_snv_train_forward(self._h, genome.handle, _lib.ptr(pos), _lib.ptr(meta), n, _lib.ptr(self.blob),
                                                 _lib.ptr(logp), _lib.current_stream()))
            _lib.check(L.mural_ce_sum_grad(_lib.ptr(logp), _lib.ptr(meta), n, self.model.n_class, _lib.ptr(self.loss_dev), _lib.ptr(dlogp),
                                           _lib.current_stream()))
            _lib.check(L.mural_snv_train_backward(self._h, _lib.ptr(self.blob), _lib.ptr(dlogp), _lib.ptr(self.grads),
                                                  _lib.current_stream()))

    def _world(self):
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    def _capture(self, batch):
        """Two graphs: forward + CE + backward, and clip + optimizer.  The gradient all-reduce of a data-parallel step runs
        between them on the same stream (eagerly: NCCL stays outside the capture); with one rank the two replays are back to
        back.  Static buffers hold the batch; dropout position, lr and the optimizer step are read from device memory."""
        n = len(batch)
        g = {"n": n, "genome": batch.genome, "pos": batch.pos.clone(), "meta": batch.meta.clone(),
             "logp": torch.empty((n, self.model.n_class), dtype=torch.float32, device=self.device),
             "world": self._world()}
        g["dlogp"] = torch.empty_like(g["logp"])
        c0 = self._cont_of(batch)
        g["cont"] = c0.clone() if c0 is not None else None   # static buffer of the graph, refilled per step
        torch.cuda.synchronize(self.device)
        g["fb"], g["opt"] = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        # thread-local capture mode: CUDA calls of other threads (NCCL watchdog, clock sampler) must not invalidate the capture
        with torch.cuda.graph(g["fb"], capture_error_mode="thread_local"):
            self._launch_forward_backward(g["genome"], g["pos"], g["meta"], n, g["logp"], g["dlogp"], g["cont"])
        with torch.cuda.graph(g["opt"], capture_error_mode="thread_local"):
            self._launch_optimizer(g["world"])
        return g

    # ---- fused step
    def step(self, batch):
        """forward + CE(sum) + backward + all-reduce + clip + optimizer; returns the log-probs (loss accumulates in
        self.loss_dev, read it with .item() once per print interval).  On the graph path the returned tensor is the graph's
        static output buffer: it is overwritten by the next step of the same batch size — clone it to keep it."""
        n = len(batch)
        if n < 2:
            return None                                      # training.py:415: batches of one site are skipped
        _check_attached(self)
        self._note_batch(n)
        if self.use_graph:
            g = self._graphs.get(n)
            if g is None and self._graph_warm.get(n, 0) >= 1 and len(self._graphs) < self.MAX_GRAPHS:
                # capture does not execute: the captured step runs at the replay below
                g = self._graphs[n] = self._capture(batch)
            if g is not None and g["genome"] is batch.genome and g["world"] == self._world():
                g["pos"].copy_(batch.pos)
                g["meta"].copy_(batch.meta)
                if g["cont"] is not None:
                    g["cont"].copy_(self._cont_of(batch))
                if self._lr_on_dev != self.lr:
                    self._lr_dev.fill_(self.lr)
                    self._lr_on_dev = self.lr
                g["fb"].replay()
                self.n_forward += 1
                self.all_reduce_grads()
                g["opt"].replay()
                self.opt_step += 1
                self.model.mark_dirty()
                return g["logp"]
            self._graph_warm[n] = self._graph_warm.get(n, 0) + 1
        logp = self.forward(batch)
        dlogp = torch.empty_like(logp)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(logp), _lib.ptr(batch.meta), n, self.model.n_class, _lib.ptr(self.loss_dev),
                                                    _lib.ptr(dlogp), _lib.current_stream()))
        self.backward(dlogp)
        world = self.all_reduce_grads()
        self.apply(world)
        return logp

    def sync_buffers(self):
        """Data parallel: BatchNorm statistics are rank-local during training (DDP default); average the running
        statistics over ranks so that the checkpoint rank 0 writes does not carry rank-0 statistics only."""
        d = torch.distributed
        if d.is_available() and d.is_initialized() and d.get_world_size() > 1 and self.n_blob > self.n_trainable:
            buf = self.blob[self.n_trainable:]
            d.all_reduce(buf, op=d.ReduceOp.SUM)
            buf.div_(d.get_world_size())
            self.model.mark_dirty()

    def sync_counters(self):
        """num_batches_tracked of every BatchNorm follows the number of training forwards (checkpoint contract)."""
        d = self.n_forward - self._tracked_synced
        if d:
            for mod in self.model.modules():
                if isinstance(mod, torch.nn.BatchNorm1d) and mod.num_batches_tracked is not None:
                    mod.num_batches_tracked += d
            self._tracked_synced = self.n_forward


def _rebind_views(st, layout):
    sd = dict(st.model.named_parameters())
    sd.update(dict(st.model.named_buffers()))
    ts = [sd[name] for name, _, _, _ in layout]
    st._detached = any(t.device != st.device or t.dtype != torch.float32 for t in ts)
    if st._detached:
        return
    for (name, off, num, _), t in zip(layout, ts):
        view = st.blob[off:off + num].view(t.shape)
        if t.data_ptr() != view.data_ptr():
            view.copy_(t.detach())
            t.data = view
    if hasattr(st, "_graphs"):
        st._graphs.clear()
        st._graph_warm.clear()
    st.model.mark_dirty()


def _check_attached(st):
    if getattr(st, "_detached", False):
        raise RuntimeError("the model was moved off %s (or cast) after its training state was created; move it back before stepping" % st.device)


class _TrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, batch, *params):
        ctx.state = state
        return state.forward(batch)

    @staticmethod
    def backward(ctx, dlogp):
        st = ctx.state
        st.backward(dlogp)
        return (None, None) + tuple(g.clone() for g in st.grad_views)


def network2_train_forward(model, local_input, distal_input):
    """Network2.forward in train() mode (differentiable)."""
    if not isinstance(distal_input, SiteBatch):
        raise NotImplementedError("mural_b200 training consumes SiteBatch-es (site records); tensor inputs are eval-only")
    st = getattr(model, "_train_state", None)
    if st is None:
        st = TrainState(model)
    return _TrainFn.apply(st, distal_input, *st.params)


class IndelTrainState:
    """Training step of MuRaL-indel `UNet_Small` (training.py:404-452 with model_type 'indel') on the native
    tape (csrc/indel_train.cu).  Same protocol as `TrainState`: the model's parameters and BatchNorm buffers become views of one
    flat fp32 buffer, gradients land in one flat buffer (the single all-reduce of a data-parallel step), clip + optimizer are
    the fused kernels.  The loss is CrossEntropyLoss(sum) on the Softplus outputs, as in the reference."""

    MAX_GRAPHS = 4

    def __init__(self, model, distal_radius, optim="Adam", lr=1e-3, weight_decay=0.0, max_norm=10.0, seed=0, grad_average=False,
                 use_graph=True):
        """use_graph: capture the step of a batch size in CUDA graphs (forward + CE + backward | clip + optimizer, the data-parallel
        all-reduce eagerly between them) after one eager step of that size and replay them; the dropout stream position, the
        learning rate and the optimizer step are read from device memory, so replays stay in step with the eager form."""
        L = _lib.lib()
        self.use_graph = bool(use_graph)
        self._graphs, self._graph_warm = {}, {}
        self.model = model
        self.device = model.out_fc[2].weight.device
        if self.device.type != "cuda":
            raise RuntimeError("mural_b200 training runs on CUDA only")
        if optim not in OPTIMIZERS:
            raise ValueError("Error: unsupported optimization method %s" % optim)
        self.kind, self.lr, self.weight_decay, self.max_norm = OPTIMIZERS[optim], float(lr), float(weight_decay), float(max_norm)
        self.distal_radius = int(distal_radius)
        h = model._handle(self.distal_radius)
        self.n_blob = int(L.mural_indel_model_n_params(h))
        self.blob = torch.empty(self.n_blob, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(self.n_blob, dtype=torch.float32, device=self.device)
        sd = dict(model.named_parameters())
        sd.update(dict(model.named_buffers()))
        self.params, self.grad_views, self.n_trainable = [], [], 0
        for i in range(L.mural_indel_model_n_tensors(h)):
            name, off, num, buf = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
            _lib.check(L.mural_indel_model_tensor(h, i, C.byref(name), C.byref(off), C.byref(num), C.byref(buf)))
            t = sd[name.value.decode()]
            view = self.blob[off.value:off.value + num.value].view(t.shape)
            view.copy_(t.detach())
            t.data = view                                     # parameter / running statistic now lives inside the flat blob
            if not buf.value:
                self.params.append(t)
                self.grad_views.append(self.grads[off.value:off.value + num.value].view(t.shape))
                self.n_trainable = max(self.n_trainable, off.value + num.value)   # trainable tensors come first in the layout
        self.m = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
        self.v = torch.zeros_like(self.m) if self.kind != 2 else None
        self.vmax = torch.zeros_like(self.m) if self.kind == 1 else None
        self.scratch = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.loss_dev = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
        self._lr_on_dev = float(lr)
        self._opt_step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.opt_step = 0
        self.n_forward = 0
        self._tracked_synced = 0
        self.grad_average = grad_average
        t = C.c_void_p()
        _lib.check(L.mural_indel_train_create(h, C.byref(t)))
        self._h = t
        _lib.check(L.mural_indel_train_set_dropout(self._h, float(model.out_fc[1].p), int(seed)))
        model._train_state = self

    def set_dropout(self, p_fc, seed=0):
        _lib.check(_lib.lib().mural_indel_train_set_dropout(self._h, float(p_fc), int(seed)))

    def rebind(self):
        """See TrainState.rebind."""
        L = _lib.lib()
        h = self.model._handle(self.distal_radius)
        layout = []
        for i in range(L.mural_indel_model_n_tensors(h)):
            name, off, num, buf = C.c_char_p(), C.c_int64(), C.c_int64(), C.c_int32()
            _lib.check(L.mural_indel_model_tensor(h, i, C.byref(name), C.byref(off), C.byref(num), C.byref(buf)))
            layout.append((name.value.decode(), off.value, num.value, buf.value))
        _rebind_views(self, layout)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().mural_indel_train_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def forward(self, batch):
        """batch: SiteBatch (windows gathered on the device) or the reference's one-hot tensor [B, 4, 2R]."""
        L = _lib.lib()
        _check_attached(self)
        with torch.cuda.device(self.device):
            if isinstance(batch, SiteBatch):
                n = len(batch)
                out = torch.empty((n, self.model.n_class), dtype=torch.float32, device=self.device)
                _lib.check(L.mural_indel_train_forward(self._h, batch.genome.handle, _lib.ptr(batch.pos), _lib.ptr(batch.meta), n,
                                                       _lib.ptr(self.blob), _lib.ptr(out), _lib.current_stream()))
            else:
                x = batch.to(self.device, torch.float32).contiguous()
                n = x.shape[0]
                out = torch.empty((n, self.model.n_class), dtype=torch.float32, device=self.device)
                _lib.check(L.mural_indel_train_forward_tensors(self._h, _lib.ptr(x), n, x.shape[2], _lib.ptr(self.blob), _lib.ptr(out),
                                                               _lib.current_stream()))
        self.n_forward += 1
        self.model.mark_dirty()
        return out

    def backward(self, dout):
        dout = dout.contiguous().to(torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_indel_train_backward(self._h, _lib.ptr(self.blob), _lib.ptr(dout), _lib.ptr(self.grads),
                                                             _lib.current_stream()))
        return self.grads

    def _launch_optimizer(self, world):
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_optimizer_step_dev(self.kind, _lib.ptr(self.blob), _lib.ptr(self.grads), _lib.ptr(self.m),
                                                           _lib.ptr(self.v), _lib.ptr(self.vmax), self.n_trainable, _lib.ptr(self._lr_dev),
                                                           self.weight_decay, _lib.ptr(self._opt_step_dev), self.max_norm,
                                                           1.0 / world if self.grad_average else 1.0, _lib.ptr(self.scratch),
                                                           _lib.current_stream()))

    def _world(self):
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    def _capture(self, batch):
        n = len(batch)
        L = _lib.lib()
        g = {"genome": batch.genome, "pos": batch.pos.clone(), "meta": batch.meta.clone(), "world": self._world(),
             "out": torch.empty((n, self.model.n_class), dtype=torch.float32, device=self.device)}
        g["dout"] = torch.empty_like(g["out"])
        torch.cuda.synchronize(self.device)
        g["fb"], g["opt"] = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g["fb"], capture_error_mode="thread_local"):
            with torch.cuda.device(self.device):
                _lib.check(L.mural_indel_train_forward(self._h, g["genome"].handle, _lib.ptr(g["pos"]), _lib.ptr(g["meta"]), n,
                                                       _lib.ptr(self.blob), _lib.ptr(g["out"]), _lib.current_stream()))
                _lib.check(L.mural_ce_sum_grad(_lib.ptr(g["out"]), _lib.ptr(g["meta"]), n, self.model.n_class, _lib.ptr(self.loss_dev),
                                               _lib.ptr(g["dout"]), _lib.current_stream()))
                _lib.check(L.mural_indel_train_backward(self._h, _lib.ptr(self.blob), _lib.ptr(g["dout"]), _lib.ptr(self.grads),
                                                        _lib.current_stream()))
        with torch.cuda.graph(g["opt"], capture_error_mode="thread_local"):
            self._launch_optimizer(g["world"])
        return g

    def step(self, batch):
        """forward + CE(sum) + backward + all-reduce + clip + optimizer on a SiteBatch (labels in its meta).  On the graph path the
        returned tensor is the graph's static output buffer (overwritten by the next step of the same batch size)."""
        n = len(batch)
        if n < 2:
            return None                                      # training.py:415
        if self.use_graph and isinstance(batch, SiteBatch):
            _check_attached(self)
            g = self._graphs.get(n)
            if g is None and self._graph_warm.get(n, 0) >= 1 and len(self._graphs) < self.MAX_GRAPHS:
                g = self._graphs[n] = self._capture(batch)
            if g is not None and g["genome"] is batch.genome and g["world"] == self._world():
                g["pos"].copy_(batch.pos)
                g["meta"].copy_(batch.meta)
                if self._lr_on_dev != self.lr:
                    self._lr_dev.fill_(self.lr)
                    self._lr_on_dev = self.lr
                g["fb"].replay()
                self.n_forward += 1
                if g["world"] > 1:
                    torch.distributed.all_reduce(self.grads[:self.n_trainable], op=torch.distributed.ReduceOp.SUM)
                g["opt"].replay()
                self.opt_step += 1
                self.model.mark_dirty()
                return g["out"]
            self._graph_warm[n] = self._graph_warm.get(n, 0) + 1
        out = self.forward(batch)
        dout = torch.empty_like(out)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(out), _lib.ptr(batch.meta), n, self.model.n_class, _lib.ptr(self.loss_dev),
                                                    _lib.ptr(dout), _lib.current_stream()))
        self.backward(dout)
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(self.grads[:self.n_trainable], op=torch.distributed.ReduceOp.SUM)
            world = torch.distributed.get_world_size()
        self.opt_step += 1
        if self._lr_on_dev != self.lr:
            self._lr_dev.fill_(self.lr)
            self._lr_on_dev = self.lr
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mural_optimizer_step_dev(self.kind, _lib.ptr(self.blob), _lib.ptr(self.grads), _lib.ptr(self.m),
                                                           _lib.ptr(self.v), _lib.ptr(self.vmax), self.n_trainable, _lib.ptr(self._lr_dev),
                                                           self.weight_decay, _lib.ptr(self._opt_step_dev), self.max_norm,
                                                           1.0 / world if self.grad_average else 1.0, _lib.ptr(self.scratch),
                                                           _lib.current_stream()))
        self.model.mark_dirty()
        return out

    def sync_buffers(self):
        """Data parallel: average the rank-local BatchNorm running statistics over ranks (see TrainState.sync_buffers)."""
        d = torch.distributed
        if d.is_available() and d.is_initialized() and d.get_world_size() > 1 and self.n_blob > self.n_trainable:
            buf = self.blob[self.n_trainable:]
            d.all_reduce(buf, op=d.ReduceOp.SUM)
            buf.div_(d.get_world_size())
            self.model.mark_dirty()

    def sync_counters(self):
        """num_batches_tracked follows the number of BatchNorm calls: one per training forward, two for the reverse-strand
        stem (it is applied to the window and to its reverse complement, model_indel.py:155)."""
        d = self.n_forward - self._tracked_synced
        if d:
            stem_bn = self.model.conv[1] if getattr(self.model, "use_reverse", None) else None
            for mod in self.model.modules():
                if isinstance(mod, torch.nn.BatchNorm1d) and mod.num_batches_tracked is not None:
                    mod.num_batches_tracked += 2 * d if mod is stem_bn else d
            self._tracked_synced = self.n_forward


class _IndelTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, batch, *params):
        ctx.state = state
        return state.forward(batch)

    @staticmethod
    def backward(ctx, dout):
        st = ctx.state
        st.backward(dout)
        return (None, None) + tuple(g.clone() for g in st.grad_views)


def unet_train_forward(model, distal_input, distal_radius=None):
    """UNet_Small.forward in train() mode (differentiable): the reference's loop body runs unchanged on it."""
    if distal_radius is None:
        if isinstance(distal_input, SiteBatch):
            raise ValueError("distal_radius is required with a SiteBatch")
        distal_radius = distal_input.shape[2] // 2
    st = getattr(model, "_train_state", None)
    if st is None or st.distal_radius != int(distal_radius):
        st = IndelTrainState(model, distal_radius)
    return _IndelTrainFn.apply(st, distal_input, *st.params)


def load_pretrained(model, model_state, train_all=True, init_fc_with_pretrained=True, model_type="snv"):
    """Transfer-learning initialisation, training.py:289-320: load the pretrained state dict on the CPU copy (strict keys,
    `.layer.N` aliases included), move back, then the two switches.  As in the reference both switches only work when
    True for Network2 / UNet_Small: the partial-freeze and fc re-initialisation branches address `model.distal_fc`, which
    neither model defines (the SNV CLI forces train_all, mural_snv.py:103-106), so they end in the same AttributeError
    here; INDEL exits with the reference's messages (training.py:307,318)."""
    from .nn_utils import weights_init
    device = next(model.parameters()).device
    model.to(torch.device("cpu"))
    model.load_state_dict(model_state)
    model.to(device)
    if train_all:
        for param in model.parameters():
            param.requires_grad = True
    else:
        if model_type == "indel":
            raise SystemExit("Error: --train_all need used in commend line for INDEL, transfer learning for INDEL model need fine tune "
                             "all parameters !")
        for param in model.parameters():
            param.requires_grad = False
        model.local_fc[-1].weight.requires_grad = True
        model.local_fc[-1].bias.requires_grad = True
        model.distal_fc[-1].weight.requires_grad = True
        model.distal_fc[-1].bias.requires_grad = True
    if not init_fc_with_pretrained:
        if model_type == "indel":
            raise SystemExit("Error: --init_fc_with_pretrained need used in commend line for INDEL, transfer learning for INDEL model "
                             "need fine tune all parameters !")
        model.local_fc[-1].apply(weights_init)
        model.distal_fc[-1].apply(weights_init)
    return model


def validate_epoch(model, dataset, n_class=4, pred_batch_size=4096, segment_indices=None, kmer_list=(3, 5, 7), win_size_list=(100000, 500000),
                   printer=print):
    """The validation half of an epoch (training.py:454-520) on the device: prediction in emission order
    (`model_predict_m`), summed cross-entropy, then the reference's Evaluator metrics before calibration (k-mer
    correlations, regional score, regional window correlations) through mural_b200.evaluation — no pandas group-bys, no
    per-row python.  Calibrator fitting (JAX in the reference) stays with the caller.  Returns a dict."""
    import pandas as pd
    from .calibration import calibrate
    from .data import generate_site_batches
    from .evaluation import EvalData, Evaluator
    from .nn_utils import model_predict_m
    indel = dataset.model_type == "indel"
    if indel and tuple(kmer_list) == (3, 5, 7):
        kmer_list = (2, 4, 6)                                 # training.py:495: even k for indel (data_local has no centre column)
    segs = np.arange(len(dataset)) if segment_indices is None else np.asarray(segment_indices)
    was_training = model.training

    class _Batches:                                          # carries the window radius UNet_Small.forward needs with site records
        distal_radius = dataset.distal_radius

        def __iter__(self):
            return iter(generate_site_batches(dataset, 1 << 30, pred_batch_size, shuffle=False, segment_indices=segs))
    pred_y, total_loss = model_predict_m(model, _Batches(), None, dataset.genome.device, n_class, model_type=dataset.model_type)
    rows = np.concatenate([np.arange(dataset.batch_offsets[i], dataset.batch_offsets[i + 1]) for i in segs]) if len(segs) else np.zeros(0, np.int64)
    pos = torch.from_numpy(dataset.pos[rows]).to(pred_y.device)
    meta = torch.from_numpy(dataset.meta[rows]).to(pred_y.device)
    flank = dataset.genome.encode_local(pos, meta, dataset.local_radius, 1, dataset.model_type)   # data_local's us*/[mid]/ds* columns (prepare_local_data :400)
    if indel:                                                # EvalData's layout has a centre column; indel frames have none (evaluation.py)
        r = dataset.local_radius
        flank = torch.cat([flank[:, :r], torch.zeros_like(flank[:, :1]), flank[:, r:]], dim=1)
    prob = calibrate(pred_y.contiguous())                                        # F.softmax(pred_y, dim=1), training.py:463
    ev = Evaluator(EvalData(flank, meta, prob, f32=True), None, n_class, printer=printer)
    ev.evaluate_kmer(list(kmer_list))
    n = int(pred_y.shape[0])
    if n >= 10:
        ev.evaluate_regional_score(n, list(kmer_list)[:2])
    names, start, end, strand = dataset.position_info()
    chr_pos = pd.DataFrame({"chrom": np.asarray(names)[rows], "start": np.asarray(start)[rows], "end": np.asarray(end)[rows],
                            "strand": np.asarray(strand)[rows]})
    ev.evaluate_regional_corr(chr_pos, list(win_size_list))
    printer('Validation Loss: ', total_loss / max(1, n))
    if was_training:
        model.train()
    return {"valid_loss": total_loss / max(1, n), "valid_size": n, **ev.metrics}


def _steps_this_epoch(dataset, segs, batch_size, world, rank):
    """Data parallel: rank r trains on segments segs[r::world]; every rank must enter the gradient all-reduce the same number
    of times, so the number of steps is the minimum over ranks of the batches a rank will draw (full batches + a tail of at
    least 2 sites, training.py:415)."""
    mine = segs[rank::world]
    n = int(sum(int(dataset.batch_sizes[i]) for i in mine))
    steps = n // batch_size + (1 if n % batch_size >= 2 else 0)
    if world > 1:
        d = torch.distributed
        t = torch.tensor([steps], dtype=torch.int64, device=dataset.genome.device if d.get_backend() == "nccl" else "cpu")
        d.all_reduce(t, op=d.ReduceOp.MIN)
        steps = int(t.item())
    return mine, steps


def train_epochs(model, dataset, epochs, batch_size, sampled_segments=10, optim="Adam", lr=1e-3, weight_decay=0.0, LR_gamma=0.5,
                 min_lr=1e-6, restart_lr=1e-4, seed=0, print_every=1000, segment_indices=None, valid_indices=None, history=None,
                 pred_batch_size=4096, printer=print, lr_scheduler="StepLR", config=None):
    """The hot loop of training.py:387-452 on site records; returns per-epoch mean losses.  With `valid_indices` (segment
    indices of `dataset` held out for validation, as `random_split` does at training.py:152-168) every epoch ends with
    `validate_epoch` and its dict is appended to `history`.  `lr_scheduler`: 'StepLR' | 'StepLR2' | 'ROP' (training.py:364-371;
    StepLR2 restarts every epoch at restart_lr :396-398, ROP steps on the validation loss :553-554); a reference `config` dict
    (optim, learning_rate, weight_decay, LR_gamma, lr_scheduler, min_lr, restart_lr, batch_size) overrides the keyword
    arguments.  With torch.distributed initialised the training segments are dealt round-robin to the ranks and every rank
    runs the same number of steps per epoch.  Calibrator fitting and checkpoint bookkeeping stay with the caller (out of
    scope)."""
    from .data import generate_site_batches
    if config is not None:
        optim, lr, weight_decay = config.get("optim", optim), config.get("learning_rate", lr), config.get("weight_decay", weight_decay)
        LR_gamma, lr_scheduler = config.get("LR_gamma", LR_gamma), config.get("lr_scheduler", lr_scheduler)
        min_lr, restart_lr, batch_size = config.get("min_lr", min_lr), config.get("restart_lr", restart_lr), config.get("batch_size", batch_size)
    if optim not in OPTIMIZERS:
        raise ValueError("Error: unsupported optimization method %s" % optim)
    if len(dataset.label) and int(dataset.label.max()) >= model.n_class:
        raise ValueError("labels must be < n_class=%d (CrossEntropyLoss fails on out-of-range targets); max label %d" %
                         (model.n_class, int(dataset.label.max())))
    st = getattr(model, "_train_state", None)
    if st is not None and (st.kind != OPTIMIZERS[optim] or st.weight_decay != float(weight_decay) or st.opt_step == 0):
        st = None                       # a state auto-created by a train-mode forward (Adam, lr 1e-3) must not shadow the arguments
    if st is None:
        st = (IndelTrainState(model, dataset.distal_radius, optim, lr, weight_decay, seed=seed) if dataset.model_type == "indel"
              else TrainState(model, optim, lr, weight_decay, seed=seed))
    st.lr = float(lr)
    d = torch.distributed
    world = d.get_world_size() if d.is_available() and d.is_initialized() else 1
    rank = d.get_rank() if world > 1 else 0
    segs_all = np.arange(len(dataset)) if segment_indices is None else np.asarray(segment_indices)
    segs, n_steps = _steps_this_epoch(dataset, segs_all, batch_size, world, rank)
    train_size = int(sum(int(dataset.batch_sizes[i]) for i in segs_all))
    sched = make_scheduler({"lr_scheduler": lr_scheduler, "learning_rate": lr, "batch_size": batch_size, "LR_gamma": LR_gamma,
                            "min_lr": min_lr, "restart_lr": restart_lr}, max(train_size, batch_size))
    per_batch = lr_scheduler in ("StepLR", "StepLR2")
    losses = []
    for epoch in range(epochs):
        model.train()
        st.loss_dev.zero_()
        n_sites = k = 0
        if epoch > 0 and lr_scheduler == "StepLR2":
            sched.lr = st.lr = restart_lr                   # training.py:396-398
        for batch in generate_site_batches(dataset, sampled_segments, batch_size, shuffle=True, seed=seed + epoch, segment_indices=segs):
            if len(batch) < 2:
                continue
            if k >= n_steps:
                break                                        # keeps the ranks' all-reduce counts equal
            st.step(batch)
            k += 1
            n_sites += len(batch)
            if per_batch:
                st.lr = sched.step()
        losses.append(float(st.loss_dev.item()) / max(1, n_sites))
        st.sync_counters()
        st.sync_buffers()
        if valid_indices is not None:
            h = validate_epoch(model, dataset, model.n_class, pred_batch_size, valid_indices, printer=printer)
            h["epoch"], h["train_loss"] = epoch, losses[-1]
            if lr_scheduler == "ROP":
                st.lr = sched.step(h["valid_loss"])          # training.py:553-554
            if history is not None:
                history.append(h)
    return losses
