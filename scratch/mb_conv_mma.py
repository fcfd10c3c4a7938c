"""Isolated timing of the conv / wgrad kernels through the parity hooks (CUDA events, L2-sized working sets)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mural_b200 import _lib
lib = _lib.lib()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n, L in ((4096, 134), (4096, 67), (4096, 20), (128, 134), (128, 7)):
    x = torch.randn(n, L, 32, device="cuda"); r1 = torch.randn_like(x); r2 = torch.randn_like(x); out = torch.empty_like(x)
    Wt = torch.randn(3, 32, 32, device="cuda") * .2; b = torch.randn(32, device="cuda"); a = torch.randn(32, device="cuda"); bb = torch.randn(32, device="cuda")
    G = torch.zeros(32 * 32 * 3 + 32, device="cuda")
    res = []
    for impl in (0, 1, 2):
        us = t(lambda: _lib.check(lib.mural_conv32_layer(_lib.ptr(x), _lib.ptr(out), _lib.ptr(r1), _lib.ptr(r2), n, L, _lib.ptr(Wt), _lib.ptr(b), _lib.ptr(a), _lib.ptr(bb), 1, 0, impl, _lib.current_stream())))
        res.append("conv impl%d %.1f us" % (impl, us))
    for impl in (0, 1):
        us = t(lambda: _lib.check(lib.mural_conv32_wgrad(_lib.ptr(x), _lib.ptr(r1), n, L, 1, _lib.ptr(a), _lib.ptr(bb), _lib.ptr(G), ctypes.c_void_p(G.data_ptr() + 4 * 3072), impl, _lib.current_stream())))
        res.append("wgrad impl%d %.1f us" % (impl, us))
    rows = n * L
    print("n=%d L=%d rows=%d (%.0f MB per tensor): " % (n, L, rows, rows * 128 / 1e6) + " | ".join(res))
