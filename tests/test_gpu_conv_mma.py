"""GPU: the tensor-core conv kernels of the fp32-equivalent / training paths (csrc/snv_conv_mma.cu) vs torch fp64:
one BatchNorm-affine -> Conv1d(32,32,3,padding=1) layer in the [n*L, 32] row layout with per-site zero padding applied AFTER the
affine (nn.Sequential(BatchNorm1d, Conv1d(padding)), model_snv.py:350-353,804), residuals, ReLU — and its weight gradient.
Ragged shapes: site lengths that do not divide the 128-row tile, 1-row sites, a single site, tails."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(48, 67), (48, 134), (12, 667), (1, 7), (5, 1), (3, 2), (257, 8), (64, 20), (7, 129), (2, 128), (31, 23)]


def _ref_conv(x, W, bias, a, b, relu_in, relu_out, r1, r2):
    """x [n, L, 32] -> [n, L, 32] in float64."""
    u = (torch.relu(x) if relu_in else x) * a + b
    y = torch.nn.functional.conv1d(u.transpose(1, 2), W, bias, padding=1).transpose(1, 2)
    if r1 is not None:
        y = y + r1
    if r2 is not None:
        y = y + r2
    return torch.relu(y) if relu_out else y


@pytest.mark.parametrize("n,L", SHAPES)
def test_conv32_layer_impls_vs_fp64(n, L):
    from mural_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(n * 1000 + L)
    x = torch.randn(n, L, 32, generator=g, dtype=torch.float64) * 3
    W = torch.randn(32, 32, 3, generator=g, dtype=torch.float64) * 0.2          # [co][ci][tap]
    bias, a, b = (torch.randn(32, generator=g, dtype=torch.float64) for _ in range(3))
    r1, r2 = torch.randn(n, L, 32, generator=g, dtype=torch.float64), torch.randn(n, L, 32, generator=g, dtype=torch.float64)
    dev = lambda t: t.to(torch.float32).cuda().contiguous()
    Wt = dev(W.permute(2, 1, 0))                                                # [tap][ci][co]
    xd, bd, ad, bbd, r1d, r2d = dev(x), dev(bias), dev(a), dev(b), dev(r1), dev(r2)
    f64 = lambda t: t.to(torch.float64).cpu()
    for relu_in, relu_out, use_r in ((1, 0, 2), (0, 1, 0), (1, 0, 1)):
        ref = _ref_conv(f64(xd), f64(Wt).permute(2, 1, 0), f64(bd), f64(ad), f64(bbd), relu_in, relu_out,
                        f64(r1d) if use_r >= 1 else None, f64(r2d) if use_r >= 2 else None)
        scale = float(ref.abs().max())
        for impl, tol in ((0, 2e-6), (1, 3e-5), (2, 2e-6)):
            out = torch.full((n, L, 32), float("nan"), dtype=torch.float32, device="cuda")
            _lib.check(lib.mural_conv32_layer(_lib.ptr(xd), _lib.ptr(out), _lib.ptr(r1d) if use_r >= 1 else None,
                                              _lib.ptr(r2d) if use_r >= 2 else None, n, L, _lib.ptr(Wt), _lib.ptr(bd), _lib.ptr(ad), _lib.ptr(bbd),
                                              relu_in, relu_out, impl, _lib.current_stream()))
            err = float((out.cpu().to(torch.float64) - ref).abs().max()) / scale
            assert err < tol, (n, L, impl, relu_in, relu_out, use_r, err)


@pytest.mark.parametrize("n,L", SHAPES)
def test_conv32_wgrad_vs_fp64(n, L):
    from mural_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(n * 77 + L)
    x = torch.randn(n, L, 32, generator=g, dtype=torch.float64) * 2
    dy = torch.randn(n, L, 32, generator=g, dtype=torch.float64)
    a, b = torch.randn(32, generator=g, dtype=torch.float64), torch.randn(32, generator=g, dtype=torch.float64)
    dev = lambda t: t.to(torch.float32).cuda().contiguous()
    xd, dyd, ad, bd = dev(x), dev(dy), dev(a), dev(b)
    f64 = lambda t: t.to(torch.float64).cpu()
    for relu_in in (0, 1):
        W = torch.zeros(32, 32, 3, dtype=torch.float64, requires_grad=True)
        bias = torch.zeros(32, dtype=torch.float64, requires_grad=True)
        y = _ref_conv(f64(xd), W, bias, f64(ad), f64(bd), relu_in, 0, None, None)
        (y * f64(dyd)).sum().backward()
        for impl in (0, 1):
            G = torch.zeros(32 * 32 * 3 + 32, dtype=torch.float32, device="cuda")
            _lib.check(lib.mural_conv32_wgrad(_lib.ptr(xd), _lib.ptr(dyd), n, L, relu_in, _lib.ptr(ad), _lib.ptr(bd), _lib.ptr(G),
                                              C_ptr(G, 32 * 32 * 3), impl, _lib.current_stream()))
            gw = G[:32 * 32 * 3].view(32, 32, 3).cpu().to(torch.float64)
            gb = G[32 * 32 * 3:].cpu().to(torch.float64)
            sw, sb = max(1e-6, float(W.grad.abs().max())), max(1e-6, float(bias.grad.abs().max()))
            assert float((gw - W.grad).abs().max()) / sw < 3e-5, (n, L, impl, relu_in)
            assert float((gb - bias.grad).abs().max()) / sb < 3e-5, (n, L, impl, relu_in)


def C_ptr(t, offset_elems):
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())
