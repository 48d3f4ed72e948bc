"""fp32 first-layer convolution on the tensor-core im2col path (mnb_fconv2d_fwd_tc / _wgrad_tc) against an fp64
convolution: the 3-piece bf16 split keeps fp32 accuracy (tolerance 2e-6 of the tensor's max, SURVEY 8c: 1e-5)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as TF

from tests.oracle_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

SHAPES = [  # B, C, H, W, K, R, bias
    (8, 3, 32, 32, 256, 5, True),     # NIN-GC first conv
    (8, 3, 32, 32, 192, 5, True),     # NIN first conv
    (5, 3, 32, 32, 64, 3, False),     # CIFAR ResNet first conv
    (3, 1, 16, 16, 24, 7, True),
    (2, 5, 8, 16, 40, 5, True),       # 125 im2col columns
    (160, 3, 32, 32, 256, 5, True),   # more tiles than CTAs: persistent loop, both accumulators
    (2, 4, 64, 64, 16, 3, False),
    (1, 2, 1, 128, 8, 1, True),
]


@pytest.mark.parametrize("shape", SHAPES, ids=[str(s) for s in SHAPES])
def test_float_conv_tc_matches_fp64(shape):
    from micronet_b200 import _lib as L
    from micronet_b200.fused import EngineFloatConv2d
    B, C, H, W, K, R, has_bias = shape
    torch.manual_seed(sum(shape[:6]))
    conv = EngineFloatConv2d(C, K, R, 1, R // 2, bias=has_bias).to(DEV)
    x = (torch.randn(B, C, H, W) * 1.7 + 0.3).to(DEV)
    go = torch.randn(B, K, H, W, device=DEV)
    xr = x.double().requires_grad_(True)
    wr = conv.weight.detach().double().requires_grad_(True)
    br = conv.bias.detach().double().requires_grad_(True) if has_bias else None
    yr = TF.conv2d(xr, wr, br, 1, R // 2)
    yr.backward(go.double())

    xg = x.clone().requires_grad_(True)
    y = conv(xg)
    y.backward(go)
    L.tc_check()
    assert rel_err(y.detach(), yr.detach()) < 2e-6
    assert rel_err(conv.weight.grad, wr.grad) < 2e-6
    assert rel_err(xg.grad, xr.grad) < 1e-5       # ATen data gradient (not on the hot path: first layer)
    if has_bias:
        assert rel_err(conv.bias.grad, br.grad) < 2e-6


def test_float_conv_falls_back_outside_cover():
    """48 x 48 rows do not tile into 128 positions: the module then runs ATen's convolution"""
    from micronet_b200.fused import EngineFloatConv2d
    torch.manual_seed(0)
    conv = EngineFloatConv2d(3, 16, 3, 1, 1).to(DEV)
    x = torch.randn(2, 3, 48, 48, device=DEV)
    y = conv(x)
    assert rel_err(y.detach(), TF.conv2d(x, conv.weight, conv.bias, 1, 1).detach()) < 1e-6
    y.sum().backward()
    assert conv.weight.grad is not None and torch.isfinite(conv.weight.grad).all()
