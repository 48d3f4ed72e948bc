"""Fused DoReFa producer (fused.BatchNormReluQuant2d -> mnb_bn_relu_quant_pack_fwd): BatchNorm2d + ReLU + the next conv's
activation quantizer + operand packing in one pass, against the un-fused engine path and a torch restatement.

BatchNorm's fp32 arithmetic is not bit-reproducible between implementations (fma vs mul+add, statistics summed in another
order), so integer levels are compared with the tie-excuse rule of SURVEY 7.2.1: a level may differ by one only where the
pre-rounding value sits within 1e-4 of a rounding tie, on at most 1e-4 of the elements."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as TF

from tests.oracle_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _unpack(planes, B, C, H, W):
    t = planes.view(torch.bfloat16).view(B, C // 8, H, W, 8).float()
    return t.permute(0, 1, 4, 2, 3).reshape(B, C, H, W)


def _shuffle(x, groups):
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).contiguous().view(b, c, h, w)


@pytest.mark.parametrize("a_bits", [4, 8])
@pytest.mark.parametrize("sg", [1, 4])
def test_producer_levels_and_mask(a_bits, sg):
    from micronet_b200.fused import BatchNormReluQuant2d
    B, C, H, W = 6, 64, 16, 16
    torch.manual_seed(a_bits * 10 + sg)
    bn = BatchNormReluQuant2d(C).to(DEV).train()
    bn.a_bits, bn.out_shuffle_groups = a_bits, sg
    ref = nn.BatchNorm2d(C).to(DEV).train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C) + 0.5); bn.bias.copy_(torch.randn(C) * 0.5)
        ref.weight.copy_(bn.weight); ref.bias.copy_(bn.bias)
    x = (torch.randn(B, C, H, W) * 4 + 1).to(DEV).requires_grad_(True)
    y = bn(x)
    packed, bits = y._mnb_pk_q
    assert bits == a_bits
    lev = _unpack(packed, B, C, H, W)
    xr = x.detach().clone().requires_grad_(True)
    yb = ref(xr)
    yr = TF.relu(yb)
    n = float(2 ** a_bits - 1)
    pre = torch.clamp(yr * 0.1, 0, 1) * n
    want = torch.floor(pre + 0.5)
    if sg > 1:
        want, pre = _shuffle(want, sg), _shuffle(pre, sg)
    diff = (lev - want.detach()).abs()
    near_tie = ((pre.detach() + 0.5) - torch.floor(pre.detach() + 0.5)).abs().minimum(
        1 - ((pre.detach() + 0.5) - torch.floor(pre.detach() + 0.5)).abs()) < 1e-3
    assert diff.max().item() <= 1.0
    assert ((diff > 0) & ~near_tie).sum().item() == 0, "a level differs away from a rounding tie"
    assert (diff > 0).sum().item() <= max(2, int(1e-4 * diff.numel()))
    assert torch.equal(bn.running_mean, ref.running_mean) or rel_err(bn.running_mean, ref.running_mean) <= 1e-6
    # backward: g is what the consuming conv's data gradient delivers (already x 0.1), in the shuffled channel order
    g = torch.randn(B, C, H, W, device=DEV)
    y.backward(g)
    g_own = g if sg == 1 else g.view(B, C // sg, sg, H, W).transpose(1, 2).contiguous().view(B, C, H, W)   # inverse shuffle
    mask = ((yb > 0) & (yb * 0.1 <= 1)).float()
    (yb * (g_own * mask).detach()).sum().backward()
    # an element whose mask bit sits exactly on a boundary (bn = 0 or 10 to within rounding) may differ: at most two do
    d = (x.grad - xr.grad).abs()
    assert (d > 1e-4 * xr.grad.abs().max()).sum().item() <= 2
    assert rel_err(bn.weight.grad, ref.weight.grad) <= 1e-3 and rel_err(bn.bias.grad, ref.bias.grad) <= 1e-3


def _block(cin, cout, k, groups=1, shuffle=0, sgroups=1):
    from harness.models import ConvBNReLU
    return ConvBNReLU(cin, cout, k, 1, k // 2, groups=groups, channel_shuffle=shuffle, shuffle_groups=sgroups)


@pytest.mark.parametrize("a_bits", [4, 8])
def test_fused_blocks_match_the_unfused_engine(a_bits):
    import micronet_b200 as E
    from micronet_b200.fused import BatchNormReluQuant2d
    torch.manual_seed(3)
    base = nn.Sequential(_block(3, 64, 5), _block(64, 64, 1, groups=2), _block(64, 128, 3, groups=4, shuffle=1, sgroups=2),
                         _block(128, 128, 1, groups=4, shuffle=1, sgroups=4))
    for m in base.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.xavier_uniform_(m.weight)
    plain = E.dorefa.prepare(copy.deepcopy(base), inplace=True, a_bits=a_bits, w_bits=a_bits).to(DEV).train()
    fused = E.dorefa.prepare(copy.deepcopy(base), inplace=True, a_bits=a_bits, w_bits=a_bits, fuse=True).to(DEV).train()
    assert sum(isinstance(m, BatchNormReluQuant2d) for m in fused.modules()) == 3
    x = torch.randn(8, 3, 16, 16, device=DEV)
    go = torch.randn(8, 128, 16, 16, device=DEV)
    outs = []
    for net in (plain, fused):
        y = net(x)
        y.backward(go)
        outs.append(y.detach())
    # a handful of activation levels may sit on the other side of a rounding tie (BatchNorm arithmetic): directions agree
    cos = TF.cosine_similarity(outs[0].flatten(), outs[1].flatten(), dim=0).item()
    assert cos > 0.9995, cos
    gp = dict(plain.named_parameters())
    for n, p in fused.named_parameters():
        if n.endswith("conv.bias"):
            continue      # bias in front of a training-mode BatchNorm: mathematically zero gradient, both sides hold noise
        c = TF.cosine_similarity(p.grad.flatten(), gp[n].grad.flatten(), dim=0).item()
        assert c > 0.995, (n, c)
    assert plain.state_dict().keys() == fused.state_dict().keys()
    from micronet_b200 import _lib as L
    L.tc_check()
