"""Fused BatchNorm2d + binarizer (micronet_b200.fused) against nn.BatchNorm2d (fp64, CPU) followed by the
oracle's binarizer (WB:11-36 / WB:79-94): outputs, saturate-STE gradients, parameter gradients, running stats."""
import copy

import pytest
import torch
import torch.nn as nn

from oracle import reference_port as O
from tests.oracle_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

SHAPES = [  # B, C, H, W  (NIN-GC planes + ragged ones that take the scalar kernels)
    (8, 192, 32, 32), (8, 96, 16, 16), (16, 192, 8, 8), (3, 10, 1, 1), (5, 7, 3, 5), (2, 33, 6, 6), (64, 160, 32, 32),
]


def _pair(c, seed):
    torch.manual_seed(seed)
    bn = nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c) + 0.5)
        bn.bias.copy_(torch.randn(c) * 0.3)
        bn.running_mean.copy_(torch.randn(c) * 0.1)
        bn.running_var.copy_(torch.rand(c) + 0.5)
    return bn


def _reference(bn, x, go):
    """fp64 BatchNorm + the oracle's binarizer; returns y, bn output, dx, dgamma, dbeta"""
    ref = copy.deepcopy(bn).double()
    xr = x.double().requires_grad_(True)
    pre = ref(xr)
    y = O.wb_binarize_activation(pre)
    y.backward(go.double())
    return y.detach(), pre.detach(), xr.grad, ref.weight.grad, ref.bias.grad, ref


@pytest.mark.parametrize("shape", SHAPES, ids=[str(s) for s in SHAPES])
@pytest.mark.parametrize("training", [True, False], ids=["train", "eval"])
def test_fused_bn_binarize_matches_bn_then_oracle_binarizer(shape, training):
    from micronet_b200.fused import BatchNormBinarize2d
    B, C, H, W = shape
    bn = _pair(C, sum(shape))
    bn.train(training)
    x = torch.randn(B, C, H, W) * 1.5 + 0.2
    go = torch.randn(B, C, H, W)
    y_r, pre, dx_r, dg_r, db_r, ref = _reference(bn, x, go)

    fused = BatchNormBinarize2d(C)
    fused.load_state_dict(bn.state_dict())
    fused = fused.to(DEV).train(training)
    xg = x.to(DEV).requires_grad_(True)
    y = fused(xg)
    y.backward(go.to(DEV))

    # sign(): only elements whose batch-norm output is within fp32 rounding of 0 may differ
    differ = (y.detach().cpu().double() != y_r)
    assert not (differ & (pre.abs() > 1e-5)).any()
    assert set(torch.unique(y.detach()).tolist()) <= {-1.0, 1.0}
    # saturate STE: elements within rounding of |bn| = 1 may take the other branch; mask them out of dx
    edge = ((pre.abs() - 1).abs() < 1e-5)
    n_edge = int(edge.sum())
    assert n_edge <= 1 + x.numel() // 20000
    if n_edge == 0:
        assert rel_err(xg.grad, dx_r) < 2e-5
        assert rel_err(fused.weight.grad, dg_r) < 2e-5 and rel_err(fused.bias.grad, db_r) < 2e-5
    else:  # an edge element moves the channel sums by at most |go| each
        assert rel_err(fused.bias.grad, db_r) < 1e-3
    if training:
        assert rel_err(fused.running_mean, ref.running_mean) < 1e-6
        assert rel_err(fused.running_var, ref.running_var) < 1e-6
        assert int(fused.num_batches_tracked) == int(ref.num_batches_tracked) == 1
    else:
        assert torch.equal(fused.running_mean.cpu(), bn.running_mean)


def test_prepare_fuse_bn_rewrites_pairs_and_keeps_state_dict_layout():
    import micronet_b200 as E
    from harness import models as zoo
    from micronet_b200.fused import BatchNormBinarize2d
    torch.manual_seed(0)
    base = zoo.NINGC()
    plain = E.wbwtab.prepare(base, A=2, W=3)
    fused = E.wbwtab.prepare(base, A=2, W=3, fuse_bn=True)
    assert list(plain.state_dict().keys()) == list(fused.state_dict().keys())
    n_aq = sum(isinstance(m, E.wbwtab.ActivationQuantizer) for m in plain.modules())
    assert sum(isinstance(m, BatchNormBinarize2d) for m in fused.modules()) == n_aq
    assert not any(isinstance(m, E.wbwtab.ActivationQuantizer) for m in fused.modules())
    # A != 2 keeps the ReLU path untouched
    relu = E.wbwtab.prepare(base, A=32, W=2, fuse_bn=True)
    assert not any(isinstance(m, BatchNormBinarize2d) for m in relu.modules())


def test_fused_model_step_matches_unfused_engine_model():
    """whole NIN-GC step, fused vs unfused engine models: same loss and (up to sign flips at |bn| ~ 0)
    the same gradients"""
    import micronet_b200 as E
    from harness import models as zoo
    torch.manual_seed(1)
    base = zoo.NINGC()
    zoo.init_like_reference(base)
    x = torch.randn(16, 3, 32, 32).to(DEV)
    t = torch.randint(0, 10, (16,)).to(DEV)
    out = {}
    for name, kw in (("plain", {}), ("fused", {"fuse_bn": True})):
        m = E.wbwtab.prepare(base, A=2, W=3, **kw).to(DEV).train()
        loss = nn.functional.cross_entropy(m(x), t)
        loss.backward()
        out[name] = (loss.item(), {n: p.grad.clone() for n, p in m.named_parameters()},
                     {n: b.clone() for n, b in m.named_buffers()})
    assert abs(out["plain"][0] - out["fused"][0]) < 2e-3 * abs(out["plain"][0])
    for n, b in out["plain"][2].items():
        if b.dtype.is_floating_point:
            assert rel_err(out["fused"][2][n], b) < 1e-3, n
    worst = max(rel_err(out["fused"][1][n], g) for n, g in out["plain"][1].items())
    assert worst < 5e-2, worst
