"""IAO activation-only wrappers (IAO:1160-1498: QuantReLU / QuantLeakyReLU / QuantSigmoid / QuantMaxPool2d /
QuantAvgPool2d / QuantAdaptiveAvgPool2d / QuantAdd) on the engine's fused quantizer kernel against the oracle:
observer state and scale bit-exact, outputs and input gradients within 1e-5, over training steps (EMA observers,
first-call branch) and one eval step; symmetric and asymmetric quantizers, QAT and PTQ (percentile) observers."""
import copy

import pytest
import torch
import torch.nn as nn

from tests.oracle_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

OPS = {
    "relu": (lambda E, kw: E.iao.QuantReLU(**kw), lambda: nn.ReLU()),
    "leaky_relu": (lambda E, kw: E.iao.QuantLeakyReLU(negative_slope=0.1, **kw), lambda: nn.LeakyReLU(0.1)),
    "sigmoid": (lambda E, kw: E.iao.QuantSigmoid(**kw), lambda: nn.Sigmoid()),
    "max_pool": (lambda E, kw: E.iao.QuantMaxPool2d(3, stride=2, padding=1, **kw), lambda: nn.MaxPool2d(3, 2, 1)),
    "avg_pool": (lambda E, kw: E.iao.QuantAvgPool2d(2, stride=2, **kw), lambda: nn.AvgPool2d(2, 2)),
    "adaptive_avg_pool": (lambda E, kw: E.iao.QuantAdaptiveAvgPool2d((1, 1), **kw), lambda: nn.AdaptiveAvgPool2d((1, 1))),
}
QUANT = {
    "sym8": dict(a_bits=8, q_type=0),
    "asym8": dict(a_bits=8, q_type=1),
    "sym4": dict(a_bits=4, q_type=0),
    "ptq": dict(a_bits=8, q_type=0, ptq=True, percentile=0.99),
}


def _steps(seed, shape, n=3):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(shape, generator=g) * (2.0 + 0.5 * i), torch.randn(shape, generator=g)) for i in range(n)]


def _check_state(e, o):
    es, os_ = e.state_dict(), o.state_dict()
    eq, oq = e.activation_quantizer, o.activation_quantizer
    assert torch.equal(eq.scale.cpu().view(-1), oq.scale.view(-1)), "scale"
    assert torch.equal(eq.zero_point.cpu().view(-1), oq.zero_point.view(-1)), "zero_point"
    assert torch.equal(eq.observer.min_val.cpu().view(-1), oq.observer.min_val.view(-1)), "observer.min_val"
    assert torch.equal(eq.observer.max_val.cpu().view(-1), oq.observer.max_val.view(-1)), "observer.max_val"
    assert len(es) >= 4 and len(os_) >= 4


@pytest.mark.parametrize("quant", sorted(QUANT))
@pytest.mark.parametrize("op", sorted(OPS))
def test_quantized_activation_wrapper_matches_oracle(op, quant):
    import micronet_b200 as E
    from oracle import reference_port as O
    kw = QUANT[quant]
    e = OPS[op][0](E, kw).to(DEV)
    o = O.IaoQuantThenOp(OPS[op][1](), **kw)
    for i, (x, go_full) in enumerate(_steps(hash((op, quant)) % 1000, (3, 8, 10, 10))):
        training = i < 2
        e.train(training); o.train(training)
        xe = x.to(DEV).requires_grad_(True)
        xo = x.clone().requires_grad_(True)
        ye, yo = e(xe), o(xo)
        assert rel_err(ye.detach(), yo.detach()) <= 1e-5, (op, quant, i)
        go = go_full if go_full.shape == yo.shape else torch.randn(yo.shape, generator=torch.Generator().manual_seed(i))
        ye.backward(go.to(DEV)); yo.backward(go)
        assert rel_err(xe.grad, xo.grad) <= 1e-5, (op, quant, i)
        _check_state(e, o)


@pytest.mark.parametrize("quant", sorted(QUANT))
def test_quant_add_matches_oracle(quant):
    import micronet_b200 as E
    from oracle import reference_port as O
    kw = QUANT[quant]
    e = E.iao.QuantAdd(**kw).to(DEV)
    o = O.IaoQuantAdd(**kw)
    steps_a, steps_b = _steps(31, (2, 8, 6, 6)), _steps(32, (2, 8, 6, 6))
    for i, ((a, go), (b, _)) in enumerate(zip(steps_a, steps_b)):
        training = i < 2
        e.train(training); o.train(training)
        b = torch.relu(b) * 1.7          # the shortcut branch has a different range than the residual branch
        ae, be = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        ao, bo = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ye, yo = e(ae, be), o(ao, bo)
        assert rel_err(ye.detach(), yo.detach()) <= 1e-5, (quant, i)
        ye.backward(go.to(DEV)); yo.backward(go)
        assert rel_err(ae.grad, ao.grad) <= 1e-5 and rel_err(be.grad, bo.grad) <= 1e-5, (quant, i)
        _check_state(e, o)
        for name in ("observer_res", "observer_shortcut"):
            eo, oo = getattr(e, name), getattr(o, name)
            assert torch.equal(eo.min_val.cpu().view(-1), oo.min_val.view(-1)), name
            assert torch.equal(eo.max_val.cpu().view(-1), oo.max_val.view(-1)), name
