"""IAO PTQ inference (BASELINE.json configs[4], iao/main.py:109-142 calibration + :511-519 eval): the frozen fast path
(iao.freeze_inference: weights folded / quantized / packed once, ReLUs folded into the operand packer and the QuantAdd
kernel, observers of QuantAdd frozen) must be bit-identical to the plain eval forward of the engine, and the engine's eval
logits must match the CPU oracle's on the same calibration + input (2 x 3 x 224 x 224, full-width ResNet-18)."""
import copy

import pytest
import torch

from tests.oracle_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PTQ = dict(a_bits=8, w_bits=8, q_type=0, q_level=0, weight_observer=0, bn_fuse=True, pretrained_model=True, ptq=True,
           percentile=0.999999)


def _prepared(widths, hw, seed=4):
    from harness import models as zoo, train as H
    torch.manual_seed(seed)
    base = zoo.init_like_reference(zoo.ResNet(widths=widths))
    with torch.no_grad():                      # a "pretrained" model: non-trivial BatchNorm statistics
        for m in base.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
    eng = H.prepare_engine(copy.deepcopy(base), "iao", **PTQ).to(DEV)
    ora = H.prepare_oracle(copy.deepcopy(base), "iao", **PTQ)
    g = torch.Generator().manual_seed(seed + 1)
    calib = torch.randn(2, 3, hw, hw, generator=g)
    x = torch.randn(2, 3, hw, hw, generator=g)
    return eng, ora, calib, x


def test_frozen_inference_is_bit_identical_and_graph_safe():
    from micronet_b200 import iao, _lib as L
    eng, _, calib, x = _prepared((16, 32, 64, 128), 64)
    with torch.no_grad():
        eng.train(); eng(calib.to(DEV)); eng.eval()
        plain = eng(x.to(DEV)).clone()
        iao.freeze_inference(eng)
        n_identity = sum(isinstance(m, torch.nn.Identity) for m in eng.modules())
        frozen1 = eng(x.to(DEV)).clone()
        frozen2 = eng(x.to(DEV)).clone()          # second call: cached weights / images
        iao.freeze_inference(eng, enable=False)
        back = eng(x.to(DEV)).clone()
    assert torch.equal(plain, frozen1) and torch.equal(plain, frozen2) and torch.equal(plain, back)
    assert n_identity > sum(isinstance(m, torch.nn.Identity) for m in eng.modules())   # the ReLUs came back
    L.tc_check()


def test_eval_logits_match_the_oracle_at_224():
    from micronet_b200 import iao, _lib as L
    eng, ora, calib, x = _prepared((64, 128, 256, 512), 224)
    with torch.no_grad():
        eng.train(); eng(calib.to(DEV)); eng.eval()
        ora.train(); ora(calib); ora.eval()
        iao.freeze_inference(eng)
        ye = eng(x.to(DEV)).cpu()
        yo = ora(x)
    # the calibrated activation ranges are percentile statistics of fp32 conv outputs: identical selection rule on both
    # sides; a handful of activation levels sit on the other side of a rounding tie (conv summation order)
    assert rel_err(ye, yo) <= 2e-3, rel_err(ye, yo)
    for (n, be), (_, bo) in zip(sorted(eng.state_dict().items()), sorted(ora.state_dict().items())):
        if n.endswith("activation_quantizer.scale"):
            assert rel_err(be, bo) <= 1e-5, n
    L.tc_check()
