"""Inference-graph converters (micronet_b200.bn_fuse) against fixtures produced by the reference's own
bn_fuse scripts (tests/golden/make_golden_bnfuse.py): the converted model must have the same module types in the same
places and a bit-identical state_dict.  Host logic only - runs without a GPU."""
import numpy as np
import pytest
import torch

from tests.oracle_util import load_golden

CFG = [16, 16, 16, 32, 32, 32, 64, 64]
KINDS = {"QuantConv2d", "Conv2d", "Identity", "BatchNorm2d", "QuantBNFuseConv2d", "ActivationQuantizer"}


def _types(model):
    return {n: type(m).__name__ for n, m in model.named_modules() if type(m).__name__ in KINDS}


def _gold_types(gold):
    out = {}
    for item in gold["types"]:
        n, t = str(item).split(":")
        if t in KINDS:
            out[n] = t
    return out


def _float_model(gold):
    from harness import models as zoo
    m = zoo.NINGC(CFG)
    m.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("init.")})
    return m


def _compare_state(model, gold):
    want = {k[6:]: v for k, v in gold.items() if k.startswith("fused.")}
    got = model.state_dict()
    assert set(got.keys()) == set(want.keys())
    for k, v in want.items():
        assert np.array_equal(got[k].detach().cpu().numpy(), v), k


def converted_wbwtab(gold, W):
    import micronet_b200 as E
    from micronet_b200 import bn_fuse
    inf = E.wbwtab.prepare(_float_model(gold), inplace=True, A=2, W=W, quant_inference=True)
    return bn_fuse.wbwtab_model_bn_fuse(inf, W=W, inplace=True)


def converted_iao(gold, q_type, q_level):
    import micronet_b200 as E
    from micronet_b200 import bn_fuse
    m = E.iao.prepare(_float_model(gold), inplace=True, a_bits=8, w_bits=8, q_type=q_type, q_level=q_level,
                      weight_observer=0, bn_fuse=True, pretrained_model=True)
    m.load_state_dict({k[11:]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("calibrated.")})
    return bn_fuse.iao_model_bn_fuse(m)


@pytest.mark.parametrize("W", [2, 3])
def test_wbwtab_converter_matches_the_reference_script(W):
    gold = load_golden("bnfuse", f"wbwtab_W{W}")
    inf = converted_wbwtab(gold, W)
    assert _types(inf) == _gold_types(gold)
    _compare_state(inf, gold)
    assert all(m.quant_inference for m in inf.modules() if type(m).__name__ == "QuantConv2d")


@pytest.mark.parametrize("q", [(0, 0), (1, 1)], ids=["sym_per_channel", "asym_per_layer"])
def test_iao_converter_matches_the_reference_script(q):
    gold = load_golden("bnfuse", f"iao_t{q[0]}_l{q[1]}")
    inf = converted_iao(gold, *q)
    assert _types(inf) == _gold_types(gold)
    _compare_state(inf, gold)
    assert not any(type(m).__name__ == "QuantBNFuseConv2d" for m in inf.modules())
