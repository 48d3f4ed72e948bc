"""Converted (BN-fused, quant_inference=True) models of micronet_b200.bn_fuse run on the CUDA engine against the eval
output of the reference's own converted model (fixtures of tests/golden/make_golden_bnfuse.py)."""
import pytest
import torch

from tests.oracle_util import load_golden, rel_err
from tests.test_bn_fuse_cpu import converted_iao, converted_wbwtab

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("W", [2, 3])
def test_wbwtab_converted_model_inference(W):
    gold = load_golden("bnfuse", f"wbwtab_W{W}")
    inf = converted_wbwtab(gold, W).to(DEV).eval()
    with torch.no_grad():
        y = inf(torch.from_numpy(gold["x"]).to(DEV))
    assert rel_err(y, gold["y"]) <= 1e-5, rel_err(y, gold["y"])


@pytest.mark.parametrize("q", [(0, 0), (1, 1)], ids=["sym_per_channel", "asym_per_layer"])
def test_iao_converted_model_inference(q):
    from micronet_b200 import _lib as L
    gold = load_golden("bnfuse", f"iao_t{q[0]}_l{q[1]}")
    inf = converted_iao(gold, *q).to(DEV).eval()
    with torch.no_grad():
        y = inf(torch.from_numpy(gold["x"]).to(DEV))
    # 8-bit activation levels of 11 layers: one level on the other side of a rounding tie moves a logit by ~1e-4 relative
    assert rel_err(y, gold["y"]) <= 2e-4, rel_err(y, gold["y"])
    L.tc_check()
