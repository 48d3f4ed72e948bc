"""Packed operands from the BatchNorm + binarizer producer (fused.BNSignFn -> mnb_bn_sign_fwd_packed) to the forward of
the next convolution on the packed-operand tensor-core family (mnb_pk.cu): the producer's bf16 plane must be exactly
the +-1 output in [B][C/8][H][W][8] order (producer's OUTPUT channel order), and the convolution fed by it must equal
an fp64 convolution of the fp32 tensor; the backward of such a layer must match the un-packed path."""
import os

import pytest
import torch
import torch.nn.functional as TF

from tests.oracle_util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"

# B, C, H, W, K, R, groups, shuffle groups of the producer
CASES = [(4, 256, 32, 32, 256, 1, 2, 1), (4, 256, 32, 32, 256, 1, 2, 2), (4, 512, 16, 16, 512, 1, 4, 16),
         (5, 1024, 8, 8, 1024, 1, 8, 32), (4, 256, 16, 16, 512, 3, 16, 2), (4, 512, 8, 8, 1024, 3, 32, 4),
         (3, 64, 16, 16, 32, 3, 1, 1)]


def _shuffle(x, groups):
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).contiguous().view(b, c, h, w)


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_packed_producer_and_conv_match_the_unpacked_path(case):
    from micronet_b200 import _lib as L, functional as F_
    from micronet_b200.fused import BatchNormBinarize2d
    B, C, H, W, K, R, G, sg = case
    torch.manual_seed(sum(case))
    bn = BatchNormBinarize2d(C).to(DEV).train()
    bn.out_shuffle_groups = sg
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C) + 0.5)
        bn.bias.copy_(torch.randn(C) * 0.3)
    x = (torch.randn(B, C, H, W) * 1.5).to(DEV)
    L.USE_PACKED = True       # opt-in path (MNB_PACKED_OPERANDS=1)
    try:
        y = bn(x)
    finally:
        L.USE_PACKED = False
    packed = getattr(y, "_mnb_packed", None)
    assert packed is not None, "producer did not emit the packed operand"
    # packed tensor = the +-1 plane in [B][C/8][H][W][8] order
    want = y.detach().view(B, C // 8, 8, H, W).permute(0, 1, 3, 4, 2).contiguous().to(torch.bfloat16).flatten()
    assert torch.equal(packed, want)
    w_int = torch.randint(-1, 2, (K, C // G, R, R), dtype=torch.int16).to(DEV)
    w_scale = (torch.rand(K) * 0.02 + 0.001).to(DEV)
    bias = torch.randn(K).to(DEV)
    wq = w_int.float() * w_scale.view(-1, 1, 1, 1)
    out = F_.quant_conv2d(y, wq, bias, w_int, w_scale, None, (1, 1), (R // 2, R // 2), (1, 1), G)
    L.tc_check()
    ref = TF.conv2d(y.detach().double().cpu(), wq.double().cpu(), bias.double().cpu(), 1, R // 2, 1, G)
    assert rel_err(out.detach(), ref) < 2e-6


@pytest.mark.parametrize("case", CASES[:5], ids=[str(c) for c in CASES[:5]])
def test_block_with_packed_forward_has_the_same_gradients(case):
    """BatchNorm+binarizer -> conv three ways: (base) everything on the fused kernels, (fwd) forward through the producer's
    packed plane and backward on the fused kernels (MNB_PACKED_OPERANDS=1), (both) the default since round 2 - forward AND
    backward on the packed-operand family with producer-written operands (MNB_PK_WBWTAB=1).  Same +-1 operand in the
    forward: identical outputs; gradients agree to the accuracy of the backward's operand split."""
    from micronet_b200 import _lib as L, functional as F_
    from micronet_b200.fused import BatchNormBinarize2d
    B, C, H, W, K, R, G, sg = case
    torch.manual_seed(sum(case) + 1)
    x0 = (torch.randn(B, C, H, W) * 1.5).to(DEV)
    w_int = torch.randint(-1, 2, (K, C // G, R, R), dtype=torch.int16).to(DEV)
    w_scale = (torch.rand(K) * 0.02 + 0.001).to(DEV)
    go = torch.randn(B, K, H, W).to(DEV)
    res = {}
    saved = (L.USE_PACKED, L.PK_WBWTAB)
    for name, packed, wb in (("base", False, False), ("fwd", True, False), ("both", False, True)):
        L.USE_PACKED, L.PK_WBWTAB = packed, wb
        try:
            torch.manual_seed(5)
            bn = BatchNormBinarize2d(C).to(DEV).train()
            bn.out_shuffle_groups = sg
            x = x0.clone().requires_grad_(True)
            wq = (w_int.float() * w_scale.view(-1, 1, 1, 1)).requires_grad_(True)
            F_.TIMER = F_.KernelTimer()
            out = F_.quant_conv2d(bn(x), wq, None, w_int, w_scale, None, (1, 1), (R // 2, R // 2), (1, 1), G)
            out.backward(go)
            torch.cuda.synchronize()
            kinds = {k for k, _, _, _ in F_.TIMER.records}
            res[name] = (out.detach(), x.grad, wq.grad, bn.weight.grad, kinds)
        finally:
            L.USE_PACKED, L.PK_WBWTAB = saved
            F_.TIMER = None
    if R == 1:   # (3x3 layers take the packed-operand family in every configuration: they win there even when they pack themselves)
        assert "fwd_pk" not in res["base"][4], res["base"][4]
    assert "fwd_pk" in res["fwd"][4] and {"fwd_pk", "dgrad_pk", "wgrad_pk"} <= res["both"][4], (res["fwd"][4], res["both"][4])
    assert torch.equal(res["fwd"][0], res["base"][0]) or rel_err(res["fwd"][0], res["base"][0]) <= 1e-6
    assert rel_err(res["both"][0], res["base"][0]) <= 1e-6
    for a, b in zip(res["fwd"][1:4], res["base"][1:4]):
        assert rel_err(a, b) <= 1e-6
    for a, b in zip(res["both"][1:4], res["base"][1:4]):
        assert rel_err(a, b) <= 1e-5     # two-piece dy operand (MNB_PK_TERMS_BWD=2): ~3e-6, inside the 1e-5 contract
    L.tc_check()
