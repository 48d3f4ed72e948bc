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
        lib = L.load()
        iao.freeze_inference(eng, handoff=False)
        eng(x.to(DEV))
        n0 = lib.mnb_launch_count()
        nohand = eng(x.to(DEV)).clone()
        launches_nohand = lib.mnb_launch_count() - n0
        iao.freeze_inference(eng)                  # producers write their consumer's operand plane
        n_identity = sum(isinstance(m, torch.nn.Identity) for m in eng.modules())
        frozen1 = eng(x.to(DEV)).clone()
        n0 = lib.mnb_launch_count()
        frozen2 = eng(x.to(DEV)).clone()          # second call: cached weights / images
        launches_hand = lib.mnb_launch_count() - n0
        iao.freeze_inference(eng, enable=False)
        back = eng(x.to(DEV)).clone()
    assert torch.equal(plain, nohand)
    assert torch.equal(plain, frozen1) and torch.equal(plain, frozen2) and torch.equal(plain, back)
    assert n_identity > sum(isinstance(m, torch.nn.Identity) for m in eng.modules())   # the ReLUs came back
    # 8 residual blocks: conv1 -> conv2 hand-offs (8 pack launches gone) and 7 QuantAdd -> next conv1 hand-offs
    assert launches_hand <= launches_nohand - 15, (launches_hand, launches_nohand)
    L.tc_check()


def test_conv_epilogue_writes_the_consumers_plane():
    """mnb_pk_conv_post / mnb_quant_add_pack_fwd against the separate kernels: same fp32 result, same operand plane"""
    import ctypes as C
    from micronet_b200 import _lib as L, functional as F_, pk as PK
    torch.manual_seed(5)
    B, Cc, H, W, K = 4, 64, 16, 16, 128
    x = torch.randn(B, Cc, H, W, device=DEV) * 3
    w_int = torch.randint(-127, 128, (K, Cc, 3, 3), dtype=torch.int16, device=DEV)
    w_scale = torch.rand(K, device=DEV) * 0.01 + 0.001
    bias = torch.randn(K, device=DEV)

    def iao_spec(scale):
        bufs = dict(scale=torch.tensor([scale]), zero_point=torch.zeros(1), obs_min=torch.tensor([-127.5 * scale]),
                    obs_max=torch.tensor([127.5 * scale]))
        return F_.ActSpec(L.ACT_IAO, qmin=-128, qmax=127, q_type=0, **{k: v.to(DEV) for k, v in bufs.items()})

    spec, nxt = iao_spec(0.05), iao_spec(0.11)
    sh = L.ConvShape(B, Cc, H, W, K, 3, 3, 1, 1, 1, 1, 1, 1, 1)
    x_pk, _ = PK.pack_act(x, spec.struct(), 1)
    w_img = PK.pack_weight(sh, 0, 1, 1, w_int=w_int)
    y_ref = torch.empty(B, K, H, W, device=DEV)
    L.check(PK.conv(sh, 0, x_pk, 1, w_img, 1, y_ref, n_scale=w_scale, a_scale=spec.scale, bias=bias), "conv")
    for relu in (False, True):
        for split in (False, True):
            want, _ = PK.pack_act(y_ref, nxt.struct(), 1, phase_split=split, relu=relu)
            for with_out in (True, False):
                y = torch.empty_like(y_ref) if with_out else None
                plane = PK.consumer_plane(B, K, H, W, DEV)
                L.check(PK.conv_post(sh, x_pk, 1, w_img, 1, y, nxt.struct(), plane, relu, split, n_scale=w_scale,
                                     a_scale=spec.scale, bias=bias), "conv_post")
                assert torch.equal(plane, want), (relu, split, with_out)
                if with_out:
                    assert torch.equal(y, y_ref)
    # QuantAdd + pack
    a, b = torch.randn(B, K, H, W, device=DEV) * 4, torch.randn(B, K, H, W, device=DEV) * 4
    add = iao_spec(0.07)
    lib = L.load()
    for relu in (False, True):
        ref = F_.QuantAddFn.apply(a, b, add, relu)
        for split in (False, True):
            want, _ = PK.pack_act(ref, nxt.struct(), 1, phase_split=split)
            out = torch.empty_like(a)
            plane = PK.consumer_plane(B, K, H, W, DEV)
            qp, cqp = add.struct(), nxt.struct()
            post = L.PkPost(C.pointer(cqp), 0, 1 if split else 0, plane.data_ptr())
            L.check(lib.mnb_quant_add_pack_fwd(a.data_ptr(), b.data_ptr(), B, K, H, W, C.byref(qp), 1 if relu else 0,
                                               out.data_ptr(), C.byref(post), L.stream()), "quant_add_pack")
            assert torch.equal(out, ref) and torch.equal(plane, want), (relu, split)
    L.tc_check()


def test_eval_forward_matches_the_oracle_at_224():
    """full-width ResNet-18 at 2 x 3 x 224 x 224.  A PTQ calibration on two images is chaotic (the percentile range of a
    layer is set by single extreme elements, one activation level on the other side of a rounding tie moves the next
    layer's range by 1e-3 ...), CPU <-> CPU as much as CPU <-> GPU, so the comparison is made in two teacher-forced steps:

    * calibration: the buffers of the layers in front of the first level flip (stem + first block) agree to 1e-5 after both
      sides calibrated on their own; every calibrated scale agrees to 5 %;
    * evaluation: the engine is given the ORACLE's calibrated state_dict; every quantized module, fed the oracle's input of
      that module, reproduces the oracle's output to 1e-5 except at the few elements an input level on the other side of
      a rounding tie reaches (at most 2e-3 of the elements, each by at most 2 % of the tensor's range), and the logits of
      the whole frozen model stay within 1e-2."""
    from micronet_b200 import iao, _lib as L
    from tests.test_gpu_parity import QUANT_TYPES
    eng, ora, calib, x = _prepared((64, 128, 256, 512), 224)
    with torch.no_grad():
        eng.train(); eng(calib.to(DEV)); eng.eval()
        ora.train(); ora(calib); ora.eval()
        so, se = ora.state_dict(), eng.state_dict()
        assert so.keys() == se.keys()
        for k in so:
            if not so[k].dtype.is_floating_point:
                continue
            if k.startswith(("conv1.", "conv2_x.0.residual_function.0.")):
                assert rel_err(se[k], so[k]) <= 1e-5, (k, rel_err(se[k], so[k]))
            elif k.endswith("quantizer.scale"):
                assert rel_err(se[k], so[k]) <= 5e-2, (k, rel_err(se[k], so[k]))
        eng.load_state_dict(so)
        names = [n for n, m in eng.named_modules() if type(m).__name__ in QUANT_TYPES and not n.endswith("activation_quantizer")]
        cap = {}
        om = dict(ora.named_modules())
        hooks = [om[n].register_forward_hook(lambda mod, inp, out, n=n: cap.__setitem__(n, ([t.detach().clone() for t in inp], out.detach().clone())))
                 for n in names]
        yo = ora(x)
        for h in hooks:
            h.remove()
        em = dict(eng.named_modules())
        bad = []
        for n in names:
            xin, yref = cap[n]
            ye = em[n](*[t.to(DEV) for t in xin]).cpu()
            d = (ye - yref).abs() / yref.abs().max()
            frac, worst = (d > 1e-5).float().mean().item(), d.max().item()
            if frac > 2e-3 or worst > 2e-2:
                bad.append(f"{n}: {frac:.2e} of the outputs differ, worst {worst:.2e}")
        assert not bad, "\n".join(bad)
        iao.freeze_inference(eng)
        ye = eng(x.to(DEV)).cpu()
    assert rel_err(ye, yo) <= 1e-2, rel_err(ye, yo)
    L.tc_check()
