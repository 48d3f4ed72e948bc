"""CPU suite for the host side: the C-ABI library loads and exports every symbol the header
declares (no compute calls - there is no GPU here), prepare() rewrites model trees exactly like
the reference-pinned oracle, state_dict layouts round-trip with reference checkpoints, CPU tensors
are refused loudly, and the N>1 gradient all-reduce works over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from micronet_b200 import _lib as L, build
    build.build()
    header = open(os.path.join(ROOT, "include", "micronet_b200.h")).read()
    declared = set(re.findall(r"\b(mnb_[a-z0-9_]+)\s*\(", header))
    declared -= {"mnb_stream_t"}
    assert len(declared) >= 20
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/micronet_b200.h but not exported"
    assert declared == set(L.PROTOTYPES), declared ^ set(L.PROTOTYPES)
    assert L.load().mnb_version() >= 100


def test_bad_arguments_are_rejected_without_a_gpu():
    import ctypes as C
    from micronet_b200 import _lib as L
    lib = L.load()
    sh = L.ConvShape(1, 6, 8, 8, 4, 3, 3, 1, 1, 1, 1, 1, 1, 4)  # 6 channels not divisible by 4 groups
    ops = L.ConvOperands()
    assert lib.mnb_conv2d_fwd(C.byref(sh), C.byref(ops), None, None) == -1
    assert b"groups" in lib.mnb_last_error()
    qp = L.ActQParams(L.ACT_DOREFA, 1, 0, 0, 0, None, None, None, None)  # 1-bit DoReFa: unsupported (DF:39-41)
    assert lib.mnb_act_quant_fwd(None, 0, C.byref(qp), None, None, None, None) == -1


def test_cpu_tensors_fail_loudly():
    import micronet_b200 as E
    conv = E.dorefa.QuantConv2d(4, 4, 3, padding=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        conv(torch.randn(1, 4, 8, 8))
    with pytest.raises(AssertionError):
        E.dorefa.QuantConv2d(4, 4, 3, a_bits=1).activation_quantizer.spec()


def test_fused_entry_points_validate_before_launching():
    """argument checks and geometry cover of the producer / first-layer entry points run on the host"""
    import ctypes as C
    from micronet_b200 import _lib as L
    lib = L.load()
    fake = 4096  # never dereferenced: every call below must be refused before a launch
    assert lib.mnb_bn_sign_fwd(fake, 2, 4, 16, fake, fake, fake, fake, 3, fake, fake, None) == -1
    assert b"shuffle groups" in lib.mnb_last_error()
    assert lib.mnb_bn_sign_bwd(fake, fake, fake, 2, 4, 16, fake, fake, fake, 1, 1, fake, fake, fake, None, None, None) == -1
    assert lib.mnb_maxpool2d_fwd(fake, 1, 4, 8, 8, 16, 2, 0, 1, fake, fake, None) == -1      # kernel > 15
    assert lib.mnb_maxpool2d_bwd(fake, fake, 1, 4, 8, 8, 3, 2, 2, 1, fake, None) == -1       # pad > kernel / 2
    assert lib.mnb_bn_sign_pool_fwd(fake, 1, 4, 7, 8, fake, fake, fake, fake, 1, fake, fake, fake, None) == L.E_UNSUPPORTED
    assert lib.mnb_bn_sign_pool_fwd(fake, 1, 4, 8, 12, fake, fake, fake, fake, 1, fake, fake, fake, None) == L.E_UNSUPPORTED
    assert lib.mnb_bn_batch_stats(fake, 1, 4, 1, 1e-5, 0.1, fake, fake, None, fake, fake, None) == -1   # one value per channel
    first = L.ConvShape(8, 3, 32, 32, 256, 5, 5, 1, 1, 2, 2, 1, 1, 1)
    need = lib.mnb_fconv2d_wgrad_tc_scratch_bytes(C.byref(first))
    assert need == min(8 * 8, 148) * 256 * 75 * 4          # one partial dw per CTA, one CTA per 128-position tile
    for bad in (L.ConvShape(8, 3, 32, 32, 256, 5, 5, 1, 1, 2, 2, 1, 1, 3),      # groups
                L.ConvShape(8, 3, 32, 32, 256, 5, 5, 2, 2, 2, 2, 1, 1, 1),      # stride
                L.ConvShape(8, 8, 32, 32, 256, 5, 5, 1, 1, 2, 2, 1, 1, 1),      # C*R*S = 200 > 128
                L.ConvShape(8, 3, 48, 48, 64, 3, 3, 1, 1, 1, 1, 1, 1, 1),       # rows do not tile into 128 positions
                L.ConvShape(8, 3, 32, 32, 512, 3, 3, 1, 1, 1, 1, 1, 1, 1)):     # more than 256 output channels
        assert lib.mnb_fconv2d_wgrad_tc_scratch_bytes(C.byref(bad)) == -1
        assert lib.mnb_fconv2d_fwd_tc(C.byref(bad), fake, fake, None, fake, fake, None) == L.E_UNSUPPORTED


def test_fuse_pass_rewrites_the_prepared_graph_only():
    """wbwtab.prepare(fuse_bn=True): BN+binarizer pairs, pools, shuffles and the first float conv are rewritten on
    the engine's copy; state_dict layout and the user's model stay as they were"""
    import torch.nn as nn
    import micronet_b200 as E
    from harness import models as zoo
    from micronet_b200.fused import BatchNormBinarize2d, EngineFloatConv2d, EngineMaxPool2d
    torch.manual_seed(0)
    base = zoo.NINGC()
    plain = E.wbwtab.prepare(base, A=2, W=3)
    fused = E.wbwtab.prepare(base, A=2, W=3, fuse_bn=True)
    assert list(plain.state_dict().keys()) == list(fused.state_dict().keys())
    for (k, a), b in zip(plain.state_dict().items(), fused.state_dict().values()):
        assert torch.equal(a, b), k
    n_aq = sum(isinstance(m, E.wbwtab.ActivationQuantizer) for m in plain.modules())
    assert sum(isinstance(m, BatchNormBinarize2d) for m in fused.modules()) == n_aq
    assert not any(isinstance(m, E.wbwtab.ActivationQuantizer) for m in fused.modules())
    # both 2x2 pools follow a fused BN+binarizer and are absorbed by it
    assert not any(isinstance(m, nn.MaxPool2d) for m in fused.modules())
    assert sum(bool(getattr(m, "pool2", False)) for m in fused.modules()) == 2
    # every channel shuffle moved into its producer: flags cleared on the copy, groups recorded upstream
    assert not any(getattr(m, "channel_shuffle_flag", 0) for m in fused.modules())
    moved = sorted(m.out_shuffle_groups for m in fused.modules() if getattr(m, "out_shuffle_groups", 1) > 1)
    assert moved == sorted(m.shuffle_groups for m in plain.modules() if getattr(m, "channel_shuffle_flag", 0))
    assert any(getattr(m, "channel_shuffle_flag", 0) for m in base.modules()), "the user's model is left untouched"
    # the covered plain conv (3 -> 256, 5x5) goes to the fp32 tensor-core kernels; the 1024-channel classifier conv behind
    # the last binarizer to the packed-operand family (+-1 input), falling back to the stock conv without that tag
    from micronet_b200.fused import EnginePmConv2d
    assert [n for n, c in fused.named_modules() if isinstance(c, EngineFloatConv2d)] == ["model.0.conv"]
    assert type(dict(fused.named_modules())["model.10.conv"]) is EnginePmConv2d
    # producers whose only reader takes their bf16 plane skip the fp32 output: every un-pooled BatchNorm + binarizer here
    assert sum(bool(getattr(m, "plane_only", False)) for m in fused.modules()) == 6
    assert not any(getattr(m, "plane_only", False) and m.pool2 for m in fused.modules() if isinstance(m, BatchNormBinarize2d))
    # A != 2 keeps the ReLU path: nothing to fuse
    relu = E.wbwtab.prepare(base, A=32, W=2, fuse_bn=True)
    assert not any(isinstance(m, BatchNormBinarize2d) for m in relu.modules())
    # dorefa: pools (3x3 / stride 2 here: not absorbed) and the shuffles behind them
    nin = E.dorefa.prepare(zoo.NIN(), a_bits=8, w_bits=8, fuse=True)
    assert sum(isinstance(m, EngineMaxPool2d) for m in nin.modules()) == 2
    gc = E.dorefa.prepare(zoo.NINGC(), a_bits=4, w_bits=4, fuse=True)
    assert sorted(m.out_shuffle_groups for m in gc.modules() if isinstance(m, EngineMaxPool2d)) == [2, 4]
    with pytest.raises(RuntimeError, match="CUDA"):
        fused(torch.randn(2, 3, 32, 32))


def _types(model):
    return [(n, type(m).__name__) for n, m in model.named_modules()]


@pytest.mark.parametrize("scheme", ["wbwtab", "dorefa", "iao", "iao_bnfuse"])
def test_prepare_rewrites_like_the_oracle(scheme):
    import micronet_b200 as E
    from harness import models as zoo
    from oracle import reference_port as O
    rename = {"WbQuantConv2d": "QuantConv2d", "DorefaQuantConv2d": "QuantConv2d", "IaoQuantConv2d": "QuantConv2d",
              "IaoQuantBNFuseConv2d": "QuantBNFuseConv2d", "IaoQuantLinear": "QuantLinear",
              "DorefaQuantLinear": "QuantLinear", "WbActivationQuantizer": "ActivationQuantizer",
              "IaoQuantAdd": "QuantAdd"}
    quant_types = set(rename.values())
    if scheme == "wbwtab":
        e, o = E.wbwtab.prepare(zoo.NINGC(), A=2, W=3), O.prepare_wbwtab(zoo.NINGC(), A=2, W=3)
    elif scheme == "dorefa":
        e, o = E.dorefa.prepare(zoo.NIN(), a_bits=4, w_bits=4), O.prepare_dorefa(zoo.NIN(), a_bits=4, w_bits=4)
    elif scheme == "iao":
        e, o = E.iao.prepare(zoo.NINGC()), O.prepare_iao(zoo.NINGC(), add_type=zoo.Add)
    else:
        e = E.iao.prepare(zoo.resnet18(), bn_fuse=True)
        o = O.prepare_iao(zoo.resnet18(), bn_fuse=True, add_type=zoo.Add)
    et = {n: t for n, t in _types(e) if t in quant_types and not n.endswith("_quantizer")}
    ot = {n: rename[t] for n, t in _types(o) if t in rename}
    assert et == ot and len(et) > 0
    # identical parameter / buffer names and shapes (reference checkpoints must load)
    es, os_ = e.state_dict(), o.state_dict()
    assert list(es.keys()) == list(os_.keys())
    for k in es:
        assert es[k].shape == os_[k].shape and es[k].dtype == os_[k].dtype, k


def test_state_dict_matches_reference_checkpoint_layout():
    """keys/shapes recorded from the REFERENCE's own state_dict (golden fixture)."""
    import micronet_b200 as E
    from harness import models as zoo
    from tests.golden.cases import MODEL_CASES
    from tests.oracle_util import load_golden
    for case in MODEL_CASES:
        gold = load_golden("model", case["name"])
        m = {"nin_gc": zoo.NINGC, "nin": zoo.NIN}.get(case["model"])
        m = m(case["cfg"]) if m else zoo.ResNet(widths=tuple(case["cfg"]))
        mod = {"wbwtab": E.wbwtab, "dorefa": E.dorefa, "iao": E.iao}[case["scheme"]]
        m = mod.prepare(m, inplace=True, **case["prepare"])
        final = {k[6:]: v for k, v in gold.items() if k.startswith("final.")}
        sd = m.state_dict()
        assert set(sd) == set(final), set(sd) ^ set(final)
        for k, v in final.items():
            assert tuple(sd[k].shape) == tuple(v.shape), k
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in final.items()})


def test_quant_test_manual_surface():
    """constructor surface exercised by the reference's install check (micronet/__init__.py:6-123)."""
    import micronet_b200 as E
    E.wbwtab.QuantConv2d(1, 10, kernel_size=5, W=3)
    E.wbwtab.ActivationQuantizer(A=2)
    E.dorefa.QuantConv2d(1, 10, kernel_size=5, a_bits=8, w_bits=8)
    E.dorefa.QuantLinear(320, 50, a_bits=8, w_bits=8)
    E.iao.QuantConv2d(1, 10, kernel_size=5, a_bits=8, w_bits=8, q_type=1, q_level=1, weight_observer=1)
    E.iao.QuantBNFuseConv2d(10, 20, kernel_size=5, bn_fuse_calib=True, pretrained_model=True, qaft=False)
    E.iao.QuantLinear(320, 50, ptq=True, percentile=0.999)
    E.iao.QuantMaxPool2d(kernel_size=2, a_bits=8)
    E.iao.QuantReLU(inplace=True)
    E.iao.QuantAdd(a_bits=8, q_type=1)


WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from micronet_b200.parallel import FlatGradBucket
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
bucket = FlatGradBucket(net.parameters())
g = torch.Generator().manual_seed(1)
x = torch.randn(8, 5, generator=g); t = torch.randint(0, 3, (8,), generator=g)
shard = slice(rank * 8 // world, (rank + 1) * 8 // world)
bucket.zero()
torch.nn.functional.cross_entropy(net(x[shard]), t[shard]).backward()
bucket.all_reduce()
ref = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
ref.load_state_dict(net.state_dict())
torch.nn.functional.cross_entropy(ref(x), t).backward()
for p, q in zip(net.parameters(), ref.parameters()):
    assert torch.allclose(p.grad, q.grad, atol=1e-6), (p.grad, q.grad)
assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_flat_bucket_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
