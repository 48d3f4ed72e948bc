"""QAT step harness shared by bench.py, smoke() and the tests.

Re-creates the reference's training step (wbwtab/main.py:70-98: forward, CrossEntropy, zero_grad,
backward, Adam step; init wbwtab/main.py:309-317; one Adam param-group per tensor :331-339) on
synthetic CIFAR-10-shaped data - there is no dataset / network in this environment."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import models as zoo

# the BASELINE.json configs, as (model factory, scheme, prepare kwargs, weight decay)
WORKLOADS = {
    # configs[1]: the configuration the headline metric is quoted on
    "nin_gc_wbwtab_w3a2": dict(model="nin_gc", scheme="wbwtab", prepare=dict(W=3, A=2), wd=0.0, hw=32,
                               engine_extra=dict(fuse_bn=True)),
    "nin_dorefa_w8a8": dict(model="nin", scheme="dorefa", prepare=dict(a_bits=8, w_bits=8), wd=1e-5, hw=32,
                            engine_extra=dict(fuse=True)),
    "resnet18_iao_w8a8_bnfuse": dict(model="resnet18", scheme="iao",
                                     prepare=dict(a_bits=8, w_bits=8, q_type=0, q_level=0, weight_observer=0,
                                                  bn_fuse=True), wd=1e-5, hw=32),
    "nin_gc_dorefa_w4a4": dict(model="nin_gc", scheme="dorefa", prepare=dict(a_bits=4, w_bits=4), wd=1e-5, hw=32,
                               engine_extra=dict(fuse=True)),
    # configs[4]: IAO int8 PTQ inference, ResNet-18 at 224x224 (iao/main.py:109-142 calibrate, :511-519 eval):
    # prepare with ptq=True (HistogramObserver activations), 2 calibration batches in train mode / no_grad, then eval()
    "resnet18_iao_ptq_224": dict(model="resnet18", scheme="iao",
                                 prepare=dict(a_bits=8, w_bits=8, q_type=0, q_level=0, weight_observer=0, bn_fuse=True,
                                              pretrained_model=True, ptq=True, percentile=0.999999),
                                 wd=0.0, hw=224, inference=True, batch=64, calib_batches=2),
}


def build_float_model(name, seed=1):
    torch.manual_seed(seed)
    m = {"nin": zoo.NIN, "nin_gc": zoo.NINGC, "resnet18": zoo.resnet18}[name]()
    return zoo.init_like_reference(m)


def prepare_engine(model, scheme, **kw):
    """``prepare`` of the engine; workload entries may carry ``engine_extra`` kwargs (engine-only
    extensions such as BN+binarizer fusion) that the reference / oracle ``prepare`` does not know."""
    import micronet_b200 as E
    return {"wbwtab": E.wbwtab, "dorefa": E.dorefa, "iao": E.iao}[scheme].prepare(model, inplace=True, **kw)


def prepare_oracle(model, scheme, **kw):
    from oracle import reference_port as O  # CPU checker / baseline only
    if scheme == "wbwtab":
        return O.prepare_wbwtab(model, inplace=True, **kw)
    if scheme == "dorefa":
        return O.prepare_dorefa(model, inplace=True, **kw)
    return O.prepare_iao(model, inplace=True, add_type=zoo.Add, **kw)


def make_optimizer(model, lr=0.01, wd=0.0):
    groups = [{"params": [p], "lr": lr, "weight_decay": wd} for p in model.parameters()]
    return torch.optim.Adam(groups, lr=lr, weight_decay=wd)


def synthetic_batch(batch, hw, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, 3, hw, hw, generator=g)
    t = torch.randint(0, 10, (batch,), generator=g)
    if pin:
        x, t = x.pin_memory(), t.pin_memory()
    return x.to(device), t.to(device)


class QatStepper:
    """one QAT step = forward + CE loss + zero_grad + backward (+ grad all-reduce) + Adam.

    ``flat=True`` (engine on CUDA): parameters / gradients live in flat buckets, the gradient
    all-reduce is one NCCL call and Adam is one fused launch (micronet_b200.FlatAdam); otherwise the
    reference's own ``torch.optim.Adam`` with one param group per tensor."""

    def __init__(self, model, lr=0.01, wd=0.0, flat=False):
        self.model = model
        self.crit = nn.CrossEntropyLoss()
        if flat:
            from micronet_b200 import FlatAdam
            self.opt = FlatAdam(model.parameters(), lr=lr, weight_decay=wd)
        else:
            self.opt = make_optimizer(model, lr, wd)
        self.flat = flat
        self.steps = 0
        self.check_every = 50   # synchronising look at the tensor-core kernels' timeout flag (0: never)

    def step(self, x, t):
        self.model.train()
        out = self.model(x)
        loss = self.crit(out, t)
        self.opt.zero_grad()
        loss.backward()
        if self.flat:
            self.opt.all_reduce()
        self.opt.step()
        self.steps += 1
        if self.flat and self.check_every and self.steps % self.check_every == 0:
            from micronet_b200 import _lib as L
            L.tc_check()   # a bounded pipeline wait that gave up means garbage results: raise instead of training on
        return loss


class InferStepper:
    """PTQ inference (BASELINE.json configs[4]): ``calibrate`` = the reference's ptq_calibration loop
    (iao/main.py:109-142: train mode, no_grad, observers + scale update only), then ``eval()``; a step is one
    forward pass of a batch."""

    def __init__(self, model):
        self.model = model

    @torch.no_grad()
    def calibrate(self, batches):
        self.model.train()
        for x in batches:
            self.model(x)
        self.model.eval()
        if any(type(m).__module__.startswith("micronet_b200") for m in self.model.modules()):
            from micronet_b200 import iao
            iao.freeze_inference(self.model)
        return self

    @torch.no_grad()
    def step(self, x, t=None):
        return self.model(x)
