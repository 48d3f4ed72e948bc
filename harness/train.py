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
    reference's own ``torch.optim.Adam`` with one param group per tensor.

    ``graph=True`` (engine only): after ``graph_warmup`` eager steps (first-call observer branches, lazy scratch
    allocations) the forward + loss + zero_grad + backward of a step is captured ONCE into a CUDA graph and replayed:
    a ResNet-18 BN-fuse step is ~1500 launches whose Python / ctypes issue time exceeds their GPU time.  The
    all-reduce and the Adam launch stay outside the graph.  Inputs are copied into static tensors."""

    def __init__(self, model, lr=0.01, wd=0.0, flat=False, graph=False, graph_warmup=3):
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
        self.graph_wanted = bool(graph and flat)
        self.graph_warmup = graph_warmup
        self.graph = None
        self.graph_error = None

    def _fwd_bwd(self, x, t):
        out = self.model(x)
        loss = self.crit(out, t)
        self.opt.zero_grad()
        loss.backward()
        return loss

    def _capture(self, x, t):
        try:
            self.static_x, self.static_t = x.clone(), t.clone()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.static_loss = self._fwd_bwd(self.static_x, self.static_t)
            self.graph = g
        except Exception as e:  # capture is an optimisation: fall back to eager launches
            self.graph, self.graph_wanted = None, False
            self.graph_error = f"{type(e).__name__}: {e}"[:300]
            torch.cuda.synchronize()

    def step(self, x, t):
        self.model.train()
        if self.graph is None and self.graph_wanted and self.steps >= self.graph_warmup:
            self._capture(x, t)
        if self.graph is not None:
            self.static_x.copy_(x, non_blocking=True)
            self.static_t.copy_(t, non_blocking=True)
            self.graph.replay()
            loss = self.static_loss
        else:
            loss = self._fwd_bwd(x, t)
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

    def __init__(self, model, graph=False):
        self.model = model
        self.graph_wanted, self.graph, self.graph_error, self.steps = bool(graph), None, None, 0

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
        if self.graph is None and self.graph_wanted and self.steps >= 2:
            try:
                self.static_x = x.clone()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.static_out = self.model(self.static_x)
                self.graph = g
            except Exception as e:
                self.graph, self.graph_wanted = None, False
                self.graph_error = f"{type(e).__name__}: {e}"[:300]
                torch.cuda.synchronize()
        self.steps += 1
        if self.graph is not None and x.shape == self.static_x.shape:
            self.static_x.copy_(x, non_blocking=True)
            self.graph.replay()
            return self.static_out
        return self.model(x)
