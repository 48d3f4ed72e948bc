"""debug: per-parameter gradient differences between runs / between fused and unfused DoReFa NIN-GC models"""
import torch
import torch.nn as nn

import micronet_b200 as E
from harness import models as zoo

DEV = "cuda"


def run(kw, seed=2):
    torch.manual_seed(seed)
    base = zoo.NINGC()
    zoo.init_like_reference(base)
    x = torch.randn(8, 3, 32, 32).to(DEV)
    t = torch.randint(0, 10, (8,)).to(DEV)
    m = E.dorefa.prepare(base, a_bits=4, w_bits=4, **kw).to(DEV).train()
    acts = {}
    for n, mod in m.named_modules():
        if n.count(".") == 1 and n.startswith("model."):
            mod.register_full_backward_hook(lambda mod, gi, go, n=n: acts.__setitem__(n, go[0].detach().clone()))
    loss = nn.functional.cross_entropy(m(x), t)
    loss.backward()
    return loss.item(), {n: p.grad.clone() for n, p in m.named_parameters()}, acts


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def shuffle(x, groups):
    b, c, h, w = x.shape
    return x.view(b, groups, c // groups, h, w).transpose(1, 2).contiguous().view(b, c, h, w)


a = run({})
b = run({})
c = run({"fuse": True})
print("loss", a[0], b[0], c[0])
for n in a[1]:
    print(f"{n:28s} plain-vs-plain {rel(a[1][n], b[1][n]):.2e}   fused-vs-plain {rel(c[1][n], a[1][n]):.2e}")
for n in sorted(a[2], key=lambda s: int(s.split('.')[1])):
    ga, gc = a[2][n], c[2][n]
    r = rel(gc, ga) if ga.shape == gc.shape else float("nan")
    extra = ""
    if n in ("model.3", "model.7"):
        g = 2 if n == "model.3" else 4
        extra = f"  (fused grad_out un-shuffled vs plain: {rel(gc, shuffle(ga, g)):.2e})"
    print(f"grad_out {n:10s} run-to-run {rel(ga, b[2][n]):.2e}  fused-vs-plain {r:.2e}{extra}")
