"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle (GPU) tests."""
import os

import numpy as np
import torch

from oracle import reference_port as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ORACLE_CLASSES = {
    ("dorefa", "conv"): O.DorefaQuantConv2d,
    ("dorefa", "linear"): O.DorefaQuantLinear,
    ("wbwtab", "conv"): O.WbQuantConv2d,
    ("iao", "conv"): O.IaoQuantConv2d,
    ("iao", "bnfuse"): O.IaoQuantBNFuseConv2d,
    ("iao", "linear"): O.IaoQuantLinear,
    ("iao", "convT"): O.IaoQuantConvTranspose2d,
}


def load_golden(prefix, name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{prefix}_{name}.npz")))


def build_from_golden(cls, case, gold, device="cpu"):
    mod = cls(*case["args"], **case["kwargs"])
    state = {k[len("init."):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("init.")}
    mod.load_state_dict(state, strict=True)
    return mod.to(device)


def rel_err(a, b):
    """max |a-b| / max(|b|) — the per-tensor criterion of SURVEY.md §8(c)."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    den = b.abs().max().item()
    if den == 0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / den


def run_layer_steps(mod, case, gold, device="cpu"):
    """Replay the golden case through ``mod``; yields per-step dicts of results."""
    steps = case["train_steps"] + case["eval_steps"]
    for i in range(steps):
        training = i < case["train_steps"]
        mod.train(training)
        x = torch.from_numpy(gold[f"s{i}.x"]).to(device).requires_grad_(True)
        go = torch.from_numpy(gold[f"s{i}.go"]).to(device)
        y = mod(x)
        res = {"y": y.detach()}
        if training:
            mod.zero_grad()
            y.backward(go)
            res["dx"] = x.grad.detach()
            for n, p in mod.named_parameters():
                res[f"d.{n}"] = p.grad.detach()
        for n, t in mod.state_dict().items():
            res[f"state.{n}"] = t.detach().clone()
        yield i, res
