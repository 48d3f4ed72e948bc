"""Producer-side fusion (first step): BatchNorm2d + binarizing ActivationQuantizer.

A wbwtab-prepared block is ``conv -> bn -> ActivationQuantizer(A=2)`` (nin_gc.py:53-59 with the ReLU
swapped by WB:319-322).  ``fuse_bn_binarize`` rewrites every such sibling pair into one
``BatchNormBinarize2d`` (a ``nn.BatchNorm2d`` subclass: same parameters / buffers / state_dict keys) followed
by ``nn.Identity``; semantics are unchanged: training-mode batch statistics, running-stat updates with the
module's momentum, sign() with 0 -> +1 and the saturate STE |bn| < 1 in the backward pass."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L


def _channel_stats3(x):
    lib = L.load()
    b, c = x.shape[0], x.shape[1]
    hw = x.numel() // (b * c)
    out = torch.empty(3 * c, dtype=torch.float32, device=x.device)
    L.check(lib.mnb_channel_stats(x.data_ptr(), b, c, hw, 2, out.data_ptr(), L.scratch(x.device, c).data_ptr(),
                                  L.stream()), "channel_stats")
    return out[:c], out[c:2 * c], out[2 * c:]


class BNSignFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, mean, invstd, training):
        lib = L.load()
        x = x.contiguous()
        b, c = x.shape[0], x.shape[1]
        hw = x.numel() // (b * c)
        y = torch.empty_like(x)
        bits = torch.empty((x.numel() + 31) // 32, dtype=torch.int32, device=x.device)
        L.check(lib.mnb_bn_sign_fwd(x.data_ptr(), b, c, hw, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                    beta.data_ptr(), y.data_ptr(), bits.data_ptr(), L.stream()), "bn_sign_fwd")
        ctx.save_for_backward(x, gamma, mean, invstd)
        ctx.bits, ctx.training = bits, training
        return y

    @staticmethod
    def backward(ctx, g):
        lib = L.load()
        x, gamma, mean, invstd = ctx.saved_tensors
        g = g.contiguous()
        b, c = x.shape[0], x.shape[1]
        hw = x.numel() // (b * c)
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        L.check(lib.mnb_bn_sign_bwd(g.data_ptr(), ctx.bits.data_ptr(), x.data_ptr(), b, c, hw, mean.data_ptr(),
                                    invstd.data_ptr(), gamma.data_ptr(), 1 if ctx.training else 0, dx.data_ptr(),
                                    dgamma.data_ptr(), dbeta.data_ptr(), L.scratch(x.device, c).data_ptr(),
                                    L.stream()), "bn_sign_bwd")
        return dx, dgamma, dbeta, None, None, None


class BatchNormBinarize2d(nn.BatchNorm2d):
    """nn.BatchNorm2d followed by wbwtab's binarizing ActivationQuantizer, as two kernels forward
    (statistics, normalise+sign) and two backward (reductions, apply)."""

    def forward(self, input):
        L.require_cuda(input, self.weight)
        assert self.affine and self.track_running_stats and self.momentum is not None, \
            "BatchNormBinarize2d supports affine BN with running statistics and a float momentum"
        use_batch = self.training
        if use_batch:
            mean, var_b, var_u = _channel_stats3(input.detach().contiguous())
            with torch.no_grad():
                m = self.momentum
                self.running_mean.mul_(1 - m).add_(mean, alpha=m)
                self.running_var.mul_(1 - m).add_(var_u, alpha=m)
                self.num_batches_tracked.add_(1)
            invstd = torch.rsqrt(var_b + self.eps)
        else:
            mean = self.running_mean
            invstd = torch.rsqrt(self.running_var + self.eps)
        return BNSignFn.apply(input, self.weight, self.bias, mean.contiguous(), invstd.contiguous(), use_batch)


def fuse_bn_binarize(module: nn.Module) -> nn.Module:
    """in place: every (BatchNorm2d, wbwtab.ActivationQuantizer(A=2)) sibling pair -> (BatchNormBinarize2d, Identity)"""
    from .wbwtab import ActivationQuantizer
    prev_name, prev = None, None
    for name, child in list(module.named_children()):
        if (isinstance(child, ActivationQuantizer) and child.A == 2 and type(prev) is nn.BatchNorm2d
                and prev.affine and prev.track_running_stats and prev.momentum is not None):
            fused = BatchNormBinarize2d(prev.num_features, eps=prev.eps, momentum=prev.momentum)
            fused.weight, fused.bias = prev.weight, prev.bias
            fused.running_mean, fused.running_var = prev.running_mean, prev.running_var
            fused.num_batches_tracked = prev.num_batches_tracked
            fused.train(prev.training)
            module._modules[prev_name] = fused
            module._modules[name] = nn.Identity()
        else:
            fuse_bn_binarize(child)
        prev_name, prev = name, module._modules[name]
    return module
