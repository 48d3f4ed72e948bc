"""GPU timing probe of the tensor-core fwd / dgrad kernel with parts switched off (MNB_TC_DEBUG): debug only"""
import os
import subprocess
import sys

import torch

SHAPES = [(256, 256, 32, 32, 256, 1, 2), (256, 512, 16, 16, 512, 1, 4), (256, 256, 16, 16, 512, 3, 16)]


def run():
    from micronet_b200 import functional as F_
    dev = torch.device("cuda")
    out = []
    for (B, C, H, W, K, R, G) in SHAPES:
        g = torch.Generator().manual_seed(1)
        x = (torch.randint(0, 2, (B, C, H, W), generator=g).float() * 2 - 1).to(dev).requires_grad_(True)
        w_int = torch.randint(-1, 2, (K, C // G, R, R), generator=g, dtype=torch.int16).to(dev)
        w_scale = (torch.rand(K, generator=g) * 0.02 + 0.001).to(dev)
        wq = (w_int.float() * w_scale.view(-1, 1, 1, 1))
        go = torch.randn(B, K, H, W, generator=g).to(dev)
        y = F_.quant_conv2d(x, wq, None, w_int, w_scale, None, (1, 1), (R // 2, R // 2), (1, 1), G)

        def fwd():
            return F_.quant_conv2d(x, wq, None, w_int, w_scale, None, (1, 1), (R // 2, R // 2), (1, 1), G)

        def bwd():
            torch.autograd.grad(y, x, go, retain_graph=True)
        for name, fn in (("fwd", fwd), ("dgrad", bwd)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out.append(f"{name}{R}x{R}g{G}:{e0.elapsed_time(e1) * 100:.0f}")
    print(f"dbg={os.environ.get('MNB_TC_DEBUG', '0'):>2s}", " ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for mask in (0, 1, 2, 4, 8, 6, 14, 15):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MNB_TC_DEBUG=str(mask)), timeout=120)
