"""GPU timing probe for the fp32 first-layer tensor-core conv (debug): python harness/fconv_probe.py"""
import ctypes as C
import os
import subprocess
import sys

import torch


def run():
    from micronet_b200 import _lib as L
    lib = L.load()
    B, Cc, H, W, K, R = 256, 3, 32, 32, 256, 5
    dev = torch.device("cuda")
    x = torch.randn(B, Cc, H, W, device=dev)
    w = torch.randn(K, Cc, R, R, device=dev) * 0.1
    b = torch.randn(K, device=dev)
    y = torch.empty(B, K, H, W, device=dev)
    dy = torch.randn(B, K, H, W, device=dev)
    dw = torch.empty_like(w)
    sh = L.ConvShape(B, Cc, H, W, K, R, R, 1, 1, R // 2, R // 2, 1, 1, 1)
    scratch = torch.empty(int(lib.mnb_fconv2d_wgrad_tc_scratch_bytes(C.byref(sh))), dtype=torch.uint8, device=dev)
    err = L.tc_err_flag(dev)

    def fwd():
        L.check(lib.mnb_fconv2d_fwd_tc(C.byref(sh), x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), err.data_ptr(),
                                       L.stream()), "fwd")

    def wg():
        L.check(lib.mnb_fconv2d_wgrad_tc(C.byref(sh), dy.data_ptr(), x.data_ptr(), dw.data_ptr(), scratch.data_ptr(),
                                         err.data_ptr(), L.stream()), "wgrad")
    out = []
    for name, fn in (("fwd", fwd), ("wgrad", wg)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) * 100:.1f}us")
    print(os.environ.get("MNB_FCONV_DEBUG", "0"), " ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for mask in (0, 1, 2, 4, 8, 3, 7, 9, 15):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MNB_FCONV_DEBUG=str(mask)), timeout=120)
