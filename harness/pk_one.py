"""Run each packed-operand kernel of ONE layer a few times (ncu target):  python -m harness.pk_one conv2_x [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from harness.pk_probe import LAYERS
from micronet_b200 import _lib as L, pk as PK, functional as F_

key = sys.argv[1] if len(sys.argv) > 1 else "conv2_x"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(os.environ.get("PK_BATCH", "256"))
TD = int(os.environ.get("PK_DY_TERMS", "3"))   # pieces of the gradient operand (the QAT backward runs 2)
dev = torch.device("cuda:0")
for name, Cc, H, W, K, R, st, pad, G in LAYERS:
    if key not in name:
        continue
    sh = L.ConvShape(B, Cc, H, W, K, R, R, st, st, pad, pad, 1, 1, G)
    P, Q = (H + 2 * pad - R) // st + 1, (W + 2 * pad - R) // st + 1
    x = torch.randn(B, Cc, H, W, device=dev); dy = torch.randn(B, K, P, Q, device=dev)
    w_int = torch.randint(-127, 128, (K, Cc // G, R, R), dtype=torch.int16, device=dev)
    w_f = torch.randn(K, Cc // G, R, R, device=dev) * 0.05
    w_scale = torch.rand(K, device=dev) * 0.01 + 0.001
    y = torch.empty(B, K, P, Q, device=dev); dx = torch.empty_like(x); dw = torch.empty(K, Cc // G, R, R, device=dev)
    qp = F_.ActSpec(L.ACT_DOREFA, bits=8).struct()
    xq = PK.pack_act(x, qp, 1, phase_split=st == 2)[0]; xr = PK.pack_act(x, None, 3, phase_split=st == 2)[0]
    dyp = PK.pack_act(dy, None, TD)[0]
    i11 = PK.pack_weight(sh, 0, 1, 1, w_int=w_int); i33 = PK.pack_weight(sh, 0, 3, 3, w_f32=w_f)
    idg = PK.pack_weight(sh, 1, TD, 1, w_int=w_int)
    for _ in range(reps):
        L.check(PK.conv(sh, 0, xq, 1, i11, 1, y, n_scale=w_scale), "c")
        L.check(PK.conv(sh, 0, xr, 3, i33, 3, y), "c")
        L.check(PK.conv(sh, 1, dyp, TD, idg, 1, dx), "c")
        L.check(PK.wgrad(sh, dyp, TD, xq, 1, dw), "w")
    torch.cuda.synchronize(); L.tc_check(); print("ran", name)
