"""Hardware self-tests of the sm_100a building blocks (tcgen05 descriptors, TMA geometry)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _err_flag():
    return torch.zeros(1, dtype=torch.int32, device=DEV)


@pytest.mark.parametrize("int8", [0, 1])
@pytest.mark.parametrize("N,K", [(16, 32), (32, 64), (64, 32), (128, 64), (48, 64)])
def test_umma_descriptor_selftest(N, K, int8):
    from micronet_b200 import _lib as L
    lib = L.load()
    g = torch.Generator().manual_seed(N * 1000 + K + int8)
    lo, hi = (-128, 128) if int8 else (-16, 17)
    A = torch.randint(lo, hi, (128, K), generator=g).float()
    B = torch.randint(lo, hi, (N, K), generator=g).float()
    want = A.double() @ B.double().t()
    Ad, Bd = A.to(DEV), B.to(DEV)
    D = torch.full((128, N), float("nan"), device=DEV)
    err = _err_flag()
    L.check(lib.mnb_selftest_umma(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, int8, err.data_ptr(),
                                  L.stream()), "selftest_umma")
    torch.cuda.synchronize()
    assert err.item() == 0, f"bounded wait timed out (code {err.item()})"
    assert torch.equal(D.cpu().double(), want), (D.cpu()[:2, :8], want[:2, :8])


@pytest.mark.parametrize("coord", [(0, 0, 0), (-2, -1, 0), (16, 8, 4), (3, 2, 5), (-1, 7, 1)])
def test_tma_box_and_oob_fill(coord):
    from micronet_b200 import _lib as L
    lib = L.load()
    Wd, Hd, Cd = 20, 10, 6
    box = (8, 4, 2)
    src = torch.arange(Wd * Hd * Cd, dtype=torch.float32).reshape(Cd, Hd, Wd) + 1.0
    pad = torch.zeros(Cd + 8, Hd + 16, Wd + 32)
    pad[:Cd, 8:8 + Hd, 16:16 + Wd] = src
    want = pad[coord[2]:coord[2] + box[2], 8 + coord[1]:8 + coord[1] + box[1], 16 + coord[0]:16 + coord[0] + box[0]]
    out = torch.full((box[2], box[1], box[0]), float("nan"), device=DEV)
    err = _err_flag()
    dims = (C.c_int64 * 3)(Wd, Hd, Cd)
    bx = (C.c_int32 * 3)(*box)
    cd = (C.c_int32 * 3)(*coord)
    sd = src.to(DEV)
    L.check(lib.mnb_selftest_tma3d(sd.data_ptr(), dims, bx, cd, out.data_ptr(), err.data_ptr(), L.stream()),
            "selftest_tma3d")
    torch.cuda.synchronize()
    assert err.item() == 0, f"bounded wait timed out (code {err.item()})"
    assert torch.equal(out.cpu(), want), (out.cpu(), want)
