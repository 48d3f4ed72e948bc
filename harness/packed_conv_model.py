"""CPU model of the index arithmetic of the experimental packed-operand forward conv (mnb_conv_packed.cu).

It replays, in numpy, exactly what the kernel asks the hardware to do - the 5-D TMA box into shared memory, the
UMMA K-major no-swizzle addressing of A and B from the descriptor start addresses the MMA warp computes, the
accumulator-row -> output-pixel mapping of the epilogue - and compares the result with a plain convolution.  It checks
the arithmetic, not the hardware semantics (those are pinned by mnb_selftest_umma / mnb_selftest_tma3d).

    python harness/packed_conv_model.py"""
import numpy as np


def plan(B, Cin, Cout, H, W, R, G):
    pad = R // 2
    cin_g, cout_g = Cin // G, Cout // G
    BW = W + 2 * pad
    TH = min(H, 128 // BW)
    THH = TH + 2 * pad
    TB = 1
    if pad == 0 and TH == H:
        TB = max(1, min(B, 128 // (H * W)))
    npos = TB * THH * BW
    halo = (R - 1) * BW + (R - 1)
    sg = next(s for s in range(G, 0, -1) if G % s == 0 and s * cout_g <= 256)
    return dict(pad=pad, cin_g=cin_g, cout_g=cout_g, BW=BW, TH=TH, THH=THH, TB=TB, npos=npos, halo=halo, slab=sg,
                row_tiles=(H + TH - 1) // TH)


def run(B, Cin, Cout, H, W, R, G, seed=0):
    rng = np.random.default_rng(seed)
    p = plan(B, Cin, Cout, H, W, R, G)
    x = rng.choice([-1.0, 1.0], size=(B, Cin, H, W))
    w = rng.integers(-1, 2, size=(Cout, Cin // G, R, R)).astype(np.float64)
    # global packed tensor xp[b][c8][h][w][8]
    xp = x.reshape(B, Cin // 8, 8, H, W).transpose(0, 1, 3, 4, 2).copy()
    # packed weights [g][tap][c8][n][e]
    cin_g, cout_g, RS = p["cin_g"], p["cout_g"], R * R
    wp = np.zeros((G, RS, cin_g // 8, cout_g, 8))
    for g in range(G):
        for tap in range(RS):
            for c in range(cin_g):
                wp[g, tap, c // 8, :, c % 8] = w[g * cout_g:(g + 1) * cout_g, c, tap // R, tap % R]
    y = np.zeros((B, Cout, H, W))
    c8g = cin_g // 8
    n_slabs = G // p["slab"]
    for slab in range(n_slabs):
        g_first = slab * p["slab"]
        slab_c8 = p["slab"] * c8g
        b_flat = wp[g_first:g_first + p["slab"]].reshape(-1)        # resident B image, 1 unit = one bf16
        for bt in range((B + p["TB"] - 1) // p["TB"]):
            for rt in range(p["row_tiles"]):
                # ---- TMA box (8, BW, THH, TB, slab_c8) at coords (0, -pad, h0 - pad, b0, g_first * c8g), zero fill
                box = np.zeros((slab_c8, p["TB"], p["THH"], p["BW"], 8))
                for c8 in range(slab_c8):
                    for tb in range(p["TB"]):
                        for row in range(p["THH"]):
                            for col in range(p["BW"]):
                                b, h, wc = bt * p["TB"] + tb, rt * p["TH"] - p["pad"] + row, col - p["pad"]
                                if b < B and 0 <= h < H and 0 <= wc < W:
                                    box[c8, tb, row, col] = xp[b, g_first * c8g + c8, h, wc]
                a_flat = np.concatenate([box.reshape(-1), np.zeros((p["halo"] + 128) * 8)])   # + slack rows
                # ---- MMAs: units of 16 bytes = 8 bf16; A rows are positions, LBO = npos, SBO = 8 rows
                acc = np.zeros((128, p["slab"] * cout_g))
                for gi in range(p["slab"]):
                    a_g = gi * c8g * p["npos"]
                    b_g = gi * (RS * c8g * cout_g)
                    for r in range(R):
                        for s2 in range(R):
                            for j in range(cin_g // 16):
                                a_start = a_g + r * p["BW"] + s2 + j * 2 * p["npos"]
                                b_start = b_g + (r * R + s2) * (c8g * cout_g) + j * 2 * cout_g
                                for kc in range(2):                                   # two 8-wide K chunks per K16 step
                                    A = a_flat.reshape(-1, 8)[a_start + kc * p["npos"]: a_start + kc * p["npos"] + 128]
                                    Bm = b_flat.reshape(-1, 8)[b_start + kc * cout_g: b_start + kc * cout_g + cout_g]
                                    acc[:, gi * cout_g:(gi + 1) * cout_g] += A @ Bm.T
                # ---- epilogue: accumulator row = position of the zero-padded tile
                for pos in range(128):
                    tb = pos // (p["THH"] * p["BW"])
                    rem = pos - tb * p["THH"] * p["BW"]
                    th, wc = rem // p["BW"], rem % p["BW"]
                    b, h = bt * p["TB"] + tb, rt * p["TH"] + th
                    if tb < p["TB"] and th < p["TH"] and wc < W and b < B and h < H:
                        y[b, g_first * cout_g:(g_first + p["slab"]) * cout_g, h, wc] = acc[pos]
    # reference: plain grouped correlation
    ref = np.zeros_like(y)
    xpad = np.pad(x, ((0, 0), (0, 0), (p["pad"],) * 2, (p["pad"],) * 2))
    for g in range(G):
        for n in range(cout_g):
            k = g * cout_g + n
            for c in range(cin_g):
                for r in range(R):
                    for s2 in range(R):
                        ref[:, k] += w[k, c, r, s2] * xpad[:, g * cin_g + c, r:r + H, s2:s2 + W]
    err = np.abs(y - ref).max()
    print(f"B={B} C={Cin} K={Cout} {H}x{W} R={R} G={G}: tile TH={p['TH']} TB={p['TB']} slab={p['slab']} max |diff| = {err}")
    assert err == 0


if __name__ == "__main__":
    run(3, 32, 32, 8, 8, 1, 2)
    run(2, 32, 64, 16, 16, 3, 2)
    run(5, 16, 16, 8, 8, 1, 1)
    run(2, 64, 32, 9, 12, 3, 4)
