#!/usr/bin/env python
"""Turn the raw-page CSV of an `ncu --set full` capture of one bench step into the committed summaries:

    ncu --set full --clock-control none -k "regex:conv_tc_kernel|wgrad_tc_kernel|bn_sign|channel_stats|tcfp32|maxpool" \
        -s 168 -c 56 -o /tmp/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline
    ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/r1b_full_raw.csv      # (the .ncu-rep itself is > 64 MiB)
    python profiles/extract_ncu.py gpurun_out/r1b_full_raw.csv r1b

writes profiles/<tag>_tc_kernels.md (per-launch table of the tensor-core conv kernels, mapped to their layers),
profiles/<tag>_fused_kernels.md (the producer-side kernels) and profiles/kernel_traffic.json (measured DRAM bytes per
launch, keyed like bench.py's roofline kernels).  The window may start anywhere in a step: forward convs are the
conv_tc launches that follow a bn_sign forward kernel, dgrads are followed by their wgrad; a step starts at the forward
conv whose BatchNorm statistics launch follows a backward kernel."""
import collections
import csv
import json
import os
import sys

LAYERS = [  # (B, C, H, W, K, R, S, sh, sw, ph, pw, dh, dw, G) of the 7 quantized NIN-GC convs at batch 256
    (256, 256, 32, 32, 256, 1, 1, 1, 1, 0, 0, 1, 1, 2), (256, 256, 32, 32, 256, 1, 1, 1, 1, 0, 0, 1, 1, 2),
    (256, 256, 16, 16, 512, 3, 3, 1, 1, 1, 1, 1, 1, 16), (256, 512, 16, 16, 512, 1, 1, 1, 1, 0, 0, 1, 1, 4),
    (256, 512, 16, 16, 512, 1, 1, 1, 1, 0, 0, 1, 1, 4), (256, 512, 8, 8, 1024, 3, 3, 1, 1, 1, 1, 1, 1, 32),
    (256, 1024, 8, 8, 1024, 1, 1, 1, 1, 0, 0, 1, 1, 8)]
SCALE_B = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
SCALE_T = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main():
    path, tag = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(path)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {n: hdr.index(n) for n in hdr}

    def val(r, n, scale=None):
        v = float(r[ix[n]].replace(",", ""))
        return v * scale.get(units[ix[n]], 1.0) if scale else v

    recs = []
    for r in body:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        recs.append({"name": name, "us": val(r, "gpu__time_duration.sum", SCALE_T),
                     "dram_mb": val(r, "dram__bytes_read.sum", SCALE_B) + val(r, "dram__bytes_write.sum", SCALE_B),
                     "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     "tensor_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                     "regs": r[ix["launch__registers_per_thread"]],
                     "smem_kb": val(r, "launch__shared_mem_per_block_dynamic")})
    fwd, bwd = [], []
    for i, rec in enumerate(recs):
        if "conv_tc_kernel" not in rec["name"]:
            continue
        if i + 1 < len(recs) and "wgrad_tc_kernel" in recs[i + 1]["name"]:
            bwd.append((i, i + 1))
        elif i > 0 and "bn_sign" in recs[i - 1]["name"] and "fwd" in recs[i - 1]["name"]:
            fwd.append(i)
    start = next(j for j, i in enumerate(fwd) if i >= 3 and "bwd" in recs[i - 3]["name"])
    assert len(fwd) >= 7 and len(bwd) >= 7, (len(fwd), len(bwd))
    table = []  # (kind, layer index, record)
    seen = set()
    for j, i in enumerate(fwd):
        layer = (j - start) % 7
        if layer not in seen:
            seen.add(layer)
            table.append(("fwd_tc", layer, recs[i]))
    table.sort(key=lambda t: t[1])
    first_full = next(j for j in range(len(bwd)) if j + 7 <= len(bwd) and recs[bwd[j][0]]["smem_kb"] and
                      abs(recs[bwd[j][1]]["dram_mb"] - 138) < 40)  # D8/W8: the 1024-channel 8x8 layer opens the backward pass
    for k in range(7):
        d, w = bwd[first_full + k]
        table.append(("dgrad_tc", 6 - k, recs[d]))
        table.append(("wgrad_tc", 6 - k, recs[w]))
    here = os.path.dirname(os.path.abspath(__file__))
    traffic = {}
    with open(os.path.join(here, f"{tag}_tc_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of the tensor-core conv kernels, one QAT step\n\n")
        f.write("NIN-GC wbwtab W-ternary/A-binary, batch 256, 1 x B200 (command in profiles/extract_ncu.py).  Algorithmic bytes = "
                "4(B*C*H*W + B*K*P*Q) + 4|W| (SURVEY.md 8d) - the forward conv after a fused 2x2 pool reads a quarter of that "
                "input; `dram` = dram__bytes_read.sum + dram__bytes_write.sum of that launch (writes that are still in the "
                "126 MB L2 when the consumer reads them never reach DRAM).\n\n")
        f.write("| # | kernel | layer (C,HxW,K,RxS,g) | time us | dram MB | algorithmic MB | dram % of peak | tensor pipe % | regs | dyn smem KB |\n")
        f.write("|---|---|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for n, (kind, layer, rec) in enumerate(table):
            L = LAYERS[layer]
            B, C, H, W, K, R, S = L[:7]
            algo = (4.0 * (B * C * H * W + B * K * H * W) + 4.0 * K * (C // L[-1]) * R * S) / 1e6
            f.write(f"| {n} | {kind} | {C},{H}x{W},{K},{R}x{S},g{L[-1]} | {rec['us']:.1f} | {rec['dram_mb']:.1f} | {algo:.1f} | "
                    f"{rec['dram_pct']:.1f} | {rec['tensor_pct']:.1f} | {rec['regs']} | {rec['smem_kb']:.0f} |\n")
            traffic[f"{kind}:{list(L)}"] = {"dram_bytes": rec["dram_mb"] * 1e6, "ncu_time_us": rec["us"],
                                            "algorithmic_bytes": algo * 1e6}
    agg = collections.OrderedDict()
    for rec in recs:
        if "conv_tc_kernel" in rec["name"] or "wgrad_tc_kernel" in rec["name"]:
            continue
        a = agg.setdefault(rec["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += rec["us"]; a[2] += rec["dram_mb"]; a[3] += rec["dram_pct"]
    with open(os.path.join(here, f"{tag}_fused_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full` of the producer-side kernels (BatchNorm + binarizer [+ pool], statistics), same capture\n\n")
        f.write("| kernel | launches in window | avg time us | avg dram MB | avg dram % of peak | achieved GB/s (dram MB / time) |\n|---|---:|---:|---:|---:|---:|\n")
        for name, (n, us, mb, pct) in agg.items():
            f.write(f"| `{name}` | {n} | {us / n:.1f} | {mb / n:.1f} | {pct / n:.1f} | {mb / us * 1e3:.0f} |\n")
    json.dump({"source": os.path.basename(path), "kernels": traffic}, open(os.path.join(here, "kernel_traffic.json"), "w"), indent=1)
    print("wrote", len(traffic), "tensor-core kernels,", len(agg), "fused kernel kinds")


if __name__ == "__main__":
    main()
