#!/usr/bin/env python
"""Turn an `ncu --set full` report of one bench step into the committed summaries:

    python profiles/extract_ncu.py gpurun_out/prof_tc.ncu-rep r1

writes profiles/<tag>_tc_kernels.md (per-launch metric table) and profiles/kernel_traffic.json
(measured DRAM bytes per launch, keyed like bench.py's roofline kernels).  The capture must hold exactly
one QAT step's tensor-core launches of the bench workload, in launch order: 7 forward convs (layers 2..8
of NIN-GC), then for layers 8..2 one dgrad followed by one wgrad."""
import csv
import io
import json
import os
import subprocess
import sys

LAYERS = [  # (B, C, H, W, K, R, S, sh, sw, ph, pw, dh, dw, G) of the 7 quantized NIN-GC convs at batch 256
    (256, 256, 32, 32, 256, 1, 1, 1, 1, 0, 0, 1, 1, 2), (256, 256, 32, 32, 256, 1, 1, 1, 1, 0, 0, 1, 1, 2),
    (256, 256, 16, 16, 512, 3, 3, 1, 1, 1, 1, 1, 1, 16), (256, 512, 16, 16, 512, 1, 1, 1, 1, 0, 0, 1, 1, 4),
    (256, 512, 16, 16, 512, 1, 1, 1, 1, 0, 0, 1, 1, 4), (256, 512, 8, 8, 1024, 3, 3, 1, 1, 1, 1, 1, 1, 32),
    (256, 1024, 8, 8, 1024, 1, 1, 1, 1, 0, 0, 1, 1, 8)]
WANT = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {w: hdr.index(w) for w in WANT if w in hdr}
    recs = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if "conv_tc_kernel" not in name and "wgrad_tc_kernel" not in name:
            continue
        recs.append({w: r[i] for w, i in idx.items()})
    kinds = ["fwd_tc"] * 7
    order = list(range(7))
    for layer in reversed(range(7)):
        kinds += ["dgrad_tc", "wgrad_tc"]
        order += [layer, layer]
    here = os.path.dirname(os.path.abspath(__file__))
    traffic = {}
    with open(os.path.join(here, f"{tag}_tc_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of the tensor-core conv kernels, one QAT step\n\n")
        f.write("NIN-GC wbwtab W-ternary/A-binary, batch 256, 1 x B200.  Algorithmic bytes = 4(B*C*H*W + B*K*P*Q) + 4|W|"
                " (SURVEY.md 8d); `dram` = dram__bytes_read.sum + dram__bytes_write.sum of that launch.\n\n")
        f.write("| # | kernel | layer (C,HxW,K,RxS,g) | time us | dram MB | algorithmic MB | dram % of peak | tensor pipe % | regs | dyn smem KB |\n")
        f.write("|---|---|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for i, rec in enumerate(recs[:len(kinds)]):
            L = LAYERS[order[i]]
            B, C, H, W, K, R, S = L[:7]
            algo = (4.0 * (B * C * H * W + B * K * H * W) + 4.0 * K * (C // L[-1]) * R * S) / 1e6
            dur = float(rec["gpu__time_duration.sum"])
            rd, wr = float(rec["dram__bytes_read.sum"]), float(rec["dram__bytes_write.sum"])
            ru = units[idx["dram__bytes_read.sum"]]
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(ru, 1.0)
            dram_mb = (rd + wr) * scale
            tu = units[idx["gpu__time_duration.sum"]]
            dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(tu, 1.0)
            f.write(f"| {i} | {kinds[i]} | {C},{H}x{W},{K},{R}x{S},g{L[-1]} | {dur_us:.1f} | {dram_mb:.1f} | {algo:.1f} | "
                    f"{float(rec['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):.1f} | "
                    f"{float(rec['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']):.1f} | "
                    f"{rec['launch__registers_per_thread']} | {float(rec['launch__shared_mem_per_block_dynamic']):.0f} |\n")
            traffic[f"{kinds[i]}:{list(L)}"] = {"dram_bytes": dram_mb * 1e6, "ncu_time_us": dur_us, "algorithmic_bytes": algo * 1e6}
    json.dump({"source": os.path.basename(rep), "kernels": traffic}, open(os.path.join(here, "kernel_traffic.json"), "w"), indent=1)
    print("wrote", len(traffic), "kernels")


if __name__ == "__main__":
    main()
