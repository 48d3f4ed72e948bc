#!/bin/bash
# XNOR kernel v2 (2 pixels per thread, constants in shared memory, border template): tests, probe, ncu pair
O=gpurun_out/r3d; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_xnor.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -12; grep -E "^E   " $O/tests.log | cut -c1-300 | head -16
timeout 200 python -m harness.xnor_probe --json $O/xnor_probe.json > $O/xnor_probe.log 2>&1
echo "== xnor probe rc=$?"; python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/r3d/xnor_probe.json")):
        print(r["layer"], "xnor", round(r["xnor_conv"],1), "+pack", round(r["xnor_pack_act"],1), "| tc", round(r["tc_conv"],1), "+pack", round(r["tc_pack_act"],1), "| out-only roof", round(r["roof_us_out_only"],1), r["bit_identical"])
except Exception as e:
    print("no probe json", e)
PY
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:^conv_kernel|pk_conv" -c 2 -o $O/xnor_vs_tc_1x1g2 -f python -m harness.xnor_probe --only "1x1 g2" --iters 1 > $O/ncu_b.log 2>&1
echo "== ncu xnor rc=$?"; ncu -i $O/xnor_vs_tc_1x1g2.ncu-rep --page raw --csv > $O/xnor_raw.csv 2>/dev/null; ls -la $O
