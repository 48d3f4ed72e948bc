#!/bin/bash
# round 2, second session, call 2: XNOR-popcount vs tensor-core forward, PK_WBWTAB A/B, headline launch list, ncu of the 1x1 dgrad
O=gpurun_out/r2z; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_xnor.py tests/test_gpu_packed_experimental.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -12; grep -E "^E   " $O/tests.log | cut -c1-300 | head -16
timeout 200 python -m harness.xnor_probe --json $O/xnor_probe.json > $O/xnor_probe.log 2>&1
echo "== xnor probe rc=$?"; python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/r2z/xnor_probe.json")):
        print(r["layer"], "xnor", round(r["xnor_conv"],1), "+pack", round(r["xnor_pack_act"],1), "| tc", round(r["tc_conv"],1), "+pack", round(r["tc_pack_act"],1), "| out-only roof", round(r["roof_us_out_only"],1), r["bit_identical"])
except Exception as e:
    print("no probe json", e)
PY
tail -3 $O/xnor_probe.log | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --kernels-json $O/kern_on.json > $O/bench_on.log 2>&1
echo "== bench PK_WBWTAB=1 rc=$?"; tail -1 $O/bench_on.log | cut -c1-160
MNB_PK_WBWTAB=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --kernels-json $O/kern_off.json > $O/bench_off.log 2>&1
echo "== bench PK_WBWTAB=0 rc=$?"; tail -1 $O/bench_off.log | cut -c1-160
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/launches_headline.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_list.log 2>&1
echo "== launch list rc=$?"; wc -l $O/launches_headline.csv
PK_DY_TERMS=2 timeout 400 ncu --set full --import-source on --clock-control none -k regex:pk_conv -c 3 -o $O/pkconv_1x1g2 -f python -m harness.pk_one "1x1 g2" 1 > $O/ncu_full.log 2>&1
echo "== ncu full rc=$?"; ls -la $O/*.ncu-rep
