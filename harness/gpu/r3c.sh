#!/bin/bash
# 8 epilogue warps in pk_conv, plane-only producers, pooled BN backward pack, head conv on pk, XNOR dispatch fix: full suite + bench
O=gpurun_out/r3c; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -14; grep -E "^E   " $O/tests.log | cut -c1-300 | head -24
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernels-json $O/kern.json > $O/bench.log 2> $O/bench.err
echo "== bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r3c/bench.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["gpu_launches"])
    for e in d.get("extra_workloads", []): print(" ", e.get("workload"), e.get("value"), e.get("ms_per_step"), e.get("error"))
except Exception as e: print("bench parse", e)
PY
MNB_PLANE_ONLY=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > $O/bench_noplane.log 2>&1
echo "== bench MNB_PLANE_ONLY=0 rc=$?"; tail -1 $O/bench_noplane.log | cut -c1-140
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/launches_headline.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_list.log 2>&1
echo "== launch list rc=$?"; wc -l $O/launches_headline.csv
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:^conv_kernel|pk_conv" -c 2 -o $O/xnor_vs_tc_1x1g2 -f python -m harness.xnor_probe --only "1x1 g2" --iters 1 > $O/ncu_b.log 2>&1
echo "== ncu xnor rc=$?"
PK_DY_TERMS=2 timeout 300 ncu --set full --import-source on --clock-control none -k "regex:pk_conv|pk_wgrad" -c 4 -o $O/pk_1x1g2 -f python -m harness.pk_one "1x1 g2" 1 > $O/ncu_c.log 2>&1
echo "== ncu pk rc=$?"; ls -la $O/*.ncu-rep
