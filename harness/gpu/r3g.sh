#!/bin/bash
# final state of round 2: full GPU suite, smoke, default bench (extras + CPU baseline), reference arm, launch list
O=gpurun_out/r3g; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -14; grep -E "^E   " $O/tests.log | cut -c1-300 | head -24
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 $O/smoke.log | cut -c1-300
timeout 400 python bench.py --steps 20 --warmup 5 --kernels-json $O/kern.json > $O/bench.log 2> $O/bench.err
echo "== bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r3g/bench.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["roofline"]["engine_conv_share_of_step"], d["roofline"]["all_conv_kernels_frac"], d["roofline"]["traffic"], d["gpu_launches"], d["clocks"], d["cpu_baseline"]["value"])
    for e in d.get("extra_workloads", []): print(" ", e.get("workload"), e.get("value"), e.get("ms_per_step"), e.get("e2e",{}).get("value"), e.get("error"))
    for r in sorted(json.load(open("gpurun_out/r3g/kern.json"))["kernels"], key=lambda r:(r["shape"][1],r["shape"][5],r["kind"])):
        s=r["shape"]; print(f"  {r['kind']:10s} C{s[1]} {s[2]}x{s[3]} K{s[4]} R{s[5]} g{s[13]} {r['avg_us']:.1f}us frac {r['frac_of_roof']:.2f}")
except Exception as e: print("bench parse", e)
PY
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1; echo "== ref arm rc=$?"; tail -1 $O/bench_ref.log | cut -c1-300
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/launches_headline.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_list.log 2>&1
echo "== launch list rc=$?"; wc -l $O/launches_headline.csv
