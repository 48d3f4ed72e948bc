#!/bin/bash
# fast divmod in the TMA / epilogue roles of pk_conv; XNOR pixels-per-thread by receptive field
O=gpurun_out/r3e; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pk.py tests/test_gpu_xnor.py tests/test_gpu_inference.py tests/test_gpu_conv_transpose.py tests/test_gpu_fused_bn.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -12; grep -E "^E   " $O/tests.log | cut -c1-300 | head -16
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernels-json $O/kern.json > $O/bench.log 2> $O/bench.err
echo "== bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r3e/bench.log").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["roofline"]["traffic"], d["gpu_launches"])
    for e in d.get("extra_workloads", []): print(" ", e.get("workload"), e.get("value"), e.get("ms_per_step"), e.get("error"))
    for r in sorted(json.load(open("gpurun_out/r3e/kern.json"))["kernels"], key=lambda r:(r["shape"][1],r["shape"][5],r["kind"])):
        s=r["shape"]; print(f"  {r['kind']:10s} C{s[1]} {s[2]}x{s[3]} K{s[4]} R{s[5]} g{s[13]} {r['avg_us']:.1f}us frac {r['frac_of_roof']:.2f}")
except Exception as e: print("bench parse", e)
PY
timeout 200 python -m harness.xnor_probe --json $O/xnor_probe.json > $O/xnor_probe.log 2>&1
echo "== xnor probe rc=$?"; python - <<'PY'
import json
try:
    for r in json.load(open("gpurun_out/r3e/xnor_probe.json")):
        print(r["layer"], "xnor", round(r["xnor_conv"],1), "+pack", round(r["xnor_pack_act"],1), "| tc", round(r["tc_conv"],1), "+pack", round(r["tc_pack_act"],1), r["bit_identical"])
except Exception as e:
    print("no probe json", e)
PY
