#!/bin/bash
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_inference.py tests/test_gpu_pk.py tests/test_gpu_fused_dorefa.py tests/test_gpu_iao_wrappers.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -8; grep -E "^E   " $O/tests.log | cut -c1-250 | head -12
timeout 300 python -m harness.debug.wgrad_real > $O/wgrad_real.log 2>&1; echo "== wgrad_real"; tail -16 $O/wgrad_real.log | cut -c1-330
timeout 300 python -m harness.pk_probe --only conv2_x --compact > $O/probe.log 2> $O/probe.txt
echo "== probe"; grep "^  resnet" $O/probe.txt | cut -c1-400
for w in resnet18_iao_ptq_224 resnet18_iao_w8a8_bnfuse; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-200
done
