#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_inference.py -m gpu -q -p no:cacheprovider > $O/infer.log 2>&1
echo "== inference tests rc=$?"; grep -E "^FAILED|passed|failed" $O/infer.log | cut -c1-200 | tail -8; grep -E "^E   " $O/infer.log | cut -c1-300 | head -20
timeout 600 python bench.py --workload resnet18_iao_ptq_224 --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_ptq.log 2>&1
echo "== bench ptq rc=$?"; tail -1 $O/bench_ptq.log | cut -c1-200
timeout 300 python -m harness.debug.wgrad_cond > $O/wgrad_cond.log 2>&1; echo "== wgrad_cond"; tail -20 $O/wgrad_cond.log
for dbg in 3 7; do
  MNB_PK_DEBUG=$dbg timeout 300 python -m harness.pk_probe --only conv2_x --compact > $O/probe_dbg$dbg.log 2> $O/probe_dbg$dbg.txt
  echo "== probe dbg=$dbg"; grep "^  resnet" $O/probe_dbg$dbg.txt | cut -c1-400
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pk_conv -c 2 -o $O/pkconv_conv2 -f python -m harness.pk_one conv2_x 1 > $O/ncu.log 2>&1
echo "== ncu rc=$?"; ls -la $O/*.ncu-rep
