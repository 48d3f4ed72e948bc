#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pk.py tests/test_gpu_inference.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -8; grep -E "^E   " $O/tests.log | cut -c1-250 | head -12
for dbg in 0 27 59; do
  MNB_PK_DEBUG=$dbg timeout 300 python -m harness.pk_probe --only "conv2_x 64->64 3x3 @32" --compact > $O/probe_dbg$dbg.log 2> $O/probe_dbg$dbg.txt
  echo "== probe dbg=$dbg"; grep "^  resnet" $O/probe_dbg$dbg.txt | cut -c1-330
done
timeout 300 python -m harness.pk_probe --only resnet --compact > $O/probe.log 2> $O/probe.txt
echo "== probe"; grep "^  resnet" $O/probe.txt | cut -c1-400
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pk_wgrad -c 1 -o $O/pkwgrad_conv2 -f python -m harness.pk_one conv2_x 1 > $O/ncu.log 2>&1
echo "== ncu rc=$?"; ls -la $O/*.ncu-rep
for w in resnet18_iao_w8a8_bnfuse nin_gc_dorefa_w4a4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-200
done
