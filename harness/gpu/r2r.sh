#!/bin/bash
O=gpurun_out/r2r; mkdir -p $O
for dbg in 0 7; do
  MNB_PK_DEBUG=$dbg timeout 300 python -m harness.pk_probe --only resnet --compact > $O/probe_dbg$dbg.log 2> $O/probe_dbg$dbg.txt
  echo "== probe dbg=$dbg"; grep "^  resnet" $O/probe_dbg$dbg.txt | cut -c1-400
done
for w in resnet18_iao_ptq_224 resnet18_iao_w8a8_bnfuse nin_gc_dorefa_w4a4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_inference.py tests/test_gpu_pk.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -8
