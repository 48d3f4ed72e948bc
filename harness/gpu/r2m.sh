#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
for dbg in 0 1 2 4 5 6; do
  MNB_PK_DEBUG=$dbg timeout 200 python -m harness.pk_probe --only conv2_x > $O/probe_dbg$dbg.log 2>&1
  echo "dbg=$dbg $(tail -1 $O/probe_dbg$dbg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('fwd_1x1','fwd_3x3','dgrad_3x1','dgrad_3x3')})")"
done
for dbg in 0 2 5; do
  MNB_PK_DEBUG=$dbg timeout 200 python -m harness.pk_probe --only conv5_x > $O/probe5_dbg$dbg.log 2>&1
  echo "conv5 dbg=$dbg $(tail -1 $O/probe5_dbg$dbg.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],1) for k in ('fwd_1x1','fwd_3x3','dgrad_3x1','dgrad_3x3')})")"
done
MNB_PK_TERMS=2 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_pk.py > $O/full_t2.log 2>&1
echo "== full T=2 rc=$?"; grep -E "^FAILED|passed|failed" $O/full_t2.log | cut -c1-200 | tail -30
grep -E "^E   " $O/full_t2.log | cut -c1-220 | head -30
MNB_PK_TERMS=2 timeout 600 python bench.py --workload resnet18_iao_w8a8_bnfuse --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_resnet_t2.log 2>&1
echo "== bench resnet T=2 rc=$?"; tail -1 $O/bench_resnet_t2.log | cut -c1-160
