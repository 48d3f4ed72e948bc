#!/bin/bash
O=gpurun_out/r2n; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_pk.py > $O/full.log 2>&1
echo "== full (no pk) rc=$?"; grep -E "^FAILED|passed|failed" $O/full.log | cut -c1-200 | tail -10
grep -E "^E   " $O/full.log | cut -c1-220 | head -20
timeout 1200 python -m pytest tests/test_gpu_pk.py -m gpu -q -p no:cacheprovider > $O/pk.log 2>&1
echo "== pk rc=$?"; grep -E "^FAILED|passed|failed" $O/pk.log | cut -c1-200 | tail -20
grep -E "^E   " $O/pk.log | cut -c1-220 | head -20
for mt in 1 0; do
  if [ $mt = 1 ]; then export MNB_PK_MT=1; else unset MNB_PK_MT; fi
  timeout 200 python -m harness.pk_probe --only conv2_x > $O/probe2_mt$mt.log 2>&1
  echo "conv2 mt=$mt $(tail -1 $O/probe2_mt$mt.log | cut -c1-400)"
  timeout 200 python -m harness.pk_probe --only conv4_x > $O/probe4_mt$mt.log 2>&1
  echo "conv4 mt=$mt $(tail -1 $O/probe4_mt$mt.log | cut -c1-400)"
done
unset MNB_PK_MT
for w in resnet18_iao_w8a8_bnfuse resnet18_iao_ptq_224; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-200
done
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_default.log 2>&1
echo "== bench default rc=$?"; tail -1 $O/bench_default.log | cut -c1-1500
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
