#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pk.py -q -p no:cacheprovider > $O/pk.log 2>&1
echo "== pk rc=$?"; grep -E "^FAILED|^E  .*Error|passed|failed" $O/pk.log | cut -c1-250 | tail -12
timeout 900 python -m harness.pk_probe --json $O/pk_probe.json > $O/pk_probe.log 2>&1; echo "== probe rc=$?"; tail -2 $O/pk_probe.log | cut -c1-200
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_pk.py > $O/full.log 2>&1
echo "== full rc=$?"; grep -E "^FAILED|passed|failed" $O/full.log | cut -c1-250 | tail -20
for w in nin_gc_wbwtab_w3a2 nin_gc_dorefa_w4a4 nin_dorefa_w8a8 resnet18_iao_w8a8_bnfuse resnet18_iao_ptq_224; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra --kernels-json $O/k_$w.json > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-200
done
for w in nin_gc_dorefa_w4a4 resnet18_iao_w8a8_bnfuse resnet18_iao_ptq_224; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_$w.log 2>&1
  echo "== launches $w rc=$?"; wc -l $O/launches_$w.csv
done
