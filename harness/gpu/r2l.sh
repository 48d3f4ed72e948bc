#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/full.log 2>&1
echo "== full rc=$?"; grep -E "^FAILED|passed|failed" $O/full.log | cut -c1-250 | tail -20
timeout 600 python -m harness.pk_probe --json $O/pk_probe.json > $O/pk_probe.log 2>&1; echo "== probe rc=$?"; tail -1 $O/pk_probe.log | cut -c1-200
for w in nin_gc_wbwtab_w3a2 nin_gc_dorefa_w4a4 nin_dorefa_w8a8 resnet18_iao_w8a8_bnfuse resnet18_iao_ptq_224; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra --kernels-json $O/k_$w.json > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-160; tail -1 $O/bench_$w.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   graph', str(d.get('cuda_graph'))[:300])"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pk_ -c 8 -o $O/one_conv2x python -m harness.pk_one conv2_x 2 > $O/ncu_one.log 2>&1; echo "== ncu one rc=$?"
MNB_PK_STAGES=2 timeout 200 python -m harness.pk_probe --only conv2_x > $O/probe_st2.log 2>&1; tail -1 $O/probe_st2.log | cut -c1-700
