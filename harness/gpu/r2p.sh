#!/bin/bash
O=gpurun_out/r2p; mkdir -p $O
for c in nin_gc_wb_binary resnet_iao_bnfuse; do
  timeout 300 python -m harness.debug.tf_detail $c > $O/tf_$c.log 2>&1; echo "== tf_detail $c"; grep -E "model.8|conv2_x.1.residual_function.3" $O/tf_$c.log | head -12
done
for tb in 2 3; do
  rm -f $O/errlog_t$tb.txt
  MNB_PK_TERMS_BWD=$tb MNB_TEST_ERRLOG=$O/errlog_t$tb.txt timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "teacher_forced or full_size" > $O/parity_t$tb.log 2>&1
  echo "== parity T_bwd=$tb rc=$?"; grep -E "^FAILED|passed|failed" $O/parity_t$tb.log | cut -c1-200 | tail -8; grep -E "^E   " $O/parity_t$tb.log | cut -c1-250 | head -12
done
timeout 1200 python -m pytest tests/test_gpu_inference.py tests/test_gpu_bn_fuse.py tests/test_gpu_tc_conv.py tests/test_gpu_fused_bn.py -m gpu -q -p no:cacheprovider > $O/misc.log 2>&1
echo "== misc rc=$?"; grep -E "^FAILED|passed|failed" $O/misc.log | cut -c1-200 | tail -12; grep -E "^E   " $O/misc.log | cut -c1-250 | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_default.log 2>&1
echo "== bench default rc=$?"; tail -1 $O/bench_default.log | cut -c1-200
for dbg in 0 2 10 18; do
  MNB_PK_DEBUG=$dbg timeout 300 python -m harness.pk_probe --only resnet --compact > $O/probe_dbg$dbg.log 2> $O/probe_dbg$dbg.txt
  echo "== probe dbg=$dbg"; grep "^  resnet" $O/probe_dbg$dbg.txt | cut -c1-400
done
MNB_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_resnet.csv python bench.py --workload resnet18_iao_w8a8_bnfuse --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_resnet.log 2>&1
echo "== ncu rc=$?"; wc -l $O/launches_resnet.csv
