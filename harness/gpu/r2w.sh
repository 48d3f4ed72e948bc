#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pk.py tests/test_gpu_inference.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -8; grep -E "^E   " $O/tests.log | cut -c1-250 | head -12
for ct in 0 2 4 8 16; do
  if [ $ct = 0 ]; then unset MNB_PK_COLTILES; else export MNB_PK_COLTILES=$ct; fi
  timeout 300 python -m harness.pk_probe --only "ptq224 conv2" --batch 64 --compact > $O/probe224_ct$ct.log 2> $O/probe224_ct$ct.txt
  echo "== 224 coltiles=$ct"; grep "^  ptq" $O/probe224_ct$ct.txt | cut -c1-330
done
for ct in 1 2; do
  MNB_PK_COLTILES=$ct timeout 300 python -m harness.pk_probe --only "conv2_x 64->64 3x3 @32" --compact > $O/probe32_ct$ct.log 2> $O/probe32_ct$ct.txt
  echo "== 32 coltiles=$ct"; grep "^  resnet" $O/probe32_ct$ct.txt | cut -c1-330
done
unset MNB_PK_COLTILES
timeout 300 python -m harness.pk_probe --only "nin" --compact > $O/probe_nin.log 2> $O/probe_nin.txt
echo "== nin layers"; grep "^  nin" $O/probe_nin.txt | cut -c1-330
for w in resnet18_iao_ptq_224 resnet18_iao_w8a8_bnfuse; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-200
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --kernels-json $O/kernels_headline.json > $O/bench_default.log 2>&1
echo "== bench default rc=$?"; tail -1 $O/bench_default.log | cut -c1-200
