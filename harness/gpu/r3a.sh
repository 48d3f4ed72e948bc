#!/bin/bash
# vectorised BatchNorm+binarizer producers (forward plane, backward gradient pieces), unsegmented short-chain data gradients
O=gpurun_out/r3a; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused_bn.py tests/test_gpu_pk.py tests/test_gpu_packed_experimental.py tests/test_gpu_xnor.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -12; grep -E "^E   " $O/tests.log | cut -c1-300 | head -16
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_headline or teacher_forced or full_size" > $O/parity.log 2>&1
echo "== parity rc=$?"; grep -E "^FAILED|passed|failed" $O/parity.log | cut -c1-200 | tail -8; grep -E "^E   " $O/parity.log | cut -c1-300 | head -16
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --kernels-json $O/kern_on.json > $O/bench_on.log 2>&1
echo "== bench rc=$?"; tail -1 $O/bench_on.log | cut -c1-160
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/launches_headline.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_list.log 2>&1
echo "== launch list rc=$?"; wc -l $O/launches_headline.csv
for w in resnet18_iao_w8a8_bnfuse nin_gc_dorefa_w4a4; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-160
done
