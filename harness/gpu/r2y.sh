#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fused_bn.py tests/test_gpu_pk.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -8; grep -E "^E   " $O/tests.log | cut -c1-300 | head -16
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_headline or teacher_forced" > $O/parity.log 2>&1
echo "== parity rc=$?"; grep -E "^FAILED|passed|failed" $O/parity.log | cut -c1-200 | tail -8; grep -E "^E   " $O/parity.log | cut -c1-300 | head -16
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra --kernels-json $O/kernels_headline.json > $O/bench_default.log 2>&1
echo "== bench default rc=$?"; tail -1 $O/bench_default.log | cut -c1-250
MNB_PK_WBWTAB=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_default_off.log 2>&1
echo "== bench default (MNB_PK_WBWTAB=0) rc=$?"; tail -1 $O/bench_default_off.log | cut -c1-250
timeout 300 python -m harness.pk_probe --only "nin-gc" --compact > $O/probe_nin.log 2> $O/probe_nin.txt
echo "== nin layers"; grep "^  nin" $O/probe_nin.txt | cut -c1-330
timeout 300 python bench.py --workload resnet18_iao_w8a8_bnfuse --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_resnet.log 2>&1
echo "== bench resnet rc=$?"; tail -1 $O/bench_resnet.log | cut -c1-200
