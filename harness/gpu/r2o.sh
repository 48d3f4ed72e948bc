#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
timeout 600 python -m harness.debug.ptq_layers 224 > $O/ptq_layers.log 2>&1; tail -70 $O/ptq_layers.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_pk.py --deselect tests/test_gpu_inference.py::test_eval_logits_match_the_oracle_at_224 > $O/full.log 2>&1
echo "== full (no pk) rc=$?"; grep -E "^FAILED|passed|failed" $O/full.log | cut -c1-200 | tail -30
grep -E "^E   " $O/full.log | cut -c1-220 | head -40
for w in nin_gc_dorefa_w4a4 nin_dorefa_w8a8; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_$w.log 2>&1
  echo "== bench $w rc=$?"; tail -1 $O/bench_$w.log | cut -c1-200
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_default.log 2>&1
echo "== bench default rc=$?"; tail -1 $O/bench_default.log | cut -c1-300
