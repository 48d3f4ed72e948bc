#!/bin/bash
# last GPU window of the round (4 minutes): the shipped state through the whole GPU suite (-x) and the headline bench
O=gpurun_out/r3h; mkdir -p $O
timeout 150 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout 60 > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -6; grep -E "^E   " $O/tests.log | cut -c1-300 | head -8
timeout 60 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --kernels-json $O/kern.json > $O/bench.log 2> $O/bench.err
echo "== bench rc=$?"; tail -1 $O/bench.log | cut -c1-400
