#!/bin/bash
# round-2 GPU call e: ncu source-level captures of the packed-operand kernels (where do the cycles go?)
O=gpurun_out/r2e; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pk_conv_kernel -s 12 -c 8 -o $O/resnet_conv python bench.py --workload resnet18_iao_w8a8_bnfuse --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu1.log 2>&1
echo "== ncu1 rc=$?"; tail -2 $O/ncu1.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pk_wgrad_kernel -s 4 -c 3 -o $O/resnet_wgrad python bench.py --workload resnet18_iao_w8a8_bnfuse --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu2.log 2>&1
echo "== ncu2 rc=$?"; tail -2 $O/ncu2.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pk_conv_kernel -s 0 -c 4 -o $O/ningc_conv python bench.py --workload nin_gc_dorefa_w4a4 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu3.log 2>&1
echo "== ncu3 rc=$?"; tail -2 $O/ncu3.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_iao_wrappers.py tests/test_gpu_pk.py -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; tail -3 $O/tests.log | cut -c1-200
ls -la $O
