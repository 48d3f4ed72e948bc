#!/bin/bash
# round-2 GPU call b: first hardware run of the packed-operand family + ncu capture of the fused-quantizer forward
O=gpurun_out/r2b; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
for k in pack_act forward_integer forward_fp32 data_gradient weight_gradient module_path; do
  timeout 400 python -m pytest tests/test_gpu_pk.py -q -k $k -p no:cacheprovider > $O/pk_$k.log 2>&1
  echo "== $k rc=$?"; tail -4 $O/pk_$k.log | cut -c1-300
done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_pk.py > $O/full.log 2>&1
echo "== full rc=$?"; tail -6 $O/full.log | cut -c1-300
MNB_PK=off timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 7 -o $O/dorefa_fwd python bench.py --workload nin_gc_dorefa_w4a4 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu.log 2>&1
echo "== ncu rc=$?"; tail -2 $O/ncu.log | cut -c1-300; ls -la $O
