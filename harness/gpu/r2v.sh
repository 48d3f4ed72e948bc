#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "fused_headline" > $O/parity.log 2>&1
echo "== parity rc=$?"; grep -E "^FAILED|passed|failed" $O/parity.log | cut -c1-200 | tail -8; grep -E "^E   " $O/parity.log | cut -c1-250 | head -12
timeout 600 ncu --set full --import-source on --clock-control none -k regex:pk_conv -c 3 -o $O/pkconv_conv2 -f python -m harness.pk_one conv2_x 1 > $O/ncu.log 2>&1
echo "== ncu rc=$?"; ls -la $O/*.ncu-rep
MNB_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_ptq.csv python bench.py --workload resnet18_iao_ptq_224 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/ncu_ptq.log 2>&1
echo "== ncu ptq rc=$?"; wc -l $O/launches_ptq.csv
