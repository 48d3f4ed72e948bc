#!/bin/bash
# head conv on the packed family, QuantConvTranspose2d, XNOR dispatch; ncu of the 3x3 grouped forward and of the XNOR / tc pair
O=gpurun_out/r3b; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused_bn.py tests/test_gpu_conv_transpose.py tests/test_gpu_xnor.py tests/test_gpu_packed_experimental.py -m gpu -q -p no:cacheprovider > $O/tests.log 2>&1
echo "== tests rc=$?"; grep -E "^FAILED|passed|failed" $O/tests.log | cut -c1-200 | tail -12; grep -E "^E   " $O/tests.log | cut -c1-300 | head -20
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "convT or fused_headline or model_first_step or teacher_forced" > $O/parity.log 2>&1
echo "== parity rc=$?"; grep -E "^FAILED|passed|failed" $O/parity.log | cut -c1-200 | tail -8; grep -E "^E   " $O/parity.log | cut -c1-300 | head -16
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 $O/smoke.log | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --kernels-json $O/kern.json > $O/bench.log 2>&1
echo "== bench rc=$?"; tail -1 $O/bench.log | cut -c1-160
timeout 300 ncu --set full --import-source on --clock-control none -k regex:pk_conv -c 1 -o $O/pkconv_3x3g16 -f python -m harness.pk_one "3x3 g16" 1 > $O/ncu_a.log 2>&1
echo "== ncu 3x3 rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k "regex:xnor::conv_kernel|pk_conv" -c 4 -o $O/xnor_vs_tc_1x1g2 -f python -m harness.xnor_probe --only "1x1 g2" --iters 1 > $O/ncu_b.log 2>&1
echo "== ncu xnor rc=$?"; ls -la $O/*.ncu-rep
