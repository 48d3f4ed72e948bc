#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
for dbg in 26 27 11 19; do
  MNB_PK_DEBUG=$dbg timeout 300 python -m harness.pk_probe --only conv2_x --compact > $O/probe_dbg$dbg.log 2> $O/probe_dbg$dbg.txt
  echo "== probe dbg=$dbg"; grep "^  resnet" $O/probe_dbg$dbg.txt | cut -c1-330
done
for st in 2 8; do
  MNB_PK_STAGES=$st timeout 300 python -m harness.pk_probe --only conv2_x --compact > $O/probe_st$st.log 2> $O/probe_st$st.txt
  echo "== probe stages=$st"; grep "^  resnet" $O/probe_st$st.txt | cut -c1-330
  MNB_PK_DEBUG=7 MNB_PK_STAGES=$st timeout 300 python -m harness.pk_probe --only conv2_x --compact > $O/probe_st${st}_dbg7.log 2> $O/probe_st${st}_dbg7.txt
  echo "== probe stages=$st dbg=7"; grep "^  resnet" $O/probe_st${st}_dbg7.txt | cut -c1-330
done
for mt in 1 2; do
  MNB_PK_MT=$mt timeout 300 python -m harness.pk_probe --only conv2_x --compact > $O/probe_mt$mt.log 2> $O/probe_mt$mt.txt
  echo "== probe mt=$mt"; grep "^  resnet" $O/probe_mt$mt.txt | cut -c1-330
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > $O/parity.log 2>&1
echo "== parity rc=$?"; grep -E "^FAILED|passed|failed" $O/parity.log | cut -c1-200 | tail -8; grep -E "^E   " $O/parity.log | cut -c1-250 | head -12
