#!/bin/bash
# round-2 GPU call f: MMA rate probe + isolated-layer timing of the packed-operand kernels
O=gpurun_out/r2f; mkdir -p $O
timeout 300 python harness/mma_rate_probe.py > $O/mma_rate.log 2>&1; echo "== rate rc=$?"; cat $O/mma_rate.log | cut -c1-120
timeout 900 python -m harness.pk_probe --json $O/pk_probe.json > $O/pk_probe.log 2>&1; echo "== probe rc=$?"; tail -12 $O/pk_probe.log | cut -c1-900
