#!/bin/bash
# which role bounds the 3x3 grouped layers of the packed-operand family?  plan knobs and role switches, CUDA events (pk_probe)
O=gpurun_out/r3f; mkdir -p $O
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 120 python -m harness.pk_probe --only "nin-gc 3x3" --compact > $O/p_$label.log 2> $O/p_$label.txt
  echo "== $label"; grep "^  nin" $O/p_$label.txt | sed 's/pack_[a-z_0-9]*=[0-9]* //g' | cut -c1-260
}
run base X=0
run mt1 MNB_PK_MT=1
run mt2 MNB_PK_MT=2
run st8 MNB_PK_STAGES=8
run st2 MNB_PK_STAGES=2
run nostore MNB_PK_DEBUG=1
run nomma MNB_PK_DEBUG=2
run nostore_nomma MNB_PK_DEBUG=3
run noepi MNB_PK_DEBUG=7
run noloadA MNB_PK_DEBUG=8
run noloadAB MNB_PK_DEBUG=24
