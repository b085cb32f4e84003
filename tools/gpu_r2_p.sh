#!/bin/bash
# segment count vs feature width on one GPU (several column panels per launch): does one wave per panel hurt wide matrices?
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for SEGS in 1 2 4; do for N in 128 256 512; do
  echo -n "segs=$SEGS N=$N: "; DGS_SPMM_SEGS=$SEGS timeout 300 python tools/exp_panels.py reddit $N 2>/dev/null | head -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_call'],4), round(d['kernel_ms'],4), round(d['fixup_ms'],4))"
done; done
