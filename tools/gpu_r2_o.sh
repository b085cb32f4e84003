#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
for WPC in 0 13 10 8; do echo "wpc=$WPC"; DGS_SDDMM_WPC=$WPC timeout 300 python tools/exp_sddmm_k.py 2>/dev/null | grep -E " 256 | 512 "; done
for ST in 3; do echo "stages=$ST wpc=8"; DGS_SDDMM_STAGES=$ST DGS_SDDMM_WPC=8 timeout 300 python tools/exp_sddmm_k.py 2>/dev/null | grep -E " 256 | 512 "; done
