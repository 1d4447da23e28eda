#!/bin/bash
cd $GRAFT_REPO_ROOT
export ECH_ONE_SETUP=1 ECH_MG_CONCURRENT_FACTOR=0
python tools/vcycle_once.py > gpurun_out/r2z_vcycle_once.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k "regex:ilu0_apply_packed|ilu0_factor_lcol" -c 9 -o gpurun_out/r2z_ilu_kernels -f python tools/vcycle_once.py > gpurun_out/r2z_ilu_ncu.log 2>&1
tail -3 gpurun_out/r2z_ilu_ncu.log; ls -la gpurun_out/r2z_ilu_kernels.ncu-rep
