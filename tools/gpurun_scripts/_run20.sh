#!/bin/bash
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for mode in cg box both; do
  ECH_GCOARSE=$mode timeout 300 $TR tools/dist_mg_check.py 256 2>&1 | grep "DIST_MG\|Error\|error" | sed "s/^/[$mode] /" >> gpurun_out/r2w_dist_mg_2gpu.log
done
ECH_GCOARSE=cg timeout 600 $TR tools/dist_mg_check.py 1408 2>&1 | grep "DIST_MG\|Error\|error" | sed "s/^/[cg 1408] /" >> gpurun_out/r2w_dist_mg_2gpu.log
cat gpurun_out/r2w_dist_mg_2gpu.log
