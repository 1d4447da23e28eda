#!/bin/bash
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for lv in 2 1 0; do
  ECH_GCOARSE_MAX_VERTICES=100000000 ECH_GCOARSE_LEVEL=$lv timeout 400 $TR tools/dist_mg_check.py 1408 2>&1 | grep "DIST_MG\|Error\|error" | sed "s/^/[level $lv] /" >> gpurun_out/r2w_dist_mg_2gpu_levels.log
done
cat gpurun_out/r2w_dist_mg_2gpu_levels.log
