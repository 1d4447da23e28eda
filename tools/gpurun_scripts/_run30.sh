#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/vcycle_once.py > gpurun_out/r2z_vcycle_once.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_vcycle_launches.csv python tools/vcycle_once.py > gpurun_out/r2z_vcycle_ncu.log 2>&1
tail -2 gpurun_out/r2z_vcycle_ncu.log; wc -l gpurun_out/r2z_vcycle_launches.csv
