#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests -q -m gpu --tb=short -x 2>&1 | grep -v "Warning\|torch.sparse\|^$" | tail -15 > gpurun_out/r2v_gpu_suite.log
cat gpurun_out/r2v_gpu_suite.log
