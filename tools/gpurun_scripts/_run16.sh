#!/bin/bash
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_multigrid.py -q -x --tb=short -s -k auxiliary 2>&1 | grep -v Warning | tail -8 > gpurun_out/r2t_mg_tests.log
cat gpurun_out/r2t_mg_tests.log
python tools/mg_profile.py 2>&1 | grep -v "Warning\|torch.sparse" > gpurun_out/r2t_mg_profile.log
cat gpurun_out/r2t_mg_profile.log
