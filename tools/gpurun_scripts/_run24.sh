#!/bin/bash
cd $GRAFT_REPO_ROOT
python tools/mg_profile.py 2>&1 | grep -v "Warning\|torch.sparse" > gpurun_out/r2x_mg_profile2.log
grep "set-up\|V-cycle total" gpurun_out/r2x_mg_profile2.log
