#!/bin/bash
cd $GRAFT_REPO_ROOT
for nb in 8 2 1; do
  echo "== ECH_ILU_NBLOCKS=$nb"
  ECH_ILU_NBLOCKS=$nb timeout 200 python -m pytest tests/test_gpu_reference_examples.py -q -x -s -k other_examples 2>&1 | grep "REFERENCE_EXAMPLE\|passed\|failed\|WARNING" | head -5
done
