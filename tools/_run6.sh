ECH_MODULE_FLAGS="-DECH_COOP2_PIPE=0" python tools/tune_coop2.py 2>/dev/null
ECH_MODULE_FLAGS="-DECH_COOP2_NBUF=1" ECH_COOP2_WAVES=0 python tools/tune_coop2.py 2>/dev/null
ECH_MODULE_FLAGS="-DECH_COOP2_NBUF=1" ECH_COOP2_WAVES=2 python tools/tune_coop2.py 2>/dev/null
ECH_MODULE_FLAGS="-DECH_COOP2_NBUF=2 -DECH_COOP_WARPS=2" ECH_COOP2_WAVES=2 python tools/tune_coop2.py 2>/dev/null
ECH_MODULE_FLAGS="-DECH_COOP2_NBUF=2 -DECH_COOP_WARPS=2" ECH_COOP2_WAVES=0 python tools/tune_coop2.py 2>/dev/null
ECH_MODULE_FLAGS="-DECH_COOP2_NBUF=1" python -m pytest "tests/test_gpu_medium.py::test_medium_mesh_every_entry_matches_cpu_port[1-259]" -q --tb=short 2>&1 | tail -3
