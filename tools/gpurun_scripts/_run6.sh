python tools/tune_coop2.py 2>/dev/null
ECH_MODULE_FLAGS="-DECH_MAXNREG=224" python tools/tune_coop2.py 2>/dev/null
ECH_MODULE_FLAGS="-DECH_MAXNREG=216" python tools/tune_coop2.py 2>/dev/null
ECH_COOP2_WAVES=3 python tools/tune_coop2.py 2>/dev/null
ECH_COOP2_WAVES=4 python tools/tune_coop2.py 2>/dev/null
ECH_COOP2_WAVES=1 python tools/tune_coop2.py 2>/dev/null
