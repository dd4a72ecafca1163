timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/fwd_timing_probe.py 2>&1 | tail -2
STMGCN_HC_TMA=0 python tools/fwd_timing_probe.py 2>&1 | tail -2
