timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/fwd_timing_probe.py 2>&1 | tail -2
STMGCN_FWD_TMA=0 python tools/fwd_timing_probe.py 2>&1 | tail -2
