python tools/hbm_write_probe.py 2>&1 | tail -3
