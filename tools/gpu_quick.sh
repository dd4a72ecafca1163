#!/bin/bash
# quick LSTM-kernel iteration on one GPU: targeted parity tests, kernel timing, role accounting (instrumented build)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "lstm or tensor_core or golden or cfg3 or bf16 or training" 2>&1 | tail -4 | cut -c1-300
python tools/lstm_time.py 2>&1 | tail -6
if [ -f st-mgcn_b200/lib/libstmgcn_b200_prof.so ]; then STMGCN_LIB_PATH=st-mgcn_b200/lib/libstmgcn_b200_prof.so python tools/tc_role_profile16.py 2>&1 | tee gpurun_out/roles16.txt | tail -12; fi
if [ "$1" = "ncu" ]; then timeout 600 ./tools/ncu_lstm16.sh 2>&1 | tail -2; fi
