"""Time the shared LSTM of one graph branch alone (cfg3 rows) -- forward (no_grad) and forward+backward -- on the current
kernels.  Usage: python tools/lstm_time.py [rows_n] [batch] [T]    (env STMGCN_LSTM_PATH=fma selects the exact-fp32 CUDA-core kernels)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "st-mgcn_b200")]
import torch
from stmgcn_b200 import ops, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
b = int(sys.argv[2]) if len(sys.argv) > 2 else 64
t = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dev = "cuda:0"
hid, lyr = 64, 3
g = torch.Generator().manual_seed(0)
xo = torch.randn(n, b, t, 1, generator=g).to(dev)
s = torch.rand(b, t, generator=g).to(dev)
ws = []
for l in range(lyr):
    in_l = 1 if l == 0 else hid
    ws += [(torch.rand(4 * hid, in_l, generator=g) - 0.5) * 0.25, (torch.rand(4 * hid, hid, generator=g) - 0.5) * 0.25,
           (torch.rand(4 * hid, generator=g) - 0.5) * 0.25, (torch.rand(4 * hid, generator=g) - 0.5) * 0.25]
ws = [w.to(dev) for w in ws]
d_top = torch.randn(n, b, hid, generator=g).to(dev) * 1e-3

def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def fwd_only():
    with torch.no_grad():
        ops.SharedLSTM.apply(xo, s, None, None, lyr, hid, False, *ws)

def fwd_bwd():
    wr = [w.detach().requires_grad_(True) for w in ws]
    sr = s.detach().requires_grad_(True)
    h, _, _ = ops.SharedLSTM.apply(xo, sr, None, None, lyr, hid, False, *wr)
    h.backward(d_top)

out = {"rows": n * b, "T": t, "path": ops.lstm_path(), "planes": ops.lstm_planes(),
       "fwd_ms": timed(fwd_only), "fwd_bwd_ms": timed(fwd_bwd)}
print(json.dumps(out))
