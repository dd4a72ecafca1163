"""Timing probe for the forward LSTM cell kernel (one graph branch at cfg3 shapes): training mode (gate tape written)
vs no_grad (no tape, no proxy fence in the epilogue).  STMGCN_DBG_SKIP_HC=1 additionally drops the h/c stores."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "st-mgcn_b200"))
from stmgcn_b200 import ops  # noqa: E402

dev = "cuda:0"
n, b, t, hid, lyr = 4096, 64, 12, 64, 3
xo = torch.randn(n, b, t, 1, device=dev)
s = torch.rand(b, t, device=dev)
ws = []
for l in range(lyr):
    in_l = 1 if l == 0 else hid
    ws += [torch.randn(4 * hid, in_l, device=dev) * 0.1, torch.randn(4 * hid, hid, device=dev) * 0.1,
           torch.randn(4 * hid, device=dev) * 0.1, torch.randn(4 * hid, device=dev) * 0.1]


def run(train):
    w = [x.clone().requires_grad_(train) for x in ws]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    with torch.set_grad_enabled(train):
        ops.SharedLSTM.apply(xo, s, None, None, lyr, hid, False, *w)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for mode in (True, False):
    run(mode)
    ts = [run(mode) for _ in range(4)]
    print("train" if mode else "no_grad", "skip_hc=" + os.environ.get("STMGCN_DBG_SKIP_HC", "0"), ["%.3f" % x for x in ts])
