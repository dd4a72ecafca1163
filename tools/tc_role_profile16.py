"""Per-role barrier-wait accounting of the second-generation LSTM kernels (lstm16.cu), instrumented build.

    STMGCN_TC_PROFILE=1 python -m stmgcn_b200.build          # -> lib/libstmgcn_b200_prof.so
    STMGCN_LIB_PATH=st-mgcn_b200/lib/libstmgcn_b200_prof.so python tools/tc_role_profile16.py

Prints, per role, the share of the role's lifetime spent blocked on each barrier class.
forward : producer (0: stage empty) | MMA (1: stage full, 2: TMEM empty) | epilogue (3: TMEM full)
backward: compute (0: dA buffer free, 1: recompute ready, 2: data-gradient accumulator ready, 3: recompute of a tile's first chunk ready)
          MMA (0: weight chunk landed, 1: dA ready, 2: recompute buffer / dgrad accumulator free, 3: A planes landed)
          producer (0: weight stage free, 1: A planes free)
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "st-mgcn_b200"))
from stmgcn_b200 import _lib, ops  # noqa: E402

ROLES = {0: "fwd producer", 1: "fwd mma", 2: "fwd epilogue", 3: "bwd compute", 4: "bwd mma", 5: "bwd producer"}


def read(reset=True):
    buf = (ctypes.c_ulonglong * 64)()
    fn = _lib.lib.stmgcn_dbg_tc_prof16
    fn.restype = ctypes.c_int32
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    fn(buf, 1 if reset else 0)
    return list(buf)


def report(tag):
    v = read()
    print(f"== {tag}")
    for role, name in ROLES.items():
        tot = v[48 + role]
        if tot == 0:
            continue
        waits = [v[role * 4 + i] / tot for i in range(4)]
        print(f"  {name:14s} lifetime {tot / 1e6:9.1f} Mcycles  wait0 {waits[0]:.3f} wait1 {waits[1]:.3f} wait2 {waits[2]:.3f} "
              f"wait3 {waits[3]:.3f}  busy {1 - sum(waits):.3f}")


def main():
    dev = "cuda:0"
    n, b, t, hid, lyr = 4096, 64, 12, 64, 3
    xo = torch.randn(n, b, t, 1, device=dev)
    s = torch.rand(b, t, device=dev)
    ws = []
    for l in range(lyr):
        in_l = 1 if l == 0 else hid
        ws += [torch.randn(4 * hid, in_l, device=dev) * 0.1, torch.randn(4 * hid, hid, device=dev) * 0.1,
               torch.randn(4 * hid, device=dev) * 0.1, torch.randn(4 * hid, device=dev) * 0.1]
    ws = [w.requires_grad_(True) for w in ws]
    d_top = torch.randn(n, b, hid, device=dev)
    for it in range(2):
        read()
        h_top, _, _ = ops.SharedLSTM.apply(xo, s, None, None, lyr, hid, False, *ws)
        report(f"forward (36 launches), iteration {it}")
        h_top.backward(d_top)
        report(f"backward (time-fused launches + 3 reductions), iteration {it}")


if __name__ == "__main__":
    main()
