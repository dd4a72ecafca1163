"""Per-role barrier-wait accounting of the tcgen05 LSTM kernels (library built with STMGCN_TC_PROFILE=1).

Runs the shared LSTM of one graph branch at cfg3 shapes and prints, per role, the share of the role's lifetime spent
blocked on each barrier class (0 = stage empty, 1 = stage/raw full, 2 = TMEM empty, 3 = TMEM full).
Usage: STMGCN_TC_PROFILE=1 python -m stmgcn_b200.build; python tools/tc_role_profile.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "st-mgcn_b200"))
from stmgcn_b200 import _lib, ops  # noqa: E402

ROLES = {0: "fwd loader", 1: "fwd mma", 2: "fwd epilogue", 3: "bwd128 compute", 4: "bwd128 mma", 5: "bwd128 epilogue",
         6: "bwd64 compute", 7: "bwd64 mma", 8: "bwd64 epilogue", 9: "wgrad loader", 10: "wgrad mma"}


def read(reset=True):
    buf = (ctypes.c_ulonglong * 64)()
    fn = _lib.lib.stmgcn_dbg_tc_prof
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
        print(f"  {name:16s} lifetime {tot / 1e6:9.1f} Mcycles  wait empty {waits[0]:.2f} full {waits[1]:.2f} "
              f"tmem_empty {waits[2]:.2f} tmem_full {waits[3]:.2f}  busy {1 - sum(waits):.2f}")


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
        report(f"backward (36 + 3 launches), iteration {it}")


if __name__ == "__main__":
    main()
