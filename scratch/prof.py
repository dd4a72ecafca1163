import sys, os, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/st-mgcn_b200")
import torch
from torch import nn
import GCN, STMGCN
from stmgcn_b200 import _lib, synth
w = synth.WORKLOADS["cfg3"]; dev = "cuda:0"
pre = GCN.Adj_Preprocessor("chebyshev", w.cheb_order)
sups = [pre.process_sparse(a).to(dev) for a in synth.make_adjacency_list(w)]
torch.manual_seed(0)
model = STMGCN.ST_MGCN(**synth.model_kwargs(w)).to(dev)
x, y = synth.make_inputs(w); x, y = x.to(dev), y.to(dev)
def step():
    model.zero_grad(); loss = nn.MSELoss()(model(obs_seq=x, sta_adj_list=sups), y); loss.backward()
for _ in range(2): step()
fn = _lib.lib.stmgcn_dbg_tc_prof; buf = (ctypes.c_ulonglong * 64)()
fn(buf, 1); step(); fn(buf, 1)
names = {0:"fwd loader",1:"fwd mma",2:"fwd epi",3:"bwd128 loader",4:"bwd128 mma",5:"bwd128 epi",6:"bwd64 loader",7:"bwd64 mma",8:"bwd64 epi",9:"wgrad loader",10:"wgrad mma"}
cls = ["empty","full","tmem_empty","tmem_full/done"]
for r, n in names.items():
    tot = buf[48 + r]
    if tot == 0: continue
    print(f"{n:14s} total {tot/1e6:9.1f} Mcyc  " + "  ".join(f"{cls[i]}={100*buf[r*4+i]/tot:5.1f}%" for i in range(4)))
