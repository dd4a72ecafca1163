"""The dominant Chebyshev SpMM launch of BASELINE configs[2] alone: Y = 2 L X - Z on (N = 4096, F = 64*64) fp32, graph 0.
Used under ncu (tools/ncu_spmm.sh) to read the launch's DRAM traffic; prints the CUDA-event time otherwise."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "st-mgcn_b200")]
import torch
import GCN
from stmgcn_b200 import ops, synth
from stmgcn_b200.graph import supports_from_dense

w = synth.WORKLOADS["cfg3"]
dev = "cuda:0"
sup = GCN.Adj_Preprocessor("chebyshev", w.cheb_order).process_sparse(synth.make_adjacency(w.n_regions, 0, w.density)).to(dev)
g = supports_from_dense(sup).graphs[0]
f = w.batch * w.lstm_hidden
x, z, y = (torch.randn(w.n_regions, f, device=dev) for _ in range(3))
for _ in range(3):
    ops.spmm_step(g, False, 2.0, x, -1.0, z, 0.0, None, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.spmm_step(g, False, 2.0, x, -1.0, z, 0.0, None, y)
e1.record()
torch.cuda.synchronize()
alg = g.nnz * 8 + (w.n_regions + 1) * 4 + 3 * w.n_regions * f * 4
us = e0.elapsed_time(e1) / 10 * 1e3
print(json.dumps({"us_per_launch": us, "algorithmic_bytes": alg, "achieved_GBps": alg / us / 1e3, "nnz": g.nnz}))
# the same launch with the gathered operand read from its bf16 copy (bf16-arithmetic mode), writing y and its bf16 copy
x16, y16 = ops.to_bf16(x), torch.empty(x.shape, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.spmm_step16(g, False, 2.0, x16, -1.0, z, 0.0, None, y, y16)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    ops.spmm_step16(g, False, 2.0, x16, -1.0, z, 0.0, None, y, y16)
e1.record()
torch.cuda.synchronize()
alg16 = g.nnz * 8 + (w.n_regions + 1) * 4 + w.n_regions * f * (2 + 4 + 4 + 2)
us16 = e0.elapsed_time(e1) / 10 * 1e3
print(json.dumps({"bf16_gather_us_per_launch": us16, "algorithmic_bytes": alg16, "achieved_GBps": alg16 / us16 / 1e3}))
